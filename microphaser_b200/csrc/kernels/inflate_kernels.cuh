// inflate_kernels.cuh — launch entry point of the device-side BGZF inflate (kernels/inflate_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// one BGZF block of a batch: compressed bytes at cbuf + coff (clen of them), inflated to out + ooff (isize bytes)
struct MphRawBlock {
  uint32_t coff, clen, ooff, isize;
};

namespace mphk {

void launch_bgzf_inflate(const uint8_t* cbuf, const MphRawBlock* blocks, uint32_t n, uint8_t* out, uint32_t* status, cudaStream_t st);

}  // namespace mphk
