// inflate_kernels.cu — BGZF inflate on the device (reference: htslib's BGZF reader under rust-htslib `bam::IndexedReader`,
// src/main.rs:74). BGZF blocks are independent raw-DEFLATE streams of at most 64 KiB of payload, so a batch of a few thousand
// of them is inflated at once: one warp per block, lane 0 runs the decoder of core/inflate_core.h with its Huffman tables in
// shared memory. The compressed batch crosses the bus (a third of the inflated bytes); the file drivers hand the inflated
// batch to the same record framing / parse as the host inflate.
#include "inflate_kernels.cuh"
#include "kernel_common.cuh"

#include "../core/inflate_core.h"

namespace mphk {

namespace {

constexpr int INF_WARPS = 4;

__global__ void __launch_bounds__(INF_WARPS * 32) k_bgzf_inflate(const uint8_t* __restrict__ cbuf, const MphRawBlock* __restrict__ blocks, uint32_t n,
                                                                uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
  __shared__ MphInflateScratch scratch[INF_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t i = blockIdx.x * INF_WARPS + warp;
  if (i >= n || lane != 0) return;
  const MphRawBlock b = blocks[i];
  const int st = b.isize ? mph_inflate_raw(cbuf + b.coff, b.clen, out + b.ooff, b.isize, &scratch[warp]) : MPH_INF_OK;
  if (st != MPH_INF_OK) atomicMax(status, (uint32_t)st);
}

}  // namespace

void launch_bgzf_inflate(const uint8_t* cbuf, const MphRawBlock* blocks, uint32_t n, uint8_t* out, uint32_t* status, cudaStream_t st) {
  if (n) MPH_LAUNCH(k_bgzf_inflate, ((n + INF_WARPS - 1) / INF_WARPS, INF_WARPS * 32, 0, st), cbuf, blocks, n, out, status);
}

}  // namespace mphk
