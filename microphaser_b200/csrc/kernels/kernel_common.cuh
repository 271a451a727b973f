// kernel_common.cuh — constants and small device helpers shared by the phasing kernels (internal to csrc/kernels).
#pragma once
#include <algorithm>

#include "phase_kernels.cuh"

#include "../core/phase_core.h"
#include "../core/replay_core.h"

namespace mphk {

// Every launch goes through this macro: the calling thread's count is what mph_timing.kernel_launches (the bench line's
// "gpu_launches") reports, so the figure is the number of kernels actually launched, not an estimate.
extern thread_local uint64_t g_kernel_launches;
#define MPH_LAUNCH_UNPACK(...) __VA_ARGS__
#define MPH_LAUNCH(kern, cfg, ...) \
  do { kern<<<MPH_LAUNCH_UNPACK cfg>>>(__VA_ARGS__); ++::mphk::g_kernel_launches; } while (0)

namespace detail {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int K2_WARPS = 1;  // warps (chunks) per CTA of the window kernels: one, so that a finished chunk frees its slot at once
                             // (measured on B200: 4 -> 1.62 ms, 2 -> 1.59 ms, 1 -> 1.49 ms per whole-exome shard)
constexpr int K2_TABLE = 32;
constexpr int MAX_SEQ_CAP = 256;  // bytes per assembled sequence kept in local memory

__device__ __forceinline__ void raise(const DeviceBatch& d, uint32_t bits) {
  if (bits) atomicOr(&d.counters[CTR_ERR], bits);
}

constexpr int K2_LIST = 64;  // listed reads per warp between two flushes (phase B of the window kernels)

// order of the reference's BTreeMap<(haplotype, frame), count> (:383,434) on the packed key
__device__ __forceinline__ bool hist_less(const MphHist& a, const MphHist& b) {
  if (a.hap != b.hap) return a.hap < b.hap;
  const uint32_t fa = a.frame & 0x7FFFFFFFu, fb = b.frame & 0x7FFFFFFFu;
  if (fa != fb) return fa < fb;
  return (a.frame >> 31) < (b.frame >> 31);
}

}  // namespace detail
}  // namespace mphk
