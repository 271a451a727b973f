// phase_kernels.cu — hand-written sm_100a kernels of the per-window phasing path.
//
//   K1 k_allele_call   thread per read: supports_variant / bad_quality bitmasks over the variants
//                      inside the read (reference src/microphasing.rs:78-139). Coalesced SoA header
//                      loads; packed 4-bit bases are touched only where a variant lies.
//   K2 k_window_hist   CTA per chunk of consecutive windows of one exon, warp per window: candidate
//                      reads are a contiguous index range (binary search on the sorted starts);
//                      the common pair (no allele call, no bad base) costs one membership test and
//                      two ballots; the rest evaluates the closed form of the ObservationMatrix
//                      (:157-343) and is histogrammed with warp-aggregated inserts into a per-warp
//                      shared-memory table (:383-411).
//   K3 k_assemble      warp per chunk, thread per window: haplotype sequence walk (:458-603), stop
//                      codon test (:42-76, :694-697), "interesting window" flag.
//   K4 k_flag_count / k_block_scan / k_scatter   stable compaction of the interesting windows so the
//                      host sees them in the reference's order.
//   K5 k_live_depth    statistics only: sum of depth over the windows the reference reaches.
//
// No tensor cores: nothing here is a dense contraction; the path is integer / byte work bounded by
// HBM and L2 bandwidth and by instruction issue in K2.
#include "phase_kernels.cuh"

#include "../core/phase_core.h"

namespace mphk {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int K2_WARPS = 8;
constexpr int K2_TABLE = 32;
constexpr int MAX_SEQ_CAP = 256;  // bytes per assembled sequence kept in local memory

__device__ __forceinline__ void raise(const DeviceBatch& d, uint32_t bits) {
  if (bits) atomicOr(&d.counters[CTR_ERR], bits);
}

// ------------------------------------------------------------------ K1
__global__ void __launch_bounds__(256) k_allele_call(const DeviceBatch d) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.n_reads) return;
  const uint32_t nv = d.read_nv[r];
  const uint32_t hf = d.read_flags[r];
  MphCall c;
  c.S = 0;
  c.B = 0;
  if (nv) {
    MphRead rd;
    rd.start = d.read_start[r];
    rd.end = d.read_end[r];
    rd.vlo = d.read_vlo[r];
    rd.l_seq = d.read_lseq[r];
    rd.nv = nv;
    rd.n_cig = d.read_ncig[r];
    const uint8_t* bases = d.bases + (size_t)d.read_seq_off[r] * 16;
    const uint32_t* cig = d.cigars + d.read_cig_off[r];
    c = mph_call_read(rd, bases, cig, d.vars);
  }
  d.call_S[r] = c.S;
  d.call_B[r] = c.B;
  d.call_flags[r] = (uint8_t)((((c.S | c.B) != 0) ? 1u : 0u) | ((hf & MPH_RF_PARTNER) ? 2u : 0u));
  if (hf & MPH_RF_OVERFLOW) raise(d, MPH_E_VARS_PER_WINDOW);
}

// ------------------------------------------------------------------ K2
__device__ __forceinline__ uint32_t partner_lookup(const DeviceBatch& d, uint32_t r) {
  uint32_t lo = 0, hi = d.n_pairs;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (d.pairs[mid].x < r) lo = mid + 1;
    else hi = mid;
  }
  return (lo < d.n_pairs && d.pairs[lo].x == r) ? d.pairs[lo].y : NONE;
}

__device__ __forceinline__ bool hist_less(const MphHist& a, const MphHist& b) {
  if (a.hap != b.hap) return a.hap < b.hap;
  const uint32_t fa = a.frame & 0x7FFFFFFFu, fb = b.frame & 0x7FFFFFFFu;
  if (fa != fb) return fa < fb;
  return (a.frame >> 31) < (b.frame >> 31);
}

__global__ void __launch_bounds__(K2_WARPS * 32) k_window_hist(const DeviceBatch d) {
  __shared__ MphHist table[K2_WARPS][K2_TABLE];
  __shared__ MphSegment s_seg;
  __shared__ unsigned long long cta_depth;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const MphChunk ch = d.chunks[blockIdx.x];
  if (threadIdx.x < sizeof(MphSegment) / 4) reinterpret_cast<uint32_t*>(&s_seg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&d.segs[ch.seg])[threadIdx.x];
  if (threadIdx.x == 0) cta_depth = 0;
  __syncthreads();
  const MphSegment& sg = s_seg;
  const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
  const bool has_fs = (sg.flags & MPH_SF_HAS_FS) != 0;
  const uint32_t s0 = sg.off0 - sg.ceo;
  unsigned long long warp_depth = 0;
  for (uint32_t wi = warp; wi < ch.n; wi += K2_WARPS) {
    const uint32_t i = ch.i_first + wi;
    const uint32_t k = sg.k_first + i * sg.k_stride;
    const uint32_t widx = sg.win_base + i;
    const MphGeom g = mph_geom(sg, k);
    const uint32_t va = mph_var_lb(d.vars, sg.var_lo, sg.var_hi, g.s);
    const uint32_t vb = mph_var_lb(d.vars, va, sg.var_hi, g.e);
    if (vb - va > 64 && lane == 0) raise(d, MPH_E_VARS_PER_WINDOW);
    uint32_t rlo, rhi;
    mph_candidate_range(sg, d.read_start, g, &rlo, &rhi);
    uint32_t depth = 0, c0 = 0, n_keys = 0;
    for (uint32_t base = rlo; base < rhi; base += 32) {
      const uint32_t r = base + lane;
      const bool valid = r < rhi;
      uint32_t st = 0, en = 0, cf = 0;
      if (valid) {
        st = d.read_start[r];
        en = d.read_end[r];
        cf = d.call_flags[r];
      }
      bool member = false, counted = false;
      uint64_t hap = 0;
      uint32_t frame = 0;
      if (valid && en >= g.e) {
        if (cf == 0 && !has_fs) {
          // no allele call, no bad base, no duplicate qname: membership only
          if (!rev) member = (st <= s0) ? ((int64_t)st >= (int64_t)s0 - (int64_t)sg.K) : (st > sg.off0 && st - sg.off0 <= k);
          else member = (uint64_t)st + sg.K >= g.s;
          counted = member;
        } else {
          const uint32_t vlo = d.read_vlo[r];
          const uint64_t S = d.call_S[r], B = d.call_B[r];
          MphPair p;
          if (!rev) {
            p = mph_fwd_state(sg, d.vars, k, g, va, vb, st, en, vlo, S, B);
          } else {
            p.member = 0; p.bad = 0; p.hap = 0; p.frame = 0;
            const uint64_t Bx = B | (S & mph_range_mask(sg.sl_va, sg.sl_vb, vlo));
            uint32_t ke = mph_rev_entry(sg, d.vars, k, st, en, vlo, Bx);
            if (ke != NONE && (cf & 2u)) {
              // `contains` (:281-294): of two reads sharing (start, qname) only the first to enter stays
              const uint32_t q = partner_lookup(d, r);
              if (q != NONE) {
                const uint32_t qs = d.read_start[q], qe = d.read_end[q], qv = d.read_vlo[q];
                if (qs <= g.s && qe >= g.e) {
                  const uint64_t Bq = d.call_B[q] | (d.call_S[q] & mph_range_mask(sg.sl_va, sg.sl_vb, qv));
                  const uint32_t kq = mph_rev_entry(sg, d.vars, k, qs, qe, qv, Bq);
                  if (kq != NONE && (kq < ke || (kq == ke && q < r))) ke = NONE;
                }
              }
            }
            if (ke != NONE) p = mph_rev_state(sg, d.vars, k, g, va, vb, st, en, vlo, S, B, ke);
          }
          member = p.member != 0;
          counted = member && !p.bad;
          hap = p.hap;
          frame = p.frame;
        }
      }
      depth += __popc(__ballot_sync(FULL, member));
      const bool zero_key = counted && hap == 0 && frame == 0;
      c0 += __popc(__ballot_sync(FULL, zero_key));
      unsigned pending = __ballot_sync(FULL, counted && !zero_key);
      while (pending) {
        const int leader = __ffs(pending) - 1;
        const uint64_t lh = __shfl_sync(FULL, hap, leader);
        const uint32_t lf = __shfl_sync(FULL, frame, leader);
        const unsigned same = __ballot_sync(FULL, counted && !zero_key && hap == lh && frame == lf);
        if (lane == 0) {
          uint32_t t = 0;
          for (; t < n_keys; ++t)
            if (table[warp][t].hap == lh && table[warp][t].frame == lf) break;
          if (t == n_keys) {
            if (n_keys < K2_TABLE) {
              table[warp][t].hap = lh;
              table[warp][t].frame = lf;
              table[warp][t].count = 0;
              ++n_keys;
            } else {
              raise(d, MPH_E_KEYS_PER_WINDOW);
              t = K2_TABLE - 1;
            }
          }
          table[warp][t].count += __popc(same);
        }
        pending &= ~same;
      }
    }
    if (lane == 0) {
      // keys in the reference's BTreeMap order (:383,434)
      for (uint32_t a = 1; a < n_keys; ++a) {
        const MphHist key = table[warp][a];
        uint32_t b = a;
        while (b > 0 && hist_less(key, table[warp][b - 1])) {
          table[warp][b] = table[warp][b - 1];
          --b;
        }
        table[warp][b] = key;
      }
      MphWinOut wo;
      wo.depth = depth;
      wo.c0 = c0;
      wo.n_extra = n_keys;
      wo.extra_off = 0;
      if (n_keys) {
        const uint32_t off = atomicAdd(&d.counters[CTR_HIST], n_keys);
        if (off + n_keys <= d.hist_cap) {
          wo.extra_off = off;
          for (uint32_t a = 0; a < n_keys; ++a) d.hist[off + a] = table[warp][a];
        } else {
          raise(d, MPH_E_HIST_OVERFLOW);
          wo.n_extra = 0;
        }
      }
      d.win_out[widx] = wo;
      warp_depth += depth;
    }
    __syncwarp();
  }
  if (lane == 0 && warp_depth) atomicAdd(&cta_depth, warp_depth);
  __syncthreads();
  if (threadIdx.x == 0 && cta_depth) atomicAdd(d.sum_depth, cta_depth);
}

// ------------------------------------------------------------------ K3
__device__ void assemble_entry(const DeviceBatch& d, const MphSegment& sg, const MphGeom& g, uint32_t va, uint32_t vb, uint64_t hap,
                               bool boundary, uint8_t* seq, uint8_t* germ, MphHap* out, uint32_t* err) {
  const uint32_t cap = d.seq_cap;
  if (va == vb) {
    *err |= mph_plain_window(sg, g, d.ref, out);
    if (boundary && !(*err & MPH_E_REF_RANGE)) {
      const uint32_t off = atomicAdd(&d.counters[CTR_SEQ], 2 * cap);
      if (off + 2 * cap <= d.seq_cap_bytes) {
        const uint8_t* p = d.ref + sg.ref_off + (g.s - sg.ref_pos0);
        const uint32_t len = g.e - g.s;
        for (uint32_t t = 0; t < len && t < cap; ++t) {
          const uint8_t c = p[t];
          d.seq[off + t] = c;
          d.seq[off + cap + t] = c;
        }
        if (len > cap) out->flags |= MPH_HF_OVERFLOW;
        out->seq_off = off;
        out->flags |= MPH_HF_SEQ;
      } else {
        *err |= MPH_E_SEQ_OVERFLOW;
      }
    }
    return;
  }
  *err |= mph_assemble(sg, g, d.vars, va, vb, d.ref, d.ins_bytes, hap, seq, germ, cap, out);
  if (boundary || out->n_som > 0) {
    const uint32_t off = atomicAdd(&d.counters[CTR_SEQ], 2 * cap);
    if (off + 2 * cap <= d.seq_cap_bytes) {
      const uint32_t sl = out->seq_len < cap ? out->seq_len : cap, gl = out->germ_len < cap ? out->germ_len : cap;
      for (uint32_t t = 0; t < sl; ++t) d.seq[off + t] = seq[t];
      for (uint32_t t = 0; t < gl; ++t) d.seq[off + cap + t] = germ[t];
      out->seq_off = off;
      out->flags |= MPH_HF_SEQ;
    } else {
      *err |= MPH_E_SEQ_OVERFLOW;
    }
  }
}

__global__ void __launch_bounds__(256) k_assemble(const DeviceBatch d) {
  const uint32_t chunk = blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (chunk >= d.n_chunks) return;
  const MphChunk ch = d.chunks[chunk];
  if (lane >= ch.n) return;
  const MphSegment sg = d.segs[ch.seg];
  const uint32_t i = ch.i_first + lane;
  const uint32_t k = sg.k_first + i * sg.k_stride;
  const uint32_t widx = sg.win_base + i;
  const MphGeom g = mph_geom(sg, k);
  const uint32_t va = mph_var_lb(d.vars, sg.var_lo, sg.var_hi, g.s);
  const uint32_t vb = mph_var_lb(d.vars, va, sg.var_hi, g.e);
  const MphWinOut wo = d.win_out[widx];
  const bool boundary = i == 0 || i + 1 == sg.n_win || (sg.flags & MPH_SF_HAS_FS);
  uint8_t seq[MAX_SEQ_CAP], germ[MAX_SEQ_CAP];
  uint32_t err = 0;
  MphHap h0;
  assemble_entry(d, sg, g, va, vb, 0, boundary, seq, germ, &h0, &err);
  d.hap0[widx] = h0;
  for (uint32_t x = 0; x < wo.n_extra; ++x) {
    MphHap hx;
    assemble_entry(d, sg, g, va, vb, d.hist[wo.extra_off + x].hap, boundary, seq, germ, &hx, &err);
    d.hapx[wo.extra_off + x] = hx;
  }
  d.win_flag[widx] = (va != vb || (h0.flags & MPH_HF_STOP) || boundary || wo.n_extra > 0) ? 1 : 0;
  raise(d, err);
}

// ------------------------------------------------------------------ K4: stable compaction
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) k_flag_count(const DeviceBatch d) {
  const uint32_t w = blockIdx.x * SCAN_THREADS + threadIdx.x;
  const int f = (w < d.n_windows) ? d.win_flag[w] : 0;
  const int n = __syncthreads_count(f);
  if (threadIdx.x == 0) d.block_counts[blockIdx.x] = (uint32_t)n;
}

// exclusive scan of block_counts in place (one CTA), total -> counters[CTR_NIW]
__global__ void __launch_bounds__(SCAN_THREADS) k_block_scan(const DeviceBatch d, uint32_t n_blocks) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_blocks; base += SCAN_THREADS) {
    const uint32_t idx = base + threadIdx.x;
    const uint32_t v = idx < n_blocks ? d.block_counts[idx] : 0;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t s = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const uint32_t before = carry + (warp ? warp_sums[warp - 1] : 0) + (x - v);
    if (idx < n_blocks) d.block_counts[idx] = before;
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) d.counters[CTR_NIW] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scatter(const DeviceBatch d) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t w = blockIdx.x * SCAN_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool f = (w < d.n_windows) && d.win_flag[w];
  const unsigned bal = __ballot_sync(FULL, f);
  if (lane == 0) warp_sums[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    uint32_t s = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  if (f) {
    const uint32_t pos = d.block_counts[blockIdx.x] + (warp ? warp_sums[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1));
    d.iw[pos] = w;
    d.iw_out[pos] = d.win_out[w];
    d.iw_hap0[pos] = d.hap0[w];
  }
}

// ------------------------------------------------------------------ K5: statistics
__global__ void __launch_bounds__(256) k_live_depth(const DeviceBatch d) {
  const uint32_t chunk = blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long v = 0;
  if (chunk < d.n_chunks) {
    const MphChunk ch = d.chunks[chunk];
    if (lane < ch.n) {
      const uint32_t i = ch.i_first + lane;
      if (i < d.seg_live[ch.seg]) v = d.win_out[d.segs[ch.seg].win_base + i].depth;
    }
  }
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  if (lane == 0 && v) atomicAdd(d.live_depth, v);
}

}  // namespace

void launch_allele_call(const DeviceBatch& d, cudaStream_t st) {
  if (d.n_reads) k_allele_call<<<(d.n_reads + 255) / 256, 256, 0, st>>>(d);
}
void launch_window_hist(const DeviceBatch& d, cudaStream_t st) {
  if (d.n_chunks) k_window_hist<<<d.n_chunks, K2_WARPS * 32, 0, st>>>(d);
}
void launch_assemble(const DeviceBatch& d, cudaStream_t st) {
  if (d.n_chunks) k_assemble<<<(d.n_chunks + 7) / 8, 256, 0, st>>>(d);
}
void launch_compact(const DeviceBatch& d, cudaStream_t st) {
  const uint32_t nb = (d.n_windows + SCAN_THREADS - 1) / SCAN_THREADS;
  if (nb) k_flag_count<<<nb, SCAN_THREADS, 0, st>>>(d);
  k_block_scan<<<1, SCAN_THREADS, 0, st>>>(d, nb);
  if (nb) k_scatter<<<nb, SCAN_THREADS, 0, st>>>(d);
}
void launch_live_depth(const DeviceBatch& d, cudaStream_t st) {
  if (d.n_chunks) k_live_depth<<<(d.n_chunks + 7) / 8, 256, 0, st>>>(d);
}
int kernel_launch_count() { return 6; }

}  // namespace mphk
