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
#include "../core/replay_core.h"

namespace mphk {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int K2_WARPS = 4;
constexpr int K2_TABLE = 32;
constexpr int MAX_SEQ_CAP = 256;  // bytes per assembled sequence kept in local memory

__device__ __forceinline__ void raise(const DeviceBatch& d, uint32_t bits) {
  if (bits) atomicOr(&d.counters[CTR_ERR], bits);
}

// ------------------------------------------------------------------ K1
__global__ void __launch_bounds__(256) k_allele_call(const DeviceBatch d) {
  const uint32_t e = d.vr0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.vr1) return;
  const uint32_t r = d.vr_read[e];
  MphRead rd;
  rd.start = d.read_start[r];
  rd.end = d.read_end[r];
  rd.vlo = d.vr_vlo[e];
  rd.l_seq = d.vr_lseq[e];
  rd.nv = d.vr_nv[e];
  rd.n_cig = d.vr_ncig[e];
  const uint8_t* bases = d.bases + (size_t)d.vr_seq_off[e] * 16;
  const uint32_t* cig = d.cigars + d.vr_cig_off[e];
  const MphCall c = mph_call_read(rd, bases, cig, d.vars, d.mode == 0);
  d.call_S[r] = c.S;
  d.call_B[r] = c.B;
  d.call_flags[r] = (uint8_t)(((c.S | c.B) != 0) ? 1u : 0u);  // bit0: the read needs the full pair evaluation
  d.read_vlo[r] = rd.vlo;
  d.read_nv[r] = (uint8_t)rd.nv;
  d.read_vr[r] = e;
  if (d.read_flags[r] & MPH_RF_OVERFLOW) raise(d, MPH_E_VARS_PER_WINDOW);
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

// One (read, window) pair with the read flagged by K1 (allele calls, bad bases or a duplicate qname).
__device__ __forceinline__ MphPair eval_flagged(const DeviceBatch& d, const MphSegment& sg, bool rev, uint32_t k, const MphGeom& g, uint32_t va,
                                                uint32_t vb, uint32_t r, uint32_t st, uint32_t en, uint32_t cf, uint32_t vlo, uint64_t S, uint64_t B) {
  MphPair p;
  if (!rev) return mph_fwd_state(sg, d.vars, k, g, va, vb, st, en, vlo, S, B);
  p.member = 0; p.bad = 0; p.hap = 0; p.frame = 0;
  if (st > g.s || en < g.e) return p;
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
  return p;
}

// Wide variant: one warp per window, lanes over the candidate reads, keys in a 32-entry table.
// Used only for windows whose key count overflows the per-lane table of k_window_hist.
__device__ void window_hist_warp(const DeviceBatch& d, const MphSegment& sg, uint32_t i, uint32_t code, MphHist* table, int lane) {
  const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
  const uint32_t k = sg.k_first + i * sg.k_stride;
  const uint32_t widx = sg.win_base + i;
  const MphGeom g = mph_geom(sg, k);
  const uint32_t va = mph_var_lb(d.vars, sg.var_lo, sg.var_hi, g.s);
  const uint32_t vb = mph_var_lb(d.vars, va, sg.var_hi, g.e);
  uint32_t rlo, rhi;
  mph_candidate_range(sg, d.read_start, g, &rlo, &rhi);
  uint32_t depth = 0, c0 = 0, n_keys = 0;
  for (uint32_t base = rlo; base < rhi; base += 32) {
    const uint32_t r = base + lane;
    bool member = false, counted = false;
    uint64_t hap = 0;
    uint32_t frame = 0;
    if (r < rhi) {
      const uint32_t st = d.read_start[r], en = d.read_end[r];
      if (en >= g.e) {
        const uint32_t cf = d.call_flags[r] | ((d.read_flags[r] & MPH_RF_PARTNER) ? 2u : 0u);
        const MphPair p = eval_flagged(d, sg, rev, k, g, va, vb, r, st, en, cf, d.read_vlo[r], d.call_S[r], d.call_B[r]);
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
          if (table[t].hap == lh && table[t].frame == lf) break;
        if (t == n_keys) {
          if (n_keys < K2_TABLE) {
            table[t].hap = lh;
            table[t].frame = lf;
            table[t].count = 0;
            ++n_keys;
          } else {
            raise(d, MPH_E_KEYS_PER_WINDOW);
            t = K2_TABLE - 1;
          }
        }
        table[t].count += __popc(same);
      }
      pending &= ~same;
    }
  }
  if (lane == 0) {
    for (uint32_t a = 1; a < n_keys; ++a) {  // keys in the reference's BTreeMap order (:383,434)
      const MphHist key = table[a];
      uint32_t b = a;
      while (b > 0 && hist_less(key, table[b - 1])) {
        table[b] = table[b - 1];
        --b;
      }
      table[b] = key;
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
        for (uint32_t a = 0; a < n_keys; ++a) {
          d.hist[off + a] = table[a];
          d.hist_win[off + a] = code;
        }
      } else {
        raise(d, MPH_E_HIST_OVERFLOW);
        wo.n_extra = 0;
      }
    }
    d.win_out[widx] = wo;
    atomicAdd(d.sum_depth, (unsigned long long)depth);
    MphHap h0;
    const uint32_t err = mph_plain_hap(sg, g, d.stopmap, d.ref, vb - va, &h0);
    d.win_flag[widx] = 1;  // it has extra keys, hence it is interesting
    d.hap0[widx] = h0;
    raise(d, err);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(K2_WARPS * 32) k_window_hist_wide(const DeviceBatch d) {
  __shared__ MphHist table[K2_WARPS][K2_TABLE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = d.counters[CTR_OVF];
  for (uint32_t o = blockIdx.x * K2_WARPS + warp; o < n; o += gridDim.x * K2_WARPS) {
    const uint32_t code = d.ovf_list[o];
    const MphChunk ch = d.chunks[code >> 5];
    const MphSegment sg = d.segs[ch.seg];
    window_hist_warp(d, sg, ch.i_first + (code & 31u), code, table[warp], lane);
  }
}

// Main K2: one warp per chunk (<= 32 consecutive windows of one exon), two phases.
//  A (lane = read): walk the union of the windows' candidate ranges 32 reads at a time. A read
//    without allele calls / bad bases / duplicate qname is an observation of a *contiguous* run of
//    the chunk's windows (membership is monotone in the iteration number), so it contributes a
//    +1/-1 pair to a 33-entry difference array instead of 32 separate tests. Reads that carry an
//    allele call go to a short list, and so do the rare reads that need the full closed form.
//  B (lane = window): prefix-sum the difference array -> depth and the (hap 0, frame 0) count; then
//    only the listed reads are broadcast and evaluated per window, and their haplotype keys go to
//    per-lane shared-memory tables.
constexpr int K2_LANE_KEYS = 4;
constexpr int K2_LIST = 64;

__global__ void __launch_bounds__(K2_WARPS * 32) k_window_hist(const DeviceBatch d) {
  // per-lane key tables, [key][lane] so that a warp touches 32 distinct banks
  __shared__ uint64_t t_hap[K2_WARPS][K2_LANE_KEYS][32];
  __shared__ uint32_t t_cnt[K2_WARPS][K2_LANE_KEYS][32];
  __shared__ uint32_t t_frm[K2_WARPS][K2_LANE_KEYS][32];
  __shared__ MphSegment s_seg[K2_WARPS];
  __shared__ uint32_t s_s[K2_WARPS][32], s_e[K2_WARPS][32];
  __shared__ int s_add[K2_WARPS][34];
  __shared__ uint32_t s_list[K2_WARPS][K2_LIST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t chunk = d.c0 + blockIdx.x * K2_WARPS + warp;
  if (chunk >= d.c1) return;
  const MphChunk ch = d.chunks[chunk];
  if (lane < (int)(sizeof(MphSegment) / 4)) reinterpret_cast<uint32_t*>(&s_seg[warp])[lane] = reinterpret_cast<const uint32_t*>(&d.segs[ch.seg])[lane];
  s_add[warp][lane] = 0;
  if (lane < 2) s_add[warp][32 + lane] = 0;
  __syncwarp();
  const MphSegment& sg = s_seg[warp];
  if (sg.flags & MPH_SF_REPLAY) return;  // the whole transcript goes through k_replay
  const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
  const bool has_fs = (sg.flags & MPH_SF_HAS_FS) != 0;
  const int n = (int)ch.n;
  const bool active = lane < n;
  const uint32_t i = ch.i_first + (active ? lane : 0);
  const uint32_t k = sg.k_first + i * sg.k_stride;
  const MphGeom g = mph_geom(sg, k);
  const uint32_t va = mph_var_lb(d.vars, ch.va0, ch.vb1, g.s);
  const uint32_t vb = mph_var_lb(d.vars, va, ch.vb1, g.e);
  const uint32_t nvar = vb - va;
  if (active && nvar > 64) raise(d, MPH_E_VARS_PER_WINDOW);
  const bool chunk_has_var = ch.vb1 > ch.va0;
  s_s[warp][lane] = g.s;
  s_e[warp][lane] = g.e;
  const uint32_t s0 = sg.off0 - sg.ceo;
  const uint32_t rlo = ch.rlo, rhi = ch.rhi;  // union of the lanes' candidate ranges, resolved by the packer
  const uint32_t my_s = active ? g.s : 0u;
  const uint32_t my_e = active ? g.e : 0xFFFFFFFFu;  // inactive lanes: nothing encloses e = 0xFFFFFFFF
  const int64_t c1_lo = (int64_t)s0 - (int64_t)sg.K;  // forward: class-1 reads (offered at iteration 0)
  uint32_t depth_x = 0, n_keys = 0;  // depth_x: observations counted in phase B (complex reads)
  int c0_adj = 0;
  bool overflow = false;
  __syncwarp();
  auto add_key = [&](uint64_t hap, uint32_t frame) {
    uint32_t t = 0;
    for (; t < n_keys; ++t)
      if (t_hap[warp][t][lane] == hap && t_frm[warp][t][lane] == frame) break;
    if (t == n_keys) {
      if (n_keys == K2_LANE_KEYS || d.force_wide) { overflow = true; return; }
      t_hap[warp][t][lane] = hap;
      t_frm[warp][t][lane] = frame;
      t_cnt[warp][t][lane] = 0;
      ++n_keys;
    }
    t_cnt[warp][t][lane] += 1;
  };
  // phase B body for the listed reads (lane = window)
  auto process_list = [&](uint32_t list_n) {
    for (uint32_t x = 0; x < list_n; ++x) {
      const uint32_t code = s_list[warp][x];
      const uint32_t r = code & 0x7FFFFFFFu;
      const uint32_t st = d.read_start[r], en = d.read_end[r], vlo = d.read_vlo[r];
      const uint64_t S = d.call_S[r];
      if (!(code >> 31)) {
        // already counted as a plain observation; windows with variants still need its haplotype
        if (nvar == 0) continue;
        bool member;
        if (!rev) member = en >= my_e && ((st <= s0) ? ((int64_t)st >= c1_lo) : (st > sg.off0 && st - sg.off0 <= k));
        else member = st <= my_s && en >= my_e && (uint64_t)st + sg.K >= my_s;
        if (!member) continue;
        const uint64_t bits = mph_window_bits(S, vlo, va, nvar);
        const uint64_t hap = rev ? bits : (mph_bitrev64(bits) >> (64 - nvar));
        if (hap != 0) {
          c0_adj -= 1;
          add_key(hap, 0);
        }
      } else if (en >= my_e && st <= my_s) {
        const uint32_t cf = d.call_flags[r] | ((d.read_flags[r] & MPH_RF_PARTNER) ? 2u : 0u);
        const MphPair p = eval_flagged(d, sg, rev, k, g, va, vb, r, st, en, cf, vlo, S, d.call_B[r]);
        depth_x += p.member;
        if (p.member && !p.bad) {
          if (p.hap == 0 && p.frame == 0) c0_adj += 1;
          else add_key(p.hap, p.frame);
        }
      }
    }
  };
  // ---- phase A (lane = read)
  uint32_t list_n = 0;
  const uint32_t* ss = s_s[warp];
  const uint32_t* se = s_e[warp];
  for (uint32_t base = rlo; base < rhi; base += 32) {
    const uint32_t r = base + lane;
    int cls = 0, ilo = 1, ihi = 0;
    if (r < rhi) {
      const uint32_t st = d.read_start[r], en = d.read_end[r];
      const uint32_t cf = d.call_flags[r] | ((d.read_flags[r] & MPH_RF_PARTNER) ? 2u : 0u);
      if (cf == 0 && !has_fs) {
        cls = 1;
      } else {
        const uint32_t vlo = d.read_vlo[r];
        const uint64_t S = d.call_S[r], B = d.call_B[r];
        const uint64_t Bx = B | (S & mph_range_mask(sg.sl_va, sg.sl_vb, vlo));
        if (Bx == 0 && !(cf & 2u) && !has_fs) cls = (S != 0 && chunk_has_var) ? 2 : 1;
        else cls = 3;
      }
      if (cls != 3) {
        // run of windows [ilo, ihi] of this chunk at which the read is an observation
        if (!rev) {
          int a = 0, b = n;  // e non-decreasing: windows with e <= en form a prefix
          while (a < b) { const int m = (a + b) >> 1; if (se[m] <= en) a = m + 1; else b = m; }
          ihi = a - 1;
          if (st <= s0) {
            ilo = ((int64_t)st >= c1_lo) ? 0 : n;
          } else if (st <= sg.off0) {
            ilo = n;
          } else {
            const uint32_t k_ins = st - sg.off0;  // first window with k >= k_ins
            const uint32_t t = k_ins > sg.k_first ? (k_ins - sg.k_first + sg.k_stride - 1) / sg.k_stride : 0;
            ilo = t > ch.i_first ? (int)min(t - ch.i_first, (uint32_t)n) : 0;
          }
        } else {
          int a = 0, b = n;  // s non-increasing: windows with s >= st form a prefix
          while (a < b) { const int m = (a + b) >> 1; if (ss[m] >= st) a = m + 1; else b = m; }
          ihi = a - 1;
          a = 0; b = n;      // first window with e <= en
          while (a < b) { const int m = (a + b) >> 1; if (se[m] > en) a = m + 1; else b = m; }
          ilo = a;
          const uint64_t lim = (uint64_t)st + sg.K;
          a = 0; b = n;      // first window with s <= st + K
          while (a < b) { const int m = (a + b) >> 1; if ((uint64_t)ss[m] > lim) a = m + 1; else b = m; }
          if (a > ilo) ilo = a;
        }
      }
    }
    const bool counted = (cls == 1 || cls == 2) && ilo <= ihi;
    // warp-aggregated updates of the difference array (one writer per distinct index)
    {
      const int key = counted ? ilo : 33;
      const unsigned m = __match_any_sync(FULL, key);
      if (counted && lane == __ffs(m) - 1) s_add[warp][ilo] += __popc(m);
      __syncwarp();
      const int key2 = counted ? ihi + 1 : 33;
      const unsigned m2 = __match_any_sync(FULL, key2);
      if (counted && lane == __ffs(m2) - 1) s_add[warp][ihi + 1] -= __popc(m2);
      __syncwarp();
    }
    const bool need = (cls == 2 && counted) || cls == 3;
    const unsigned nm = __ballot_sync(FULL, need);
    if (nm) {
      if (list_n + __popc(nm) > K2_LIST) {
        process_list(list_n);
        list_n = 0;
        __syncwarp();
      }
      if (need) s_list[warp][list_n + __popc(nm & ((1u << lane) - 1))] = r | (cls == 3 ? 0x80000000u : 0u);
      list_n += __popc(nm);
      __syncwarp();
    }
  }
  process_list(list_n);
  // ---- phase B: prefix sum of the difference array
  int run = s_add[warp][lane];
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL, run, o);
    if (lane >= o) run += y;
  }
  const uint32_t depth = (uint32_t)run + depth_x;
  const uint32_t c0 = (uint32_t)(run + c0_adj);
  // each lane sorts its keys (reference BTreeMap order :383,434)
  for (uint32_t a = 1; a < n_keys; ++a) {
    MphHist key;
    key.hap = t_hap[warp][a][lane]; key.frame = t_frm[warp][a][lane]; key.count = t_cnt[warp][a][lane];
    uint32_t b = a;
    while (b > 0) {
      MphHist prev;
      prev.hap = t_hap[warp][b - 1][lane]; prev.frame = t_frm[warp][b - 1][lane]; prev.count = t_cnt[warp][b - 1][lane];
      if (!hist_less(key, prev)) break;
      t_hap[warp][b][lane] = prev.hap; t_frm[warp][b][lane] = prev.frame; t_cnt[warp][b][lane] = prev.count;
      --b;
    }
    t_hap[warp][b][lane] = key.hap; t_frm[warp][b][lane] = key.frame; t_cnt[warp][b][lane] = key.count;
  }
  const bool ovf = active && overflow;
  const uint32_t mine = (active && !ovf) ? n_keys : 0;
  // warp-aggregated allocation in the key arena
  uint32_t incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t total = __shfl_sync(FULL, incl, 31);
  uint32_t base_off = 0;
  if (lane == 0 && total) base_off = atomicAdd(&d.counters[CTR_HIST], total);
  base_off = __shfl_sync(FULL, base_off, 0);
  const bool fits = base_off + total <= d.hist_cap;
  if (lane == 0 && total && !fits) raise(d, MPH_E_HIST_OVERFLOW);
  if (active && !ovf) {
    MphWinOut wo;
    wo.depth = depth;
    wo.c0 = c0;
    wo.n_extra = fits ? mine : 0;
    wo.extra_off = base_off + incl - mine;
    if (fits)
      for (uint32_t a = 0; a < mine; ++a) {
        MphHist h;
        h.hap = t_hap[warp][a][lane]; h.frame = t_frm[warp][a][lane]; h.count = t_cnt[warp][a][lane];
        d.hist[wo.extra_off + a] = h;
        d.hist_win[wo.extra_off + a] = (chunk << 5) | (uint32_t)lane;
      }
    const uint32_t widx = sg.win_base + i;
    d.win_out[widx] = wo;
    // Haplotype 0 needs no sequence walk (no variant is applied): stop test from the bitmap, and the
    // "interesting" decision. Only the extra keys (hap != 0) go through K3.
    MphHap h0;
    const uint32_t err = mph_plain_hap(sg, g, d.stopmap, d.ref, nvar, &h0);
    const bool boundary = mph_is_boundary(sg, i);
    const bool interesting = nvar > 0 || (h0.flags & MPH_HF_STOP) || boundary || wo.n_extra > 0;
    d.win_flag[widx] = interesting ? 1 : 0;
    if (interesting) d.hap0[widx] = h0;
    raise(d, err);
  }
  if (ovf) {
    const uint32_t o = atomicAdd(&d.counters[CTR_OVF], 1u);
    d.ovf_list[o] = (chunk << 5) | (uint32_t)lane;
  }
  unsigned long long dsum = (active && !ovf) ? depth : 0;
  for (int o = 16; o; o >>= 1) dsum += __shfl_down_sync(FULL, dsum, o);
  if (lane == 0 && dsum) atomicAdd(d.sum_depth, dsum);
}

// ------------------------------------------------------------------ serial replay
// One warp per irregular transcript. The matrix operations of one transcript are a strict sequence
// (core/replay_core.h: mph_replay_tx is the single-threaded statement of the same steps and what the
// CPU emulator runs); within one step the observations are independent, so the lanes stride over
// them. The observation list lives in shared memory (it spills to a global scratch slice if a
// window is deeper than RP_OBS), the segment's variant positions are cached in shared memory, and
// the read cursors move incrementally, so an iteration costs one global round trip.
constexpr int RP_WARPS = 1;
constexpr int RP_OBS = 512;
constexpr int RP_READS = 640;  // reads of one exon's candidate range cached in shared memory (single-exon units)
constexpr int RP_POS = 256;
constexpr int RP_STOPW = 24;  // stop-codon bitmap words cached per segment

struct RpShared {
  uint64_t o_hap[RP_OBS];
  uint32_t o_read[RP_OBS], o_key[RP_OBS], o_frame[RP_OBS];
  uint8_t o_flags[RP_OBS];
  uint32_t rs[RP_READS], re[RP_READS];
  uint8_t rf[RP_READS], im[RP_READS];
  uint32_t pos[RP_POS];
  uint32_t stop[RP_STOPW];
  uint32_t dq[MPH_RP_MAXCOLS], dqpos[MPH_RP_MAXCOLS];
  MphHist table[MPH_RP_KEYS];
  MphSegment sg;
};

// first index in [lo, hi) with a[idx] >= target (hi if none): the tile around the previous answer is tried before a binary search;
// two lower bounds over the same array with their loads issued together (one round trip in the common case)
__device__ __forceinline__ void warp_lb2(const uint32_t* __restrict__ a, uint32_t lo, uint32_t hi, uint32_t t0, uint32_t t1, uint32_t* c0, uint32_t* c1,
                                         int lane) {
  uint32_t g0 = *c0 > lo + 16u ? *c0 - 16u : lo, g1 = *c1 > lo + 16u ? *c1 - 16u : lo;
  if (g0 > hi) g0 = hi;
  if (g1 > hi) g1 = hi;
  const uint32_t i0 = g0 + lane, i1 = g1 + lane;
  const uint32_t v0 = i0 < hi ? a[i0] : 0xFFFFFFFFu, v1 = i1 < hi ? a[i1] : 0xFFFFFFFFu;
  const uint32_t p0 = g0 > lo ? a[g0 - 1] : 0u, p1 = g1 > lo ? a[g1 - 1] : 0u;
  const unsigned b0 = __ballot_sync(FULL, i0 >= hi || v0 >= t0), b1 = __ballot_sync(FULL, i1 >= hi || v1 >= t1);
  *c0 = ((g0 == lo || p0 < t0) && b0) ? g0 + (uint32_t)__ffs(b0) - 1u : mph_u32_lb(a, lo, hi, t0);
  *c1 = ((g1 == lo || p1 < t1) && b1) ? g1 + (uint32_t)__ffs(b1) - 1u : mph_u32_lb(a, lo, hi, t1);
}

__global__ void __launch_bounds__(RP_WARPS * 32) k_replay(const DeviceBatch d) {
  __shared__ RpShared sh_all[RP_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ti = d.rp0 + blockIdx.x * RP_WARPS + warp;
  if (ti >= d.rp1) return;
  RpShared& sh = sh_all[warp];
  const MphReplayTx t = d.replay[ti];
  MphReplayCtx c;
  c.read_start = d.read_start; c.read_end = d.read_end; c.read_flags = d.read_flags;
  c.read_vlo = d.read_vlo; c.read_nv = d.read_nv; c.read_vr = d.read_vr;
  c.vr_seq_off = d.vr_seq_off; c.vr_cig_off = d.vr_cig_off; c.vr_lseq = d.vr_lseq; c.vr_ncig = d.vr_ncig;
  c.bases = d.bases; c.cigars = d.cigars; c.call_S = d.call_S; c.call_B = d.call_B;
  c.pairs = reinterpret_cast<const uint32_t*>(d.pairs); c.n_pairs = d.n_pairs;
  c.vars = d.vars; c.segs = d.segs; c.seg_chunk0 = d.seg_chunk0; c.stopmap = d.stopmap; c.ref = d.ref;
  c.dq_init = d.dq_init;
  c.mode = 0; c.tx_id_bytes = nullptr; c.tx_id_off = nullptr; c.win_depth = nullptr; c.win_id = nullptr; c.o_last = nullptr; c.seg_err = d.seg_err;
  c.o_read = nullptr; c.o_hap = nullptr; c.o_frame = nullptr; c.o_flags = nullptr; c.o_inmat = d.o_inmat;
  c.win_out = d.win_out; c.hist = d.hist; c.hist_win = d.hist_win; c.hist_cap = d.hist_cap;
  c.hap0 = d.hap0; c.win_flag = d.win_flag; c.win_voff = d.win_voff; c.vlist = d.vlist; c.vlist_cap = d.vlist_cap;
  c.counters = d.counters; c.sum_depth = d.sum_depth;
  // observation list: shared memory first, generic pointers so that it can move to the global scratch slice
  uint64_t* o_hap = sh.o_hap;
  uint32_t *o_read = sh.o_read, *o_key = sh.o_key, *o_frame = sh.o_frame;
  uint8_t* o_flags = sh.o_flags;
  uint32_t o_cap = RP_OBS;
  // read-side arrays, indexed by the global read number: global memory, or (single-exon units whose candidate
  // range fits) shared-memory copies addressed through shifted pointers
  const uint32_t* rs = d.read_start;
  const uint32_t* re = d.read_end;
  const uint8_t* rf = d.read_flags;
  uint8_t* in_mat = d.o_inmat + t.obs_off - t.read_lo;
  uint32_t r_lo = t.read_lo, r_hi = t.read_hi;
  bool reads_cached = false;
  if (t.seg_hi - t.seg_lo == 1) {
    const MphSegment sg0 = d.segs[t.seg_lo];
    const MphGeom ga = mph_geom(sg0, 0), gz = mph_geom(sg0, sg0.n_iter ? sg0.n_iter - 1 : 0);
    const bool rv = (sg0.flags & MPH_SF_REVERSE) != 0;
    const uint32_t s_lo = rv ? min(ga.s, gz.s) : ga.s, s_hi = max(ga.s, gz.s);
    const uint32_t c_lo = mph_u32_lb(d.read_start, t.read_lo, t.read_hi, s_lo > sg0.K ? s_lo - sg0.K : 0u);
    const uint32_t c_hi = mph_u32_lb(d.read_start, c_lo, t.read_hi, s_hi + 1u);
    if (c_hi - c_lo <= RP_READS) {
      for (uint32_t x = lane; x < c_hi - c_lo; x += 32) {
        const uint32_t r = c_lo + x;
        sh.rs[x] = d.read_start[r]; sh.re[x] = d.read_end[r]; sh.rf[x] = d.read_flags[r]; sh.im[x] = 0;
      }
      rs = sh.rs - c_lo; re = sh.re - c_lo; rf = sh.rf - c_lo; in_mat = sh.im - c_lo;
      c.read_start = rs; c.read_end = re; c.read_flags = rf;
      r_lo = c_lo; r_hi = c_hi;
      reads_cached = true;
    }
  }
  if (!reads_cached)
    for (uint32_t x = lane; x < t.obs_cap; x += 32) in_mat[t.read_lo + x] = 0;
  uint32_t err = 0, ncols = t.dq_n <= MPH_RP_MAXCOLS ? t.dq_n : 0u, n_obs = 0;
  if (t.dq_n > MPH_RP_MAXCOLS) err |= MPH_E_VARS_PER_WINDOW;
  for (uint32_t j = lane; j < ncols; j += 32) sh.dq[j] = d.dq_init[t.dq_off + j];
  uint64_t last_window_vars = t.last_vars;
  uint32_t cur_lo = r_lo, cur_hi = r_lo;
  uint32_t vl_cur = 0, vl_end = 0;        // slice of the column-list arena owned by this warp
  unsigned long long depth_sum = 0;       // lane 0
  uint32_t key_bound = 0;                 // forward: min end over the observations, reverse: max start (cleanup is skipped when it cannot remove anything)
  for (uint32_t j = lane; j < ncols; j += 32) sh.dqpos[j] = d.vars[sh.dq[j]].pos;
  __syncwarp();

  bool panicked = false;  // uniform: the reference panics here (drain out of range, inverted BTreeMap range)
  auto shrink_left = [&](uint64_t n) -> bool {  // :220-229
    if (n > ncols) { panicked = true; return false; }
    if (n) {
      const uint32_t a = lane + (uint32_t)n, b = lane + 32 + (uint32_t)n;
      const uint32_t x0 = a < ncols ? sh.dq[a] : 0u, x1 = b < ncols ? sh.dq[b] : 0u;
      const uint32_t y0 = a < ncols ? sh.dqpos[a] : 0u, y1 = b < ncols ? sh.dqpos[b] : 0u;
      __syncwarp();
      if (a < ncols) { sh.dq[lane] = x0; sh.dqpos[lane] = y0; }
      if (b < ncols) { sh.dq[lane + 32] = x1; sh.dqpos[lane + 32] = y1; }
      ncols -= (uint32_t)n;
      const uint64_t mask = ncols >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << ncols) - 1);
      for (uint32_t o = lane; o < n_obs; o += 32) o_hap[o] &= mask;
      __syncwarp();
    }
    return true;
  };

  for (uint32_t si = t.seg_lo; si < t.seg_hi; ++si) {
    __syncwarp();
    if (lane < (int)(sizeof(MphSegment) / 4)) reinterpret_cast<uint32_t*>(&sh.sg)[lane] = reinterpret_cast<const uint32_t*>(&d.segs[si])[lane];
    __syncwarp();
    const MphSegment& sg = sh.sg;
    const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
    const bool is_short = (sg.flags & MPH_SF_SHORT) != 0;
    // variant positions of the exon: every window of the segment lies inside [exon_start, exon_end]
    const uint32_t vs_lo = mph_var_lb(d.vars, sg.var_lo, sg.var_hi, sg.exon_start);
    const uint32_t vs_hi = mph_var_lb(d.vars, vs_lo, sg.var_hi, sg.exon_end + 1u);
    const uint32_t n_pos = vs_hi - vs_lo;
    const bool cached = n_pos <= RP_POS;
    if (cached)
      for (uint32_t x = lane; x < n_pos; x += 32) sh.pos[x] = d.vars[vs_lo + x].pos;
    // stop-codon bitmap words of the segment's reference slice
    const uint32_t sw0 = sg.ref_off >> 5, sw_n = ((sg.ref_off + sg.ref_len) >> 5) - sw0 + 3;
    const bool stop_cached = sw_n <= RP_STOPW;
    if (stop_cached && lane < (int)sw_n) sh.stop[lane] = d.stopmap[sw0 + lane];
    const uint32_t* stopmap_seg = stop_cached ? sh.stop - sw0 : d.stopmap;
    __syncwarp();
    auto lbpos = [&](uint64_t x) -> uint32_t {  // first variant index with pos >= x
      if (!cached || x < sg.exon_start || x > (uint64_t)sg.exon_end + 1u) return mph_var_lb(d.vars, sg.var_lo, sg.var_hi, (uint32_t)x);
      uint32_t lo = 0, hi = n_pos;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sh.pos[mid] < (uint32_t)x) lo = mid + 1;
        else hi = mid;
      }
      return vs_lo + lo;
    };
    if (!shrink_left(last_window_vars)) {  // :1024
      if (lane == 0) d.seg_err[si] = 1;
      break;
    }
    last_window_vars = 0;
    uint64_t old_offset = sg.off0, old_end = (uint64_t)sg.off0 + sg.ewl;
    bool reached_end = false;
    uint32_t prev_va = vs_lo, prev_vb = vs_lo;
    uint32_t emit_k = sg.k_first, emit_i = 0;  // next enumerated window
    for (uint32_t k = 0; k < sg.n_iter; ++k) {
      const uint64_t offset = rev ? (uint64_t)sg.off0 - k : (uint64_t)sg.off0 + k;
      const MphGeom g = mph_geom(sg, k);
      const uint64_t rest = rev ? offset - sg.exon_start : sg.exon_end - (offset + sg.ewl);
      const bool is_last_exon_window = rest < 3, is_first_exon_window = k == 0;
      // for k >= 1 old_offset / old_end are the previous window's start / end, so every variant_tree.range count
      // (:1119-1170) is a difference of the previous and the current lower bounds; an inverted range panics
      uint32_t va, vb;
      if (k == 0 || !cached) { va = lbpos(g.s); vb = lbpos(g.e); }
      else {
        va = prev_va; vb = prev_vb;
        while (va < vs_hi && sh.pos[va - vs_lo] < g.s) ++va;
        while (va > vs_lo && sh.pos[va - 1 - vs_lo] >= g.s) --va;
        while (vb < vs_hi && sh.pos[vb - vs_lo] < g.e) ++vb;
        while (vb > vs_lo && sh.pos[vb - 1 - vs_lo] >= g.e) --vb;
      }
      const uint64_t nvars = vb - va;
      uint64_t added_vars, deleted_vars;
      if (is_first_exon_window) added_vars = nvars;
      else if (is_short || reached_end) added_vars = 0;
      else if (g.s > old_offset) { if (old_end > g.e) panicked = true; added_vars = vb - prev_vb; }
      else { if (g.s > old_offset) panicked = true; added_vars = prev_va - va; }
      if (offset == old_offset || is_short) deleted_vars = 0;
      else if (g.s > old_offset) deleted_vars = va - prev_va;
      else { if (g.e > old_end) panicked = true; deleted_vars = prev_vb - vb; }
      prev_va = va; prev_vb = vb;
      if (is_last_exon_window) reached_end = true;
      if (panicked) {
        if (lane == 0) d.seg_err[si] = k + 1;
        break;
      }
      // cleanup_reads (:259-278): in-place compaction, a tile is read completely before it is written
      if (n_obs && (rev ? key_bound > g.s : key_bound < g.e)) {
        uint32_t w = 0, kb = rev ? 0u : 0xFFFFFFFFu;
        for (uint32_t base = 0; base < n_obs; base += 32) {
          const uint32_t o = base + lane;
          const bool valid = o < n_obs;
          uint32_t r = 0, key = 0, fr = 0;
          uint64_t hp = 0;
          uint8_t fl = 0;
          if (valid) { r = o_read[o]; key = o_key[o]; hp = o_hap[o]; fr = o_frame[o]; fl = o_flags[o]; }
          const bool keep = valid && (rev ? key < g.s + 1u : key >= g.e);
          if (valid && !keep) in_mat[r] = 0;
          const unsigned bal = __ballot_sync(FULL, keep);
          __syncwarp();
          if (keep) {
            const uint32_t p = w + __popc(bal & ((1u << lane) - 1));
            o_read[p] = r; o_key[p] = key; o_hap[p] = hp; o_frame[p] = fr; o_flags[p] = fl;
            kb = rev ? max(kb, key) : min(kb, key);
          }
          w += __popc(bal);
          __syncwarp();
        }
        n_obs = w;
        key_bound = rev ? __reduce_max_sync(FULL, kb) : __reduce_min_sync(FULL, kb);
      }
      if (!shrink_left(deleted_vars)) {
        if (lane == 0) d.seg_err[si] = k + 1;
        break;
      }
      // candidate reads (:1191-1249) and push_read (:297-343)
      {
        const bool wide = rev || offset == (uint64_t)sg.exon_start + sg.ceo;
        const uint32_t lo = wide ? (g.s > sg.K ? g.s - sg.K : 0u) : g.s;
        warp_lb2(rs, r_lo, r_hi, lo, g.s + 1u, &cur_lo, &cur_hi, lane);
        if (cur_hi < cur_lo) cur_hi = cur_lo;
        for (uint32_t base4 = cur_lo; base4 < cur_hi; base4 += 128) {
         // the loads of four tiles are issued before any of them is processed
         uint32_t en4[4];
         uint8_t im4[4], rf4[4];
#pragma unroll
         for (int j = 0; j < 4; ++j) {
           const uint32_t rj = base4 + 32 * j + lane;
           const bool in = rj < cur_hi;
           en4[j] = in ? re[rj] : 0u;
           im4[j] = (in && rev) ? in_mat[rj] : (uint8_t)0;
           rf4[j] = (in && rev) ? rf[rj] : (uint8_t)0;
         }
#pragma unroll
         for (int j = 0; j < 4; ++j) {
          const uint32_t base = base4 + 32 * j;
          if (base >= cur_hi) break;
          const uint32_t r = base + lane;
          bool cand = r < cur_hi && en4[j] >= g.e;
          if (cand && rev) {
            // `contains` (:281-294): the read itself or the read sharing its (start, qname) is in the matrix already
            bool dup = im4[j] != 0;
            if (!dup && (rf4[j] & MPH_RF_PARTNER)) {
              const uint32_t q = mph_rp_partner(c, r);
              if (q != NONE && q >= r_lo && q < r_hi) {
                dup = in_mat[q] != 0;
                if (!dup && q < r && q >= cur_lo && re[q] >= g.e) {  // offered just before r in this same iteration
                  uint64_t hq = 0;
                  uint32_t fq = 0;
                  uint8_t lq = 0;
                  for (uint32_t i = 0; i < ncols; ++i) mph_rp_update(c, t, q, i, sh.dq[ncols - 1 - i], &hq, &fq, &lq, &err);
                  dup = !(lq & 1);
                }
              }
            }
            cand = !dup;
          }
          uint64_t hap = 0;
          uint32_t frame = 0;
          uint8_t fl = 0;
          if (cand) {
            for (uint32_t i = 0; i < ncols; ++i) mph_rp_update(c, t, r, i, sh.dq[ncols - 1 - i], &hap, &frame, &fl, &err);
            if (fl & 1) cand = false;  // rejected at push (:338)
          }
          const unsigned bal = __ballot_sync(FULL, cand);
          const uint32_t n_in = __popc(bal);
          bool drop = false;  // uniform
          if (n_in && n_obs + n_in > o_cap) {
            if (o_hap == sh.o_hap && n_obs + n_in <= t.obs_cap) {  // move the list to its global scratch slice
              uint64_t* g_hap = d.o_hap + t.obs_off;
              uint32_t *g_read = d.o_read + t.obs_off, *g_key = d.o_key + t.obs_off, *g_frame = d.o_frame + t.obs_off;
              uint8_t* g_flags = d.o_flags + t.obs_off;
              for (uint32_t o = lane; o < n_obs; o += 32) { g_hap[o] = o_hap[o]; g_read[o] = o_read[o]; g_key[o] = o_key[o]; g_frame[o] = o_frame[o]; g_flags[o] = o_flags[o]; }
              o_hap = g_hap; o_read = g_read; o_key = g_key; o_frame = g_frame; o_flags = g_flags;
              o_cap = t.obs_cap;
              __syncwarp();
            } else {
              err |= MPH_E_REPLAY_INPUT;
              drop = true;
            }
          }
          if (n_in && !drop) {
            uint32_t kk = rev ? 0u : 0xFFFFFFFFu;
            if (cand) {
              const uint32_t p = n_obs + __popc(bal & ((1u << lane) - 1));
              kk = rev ? rs[r] : en4[j];
              o_read[p] = r; o_key[p] = kk; o_hap[p] = hap; o_frame[p] = frame; o_flags[p] = fl;
              in_mat[r] = 1;
            }
            if (rev) key_bound = n_obs ? max(key_bound, __reduce_max_sync(FULL, kk)) : __reduce_max_sync(FULL, kk);
            else key_bound = n_obs ? min(key_bound, __reduce_min_sync(FULL, kk)) : __reduce_min_sync(FULL, kk);
            n_obs += n_in;
          }
          __syncwarp();
         }
        }
      }
      // newly collected variants (:1280-1296) and extend_right (:232-256)
      {
        const uint64_t skip = nvars - added_vars;  // wraps like the release build: nothing is added then
        const uint32_t n_new = skip <= nvars ? (uint32_t)(nvars - skip) : 0u;
        if (n_new) {
          if (ncols + n_new > MPH_RP_MAXCOLS) { err |= MPH_E_VARS_PER_WINDOW; break; }
          for (uint32_t o = lane; o < n_obs; o += 32) {
            uint64_t hap = o_hap[o] << (n_new & 63u);
            uint32_t frame = o_frame[o];
            uint8_t fl = o_flags[o];
            const uint32_t r = o_read[o];
            for (uint32_t i = 0; i < n_new; ++i) {
              const uint32_t x = (uint32_t)skip + (n_new - 1 - i);
              mph_rp_update(c, t, r, i, rev ? vb - 1 - x : va + x, &hap, &frame, &fl, &err);
            }
            o_hap[o] = hap; o_frame[o] = frame; o_flags[o] = fl;
          }
          for (uint32_t x = (uint32_t)skip + lane; x < (uint32_t)nvars; x += 32) {
            const uint32_t v = rev ? vb - 1 - x : va + x;
            sh.dq[ncols + (x - (uint32_t)skip)] = v;
            sh.dqpos[ncols + (x - (uint32_t)skip)] = d.vars[v].pos;
          }
          ncols += n_new;
          __syncwarp();
        }
      }
      last_window_vars = nvars;
      if (k == emit_k && emit_i < sg.n_win) {
        // histogram of print_haplotypes (:383-411) and the window's outputs
        const uint32_t i = emit_i;
        emit_k += sg.k_stride;
        ++emit_i;
        const uint32_t widx = sg.win_base + i;
        uint32_t n_keys = 0, c0 = 0;
        for (uint32_t base = 0; base < n_obs; base += 32) {
          const uint32_t o = base + lane;
          const bool valid = o < n_obs && !(o_flags[o] & 1);
          const uint64_t hap = valid ? o_hap[o] : 0;
          const uint32_t fr = valid ? o_frame[o] : 0;
          const bool zero_key = valid && hap == 0 && fr == 0;
          c0 += __popc(__ballot_sync(FULL, zero_key));
          unsigned pending = __ballot_sync(FULL, valid && !zero_key);
          while (pending) {
            const int leader = __ffs(pending) - 1;
            const uint64_t lh = __shfl_sync(FULL, hap, leader);
            const uint32_t lf = __shfl_sync(FULL, fr, leader);
            const unsigned same = __ballot_sync(FULL, valid && !zero_key && hap == lh && fr == lf);
            if (lane == 0) {
              uint32_t x = 0;
              for (; x < n_keys; ++x)
                if (sh.table[x].hap == lh && sh.table[x].frame == lf) break;
              if (x == n_keys) {
                if (n_keys < MPH_RP_KEYS) {
                  sh.table[x].hap = lh; sh.table[x].frame = lf; sh.table[x].count = 0;
                  ++n_keys;
                } else {
                  err |= MPH_E_KEYS_PER_WINDOW;
                  x = MPH_RP_KEYS - 1;
                }
              }
              sh.table[x].count += __popc(same);
            }
            pending &= ~same;
          }
        }
        n_keys = __shfl_sync(FULL, n_keys, 0);
        if (vl_cur + ncols + 1 > vl_end) {  // a fresh slice of the column-list arena (one atomic per ~256 entries)
          uint32_t got = 0;
          const uint32_t want = max(256u, ncols + 1);
          if (lane == 0) got = atomicAdd(&d.counters[CTR_VLIST], want);
          vl_cur = __shfl_sync(FULL, got, 0);
          vl_end = vl_cur + want;
        }
        const uint32_t voff = vl_cur;
        vl_cur += ncols + 1;
        if (lane == 0) {
          for (uint32_t a = 1; a < n_keys; ++a) {
            const MphHist key = sh.table[a];
            uint32_t b = a;
            while (b > 0 && hist_less(key, sh.table[b - 1])) { sh.table[b] = sh.table[b - 1]; --b; }
            sh.table[b] = key;
          }
          MphWinOut wo;
          wo.depth = n_obs; wo.c0 = c0; wo.n_extra = n_keys; wo.extra_off = 0;
          if (n_keys) {
            const uint32_t off = atomicAdd(&d.counters[CTR_HIST], n_keys);
            if (off + n_keys <= d.hist_cap) {
              wo.extra_off = off;
              const uint32_t code = ((d.seg_chunk0[si] + (i >> 5)) << 5) | (i & 31u);
              for (uint32_t a = 0; a < n_keys; ++a) { d.hist[off + a] = sh.table[a]; d.hist_win[off + a] = code; }
            } else {
              err |= MPH_E_HIST_OVERFLOW;
              wo.n_extra = 0;
            }
          }
          d.win_out[widx] = wo;
          depth_sum += n_obs;
          // variants the sequence walk visits (:473-476): j only advances while variants[j].pos == i
          uint32_t j = 0;
          for (uint32_t p = g.s; p < g.e && j < ncols; ++p)
            while (j < ncols && (rev ? sh.dqpos[ncols - 1 - j] : sh.dqpos[j]) == p) ++j;
          MphHap h0;
          err |= mph_plain_hap(sg, g, stopmap_seg, d.ref, j, &h0);
          if (ncols > 32) err |= MPH_E_VARS_PER_WINDOW;
          d.hap0[widx] = h0;
          d.win_flag[widx] = 1;
        }
        if (voff + ncols + 1 <= d.vlist_cap) {
          if (lane == 0) { d.vlist[voff] = ncols; d.win_voff[widx] = voff; }
          for (uint32_t j = lane; j < ncols; j += 32) d.vlist[voff + 1 + j] = rev ? sh.dq[ncols - 1 - j] : sh.dq[j];
        } else {
          err |= MPH_E_VLIST_OVERFLOW;
        }
        __syncwarp();
      }
      old_offset = g.s;
      old_end = g.e;
      if (is_short) break;
    }
    if (panicked || __any_sync(FULL, (err & (MPH_E_REPLAY_PANIC | MPH_E_VARS_PER_WINDOW)) != 0)) break;
  }
  if (lane == 0 && depth_sum) atomicAdd(d.sum_depth, depth_sum);
  raise(d, err);
}

// Normal mode: the same replay with the normal-mode matrix (every re-offered copy is kept, entries are
// (read, haplotype, copies)). The normal-mode residue writes a record for every window, so this path is far from the
// critical one; lane 0 of a warp runs the single-threaded statement of core/replay_core.h per unit.
__global__ void __launch_bounds__(64) k_replay_normal(const DeviceBatch d) {
  const uint32_t ti = d.rp0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (ti >= d.rp1 || (threadIdx.x & 31)) return;
  MphReplayCtx c;
  c.read_start = d.read_start; c.read_end = d.read_end; c.read_flags = d.read_flags;
  c.read_vlo = d.read_vlo; c.read_nv = d.read_nv; c.read_vr = d.read_vr;
  c.vr_seq_off = d.vr_seq_off; c.vr_cig_off = d.vr_cig_off; c.vr_lseq = d.vr_lseq; c.vr_ncig = d.vr_ncig;
  c.bases = d.bases; c.cigars = d.cigars; c.call_S = d.call_S; c.call_B = d.call_B;
  c.pairs = reinterpret_cast<const uint32_t*>(d.pairs); c.n_pairs = d.n_pairs;
  c.vars = d.vars; c.segs = d.segs; c.seg_chunk0 = d.seg_chunk0; c.stopmap = d.stopmap; c.ref = d.ref; c.dq_init = d.dq_init;
  c.mode = 1; c.tx_id_bytes = d.tx_id_bytes; c.tx_id_off = d.tx_id_off; c.win_depth = d.win_depth; c.win_id = d.win_id; c.o_last = d.o_key;
  c.o_read = d.o_read; c.o_hap = d.o_hap; c.o_frame = d.o_frame; c.o_flags = d.o_flags; c.o_inmat = d.o_inmat;
  c.win_out = d.win_out; c.hist = d.hist; c.hist_win = d.hist_win; c.hist_cap = d.hist_cap;
  c.hap0 = d.hap0; c.win_flag = d.win_flag; c.win_voff = d.win_voff; c.vlist = d.vlist; c.vlist_cap = d.vlist_cap;
  c.seg_err = d.seg_err; c.counters = d.counters; c.sum_depth = d.sum_depth;
  mph_replay_tx(c, d.replay[ti]);
}

// ------------------------------------------------------------------ K3
// One thread per extra histogram key (haplotype != 0): the sequence walk of print_haplotypes
// (:458-603) into thread-local buffers, then the stop test; the bytes are kept only for haplotypes
// that can be written (n_somatic > 0) or merged across a splice junction (boundary windows).
__global__ void __launch_bounds__(128) k_assemble(const DeviceBatch d) {
  const uint32_t n = min(d.counters[CTR_HIST], d.hist_cap);
  uint8_t seq[MAX_SEQ_CAP], germ[MAX_SEQ_CAP];
  const uint32_t cap = d.seq_cap;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    const uint64_t hap = d.hist[x].hap;
    if (hap == 0) continue;  // a (hap 0, frame != 0) key shares the window's haplotype-0 record
    const uint32_t code = d.hist_win[x];
    const MphChunk ch = d.chunks[code >> 5];
    const MphSegment sg = d.segs[ch.seg];
    const uint32_t i = ch.i_first + (code & 31u);
    const uint32_t k = sg.k_first + i * sg.k_stride;
    const MphGeom g = mph_geom(sg, k);
    const uint32_t va = mph_var_lb(d.vars, ch.va0, ch.vb1, g.s);
    const uint32_t vb = mph_var_lb(d.vars, va, ch.vb1, g.e);
    const bool replayed = (sg.flags & MPH_SF_REPLAY) != 0;
    const bool boundary = mph_is_boundary(sg, i) || replayed;
    MphHap out;
    uint32_t err;
    if (!replayed) {
      err = mph_assemble(sg, g, d.vars, va, vb, d.ref, d.ins_bytes, hap, seq, germ, cap, &out);
    } else {
      // the walk runs over the matrix columns the replay recorded, gathered into a dense array
      MphVar cols[MPH_RP_MAXCOLS];
      const uint32_t off = d.win_voff[sg.win_base + i];
      const uint32_t ncol = off == NONE ? 0u : min(d.vlist[off], (uint32_t)MPH_RP_MAXCOLS);
      for (uint32_t j = 0; j < ncol; ++j) cols[j] = d.vars[d.vlist[off + 1 + j]];
      err = mph_assemble(sg, g, cols, 0, ncol, d.ref, d.ins_bytes, hap, seq, germ, cap, &out);
    }
    // record id (:667-675): SHA-1 over the debug rendering of the assembled bytes, hashed while they are in registers / L1
    if ((out.n_som > 0 || (sg.flags & MPH_SF_HAS_FS)) && out.seq_len <= cap) {
      const uint32_t t0 = d.tx_id_off[sg.tx];
      out.id64 = mph_record_id64(seq, out.seq_len, d.tx_id_bytes + t0, d.tx_id_off[sg.tx + 1] - t0, g.s);
      out.flags |= MPH_HF_ID;
    }
    if (boundary || out.n_som > 0) {
      const uint32_t off = atomicAdd(&d.counters[CTR_SEQ], 2 * cap);
      if (off + 2 * cap <= d.seq_cap_bytes) {
        const uint32_t sl = out.seq_len < cap ? out.seq_len : cap, gl = out.germ_len < cap ? out.germ_len : cap;
        for (uint32_t t = 0; t < sl; ++t) d.seq[off + t] = seq[t];
        for (uint32_t t = 0; t < gl; ++t) d.seq[off + cap + t] = germ[t];
        out.seq_off = off;
        out.flags |= MPH_HF_SEQ;
      } else {
        err |= MPH_E_SEQ_OVERFLOW;
      }
    }
    d.hapx[x] = out;
    raise(d, err);
  }
}

// ------------------------------------------------------------------ normal mode (src/normal_microphasing.rs)
// The matrix of the normal mode keeps every copy of a re-offered read on the reverse strand
// (cleanup_reads(splice_side_offset) :1001, no `contains`), so an observation count is a ramp in the
// iteration number: a read contributes k - kc + 1 copies from its entry iteration kc until the window
// start reaches its own start, where the older copies are dropped. Phase A therefore keeps two
// difference arrays (slope and constant); the histogram only needs the reads that support an allele.
constexpr int KN_LANE_KEYS = 6;

__device__ __forceinline__ void normal_window_out(const DeviceBatch& d, const MphSegment& sg, const MphGeom& g, uint32_t widx, uint32_t nvar,
                                                  uint32_t depth) {
  MphHap h0;
  const uint32_t err = mph_nrm_plain(sg, g, d.ref, nvar, &h0);
  d.win_depth[widx] = depth | ((nvar == 0 && (h0.flags & MPH_NF_STOP)) ? 0x80000000u : 0u);
  // every window's reference haplotype is written by the host (:509-645): its record id is hashed here
  unsigned long long id = 0;
  if (!(h0.flags & MPH_NF_REFRANGE)) {
    const uint32_t t0 = d.tx_id_off[sg.tx];
    id = mph_record_id64(d.ref + sg.ref_off + (g.s - sg.ref_pos0), g.e - g.s, d.tx_id_bytes + t0, d.tx_id_off[sg.tx + 1] - t0, g.s);
  }
  d.win_id[widx] = id;
  d.win_flag[widx] = nvar > 0 ? 1 : 0;
  if (nvar) d.hap0[widx] = h0;
  raise(d, err);
}

// entry iteration of a reverse-strand read (first k with s(k) <= start + K and e(k) <= end), independent of the window
__device__ __forceinline__ uint32_t normal_rev_kc(const MphSegment& sg, uint32_t st, uint32_t en) {
  const uint64_t lim = (uint64_t)st + sg.K;
  uint32_t kc = sg.off0 > lim ? (uint32_t)(sg.off0 - lim) : 0;
  if (mph_geom(sg, 0).e > en) {
    const uint64_t t = (uint64_t)sg.off0 + sg.ewl;
    const uint32_t k2 = t > en ? (uint32_t)(t - en) : 1;
    if (k2 > kc) kc = k2;
    if (kc == 0) kc = 1;
  }
  kc = kc >= 2 ? kc - 2 : 0;
  for (; kc < sg.n_iter; ++kc) {
    const MphGeom gc = mph_geom(sg, kc);
    if (gc.s <= lim && gc.e <= en) break;
  }
  return kc;
}

__device__ void window_hist_warp_normal(const DeviceBatch& d, const MphSegment& sg, uint32_t i, uint32_t code, MphHist* table, int lane) {
  const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
  const uint32_t k = sg.k_first + i * sg.k_stride;
  const uint32_t widx = sg.win_base + i;
  const MphGeom g = mph_geom(sg, k);
  const uint32_t va = mph_var_lb(d.vars, sg.var_lo, sg.var_hi, g.s);
  const uint32_t vb = mph_var_lb(d.vars, va, sg.var_hi, g.e);
  uint32_t rlo, rhi;
  mph_candidate_range(sg, d.read_start, g, &rlo, &rhi);
  uint32_t depth = 0, c0 = 0, n_keys = 0;
  for (uint32_t base = rlo; base < rhi; base += 32) {
    const uint32_t r = base + lane;
    uint32_t copies = 0, kp_last = k, vlo = 0;
    uint64_t S = 0;
    if (r < rhi) {
      const uint32_t st = d.read_start[r], en = d.read_end[r];
      if (!rev) {
        const uint32_t kp = mph_nrm_fwd_entry(sg, k, g, st, en);
        if (kp != NONE) { copies = 1; kp_last = kp; }
      } else {
        uint32_t kc;
        copies = mph_nrm_rev_copies(sg, k, g, st, en, &kc);
      }
      if (copies) { S = d.call_S[r]; vlo = d.read_vlo[r]; }
    }
    uint32_t cs = copies;
    for (int o = 16; o; o >>= 1) cs += __shfl_xor_sync(FULL, cs, o);
    depth += cs;
    const uint32_t plain = (S == 0) ? copies : 0;
    uint32_t ps = plain;
    for (int o = 16; o; o >>= 1) ps += __shfl_xor_sync(FULL, ps, o);
    c0 += ps;
    uint32_t todo = (S != 0) ? copies : 0, maxc = todo;
    for (int o = 16; o; o >>= 1) maxc = max(maxc, __shfl_xor_sync(FULL, maxc, o));
    for (uint32_t c = 0; c < maxc; ++c) {
      const bool have = c < todo;
      uint64_t hap = 0;
      if (have) hap = mph_nrm_hap(sg, d.vars, rev ? k - c : kp_last, k, va, vb, vlo, S);
      c0 += __popc(__ballot_sync(FULL, have && hap == 0));
      unsigned pending = __ballot_sync(FULL, have && hap != 0);
      while (pending) {
        const int leader = __ffs(pending) - 1;
        const uint64_t lh = __shfl_sync(FULL, hap, leader);
        const unsigned same = __ballot_sync(FULL, have && hap == lh);
        if (lane == 0) {
          uint32_t t = 0;
          for (; t < n_keys; ++t)
            if (table[t].hap == lh) break;
          if (t == n_keys) {
            if (n_keys < K2_TABLE) {
              table[t].hap = lh;
              table[t].frame = 0;
              table[t].count = 0;
              ++n_keys;
            } else {
              raise(d, MPH_E_KEYS_PER_WINDOW);
              t = K2_TABLE - 1;
            }
          }
          table[t].count += __popc(same);
        }
        pending &= ~same;
      }
    }
  }
  if (lane == 0) {
    for (uint32_t a = 1; a < n_keys; ++a) {  // ascending haplotype value (VecMap iteration order :383)
      const MphHist key = table[a];
      uint32_t b = a;
      while (b > 0 && key.hap < table[b - 1].hap) {
        table[b] = table[b - 1];
        --b;
      }
      table[b] = key;
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
        for (uint32_t a = 0; a < n_keys; ++a) {
          d.hist[off + a] = table[a];
          d.hist_win[off + a] = code;
        }
      } else {
        raise(d, MPH_E_HIST_OVERFLOW);
        wo.n_extra = 0;
      }
    }
    d.win_out[widx] = wo;
    atomicAdd(d.sum_depth, (unsigned long long)depth);
    normal_window_out(d, sg, g, widx, vb - va, depth);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(K2_WARPS * 32) k_window_hist_wide_normal(const DeviceBatch d) {
  __shared__ MphHist table[K2_WARPS][K2_TABLE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = d.counters[CTR_OVF];
  for (uint32_t o = blockIdx.x * K2_WARPS + warp; o < n; o += gridDim.x * K2_WARPS) {
    const uint32_t code = d.ovf_list[o];
    const MphChunk ch = d.chunks[code >> 5];
    const MphSegment sg = d.segs[ch.seg];
    window_hist_warp_normal(d, sg, ch.i_first + (code & 31u), code, table[warp], lane);
  }
}

__global__ void __launch_bounds__(K2_WARPS * 32) k_window_hist_normal(const DeviceBatch d) {
  __shared__ uint64_t t_hap[K2_WARPS][KN_LANE_KEYS][32];
  __shared__ uint32_t t_cnt[K2_WARPS][KN_LANE_KEYS][32];
  __shared__ MphSegment s_seg[K2_WARPS];
  __shared__ uint32_t s_s[K2_WARPS][32], s_e[K2_WARPS][32];
  __shared__ int s_add[K2_WARPS][34], s_slope[K2_WARPS][34];
  __shared__ uint32_t s_list[K2_WARPS][K2_LIST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t chunk = d.c0 + blockIdx.x * K2_WARPS + warp;
  if (chunk >= d.c1) return;
  const MphChunk ch = d.chunks[chunk];
  if (lane < (int)(sizeof(MphSegment) / 4)) reinterpret_cast<uint32_t*>(&s_seg[warp])[lane] = reinterpret_cast<const uint32_t*>(&d.segs[ch.seg])[lane];
  s_add[warp][lane] = 0;
  s_slope[warp][lane] = 0;
  if (lane < 2) { s_add[warp][32 + lane] = 0; s_slope[warp][32 + lane] = 0; }
  __syncwarp();
  const MphSegment& sg = s_seg[warp];
  if (sg.flags & MPH_SF_REPLAY) return;  // the whole transcript goes through k_replay_normal
  const bool rev = (sg.flags & MPH_SF_REVERSE) != 0;
  const int n = (int)ch.n;
  const bool active = lane < n;
  const uint32_t i = ch.i_first + (active ? lane : 0);
  const uint32_t k = sg.k_first + i * sg.k_stride;
  const MphGeom g = mph_geom(sg, k);
  const uint32_t va = mph_var_lb(d.vars, ch.va0, ch.vb1, g.s);
  const uint32_t vb = mph_var_lb(d.vars, va, ch.vb1, g.e);
  const uint32_t nvar = vb - va;
  if (active && nvar > 64) raise(d, MPH_E_VARS_PER_WINDOW);
  const bool chunk_has_var = ch.vb1 > ch.va0;
  s_s[warp][lane] = g.s;
  s_e[warp][lane] = g.e;
  const uint32_t s0 = sg.off0 - sg.ceo;
  const uint32_t rlo = ch.rlo, rhi = ch.rhi;
  const int64_t c1_lo = (int64_t)s0 - (int64_t)sg.K;
  const int k_base = (int)(sg.k_first + ch.i_first * sg.k_stride);  // iteration number of lane 0
  uint32_t n_keys = 0;
  int c0_adj = 0;
  bool overflow = false;
  __syncwarp();
  auto add_key = [&](uint64_t hap) {
    uint32_t t = 0;
    for (; t < n_keys; ++t)
      if (t_hap[warp][t][lane] == hap) break;
    if (t == n_keys) {
      if (n_keys == KN_LANE_KEYS || d.force_wide) { overflow = true; return; }
      t_hap[warp][t][lane] = hap;
      t_cnt[warp][t][lane] = 0;
      ++n_keys;
    }
    t_cnt[warp][t][lane] += 1;
  };
  // phase B body (lane = window): copies of the listed reads, one haplotype per copy
  auto process_list = [&](uint32_t list_n) {
    if (!active || nvar == 0) return;
    for (uint32_t x = 0; x < list_n; ++x) {
      const uint32_t r = s_list[warp][x];
      const uint32_t st = d.read_start[r], en = d.read_end[r], vlo = d.read_vlo[r];
      const uint64_t S = d.call_S[r];
      if (!rev) {
        const uint32_t kp = mph_nrm_fwd_entry(sg, k, g, st, en);
        if (kp == NONE) continue;
        const uint64_t hap = mph_nrm_hap(sg, d.vars, kp, k, va, vb, vlo, S);
        if (hap) { c0_adj -= 1; add_key(hap); }
      } else {
        uint32_t kc;
        const uint32_t copies = mph_nrm_rev_copies(sg, k, g, st, en, &kc);
        for (uint32_t c = 0; c < copies; ++c) {
          const uint64_t hap = mph_nrm_hap(sg, d.vars, k - c, k, va, vb, vlo, S);
          if (hap) { c0_adj -= 1; add_key(hap); }
        }
      }
    }
  };
  // ---- phase A (lane = read)
  uint32_t list_n = 0;
  const uint32_t* ss = s_s[warp];
  const uint32_t* se = s_e[warp];
  for (uint32_t base = rlo; base < rhi; base += 32) {
    const uint32_t r = base + lane;
    bool need = false;
    if (r < rhi) {
      const uint32_t st = d.read_start[r], en = d.read_end[r];
      int ilo = 1, ihi = 0;
      if (!rev) {
        int a = 0, b = n;  // e non-decreasing: windows with e <= en form a prefix
        while (a < b) { const int m = (a + b) >> 1; if (se[m] <= en) a = m + 1; else b = m; }
        ihi = a - 1;
        if (st <= s0) {
          ilo = ((int64_t)st >= c1_lo) ? 0 : n;
        } else if (st <= sg.off0) {
          ilo = n;
        } else {
          const uint32_t k_ins = st - sg.off0;
          const uint32_t t = k_ins > sg.k_first ? (k_ins - sg.k_first + sg.k_stride - 1) / sg.k_stride : 0;
          ilo = t > ch.i_first ? (int)min(t - ch.i_first, (uint32_t)n) : 0;
        }
        if (ilo <= ihi) {
          atomicAdd(&s_add[warp][ilo], 1);
          atomicAdd(&s_add[warp][ihi + 1], -1);
          need = true;
        }
      } else {
        const uint32_t kc = normal_rev_kc(sg, st, en);
        const uint32_t t = kc > sg.k_first ? (kc - sg.k_first + sg.k_stride - 1) / sg.k_stride : 0;
        ilo = t > ch.i_first ? (int)min(t - ch.i_first, (uint32_t)n) : 0;
        int a = 0, b = n;  // s non-increasing: windows with s >= st form a prefix
        while (a < b) { const int m = (a + b) >> 1; if (ss[m] >= st) a = m + 1; else b = m; }
        ihi = a - 1;
        a = 0; b = n;      // windows with s > st
        while (a < b) { const int m = (a + b) >> 1; if (ss[m] > st) a = m + 1; else b = m; }
        const int iramp = a - 1;
        if (ilo <= iramp) {  // copies(j) = stride * j + (k_base - kc + 1)
          const int cst = k_base - (int)kc + 1;
          atomicAdd(&s_slope[warp][ilo], (int)sg.k_stride);
          atomicAdd(&s_slope[warp][iramp + 1], -(int)sg.k_stride);
          atomicAdd(&s_add[warp][ilo], cst);
          atomicAdd(&s_add[warp][iramp + 1], -cst);
          need = true;
        }
        const int elo = max(ilo, iramp + 1);
        if (elo <= ihi) {  // window start == read start: the older copies have just been dropped
          atomicAdd(&s_add[warp][elo], 1);
          atomicAdd(&s_add[warp][ihi + 1], -1);
          need = true;
        }
      }
      need = need && chunk_has_var && (d.call_flags[r] & 1u) && d.call_S[r] != 0;
    }
    const unsigned nm = __ballot_sync(FULL, need);
    if (nm) {
      if (list_n + __popc(nm) > K2_LIST) {
        __syncwarp();
        process_list(list_n);
        list_n = 0;
        __syncwarp();
      }
      if (need) s_list[warp][list_n + __popc(nm & ((1u << lane) - 1))] = r;
      list_n += __popc(nm);
      __syncwarp();
    }
  }
  __syncwarp();
  process_list(list_n);
  // ---- phase B: prefix sums of the two difference arrays
  int run = s_add[warp][lane], slope = s_slope[warp][lane];
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(FULL, run, o), z = __shfl_up_sync(FULL, slope, o);
    if (lane >= o) { run += y; slope += z; }
  }
  const uint32_t depth = (uint32_t)(run + slope * lane);
  const uint32_t c0 = (uint32_t)((int)depth + c0_adj);
  for (uint32_t a = 1; a < n_keys; ++a) {  // ascending haplotype value
    const uint64_t kh = t_hap[warp][a][lane];
    const uint32_t kc = t_cnt[warp][a][lane];
    uint32_t b = a;
    while (b > 0 && kh < t_hap[warp][b - 1][lane]) {
      t_hap[warp][b][lane] = t_hap[warp][b - 1][lane];
      t_cnt[warp][b][lane] = t_cnt[warp][b - 1][lane];
      --b;
    }
    t_hap[warp][b][lane] = kh;
    t_cnt[warp][b][lane] = kc;
  }
  const bool ovf = active && overflow;
  const uint32_t mine = (active && !ovf) ? n_keys : 0;
  uint32_t incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += y;
  }
  const uint32_t total = __shfl_sync(FULL, incl, 31);
  uint32_t base_off = 0;
  if (lane == 0 && total) base_off = atomicAdd(&d.counters[CTR_HIST], total);
  base_off = __shfl_sync(FULL, base_off, 0);
  const bool fits = base_off + total <= d.hist_cap;
  if (lane == 0 && total && !fits) raise(d, MPH_E_HIST_OVERFLOW);
  if (active && !ovf) {
    MphWinOut wo;
    wo.depth = depth;
    wo.c0 = c0;
    wo.n_extra = fits ? mine : 0;
    wo.extra_off = base_off + incl - mine;
    if (fits)
      for (uint32_t a = 0; a < mine; ++a) {
        MphHist h;
        h.hap = t_hap[warp][a][lane]; h.frame = 0; h.count = t_cnt[warp][a][lane];
        d.hist[wo.extra_off + a] = h;
        d.hist_win[wo.extra_off + a] = (chunk << 5) | (uint32_t)lane;
      }
    const uint32_t widx = sg.win_base + i;
    d.win_out[widx] = wo;
    normal_window_out(d, sg, g, widx, nvar, depth);
  }
  if (ovf) {
    const uint32_t o = atomicAdd(&d.counters[CTR_OVF], 1u);
    d.ovf_list[o] = (chunk << 5) | (uint32_t)lane;
  }
  unsigned long long dsum = (active && !ovf) ? depth : 0;
  for (int o = 16; o; o >>= 1) dsum += __shfl_down_sync(FULL, dsum, o);
  if (lane == 0 && dsum) atomicAdd(d.sum_depth, dsum);
}

// thread per haplotype key != 0: the sequence walk of the normal-mode print_haplotypes (:403-507);
// every haplotype of a variant window is written by the host, so every sequence is kept
__global__ void __launch_bounds__(128) k_assemble_normal(const DeviceBatch d) {
  const uint32_t n = min(d.counters[CTR_HIST], d.hist_cap);
  uint8_t seq[MAX_SEQ_CAP];
  const uint32_t cap = d.seq_cap;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    const MphHist key = d.hist[x];
    const uint32_t code = d.hist_win[x];
    const MphChunk ch = d.chunks[code >> 5];
    const MphSegment sg = d.segs[ch.seg];
    const uint32_t i = ch.i_first + (code & 31u);
    const uint32_t k = sg.k_first + i * sg.k_stride;
    const MphGeom g = mph_geom(sg, k);
    const uint32_t va = mph_var_lb(d.vars, ch.va0, ch.vb1, g.s);
    const uint32_t vb = mph_var_lb(d.vars, va, ch.vb1, g.e);
    const uint32_t depth = d.win_out[sg.win_base + i].depth;
    MphHap out;
    uint32_t err;
    if (!(sg.flags & MPH_SF_REPLAY)) {
      err = mph_nrm_assemble(sg, g, d.vars, va, vb, d.ref, d.ins_bytes, key.hap, key.count == depth, seq, cap, &out);
    } else {
      MphVar cols[MPH_RP_MAXCOLS];
      const uint32_t off = d.win_voff[sg.win_base + i];
      const uint32_t ncol = off == NONE ? 0u : min(d.vlist[off], (uint32_t)MPH_RP_MAXCOLS);
      for (uint32_t j = 0; j < ncol; ++j) cols[j] = d.vars[d.vlist[off + 1 + j]];
      err = mph_nrm_assemble(sg, g, cols, 0, ncol, d.ref, d.ins_bytes, key.hap, key.count == depth, seq, cap, &out);
    }
    if (out.seq_len <= cap) {
      const uint32_t t0 = d.tx_id_off[sg.tx];
      out.id64 = mph_record_id64(seq, out.seq_len, d.tx_id_bytes + t0, d.tx_id_off[sg.tx + 1] - t0, g.s);
      out.flags |= MPH_NF_ID;
    }
    const uint32_t off = atomicAdd(&d.counters[CTR_SEQ], cap);
    if (off + cap <= d.seq_cap_bytes) {
      const uint32_t sl = out.seq_len < cap ? out.seq_len : cap;
      for (uint32_t t = 0; t < sl; ++t) d.seq[off + t] = seq[t];
      out.seq_off = off;
      out.flags |= MPH_NF_SEQ;
    } else {
      err |= MPH_E_SEQ_OVERFLOW;
    }
    d.hapx[x] = out;
    raise(d, err);
  }
}

// ------------------------------------------------------------------ K4: stable compaction
constexpr int SCAN_THREADS = 1024;

__global__ void __launch_bounds__(SCAN_THREADS) k_flag_count(const DeviceBatch d) {
  const uint32_t w = d.w0 + blockIdx.x * SCAN_THREADS + threadIdx.x;
  const int f = (w < d.w1) ? d.win_flag[w] : 0;
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
  const uint32_t w = d.w0 + blockIdx.x * SCAN_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool f = (w < d.w1) && d.win_flag[w];
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
    if (d.win_voff) d.iw_voff[pos] = d.win_voff[w];
  }
}

// ------------------------------------------------------------------ K5: statistics
__global__ void __launch_bounds__(256) k_live_depth(const DeviceBatch d) {
  const uint32_t chunk = d.c0 + blockIdx.x * 8 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long v = 0;
  if (chunk < d.c1) {
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
  if (d.vr1 > d.vr0) k_allele_call<<<(d.vr1 - d.vr0 + 255) / 256, 256, 0, st>>>(d);
}
void launch_window_hist(const DeviceBatch& d, cudaStream_t st) {
  if (d.c1 <= d.c0) return;
  const uint32_t nc = d.c1 - d.c0;
  if (d.mode == 1) {
    k_window_hist_normal<<<(nc + K2_WARPS - 1) / K2_WARPS, K2_WARPS * 32, 0, st>>>(d);
    k_window_hist_wide_normal<<<148, K2_WARPS * 32, 0, st>>>(d);
    return;
  }
  k_window_hist<<<(nc + K2_WARPS - 1) / K2_WARPS, K2_WARPS * 32, 0, st>>>(d);
  // windows with more distinct haplotypes than a lane table holds (rare): one warp per window
  k_window_hist_wide<<<148, K2_WARPS * 32, 0, st>>>(d);
}
void launch_replay(const DeviceBatch& d, cudaStream_t st) {
  if (d.rp1 > d.rp0 && d.mode == 1) k_replay_normal<<<(d.rp1 - d.rp0 + 1) / 2, 64, 0, st>>>(d);
  else if (d.rp1 > d.rp0) k_replay<<<(d.rp1 - d.rp0 + RP_WARPS - 1) / RP_WARPS, RP_WARPS * 32, 0, st>>>(d);
}
void launch_assemble(const DeviceBatch& d, cudaStream_t st) {
  if (d.c1 > d.c0 && d.mode == 1) k_assemble_normal<<<148 * 8, 128, 0, st>>>(d);
  else if (d.c1 > d.c0) k_assemble<<<148 * 8, 128, 0, st>>>(d);  // grid-stride over the key arena (its size lives on the device)
}
void launch_compact(const DeviceBatch& d, cudaStream_t st) {
  const uint32_t nb = (d.w1 - d.w0 + SCAN_THREADS - 1) / SCAN_THREADS;
  if (nb) k_flag_count<<<nb, SCAN_THREADS, 0, st>>>(d);
  k_block_scan<<<1, SCAN_THREADS, 0, st>>>(d, nb);
  if (nb) k_scatter<<<nb, SCAN_THREADS, 0, st>>>(d);
}
void launch_live_depth(const DeviceBatch& d, cudaStream_t st) {
  if (d.c1 > d.c0) k_live_depth<<<(d.c1 - d.c0 + 7) / 8, 256, 0, st>>>(d);
}
int kernel_launch_count() { return 7; }

}  // namespace mphk
