// phase_core.h — the per-item arithmetic of the phasing kernels, written once as
// host+device inline functions. The CUDA kernels (csrc/kernels/*.cu) wrap these in the
// parallel structure (warp per window, match_any histogram, compaction); the test-only CPU
// emulator (tests/emu) wraps the same functions in plain loops so the closed forms below can be
// diffed against the oracle without a GPU.
//
// What is restated here, with the reference lines it replaces (src/microphasing.rs):
//   mph_geom          window 4-tuple (splice_side_offset, splice_end, splice_gap, splice_pos)  :1050-1111
//   mph_call_read     supports_variant :95-139 + bad_quality :78-93 for every variant inside a read
//   mph_fwd_state /   closed form of the ObservationMatrix bookkeeping — cleanup_reads :259-278,
//   mph_rev_*         shrink_left :220-229, push_read :297-343, extend_right :232-256,
//                     update_haplotype :157-197 — for one (read, window) pair
//   mph_assemble      the haplotype sequence walk of print_haplotypes :458-603
//   mph_has_stop      has_stop_codon :42-76
#pragma once
#include "layout.h"

typedef struct {
  uint32_t s, e, gap, spos;
} MphGeom;

// :1050-1111. k-th iteration of a segment (read_through is always false once `valid` held, :1041-1045)
MPH_HD MphGeom mph_geom(const MphSegment& g, uint32_t k) {
  const bool rev = (g.flags & MPH_SF_REVERSE) != 0, shrt = (g.flags & MPH_SF_SHORT) != 0;
  const uint32_t ceo = g.ceo, ewl = g.ewl;
  MphGeom o;
  if (!rev) {
    const uint32_t off = g.off0 + k;
    const uint32_t rest = g.exon_end - (off + ewl);
    const bool first = k == 0, last = rest < 3;
    if (shrt || (first && last)) { o.s = off - ceo; o.e = off + ewl + rest; o.gap = ceo + rest; o.spos = 2; }
    else if (first) { o.s = off - ceo; o.e = off + ewl; o.gap = ceo; o.spos = 1; }
    else if (last) { o.s = off; o.e = off + ewl + rest; o.gap = rest; o.spos = 0; }
    else { o.s = off; o.e = off + ewl; o.gap = 0; o.spos = 0; }
  } else {
    const uint32_t off = g.off0 - k;
    const uint32_t rest = off - g.exon_start;
    const bool first = k == 0, last = rest < 3;
    if (shrt) { o.s = off - rest; o.e = off + ewl + ceo; o.gap = ceo + rest; o.spos = 2; }
    else if (first) { o.s = off; o.e = off + ewl + ceo; o.gap = ceo; o.spos = 0; }
    else if (last) { o.s = off - rest; o.e = off + ewl; o.gap = rest; o.spos = 1; }
    else { o.s = off; o.e = off + ewl; o.gap = 0; o.spos = 0; }
  }
  return o;
}

// first variant index in [lo, hi) with pos >= p
MPH_HD uint32_t mph_var_lb(const MphVar* vars, uint32_t lo, uint32_t hi, uint32_t p) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (vars[mid].pos < p) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// first index in [lo, hi) with a[idx] >= p
MPH_HD uint32_t mph_u32_lb(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t p) {
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (a[mid] < p) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// bits j with ia <= vlo + j < ib
MPH_HD uint64_t mph_range_mask(uint32_t ia, uint32_t ib, uint32_t vlo) {
  if (ib <= ia) return 0;
  int64_t lo = (int64_t)ia - (int64_t)vlo, hi = (int64_t)ib - (int64_t)vlo;
  if (lo < 0) lo = 0;
  if (hi > 64) hi = 64;
  if (hi <= lo) return 0;
  const uint64_t m = (hi - lo >= 64) ? ~(uint64_t)0 : (((uint64_t)1 << (hi - lo)) - 1);
  return m << lo;
}

MPH_HD uint64_t mph_bitrev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __brevll(x);
#else
  x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
  return (x >> 32) | (x << 32);
#endif
}

// ------------------------------------------------------------------ K1: allele calls
// CigarStringView::read_pos(ref_pos, false, false) over BAM-encoded ops; n_cig == 0 means one M.
// Returns 1 and *qpos on Some, 0 on None / Err (the caller treats both as "no support", :106-110).
MPH_HD int mph_read_pos(const uint32_t* cig, uint32_t n_cig, uint32_t l_seq, uint32_t start, uint32_t ref_pos,
                        uint32_t* qpos_out) {
  if (n_cig == 0) {
    if (ref_pos >= start && ref_pos - start < l_seq) { *qpos_out = ref_pos - start; return 1; }
    return 0;
  }
  uint32_t j = 0;
  for (uint32_t i = 0; i < n_cig; ++i) {
    const uint32_t op = cig[i] & 15u;
    if (op == 0 || op == 8 || op == 7 || op == 1 || op == 4) { j = i; break; }  // M X = I S
    if (op == 2 || op == 3) return 0;                                          // D / N first: Err
    if (op == 5 && i > 0 && i + 1 < n_cig) return 0;                           // H in the middle: Err
    if ((op == 6 || op == 5) && i + 1 == n_cig) return 0;                      // only P / H: None
  }
  uint64_t rpos = start, qpos = 0;
  while (rpos <= ref_pos && j < n_cig) {
    const uint32_t op = cig[j] & 15u;
    const uint64_t l = cig[j] >> 4;
    if (op == 0 || op == 8 || op == 7) {
      if (rpos + l > ref_pos) { *qpos_out = (uint32_t)(qpos + (ref_pos - rpos)); return 1; }
      rpos += l; qpos += l; ++j;
    } else if (op == 4 || op == 1) { qpos += l; ++j; }
    else if (op == 3 || op == 2) { rpos += l; ++j; }
    else if (op == 6) { ++j; }
    else { return 0; }  // H: Err in the middle, None at the end
  }
  return 0;
}

// Packed read record (16-B aligned in the `bases` arena), what K1 needs of one alignment:
//   byte 0  format: bit 0 = bases are 2-bit codes (A C G T only; else BAM 4-bit codes), bit 1 = the positions with
//           qual < 10 are a list (else a bitmask); bit 2 = column record of an ungapped read (one byte per variant
//           inside the read, see mph_call_read) - then nothing else follows
//   byte 1  list length (format bit 1)
//   then    the bases: ceil(l_seq/4) or ceil(l_seq/2) B, first base in the high bits
//   then    the low-quality positions: `list length` bytes (one position each), or ceil(l_seq/8) B with bit i <=> base i
MPH_HD uint32_t mph_rec_bases_bytes(uint8_t fmt, uint32_t l_seq) { return (fmt & 1u) ? (l_seq + 3u) >> 2 : (l_seq + 1u) >> 1; }
// qual[i] < 10 (:82-84,99-101)
MPH_HD bool mph_rec_low(const uint8_t* rec, uint32_t l_seq, uint32_t i) {
  const uint8_t fmt = rec[0];
  const uint8_t* lq = rec + 2 + mph_rec_bases_bytes(fmt, l_seq);
  if (fmt & 2u) {
    const uint32_t n = rec[1];
    for (uint32_t x = 0; x < n; ++x)
      if (lq[x] == i) return true;
    return false;
  }
  return (lq[i >> 3] >> (i & 7u)) & 1u;
}
// BAM 4-bit code of base i
MPH_HD uint8_t mph_rec_base4(const uint8_t* rec, uint32_t i) {
  const uint8_t* sq = rec + 2;
  if (rec[0] & 1u) return (uint8_t)(1u << ((sq[i >> 2] >> (6u - 2u * (i & 3u))) & 3u));
  const uint8_t b = sq[i >> 1];
  return (i & 1u) ? (uint8_t)(b & 15u) : (uint8_t)(b >> 4);
}

// supports_variant / bad_quality for every variant with pos in [start, end) of one read.
// `bases` points at the read's packed record (above).
MPH_HD MphCall mph_call_read(const MphRead& r, const uint8_t* bases, const uint32_t* cig, const MphVar* vars, bool use_qual = true) {
  MphCall c;
  c.S = 0;
  c.B = 0;
  const uint32_t nv = r.nv;
  if (nv == 0) return c;
  if (bases[0] & 4u) {
    // column record of an ungapped read: byte 1 + j holds the base at variant j's position (BAM 4-bit code), bit 4 =
    // its quality is below 10, bit 5 = the position is inside the read. Query position = reference offset here, so the
    // quality test (:82-84,99-101) and read_pos (:106-110) index the same base; insertions / deletions need a CIGAR
    // operation of their kind (:120-135), which a single-M read does not have.
    for (uint32_t j = 0; j < nv; ++j) {
      const MphVar v = vars[r.vlo + j];
      if (v.kind != MPH_SNV) continue;
      const uint8_t col = bases[1 + j];
      if (!(col & 32u)) continue;
      if (use_qual && (col & 16u)) c.B |= (uint64_t)1 << j;
      else if ((col & 15u) == v.alt4) c.S |= (uint64_t)1 << j;
    }
    return c;
  }
  for (uint32_t j = 0; j < nv; ++j) {
    const MphVar v = vars[r.vlo + j];
    bool sup = false;
    if (v.kind == MPH_SNV) {
      const uint32_t rel = v.pos - r.start;  // raw reference offset indexes the *query* qualities (:82-84,99-101)
      bool low = false;
      if (use_qual && rel < r.l_seq) low = mph_rec_low(bases, r.l_seq, rel);  // normal mode has no base-quality test (normal_microphasing.rs:43-52)
      if (low) {
        c.B |= (uint64_t)1 << j;
      } else {
        uint32_t q;
        if (mph_read_pos(cig, r.n_cig, r.l_seq, r.start, v.pos, &q) && q < r.l_seq) {
          sup = mph_rec_base4(bases, q) == v.alt4;
        }
      }
    } else {
      const uint32_t want = v.kind == MPH_INS ? 1u : 2u;  // BAM op codes: I = 1, D = 2
      for (uint32_t i = 0; i < r.n_cig; ++i)
        if ((cig[i] & 15u) == want && (cig[i] >> 4) == v.len) { sup = true; break; }
    }
    if (sup) c.S |= (uint64_t)1 << j;
  }
  return c;
}

// windows whose haplotype list a splice-junction merge can read (:1401-1405,1497-1510): the first and the last window
// of an exon (the last-but-one when the next exon stores its only window on the "previous exon" side) - but only at
// junctions that can produce a record, i.e. with a variant in one of the two windows (the packer decides; a merge of
// reference-only lists writes nothing, :1801-1810,1887) - and every window of a short exon or when frameshifts keep
// several reading frames alive
MPH_HD bool mph_is_boundary(const MphSegment& g, uint32_t i) {
  if (g.flags & (MPH_SF_HAS_FS | MPH_SF_SHORT)) return true;
  if (i == 0 && (g.flags & MPH_SF_JOIN_HEAD)) return true;
  if (!(g.flags & MPH_SF_JOIN_TAIL)) return false;
  if (i + 1 == g.n_win) return true;
  return (g.flags & MPH_SF_KEEP_PENULT) && i + 2 == g.n_win;
}

// Reads that can be observations of window gk: start <= s, and start not further left than the
// candidate range of the reference (:1191-1249) nor than the longest alignment allows.
MPH_HD void mph_candidate_range(const MphSegment& g, const uint32_t* read_start, const MphGeom& gk, uint32_t* rlo, uint32_t* rhi) {
  const bool rev = (g.flags & MPH_SF_REVERSE) != 0;
  int64_t lo = rev ? (int64_t)gk.s - (int64_t)g.K : (int64_t)(g.off0 - g.ceo) - (int64_t)g.K;
  const int64_t lo2 = (int64_t)gk.e - (int64_t)g.max_span;
  if (lo2 > lo) lo = lo2;
  if (lo < 0) lo = 0;
  *rlo = mph_u32_lb(read_start, g.read_lo, g.read_hi, (uint32_t)lo);
  *rhi = mph_u32_lb(read_start, *rlo, g.read_hi, gk.s + 1u);
}

// ------------------------------------------------------------------ K2: one (read, window) pair
typedef struct {
  uint32_t member;  // the read is an observation of the matrix at this window (counts towards depth)
  uint32_t bad;     // obs.bad_qual
  uint64_t hap;     // obs.haplotype
  uint32_t frame;   // obs.frame.0 | (obs.frame.1 != 0) << 31
} MphPair;

// bit j of the result <-> window variant j (list order), from the read's S mask
MPH_HD uint64_t mph_window_bits(uint64_t S, uint32_t vlo, uint32_t va, uint32_t n) {
  if (n == 0 || va < vlo || va - vlo >= 64) return 0;
  uint64_t b = S >> (va - vlo);
  if (n < 64) b &= (((uint64_t)1 << n) - 1);
  return b;
}

// obs.frame accumulators over the variants the observation has evaluated so far (:172-174,188)
MPH_HD uint32_t mph_frames(const MphVar* vars, uint32_t ia, uint32_t ib, uint32_t vlo, uint64_t S) {
  uint32_t f0 = 0, f1nz = 0;
  for (uint32_t j = ia; j < ib; ++j) {
    const uint32_t fs = (vars[j].flags & MPH_VF_FS_MASK) >> MPH_VF_FS_SHIFT;
    if (fs) {
      if (vars[j].pos != 0) f1nz = 1;
      if (j >= vlo && j - vlo < 64 && ((S >> (j - vlo)) & 1)) f0 += fs;
    }
  }
  return f0 | (f1nz << 31);
}

// Forward strand: the read is offered exactly once per segment — at iteration 0 if
// start in [s0-K, s0], else at the iteration whose window starts at `start` (:1227-1248); a read
// that turns bad_qual while the variants retained from the previous window are evaluated is dropped
// for good (:324-335); afterwards badness is sticky (:192-195) and the read stays while
// end_pos >= splice_end (:259-273).
MPH_HD MphPair mph_fwd_state(const MphSegment& g, const MphVar* vars, uint32_t k, const MphGeom& gk, uint32_t va, uint32_t vb,
                             uint32_t start, uint32_t end, uint32_t vlo, uint64_t S, uint64_t B) {
  MphPair o;
  o.member = 0; o.bad = 0; o.hap = 0; o.frame = 0;
  if (end < gk.e) return o;
  const uint32_t s0 = g.off0 - g.ceo;
  uint32_t k_ins;
  if (start <= s0) {
    if ((int64_t)start < (int64_t)s0 - (int64_t)g.K) return o;
    k_ins = 0;
  } else {
    if (start <= g.off0) return o;  // starts in (s0, off0] are never offered when ceo > 0
    k_ins = start - g.off0;
    if (k_ins > k) return o;
  }
  const uint64_t Bx = B | (S & mph_range_mask(g.sl_va, g.sl_vb, vlo));
  const bool has_fs = (g.flags & MPH_SF_HAS_FS) != 0;
  if (Bx || has_fs) {
    const MphGeom gi = k_ins == k ? gk : mph_geom(g, k_ins);
    const uint32_t ia = mph_var_lb(vars, g.var_lo, g.var_hi, gi.s);
    if (Bx) {
      if (k_ins > 0) {
        const uint32_t ib = mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, k_ins - 1).e);
        if (Bx & mph_range_mask(ia, ib, vlo)) return o;  // rejected at push
      }
      o.bad = (Bx & mph_range_mask(ia, vb, vlo)) != 0;
    }
    if (has_fs) o.frame = mph_frames(vars, ia, vb, vlo, S);
  }
  o.member = 1;
  if (!o.bad) {
    const uint32_t n = vb - va;
    const uint64_t bits = mph_window_bits(S, vlo, va, n);
    o.hap = n ? (mph_bitrev64(bits) >> (64 - n)) : 0;  // newest variant = bit 0 (:324,248)
  }
  return o;
}

// Reverse strand: every read with start in [s-K, s] is re-offered at every iteration (:1191-1226);
// it enters at the first iteration where it encloses the window and none of the variants retained
// from the previous window is bad for it; `contains` (:281-294) keeps it from entering twice.
// Returns the entry iteration or 0xFFFFFFFF.
MPH_HD uint32_t mph_rev_entry(const MphSegment& g, const MphVar* vars, uint32_t k, uint32_t start, uint32_t end, uint32_t vlo,
                              uint64_t Bx) {
  const uint64_t lim = (uint64_t)start + g.K;
  uint32_t kc = g.off0 > lim ? (uint32_t)(g.off0 - lim) : 0;
  if (mph_geom(g, 0).e > end) {
    const uint64_t t = (uint64_t)g.off0 + g.ewl;
    const uint32_t k2 = t > end ? (uint32_t)(t - end) : 1;
    if (k2 > kc) kc = k2;
    if (kc == 0) kc = 1;
  }
  // in the last three iterations of the exon s(k') = exon.start < off0 - k' (:1105-1106), so the
  // start-range condition can become true up to two iterations before the linear guess
  kc = kc >= 2 ? kc - 2 : 0;
  for (; kc <= k; ++kc) {
    const MphGeom gc = mph_geom(g, kc);
    if (gc.s <= lim && gc.e <= end && start <= gc.s) break;
  }
  if (kc > k) return 0xFFFFFFFFu;
  if (!Bx) return kc;
  for (uint32_t ke = kc; ke <= k; ++ke) {
    if (ke == 0) return 0;
    const uint32_t ia = mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, ke - 1).s);
    const uint32_t ib = mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, ke).e);
    if (!(Bx & mph_range_mask(ia, ib, vlo))) return ke;
  }
  return 0xFFFFFFFFu;
}

MPH_HD MphPair mph_rev_state(const MphSegment& g, const MphVar* vars, uint32_t k, const MphGeom& gk, uint32_t va, uint32_t vb,
                             uint32_t start, uint32_t end, uint32_t vlo, uint64_t S, uint64_t B, uint32_t ke) {
  MphPair o;
  o.member = 1; o.bad = 0; o.hap = 0; o.frame = 0;
  (void)end; (void)k;
  const uint64_t Bx = B | (S & mph_range_mask(g.sl_va, g.sl_vb, vlo));
  const bool has_fs = (g.flags & MPH_SF_HAS_FS) != 0;
  if (Bx || has_fs) {
    const uint32_t ib = mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, ke).e);
    if (Bx) o.bad = (Bx & mph_range_mask(va, ib, vlo)) != 0;
    if (has_fs) o.frame = mph_frames(vars, va, ib, vlo, S);
  }
  if (!o.bad) o.hap = mph_window_bits(S, vlo, va, vb - va);  // bit j <-> variants[j] (:486-489)
  (void)gk; (void)start;
  return o;
}

// ------------------------------------------------------------------ record ids (SHA-1, FIPS 180-1)
// The reference names a record by the first 15 hex digits of sha1(format!("{:?}{}{}", seq, transcript.id, offset))
// (:667-675; `{:?}` of a byte vector prints "[65, 84, 71]"). The message is generated and hashed on the fly,
// 64 bytes at a time, so that no per-thread message buffer is needed.
typedef struct {
  uint32_t h[5];
  uint32_t w[16];  // message words of the current block (big-endian), written one whole word at a time
  uint32_t acc;    // the bytes of the word being filled
  uint32_t len;    // bytes pushed so far
} MphSha1;

MPH_HD uint32_t mph_rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }

MPH_HD void mph_sha1_init(MphSha1* s) {
  s->h[0] = 0x67452301u; s->h[1] = 0xEFCDAB89u; s->h[2] = 0x98BADCFEu; s->h[3] = 0x10325476u; s->h[4] = 0xC3D2E1F0u;
  s->acc = 0;
  s->len = 0;
}

MPH_HD void mph_sha1_block(MphSha1* s) {
  uint32_t w[16];
  for (int i = 0; i < 16; ++i) w[i] = s->w[i];
  uint32_t a = s->h[0], b = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4];
#define MPH_SHA1_ROUND(i_, f_, k_)                                                                      \
  do {                                                                                                  \
    uint32_t wi_;                                                                                       \
    if ((i_) < 16) wi_ = w[(i_)];                                                                       \
    else {                                                                                              \
      wi_ = mph_rol(w[((i_) + 13) & 15] ^ w[((i_) + 8) & 15] ^ w[((i_) + 2) & 15] ^ w[(i_) & 15], 1);   \
      w[(i_) & 15] = wi_;                                                                               \
    }                                                                                                   \
    const uint32_t t_ = mph_rol(a, 5) + (f_) + e + (k_) + wi_;                                          \
    e = d; d = c; c = mph_rol(b, 30); b = a; a = t_;                                                    \
  } while (0)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < 20; ++i) MPH_SHA1_ROUND(i, (b & c) | (~b & d), 0x5A827999u);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 20; i < 40; ++i) MPH_SHA1_ROUND(i, b ^ c ^ d, 0x6ED9EBA1u);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 40; i < 60; ++i) MPH_SHA1_ROUND(i, (b & c) | (b & d) | (c & d), 0x8F1BBCDCu);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 60; i < 80; ++i) MPH_SHA1_ROUND(i, b ^ c ^ d, 0xCA62C1D6u);
#undef MPH_SHA1_ROUND
  s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e;
}

// One byte: shifted into the word accumulator; a full word is stored once (one local-memory store per four bytes on the
// device - the first version did a read-modify-write of w[] per byte, a third of K3's samples in the round-2 ncu capture).
MPH_HD void mph_sha1_byte(MphSha1* s, uint8_t v) {
  s->acc = (s->acc << 8) | v;
  s->len += 1;
  if ((s->len & 3u) == 0) {
    s->w[((s->len >> 2) - 1u) & 15u] = s->acc;
    if ((s->len & 63u) == 0) mph_sha1_block(s);
  }
}

MPH_HD void mph_sha1_decimal(MphSha1* s, uint32_t v) {
  uint8_t dig[10];
  int n = 0;
  do { dig[n++] = (uint8_t)('0' + v % 10u); v /= 10u; } while (v);
  while (n) mph_sha1_byte(s, dig[--n]);
}

// decimal rendering of one byte (0..255) without a digit buffer
MPH_HD void mph_sha1_decimal_u8(MphSha1* s, uint32_t v) {
  if (v >= 100u) {
    const uint32_t h = v >= 200u ? 2u : 1u;
    mph_sha1_byte(s, (uint8_t)('0' + h));
    v -= 100u * h;
    const uint32_t t = (v * 205u) >> 11;  // v / 10 for v < 100
    mph_sha1_byte(s, (uint8_t)('0' + t));
    mph_sha1_byte(s, (uint8_t)('0' + (v - 10u * t)));
  } else if (v >= 10u) {
    const uint32_t t = (v * 205u) >> 11;
    mph_sha1_byte(s, (uint8_t)('0' + t));
    mph_sha1_byte(s, (uint8_t)('0' + (v - 10u * t)));
  } else {
    mph_sha1_byte(s, (uint8_t)('0' + v));
  }
}

// padding, length and the last block(s); returns the leading 64 bits of the digest
MPH_HD uint64_t mph_sha1_finish64(MphSha1* s) {
  const uint32_t bits = s->len * 8u;  // messages are far below 2^29 bytes
  mph_sha1_byte(s, 0x80);
  while (s->len & 3u) mph_sha1_byte(s, 0);
  uint32_t idx = (s->len >> 2) & 15u;  // next word of the block (0: the block was just hashed)
  if (idx > 14u) {                     // no room for the 64-bit length: pad this block out and start another
    s->w[15] = 0;
    mph_sha1_block(s);
    idx = 0;
  }
  for (uint32_t i = idx; i < 15u; ++i) s->w[i] = 0;
  s->w[15] = bits;
  mph_sha1_block(s);
  return ((uint64_t)s->h[0] << 32) | s->h[1];
}

// leading 64 bits of sha1(format!("{:?}{}{}", seq, transcript_id, offset))
MPH_HD uint64_t mph_record_id64(const uint8_t* seq, uint32_t n, const uint8_t* tx_id, uint32_t tx_len, uint32_t offset) {
  MphSha1 s;
  mph_sha1_init(&s);
  mph_sha1_byte(&s, '[');
  for (uint32_t i = 0; i < n; ++i) {
    if (i) { mph_sha1_byte(&s, ','); mph_sha1_byte(&s, ' '); }
    mph_sha1_decimal_u8(&s, seq[i]);
  }
  mph_sha1_byte(&s, ']');
  for (uint32_t i = 0; i < tx_len; ++i) mph_sha1_byte(&s, tx_id[i]);
  mph_sha1_decimal(&s, offset);
  return mph_sha1_finish64(&s);
}

// ------------------------------------------------------------------ K3: haplotype assembly
MPH_HD bool mph_is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }
MPH_HD uint8_t mph_lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }
MPH_HD uint8_t mph_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// has_stop_codon :42-76 — case-sensitive, forward scans codons from the front, reverse from the back
MPH_HD bool mph_has_stop(const uint8_t* p, uint32_t n, bool forward) {
  if (n < 3) return false;
  if (forward) {
    for (uint32_t c = 0; c + 3 <= n; c += 3) {
      if (p[c] == 'T' && ((p[c + 1] == 'G' && p[c + 2] == 'A') || (p[c + 1] == 'A' && (p[c + 2] == 'G' || p[c + 2] == 'A')))) return true;
    }
    return false;
  }
  uint32_t c = n - 3;
  for (;;) {
    if (p[c + 2] == 'A' && ((p[c] == 'T' && p[c + 1] == 'C') || (p[c] == 'C' && p[c + 1] == 'T') || (p[c] == 'T' && p[c + 1] == 'T'))) return true;
    if (c < 3) return false;
    c -= 3;
  }
}

// neopeptide slice of the assembled sequence (:686-693): [a, b)
MPH_HD void mph_neo_slice(const MphGeom& g, uint32_t seq_len, uint32_t window_len, bool insertion, uint32_t* a, uint32_t* b) {
  const uint32_t twl = seq_len < window_len ? seq_len : window_len;  // this_window_len :651-654
  if (g.spos == 1) { *a = g.gap < seq_len ? g.gap : seq_len; *b = seq_len; }
  else if (g.spos == 0) { *a = 0; *b = insertion ? seq_len : twl; }
  else { *a = 0; *b = seq_len; }
}

// Walk of print_haplotypes :458-603 for one haplotype. seq / germ must hold `cap` bytes each;
// bytes beyond cap are dropped and MPH_HF_OVERFLOW is raised. Returns MPH_E_* error bits.
MPH_HD uint32_t mph_assemble(const MphSegment& g, const MphGeom& gk, const MphVar* vars, uint32_t va, uint32_t vb,
                             const uint8_t* ref_arena, const uint8_t* ins_arena, uint64_t hap, uint8_t* seq, uint8_t* germ,
                             uint32_t cap, MphHap* out) {
  const bool rev = (g.flags & MPH_SF_REVERSE) != 0;
  const uint32_t n = vb - va;
  uint32_t err = 0, sl = 0, gl = 0, flags = 0, n_var = 0, n_som = 0, n_prof = 0, brk = 0;
  uint64_t profile = 0;
  const uint8_t* ref = ref_arena + g.ref_off;
#define MPH_REF(i_, dst_)                                          \
  do {                                                             \
    const uint64_t ri_ = (uint64_t)(i_) - g.ref_pos0;              \
    if ((i_) < g.ref_pos0 || ri_ >= g.ref_len) { flags |= MPH_HF_REFRANGE; dst_ = 'N'; } \
    else dst_ = ref[ri_];                                          \
  } while (0)
#define MPH_PUSH(buf_, len_, c_)           \
  do {                                     \
    if ((len_) < cap) (buf_)[len_] = (c_); \
    else flags |= MPH_HF_OVERFLOW;         \
    ++(len_);                              \
  } while (0)
  uint64_t i = gk.s;
  uint32_t j = 0;
  const uint64_t window_end = gk.e;
  while (i < window_end) {
    while (j < n && i == vars[va + j].pos) {
      const MphVar v = vars[va + j];
      const uint32_t bit_pos = rev ? j : n - 1 - j;
      const bool germline = (v.flags & MPH_VF_GERMLINE) != 0;
      if ((hap >> (bit_pos & 63)) & 1) {
        uint8_t r;
        if (v.kind == MPH_SNV) {
          MPH_REF(i, r);
          const uint8_t a = mph_is_upper(r) ? mph_lower(v.alt) : v.alt;  // switch_ascii_case :26-32
          MPH_PUSH(germ, gl, germline ? a : r);
          MPH_PUSH(seq, sl, a);
          i += 1;
        } else if (v.kind == MPH_INS) {
          MPH_REF(i, r);
          const bool up = mph_is_upper(r);
          for (uint32_t t = 0; t <= v.len; ++t) {
            const uint8_t c0 = ins_arena[v.ins_off + t];
            const uint8_t c = up ? mph_lower(c0) : mph_upper(c0);  // switch_ascii_case_vec :34-40
            if (germline) MPH_PUSH(germ, gl, c);
            MPH_PUSH(seq, sl, c);
          }
          if (!germline) flags |= MPH_HF_INDEL;
          flags |= MPH_HF_INSERTION;
          i += 1;
        } else {
          if (rev && (uint64_t)v.pos + v.len - 1 >= window_end) { brk = 1; break; }  // :549-552 (j stays stuck)
          MPH_REF(i, r);
          if (germline || i == window_end - 1) {
            MPH_PUSH(germ, gl, r);
          } else {
            for (uint64_t t = i; t < i + v.len + 1; ++t) {
              uint8_t rr;
              MPH_REF(t, rr);
              MPH_PUSH(germ, gl, rr);
            }
            flags |= MPH_HF_INDEL;
          }
          MPH_PUSH(seq, sl, r);
          i += (uint64_t)v.len + 1;
        }
        const uint64_t code = germline ? 1 : 2;
        if (n_prof < 32) profile |= code << (2 * n_prof);
        if (!germline) ++n_som;
        ++n_var;
      }
      ++n_prof;
      ++j;
    }
    if (i < window_end) {
      uint8_t r;
      MPH_REF(i, r);
      MPH_PUSH(seq, sl, r);
      MPH_PUSH(germ, gl, r);
      i += 1;
    }
  }
#undef MPH_REF
#undef MPH_PUSH
  if (n_prof > 32) err |= MPH_E_VARS_PER_WINDOW;
  bool eq = sl == gl;
  const uint32_t lim = sl < cap ? sl : cap;
  for (uint32_t t = 0; eq && t < lim; ++t) eq = seq[t] == germ[t];
  if (eq) flags |= MPH_HF_GERM_EQ;
  uint32_t a, b;
  mph_neo_slice(gk, sl < cap ? sl : cap, g.ewl, (flags & MPH_HF_INSERTION) != 0, &a, &b);
  if (mph_has_stop(seq + a, b - a, !rev)) flags |= MPH_HF_STOP;
  {
    // normal_peptide vs neopeptide (:677-693) when no reading frame but the main one is open: germline_seq is cleared
    // for a somatic indel + insertion only (:624-631)
    const bool indel = (flags & MPH_HF_INDEL) != 0, insertion = (flags & MPH_HF_INSERTION) != 0;
    const uint32_t slc = sl < cap ? sl : cap, glc = (indel && insertion) ? 0u : (gl < cap ? gl : cap);
    const uint32_t twl = slc < g.ewl ? slc : g.ewl, nwl = indel ? (glc < g.ewl ? glc : g.ewl) : twl;
    uint32_t na = 0, nb = 0, pa = 0, pb = slc;
    bool bad = false;
    if (glc != 0) {
      if (gk.spos == 1) { na = gk.gap; nb = glc; bad = bad || gk.gap > glc; }
      else if (gk.spos == 0) { nb = nwl; bad = bad || nwl > glc; }
      else nb = glc;
    }
    if (gk.spos == 1) { pa = gk.gap; bad = bad || gk.gap > slc; }
    else if (gk.spos == 0 && !insertion) pb = twl;
    if (bad) {
      flags |= MPH_HF_SLICE_ERR;
    } else {
      bool diff = (nb - na) != (pb - pa);
      for (uint32_t t = 0; !diff && t < nb - na; ++t) diff = germ[na + t] != seq[pa + t];
      if (diff) flags |= MPH_HF_PEPDIFF;
    }
  }
  out->flags = flags;
  out->seq_len = (uint16_t)sl;
  out->germ_len = (uint16_t)gl;
  out->n_var = (uint8_t)n_var;
  out->n_som = (uint8_t)n_som;
  out->n_prof = (uint8_t)(n_prof < 255 ? n_prof : 255);
  out->brk = (uint8_t)brk;
  out->seq_off = 0xFFFFFFFFu;
  out->profile = profile;
  out->id64 = 0;
  return err;
}

// Haplotype 0 of any window: no variant is applied, so seq == germline_seq == refseq[s..e)
// (:464-471, :594-599) and the walk visits all nvar variants without touching the sequence. Only
// the stop test remains; it reads the packer's stop-codon bitmap (bit t <=> the three reference
// bytes starting at arena byte t spell a stop codon for the segment's strand).
MPH_HD uint32_t mph_plain_hap(const MphSegment& g, const MphGeom& gk, const uint32_t* stopmap, const uint8_t* ref_arena, uint32_t nvar,
                              MphHap* out) {
  uint32_t err = 0;
  const uint32_t len = gk.e - gk.s;
  uint32_t flags = MPH_HF_GERM_EQ;
  if (gk.s < g.ref_pos0 || (uint64_t)gk.e - g.ref_pos0 > g.ref_len) {
    flags |= MPH_HF_REFRANGE;
  } else {
    uint32_t a, b;
    mph_neo_slice(gk, len, g.ewl, false, &a, &b);
    const bool fwd = (g.flags & MPH_SF_REVERSE) == 0;
    if (b >= a + 3) {
      if (len <= 61) {
        // codon starts q in [a, b-3] with q = a (mod 3) forward / q = b (mod 3) reverse (:42-76)
        const uint64_t bit0 = (uint64_t)g.ref_off + (gk.s - g.ref_pos0);
        const uint32_t w = (uint32_t)(bit0 >> 5), sh = (uint32_t)(bit0 & 31);
        const uint64_t lo = stopmap[w] | ((uint64_t)stopmap[w + 1] << 32);
        uint64_t bits = lo >> sh;
        if (sh) bits |= (uint64_t)stopmap[w + 2] << (64 - sh);
        const uint32_t phase = (fwd ? a : b) % 3;
        uint64_t mask = 0x9249249249249249ull << phase;
        mask &= ~(uint64_t)0 << a;
        mask &= ~(uint64_t)0 >> (64 - (b - 2));
        if (bits & mask) flags |= MPH_HF_STOP;
      } else {
        const uint8_t* p = ref_arena + g.ref_off + (gk.s - g.ref_pos0);
        if (mph_has_stop(p + a, b - a, fwd)) flags |= MPH_HF_STOP;
      }
    }
  }
  out->flags = flags;
  out->seq_len = (uint16_t)len;
  out->germ_len = (uint16_t)len;
  out->n_var = 0; out->n_som = 0; out->n_prof = (uint8_t)(nvar < 255 ? nvar : 255); out->brk = 0;
  out->seq_off = 0xFFFFFFFFu;
  out->profile = 0;
  out->id64 = 0;
  if (nvar > 32) err |= MPH_E_VARS_PER_WINDOW;
  return err;
}

// byte-wise variant of the above for a window without variants (kept for cross-checks)
MPH_HD uint32_t mph_plain_window(const MphSegment& g, const MphGeom& gk, const uint8_t* ref_arena, MphHap* out) {
  uint32_t err = 0;
  const uint32_t len = gk.e - gk.s;
  uint32_t flags = MPH_HF_GERM_EQ;
  if (gk.s < g.ref_pos0 || (uint64_t)gk.e - g.ref_pos0 > g.ref_len) {
    flags |= MPH_HF_REFRANGE;
  } else {
    const uint8_t* p = ref_arena + g.ref_off + (gk.s - g.ref_pos0);
    uint32_t a, b;
    mph_neo_slice(gk, len, g.ewl, false, &a, &b);
    if (mph_has_stop(p + a, b - a, (g.flags & MPH_SF_REVERSE) == 0)) flags |= MPH_HF_STOP;
  }
  out->flags = flags;
  out->seq_len = (uint16_t)len;
  out->germ_len = (uint16_t)len;
  out->n_var = 0; out->n_som = 0; out->n_prof = 0; out->brk = 0;
  out->seq_off = 0xFFFFFFFFu;
  out->profile = 0;
  out->id64 = 0;
  return err;
}

// ================================================================================================
// `normal` mode (reference src/normal_microphasing.rs; line numbers below refer to that file).
// Same window geometry; what differs in the matrix: no quality / mapq filters, no
// `contains`, push_read numbers the variants oldest-first (:317) while extend_right numbers the new
// ones newest-first (:260), and on the reverse strand cleanup_reads(splice_side_offset) (:1001)
// together with re-offering every read each iteration (:942-967) inserts a read once per iteration.
// ================================================================================================

// Number of copies of a read in the matrix at iteration k of a reverse-strand segment, and the first
// iteration kc at which a copy was pushed (copies exist for every iteration kc..k).
MPH_HD uint32_t mph_nrm_rev_copies(const MphSegment& g, uint32_t k, const MphGeom& gk, uint32_t start, uint32_t end, uint32_t* kc_out) {
  *kc_out = k;
  if (start > gk.s || end < gk.e) return 0;
  const uint64_t lim = (uint64_t)start + g.K;
  if (start == gk.s) return lim >= gk.s ? 1u : 0u;  // older copies were just removed (keys >= s, :275-278)
  uint32_t kc = g.off0 > lim ? (uint32_t)(g.off0 - lim) : 0;
  if (mph_geom(g, 0).e > end) {
    const uint64_t t = (uint64_t)g.off0 + g.ewl;
    const uint32_t k2 = t > end ? (uint32_t)(t - end) : 1;
    if (k2 > kc) kc = k2;
    if (kc == 0) kc = 1;
  }
  kc = kc >= 2 ? kc - 2 : 0;
  for (; kc <= k; ++kc) {
    const MphGeom gc = mph_geom(g, kc);
    if (gc.s <= lim && gc.e <= end) break;
  }
  if (kc > k) return 0;
  *kc_out = kc;
  return k - kc + 1;
}

// obs.haplotype at iteration k of an observation pushed at iteration kp (:238-267,301-331).
// Storage order of the variants is ascending position (ALT order reversed on the reverse strand);
// the matrix adds them in ascending index order on the forward strand, descending on the reverse strand.
MPH_HD uint64_t mph_nrm_hap(const MphSegment& g, const MphVar* vars, uint32_t kp, uint32_t k, uint32_t va, uint32_t vb, uint32_t vlo, uint64_t S) {
  const bool rev = (g.flags & MPH_SF_REVERSE) != 0;
  if (S == 0) return 0;
  auto sbit = [&](uint32_t idx) -> uint64_t { return (idx >= vlo && idx - vlo < 64) ? ((S >> (idx - vlo)) & 1) : 0; };
  // deque at the push: what the previous iteration held minus what has just left (empty at iteration 0)
  uint32_t a0, b0;
  if (kp == 0) {
    a0 = b0 = rev ? mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, 0).e) : mph_var_lb(vars, g.var_lo, g.var_hi, mph_geom(g, 0).s);
  } else {
    const MphGeom gp = mph_geom(g, kp), gq = mph_geom(g, kp - 1);
    a0 = mph_var_lb(vars, g.var_lo, g.var_hi, rev ? gq.s : gp.s);
    b0 = mph_var_lb(vars, g.var_lo, g.var_hi, rev ? gp.e : gq.e);
    if (b0 < a0) b0 = a0;
  }
  const uint32_t m = b0 - a0;
  const uint32_t A = rev ? (a0 > va ? a0 - va : 0) : (vb > b0 ? vb - b0 : 0);
  const uint32_t Del = rev ? (b0 > vb ? b0 - vb : 0) : (va > a0 ? va - a0 : 0);
  uint64_t hap = 0;
  // push-time columns: bit i <-> i-th oldest; it survives the masks while Del < m - i
  for (uint32_t i = 0; i < m && i + Del < m; ++i) {
    const uint32_t idx = rev ? b0 - 1 - i : a0 + i;
    const uint32_t pos = i + A;
    if (pos < 64 && sbit(idx)) hap |= (uint64_t)1 << pos;
  }
  // columns added afterwards: the j-th added sits at bit A - j and survives while Del < m + j
  for (uint32_t j = 1; j <= A; ++j) {
    if (Del >= m + j) continue;
    const uint32_t idx = rev ? a0 - j : b0 + j - 1;
    const uint32_t pos = A - j;
    if (pos < 64 && sbit(idx)) hap |= (uint64_t)1 << pos;
  }
  return hap;
}

enum { MPH_NF_STOP = 1, MPH_NF_INSERTION = 4, MPH_NF_SEQ = 8, MPH_NF_OVERFLOW = 32, MPH_NF_REFRANGE = 64, MPH_NF_ID = 128 };

// Sequence walk of the normal-mode print_haplotypes (:403-507) for one haplotype; `all_reads` = the
// haplotype is carried by every observation (freq == 1, :422). Fills seq (cap bytes) and out.
MPH_HD uint32_t mph_nrm_assemble(const MphSegment& g, const MphGeom& gk, const MphVar* vars, uint32_t va, uint32_t vb, const uint8_t* ref_arena,
                                 const uint8_t* ins_arena, uint64_t hap, bool all_reads, uint8_t* seq, uint32_t cap, MphHap* out) {
  const bool rev = (g.flags & MPH_SF_REVERSE) != 0;
  const uint32_t n = vb - va;
  uint32_t err = 0, sl = 0, flags = 0, n_var = 0, n_som = 0, n_prof = 0;
  uint64_t profile = 0;
  const uint8_t* ref = ref_arena + g.ref_off;
#define MPH_NREF(i_, dst_)                                                            \
  do {                                                                                \
    const uint64_t ri_ = (uint64_t)(i_) - g.ref_pos0;                                 \
    if ((i_) < g.ref_pos0 || ri_ >= g.ref_len) { flags |= MPH_HF_REFRANGE; dst_ = 'N'; } \
    else dst_ = ref[ri_];                                                             \
  } while (0)
#define MPH_NPUSH(c_)                        \
  do {                                       \
    if (sl < cap) seq[sl] = (c_);            \
    else flags |= MPH_NF_OVERFLOW;           \
    ++sl;                                    \
  } while (0)
  uint64_t i = gk.s, window_end = gk.e;
  uint32_t j = 0;
  if (n == 0) {
    for (; i < window_end; ++i) { uint8_t r; MPH_NREF(i, r); MPH_NPUSH(r); }
  } else {
    while (i < window_end) {
      while (j < n && i == vars[va + j].pos) {
        if (all_reads && !(vars[va + j].flags & MPH_VF_GERMLINE)) {  // :422-426
          ++j;
          ++n_prof;
          continue;
        }
        if ((hap >> (j & 63)) & 1) {
          if (j + 1 < n && i == vars[va + j + 1].pos) ++j;  // :429-431
          const MphVar v = vars[va + j];
          uint8_t r;
          MPH_NREF(i, r);
          if (v.kind == MPH_SNV) {
            MPH_NPUSH(mph_is_upper(r) ? mph_lower(v.alt) : v.alt);
            i += 1;
          } else if (v.kind == MPH_INS) {
            const bool up = mph_is_upper(r);
            for (uint32_t t = 0; t <= v.len; ++t) {
              const uint8_t c0 = ins_arena[v.ins_off + t];
              MPH_NPUSH(up ? mph_lower(c0) : mph_upper(c0));
            }
            flags |= MPH_NF_INSERTION;
            i += 1;
          } else {
            MPH_NPUSH(r);
            i += (uint64_t)v.len + 1;
            window_end += (uint64_t)v.len + 1;  // :457
          }
          const uint64_t code = (v.flags & MPH_VF_GERMLINE) ? 1 : 2;
          if (n_prof < 32) profile |= code << (2 * n_prof);
          if (code == 2) ++n_som;
          ++n_var;
        }
        ++n_prof;
        ++j;
      }
      uint8_t r;
      MPH_NREF(i, r);  // :476 — unconditional, also right after a variant at the last window position
      MPH_NPUSH(r);
      i += 1;
    }
  }
#undef MPH_NREF
#undef MPH_NPUSH
  if (n_prof > 32) err |= MPH_E_VARS_PER_WINDOW;
  // peptide slice (:485-492) and the first / last codon stop test (:493-507)
  const uint32_t len = sl < cap ? sl : cap;
  const uint32_t twl = len < g.ewl ? len : g.ewl;
  uint32_t a = 0, b = len;
  if (gk.spos == 1) a = gk.gap < len ? gk.gap : len;
  else if (gk.spos == 0 && !(flags & MPH_NF_INSERTION)) b = twl;
  if (b >= a + 3) {
    const uint8_t* p = seq + a;
    const uint32_t pl = b - a;
    bool stop;
    if (!rev) stop = p[0] == 'T' && ((p[1] == 'G' && p[2] == 'A') || (p[1] == 'A' && (p[2] == 'G' || p[2] == 'A')));
    else stop = p[pl - 1] == 'A' && ((p[pl - 3] == 'T' && (p[pl - 2] == 'C' || p[pl - 2] == 'T')) || (p[pl - 3] == 'C' && p[pl - 2] == 'T'));
    if (stop) flags |= MPH_NF_STOP;
  }
  out->flags = flags;
  out->seq_len = (uint16_t)sl;
  out->germ_len = 0;
  out->n_var = (uint8_t)n_var;
  out->n_som = (uint8_t)n_som;
  out->n_prof = (uint8_t)(n_prof < 255 ? n_prof : 255);
  out->brk = 0;
  out->seq_off = 0xFFFFFFFFu;
  out->profile = profile;
  out->id64 = 0;
  return err;
}

// haplotype 0 of a window without variants in normal mode: seq = refseq[s..e) (:409-412), first / last codon stop test
MPH_HD uint32_t mph_nrm_plain(const MphSegment& g, const MphGeom& gk, const uint8_t* ref_arena, uint32_t nv, MphHap* out) {
  uint32_t err = 0, flags = 0;
  const uint32_t len = gk.e - gk.s;
  if (gk.s < g.ref_pos0 || (uint64_t)gk.e - g.ref_pos0 > g.ref_len) {
    flags |= MPH_HF_REFRANGE;
  } else {
    const uint8_t* q = ref_arena + g.ref_off + (gk.s - g.ref_pos0);
    const uint32_t twl = len < g.ewl ? len : g.ewl;
    uint32_t a = 0, b = len;
    if (gk.spos == 1) a = gk.gap < len ? gk.gap : len;
    else if (gk.spos == 0) b = twl;
    if (b >= a + 3) {
      const uint8_t* p = q + a;
      const uint32_t pl = b - a;
      bool stop;
      if (!(g.flags & MPH_SF_REVERSE)) stop = p[0] == 'T' && ((p[1] == 'G' && p[2] == 'A') || (p[1] == 'A' && (p[2] == 'G' || p[2] == 'A')));
      else stop = p[pl - 1] == 'A' && ((p[pl - 3] == 'T' && (p[pl - 2] == 'C' || p[pl - 2] == 'T')) || (p[pl - 3] == 'C' && p[pl - 2] == 'T'));
      if (stop) flags |= MPH_NF_STOP;
    }
  }
  out->flags = flags;
  out->seq_len = (uint16_t)len;
  out->germ_len = 0;
  out->n_var = 0; out->n_som = 0; out->brk = 0;
  out->n_prof = (uint8_t)(nv < 255 ? nv : 255);  // every window variant is visited and left unset
  out->seq_off = 0xFFFFFFFFu;
  out->profile = 0;
  out->id64 = 0;
  return err;
}

// Forward strand, normal mode: iteration at which the read is pushed (:974-1003) or 0xFFFFFFFF if it
// is not an observation of window k. One copy only: s(k) grows strictly, a read is fetched once.
MPH_HD uint32_t mph_nrm_fwd_entry(const MphSegment& g, uint32_t k, const MphGeom& gk, uint32_t start, uint32_t end) {
  if (end < gk.e) return 0xFFFFFFFFu;
  const uint32_t s0 = g.off0 - g.ceo;
  if (start <= s0) return (int64_t)start < (int64_t)s0 - (int64_t)g.K ? 0xFFFFFFFFu : 0u;
  if (start <= g.off0) return 0xFFFFFFFFu;  // starts in (s0, off0] are never fetched when ceo > 0
  const uint32_t k_ins = start - g.off0;
  return k_ins > k ? 0xFFFFFFFFu : k_ins;
}
