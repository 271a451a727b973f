// records_host.hpp — host side of the device-built records (core/record_core.h): renders the text columns of an MphRec
// (reference IDRecord, src/common.rs:350-373) and puts host-built and device-built records into one ordered stream.
// Only formatting happens here: which variants a record lists, its counts, frequency and sequences were decided by the
// record kernels.
#pragma once
#include <charconv>

#include "residue.hpp"

namespace mph {

// what rendering needs besides the record itself: the batch's variants (positions, amino-acid changes), the bytes the
// record's seq_off points into and the second sources of merged records
struct RenderCtx {
  const MphVar* vars = nullptr;
  const std::vector<std::string>* var_prot = nullptr;
  const uint8_t* seq = nullptr;
  const MphRecSrc* aux = nullptr;
  const uint8_t* ref = nullptr;  // the batch's reference arena (normal mode: records of reference windows ship no bytes)
};
inline RenderCtx render_ctx(const Batch& b, const PhaseRaw& raw) {
  RenderCtx c;
  c.vars = b.vars.data(); c.var_prot = &b.var_prot; c.seq = raw.rec_seq.data(); c.aux = raw.rec_aux.data(); c.ref = b.ref.data();
  return c;
}

namespace detail {

inline void append_u64(std::string& dst, uint64_t v) {
  char buf[24];
  auto r = std::to_chars(buf, buf + sizeof buf, v);
  dst.append(buf, r.ptr);
}

// variant_sites of one source window (:757-769): 1-based positions of its distinct variant sites
inline void append_sites(const RenderCtx& b, uint32_t var_ref, uint32_t n_win, std::string& dst, uint32_t one_based = 1) {
  bool first = true;
  for (uint32_t c = 0; c < n_win; ++c) {
    const MphVar& v = b.vars[var_ref + c];
    if (c != 0 && v.pos == b.vars[var_ref + c - 1].pos) continue;
    if (!first) dst.push_back('|');
    first = false;
    append_u64(dst, uint64_t(v.pos) + one_based);
  }
}

// positions / amino-acid changes of the variants a source contributes (:733-749, common.rs:399-478)
inline void append_lists(const RenderCtx& b, uint32_t var_ref, uint64_t profile, uint32_t n_prof, uint32_t keep, InfoRecord& o, bool& fs, bool& fsa,
                         bool& fg, bool& fga, uint32_t one_based = 1) {
  for (uint32_t c = 0; c < n_prof && c < 32; ++c) {
    const unsigned code = unsigned((profile >> (2 * c)) & 3);
    if (!code || !((keep >> c) & 1u)) continue;
    const uint32_t vi = var_ref + c;
    std::string& pos = code == 2 ? o.somatic_positions : o.germline_positions;
    std::string& aa = code == 2 ? o.somatic_aa_change : o.germline_aa_change;
    bool& fp = code == 2 ? fs : fg;
    bool& fa = code == 2 ? fsa : fga;
    if (!fp) pos.push_back('|');
    fp = false;
    append_u64(pos, uint64_t(b.vars[vi].pos) + one_based);
    if (!fa) aa.push_back('|');
    fa = false;
    aa += (*b.var_prot)[vi];
  }
}

}  // namespace detail

// text form of one device-built record
// text form of a record of the `normal` mode (IDRecord of src/normal_microphasing.rs:80-102): 0-based positions, the site list
// stops with the visited variants (:509-560), a merged record's lists are its two sources' strings put together as they are
// (:105-146: plain `+`, no separator) and its counts are 32-bit (:148-179 may wrap them)
inline OutRecord render_record_normal(const RenderCtx& b, const MphRec& r) {
  OutRecord o;
  InfoRecord& info = o.info;
  static const char* hx = "0123456789abcdef";
  info.id.resize(16);
  for (int q = 0; q < 15; ++q) info.id[q] = hx[(r.id64 >> (60 - 4 * q)) & 15];
  info.id[15] = (r.flags & MPH_RC_REVERSE) ? 'R' : 'F';
  info.tx = r.tx;
  info.offset = r.offset;
  info.frame = 0;
  info.freq = r.freq;
  info.depth = r.depth;
  info.nvar = r.nvar;
  info.nsomatic = r.nsomatic;
  info.nvariant_sites = r.nsites;
  info.nsomvariant_sites = r.nsomsites;
  auto source = [&](uint32_t var_ref, uint64_t profile, uint32_t n_prof, uint32_t n_win) {
    bool fs = true, fsa = true, fg = true, fga = true;
    std::string sp, sa, gp, ga, st;
    InfoRecord tmp;
    detail::append_lists(b, var_ref, profile, std::min(n_prof, n_win), 0xFFFFFFFFu, tmp, fs, fsa, fg, fga, 0);
    detail::append_sites(b, var_ref, std::min(n_prof, n_win), st, 0);
    info.somatic_positions += tmp.somatic_positions; info.somatic_aa_change += tmp.somatic_aa_change;
    info.germline_positions += tmp.germline_positions; info.germline_aa_change += tmp.germline_aa_change;
    info.variant_sites += st;
  };
  source(r.var_ref, r.profile, r.n_prof, r.n_win);
  if (r.flags & MPH_RC_MERGED) {
    const MphRecSrc& x = b.aux[r.aux];
    source(x.var_ref, x.profile, x.n_prof, x.n_win);
    info.nvar = r.keep;       // 32-bit counts of a merged record
    info.nsomatic = x.keep;
  }
  const char* mp = reinterpret_cast<const char*>((r.flags & MPH_RC_REFSEQ) ? b.ref : b.seq) + r.seq_off;
  info.mutant_sequence.assign(mp, r.neo_len);
  if (r.flags & MPH_RC_HAS_MT) {
    o.has_mt = true;
    o.mt_same = r.mt_len == r.neo_len;
    if (!o.mt_same) o.mt.assign(mp, r.mt_len);
  }
  return o;
}

inline OutRecord render_record(const RenderCtx& b, const MphRec& r) {
  if (r.flags & MPH_RC_NORMAL) return render_record_normal(b, r);
  OutRecord o;
  InfoRecord& info = o.info;
  static const char* hx = "0123456789abcdef";
  info.id.resize(16);
  for (int q = 0; q < 15; ++q) info.id[q] = hx[(r.id64 >> (60 - 4 * q)) & 15];
  info.id[15] = (r.flags & MPH_RC_REVERSE) ? 'R' : 'F';
  info.tx = r.tx;
  info.offset = r.offset;
  info.frame = 0;
  info.freq = r.freq;
  info.depth = r.depth;
  info.nvar = r.nvar;
  info.nsomatic = r.nsomatic;
  info.nvariant_sites = r.nsites;
  info.nsomvariant_sites = r.nsomsites;
  bool fs = true, fsa = true, fg = true, fga = true;
  detail::append_lists(b, r.var_ref, r.profile, r.n_prof, r.keep, info, fs, fsa, fg, fga);
  detail::append_sites(b, r.var_ref, r.n_win, info.variant_sites);
  if (r.flags & MPH_RC_MERGED) {
    const MphRecSrc& x = b.aux[r.aux];
    detail::append_lists(b, x.var_ref, x.profile, x.n_prof, x.keep, info, fs, fsa, fg, fga);
    // "self|rec" with one leading and one trailing '|' removed (common.rs:497-503)
    std::string& vr = info.variant_sites;
    vr.push_back('|');
    detail::append_sites(b, x.var_ref, x.n_win, vr);
    if (!vr.empty() && vr.front() == '|') vr.erase(0, 1);
    if (!vr.empty() && vr.back() == '|') vr.pop_back();
  }
  const uint32_t mut_n = std::max(r.neo_len, r.mt_len), nrm_n = std::max(r.norm_len, r.wt_len);
  const char* mp = reinterpret_cast<const char*>(b.seq) + r.seq_off;
  const char* np = mp + mut_n;
  (void)nrm_n;
  info.mutant_sequence.assign(mp, r.neo_len);
  info.normal_sequence.assign(np, r.norm_len);
  if (r.flags & MPH_RC_HAS_MT) {
    o.has_mt = true;
    o.mt_same = r.mt_len == r.neo_len;
    if (!o.mt_same) o.mt.assign(mp, r.mt_len);
  }
  if (r.flags & MPH_RC_HAS_WT) {
    o.has_wt = true;
    o.wt_same = r.wt_len == r.norm_len;
    if (!o.wt_same) o.wt.assign(np, r.wt_len);
  }
  return o;
}

// host-built records (host-class transcripts, ascending transcript) and device-built records (device class, ascending
// transcript) as one stream in transcript order; a transcript belongs to exactly one class
// (restricted to the device-built records of transcripts [tx_lo, tx_hi))
inline std::vector<OutRecord> ordered_records(const Batch& b, const PhaseRaw& raw, std::vector<OutRecord>&& host_recs, uint32_t tx_lo = 0,
                                              uint32_t tx_hi = 0xFFFFFFFFu) {
  auto by_tx = [](const MphRec& r, uint32_t t) { return r.tx < t; };
  size_t d = size_t(std::lower_bound(raw.recs.begin(), raw.recs.end(), tx_lo, by_tx) - raw.recs.begin());
  const size_t d_end = size_t(std::lower_bound(raw.recs.begin() + long(d), raw.recs.end(), tx_hi, by_tx) - raw.recs.begin());
  if (d == d_end) return std::move(host_recs);
  const RenderCtx rc = render_ctx(b, raw);
  std::vector<OutRecord> out;
  out.reserve(host_recs.size() + (d_end - d));
  size_t h = 0;
  while (h < host_recs.size() || d < d_end) {
    const bool take_dev = h == host_recs.size() || (d < d_end && raw.recs[d].tx < host_recs[h].info.tx);
    if (take_dev) out.push_back(render_record(rc, raw.recs[d++]));
    else out.push_back(std::move(host_recs[h++]));
  }
  return out;
}

}  // namespace mph
