// emu_pipeline.hpp — TEST INFRASTRUCTURE ONLY. A CPU stand-in for the device side of
// mph_phase_batch: the same per-item functions as the CUDA kernels (csrc/core/phase_core.h)
// driven by plain loops. It exists so the closed-form window logic, the packer and the host
// residue can be diffed against the oracle in the GPU-less build container. The product library
// never includes this file; on a GPU box the parity tests call the CUDA path through the C ABI.
#pragma once
#include <map>

#include "../../microphaser_b200/csrc/host/residue.hpp"

namespace mphemu {

using namespace mph;

inline uint32_t partner_of(const Batch& b, uint32_t r) {
  for (size_t i = 0; i < b.partner_a.size(); ++i) {
    if (b.partner_a[i] == r) return b.partner_b[i];
    if (b.partner_b[i] == r) return b.partner_a[i];
  }
  return 0xFFFFFFFFu;
}

// what K1 expands on the device: per-read (vlo, nv, side-table entry) from the compact table
struct PerRead {
  std::vector<uint32_t> vlo, vr;
  std::vector<uint8_t> nv;
  explicit PerRead(const Batch& b) : vlo(b.read_start.size(), 0), vr(b.read_start.size(), 0xFFFFFFFFu), nv(b.read_start.size(), 0) {
    for (size_t e = 0; e < b.vr_read.size(); ++e) { vlo[b.vr_read[e]] = b.vr_vlo[e]; nv[b.vr_read[e]] = b.vr_nv[e]; vr[b.vr_read[e]] = uint32_t(e); }
  }
};

inline std::vector<MphCall> allele_calls(const Batch& b, bool use_qual) {
  std::vector<MphCall> calls(b.read_start.size());
  for (auto& c : calls) { c.S = 0; c.B = 0; }
  for (size_t e = 0; e < b.vr_read.size(); ++e) {
    const uint32_t r = b.vr_read[e];
    MphRead rd{b.read_start[r], b.read_end[r], b.vr_vlo[e], b.vr_lseq[e], b.vr_nv[e], b.vr_ncig[e]};
    calls[r] = mph_call_read(rd, b.bases.data() + size_t(b.vr_seq_off[e]), b.cigars.data() + b.vr_cig_off[e], b.vars.data(), use_qual);
  }
  return calls;
}

// serial replay of the irregular transcripts (core/replay_core.h) into per-window staging arrays
struct ReplayStage {
  std::vector<MphWinOut> out;
  std::vector<MphHap> hap0;
  std::vector<uint8_t> flag;
  std::vector<uint32_t> voff, vlist, histwin;
  std::vector<MphHist> hist;
};

inline void run_replay(const Batch& b, const std::vector<MphCall>& calls, const PerRead& pr_, int mode, PhaseRaw& raw, ReplayStage& st) {
  if (b.replay.empty()) return;
  const size_t nr = b.read_start.size();
  st.out.resize(b.n_windows); st.hap0.resize(b.n_windows); st.flag.resize(b.n_windows); st.voff.assign(b.n_windows, 0xFFFFFFFFu);
  std::vector<uint64_t> S(nr), B(nr);
  for (size_t r = 0; r < nr; ++r) { S[r] = calls[r].S; B[r] = calls[r].B; }
  std::vector<std::pair<uint32_t, uint32_t>> pr;
  for (auto& e : b.pair_edges) pr.push_back(e);
  std::sort(pr.begin(), pr.end());
  std::vector<uint32_t> pairs;
  for (auto& x : pr) { pairs.push_back(x.first); pairs.push_back(x.second); }
  std::vector<uint32_t> o_read(b.replay_obs + 1), o_frame(b.replay_obs + 1), o_last(b.replay_obs + 1);
  std::vector<uint64_t> o_hap(b.replay_obs + 1);
  std::vector<uint8_t> o_flags(b.replay_obs + 1), o_inmat(b.replay_obs + 1);
  st.hist.resize(size_t(b.n_windows) * 8 + 1024);
  st.histwin.resize(st.hist.size());
  st.vlist.resize(size_t(b.n_windows) * 66 + 1024);
  uint32_t counters[8] = {0};
  unsigned long long sum_depth = 0;
  raw.seg_err.assign(b.segs.size(), 0);
  if (mode == 1 && raw.win_depth.size() != b.n_windows) { raw.win_depth.assign(b.n_windows, 0); raw.win_id.assign(b.n_windows, 0); }
  MphReplayCtx c;
  c.seg_err = raw.seg_err.data();
  c.read_start = b.read_start.data(); c.read_end = b.read_end.data(); c.read_vlo = pr_.vlo.data(); c.read_nv = pr_.nv.data(); c.read_vr = pr_.vr.data();
  c.vr_seq_off = b.vr_seq_off.data(); c.vr_cig_off = b.vr_cig_off.data(); c.vr_lseq = b.vr_lseq.data(); c.vr_ncig = b.vr_ncig.data();
  c.read_flags = b.read_flags.data(); c.bases = b.bases.data(); c.cigars = b.cigars.data(); c.call_S = S.data(); c.call_B = B.data();
  c.pairs = pairs.data(); c.n_pairs = uint32_t(pairs.size() / 2); c.vars = b.vars.data(); c.segs = b.segs.data(); c.seg_chunk0 = b.seg_chunk0.data();
  c.stopmap = b.stopmap.data(); c.ref = b.ref.data(); c.dq_init = b.replay_dq.data();
  c.batch = getenv("MPH_EMU_NO_FOLD") ? 0u : 1u;  // the kernels fold; MPH_EMU_NO_FOLD=1 runs the literal per-iteration replay
  c.mode = uint32_t(mode); c.tx_id_bytes = b.tx_id_bytes.data(); c.tx_id_off = b.tx_id_off.data();
  c.win_depth = mode == 1 ? raw.win_depth.data() : nullptr;
  c.win_id = mode == 1 ? reinterpret_cast<unsigned long long*>(raw.win_id.data()) : nullptr;
  c.o_last = o_last.data();
  c.o_read = o_read.data(); c.o_hap = o_hap.data(); c.o_frame = o_frame.data(); c.o_flags = o_flags.data(); c.o_inmat = o_inmat.data();
  c.win_out = st.out.data(); c.hist = st.hist.data(); c.hist_win = st.histwin.data(); c.hist_cap = uint32_t(st.hist.size());
  c.hap0 = st.hap0.data(); c.win_flag = st.flag.data(); c.win_voff = st.voff.data(); c.vlist = st.vlist.data(); c.vlist_cap = uint32_t(st.vlist.size());
  c.counters = counters; c.sum_depth = &sum_depth;
  for (const MphReplayTx& t : b.replay) mph_replay_tx(c, t);
  raw.err |= counters[MPH_RP_CTR_ERR];
  raw.sum_depth += sum_depth;
}

// normal mode (reference src/normal_microphasing.rs)
inline PhaseRaw phase_normal(const Batch& b) {
  PhaseRaw raw;
  const size_t nr = b.read_start.size();
  const std::vector<MphCall> calls = allele_calls(b, false);
  const PerRead per_read(b);
  const std::vector<uint32_t>& read_vlo = per_read.vlo;
  for (size_t r = 0; r < nr; ++r)
    if (b.read_flags[r] & MPH_RF_OVERFLOW) raw.err |= MPH_E_VARS_PER_WINDOW;
  raw.win_depth.assign(b.n_windows, 0);
  raw.win_id.assign(b.n_windows, 0);
  ReplayStage rp;
  run_replay(b, calls, per_read, 1, raw, rp);
  std::vector<uint8_t> seqbuf(b.seq_cap);
  // device-class transcripts (core/record_core.h): full per-window arrays, as on the device
  std::vector<MphWinOut> d_win_out(b.n_windows);
  std::vector<MphHap> d_hap0(b.n_windows);
  for (const MphChunk& ch : b.chunks) {
    const MphSegment& sg = b.segs[ch.seg];
    const bool rev = sg.flags & MPH_SF_REVERSE;
    for (uint32_t i = ch.i_first; i < ch.i_first + ch.n; ++i) {
      const uint32_t k = sg.k_first + i * sg.k_stride;
      const uint32_t widx = sg.win_base + i;
      const MphGeom g = mph_geom(sg, k);
      if (sg.flags & MPH_SF_REPLAY) {
        // win_depth / win_id were written by the replay; windows with matrix columns also get their haplotypes assembled
        if (!rp.flag[widx]) continue;
        MphWinOut wo = rp.out[widx];
        const uint32_t* list = rp.voff[widx] == 0xFFFFFFFFu ? nullptr : &rp.vlist[rp.voff[widx]];
        const uint32_t ncol = list ? list[0] : 0;
        std::vector<MphVar> gathered(ncol + 1);
        for (uint32_t j = 0; j < ncol; ++j) gathered[j] = b.vars[list[1 + j]];
        const uint32_t src = wo.extra_off;
        wo.extra_off = uint32_t(raw.hist.size());
        for (uint32_t x = 0; x < wo.n_extra; ++x) {
          const MphHist e = rp.hist[src + x];
          MphHap hx;
          raw.err |= mph_nrm_assemble(sg, g, gathered.data(), 0, ncol, b.ref.data(), b.ins_bytes.data(), e.hap, e.count == wo.depth, seqbuf.data(), b.seq_cap, &hx);
          if (hx.seq_len <= b.seq_cap) {
            hx.id64 = mph_record_id64(seqbuf.data(), hx.seq_len, b.tx_id_bytes.data() + b.tx_id_off[sg.tx], b.tx_id_off[sg.tx + 1] - b.tx_id_off[sg.tx], g.s);
            hx.flags |= MPH_NF_ID;
          }
          hx.seq_off = uint32_t(raw.seq.size());
          raw.seq.resize(raw.seq.size() + b.seq_cap, 0);
          memcpy(&raw.seq[hx.seq_off], seqbuf.data(), std::min<uint32_t>(hx.seq_len, b.seq_cap));
          hx.flags |= MPH_NF_SEQ;
          raw.hist.push_back(e);
          raw.hapx.push_back(hx);
        }
        raw.iw.push_back(widx);
        raw.iw_out.push_back(wo);
        raw.iw_hap0.push_back(rp.hap0[widx]);
        raw.iw_voff.resize(raw.iw.size(), 0xFFFFFFFFu);
        if (list) {
          raw.iw_voff.back() = uint32_t(raw.vlist.size());
          raw.vlist.insert(raw.vlist.end(), list, list + 1 + ncol);
        }
        continue;
      }
      const uint32_t va = mph_var_lb(b.vars.data(), sg.var_lo, sg.var_hi, g.s);
      const uint32_t vb = mph_var_lb(b.vars.data(), sg.var_lo, sg.var_hi, g.e);
      const uint32_t nv = vb - va;
      if (nv > 64) raw.err |= MPH_E_VARS_PER_WINDOW;
      uint32_t rlo, rhi;
      mph_candidate_range(sg, b.read_start.data(), g, &rlo, &rhi);
      uint32_t depth = 0;
      std::map<uint64_t, uint32_t> hist;
      for (uint32_t r = rlo; r < rhi; ++r) {
        const uint32_t st = b.read_start[r], en = b.read_end[r], vlo = read_vlo[r];
        if (!rev) {
          const uint32_t kp = mph_nrm_fwd_entry(sg, k, g, st, en);
          if (kp == 0xFFFFFFFFu) continue;
          ++depth;
          if (nv) hist[mph_nrm_hap(sg, b.vars.data(), kp, k, va, vb, vlo, calls[r].S)] += 1;
        } else {
          uint32_t kc;
          const uint32_t copies = mph_nrm_rev_copies(sg, k, g, st, en, &kc);
          if (!copies) continue;
          depth += copies;
          if (!nv) continue;
          if (calls[r].S == 0) { hist[0] += copies; continue; }
          for (uint32_t c = 0; c < copies; ++c) hist[mph_nrm_hap(sg, b.vars.data(), k - c, k, va, vb, vlo, calls[r].S)] += 1;
        }
      }
      raw.sum_depth += depth;
      MphHap h0;
      raw.err |= mph_nrm_plain(sg, g, b.ref.data(), nv, &h0);
      raw.win_depth[widx] = depth | ((nv == 0 && (h0.flags & MPH_NF_STOP)) ? 0x80000000u : 0u);
      if (!(h0.flags & MPH_NF_REFRANGE))
        raw.win_id[widx] = mph_record_id64(b.ref.data() + sg.ref_off + (g.s - sg.ref_pos0), g.e - g.s, b.tx_id_bytes.data() + b.tx_id_off[sg.tx],
                                          b.tx_id_off[sg.tx + 1] - b.tx_id_off[sg.tx], g.s);
      if (!nv) continue;
      MphWinOut wo;
      wo.depth = depth;
      wo.c0 = 0;
      wo.extra_off = uint32_t(raw.hist.size());
      wo.n_extra = 0;
      for (auto& kv : hist) {
        if (kv.first == 0) { wo.c0 = kv.second; continue; }
        MphHap hx;
        raw.err |= mph_nrm_assemble(sg, g, b.vars.data(), va, vb, b.ref.data(), b.ins_bytes.data(), kv.first, kv.second == depth, seqbuf.data(), b.seq_cap, &hx);
        if (hx.seq_len <= b.seq_cap) {
          hx.id64 = mph_record_id64(seqbuf.data(), hx.seq_len, b.tx_id_bytes.data() + b.tx_id_off[sg.tx], b.tx_id_off[sg.tx + 1] - b.tx_id_off[sg.tx], g.s);
          hx.flags |= MPH_NF_ID;
        }
        hx.seq_off = uint32_t(raw.seq.size());
        raw.seq.resize(raw.seq.size() + b.seq_cap, 0);
        memcpy(&raw.seq[hx.seq_off], seqbuf.data(), std::min<uint32_t>(hx.seq_len, b.seq_cap));
        hx.flags |= MPH_NF_SEQ;
        raw.hist.push_back(MphHist{kv.first, kv.second, 0});
        raw.hapx.push_back(hx);
        ++wo.n_extra;
      }
      if (sg.flags & MPH_SF_DEVREC) {
        d_win_out[widx] = wo;
        d_hap0[widx] = h0;
        continue;
      }
      raw.iw.push_back(widx);
      raw.iw_out.push_back(wo);
      raw.iw_hap0.push_back(h0);
    }
  }
  if (!b.replay.empty()) raw.iw_voff.resize(raw.iw.size(), 0xFFFFFFFFu);
  // record kernels of the normal mode: first window whose haplotypes all stop, then the records of the live windows
  {
    MphRecCtx c;
    c.segs = b.segs.data(); c.vars = b.vars.data(); c.ref = b.ref.data(); c.win_out = d_win_out.data(); c.hap0 = d_hap0.data();
    c.hist = raw.hist.data(); c.hapx = raw.hapx.data(); c.seq = raw.seq.data(); c.seq_cap = b.seq_cap;
    c.tx_id_bytes = b.tx_id_bytes.data(); c.tx_id_off = b.tx_id_off.data();
    MphNrmCtx n;
    n.win_depth = raw.win_depth.data();
    n.win_id = reinterpret_cast<const unsigned long long*>(raw.win_id.data());
    std::vector<uint32_t> tx_stop(b.txs.size(), 0xFFFFFFFFu);
    for (const MphSegment& sg : b.segs) {
      if (!(sg.flags & MPH_SF_DEVREC)) continue;
      for (uint32_t i = 0; i < sg.n_win; ++i)
        if (mph_nrc_window_stops(c, n, sg, i, sg.win_base + i)) tx_stop[sg.tx] = std::min(tx_stop[sg.tx], sg.win_base + i);
    }
    uint32_t err = 0;
    for (size_t si = 0; si < b.segs.size(); ++si) {
      const MphSegment& sg = b.segs[si];
      if (!(sg.flags & MPH_SF_DEVREC)) continue;
      for (uint32_t i = 0; i < sg.n_win; ++i) {
        const uint32_t widx = sg.win_base + i;
        if (widx > tx_stop[sg.tx]) break;
        raw.dev_windows += 1;
        raw.dev_read_windows += raw.win_depth[widx] & 0x7FFFFFFFu;
        uint32_t bytes = 0;
        const uint32_t cnt = mph_nrc_window_count(c, n, sg, i, widx, &bytes, &err);
        const size_t r0 = raw.recs.size(), s0 = raw.rec_seq.size();
        raw.recs.resize(r0 + cnt);
        raw.rec_seq.resize(s0 + bytes);
        if (mph_nrc_window_emit(c, n, sg, i, widx, raw.recs.data() + r0, raw.rec_seq.data(), uint32_t(s0), &err) != cnt) err |= MPH_E_INTERNAL;
        const bool junction = i == 0 && !(sg.flags & MPH_SF_FIRST_EXON) && widx < tx_stop[sg.tx] && si > 0 && b.segs[si - 1].tx == sg.tx;
        if (junction) {
          const MphSegment& sp = b.segs[si - 1];
          const uint32_t ub = mph_nrc_merge_t<MphSerialOps>(c, n, sp, sg, b.window_len, nullptr, nullptr, nullptr, 0, 0, 0, &err);
          if (ub) {
            std::vector<MphRec> mr(ub);
            std::vector<MphRecSrc> ma(ub);
            const size_t sb = raw.rec_seq.size(), ab = raw.rec_aux.size();
            raw.rec_seq.resize(sb + size_t(ub) * MPH_RC_SEQ_SLOT, 0);
            const uint32_t nm = mph_nrc_merge_t<MphSerialOps>(c, n, sp, sg, b.window_len, mr.data(), ma.data(), raw.rec_seq.data(), uint32_t(ab), uint32_t(sb), ub, &err);
            for (uint32_t x = 0; x < nm; ++x) mph_rc_merged_id(c, &mr[x], raw.rec_seq.data() + mr[x].seq_off, b.window_len);
            raw.rec_aux.insert(raw.rec_aux.end(), ma.begin(), ma.begin() + nm);
            const size_t rb = raw.recs.size();
            raw.recs.resize(rb + nm);
            for (uint32_t x = 0; x < nm; ++x) raw.recs[rb + mr[x].rank] = mr[x];
          }
        }
      }
    }
    raw.err |= err;
  }
  return raw;
}

inline PhaseRaw phase_somatic(const Batch& b) {
  PhaseRaw raw;
  const size_t nr = b.read_start.size();
  // K1
  const std::vector<MphCall> calls = allele_calls(b, true);
  const PerRead per_read(b);
  const std::vector<uint32_t>& read_vlo = per_read.vlo;
  for (size_t r = 0; r < nr; ++r)
    if (b.read_flags[r] & MPH_RF_OVERFLOW) raw.err |= MPH_E_VARS_PER_WINDOW;
  ReplayStage rp;
  run_replay(b, calls, per_read, 0, raw, rp);
  std::vector<MphWinOut>& rp_out = rp.out;
  std::vector<MphHap>& rp_hap0 = rp.hap0;
  std::vector<uint32_t>&rp_voff = rp.voff, &rp_vlist = rp.vlist;
  std::vector<MphHist>& rp_hist = rp.hist;
  // K2 + K3 + K4
  std::vector<uint8_t> seqbuf(b.seq_cap), germbuf(b.seq_cap);
  // device-class transcripts (core/record_core.h): per-window arrays and their own sequence arena, as on the device
  std::vector<MphWinOut> d_win_out(b.n_windows);
  std::vector<MphHap> d_hap0(b.n_windows);
  std::vector<uint8_t> d_flag(b.n_windows, 0), d_seq;
  for (const MphChunk& ch : b.chunks) {
    const MphSegment& sg = b.segs[ch.seg];
    const bool rev = sg.flags & MPH_SF_REVERSE;
    for (uint32_t i = ch.i_first; i < ch.i_first + ch.n; ++i) {
      const uint32_t k = sg.k_first + i * sg.k_stride;
      const uint32_t widx = sg.win_base + i;
      const MphGeom g = mph_geom(sg, k);
      if (sg.flags & MPH_SF_REPLAY) {
        // K3 on the replayed window: the walk runs over the matrix columns, gathered into a dense array
        MphWinOut wo = rp_out[widx];
        const uint32_t* list = rp_voff[widx] == 0xFFFFFFFFu ? nullptr : &rp_vlist[rp_voff[widx]];
        const uint32_t ncol = list ? list[0] : 0;
        std::vector<MphVar> gathered(ncol + 1);
        for (uint32_t j = 0; j < ncol; ++j) gathered[j] = b.vars[list[1 + j]];
        const uint32_t src = wo.extra_off;
        wo.extra_off = uint32_t(raw.hist.size());
        for (uint32_t x = 0; x < wo.n_extra; ++x) {
          const MphHist e = rp_hist[src + x];
          MphHap hx;
          memset(&hx, 0, sizeof hx);
          if (e.hap != 0) {
            raw.err |= mph_assemble(sg, g, gathered.data(), 0, ncol, b.ref.data(), b.ins_bytes.data(), e.hap, seqbuf.data(), germbuf.data(), b.seq_cap, &hx);
            if ((hx.n_som > 0 || (sg.flags & MPH_SF_HAS_FS)) && hx.seq_len <= b.seq_cap) {
              hx.id64 = mph_record_id64(seqbuf.data(), hx.seq_len, b.tx_id_bytes.data() + b.tx_id_off[sg.tx], b.tx_id_off[sg.tx + 1] - b.tx_id_off[sg.tx], g.s);
              hx.flags |= MPH_HF_ID;
            }
            hx.seq_off = uint32_t(raw.seq.size());
            raw.seq.resize(raw.seq.size() + 2 * b.seq_cap, 0);
            memcpy(&raw.seq[hx.seq_off], seqbuf.data(), std::min<uint32_t>(hx.seq_len, b.seq_cap));
            memcpy(&raw.seq[hx.seq_off + b.seq_cap], germbuf.data(), std::min<uint32_t>(hx.germ_len, b.seq_cap));
            hx.flags |= MPH_HF_SEQ;
          }
          raw.hist.push_back(e);
          raw.hapx.push_back(hx);
        }
        raw.iw.push_back(widx);
        raw.iw_out.push_back(wo);
        raw.iw_hap0.push_back(rp_hap0[widx]);
        raw.iw_voff.resize(raw.iw.size(), 0xFFFFFFFFu);
        if (list) {
          raw.iw_voff.back() = uint32_t(raw.vlist.size());
          raw.vlist.insert(raw.vlist.end(), list, list + 1 + ncol);
        }
        continue;
      }
      const uint32_t va = mph_var_lb(b.vars.data(), sg.var_lo, sg.var_hi, g.s);
      const uint32_t vb = mph_var_lb(b.vars.data(), sg.var_lo, sg.var_hi, g.e);
      if (vb - va > 64) raw.err |= MPH_E_VARS_PER_WINDOW;
      uint32_t rlo, rhi;
      mph_candidate_range(sg, b.read_start.data(), g, &rlo, &rhi);
      uint32_t depth = 0;
      std::map<std::tuple<uint64_t, uint32_t, uint32_t>, uint32_t> hist;  // (hap, frame0, f1nz) -> count
      for (uint32_t r = rlo; r < rhi; ++r) {
        MphPair p;
        const uint32_t st = b.read_start[r], en = b.read_end[r], vlo = read_vlo[r];
        if (!rev) {
          p = mph_fwd_state(sg, b.vars.data(), k, g, va, vb, st, en, vlo, calls[r].S, calls[r].B);
        } else {
          p.member = 0;
          if (st <= g.s && en >= g.e) {
            const uint64_t Bx = calls[r].B | (calls[r].S & mph_range_mask(sg.sl_va, sg.sl_vb, vlo));
            uint32_t ke = mph_rev_entry(sg, b.vars.data(), k, st, en, vlo, Bx);
            if (ke != 0xFFFFFFFFu && (b.read_flags[r] & MPH_RF_PARTNER)) {
              const uint32_t q = partner_of(b, r);
              const uint64_t Bq = calls[q].B | (calls[q].S & mph_range_mask(sg.sl_va, sg.sl_vb, read_vlo[q]));
              uint32_t kq = 0xFFFFFFFFu;
              if (b.read_start[q] <= g.s && b.read_end[q] >= g.e) kq = mph_rev_entry(sg, b.vars.data(), k, b.read_start[q], b.read_end[q], read_vlo[q], Bq);
              if (kq != 0xFFFFFFFFu && (kq < ke || (kq == ke && q < r))) ke = 0xFFFFFFFFu;
            }
            if (ke != 0xFFFFFFFFu) p = mph_rev_state(sg, b.vars.data(), k, g, va, vb, st, en, vlo, calls[r].S, calls[r].B, ke);
          }
        }
        if (!p.member) continue;
        ++depth;
        if (p.bad) continue;
        hist[{p.hap, p.frame & 0x7FFFFFFFu, p.frame >> 31}] += 1;
      }
      raw.sum_depth += depth;
      MphWinOut wo;
      wo.depth = depth;
      wo.c0 = 0;
      wo.extra_off = uint32_t(raw.hist.size());
      wo.n_extra = 0;
      std::vector<MphHist> extras;
      for (auto& kv : hist) {
        if (std::get<0>(kv.first) == 0 && std::get<1>(kv.first) == 0 && std::get<2>(kv.first) == 0) { wo.c0 = kv.second; continue; }
        extras.push_back(MphHist{std::get<0>(kv.first), kv.second, std::get<1>(kv.first) | (std::get<2>(kv.first) << 31)});
      }
      wo.n_extra = uint32_t(extras.size());
      // K3
      const bool boundary = mph_is_boundary(sg, i);
      const bool devrec = (sg.flags & MPH_SF_DEVREC) != 0;
      std::vector<uint8_t>& seq_arena = devrec ? d_seq : raw.seq;
      auto assemble = [&](uint64_t hap, MphHap* out) {
        if (hap == 0) {
          raw.err |= mph_plain_hap(sg, g, b.stopmap.data(), b.ref.data(), vb - va, out);
          return;
        }
        raw.err |= mph_assemble(sg, g, b.vars.data(), va, vb, b.ref.data(), b.ins_bytes.data(), hap, seqbuf.data(), germbuf.data(), b.seq_cap, out);
        if ((out->n_som > 0 || (sg.flags & MPH_SF_HAS_FS)) && out->seq_len <= b.seq_cap) {
          out->id64 = mph_record_id64(seqbuf.data(), out->seq_len, b.tx_id_bytes.data() + b.tx_id_off[sg.tx], b.tx_id_off[sg.tx + 1] - b.tx_id_off[sg.tx], g.s);
          out->flags |= MPH_HF_ID;
        }
        if (boundary || out->n_som > 0) {
          out->seq_off = uint32_t(seq_arena.size());
          seq_arena.resize(seq_arena.size() + 2 * b.seq_cap, 0);
          memcpy(&seq_arena[out->seq_off], seqbuf.data(), std::min<uint32_t>(out->seq_len, b.seq_cap));
          memcpy(&seq_arena[out->seq_off + b.seq_cap], germbuf.data(), std::min<uint32_t>(out->germ_len, b.seq_cap));
          out->flags |= MPH_HF_SEQ;
        }
      };
      MphHap h0;
      assemble(0, &h0);
      bool interesting = va != vb || (h0.flags & MPH_HF_STOP) || boundary || wo.n_extra > 0;
      for (auto& e : extras) {
        MphHap hx;
        assemble(e.hap, &hx);
        raw.hist.push_back(e);
        raw.hapx.push_back(hx);
      }
      if (devrec) {
        d_win_out[widx] = wo;
        d_hap0[widx] = h0;
        d_flag[widx] = interesting ? 2 : 0;
      } else if (interesting) {
        raw.iw.push_back(widx);
        raw.iw_out.push_back(wo);
        raw.iw_hap0.push_back(h0);
      }
    }
  }
  if (!b.replay.empty()) raw.iw_voff.resize(raw.iw.size(), 0xFFFFFFFFu);
  // record kernels (kernels/record_kernels.cu): first removing window per transcript, then the records of the live windows
  {
    MphRecCtx c;
    c.segs = b.segs.data(); c.vars = b.vars.data(); c.ref = b.ref.data(); c.win_out = d_win_out.data(); c.hap0 = d_hap0.data();
    c.hist = raw.hist.data(); c.hapx = raw.hapx.data(); c.seq = d_seq.data(); c.seq_cap = b.seq_cap;
    c.tx_id_bytes = b.tx_id_bytes.data(); c.tx_id_off = b.tx_id_off.data();
    std::vector<uint32_t> tx_stop(b.txs.size(), 0xFFFFFFFFu), stopq(b.n_windows, 0xFFFFFFFFu);
    for (const MphSegment& sg : b.segs) {
      if (!(sg.flags & MPH_SF_DEVREC)) continue;
      for (uint32_t i = 0; i < sg.n_win; ++i) {
        const uint32_t widx = sg.win_base + i;
        if (d_flag[widx] != 2) continue;
        const uint32_t q = mph_rc_window_stop(c, sg, i, widx);
        if (q != 0xFFFFFFFFu) { stopq[widx] = q; tx_stop[sg.tx] = std::min(tx_stop[sg.tx], widx); }
      }
    }
    uint32_t err = 0;
    for (size_t si = 0; si < b.segs.size(); ++si) {
      const MphSegment& sg = b.segs[si];
      if (!(sg.flags & MPH_SF_DEVREC)) continue;
      for (uint32_t i = 0; i < sg.n_win; ++i) {
        const uint32_t widx = sg.win_base + i;
        if (widx > tx_stop[sg.tx]) break;
        raw.dev_windows += 1;
        raw.dev_read_windows += d_win_out[widx].depth;
        if (d_flag[widx] != 2) continue;
        uint32_t bytes = 0;
        const uint32_t n = mph_rc_window_count(c, sg, i, widx, stopq[widx], &bytes, &err);
        const size_t r0 = raw.recs.size(), s0 = raw.rec_seq.size();
        raw.recs.resize(r0 + n);
        raw.rec_seq.resize(s0 + bytes);
        const uint32_t wrote = mph_rc_window_emit(c, sg, i, widx, stopq[widx], raw.recs.data() + r0, raw.rec_seq.data(), uint32_t(s0), &err);
        if (wrote != n) err |= MPH_E_INTERNAL;
        const bool junction = i == 0 && !(sg.flags & MPH_SF_FIRST_EXON) && (sg.flags & MPH_SF_JOIN_HEAD) && widx < tx_stop[sg.tx] && si > 0 &&
                              b.segs[si - 1].tx == sg.tx;
        if (junction) {
          const MphSegment& sp = b.segs[si - 1];
          const uint32_t ub = mph_rc_merge(c, sp, sg, b.window_len, nullptr, nullptr, nullptr, 0, 0, 0, &err);
          if (getenv("MPH_EMU_JSTAT")) fprintf(stderr, "J %u %u %u\n", mph_rc_nkeys(d_win_out[sg.win_base]), mph_rc_nkeys(d_win_out[sp.win_base + sp.n_win - 1]), ub);
          if (ub) {
            std::vector<MphRec> mr(ub);
            std::vector<MphRecSrc> ma(ub);
            const size_t sb = raw.rec_seq.size(), ab = raw.rec_aux.size();
            raw.rec_seq.resize(sb + size_t(ub) * MPH_RC_SEQ_SLOT, 0);
            const uint32_t nm = mph_rc_merge(c, sp, sg, b.window_len, mr.data(), ma.data(), raw.rec_seq.data(), uint32_t(ab), uint32_t(sb), ub, &err);
            for (uint32_t x = 0; x < nm; ++x) mph_rc_merged_id(c, &mr[x], raw.rec_seq.data() + mr[x].seq_off, b.window_len);
            raw.rec_aux.insert(raw.rec_aux.end(), ma.begin(), ma.begin() + nm);
            const size_t rb = raw.recs.size();
            raw.recs.resize(rb + nm);
            for (uint32_t x = 0; x < nm; ++x) raw.recs[rb + mr[x].rank] = mr[x];
          }
        }
      }
    }
    raw.err |= err;
  }
  return raw;
}

inline PhaseRaw phase(const Batch& b) { return b.mode == 1 ? phase_normal(b) : phase_somatic(b); }

}  // namespace mphemu
