// batch.hpp — host-side batch: the structure-of-arrays buffers that cross the C ABI
// (include/microphaser_gpu.h: mph_batch_in) plus the host-only metadata needed to print records.
//
// The Packer restates the data-independent part of the reference's per-gene set-up and exon loop
// (reference src/microphasing.rs:894-1029): read selection (mapq >= 5, :910), max_read_len (:913),
// variant_tree flattening (:932-942), per-exon current_exon_offset / is_short_exon /
// exon_window_len / start offset (:989-1019) and the exon_rest carry between exons (:1386-1400).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../core/layout.h"
#include "../core/replay_core.h"

namespace mph {

struct Fatal : std::runtime_error {  // the reference panics (exit status 101)
  using std::runtime_error::runtime_error;
};
struct Unsupported : std::runtime_error {  // input needs the serial path that is not built yet
  using std::runtime_error::runtime_error;
};

struct HostVariant {
  uint32_t pos = 0, len = 0;
  uint8_t kind = MPH_SNV, alt = 0;
  bool germline = true;
  std::string ins;          // whole ALT allele for insertions
  std::string prot_change;  // common.rs:23-33
  uint32_t fs() const {     // common.rs:215-221
    if (kind == MPH_SNV) return 0;
    if (kind == MPH_DEL) return len % 3;
    return (3 - ((uint32_t(ins.size()) - 1) % 3)) % 3;
  }
};

struct HostRead {  // what the host parser hands over per alignment record
  uint32_t start = 0, end = 0, l_seq = 0;
  const uint8_t* seq4 = nullptr;  // BAM 4-bit bases
  const uint8_t* qual = nullptr;  // raw phred
  const uint32_t* cigar = nullptr;
  uint32_t n_cigar = 0;
  uint64_t qname_hash = 0;
};

struct HostExon {
  uint32_t start, end, frame;
};
struct HostTranscript {
  std::string id;
  bool reverse = false;
  std::vector<HostExon> exons;
};
struct HostGene {
  std::string id, name, chrom;
  uint32_t start = 0, end = 0;
  std::vector<HostTranscript> transcripts;
};

struct TxMeta {
  std::string id;
  uint32_t gene;  // index into Batch::genes
  bool reverse;
  uint32_t seg_lo, seg_hi;
};
struct GeneMeta {
  std::string id, name, chrom;
  uint32_t start, end;
  uint32_t var_lo, var_hi, read_lo, read_hi;
};

// sizes of every per-gene-contiguous array after a gene has been packed: a batch can be cut at any gene boundary
struct GeneMark {
  uint64_t reads = 0, vr = 0, bases = 0, cigars = 0, vars = 0, ins = 0, segs = 0, chunks = 0, ref = 0, windows = 0, txs = 0, replay = 0, dq = 0, partners = 0;
  uint64_t runs = 0, span_exc = 0, flag_exc = 0, vruns = 0, ncig_exc = 0, mspan_exc = 0;
};

struct Batch {
  uint32_t window_len = 27;
  std::vector<GeneMark> marks{GeneMark{}};  // marks[g + 1]: sizes after gene g
  int mode = 0;  // 0 somatic (src/microphasing.rs), 1 normal (src/normal_microphasing.rs)
  // reads (SoA)
  std::vector<uint32_t> read_start, read_end;
  std::vector<uint8_t> read_flags;
  // what crosses the bus instead of the three arrays above (9 B per read): starts are sorted within a gene and the
  // reference span of a short read fits a byte, so a read ships as 2 B - the distance to the previous read's start and
  // its span - plus a run table (a new run wherever the distance exceeds a byte, and at every gene) and two exception
  // lists (spans >= 255, non-zero flags). K0 (k_read_decode) rebuilds read_start / read_end / read_flags on the device.
  std::vector<uint8_t> rd_delta, rd_span;
  struct U2 { uint32_t x, y; };
  std::vector<U2> rd_runs;      // (first read, its start)
  std::vector<U2> rd_span_exc;  // (read, end)
  std::vector<U2> rd_flag_exc;  // (read, flags)
  // Second bus form of the spans: short reads of one sequencing run share one reference span unless they carry an indel or
  // a clip, so the batch ships ONE span (the commonest of the first gene with a few reads) and lists the reads that differ
  // (read, end) - 1 B per read less on the bus. The packer writes both forms; bus_span_mode() picks per batch (trimmed
  // reads of all lengths keep the byte array).
  uint32_t modal_span = 0;
  bool modal_set = false;
  std::vector<U2> rd_mspan_exc;  // (read, end) of every read whose span is not modal_span
  // bus form of the side table (9 B per entry instead of 21): the entries are sorted by read, their first-variant index
  // never decreases within a gene, and the offsets of the packed bases / CIGARs are running sums of the record sizes, so an
  // entry ships as read distance (u16), variant-index distance (u8), record bytes (u16), CIGAR ops (u8; 255 = see vs_ncig_exc)
  // next to vr_lseq / vr_nv; a run table restarts the sums at every gene and wherever a distance does not fit.
  std::vector<uint16_t> vs_read_d, vs_size;
  std::vector<uint8_t> vs_vlo_d, vs_ncig;
  struct VRun { uint32_t entry, read, vlo, seq_off, cig_off; };
  std::vector<VRun> vs_runs;
  std::vector<U2> vs_ncig_exc;  // (entry, CIGAR ops) for entries with 255 and more
  // compact side table of the reads K1 has work for (variants inside the alignment; every read of a gene with replayed
  // transcripts): read index, first variant index, offsets of the packed bases / CIGAR, lengths, variant count
  std::vector<uint32_t> vr_read, vr_vlo, vr_seq_off, vr_cig_off;
  std::vector<uint16_t> vr_lseq, vr_ncig;
  std::vector<uint8_t> vr_nv;
  std::vector<uint32_t> partner_a, partner_b;  // sorted pairs (a < b) of reads with identical (start, qname), groups of exactly two
  // directed edges read -> next read with the same (start, qname): both directions of every pair above, and a cycle through
  // every larger group (those genes go through the serial replay, which walks the cycle for `contains`, :281-294)
  std::vector<std::pair<uint32_t, uint32_t>> pair_edges;
  std::vector<uint8_t> bases;                  // 16-B aligned packed records: 4-bit bases + (qual<10) bits
  std::vector<uint32_t> cigars;
  // variants
  std::vector<MphVar> vars;
  std::vector<uint8_t> ins_bytes;
  std::vector<std::string> var_prot;  // host only
  // geometry
  std::vector<MphSegment> segs;
  std::vector<MphChunk> chunks;
  // per segment: the union of its windows' candidate reads and variants (the read-run kernel visits every (segment, read) pair
  // of these ranges once); seg_work_off = running sum of the range lengths, n_segs + 1 entries
  std::vector<MphSegWork> seg_work;
  std::vector<uint32_t> seg_work_off{0};
  std::vector<uint8_t> ref;  // per-segment reference slices
  std::vector<uint32_t> stopmap;  // 1 bit per ref byte: a stop codon (for the slice's strand) starts here; 2 words of slack
  uint64_t n_windows = 0;
  uint32_t seq_cap = 64;  // bytes per assembled sequence slot
  // transcripts that go through the serial replay (core/replay_core.h) and the size of their observation scratch
  std::vector<MphReplayTx> replay;
  std::vector<uint32_t> replay_dq;  // initial matrix columns of the replay units
  uint64_t replay_obs = 0;
  std::vector<uint32_t> seg_chunk0;  // per segment: index of its first chunk
  // transcript ids for the record ids hashed on the device (:667-675): byte arena + n_tx + 1 offsets
  std::vector<uint8_t> tx_id_bytes;
  std::vector<uint32_t> tx_id_off{0};
  // host-only
  std::vector<TxMeta> txs;
  std::vector<GeneMeta> genes;

  uint64_t n_reads() const { return read_start.size(); }
};

// which form of the spans crosses the bus: 1 = one span per batch + the reads that differ (8 B each), 0 = a byte per read
// (+ the reads with spans >= 255). MPH_BUS_SPAN_BYTES=1 forces the byte form (test hook).
inline int bus_span_mode(const Batch& b) {
  static const bool force_bytes = getenv("MPH_BUS_SPAN_BYTES") != nullptr;
  return (!force_bytes && b.modal_set && b.rd_mspan_exc.size() * 8 <= b.rd_span.size()) ? 1 : 0;
}

// host statement of K0 (k_read_decode / k_read_patch): used by the CPU checks of the bus encoding
inline void decode_reads(const Batch& b, std::vector<uint32_t>& start, std::vector<uint32_t>& end, std::vector<uint8_t>& flags, int form = -1) {
  const size_t n = b.rd_delta.size();
  start.assign(n, 0); end.assign(n, 0); flags.assign(n, 0);
  const int mode = form < 0 ? bus_span_mode(b) : form;
  for (size_t j = 0; j < b.rd_runs.size(); ++j) {
    const size_t lo = b.rd_runs[j].x, hi = j + 1 < b.rd_runs.size() ? b.rd_runs[j + 1].x : n;
    uint32_t pos = b.rd_runs[j].y;
    for (size_t r = lo; r < hi; ++r) {
      pos += b.rd_delta[r];
      start[r] = pos;
      end[r] = pos + (mode ? b.modal_span : uint32_t(b.rd_span[r]));
    }
  }
  for (auto& e : (mode ? b.rd_mspan_exc : b.rd_span_exc)) end[e.x] = e.y;
  for (auto& e : b.rd_flag_exc) flags[e.x] = uint8_t(e.y);
}

// host statement of k_side_decode: the side table from its bus form
inline void decode_side_table(const Batch& b, std::vector<uint32_t>& read, std::vector<uint32_t>& vlo, std::vector<uint32_t>& seq_off,
                              std::vector<uint32_t>& cig_off, std::vector<uint16_t>& ncig) {
  const size_t n = b.vs_read_d.size();
  read.assign(n, 0); vlo.assign(n, 0); seq_off.assign(n, 0); cig_off.assign(n, 0); ncig.assign(n, 0);
  for (size_t j = 0; j < b.vs_runs.size(); ++j) {
    const size_t lo = b.vs_runs[j].entry, hi = j + 1 < b.vs_runs.size() ? b.vs_runs[j + 1].entry : n;
    uint32_t r = b.vs_runs[j].read, v = b.vs_runs[j].vlo, so = b.vs_runs[j].seq_off, co = b.vs_runs[j].cig_off;
    for (size_t e = lo; e < hi; ++e) {
      r += b.vs_read_d[e]; v += b.vs_vlo_d[e];
      read[e] = r; vlo[e] = v; seq_off[e] = so; cig_off[e] = co; ncig[e] = b.vs_ncig[e];
      so += b.vs_size[e];
      co += b.vs_ncig[e] < 255 ? b.vs_ncig[e] : 0u;  // exceptions patch ncig and are followed by a fresh run
    }
  }
  for (auto& x : b.vs_ncig_exc) ncig[x.x] = uint16_t(x.y);
}

inline uint8_t base_code(uint8_t c) {
  static const char* dec = "=ACMGRSVTWYHKDBN";
  for (int i = 0; i < 16; ++i)
    if (uint8_t(dec[i]) == c) return uint8_t(i);
  return 0xFF;
}

class Packer {
 public:
  explicit Packer(uint32_t window_len, int mode = 0, uint32_t chunk_windows = 32) : chunk_windows_(chunk_windows) {
    b_.window_len = window_len;
    b_.mode = mode;
  }

  // `reads`: the records bam::RecordBuffer holds for this gene after the mapq filter, in file order;
  // `max_read_len`: max seq().len() over them (:913-915); `sites`: variant_tree in ascending position
  // order, each site holding its ALT alleles in VCF order (:937); `refseq`: bases of
  // [gene.start, gene.end + 100).
  void add_gene(const HostGene& g, const std::vector<HostRead>& reads, uint32_t max_read_len,
                const std::vector<std::vector<HostVariant>>& sites, std::vector<uint8_t> refseq) {
    // The order of the ALT alleles of a multi-allelic site follows the transcript's strand (print_haplotypes :373-379), and the
    // packed variant list has one order per gene. A gene with such a site and transcripts on both strands is packed as several
    // pseudo-genes, one per run of consecutive transcripts of the same strand (the transcripts keep their GTF order, all state
    // of the reference's loop is per transcript :953-971); each carries its own copy of the gene's reads and variants.
    bool multi = false;
    for (auto& site : sites) multi = multi || site.size() > 1;
    int strand = -1;
    bool mixed = false;
    for (auto& t : g.transcripts) {
      if (t.exons.empty()) continue;
      if (strand < 0) strand = t.reverse ? 1 : 0;
      else if (strand != (t.reverse ? 1 : 0)) mixed = true;
    }
    if (!(multi && mixed)) return add_gene_one(g, reads, max_read_len, sites, std::move(refseq));
    size_t i = 0;
    while (i < g.transcripts.size()) {
      HostGene part;
      part.id = g.id; part.name = g.name; part.chrom = g.chrom; part.start = g.start; part.end = g.end;
      int run = -1;
      for (; i < g.transcripts.size(); ++i) {
        const HostTranscript& t = g.transcripts[i];
        if (!t.exons.empty()) {
          const int st = t.reverse ? 1 : 0;
          if (run < 0) run = st;
          else if (run != st) break;
        }
        part.transcripts.push_back(t);
      }
      add_gene_one(part, reads, max_read_len, sites, refseq);
    }
  }

 private:
  void add_gene_one(const HostGene& g, const std::vector<HostRead>& reads, uint32_t max_read_len,
                    const std::vector<std::vector<HostVariant>>& sites, std::vector<uint8_t> refseq) {
    const uint32_t wl = b_.window_len;
    if (g.end > 0x7FF00000u) throw Unsupported("gene " + g.id + ": coordinates beyond 2^31 (the kernels use signed 32-bit window arithmetic)");
    GeneMeta gm;
    gm.id = g.id; gm.name = g.name; gm.chrom = g.chrom; gm.start = g.start; gm.end = g.end;
    const uint32_t gi = uint32_t(b_.genes.size());
    // strand decides the ALT order inside a multi-allelic site (print_haplotypes :373-379)
    int strand = -1;
    bool mixed = false, any_reverse = false;
    for (auto& t : g.transcripts) {
      if (t.exons.empty()) continue;
      any_reverse = any_reverse || t.reverse;
      if (strand < 0) strand = t.reverse ? 1 : 0;
      else if (strand != (t.reverse ? 1 : 0)) mixed = true;
    }
    bool multi = false;
    gm.var_lo = uint32_t(b_.vars.size());
    uint32_t max_del = 0, max_ins = 0, n_del = 0;
    bool has_fs = false;
    for (auto& site : sites) {
      if (site.size() > 1) multi = true;
      for (size_t a = 0; a < site.size(); ++a) {
        const HostVariant& hv = site[strand == 1 ? site.size() - 1 - a : a];
        MphVar v;
        memset(&v, 0, sizeof v);
        v.pos = hv.pos; v.len = hv.len; v.kind = hv.kind; v.alt = hv.alt;
        v.flags = uint8_t((hv.germline ? MPH_VF_GERMLINE : 0) | (hv.fs() << MPH_VF_FS_SHIFT));
        v.alt4 = hv.kind == MPH_SNV ? base_code(hv.alt) : 0xFF;
        if (hv.kind == MPH_INS) {
          v.ins_off = uint32_t(b_.ins_bytes.size());
          b_.ins_bytes.insert(b_.ins_bytes.end(), hv.ins.begin(), hv.ins.end());
          if (hv.len + 1 > max_ins) max_ins = hv.len + 1;
        }
        if (hv.kind == MPH_DEL && hv.len > max_del) max_del = hv.len;
        if (hv.kind == MPH_DEL) ++n_del;
        if (hv.fs()) has_fs = true;
        b_.vars.push_back(v);
        b_.var_prot.push_back(hv.prot_change);
      }
    }
    gm.var_hi = uint32_t(b_.vars.size());
    if (mixed && multi) throw std::logic_error("internal: a gene with multi-allelic sites and both strands reached the packer unsplit");
    // assembled sequences can grow by insertions / deleted reference bases
    uint32_t need = wl + 8 + 2 * (max_ins + max_del);
    need = (need + 15u) & ~15u;
    if (b_.mode == 1) need += 16;  // the extra reference base after every variant block (normal_microphasing.rs:476)
    if (need > b_.seq_cap) b_.seq_cap = need;

    // reads
    gm.read_lo = uint32_t(b_.read_start.size());
    uint32_t vcur = gm.var_lo, max_span = 0;
    std::unordered_map<uint64_t, uint32_t> seen;  // (start, qname) -> first read index
    std::map<uint32_t, std::vector<uint32_t>> big_groups;  // first read -> third and later reads of its group
    const size_t partners_mark = b_.partner_a.size();
    const size_t bases_mark = b_.bases.size(), cigars_mark = b_.cigars.size(), vr_mark = b_.vr_read.size();
    bool full_records = false;  // set for the re-pack of a gene with replayed transcripts: the replay looks at columns outside a read's own variants
    auto add_vr = [&](uint32_t idx, uint32_t vlo, uint32_t nv, const HostRead& r) {
      const size_t off = b_.bases.size();  // byte offset: the records are not aligned (K1 reads them byte-wise)
      static const bool force_wide_format = getenv("MPH_PACK_WIDE") != nullptr;  // test hook: 4-bit bases + bitmask for every read
      const bool single_m_read = r.n_cigar == 1 && (r.cigar[0] & 15u) == 0 && (r.cigar[0] >> 4) == r.l_seq;
      if (single_m_read && !full_records && !force_wide_format) {
        // column record (core/phase_core.h, format bit 2): an ungapped read's reference-coordinate map is the identity shift,
        // so the only bytes K1 can ever look at are the bases at the variant positions inside it: one byte per variant
        // (BAM 4-bit base, bit 4 = quality < 10, bit 5 = the position is inside the read) instead of the whole read
        b_.bases.resize(off + 1 + nv, 0);
        uint8_t* rec = &b_.bases[off];
        rec[0] = 4;
        for (uint32_t j = 0; j < nv; ++j) {
          const uint32_t rel = b_.vars[vlo + j].pos - r.start;
          if (rel >= r.l_seq) continue;
          const uint8_t c4 = (rel & 1u) ? (r.seq4[rel >> 1] & 15u) : (r.seq4[rel >> 1] >> 4);
          rec[1 + j] = uint8_t(c4 | ((b_.mode == 0 && r.qual[rel] < 10) ? 16u : 0u) | 32u);
        }
        b_.vr_read.push_back(idx);
        b_.vr_vlo.push_back(vlo);
        b_.vr_seq_off.push_back(uint32_t(off));
        b_.vr_lseq.push_back(uint16_t(r.l_seq));
        b_.vr_nv.push_back(uint8_t(nv));
        b_.vr_cig_off.push_back(0);
        b_.vr_ncig.push_back(0);
        return;
      }
      // packed record (core/phase_core.h): 2-bit bases when the read has only A C G T, the positions with
      // qual < 10 as a short list when there are few of them (the normal mode never tests qualities: empty list)
      bool acgt = !force_wide_format;
      for (uint32_t i = 0; i < r.l_seq && acgt; ++i) {
        const uint8_t c4 = (i & 1u) ? (r.seq4[i >> 1] & 15u) : (r.seq4[i >> 1] >> 4);
        acgt = c4 == 1 || c4 == 2 || c4 == 4 || c4 == 8;
      }
      uint32_t n_low = 0;
      if (b_.mode == 0)
        for (uint32_t i = 0; i < r.l_seq; ++i) n_low += r.qual[i] < 10;
      const bool low_list = b_.mode == 1 || (!force_wide_format && r.l_seq <= 256 && n_low <= 255 && n_low < (r.l_seq + 7u) / 8);
      const uint8_t fmt = uint8_t((acgt ? 1u : 0u) | (low_list ? 2u : 0u));
      const size_t nb = mph_rec_bases_bytes(fmt, r.l_seq), nq = low_list ? (b_.mode == 1 ? 0u : n_low) : (r.l_seq + 7u) / 8;
      b_.bases.resize(off + 2 + nb + nq, 0);
      uint8_t* rec = &b_.bases[off];
      rec[0] = fmt;
      rec[1] = uint8_t(low_list && b_.mode == 0 ? n_low : 0);
      if (acgt) {
        for (uint32_t i = 0; i < r.l_seq; ++i) {
          const uint8_t c4 = (i & 1u) ? (r.seq4[i >> 1] & 15u) : (r.seq4[i >> 1] >> 4);
          const uint8_t c2 = c4 == 1 ? 0 : (c4 == 2 ? 1 : (c4 == 4 ? 2 : 3));
          rec[2 + (i >> 2)] |= uint8_t(c2 << (6u - 2u * (i & 3u)));
        }
      } else {
        memcpy(rec + 2, r.seq4, nb);
      }
      if (b_.mode == 0) {
        uint8_t* lq = rec + 2 + nb;
        uint32_t x = 0;
        for (uint32_t i = 0; i < r.l_seq; ++i)
          if (r.qual[i] < 10) {
            if (low_list) lq[x++] = uint8_t(i);
            else lq[i >> 3] |= uint8_t(1u << (i & 7));
          }
      }
      b_.vr_read.push_back(idx);
      b_.vr_vlo.push_back(vlo);
      b_.vr_seq_off.push_back(uint32_t(off));
      b_.vr_lseq.push_back(uint16_t(r.l_seq));
      b_.vr_nv.push_back(uint8_t(nv));
      const bool single_m = r.n_cigar == 1 && (r.cigar[0] & 15u) == 0 && (r.cigar[0] >> 4) == r.l_seq;
      if (!single_m) {
        b_.vr_cig_off.push_back(uint32_t(b_.cigars.size()));
        b_.vr_ncig.push_back(uint16_t(r.n_cigar));
        b_.cigars.insert(b_.cigars.end(), r.cigar, r.cigar + r.n_cigar);
      } else {
        b_.vr_cig_off.push_back(0);
        b_.vr_ncig.push_back(0);
      }
    };
    std::vector<std::pair<uint32_t, uint32_t>> read_vrange;  // (vlo, nv) of every read of this gene, for the re-pack of replay genes
    read_vrange.reserve(reads.size());
    for (auto& r : reads) {
      const uint32_t idx = uint32_t(b_.read_start.size());
      while (vcur < gm.var_hi && b_.vars[vcur].pos < r.start) ++vcur;
      uint32_t ve = vcur;
      while (ve < gm.var_hi && b_.vars[ve].pos < r.end) ++ve;
      uint32_t nv = ve - vcur;
      uint8_t flags = 0;
      if (nv > 64) { nv = 64; flags |= MPH_RF_OVERFLOW; }
      read_vrange.emplace_back(vcur, nv);
      b_.read_start.push_back(r.start);
      b_.read_end.push_back(r.end);
      if (r.l_seq > 0xFFFF) throw Unsupported("read longer than 65535 bases");
      if (nv > 0) add_vr(idx, vcur, nv, r);
      if (any_reverse && b_.mode == 0) {  // `contains` only bites on the reverse strand (keys are read starts, :328-331); normal mode has none
        const uint64_t key = r.qname_hash * 0x9E3779B97F4A7C15ull ^ (uint64_t(r.start) << 1);
        auto it = seen.find(key);
        if (it == seen.end()) {
          seen.emplace(key, idx);
        } else {
          const uint32_t first = it->second;
          flags |= MPH_RF_PARTNER;
          if (b_.read_flags[first] & MPH_RF_PARTNER) {
            big_groups[first].push_back(idx);  // a third (or later) read of the group
          } else {
            b_.read_flags[first] |= MPH_RF_PARTNER;
            b_.partner_a.push_back(first);
            b_.partner_b.push_back(idx);
          }
        }
      }
      b_.read_flags.push_back(flags);
      if (r.end - r.start > max_span) max_span = r.end - r.start;
    }
    gm.read_hi = uint32_t(b_.read_start.size());
    // `contains` edges: pairs in both directions; a group of three or more becomes a cycle and sends the gene's reverse-strand
    // transcripts through the serial replay (the closed form of `contains` knows one partner per read)
    bool force_replay = false;
    for (size_t pi = partners_mark; pi < b_.partner_a.size();) {
      const uint32_t a = b_.partner_a[pi], bq = b_.partner_b[pi];
      auto bg = big_groups.find(a);
      if (bg == big_groups.end()) {
        b_.pair_edges.emplace_back(a, bq);
        b_.pair_edges.emplace_back(bq, a);
        ++pi;
        continue;
      }
      std::vector<uint32_t> members{a, bq};
      members.insert(members.end(), bg->second.begin(), bg->second.end());
      for (size_t m = 0; m < members.size(); ++m) b_.pair_edges.emplace_back(members[m], members[(m + 1) % members.size()]);
      force_replay = true;
      b_.partner_a.erase(b_.partner_a.begin() + long(pi));
      b_.partner_b.erase(b_.partner_b.begin() + long(pi));
    }

    // transcripts -> segments
    const uint32_t ref_end = g.start + uint32_t(refseq.size());
    // somatic: one deletion can pull the walk past the window end (:560-563); normal: every applied
    // deletion moves window_end (normal_microphasing.rs:457), so the overhang adds up
    const uint32_t margin = b_.mode == 1 ? (max_del + 1) * std::min<uint32_t>(n_del, wl) + 2 : max_del + 2;
    bool gene_replay = false;
    for (auto& t : g.transcripts) {
      if (t.exons.empty()) continue;  // is_coding (:947)
      TxMeta tm;
      tm.id = t.id; tm.gene = gi; tm.reverse = t.reverse;
      tm.seg_lo = uint32_t(b_.segs.size());
      const uint32_t txi = uint32_t(b_.txs.size());
      uint64_t exon_rest = 0;
      uint32_t exon_count = 0;
      const size_t exon_number = t.exons.size();
      bool tx_replay = force_replay && t.reverse && b_.mode == 0;
      for (auto& ex : t.exons) {
        if (ex.start > ex.end) continue;  // :981
        exon_count += 1;
        const uint64_t exon_len = ex.end - ex.start;
        // somatic: the first exon takes the GTF frame column (microphasing.rs:989-995); normal: exon_rest only (normal_microphasing.rs:739-742)
        const uint64_t ceo = (exon_count == 1 && b_.mode == 0) ? ex.frame : (exon_rest == 0 ? 0 : 3 - exon_rest);
        const bool is_short = exon_len < 3 ? true : uint64_t(wl) >= exon_len - ceo - (3 - ceo) % 3;
        if (ceo > exon_len || ceo > 2) throw Unsupported("transcript " + t.id + ": exon offset " + std::to_string(ceo) + " does not fit the exon");
        uint64_t ewl = !is_short ? wl : (exon_len - ceo) - ((exon_len - ceo) % 3);
        if (ewl == 0) ewl = exon_len;
        exon_rest = 0;
        if (max_read_len < ewl) continue;  // :1046 — loop breaks before touching the matrix
        MphSegment sg;
        memset(&sg, 0, sizeof sg);
        sg.exon_start = ex.start; sg.exon_end = ex.end;
        sg.ewl = uint32_t(ewl); sg.ceo = uint32_t(ceo);
        sg.flags = (t.reverse ? MPH_SF_REVERSE : 0) | (is_short ? MPH_SF_SHORT : 0) | (exon_count == 1 ? MPH_SF_FIRST_EXON : 0) |
                   (exon_count == exon_number ? MPH_SF_LAST_EXON : 0) | (has_fs ? MPH_SF_HAS_FS : 0);
        if (t.reverse) {
          if (uint64_t(ex.end) < ewl + ceo) throw Unsupported("transcript " + t.id + ": exon offset underflow");
          sg.off0 = uint32_t(ex.end - ewl - ceo);
          if (sg.off0 < ex.start) continue;  // !valid at the first iteration
          sg.n_iter = is_short ? 1 : sg.off0 - ex.start + 1;
        } else {
          sg.off0 = uint32_t(ex.start + ceo);
          if (uint64_t(sg.off0) + ewl > ex.end) continue;
          sg.n_iter = is_short ? 1 : uint32_t(ex.end - ewl - sg.off0 + 1);
        }
        // reference quirk: at the last iteration of a reverse-strand exon `offset == old_offset`
        // (:1159) suppresses the deletion of the variants at exon.start + ewl, which then stay in
        // the matrix for the rest of the transcript. Only the serial replay reproduces that.
        if (t.reverse && !is_short && sg.n_iter >= 2) {
          const uint32_t pstar = ex.start + uint32_t(ewl);
          const uint32_t vi = mph_var_lb(b_.vars.data(), gm.var_lo, gm.var_hi, pstar);
          if (vi < gm.var_hi && b_.vars[vi].pos == pstar) tx_replay = true;
          // second quirk: when the first window is already a last window (rest < 3) the later iterations move the
          // window start down to exon.start, but `reached_end` (:1136-1139) keeps the variants entering there out of
          // the matrix while last_window_vars still counts them
          if (sg.off0 - ex.start < 3) {
            const uint32_t v0 = mph_var_lb(b_.vars.data(), gm.var_lo, gm.var_hi, ex.start);
            if (v0 < gm.var_hi && b_.vars[v0].pos < sg.off0) tx_replay = true;
          }
        }
        sg.K = uint32_t(max_read_len - ewl);
        if (!t.reverse && uint64_t(sg.off0 - sg.ceo) < sg.K) throw Fatal("range start is greater than range end in BTreeMap");
        sg.read_lo = gm.read_lo; sg.read_hi = gm.read_hi;
        sg.var_lo = gm.var_lo; sg.var_hi = gm.var_hi;
        sg.max_span = max_span;
        if (exon_count == 1 && b_.mode == 0) {  // start-loss positions (:1305-1316); normal mode has none
          const uint32_t lo = t.reverse ? (ex.end >= 3 ? ex.end - 3 : 0) : ex.start;
          const uint32_t hi = t.reverse ? ex.end : ex.start + 3;
          sg.sl_va = mph_var_lb(b_.vars.data(), gm.var_lo, gm.var_hi, lo);
          sg.sl_vb = mph_var_lb(b_.vars.data(), gm.var_lo, gm.var_hi, hi);
        }
        // windows: main-ORF iterations only, unless the gene has frameshifting variants
        if (is_short) { sg.k_first = 0; sg.k_stride = 1; sg.n_win = 1; }
        else if (has_fs) { sg.k_first = 0; sg.k_stride = 1; sg.n_win = sg.n_iter; }
        else {
          sg.k_stride = 3;
          sg.k_first = t.reverse ? (3 - uint32_t(ewl % 3)) % 3 : 0;  // coding_shift % 3 == ceo % 3 (:1372-1381)
          sg.n_win = sg.k_first < sg.n_iter ? (sg.n_iter - 1 - sg.k_first) / 3 + 1 : 0;
        }
        // exon_rest after this exon = rest at the last main-ORF window (:1386-1400)
        {
          uint32_t k_last_main;
          bool any = true;
          if (is_short) k_last_main = 0;
          else {
            const uint32_t kf = t.reverse ? (3 - uint32_t(ewl % 3)) % 3 : 0;
            if (kf >= sg.n_iter) any = false;
            k_last_main = any ? kf + ((sg.n_iter - 1 - kf) / 3) * 3 : 0;
          }
          if (any) {
            exon_rest = t.reverse ? (sg.off0 - k_last_main) - ex.start : ex.end - (sg.off0 + k_last_main + ewl);
            if (ewl < 3) exon_rest = ewl;
          }
        }
        // reference slice: the exon plus the overhang a deletion can reach (:560-563)
        // An exon that sticks out of [gene.start, gene.end + 100) has no bytes there: the reference panics on the slice
        // (:464-471) when - and only when - its loop gets to such a window, so the slice is clipped here and the windows
        // outside raise MPH_HF_REFRANGE when they are reached.
        const uint32_t slice_lo = std::min(std::max(ex.start, g.start), ref_end);
        const uint32_t slice_end = std::max<uint64_t>(slice_lo, std::min<uint64_t>(uint64_t(ex.end) + margin, ref_end));
        sg.ref_pos0 = slice_lo;
        sg.ref_off = uint32_t(b_.ref.size());
        sg.ref_len = slice_end - slice_lo;
        b_.ref.insert(b_.ref.end(), refseq.begin() + (slice_lo - g.start), refseq.begin() + (slice_end - g.start));
        b_.stopmap.resize(b_.ref.size() / 32 + 4, 0);
        for (uint32_t x = 0; x + 3 <= sg.ref_len; ++x) {  // case-sensitive like has_stop_codon (:42-76)
          const uint8_t* c = &b_.ref[sg.ref_off + x];
          const bool stop = t.reverse ? (c[2] == 'A' && ((c[0] == 'T' && (c[1] == 'C' || c[1] == 'T')) || (c[0] == 'C' && c[1] == 'T')))
                                      : (c[0] == 'T' && ((c[1] == 'G' && c[2] == 'A') || (c[1] == 'A' && (c[2] == 'G' || c[2] == 'A'))));
          if (stop) b_.stopmap[(sg.ref_off + x) >> 5] |= 1u << ((sg.ref_off + x) & 31);
        }
        sg.tx = txi;
        sg.win_base = uint32_t(b_.n_windows);
        b_.n_windows += sg.n_win;
        const uint32_t si = uint32_t(b_.segs.size());
        // carry-over of observations from the previous exon needs the serial path
        if (si > tm.seg_lo && carries_over(b_.segs[si - 1], sg)) tx_replay = true;
        if (si > tm.seg_lo && sg.n_win == 1 && !is_short) b_.segs[si - 1].flags |= MPH_SF_KEEP_PENULT;
        b_.segs.push_back(sg);
        b_.seg_chunk0.push_back(uint32_t(b_.chunks.size()));
        for (uint32_t i = 0; i < sg.n_win; i += chunk_windows_) {
          MphChunk c;
          c.seg = si; c.i_first = i; c.n = std::min(chunk_windows_, sg.n_win - i); c.pad = 0;
          // s and e are monotone in the iteration number (non-decreasing forward, non-increasing reverse)
          const MphGeom ga = mph_geom(sg, sg.k_first + i * sg.k_stride), gz = mph_geom(sg, sg.k_first + (i + c.n - 1) * sg.k_stride);
          const uint32_t s_min = std::min(ga.s, gz.s), s_max = std::max(ga.s, gz.s), e_min = std::min(ga.e, gz.e), e_max = std::max(ga.e, gz.e);
          int64_t lo = t.reverse ? int64_t(s_min) - int64_t(sg.K) : int64_t(sg.off0 - sg.ceo) - int64_t(sg.K);
          lo = std::max<int64_t>(lo, int64_t(e_min) - int64_t(sg.max_span));
          lo = std::max<int64_t>(lo, 0);
          c.rlo = mph_u32_lb(b_.read_start.data(), sg.read_lo, sg.read_hi, uint32_t(lo));
          c.rhi = mph_u32_lb(b_.read_start.data(), c.rlo, sg.read_hi, s_max + 1u);
          c.va0 = mph_var_lb(b_.vars.data(), sg.var_lo, sg.var_hi, s_min);
          c.vb1 = mph_var_lb(b_.vars.data(), c.va0, sg.var_hi, e_max);
          b_.chunks.push_back(c);
        }
        {
          MphSegWork sw;
          sw.rlo = sw.rhi = sw.va0 = sw.vb1 = 0;
          const uint32_t c_first = b_.seg_chunk0.back(), c_end = uint32_t(b_.chunks.size());
          for (uint32_t c = c_first; c < c_end; ++c) {  // both unions are intervals: s and e are monotone in the iteration number
            const MphChunk& ch = b_.chunks[c];
            if (c == c_first) { sw.rlo = ch.rlo; sw.rhi = ch.rhi; sw.va0 = ch.va0; sw.vb1 = ch.vb1; }
            else { sw.rlo = std::min(sw.rlo, ch.rlo); sw.rhi = std::max(sw.rhi, ch.rhi); sw.va0 = std::min(sw.va0, ch.va0); sw.vb1 = std::max(sw.vb1, ch.vb1); }
          }
          b_.seg_work.push_back(sw);
        }
      }
      tm.seg_hi = uint32_t(b_.segs.size());
      // junctions that can produce records: a variant inside one of the two windows whose lists the merge reads
      {
        auto window_has_var = [&](const MphSegment& sg, uint32_t i) {
          if (i >= sg.n_win) return false;
          const MphGeom gw = mph_geom(sg, sg.k_first + i * sg.k_stride);
          const uint32_t a = mph_var_lb(b_.vars.data(), sg.var_lo, sg.var_hi, gw.s);
          return a < sg.var_hi && b_.vars[a].pos < gw.e;
        };
        for (uint32_t si = tm.seg_lo; si < tm.seg_hi; ++si) {
          MphSegment& sg = b_.segs[si];
          if (si == tm.seg_lo) sg.flags |= MPH_SF_JOIN_HEAD;      // no junction before the first exon; its list is kept as before
          if (si + 1 == tm.seg_hi) sg.flags |= MPH_SF_JOIN_TAIL;  // nor after the last one
          if (si == tm.seg_lo) continue;
          MphSegment& pa = b_.segs[si - 1];
          const bool always = tx_replay || ((sg.flags | pa.flags) & (MPH_SF_HAS_FS | MPH_SF_SHORT)) || sg.n_win == 0 || pa.n_win == 0;
          bool active = always || window_has_var(sg, 0) || window_has_var(pa, pa.n_win - 1);
          if (!active && (pa.flags & MPH_SF_KEEP_PENULT) && pa.n_win >= 2) active = window_has_var(pa, pa.n_win - 2);
          if (active) { sg.flags |= MPH_SF_JOIN_HEAD; pa.flags |= MPH_SF_JOIN_TAIL; }
        }
      }
      // device class (core/record_core.h): the records of this transcript are built by the record kernels
      {
        static const bool no_devrec = getenv("MPH_NO_DEVREC") != nullptr;  // test hook: everything through the host residue
        bool dev = !no_devrec && !tx_replay && !has_fs && need <= 240 && wl <= 32 && tm.seg_hi > tm.seg_lo &&
                   size_t(tm.seg_hi - tm.seg_lo) == size_t(exon_count);  // no exon was skipped (:1043-1048)
        for (uint32_t si = tm.seg_lo; dev && si < tm.seg_hi; ++si) {
          const MphSegment& sg = b_.segs[si];
          dev = !(sg.flags & MPH_SF_SHORT) && sg.n_win >= 2 && sg.k_first == 0 && sg.k_stride == 3;
        }
        if (dev)
          for (uint32_t si = tm.seg_lo; si < tm.seg_hi; ++si) b_.segs[si].flags |= MPH_SF_DEVREC;
      }
      if (tx_replay && tm.seg_hi > tm.seg_lo) {
        // split the transcript into units at the exon boundaries no observation crosses; the matrix columns at a
        // unit start come from a read-free pass over the window loop's column bookkeeping (:1119-1178,1280-1296)
        std::vector<uint32_t> dq;
        uint64_t last_vars = 0;
        bool broken = false;  // the reference panics in this transcript (drain out of range): keep the rest in one unit
        MphReplayTx rt;
        memset(&rt, 0, sizeof rt);
        auto open_unit = [&](uint32_t si) {
          rt.seg_lo = si; rt.read_lo = gm.read_lo; rt.read_hi = gm.read_hi;
          // normal mode keeps every copy of a re-offered read; copies with equal haplotypes share an entry
          rt.obs_off = uint32_t(b_.replay_obs); rt.obs_cap = b_.mode == 1 ? 4 * (gm.read_hi - gm.read_lo) + 1024 : gm.read_hi - gm.read_lo;
          rt.sl_va = b_.segs[tm.seg_lo].sl_va; rt.sl_vb = b_.segs[tm.seg_lo].sl_vb;
          if (!(b_.segs[tm.seg_lo].flags & MPH_SF_FIRST_EXON)) rt.sl_va = rt.sl_vb = 0;
          rt.dq_off = uint32_t(b_.replay_dq.size()); rt.dq_n = uint32_t(dq.size()); rt.last_vars = uint32_t(last_vars);
          b_.replay_dq.insert(b_.replay_dq.end(), dq.begin(), dq.end());
          b_.replay_obs += rt.obs_cap;
          if (b_.replay_obs > 0xFFFFFF00ull) throw Unsupported("batch too large: split it into gene ranges");
        };
        auto close_unit = [&](uint32_t si_end) {
          rt.seg_hi = si_end;
          b_.replay.push_back(rt);
        };
        open_unit(tm.seg_lo);
        for (uint32_t si = tm.seg_lo; si < tm.seg_hi; ++si) {
          const MphSegment& sg = b_.segs[si];
          if (si > tm.seg_lo && !broken && !carries_over(b_.segs[si - 1], sg)) {
            close_unit(si);
            open_unit(si);
          }
          if (broken) continue;
          const bool rev = (sg.flags & MPH_SF_REVERSE) != 0, is_short = (sg.flags & MPH_SF_SHORT) != 0;
          auto shrink = [&](uint64_t n) {
            if (n > dq.size()) { broken = true; return; }
            dq.erase(dq.begin(), dq.begin() + long(n));
          };
          shrink(last_vars);
          last_vars = 0;
          uint64_t old_offset = sg.off0, old_end = uint64_t(sg.off0) + sg.ewl;
          bool reached_end = false;
          for (uint32_t k = 0; k < sg.n_iter && !broken; ++k) {
            const uint64_t offset = rev ? uint64_t(sg.off0) - k : uint64_t(sg.off0) + k;
            const MphGeom g = mph_geom(sg, k);
            const uint64_t rest = rev ? offset - sg.exon_start : sg.exon_end - (offset + sg.ewl);
            auto cnt = [&](uint64_t a, uint64_t c2) -> uint64_t {
              if (a > c2) { broken = true; return 0; }
              const uint32_t ia = mph_var_lb(b_.vars.data(), sg.var_lo, sg.var_hi, uint32_t(a));
              return mph_var_lb(b_.vars.data(), ia, sg.var_hi, uint32_t(c2)) - ia;
            };
            const uint32_t va = mph_var_lb(b_.vars.data(), sg.var_lo, sg.var_hi, g.s);
            const uint32_t vb = mph_var_lb(b_.vars.data(), va, sg.var_hi, g.e);
            const uint64_t nvars = vb - va;
            uint64_t added, deleted;
            if (k == 0) added = nvars;
            else if (is_short || reached_end) added = 0;
            else if (g.s > old_offset) added = cnt(old_end, g.e);
            else added = cnt(g.s, old_offset);
            if (offset == old_offset || is_short) deleted = 0;
            else if (g.s > old_offset) deleted = cnt(old_offset, g.s);
            else deleted = cnt(g.e, old_end);
            if (rest < 3) reached_end = true;
            if (broken) break;
            shrink(deleted);
            const uint64_t skip = nvars - added;
            if (skip <= nvars)
              for (uint64_t x = skip; x < nvars; ++x) dq.push_back(rev ? vb - 1 - uint32_t(x) : va + uint32_t(x));
            last_vars = nvars;
            old_offset = g.s;
            old_end = g.e;
            if (is_short) break;
          }
        }
        close_unit(tm.seg_hi);
        for (uint32_t si = tm.seg_lo; si < tm.seg_hi; ++si) b_.segs[si].flags |= MPH_SF_REPLAY;
        gene_replay = true;
      }
      b_.tx_id_bytes.insert(b_.tx_id_bytes.end(), tm.id.begin(), tm.id.end());
      b_.tx_id_off.push_back(uint32_t(b_.tx_id_bytes.size()));
      b_.txs.push_back(std::move(tm));
    }
    if (gene_replay) {
      // the replay evaluates matrix columns outside a read's own variant range: every read of the gene ships its bases and CIGAR
      b_.bases.resize(bases_mark);
      b_.cigars.resize(cigars_mark);
      b_.vr_read.resize(vr_mark); b_.vr_vlo.resize(vr_mark); b_.vr_seq_off.resize(vr_mark); b_.vr_cig_off.resize(vr_mark);
      b_.vr_lseq.resize(vr_mark); b_.vr_ncig.resize(vr_mark); b_.vr_nv.resize(vr_mark);
      uint32_t idx = gm.read_lo;
      size_t q = 0;
      full_records = true;
      for (auto& r : reads) {
        add_vr(idx, read_vrange[q].first, read_vrange[q].second, r);
        ++idx;
        ++q;
      }
    }
    // bus encoding of this gene's side-table entries (see Batch::vs_read_d)
    {
      uint32_t cig_sum = uint32_t(cigars_mark);
      bool after_exception = false;  // the device sums the 8-bit op counts: an entry with 255+ ops is followed by a fresh run
      for (size_t e = vr_mark; e < b_.vr_read.size(); ++e) {
        const uint32_t size = uint32_t((e + 1 < b_.vr_read.size() ? b_.vr_seq_off[e + 1] : b_.bases.size()) - b_.vr_seq_off[e]);
        if (size > 0xFFFFu) throw Unsupported("read record larger than 64 KB");
        const bool fresh = e == vr_mark || after_exception || b_.vr_read[e] - b_.vr_read[e - 1] > 0xFFFFu || b_.vr_vlo[e] < b_.vr_vlo[e - 1] ||
                           b_.vr_vlo[e] - b_.vr_vlo[e - 1] > 255u;
        if (fresh) b_.vs_runs.push_back(Batch::VRun{uint32_t(e), b_.vr_read[e], b_.vr_vlo[e], b_.vr_seq_off[e], cig_sum});
        b_.vs_read_d.push_back(uint16_t(fresh ? 0u : b_.vr_read[e] - b_.vr_read[e - 1]));
        b_.vs_vlo_d.push_back(uint8_t(fresh ? 0u : b_.vr_vlo[e] - b_.vr_vlo[e - 1]));
        b_.vs_size.push_back(uint16_t(size));
        const uint32_t nc = b_.vr_ncig[e];
        b_.vs_ncig.push_back(uint8_t(nc < 255u ? nc : 255u));
        if (nc >= 255u) b_.vs_ncig_exc.push_back(Batch::U2{uint32_t(e), nc});
        after_exception = nc >= 255u;
        b_.vr_cig_off[e] = cig_sum;  // running sum (single-M entries ship no CIGAR and never read theirs)
        cig_sum += nc;
      }
    }
    // bus encoding of this gene's reads (see Batch::rd_delta)
    if (!b_.modal_set && gm.read_hi - gm.read_lo >= 16) {  // the batch's span: the commonest one of this gene
      std::vector<uint32_t> spans;
      for (uint32_t r = gm.read_lo; r < gm.read_hi; ++r) spans.push_back(b_.read_end[r] - b_.read_start[r]);
      std::sort(spans.begin(), spans.end());
      size_t best = 0;
      for (size_t i = 0, j = 0; i < spans.size(); i = j) {
        while (j < spans.size() && spans[j] == spans[i]) ++j;
        if (j - i > best) { best = j - i; b_.modal_span = spans[i]; }
      }
      b_.modal_set = true;
    }
    for (uint32_t r = gm.read_lo; r < gm.read_hi; ++r) {
      uint32_t delta = 0;
      if (r == gm.read_lo || b_.read_start[r] - b_.read_start[r - 1] > 255u) b_.rd_runs.push_back(Batch::U2{r, b_.read_start[r]});
      else delta = b_.read_start[r] - b_.read_start[r - 1];
      b_.rd_delta.push_back(uint8_t(delta));
      const uint32_t span = b_.read_end[r] - b_.read_start[r];
      b_.rd_span.push_back(uint8_t(span < 255u ? span : 255u));
      if (span >= 255u) b_.rd_span_exc.push_back(Batch::U2{r, b_.read_end[r]});
      if (!b_.modal_set || span != b_.modal_span) b_.rd_mspan_exc.push_back(Batch::U2{r, b_.read_end[r]});
      if (b_.read_flags[r]) b_.rd_flag_exc.push_back(Batch::U2{r, b_.read_flags[r]});
    }
    // work items of the read-run kernel: (segment, read) pairs; replayed transcripts take none (k_replay does their windows)
    for (size_t si = b_.seg_work_off.size() - 1; si < b_.segs.size(); ++si) {
      if (b_.segs[si].flags & MPH_SF_REPLAY) b_.seg_work[si].rhi = b_.seg_work[si].rlo;
      const uint64_t next = uint64_t(b_.seg_work_off.back()) + (b_.seg_work[si].rhi - b_.seg_work[si].rlo);
      if (next > 0xFFFFFF00ull) throw Unsupported("batch too large: split it into gene ranges");
      b_.seg_work_off.push_back(uint32_t(next));
    }
    b_.genes.push_back(std::move(gm));
    GeneMark mk;
    mk.reads = b_.read_start.size(); mk.vr = b_.vr_read.size(); mk.bases = b_.bases.size(); mk.cigars = b_.cigars.size(); mk.vars = b_.vars.size(); mk.ins = b_.ins_bytes.size();
    mk.segs = b_.segs.size(); mk.chunks = b_.chunks.size(); mk.ref = b_.ref.size(); mk.windows = b_.n_windows; mk.txs = b_.txs.size();
    mk.replay = b_.replay.size(); mk.dq = b_.replay_dq.size(); mk.partners = b_.partner_a.size();
    mk.runs = b_.rd_runs.size(); mk.span_exc = b_.rd_span_exc.size(); mk.flag_exc = b_.rd_flag_exc.size(); mk.mspan_exc = b_.rd_mspan_exc.size();
    mk.vruns = b_.vs_runs.size(); mk.ncig_exc = b_.vs_ncig_exc.size();
    b_.marks.push_back(mk);
  }

 public:
  Batch& batch() { return b_; }

 private:
  // An observation survives into the next exon when it still passes cleanup_reads there
  // (:259-278); exome introns are longer than a read, so this is rare, and the closed form of
  // phase_core.h does not model it.
  bool carries_over(const MphSegment& a, const MphSegment& bseg) const {
    if (a.n_iter == 0) return false;
    const MphGeom ga = mph_geom(a, a.n_iter - 1), gb = mph_geom(bseg, 0);
    const bool rev = (a.flags & MPH_SF_REVERSE) != 0;
    // a carried read encloses the last window of A and still passes B's first cleanup:
    //   forward: start <= s_A and end >= e_B ; reverse: start <= s_B and end >= e_A
    const uint32_t need_end = rev ? ga.e : gb.e;
    const uint32_t max_start = rev ? std::min(ga.s, gb.s) : ga.s;
    const uint32_t min_start = need_end > a.max_span ? need_end - a.max_span : 0;
    if (min_start > max_start) return false;
    auto lo = std::lower_bound(b_.read_start.begin() + a.read_lo, b_.read_start.begin() + a.read_hi, min_start) - b_.read_start.begin();
    for (uint32_t r = uint32_t(lo); r < a.read_hi && b_.read_start[r] <= max_start; ++r)
      if (b_.read_end[r] >= need_end && b_.read_end[r] >= ga.e) return true;
    return false;
  }

  Batch b_;
  uint32_t chunk_windows_;
};

}  // namespace mph
