// oracle_normal.hpp — TEST INFRASTRUCTURE ONLY.
//
// Statement-by-statement CPU restatement of the reference's `normal` (healthy peptidome) phasing path,
// reference src/normal_microphasing.rs (all line numbers below refer to that file):
//   supports_variant :43-78, IDRecord::{update,add_freq} :104-180, Observation :188-216,
//   ObservationMatrix :218-648, phase_gene :650-1279, phase :1281-1440.
// The differences from the somatic path that change results are reproduced
// literally: no mapq / base-quality filters, no `contains`, oldest-first bit order in push_read vs
// newest-first in extend_right, cleanup_reads(splice_side_offset) on the reverse strand (reads are
// inserted repeatedly), VecMap histogram, freq = count / nrows (NaN without reads), the `j`-indexed
// bit test and same-position skip in the sequence walk, the extra reference base after each variant
// block, first/last-codon stop test, every haplotype of every non-short window written.
//
// Pinned against the reference's expected FASTA of test_forward_germline and
// splice_test_forward_germline (tests/golden/*_normal). The TSV is written by the reference's tests
// but never diffed, and there is no live reverse-strand fixture: TSV contents and reverse-strand
// `normal` behaviour are parity unpinned.
#pragma once
#include <cmath>
#include <limits>

#include "oracle_common.hpp"

namespace oracle {
namespace normal {

using mphio::BamRecord;

inline bool is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }
inline uint8_t lower(uint8_t c) { return is_upper(c) ? uint8_t(c + 32) : c; }
inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? uint8_t(c - 32) : c; }
inline uint8_t switch_ascii_case(uint8_t c, uint8_t r) { return is_upper(r) ? lower(c) : c; }

// :43-78
inline bool supports_variant(const BamRecord& read, const Variant& variant) {
  switch (variant.kind) {
    case Variant::SNV: {
      uint32_t p;
      if (mphio::cigar_read_pos(read.cigar, read.pos, int64_t(uint32_t(variant.pos)), &p) != 1) return false;
      if (p >= read.l_seq) throw Panic("index out of bounds: read.seq()[p]");
      return read.base(p) == variant.alt;
    }
    case Variant::Insertion:
      for (uint32_t c : read.cigar)
        if ((c & 15) == mphio::C_I && (c >> 4) == uint32_t(variant.len)) return true;
      return false;
    default:
      for (uint32_t c : read.cigar)
        if ((c & 15) == mphio::C_D && (c >> 4) == uint32_t(variant.len)) return true;
      return false;
  }
}

// :80-180
struct NRecord {
  std::string id, transcript, gene_id, gene_name, chrom;
  uint64_t offset = 0, frame = 0;
  double freq = 0;
  uint32_t depth = 0, nvar = 0, nsomatic = 0, nvariant_sites = 0, nsomvariant_sites = 0;
  std::string strand, variant_sites, somatic_positions, somatic_aa_change, germline_positions, germline_aa_change, peptide_sequence;

  static const std::vector<std::string>& header() {
    static const std::vector<std::string> h = {
        "id", "transcript", "gene_id", "gene_name", "chrom", "offset", "frame", "freq", "depth", "nvar", "nsomatic", "nvariant_sites",
        "nsomvariant_sites", "strand", "variant_sites", "somatic_positions", "somatic_aa_change", "germline_positions",
        "germline_aa_change", "peptide_sequence"};
    return h;
  }
  std::vector<std::string> fields() const {
    return {id, transcript, gene_id, gene_name, chrom, std::to_string(offset), std::to_string(frame), mphfmt::format_f64(freq),
            std::to_string(depth), std::to_string(nvar), std::to_string(nsomatic), std::to_string(nvariant_sites),
            std::to_string(nsomvariant_sites), strand, variant_sites, somatic_positions, somatic_aa_change, germline_positions,
            germline_aa_change, peptide_sequence};
  }
  // :105-146
  NRecord update(const NRecord& rec, uint64_t off, const Bytes& seq) const {
    NRecord o;
    o.id = mphfmt::record_id(seq.data(), seq.size(), transcript, off, strand.empty() ? '?' : strand[0]);
    o.transcript = transcript; o.gene_id = gene_id; o.gene_name = gene_name; o.chrom = chrom;
    o.offset = off + offset;
    o.frame = frame;
    o.freq = freq * rec.freq;
    o.depth = depth;
    o.nvar = nvar + rec.nvar;
    o.nsomatic = nsomatic + rec.nsomatic;
    o.nvariant_sites = nvariant_sites + rec.nvariant_sites;
    o.nsomvariant_sites = nsomvariant_sites + rec.nsomvariant_sites;
    o.strand = strand;
    o.variant_sites = variant_sites + rec.variant_sites;
    o.somatic_positions = somatic_positions + rec.somatic_positions;
    o.somatic_aa_change = somatic_aa_change + rec.somatic_aa_change;
    o.germline_positions = germline_positions + rec.germline_positions;
    o.germline_aa_change = germline_aa_change + rec.germline_aa_change;
    o.peptide_sequence.assign(seq.begin(), seq.end());
    return o;
  }
  // :148-179
  NRecord add_freq(double f) const {
    NRecord o = *this;
    const uint32_t new_nvar = f > 0.0 ? nvar - 1 : nvar;  // u32 subtraction wraps in the release build
    o.nsomatic = new_nvar < nsomatic ? nsomatic - 1 : nsomatic;
    o.nvar = new_nvar;
    o.freq = freq + f;
    return o;
  }
};

struct HaplotypeSeq {
  Bytes sequence;
  NRecord record;
};

struct Observation {
  ReadPtr read;
  uint64_t haplotype = 0;
  // :195-215
  void update_haplotype(size_t i, const Variant& variant) {
    if (uint64_t(int64_t(read->pos)) > variant.pos) throw Panic("bug: read starts right of variant");
    if (supports_variant(*read, variant)) haplotype |= uint64_t(1) << (i & 63);
  }
};

struct Writers {
  FastaWriter fasta;  // stdout
  TsvWriter tsv;      // --tsv
};

struct Stats {
  uint64_t windows = 0, read_windows = 0;
};
inline Stats& stats() {
  static Stats s;
  return s;
}

struct ObservationMatrix {
  std::map<uint64_t, std::vector<Observation>> observations;
  std::deque<Variant> variants;

  uint32_t ncols() const { return uint32_t(variants.size()); }
  size_t nrows() const {
    size_t n = 0;
    for (auto& kv : observations) n += kv.second.size();
    return n;
  }
  // :238-247
  void shrink_left(size_t k) {
    if (k > variants.size()) throw Panic("drain: range end out of bounds");
    variants.erase(variants.begin(), variants.begin() + long(k));
    const uint32_t nc = ncols();
    const uint64_t mask = nc >= 64 ? ~uint64_t(0) : ((uint64_t(1) << nc) - 1);
    for (auto& kv : observations)
      for (auto& obs : kv.second) obs.haplotype &= mask;
  }
  // :250-267
  void extend_right(const std::vector<Variant>& new_variants) {
    const size_t k = new_variants.size();
    if (k > 0)
      for (auto& kv : observations)
        for (auto& obs : kv.second) obs.haplotype <<= (k & 63);
    for (auto& kv : observations)
      for (auto& obs : kv.second) {
        size_t i = 0;
        for (auto it = new_variants.rbegin(); it != new_variants.rend(); ++it, ++i) obs.update_haplotype(i, *it);
      }
    for (auto& v : new_variants) variants.push_back(v);
  }
  // :270-283
  void cleanup_reads(uint64_t interval_end, bool reverse) {
    auto it = observations.lower_bound(interval_end);
    if (!reverse) observations.erase(observations.begin(), it);
    else observations.erase(it, observations.end());
  }
  // :301-331 — no `contains`, oldest variant = bit 0
  void push_read(const ReadPtr& read, uint64_t interval_end, uint64_t interval_start, bool reverse) {
    const uint64_t end_pos = uint64_t(read->end_pos()), start_pos = uint64_t(int64_t(read->pos));
    if (end_pos >= interval_end && start_pos <= interval_start) {
      Observation obs;
      obs.read = read;
      size_t i = 0;
      for (auto it = variants.begin(); it != variants.end(); ++it, ++i) obs.update_haplotype(i, *it);
      observations[reverse ? start_pos : end_pos].push_back(std::move(obs));
    }
  }

  static uint8_t ref_at(const Bytes& refseq, uint64_t idx) {
    if (idx >= refseq.size()) throw Panic("index out of bounds: refseq");
    return refseq[size_t(idx)];
  }

  // :341-647
  std::vector<HaplotypeSeq> print_haplotypes(const Gene& gene, const Transcript& transcript, uint64_t offset, uint64_t splice_end,
                                             uint64_t splice_pos, uint64_t splice_gap, uint64_t window_len, const Bytes& refseq, Writers& w,
                                             bool is_short_exon, uint64_t frame) const {
    std::vector<const Variant*> variants_v;
    for (auto& v : variants) variants_v.push_back(&v);
    const bool rev = transcript.strand == Strand::Reverse;
    if (rev) std::reverse(variants_v.begin(), variants_v.end());
    // VecMap<usize>: iteration in ascending key order
    std::map<uint64_t, uint64_t> haplotypes;
    for (auto& kv : observations)
      for (auto& obs : kv.second) haplotypes[obs.haplotype] += 1;
    const char* strand = rev ? "Reverse" : "Forward";
    std::vector<HaplotypeSeq> haplotypes_vec;
    if (haplotypes.empty()) haplotypes[0] = 0;
    const uint64_t gs = gene.start();
    Bytes seq;
    for (auto& hk : haplotypes) {
      const uint64_t haplotype = hk.first, count = hk.second;
      seq.clear();
      bool insertion = false;
      uint32_t n_somatic = 0, n_variants = 0;
      const double freq = double(count) / double(nrows());
      const uint32_t depth = uint32_t(nrows());
      uint64_t i = offset;
      size_t j = 0;
      uint64_t window_end = splice_end;
      std::vector<int> variant_profile;
      if (variants_v.empty()) {
        if (offset - gs > window_end - gs || window_end - gs > refseq.size()) throw Panic("slice index out of range: refseq");
        seq.insert(seq.end(), refseq.begin() + long(offset - gs), refseq.begin() + long(window_end - gs));
      } else {
        while (i < window_end) {
          while (j < variants_v.size() && i == variants_v[j]->pos) {
            if (std::fabs(freq - 1.0) < std::numeric_limits<double>::epsilon() && !variants_v[j]->is_germline()) {
              j += 1;
              variant_profile.push_back(0);
              continue;
            }
            if ((haplotype & (uint64_t(1) << (j & 63))) != 0) {
              if (j + 1 < variants_v.size() && i == variants_v[j + 1]->pos) j += 1;
              const Variant& vj = *variants_v[j];
              if (vj.kind == Variant::SNV) {
                seq.push_back(switch_ascii_case(vj.alt, ref_at(refseq, i - gs)));
                i += 1;
              } else if (vj.kind == Variant::Insertion) {
                const uint8_t r = ref_at(refseq, i - gs);
                for (uint8_t c : vj.seq) seq.push_back(is_upper(r) ? lower(c) : upper(c));
                insertion = true;
                i += 1;
              } else {
                seq.push_back(ref_at(refseq, i - gs));
                i += vj.len + 1;
                window_end += vj.len + 1;
              }
              if (!vj.is_germline()) { n_somatic += 1; variant_profile.push_back(2); }
              else variant_profile.push_back(1);
              n_variants += 1;
            } else {
              variant_profile.push_back(0);
            }
            j += 1;
          }
          seq.push_back(ref_at(refseq, i - gs));  // :476 — unconditional
          i += 1;
        }
      }
      const uint64_t this_window_len = seq.size() < window_len ? uint64_t(seq.size()) : window_len;
      auto slice = [&](uint64_t a, uint64_t e) -> std::string {
        if (a > e || e > seq.size()) throw Panic("slice index out of range");
        return std::string(seq.begin() + long(a), seq.begin() + long(e));
      };
      std::string peptide;
      if (splice_pos == 1) peptide = slice(splice_gap, seq.size());
      else if (splice_pos == 0) peptide = insertion ? slice(0, seq.size()) : slice(0, this_window_len);
      else peptide = slice(0, seq.size());
      auto starts = [&](const char* c) { return peptide.size() >= 3 && peptide.compare(0, 3, c) == 0; };
      auto ends = [&](const char* c) { return peptide.size() >= 3 && peptide.compare(peptide.size() - 3, 3, c) == 0; };
      const bool stop_gain = rev ? (ends("TCA") || ends("CTA") || ends("TTA")) : (starts("TGA") || starts("TAG") || starts("TAA"));
      if (stop_gain && splice_pos != 2) continue;
      const std::string fasta_id = mphfmt::record_id(seq.data(), seq.size(), transcript.id, offset, strand[0]);
      uint32_t n_variantsites = 0, n_som_variantsites = 0;
      std::vector<std::string> s_pc, g_pc, s_pos, g_pos, sites;
      for (size_t c = 0; c < variants_v.size(); ++c) {
        if (c < variant_profile.size()) {
          if (variant_profile[c] == 2) { s_pos.push_back(std::to_string(variants_v[c]->pos)); s_pc.push_back(variants_v[c]->prot_change()); }
          else if (variant_profile[c] == 1) { g_pos.push_back(std::to_string(variants_v[c]->pos)); g_pc.push_back(variants_v[c]->prot_change()); }
          if (c == 0 || variants_v[c]->pos != variants_v[c - 1]->pos) {
            n_variantsites += 1;
            sites.push_back(std::to_string(variants_v[c]->pos));
            if (!variants_v[c]->is_germline()) n_som_variantsites += 1;
          }
        }
      }
      NRecord record;
      record.id = fasta_id; record.transcript = transcript.id; record.gene_id = gene.id; record.gene_name = gene.name; record.chrom = gene.chrom;
      record.offset = offset; record.frame = frame; record.freq = freq; record.depth = depth; record.nvar = n_variants;
      record.nsomatic = n_somatic; record.nvariant_sites = n_variantsites; record.nsomvariant_sites = n_som_variantsites;
      record.strand = strand;
      record.variant_sites = IDRecord::join(sites);
      record.somatic_positions = IDRecord::join(s_pos);
      record.somatic_aa_change = IDRecord::join(s_pc);
      record.germline_positions = IDRecord::join(g_pos);
      record.germline_aa_change = IDRecord::join(g_pc);
      record.peptide_sequence = peptide;
      HaplotypeSeq hs;
      hs.sequence = seq;
      hs.record = record;
      hs.record.peptide_sequence.assign(seq.begin(), seq.end());
      haplotypes_vec.push_back(std::move(hs));
      if (!is_short_exon) {
        if (splice_pos == 1) {
          if (splice_gap > seq.size()) throw Panic("slice index out of range");
          w.fasta.write(record.id, seq.data() + splice_gap, seq.size() - size_t(splice_gap));
        } else if (splice_pos == 0) {
          if (window_len > seq.size()) throw Panic("slice index out of range");
          w.fasta.write(record.id, seq.data(), size_t(window_len));
        }
        w.tsv.row(NRecord::header(), record.fields());
      }
    }
    return haplotypes_vec;
  }
};

template <class Map>
inline size_t count_range(const Map& m, uint64_t a, uint64_t b) {
  if (a > b) throw Panic("range start is greater than range end in BTreeMap");
  size_t n = 0;
  for (auto it = m.lower_bound(a); it != m.end() && it->first < b; ++it) n += it->second.size();
  return n;
}

// :650-1279
inline void phase_gene(const Gene& gene, BamRecordBuffer& read_buffer, VcfRecordBuffer& variant_buffer, mphio::FastaIndexed& fasta, Writers& w,
                       uint64_t window_len, Bytes& refseq, bool unsupported_allele_warning_only) {
  fasta.fetch(gene.chrom, gene.start(), gene.end() + 100, refseq);
  std::map<uint64_t, std::vector<Variant>> variant_tree;
  std::map<uint64_t, std::vector<ReadPtr>> read_tree;
  read_buffer.fetch(gene.chrom, gene.start(), gene.end());
  uint64_t max_read_len = 0;
  for (auto& rec : read_buffer.inner) {
    if (uint64_t(rec->l_seq) > max_read_len) max_read_len = rec->l_seq;
    read_tree[uint64_t(int64_t(rec->pos))].push_back(rec);
  }
  variant_buffer.fetch(gene.chrom, gene.start(), gene.end());
  for (auto& rec : variant_buffer.ring) variant_tree[uint64_t(rec.pos)] = variants_from_record(*variant_buffer.vcf, rec, unsupported_allele_warning_only);

  for (const Transcript& transcript : gene.transcripts) {
    if (!transcript.is_coding()) continue;
    const bool fwd = transcript.strand == Strand::Forward;
    const size_t exon_number = transcript.exons.size();
    ObservationMatrix observations;
    std::map<uint64_t, uint64_t> frameshifts;
    if (fwd) frameshifts[0] = 0;
    else frameshifts[gene.end()] = 0;
    uint64_t exon_rest = 0;
    std::vector<HaplotypeSeq> prev_hap_vec, hap_vec;
    size_t last_window_vars = 0;
    for (size_t exon_count = 0; exon_count < transcript.exons.size(); ++exon_count) {
      const Interval& exon = transcript.exons[exon_count];
      if (frameshifts.empty()) break;
      if (exon.start > exon.end) continue;
      const bool is_last_exon = exon_count == exon_number - 1;
      const bool is_first_exon = exon_count == 0;
      const uint64_t exon_len = exon.end - exon.start;
      const uint64_t current_exon_offset = exon_rest == 0 ? 0 : 3 - exon_rest;
      const bool is_short_exon = exon_len < 3 ? true : window_len >= exon_len - current_exon_offset - (3 - current_exon_offset) % 3;
      uint64_t exon_window_len = !is_short_exon ? window_len : (exon_len - current_exon_offset) - ((exon_len - current_exon_offset) % 3);
      if (exon_window_len == 0) exon_window_len = exon_len;
      exon_rest = 0;
      uint64_t offset = !fwd ? exon.end - exon_window_len - current_exon_offset : exon.start + current_exon_offset;
      bool reached_end = false;
      uint64_t old_offset = offset;
      uint64_t old_end = old_offset + exon_window_len;
      observations.shrink_left(last_window_vars);
      last_window_vars = 0;
      bool is_first_exon_window = true;
      for (;;) {
        if (frameshifts.empty()) break;
        const bool valid = !fwd ? offset >= exon.start : offset + exon_window_len <= exon.end;
        if (!valid) break;
        if (max_read_len < exon_window_len) break;
        const uint64_t rest = fwd ? exon.end - (offset + exon_window_len) : offset - exon.start;
        const bool is_last_exon_window = rest < 3;
        uint64_t splice_side_offset, splice_end, splice_gap, splice_pos;
        if (fwd) {
          if (is_short_exon || (is_first_exon_window && is_last_exon_window)) {
            splice_side_offset = offset - current_exon_offset; splice_end = offset + exon_window_len + rest;
            splice_gap = current_exon_offset + rest; splice_pos = 2;
          } else if (is_first_exon_window) {
            splice_side_offset = offset - current_exon_offset; splice_end = offset + exon_window_len; splice_gap = current_exon_offset; splice_pos = 1;
          } else if (is_last_exon_window) {
            splice_side_offset = offset; splice_end = offset + exon_window_len + rest; splice_gap = rest; splice_pos = 0;
          } else {
            splice_side_offset = offset; splice_end = offset + exon_window_len; splice_gap = 0; splice_pos = 0;
          }
        } else {
          if (is_short_exon) {
            splice_side_offset = offset - rest; splice_end = offset + exon_window_len + current_exon_offset;
            splice_gap = current_exon_offset + rest; splice_pos = 2;
          } else if (is_first_exon_window) {
            splice_side_offset = offset; splice_end = offset + exon_window_len + current_exon_offset; splice_gap = current_exon_offset; splice_pos = 0;
          } else if (is_last_exon_window) {
            splice_side_offset = offset - rest; splice_end = offset + exon_window_len; splice_gap = rest; splice_pos = 1;
          } else {
            splice_side_offset = offset; splice_end = offset + exon_window_len; splice_gap = 0; splice_pos = 0;
          }
        }
        const size_t nvars = count_range(variant_tree, splice_side_offset, splice_end);
        last_window_vars = nvars;
        size_t added_vars;
        if (is_first_exon_window) added_vars = nvars;
        else if (is_short_exon) added_vars = 0;
        else if (reached_end) added_vars = 0;
        else if (splice_side_offset > old_offset) added_vars = count_range(variant_tree, old_end, splice_end);
        else added_vars = count_range(variant_tree, splice_side_offset, old_offset);
        size_t deleted_vars;
        if (offset == old_offset || is_short_exon) deleted_vars = 0;
        else if (splice_side_offset > old_offset) deleted_vars = count_range(variant_tree, old_offset, splice_side_offset);
        else deleted_vars = count_range(variant_tree, splice_end, old_end);
        if (is_last_exon_window) reached_end = true;
        std::vector<ReadPtr> reads;
        {
          uint64_t lo, hi;
          const bool wide = !fwd || offset == exon.start + current_exon_offset;
          if (wide) { lo = splice_side_offset - (max_read_len - exon_window_len); hi = splice_side_offset + 1; }
          else { lo = splice_side_offset; hi = splice_side_offset + 1; }
          if (lo > hi) throw Panic("range start is greater than range end in BTreeMap");
          for (auto it = read_tree.lower_bound(lo); it != read_tree.end() && it->first < hi; ++it)
            for (auto& r : it->second) reads.push_back(r);
        }
        const bool reverse = !fwd;
        if (reverse) observations.cleanup_reads(splice_side_offset, reverse);  // :1001 — no "+ 1"
        else observations.cleanup_reads(splice_end, reverse);
        observations.shrink_left(deleted_vars);
        for (auto& read : reads) observations.push_read(read, splice_end, splice_side_offset, reverse);
        std::vector<Variant> variants;
        {
          if (splice_side_offset > splice_end) throw Panic("range start is greater than range end in BTreeMap");
          std::vector<const std::vector<Variant>*> groups;
          for (auto it = variant_tree.lower_bound(splice_side_offset); it != variant_tree.end() && it->first < splice_end; ++it) groups.push_back(&it->second);
          if (reverse) std::reverse(groups.begin(), groups.end());
          const size_t skip = nvars - added_vars;
          size_t idx = 0;
          for (auto g : groups)
            for (auto& v : *g)
              if (idx++ >= skip) variants.push_back(v);
        }
        for (const Variant& variant : variants) {  // :1039-1049 — no "% 3", no strand split
          const uint64_t s = variant.frameshift();
          if (s > 0) {
            std::vector<uint64_t> previous;
            for (auto& kv : frameshifts) previous.push_back(kv.second + s);
            for (uint64_t s_ : previous) frameshifts[variant.end_pos()] = s_;
          }
        }
        observations.extend_right(variants);
        uint64_t stopped_frameshift = 3;
        std::vector<std::pair<uint64_t, uint64_t>> active;
        if (fwd) { for (auto it = frameshifts.begin(); it != frameshifts.end() && it->first < offset; ++it) active.push_back(*it); }
        else { for (auto it = frameshifts.lower_bound(offset + exon_window_len); it != frameshifts.end(); ++it) active.push_back(*it); }
        uint64_t frameshift_count = 0;
        bool main_orf = false;
        for (auto& kf : active) {
          const uint64_t key = kf.first, frameshift = kf.second;
          if (frameshift == 0) main_orf = true;
          frameshift_count += 1;
          const uint64_t coding_shift = fwd ? offset - exon.start : exon.end - offset;
          const bool has_frameshift = frameshift > 0;
          if (coding_shift % 3 == (frameshift + current_exon_offset) % 3 || is_short_exon) {
            if (!has_frameshift) {
              exon_rest = fwd ? exon.end - (offset + exon_window_len) : offset - exon.start;
              if (exon_window_len < 3) exon_rest = exon_window_len;
            }
            if (frameshift == 0) {
              stats().windows += 1;
              stats().read_windows += observations.nrows();
            }
            auto res = observations.print_haplotypes(gene, transcript, splice_side_offset, splice_end, splice_pos, splice_gap, exon_window_len,
                                                     refseq, w, is_short_exon, frameshift);
            if (res.empty()) stopped_frameshift = key;
            if (exon_rest < 3 && (!is_short_exon || is_first_exon)) prev_hap_vec = std::move(res);
            else hap_vec = std::move(res);
          }
        }
        if (frameshift_count == 0 || !main_orf) {
          frameshifts.clear();
          break;
        }
        frameshifts.erase(stopped_frameshift);  // :1130 — unconditional
        if (frameshifts.empty()) break;
        const bool at_splice_side = fwd ? offset - current_exon_offset == exon.start : offset + exon_window_len + current_exon_offset == exon.end;
        is_first_exon_window = false;
        if (at_splice_side && !is_first_exon) {  // :1145-1250
          const std::vector<HaplotypeSeq>& first_hap_vec = fwd ? hap_vec : prev_hap_vec;
          const std::vector<HaplotypeSeq>& sec_hap_vec = fwd ? prev_hap_vec : hap_vec;
          std::map<std::pair<uint64_t, Bytes>, std::pair<Bytes, NRecord>> output_map;
          std::vector<HaplotypeSeq> new_hap_vec;
          for (const HaplotypeSeq& hapseq : first_hap_vec) {
            const Bytes& sequence = hapseq.sequence;
            const NRecord& record = hapseq.record;
            for (const HaplotypeSeq& prev_hapseq : sec_hap_vec) {
              Bytes prev_sequence = prev_hapseq.sequence;
              const NRecord& prev_record = prev_hapseq.record;
              prev_sequence.insert(prev_sequence.end(), sequence.begin(), sequence.end());
              if (is_short_exon) {
                HaplotypeSeq nh;
                nh.sequence = prev_sequence;
                nh.record = prev_record.update(record, 0, prev_sequence);
                new_hap_vec.push_back(std::move(nh));
              }
              uint64_t splice_offset = 3;
              if (!fwd && exon_rest < 3) splice_offset += exon_rest;
              size_t end_offset = 3;
              if (is_last_exon_window) end_offset = 0;
              if (uint64_t(prev_sequence.size()) < 2 * window_len) {
                if (fwd) splice_offset = 0;
                else end_offset = 0;
              }
              while (splice_offset + window_len <= uint64_t(prev_sequence.size() - end_offset)) {
                if (splice_offset + window_len > prev_sequence.size()) throw Panic("slice index out of range");
                Bytes out_seq(prev_sequence.begin() + long(splice_offset), prev_sequence.begin() + long(splice_offset + window_len));
                NRecord out_record = prev_record.update(record, splice_offset, out_seq);
                auto id_tuple = std::make_pair(splice_offset, out_seq);
                auto it = output_map.find(id_tuple);
                const double old_freq = it != output_map.end() ? it->second.second.freq : 0.0;
                output_map[id_tuple] = std::make_pair(out_seq, out_record.add_freq(old_freq));
                splice_offset += 3;
              }
            }
          }
          if (is_short_exon && !is_last_exon) {
            prev_hap_vec = std::move(new_hap_vec);
          } else {
            for (auto& kv : output_map) {
              const NRecord& out_record = kv.second.second;
              const Bytes& out_seq = kv.second.first;
              if (window_len > out_seq.size()) throw Panic("slice index out of range");
              w.fasta.write(out_record.id, out_seq.data(), size_t(window_len));
              w.tsv.row(NRecord::header(), out_record.fields());
            }
          }
        }
        old_offset = splice_side_offset;
        old_end = splice_end;
        if (fwd) offset += 1;
        else offset -= 1;
        if (frameshifts.empty()) break;
        if (is_short_exon) break;  // :1266-1269 (inside `loop`)
      }
    }
  }
}

// :1281-1440 — like the somatic driver but without three_prime_utr handling
inline void phase(mphio::FastaIndexed& fasta, std::istream& gtf, mphio::VcfFile& vcf, mphio::BamFile& bam, Writers& w, uint64_t window_len,
                  bool unsupported_allele_warning_only) {
  BamRecordBuffer read_buffer;
  read_buffer.load(bam);
  VcfRecordBuffer variant_buffer;
  variant_buffer.vcf = &vcf;
  Bytes refseq;
  std::unique_ptr<Gene> gene;
  bool start_codon_found = false;
  auto phase_last_gene = [&](const Gene& g) {
    if (g.biotype == "protein_coding") phase_gene(g, read_buffer, variant_buffer, fasta, w, window_len, refseq, unsupported_allele_warning_only);
  };
  auto need = [](const mphio::GtfRecord& r, const char* k, const char* msg) -> const std::string& {
    const std::string* v = r.get(k);
    if (!v) throw Panic(msg);
    return *v;
  };
  std::string last_chrom = "not_yet_set";
  uint64_t last_start = 0;
  std::string line;
  mphio::GtfRecord record;
  while (std::getline(gtf, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!mphio::parse_gtf_line(line, record)) continue;
    const std::string& ft = record.feature;
    if (ft == "gene") {
      if (gene) {
        phase_last_gene(*gene);
        last_chrom = gene->chrom;
        last_start = gene->start();
      }
      const std::string& gene_name = need(record, "gene_name", "missing gene_name in GTF");
      if (last_chrom == record.seqname && !(last_start <= record.start))
        throw Panic("Your GTF file is not sorted correctly. Gene " + gene_name + " starts at " + std::to_string(record.start) +
                    ", while previous gene record started at " + std::to_string(last_start) + ".");
      gene.reset(new Gene{need(record, "gene_id", "missing gene_id in GTF"), gene_name, record.seqname,
                          need(record, "gene_biotype", "missing gene_biotype in GTF"), Interval::make(record.start - 1, record.end, record.frame), {}});
    } else if (ft == "transcript") {
      start_codon_found = false;
      if (!gene) throw Panic("no gene record before transcript in GTF");
      Transcript t;
      t.id = need(record, "transcript_id", "missing transcript_id attribute in GTF");
      t.biotype = need(record, "transcript_biotype", "missing transcript_biotype in GTF");
      if (record.strand == '+') t.strand = Strand::Forward;
      else if (record.strand == '-') t.strand = Strand::Reverse;
      else throw Panic("missing strand information in GTF");
      gene->transcripts.push_back(std::move(t));
    } else if (ft == "CDS") {
      if (!gene) throw Panic("no gene record before exon in GTF");
      if (gene->transcripts.empty()) throw Panic("no transcript record before exon in GTF");
      gene->transcripts.back().exons.push_back(Interval::make(record.start - 1, record.end, record.frame));
    } else if (ft == "start_codon") {
      if (start_codon_found) continue;
      start_codon_found = true;
      if (!gene) throw Panic("no gene record before start_codon in GTF");
      if (gene->transcripts.empty()) throw Panic("no transcript record before start codon in GTF");
      if (gene->transcripts.back().exons.empty()) throw Panic("no exon record before start codon in GTF");
      if (record.strand == '+') gene->transcripts.back().exons.back().start = record.start - 1;
      else gene->transcripts.back().exons.back().end = record.end;
    }
  }
  if (gene) phase_last_gene(*gene);
}

}  // namespace normal
}  // namespace oracle
