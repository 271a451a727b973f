// oracle_somatic.hpp — TEST INFRASTRUCTURE ONLY.
//
// Statement-by-statement CPU restatement of the reference's `somatic` phasing path,
// reference src/microphasing.rs (all line numbers below refer to that file):
//   has_stop_codon :42-76, bad_quality :78-93, supports_variant :95-139,
//   Observation::update_haplotype :157-197, ObservationMatrix :200-880,
//   phase_gene :882-1941, phase :1943-2131.
// Ordered maps are std::map wherever the reference uses BTreeMap so iteration orders agree.
// Unsigned arithmetic wraps like the reference's release build; conditions that panic in
// both debug and release builds (slice/drain/range bounds, unwrap on None) throw Panic.
//
// Pinned against the reference's own expected outputs (tests/golden/*/expected): see
// tests/test_oracle_golden.py.  Parity unpinned (no live reference fixture): frame > 0
// (frameshift) logic, RecordBuffer edge cases, > 64 variants per window.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "oracle_common.hpp"

namespace oracle {
namespace somatic {

using mphio::BamRecord;

// optional per-window trace for kernel debugging (tests only)
struct Trace {
  FILE* f = nullptr;
};
inline Trace& trace() {
  static Trace t;
  return t;
}

inline bool bitvector_is_set(uint64_t b, size_t k) { return (b & (uint64_t(1) << (k & 63))) != 0; }

inline uint8_t ascii_lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? uint8_t(c + 32) : c; }
inline uint8_t ascii_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? uint8_t(c - 32) : c; }
inline bool is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }

// :26-32
inline uint8_t switch_ascii_case(uint8_t c, uint8_t r) { return is_upper(r) ? ascii_lower(c) : c; }
// :34-40
inline Bytes switch_ascii_case_vec(const Bytes& v, uint8_t r) {
  Bytes o(v);
  if (is_upper(r)) for (auto& c : o) c = ascii_lower(c);
  else for (auto& c : o) c = ascii_upper(c);
  return o;
}

inline bool starts_with_at(const std::string& s, size_t c, const char* codon) {
  return s.size() >= c + 3 && s[c] == codon[0] && s[c + 1] == codon[1] && s[c + 2] == codon[2];
}
// :42-76
inline bool has_stop_codon(const std::string& peptide, bool forward) {
  if (peptide.size() < 3) return false;
  if (!forward) {
    static const char* codonlist[3] = {"TCA", "CTA", "TTA"};
    size_t c = peptide.size() - 3;
    for (;;) {
      for (auto codon : codonlist)
        if (starts_with_at(peptide, c, codon)) return true;
      if (c < 3) return false;
      c -= 3;
    }
  } else {
    static const char* codonlist[3] = {"TGA", "TAG", "TAA"};
    size_t c = 0;
    while (c < peptide.size()) {
      for (auto codon : codonlist)
        if (starts_with_at(peptide, c, codon)) return true;
      c += 3;
    }
    return false;
  }
}

// :78-93
inline bool bad_quality(const BamRecord& read, const Variant& variant) {
  if (variant.kind == Variant::SNV) {
    uint64_t relative_pos = variant.pos - uint64_t(int64_t(read.pos));
    if (relative_pos < uint64_t(read.qual.size())) {
      if (read.qual[size_t(relative_pos)] < 10) return true;
    }
  }
  return false;
}

// :95-139
inline bool supports_variant(const BamRecord& read, const Variant& variant) {
  switch (variant.kind) {
    case Variant::SNV: {
      uint64_t relative_pos = variant.pos - uint64_t(int64_t(read.pos));
      if (relative_pos < uint64_t(read.qual.size())) {
        if (read.qual[size_t(relative_pos)] < 10) return false;
      }
      uint32_t p;
      // `pos as u32` truncation of the reference is irrelevant for coordinates < 2^32
      int rc = mphio::cigar_read_pos(read.cigar, read.pos, int64_t(uint32_t(variant.pos)), &p);
      if (rc != 1) return false;
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

// :141-145
struct HaplotypeSeq {
  Bytes sequence;
  IDRecord record;
};

// :147-198
struct Observation {
  ReadPtr read;
  uint64_t haplotype = 0;
  uint64_t frame0 = 0, frame1 = 0;
  bool bad_qual = false;
  bool start_loss = false;

  void update_haplotype(size_t i, const Variant& variant, bool has_start_loss) {
    if (uint64_t(int64_t(read->pos)) > variant.pos) throw Panic("bug: read starts right of variant");
    if (variant.frameshift() > 0) frame1 += variant.pos;
    if (supports_variant(*read, variant)) {
      if (has_start_loss) start_loss = true;
      haplotype |= uint64_t(1) << (i & 63);
      frame0 += variant.frameshift();
    }
    if (bad_quality(*read, variant) || bad_qual || start_loss) {
      haplotype = 0;
      bad_qual = true;
    }
  }
};

inline bool contains_pos(const std::vector<uint64_t>& v, uint64_t p) { return std::find(v.begin(), v.end(), p) != v.end(); }

struct Writers {
  FastaWriter fasta;   // stdout: mutant windows
  TsvWriter tsv;       // --tsv
  FastaWriter normal;  // --normal-output
};

using FrameFreqs = std::map<uint64_t, std::pair<double, bool>>;

// :200-880
struct ObservationMatrix {
  std::map<uint64_t, std::vector<Observation>> observations;
  std::deque<Variant> variants;

  uint32_t ncols() const { return uint32_t(variants.size()); }
  size_t nrows() const {
    size_t n = 0;
    for (auto& kv : observations) n += kv.second.size();
    return n;
  }

  // :220-229
  void shrink_left(size_t k) {
    if (k > variants.size()) throw Panic("drain: range end out of bounds");
    variants.erase(variants.begin(), variants.begin() + long(k));
    uint32_t nc = ncols();
    uint64_t mask = nc >= 64 ? ~uint64_t(0) : ((uint64_t(1) << nc) - 1);
    for (auto& kv : observations)
      for (auto& obs : kv.second) obs.haplotype &= mask;
  }

  // :232-256
  void extend_right(const std::vector<Variant>& new_variants, const std::vector<uint64_t>& start_loss) {
    size_t k = new_variants.size();
    if (k > 0)
      for (auto& kv : observations)
        for (auto& obs : kv.second) obs.haplotype <<= (k & 63);
    for (auto& kv : observations)
      for (auto& obs : kv.second) {
        size_t i = 0;
        for (auto it = new_variants.rbegin(); it != new_variants.rend(); ++it, ++i)
          obs.update_haplotype(i, *it, contains_pos(start_loss, it->pos));
      }
    for (auto& v : new_variants) variants.push_back(v);
  }

  // :259-278
  void cleanup_reads(uint64_t interval_end, bool reverse) {
    auto it = observations.lower_bound(interval_end);
    if (!reverse) {
      observations.erase(observations.begin(), it);  // keep keys >= interval_end
    } else {
      observations.erase(it, observations.end());  // keep keys < interval_end
    }
  }

  // :281-294
  bool contains(const BamRecord& read) const {
    uint64_t pos = uint64_t(int64_t(read.pos));
    auto it = observations.find(pos);
    if (it != observations.end()) {
      for (auto& obs : it->second)
        if (obs.read->qname == read.qname) return true;
      return false;
    }
    return false;
  }

  // :297-343
  void push_read(const ReadPtr& read, uint64_t interval_end, uint64_t interval_start, bool reverse,
                 const std::vector<uint64_t>& start_loss) {
    uint64_t end_pos = uint64_t(read->end_pos());
    uint64_t start_pos = uint64_t(int64_t(read->pos));
    if (end_pos >= interval_end && start_pos <= interval_start && !contains(*read)) {
      Observation obs;
      obs.read = read;
      size_t i = 0;
      for (auto it = variants.rbegin(); it != variants.rend(); ++it, ++i)
        obs.update_haplotype(i, *it, contains_pos(start_loss, it->pos));
      uint64_t pos = reverse ? start_pos : end_pos;
      if (obs.bad_qual) return;
      observations[pos].push_back(std::move(obs));
    }
  }

  static const uint8_t& ref_at(const Bytes& refseq, uint64_t idx) {
    if (idx >= refseq.size()) throw Panic("index out of bounds: refseq");
    return refseq[size_t(idx)];
  }
  static void ref_extend(Bytes& dst, const Bytes& refseq, uint64_t a, uint64_t b) {
    if (a > b || b > refseq.size()) throw Panic("slice index out of range: refseq");
    dst.insert(dst.end(), refseq.begin() + long(a), refseq.begin() + long(b));
  }

  // :353-879
  std::pair<std::vector<HaplotypeSeq>, FrameFreqs> print_haplotypes(
      const Gene& gene, const Transcript& transcript, uint64_t offset, uint64_t splice_end, uint64_t splice_pos,
      uint64_t splice_gap, uint64_t exon_end, uint64_t exon_start, uint64_t window_len, const Bytes& refseq, Writers& w,
      bool is_short_exon, uint64_t frame_in, FrameFreqs frameshift_frequencies, bool is_first_exon_window) const {
    std::vector<const Variant*> variants;
    for (auto& v : variants_deque()) variants.push_back(&v);
    const bool reverse_strand = transcript.strand == Strand::Reverse;
    if (reverse_strand) std::reverse(variants.begin(), variants.end());
    uint64_t frame = frame_in;
    uint64_t frame_depth = 0;
    // count haplotypes :383-411
    std::map<std::pair<uint64_t, uint64_t>, uint64_t> haplotypes;
    for (auto& kv : observations)
      for (auto& obs : kv.second) {
        if (obs.bad_qual) continue;
        if (frame > 0 && obs.frame0 != frame && obs.frame1 != 0) continue;
        frame_depth += 1;
        if (frame > 0) haplotypes[{obs.haplotype, frame}] += 1;
        else haplotypes[{obs.haplotype, obs.frame0}] += 1;
      }
    const char* strand = reverse_strand ? "Reverse" : "Forward";
    const bool has_frameshift = frame > 0;
    std::vector<HaplotypeSeq> haplotypes_vec;
    if (haplotypes.empty()) haplotypes[{0, 0}] = 0;
    uint64_t shift_in_window = 0;

    if (trace().f) {
      fprintf(trace().f, "W\t%s\t%llu\t%llu\t%llu\t%zu\t%llu\t%zu", transcript.id.c_str(), (unsigned long long)offset,
              (unsigned long long)splice_end, (unsigned long long)frame_in, nrows(), (unsigned long long)frame_depth,
              variants.size());
      for (auto& h : haplotypes)
        fprintf(trace().f, "\t%llu:%llu:%llu", (unsigned long long)h.first.first, (unsigned long long)h.first.second,
                (unsigned long long)h.second);
      fputc('\n', trace().f);
    }

    Bytes seq, germline_seq;
    for (auto& hk : haplotypes) {
      const uint64_t haplotype = hk.first.first;
      const uint64_t haplotype_frame = hk.first.second;
      const uint64_t count = hk.second;
      bool indel = false, insertion = false, shift_is_set = false;
      seq.clear();
      germline_seq.clear();
      uint32_t n_somatic = 0, n_variants = 0;
      double freq = count == 0 ? 0.0 : double(count) / double(frame_depth);
      uint32_t depth = uint32_t(nrows());
      uint64_t i = offset;
      size_t j = 0;
      const uint64_t window_end = splice_end;
      std::vector<int> variant_profile;
      const uint64_t gs = gene.start();
      if (variants.empty()) {
        ref_extend(germline_seq, refseq, offset - gs, window_end - gs);
        ref_extend(seq, refseq, offset - gs, window_end - gs);
      } else {
        while (i < window_end) {
          bool broke = false;
          while (j < variants.size() && i == variants[j]->pos) {
            const Variant& vj = *variants[j];
            shift_in_window = shift_in_window > 0 ? shift_in_window : vj.frameshift();
            size_t bit_pos = reverse_strand ? j : variants.size() - 1 - j;
            if (bitvector_is_set(haplotype, bit_pos)) {
              if (shift_in_window > 0) {
                shift_is_set = true;
                frameshift_frequencies[vj.frameshift()] = {freq, !vj.is_germline()};
                frameshift_frequencies[0] = {1.0 - freq, false};
              }
              if (vj.kind == Variant::SNV) {
                uint8_t r = ref_at(refseq, i - gs);
                if (vj.is_germline()) germline_seq.push_back(switch_ascii_case(vj.alt, r));
                else germline_seq.push_back(r);
                seq.push_back(switch_ascii_case(vj.alt, r));
                i += 1;
              } else if (vj.kind == Variant::Insertion) {
                uint8_t r = ref_at(refseq, i - gs);
                Bytes sw = switch_ascii_case_vec(vj.seq, r);
                if (vj.is_germline()) germline_seq.insert(germline_seq.end(), sw.begin(), sw.end());
                else indel = true;
                seq.insert(seq.end(), sw.begin(), sw.end());
                insertion = true;
                i += 1;
              } else {
                if (reverse_strand && vj.end_pos() >= window_end) {
                  broke = true;  // :549-552 `break` leaves only the inner while loop
                  break;
                }
                if (vj.is_germline() || i == window_end - 1) {
                  germline_seq.push_back(ref_at(refseq, i - gs));
                } else {
                  ref_extend(germline_seq, refseq, i - gs, i + vj.len + 1 - gs);
                  indel = true;
                }
                seq.push_back(ref_at(refseq, i - gs));
                i += vj.len + 1;
              }
              if (!vj.is_germline()) { n_somatic += 1; variant_profile.push_back(2); }
              else variant_profile.push_back(1);
              n_variants += 1;
            } else {
              variant_profile.push_back(0);
            }
            j += 1;
          }
          (void)broke;
          if (i < window_end) {
            uint8_t r = ref_at(refseq, i - gs);
            seq.push_back(r);
            germline_seq.push_back(r);
            i += 1;
          }
        }
      }
      // :604-631
      double frame_frequency = freq;
      if (shift_is_set && frame == 0) frame = shift_in_window;
      frameshift_frequencies.emplace(frame, std::make_pair(0.0, false));
      if (shift_in_window == 0) frame_frequency = freq * frameshift_frequencies.at(frame).first;
      if (shift_in_window == 0 && haplotype_frame > 0 && frame == 0) frame_frequency = 0.0;
      if ((indel && insertion) ||
          (shift_in_window == 0 && (frameshift_frequencies.at(frame).second || (has_frameshift && germline_seq != seq)))) {
        germline_seq.clear();
      }
      uint64_t this_window_len = seq.size() < window_len ? uint64_t(seq.size()) : window_len;
      uint64_t normal_window_len = indel ? (germline_seq.size() < window_len ? uint64_t(germline_seq.size()) : window_len) : this_window_len;
      std::string fasta_id = mphfmt::record_id(seq.data(), seq.size(), transcript.id, offset, strand[0]);
      auto slice = [](const Bytes& b, uint64_t a, uint64_t e) -> std::string {
        if (a > e || e > b.size()) throw Panic("slice index out of range");
        return std::string(b.begin() + long(a), b.begin() + long(e));
      };
      // :677-693
      std::string normal_peptide;
      if (germline_seq.empty()) normal_peptide = "";
      else if (splice_pos == 1) normal_peptide = slice(germline_seq, splice_gap, germline_seq.size());
      else if (splice_pos == 0) normal_peptide = slice(germline_seq, 0, normal_window_len);
      else normal_peptide = slice(germline_seq, 0, germline_seq.size());
      std::string neopeptide;
      if (splice_pos == 1) neopeptide = slice(seq, splice_gap, seq.size());
      else if (splice_pos == 0) neopeptide = insertion ? slice(seq, 0, seq.size()) : slice(seq, 0, this_window_len);
      else neopeptide = slice(seq, 0, seq.size());
      bool stop_gain = has_stop_codon(neopeptide, !reverse_strand);
      bool remove_peptide = false;
      if (stop_gain && splice_pos != 2 && (window_len == this_window_len || indel) && !is_first_exon_window &&
          ((normal_peptide != neopeptide) || !indel || std::fabs(freq - 1.0) < std::numeric_limits<double>::epsilon())) {
        remove_peptide = true;
        if (frame == 0) frameshift_frequencies[frame] = {0.0, false};
        else frameshift_frequencies.erase(frame);
      }
      // :720-764
      uint32_t n_variantsites = 0, n_som_variantsites = 0;
      std::vector<std::string> somatic_p_changes, germline_p_changes, somatic_var_pos, germline_var_pos, variantsites_pos;
      for (size_t c = 0; c < variants.size(); ++c) {
        if (c < variant_profile.size()) {
          if (variant_profile[c] == 2) {
            somatic_var_pos.push_back(std::to_string(variants[c]->pos + 1));
            somatic_p_changes.push_back(variants[c]->prot_change());
          } else if (variant_profile[c] == 1) {
            germline_var_pos.push_back(std::to_string(variants[c]->pos + 1));
            germline_p_changes.push_back(variants[c]->prot_change());
          }
        }
        if (c == 0 || variants[c]->pos != variants[c - 1]->pos) {
          n_variantsites += 1;
          variantsites_pos.push_back(std::to_string(variants[c]->pos + 1));
          if (!variants[c]->is_germline()) n_som_variantsites += 1;
        }
      }
      uint64_t inframe_offset = splice_pos == 0 ? offset + 1 : offset + 1 + splice_gap;
      IDRecord record;
      record.id = fasta_id; record.transcript = transcript.id; record.gene_id = gene.id; record.gene_name = gene.name;
      record.chrom = gene.chrom; record.offset = inframe_offset; record.frame = frame; record.freq = frame_frequency;
      record.depth = depth; record.nvar = n_variants; record.nsomatic = n_somatic; record.nvariant_sites = n_variantsites;
      record.nsomvariant_sites = n_som_variantsites; record.strand = strand;
      record.variant_sites = IDRecord::join(variantsites_pos);
      record.somatic_positions = IDRecord::join(somatic_var_pos);
      record.somatic_aa_change = IDRecord::join(somatic_p_changes);
      record.germline_positions = IDRecord::join(germline_var_pos);
      record.germline_aa_change = IDRecord::join(germline_p_changes);
      record.normal_sequence = normal_peptide;
      record.mutant_sequence = neopeptide;
      // :796-832  (rest/start are debug-only)
      (void)exon_end; (void)exon_start;
      HaplotypeSeq hap_seq;
      hap_seq.record = record;
      hap_seq.record.normal_sequence.assign(germline_seq.begin(), germline_seq.end());
      hap_seq.record.mutant_sequence.assign(seq.begin(), seq.end());
      if (!remove_peptide || frame == 0) haplotypes_vec.push_back(std::move(hap_seq));
      // :839-875
      if ((record.nsomatic > 0 || has_frameshift) && !is_short_exon && germline_seq != seq && record.freq > 0.0 &&
          (!stop_gain || has_frameshift)) {
        if (splice_pos == 1) {
          if (splice_gap > seq.size()) throw Panic("slice index out of range");
          w.fasta.write(record.id, seq.data() + splice_gap, seq.size() - size_t(splice_gap));
        } else if (splice_pos == 0) {
          if (this_window_len > seq.size()) throw Panic("slice index out of range");
          w.fasta.write(record.id, seq.data(), size_t(this_window_len));
        }
        if (!germline_seq.empty()) {
          if (splice_pos == 1) {
            if (splice_gap > germline_seq.size()) throw Panic("slice index out of range");
            w.normal.write(record.id, germline_seq.data() + splice_gap, germline_seq.size() - size_t(splice_gap));
          } else if (splice_pos == 0) {
            if (this_window_len > germline_seq.size()) throw Panic("slice index out of range");
            w.normal.write(record.id, germline_seq.data(), size_t(this_window_len));
          }
        }
        w.tsv.row(IDRecord::header(), record.fields());
      }
    }
    return {std::move(haplotypes_vec), std::move(frameshift_frequencies)};
  }

  const std::deque<Variant>& variants_deque() const { return variants; }
};

template <class Map>
inline size_t count_range(const Map& m, uint64_t a, uint64_t b) {
  if (a > b) throw Panic("range start is greater than range end in BTreeMap");
  size_t n = 0;
  for (auto it = m.lower_bound(a); it != m.end() && it->first < b; ++it) n += it->second.size();
  return n;
}

struct GeneInputs {
  BamRecordBuffer* read_buffer;
  VcfRecordBuffer* variant_buffer;
  mphio::FastaIndexed* fasta;
};

// statistics for bench.py's cpu_baseline (windows = main-ORF print_haplotypes calls)
struct Stats {
  uint64_t windows = 0, read_windows = 0, print_calls = 0;
};
inline Stats& stats() {
  static Stats s;
  return s;
}

// :882-1941
inline void phase_gene(const Gene& gene, GeneInputs& in, Writers& w, uint64_t window_len, Bytes& refseq,
                       bool unsupported_allele_warning_only) {
  const uint64_t end_overflow = 100;
  in.fasta->fetch(gene.chrom, gene.start(), gene.end() + end_overflow, refseq);
  std::map<uint64_t, std::vector<Variant>> variant_tree;
  std::map<uint64_t, std::vector<ReadPtr>> read_tree;
  in.read_buffer->fetch(gene.chrom, gene.start(), gene.end());
  uint64_t max_read_len = 0;
  for (auto& rec : in.read_buffer->inner) {
    if (rec->mapq < 5) continue;
    if (uint64_t(rec->l_seq) > max_read_len) max_read_len = rec->l_seq;
    read_tree[uint64_t(int64_t(rec->pos))].push_back(rec);
  }
  in.variant_buffer->fetch(gene.chrom, gene.start(), gene.end());
  for (auto& rec : in.variant_buffer->ring)
    variant_tree[uint64_t(rec.pos)] = variants_from_record(*in.variant_buffer->vcf, rec, unsupported_allele_warning_only);

  for (const Transcript& transcript : gene.transcripts) {
    if (!transcript.is_coding()) continue;
    const bool fwd = transcript.strand == Strand::Forward;
    const size_t exon_number = transcript.exons.size();
    ObservationMatrix observations;
    std::map<uint64_t, uint64_t> frameshifts;
    std::vector<uint64_t> deletions;
    if (fwd) frameshifts[0] = 0;
    else frameshifts[gene.end()] = 0;
    uint64_t exon_rest = 0;
    std::vector<HaplotypeSeq> prev_hap_vec, hap_vec;
    FrameFreqs frameshift_frequencies;
    frameshift_frequencies[0] = {1.0, false};
    std::vector<uint64_t> start_loss;
    size_t last_window_vars = 0;
    uint64_t exon_count = 0;
    for (const Interval& exon : transcript.exons) {
      if (frameshifts.empty()) break;
      if (exon.start > exon.end) continue;
      exon_count += 1;
      const uint64_t exon_len = exon.end - exon.start;
      const uint64_t current_exon_offset = exon_count == 1 ? exon.frame : (exon_rest == 0 ? 0 : 3 - exon_rest);
      const bool is_last_exon = exon_count == exon_number;
      const bool is_first_exon = exon_count == 1;
      const bool is_short_exon =
          exon_len < 3 ? true : window_len >= exon_len - current_exon_offset - (3 - current_exon_offset) % 3;
      uint64_t exon_window_len =
          !is_short_exon ? window_len : (exon_len - current_exon_offset) - ((exon_len - current_exon_offset) % 3);
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
        const bool read_through = is_last_exon && !valid;
        if (!valid) break;
        if (max_read_len < exon_window_len) break;
        const uint64_t rest = fwd ? exon.end - (offset + exon_window_len) : offset - exon.start;
        const bool is_last_exon_window = rest < 3;
        uint64_t splice_side_offset, splice_end, splice_gap, splice_pos;
        if (fwd) {
          if (is_short_exon) {
            splice_side_offset = offset - current_exon_offset; splice_end = offset + exon_window_len + rest;
            splice_gap = current_exon_offset + rest; splice_pos = 2;
          } else if (is_first_exon_window) {
            if (is_last_exon_window) {
              splice_side_offset = offset - current_exon_offset; splice_end = offset + exon_window_len + rest;
              splice_gap = current_exon_offset + rest; splice_pos = 2;
            } else {
              splice_side_offset = offset - current_exon_offset; splice_end = offset + exon_window_len;
              splice_gap = current_exon_offset; splice_pos = 1;
            }
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
            splice_side_offset = offset; splice_end = offset + exon_window_len + current_exon_offset;
            splice_gap = current_exon_offset; splice_pos = 0;
          } else if (is_last_exon_window) {
            splice_side_offset = offset - rest; splice_end = offset + exon_window_len; splice_gap = rest; splice_pos = 1;
          } else {
            splice_side_offset = offset; splice_end = offset + exon_window_len; splice_gap = 0; splice_pos = 0;
          }
        }
        // :1119-1178
        const size_t nvars = count_range(variant_tree, splice_side_offset, splice_end);
        last_window_vars = nvars;
        size_t added_vars;
        if (is_first_exon_window) added_vars = nvars;
        else if (is_short_exon && !read_through) added_vars = 0;
        else if (reached_end && !read_through) added_vars = 0;
        else if (splice_side_offset > old_offset) added_vars = count_range(variant_tree, old_end, splice_end);
        else added_vars = count_range(variant_tree, splice_side_offset, old_offset);
        size_t deleted_vars;
        if (offset == old_offset || (is_short_exon && !read_through)) deleted_vars = 0;
        else if (splice_side_offset > old_offset) deleted_vars = count_range(variant_tree, old_offset, splice_side_offset);
        else deleted_vars = count_range(variant_tree, splice_end, old_end);
        if (is_last_exon_window && !read_through) reached_end = true;

        // :1191-1249
        std::vector<ReadPtr> reads;
        {
          uint64_t lo, hi;
          bool wide = !fwd || offset == exon.start + current_exon_offset;
          if (wide) { lo = splice_side_offset - (max_read_len - exon_window_len); hi = splice_side_offset + 1; }
          else { lo = splice_side_offset; hi = splice_side_offset + 1; }
          if (lo > hi) throw Panic("range start is greater than range end in BTreeMap");
          for (auto it = read_tree.lower_bound(lo); it != read_tree.end() && it->first < hi; ++it)
            for (auto& r : it->second) reads.push_back(r);
        }
        const bool reverse = !fwd;
        if (reverse) observations.cleanup_reads(splice_side_offset + 1, reverse);
        else observations.cleanup_reads(splice_end, reverse);
        observations.shrink_left(deleted_vars);
        for (auto& read : reads) observations.push_read(read, splice_end, splice_side_offset, reverse, start_loss);

        // :1280-1296
        std::vector<Variant> variants;
        {
          if (splice_side_offset > splice_end) throw Panic("range start is greater than range end in BTreeMap");
          std::vector<const std::vector<Variant>*> groups;
          for (auto it = variant_tree.lower_bound(splice_side_offset); it != variant_tree.end() && it->first < splice_end; ++it)
            groups.push_back(&it->second);
          if (reverse) std::reverse(groups.begin(), groups.end());
          size_t skip = nvars - added_vars, idx = 0;  // usize subtraction wraps in the release build
          for (auto g : groups)
            for (auto& v : *g) {
              if (idx++ >= skip) variants.push_back(v);
            }
        }
        // :1299-1342
        for (const Variant& variant : variants) {
          bool is_start_loss = fwd ? (is_first_exon && variant.pos >= exon.start && variant.pos < exon.start + 3)
                                   : (is_first_exon && variant.pos < exon.end && variant.pos >= exon.end - 3);
          if (is_start_loss) start_loss.push_back(variant.pos);
          if (variant.kind == Variant::Deletion) deletions.push_back(fwd ? variant.end_pos() : variant.pos);
          uint64_t s = variant.frameshift();
          if ((s % 3) > 0) {
            std::vector<uint64_t> previous;
            for (auto& kv : frameshifts) previous.push_back(kv.second + s);
            for (uint64_t s_ : previous) {
              if (fwd) frameshifts[variant.end_pos()] = s_ % 3;
              else frameshifts[variant.pos] = s_ % 3;
            }
          }
        }
        observations.extend_right(variants, start_loss);
        uint64_t stopped_frameshift = 3;
        // active frameshifts :1347-1350 — snapshot (the map is not mutated while iterating)
        std::vector<std::pair<uint64_t, uint64_t>> active;
        if (fwd) { for (auto it = frameshifts.begin(); it != frameshifts.end() && it->first < offset; ++it) active.push_back(*it); }
        else { for (auto it = frameshifts.lower_bound(offset + exon_window_len); it != frameshifts.end(); ++it) active.push_back(*it); }
        bool closed_deletion = deletions.empty() ? false : (fwd ? deletions[0] < offset : deletions[0] >= offset + exon_window_len);
        uint64_t frameshift_count = 0;
        bool main_orf = false;
        for (auto& kf : active) {
          const uint64_t key = kf.first, frameshift = kf.second;
          frameshift_count += 1;
          if (frameshift == 0) main_orf = true;
          const uint64_t coding_shift = fwd ? offset - exon.start : exon.end - offset;
          const bool has_frameshift = frameshift > 0;
          if (coding_shift % 3 == (frameshift + current_exon_offset) % 3 || (is_short_exon && !read_through)) {
            if (!has_frameshift && !read_through) {
              exon_rest = fwd ? exon.end - (offset + exon_window_len) : offset - exon.start;
              if (exon_window_len < 3) exon_rest = exon_window_len;
            }
            if (frameshift == 0) {
              stats().windows += 1;
              stats().read_windows += observations.nrows();
            }
            stats().print_calls += 1;
            auto res = observations.print_haplotypes(gene, transcript, splice_side_offset, splice_end, splice_pos, splice_gap,
                                                     exon.end, exon.start, exon_window_len, refseq, w, is_short_exon, frameshift,
                                                     frameshift_frequencies, is_first_exon_window);
            frameshift_frequencies = std::move(res.second);
            if (res.first.empty() || !frameshift_frequencies.count(frameshift)) stopped_frameshift = key;
            if (closed_deletion) deletions.clear();
            if (exon_rest < 3 && (!is_short_exon || is_first_exon) && !has_frameshift && !read_through) prev_hap_vec = std::move(res.first);
            else hap_vec = std::move(res.first);
            if (frameshift != 0 && frameshift_frequencies.count(frameshift) && frameshift_frequencies.at(frameshift).first == 0.0)
              stopped_frameshift = key;
          }
        }
        if (frameshift_count == 0 || !main_orf || !frameshift_frequencies.count(0)) {
          frameshifts.clear();
          break;
        }
        if (stopped_frameshift != 3) {
          auto it = frameshifts.find(stopped_frameshift);
          if (it == frameshifts.end()) throw Panic("called `Option::unwrap()` on a `None` value (frameshifts.get)");
          if (it->second != 0) frameshifts.erase(it);
        }
        if (frameshifts.empty()) break;
        if (frameshift_frequencies.at(0).first == 0.0 && frameshifts.size() == 1) {
          frameshifts.clear();
          break;
        }
        const bool at_splice_side = fwd ? offset - current_exon_offset == exon.start
                                        : offset + exon_window_len + current_exon_offset == exon.end;
        is_first_exon_window = false;
        // :1505-1908 splice-junction merge
        if (at_splice_side && !is_first_exon) {
          const std::vector<HaplotypeSeq>& first_hap_vec = fwd ? hap_vec : prev_hap_vec;
          const std::vector<HaplotypeSeq>& sec_hap_vec = fwd ? prev_hap_vec : hap_vec;
          struct OutVal { Bytes mt; IDRecord rec; Bytes wt; };
          std::map<std::tuple<uint64_t, Bytes, Bytes>, OutVal> output_map;
          std::vector<HaplotypeSeq> new_hap_vec;
          for (const HaplotypeSeq& hapseq : first_hap_vec) {
            const IDRecord& record = hapseq.record;
            const std::string& wt_sequence = record.normal_sequence;
            const std::string& mt_sequence = record.mutant_sequence;
            for (const HaplotypeSeq& prev_hapseq : sec_hap_vec) {
              const IDRecord& prev_record = prev_hapseq.record;
              const std::string& prev_wt_sequence = prev_record.normal_sequence;
              const std::string& prev_mt_sequence = prev_record.mutant_sequence;
              const std::string new_wt = prev_wt_sequence + wt_sequence;
              const Bytes new_wt_sequence(new_wt.begin(), new_wt.end());
              std::vector<std::string> new_mt_sequences;
              if (wt_sequence != mt_sequence) {
                new_mt_sequences.push_back(prev_wt_sequence + mt_sequence);
                if (prev_wt_sequence != prev_mt_sequence) {
                  new_mt_sequences.push_back(prev_mt_sequence + wt_sequence);
                  new_mt_sequences.push_back(prev_mt_sequence + mt_sequence);
                }
              } else {
                new_mt_sequences.push_back(prev_mt_sequence + mt_sequence);
              }
              const double eps = std::numeric_limits<double>::epsilon();
              if (is_short_exon && !is_last_exon) {
                double out_freq = std::fabs(record.freq - prev_record.freq) < eps ? record.freq : record.freq * prev_record.freq;
                HaplotypeSeq nh;
                nh.record = prev_record.update(record, 0, record.frame, out_freq, new_wt_sequence, new_wt_sequence, window_len);
                new_hap_vec.push_back(std::move(nh));
              }
              for (const std::string& new_mt : new_mt_sequences) {
                const Bytes new_mt_sequence(new_mt.begin(), new_mt.end());
                if (is_short_exon && !is_last_exon) {
                  double out_freq = std::fabs(record.freq - prev_record.freq) < eps ? record.freq : record.freq * prev_record.freq;
                  HaplotypeSeq nh;
                  nh.record = prev_record.update(record, 0, record.frame, out_freq, new_wt_sequence, new_mt_sequence, window_len);
                  new_hap_vec.push_back(std::move(nh));
                  continue;
                }
                std::vector<std::pair<uint64_t, uint64_t>> active2;
                if (fwd) { for (auto it = frameshifts.begin(); it != frameshifts.end() && it->first < offset; ++it) active2.push_back(*it); }
                else { for (auto it = frameshifts.lower_bound(offset + exon_window_len); it != frameshifts.end(); ++it) active2.push_back(*it); }
                for (auto& pf : active2) {
                  const uint64_t pos = pf.first, frameshift = pf.second;
                  frameshift_frequencies.emplace(frameshift, std::make_pair(0.0, false));
                  const bool shift_in_window = fwd ? pos >= prev_record.offset : pos < record.offset + exon_window_len;
                  const bool somatic_shift = frameshift_frequencies.at(frameshift).second;
                  const double frameshift_freq = frameshift_frequencies.at(frameshift).first;
                  const double ff0 = frameshift_frequencies.at(0).first;
                  const double main_orf_freq = ff0 == 0.0 ? frameshift_freq : ff0;
                  const double shift_orf_freq = shift_in_window ? frameshift_freq : (ff0 == 0.0 ? frameshift_freq : ff0);
                  const double variant_freq_record = fwd ? record.freq / main_orf_freq : record.freq / shift_orf_freq;
                  const double variant_freq_prev_record = fwd ? prev_record.freq / shift_orf_freq : prev_record.freq / main_orf_freq;
                  const double freq_record = ff0 == 0.0 ? frameshift_freq : variant_freq_record * frameshift_freq;
                  const double freq_prev_record = ff0 == 0.0 ? frameshift_freq : variant_freq_prev_record * frameshift_freq;
                  const double out_freq = std::fabs(record.freq - prev_record.freq) < eps ? freq_record : freq_record * freq_prev_record;
                  const uint64_t out_shift = shift_in_window ? 0 : frameshift;
                  uint64_t splice_offset = 3 - out_shift;
                  if (!fwd && exon_rest < 3) splice_offset += exon_rest;
                  size_t end_offset = 3 + size_t(out_shift);
                  if (is_last_exon_window) end_offset = 0;
                  if (uint64_t(new_mt_sequence.size()) < 2 * window_len) {
                    if (fwd) splice_offset = 0;
                    else end_offset = 0;
                  }
                  // `new_mt_sequence.len() - end_offset` is usize arithmetic: wraps in release
                  while (splice_offset + window_len <= uint64_t(new_mt_sequence.size() - end_offset)) {
                    auto sub = [](const Bytes& b, uint64_t a, uint64_t e) -> Bytes {
                      if (a > e || e > b.size()) throw Panic("slice index out of range");
                      return Bytes(b.begin() + long(a), b.begin() + long(e));
                    };
                    Bytes out_wt_seq;
                    if (splice_offset + window_len <= uint64_t(new_wt_sequence.size())) {
                      if (fwd) out_wt_seq = sub(new_wt_sequence, splice_offset, splice_offset + window_len);
                      else out_wt_seq = sub(new_wt_sequence, new_wt_sequence.size() - end_offset - size_t(window_len),
                                            new_wt_sequence.size() - end_offset);
                    }
                    Bytes out_mt_seq;
                    if (fwd) out_mt_seq = sub(new_mt_sequence, splice_offset, splice_offset + window_len);
                    else out_mt_seq = sub(new_mt_sequence, new_mt_sequence.size() - end_offset - size_t(window_len),
                                          new_mt_sequence.size() - end_offset);
                    if (out_shift > 0 && out_wt_seq == out_mt_seq && somatic_shift) out_wt_seq.clear();
                    if (out_wt_seq == out_mt_seq || (out_wt_seq.empty() && frameshift == 0)) {
                      if (fwd) splice_offset += 3;
                      else end_offset += 3;
                      continue;
                    }
                    const uint64_t out_offset = fwd ? splice_offset : uint64_t(end_offset);
                    IDRecord out_record = fwd ? prev_record.update(record, out_offset, frameshift, out_freq, out_wt_seq, out_mt_seq, window_len)
                                              : record.update(prev_record, out_offset, frameshift, out_freq, out_wt_seq, out_mt_seq, window_len);
                    auto id_tuple = std::make_tuple(out_offset, out_mt_seq, out_wt_seq);
                    auto itx = output_map.find(id_tuple);
                    double old_freq = itx != output_map.end() ? itx->second.rec.freq : 0.0;
                    output_map[id_tuple] = OutVal{out_mt_seq, out_record.add_freq(old_freq), out_wt_seq};
                    if (fwd) splice_offset += 3;
                    else end_offset += 3;
                  }
                }
              }
            }
          }
          if (is_short_exon && !is_last_exon) {
            prev_hap_vec = std::move(new_hap_vec);
          } else {
            for (auto& kv : output_map) {
              const IDRecord& out_record = kv.second.rec;
              const Bytes& out_mt_seq = kv.second.mt;
              const Bytes& out_wt_seq = kv.second.wt;
              if (out_mt_seq != out_wt_seq) {
                if (window_len > out_mt_seq.size()) throw Panic("slice index out of range");
                w.fasta.write(out_record.id, out_mt_seq.data(), size_t(window_len));
                if (!out_wt_seq.empty()) {
                  if (window_len > out_wt_seq.size()) throw Panic("slice index out of range");
                  w.normal.write(out_record.id, out_wt_seq.data(), size_t(window_len));
                }
                w.tsv.row(IDRecord::header(), out_record.fields());
              }
            }
            if (is_short_exon) prev_hap_vec = std::move(new_hap_vec);
          }
        }
        old_offset = splice_side_offset;
        old_end = splice_end;
        if (fwd) offset += 1;
        else offset -= 1;
        if (frameshifts.empty()) break;
        if (is_short_exon) break;  // :1928-1931 — still inside `loop`: a short exon gets exactly one window
      }
    }
  }
}

// :1943-2131
inline void phase(mphio::FastaIndexed& fasta, std::istream& gtf, mphio::VcfFile& vcf, mphio::BamFile& bam, Writers& w,
                  uint64_t window_len, bool unsupported_allele_warning_only) {
  BamRecordBuffer read_buffer;
  read_buffer.load(bam);
  VcfRecordBuffer variant_buffer;
  variant_buffer.vcf = &vcf;
  Bytes refseq;
  GeneInputs in{&read_buffer, &variant_buffer, &fasta};
  std::unique_ptr<Gene> gene;
  bool start_codon_found = false, three_prime_found = false;
  auto phase_last_gene = [&](const Gene& g) {
    if (g.biotype == "protein_coding") phase_gene(g, in, w, window_len, refseq, unsupported_allele_warning_only);
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
      if (last_chrom == record.seqname) {
        if (!(last_start <= record.start))
          throw Panic("Your GTF file is not sorted correctly. Gene " + gene_name + " starts at " + std::to_string(record.start) +
                      ", while previous gene record started at " + std::to_string(last_start) + ".");
      }
      gene.reset(new Gene{need(record, "gene_id", "missing gene_id in GTF"), gene_name, record.seqname,
                          need(record, "gene_biotype", "missing gene_biotype in GTF"),
                          Interval::make(record.start - 1, record.end, record.frame), {}});
    } else if (ft == "transcript") {
      start_codon_found = false;
      three_prime_found = false;
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
    } else if (ft == "three_prime_utr") {
      if (!gene) throw Panic("no gene record before exon in GTF");
      if (gene->transcripts.empty()) throw Panic("no transcript record before exon in GTF");
      if (three_prime_found) {
        gene->transcripts.back().exons.push_back(Interval::make(record.start - 1, record.end, record.frame));
      } else {
        three_prime_found = true;
        if (gene->transcripts.back().exons.empty()) throw Panic("no exon record before start codon in GTF");
        if (record.strand == '+') gene->transcripts.back().exons.back().end = record.end;
        else gene->transcripts.back().exons.back().start = record.start - 1;
      }
    }
  }
  if (gene) phase_last_gene(*gene);
}

}  // namespace somatic
}  // namespace oracle
