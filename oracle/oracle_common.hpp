// oracle_common.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the reference's domain types and I/O-buffer semantics:
//   * Variant / Annotation / Gene / Transcript / Interval / IDRecord  — reference src/common.rs:15-569
//   * bam::RecordBuffer / bcf::buffer::RecordBuffer fetch semantics    — rust-htslib 0.36 (crate not
//     vendored under /root/reference; restated from its published behaviour)
//   * fasta::Writer / csv::Writer output formats                       — SURVEY.md Appendix B
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or execute anything under oracle/. The product path never does.
#pragma once
#include <cstdint>
#include <cstdio>
#include <deque>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../microphaser_b200/csrc/io/fmt_util.hpp"
#include "../microphaser_b200/csrc/io/hts_io.hpp"

namespace oracle {

// Rust panic (exit status 101 in the reference binary)
struct Panic : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// Err(..) bubbled to main (exit status 1)
struct Failure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

using Bytes = std::vector<uint8_t>;

inline void log_warn(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); }

// ------------------------------------------------------------------ Variant (common.rs:38-222)
struct Variant {
  enum Kind { SNV, Insertion, Deletion } kind;
  uint64_t pos;
  uint8_t alt = 0;   // SNV
  Bytes seq;         // Insertion: whole ALT allele incl. anchor base
  uint64_t len = 0;  // Insertion: alt.len()-1 ; Deletion: ref.len()-1 or |SVLEN|
  bool is_germline_ = true;
  std::string prot_change_;

  uint64_t end_pos() const {  // common.rs:185-191
    return kind == Deletion ? pos + len - 1 : pos;
  }
  bool is_germline() const { return is_germline_; }
  const std::string& prot_change() const { return prot_change_; }
  uint64_t frameshift() const {  // common.rs:215-221
    switch (kind) {
      case SNV: return 0;
      case Deletion: return len % 3;
      default: return (3 - ((uint64_t(seq.size()) - 1) % 3)) % 3;
    }
  }
};

// Variant::new (common.rs:71-175) + Annotation::new (common.rs:21-35)
inline std::vector<Variant> variants_from_record(const mphio::VcfFile& vcf, const mphio::VcfRecord& rec,
                                                 bool unsupported_allele_warning_only) {
  auto warn_or_error = [&](const std::string& msg) {
    if (unsupported_allele_warning_only) log_warn(msg);
    else { log_warn(msg); throw Panic(msg); }
  };
  // rec.info(b"SOMATIC").flag().unwrap_or(false): Err when the tag is not a declared Flag
  bool is_germline = !(vcf.somatic_defined && rec.somatic_flag);
  // Annotation::new: Err(_) -> "" when ANN is undeclared; Ok(None).unwrap() panics when declared but absent
  std::string info;
  if (vcf.ann_defined) {
    if (!rec.has_ann) throw Panic("called `Option::unwrap()` on a `None` value (INFO/ANN declared but absent)");
    info = rec.ann_first;
  }
  std::string pc;
  if (!info.empty()) {
    auto fields = mphio::split(info, '|');
    for (auto& f : fields)
      if (f.find("p.") != std::string::npos) { pc = f; break; }
  }
  uint64_t pos = uint64_t(rec.pos);
  std::vector<Variant> out;
  const std::string& refallele = rec.ref;
  for (const std::string& a : rec.alts) {
    if (a.size() == 1 && refallele.size() > 1) {
      Variant v{Variant::Deletion, pos};
      v.len = refallele.size() - 1;
      v.is_germline_ = is_germline; v.prot_change_ = pc;
      out.push_back(v);
    } else if (a.size() > 1 && refallele.size() == 1) {
      if (a[0] == '<') {
        if (a == "<DEL>") {
          std::string err;
          uint64_t l = 0;
          bool ok = false;
          if (!vcf.svlen_defined) {
            err = "Encountered rust_htslib error when trying to access 'SVLEN' tag for '<DEL>' alternative allele on contig " +
                  std::to_string(rec.rid) + " at position " + std::to_string(pos);
          } else if (!rec.has_svlen) {
            err = "Found no 'SVLEN' info tag for <DEL> alternative allele at chr " + std::to_string(rec.rid) + " pos " + std::to_string(pos);
          } else if (rec.svlen.size() > 1) {
            err = "microphaser does not handle multiallelic records. Please normalize, e.g. with `bcftools norm -m-`.";
          } else if (rec.svlen[0] == INT64_MIN) {
            err = "Found no 'SVLEN' info tag for <DEL> alternative allele on contig " + std::to_string(rec.rid) + " at pos " + std::to_string(pos);
          } else {
            l = uint64_t(rec.svlen[0] < 0 ? -rec.svlen[0] : rec.svlen[0]);
            ok = true;
          }
          if (ok) {
            Variant v{Variant::Deletion, pos};
            v.len = l; v.is_germline_ = is_germline; v.prot_change_ = pc;
            out.push_back(v);
          } else {
            warn_or_error(err);
          }
        } else {
          warn_or_error("Alternative allele type '" + a + "' not yet supported, but found on contig " + std::to_string(rec.rid) +
                        " at position " + std::to_string(pos) + ".");
        }
      } else {
        Variant v{Variant::Insertion, pos};
        v.seq.assign(a.begin(), a.end());
        v.len = a.size() - 1;
        v.is_germline_ = is_germline; v.prot_change_ = pc;
        out.push_back(v);
      }
    } else if (a.size() == 1 && refallele.size() == 1) {
      Variant v{Variant::SNV, pos};
      v.alt = uint8_t(a[0]);
      v.is_germline_ = is_germline; v.prot_change_ = pc;
      out.push_back(v);
    } else {
      log_warn("Unsupported variant " + refallele + " -> " + a);
    }
  }
  return out;
}

// ------------------------------------------------------------------ Gene model (common.rs:224-348)
struct Interval {
  uint64_t start, end, frame;
  static Interval make(uint64_t s, uint64_t e, const std::string& frame) {
    uint64_t f = 0;
    if (frame != ".") {
      size_t used = 0;
      try { f = std::stoull(frame, &used); } catch (...) { used = 0; }
      if (used != frame.size() || frame.empty()) throw Panic("called `Result::unwrap()` on an `Err` value: ParseIntError (GTF frame)");
    }
    return Interval{s, e, f};
  }
};
enum class Strand { Forward, Reverse };
struct Transcript {
  std::string id, biotype;
  Strand strand;
  std::vector<Interval> exons;
  bool is_coding() const { return !exons.empty(); }
};
struct Gene {
  std::string id, name, chrom, biotype;
  Interval interval;
  std::vector<Transcript> transcripts;
  uint64_t start() const { return interval.start; }
  uint64_t end() const { return interval.end; }
};

// ------------------------------------------------------------------ writers
struct FastaWriter {  // bio::io::fasta::Writer::write(id, None, seq)
  FILE* f;
  void write(const std::string& id, const uint8_t* seq, size_t n) {
    fputc('>', f);
    fwrite(id.data(), 1, id.size(), f);
    fputc('\n', f);
    fwrite(seq, 1, n, f);
    fputc('\n', f);
  }
};

struct TsvWriter {  // csv::WriterBuilder::new().delimiter(b'\t') + serde header-on-first-row
  FILE* f;
  bool has_headers = true;
  bool header_written = false;
  void row(const std::vector<std::string>& header, const std::vector<std::string>& fields) {
    if (has_headers && !header_written) {
      header_written = true;
      emit(header);
    }
    emit(fields);
  }
  void emit(const std::vector<std::string>& fields) {
    std::string line;
    if (fields.size() == 1 && fields[0].empty()) line = "\"\"";
    for (size_t i = 0; i < fields.size() && !(fields.size() == 1 && fields[0].empty()); ++i) {
      if (i) line.push_back('\t');
      mphfmt::csv_field(fields[i], '\t', line);
    }
    line.push_back('\n');
    fwrite(line.data(), 1, line.size(), f);
  }
};

// ------------------------------------------------------------------ IDRecord (common.rs:350-569)
struct IDRecord {
  std::string id, transcript, gene_id, gene_name, chrom;
  uint64_t offset = 0, frame = 0;
  double freq = 0;
  uint32_t depth = 0, nvar = 0, nsomatic = 0, nvariant_sites = 0, nsomvariant_sites = 0;
  std::string strand, variant_sites, somatic_positions, somatic_aa_change, germline_positions, germline_aa_change,
      normal_sequence, mutant_sequence;

  static const std::vector<std::string>& header() {
    static const std::vector<std::string> h = {
        "id", "transcript", "gene_id", "gene_name", "chrom", "offset", "frame", "freq", "depth", "nvar", "nsomatic",
        "nvariant_sites", "nsomvariant_sites", "strand", "variant_sites", "somatic_positions", "somatic_aa_change",
        "germline_positions", "germline_aa_change", "normal_sequence", "mutant_sequence"};
    return h;
  }
  std::vector<std::string> fields() const {
    return {id, transcript, gene_id, gene_name, chrom, std::to_string(offset), std::to_string(frame),
            mphfmt::format_f64(freq), std::to_string(depth), std::to_string(nvar), std::to_string(nsomatic),
            std::to_string(nvariant_sites), std::to_string(nsomvariant_sites), strand, variant_sites, somatic_positions,
            somatic_aa_change, germline_positions, germline_aa_change, normal_sequence, mutant_sequence};
  }

  static uint64_t parse_u64(const std::string& p) {
    if (p.empty()) throw Panic("ParseIntError: empty");
    uint64_t v = 0;
    for (char c : p) {
      if (c < '0' || c > '9') throw Panic("ParseIntError: invalid digit");
      v = v * 10 + uint64_t(c - '0');
    }
    return v;
  }
  static const std::string& at(const std::vector<std::string>& v, size_t i) {
    if (i >= v.size()) throw Panic("index out of bounds");
    return v[i];
  }
  static std::string join(const std::vector<std::string>& v) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) {
      if (i) s.push_back('|');
      s += v[i];
    }
    return s;
  }

  // IDRecord::update — common.rs:376-526
  IDRecord update(const IDRecord& rec, uint64_t offset_, uint64_t frame_, double freq_, const Bytes& wt_seq,
                  const Bytes& mt_seq, uint64_t wlen) const {
    std::string fasta_id = mphfmt::record_id(mt_seq.data(), mt_seq.size(), transcript, offset_, strand.empty() ? '?' : strand[0]);
    auto somatic_positions_v = mphio::split(somatic_positions, '|');
    auto somatic_aa = mphio::split(somatic_aa_change, '|');
    auto other_somatic_aa = mphio::split(rec.somatic_aa_change, '|');
    auto germline_positions_v = mphio::split(germline_positions, '|');
    auto germline_aa = mphio::split(germline_aa_change, '|');
    auto other_germline_aa = mphio::split(rec.germline_aa_change, '|');
    std::vector<std::string> s_p, g_p, s_aa, g_aa;
    uint32_t nvariants = 0, nsom = 0;
    size_t c = 0;
    const uint64_t window_len = wlen;
    const bool fwd = strand == "Forward";
    for (auto& p : somatic_positions_v) {
      if (p.empty()) break;
      uint64_t pv = parse_u64(p);
      bool active = fwd ? (offset + offset_ <= pv) : (offset + window_len - offset_ >= pv);
      if (active) { s_p.push_back(p); s_aa.push_back(at(somatic_aa, c)); nsom += 1; nvariants += 1; }
      c += 1;
    }
    c = 0;
    for (auto& p : mphio::split(rec.somatic_positions, '|')) {
      if (p.empty()) break;
      uint64_t pv = parse_u64(p);
      bool active = fwd ? (rec.offset + offset_ >= pv) : (rec.offset + window_len - 3 - offset_ <= pv);
      if (active) { s_p.push_back(p); s_aa.push_back(at(other_somatic_aa, c)); nsom += 1; nvariants += 1; }
      c += 1;
    }
    c = 0;
    for (auto& p : germline_positions_v) {
      if (p.empty()) break;
      if (offset + offset_ <= parse_u64(p)) { g_p.push_back(p); g_aa.push_back(at(germline_aa, c)); nvariants += 1; }
      c += 1;
    }
    c = 0;
    for (auto& p : mphio::split(rec.germline_positions, '|')) {
      if (p.empty()) break;
      if (rec.offset >= parse_u64(p) - offset_) { g_p.push_back(p); g_aa.push_back(at(other_germline_aa, c)); nvariants += 1; }
      c += 1;
    }
    uint64_t new_offset = fwd ? offset + offset_ : rec.offset + window_len + 3 - offset_;
    uint32_t new_depth = (rec.depth == 0 || depth == 0) ? 0 : (rec.depth + depth) / 2;
    std::string vr = variant_sites + "|" + rec.variant_sites;
    if (!vr.empty() && vr.front() == '|') vr = vr.substr(1);
    if (!vr.empty() && vr.back() == '|') vr.pop_back();
    IDRecord o;
    o.id = fasta_id; o.transcript = transcript; o.gene_id = gene_id; o.gene_name = gene_name; o.chrom = chrom;
    o.offset = new_offset; o.frame = frame_; o.freq = freq_; o.depth = new_depth; o.nvar = nvariants; o.nsomatic = nsom;
    o.nvariant_sites = nvariant_sites + rec.nvariant_sites;
    o.nsomvariant_sites = nsomvariant_sites + rec.nsomvariant_sites;
    o.strand = strand; o.variant_sites = vr;
    o.somatic_positions = join(s_p); o.somatic_aa_change = join(s_aa);
    o.germline_positions = join(g_p); o.germline_aa_change = join(g_aa);
    o.normal_sequence.assign(wt_seq.begin(), wt_seq.end());
    o.mutant_sequence.assign(mt_seq.begin(), mt_seq.end());
    return o;
  }

  // IDRecord::add_freq — common.rs:528-568
  IDRecord add_freq(double f) const {
    IDRecord o = *this;
    uint32_t new_nvar = nvar == 0 ? nvar : (f > 0.0 ? nvar - 1 : nvar);
    uint32_t new_somatic = new_nvar < nsomatic ? nsomatic - 1 : nsomatic;
    o.freq = freq > 0.5 ? freq : freq + f;
    o.nvar = new_nvar;
    o.nsomatic = new_somatic;
    return o;
  }
};

// ------------------------------------------------------------------ I/O buffers
using ReadPtr = std::shared_ptr<const mphio::BamRecord>;

// bam::RecordBuffer::new(reader, false) + fetch(chrom, start, end): keeps the mapped records of `chrom` with
// pos < end in file order; one look-ahead "overflow" record is carried into the next fetch; records right of `end`
// left over from an earlier, wider fetch stay buffered. After an indexed re-fetch the reader yields every record that
// OVERLAPS `start` (htslib's iterator returns records with pos < region end and end_pos > region start), and the
// buffer's loop has no start test of its own - only bcf::buffer has one, its reader being un-indexed - so reads that
// begin left of the gene and reach into it are in the buffer; without a re-fetch, records left of `start` are dropped
// from the front.
struct BamRecordBuffer {
  std::vector<std::vector<ReadPtr>> by_tid;  // mapped+unmapped records in file order, per tid
  const mphio::BamFile* bam = nullptr;
  std::deque<ReadPtr> inner;
  ReadPtr overflow;
  int cur_tid = -1;
  size_t cursor = 0;  // next record index in by_tid[cur_tid]

  void load(mphio::BamFile& b) {
    bam = &b;
    by_tid.assign(b.ref_names.size(), {});
    mphio::BamRecord r;
    while (b.next(r)) {
      if (r.tid < 0 || size_t(r.tid) >= by_tid.size()) continue;
      by_tid[r.tid].push_back(std::make_shared<mphio::BamRecord>(r));
    }
  }
  void fetch(const std::string& chrom, uint64_t start, uint64_t end) {
    if (overflow) { inner.push_back(overflow); overflow.reset(); }
    auto it = bam->tid_of.find(chrom);
    if (it == bam->tid_of.end()) throw Failure("sequence " + chrom + " not found in BAM header");
    int tid = it->second;
    const auto& v = by_tid[tid];
    bool refetched = false;
    if (inner.empty() || uint64_t(inner.back()->pos) < start || inner.front()->tid != tid || uint64_t(inner.front()->pos) > start) {
      // indexed re-fetch: position the reader at the first record that can overlap `start`
      refetched = true;
      inner.clear();
      cur_tid = tid;
      cursor = 0;
      while (cursor < v.size() && (v[cursor]->is_unmapped() || uint64_t(v[cursor]->end_pos()) <= start) &&
             uint64_t(v[cursor]->pos) < start)
        ++cursor;
    } else {
      while (!inner.empty() && uint64_t(inner.front()->pos) < start) inner.pop_front();
    }
    while (cur_tid == tid && cursor < v.size()) {
      const ReadPtr& r = v[cursor++];
      if (r->is_unmapped()) continue;
      uint64_t pos = uint64_t(r->pos);
      if (pos >= end) { overflow = r; break; }
      if (pos >= start || (refetched && uint64_t(r->end_pos()) > start)) inner.push_back(r);
    }
  }
};

// bcf::Reader (streaming, un-indexed) + bcf::buffer::RecordBuffer::fetch(chrom, start, end)
struct VcfRecordBuffer {
  mphio::VcfFile* vcf = nullptr;
  std::deque<mphio::VcfRecord> ring, ring2;
  bool have_overflow = false;
  mphio::VcfRecord overflow;

  void drain_left(int rid, uint64_t start) {
    while (!ring.empty() && ring.front().rid == rid && uint64_t(ring.front().pos) < start) ring.pop_front();
  }
  void fetch(const std::string& chrom, uint64_t start, uint64_t end) {
    int rid = vcf->name2rid(chrom);
    if (rid < 0) throw Failure("contig " + chrom + " not found in VCF header");
    bool has_last = !ring.empty(), has_next = !ring2.empty();
    if (has_last) {
      if (ring.back().rid != rid) { ring.swap(ring2); ring2.clear(); }
      else drain_left(rid, start);
    } else if (has_next) {
      ring.swap(ring2); ring2.clear();
      drain_left(rid, start);
    }
    if (!ring2.empty()) return;
    if (have_overflow) {
      uint64_t pos = uint64_t(overflow.pos);
      if (pos >= start) {
        if (pos <= end) { ring.push_back(overflow); have_overflow = false; }
        else return;
      } else {
        have_overflow = false;
      }
    }
    mphio::VcfRecord rec;
    while (vcf->next(rec)) {
      uint64_t pos = uint64_t(rec.pos);
      if (rec.rid == rid) {
        if (pos >= end) { overflow = rec; have_overflow = true; break; }
        else if (pos >= start) ring.push_back(rec);
      } else if (rec.rid > rid) {
        ring2.push_back(rec);
        break;
      }
    }
  }
};

}  // namespace oracle
