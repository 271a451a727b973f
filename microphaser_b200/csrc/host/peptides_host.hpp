// peptides_host.hpp — host side of `filter` and `build_reference` (reference src/peptides.rs).
// Translation and the normal-peptidome membership test run on the device (kernels/peptide_kernels.cu)
// through the two callables below; what stays here is the row grouping (:246-254,292-402,534-571),
// the maximum-likelihood frequency and credible interval (:188-218,405-481,594-664; tiny f64 work)
// and the writers. statrs 0.15.0 Binomial::pmf and bio 0.34 LogProb Simpson integration are restated
// from their published behaviour (crates not vendored under /root/reference).
#pragma once
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "../io/fmt_util.hpp"
#include "../io/hts_io.hpp"
#include "batch.hpp"

namespace mph {
namespace pep {

// translate(sequences, frames) -> peptides ('?' + bad flag for an unknown codon); device call
using TranslateFn = std::function<void(const std::vector<std::string>& nt, const std::vector<int8_t>& frame, std::vector<std::string>& aa,
                                       std::vector<uint8_t>& bad)>;
// probe(queries of equal length k) -> hit flags against the loaded set; device call
using ProbeFn = std::function<void(const std::vector<std::string>& queries, uint32_t k, std::vector<uint8_t>& hit)>;

struct InfoRow {  // src/common.rs:350-373
  std::string id, transcript, gene_id, gene_name, chrom;
  uint64_t offset = 0, frame = 0;
  double freq = 0;
  uint32_t depth = 0, nvar = 0, nsomatic = 0, nvariant_sites = 0, nsomvariant_sites = 0;
  std::string strand, variant_sites, somatic_positions, somatic_aa_change, germline_positions, germline_aa_change, normal_sequence, mutant_sequence;
};

inline uint64_t to_u64(const std::string& p) {
  if (p.empty()) throw std::runtime_error("CSV deserialize error: cannot parse integer from empty string");
  uint64_t v = 0;
  for (char c : p) {
    if (c < '0' || c > '9') throw std::runtime_error("CSV deserialize error: invalid digit found in string");
    v = v * 10 + uint64_t(c - '0');
  }
  return v;
}
inline double to_f64(const std::string& s) {
  if (s == "NaN") return NAN;
  if (s == "inf") return INFINITY;
  if (s == "-inf") return -INFINITY;
  size_t used = 0;
  double v = 0;
  try { v = std::stod(s, &used); } catch (...) { used = 0; }
  if (used != s.size() || s.empty()) throw std::runtime_error("CSV deserialize error: invalid float literal");
  return v;
}

// csv::Reader with '\t' delimiter (quoted fields, doubled quotes)
inline std::vector<std::vector<std::string>> read_tsv(const std::string& path) {
  std::vector<uint8_t> d = mphio::read_file(path);
  std::vector<std::vector<std::string>> rows;
  std::vector<std::string> cur;
  std::string f;
  bool inq = false, any = false;
  for (size_t i = 0; i < d.size(); ++i) {
    const char c = char(d[i]);
    if (inq) {
      if (c == '"') {
        if (i + 1 < d.size() && d[i + 1] == '"') { f.push_back('"'); ++i; }
        else inq = false;
      } else f.push_back(c);
      continue;
    }
    if (c == '"' && f.empty()) { inq = true; any = true; }
    else if (c == '\t') { cur.push_back(f); f.clear(); any = true; }
    else if (c == '\n' || c == '\r') {
      if (c == '\r' && i + 1 < d.size() && d[i + 1] == '\n') ++i;
      if (any || !f.empty()) { cur.push_back(f); rows.push_back(cur); }
      cur.clear(); f.clear(); any = false;
    } else { f.push_back(c); any = true; }
  }
  if (any || !f.empty()) { cur.push_back(f); rows.push_back(cur); }
  return rows;
}

// ---- numerics ---------------------------------------------------------------------------------
inline double ln_factorial(uint64_t x) {  // statrs: table of 171 factorials, ln_gamma beyond
  static const std::vector<double> cache = [] {
    std::vector<double> c(171);
    double f = 1.0;
    c[0] = 1.0;
    for (int i = 1; i <= 170; ++i) { f *= double(i); c[size_t(i)] = f; }
    return c;
  }();
  return x <= 170 ? std::log(cache[size_t(x)]) : std::lgamma(double(x) + 1.0);
}
inline double binomial_pmf(double p, uint64_t n, uint64_t x) {
  if (x > n) return 0.0;
  if (p == 0.0) return x == 0 ? 1.0 : 0.0;
  if (std::fabs(p - 1.0) <= 4 * std::numeric_limits<double>::epsilon()) return x == n ? 1.0 : 0.0;
  return std::exp(ln_factorial(n) - ln_factorial(x) - ln_factorial(n - x) + double(x) * std::log(p) + double(n - x) * std::log(1.0 - p));
}
inline double density(const std::vector<double>& alt, const std::vector<uint32_t>& depth, double theta) {  // :188-201
  double prob = 1.0;
  for (size_t i = 0; i < alt.size(); ++i) {
    if (!(theta >= 0.0 && theta <= 1.0)) throw Fatal("called `Result::unwrap()` on an `Err` value: BadParams");
    const double a = std::round(alt[i]);
    const uint64_t k = a <= 0 ? 0 : (a >= 1.8446744073709552e19 ? UINT64_MAX : uint64_t(a));
    prob *= binomial_pmf(theta, depth[i], k);
  }
  return prob;
}
inline double ln_sum_exp(const std::vector<double>& probs) {
  if (probs.empty()) return -INFINITY;
  double pmax = probs[0];
  size_t imax = 0;
  for (size_t i = 1; i < probs.size(); ++i)
    if (probs[i] > pmax) { pmax = probs[i]; imax = i; }
  if (pmax == -INFINITY) return -INFINITY;
  if (pmax == INFINITY) return INFINITY;
  double s = 0.0;
  for (size_t i = 0; i < probs.size(); ++i)
    if (i != imax) s += std::exp(probs[i] - pmax);
  return pmax + std::log1p(s);
}
template <class D>
inline double ln_simpson(D fn, double a, double b, size_t n) {
  std::vector<double> probs;
  const double step = (b - a) / double(n - 1);
  for (size_t i = 1; i + 1 < n; ++i) probs.push_back(fn(a + double(i) * step) + std::log(double(2 + (i % 2) * 2)));
  probs.push_back(fn(a));
  probs.push_back(fn(b));
  return ln_sum_exp(probs) + std::log(b - a) - std::log(double(n - 1)) - std::log(3.0);
}

struct Interval {
  uint64_t ml;
  double a, b;
};
// ML on a 101-point grid + 95% credible interval; `final_pass` = the search variant of the final flush (:594-664)
inline Interval credible_interval(const std::vector<double>& fr, const std::vector<uint32_t>& dp, bool final_pass) {
  uint64_t ml = 0;
  double best = 0;
  for (uint64_t t = 0; t < 101; ++t) {
    const double prob = density(fr, dp, double(t) * 0.01);
    if (std::isnan(prob) || (t > 0 && std::isnan(best))) throw Fatal("called `Option::unwrap()` on a `None` value (partial_cmp)");
    if (t == 0 || !(prob < best)) { best = prob; ml = t; }
  }
  const double r = ln_simpson([&](double v) { return std::log(density(fr, dp, v)); }, 0.0, 1.0, 99);
  double a = ml < 10 ? 0.0 : double(ml - 10) * 0.01, b = ml > 90 ? 1.0 : double(ml + 10) * 0.01;
  double p = std::log(0.0);
  const double l95 = std::log(0.95), l96 = std::log(0.96);
  auto integrate = [&] { return ln_simpson([&](double v) { return std::log(density(fr, dp, v)) - r; }, a, b, 11); };
  if (!final_pass) {
    double a_old = double(ml) * 0.01, b_old = double(ml) * 0.01;
    for (int counter = 0; counter < 50; ++counter) {
      if (p < l95) { a_old = a; a = a < 0.1 ? 0.0 : (a - 0.1); b_old = b; b = b > 0.9 ? 1.0 : (b + 0.1); }
      if (p > l96) { a += (a_old - a) / 2.0; b -= (b - b_old) / 2.0; }
      p = integrate();
      if (p >= l95 && p < l96) break;
    }
  } else {
    double a_r = double(ml) * 0.01, a_l = 0.0, b_r = 1.0, b_l = double(ml) * 0.01;
    for (int counter = 0; counter < 10; ++counter) {
      if (p < l95) { a_r = a; a = a < 0.1 ? 0.0 : a - ((a - a_l) / 2.0); b_l = b; b = b > 0.9 ? 1.0 : b + ((b_r - b) / 2.0); }
      if (p > l96) { a_l = a; a += (a_r - a) / 2.0; b_r = b; b -= (b - b_l) / 2.0; }
      p = integrate();
      if (p >= l95 && p < l96) break;
    }
  }
  return Interval{ml, a, b};
}

inline void put_fasta(FILE* f, const std::string& id, const std::string& seq) {
  fputc('>', f);
  fwrite(id.data(), 1, id.size(), f);
  fputc('\n', f);
  fwrite(seq.data(), 1, seq.size(), f);
  fputc('\n', f);
}
inline void put_row(FILE* f, const std::vector<std::string>& fields) {
  std::string line;
  for (size_t i = 0; i < fields.size(); ++i) {
    if (i) line.push_back('\t');
    mphfmt::csv_field(fields[i], '\t', line);
  }
  line.push_back('\n');
  fwrite(line.data(), 1, line.size(), f);
}

// bincode 1: HashSet<Vec<u8>> = u64 LE count, then u64 LE length + bytes per item
inline std::vector<std::string> read_peptide_set(const std::string& path) {
  std::vector<uint8_t> d = mphio::read_file(path);
  auto u64 = [&](size_t o) {
    if (o + 8 > d.size()) throw Fatal("called `Result::unwrap()` on an `Err` value: Io(UnexpectedEof)");
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= uint64_t(d[o + size_t(i)]) << (8 * i);
    return v;
  };
  std::vector<std::string> items;
  size_t o = 0;
  const uint64_t n = u64(o);
  o += 8;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t l = u64(o);
    o += 8;
    if (o + l > d.size()) throw Fatal("called `Result::unwrap()` on an `Err` value: Io(UnexpectedEof)");
    items.emplace_back(reinterpret_cast<const char*>(d.data() + o), size_t(l));
    o += size_t(l);
  }
  return items;
}

struct FilterPaths {
  std::string tsv_in, normal_out, tsv_out, removed_tsv, removed_fasta;
  FILE* fasta_out = stdout;
};

// `filter` (:234-709). The reference interleaves grouping, the membership test and writing; here the grouping
// runs first and records every output entry in order, the membership of all entries is resolved by ONE device
// probe, then the entries are written in the recorded order.
inline void run_filter(const FilterPaths& io, size_t peptide_length, const TranslateFn& translate, const ProbeFn& probe) {
  auto open_out = [](const std::string& p) {
    FILE* f = fopen(p.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + p);
    return f;
  };
  FILE* tsv_f = open_out(io.tsv_out);
  FILE* removed_f = open_out(io.removed_tsv);
  FILE* removed_fa = open_out(io.removed_fasta);
  FILE* normal_f = open_out(io.normal_out);
  static const std::vector<std::string> header = {
      "id", "transcript", "gene_id", "gene_name", "chrom", "offset", "frame", "freq", "credible_interval", "depth", "nvar", "nsomatic",
      "nvariant_sites", "nsomvariant_sites", "strand", "variant_sites", "somatic_positions", "somatic_aa_change", "germline_positions",
      "germline_aa_change", "normal_sequence", "mutant_sequence", "normal_peptide", "tumor_peptide"};
  put_row(tsv_f, header);  // always written (:255-258)

  // ---- rows + device translation of every mutant / normal window
  auto raw = read_tsv(io.tsv_in);
  std::vector<InfoRow> rows;
  for (size_t ri = 1; ri < raw.size(); ++ri) {
    const auto& c = raw[ri];
    if (c.size() != 21) throw std::runtime_error("CSV error: record has " + std::to_string(c.size()) + " fields, expected 21");
    InfoRow r;
    r.id = c[0]; r.transcript = c[1]; r.gene_id = c[2]; r.gene_name = c[3]; r.chrom = c[4];
    r.offset = to_u64(c[5]); r.frame = to_u64(c[6]); r.freq = to_f64(c[7]); r.depth = uint32_t(to_u64(c[8])); r.nvar = uint32_t(to_u64(c[9]));
    r.nsomatic = uint32_t(to_u64(c[10])); r.nvariant_sites = uint32_t(to_u64(c[11])); r.nsomvariant_sites = uint32_t(to_u64(c[12]));
    r.strand = c[13]; r.variant_sites = c[14]; r.somatic_positions = c[15]; r.somatic_aa_change = c[16]; r.germline_positions = c[17];
    r.germline_aa_change = c[18]; r.normal_sequence = c[19]; r.mutant_sequence = c[20];
    rows.push_back(std::move(r));
  }
  std::vector<std::string> nt;
  std::vector<int8_t> frames;
  for (auto& r : rows) {
    const int8_t fr = (!r.id.empty() && r.id.back() == 'F') ? 1 : -1;
    nt.push_back(r.mutant_sequence); frames.push_back(fr);
    nt.push_back(r.normal_sequence); frames.push_back(fr);
  }
  std::vector<std::string> aa;
  std::vector<uint8_t> bad;
  translate(nt, frames, aa, bad);

  // ---- grouping; every output entry is recorded with its region's ML / interval
  struct Entry {
    InfoRow row;
    std::string tumor_p, normal_p, ci;
  };
  std::vector<Entry> entries;  // in output order
  using Key = std::tuple<uint64_t, std::string, std::string>;
  struct Pending {
    InfoRow row;
    std::string tumor_p, normal_p;
  };
  std::map<Key, std::vector<double>> frequencies;
  std::map<Key, std::vector<uint32_t>> depth;
  std::map<Key, std::vector<Pending>> records;
  auto flush = [&](bool final_pass) {
    for (auto& kv : records) {
      const Interval iv = credible_interval(frequencies.at(kv.first), depth.at(kv.first), final_pass);
      char ci[64];
      snprintf(ci, sizeof ci, "%.2f-%.2f", iv.a, iv.b);
      for (auto& p : kv.second) {
        Entry e{p.row, p.tumor_p, p.normal_p, ci};
        e.row.freq = e.row.depth == 0 ? 0.0 : double(iv.ml) * 0.01;
        entries.push_back(std::move(e));
      }
    }
  };
  std::tuple<std::string, std::string, std::string> current{"", "", ""};
  std::pair<std::string, std::string> region_sites{"", ""};
  std::unordered_set<std::string> seen_peptides;
  std::map<std::pair<std::string, uint64_t>, size_t> stop_gained;
  for (size_t ri = 0; ri < rows.size(); ++ri) {
    const InfoRow& row = rows[ri];
    if (row.mutant_sequence.size() < 2) throw Fatal("attempt to subtract with overflow / slice index (to_protein on a sequence shorter than 2)");
    if (bad[2 * ri] || (!row.normal_sequence.empty() && bad[2 * ri + 1])) throw Fatal("called `Result::unwrap()` on an `Err` value: () (unknown codon)");
    size_t som_pos = 0;
    if (!row.somatic_positions.empty() && row.somatic_positions.find('|') == std::string::npos) {
      for (char ch : row.somatic_positions)
        if (ch < '0' || ch > '9') throw Fatal("called `Result::unwrap()` on an `Err` value: ParseIntError");
      som_pos = size_t(to_u64(row.somatic_positions));
    }
    const size_t offset = size_t(row.offset);
    const std::string& tumor_peptide = aa[2 * ri];
    const std::string normal_peptide = row.normal_sequence.empty() ? std::string() : aa[2 * ri + 1];
    const std::pair<std::string, uint64_t> check{row.transcript, row.frame};
    auto sg = stop_gained.find(check);
    if (sg != stop_gained.end()) {
      const bool downstream = row.strand == "Forward" ? offset > sg->second : (row.strand == "Reverse" ? offset < sg->second : false);
      if (downstream) continue;
    }
    if (tumor_peptide.find('X') != std::string::npos && (std::fabs(row.freq - 1.0) < std::numeric_limits<double>::epsilon() || row.frame > 0))
      stop_gained[check] = offset;
    size_t i = 0;
    while (i + peptide_length <= tumor_peptide.size()) {
      const std::string tumor_pep = tumor_peptide.substr(i, peptide_length);
      if (tumor_pep.find('X') != std::string::npos) break;
      const std::string normal_pep = normal_peptide.size() >= i + peptide_length ? normal_peptide.substr(i, peptide_length) : normal_peptide;
      if (normal_pep.empty() && som_pos > 0) {
        if (row.strand == "Forward") {
          if (((i + peptide_length) * 3) + offset <= som_pos) { i += 1; continue; }
        } else if (row.strand == "Reverse") {
          if ((tumor_peptide.size() - (i + peptide_length)) * 3 + offset > som_pos) { i += 1; continue; }
        }
      }
      i += 1;
      if (tumor_pep == normal_pep) continue;
      const std::pair<std::string, std::string> current_sites{row.transcript, row.variant_sites};
      const std::tuple<std::string, std::string, std::string> cur3{row.transcript, row.somatic_positions, row.germline_positions};
      if (cur3 == current) {
        if (seen_peptides.count(tumor_pep)) continue;
      } else {
        current = cur3;
        seen_peptides.clear();
      }
      seen_peptides.insert(tumor_pep);
      Pending value{row, tumor_pep, normal_pep};
      value.row.id = std::to_string(i) + "_" + row.id;
      const Key key{row.frame, row.somatic_positions, row.germline_positions};
      const double alt = row.freq * double(row.depth);
      if (current_sites != region_sites) {
        flush(false);
        frequencies.clear(); depth.clear(); records.clear();
        frequencies[key] = {alt};
        depth[key] = {row.depth};
        records[key].push_back(std::move(value));
        region_sites = current_sites;
      } else {
        // entry(key).or_insert_with(|| vec![x]).push(x): a key new to the region receives the value twice (:557-568)
        const bool fresh = !records.count(key);
        for (int rep = 0; rep < (fresh ? 2 : 1); ++rep) {
          frequencies[key].push_back(alt);
          depth[key].push_back(row.depth);
          records[key].push_back(value);
        }
      }
    }
  }
  flush(true);

  // ---- ONE device probe for all entries, then the writes in order
  std::vector<std::string> queries;
  for (auto& e : entries) queries.push_back(e.tumor_p);
  std::vector<uint8_t> hit;
  probe(queries, uint32_t(peptide_length), hit);
  bool removed_header = false;
  for (size_t x = 0; x < entries.size(); ++x) {
    const Entry& e = entries[x];
    const InfoRow& r = e.row;
    const std::vector<std::string> fields = {r.id, r.transcript, r.gene_id, r.gene_name, r.chrom, std::to_string(r.offset), std::to_string(r.frame),
                                             mphfmt::format_f64(r.freq), e.ci, std::to_string(r.depth), std::to_string(r.nvar),
                                             std::to_string(r.nsomatic), std::to_string(r.nvariant_sites), std::to_string(r.nsomvariant_sites),
                                             r.strand, r.variant_sites, r.somatic_positions, r.somatic_aa_change, r.germline_positions,
                                             r.germline_aa_change, r.normal_sequence, r.mutant_sequence, e.normal_p, e.tumor_p};
    if (hit[x]) {
      put_fasta(removed_fa, r.id, e.tumor_p);
      if (!removed_header) { put_row(removed_f, header); removed_header = true; }
      put_row(removed_f, fields);
    } else {
      put_fasta(io.fasta_out, r.id, e.tumor_p);
      if (!e.normal_p.empty()) put_fasta(normal_f, r.id, e.normal_p);
      put_row(tsv_f, fields);
    }
  }
  fclose(tsv_f); fclose(removed_f); fclose(removed_fa); fclose(normal_f);
  fflush(io.fasta_out);
}

// bio::io::fasta::Reader: id = header up to the first whitespace, sequence lines concatenated
inline std::vector<std::pair<std::string, std::string>> read_fasta_records(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open " + path);
  std::vector<std::pair<std::string, std::string>> out;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty() && line[0] == '>') {
      size_t e = line.find_first_of(" \t", 1);
      out.emplace_back(line.substr(1, e == std::string::npos ? std::string::npos : e - 1), std::string());
    } else if (!out.empty()) {
      out.back().second += line;
    }
  }
  return out;
}

// `build_reference` (:148-186): every codon-step window of every record is translated on the device; the distinct
// peptides come back from the device hash set (`dedupe`) for the bincode file, whose item order is arbitrary.
using DedupeFn = std::function<std::vector<std::string>(const std::vector<std::string>& peptides, uint32_t k)>;
inline void run_build_reference(const std::string& reference_fasta, const std::string& binary_out, FILE* fasta_out, size_t peptide_length,
                                const TranslateFn& translate, const DedupeFn& dedupe) {
  FILE* bin = fopen(binary_out.c_str(), "wb");
  if (!bin) throw std::runtime_error("cannot create " + binary_out);
  auto recs = read_fasta_records(reference_fasta);
  std::vector<std::string> nt, ids;
  std::vector<int8_t> frames;
  const size_t base_length = peptide_length * 3;
  for (auto& r : recs) {
    const int8_t fr = (!r.first.empty() && r.first.back() == 'F') ? 1 : -1;
    for (size_t i = 0; i + base_length <= r.second.size(); i += 3) {
      nt.push_back(r.second.substr(i, base_length));
      frames.push_back(fr);
      ids.push_back(r.first);
    }
  }
  std::vector<std::string> aa;
  std::vector<uint8_t> bad;
  translate(nt, frames, aa, bad);
  for (size_t x = 0; x < aa.size(); ++x) {
    if (bad[x]) throw Fatal("called `Result::unwrap()` on an `Err` value: () (unknown codon)");
    put_fasta(fasta_out, ids[x], aa[x]);
  }
  fflush(fasta_out);
  std::vector<std::string> distinct = dedupe(aa, uint32_t(peptide_length));
  auto w64 = [&](uint64_t v) {
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = uint8_t(v >> (8 * i));
    fwrite(b, 1, 8, bin);
  };
  w64(distinct.size());
  for (auto& p : distinct) {
    w64(p.size());
    fwrite(p.data(), 1, p.size(), bin);
  }
  fclose(bin);
}

}  // namespace pep
}  // namespace mph
