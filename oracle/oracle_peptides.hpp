// oracle_peptides.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's `build_reference` and `filter` sub-commands,
// reference src/peptides.rs (line numbers refer to that file): make_pairs / to_aminoacid /
// to_protein :85-146, build :148-186, density / prob_func :188-218, filter :234-709.
// Third-party arithmetic restated from published behaviour (crates are not vendored under
// /root/reference): statrs 0.15.0 `Binomial::pmf` (ln_binomial via a 171-entry factorial table,
// ln_gamma beyond), bio 0.34 `LogProb::{ln_simpsons_integrate_exp, ln_sum_exp}`, bincode 1 layout
// of HashSet<Vec<u8>> (u64 LE count, then per item u64 LE length + bytes), csv 1 reader/writer.
// Pinned by the reference fixtures test_build, test_filter, test_filter_long, test_filter_fs
// (tests/golden/test_*). The `ref_set.contains` *hit* path is not exercised by any reference
// fixture (their sets hold 4-mers while -l 9 probes 9-mers): parity unpinned, covered here by
// synthetic cases only.
#pragma once
#include <cmath>
#include <fstream>
#include <set>
#include <unordered_set>

#include "oracle_common.hpp"

namespace oracle {
namespace peptides {

// :85-126 — 'X' marks a stop codon; an unknown codon is an Err that the callers unwrap (panic)
inline char to_aminoacid(const char* c) {
  static const char* table[][7] = {
      {"I", "ATT", "ATC", "ATA"}, {"L", "CTT", "CTC", "CTA", "CTG", "TTA", "TTG"}, {"V", "GTT", "GTC", "GTA", "GTG"}, {"F", "TTT", "TTC"},
      {"M", "ATG"}, {"C", "TGT", "TGC"}, {"A", "GCT", "GCC", "GCA", "GCG"}, {"G", "GGT", "GGC", "GGA", "GGG"}, {"P", "CCT", "CCC", "CCA", "CCG"},
      {"T", "ACT", "ACC", "ACA", "ACG"}, {"S", "TCT", "TCC", "TCA", "TCG", "AGT", "AGC"}, {"Y", "TAT", "TAC"}, {"W", "TGG"}, {"Q", "CAA", "CAG"},
      {"N", "AAT", "AAC"}, {"H", "CAT", "CAC"}, {"E", "GAA", "GAG"}, {"D", "GAT", "GAC"}, {"K", "AAA", "AAG"},
      {"R", "CGT", "CGC", "CGA", "CGG", "AGA", "AGG"}, {"X", "TAA", "TAG", "TGA"}};
  for (auto& row : table)
    for (int i = 1; i < 7 && row[i]; ++i)
      if (row[i][0] == c[0] && row[i][1] == c[1] && row[i][2] == c[2]) return row[0][0];
  return 0;
}

// bio::alphabets::dna::revcomp (IUPAC aware; only the letters that can reach a codon table hit matter)
inline char complement(char c) {
  switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    case 'N': return 'N'; case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W';
    case 'K': return 'M'; case 'M': return 'K'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    default: return c;
  }
}

// :128-146
inline std::string to_protein(const std::string& s, int frame) {
  std::string r = s;
  for (auto& c : r)
    if (c >= 'a' && c <= 'z') c = char(c - 32);
  if (frame < 0) {
    std::string rc(r.rbegin(), r.rend());
    for (auto& c : rc) c = complement(c);
    r = rc;
    frame = -frame;
  }
  std::string p;
  size_t i = size_t(frame) - 1;
  const size_t lim = r.size() - 2;  // usize arithmetic: wraps for len < 2
  while (i < lim) {
    if (i + 3 > r.size()) throw Panic("range end index out of range for slice");
    const char aa = to_aminoacid(r.data() + i);
    if (!aa) throw Panic("called `Result::unwrap()` on an `Err` value: () (unknown codon)");
    p.push_back(aa);
    i += 3;
  }
  return p;
}

// bio::io::fasta::Reader — id = header up to the first whitespace, sequence lines concatenated
struct FastaRec {
  std::string id, seq;
};
inline std::vector<FastaRec> read_fasta(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw Failure("cannot open " + path);
  std::vector<FastaRec> out;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty() && line[0] == '>') {
      FastaRec r;
      size_t e = line.find_first_of(" \t", 1);
      r.id = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
      out.push_back(r);
    } else if (!out.empty()) {
      out.back().seq += line;
    }
  }
  return out;
}

inline void write_u64(FILE* f, uint64_t v) {
  uint8_t b[8];
  for (int i = 0; i < 8; ++i) b[i] = uint8_t(v >> (8 * i));
  fwrite(b, 1, 8, f);
}

// :148-186
inline void build(const std::string& reference_fasta, const std::string& binary_out, FILE* fasta_out, size_t peptide_length) {
  std::set<std::string> ref_set;  // written sorted: the reference's HashSet order is arbitrary and never diffed
  FILE* bin = fopen(binary_out.c_str(), "wb");
  if (!bin) throw Failure("cannot create " + binary_out);
  FastaWriter fw{fasta_out};
  for (auto& record : read_fasta(reference_fasta)) {
    const int frame = (!record.id.empty() && record.id.back() == 'F') ? 1 : -1;
    const size_t base_length = peptide_length * 3;
    size_t i = 0;
    while (i + base_length <= record.seq.size()) {
      const std::string pepseq = to_protein(record.seq.substr(i, base_length), frame);
      fw.write(record.id, reinterpret_cast<const uint8_t*>(pepseq.data()), pepseq.size());
      ref_set.insert(pepseq);
      i += 3;
    }
  }
  write_u64(bin, ref_set.size());
  for (auto& p : ref_set) {
    write_u64(bin, p.size());
    fwrite(p.data(), 1, p.size(), bin);
  }
  fclose(bin);
}

inline std::unordered_set<std::string> load_set(const std::string& path) {
  std::vector<uint8_t> d = mphio::read_file(path);
  auto u64 = [&](size_t o) {
    if (o + 8 > d.size()) throw Panic("called `Result::unwrap()` on an `Err` value: Io(UnexpectedEof)");
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= uint64_t(d[o + i]) << (8 * i);
    return v;
  };
  std::unordered_set<std::string> s;
  size_t o = 0;
  const uint64_t n = u64(o);
  o += 8;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t l = u64(o);
    o += 8;
    if (o + l > d.size()) throw Panic("called `Result::unwrap()` on an `Err` value: Io(UnexpectedEof)");
    s.emplace(reinterpret_cast<const char*>(d.data() + o), size_t(l));
    o += l;
  }
  return s;
}

// ---- statrs 0.15.0 -------------------------------------------------------------------------
inline double ln_factorial(uint64_t x) {
  static double cache[171];
  static bool init = false;
  if (!init) {
    double f = 1.0;
    cache[0] = 1.0;
    for (int i = 1; i <= 170; ++i) { f *= double(i); cache[i] = f; }
    init = true;
  }
  if (x <= 170) return std::log(cache[x]);
  return std::lgamma(double(x) + 1.0);
}
inline double binomial_pmf(double p, uint64_t n, uint64_t x) {
  if (x > n) return 0.0;
  if (p == 0.0) return x == 0 ? 1.0 : 0.0;
  if (std::fabs(p - 1.0) <= 4 * std::numeric_limits<double>::epsilon()) return x == n ? 1.0 : 0.0;  // ulps_eq!(p, 1.0)
  const double lb = ln_factorial(n) - ln_factorial(x) - ln_factorial(n - x);
  return std::exp(lb + double(x) * std::log(p) + double(n - x) * std::log(1.0 - p));
}
// :188-201
inline double density(const std::vector<double>& alt, const std::vector<uint32_t>& depth, double theta) {
  double prob = 1.0;
  for (size_t i = 0; i < alt.size(); ++i) {
    if (!(theta >= 0.0 && theta <= 1.0)) throw Panic("called `Result::unwrap()` on an `Err` value: BadParams");
    const double a = std::round(alt[i]);
    const uint64_t k = a <= 0 ? 0 : (a >= 1.8446744073709552e19 ? UINT64_MAX : uint64_t(a));  // `as u64` saturates, NaN -> 0
    prob *= binomial_pmf(theta, depth[i], k);
  }
  return prob;
}

// ---- bio 0.34 LogProb --------------------------------------------------------------------
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
inline double ln_simpsons_integrate_exp(D density_fn, double a, double b, size_t n) {
  std::vector<double> probs;
  const double step = (b - a) / double(n - 1);  // itertools_num::linspace
  for (size_t i = 1; i + 1 < n; ++i) {
    const double v = a + double(i) * step;
    const double weight = double(2 + (i % 2) * 2);
    probs.push_back(density_fn(v) + std::log(weight));
  }
  probs.push_back(density_fn(a));
  probs.push_back(density_fn(b));
  const double width = b - a;
  return ln_sum_exp(probs) + std::log(width) - std::log(double(n - 1)) - std::log(3.0);
}

// ---- csv reader (quotes doubled, fields may be quoted) ---------------------------------
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

inline double parse_f64(const std::string& s) {
  if (s == "NaN") return NAN;
  if (s == "inf") return INFINITY;
  if (s == "-inf") return -INFINITY;
  size_t used = 0;
  double v = 0;
  try { v = std::stod(s, &used); } catch (...) { used = 0; }
  if (used != s.size() || s.empty()) throw Failure("CSV deserialize error: invalid float literal");
  return v;
}

struct FilteredRow {
  IDRecord idr;
  std::string tumor_p, normal_p;
};

// :234-709
inline void filter(const std::string& reference_bin, const std::string& tsv_in, FILE* fasta_out, const std::string& normal_out,
                   const std::string& tsv_out, const std::string& removed_tsv, const std::string& removed_fasta, size_t peptide_length) {
  auto open_out = [](const std::string& p) {
    FILE* f = fopen(p.c_str(), "wb");
    if (!f) throw Failure("cannot create " + p);
    return f;
  };
  // writers are created before anything is read (src/main.rs:188-201)
  TsvWriter tsv_writer{open_out(tsv_out)};
  tsv_writer.has_headers = false;
  TsvWriter removed_writer{open_out(removed_tsv)};
  FastaWriter removed_fasta_writer{open_out(removed_fasta)};
  FastaWriter fasta_writer{fasta_out};
  FastaWriter normal_writer{open_out(normal_out)};
  const std::unordered_set<std::string> ref_set = load_set(reference_bin);
  using Key = std::tuple<uint64_t, std::string, std::string>;
  std::tuple<std::string, std::string, std::string> current{"", "", ""}, current_variant{"", "", ""};
  std::pair<std::string, std::string> region_sites{"", ""};
  std::map<Key, std::vector<double>> frequencies;
  std::map<Key, std::vector<uint32_t>> depth;
  std::map<Key, std::vector<FilteredRow>> records;
  std::unordered_set<std::string> seen_peptides;
  std::map<std::pair<std::string, uint64_t>, size_t> stop_gained;
  static const std::vector<std::string> out_header = {
      "id", "transcript", "gene_id", "gene_name", "chrom", "offset", "frame", "freq", "credible_interval", "depth", "nvar", "nsomatic",
      "nvariant_sites", "nsomvariant_sites", "strand", "variant_sites", "somatic_positions", "somatic_aa_change", "germline_positions",
      "germline_aa_change", "normal_sequence", "mutant_sequence", "normal_peptide", "tumor_peptide"};
  tsv_writer.emit(out_header);

  auto fields_of = [&](const IDRecord& r, const std::string& ci, const std::string& np, const std::string& tp) {
    return std::vector<std::string>{r.id, r.transcript, r.gene_id, r.gene_name, r.chrom, std::to_string(r.offset), std::to_string(r.frame),
                                    mphfmt::format_f64(r.freq), ci, std::to_string(r.depth), std::to_string(r.nvar), std::to_string(r.nsomatic),
                                    std::to_string(r.nvariant_sites), std::to_string(r.nsomvariant_sites), r.strand, r.variant_sites,
                                    r.somatic_positions, r.somatic_aa_change, r.germline_positions, r.germline_aa_change, r.normal_sequence,
                                    r.mutant_sequence, np, tp};
  };
  auto fmt2 = [](double v) {
    char b[64];
    snprintf(b, sizeof b, "%.2f", v);
    return std::string(b);
  };
  // flush of one region: ML frequency, credible interval, membership test, writes. `final_pass` selects the
  // second copy of the search loop (:594-707), which differs from the first (:405-533).
  auto flush = [&](bool final_pass) {
    for (auto& kv : records) {
      const Key& key = kv.first;
      const std::vector<double>& fr = frequencies.at(key);
      const std::vector<uint32_t>& dp = depth.at(key);
      // prob_func + max_by (last maximum wins on ties; partial_cmp().unwrap() panics on NaN)
      uint64_t ml = 0;
      double best = 0;
      for (uint64_t t = 0; t < 101; ++t) {
        const double theta = double(t) * 0.01;
        const double prob = density(fr, dp, theta);
        if (std::isnan(prob) || (t > 0 && std::isnan(best))) throw Panic("called `Option::unwrap()` on a `None` value (partial_cmp)");
        if (t == 0 || !(prob < best)) { best = prob; ml = t; }
      }
      const double r = ln_simpsons_integrate_exp([&](double v) { return std::log(density(fr, dp, v)); }, 0.0, 1.0, 99);
      double a = ml < 10 ? 0.0 : double(ml - 10) * 0.01;
      double b = ml > 90 ? 1.0 : double(ml + 10) * 0.01;
      double p = std::log(0.0);
      const double l95 = std::log(0.95), l96 = std::log(0.96);
      if (!final_pass) {
        double a_old = double(ml) * 0.01, b_old = double(ml) * 0.01;
        int counter = 0;
        for (;;) {
          if (counter == 50) break;
          if (p < l95) {
            a_old = a;
            a = a < 0.1 ? 0.0 : (a - 0.1);
            b_old = b;
            b = b > 0.9 ? 1.0 : (b + 0.1);
          }
          if (p > l96) {
            a += (a_old - a) / 2.0;
            b -= (b - b_old) / 2.0;
          }
          p = ln_simpsons_integrate_exp([&](double v) { return std::log(density(fr, dp, v)) - r; }, a, b, 11);
          if (p >= l95 && p < l96) break;
          counter += 1;
        }
      } else {
        double a_r = double(ml) * 0.01, a_l = 0.0, b_r = 1.0, b_l = double(ml) * 0.01;
        int counter = 0;
        for (;;) {
          if (counter == 10) break;
          if (p < l95) {
            a_r = a;
            a = a < 0.1 ? 0.0 : a - ((a - a_l) / 2.0);
            b_l = b;
            b = b > 0.9 ? 1.0 : b + ((b_r - b) / 2.0);
          }
          if (p > l96) {
            a_l = a;
            a += (a_r - a) / 2.0;
            b_r = b;
            b -= (b - b_l) / 2.0;
          }
          p = ln_simpsons_integrate_exp([&](double v) { return std::log(density(fr, dp, v)) - r; }, a, b, 11);
          if (p >= l95 && p < l96) break;
          counter += 1;
        }
      }
      for (auto& e : kv.second) {
        IDRecord out_row = e.idr;
        out_row.freq = out_row.depth == 0 ? 0.0 : double(ml) * 0.01;
        const std::string ci = fmt2(a) + "-" + fmt2(b);
        const auto fields = fields_of(out_row, ci, e.normal_p, e.tumor_p);
        if (ref_set.count(e.tumor_p)) {
          removed_fasta_writer.write(out_row.id, reinterpret_cast<const uint8_t*>(e.tumor_p.data()), e.tumor_p.size());
          removed_writer.row(out_header, fields);
        } else {
          fasta_writer.write(out_row.id, reinterpret_cast<const uint8_t*>(e.tumor_p.data()), e.tumor_p.size());
          if (!e.normal_p.empty()) normal_writer.write(out_row.id, reinterpret_cast<const uint8_t*>(e.normal_p.data()), e.normal_p.size());
          tsv_writer.emit(fields);
        }
      }
    }
  };

  auto rows = read_tsv(tsv_in);
  for (size_t ri = 1; ri < rows.size(); ++ri) {  // first row = header
    const auto& c = rows[ri];
    if (c.size() != 21) throw Failure("CSV error: record has " + std::to_string(c.size()) + " fields, expected 21");
    IDRecord row;
    row.id = c[0]; row.transcript = c[1]; row.gene_id = c[2]; row.gene_name = c[3]; row.chrom = c[4];
    row.offset = IDRecord::parse_u64(c[5]); row.frame = IDRecord::parse_u64(c[6]); row.freq = parse_f64(c[7]);
    row.depth = uint32_t(IDRecord::parse_u64(c[8])); row.nvar = uint32_t(IDRecord::parse_u64(c[9])); row.nsomatic = uint32_t(IDRecord::parse_u64(c[10]));
    row.nvariant_sites = uint32_t(IDRecord::parse_u64(c[11])); row.nsomvariant_sites = uint32_t(IDRecord::parse_u64(c[12]));
    row.strand = c[13]; row.variant_sites = c[14]; row.somatic_positions = c[15]; row.somatic_aa_change = c[16];
    row.germline_positions = c[17]; row.germline_aa_change = c[18]; row.normal_sequence = c[19]; row.mutant_sequence = c[20];
    size_t som_pos = 0;
    if (!row.somatic_positions.empty() && row.somatic_positions.find('|') == std::string::npos) som_pos = size_t(IDRecord::parse_u64(row.somatic_positions));
    const std::string& orientation = row.strand;
    const size_t offset = size_t(row.offset);
    const int frame = (!row.id.empty() && row.id.back() == 'F') ? 1 : -1;
    const std::string tumor_peptide = to_protein(row.mutant_sequence, frame);
    const std::string normal_peptide = row.normal_sequence.empty() ? std::string() : to_protein(row.normal_sequence, frame);
    size_t i = 0;
    const std::pair<std::string, uint64_t> check{row.transcript, row.frame};
    auto sg = stop_gained.find(check);
    if (sg != stop_gained.end()) {
      const bool downstream = orientation == "Forward" ? offset > sg->second : (orientation == "Reverse" ? offset < sg->second : false);
      if (downstream) continue;
    }
    if (tumor_peptide.find('X') != std::string::npos && (std::fabs(row.freq - 1.0) < std::numeric_limits<double>::epsilon() || row.frame > 0))
      stop_gained[check] = offset;
    while (i + peptide_length <= tumor_peptide.size()) {
      const std::string tumor_pep = tumor_peptide.substr(i, peptide_length);
      if (tumor_pep.find('X') != std::string::npos) break;
      const std::string normal_pep = normal_peptide.size() >= i + peptide_length ? normal_peptide.substr(i, peptide_length) : normal_peptide;
      if (normal_pep.empty() && som_pos > 0) {
        if (orientation == "Forward") {
          if (((i + peptide_length) * 3) + offset <= som_pos) { i += 1; continue; }
        } else if (orientation == "Reverse") {
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
      if (current_variant == std::make_tuple(std::string(), std::string(), std::string())) current_variant = cur3;
      seen_peptides.insert(tumor_pep);
      IDRecord row2 = row;
      row2.id = std::to_string(i) + "_" + row2.id;
      const uint64_t frameshift = row2.frame;
      const double current_freq = row2.freq;
      const uint32_t current_depth = row2.depth;
      FilteredRow value{row2, tumor_pep, normal_pep};
      const Key key{frameshift, row.somatic_positions, row.germline_positions};
      if (current_sites != region_sites) {
        flush(false);
        frequencies.clear();
        frequencies[key] = {current_freq * double(current_depth)};
        depth.clear();
        depth[key] = {current_depth};
        records.clear();
        records[key] = {value};
        region_sites = current_sites;
      } else {
        // entry(..).or_insert_with(|| vec![x]).push(x): a fresh key receives the value twice
        auto itd = depth.find(key);
        if (itd == depth.end()) depth[key] = {current_depth, current_depth};
        else itd->second.push_back(current_depth);
        auto itf = frequencies.find(key);
        if (itf == frequencies.end()) frequencies[key] = {current_freq * double(current_depth), current_freq * double(current_depth)};
        else itf->second.push_back(current_freq * double(current_depth));
        auto itr = records.find(key);
        if (itr == records.end()) records[key] = {value, value};
        else itr->second.push_back(value);
      }
    }
  }
  flush(true);
  fclose(tsv_writer.f);
  fclose(removed_writer.f);
  fclose(removed_fasta_writer.f);
  fclose(normal_writer.f);
}

}  // namespace peptides
}  // namespace oracle
