// writer.hpp — the three output streams of `somatic` in the reference's byte format
// (SURVEY.md Appendix B): mutant FASTA on stdout, wild-type FASTA (--normal-output) and the
// info TSV (--tsv; header from the IDRecord field order, src/common.rs:351-373, written only
// together with the first row).
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../io/fmt_util.hpp"
#include "residue_normal.hpp"

namespace mph {

struct Outputs {
  FILE* fasta = nullptr;
  FILE* tsv = nullptr;
  FILE* normal = nullptr;
  bool header_written = false;
};

inline void write_fasta(FILE* f, std::string_view id, std::string_view seq) {
  fputc('>', f);
  fwrite(id.data(), 1, id.size(), f);
  fputc('\n', f);
  fwrite(seq.data(), 1, seq.size(), f);
  fputc('\n', f);
}

// `normal` mode: FASTA on stdout and the 20-column TSV (IDRecord of src/normal_microphasing.rs:80-102)
inline void write_records_normal(const Batch& b, const std::vector<OutRecord>& recs, Outputs& o) {
  static const char* header =
      "id\ttranscript\tgene_id\tgene_name\tchrom\toffset\tframe\tfreq\tdepth\tnvar\tnsomatic\tnvariant_sites\tnsomvariant_sites\t"
      "strand\tvariant_sites\tsomatic_positions\tsomatic_aa_change\tgermline_positions\tgermline_aa_change\tpeptide_sequence\n";
  std::string line;
  for (const OutRecord& r : recs) {
    if (r.has_mt) write_fasta(o.fasta, r.info.id, r.mt_str());
    if (!o.header_written) {
      fputs(header, o.tsv);
      o.header_written = true;
    }
    const TxMeta& tm = b.txs[r.info.tx];
    const GeneMeta& gm = b.genes[tm.gene];
    const std::string fields[20] = {r.info.id, tm.id, gm.id, gm.name, gm.chrom, std::to_string(r.info.offset), std::to_string(r.info.frame),
                                    mphfmt::format_f64(r.info.freq), std::to_string(r.info.depth), std::to_string(r.info.nvar),
                                    std::to_string(r.info.nsomatic), std::to_string(r.info.nvariant_sites),
                                    std::to_string(r.info.nsomvariant_sites), tm.reverse ? "Reverse" : "Forward", r.info.variant_sites,
                                    r.info.somatic_positions, r.info.somatic_aa_change, r.info.germline_positions,
                                    r.info.germline_aa_change, r.info.mutant_sequence};
    line.clear();
    for (int i = 0; i < 20; ++i) {
      if (i) line.push_back('\t');
      mphfmt::csv_field(fields[i], '\t', line);
    }
    line.push_back('\n');
    fwrite(line.data(), 1, line.size(), o.tsv);
  }
}

inline void write_records(const Batch& b, const std::vector<OutRecord>& recs, Outputs& o) {
  if (b.mode == 1) return write_records_normal(b, recs, o);
  static const char* header =
      "id\ttranscript\tgene_id\tgene_name\tchrom\toffset\tframe\tfreq\tdepth\tnvar\tnsomatic\tnvariant_sites\tnsomvariant_sites\t"
      "strand\tvariant_sites\tsomatic_positions\tsomatic_aa_change\tgermline_positions\tgermline_aa_change\tnormal_sequence\t"
      "mutant_sequence\n";
  std::string line;
  for (const OutRecord& r : recs) {
    if (r.has_mt) write_fasta(o.fasta, r.info.id, r.mt_str());
    if (r.has_wt) write_fasta(o.normal, r.info.id, r.wt_str());
    if (!o.header_written) {
      fputs(header, o.tsv);
      o.header_written = true;
    }
    const TxMeta& tm = b.txs[r.info.tx];
    const GeneMeta& gm = b.genes[tm.gene];
    const std::string fields[21] = {r.info.id, tm.id, gm.id, gm.name, gm.chrom, std::to_string(r.info.offset), std::to_string(r.info.frame),
                                    mphfmt::format_f64(r.info.freq), std::to_string(r.info.depth), std::to_string(r.info.nvar),
                                    std::to_string(r.info.nsomatic), std::to_string(r.info.nvariant_sites),
                                    std::to_string(r.info.nsomvariant_sites), tm.reverse ? "Reverse" : "Forward", r.info.variant_sites,
                                    r.info.somatic_positions, r.info.somatic_aa_change, r.info.germline_positions,
                                    r.info.germline_aa_change, r.info.normal_sequence, r.info.mutant_sequence};
    line.clear();
    for (int i = 0; i < 21; ++i) {
      if (i) line.push_back('\t');
      mphfmt::csv_field(fields[i], '\t', line);
    }
    line.push_back('\n');
    fwrite(line.data(), 1, line.size(), o.tsv);
  }
}

}  // namespace mph
