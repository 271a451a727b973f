// oracle_main.cpp — TEST INFRASTRUCTURE ONLY.
// Command-line front end of the CPU oracle with the reference's CLI surface
// (reference src/main.rs:34-265, src/somatic_cli.yaml, germline_cli.yaml, filter_cli.yaml,
// build_ref_cli.yaml): GTF on stdin, mutant FASTA on stdout.
//   oracle somatic <tumor.bam> -r ref.fa -b variants.vcf [-t info.tsv] [-n normal.fasta] [-w 27] [-u]
//   oracle normal  <normal.bam> -r ref.fa -b variants.vcf [-t info.tsv] [-w 27] [-u]
//   oracle build_reference -r normal.fa -o peptides.bin [-l 9]
//   oracle filter -t info.tsv -r peptides.bin [-o ..] [-s ..] [-p ..] [-n ..] [-l 9]
// Exit status: 0 ok, 1 on Err (Failure / I/O error), 101 on a restated Rust panic.
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "oracle_normal.hpp"
#include "oracle_peptides.hpp"
#include "oracle_somatic.hpp"

namespace {

struct Args {
  std::map<std::string, std::string> opt;
  std::vector<std::string> pos;
  std::set<std::string> flags;
};

// long name -> (short, takes_value)
struct Spec {
  const char* lng;
  char shrt;
  bool value;
};

Args parse(int argc, char** argv, int first, const std::vector<Spec>& specs) {
  Args a;
  for (int i = first; i < argc; ++i) {
    std::string s = argv[i];
    const Spec* sp = nullptr;
    std::string val;
    bool have_val = false;
    if (s.rfind("--", 0) == 0) {
      std::string name = s.substr(2);
      size_t eq = name.find('=');
      if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); have_val = true; }
      for (auto& x : specs) if (name == x.lng) sp = &x;
      if (!sp) throw oracle::Failure("unknown option " + s);
    } else if (s.size() >= 2 && s[0] == '-' && !(s[1] >= '0' && s[1] <= '9')) {
      for (auto& x : specs) if (s[1] == x.shrt) sp = &x;
      if (!sp) throw oracle::Failure("unknown option " + s);
      if (s.size() > 2) { val = s.substr(s[2] == '=' ? 3 : 2); have_val = true; }
    } else {
      a.pos.push_back(s);
      continue;
    }
    if (sp->value) {
      if (!have_val) {
        if (i + 1 >= argc) throw oracle::Failure(std::string("option --") + sp->lng + " needs a value");
        val = argv[++i];
      }
      a.opt[sp->lng] = val;
    } else {
      a.flags.insert(sp->lng);
    }
  }
  return a;
}

FILE* open_out(const std::string& p) {
  FILE* f = fopen(p.c_str(), "wb");
  if (!f) throw oracle::Failure("cannot create " + p);
  return f;
}

int run(int argc, char** argv) {
  if (argc < 2) return 0;
  std::string sub = argv[1];
  const char* tr = getenv("MPH_ORACLE_TRACE");
  if (tr && *tr) oracle::somatic::trace().f = fopen(tr, "w");
  const char* st = getenv("MPH_ORACLE_STATS");
  auto t0 = std::chrono::steady_clock::now();
  if (sub == "somatic") {
    Args a = parse(argc, argv, 2, {{"ref", 'r', true}, {"variants", 'b', true}, {"window-len", 'w', true}, {"tsv", 't', true},
                                   {"normal-output", 'n', true}, {"unsupported-allele-warning-only", 'u', false}, {"verbose", 'v', false}});
    if (a.pos.size() != 1 || !a.opt.count("ref") || !a.opt.count("variants")) throw oracle::Failure("usage: somatic <bam> -r REF -b VCF");
    mphio::BamFile bam(a.pos[0]);
    mphio::VcfFile vcf(a.opt["variants"]);
    mphio::FastaIndexed fasta(a.opt["ref"]);
    std::string tsv = a.opt.count("tsv") ? a.opt["tsv"] : "info.tsv";
    std::string nrm = a.opt.count("normal-output") ? a.opt["normal-output"] : "normal.fasta";
    uint64_t wl = a.opt.count("window-len") ? std::stoull(a.opt["window-len"]) : 27;
    oracle::somatic::Writers w{{stdout}, {open_out(tsv)}, {open_out(nrm)}};
    auto t1 = std::chrono::steady_clock::now();
    oracle::somatic::phase(fasta, std::cin, vcf, bam, w, wl, a.flags.count("unsupported-allele-warning-only") != 0);
    auto t2 = std::chrono::steady_clock::now();
    fclose(w.tsv.f);
    fclose(w.normal.f);
    if (st && *st) {
      FILE* sf = fopen(st, "w");
      auto& s = oracle::somatic::stats();
      fprintf(sf, "{\"windows\": %llu, \"read_windows\": %llu, \"print_calls\": %llu, \"load_s\": %.6f, \"phase_s\": %.6f}\n",
              (unsigned long long)s.windows, (unsigned long long)s.read_windows, (unsigned long long)s.print_calls,
              std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
      fclose(sf);
    }
    return 0;
  }
  if (sub == "normal") {
    Args a = parse(argc, argv, 2, {{"ref", 'r', true}, {"variants", 'b', true}, {"window-len", 'w', true}, {"tsv", 't', true},
                                   {"unsupported-allele-warning-only", 'u', false}, {"verbose", 'v', false}});
    if (a.pos.size() != 1 || !a.opt.count("ref") || !a.opt.count("variants")) throw oracle::Failure("usage: normal <bam> -r REF -b VCF");
    mphio::BamFile bam(a.pos[0]);
    mphio::VcfFile vcf(a.opt["variants"]);
    mphio::FastaIndexed fasta(a.opt["ref"]);
    std::string tsv = a.opt.count("tsv") ? a.opt["tsv"] : "info.tsv";
    uint64_t wl = a.opt.count("window-len") ? std::stoull(a.opt["window-len"]) : 27;
    oracle::normal::Writers w{{stdout}, {open_out(tsv)}};
    auto t1 = std::chrono::steady_clock::now();
    oracle::normal::phase(fasta, std::cin, vcf, bam, w, wl, a.flags.count("unsupported-allele-warning-only") != 0);
    auto t2 = std::chrono::steady_clock::now();
    fclose(w.tsv.f);
    if (st && *st) {
      FILE* sf = fopen(st, "w");
      auto& s = oracle::normal::stats();
      fprintf(sf, "{\"windows\": %llu, \"read_windows\": %llu, \"load_s\": %.6f, \"phase_s\": %.6f}\n", (unsigned long long)s.windows,
              (unsigned long long)s.read_windows, std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
      fclose(sf);
    }
    return 0;
  }
  if (sub == "build_reference") {
    Args a = parse(argc, argv, 2, {{"reference", 'r', true}, {"output", 'o', true}, {"peptide-length", 'l', true}, {"verbose", 'v', false}});
    if (!a.opt.count("reference") || !a.opt.count("output")) throw oracle::Failure("usage: build_reference -r FASTA -o BIN");
    size_t l = a.opt.count("peptide-length") ? std::stoul(a.opt["peptide-length"]) : 9;
    oracle::peptides::build(a.opt["reference"], a.opt["output"], stdout, l);
    return 0;
  }
  if (sub == "filter") {
    Args a = parse(argc, argv, 2, {{"tsv", 't', true}, {"reference", 'r', true}, {"tsv-output", 'o', true}, {"similar-removed", 's', true},
                                   {"removed-peptides", 'p', true}, {"normal-output", 'n', true}, {"peptide-length", 'l', true},
                                   {"verbose", 'v', false}});
    auto get = [&](const char* a1, const char* def) { return a.opt.count(a1) ? a.opt[a1] : std::string(def); };
    if (!a.opt.count("tsv") || !a.opt.count("reference")) throw oracle::Failure("usage: filter -t TSV -r BIN");
    size_t l = a.opt.count("peptide-length") ? std::stoul(a.opt["peptide-length"]) : 9;
    oracle::peptides::filter(a.opt["reference"], a.opt["tsv"], stdout, get("normal-output", "normal.filtered.fa"),
                             get("tsv-output", "info.filtered.tsv"), get("similar-removed", "info.removed.tsv"),
                             get("removed-peptides", "peptides.removed.fasta"), l);
    return 0;
  }
  throw oracle::Failure("unknown subcommand " + sub);
}

}  // namespace

int main(int argc, char** argv) {
  try {
    int rc = run(argc, argv);
    fflush(stdout);
    return rc;
  } catch (const oracle::Panic& e) {
    fflush(stdout);
    fprintf(stderr, "thread 'main' panicked at '%s'\n", e.what());
    return 101;
  } catch (const std::exception& e) {
    fflush(stdout);
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
