// TEST INFRASTRUCTURE ONLY — dumps what the product's I/O and formatting helpers (csrc/io/hts_io.hpp, fmt_util.hpp,
// core/phase_core.h) make of their inputs, as text, so tests/test_io_independent.py can compare them with independent
// Python implementations (tests/golden/bamlite.py, repr(), hashlib, csv). The oracle shares these headers with the
// product, so agreement between the two says nothing about them; this does.
//   io_dump bam <file> [threads]   one line per record: tid pos end mapq flag l_seq qname cigar seq qualhex probes...
//   io_dump fmt                    stdin: "f <hex u64>" | "id <tx> <offset> <strand> <hexbytes>" | "csv <hexbytes>"
#include <cinttypes>
#include <cstdio>
#include <iostream>
#include <string>

#include "../../microphaser_b200/csrc/core/phase_core.h"
#include "../../microphaser_b200/csrc/host/ingest.hpp"
#include "../../microphaser_b200/csrc/io/fmt_util.hpp"
#include "../../microphaser_b200/csrc/io/hts_io.hpp"

static std::string unhex(const std::string& h) {
  std::string s;
  for (size_t i = 0; i + 1 < h.size(); i += 2) s.push_back(char(std::stoi(h.substr(i, 2), nullptr, 16)));
  return s;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "bam") {
    mphio::BamFile bam(argv[2], argc > 3 ? unsigned(atoi(argv[3])) : 1u);
    mphio::BamRecord r;
    static const char* ops = "MIDNSHP=X";
    while (bam.next(r)) {
      printf("%d %d %lld %u %u %u %s ", r.tid, r.pos, (long long)r.end_pos(), unsigned(r.mapq), unsigned(r.flag), r.l_seq, r.qname.c_str());
      if (r.cigar.empty()) printf("*");
      for (uint32_t c : r.cigar) printf("%u%c", c >> 4, ops[c & 15]);
      printf(" ");
      for (uint32_t i = 0; i < r.l_seq; ++i) putchar(r.base(i));
      printf(" ");
      for (uint8_t q : r.qual) printf("%02x", q);
      // CIGAR walk at a few reference positions: the host statement and the kernels' statement of read_pos
      for (int64_t probe : {int64_t(r.pos), int64_t(r.pos) + 7, int64_t(r.pos) + 50, r.end_pos() - 1, r.end_pos()}) {
        uint32_t q1 = 0, q2 = 0;
        const int a = mphio::cigar_read_pos(r.cigar, r.pos, probe, &q1);
        const int b = mph_read_pos(r.cigar.data(), uint32_t(r.cigar.size()), r.l_seq, uint32_t(r.pos), uint32_t(probe), &q2);
        printf(" %d:%u:%d:%u", a, a == 1 ? q1 : 0u, b, b == 1 ? q2 : 0u);
      }
      printf("\n");
    }
    return 0;
  }
  if (argc >= 2 && std::string(argv[1]) == "fmt") {
    std::string kind;
    while (std::cin >> kind) {
      if (kind == "f") {
        std::string h;
        std::cin >> h;
        const uint64_t bits = std::stoull(h, nullptr, 16);
        double v;
        memcpy(&v, &bits, 8);
        printf("%s\n", mphfmt::format_f64(v).c_str());
      } else if (kind == "id") {
        std::string tx, strand, hex;
        uint64_t off;
        std::cin >> tx >> off >> strand >> hex;
        const std::string seq = unhex(hex);
        printf("%s\n", mphfmt::record_id(reinterpret_cast<const uint8_t*>(seq.data()), seq.size(), tx, off, strand[0]).c_str());
      } else if (kind == "csv") {
        std::string hex, out;
        std::cin >> hex;
        mphfmt::csv_field(unhex(hex == "-" ? "" : hex), '\t', out);
        for (unsigned char c : out) printf("%02x", c);
        printf("\n");
      }
    }
    return 0;
  }
  return 2;
}
