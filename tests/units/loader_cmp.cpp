// TEST INFRASTRUCTURE ONLY — the threaded BAM loader (records point into the kept inflated batches, header parse on a pool) against
// the sequential loader (records copied into arenas) on the same file, record by record.
#include <cstdio>
#include "../../microphaser_b200/csrc/host/ingest.hpp"
int main(int argc, char** argv) {
  // sequential loader with zlib, threaded loader with the reader's own DEFLATE decoder (io/fast_inflate.hpp)
  mphio::BamFile b1(argv[1], 1);
  mph::ReadBuffer r1(b1);
  mphio::fast_inflate_enabled().store(true);
  mphio::BamFile b8(argv[1], 8);
  mph::ReadBuffer r8(b8);
  printf("records %zu %zu\n", r1.n_records(), r8.n_records());
  size_t bad = 0, n = 0;
  for (auto& name : b1.ref_names) {
    auto a = r1.fetch(name, 0, 0xFFFFFFFFull);  // copy of the deque
    const auto& b = r8.fetch(name, 0, 0xFFFFFFFFull);
    if (a.size() != b.size()) { printf("contig %s: %zu vs %zu\n", name.c_str(), a.size(), b.size()); ++bad; continue; }
    for (size_t i = 0; i < a.size(); ++i, ++n) {
      const auto& x = *a[i]; const auto& y = *b[i];
      bool ok = x.tid == y.tid && x.pos == y.pos && x.end == y.end && x.l_seq == y.l_seq && x.n_cigar == y.n_cigar && x.flag == y.flag && x.mapq == y.mapq && x.qname_hash == y.qname_hash;
      ok = ok && memcmp(x.seq_p, y.seq_p, (x.l_seq + 1) / 2) == 0 && memcmp(x.qual_p, y.qual_p, x.l_seq) == 0 && memcmp(x.cig_p, y.cig_p, 4 * x.n_cigar) == 0;
      if (!ok) ++bad;
    }
  }
  printf("compared %zu records, %zu differences\n", n, bad);
  return bad != 0;
}
