// TEST INFRASTRUCTURE ONLY — times the host ingest of the file driver (BAM load, per-gene fetch, packing) without a GPU.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>
#include "../../microphaser_b200/csrc/host/ingest.hpp"
int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ingest_bench DIR [io_threads]\n"); return 2; }
  const std::string d = argv[1];
  const unsigned thr = argc > 2 ? unsigned(atoi(argv[2])) : std::max(1u, std::thread::hardware_concurrency());
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  mphio::fast_inflate_enabled().store(getenv("MPH_ZLIB_INFLATE") == nullptr);
  for (int rep = 0; rep < 3; ++rep) {
    const auto t0 = now();
    mphio::BamFile bam(d + "/reads.bam", thr);
    mphio::VcfFile vcf(d + "/variants.vcf");
    mphio::FastaIndexed fasta(d + "/ref.fa");
    std::ifstream gf(d + "/annotation.gtf");
    mph::IngestOptions io;
    const auto t1 = now();
    mph::ReadBufferLoader loader(bam);
    double load_ms = 0;
    std::vector<mph::GeneInput> genes = mph::ingest_genes_with(gf, [&]() -> mph::ReadBuffer& { mph::ReadBuffer& r = loader.get(); load_ms = ms(t1, now()); return r; }, vcf, fasta, io, true);
    const auto t2 = t1 + std::chrono::microseconds(int64_t(load_ms * 1000));
    const auto t3 = now();
    uint64_t n_reads = 0;
    for (auto& g : genes) n_reads += g.n_reads;
    mph::Packer packer(27, 0);
    mph::pack_genes(genes, 0, genes.size(), packer);
    const auto t4 = now();
    printf("open %.1f ms, BAM load %.1f ms, per-gene fetch %.1f ms (%zu genes, %llu reads), packing (1 thread) %.1f ms, windows %llu, total ingest %.1f ms\n", ms(t0, t1), ms(t1, t2),
           ms(t2, t3), genes.size(), (unsigned long long)n_reads, ms(t3, t4), (unsigned long long)packer.batch().n_windows, ms(t0, t3));
  }
}
