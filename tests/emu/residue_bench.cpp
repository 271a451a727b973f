// TEST INFRASTRUCTURE ONLY — times the host residue on a synthetic batch (device side emulated).
#include <chrono>
#include <cstdio>
#include "../../microphaser_b200/csrc/host/synth_native.hpp"
#include "emu_pipeline.hpp"
int main(int argc, char** argv) {
  mph::SynthParams sp;
  sp.n_transcripts = argc > 1 ? atoi(argv[1]) : 1000;
  sp.coverage = 100.0;
  mph::Packer packer(27);
  mph::synth_into(packer, sp);
  mph::Batch& b = packer.batch();
  auto t0 = std::chrono::steady_clock::now();
  mph::PhaseRaw raw = mphemu::phase(b);
  auto t1 = std::chrono::steady_clock::now();
  std::vector<mph::OutRecord> recs;
  mph::ResidueStats st;
  for (int rep = 0; rep < 3; ++rep) {
    recs.clear();
    mph::Residue r(b, raw);
    auto a = std::chrono::steady_clock::now();
    r.run(0, uint32_t(b.txs.size()), recs, st);
    auto c = std::chrono::steady_clock::now();
    printf("residue %.1f ms, %zu records, %zu interesting windows of %llu\n", std::chrono::duration<double, std::milli>(c - a).count(), recs.size(), raw.iw.size(), (unsigned long long)b.n_windows);
  }
  printf("emu phase %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
}
