// TEST / BENCH INFRASTRUCTURE — writes the synthetic workload of bench.py as files (FASTA + .fai, GTF, VCF, BAM) for the
// CPU oracle, without loading the CUDA library: the generator is a host-only header of the product tree
// (csrc/host/synth_files.hpp), compiled here into a stand-alone tool.
//   mph_synth_files <dir> <seed> <n_transcripts> <coverage> <germline/kb> <somatic/kb> <ins_frac> <del_frac>
#include <cstdio>
#include <cstdlib>

#include "../microphaser_b200/csrc/host/synth_files.hpp"

int main(int argc, char** argv) {
  if (argc < 9) {
    fprintf(stderr, "usage: mph_synth_files <dir> <seed> <n_transcripts> <coverage> <germline/kb> <somatic/kb> <ins_frac> <del_frac>\n");
    return 2;
  }
  mph::SynthParams sp;
  sp.seed = strtoull(argv[2], nullptr, 0);
  sp.n_transcripts = uint32_t(atoi(argv[3]));
  sp.coverage = atof(argv[4]);
  sp.germline_per_kb = atof(argv[5]);
  sp.somatic_per_kb = atof(argv[6]);
  sp.ins_var_frac = atof(argv[7]);
  sp.del_var_frac = atof(argv[8]);
  try {
    const mph::SynthFileStats st = mph::synth_write_files(sp, 27, argv[1]);
    printf("{\"reads\": %llu, \"variants\": %llu, \"transcripts\": %llu}\n", (unsigned long long)st.reads, (unsigned long long)st.variants,
           (unsigned long long)st.transcripts);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
