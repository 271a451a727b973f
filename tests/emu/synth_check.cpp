// TEST INFRASTRUCTURE ONLY — the native synthetic workload (the one bench.py packs with mph_synth_batch) through the
// product's packer and host residue with the device side emulated (emu_pipeline.hpp), written the way mph_result_write
// does, plus the same workload as files for the oracle. Lets the GPU-less container check that the packed batch and
// the files describe the same genes / reads / variants, and diff config-sized workloads against the oracle.
//   synth_check <out_dir> <files_dir|-> <mode 0|1> <seed> <n_tx> <coverage> <germ/kb> <som/kb> <ins_frac> <del_frac>
#include <cstdio>
#include <cstdlib>

#include "../../microphaser_b200/csrc/host/synth_files.hpp"
#include "../../microphaser_b200/csrc/host/records_host.hpp"
#include "../../microphaser_b200/csrc/host/writer.hpp"
#include "emu_pipeline.hpp"

int main(int argc, char** argv) {
  if (argc < 11) return 2;
  const std::string out = argv[1], files = argv[2];
  const int mode = atoi(argv[3]);
  mph::SynthParams sp;
  sp.seed = strtoull(argv[4], nullptr, 0);
  sp.n_transcripts = uint32_t(atoi(argv[5]));
  sp.coverage = atof(argv[6]);
  sp.germline_per_kb = atof(argv[7]);
  sp.somatic_per_kb = atof(argv[8]);
  sp.ins_var_frac = atof(argv[9]);
  sp.del_var_frac = atof(argv[10]);
  try {
    if (files != "-") mph::synth_write_files(sp, 27, files);
    mph::Packer packer(27, mode);
    mph::synth_into(packer, sp);
    mph::Batch& b = packer.batch();
    {  // the 2-byte bus encoding of the reads must decode to the arrays the kernels use
      std::vector<uint32_t> ds, de;
      std::vector<uint8_t> df;
      for (int form = 0; form <= (b.modal_set ? 1 : 0); ++form) {  // both bus forms of the spans: a byte per read / one span + exceptions
        mph::decode_reads(b, ds, de, df, form);
        if (ds != b.read_start || de != b.read_end || df != b.read_flags) { fprintf(stderr, "read encoding (span form %d) does not round-trip\n", form); return 4; }
      }
      fprintf(stderr, "bus form of the spans: %d (span %u, %zu of %zu reads listed)\n", mph::bus_span_mode(b), b.modal_span, b.rd_mspan_exc.size(), b.rd_span.size());
      std::vector<uint32_t> vr, vv, vso, vco;
      std::vector<uint16_t> vn;
      mph::decode_side_table(b, vr, vv, vso, vco, vn);
      bool ok = vr == b.vr_read && vv == b.vr_vlo && vso == b.vr_seq_off && vn == b.vr_ncig;
      for (size_t e = 0; ok && e < vn.size(); ++e) ok = vn[e] == 0 || vco[e] == b.vr_cig_off[e];
      if (!ok) { fprintf(stderr, "side-table encoding does not round-trip\n"); return 4; }
    }
    mph::PhaseRaw raw = mphemu::phase(b);
    if (raw.err) { fprintf(stderr, "device error bits %u\n", raw.err); return 3; }
    std::vector<mph::OutRecord> recs;
    mph::ResidueStats st;
    if (mode == 1) { mph::ResidueNormal r(b, raw); r.run(0, uint32_t(b.txs.size()), recs, st); recs = mph::ordered_records(b, raw, std::move(recs)); }
    else { mph::Residue r(b, raw); r.run(0, uint32_t(b.txs.size()), recs, st); recs = mph::ordered_records(b, raw, std::move(recs)); }
    mph::Outputs o;
    o.fasta = fopen((out + "/out.fa").c_str(), "wb");
    o.tsv = fopen((out + "/out.tsv").c_str(), "wb");
    o.normal = mode == 0 ? fopen((out + "/out.normal.fa").c_str(), "wb") : nullptr;
    mph::write_records(b, recs, o);
    fclose(o.fasta); fclose(o.tsv);
    if (o.normal) fclose(o.normal);
    printf("{\"reads\": %zu, \"variants\": %zu, \"windows\": %llu, \"records\": %zu, \"replay_units\": %zu, \"device_records\": %zu}\n", b.read_start.size(), b.vars.size(),
           (unsigned long long)b.n_windows, recs.size(), b.replay.size(), raw.recs.size());
  } catch (const mph::Unsupported& e) {
    fprintf(stderr, "unsupported: %s\n", e.what());
    return 3;
  } catch (const mph::Fatal& e) {
    fprintf(stderr, "panic: %s\n", e.what());
    return 101;
  }
  return 0;
}
