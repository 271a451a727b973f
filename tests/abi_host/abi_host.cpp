// TEST INFRASTRUCTURE — a stand-in for the Rust host of INTEGRATION.md: it parses the input files with the product's own
// readers, flattens every gene into the plain arrays of `mph_gene_in`, and goes through the C ABI only
// (mph_packer_create / add_gene / finish, mph_phase_batch, mph_result_write). Exercises the marshalling of
// mph_packer_add_gene with real data; the CLI reaches the packer without that step.
//   abi_host <somatic|normal> <reads.bam> <ref.fa> <variants.vcf> <annotation.gtf> <out.fa> <out.tsv> [<out.normal.fa>]
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <fstream>

#include "../../include/microphaser_gpu.h"
#include "../../microphaser_b200/csrc/host/ingest.hpp"

static int die(mph_ctx* ctx, int rc, const char* what) {
  fprintf(stderr, "%s: %s\n", what, mph_last_error(ctx));
  return rc == MPH_ERR_PANIC ? 101 : (rc == MPH_ERR_UNSUPPORTED ? 3 : 1);
}

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const int mode = std::string(argv[1]) == "normal" ? 1 : 0;
  try {
    mphio::BamFile bam(argv[2]);
    mphio::FastaIndexed fasta(argv[3]);
    mphio::VcfFile vcf(argv[4]);
    std::ifstream gtf(argv[5]);
    mph::IngestOptions io;
    io.mode = mode;
    if (mode == 1) io.min_mapq = 0;
    mph::ReadBuffer reads(bam);
    std::vector<mph::GeneInput> genes = mph::ingest_genes(gtf, reads, vcf, fasta, io);
    mph_ctx* ctx = nullptr;
    int rc = mph_ctx_create(0, &ctx);
    if (rc != MPH_OK) return die(nullptr, rc, "mph_ctx_create");
    mph_packer* pk = nullptr;
    rc = mph_packer_create(27, mode, &pk);
    if (rc != MPH_OK) return die(nullptr, rc, "mph_packer_create");
    for (mph::GeneInput& gi : genes) {
      std::vector<const char*> tx_id;
      std::vector<uint8_t> tx_rev;
      std::vector<uint32_t> tx_off{0}, ex_s, ex_e, ex_f;
      for (auto& t : gi.gene.transcripts) {
        tx_id.push_back(t.id.c_str());
        tx_rev.push_back(t.reverse ? 1 : 0);
        for (auto& e : t.exons) { ex_s.push_back(e.start); ex_e.push_back(e.end); ex_f.push_back(e.frame); }
        tx_off.push_back(uint32_t(ex_s.size()));
      }
      std::vector<uint32_t> r_start, r_end, r_lseq, r_soff, r_qoff, r_coff{0}, cig;
      std::vector<uint64_t> r_hash;
      std::vector<uint8_t> seq4, qual;
      for (auto& r : gi.reads) {
        r_start.push_back(r.start); r_end.push_back(r.end); r_lseq.push_back(r.l_seq); r_hash.push_back(r.qname_hash);
        r_soff.push_back(uint32_t(seq4.size()));
        seq4.insert(seq4.end(), r.seq4, r.seq4 + (r.l_seq + 1) / 2);
        r_qoff.push_back(uint32_t(qual.size()));
        qual.insert(qual.end(), r.qual, r.qual + r.l_seq);
        cig.insert(cig.end(), r.cigar, r.cigar + r.n_cigar);
        r_coff.push_back(uint32_t(cig.size()));
      }
      std::vector<uint32_t> v_pos, v_len, v_ioff{0};
      std::vector<uint8_t> v_kind, v_germ, v_alt, ins;
      std::vector<const char*> v_prot;
      for (auto& site : gi.sites)
        for (auto& v : site) {
          v_pos.push_back(v.pos); v_len.push_back(v.len); v_kind.push_back(v.kind); v_germ.push_back(v.germline ? 1 : 0); v_alt.push_back(v.alt);
          ins.insert(ins.end(), v.ins.begin(), v.ins.end());
          v_ioff.push_back(uint32_t(ins.size()));
          v_prot.push_back(v.prot_change.c_str());
        }
      uint8_t dummy8 = 0;
      uint32_t dummy32 = 0;
      mph_gene_in g;
      memset(&g, 0, sizeof g);
      g.gene_id = gi.gene.id.c_str(); g.gene_name = gi.gene.name.c_str(); g.chrom = gi.gene.chrom.c_str();
      g.gene_start = gi.gene.start; g.gene_end = gi.gene.end;
      g.refseq = gi.refseq.data(); g.refseq_len = uint32_t(gi.refseq.size());
      g.n_tx = uint32_t(tx_id.size()); g.tx_id = tx_id.data(); g.tx_reverse = tx_rev.data(); g.tx_exon_off = tx_off.data();
      g.exon_start = ex_s.empty() ? &dummy32 : ex_s.data(); g.exon_end = ex_e.empty() ? &dummy32 : ex_e.data(); g.exon_frame = ex_f.empty() ? &dummy32 : ex_f.data();
      g.n_reads = uint32_t(r_start.size()); g.max_read_len = gi.max_read_len;
      g.read_start = r_start.data(); g.read_end = r_end.data(); g.read_lseq = r_lseq.data(); g.read_qname_hash = r_hash.data();
      g.read_seq_off = r_soff.data(); g.seq4 = seq4.empty() ? &dummy8 : seq4.data(); g.read_qual_off = r_qoff.data(); g.qual = qual.empty() ? &dummy8 : qual.data();
      g.read_cigar_off = r_coff.data(); g.cigar = cig.empty() ? &dummy32 : cig.data();
      g.n_vars = uint32_t(v_pos.size());
      g.var_pos = v_pos.data(); g.var_kind = v_kind.data(); g.var_germline = v_germ.data(); g.var_alt = v_alt.data(); g.var_len = v_len.data();
      g.var_ins_off = v_ioff.data(); g.ins_bytes = ins.empty() ? &dummy8 : ins.data(); g.var_prot_change = v_prot.data();
      rc = mph_packer_add_gene(pk, &g);
      if (rc != MPH_OK) return die(nullptr, rc, "mph_packer_add_gene");
    }
    mph_batch* batch = nullptr;
    rc = mph_packer_finish(pk, 1, &batch);
    if (rc != MPH_OK) return die(nullptr, rc, "mph_packer_finish");
    mph_packer_destroy(pk);
    mph_result* res = nullptr;
    rc = mph_phase_batch(ctx, batch, &res);
    if (rc != MPH_OK) return die(ctx, rc, "mph_phase_batch");
    const int fd_fa = open(argv[6], O_WRONLY | O_CREAT | O_TRUNC, 0644), fd_tsv = open(argv[7], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    const int fd_n = argc > 8 ? open(argv[8], O_WRONLY | O_CREAT | O_TRUNC, 0644) : -1;
    int hw = 0;
    rc = mph_result_write(res, fd_fa, fd_tsv, fd_n, &hw);
    if (rc != MPH_OK) return die(nullptr, rc, "mph_result_write");
    close(fd_fa); close(fd_tsv);
    if (fd_n >= 0) close(fd_n);
    mph_result_destroy(res);
    mph_batch_destroy(batch);
    mph_ctx_destroy(ctx);
    return 0;
  } catch (const mph::Fatal& e) {
    fprintf(stderr, "thread 'main' panicked at '%s'\n", e.what());
    return 101;
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
