// synth_files.hpp — writes the native synthetic workload as real files (FASTA + .fai, sorted GTF,
// VCF with the SOMATIC flag, coordinate-sorted BGZF BAM), so the reference-shaped CLI and the CPU
// oracle can be run on exactly the genes / reads / variants the bench packs natively.
#pragma once
#include <zlib.h>

#include <cstdio>
#include <map>

#include "synth_native.hpp"

namespace mph {

namespace synthio {

inline void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back(uint8_t(x >> (8 * i))); }
inline void put16(std::vector<uint8_t>& v, uint16_t x) { v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); }

inline uint32_t reg2bin(uint32_t beg, uint32_t end) {
  --end;
  if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
  if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
  if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
  if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
  if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
  return 0;
}

struct BgzfWriter {
  FILE* f;
  std::vector<uint8_t> buf;
  explicit BgzfWriter(const std::string& path) : f(fopen(path.c_str(), "wb")) {
    if (!f) throw std::runtime_error("cannot create " + path);
  }
  void flush_block(const uint8_t* p, size_t n) {
    std::vector<uint8_t> comp(compressBound(uLong(n)) + 64);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<uint8_t*>(p);
    zs.avail_in = uInt(n);
    zs.next_out = comp.data();
    zs.avail_out = uInt(comp.size());
    deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const uint16_t bsize = uint16_t(clen + 25);
    const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    fwrite(hdr, 1, 16, f);
    fwrite(&bsize, 2, 1, f);
    fwrite(comp.data(), 1, clen, f);
    const uint32_t crc = uint32_t(crc32(crc32(0, nullptr, 0), p, uInt(n))), isize = uint32_t(n);
    fwrite(&crc, 4, 1, f);
    fwrite(&isize, 4, 1, f);
  }
  void write(const std::vector<uint8_t>& v) {
    buf.insert(buf.end(), v.begin(), v.end());
    size_t o = 0;
    while (buf.size() - o >= 0xFF00) { flush_block(buf.data() + o, 0xFF00); o += 0xFF00; }
    buf.erase(buf.begin(), buf.begin() + long(o));
  }
  void close() {
    if (!buf.empty()) flush_block(buf.data(), buf.size());
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, f);
    fclose(f);
    f = nullptr;
  }
};

}  // namespace synthio

struct SynthFileStats {
  uint64_t reads = 0, variants = 0, transcripts = 0;
};

inline SynthFileStats synth_write_files(const SynthParams& sp, uint32_t window_len, const std::string& dir) {
  using namespace synthio;
  SynthFileStats st;
  std::vector<std::string> contig_names;
  std::map<std::string, std::string> contig_seq;
  std::vector<std::vector<uint8_t>> bam_recs;
  std::string gtf, vcf;
  synth_generate(sp, window_len, true, [&](SynthGene& sg) {
    const HostGene& g = sg.gene;
    if (!contig_seq.count(g.chrom)) contig_names.push_back(g.chrom);
    std::string& cs = contig_seq[g.chrom];
    const int tid = int(std::find(contig_names.begin(), contig_names.end(), g.chrom) - contig_names.begin());
    cs.resize(size_t(g.start) + sg.ref.size(), 'A');
    memcpy(&cs[g.start], sg.ref.data(), sg.ref.size());
    // GTF (sorted by construction)
    const HostTranscript& t = g.transcripts[0];
    const char strand = t.reverse ? '-' : '+';
    char buf[1024];
    std::string ga = "gene_id \"" + g.id + "\"; gene_version \"1\"; gene_name \"" + g.name + "\"; gene_source \"synth\"; gene_biotype \"protein_coding\";";
    std::string ta = ga + " transcript_id \"" + t.id + "\"; transcript_name \"" + g.name + "-201\"; transcript_biotype \"protein_coding\";";
    snprintf(buf, sizeof buf, "%s\tsynth\tgene\t%u\t%u\t.\t%c\t.\t%s\n", g.chrom.c_str(), g.start + 1, g.end, strand, ga.c_str());
    gtf += buf;
    snprintf(buf, sizeof buf, "%s\tsynth\ttranscript\t%u\t%u\t.\t%c\t.\t%s\n", g.chrom.c_str(), g.start + 1, g.end, strand, ta.c_str());
    gtf += buf;
    std::vector<HostExon> ex = sg.cds;
    if (t.reverse) std::reverse(ex.begin(), ex.end());
    uint32_t consumed = 0;
    for (size_t i = 0; i < ex.size(); ++i) {
      const uint32_t frame = i == 0 ? 0 : (3 - consumed % 3) % 3;
      snprintf(buf, sizeof buf, "%s\tsynth\tCDS\t%u\t%u\t.\t%c\t%u\t%s exon_number \"%zu\";\n", g.chrom.c_str(), ex[i].start + 1, ex[i].end, strand, frame, ta.c_str(), i + 1);
      gtf += buf;
      if (i == 0) {
        if (t.reverse) snprintf(buf, sizeof buf, "%s\tsynth\tstart_codon\t%u\t%u\t.\t%c\t0\t%s\n", g.chrom.c_str(), ex[i].end - 2, ex[i].end, strand, ta.c_str());
        else snprintf(buf, sizeof buf, "%s\tsynth\tstart_codon\t%u\t%u\t.\t%c\t0\t%s\n", g.chrom.c_str(), ex[i].start + 1, ex[i].start + 3, strand, ta.c_str());
        gtf += buf;
      }
      consumed += ex[i].end - ex[i].start;
    }
    const HostExon last = ex.back();
    if (t.reverse) snprintf(buf, sizeof buf, "%s\tsynth\tthree_prime_utr\t%u\t%u\t.\t%c\t.\t%s\n", g.chrom.c_str(), last.start - sg.tail_len + 1, last.start, strand, ta.c_str());
    else snprintf(buf, sizeof buf, "%s\tsynth\tthree_prime_utr\t%u\t%u\t.\t%c\t.\t%s\n", g.chrom.c_str(), last.end + 1, last.end + sg.tail_len, strand, ta.c_str());
    gtf += buf;
    ++st.transcripts;
    // VCF
    for (auto& site : sg.sites)
      for (auto& v : site) {
        auto up = [](uint8_t c) { return c >= 'a' ? char(c - 32) : char(c); };
        std::string ref_allele(1, up(sg.ref[v.pos - g.start])), alt_allele(1, char(v.alt));
        if (v.kind == MPH_DEL) {
          for (uint32_t x = 1; x <= v.len; ++x) ref_allele.push_back(up(sg.ref[v.pos - g.start + x]));
          alt_allele = ref_allele.substr(0, 1);
        } else if (v.kind == MPH_INS) {
          alt_allele = v.ins;
        }
        snprintf(buf, sizeof buf, "%s\t%u\t.\t%s\t%s\t100\t.\tDP=100%s\n", g.chrom.c_str(), v.pos + 1, ref_allele.c_str(), alt_allele.c_str(), v.germline ? "" : ";SOMATIC");
        vcf += buf;
        ++st.variants;
      }
    // BAM records
    for (size_t i = 0; i < sg.reads.size(); ++i) {
      const HostRead& r = sg.reads[i];
      std::vector<uint8_t> body;
      put32(body, uint32_t(tid));
      put32(body, r.start);
      const std::string& qn = sg.qnames[i];
      body.push_back(uint8_t(qn.size() + 1));
      body.push_back(60);
      put16(body, uint16_t(reg2bin(r.start, r.end > r.start ? r.end : r.start + 1)));
      put16(body, uint16_t(r.n_cigar));
      put16(body, 0);
      put32(body, r.l_seq);
      put32(body, 0xFFFFFFFFu);
      put32(body, 0xFFFFFFFFu);
      put32(body, 0);
      body.insert(body.end(), qn.begin(), qn.end());
      body.push_back(0);
      for (uint32_t c = 0; c < r.n_cigar; ++c) put32(body, r.cigar[c]);
      body.insert(body.end(), r.seq4, r.seq4 + (r.l_seq + 1) / 2);
      body.insert(body.end(), r.qual, r.qual + r.l_seq);
      std::vector<uint8_t> rec;
      put32(rec, uint32_t(body.size()));
      rec.insert(rec.end(), body.begin(), body.end());
      bam_recs.push_back(std::move(rec));
      ++st.reads;
    }
  });
  // FASTA + .fai
  {
    FILE* f = fopen((dir + "/ref.fa").c_str(), "wb");
    FILE* fai = fopen((dir + "/ref.fa.fai").c_str(), "wb");
    if (!f || !fai) throw std::runtime_error("cannot create FASTA in " + dir);
    uint64_t off = 0;
    for (auto& name : contig_names) {
      std::string& cs = contig_seq[name];
      cs.resize(cs.size() + 1000, 'A');
      off += fprintf(f, ">%s\n", name.c_str());
      fprintf(fai, "%s\t%zu\t%llu\t60\t61\n", name.c_str(), cs.size(), (unsigned long long)off);
      for (size_t i = 0; i < cs.size(); i += 60) {
        const size_t n = std::min<size_t>(60, cs.size() - i);
        fwrite(cs.data() + i, 1, n, f);
        fputc('\n', f);
        off += n + 1;
      }
    }
    fclose(f);
    fclose(fai);
  }
  {
    FILE* f = fopen((dir + "/annotation.gtf").c_str(), "wb");
    fwrite(gtf.data(), 1, gtf.size(), f);
    fclose(f);
    f = fopen((dir + "/variants.vcf").c_str(), "wb");
    fputs("##fileformat=VCFv4.2\n", f);
    for (auto& name : contig_names) fprintf(f, "##contig=<ID=%s,length=%zu>\n", name.c_str(), contig_seq[name].size());
    fputs("##INFO=<ID=DP,Number=1,Type=Integer,Description=\"depth\">\n##INFO=<ID=SOMATIC,Number=0,Type=Flag,Description=\"Somatic variant\">\n", f);
    fputs("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n", f);
    fwrite(vcf.data(), 1, vcf.size(), f);
    fclose(f);
  }
  {
    BgzfWriter w(dir + "/reads.bam");
    std::vector<uint8_t> hdr = {'B', 'A', 'M', 1};
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (auto& name : contig_names) text += "@SQ\tSN:" + name + "\tLN:" + std::to_string(contig_seq[name].size()) + "\n";
    put32(hdr, uint32_t(text.size()));
    hdr.insert(hdr.end(), text.begin(), text.end());
    put32(hdr, uint32_t(contig_names.size()));
    for (auto& name : contig_names) {
      put32(hdr, uint32_t(name.size() + 1));
      hdr.insert(hdr.end(), name.begin(), name.end());
      hdr.push_back(0);
      put32(hdr, uint32_t(contig_seq[name].size()));
    }
    w.write(hdr);
    for (auto& r : bam_recs) w.write(r);
    w.close();
  }
  return st;
}

}  // namespace mph
