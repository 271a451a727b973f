// synth_native.hpp — native generator of exome-shaped synthetic workloads (bench only).
// Same shapes and distributions as microphaser_b200/synth.py, but
// it feeds the Packer directly instead of writing BAM/VCF/GTF/FASTA files, so a whole-exome batch
// (20 000 transcripts, 100x, ~40 M reads) is built in seconds. Parity of this path is covered at
// small scale by the file-based generator + oracle, and at full scale by size-independent checks
// (tests/test_gpu_properties.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <string>

#include "batch.hpp"

namespace mph {

struct SynthParams {
  uint64_t seed = 0x4D500003ull;
  uint32_t n_transcripts = 450, exons = 8, exon_min = 90, exon_max = 250, read_len = 150;
  double coverage = 30.0, germline_per_kb = 1.0, somatic_per_kb = 1.0, lowq_frac = 0.02, indel_read_frac = 0.03;
  double ins_var_frac = 0.0, del_var_frac = 0.0;  // share of the variants that are 1-6 nt insertions / deletions (config C4)
};

struct Rng {  // splitmix64 / xoshiro256**
  uint64_t s[4];
  explicit Rng(uint64_t seed) {
    for (auto& x : s) {
      seed += 0x9E3779B97F4A7C15ull;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      x = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint32_t below(uint32_t n) { return uint32_t((next() >> 32) * uint64_t(n) >> 32); }
  uint32_t range(uint32_t lo, uint32_t hi) { return lo + below(hi - lo + 1); }  // inclusive
  double uniform() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t poisson(double lam) {
    if (lam <= 0) return 0;
    const double l = std::exp(-lam);
    uint32_t k = 0;
    double p = 1.0;
    for (;;) {
      p *= uniform();
      if (p <= l) return k;
      ++k;
    }
  }
};

// what the generator hands over per gene
struct SynthGene {
  HostGene gene;
  std::vector<HostExon> cds;  // genomic order, CDS only (without stop codon + UTR)
  uint32_t tail_len = 63;     // stop codon + 3' UTR appended to the last exon
  std::vector<HostRead> reads;
  std::vector<std::string> qnames;
  std::vector<std::vector<HostVariant>> sites;
  std::vector<uint8_t> ref;   // [gene.start, gene.end + 100)
  uint32_t max_read_len = 0;
};

// Generates gene after gene and calls sink(SynthGene&). With all_bases every read carries real
// bases / qualities (needed to write a BAM); otherwise only reads that overlap a variant do.
// The genes, variants and alignments do not depend on all_bases: everything structural comes from one
// generator, the bases / qualities / carried alleles of a read from a generator seeded per read.
template <class Sink>
inline void synth_generate(const SynthParams& sp, uint32_t window_len, bool all_bases, Sink&& sink) {
  Rng rng(sp.seed);
  static const char B[4] = {'A', 'C', 'G', 'T'};
  auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; };
  auto code4 = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; };
  const uint32_t L = sp.read_len;
  std::vector<uint8_t> dummy_seq((L + 1) / 2, 0x11), dummy_qual(L, 35);
  uint32_t pos = 10000;
  std::string chrom = "chrS1";
  uint32_t genes_on_contig = 0;
  for (uint32_t gi = 0; gi < sp.n_transcripts; ++gi) {
    if (genes_on_contig == 850) {  // ~24 contigs for a whole exome
      genes_on_contig = 0;
      chrom = "chrS" + std::to_string(gi / 850 + 1);
      pos = 10000;
    }
    ++genes_on_contig;
    const bool reverse = (gi & 1) != 0;
    const uint32_t n_ex = sp.exons;
    std::vector<uint32_t> lens(n_ex);
    uint32_t total = 0;
    for (auto& l : lens) { l = rng.range(sp.exon_min, sp.exon_max); total += l; }
    lens[n_ex - 1] += (3 - total % 3) % 3;
    total += (3 - total % 3) % 3;
    // coding sequence in transcript direction: ATG + stop-free codons, then stop + 60 nt UTR
    std::string coding;
    coding.reserve(total + 63);
    coding = "ATG";
    while (coding.size() < total) {
      char c[3];
      do {
        for (auto& x : c) x = B[rng.below(4)];
      } while (c[0] == 'T' && ((c[1] == 'A' && (c[2] == 'A' || c[2] == 'G')) || (c[1] == 'G' && c[2] == 'A')));
      coding.append(c, 3);
    }
    static const char* stops[3] = {"TAA", "TAG", "TGA"};
    std::string tail = stops[rng.below(3)];
    for (int i = 0; i < 60; ++i) tail.push_back(B[rng.below(4)]);
    // genomic layout
    const uint32_t gstart = pos;
    pos += 200;
    std::vector<HostExon> gex;  // genomic order, already with the three_prime_utr extension applied
    std::vector<uint8_t> ref;
    ref.assign(200, 'A');
    auto emit = [&](const std::string& s) { ref.insert(ref.end(), s.begin(), s.end()); pos += uint32_t(s.size()); };
    auto revc = [&](const std::string& s) { std::string r(s.rbegin(), s.rend()); for (auto& c : r) c = comp(c); return r; };
    std::vector<std::string> pieces;
    {
      size_t o = 0;
      for (uint32_t l : lens) { pieces.push_back(coding.substr(o, l)); o += l; }
    }
    for (uint32_t i = 0; i < n_ex; ++i) {
      const uint32_t txi = reverse ? n_ex - 1 - i : i;
      if (reverse && i == 0) {
        const uint32_t s = pos;
        emit(revc(tail));
        emit(revc(pieces[txi]));
        gex.push_back(HostExon{s, pos, 0});
      } else if (!reverse && i == n_ex - 1) {
        const uint32_t s = pos;
        emit(pieces[txi]);
        emit(tail);
        gex.push_back(HostExon{s, pos, 0});
      } else {
        const uint32_t s = pos;
        emit(reverse ? revc(pieces[txi]) : pieces[txi]);
        gex.push_back(HostExon{s, pos, 0});
      }
      if (i + 1 < n_ex) {
        const uint32_t il = rng.range(300, 5000);
        ref.insert(ref.end(), il, 'a');
        pos += il;
      }
    }
    ref.insert(ref.end(), 300, 'A');
    pos += 200;
    const uint32_t gend = pos;  // refseq covers [gstart, gend + 100)
    pos += 600;
    HostGene g;
    g.id = "ENSG" + std::to_string(100000000 + gi);
    g.name = "G" + std::to_string(gi);
    g.chrom = chrom;
    g.start = gstart;
    g.end = gend;
    HostTranscript t;
    t.id = "ENST" + std::to_string(100000000 + gi);
    t.reverse = reverse;
    // exons in transcript order; frame column of the CDS rows (only the first is used, :989-990)
    if (reverse) for (auto it = gex.rbegin(); it != gex.rend(); ++it) t.exons.push_back(*it);
    else t.exons = gex;
    g.transcripts.push_back(t);
    // variants per exon: SNVs, and with ins_var_frac / del_var_frac short insertions / deletions
    struct V { uint32_t pos; char alt; bool somatic; uint8_t hap; uint8_t kind; uint32_t len; std::string ins; };
    std::vector<V> vs;
    for (auto& e : gex) {
      const double kb = (e.end - e.start) / 1000.0;
      const uint32_t ng = rng.poisson(kb * sp.germline_per_kb), ns = rng.poisson(kb * sp.somatic_per_kb);
      for (uint32_t x = 0; x < ng + ns; ++x) {
        uint32_t vp = rng.range(e.start, e.end - 1);
        // (a variant exactly window_len after the start of a reverse-strand exon triggers the reference's
        // stale-column quirk: such transcripts go through the serial replay kernel)
        const double kind_draw = rng.uniform();
        const uint32_t ilen = rng.range(1, 6);
        const char r = char(ref[vp - gstart]);
        char a;
        do a = B[rng.below(4)]; while (a == r);
        V v{vp, a, x >= ng, uint8_t(rng.below(3)), MPH_SNV, 0, std::string()};
        char ib[6];
        for (auto& c : ib) c = B[rng.below(4)];
        // indels stay 36 nt clear of the exon ends: the reference's junction merge panics on a window pair whose
        // wild-type side is shorter than the window (usize underflow at :1775-1790), so a workload with indels at
        // the junctions cannot be run to completion by the reference itself
        const bool interior = vp >= e.start + 36 && vp + 42 < e.end;
        if (kind_draw < sp.ins_var_frac && interior) {
          v.kind = MPH_INS; v.len = ilen;
          v.ins.assign(1, r >= 'a' ? char(r - 32) : r);
          v.ins.append(ib, ilen);
        } else if (kind_draw >= sp.ins_var_frac && kind_draw < sp.ins_var_frac + sp.del_var_frac && interior) {
          v.kind = MPH_DEL; v.len = ilen;
        }
        vs.push_back(std::move(v));
      }
    }
    std::stable_sort(vs.begin(), vs.end(), [](const V& a, const V& b) { return a.pos < b.pos; });
    vs.erase(std::unique(vs.begin(), vs.end(), [](const V& a, const V& b) { return a.pos == b.pos; }), vs.end());
    std::vector<std::vector<HostVariant>> sites;
    for (auto& v : vs) {
      HostVariant hv;
      hv.pos = v.pos; hv.kind = v.kind; hv.alt = uint8_t(v.alt); hv.germline = !v.somatic; hv.len = v.len;
      if (v.kind == MPH_INS) hv.ins = v.ins;
      sites.push_back({hv});
    }
    // reads
    struct R { uint32_t start; uint32_t cig[3]; uint32_t ncig; uint32_t end; uint32_t id; };
    std::vector<R> rs;
    for (auto& e : gex) {
      const uint32_t a = e.start > L + gstart ? e.start - L : gstart, b = e.end;
      const uint32_t n = uint32_t(std::lround(sp.coverage * double(b - a) / L));
      for (uint32_t x = 0; x < n; ++x) {
        R r;
        r.start = rng.range(a, b);
        r.ncig = 1;
        r.cig[0] = L << 4;
        r.end = r.start + L;
        if (rng.uniform() < sp.indel_read_frac) {
          const uint32_t at = rng.range(10, L - 20), il = rng.range(1, 6);
          r.ncig = 3;
          if (rng.below(2)) { r.cig[0] = at << 4; r.cig[1] = (il << 4) | 1; r.cig[2] = (L - at - il) << 4; r.end = r.start + L - il; }
          else { r.cig[0] = at << 4; r.cig[1] = (il << 4) | 2; r.cig[2] = (L - at) << 4; r.end = r.start + L + il; }
        }
        r.id = uint32_t(rs.size());
        rs.push_back(r);
      }
    }
    std::sort(rs.begin(), rs.end(), [](const R& a, const R& b) { return a.start < b.start || (a.start == b.start && a.id < b.id); });
    // bases / qualities only for reads that can overlap a variant (the packer ships nothing else); each read draws
    // from its own generator, so the result does not depend on which reads are skipped
    std::vector<std::vector<uint8_t>> seqs, quals;
    std::vector<HostRead> hr(rs.size());
    seqs.reserve(rs.size() / 3);
    quals.reserve(rs.size() / 3);
    auto carries = [](const V& v, uint32_t hap, double take) { return v.somatic ? take < 0.3 : (v.hap == 2 || v.hap == hap); };
    for (size_t i = 0; i < rs.size(); ++i) {
      R& r = rs[i];
      HostRead& h = hr[i];
      auto lo = std::lower_bound(vs.begin(), vs.end(), r.start, [](const V& v, uint32_t p) { return v.pos < p; });
      const bool near_var = lo != vs.end() && lo->pos < r.start + L + 8;
      if (near_var || all_bases) {
        Rng rr(sp.seed ^ (0x9E3779B97F4A7C15ull * ((uint64_t(gi) << 32) | (r.id + 1u))));
        const uint32_t hap = rr.below(2);
        const double take = rr.uniform();
        // a read without a private indel takes the first short insertion / deletion allele it carries into its CIGAR
        const V* indel = nullptr;
        if (r.ncig == 1)
          for (auto vi = lo; vi != vs.end() && vi->pos + 20 < r.start + L; ++vi)
            if (vi->kind != MPH_SNV && vi->pos >= r.start + 10 && carries(*vi, hap, take)) { indel = &*vi; break; }
        if (indel) {
          const uint32_t at = indel->pos - r.start + 1;
          r.ncig = 3;
          r.cig[0] = at << 4;
          if (indel->kind == MPH_INS) { r.cig[1] = (indel->len << 4) | 1; r.cig[2] = (L - at - indel->len) << 4; r.end = r.start + L - indel->len; }
          else { r.cig[1] = (indel->len << 4) | 2; r.cig[2] = (L - at) << 4; r.end = r.start + L + indel->len; }
        }
        std::vector<uint8_t> s4((L + 1) / 2, 0), q(L);
        // walk the alignment: query index -> reference position
        uint32_t qi = 0, rp = r.start;
        for (uint32_t c = 0; c < r.ncig; ++c) {
          const uint32_t op = r.cig[c] & 15, len = r.cig[c] >> 4;
          if (op == 0) {
            for (uint32_t x = 0; x < len; ++x, ++qi, ++rp) {
              char base = char(ref[rp - gstart]);
              if (base >= 'a') base = char(base - 32);
              auto vi = std::lower_bound(vs.begin(), vs.end(), rp, [](const V& v, uint32_t p) { return v.pos < p; });
              if (vi != vs.end() && vi->pos == rp && vi->kind == MPH_SNV && carries(*vi, hap, take)) base = vi->alt;
              s4[qi >> 1] |= uint8_t(code4(base) << ((qi & 1) ? 0 : 4));
            }
          } else if (op == 1) {
            for (uint32_t x = 0; x < len; ++x, ++qi) {
              const char base = indel && indel->kind == MPH_INS ? indel->ins[1 + x] : B[rr.below(4)];
              s4[qi >> 1] |= uint8_t(code4(base) << ((qi & 1) ? 0 : 4));
            }
          } else {
            rp += len;
          }
        }
        for (uint32_t x = 0; x < L; ++x) q[x] = rr.uniform() < sp.lowq_frac ? uint8_t(rr.range(2, 9)) : uint8_t(rr.range(30, 40));
        seqs.push_back(std::move(s4));
        quals.push_back(std::move(q));
        h.seq4 = seqs.back().data();
        h.qual = quals.back().data();
      } else {
        h.seq4 = dummy_seq.data();
        h.qual = dummy_qual.data();
      }
      h.start = r.start; h.end = r.end; h.l_seq = L; h.n_cigar = r.ncig; h.cigar = rs[i].cig;
      h.qname_hash = (uint64_t(gi) << 32) | r.id;
    }
    for (size_t i = 0; i < rs.size(); ++i) hr[i].cigar = rs[i].cig;
    ref.resize(size_t(gend) + 100 - gstart, 'A');
    SynthGene sgene;
    sgene.gene = g;
    sgene.cds = gex;
    if (reverse) sgene.cds.front().start += 63;
    else sgene.cds.back().end -= 63;
    sgene.reads = std::move(hr);
    if (all_bases) {
      sgene.qnames.resize(rs.size());
      for (size_t i = 0; i < rs.size(); ++i) sgene.qnames[i] = "r" + std::to_string(gi) + "_" + std::to_string(rs[i].id);
    }
    sgene.sites = std::move(sites);
    sgene.ref = std::move(ref);
    sgene.max_read_len = L;
    sink(sgene);
  }
}

inline void synth_into(Packer& packer, const SynthParams& sp) {
  const bool normal_mode = packer.batch().mode == 1;
  synth_generate(sp, packer.batch().window_len, false, [&](SynthGene& sg) {
    if (normal_mode) {
      // the normal mode ignores three_prime_utr rows (src/normal_microphasing.rs:1340-1432): its exons are the CDS rows as they are
      HostTranscript& t = sg.gene.transcripts[0];
      t.exons.clear();
      if (t.reverse) for (auto it = sg.cds.rbegin(); it != sg.cds.rend(); ++it) t.exons.push_back(*it);
      else t.exons = sg.cds;
    }
    packer.add_gene(sg.gene, sg.reads, sg.max_read_len, sg.sites, std::move(sg.ref));
  });
}

}  // namespace mph
