// TEST INFRASTRUCTURE ONLY — the DEFLATE decoder the GPU runs per BGZF block (core/inflate_core.h) and the host reader's
// fast decoder (io/fast_inflate.hpp) against zlib:
// every block of the BGZF files named on the command line, plus streams zlib writes with stored / fixed / dynamic blocks at
// every level from inputs of several kinds, plus corrupted streams (must return an error, never crash or overrun).
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../microphaser_b200/csrc/core/inflate_core.h"
#include "../../microphaser_b200/csrc/io/fast_inflate.hpp"

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& in, int level, int strategy) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
  std::vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
  zs.next_in = const_cast<uint8_t*>(in.data()); zs.avail_in = uInt(in.size());
  zs.next_out = out.data(); zs.avail_out = uInt(out.size());
  deflate(&zs, Z_FINISH);
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return out;
}

int main(int argc, char** argv) {
  MphInflateScratch sc;
  static mphio::FastInflate fast;
  // 1. BGZF files, block by block
  size_t n_blocks = 0, n_bytes = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = fopen(argv[a], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
    std::vector<uint8_t> hdr(18), cbuf, want, got;
    while (fread(hdr.data(), 1, 18, f) == 18) {
      const unsigned xlen = hdr[10] | (hdr[11] << 8);
      std::vector<uint8_t> extra(xlen);
      memcpy(extra.data(), hdr.data() + 12, 6);
      if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, f) != xlen - 6) break;
      const unsigned bsize = extra[4] | (extra[5] << 8);
      const size_t clen = size_t(bsize) + 1 - 12 - xlen - 8;
      cbuf.resize(clen + 8);
      if (fread(cbuf.data(), 1, clen + 8, f) != clen + 8) break;
      uint32_t isize;
      memcpy(&isize, cbuf.data() + clen + 4, 4);
      want.assign(isize, 0);
      got.assign(isize + 16, 0xAB);
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      inflateInit2(&zs, -15);
      zs.next_in = cbuf.data(); zs.avail_in = uInt(clen); zs.next_out = want.data(); zs.avail_out = isize;
      const int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      CHECK(rc == Z_STREAM_END || isize == 0);
      const int st = mph_inflate_raw(cbuf.data(), uint32_t(clen), got.data(), isize, &sc);
      CHECK(st == MPH_INF_OK);
      CHECK(memcmp(got.data(), want.data(), isize) == 0);
      for (int t = 0; t < 16; ++t) CHECK(got[isize + t] == 0xAB);
      got.assign(isize + 16, 0xAB);
      CHECK(fast.run(cbuf.data(), clen, got.data(), isize));
      CHECK(memcmp(got.data(), want.data(), isize) == 0);
      for (int t = 0; t < 16; ++t) CHECK(got[isize + t] == 0xAB);
      ++n_blocks;
      n_bytes += isize;
    }
    fclose(f);
  }
  // 2. zlib-written streams: levels 0 (stored) .. 9, fixed codes, run-length and Huffman-only strategies
  unsigned seed = 12345;
  auto rnd = [&] { seed = seed * 1103515245u + 12345u; return (seed >> 16) & 0x7FFF; };
  size_t n_streams = 0;
  for (size_t n : {size_t(0), size_t(1), size_t(2), size_t(257), size_t(65280), size_t(40000)})
    for (int kind = 0; kind < 5; ++kind) {
      std::vector<uint8_t> in(n);
      for (size_t i = 0; i < n; ++i) {
        if (kind == 0) in[i] = uint8_t(rnd());                              // incompressible
        else if (kind == 1) in[i] = "ACGT"[rnd() & 3];                      // sequence-like
        else if (kind == 2) in[i] = uint8_t(i % 7 == 0 ? rnd() : 'x');      // long matches
        else if (kind == 3) in[i] = 0;                                      // one distance, maximal lengths
        else in[i] = uint8_t((i * 2654435761u) >> 13);                      // structured
      }
      for (int level : {0, 1, 6, 9})
        for (int strategy : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_RLE, Z_HUFFMAN_ONLY}) {
          const std::vector<uint8_t> c = deflate_raw(in, level, strategy);
          std::vector<uint8_t> got(n + 8, 0xCD);
          const int st = mph_inflate_raw(c.data(), uint32_t(c.size()), got.data(), uint32_t(n), &sc);
          CHECK(st == MPH_INF_OK);
          CHECK(n == 0 || memcmp(got.data(), in.data(), n) == 0);
          for (int t = 0; t < 8; ++t) CHECK(got[n + t] == 0xCD);
          {
            std::vector<uint8_t> g3(n + 8, 0xCD);
            CHECK(fast.run(c.data(), c.size(), g3.data(), n));
            CHECK(n == 0 || memcmp(g3.data(), in.data(), n) == 0);
            for (int t = 0; t < 8; ++t) CHECK(g3[n + t] == 0xCD);
            if (n) {
              CHECK(!fast.run(c.data(), c.size(), g3.data(), n - 1));
              CHECK(!fast.run(c.data(), c.size(), g3.data(), n + 1) || false);
              g3[n] = 0xCD;
              CHECK(!fast.run(c.data(), c.size() / 2, g3.data(), n) || c.size() < 2);
            }
          }
          ++n_streams;
          // wrong output size, truncated input, flipped bits: an error or (for a flip that keeps the stream valid) any
          // result, but never a write outside [0, n)
          if (n) {
            CHECK(mph_inflate_raw(c.data(), uint32_t(c.size()), got.data(), uint32_t(n - 1), &sc) != MPH_INF_OK);
            CHECK(mph_inflate_raw(c.data(), uint32_t(c.size()), got.data(), uint32_t(n + 1), &sc) != MPH_INF_OK);
            got[n + 1] = 0xCD;
            CHECK(mph_inflate_raw(c.data(), uint32_t(c.size() / 2), got.data(), uint32_t(n), &sc) != MPH_INF_OK || c.size() < 2);
            std::vector<uint8_t> bad = c;
            for (int t = 0; t < 8 && !bad.empty(); ++t) {
              bad[rnd() % bad.size()] ^= uint8_t(1u << (rnd() & 7));
              std::vector<uint8_t> g2(n + 8, 0xEF);
              mph_inflate_raw(bad.data(), uint32_t(bad.size()), g2.data(), uint32_t(n), &sc);
              for (int u = 0; u < 8; ++u) CHECK(g2[n + u] == 0xEF);
              fast.run(bad.data(), bad.size(), g2.data(), n);
              for (int u = 0; u < 8; ++u) CHECK(g2[n + u] == 0xEF);
            }
          }
        }
    }
  if (g_fail) { fprintf(stderr, "%d checks failed\n", g_fail); return 1; }
  printf("inflate ok: %zu BGZF blocks (%zu bytes), %zu zlib streams\n", n_blocks, n_bytes, n_streams);
  return 0;
}
