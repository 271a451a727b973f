// inflate_core.h — raw DEFLATE (RFC 1951) decoder for one BGZF block, written so that the same code runs as one GPU thread per
// block (kernels/inflate_kernels.cu) and on the CPU (tests pin it to zlib, block by block).
//
// A BGZF file (SAM/BAM spec section 4.1) is a series of independent gzip members of at most 64 KiB of payload each; the
// reference reads them through htslib (rust-htslib `bam::IndexedReader`, src/main.rs:74). Because the blocks are independent,
// a whole batch of them inflates in parallel - one decoder per block, no state shared between blocks.
//
// Decoder: 64-bit bit buffer refilled bytewise; canonical Huffman codes decoded through a primary lookup table of
// 2^MPH_INF_LBITS entries for literal / length codes and 2^MPH_INF_DBITS for distance codes (entry = symbol << 4 | code
// length), longer codes through the count / symbol arrays of the canonical code (the method of zlib's contrib/puff, restated);
// stored, fixed and dynamic blocks. Every read and write is bounds-checked: a corrupt block returns an error, never writes
// outside its output slice.
#pragma once
#include <stdint.h>

#ifndef MPH_HD
#ifdef __CUDACC__
#define MPH_HD __host__ __device__ inline
#else
#define MPH_HD inline
#endif
#endif

#define MPH_INF_LBITS 10
#define MPH_INF_DBITS 8

enum {
  MPH_INF_OK = 0,
  MPH_INF_TRUNCATED = 1,   // ran out of input
  MPH_INF_OVERRUN = 2,     // more output than the block's ISIZE
  MPH_INF_BAD_BLOCK = 3,   // reserved block type / stored length mismatch
  MPH_INF_BAD_CODE = 4,    // invalid Huffman code / code lengths / symbol
  MPH_INF_BAD_DIST = 5,    // distance reaches before the start of the block
  MPH_INF_SHORT = 6        // stream ended before ISIZE bytes were produced
};

// decoding tables of one Huffman code; the caller provides the storage (shared memory on the device)
typedef struct {
  uint16_t* fast;     // 2^bits entries: symbol << 4 | length, 0 = longer than `bits` (or unused)
  uint16_t* symbol;   // symbols in canonical order
  uint16_t count[16]; // codes per length
  int bits;
} MphHuff;

typedef struct {
  const uint8_t* in;
  uint32_t in_len, in_pos;
  uint64_t bitbuf;
  int bitcnt;
} MphBits;

MPH_HD void mph_bits_refill(MphBits* b) {
  while (b->bitcnt <= 56 && b->in_pos < b->in_len) {
    b->bitbuf |= (uint64_t)b->in[b->in_pos++] << b->bitcnt;
    b->bitcnt += 8;
  }
}
// n <= 32 bits, LSB first; *ok = false when the input is exhausted
MPH_HD uint32_t mph_bits_get(MphBits* b, int n, bool* ok) {
  if (b->bitcnt < n) {
    mph_bits_refill(b);
    if (b->bitcnt < n) { *ok = false; return 0; }
  }
  const uint32_t v = (uint32_t)(b->bitbuf & (((uint64_t)1 << n) - 1));
  b->bitbuf >>= n;
  b->bitcnt -= n;
  return v;
}

// canonical code from the code lengths of n symbols; returns false for an over-subscribed set (incomplete sets are legal
// for the distance code of a block with a single distance, RFC 1951 3.2.7)
MPH_HD bool mph_huff_build(MphHuff* h, const uint8_t* length, int n) {
  for (int l = 0; l < 16; ++l) h->count[l] = 0;
  for (int s = 0; s < n; ++s) h->count[length[s]]++;
  const int fast_n = 1 << h->bits;
  for (int i = 0; i < fast_n; ++i) h->fast[i] = 0;
  if (h->count[0] == n) return true;  // no codes: legal, decoding any symbol fails
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= h->count[l];
    if (left < 0) return false;
  }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h->count[l]);
  for (int s = 0; s < n; ++s)
    if (length[s]) h->symbol[offs[length[s]]++] = (uint16_t)s;
  // primary table: the codes of length <= bits, bit-reversed (DEFLATE packs Huffman codes MSB first into an LSB-first stream)
  uint32_t code = 0;
  int idx = 0;
  for (int l = 1; l <= h->bits; ++l) {
    for (int k = 0; k < h->count[l]; ++k, ++idx, ++code) {
      uint32_t rev = 0;
      for (int t = 0; t < l; ++t) rev |= ((code >> t) & 1u) << (l - 1 - t);
      const uint16_t e = (uint16_t)((h->symbol[idx] << 4) | l);
      for (uint32_t x = rev; x < (uint32_t)fast_n; x += 1u << l) h->fast[x] = e;
    }
    code <<= 1;
  }
  return true;
}

// one symbol; -1 = invalid code, -2 = out of input
MPH_HD int mph_huff_decode(MphBits* b, const MphHuff* h) {
  if (b->bitcnt < 15) mph_bits_refill(b);
  const uint16_t e = h->fast[b->bitbuf & ((1u << h->bits) - 1)];
  if (e) {
    const int l = e & 15;
    if (l > b->bitcnt) return -2;
    b->bitbuf >>= l;
    b->bitcnt -= l;
    return e >> 4;
  }
  // longer than the primary table: walk the canonical code one bit at a time
  int code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; ++l) {
    if (b->bitcnt < 1) return -2;
    code |= (int)(b->bitbuf & 1);
    b->bitbuf >>= 1;
    b->bitcnt -= 1;
    const int cnt = h->count[l];
    if (code - cnt < first) return h->symbol[index + (code - first)];
    index += cnt;
    first += cnt;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// length / distance bases and extra bits (RFC 1951 3.2.5), order of the code-length code lengths (3.2.7): one copy for the
// host, one in the device's constant memory
#define MPH_INF_TABLES(Q, P)                                                                                                                        \
  Q uint16_t P##lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};      \
  Q uint8_t P##lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};                                    \
  Q uint16_t P##dbase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,                                  \
                             193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};                               \
  Q uint8_t P##dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};                         \
  Q uint8_t P##order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
MPH_INF_TABLES(static const, mph_inf_h_)
#ifdef __CUDACC__
MPH_INF_TABLES(static __constant__, mph_inf_d_)
#endif

// scratch of one decoder: 2 * 2^LBITS + 2 * 2^DBITS + 2 * (288 + 32) bytes
typedef struct {
  uint16_t lfast[1 << MPH_INF_LBITS];
  uint16_t dfast[1 << MPH_INF_DBITS];
  uint16_t lsym[288];
  uint16_t dsym[32];
} MphInflateScratch;

// inflates one raw DEFLATE stream of in_len bytes into exactly out_len bytes
MPH_HD int mph_inflate_raw(const uint8_t* in, uint32_t in_len, uint8_t* out, uint32_t out_len, MphInflateScratch* sc) {
#ifdef __CUDA_ARCH__
  const uint16_t* lbase = mph_inf_d_lbase; const uint8_t* lext = mph_inf_d_lext; const uint16_t* dbase = mph_inf_d_dbase;
  const uint8_t* dext = mph_inf_d_dext; const uint8_t* order = mph_inf_d_order;
#else
  const uint16_t* lbase = mph_inf_h_lbase; const uint8_t* lext = mph_inf_h_lext; const uint16_t* dbase = mph_inf_h_dbase;
  const uint8_t* dext = mph_inf_h_dext; const uint8_t* order = mph_inf_h_order;
#endif
  MphBits b;
  b.in = in; b.in_len = in_len; b.in_pos = 0; b.bitbuf = 0; b.bitcnt = 0;
  MphHuff lh, dh;
  lh.fast = sc->lfast; lh.symbol = sc->lsym; lh.bits = MPH_INF_LBITS;
  dh.fast = sc->dfast; dh.symbol = sc->dsym; dh.bits = MPH_INF_DBITS;
  uint32_t o = 0;
  bool ok = true;
  for (;;) {
    const uint32_t last = mph_bits_get(&b, 1, &ok);
    const uint32_t type = mph_bits_get(&b, 2, &ok);
    if (!ok) return MPH_INF_TRUNCATED;
    if (type == 0) {
      // stored: skip to the byte boundary, LEN, NLEN, bytes
      const int drop = b.bitcnt & 7;
      b.bitbuf >>= drop;
      b.bitcnt -= drop;
      const uint32_t len = mph_bits_get(&b, 16, &ok), nlen = mph_bits_get(&b, 16, &ok);
      if (!ok) return MPH_INF_TRUNCATED;
      if ((len ^ 0xFFFFu) != nlen) return MPH_INF_BAD_BLOCK;
      if (len > out_len - o) return MPH_INF_OVERRUN;
      for (uint32_t i = 0; i < len; ++i) {
        const uint32_t v = mph_bits_get(&b, 8, &ok);
        if (!ok) return MPH_INF_TRUNCATED;
        out[o++] = (uint8_t)v;
      }
    } else if (type == 3) {
      return MPH_INF_BAD_BLOCK;
    } else {
      uint8_t lengths[320];
      if (type == 1) {
        int s = 0;
        for (; s < 144; ++s) lengths[s] = 8;
        for (; s < 256; ++s) lengths[s] = 9;
        for (; s < 280; ++s) lengths[s] = 7;
        for (; s < 288; ++s) lengths[s] = 8;
        mph_huff_build(&lh, lengths, 288);
        for (s = 0; s < 30; ++s) lengths[s] = 5;
        mph_huff_build(&dh, lengths, 30);
      } else {
        const int nlen = (int)mph_bits_get(&b, 5, &ok) + 257, ndist = (int)mph_bits_get(&b, 5, &ok) + 1, ncode = (int)mph_bits_get(&b, 4, &ok) + 4;
        if (!ok) return MPH_INF_TRUNCATED;
        if (nlen > 286 || ndist > 30) return MPH_INF_BAD_CODE;
        int i = 0;
        for (; i < ncode; ++i) lengths[order[i]] = (uint8_t)mph_bits_get(&b, 3, &ok);
        if (!ok) return MPH_INF_TRUNCATED;
        for (; i < 19; ++i) lengths[order[i]] = 0;
        // the code-length code is decoded through the distance tables' storage (19 symbols, at most 7 bits)
        MphHuff ch;
        ch.fast = sc->dfast; ch.symbol = sc->dsym; ch.bits = 7;
        if (!mph_huff_build(&ch, lengths, 19)) return MPH_INF_BAD_CODE;
        i = 0;
        while (i < nlen + ndist) {
          const int sym = mph_huff_decode(&b, &ch);
          if (sym == -2) return MPH_INF_TRUNCATED;
          if (sym < 0) return MPH_INF_BAD_CODE;
          if (sym < 16) {
            lengths[i++] = (uint8_t)sym;
          } else {
            int rep, val = 0;
            if (sym == 16) {
              if (i == 0) return MPH_INF_BAD_CODE;
              val = lengths[i - 1];
              rep = 3 + (int)mph_bits_get(&b, 2, &ok);
            } else if (sym == 17) {
              rep = 3 + (int)mph_bits_get(&b, 3, &ok);
            } else {
              rep = 11 + (int)mph_bits_get(&b, 7, &ok);
            }
            if (!ok) return MPH_INF_TRUNCATED;
            if (i + rep > nlen + ndist) return MPH_INF_BAD_CODE;
            while (rep--) lengths[i++] = (uint8_t)val;
          }
        }
        if (lengths[256] == 0) return MPH_INF_BAD_CODE;  // no end-of-block code
        if (!mph_huff_build(&lh, lengths, nlen)) return MPH_INF_BAD_CODE;
        if (!mph_huff_build(&dh, lengths + nlen, ndist)) return MPH_INF_BAD_CODE;
      }
      for (;;) {
        int sym = mph_huff_decode(&b, &lh);
        if (sym == -2) return MPH_INF_TRUNCATED;
        if (sym < 0) return MPH_INF_BAD_CODE;
        if (sym < 256) {
          if (o >= out_len) return MPH_INF_OVERRUN;
          out[o++] = (uint8_t)sym;
        } else if (sym == 256) {
          break;
        } else {
          sym -= 257;
          if (sym >= 29) return MPH_INF_BAD_CODE;
          const uint32_t len = lbase[sym] + mph_bits_get(&b, lext[sym], &ok);
          const int ds = mph_huff_decode(&b, &dh);
          if (ds == -2) return MPH_INF_TRUNCATED;
          if (ds < 0 || ds >= 30) return MPH_INF_BAD_CODE;
          const uint32_t dist = dbase[ds] + mph_bits_get(&b, dext[ds], &ok);
          if (!ok) return MPH_INF_TRUNCATED;
          if (dist > o) return MPH_INF_BAD_DIST;
          if (len > out_len - o) return MPH_INF_OVERRUN;
          const uint8_t* src = out + o - dist;
          for (uint32_t i = 0; i < len; ++i) out[o + i] = src[i];  // overlapping copies repeat the pattern, byte by byte
          o += len;
        }
      }
    }
    if (last) break;
  }
  return o == out_len ? MPH_INF_OK : MPH_INF_SHORT;
}
