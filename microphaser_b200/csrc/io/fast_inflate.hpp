// fast_inflate.hpp — raw DEFLATE (RFC 1951) decoder of the file drivers' BGZF reader (host side).
//
// The BAM load of the file drivers is bound by inflate (zlib: ~200 MB/s per core). BGZF blocks are small (<= 64 KiB out),
// independent and their inflated size is known, which a decoder can exploit: one table look-up per symbol with the extra-bit
// counts and bases inside the table entry (11-bit literal / length table + sub-tables, 8-bit distance table + sub-tables),
// a 64-bit bit buffer refilled with one unaligned load, word-wise match copies, and a fast loop that runs while both the
// input and the output have a safety margin, followed by a bounds-checked loop for the rest. The technique is the one
// libdeflate documents; the code is written for this reader and checked against zlib block by block
// (tests/units/inflate_check.cpp). Any error makes the caller inflate that block with zlib instead.
#pragma once
#include <cstdint>
#include <cstring>

namespace mphio {

class FastInflate {
 public:
  // inflates in[0, in_len) into exactly out[0, out_len); false on any inconsistency (nothing outside out[0, out_len) is written)
  bool run(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    in_ = in; in_end_ = in + in_len; out0_ = out; out_ = out; out_end_ = out + out_len;
    bitbuf_ = 0; bitcnt_ = 0;
    for (;;) {
      need(3);
      if (bitcnt_ < 3) return false;
      const uint32_t last = take(1), type = take(2);
      if (type == 0) {
        // stored: byte boundary, LEN, NLEN, bytes; the bit buffer may hold whole bytes read ahead
        drop(bitcnt_ & 7);
        need(32);
        if (bitcnt_ < 32) return false;
        const uint32_t len = take(16), nlen = take(16);
        if ((len ^ 0xFFFFu) != nlen) return false;
        if (len > size_t(out_end_ - out_)) return false;
        uint32_t n = len;
        while (n && bitcnt_ >= 8) { *out_++ = uint8_t(take(8)); --n; }
        if (n > size_t(in_end_ - in_)) return false;
        memcpy(out_, in_, n);
        out_ += n; in_ += n;
      } else if (type == 1) {
        if (!fixed_ready_) build_fixed();
        if (!codes(fixed_l_, fixed_d_)) return false;
      } else if (type == 2) {
        if (!dynamic_header()) return false;
        if (!codes(ltab_, dtab_)) return false;
      } else {
        return false;
      }
      if (last) break;
    }
    return out_ == out_end_;
  }

 private:
  static constexpr int LBITS = 11, DBITS = 8;
  static constexpr uint32_t F_LITERAL = 0x8000, F_EOB = 0x4000, F_SUB = 0x2000, F_VALID = 0x1000;
  // entry: bits 0-4 code length (total), bits 8-11 extra bits / sub-table bits, flags in 12-15, bits 16-31 literal / base / sub-table offset

  const uint8_t *in_ = nullptr, *in_end_ = nullptr;
  uint8_t *out0_ = nullptr, *out_ = nullptr, *out_end_ = nullptr;
  uint64_t bitbuf_ = 0;
  int bitcnt_ = 0;
  uint32_t ltab_[(1 << LBITS) + 288 * 16], dtab_[(1 << DBITS) + 32 * 128];
  uint32_t fixed_l_[(1 << LBITS) + 16], fixed_d_[(1 << DBITS) + 16];
  bool fixed_ready_ = false;

  void need(int n) {  // bytewise, bounds-checked
    while (bitcnt_ < n && in_ < in_end_) { bitbuf_ |= uint64_t(*in_++) << bitcnt_; bitcnt_ += 8; }
  }
  uint32_t take(int n) {
    const uint32_t v = uint32_t(bitbuf_ & ((uint64_t(1) << n) - 1));
    bitbuf_ >>= n; bitcnt_ -= n;
    return v;
  }
  void drop(int n) { bitbuf_ >>= n; bitcnt_ -= n; }
  // the word-wise refill of the fast loop leaves stream bits above bitcnt_ in the buffer (they are the bytes at in_ and are
  // OR-ed in again by the next refill); the bytewise paths want them gone
  void clean() { bitbuf_ &= bitcnt_ >= 64 ? ~uint64_t(0) : ((uint64_t(1) << bitcnt_) - 1); }

  // builds a decoding table from code lengths; `kind` 0 = literal / length alphabet, 1 = distance alphabet
  static bool build(uint32_t* tab, int tbits, const uint8_t* len, int n, int kind) {
    static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    int count[16] = {0};
    for (int s = 0; s < n; ++s) count[len[s]]++;
    const int main_n = 1 << tbits;
    for (int i = 0; i < main_n; ++i) tab[i] = 0;
    if (count[0] == n) return true;
    int left = 1, maxlen = 0;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
      if (count[l]) maxlen = l;
    }
    uint32_t next_code[16], code = 0;
    for (int l = 1; l < 16; ++l) { next_code[l] = code; code = (code + uint32_t(count[l])) << 1; }
    const int sub_bits = maxlen > tbits ? maxlen - tbits : 0;
    int sub_next = main_n;
    for (int s = 0; s < n; ++s) {
      const int l = len[s];
      if (!l) continue;
      const uint32_t c = next_code[l]++;
      uint32_t rev = 0;
      for (int t = 0; t < l; ++t) rev |= ((c >> t) & 1u) << (l - 1 - t);
      uint32_t e;
      if (kind == 0) {
        if (s < 256) e = (uint32_t(s) << 16) | F_LITERAL | F_VALID | uint32_t(l);
        else if (s == 256) e = F_EOB | F_VALID | uint32_t(l);
        else if (s - 257 < 29) e = (uint32_t(lbase[s - 257]) << 16) | (uint32_t(lext[s - 257]) << 8) | F_VALID | uint32_t(l);
        else e = uint32_t(l);  // 286 / 287: not valid in a stream
      } else {
        if (s < 30) e = (uint32_t(dbase[s]) << 16) | (uint32_t(dext[s]) << 8) | F_VALID | uint32_t(l);
        else e = uint32_t(l);
      }
      if (l <= tbits) {
        for (uint32_t x = rev; x < uint32_t(main_n); x += 1u << l) tab[x] = e;
      } else {
        const uint32_t prefix = rev & uint32_t(main_n - 1);
        if (!(tab[prefix] & F_SUB)) {
          tab[prefix] = (uint32_t(sub_next) << 16) | (uint32_t(sub_bits) << 8) | F_SUB;
          for (int i = 0; i < (1 << sub_bits); ++i) tab[sub_next + i] = 0;
          sub_next += 1 << sub_bits;
        }
        const uint32_t base = tab[prefix] >> 16;
        for (uint32_t x = rev >> tbits; x < (1u << sub_bits); x += 1u << (l - tbits)) tab[base + x] = e;
      }
    }
    return true;
  }

  void build_fixed() {
    uint8_t len[288];
    int s = 0;
    for (; s < 144; ++s) len[s] = 8;
    for (; s < 256; ++s) len[s] = 9;
    for (; s < 280; ++s) len[s] = 7;
    for (; s < 288; ++s) len[s] = 8;
    build(fixed_l_, LBITS, len, 288, 0);
    for (s = 0; s < 30; ++s) len[s] = 5;
    build(fixed_d_, DBITS, len, 30, 1);
    fixed_ready_ = true;
  }

  bool dynamic_header() {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    need(14);
    if (bitcnt_ < 14) return false;
    const int nlen = int(take(5)) + 257, ndist = int(take(5)) + 1, ncode = int(take(4)) + 4;
    if (nlen > 286 || ndist > 30) return false;
    uint8_t cl[19] = {0};
    for (int i = 0; i < ncode; ++i) {
      need(3);
      if (bitcnt_ < 3) return false;
      cl[order[i]] = uint8_t(take(3));
    }
    uint32_t ctab[128 + 16];
    if (!build_plain(ctab, 7, cl, 19)) return false;
    uint8_t lengths[320];
    int i = 0;
    while (i < nlen + ndist) {
      need(7 + 7);
      const uint32_t e = ctab[bitbuf_ & 127];
      const int l = int(e & 31);
      if (!l || l > bitcnt_) return false;
      drop(l);
      const int sym = int(e >> 16);
      if (sym < 16) { lengths[i++] = uint8_t(sym); continue; }
      int rep, val = 0;
      if (sym == 16) {
        if (i == 0 || bitcnt_ < 2) return false;
        val = lengths[i - 1];
        rep = 3 + int(take(2));
      } else if (sym == 17) {
        if (bitcnt_ < 3) return false;
        rep = 3 + int(take(3));
      } else {
        if (bitcnt_ < 7) return false;
        rep = 11 + int(take(7));
      }
      if (i + rep > nlen + ndist) return false;
      while (rep--) lengths[i++] = uint8_t(val);
    }
    if (lengths[256] == 0) return false;
    return build(ltab_, LBITS, lengths, nlen, 0) && build(dtab_, DBITS, lengths + nlen, ndist, 1);
  }

  // plain symbol table (code-length code): entry = symbol << 16 | length, codes of at most tbits bits
  static bool build_plain(uint32_t* tab, int tbits, const uint8_t* len, int n) {
    int count[16] = {0};
    for (int s = 0; s < n; ++s) count[len[s]]++;
    for (int i = 0; i < (1 << tbits); ++i) tab[i] = 0;
    if (count[0] == n) return true;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
      if (count[l] && l > tbits) return false;
    }
    uint32_t next_code[16], code = 0;
    for (int l = 1; l < 16; ++l) { next_code[l] = code; code = (code + uint32_t(count[l])) << 1; }
    for (int s = 0; s < n; ++s) {
      const int l = len[s];
      if (!l) continue;
      const uint32_t c = next_code[l]++;
      uint32_t rev = 0;
      for (int t = 0; t < l; ++t) rev |= ((c >> t) & 1u) << (l - 1 - t);
      for (uint32_t x = rev; x < (1u << tbits); x += 1u << l) tab[x] = (uint32_t(s) << 16) | uint32_t(l);
    }
    return true;
  }

  static uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

  // the symbols of one block with the given tables
  bool codes(const uint32_t* lt, const uint32_t* dt) {
    const uint64_t lmask = (uint64_t(1) << LBITS) - 1, dmask = (uint64_t(1) << DBITS) - 1;
    // ---- fast loop: at least 8 readable input bytes and 258 + 8 writable output bytes on every iteration
    while (in_end_ - in_ >= 8 && out_end_ - out_ >= 266) {
      bitbuf_ |= load64(in_) << bitcnt_;
      in_ += (63 - bitcnt_) >> 3;
      bitcnt_ |= 56;  // 56..63 valid bits
      uint32_t e = lt[bitbuf_ & lmask];
      if (e & F_SUB) e = lt[(e >> 16) + ((bitbuf_ >> LBITS) & ((1u << ((e >> 8) & 15)) - 1))];
      if (!(e & F_VALID)) return false;
      drop(int(e & 31));
      if (e & F_LITERAL) {
        *out_++ = uint8_t(e >> 16);
        // up to three more literals out of the same refill (15 + 3 x 11 bits fit the 56 that are there; entries that
        // point to a sub-table are left to the next iteration)
        e = lt[bitbuf_ & lmask];
        if ((e & (F_LITERAL | F_SUB)) == F_LITERAL) {
          drop(int(e & 31));
          *out_++ = uint8_t(e >> 16);
          e = lt[bitbuf_ & lmask];
          if ((e & (F_LITERAL | F_SUB)) == F_LITERAL) {
            drop(int(e & 31));
            *out_++ = uint8_t(e >> 16);
            e = lt[bitbuf_ & lmask];
            if ((e & (F_LITERAL | F_SUB)) == F_LITERAL) {
              drop(int(e & 31));
              *out_++ = uint8_t(e >> 16);
            }
          }
        }
        continue;
      }
      if (e & F_EOB) { clean(); return true; }
      // length (<= 15 + 5 bits gone, >= 36 left), then distance (<= 15 + 13 bits)
      const int lx = int((e >> 8) & 15);
      const uint32_t len = (e >> 16) + uint32_t(bitbuf_ & ((1u << lx) - 1));
      drop(lx);
      uint32_t d = dt[bitbuf_ & dmask];
      if (d & F_SUB) d = dt[(d >> 16) + ((bitbuf_ >> DBITS) & ((1u << ((d >> 8) & 15)) - 1))];
      if (!(d & F_VALID)) return false;
      drop(int(d & 31));
      const int dx = int((d >> 8) & 15);
      const uint32_t dist = (d >> 16) + uint32_t(bitbuf_ & ((1u << dx) - 1));
      drop(dx);
      if (dist > size_t(out_ - out0_)) return false;
      const uint8_t* src = out_ - dist;
      uint8_t* dst = out_;
      out_ += len;
      if (dist >= 8) {
        // words; the last one may run past the match, never past the margin checked above
        const uint8_t* end = dst + len;
        do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < end);
      } else if (dist == 1) {
        memset(dst, *src, len);
      } else {
        for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
      }
    }
    // ---- careful loop for the tail: every read and write is checked
    clean();
    for (;;) {
      need(48);
      uint32_t e = lt[bitbuf_ & lmask];
      if (e & F_SUB) e = lt[(e >> 16) + ((bitbuf_ >> LBITS) & ((1u << ((e >> 8) & 15)) - 1))];
      if (!(e & F_VALID) || int(e & 31) > bitcnt_) return false;
      drop(int(e & 31));
      if (e & F_LITERAL) {
        if (out_ >= out_end_) return false;
        *out_++ = uint8_t(e >> 16);
        continue;
      }
      if (e & F_EOB) return true;
      const int lx = int((e >> 8) & 15);
      if (lx > bitcnt_) return false;
      const uint32_t len = (e >> 16) + uint32_t(bitbuf_ & ((1u << lx) - 1));
      drop(lx);
      uint32_t d = dt[bitbuf_ & dmask];
      if (d & F_SUB) d = dt[(d >> 16) + ((bitbuf_ >> DBITS) & ((1u << ((d >> 8) & 15)) - 1))];
      if (!(d & F_VALID) || int(d & 31) > bitcnt_) return false;
      drop(int(d & 31));
      const int dx = int((d >> 8) & 15);
      if (dx > bitcnt_) return false;
      const uint32_t dist = (d >> 16) + uint32_t(bitbuf_ & ((1u << dx) - 1));
      drop(dx);
      if (dist > size_t(out_ - out0_) || len > size_t(out_end_ - out_)) return false;
      const uint8_t* src = out_ - dist;
      for (uint32_t i = 0; i < len; ++i) out_[i] = src[i];
      out_ += len;
    }
  }
};

}  // namespace mphio
