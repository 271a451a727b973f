// fmt_util.hpp — output-format primitives that are part of the reference's byte contract
// (SURVEY.md Appendix B): SHA-1 record ids, Rust `{:?}` rendering of a byte vector,
// ryu-style shortest round-trip f64 text (csv+serde), csv field quoting.
// Third-party crates restated (not vendored under /root/reference): sha1 0.6, ryu (via csv 1.x).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace mphfmt {

// ---------------------------------------------------------------- SHA-1 (FIPS 180-1)
struct Sha1 {
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  uint8_t buf[64];
  uint64_t len = 0;
  size_t fill = 0;

  static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
  void block(const uint8_t* p) {
    uint32_t w[16];  // rolling message schedule
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t(p[4 * i]) << 24) | (uint32_t(p[4 * i + 1]) << 16) | (uint32_t(p[4 * i + 2]) << 8) | p[4 * i + 3];
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    auto sched = [&](int i) { return w[i & 15] = rol(w[(i + 13) & 15] ^ w[(i + 8) & 15] ^ w[(i + 2) & 15] ^ w[i & 15], 1); };
    // one round with the five working variables passed in rotated order, so that no value has to be moved
    auto r0 = [&](uint32_t v, uint32_t& x, uint32_t y, uint32_t z, uint32_t& t, int i) { t += rol(v, 5) + ((x & y) | (~x & z)) + 0x5A827999u + (i < 16 ? w[i] : sched(i)); x = rol(x, 30); };
    auto r1 = [&](uint32_t v, uint32_t& x, uint32_t y, uint32_t z, uint32_t& t, int i) { t += rol(v, 5) + (x ^ y ^ z) + 0x6ED9EBA1u + sched(i); x = rol(x, 30); };
    auto r2 = [&](uint32_t v, uint32_t& x, uint32_t y, uint32_t z, uint32_t& t, int i) { t += rol(v, 5) + ((x & y) | (x & z) | (y & z)) + 0x8F1BBCDCu + sched(i); x = rol(x, 30); };
    auto r3 = [&](uint32_t v, uint32_t& x, uint32_t y, uint32_t z, uint32_t& t, int i) { t += rol(v, 5) + (x ^ y ^ z) + 0xCA62C1D6u + sched(i); x = rol(x, 30); };
#define MPH_SHA1_R5(R, i) R(a, b, c, d, e, i); R(e, a, b, c, d, i + 1); R(d, e, a, b, c, i + 2); R(c, d, e, a, b, i + 3); R(b, c, d, e, a, i + 4);
    MPH_SHA1_R5(r0, 0) MPH_SHA1_R5(r0, 5) MPH_SHA1_R5(r0, 10) MPH_SHA1_R5(r0, 15)
    MPH_SHA1_R5(r1, 20) MPH_SHA1_R5(r1, 25) MPH_SHA1_R5(r1, 30) MPH_SHA1_R5(r1, 35)
    MPH_SHA1_R5(r2, 40) MPH_SHA1_R5(r2, 45) MPH_SHA1_R5(r2, 50) MPH_SHA1_R5(r2, 55)
    MPH_SHA1_R5(r3, 60) MPH_SHA1_R5(r3, 65) MPH_SHA1_R5(r3, 70) MPH_SHA1_R5(r3, 75)
#undef MPH_SHA1_R5
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
  }
  void update(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    len += n;
    while (n) {
      size_t take = 64 - fill < n ? 64 - fill : n;
      memcpy(buf + fill, p, take);
      fill += take; p += take; n -= take;
      if (fill == 64) { block(buf); fill = 0; }
    }
  }
  void finish() {  // pads and absorbs the last block(s); h[] is the digest afterwards
    const uint64_t bits = len * 8;
    buf[fill++] = 0x80;
    if (fill > 56) {
      memset(buf + fill, 0, 64 - fill);
      block(buf);
      fill = 0;
    }
    memset(buf + fill, 0, 56 - fill);
    for (int i = 0; i < 8; ++i) buf[56 + i] = uint8_t(bits >> (56 - 8 * i));
    block(buf);
    fill = 0;
  }
  std::string hexdigest() {
    finish();
    static const char* hx = "0123456789abcdef";
    std::string s(40, '0');
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 4; ++j) {
        uint8_t v = uint8_t(h[i] >> (24 - 8 * j));
        s[8 * i + 2 * j] = hx[v >> 4];
        s[8 * i + 2 * j + 1] = hx[v & 15];
      }
    return s;
  }
};

// Rust `format!("{:?}", &Vec<u8>)` -> "[65, 84, 71]"
inline void debug_bytes(const uint8_t* p, size_t n, std::string& out) {
  out.push_back('[');
  char tmp[8];
  for (size_t i = 0; i < n; ++i) {
    if (i) out.append(", ");
    auto r = std::to_chars(tmp, tmp + 8, unsigned(p[i]));
    out.append(tmp, r.ptr);
  }
  out.push_back(']');
}

// record id = first 15 hex chars of sha1(format!("{:?}{}{}", seq, transcript_id, offset))
// followed by the first character of the strand name
// (reference src/microphasing.rs:667-675, src/common.rs:387-395).
inline std::string record_id(const uint8_t* seq, size_t n, const std::string& transcript, uint64_t offset, char strand_initial) {
  // the message is streamed into the hash in pieces; "[65, 84, 71]" is rendered 3 characters + ", " per byte at most
  Sha1 sh;
  char piece[5 * 64 + 2];
  size_t fill = 0;
  piece[fill++] = '[';
  for (size_t i = 0; i < n; ++i) {
    if (fill + 6 > sizeof piece) { sh.update(piece, fill); fill = 0; }
    if (i) { piece[fill++] = ','; piece[fill++] = ' '; }
    const unsigned v = seq[i];
    if (v >= 100) piece[fill++] = char('0' + v / 100);
    if (v >= 10) piece[fill++] = char('0' + v / 10 % 10);
    piece[fill++] = char('0' + v % 10);
  }
  if (fill + 1 > sizeof piece) { sh.update(piece, fill); fill = 0; }
  piece[fill++] = ']';
  sh.update(piece, fill);
  sh.update(transcript.data(), transcript.size());
  char num[24];
  auto r = std::to_chars(num, num + sizeof num, offset);
  sh.update(num, size_t(r.ptr - num));
  sh.finish();
  static const char* hx = "0123456789abcdef";
  std::string id(16, strand_initial);
  for (int q = 0; q < 15; ++q) id[q] = hx[(sh.h[q / 8] >> (28 - 4 * (q % 8))) & 15];
  return id;
}

// ryu::Buffer::format(f64) as used by csv's serde serializer: shortest round-trip digits,
// layout rules of ryu's pretty printer (src/pretty/mod.rs): plain decimal when the decimal
// exponent kk is in (-5, 16], otherwise d[.ddd]e[-]X; NaN / inf / -inf for non-finite.
inline std::string format_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  std::string out;
  if (std::signbit(v)) {
    out.push_back('-');
    v = -v;
  }
  if (v == 0.0) {
    out += "0.0";
    return out;
  }
  char sci[64];
  auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
  // d[.ddd]e[+-]XX  (shortest round-trip digits)
  std::string s(sci, r.ptr);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  int exp10 = std::stoi(s.substr(epos + 1));
  std::string digits;
  for (char c : mant)
    if (c != '.') digits.push_back(c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  int length = int(digits.size());
  int k = exp10 - (length - 1);  // value = digits * 10^k
  int kk = length + k;           // 10^(kk-1) <= v < 10^kk
  if (0 <= k && kk <= 16) {
    out += digits;
    out.append(size_t(k), '0');
    out += ".0";
  } else if (0 < kk && kk <= 16) {
    out.append(digits, 0, size_t(kk));
    out.push_back('.');
    out.append(digits, size_t(kk), std::string::npos);
  } else if (-5 < kk && kk <= 0) {
    out += "0.";
    out.append(size_t(-kk), '0');
    out += digits;
  } else {
    out.push_back(digits[0]);
    if (length > 1) {
      out.push_back('.');
      out.append(digits, 1, std::string::npos);
    }
    out.push_back('e');
    out += std::to_string(kk - 1);
  }
  return out;
}

// csv::Writer default QuoteStyle::Necessary with delimiter '\t': quote when the field
// contains the delimiter, a quote, CR or LF; an empty field is written bare except when
// it is the only field of a record. Quotes are doubled.
inline void csv_field(const std::string& f, char delim, std::string& out) {
  bool need = false;
  for (char c : f)
    if (c == delim || c == '"' || c == '\n' || c == '\r') {
      need = true;
      break;
    }
  if (!need) {
    out += f;
    return;
  }
  out.push_back('"');
  for (char c : f) {
    if (c == '"') out.push_back('"');
    out.push_back(c);
  }
  out.push_back('"');
}

}  // namespace mphfmt
