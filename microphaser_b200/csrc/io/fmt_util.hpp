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
    uint32_t w[80];
    for (int i = 0; i < 16; ++i) w[i] = (uint32_t(p[4 * i]) << 24) | (uint32_t(p[4 * i + 1]) << 16) | (uint32_t(p[4 * i + 2]) << 8) | p[4 * i + 3];
    for (int i = 16; i < 80; ++i) w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
    for (int i = 0; i < 80; ++i) {
      uint32_t f, k;
      if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
      else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
      else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
      else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
      uint32_t t = rol(a, 5) + f + e + k + w[i];
      e = d; d = c; c = rol(b, 30); b = a; a = t;
    }
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
  std::string hexdigest() {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; ++i) lb[i] = uint8_t(bits >> (56 - 8 * i));
    update(lb, 8);
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
  std::string s;
  s.reserve(5 * n + 48);
  debug_bytes(seq, n, s);
  s += transcript;
  s += std::to_string(offset);
  Sha1 sh;
  sh.update(s.data(), s.size());
  std::string id = sh.hexdigest().substr(0, 15);
  id.push_back(strand_initial);
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
