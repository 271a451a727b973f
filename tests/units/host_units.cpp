// TEST INFRASTRUCTURE ONLY — unit checks of small host-side building blocks of the product:
// InlineStr (csrc/host/residue.hpp) and the host SHA-1 / record id (csrc/io/fmt_util.hpp).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../microphaser_b200/csrc/host/residue.hpp"
#include "../../microphaser_b200/csrc/core/phase_core.h"

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

template <size_t N>
static void inline_str_checks() {
  using S = mph::InlineStr<N>;
  const std::string small(N / 2, 'a'), edge(N, 'b'), big(N + 1, 'c'), huge(5 * N + 3, 'd');
  for (const std::string* src : {&small, &edge, &big, &huge}) {
    S s(*src);
    CHECK(s.size() == src->size());
    CHECK(std::string_view(s) == *src);
    CHECK(s.c_str()[s.size()] == 0);
    CHECK(std::string(s) == *src);
    S copy(s);
    CHECK(copy == s && copy.data() != s.data());
    S moved(std::move(copy));
    CHECK(moved == s);
    CHECK(copy.empty() && copy.c_str()[0] == 0);
    // every transition between the four sizes, by copy and by move assignment
    for (const std::string* dst : {&small, &edge, &big, &huge}) {
      S a(*src), b(*dst);
      a = b;
      CHECK(std::string_view(a) == *dst && a.c_str()[a.size()] == 0);
      S c(*src), d(*dst);
      c = std::move(d);
      CHECK(std::string_view(c) == *dst && c.c_str()[c.size()] == 0);
      CHECK(d.empty());
      d = *src;  // a moved-from string is reusable
      CHECK(std::string_view(d) == *src);
      S e(*src);
      e.assign(dst->data(), dst->size());
      CHECK(std::string_view(e) == *dst);
    }
    S& self = s;
    s = self;
    CHECK(std::string_view(s) == *src);
    s.clear();
    CHECK(s.empty() && s.c_str()[0] == 0);
  }
  S r;
  r.resize(16);
  for (int i = 0; i < 16; ++i) r[i] = char('a' + i);
  CHECK(std::string_view(r) == "abcdefghijklmnop");
  r.resize(N + 9);
  CHECK(r.size() == N + 9 && r[0] == 0 && r.c_str()[N + 9] == 0);
  std::vector<S> v;
  for (int i = 0; i < 100; ++i) v.emplace_back(std::string(size_t(i), char('A' + i % 26)));  // growth moves the elements
  for (int i = 0; i < 100; ++i) CHECK(std::string_view(v[size_t(i)]) == std::string(size_t(i), char('A' + i % 26)));
  CHECK(S("x") != S("y") && S("same") == S("same") && S() == S(""));
}

static void sha1_checks() {
  // FIPS 180-1 / RFC 3174 vectors and the block-boundary lengths of the padding
  struct { const char* msg; const char* hex; } kat[] = {
      {"", "da39a3ee5e6b4b0d3255bfef95601890afd80709"},
      {"abc", "a9993e364706816aba3e25717850c26c9cd0d89d"},
      {"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", "84983e441c3bd26ebaae4aa1f95129e5e54670f1"},
  };
  for (auto& k : kat) {
    mphfmt::Sha1 h;
    h.update(k.msg, strlen(k.msg));
    CHECK(h.hexdigest() == k.hex);
  }
  struct { size_t n; const char* hex; } rep[] = {
      {55, "c1c8bbdc22796e28c0e15163d20899b65621d65a"}, {56, "c2db330f6083854c99d4b5bfb6e8f29f201be699"},
      {63, "03f09f5b158a7a8cdad920bddc29b81c18a551f5"}, {64, "0098ba824b5c16427bd7a1122a5a442a25ec644d"},
      {65, "11655326c708d70319be2610e8a57d9a5b959d3b"}, {119, "ee971065aaa017e0632a8ca6c77bb3bf8b1dfc56"},
      {120, "f34c1488385346a55709ba056ddd08280dd4c6d6"}, {128, "ad5b3fdbcb526778c2839d2f151ea753995e26a0"},
  };
  for (auto& k : rep) {
    const std::string m(k.n, 'a');
    mphfmt::Sha1 whole;
    whole.update(m.data(), m.size());
    CHECK(whole.hexdigest() == k.hex);
    mphfmt::Sha1 pieces;  // the same message fed in uneven pieces
    for (size_t off = 0, step = 1; off < m.size(); off += step, step = step * 2 + 1) pieces.update(m.data() + off, std::min(step, m.size() - off));
    CHECK(pieces.hexdigest() == k.hex);
  }
  // record id = sha1("[65, 67, ...]" + transcript + offset)[0..15] + strand; long sequences cross the piece buffer
  for (size_t n : {size_t(0), size_t(1), size_t(27), size_t(64), size_t(65), size_t(300)}) {
    std::string seq;
    for (size_t i = 0; i < n; ++i) seq.push_back("ACGTN"[i % 5]);
    std::string msg = "[";
    for (size_t i = 0; i < n; ++i) { if (i) msg += ", "; msg += std::to_string(unsigned(uint8_t(seq[i]))); }
    msg += "]ENST00000400000.7" + std::to_string(123456 + n);
    mphfmt::Sha1 h;
    h.update(msg.data(), msg.size());
    const std::string want = h.hexdigest().substr(0, 15) + "R";
    CHECK(mphfmt::record_id(reinterpret_cast<const uint8_t*>(seq.data()), seq.size(), "ENST00000400000.7", 123456 + n, 'R') == want);
  }
}

// the streaming SHA-1 the kernels use for record ids (core/phase_core.h: word accumulator, explicit padding) against the
// host SHA-1 on the same message, over every message length around the block and padding boundaries, bytes that render
// as 1, 2 and 3 decimal digits included
static void device_sha1_checks() {
  for (size_t n = 0; n <= 140; ++n)
    for (size_t tl : {size_t(0), size_t(1), size_t(7), size_t(17), size_t(18), size_t(19), size_t(20), size_t(33)}) {
      std::vector<uint8_t> seq(n);
      for (size_t i = 0; i < n; ++i) seq[i] = uint8_t((i * 37 + n * 11 + tl) % 7 == 0 ? (i * 29 + n) & 0xFF : "ACGTacgtN"[(i + n) % 9]);
      std::string tx;
      for (size_t i = 0; i < tl; ++i) tx.push_back("ENST0123456789._"[(i * 5 + n) % 16]);
      const uint32_t offset = uint32_t(n * 7919u + tl * 104729u + (n % 3 == 0 ? 4000000000u : 0u));
      std::string msg = "[";
      for (size_t i = 0; i < n; ++i) { if (i) msg += ", "; msg += std::to_string(unsigned(seq[i])); }
      msg += "]" + tx + std::to_string(offset);
      mphfmt::Sha1 h;
      h.update(msg.data(), msg.size());
      const std::string hex = h.hexdigest().substr(0, 16);
      const uint64_t got = mph_record_id64(seq.data(), uint32_t(n), reinterpret_cast<const uint8_t*>(tx.data()), uint32_t(tl), offset);
      char buf[17];
      snprintf(buf, sizeof buf, "%016llx", (unsigned long long)got);
      CHECK(hex == buf);
    }
}

int main() {
  device_sha1_checks();
  inline_str_checks<23>();
  inline_str_checks<39>();
  inline_str_checks<3>();
  sha1_checks();
  if (g_fail) { fprintf(stderr, "%d checks failed\n", g_fail); return 1; }
  puts("host units ok");
  return 0;
}
