// peptide_kernels.cu — secondary path (reference src/peptides.rs): codon translation with a
// constant-memory codon table (to_protein :128-146, make_pairs :85-117) and the normal-peptidome
// membership test (`ref_set.contains(tumor_peptide)` :502,684) as an open-addressing hash set of
// 5-bit packed peptides probed on the device.
#include "peptide_kernels.cuh"
#include "kernel_common.cuh"

namespace mphk {

namespace {

// index = 16*b0 + 4*b1 + b2 with A=0 C=1 G=2 T=3; 'X' = stop (:108)
__constant__ char c_codon[65] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVVXYXYSSSSXCWCLFLF";

__device__ __forceinline__ int base2(uint8_t c, bool complement) {
  c &= 0xDF;  // to_ascii_uppercase for letters (:129)
  int b;
  switch (c) {
    case 'A': b = 0; break;
    case 'C': b = 1; break;
    case 'G': b = 2; break;
    case 'T': b = 3; break;
    default: return -1;
  }
  return complement ? 3 - b : b;
}

// one thread per sequence; frame +1: codons from the start, frame -1: codons of the reverse complement (:131-135)
__global__ void __launch_bounds__(256) k_translate(const uint8_t* __restrict__ nt, const uint64_t* __restrict__ off, const int8_t* __restrict__ frame,
                                                   uint64_t n, uint8_t* __restrict__ aa, const uint64_t* __restrict__ aa_off, uint8_t* __restrict__ bad) {
  const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint8_t* p = nt + off[s];
  const uint64_t len = off[s + 1] - off[s];
  const bool rc = frame[s] < 0;
  uint8_t* o = aa + aa_off[s];
  const uint64_t n_aa = aa_off[s + 1] - aa_off[s];
  uint8_t any_bad = 0;
  for (uint64_t j = 0; j < n_aa; ++j) {
    const uint64_t i = 3 * j;
    int b0, b1, b2;
    if (!rc) { b0 = base2(p[i], false); b1 = base2(p[i + 1], false); b2 = base2(p[i + 2], false); }
    else { b0 = base2(p[len - 1 - i], true); b1 = base2(p[len - 2 - i], true); b2 = base2(p[len - 3 - i], true); }
    if ((b0 | b1 | b2) < 0) { o[j] = '?'; any_bad = 1; }  // unknown codon: the reference unwraps an Err (:141)
    else o[j] = (uint8_t)c_codon[16 * b0 + 4 * b1 + b2];
  }
  bad[s] = any_bad;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// 5 bits per letter ('A'..'Z' -> 1..26), k <= 12; 0 = not representable (cannot equal any translated peptide)
__device__ __forceinline__ uint64_t pack_peptide(const uint8_t* p, uint32_t k) {
  uint64_t key = 0;
  for (uint32_t i = 0; i < k; ++i) {
    const uint8_t c = p[i];
    if (c < 'A' || c > 'Z') return 0;
    key = (key << 5) | (uint64_t)(c - 'A' + 1);
  }
  return key;
}

__global__ void __launch_bounds__(256) k_set_insert(const uint8_t* __restrict__ peptides, uint32_t k, uint64_t n, unsigned long long* table, uint64_t mask,
                                                    unsigned long long* n_distinct) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = pack_peptide(peptides + i * k, k);
  if (!key) return;
  uint64_t slot = mix64(key) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&table[slot], 0ull, (unsigned long long)key);
    if (prev == 0ull) { atomicAdd(n_distinct, 1ull); return; }
    if (prev == key) return;
    slot = (slot + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) k_set_probe(const uint8_t* __restrict__ queries, uint32_t k, uint64_t n, const unsigned long long* __restrict__ table,
                                                   uint64_t mask, uint8_t* __restrict__ hit) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = pack_peptide(queries + i * k, k);
  uint8_t h = 0;
  if (key) {
    uint64_t slot = mix64(key) & mask;
    for (;;) {
      const unsigned long long v = table[slot];
      if (v == key) { h = 1; break; }
      if (v == 0ull) break;
      slot = (slot + 1) & mask;
    }
  }
  hit[i] = h;
}

// distinct items back as bytes (for the bincode HashSet file of build_reference :183)
__global__ void __launch_bounds__(256) k_set_export(const unsigned long long* __restrict__ table, uint64_t slots, uint32_t k, uint8_t* out,
                                                    unsigned long long* cursor) {
  const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= slots) return;
  unsigned long long key = table[s];
  if (!key) return;
  const unsigned long long at = atomicAdd(cursor, 1ull);
  uint8_t* o = out + at * k;
  for (int i = (int)k - 1; i >= 0; --i) { o[i] = (uint8_t)('A' + (key & 31) - 1); key >>= 5; }
}

// ---- peptides longer than 12 letters (MHC-II runs use 13-25): the slots hold 1 + the index of a peptide in the set's own
// byte array, which stays on the device; equality is a byte comparison against that immutable array, so the insert needs
// only a 32-bit compare-and-swap on the slot. Any k works.
__device__ __forceinline__ uint64_t hash_bytes(const uint8_t* p, uint32_t k) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint32_t i = 0; i < k; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  return mix64(h);
}
__device__ __forceinline__ bool same_bytes(const uint8_t* a, const uint8_t* b, uint32_t k) {
  for (uint32_t i = 0; i < k; ++i)
    if (a[i] != b[i]) return false;
  return true;
}

__global__ void __launch_bounds__(256) k_set_insert_long(const uint8_t* __restrict__ peptides, uint32_t k, uint64_t n, uint32_t* table, uint64_t mask,
                                                         unsigned long long* n_distinct) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* me = peptides + i * k;
  uint64_t slot = hash_bytes(me, k) & mask;
  for (;;) {
    const uint32_t prev = atomicCAS(&table[slot], 0u, (uint32_t)(i + 1));
    if (prev == 0u) { atomicAdd(n_distinct, 1ull); return; }
    if (same_bytes(peptides + (uint64_t)(prev - 1) * k, me, k)) return;
    slot = (slot + 1) & mask;
  }
}

__global__ void __launch_bounds__(256) k_set_probe_long(const uint8_t* __restrict__ queries, uint32_t k, uint64_t n, const uint8_t* __restrict__ peptides,
                                                        const uint32_t* __restrict__ table, uint64_t mask, uint8_t* __restrict__ hit) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* me = queries + i * k;
  uint64_t slot = hash_bytes(me, k) & mask;
  uint8_t h = 0;
  for (;;) {
    const uint32_t v = table[slot];
    if (v == 0u) break;
    if (same_bytes(peptides + (uint64_t)(v - 1) * k, me, k)) { h = 1; break; }
    slot = (slot + 1) & mask;
  }
  hit[i] = h;
}

__global__ void __launch_bounds__(256) k_set_export_long(const uint32_t* __restrict__ table, uint64_t slots, const uint8_t* __restrict__ peptides, uint32_t k,
                                                         uint8_t* out, unsigned long long* cursor) {
  const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= slots) return;
  const uint32_t v = table[s];
  if (!v) return;
  const unsigned long long at = atomicAdd(cursor, 1ull);
  const uint8_t* src = peptides + (uint64_t)(v - 1) * k;
  uint8_t* o = out + at * k;
  for (uint32_t i = 0; i < k; ++i) o[i] = src[i];
}

}  // namespace

void launch_set_insert_long(const uint8_t* peptides, uint32_t k, uint64_t n, uint32_t* table, uint64_t mask, unsigned long long* n_distinct, cudaStream_t st) {
  if (n) MPH_LAUNCH(k_set_insert_long, ((unsigned)((n + 255) / 256), 256, 0, st), peptides, k, n, table, mask, n_distinct);
}
void launch_set_probe_long(const uint8_t* queries, uint32_t k, uint64_t n, const uint8_t* peptides, const uint32_t* table, uint64_t mask, uint8_t* hit,
                           cudaStream_t st) {
  if (n) MPH_LAUNCH(k_set_probe_long, ((unsigned)((n + 255) / 256), 256, 0, st), queries, k, n, peptides, table, mask, hit);
}
void launch_set_export_long(const uint32_t* table, uint64_t slots, const uint8_t* peptides, uint32_t k, uint8_t* out, unsigned long long* cursor,
                            cudaStream_t st) {
  if (slots) MPH_LAUNCH(k_set_export_long, ((unsigned)((slots + 255) / 256), 256, 0, st), table, slots, peptides, k, out, cursor);
}

void launch_translate(const uint8_t* nt, const uint64_t* off, const int8_t* frame, uint64_t n, uint8_t* aa, const uint64_t* aa_off, uint8_t* bad,
                      cudaStream_t st) {
  if (n) MPH_LAUNCH(k_translate, ((unsigned)((n + 255) / 256), 256, 0, st), nt, off, frame, n, aa, aa_off, bad);
}
void launch_set_insert(const uint8_t* peptides, uint32_t k, uint64_t n, unsigned long long* table, uint64_t mask, unsigned long long* n_distinct,
                       cudaStream_t st) {
  if (n) MPH_LAUNCH(k_set_insert, ((unsigned)((n + 255) / 256), 256, 0, st), peptides, k, n, table, mask, n_distinct);
}
void launch_set_probe(const uint8_t* queries, uint32_t k, uint64_t n, const unsigned long long* table, uint64_t mask, uint8_t* hit, cudaStream_t st) {
  if (n) MPH_LAUNCH(k_set_probe, ((unsigned)((n + 255) / 256), 256, 0, st), queries, k, n, table, mask, hit);
}
void launch_set_export(const unsigned long long* table, uint64_t slots, uint32_t k, uint8_t* out, unsigned long long* cursor, cudaStream_t st) {
  if (slots) MPH_LAUNCH(k_set_export, ((unsigned)((slots + 255) / 256), 256, 0, st), table, slots, k, out, cursor);
}

}  // namespace mphk
