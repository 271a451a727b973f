// peptide_kernels.cuh — launch entry points of the secondary (filter / build_reference) kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mphk {

void launch_translate(const uint8_t* nt, const uint64_t* off, const int8_t* frame, uint64_t n, uint8_t* aa, const uint64_t* aa_off, uint8_t* bad,
                      cudaStream_t st);
void launch_set_insert(const uint8_t* peptides, uint32_t k, uint64_t n, unsigned long long* table, uint64_t mask, unsigned long long* n_distinct,
                       cudaStream_t st);
void launch_set_probe(const uint8_t* queries, uint32_t k, uint64_t n, const unsigned long long* table, uint64_t mask, uint8_t* hit, cudaStream_t st);
void launch_set_export(const unsigned long long* table, uint64_t slots, uint32_t k, uint8_t* out, unsigned long long* cursor, cudaStream_t st);
// peptides longer than 12 letters: slots index the set's own byte array (peptide_kernels.cu)
void launch_set_insert_long(const uint8_t* peptides, uint32_t k, uint64_t n, uint32_t* table, uint64_t mask, unsigned long long* n_distinct, cudaStream_t st);
void launch_set_probe_long(const uint8_t* queries, uint32_t k, uint64_t n, const uint8_t* peptides, const uint32_t* table, uint64_t mask, uint8_t* hit,
                           cudaStream_t st);
void launch_set_export_long(const uint32_t* table, uint64_t slots, const uint8_t* peptides, uint32_t k, uint8_t* out, unsigned long long* cursor,
                            cudaStream_t st);

}  // namespace mphk
