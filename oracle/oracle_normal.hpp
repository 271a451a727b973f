// placeholder — filled in below
#pragma once
#include "oracle_common.hpp"
namespace oracle { namespace normal {
struct Writers { FastaWriter fasta; TsvWriter tsv; };
inline void phase(mphio::FastaIndexed&, std::istream&, mphio::VcfFile&, mphio::BamFile&, Writers&, uint64_t, bool) { throw Failure("normal: not implemented"); }
}}
