// placeholder — filled in below
#pragma once
#include "oracle_common.hpp"
namespace oracle { namespace peptides {
inline void build(const std::string&, const std::string&, FILE*, size_t) { throw Failure("build: not implemented"); }
inline void filter(const std::string&, const std::string&, FILE*, const std::string&, const std::string&, const std::string&, const std::string&, size_t) { throw Failure("filter: not implemented"); }
}}
