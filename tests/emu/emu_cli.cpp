// TEST INFRASTRUCTURE ONLY — the product's host code (parsers, packer, residue, writers) with
// the device side replaced by tests/emu/emu_pipeline.hpp. Never shipped, never loaded by the package.
#include "../../microphaser_b200/csrc/host/cli.hpp"
#include "emu_pipeline.hpp"

int main(int argc, char** argv) {
  return mph::cli_main(argc, argv, [](const mph::Batch& b) { return mphemu::phase(b); });
}
