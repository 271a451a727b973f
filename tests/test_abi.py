"""The C-ABI library loads and exports every symbol include/microphaser_gpu.h declares; without a
CUDA device the product path fails loudly instead of falling back to a CPU implementation."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "microphaser_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mph_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(product):
    import microphaser_b200 as m
    lib = ctypes.CDLL(product[0])
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in the header but not exported" % s
    assert sorted(m.EXPORTS) == syms


def test_no_cpu_fallback_without_device(product):
    import microphaser_b200 as m
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(m.MphError) as e:
        m.Context(0)
    assert e.value.code == m.MPH_ERR_CUDA


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under microphaser_b200/ or include/ may include or call it."""
    bad = []
    for root in ("microphaser_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, root)):
            if "_lib" in d or "_build" in d or "__pycache__" in d:
                continue
            for f in files:
                if f.endswith((".so", ".o", ".pyc")) or f == "microphaser":
                    continue
                text = open(os.path.join(d, f), errors="replace").read()
                if re.search(r'#include\s+"[^"]*oracle', text) or re.search(r"\boracle/", text) or "tests/emu" in text.replace("tests/emu)", ""):
                    if "emu" in text and root == "microphaser_b200" and f in ("phase_core.h", "layout.h", "cli.hpp"):
                        # these headers only *mention* the emulator in comments
                        if not re.search(r'#include\s+"[^"]*(oracle|emu)', text):
                            continue
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_packer_and_synthetic_batch_are_host_only(product):
    """Packing is host work and must run without a GPU; shapes follow SURVEY.md §8(d) config C2 (scaled down)."""
    import microphaser_b200 as m
    b = m.Batch.synthetic(n_transcripts=20, coverage=30.0, pin=False, seed=7)
    v = b.view()
    assert v.n_transcripts == 20 and v.n_segments == 160
    assert v.n_windows > 20 * 8 * 20
    assert v.n_reads > 20 * 8 * 30
    assert 0 < v.h2d_bytes < v.n_reads * 120
    b.close()
