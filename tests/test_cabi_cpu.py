"""CPU-side checks of the C-ABI library: it builds, loads, exports every declared symbol, and
fails loudly (no fallback) when no GPU is present."""
import ctypes
import os
import re

import pytest

import build_native
import sgv_native as nat

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(REPO, "include", "sgvamp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sgv_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_all_symbols():
    build_native.build()
    lib = nat.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing symbol " + n
    assert sorted(names) == sorted(nat.SYMBOLS)
    assert lib.sgv_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nat.SgvError) as e:
        nat.Handle(device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "sgvamp-py_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "sgvamp_oracle" not in src, f
