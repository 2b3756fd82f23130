"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol
include/hirm_engine.h declares; creating an engine without a GPU fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from hirm import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hirm_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hirm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    L = E.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert sorted(E.EXPORTS) == syms
    assert L.hirm_engine_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EngineError) as ei:
        E.Engine()
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hierarchical-irm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), (dirpath, f)      # no product file names the test infrastructure
