"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/b2g.h
declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2g.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2g_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(b2g):
    L = b2g.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(b2g.EXPORTS) == syms


def test_no_cpu_fallback_without_gpu(b2g):
    if b2g.lib().b2g_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(b2g.B2GError, match="no CPU fallback"):
        b2g.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "block2-preview_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cpp", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle-free", ""), f
