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


@pytest.mark.parametrize("name", ["n2_su2_m30_s8.b2seq", "n2_su2_m60_s4.b2seq", "h10_sz_m40_s4.b2seq"])
def test_tile_plan_regrouping_runs_on_the_host(b2g, name):
    """b2g_debug_tiled_plan: the two-phase regrouping of a recorded H.C pair list needs no device; the number
    of work units and launches is a function of the list alone (same twice)."""
    import ctypes
    sf = b2g.load_seqfile(os.path.join(ROOT, "tests", "golden", name))
    d0, d1 = sf.as_batches(1 << 40)
    (b0, b1), keep = b2g._make_batches(d0, d1)
    got = []
    for _ in range(2):
        sec, units, launches, fp = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        rc = b2g.lib().b2g_debug_tiled_plan(ctypes.byref(b0), ctypes.byref(b1), ctypes.byref(sec), ctypes.byref(units),
                                            ctypes.byref(launches), ctypes.byref(fp))
        assert rc == 0
        got.append((units.value, launches.value, fp.value))
    assert got[0] == got[1]
    assert got[0][0] >= 2 and 2 <= got[0][1] <= 32  # at least one unit per phase; at most 8 shapes x 2 layouts x 2 phases
