#!/usr/bin/env python
"""Host-side construction time of the two-phase tile plan of a recorded pair list (no GPU): b2g_debug_tiled_plan."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2gpkg  # noqa: E402

b2g = b2gpkg.load()
paths = sys.argv[1:] or [os.path.join(ROOT, "workloads", "cr2_svp_m4000_site20.b2seq.gz"),
                         os.path.join(ROOT, "workloads", "other_configs", "c2_m500_s12.b2seq.gz")]
L = b2g.lib()
for path in paths:
    sf = b2g.load_seqfile(path)
    d0, d1 = sf.as_batches(1 << 40)
    (b0, b1), keep = b2g._make_batches(d0, d1)
    for it in range(3):
        sec, units, launches, fp = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        rc = L.b2g_debug_tiled_plan(ctypes.byref(b0), ctypes.byref(b1), ctypes.byref(sec), ctypes.byref(units),
                                    ctypes.byref(launches), ctypes.byref(fp))
        assert rc == 0, L.b2g_last_error()
        print("%s: pairs %d units %d launches %d plan %.1f ms fingerprint %016x" % (
            os.path.basename(path), sf.npairs, units.value, launches.value, sec.value * 1e3, fp.value & (2**64 - 1)))
if os.environ.get("B2G_PROF"):
    L.b2g_prof_dump(None)
