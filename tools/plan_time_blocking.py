#!/usr/bin/env python
"""Host-side planning time of a recorded blocking step (no GPU): b2g_tensor_product_execute with B2G_PLAN_ONLY."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2gpkg  # noqa: E402

b2g = b2gpkg.load()
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "workloads", "cr2_svp_m4000_blocking", "cr2_m4000_s20_call39.b2tp.gz")
tp = b2g.load_tpfile(path)
a_off, b_off, c_off, n_in, n_out = tp.offsets()
base_in, base_out = 1 << 40, 1 << 44  # fake, disjoint address ranges
terms = np.zeros(tp.nterms, dtype=b2g.TP_DTYPE)
terms["a"] = base_in + 8 * a_off
terms["b"] = base_in + 8 * b_off
terms["c"] = base_out + 8 * c_off
for k in ("am", "an", "bm", "bn", "cn", "conja", "conjb", "scale"):
    terms[k] = tp.t[k]
for it in range(3):
    t0 = time.perf_counter()
    st = b2g.tensor_product_plan(terms, b2g.DST_ZERO)
    dt = time.perf_counter() - t0
    print("terms %d entries %d clusters %d units %d plan %.1f ms (wall %.1f ms)" %
          (tp.nterms, st.entries, st.clusters, st.units, st.plan_seconds * 1e3, dt * 1e3))
if os.environ.get("B2G_PROF"):
    b2g.lib().b2g_prof_dump(None)
