#!/usr/bin/env python
"""Times one H.C matvec on the recorded pair list of several sites of the Cr2/SVP M=4000 chain
(workloads/cr2_svp_m4000_sites/, structure-only recordings, synthetic operator values):
how the same executor behaves from the tiny chain-end sectors to the heaviest mid-chain site."""
import glob
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2gpkg  # noqa: E402

b2g = b2gpkg.load()
ctx = b2g.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
if len(sys.argv) > 1:
    files = sys.argv[1:]
else:
    files = sorted(glob.glob(os.path.join(ROOT, "workloads", "cr2_svp_m4000_sites", "*.b2seq.gz")),
                   key=lambda f: int(re.search(r"_s(\d+)\.", f).group(1)))
    files.insert(4, os.path.join(ROOT, "workloads", "cr2_svp_m4000_site20.b2seq.gz"))
rows = []
for path in files:
    sf = b2g.load_seqfile(path)
    site = int(re.search(r"(?:_s|site)(\d+)\.", path).group(1))
    ops = torch.empty(max(sf.operand_doubles, 1), dtype=torch.float64, device=dev).normal_()
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, ops.data_ptr(), b2g.OPERANDS_DEVICE)
    c = torch.randn(sf.csize, dtype=torch.float64, device=dev)
    v = torch.zeros(sf.vsize, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if sf.flops > 1e11 else 20
    e0.record(stream)
    for _ in range(reps):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = plan.stats
    rows.append({"file": os.path.basename(path), "site": site, "pairs": sf.npairs, "psi": sf.csize, "operator_GB": 8e-9 * sf.operand_doubles,
                 "gflop": sf.flops * 1e-9, "ms": ms, "tflops": sf.flops / (ms * 1e-3) * 1e-12,
                 "GBps_if_streamed_once": 8e-9 * (sf.operand_doubles + 2 * sf.csize) / (ms * 1e-3),
                 "launches": int(st.launches)})
    print(json.dumps(rows[-1]), flush=True)
    if os.environ.get("SCAN_PROFILE"):
        v.zero_()
        torch.cuda.synchronize()
        prof = sorted(plan.profile(c.data_ptr(), v.data_ptr(), 1.0), key=lambda x: -x[2])
        print("   " + "; ".join(f"{n} {ms:.3f}ms u={u}" for n, fl, ms, u in prof[:12]), flush=True)
    plan.close()
    del ops, c, v
    torch.cuda.empty_cache()
tot_f, tot_t = sum(r["gflop"] for r in rows), sum(r["ms"] for r in rows)
print(json.dumps({"sites": len(rows), "sum_gflop": tot_f, "sum_ms": tot_t, "aggregate_tflops": tot_f / tot_t}))
