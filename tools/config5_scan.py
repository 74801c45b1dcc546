#!/usr/bin/env python
"""BASELINE.json config 5 (synthetic 60-orbital FCIDUMP, SU2, M=8000, ParallelRuleQC over 8 ranks) on ONE GPU:
the H_eff pair list every rank records for the mid-chain site (tools/gen_config5.sh) is executed in turn on the same
device - per-rank matvec time, operator bytes and plan construction time.  The 8-GPU matvec is bounded below by the
slowest rank (plus the all-reduce of |sigma|); that maximum and the job throughput it implies are reported as an
emulation, not as a multi-GPU measurement."""
import glob
import json
import os
import re
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2gpkg  # noqa: E402

b2g = b2gpkg.load()
ctx = b2g.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
files = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "workloads", "_gen", "config5", "rand60_m8000_s29_P8_r*.b2seq.gz")))
rows = []
for path in files:
    t0 = time.perf_counter()
    sf = b2g.load_seqfile(path)
    t_load = time.perf_counter() - t0
    rank = int(re.search(r"_r(\d+)\.", path).group(1))
    ops = torch.empty(max(sf.operand_doubles, 1), dtype=torch.float64, device=dev).normal_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, ops.data_ptr(), b2g.OPERANDS_DEVICE)
    ctx.synchronize()
    t_plan = time.perf_counter() - t0
    c = torch.randn(sf.csize, dtype=torch.float64, device=dev)
    v = torch.zeros(sf.vsize, dtype=torch.float64, device=dev)
    for _ in range(2):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = plan.stats
    rows.append({"rank": rank, "pairs": sf.npairs, "psi": sf.csize, "operator_GB": 8e-9 * sf.operand_doubles,
                 "gflop": sf.flops * 1e-9, "ms": ms, "tflops": sf.flops / (ms * 1e-3) * 1e-12,
                 "GBps_if_streamed_once": 8e-9 * (sf.operand_doubles + 2 * sf.csize) / (ms * 1e-3),
                 "workspace_GB": 8e-9 * st.workspace_doubles, "launches": int(st.launches),
                 "plan_seconds": t_plan, "load_seconds": t_load})
    print(json.dumps(rows[-1]), flush=True)
    plan.close()
    del ops, c, v, plan
    torch.cuda.empty_cache()
if rows:
    tot = sum(r["gflop"] for r in rows)
    worst = max(r["ms"] for r in rows)
    print(json.dumps({"config": "synthetic 60-orbital FCIDUMP (seed 0), SU2, C1, M=8000, site 29, ClassicParallelMPO over 8 ranks",
                      "ranks_run": len(rows), "how": "every rank's list executed in turn on one B200",
                      "sum_gflop": tot, "sum_operator_GB": sum(r["operator_GB"] for r in rows),
                      "max_rank_ms": worst, "emulated_8gpu_tflops": tot / worst if len(rows) == 8 else None,
                      "one_gpu_all_ranks_ms": sum(r["ms"] for r in rows),
                      "one_gpu_tflops": tot / sum(r["ms"] for r in rows)}))
