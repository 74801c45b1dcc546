#!/usr/bin/env python
"""Blocking step (left_contract / right_contract) of Cr2 SVP M=4000 on one B200: the recorded term list of
workloads/cr2_svp_m4000_blocking/ executed by b2g_tensor_product_execute with device-resident operands.
Reports the kernel time (CUDA events on the context stream, inside the library call), the algorithmic
bytes (8 x (source elements read once + output elements written once)) and the HBM roofline fraction.
Parity at this size: recomputation of sampled output windows with numpy from the downloaded operands, and
linearity.  Usage: python tools/blocking_bench.py [workload.b2tp] [--steps K] [--warmup W]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


DEFAULT_WORKLOAD = os.path.join(ROOT, "workloads", "cr2_svp_m4000_blocking", "cr2_m4000_s20_call39.b2tp.gz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default=DEFAULT_WORKLOAD)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check-windows", type=int, default=24)
    args = ap.parse_args()
    print(json.dumps(run(args)))


def run(args, ctx=None):
    """args: .workload .steps .warmup .check_windows; returns the result line as a dict."""
    import torch
    import b2gpkg
    b2g = b2gpkg.load()
    tp = b2g.load_tpfile(args.workload)
    a_off, b_off, c_off, n_in, n_out = tp.offsets()
    torch.cuda.set_device(0)
    gen = torch.Generator(device="cuda").manual_seed(0)
    src = torch.empty(n_in, dtype=torch.float64, device="cuda")
    CH = 1 << 27
    for lo in range(0, n_in, CH):  # chunked: no 2x temporary
        src[lo:lo + CH].uniform_(-1.0, 1.0, generator=gen)
    out = torch.zeros(n_out, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx = ctx or b2g.Context(0)
    T = tp.t
    terms = np.zeros(tp.nterms, dtype=b2g.TP_DTYPE)
    terms["a"] = src.data_ptr() + 8 * a_off
    terms["b"] = src.data_ptr() + 8 * b_off
    terms["c"] = out.data_ptr() + 8 * c_off
    for k in ("am", "an", "bm", "bn", "cn", "conja", "conjb", "scale"):
        terms[k] = T[k]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6550.0))
    ms = []
    st = None
    for it in range(args.warmup + args.steps):
        st = ctx.tensor_product_execute(terms, b2g.OPERANDS_DEVICE, b2g.DST_ZERO)
        ctx.synchronize()
        if it >= args.warmup:
            ms.append(st.kernel_ms)
    kernel_ms = float(np.mean(ms))
    alg_bytes = int(st.bytes_in + st.bytes_out)
    gbs = alg_bytes / (kernel_ms * 1e-3) * 1e-9
    distinct = 8 * (int(tp.in_sizes.sum()) + int(st.bytes_out) // 8)
    traffic, traffic_source = None, None
    try:  # DRAM bytes of the kernels of one call, from the committed ncu launch list of the H_eff blocking list
        cap = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_blocking.json")))
        if "call39" in os.path.basename(args.workload):
            traffic = 1e9 * (cap["one_call"]["dram_read_GB"] + cap["one_call"]["dram_write_GB"])
            traffic_source = ("profiles/r02_ncu_blocking.json: dram__bytes_read.sum + dram__bytes_write.sum summed over "
                              "the four kernels of one call (ncu pass of the same command)")
    except Exception:
        pass

    # ---- parity at full size: sampled output windows recomputed on the host
    rows, cols = tp.window_shapes()
    rng = np.random.default_rng(0)
    key = c_off * 4 + 0  # windows are identified by their first element
    uniq, inv = np.unique(c_off, return_inverse=True)
    pick = rng.choice(len(uniq), size=min(args.check_windows, len(uniq)), replace=False)
    worst = 0.0
    for w in pick:
        members = np.nonzero(inv == w)[0]
        r, c, cn = int(rows[members[0]]), int(cols[members[0]]), int(T["cn"][members[0]])
        if r * c > 4_000_000:
            continue
        ref = np.zeros((r, c))
        for z in members:
            am, an, bm, bn = (int(T[k][z]) for k in ("am", "an", "bm", "bn"))
            A = src[int(a_off[z]):int(a_off[z]) + am * an].cpu().numpy().reshape(am, an)
            B = src[int(b_off[z]):int(b_off[z]) + bm * bn].cpu().numpy().reshape(bm, bn)
            ref += float(T["scale"][z]) * np.kron(A.T if T["conja"][z] else A, B.T if T["conjb"][z] else B)
        base = int(uniq[w])
        got = torch.as_strided(out, (r, c), (cn, 1), base).cpu().numpy()
        worst = max(worst, float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)))
    # bilinearity: doubling every source (environment blocks AND site blocks) quadruples the blocked operators
    first = out[:min(n_out, 1 << 24)].clone()
    src.mul_(2.0)
    ctx.tensor_product_execute(terms, b2g.OPERANDS_DEVICE, b2g.DST_ZERO)
    ctx.synchronize()
    lin = float(torch.linalg.norm(out[:first.numel()] - 4.0 * first) / torch.linalg.norm(first))
    line = {
        "metric": "blocking (left/right_contract) achieved HBM GB/s at M=4000 (Cr2 SVP), distinct bytes", "value": distinct / (kernel_ms * 1e-3) * 1e-9, "unit": "GB/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms, "dtype": "f64",
        "config": {"workload": os.path.basename(args.workload), "terms": tp.nterms, "clusters": int(st.clusters),
                   "units": int(st.units), "input_doubles": n_in, "output_doubles": n_out,
                   "l2": "operands (%.1f GB) larger than L2" % ((n_in + n_out) * 8e-9)},
        # SURVEY 8(d) convention: every distinct operand read once, every output written once.  The per-term count
        # (a source block that feeds k windows counted k times) is what the kernels stream and is kept beside it.
        "roofline": {"bound": "hbm", "kernel": "b2g_blocking_kernel", "achieved": distinct / (kernel_ms * 1e-3) * 1e-9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": distinct / (kernel_ms * 1e-3) * 1e-9 / hbm_peak,
                     "algorithmic_bytes": distinct,
                     "algorithmic_bytes_definition": "8 x (distinct source elements + output window elements)",
                     "per_term": {"bytes": alg_bytes, "achieved": gbs, "frac": gbs / hbm_peak,
                                  "definition": "8 x (source elements of every term + output window elements)"},
                     "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6550 GB/s"},
        "parity": {"sampled_windows_max_rel_err": worst, "bilinearity_rel_err": lin},
        "plan_seconds_host": st.plan_seconds, "launches_per_call": int(st.launches),
    }
    del src, out
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    main()
