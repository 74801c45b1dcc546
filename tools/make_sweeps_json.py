#!/usr/bin/env python
"""profiles/r02_sweeps.json from the b2g_dmrg logs of the GPU box (gpurun_out/): complete two-site DMRG sweeps through
block2's own driver, GPU arm (b2g_host::install) and, where it ran in the same job, the reference's CPU arm.
usage: tools/make_sweeps_json.py <label>=<log> ...   (label is free text, e.g. "Cr2 SVP M=1000")"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    arms, cur = [], None
    for line in open(path, errors="replace"):
        m = re.match(r"=== (.*) ===", line)
        if m:
            cur = {"arm": m.group(1), "sweep_seconds": [], "tflop_per_sweep": [], "timers": {}}
            arms.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"Time sweep =\s+([0-9.]+) \| ([0-9.]+) ([TGM])FLOP/SWP", line)
        if m:
            scale = {"T": 1.0, "G": 1e-3, "M": 1e-6}[m.group(3)]
            cur["sweep_seconds"].append(float(m.group(1)))
            cur["tflop_per_sweep"].append(float(m.group(2)) * scale)
        if line.startswith(" | T") and cur["sweep_seconds"]:
            for k, v in re.findall(r"(T[a-z]+) = +([0-9.]+)", line):
                cur["timers"][k] = float(v)  # cumulative over the run, as the reference prints them
    summary = None
    for line in open(path, errors="replace"):
        if line.startswith('{"mode"'):
            summary = json.loads(line)
    return arms, summary


def main():
    out = {"what": "complete two-site DMRG sweeps through block2's own sweep driver on the B200 box (16 host threads); "
                   "GPU arm = b2g_host::install + GPUDMRG, CPU arm = the unmodified reference (stock TensorFunctions); "
                   "recorded runs of this round, not timed inside bench.py", "runs": []}
    for arg in sys.argv[1:]:
        label, path = arg.rsplit("=", 1)
        arms, summary = parse(os.path.join(ROOT, path))
        run = {"config": label, "log": path, "arms": arms}
        if summary:
            run["bond"] = summary.get("bond")
            run["threads"] = summary.get("threads")
            run["e_gpu"] = summary.get("e_gpu")
            keep = ("t_gpu", "t_ref", "t_plan", "t_davidson", "t_rotate", "t_contract", "t_diag", "t_iadd",
                    "resident_peak_gbytes", "resident_evicted_gbytes", "resident_uploaded_gbytes",
                    "resident_downloaded_gbytes", "max_restart_sweep_diff", "restart_sweeps")
            run["gpu_arm_host_timers"] = {k: summary[k] for k in keep if k in summary}
        gpu = [a for a in arms if a["arm"].startswith("GPU path (")]
        cpu = [a for a in arms if a["arm"].startswith("reference CPU path (")]
        if gpu and cpu and gpu[0]["sweep_seconds"] and cpu[0]["sweep_seconds"]:
            n = min(len(gpu[0]["sweep_seconds"]), len(cpu[0]["sweep_seconds"]))
            run["speedup_per_sweep"] = [cpu[0]["sweep_seconds"][i] / gpu[0]["sweep_seconds"][i] for i in range(n)]
        out["runs"].append(run)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_sweeps.json"), "w"), indent=1)
    for r in out["runs"]:
        print(r["config"], [(a["arm"][:18], a["sweep_seconds"]) for a in r["arms"]], r.get("speedup_per_sweep"))


if __name__ == "__main__":
    main()
