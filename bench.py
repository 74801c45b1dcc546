#!/usr/bin/env python
"""bench.py — Davidson H.C FP64 throughput on the Cr2/SVP M=4000 pair list.

One "step" = one matvec sigma = H_eff.c: zero sigma, replay every recorded GEMM pair of the
site through libb2g (block2 BatchGEMMSeq::operator(), core/batch_gemm.hpp:1570), and for N > 1
all-reduce the partial sigma over NCCL (ParallelTensorFunctions::operator(),
core/parallel_tensor_functions.hpp:51-55).  The pair list (shapes, windows, factors) was
recorded by the reference itself for data/CR2.SVP.FCIDUMP, SU2, two-site, M = 4000, site 20
(workloads/README.md); operator values are synthetic (seeded normal), as the contract allows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
WORKLOAD = os.path.join(ROOT, "workloads", "cr2_svp_m4000_site20.b2seq.gz")
WORKLOAD_NAME = "Cr2 SVP (data/CR2.SVP.FCIDUMP) SU2 two-site DMRG, M=4000, site 20 H_eff pair list"
METRIC = "Davidson H.C FP64 TFLOP/s at M=4000 (Cr2 SVP)"


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_replay(max_gflop: float, reps: int, warmup: int, threads: int) -> dict:
    """Times the reference's own BatchGEMMSeq::operator() (oracle/_ref/b2ref_su2 replay) on a
    bounded, seeded sample of the same pair list.  The only place bench.py executes oracle/."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b2ref_su2")
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} missing (built by __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads))
    scratch = "/tmp/b2ref_bench_scratch"
    os.makedirs(scratch, exist_ok=True)
    cmd = [exe, "replay", "--fcidump", WORKLOAD, "--threads", str(threads), "--reps", str(reps),
           "--warmup", str(warmup), "--scratch", scratch]
    if max_gflop > 0:
        cmd += ["--max-gflop", str(max_gflop)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, check=True).stdout
    line = [ln for ln in out.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


def sample_text(r: dict, threads: int) -> str:
    whole = r["pairs"] == r["pairs_total"]
    return (f"{'the full list: ' if whole else ''}{r['pairs']} of {r['pairs_total']} GEMM pairs ({r['flops'] / 1e9:.0f} GFLOP"
            f"{'' if whole else ', fixed-seed subset'}) through the reference BatchGEMMSeq::operator() (Tasked), "
            f"OpenBLAS 1 thread x {threads} OpenMP threads, {r['seconds_per_matvec']:.2f} s/pass, "
            f"{r['reps']} timed + {r['warmup']} warm-up passes, synthetic operator values (fixed seed)")


def run_reference(args) -> None:
    """The reference's own CPU executor on the SAME workload: the whole 44 415-pair list (about 3.5 s per pass on
    16 threads), every step one full matvec."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    r = reference_replay(0.0, max(args.steps, 1), max(args.warmup, 1), threads)
    sample = sample_text(r, threads)
    full = load_workload_header()
    line = {
        "impl": "reference", "metric": METRIC, "value": r["tflops"], "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_matvec"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(full),
        "cpu_baseline": {"value": r["tflops"], "unit": "TFLOP/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": r["tflops"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_workload_header():
    import b2gpkg
    return b2gpkg.load().load_seqfile(WORKLOAD)


def workload_config(full) -> dict:
    """The part of `config` both arms share (same workload, same units of work)."""
    return {"workload": WORKLOAD_NAME, "pairs": full.npairs, "wavefunction_doubles": full.csize,
            "operator_bytes": 8 * full.operand_doubles, "flop_per_step": full.flops}


def small_sector(b2g, torch, ctx, dev, stream, peaks, reps=20):
    path = os.path.join(ROOT, "workloads", "other_configs", "c2_m500_s12.b2seq.gz")
    sf = b2g.load_seqfile(path)
    gen = torch.Generator(device=dev).manual_seed(77)
    ops = torch.empty(max(sf.operand_doubles, 1), dtype=torch.float64, device=dev).normal_(0.0, 1.0, generator=gen)
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, ops.data_ptr(), b2g.OPERANDS_DEVICE)
    c = torch.randn(sf.csize, dtype=torch.float64, device=dev, generator=gen)
    v = torch.zeros(sf.vsize, dtype=torch.float64, device=dev)
    for _ in range(3):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 8.0 * (sf.operand_doubles + sf.csize + sf.vsize)
    hbm = float(peaks.get("hbm_gbs", 6550.0))
    out = {"workload": "C2 CAS cc-pVDZ SU2 M=500, site 12 H_eff pair list", "pairs": sf.npairs, "gflop": sf.flops * 1e-9,
           "ms_per_matvec": ms, "tflops": sf.flops / (ms * 1e-3) * 1e-12, "launches_per_matvec": int(plan.stats.launches),
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) * 1e-9 / hbm, "algorithmic_bytes": alg,
                        "flop_per_byte": sf.flops / alg}}
    plan.close()
    return out


def sigma_parity(np, torch, sf, full, ops, c_host, v_dev, world, dist, k=8, budget_flop=3.0e11, seed=7):
    """Checker, outside every timed region: sigma of K sampled blocks recomputed in plain numpy fp64 from the
    operands as they sit in HBM - for every pair that writes into the block,
    sigma_window += alpha1 * op(A1) . (alpha0 * c_window . op(B0)) (block2 batch_gemm.hpp:1634-1643) - and compared
    with what the CUDA path produced.  A block = a connected set of overlapping sigma windows of the SERIAL list
    (a sector block and its sub-windows), so the choice is the same on every rank; at N > 1 every rank
    recomputes its own pairs and the per-rank results are summed before the comparison with the all-reduced sigma."""
    def comps(P):
        lo = P["c1_off"].astype(np.int64)
        hi = lo + (P["m1"].astype(np.int64) - 1) * P["ldc1"] + P["n1"]
        order = np.argsort(lo, kind="stable")
        out, cur_lo, cur_hi = [], None, None
        for i in order:
            if P["m1"][i] == 0 or P["n1"][i] == 0:
                continue
            if cur_lo is None or lo[i] >= cur_hi:
                if cur_lo is not None:
                    out.append((cur_lo, cur_hi))
                cur_lo, cur_hi = int(lo[i]), int(hi[i])
            else:
                cur_hi = max(cur_hi, int(hi[i]))
        if cur_lo is not None:
            out.append((cur_lo, cur_hi))
        return out, lo, hi
    regions, flo, fhi = comps(full.p)
    fl = full.pair_flops()
    cost = [float(fl[(flo >= a) & (fhi <= b)].sum()) for a, b in regions]
    # the heaviest block always (70 % of the FLOPs of the M=4000 list sit in one block), the others by a seeded draw
    rng = np.random.default_rng(seed)
    heavy = int(np.argmax(cost))
    chosen, spent, drawn = [heavy], cost[heavy], 0.0
    for j in rng.permutation(len(regions)):
        if len(chosen) < k and j != heavy and cost[j] > 0 and drawn + cost[j] <= budget_flop:
            chosen.append(int(j))
            spent, drawn = spent + cost[j], drawn + cost[j]
    P = sf.p
    _, lo, hi = comps(P)
    b0o, a1o = sf.operand_offsets()
    view = np.lib.stride_tricks.as_strided
    cache = {}

    def block(off, rows, cols, ld):
        ext = (rows - 1) * ld + cols
        key = (int(off), int(ext))
        if key not in cache:
            cache[key] = ops[key[0]:key[0] + key[1]].cpu().numpy()
        return view(cache[key], (rows, cols), (8 * ld, 8))
    worst, checked, npairs, skipped = 0.0, 0, 0, 0
    for j in chosen:
        a, b = regions[j]
        ref = np.zeros(b - a)
        idx = np.nonzero((lo < b) & (hi > a))[0]
        for i in idx:
            if lo[i] < a or hi[i] > b:
                skipped += 1
                continue
            m0, n0, k0, m1 = int(P["m0"][i]), int(P["n0"][i]), int(P["k0"][i]), int(P["m1"][i])
            if m0 == 0 or n0 == 0 or m1 == 0:
                continue
            assert P["ta0"][i] == 0 and P["tb1"][i] == 0
            A = view(c_host[int(P["a0_off"][i]):], (m0, k0), (8 * int(P["lda0"][i]), 8))
            B = block(b0o[i], n0, k0, int(P["ldb0"][i])).T if P["tb0"][i] else block(b0o[i], k0, n0, int(P["ldb0"][i]))
            W = P["alpha0"][i] * (A @ B)
            A1 = block(a1o[i], m0, m1, int(P["lda1"][i])).T if P["ta1"][i] else block(a1o[i], m1, m0, int(P["lda1"][i]))
            win = view(ref[int(lo[i] - a):], (m1, n0), (8 * int(P["ldc1"][i]), 8))
            win += P["alpha1"][i] * (A1 @ W)
            npairs += 1
        cache.clear()
        if world > 1:
            t = torch.from_numpy(ref).to(v_dev.device)
            dist.all_reduce(t)
            ref = t.cpu().numpy()
        got = v_dev[a:b].cpu().numpy()
        den = float(np.linalg.norm(ref))
        err = float(np.linalg.norm(got - ref)) / den if den > 0 else float(np.linalg.norm(got))
        worst, checked = max(worst, err), checked + 1
    return {"blocks_checked": checked, "blocks_total": len(regions), "pairs_recomputed_this_rank": npairs,
            "pairs_outside_serial_blocks": skipped, "gflop_recomputed": spent * 1e-9, "max_rel_err": worst,
            "tolerance": 1e-11, "ok": bool(checked >= min(k, len(regions)) and worst <= 1e-11 and skipped == 0),
            "how": "numpy fp64 recomputation of sampled sigma blocks from the device operands; "
                   "seeded choice identical on every rank; summed over ranks before comparing with the all-reduced sigma"}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def fp64_gemm_peak(torch, dev) -> float:
    """cuBLAS DGEMM 8192^3 TFLOP/s on this GPU, best of 5: the FP64 roofline denominator
    (MEASURED_PEAKS.json carries only HBM and bf16 figures)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) * 1e-12


def run_b200(args) -> None:
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1:  # the numpy parity checker of every rank shares the host cores
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(max(1, host_threads() // world_env)))
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, host_threads() // world_env)))
    import numpy as np
    import torch
    import torch.distributed as dist
    import b2gpkg
    b2g = b2gpkg.load()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL prints its version banner on stdout at communicator creation; stdout (fd 1) is pointed at
    # stderr for the whole run and restored for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = b2g.Context(local)
    if world > 1:
        uid = [b2g.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    full = b2g.load_seqfile(WORKLOAD)
    # N > 1: the pair list rank r records for the same site when the reference itself parallelises the MPO over
    # ParallelRuleQC (workloads/README.md).  Default: ClassicParallelMPO (parallel_mpo.hpp:32-148) - every term of the
    # serial list on exactly one rank (the per-rank lists partition the 44 415 pairs).  B2G_BENCH_SCHEME=new: ParallelMPO
    # (NewScheme), where the terms of Partial operators are repeated on every rank (1.72x the serial FLOPs at P = 8).
    scheme = os.environ.get("B2G_BENCH_SCHEME", "classic")
    rank_dir, rank_tag = (("cr2_svp_m4000_site20_ranks_classic", "classic_") if scheme == "classic"
                          else ("cr2_svp_m4000_site20_ranks", ""))
    rank_file = os.path.join(ROOT, "workloads", rank_dir, f"cr2_m4000_s20_{rank_tag}P{world}_r{rank}.b2seq.gz")
    sharding = "serial list"
    if world > 1 and os.path.exists(rank_file):
        sf = b2g.load_seqfile(rank_file)
        sharding = ("per-rank lists recorded by the reference under ParallelRuleQC, "
                    + ("ClassicParallelMPO (each term on one rank; Partial operators reduced to their owner after blocking)"
                       if scheme == "classic" else "ParallelMPO NewScheme (Partial-operator terms repeated on every rank)"))
    elif world > 1:
        sf = full.shard(rank, world)
        sharding = "serial list split by left-operator index (emulated ParallelRuleQC ownership)"
    else:
        sf = full
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ops = torch.empty(max(sf.operand_doubles, 1), dtype=torch.float64, device=dev)
    ops.normal_(0.0, 1.0, generator=gen)
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, ops.data_ptr(), b2g.OPERANDS_DEVICE)
    gc = torch.Generator(device=dev).manual_seed(99)  # same c on every rank
    c = torch.randn(sf.csize, dtype=torch.float64, device=dev, generator=gc)
    v = torch.zeros(sf.vsize, dtype=torch.float64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.synchronize()

    def step():
        with torch.cuda.stream(stream):
            v.zero_()
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
        if world > 1:
            ctx.allreduce_sum(v.data_ptr(), sf.vsize)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    peak = fp64_gemm_peak(torch, dev) if rank == 0 else 0.0
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    launches = ctx.launches - launches0
    # kernel-only duration of the matvec launches (no memset / all-reduce), same stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            v.zero_()
        k0.record(stream)
        plan.matvec_dev(c.data_ptr(), v.data_ptr(), 1.0)
        k1.record(stream)
        ctx.synchronize()
        kms += k0.elapsed_time(k1)
    kms /= args.steps
    # per-launch durations (CUDA events between the launches, on the library's stream), averaged
    prof = {}
    for _ in range(max(args.steps, 3)):
        with torch.cuda.stream(stream):
            v.zero_()
        ctx.synchronize()
        for name, gfl, gms, units in plan.profile(c.data_ptr(), v.data_ptr(), 1.0):
            e = prof.setdefault(name, [gfl, 0.0, units, 0])
            e[1] += gms
            e[3] += 1
    kernels = [{"name": k, "gflop": e[0] * 1e-9, "ms": e[1] / e[3], "units": e[2],
                "tflops": e[0] / (e[1] / e[3] * 1e-3) * 1e-12 if e[1] > 0 else 0.0} for k, e in prof.items()]
    kernels.sort(key=lambda x: -x["ms"])
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    fl = torch.tensor([sf.flops], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    ms_step = float(t.item()) / args.steps
    executed_flops = float(fl.item())   # what all ranks actually multiplied (the parallel scheme repeats work)
    total_flops = full.flops            # the job: one H.C of the serial list, whatever N is
    value = total_flops / (ms_step * 1e-3) * 1e-12

    # end to end through the host-buffer C-ABI call (drop-in for BatchGEMMSeq::operator()):
    # c from host memory, sigma back to host, every step
    c_host = c.cpu().numpy().copy()
    v_host = np.zeros(sf.vsize)
    plan(c_host, v_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v_host[:] = 0.0
        plan(c_host, v_host)
    t_e2e = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_flops / float(te.item()) * 1e-12
    # cross-check of the two public entry points on the same list and c: host-buffer call (all-reduces inside
    # when a communicator exists) against one device-resident step (matvec + all-reduce)
    step()
    barrier()
    chk = float((torch.from_numpy(v_host).to(dev) - v).norm() / v.norm())
    # oracle check of sigma at the headline size (outside the timed regions)
    parity = sigma_parity(np, torch, sf, full, ops, c_host, v, world, dist)

    if rank == 0:
        st = plan.stats
        kernel_tflops = sf.flops / (kms * 1e-3) * 1e-12
        dom = kernels[0] if kernels else {"name": "matvec", "tflops": kernel_tflops, "ms": kms, "gflop": sf.flops * 1e-9}
        dom_share = dom["ms"] / sum(k["ms"] for k in kernels) if kernels else 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        alg_bytes = 8.0 * (sf.operand_doubles + sf.csize + sf.vsize)
        traffic = None  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        traffic_source = "profiles/r02_ncu_phase2_128x64_full.json (ncu --set full, 1-GPU list, this round's engine)"
        try:
            caps = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_phase2_128x64_full.json")))["kernels"]
            # the capture holds both operand layouts of the 128x64 phase-2 kernel: take the one that dominates here
            lay = "0>" if dom["name"].endswith("At") else "1>"  # At = phase2_kernel<Cfg, false>
            cap = next((k for k in caps if (lay + "(") in k["Kernel Name"]), caps[0])
            to_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = sum(float(cap[k].split()[0]) * to_b[cap[k].split()[1]]
                          for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(full),
                           l2="inputs (operator blocks, %.1f GB per rank) larger than L2" % (8e-9 * sf.operand_doubles),
                           parallelism="%d rank(s), %s, NCCL all-reduce of sigma" % (world, sharding),
                           executed_flop_all_ranks=executed_flops),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": 8 * sf.csize,
                    "d2h_bytes_per_step": 8 * sf.vsize, "path_check_rel": chk},
            "parity": parity,
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom["name"], "achieved": dom["tflops"], "peak": peak,
                         "unit": "TFLOP/s", "frac": dom["tflops"] / peak if peak else None, "traffic": traffic,
                         "traffic_source": traffic_source,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 tensor pipe, of measured); "
                                        "MEASURED_PEAKS.json has no FP64 entry; DMMA issue ceiling 37.05 TFLOP/s "
                                        "(profiles/r01_fp64_probe.json)",
                         "kernel_ms": dom["ms"], "kernel_gflop": dom["gflop"], "share_of_step": dom_share,
                         "whole_matvec": {"achieved": kernel_tflops, "frac": kernel_tflops / peak if peak else None,
                                          "ms": kms},
                         "algorithmic_bytes": alg_bytes,
                         "hbm_gbs_if_streamed_once": alg_bytes / (kms * 1e-3) * 1e-9,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "kernels": kernels,
            "plan": {"pairs": int(st.pairs), "arenas": int(st.arenas), "launches_per_matvec": int(st.launches),
                     "n_small": int(st.n_small), "n_large": int(st.n_large)},
        }
        if world == 1 and not args.no_cpu:
            try:
                threads = host_threads()
                r = reference_replay(0.0, 2, 1, threads)
                line["cpu_baseline"] = {"value": r["tflops"], "unit": "TFLOP/s", "cores": threads, "kind": "reference",
                                        "sample": sample_text(r, threads)}
            except Exception as exc:  # the baseline is reported, never needed by the GPU path
                line["cpu_baseline"] = {"value": None, "unit": "TFLOP/s", "cores": host_threads(), "kind": "reference",
                                        "sample": f"unavailable: {exc}"}
        if world == 1:
            # the small-sector regime (north_star: "achieved HBM GB/s for small ones"): the C2 CAS cc-pVDZ M=500
            # mid-chain list (149 004 pairs, dims <= 94, 9 GFLOP) is far below the ridge; its bound is HBM / launch
            # latency, reported as algorithmic bytes (distinct operator doubles + |c| + |sigma|) over kernel time
            try:
                line["small_sector"] = small_sector(b2g, torch, ctx, dev, stream, peaks)
            except Exception as exc:
                line["small_sector"] = {"unavailable": str(exc)}
            try:  # complete sweeps through block2's own driver: recorded runs of this round, not timed here
                rec = json.load(open(os.path.join(ROOT, "profiles", "r02_sweeps.json")))
                line["sweep"] = {"what": rec["what"], "source": "profiles/r02_sweeps.json (logs beside it)",
                                 "runs": [{"config": r["config"], "bond": r.get("bond"), "threads": r.get("threads"),
                                           "arms": [{"arm": a["arm"], "sweep_seconds": a["sweep_seconds"],
                                                     "tflop_per_sweep": a["tflop_per_sweep"]} for a in r["arms"]],
                                           "speedup_per_sweep": r.get("speedup_per_sweep")} for r in rec["runs"]]}
            except Exception:
                pass
        if world == 1:
            # the blocking step that feeds this H_eff (left_contract at the same site, same recording run):
            # HBM-bound kernels, reported beside the matvec with their own roofline
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import blocking_bench
                from types import SimpleNamespace
                bl = blocking_bench.run(SimpleNamespace(workload=blocking_bench.DEFAULT_WORKLOAD, steps=args.steps,
                                                        warmup=args.warmup, check_windows=8), ctx=ctx)
                line["blocking"] = {"workload": bl["config"]["workload"], "terms": bl["config"]["terms"],
                                    "ms": bl["ms_per_step"], "roofline": bl["roofline"], "parity": bl["parity"]}
            except Exception as exc:
                line["blocking"] = {"unavailable": str(exc)}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
