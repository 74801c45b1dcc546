#!/bin/bash
# round 2, tenth GPU job: blocking kernels with the rotating first warp, Cr2 M=1000 host profile with the pool kept,
# config 5 rank lists on one GPU, the real Cr2 M=4000 sweeps with per-site memory trace.
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 600 python -m pytest tests/test_blocking.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
RC0=$?; echo "pytest blocking rc=$RC0"; tail -2 gpurun_out/r2k_pytest.log
for W in call39 call18; do
timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_$W.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2k_blocking_$W.json 2> gpurun_out/r2k_blocking_$W.err
python -c "
import json; d=json.loads(open('gpurun_out/r2k_blocking_$W.json').read().strip().splitlines()[-1]); print('$W', d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_term']['frac'], d['parity'], d['plan_seconds_host'])"
done
timeout 600 python tools/config5_scan.py > gpurun_out/r2k_config5_scan.jsonl 2> gpurun_out/r2k_config5_scan.err
echo "config5 rc=$?"; tail -1 gpurun_out/r2k_config5_scan.jsonl; tail -2 gpurun_out/r2k_config5_scan.err
export B2G_PROF=1
B2G_PROF_FILE=gpurun_out/r2k_prof_m1000.json timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2k_cr2_m1000.log 2> gpurun_out/r2k_cr2_m1000.err
echo "cr2 m1000 rc=$?"; grep "Time sweep" gpurun_out/r2k_cr2_m1000.log; grep "^SWEEP" gpurun_out/r2k_cr2_m1000.log
rm -rf $S
M4=${1:-4000}
B2G_RESIDENT_GB=${3:-60} B2G_PROF_FILE=gpurun_out/r2k_prof_m$M4.json timeout ${2:-720} $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond $M4 --nsweeps 2 --noise-sweeps 1 --threads $T --noise 1e-5 --dsize 64 --scratch $S > gpurun_out/r2k_cr2_m$M4.log 2> gpurun_out/r2k_cr2_m$M4.err &
DPID=$!
( while kill -0 $DPID 2>/dev/null; do
    A=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
    U=$(df -BG --output=used /dev/shm | tail -1 | tr -dc 0-9)
    G=$(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits | head -1)
    echo "$(date +%s) avail_gb=$A shm_gb=$U gpu_mib=$G" >> gpurun_out/r2k_mem_m$M4.log
    if [ "$A" -lt 10 ]; then echo "WATCHDOG: MemAvailable=$A GB, stopping the run" >> gpurun_out/r2k_mem_m$M4.log; kill $DPID; fi
    sleep 5
  done ) &
wait $DPID
echo "cr2 m$M4 rc=$?"
grep "Time sweep" -A8 gpurun_out/r2k_cr2_m$M4.log | grep -v "^ --> \|^ <-- " | tail -24; tail -1 gpurun_out/r2k_cr2_m$M4.log | cut -c1-3000
grep -v "davidson n=\|site memory" gpurun_out/r2k_cr2_m$M4.err | tail -8
rm -rf $S
