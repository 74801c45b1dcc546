#!/bin/bash
# round 2, ninth GPU job: CTA-per-unit blocking kernels (parity + bench), eviction under --verify, host profile of
# Cr2 M=1000, the real Cr2 M=4000 sweeps with the out-of-memory retry, ncu --set full of the dominant kernel.
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blocking.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1
RC0=$?; echo "pytest rc=$RC0"; tail -3 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['max_rel_err'], d['parity']['ok'])
    print('blocking', d['blocking']['ms'], d['blocking']['roofline']['frac'], d['blocking']['roofline'].get('per_term',{}).get('frac'), d['blocking']['parity'])
    print('small', d['small_sector']['ms_per_matvec'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2j_bench.err').read()[-3000:])
PY
timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_call18.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2j_blocking_call18.json 2> gpurun_out/r2j_blocking_call18.err
python -c "
import json; d=json.loads(open('gpurun_out/r2j_blocking_call18.json').read().strip().splitlines()[-1]); print('call18', d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_term']['frac'], d['parity'], d['plan_seconds_host'])"
B2G_RESIDENT_GB=0.02 timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 300 --nsweeps 2 --threads $T --noise 1e-5 --verify --scratch $S > gpurun_out/r2j_c2_m300_evict_verify.log 2>&1
RC=$?; echo "c2 evict verify rc=$RC"; tail -1 gpurun_out/r2j_c2_m300_evict_verify.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'err' in k or 'evicted' in k or k in ('e_gpu',)})"
rm -rf $S
export B2G_PROF=1
B2G_PROF_FILE=gpurun_out/r2j_prof_m1000.json timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2j_cr2_m1000.log 2> gpurun_out/r2j_cr2_m1000.err
echo "cr2 m1000 rc=$?"; grep "Time sweep" gpurun_out/r2j_cr2_m1000.log; grep "^SWEEP" gpurun_out/r2j_cr2_m1000.log
rm -rf $S
if [ $RC0 -ne 0 ] || [ $RC -ne 0 ]; then echo "parity / eviction test failed: skipping M=4000"; exit 0; fi
M4=${1:-4000}
B2G_PROF_FILE=gpurun_out/r2j_prof_m$M4.json timeout ${2:-900} $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond $M4 --nsweeps 2 --noise-sweeps 1 --threads $T --noise 1e-5 --dsize 64 --scratch $S > gpurun_out/r2j_cr2_m$M4.log 2> gpurun_out/r2j_cr2_m$M4.err &
DPID=$!
( while kill -0 $DPID 2>/dev/null; do
    A=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
    U=$(df -BG --output=used /dev/shm | tail -1 | tr -dc 0-9)
    G=$(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits | head -1)
    echo "$(date +%s) avail_gb=$A shm_gb=$U gpu_mib=$G" >> gpurun_out/r2j_mem_m$M4.log
    if [ "$A" -lt 10 ]; then echo "WATCHDOG: MemAvailable=$A GB, stopping the run" >> gpurun_out/r2j_mem_m$M4.log; kill $DPID; fi
    sleep 5
  done ) &
wait $DPID
echo "cr2 m$M4 rc=$?"
unset B2G_PROF
grep "Time sweep" -A8 gpurun_out/r2j_cr2_m$M4.log | grep -v "^ --> \|^ <-- " | tail -24; tail -1 gpurun_out/r2j_cr2_m$M4.log | cut -c1-3000
tail -3 gpurun_out/r2j_cr2_m$M4.err
sort -t= -k2 -n gpurun_out/r2j_mem_m$M4.log | head -1; awk '{print $4}' gpurun_out/r2j_mem_m$M4.log | sort -t= -k2 -n | tail -1
rm -rf $S
ncu --set full --clock-control none --import-source on -k regex:phase2_kernel -s 32 -c 2 -o gpurun_out/r2j_prof_p2 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2j_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r2j_prof_p2.ncu-rep --page raw --csv > gpurun_out/r2j_prof_p2_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r2j
