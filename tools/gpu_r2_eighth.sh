#!/bin/bash
# round 2, eighth GPU job: parity of the 2-D blocking units, PDL with dynamically claimed first units (A/B),
# eviction under --verify with a tiny resident budget, then the real Cr2 M=4000 sweeps under a memory watchdog.
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blocking.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest.log
for V in pdl nopdl; do
  if [ $V = nopdl ]; then export B2G_NO_PDL=1; else unset B2G_NO_PDL; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2i_bench_$V.json 2> gpurun_out/r2i_bench_$V.err
  echo "bench $V rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench_$V.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['max_rel_err'], d['parity']['ok'])
    print('blocking', d['blocking']['ms'], d['blocking']['roofline']['frac'], d['blocking']['roofline'].get('per_term',{}).get('frac'), d['blocking']['parity'])
    print('small', d['small_sector']['ms_per_matvec'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2i_bench_$V.err').read()[-3000:])
PY
done
unset B2G_NO_PDL
timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_call18.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2i_blocking_call18.json 2> gpurun_out/r2i_blocking_call18.err
python -c "
import json; d=json.loads(open('gpurun_out/r2i_blocking_call18.json').read().strip().splitlines()[-1]); print('call18', d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_term']['frac'], d['parity'], d['plan_seconds_host'])"
# eviction: every shadow block is evicted as soon as it may be; every list checked against the reference executor
B2G_RESIDENT_GB=0.02 timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 300 --nsweeps 2 --threads $T --noise 1e-5 --verify --scratch $S > gpurun_out/r2i_c2_m300_evict_verify.log 2>&1
RC=$?; echo "c2 evict verify rc=$RC"; tail -1 gpurun_out/r2i_c2_m300_evict_verify.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'err' in k or 'resident' in k or k in ('e_gpu',)})"
rm -rf $S
B2G_RESIDENT_GB=3 timeout 200 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 1 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2i_cr2_m1000_evict.log 2> gpurun_out/r2i_cr2_m1000_evict.err
RC2=$?; echo "cr2 m1000 small budget rc=$RC2"; grep "Time sweep" gpurun_out/r2i_cr2_m1000_evict.log; grep "^SWEEP" gpurun_out/r2i_cr2_m1000_evict.log
tail -1 gpurun_out/r2i_cr2_m1000_evict.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'resident' in k})"
rm -rf $S
if [ $RC -ne 0 ] || [ $RC2 -ne 0 ]; then echo "eviction test failed: skipping M=4000"; exit 0; fi
M4=${1:-4000}
export B2G_PROF=1
B2G_PROF_FILE=gpurun_out/r2i_prof_m$M4.json timeout ${2:-1200} $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond $M4 --nsweeps 2 --noise-sweeps 1 --threads $T --noise 1e-5 --dsize 64 --scratch $S > gpurun_out/r2i_cr2_m$M4.log 2> gpurun_out/r2i_cr2_m$M4.err &
DPID=$!
( while kill -0 $DPID 2>/dev/null; do
    A=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
    U=$(df -BG --output=used /dev/shm | tail -1 | tr -dc 0-9)
    G=$(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits | head -1)
    echo "$(date +%s) avail_gb=$A shm_gb=$U gpu_mib=$G" >> gpurun_out/r2i_mem_m$M4.log
    if [ "$A" -lt 10 ]; then echo "WATCHDOG: MemAvailable=$A GB, stopping the run" >> gpurun_out/r2i_mem_m$M4.log; kill $DPID; fi
    sleep 5
  done ) &
wait $DPID
echo "cr2 m$M4 rc=$?"
grep "Time sweep" -A8 gpurun_out/r2i_cr2_m$M4.log | grep -v "^ --> " | tail -20; tail -1 gpurun_out/r2i_cr2_m$M4.log | cut -c1-3000
tail -3 gpurun_out/r2i_cr2_m$M4.err
sort -t= -k2 -n gpurun_out/r2i_mem_m$M4.log | head -1; awk '{print $3}' gpurun_out/r2i_mem_m$M4.log | sort -t= -k2 -n | tail -1
rm -rf $S
