#!/bin/bash
# round 2, twelfth GPU job: density-matrix eigenproblems through b2g_syevd (--gpu-split, cuSOLVER-backed): unit test,
# same-state parity on C2 M=500, Cr2 M=1000 profile, the real Cr2 M=4000 sweeps.
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1
RC0=$?; echo "pytest parity rc=$RC0"; tail -3 gpurun_out/r2m_pytest.log
if [ $RC0 -ne 0 ]; then echo "parity failed: skipping the sweeps"; exit 0; fi
timeout 400 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 6 --threads $T --noise 1e-5 --compare --restart-sweeps 2 --gpu-split --scratch $S > gpurun_out/r2m_c2_m500_compare_split.log 2>&1
echo "c2 compare rc=$?"; grep "RESTART\|^SWEEP" gpurun_out/r2m_c2_m500_compare_split.log; grep "Time sweep" gpurun_out/r2m_c2_m500_compare_split.log | tr '\n' ' '; echo
tail -1 gpurun_out/r2m_c2_m500_compare_split.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'syevd' in k or 'split' in k or 'diff' in k})"
rm -rf $S
export B2G_PROF=1
B2G_PROF_FILE=gpurun_out/r2m_prof_m1000.json timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --gpu-split --scratch $S > gpurun_out/r2m_cr2_m1000.log 2> gpurun_out/r2m_cr2_m1000.err
echo "cr2 m1000 rc=$?"; grep "Time sweep" -A7 gpurun_out/r2m_cr2_m1000.log | grep "Time sweep\|Tsplt"; grep "^SWEEP" gpurun_out/r2m_cr2_m1000.log
tail -1 gpurun_out/r2m_cr2_m1000.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'syevd' in k or 'split' in k})"
rm -rf $S
M4=${1:-4000}
B2G_RESIDENT_GB=${3:-60} B2G_PROF_FILE=gpurun_out/r2m_prof_m$M4.json timeout ${2:-800} $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond $M4 --nsweeps 2 --noise-sweeps 1 --threads $T --noise 1e-5 --dsize 64 --gpu-split --scratch $S > gpurun_out/r2m_cr2_m$M4.log 2> gpurun_out/r2m_cr2_m$M4.err &
DPID=$!
( while kill -0 $DPID 2>/dev/null; do
    A=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
    U=$(df -BG --output=used /dev/shm | tail -1 | tr -dc 0-9)
    G=$(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits | head -1)
    echo "$(date +%s) avail_gb=$A shm_gb=$U gpu_mib=$G" >> gpurun_out/r2m_mem_m$M4.log
    if [ "$A" -lt 10 ]; then echo "WATCHDOG: MemAvailable=$A GB, stopping the run" >> gpurun_out/r2m_mem_m$M4.log; kill $DPID; fi
    sleep 5
  done ) &
wait $DPID
echo "cr2 m$M4 rc=$?"
grep "Time sweep" -A8 gpurun_out/r2m_cr2_m$M4.log | grep -v "^ --> \|^ <-- " | tail -24; tail -1 gpurun_out/r2m_cr2_m$M4.log | cut -c1-3200
grep -v "davidson n=\|site memory" gpurun_out/r2m_cr2_m$M4.err | tail -8
rm -rf $S
