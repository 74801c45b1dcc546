#!/bin/bash
# round 2, sixth GPU job: host-side wall-clock profile (B2G_PROF) of a Cr2 M=1000 run, the same run with a small
# resident budget (forces eviction / re-upload of written-through environments), and a real Cr2 M=4000 sweep
# under a memory watchdog (scratch in /dev/shm, stale partitions removed by the reference's minimal_disk_usage).
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
M4=${1:-4000}
export B2G_PROF=1
B2G_PROF_FILE=gpurun_out/r2g_prof_m1000.json timeout 600 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2g_cr2_m1000.log 2> gpurun_out/r2g_cr2_m1000.err
echo "cr2 m1000 rc=$?"; grep "Time sweep" gpurun_out/r2g_cr2_m1000.log; grep "^SWEEP" gpurun_out/r2g_cr2_m1000.log
rm -rf $S
B2G_RESIDENT_GB=3 B2G_PROF_FILE=gpurun_out/r2g_prof_m1000_evict.json timeout 600 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 1 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2g_cr2_m1000_evict.log 2> gpurun_out/r2g_cr2_m1000_evict.err
echo "cr2 m1000 small budget rc=$?"; grep "Time sweep" gpurun_out/r2g_cr2_m1000_evict.log; grep "^SWEEP" gpurun_out/r2g_cr2_m1000_evict.log
tail -1 gpurun_out/r2g_cr2_m1000_evict.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'resident' in k})"
rm -rf $S
# ---- M=4000 with a watchdog on host memory
B2G_PROF_FILE=gpurun_out/r2g_prof_m$M4.json timeout 2400 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond $M4 --nsweeps 1 --threads $T --noise 1e-5 --dsize 64 --scratch $S > gpurun_out/r2g_cr2_m$M4.log 2> gpurun_out/r2g_cr2_m$M4.err &
DPID=$!
( while kill -0 $DPID 2>/dev/null; do
    A=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
    U=$(df -BG --output=used /dev/shm | tail -1 | tr -dc 0-9)
    G=$(nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits | head -1)
    echo "$(date +%s) avail_gb=$A shm_gb=$U gpu_mib=$G" >> gpurun_out/r2g_mem_m$M4.log
    if [ "$A" -lt 10 ]; then echo "WATCHDOG: MemAvailable=$A GB, stopping the run" >> gpurun_out/r2g_mem_m$M4.log; kill $DPID; fi
    sleep 5
  done ) &
wait $DPID
echo "cr2 m$M4 rc=$?"
grep "Time sweep" -A8 gpurun_out/r2g_cr2_m$M4.log | tail -9; tail -1 gpurun_out/r2g_cr2_m$M4.log | cut -c1-3000
tail -3 gpurun_out/r2g_cr2_m$M4.err
sort -t= -k2 -n gpurun_out/r2g_mem_m$M4.log | head -1; awk '{print $3}' gpurun_out/r2g_mem_m$M4.log | sort -t= -k2 -n | tail -1
rm -rf $S
