#!/bin/bash
# round 2, thirteenth GPU job: linear-unit size of the blocking planner (A/B), ncu DRAM bytes of the blocking kernels,
# Cr2 M=2000 sweeps with --gpu-split
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
for LM in 2048 4096 8192 16384; do
  for W in call39 call18; do
    B2G_BLK_LINMAX=$LM timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_$W.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2n_blocking_${W}_$LM.json 2> gpurun_out/r2n_blocking_${W}_$LM.err
    python -c "
import json; d=json.loads(open('gpurun_out/r2n_blocking_${W}_$LM.json').read().strip().splitlines()[-1]); print('$W linmax $LM', round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['roofline']['per_term']['frac'],3), d['plan_seconds_host'])"
  done
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b2g_blocking -s 12 -c 8 --csv --log-file gpurun_out/r2n_blocking_launches.csv python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_call39.b2tp.gz --steps 2 --warmup 3 > gpurun_out/r2n_ncu_blocking.log 2>&1
echo "ncu blocking rc=$?"; wc -l gpurun_out/r2n_blocking_launches.csv
timeout 400 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 2000 --nsweeps 2 --noise-sweeps 1 --threads $T --noise 1e-5 --dsize 48 --gpu-split --scratch $S > gpurun_out/r2n_cr2_m2000.log 2> gpurun_out/r2n_cr2_m2000.err
echo "cr2 m2000 rc=$?"; grep "Time sweep" -A7 gpurun_out/r2n_cr2_m2000.log | grep "Time sweep\|Tsplt"; grep "^SWEEP" gpurun_out/r2n_cr2_m2000.log
rm -rf $S
