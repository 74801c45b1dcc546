#!/bin/bash
# round 2, second GPU job: new device paths under --verify, converged energy parity, bench N=1, Cr2 M=1000 sweeps
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 6 --threads 4 --noise 1e-6 --verify --scratch $S > gpurun_out/r2b_n2_verify.log 2>&1
echo "n2 verify rc=$?"; tail -1 gpurun_out/r2b_n2_verify.log | cut -c1-2200
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 300 --nsweeps 2 --threads $T --noise 1e-5 --verify --scratch $S > gpurun_out/r2b_c2_m300_verify.log 2>&1
echo "c2 verify rc=$?"; tail -1 gpurun_out/r2b_c2_m300_verify.log | cut -c1-2200
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 10 --threads 4 --noise 1e-6 --dav-thrd 1e-10 --compare --scratch $S > gpurun_out/r2b_n2_compare.log 2>&1
echo "n2 compare rc=$?"; grep "^SWEEP" gpurun_out/r2b_n2_compare.log; tail -1 gpurun_out/r2b_n2_compare.log | cut -c1-400
timeout 600 $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 8 --threads $T --noise 1e-6 --dav-thrd 1e-10 --compare --scratch $S > gpurun_out/r2b_h10_compare.log 2>&1
echo "h10 compare rc=$?"; grep "^SWEEP" gpurun_out/r2b_h10_compare.log; tail -1 gpurun_out/r2b_h10_compare.log | cut -c1-400
timeout 1200 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 16 --noise-sweeps 3 --conv 1e-9 --dav-thrd 1e-10 --threads $T --noise 1e-5 --compare --scratch $S > gpurun_out/r2b_c2_m500_converged.log 2>&1
echo "c2 converged rc=$?"; grep "^SWEEP" gpurun_out/r2b_c2_m500_converged.log; tail -1 gpurun_out/r2b_c2_m500_converged.log | cut -c1-2400
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','parity','cpu_baseline','gpu_launches')})
    print(d['roofline']['whole_matvec'], d['roofline']['frac'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2b_bench_n1.err').read()[-2000:])
PY
timeout 1500 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --compare --scratch $S > gpurun_out/r2b_cr2_m1000.log 2>&1
echo "cr2 m1000 rc=$?"; grep "^SWEEP\|Time sweep" gpurun_out/r2b_cr2_m1000.log; grep "Time sweep" -A6 gpurun_out/r2b_cr2_m1000.log | tail -7; tail -1 gpurun_out/r2b_cr2_m1000.log | cut -c1-2400
