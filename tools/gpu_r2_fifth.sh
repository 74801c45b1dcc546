#!/bin/bash
# round 2, fifth GPU job: iadd block-descriptor form under --verify, Cr2 M=1000 and M=2000 sweeps (GPU arm, host timers)
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2f_pytest.log
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 6 --threads 4 --noise 1e-6 --verify --scratch $S > gpurun_out/r2f_n2_verify.log 2>&1
echo "n2 verify rc=$?"; tail -1 gpurun_out/r2f_n2_verify.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'err' in k or k in ('e_gpu','iadd_walks','iadd_entries')})"
timeout 600 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 300 --nsweeps 2 --threads $T --noise 1e-5 --verify --scratch $S > gpurun_out/r2f_c2_m300_verify.log 2>&1
echo "c2 verify rc=$?"; tail -1 gpurun_out/r2f_c2_m300_verify.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if 'err' in k or k in ('e_gpu','iadd_walks','iadd_entries','t_iadd')})"
timeout 900 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --scratch $S > gpurun_out/r2f_cr2_m1000_gpu.log 2>&1
echo "cr2 m1000 rc=$?"; grep "Time sweep" gpurun_out/r2f_cr2_m1000_gpu.log; grep "Time sweep" -A6 gpurun_out/r2f_cr2_m1000_gpu.log | tail -7; tail -1 gpurun_out/r2f_cr2_m1000_gpu.log | cut -c1-3000
timeout 1800 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 2000 --nsweeps 1 --threads $T --noise 1e-5 --dsize 48 --scratch $S > gpurun_out/r2f_cr2_m2000_gpu.log 2>&1
echo "cr2 m2000 rc=$?"; grep "Time sweep" gpurun_out/r2f_cr2_m2000_gpu.log; grep "Time sweep" -A6 gpurun_out/r2f_cr2_m2000_gpu.log | tail -7; tail -1 gpurun_out/r2f_cr2_m2000_gpu.log | cut -c1-3000
free -g | head -2; df -h /dev/shm | tail -1
