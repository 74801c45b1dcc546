#!/bin/bash
# round 2, first GPU job: box probe, device-resident environments under --verify, energy parity with a fixed seed
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
{ echo "nproc=$T"; free -g | head -2; df -h /tmp /dev/shm . | cat; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/box_probe.txt 2>&1
cat gpurun_out/box_probe.txt
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 6 --threads 4 --noise 1e-6 --verify > gpurun_out/r2_n2_verify.log 2>&1
echo "n2 verify rc=$?"; tail -1 gpurun_out/r2_n2_verify.log | cut -c1-1800
timeout 300 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 10 --threads 4 --noise 1e-6 --compare > gpurun_out/r2_n2_compare.log 2>&1
echo "n2 compare rc=$?"; grep "^SWEEP" gpurun_out/r2_n2_compare.log; tail -1 gpurun_out/r2_n2_compare.log | cut -c1-600
timeout 600 $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 8 --threads $T --noise 1e-6 --compare > gpurun_out/r2_h10_compare.log 2>&1
echo "h10 compare rc=$?"; grep "^SWEEP" gpurun_out/r2_h10_compare.log; tail -1 gpurun_out/r2_h10_compare.log | cut -c1-600
timeout 900 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 6 --threads $T --noise 1e-5 --compare > gpurun_out/r2_c2_m500_compare.log 2>&1
echo "c2 compare rc=$?"; grep "^SWEEP\|Time sweep" gpurun_out/r2_c2_m500_compare.log | tail -30; tail -1 gpurun_out/r2_c2_m500_compare.log | cut -c1-1800
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
