#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
run() { local name=$1; shift
  timeout ${TMO:-900} "$@" > gpurun_out/dmrg_$name.log 2>&1; echo "exit $?" >> gpurun_out/dmrg_$name.log
  grep '"mode"' gpurun_out/dmrg_$name.log | tail -1; }
run res_n2_verify $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 4 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --verify
run res_c2_verify $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 2 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --verify
run res_n2 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 8 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --compare
run res_c2_m1000 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 1000 --nsweeps 3 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate
grep -h "Time sweep" gpurun_out/dmrg_res_c2_m1000.log | tail -3
