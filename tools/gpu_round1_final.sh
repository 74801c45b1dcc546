#!/bin/bash
# One GPU-box pass over everything round 1 claims: parity tests, smoke, bench line (with the blocking
# roofline), blocking bench on the full-blocking list, real DMRG sweeps (C2 M=1000, Cr2 M=500) with
# H.C + Davidson + blocking + renormalisation on the GPU, --verify and --compare.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
W=workloads/cr2_svp_m4000_blocking
B2G_VERBOSE=1 timeout 600 python tools/blocking_bench.py $W/cr2_m4000_s20_call18.b2tp.gz > gpurun_out/blocking_call18.json 2> gpurun_out/blocking_call18.err; cut -c1-260 gpurun_out/blocking_call18.json
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
run() { local name=$1; shift
  timeout ${TMO:-900} "$@" > gpurun_out/dmrg_$name.log 2>&1; echo "exit $?" >> gpurun_out/dmrg_$name.log
  grep '"mode"' gpurun_out/dmrg_$name.log | tail -1; }
run final_n2_verify $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 4 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --verify
run final_h10_verify $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 2 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --verify
run final_c2_verify $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 1 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --verify
run final_n2 $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 8 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --compare
run final_h10 $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 6 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --compare
run final_c2_m1000 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 1000 --nsweeps 3 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --compare
if [ "${CR2:-1}" = "1" ]; then
run final_cr2_m500 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 500 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --gpu-contract --gpu-rotate --compare
fi
grep -h "Time sweep\|^=== " gpurun_out/dmrg_final_c2_m1000.log gpurun_out/dmrg_final_cr2_m500.log | tail -20
