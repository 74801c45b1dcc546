#!/bin/bash
# GPU box: parity tests, then block2's own DMRG driver with H.C, Davidson, blocking (left/right_contract)
# and renormalisation (left/right_rotate) on the GPU, --verify comparing every list with the reference's
# own executor on the same recorded list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
run() { # name binary args...
  local name=$1; shift
  timeout ${TMO:-600} "$@" > gpurun_out/dmrg_blk_$name.log 2>&1; echo "exit $?" >> gpurun_out/dmrg_blk_$name.log
  grep '"mode"' gpurun_out/dmrg_blk_$name.log | tail -1
}
run n2_verify $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 4 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --verify
run h10_verify $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 3 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --verify
run n2_compare $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 6 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --compare
run h10_compare $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 6 --threads $T --noise 1e-6 --gpu-contract --gpu-rotate --compare
if [ "${C2:-1}" = "1" ]; then
run c2_verify $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond ${C2M:-500} --nsweeps 2 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --verify
run c2_compare $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond ${C2M:-500} --nsweeps 3 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --compare
fi
grep -h "Time sweep" gpurun_out/dmrg_blk_*compare.log | tail -40
tail -3 gpurun_out/dmrg_blk_*.log | grep -i "error\|exit [1-9]\|b2g" | head
