#!/bin/bash
# Real DMRG sweeps through block2's own driver with the GPU executor installed, each next to the
# stock CPU path on the same FCIDUMP / seed / schedule (per-sweep energy parity, sweep wall time).
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
for dav in device host; do
  $B/b2g_dmrg_su2 --fcidump $B/data/N2.STO3G.FCIDUMP --bond 250 --nsweeps 6 --threads $T --noise 1e-6 --davidson $dav --compare > gpurun_out/dmrg_n2_$dav.log 2>&1
  tail -1 gpurun_out/dmrg_n2_$dav.log
done
$B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 6 --threads $T --noise 1e-6 --compare > gpurun_out/dmrg_h10.log 2>&1
tail -1 gpurun_out/dmrg_h10.log
if [ "${C2:-1}" = "1" ]; then
$B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond ${C2M:-500} --nsweeps ${C2S:-3} --threads $T --noise 1e-5 --compare > gpurun_out/dmrg_c2.log 2>&1
tail -1 gpurun_out/dmrg_c2.log
fi
grep -h "SWEEP\|Time sweep" gpurun_out/dmrg_*.log | tail -60
