#!/bin/bash
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag 2>&1 | head -2
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
run() { local name=$1; shift
  timeout ${TMO:-300} "$@" > gpurun_out/dmrg_$name.log 2>&1; echo "exit $?" >> gpurun_out/dmrg_$name.log
  grep '"mode"' gpurun_out/dmrg_$name.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name', {k:d[k] for k in ('bond','t_gpu','e_gpu','t_plan','t_rotate','t_contract','t_contract_record','t_contract_plan','t_contract_upload','t_contract_download','resident_hit_gbytes','max_matvec_rel_err','max_rotate_rel_err','max_contract_rel_err')})"; }
run ab_huge_verify $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 500 --nsweeps 1 --threads $T --noise 1e-5 --gpu-contract --gpu-rotate --verify
run ab_huge $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 700 --nsweeps 2 --threads $T --noise 1e-5 --dsize 48 --gpu-contract --gpu-rotate
B2G_NO_HUGEPAGE=1 run ab_nohuge $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 700 --nsweeps 2 --threads $T --noise 1e-5 --dsize 48 --gpu-contract --gpu-rotate
