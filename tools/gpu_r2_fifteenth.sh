#!/bin/bash
# round 2, fifteenth GPU job: CUDA-graph replay of small lists (parity, small-sector bench, H10 / C2 sweeps)
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blocking.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
for V in graph nograph; do
  if [ $V = nograph ]; then export B2G_NO_GRAPH=1; else unset B2G_NO_GRAPH; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2p_bench_$V.json 2> gpurun_out/r2p_bench_$V.err
  echo "bench $V rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2p_bench_$V.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['max_rel_err'], d['parity']['ok'])
    print('small', d['small_sector'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2p_bench_$V.err').read()[-3000:])
PY
done
unset B2G_NO_GRAPH
timeout 300 $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 5 --threads 8 --noise 1e-6 --compare --restart-sweeps 2 --dav-thrd 1e-10 --scratch /dev/shm/b2g_s1 > gpurun_out/r2p_h10_compare.log 2>&1
echo "h10 rc=$?"; grep "RESTART" gpurun_out/r2p_h10_compare.log; grep "Time sweep" gpurun_out/r2p_h10_compare.log | tr '\n' ' '; echo
timeout 900 python -m pytest tests/test_host_driver.py -m gpu -q -k "c2_cas or n2_device" > gpurun_out/r2p_pytest_host.log 2>&1
echo "pytest host (c2, n2) rc=$?"; tail -3 gpurun_out/r2p_pytest_host.log
