#!/bin/bash
# round 2, final job on 2 GPUs: host-driver tests (incl. the two-rank NCCL run), bench at N=2, reference arm
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 1500 python -m pytest tests/test_host_driver.py -m gpu -q > gpurun_out/r2y_pytest_host.log 2>&1
echo "pytest host driver rc=$?"; tail -4 gpurun_out/r2y_pytest_host.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2y_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e']['value'], d['parity']['max_rel_err'], d['parity']['ok'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2y_bench_n2.err').read()[-2000:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/r2y_bench_reference.json
