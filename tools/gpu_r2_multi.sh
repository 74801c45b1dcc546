#!/bin/bash
# usage: tools/gpu_r2_multi.sh N  (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2e_bench_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','e2e','gpu_launches')}); print(d['parity']); print(d['config']['parallelism'], d['config']['executed_flop_all_ranks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2e_bench_n$N.err').read()[-3000:])
PY
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_host_driver.py -m gpu -x -q -k two_rank > gpurun_out/r2e_pytest_two_rank.log 2>&1; echo "two-rank test rc=$?"; tail -5 gpurun_out/r2e_pytest_two_rank.log
fi
