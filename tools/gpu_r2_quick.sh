#!/bin/bash
# quick kernel check: parity tests + bench (no CPU leg)
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_${TAG}_pytest.log 2>&1
echo "pytest parity rc=$?"; tail -4 gpurun_out/r2_${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_${TAG}_bench.json 2> gpurun_out/r2_${TAG}_bench.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_${TAG}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['parity']['max_rel_err'], d['parity']['ok'])
    print(d['roofline']['whole_matvec'], d['roofline']['frac'])
    for k in d['kernels'][:16]: print("%-22s %7.2f ms %8.1f GF %6.2f TF/s units %d"%(k['name'],k['ms'],k['gflop'],k['tflops'],k['units']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_${TAG}_bench.err').read()[-3000:])
PY
