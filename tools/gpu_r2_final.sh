#!/bin/bash
# round 2, final GPU job: the whole -m gpu suite, smoke(), the bench line with the CPU baseline leg, memcheck
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?"; tail -4 gpurun_out/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['parity']['max_rel_err'], d['parity']['ok'])
    print(d['roofline']['frac'], d['roofline']['whole_matvec'], d['roofline']['traffic'], d['cpu_baseline'])
    print('blocking', d['blocking']['ms'], d['blocking']['roofline']['frac'], d['blocking']['roofline'].get('per_term',{}).get('frac'), d['blocking']['roofline'].get('traffic'))
    print('small', d['small_sector']['ms_per_matvec'], d['small_sector']['roofline']['frac'], 'sweep runs', len(d.get('sweep',{}).get('runs',[])))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2z_bench.err').read()[-3000:])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_blocking.py -m gpu -x -q -k "sub_windows or merged or random or bounded or synthetic or syevd" > gpurun_out/r2z_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r2z_memcheck.log | tail -5
