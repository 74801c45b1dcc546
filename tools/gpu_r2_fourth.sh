#!/bin/bash
# round 2, fourth GPU job: tile engine with ignore-src loads + chunked B fragments: parity, memcheck, bench, ncu
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest_parity.log 2>&1
echo "pytest parity rc=$?"; tail -5 gpurun_out/r2d_pytest_parity.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sub_windows or merged or random" > gpurun_out/r2d_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r2d_memcheck.log | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['parity']['max_rel_err'], d['parity']['ok'])
    print(d['roofline']['whole_matvec'], d['roofline']['frac'])
    for k in d['kernels'][:14]: print("%-22s %7.2f ms %8.1f GF %6.2f TF/s units %d"%(k['name'],k['ms'],k['gflop'],k['tflops'],k['units']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2d_bench_n1.err').read()[-3000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2d_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:phase2_kernel.*128, 64, 4, 2" -s 4 -c 2 -o gpurun_out/r2d_prof_p2 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2d_ncu_full.log 2>&1
ncu -i gpurun_out/r2d_prof_p2.ncu-rep --page raw --csv > gpurun_out/r2d_prof_p2_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r2d
