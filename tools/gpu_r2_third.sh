#!/bin/bash
# round 2, third GPU job: new tile engine (row panels, grouped phase 1, 72-wide tiles) under the parity tests and the
# bench; same-state restart parity on H10 / C2 / Cr2
mkdir -p gpurun_out
B=block2-preview_b200/host/_build
export OPENBLAS_NUM_THREADS=1
T=$(nproc)
S=/dev/shm/b2g_scratch
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py tests/test_blocking.py -m gpu -x -q > gpurun_out/r2c_pytest_parity.log 2>&1
echo "pytest parity rc=$?"; tail -12 gpurun_out/r2c_pytest_parity.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['parity']['max_rel_err'], d['parity']['ok'])
    print(d['roofline']['whole_matvec'], d['roofline']['frac'])
    for k in d['kernels']: print("%-22s %7.2f ms %8.1f GF %6.2f TF/s units %d"%(k['name'],k['ms'],k['gflop'],k['tflops'],k['units']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2c_bench_n1.err').read()[-3000:])
PY
B2G_NO_PANELS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench_n1_nopanels.json 2> gpurun_out/r2c_bench_n1_nopanels.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c_bench_n1_nopanels.json').read().strip().splitlines()[-1])
    print("NO_PANELS", d['value'], d['ms_per_step'], d['parity']['max_rel_err'])
except Exception as e:
    print('bench nopanels parse failed', e)
PY
timeout 600 $B/b2g_dmrg_sz --fcidump $B/data/H10.STO6G.R1.8.FCIDUMP --bond 500 --nsweeps 8 --threads $T --noise 1e-6 --compare --restart-sweeps 2 --scratch $S > gpurun_out/r2c_h10_compare.log 2>&1
echo "h10 compare rc=$?"; grep "SWEEP" gpurun_out/r2c_h10_compare.log; tail -1 gpurun_out/r2c_h10_compare.log | cut -c1-500
timeout 1200 $B/b2g_dmrg_su2 --fcidump $B/data/C2.CAS.PVDZ.FCIDUMP --bond 1000 --nsweeps 6 --threads $T --noise 1e-5 --compare --restart-sweeps 2 --scratch $S --dsize 16 > gpurun_out/r2c_c2_m1000_compare.log 2>&1
echo "c2 m1000 rc=$?"; grep "SWEEP\|Time sweep" gpurun_out/r2c_c2_m1000_compare.log; tail -1 gpurun_out/r2c_c2_m1000_compare.log | cut -c1-2600
timeout 1500 $B/b2g_dmrg_su2 --fcidump $B/data/CR2.SVP.FCIDUMP --occ $B/data/CR2.SVP.OCC --bond 1000 --nsweeps 2 --threads $T --noise 1e-5 --dsize 24 --compare --restart-sweeps 1 --scratch $S > gpurun_out/r2c_cr2_m1000.log 2>&1
echo "cr2 m1000 rc=$?"; grep "SWEEP\|Time sweep" gpurun_out/r2c_cr2_m1000.log; grep "Time sweep" -A6 gpurun_out/r2c_cr2_m1000.log | tail -7; tail -1 gpurun_out/r2c_cr2_m1000.log | cut -c1-2600
