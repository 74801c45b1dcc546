#!/bin/bash
# round 2, seventh GPU job: programmatic dependent launch chain in the tile engine (A/B), source-major blocking units
# (A/B), parity + memcheck, ncu launch list and one --set full capture of the dominant phase-2 kernel.
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_blocking.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2h_pytest.log
for V in pdl nopdl; do
  if [ $V = nopdl ]; then export B2G_NO_PDL=1; else unset B2G_NO_PDL; fi
  timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_bench_$V.json 2> gpurun_out/r2h_bench_$V.err
  echo "bench $V rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench_$V.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['parity']['max_rel_err'], d['parity']['ok'])
    print(d['roofline']['whole_matvec'], d['roofline']['frac'])
    print('blocking', d['blocking']['ms'], d['blocking']['roofline']['frac'], d['blocking']['roofline'].get('per_term',{}).get('frac'))
    print('small', d.get('small_sector'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2h_bench_$V.err').read()[-3000:])
PY
done
unset B2G_NO_PDL
for W in call39 call18; do
  for V in sorted nosort; do
    if [ $V = nosort ]; then export B2G_BLK_NOSORT=1; else unset B2G_BLK_NOSORT; fi
    timeout 600 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_$W.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2h_blocking_${W}_$V.json 2> gpurun_out/r2h_blocking_${W}_$V.err
    echo "blocking $W $V rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2h_blocking_${W}_$V.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_term']['frac'], d['parity'])"
  done
done
unset B2G_BLK_NOSORT
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"phase|reduce_kernel" -s 140 -c 70 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2h_ncu_bench.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/r2h_launches.csv
ncu --set full --clock-control none --import-source on -k regex:phase2_kernel -s 20 -c 2 -o gpurun_out/r2h_prof_p2 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2h_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r2h_prof_p2.ncu-rep --page raw --csv > gpurun_out/r2h_prof_p2_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r2h
