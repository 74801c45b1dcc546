#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocking.py -m gpu -x -q > gpurun_out/pytest_gpu_blocking.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_blocking.log
tail -4 gpurun_out/pytest_gpu_blocking.log
W=workloads/cr2_svp_m4000_blocking
for c in 39 18; do
timeout 600 python tools/blocking_bench.py $W/cr2_m4000_s20_call$c.b2tp.gz > gpurun_out/blocking_call$c.json 2> gpurun_out/blocking_call$c.err; echo "exit $?"; cat gpurun_out/blocking_call$c.json; tail -3 gpurun_out/blocking_call$c.err
done
if [ "${NCU:-0}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:b2g_blocking_kernel -s 2 -c 1 -o gpurun_out/prof_blocking -f python tools/blocking_bench.py $W/cr2_m4000_s20_call39.b2tp.gz --steps 1 --warmup 2 --check-windows 0 > gpurun_out/ncu_blocking.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_blocking.log
fi
