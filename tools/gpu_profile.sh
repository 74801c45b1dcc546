#!/bin/bash
# ncu passes on the GPU box: (1) launch list with per-launch device time, (2) full capture of the
# dominant kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-40} -c ${COUNT:-30} --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KREGEX:-phase2_kernel.*128, 64}" -s ${KSKIP:-6} -c 2 \
    -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
