#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocking.py -m gpu -x -q 2>&1 | tail -2
W=workloads/cr2_svp_m4000_blocking
for nt in 0 1; do
  for c in 39 18; do
    if [ $nt = 1 ]; then export B2G_BLK_NOTILE_NARROW=1; else unset B2G_BLK_NOTILE_NARROW; fi
    B2G_VERBOSE=1 timeout 300 python tools/blocking_bench.py $W/cr2_m4000_s20_call$c.b2tp.gz --steps 3 --warmup 2 --check-windows 60 2> gpurun_out/sweep.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('no_narrow_tiles=$nt call$c total_ms',round(l['ms_per_step'],3),'GB/s',round(l['value']),'parity',l['parity'])"
    grep "b2g\] blocking" gpurun_out/sweep.err | tail -1 | cut -c1-260
  done
done
