#!/bin/bash
W=workloads/cr2_svp_m4000_blocking
for cfg in "128 8192" "256 8192" "512 8192" "128 16384" "256 32768" "128 4096"; do
  set -- $cfg
  for c in 39 18; do
    B2G_BLK_CAPMIN=$1 B2G_BLK_CAPNUM=$2 B2G_VERBOSE=1 timeout 300 python tools/blocking_bench.py $W/cr2_m4000_s20_call$c.b2tp.gz --steps 2 --warmup 1 --check-windows 0 2>&1 >/dev/null | grep "b2g\] blocking" | tail -1 | sed "s/^/capmin=$1 capnum=$2 call$c /" | cut -c1-230
  done
done
