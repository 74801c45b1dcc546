#!/bin/bash
# round 2, fourteenth GPU job: blocking planner A/B - unit size not divided by the number of contributions
mkdir -p gpurun_out
for LM in 16384 32768 65536; do
  for W in call39 call18; do
    B2G_BLK_LINMAX=$LM timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_$W.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2o_blocking_${W}_$LM.json 2> gpurun_out/r2o_blocking_${W}_$LM.err
    python -c "
import json; d=json.loads(open('gpurun_out/r2o_blocking_${W}_$LM.json').read().strip().splitlines()[-1]); print('$W nodiv linmax $LM', round(d['ms_per_step'],3), round(d['roofline']['frac'],3), round(d['roofline']['per_term']['frac'],3), d['plan_seconds_host'])"
  done
done
B2G_BLK_PERDIV=1 B2G_BLK_LINMAX=16384 timeout 300 python tools/blocking_bench.py workloads/cr2_svp_m4000_blocking/cr2_m4000_s20_call39.b2tp.gz --steps 10 --warmup 3 > gpurun_out/r2o_blocking_call39_div.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2o_blocking_call39_div.json').read().strip().splitlines()[-1]); print('call39 div linmax 16384', round(d['ms_per_step'],3), round(d['roofline']['frac'],3))"
timeout 300 python -m pytest tests/test_blocking.py -m gpu -x -q 2>&1 | tail -1
