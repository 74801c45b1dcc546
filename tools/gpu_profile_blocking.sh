#!/bin/bash
# ncu on the blocking kernels: (1) launch list of both lists with DRAM bytes, (2) full capture of the
# streaming, multi-source and tile kernels of the full blocking list.
mkdir -p gpurun_out
W=workloads/cr2_svp_m4000_blocking
for c in 39 18; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b2g_blocking -s 8 -c 4 --csv \
    --log-file gpurun_out/blocking_launches_call$c.csv python tools/blocking_bench.py $W/cr2_m4000_s20_call$c.b2tp.gz --steps 1 --warmup 2 --check-windows 0 > gpurun_out/ncu_blk_list$c.log 2>&1
tail -5 gpurun_out/blocking_launches_call$c.csv
done
ncu --set full --clock-control none --import-source on -k regex:b2g_blocking -s 8 -c 4 -o gpurun_out/prof_blocking_v3 -f \
    python tools/blocking_bench.py $W/cr2_m4000_s20_call18.b2tp.gz --steps 1 --warmup 2 --check-windows 0 > gpurun_out/ncu_blk_full.log 2>&1
tail -2 gpurun_out/ncu_blk_full.log
