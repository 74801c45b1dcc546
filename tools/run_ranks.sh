#!/bin/bash
# tools/run_ranks.sh P <driver args...> : one process per GPU (rank r on device r), the reference's
# ParallelRuleQC split, host collectives over shared memory, sigma all-reduce over NCCL.
P=$1; shift
TAG=b2g_$$
for r in $(seq 1 $((P-1))); do
  "$@" --ranks $P --rank $r --device $r --shm $TAG > /tmp/${TAG}_r$r.log 2>&1 &
done
"$@" --ranks $P --rank 0 --device 0 --shm $TAG
rc=$?
wait
exit $rc
