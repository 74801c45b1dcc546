#!/bin/bash
# BASELINE.json config 5 (synthetic random-integral 60-orbital FCIDUMP, SU2, M=8000, 8 ranks): the FCIDUMP
# (tools/gen_config5_fcidump.py, SURVEY 8d recipe) and the per-rank H_eff pair lists of the mid-chain site that the
# unmodified reference records under ClassicParallelMPO / ParallelRuleQC (oracle/_ref/b2ref_su2 dump --struct).
# Needs /root/reference (oracle/_ref built); output: workloads/_gen/config5/ (git-ignored, ~18 MB per rank gzipped).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/workloads/_gen/config5
mkdir -p $OUT
[ -f $OUT/RANDOM60.FCIDUMP ] || python $ROOT/tools/gen_config5_fcidump.py $OUT/RANDOM60.FCIDUMP
P=${P:-8}; SITE=${SITE:-29}; M=${M:-8000}
cd $ROOT/oracle/_ref
for r in ${RANKS:-0 1 2 3 4 5 6 7}; do
  F=$OUT/rand60_m${M}_s${SITE}_P${P}_r$r.b2seq
  [ -f $F.gz ] && continue
  OPENBLAS_NUM_THREADS=1 ./b2ref_su2 dump --fcidump $OUT/RANDOM60.FCIDUMP --pg c1 --bond $M --sweeps 0 --site $SITE --struct \
      --threads ${THREADS:-8} --ranks $P --rank $r --classic --out $F > $OUT/dump_P${P}_r$r.log 2>&1
  tail -1 $OUT/dump_P${P}_r$r.log
  gzip -6 -f $F
done
