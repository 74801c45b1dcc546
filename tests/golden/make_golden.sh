#!/bin/bash
# Regenerates the committed fixtures from the UNMODIFIED reference (oracle/_ref/b2ref_*,
# built by `make -C oracle ref` from /root/reference/src).  Each file holds the pair list the
# reference recorded, the operand arenas, a random c, the reference's sigma = H.c, the H_eff
# diagonal, the initial ket and the reference's Davidson eigenvalue / iteration count.
# Pair order inside a file depends on the reference's OpenMP merge order; contents do not.
set -e
cd "$(dirname "$0")/../../oracle/_ref"
export OPENBLAS_NUM_THREADS=1
G=../../tests/golden
./b2ref_su2 dump --fcidump data/N2.STO3G.FCIDUMP --bond 60 --sweeps 1 --site 4 --threads 4 --noise 1e-6 --out $G/n2_su2_m60_s4.b2seq
./b2ref_sz  dump --fcidump data/H10.STO6G.R1.8.FCIDUMP --sym sz --bond 40 --sweeps 1 --site 4 --threads 4 --noise 1e-6 --out $G/h10_sz_m40_s4.b2seq
./b2ref_su2 dump --fcidump data/N2.STO3G.FCIDUMP --bond 30 --sweeps 0 --site 1 --threads 4 --out $G/n2_su2_m30_s1.b2seq
./b2ref_su2 dump --fcidump data/N2.STO3G.FCIDUMP --bond 30 --sweeps 0 --site 8 --threads 4 --out $G/n2_su2_m30_s8.b2seq
# two concurrently running ranks of the reference's parallel DMRG (ParallelMPO over ParallelRuleQC; host
# collectives through block2-preview_b200/host/b2g_shm_comm.hpp): each file holds the rank's own pair list
# and the sigma the reference all-reduced over both ranks
rm -rf /tmp/b2ref_par && mkdir -p /tmp/b2ref_par
./b2ref_su2 dump --fcidump data/N2.STO3G.FCIDUMP --bond 40 --sweeps 0 --site 4 --threads 2 --noeigs --ranks 2 --rank 1 --shm gold --scratch /tmp/b2ref_par --out $G/n2_su2_m40_s4_P2_r1.b2seq &
./b2ref_su2 dump --fcidump data/N2.STO3G.FCIDUMP --bond 40 --sweeps 0 --site 4 --threads 2 --noeigs --ranks 2 --rank 0 --shm gold --scratch /tmp/b2ref_par --out $G/n2_su2_m40_s4_P2_r0.b2seq
wait
# blocking lists: the N-th left_contract / right_contract call of a short DMRG run recorded in
# SeqTypes::Auto by the reference's own recorder, operand data, and the blocked operators the
# reference's own BatchGEMMSeq::auto_perform produced from it (ref_harness.cpp `blkdump`).
# Calls whose expressions need a SumProd temporary are refused by the harness (exit 3).  Each file also
# carries the same call in the term form of the host binding (b2g_tp_term), recorded in the same run.
./b2ref_su2 blkdump --fcidump data/N2.STO3G.FCIDUMP --bond 30 --nsweeps 2 --threads 2 --blk-call 2 --out $G/n2_su2_m30_blk2_right.b2blk
./b2ref_su2 blkdump --fcidump data/N2.STO3G.FCIDUMP --bond 30 --nsweeps 2 --threads 2 --blk-call 14 --out $G/n2_su2_m30_blk14_left.b2blk
./b2ref_sz  blkdump --fcidump data/H10.STO6G.R1.8.FCIDUMP --sym sz --bond 40 --nsweeps 2 --threads 2 --blk-call 16 --out $G/h10_sz_m40_blk16_left.b2blk
gzip -9f $G/*.b2blk
