"""CPU test of the N>1 path (world_size 2, gloo): each rank replays its shard of the pair
list (SeqFile.shard, the ParallelRuleQC-style split of the MPO terms), the partial sigma are
all-reduced, and every rank must hold the serial sigma - the algebraic identity
sum_ranks sigma_r = sigma that ParallelTensorFunctions::operator() relies on
(block2 core/parallel_tensor_functions.hpp:51-55).  The oracle stands in for the GPU here."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, path, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import b2gpkg
    from oracle import seqdump as sd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b2g = b2gpkg.load()
    sf = b2g.load_seqfile(path)
    mine = sf.shard(rank, world)
    d = sd.SeqDump(npairs=mine.npairs, csize=mine.csize, vsize=mine.vsize, max_work=mine.max_work,
                   nflop_mnk=mine.nflop_mnk, site=0, bond_dim=0, n_sites=0, ndav_ref=0, has_eigs=False, e_ref=0.0,
                   const_e=0.0, t_ref_matvec=0.0, conv_thrd=0.0, arena_sizes=mine.arena_sizes, p=mine.p,
                   arenas=mine.arenas, c=sf.c)
    part = sd.replay(d) if mine.npairs else np.zeros(sf.vsize)
    t = torch.from_numpy(part)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    counts = torch.tensor([mine.npairs, mine.nflop_mnk], dtype=torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f"sigma_{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"counts_{rank}.npy"), counts.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["h10_sz_m40_s4.b2seq", "n2_su2_m60_s4.b2seq"])
def test_sharded_replay_sums_to_serial_sigma(tmp_path, name):
    import torch.multiprocessing as mp
    import b2gpkg
    path = os.path.join(GOLDEN, name)
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, path, str(tmp_path)), nprocs=2, join=True)
    sf = b2gpkg.load().load_seqfile(path)
    for r in range(2):
        sigma = np.load(tmp_path / f"sigma_{r}.npy")
        counts = np.load(tmp_path / f"counts_{r}.npy")
        assert counts[0] == sf.npairs and counts[1] == sf.nflop_mnk      # a partition: nothing lost or doubled
        assert np.linalg.norm(sigma - sf.v_ref) < 1e-12 * np.linalg.norm(sf.v_ref)


def test_shards_partition_pairs_and_balance():
    import b2gpkg
    b2g = b2gpkg.load()
    sf = b2g.load_seqfile(os.path.join(ROOT, "workloads", "cr2_svp_m4000_site20.b2seq.gz"))
    for world in (2, 4, 8):
        shards = [sf.shard(r, world) for r in range(world)]
        assert sum(s.npairs for s in shards) == sf.npairs
        assert sum(s.nflop_mnk for s in shards) == sf.nflop_mnk
        fl = np.array([s.flops for s in shards])
        assert fl.max() / fl.mean() < 1.25
        assert max(s.operand_doubles for s in shards) < 0.75 * sf.operand_doubles
