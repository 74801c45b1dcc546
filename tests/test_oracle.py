"""CPU tests: the oracle (oracle/replay.c, oracle/davidson.py) against the golden vectors
produced by the unmodified reference, the .b2seq readers, and the host-side logic."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_FILES
from oracle import seqdump as sd


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_c_oracle_matches_reference_sigma(name):
    d = sd.load(os.path.join(GOLDEN, name))
    v = sd.replay(d)
    err = np.linalg.norm(v - d.v_ref) / np.linalg.norm(d.v_ref)
    assert err < 1e-13, err


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_numpy_restatement_matches_c_oracle(name):
    d = sd.load(os.path.join(GOLDEN, name))
    v, v2 = sd.replay(d), sd.replay_numpy(d)
    assert np.linalg.norm(v - v2) <= 1e-13 * np.linalg.norm(v)


@pytest.mark.parametrize("name", GOLDEN_FILES[:2])
def test_threaded_oracle_and_scale(name):
    d = sd.load(os.path.join(GOLDEN, name))
    v1 = sd.replay(d, nthreads=1, scale=-0.75)
    v4 = sd.replay(d, nthreads=4, scale=-0.75)
    assert np.linalg.norm(v1 + 0.75 * d.v_ref) <= 1e-13 * np.linalg.norm(d.v_ref)
    assert np.linalg.norm(v1 - v4) <= 1e-13 * np.linalg.norm(v1)


def test_recorded_pairs_follow_replay_invariants():
    """beta0 = 0, beta1 = 1, GEMM1's B is the contiguous work matrix of GEMM0."""
    for name in GOLDEN_FILES:
        P = sd.load(os.path.join(GOLDEN, name)).p
        assert (P["beta0"] == 0).all() and (P["beta1"] == 1).all()
        assert (P["k1"] == P["m0"]).all() and (P["n1"] == P["n0"]).all()
        assert (P["tb1"] == 0).all() and (P["ldb1"] == P["n0"]).all() and (P["ldc0"] == P["n0"]).all()
        assert (P["ta0"] == 0).all()


def test_product_reader_agrees_with_oracle_reader():
    import b2gpkg
    b2g = b2gpkg.load()
    for name in GOLDEN_FILES:
        path = os.path.join(GOLDEN, name)
        a, b = sd.load(path), b2g.load_seqfile(path)
        assert a.npairs == b.npairs and a.csize == b.csize and a.nflop_mnk == b.nflop_mnk
        for k in a.p:
            assert np.array_equal(a.p[k], b.p[k]), k
        assert np.array_equal(a.arenas, b.arenas) and np.array_equal(a.v_ref, b.v_ref)
        b0, b1 = b.as_batches(1 << 40)
        assert (b0["c"] == b1["b"]).all() and (b0["a"] // 8 < a.csize).all() and (b1["c"] // 8 < a.vsize).all()


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_davidson_restatement_matches_reference(name):
    from oracle import davidson as dv
    d = sd.load(os.path.join(GOLDEN, name))
    e, nd, vec = dv.davidson(lambda x: sd.replay(d, c=x), d.diag, d.ket0, conv_thrd=d.conv_thrd,
                             soft_max_iter=4000)
    assert abs(e - d.e_ref) < 1e-9, (e, d.e_ref)
    assert abs(nd - d.ndav_ref) <= 1, (nd, d.ndav_ref)
    r = sd.replay(d, c=vec) - e * vec
    assert r @ r < max(d.conv_thrd, 1e-12) * 1.01


def test_h_eff_is_symmetric_on_golden():
    d = sd.load(os.path.join(GOLDEN, "n2_su2_m30_s1.b2seq"))
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(d.csize), rng.standard_normal(d.csize)
    assert abs(y @ sd.replay(d, c=x) - x @ sd.replay(d, c=y)) < 1e-10 * np.linalg.norm(x) * np.linalg.norm(y) * 100


def test_dgemm_batch_oracle_against_numpy():
    import ctypes
    rng = np.random.default_rng(7)
    L = sd.lib()
    for _ in range(20):
        ta, tb = int(rng.integers(2)), int(rng.integers(2))
        m, n, k = (int(x) for x in rng.integers(1, 9, 3))
        lda, ldb, ldc = (k if not ta else m) + 2, (n if not tb else k) + 1, n + 3
        A = rng.standard_normal(((m if not ta else k), lda))
        B = rng.standard_normal(((k if not tb else n), ldb))
        C = rng.standard_normal((m, ldc))
        C0 = C.copy()
        al, be = 0.7, -1.3
        arr = lambda v, t: np.array([v], dtype=t)
        pa, pb, pc = arr(A.ctypes.data, np.uint64), arr(B.ctypes.data, np.uint64), arr(C.ctypes.data, np.uint64)
        i32 = lambda v: arr(v, np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        f64 = lambda v: arr(v, np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        L.b2o_dgemm_batch(ctypes.c_int64(1), i32(ta), i32(tb), i32(m), i32(n), i32(k), f64(al),
                          pa.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p)), i32(lda),
                          pb.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p)), i32(ldb), f64(be),
                          pc.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p)), i32(ldc), i32(1))
        opA = A[:, :m].T if ta else A[:, :k]
        opB = B[:, :k].T if tb else B[:, :n]
        ref = al * opA @ opB + be * C0[:, :n]
        assert np.allclose(C[:, :n], ref, atol=1e-13)
        assert np.array_equal(C[:, n:], C0[:, n:])


def test_rank_lists_of_parallel_reference_sum_to_its_allreduced_sigma():
    """Two ranks of the reference's own parallel DMRG (ParallelRuleQC / ParallelMPO) each record their
    slice of the MPO terms; the reference all-reduces sigma (parallel_tensor_functions.hpp:51-55).
    The per-rank replays of the oracle must add up to that sigma."""
    r0 = sd.load(os.path.join(GOLDEN, "n2_su2_m40_s4_P2_r0.b2seq"))
    r1 = sd.load(os.path.join(GOLDEN, "n2_su2_m40_s4_P2_r1.b2seq"))
    assert np.array_equal(r0.c, r1.c) and np.array_equal(r0.v_ref, r1.v_ref)
    assert r0.npairs != r1.npairs or not np.array_equal(r0.p["alpha1"], r1.p["alpha1"])
    s = sd.replay(r0) + sd.replay(r1)
    assert np.linalg.norm(s - r0.v_ref) < 1e-13 * np.linalg.norm(r0.v_ref)
