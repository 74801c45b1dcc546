"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
(1) the sigma vectors the unmodified reference produced (tests/golden),
(2) the C oracle on seeded synthetic pair lists, and
(3) size-independent properties (linearity, symmetry) at larger sizes."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_FILES
from oracle import seqdump as sd

pytestmark = pytest.mark.gpu
TOL = 1e-11  # north_star: H.C result <= 1e-11 relative in FP64


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_matvec_matches_reference_sigma(b2g, ctx, name):
    sf = b2g.load_seqfile(os.path.join(GOLDEN, name))
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, sf.arenas)
    v = np.zeros(sf.vsize)
    plan(sf.c, v)
    assert rel(v, sf.v_ref) < TOL
    # accumulate semantics (beta = 1) and scale
    plan(sf.c, v, -0.5)
    assert rel(v, 0.5 * sf.v_ref) < 10 * TOL
    st = plan.stats
    assert st.pairs == sf.npairs and st.nflop_mnk == sf.nflop_mnk
    plan.close()


def random_pair_list(rng, n_out=6, n_terms=12, maxdim=70, strided=True):
    """Synthetic pair list in the spirit of the reference's TestRotateTasked
    (unit_test/test_batch_gemm.cpp:88): random shapes, both transpose flags,
    strided wavefunction windows, several terms per output window."""
    from types import SimpleNamespace
    rows = []
    c_blocks, v_blocks, arenas = [], [], []
    coff = voff = aoff = 0
    for _ in range(n_out):
        m1, n0 = (int(x) for x in rng.integers(1, maxdim, 2))
        ldc1 = n0 + (int(rng.integers(0, 5)) if strided else 0)
        vlen = (m1 - 1) * ldc1 + n0
        for _ in range(int(rng.integers(1, n_terms))):
            m0, k0 = (int(x) for x in rng.integers(1, maxdim, 2))
            lda0 = k0 + (int(rng.integers(0, 5)) if strided else 0)
            clen = (m0 - 1) * lda0 + k0
            tb0, ta1 = int(rng.integers(2)), int(rng.integers(2))
            ldb0 = (k0 if tb0 else n0) + int(rng.integers(0, 3))
            b0len = ((n0 if tb0 else k0) - 1) * ldb0 + (k0 if tb0 else n0)
            lda1 = (m1 if ta1 else m0) + int(rng.integers(0, 3))
            a1len = ((m0 if ta1 else m1) - 1) * lda1 + (m1 if ta1 else m0)
            rows.append(dict(ta0=0, tb0=tb0, m0=m0, n0=n0, k0=k0, lda0=lda0, ldb0=ldb0, ldc0=n0, ta1=ta1, tb1=0,
                             m1=m1, n1=n0, k1=m0, lda1=lda1, ldb1=n0, ldc1=ldc1,
                             alpha0=float(rng.choice([1.0, -1.0, 0.5 ** 0.5])), beta0=0.0,
                             alpha1=float(rng.standard_normal()), beta1=1.0, a0_off=coff, b0_arena=0, b0_off=aoff,
                             a1_arena=0, a1_off=aoff + b0len, c1_off=voff, w_off=0))
            aoff += b0len + a1len
            coff += clen
        voff += vlen
    p = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    for k in sd.I32_NAMES:
        p[k] = p[k].astype(np.int32)
    for k in sd.I64_NAMES:
        p[k] = p[k].astype(np.int64)
    mw = int((p["m0"].astype(np.int64) * p["n0"]).max())
    nf = int((p["m0"].astype(np.int64) * p["n0"] * p["k0"] + p["m1"].astype(np.int64) * p["n1"] * p["k1"]).sum())
    d = sd.SeqDump(npairs=len(rows), csize=coff, vsize=voff, max_work=mw, nflop_mnk=nf, site=0, bond_dim=0,
                   n_sites=0, ndav_ref=0, has_eigs=False, e_ref=0.0, const_e=0.0, t_ref_matvec=0.0, conv_thrd=0.0,
                   arena_sizes=np.array([aoff], dtype=np.int64), p=p)
    d.arenas = rng.standard_normal(aoff)
    d.c = rng.standard_normal(coff)
    return d


def as_seqfile(b2g, d):
    return b2g.SeqFile(npairs=d.npairs, csize=d.csize, vsize=d.vsize, max_work=d.max_work, nflop_mnk=d.nflop_mnk,
                       site=0, bond_dim=0, n_sites=0, ndav_ref=0, e_ref=0.0, const_e=0.0, conv_thrd=0.0,
                       arena_sizes=d.arena_sizes, p=d.p, arenas=d.arenas, c=d.c)


@pytest.mark.parametrize("seed,maxdim", [(0, 9), (1, 40), (2, 70), (3, 150), (4, 300)])
def test_matvec_matches_oracle_on_random_lists(b2g, ctx, seed, maxdim):
    rng = np.random.default_rng(seed)
    d = random_pair_list(rng, maxdim=maxdim, n_out=5 if maxdim > 100 else 8)
    want = sd.replay(d, nthreads=4)
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    got = np.zeros(d.vsize)
    plan(d.c, got)
    assert rel(got, want) < TOL
    plan.close()


def shared_operator_list(rng, reps=(1, 2, 4, 5, 9), dims=((70, 33), (9, 140), (130, 66), (5, 3), (64, 64))):
    """Pair lists in which several pairs write the same sigma window THROUGH THE SAME operator block
    A1 (and some share c window and B0 too): the structure the W pre-sum merges
    (sum_p A1 W_p = A1 sum_p W_p).  Run lengths 1..9 cover the 4-pair merge limit."""
    rows, coff, voff, aoff = [], 0, 0, 0
    for (m1, n0), nrep in zip(dims, reps):
        ldc1 = n0 + 2
        for ta1 in (0, 1):
            m0 = int(rng.integers(1, 90))
            lda1 = (m1 if ta1 else m0) + 1
            a1_off, a1len = aoff, ((m0 if ta1 else m1) - 1) * lda1 + (m1 if ta1 else m0)
            aoff += a1len
            for r in range(nrep):
                k0 = int(rng.integers(1, 50))
                tb0 = int(rng.integers(2))
                ldb0 = (k0 if tb0 else n0)
                b0len = ((n0 if tb0 else k0) - 1) * ldb0 + (k0 if tb0 else n0)
                lda0 = k0 + int(rng.integers(0, 3))
                rows.append(dict(ta0=0, tb0=tb0, m0=m0, n0=n0, k0=k0, lda0=lda0, ldb0=ldb0, ldc0=n0, ta1=ta1, tb1=0,
                                 m1=m1, n1=n0, k1=m0, lda1=lda1, ldb1=n0, ldc1=ldc1, alpha0=1.0, beta0=0.0,
                                 alpha1=float(rng.standard_normal()), beta1=1.0, a0_off=coff, b0_arena=0,
                                 b0_off=aoff, a1_arena=0, a1_off=a1_off, c1_off=voff, w_off=0))
                aoff += b0len
                coff += (m0 - 1) * lda0 + k0
                if r % 3 == 2:   # duplicate the previous pair exactly except for alpha1 (shared c window and B0)
                    dup = dict(rows[-1]); dup["alpha1"] = float(rng.standard_normal()); rows.append(dup)
        voff += (m1 - 1) * ldc1 + n0
    p = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    for k in sd.I32_NAMES:
        p[k] = p[k].astype(np.int32)
    for k in sd.I64_NAMES:
        p[k] = p[k].astype(np.int64)
    mw = int((p["m0"].astype(np.int64) * p["n0"]).max())
    nf = int((p["m0"].astype(np.int64) * p["n0"] * p["k0"] + p["m1"].astype(np.int64) * p["n1"] * p["k1"]).sum())
    d = sd.SeqDump(npairs=len(rows), csize=coff, vsize=voff, max_work=mw, nflop_mnk=nf, site=0, bond_dim=0,
                   n_sites=0, ndav_ref=0, has_eigs=False, e_ref=0.0, const_e=0.0, t_ref_matvec=0.0, conv_thrd=0.0,
                   arena_sizes=np.array([aoff], dtype=np.int64), p=p)
    d.arenas = rng.standard_normal(aoff)
    d.c = rng.standard_normal(coff)
    return d


@pytest.mark.parametrize("no_panels", ["0", "1"])
def test_shared_operator_blocks_are_merged_correctly(b2g, ctx, monkeypatch, no_panels):
    """Pairs that feed one sigma window through the same operator block A1 are summed inside phase 1 (K-segments
    of one GEMM with per-segment factors); B2G_NO_PANELS=1 keeps one panel per window."""
    monkeypatch.setenv("B2G_NO_PANELS", no_panels)
    d = shared_operator_list(np.random.default_rng(33))
    want = sd.replay(d, nthreads=4)
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    for scale in (1.0, -2.5):
        got = np.zeros(d.vsize)
        plan(d.c, got, scale)
        assert rel(got, scale * want) < TOL
    plan.close()


def sigma_block_list(rng, blocks=((150, 71), (40, 200), (9, 9), (130, 81), (64, 1047 // 8), (300, 2))):
    """Sigma blocks cut into column sub-windows, row sub-windows and a full-block window, as the two-site
    wavefunction blocks of the reference are (batch_gemm.hpp:952-1022: &c(cst, 0) / &c(0, cst) with ld = block
    width); the windows of one row range share a pool of operator blocks A1, every (window, A1) has 1..5 pairs
    with different factors, operand layouts and K lengths.  Exercises the row-panel merge of phase 2 (W panels
    with column spans, layers for overlapping column ranges, block-sparse K lists) and the 72-wide tiles."""
    rows, coff, voff, aoff = [], 0, 0, 0

    def new_block(nelem):
        nonlocal aoff
        o = aoff
        aoff += nelem
        return o
    for (R, C) in blocks:
        cuts = sorted(set([0, C] + [int(x) for x in rng.integers(1, C, size=min(3, C - 1))])) if C > 1 else [0, C]
        col_wins = [(0, R, cuts[i], cuts[i + 1] - cuts[i]) for i in range(len(cuts) - 1)]
        wins = col_wins + [(0, R, 0, C)]
        if C > 3:
            wins.append((0, R, cuts[1] // 2, max(1, (C - cuts[1] // 2) // 2)))  # overlaps its neighbours
        if R > 8:
            wins += [(0, R // 2, 0, C), (R // 2, R - R // 2, 0, C)]           # row sub-windows
        pools = {}
        for (r0, m1, c0, n0) in wins:
            key = (r0, m1)
            if key not in pools:
                pools[key] = []
                for _ in range(5):
                    ta1, m0 = int(rng.integers(2)), int(rng.integers(1, 70))
                    lda1 = (m1 if ta1 else m0) + int(rng.integers(0, 2))
                    pools[key].append((new_block(((m0 if ta1 else m1) - 1) * lda1 + (m1 if ta1 else m0)), ta1, m0, lda1))
            for z in rng.choice(5, size=int(rng.integers(2, 6)), replace=False):
                a1_off, ta1, m0, lda1 = pools[key][z]
                for _ in range(int(rng.integers(1, 6))):
                    k0, tb0 = int(rng.integers(1, 60)), int(rng.integers(2))
                    ldb0 = (k0 if tb0 else n0) + int(rng.integers(0, 2))
                    lda0 = k0 + int(rng.integers(0, 3))
                    b0_off = new_block(((n0 if tb0 else k0) - 1) * ldb0 + (k0 if tb0 else n0))
                    rows.append(dict(ta0=0, tb0=tb0, m0=m0, n0=n0, k0=k0, lda0=lda0, ldb0=ldb0, ldc0=n0, ta1=ta1, tb1=0,
                                     m1=m1, n1=n0, k1=m0, lda1=lda1, ldb1=n0, ldc1=C,
                                     alpha0=float(rng.choice([1.0, -1.0, 0.5])), beta0=0.0,
                                     alpha1=float(rng.standard_normal()), beta1=1.0, a0_off=coff, b0_arena=0,
                                     b0_off=b0_off, a1_arena=0, a1_off=a1_off, c1_off=voff + r0 * C + c0, w_off=0))
                    coff += (m0 - 1) * lda0 + k0
        voff += R * C
    order = rng.permutation(len(rows))
    rows = [rows[i] for i in order]
    p = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    for k in sd.I32_NAMES:
        p[k] = p[k].astype(np.int32)
    for k in sd.I64_NAMES:
        p[k] = p[k].astype(np.int64)
    mw = int((p["m0"].astype(np.int64) * p["n0"]).max())
    nf = int((p["m0"].astype(np.int64) * p["n0"] * p["k0"] + p["m1"].astype(np.int64) * p["n1"] * p["k1"]).sum())
    d = sd.SeqDump(npairs=len(rows), csize=coff, vsize=voff, max_work=mw, nflop_mnk=nf, site=0, bond_dim=0,
                   n_sites=0, ndav_ref=0, has_eigs=False, e_ref=0.0, const_e=0.0, t_ref_matvec=0.0, conv_thrd=0.0,
                   arena_sizes=np.array([aoff], dtype=np.int64), p=p)
    d.arenas = rng.standard_normal(aoff)
    d.c = rng.standard_normal(coff)
    return d


@pytest.mark.parametrize("env", [{}, {"B2G_NO_PANELS": "1"}, {"B2G_NO_TILE72": "1"}, {"B2G_KCHUNK": "64"},
                                 {"B2G_ATOMIC_SIGMA": "1"}, {"B2G_WCAP_GB": "1e-7"}])
def test_sigma_blocks_with_sub_windows_merge_into_row_panels(b2g, ctx, monkeypatch, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for seed in (51, 52):
        d = sigma_block_list(np.random.default_rng(seed))
        want = sd.replay(d, nthreads=4)
        plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
        got = np.zeros(d.vsize)
        plan(d.c, got, 1.0)
        assert rel(got, want) < TOL, (seed, env)
        again = np.zeros(d.vsize)
        plan(d.c, again, 1.0)
        if "B2G_ATOMIC_SIGMA" not in env:
            assert np.array_equal(got, again)  # W panels are rebuilt, not re-added; fixed summation order
        plan.close()


def test_bounded_w_workspace_gives_the_same_bits(b2g, ctx, monkeypatch):
    """B2G_WCAP_GB bounds the W workspace: the row panels run slab after slab through one reused buffer
    (zeroed again per slab).  The partial sigma tiles and their summation order do not depend on the slabs, so
    sigma is bit for bit what the single-slab plan gives, also on a second replay of the same plan."""
    for name in ("n2_su2_m60_s4.b2seq", "h10_sz_m40_s4.b2seq"):
        d = sd.load(os.path.join(GOLDEN, name))
        plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
        one = np.zeros(d.vsize)
        plan(d.c, one)
        n_one = plan.stats.launches
        plan.close()
        monkeypatch.setenv("B2G_WCAP_GB", "2e-6")  # 250 doubles: nearly one slab per row panel
        plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
        many, again = np.zeros(d.vsize), np.zeros(d.vsize)
        plan(d.c, many)
        plan(d.c, again)
        assert plan.stats.launches > n_one
        plan.close()
        monkeypatch.delenv("B2G_WCAP_GB")
        assert np.array_equal(one, many) and np.array_equal(many, again), name
        assert rel(one, sd.replay(d, nthreads=4)) < TOL


def test_graph_replay_of_small_lists_takes_new_arguments(b2g, ctx, monkeypatch):
    """Small lists replay a captured CUDA graph from the third call on (first call eager, second captured); c, sigma
    and scale of every replay come through a device argument block.  Five calls with different c, different sigma
    buffers and scales against the oracle, and bit for bit against the eager route (B2G_NO_GRAPH)."""
    import torch
    d = sd.load(os.path.join(GOLDEN, "n2_su2_m60_s4.b2seq"))
    rng = np.random.default_rng(5)
    cs = [rng.standard_normal(d.csize) for _ in range(5)]
    scales = [1.0, -0.5, 2.0, 0.25, 1.0]
    want = [sc * sd.replay(d, c=c, nthreads=4) for c, sc in zip(cs, scales)]
    outs = {}
    for mode in ("graph", "eager"):
        if mode == "eager":
            monkeypatch.setenv("B2G_NO_GRAPH", "1")
        plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
        got = []
        for c, sc in zip(cs, scales):
            cd = torch.tensor(c, dtype=torch.float64, device="cuda")
            vd = torch.zeros(d.vsize, dtype=torch.float64, device="cuda")  # a fresh sigma buffer every call
            torch.cuda.synchronize()
            plan.matvec_dev(cd.data_ptr(), vd.data_ptr(), sc)
            ctx.synchronize()
            got.append(vd.cpu().numpy())
        plan.close()
        outs[mode] = got
    for k in range(5):
        assert rel(outs["graph"][k], want[k]) < TOL, k
        assert np.array_equal(outs["graph"][k], outs["eager"][k]), k


def test_matvec_is_repeatable_on_one_plan(b2g, ctx):
    """A second replay on the same plan must rebuild the W panels, not add to them."""
    d = shared_operator_list(np.random.default_rng(34))
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    a, b_ = np.zeros(d.vsize), np.zeros(d.vsize)
    plan(d.c, a)
    plan(d.c, b_)
    assert rel(a, b_) < 1e-14 and rel(a, sd.replay(d)) < TOL
    plan.close()


def test_sigma_is_bit_reproducible_and_atomic_route_agrees(b2g, ctx, monkeypatch):
    """Default route: per-chunk partial tiles summed in fixed order -> identical bits run to run.
    B2G_ATOMIC_SIGMA=1 keeps the RED.ADD route (same value within rounding)."""
    d = random_pair_list(np.random.default_rng(41), maxdim=260, n_out=5, n_terms=14)
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    runs = []
    for _ in range(3):
        out = np.zeros(d.vsize)
        plan(d.c, out)
        runs.append(out)
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    plan.close()
    monkeypatch.setenv("B2G_ATOMIC_SIGMA", "1")
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    out = np.zeros(d.vsize)
    plan(d.c, out)
    assert rel(out, runs[0]) < 1e-13 and rel(out, sd.replay(d, nthreads=4)) < TOL
    plan.close()


def test_generic_kernel_route_matches_oracle(b2g, ctx, monkeypatch):
    """B2G_FORCE_GENERIC=1 keeps the one-CTA-per-pair kernel: the anchor of the tiled path."""
    monkeypatch.setenv("B2G_FORCE_GENERIC", "1")
    d = random_pair_list(np.random.default_rng(21), maxdim=90, n_out=6)
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    assert plan.stats.launches == 1
    got = np.zeros(d.vsize)
    plan(d.c, got)
    assert rel(got, sd.replay(d, nthreads=4)) < TOL
    plan.close()


def test_empty_and_degenerate_lists(b2g, ctx):
    rng = np.random.default_rng(5)
    d = random_pair_list(rng, n_out=1, n_terms=2, maxdim=3)
    for k in d.p:
        d.p[k] = d.p[k][:0]
    d.npairs = 0
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    v = np.ones(d.vsize)
    plan(d.c, v)
    assert (v == 1).all()
    plan.close()
    # 1 x 1 x 1 pairs only
    d = random_pair_list(rng, n_out=4, n_terms=5, maxdim=2)
    plan = b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)
    got = np.zeros(d.vsize)
    plan(d.c, got)
    assert rel(got, sd.replay(d)) < TOL
    plan.close()


def test_plan_rejects_non_chained_lists(b2g, ctx):
    d = random_pair_list(np.random.default_rng(6), n_out=2, n_terms=3, maxdim=5)
    d.p["beta1"] = d.p["beta1"] * 0.0
    with pytest.raises(b2g.B2GError, match="chained"):
        b2g.SeqPlan.from_seqfile(ctx, as_seqfile(b2g, d), d.arenas)


def test_device_resident_matvec_linearity_and_determinism(b2g, ctx):
    torch = pytest.importorskip("torch")
    sf = b2g.load_seqfile(os.path.join(GOLDEN, "h10_sz_m40_s4.b2seq"))
    dev = torch.device("cuda", 0)
    ops = torch.from_numpy(sf.arenas).to(dev)
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, ops.data_ptr(), b2g.OPERANDS_DEVICE)
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(sf.csize, dtype=torch.float64, generator=g).to(dev)
    y = torch.randn(sf.csize, dtype=torch.float64, generator=g).to(dev)

    def H(vec):
        out = torch.zeros(sf.vsize, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        plan.matvec_dev(vec.data_ptr(), out.data_ptr(), 1.0)
        ctx.synchronize()
        return out

    hx, hy, hxy = H(x), H(y), H(2.0 * x - 3.0 * y)
    assert float((hxy - (2.0 * hx - 3.0 * hy)).norm() / hxy.norm()) < TOL
    # H_eff is symmetric: <y|Hx> = <x|Hy>
    assert abs(float(y @ hx - x @ hy)) < 1e-10 * float(x.norm() * y.norm())
    want = sd.replay(sd.load(os.path.join(GOLDEN, "h10_sz_m40_s4.b2seq")), c=x.cpu().numpy())
    assert rel(hx.cpu().numpy(), want) < TOL
    plan.close()


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_davidson_matches_reference(b2g, ctx, name):
    sf = b2g.load_seqfile(os.path.join(GOLDEN, name))
    plan = b2g.SeqPlan.from_seqfile(ctx, sf, sf.arenas)
    ket = sf.ket0.copy()
    e, nd = plan.davidson(sf.diag, ket, conv_thrd=sf.conv_thrd, soft_max_iter=4000)
    assert abs(e - sf.e_ref) < 1e-9, (e, sf.e_ref)          # energies within 1e-8 Ha (north_star)
    assert abs(nd - sf.ndav_ref) <= 1, (nd, sf.ndav_ref)
    v = np.zeros(sf.vsize)
    plan(ket, v)
    r = v - e * ket
    assert abs(np.linalg.norm(ket) - 1) < 1e-12 and r @ r < max(sf.conv_thrd, 1e-12) * 1.01
    plan.close()


def test_dgemm_batch_matches_oracle(b2g, ctx):
    import ctypes
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    dev = torch.device("cuda", 0)
    G = 40
    ta, tb = rng.integers(0, 2, G), rng.integers(0, 2, G)
    m, n, k = rng.integers(1, 40, G), rng.integers(1, 40, G), rng.integers(1, 40, G)
    m[:8], k[:8], n[:8] = rng.integers(1, 200, 8), 1, 1   # rows-as-AXPY groups (tensor_product lowering)
    ta[:8] = tb[:8] = 0
    lda = np.where(ta == 1, m, k) + rng.integers(0, 3, G)
    ldb = np.where(tb == 1, k, n) + rng.integers(0, 3, G)
    ldc = n + rng.integers(0, 3, G)
    gs = rng.integers(1, 4, G)
    alpha, beta = rng.standard_normal(G), rng.choice([0.0, 1.0, -0.5], G)
    ha, hb, hc, oa, ob, oc = [], [], [], [], [], []
    for g in range(G):
        for _ in range(gs[g]):
            ha.append(rng.standard_normal(((k[g] if ta[g] else m[g]) - 1) * lda[g] + (m[g] if ta[g] else k[g])))
            hb.append(rng.standard_normal(((n[g] if tb[g] else k[g]) - 1) * ldb[g] + (k[g] if tb[g] else n[g])))
            hc.append(rng.standard_normal((m[g] - 1) * ldc[g] + n[g]))
    flat = lambda lst: (np.concatenate(lst), np.cumsum([0] + [len(x) for x in lst])[:-1])
    A, oa = flat(ha); B, ob = flat(hb); C, oc = flat(hc)
    dA, dB, dC = (torch.from_numpy(x).to(dev) for x in (A, B, C))
    ctx.dgemm_batch(ta, tb, m, n, k, alpha, dA.data_ptr() + 8 * oa, lda, dB.data_ptr() + 8 * ob, ldb, beta,
                    dC.data_ptr() + 8 * oc, ldc, gs)
    got = dC.cpu().numpy()
    want = C.copy()
    L = sd.lib()
    i32p = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    f64p = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    pp = lambda base, off: (base + 8 * off).astype(np.uint64)
    pa, pb, pc = pp(A.ctypes.data, oa), pp(B.ctypes.data, ob), pp(want.ctypes.data, oc)
    vp = ctypes.POINTER(ctypes.c_void_p)
    keep = [np.ascontiguousarray(x, dtype=np.int32) for x in (ta, tb, m, n, k, lda, ldb, ldc, gs)]
    al, be = np.ascontiguousarray(alpha), np.ascontiguousarray(beta)
    L.b2o_dgemm_batch(ctypes.c_int64(G), *(x.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) for x in keep[:5]),
                      al.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), pa.ctypes.data_as(vp),
                      keep[5].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), pb.ctypes.data_as(vp),
                      keep[6].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                      be.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), pc.ctypes.data_as(vp),
                      keep[7].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                      keep[8].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
    assert np.allclose(got, want, rtol=0, atol=1e-11)


@pytest.mark.parametrize("seed,maxdim", [(51, 30), (52, 120), (53, 260)])
def test_rotation_list_with_host_operands_matches_oracle(b2g, ctx, seed, maxdim):
    """b2g_pairs_execute: the chained-pair lists tensor_rotate records (all operands real host
    pointers, every pair its own output block), results added into the host C blocks."""
    rng = np.random.default_rng(seed)
    d = random_pair_list(rng, maxdim=maxdim, n_out=10, n_terms=2)
    P = d.p
    b0o, a1o = d.operand_offsets()
    out = rng.standard_normal(d.vsize)          # pre-existing content of C (beta = 1)
    out0 = out.copy()
    w = np.zeros(len(P["m0"]), dtype=np.uint64)
    batch0 = dict(ta=P["ta0"] + 111, tb=P["tb0"] + 111, m=P["m0"], n=P["n0"], k=P["k0"], lda=P["lda0"], ldb=P["ldb0"],
                  ldc=P["ldc0"], alpha=P["alpha0"], beta=P["beta0"], a=d.c.ctypes.data + 8 * P["a0_off"],
                  b=d.arenas.ctypes.data + 8 * b0o, c=w)
    batch1 = dict(ta=P["ta1"] + 111, tb=P["tb1"] + 111, m=P["m1"], n=P["n1"], k=P["k1"], lda=P["lda1"], ldb=P["ldb1"],
                  ldc=P["ldc1"], alpha=P["alpha1"], beta=P["beta1"], a=d.arenas.ctypes.data + 8 * a1o, b=w,
                  c=out.ctypes.data + 8 * P["c1_off"])
    st = ctx.pairs_execute(batch0, batch1, d.max_work)
    want = out0 + sd.replay(d, nthreads=4)
    assert rel(out, want) < TOL
    assert st.pairs == d.npairs and st.nflop_mnk == d.nflop_mnk


def test_rank_lists_of_parallel_reference_on_gpu(b2g, ctx):
    """The two per-rank lists the reference records under ParallelRuleQC, replayed on the GPU one
    after the other into the same sigma: must equal the sigma the reference all-reduced."""
    v = None
    for r in (0, 1):
        sf = b2g.load_seqfile(os.path.join(GOLDEN, f"n2_su2_m40_s4_P2_r{r}.b2seq"))
        plan = b2g.SeqPlan.from_seqfile(ctx, sf, sf.arenas)
        if v is None:
            v = np.zeros(sf.vsize)
        plan(sf.c, v)
        plan.close()
    assert rel(v, sf.v_ref) < TOL


def test_syevd_matches_lapack_convention(b2g, ctx):
    """b2g_syevd (library-backed: cuSOLVER) against numpy.linalg.eigh: eigenvalues ascending, eigenvector k in ROW k
    of the row-major matrix (what LAPACK dsyev("V") leaves for the reference, core/matrix_functions.hpp:1672-1688),
    with a leading dimension larger than n and from several host threads at once."""
    import ctypes
    import threading
    L = b2g.lib()
    rng = np.random.default_rng(77)

    def run(n, lda, out):
        x = rng.standard_normal((n, n))
        sym = x + x.T
        a = np.zeros((n, lda))
        a[:, :n] = sym
        w = np.zeros(n)
        rc = L.b2g_syevd(ctx._h, ctypes.c_int(n), ctypes.c_void_p(a.ctypes.data), ctypes.c_int(lda),
                         ctypes.c_void_p(w.ctypes.data))
        out.append((rc, sym, a[:, :n].copy(), w))

    jobs = [(1, 1), (7, 9), (64, 64), (300, 300), (513, 520)]
    results = []
    threads = [threading.Thread(target=run, args=(n, lda, results)) for n, lda in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(results) == len(jobs)
    for rc, sym, vecs, w in results:
        assert rc == 0
        n = sym.shape[0]
        assert np.all(np.diff(w) >= 0)
        assert np.allclose(w, np.linalg.eigvalsh(sym), rtol=0, atol=1e-11 * max(1.0, np.abs(w).max()))
        # rows are eigenvectors: sym @ v_k = w_k v_k, orthonormal
        assert np.linalg.norm(vecs @ sym - w[:, None] * vecs) < 1e-10 * max(1.0, np.abs(w).max()) * n
        assert np.linalg.norm(vecs @ vecs.T - np.eye(n)) < 1e-11 * n
