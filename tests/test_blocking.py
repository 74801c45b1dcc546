"""Blocking lists (TensorFunctions::left_contract / right_contract, block2
src/core/tensor_functions.hpp:2842-2885, 2941-2984).

CPU (-m "not gpu"): the C oracle's sequential list execution (oracle/replay.c b2o_dgemm_batch)
against the blocked operators the UNMODIFIED reference produced with its own
BatchGEMMSeq::auto_perform on the same recorded list (tests/golden/*.b2blk.gz).
GPU (-m gpu): b2g_batch_execute through the C ABI against the same fixtures, against the oracle on
seeded synthetic lists (write conflicts, 2-D windows, general GEMM entries, beta != 1, overlapping
windows), bit-reproducibility, and linearity at a larger size."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import seqdump as sd

BLK_FILES = ["n2_su2_m30_blk2_right.b2blk.gz", "n2_su2_m30_blk14_left.b2blk.gz", "h10_sz_m40_blk16_left.b2blk.gz"]
TOL = 1e-13


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("name", BLK_FILES)
def test_oracle_matches_reference_blocking(b2g, name):
    bf = b2g.load_blkfile(os.path.join(GOLDEN, name))
    assert bf.nentries == int(bf.g["gp"].sum()) and bf.nflop_mnk > 0
    assert np.count_nonzero(bf.c_in) == 0  # freshly allocated operators
    inp, out = bf.inputs.copy(), bf.c_in.copy()
    a, b, c = bf.pointers(inp.ctypes.data, out.ctypes.data)
    ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, gs = bf.group_args()
    sd.batch_perform(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs)
    assert rel(out, bf.c_ref) < TOL
    assert np.array_equal(inp, bf.inputs)


def test_blocking_fixtures_carry_write_conflicts(b2g):
    """The lists are conflict-carrying (several entries per output window): that is the case
    auto_perform resolves with work arrays and b2g_batch_execute with per-window ownership."""
    bf = b2g.load_blkfile(os.path.join(GOLDEN, BLK_FILES[1]))
    _, _, c = bf.pointers(0, 0)
    assert len(np.unique(c)) < len(c)
    assert (bf.g["gp"] > 1).any() and (bf.g["k"] == 1).all()


# ----------------------------------------------------------------------------- GPU


def run_gpu(b2g, ctx, bf_args, inp, out, flags=0):
    ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs = bf_args
    return ctx.batch_execute(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs, b2g.OPERANDS_HOST, flags)


@pytest.mark.gpu
@pytest.mark.parametrize("name", BLK_FILES)
@pytest.mark.parametrize("flags", [0, 1])
def test_gpu_blocking_matches_reference(b2g, ctx, name, flags):
    bf = b2g.load_blkfile(os.path.join(GOLDEN, name))
    inp, out = bf.inputs.copy(), bf.c_in.copy()
    a, b, c = bf.pointers(inp.ctypes.data, out.ctypes.data)
    ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, gs = bf.group_args()
    st = run_gpu(b2g, ctx, (ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs), inp, out, flags)
    assert rel(out, bf.c_ref) < TOL
    assert st.entries == bf.nentries and st.nflop_mnk == bf.nflop_mnk
    assert st.clusters <= st.merged <= st.entries and st.serial_entries == 0 and st.launches >= 1
    assert np.array_equal(inp, bf.inputs)


def synthetic_list(rng, n_blocks=5, terms=6, maxdim=40, general=True, overlap=False):
    """Random conflict-carrying list in the group form of BatchGEMM<double>: AXPY row groups into
    column sub-windows of shared output blocks (what AdvancedGEMM::tensor_product emits), whole-block
    AXPYs, scaled (beta != 1) entries, alpha = 0 entries (iscale) and, optionally, general k > 1 GEMMs
    and partially overlapping windows."""
    G = {k: [] for k in ("ta", "tb", "m", "n", "k", "alpha", "lda", "ldb", "beta", "ldc", "gp")}
    A, B, C = [], [], []
    src = rng.standard_normal(200000)
    scal = rng.standard_normal(64)
    out_blocks = []
    ooff = 0
    for _ in range(n_blocks):
        rows, cols = int(rng.integers(2, maxdim)), int(rng.integers(2, maxdim))
        out_blocks.append((ooff, rows, cols))
        ooff += rows * cols
    out = rng.standard_normal(ooff)
    soff = 0

    def group(ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, entries):
        for key, val in zip(("ta", "tb", "m", "n", "k", "alpha", "lda", "ldb", "beta", "ldc", "gp"),
                            (ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, len(entries))):
            G[key].append(val)
        for (a, b, c) in entries:
            A.append(a), B.append(b), C.append(c)

    for (o, rows, cols) in out_blocks:
        # like the reference's blocked operators, one block is written either through a fixed grid
        # of sub-windows (row groups x column groups), or as a whole in one of two forms
        family = int(rng.integers(0, 3 if general else 2))
        rcut = sorted(set([0, rows] + [int(x) for x in rng.integers(1, rows, 2)]))
        ccut = sorted(set([0, cols] + [int(x) for x in rng.integers(1, cols, 2)])) if family == 0 else [0, cols]
        for _ in range(int(rng.integers(1, terms))):
            s = int(rng.integers(0, 64))
            if family <= 1:
                ri, ci = int(rng.integers(0, len(rcut) - 1)), int(rng.integers(0, len(ccut) - 1))
                r0, nr, c0, w = rcut[ri], rcut[ri + 1] - rcut[ri], ccut[ci], ccut[ci + 1] - ccut[ci]
                kind = int(rng.integers(0, 3))
                if len(ccut) == 2 and kind == 0:  # full-width row group: one contiguous AXPY (a.n == c.n branch)
                    group(111, 111, nr * cols, 1, 1, float(rng.standard_normal()), 1, 1, 1.0, 1,
                          [(("s", soff), ("k", s), ("o", o + r0 * cols))])
                    soff += nr * cols
                elif len(ccut) == 2 and kind == 1:  # scale in place (iscale): alpha = 0, A aliases C, never read
                    group(111, 111, nr * cols, 1, 1, 0.0, 1, 1, float(rng.uniform(0.5, 1.5)), 1,
                          [(("o", o + r0 * cols), ("k", s), ("o", o + r0 * cols))])
                elif len(ccut) == 2:
                    group(111, 111, nr * cols, 1, 1, float(rng.standard_normal()), 1, 1, 0.5, 1,
                          [(("s", soff), ("k", s), ("o", o + r0 * cols))])
                    soff += nr * cols
                elif kind <= 1:  # row group into a column sub-window, padded source rows
                    ld_src = w + int(rng.integers(0, 3))
                    group(111, 111, w, 1, 1, float(rng.standard_normal()), 1, 1, 1.0, 1,
                          [(("s", soff + r * ld_src), ("k", s), ("o", o + (r0 + r) * cols + c0)) for r in range(nr)])
                    soff += nr * ld_src
                else:  # transposed source (conj branch): element i of row r is a[i * nr + r]
                    group(111, 111, w, 1, 1, float(rng.standard_normal()), nr, 1, 1.0, 1,
                          [(("s", soff + r), ("k", s), ("o", o + (r0 + r) * cols + c0)) for r in range(nr)])
                    soff += w * nr
            elif int(rng.integers(2)):  # general GEMM into the whole block
                kk = int(rng.integers(2, 9))
                ta, tb = int(rng.integers(2)), int(rng.integers(2))
                lda = (rows if ta else kk) + int(rng.integers(0, 2))
                ldb = (kk if tb else cols) + int(rng.integers(0, 2))
                ea = ((kk if ta else rows) - 1) * lda + (rows if ta else kk)
                eb = ((cols if tb else kk) - 1) * ldb + (kk if tb else cols)
                group(112 if ta else 111, 112 if tb else 111, rows, cols, kk, float(rng.standard_normal()), lda, ldb,
                      float(rng.choice([1.0, 0.5])), cols, [(("s", soff), ("s", soff + ea), ("o", o))])
                soff += ea + eb
            else:  # outer product of two strided vectors (tensor_product_diagonal form)
                group(111, 112, rows, cols, 1, float(rng.standard_normal()), 2, 3, 1.0, cols,
                      [(("s", soff), ("s", soff + 2 * rows), ("o", o))])
                soff += 2 * rows + 3 * cols
        if overlap and rows > 2:  # a window overlapping the block without being one of its windows
            group(111, 111, cols + 1, 1, 1, 0.7, 1, 1, 1.0, 1, [(("s", soff), ("k", 0), ("o", o + cols - 1))])
            soff += cols + 1
    assert soff <= src.size
    return G, A, B, C, src, scal, out


def resolve(G, A, B, C, src, scal, out):
    base = {"s": src.ctypes.data, "k": scal.ctypes.data, "o": out.ctypes.data}
    conv = lambda L: np.array([base[t] + 8 * off for (t, off) in L], dtype=np.uint64)
    return (G["ta"], G["tb"], G["m"], G["n"], G["k"], G["alpha"], conv(A), G["lda"], conv(B), G["ldb"], G["beta"],
            conv(C), G["ldc"], G["gp"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,general,overlap", [(0, False, False), (1, True, False), (2, True, True),
                                                  (3, False, True), (4, True, False)])
def test_gpu_blocking_matches_oracle_on_synthetic_lists(b2g, ctx, seed, general, overlap):
    rng = np.random.default_rng(seed)
    G, A, B, C, src, scal, out0 = synthetic_list(rng, general=general, overlap=overlap)
    out_cpu, out_gpu = out0.copy(), out0.copy()
    args = resolve(G, A, B, C, src, scal, out_cpu)
    sd.batch_perform(*[args[i] for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13)])
    st = run_gpu(b2g, ctx, resolve(G, A, B, C, src, scal, out_gpu), src, out_gpu)
    assert rel(out_gpu, out_cpu) < 1e-13
    if overlap:
        assert st.serial_entries > 0
    else:
        assert st.serial_entries == 0
    # bit-reproducible
    out_again = out0.copy()
    run_gpu(b2g, ctx, resolve(G, A, B, C, src, scal, out_again), src, out_again)
    assert np.array_equal(out_again, out_gpu)


@pytest.mark.gpu
def test_gpu_blocking_rejects_aliasing_and_bad_flags(b2g, ctx):
    x = np.ones(16)
    one = np.ones(1)
    p = lambda arr, off=0: np.array([arr.ctypes.data + 8 * off], dtype=np.uint64)
    with pytest.raises(b2g.B2GError, match="aliases"):  # output read as input with alpha != 0
        ctx.batch_execute([111], [111], [8], [1], [1], [1.0], p(x), [1], p(one), [1], [1.0], p(x, 4), [1], [1])
    with pytest.raises(b2g.B2GError, match="transpose"):
        ctx.batch_execute([113], [111], [8], [1], [1], [1.0], p(x), [1], p(one), [1], [1.0], p(x, 8), [1], [1])
    # empty list and empty groups are fine
    st = ctx.batch_execute([], [], [], [], [], [], [], [], [], [], [], [], [], [])
    assert st.entries == 0
    st = ctx.batch_execute([111], [111], [0], [1], [1], [1.0], p(x), [1], p(one), [1], [1.0], p(x, 8), [1], [1])
    assert st.entries == 0 and x.sum() == 16


@pytest.mark.gpu
def test_gpu_blocking_linearity_large(b2g, ctx):
    """Size-independent property at a bond-dimension-like size: the blocked operator is linear in
    the environment operators, C(x + 2y) = C(x) + 2 C(y)."""
    rng = np.random.default_rng(7)
    rows, cols, nterm = 1500, 900, 12
    scal = rng.standard_normal(nterm)
    G = dict(ta=[111] * nterm, tb=[111] * nterm, m=[cols] * nterm, n=[1] * nterm, k=[1] * nterm,
             alpha=list(rng.standard_normal(nterm)), lda=[1] * nterm, ldb=[1] * nterm, beta=[1.0] * nterm,
             ldc=[1] * nterm, gp=[rows] * nterm)

    def run(src):
        out = np.zeros(rows * 2 * cols)
        A, B, C = [], [], []
        for t in range(nterm):
            c0 = (t % 2) * cols
            for r in range(rows):
                A.append(src.ctypes.data + 8 * ((t * rows + r) * cols))
                B.append(scal.ctypes.data + 8 * t)
                C.append(out.ctypes.data + 8 * (r * 2 * cols + c0))
        st = ctx.batch_execute(G["ta"], G["tb"], G["m"], G["n"], G["k"], G["alpha"], np.array(A, dtype=np.uint64),
                               G["lda"], np.array(B, dtype=np.uint64), G["ldb"], G["beta"],
                               np.array(C, dtype=np.uint64), G["ldc"], G["gp"], b2g.OPERANDS_HOST, b2g.DST_ZERO)
        return out, st

    x, y = rng.standard_normal(nterm * rows * cols), rng.standard_normal(nterm * rows * cols)
    cx, st = run(x)
    cy, _ = run(y)
    cxy, _ = run(x + 2.0 * y)
    assert st.entries == nterm * rows and st.merged == nterm and st.clusters == 2
    assert st.bytes_in == 8 * (nterm * rows * cols + nterm) and st.bytes_out == 8 * rows * 2 * cols
    assert rel(cxy, cx + 2.0 * cy) < 1e-13
    ref = np.zeros((rows, 2 * cols))
    for t in range(nterm):
        ref[:, (t % 2) * cols:(t % 2 + 1) * cols] += G["alpha"][t] * scal[t] * x[t * rows * cols:(t + 1) * rows * cols].reshape(rows, cols)
    assert rel(cx, ref.ravel()) < 1e-13


def synthetic_terms(rng, n_blocks=6, maxdim=30, kron=True):
    """Blocking terms in the spirit of OperatorFunctions::tensor_product: every output block has a fixed
    grid of (row group x column group) windows; each term adds scale * op(A) (x) op(B) into one window,
    with A or B a 1 x 1 block (the quantum-chemistry case) or, optionally, both larger."""
    src = rng.standard_normal(400000)
    terms, soff, ooff = [], 0, 0
    for _ in range(n_blocks):
        rgs = [int(x) for x in rng.integers(1, maxdim, int(rng.integers(1, 4)))]
        cgs = [int(x) for x in rng.integers(1, maxdim, int(rng.integers(1, 4)))]
        rows, cols = sum(rgs), sum(cgs)
        # a window belongs to one (a sector, b sector) pair: every term that writes it has the same block
        # shapes up to transposition
        kinds = rng.integers(3 if kron else 2, size=(len(rgs), len(cgs)))
        for _ in range(int(rng.integers(2, 14))):
            ri, ci = int(rng.integers(len(rgs))), int(rng.integers(len(cgs)))
            wr, wc = rgs[ri], cgs[ci]
            c = ooff + sum(rgs[:ri]) * cols + sum(cgs[:ci])
            conja, conjb = int(rng.integers(2)), int(rng.integers(2))
            kind = int(kinds[ri, ci])
            if kind == 0:    # b is 1 x 1
                am, an, bm, bn = (wc, wr, 1, 1) if conja else (wr, wc, 1, 1)
            elif kind == 1:  # a is 1 x 1
                am, an, bm, bn = (1, 1, wc, wr) if conjb else (1, 1, wr, wc)
            else:            # Kronecker: split the window as (ar*br) x (ac*bc) when it factorises
                ar = next(d for d in range(min(wr, 4), 0, -1) if wr % d == 0)
                ac = next(d for d in range(min(wc, 4), 0, -1) if wc % d == 0)
                br, bc = wr // ar, wc // ac
                am, an = (ac, ar) if conja else (ar, ac)
                bm, bn = (bc, br) if conjb else (br, bc)
            terms.append(dict(a=soff, b=soff + am * an, c=c, am=am, an=an, bm=bm, bn=bn, cn=cols, conja=conja,
                              conjb=conjb, scale=float(rng.standard_normal())))
            soff += am * an + bm * bn
        ooff += rows * cols
    assert soff <= src.size
    return terms, src, ooff


def pack_terms(b2g, terms, src, out):
    arr = np.zeros(len(terms), dtype=b2g.TP_DTYPE)
    for i, t in enumerate(terms):
        arr[i] = (src.ctypes.data + 8 * t["a"], src.ctypes.data + 8 * t["b"], out.ctypes.data + 8 * t["c"], t["am"],
                  t["an"], t["bm"], t["bn"], t["cn"], t["conja"], t["conjb"], 0, t["scale"])
    return arr


@pytest.mark.gpu
@pytest.mark.parametrize("seed,kron,zero", [(0, False, True), (1, True, True), (2, True, False), (3, False, False)])
def test_gpu_tensor_product_terms_match_oracle(b2g, ctx, seed, kron, zero):
    rng = np.random.default_rng(100 + seed)
    terms, src, osize = synthetic_terms(rng, kron=kron)
    out0 = np.zeros(osize) if zero else rng.standard_normal(osize)
    ref = out0.copy()
    sd.tensor_product_terms(terms, lambda off, n: src[off:off + n],
                            lambda off, r, c, p: np.lib.stride_tricks.as_strided(ref[off:], (r, c), (8 * p, 8)))
    out = out0.copy()
    st = ctx.tensor_product_execute(pack_terms(b2g, terms, src, out), b2g.OPERANDS_HOST, b2g.DST_ZERO if zero else 0)
    assert rel(out, ref) < 1e-13
    assert st.entries == len(terms) and st.serial_entries == 0 and st.clusters < st.merged
    out2 = out0.copy()
    ctx.tensor_product_execute(pack_terms(b2g, terms, src, out2), b2g.OPERANDS_HOST, b2g.DST_ZERO if zero else 0)
    assert np.array_equal(out, out2)  # bit-reproducible


@pytest.mark.gpu
def test_gpu_resident_map_reads_and_writes_in_place(b2g, ctx):
    """b2g_resident_map: an input inside a mapped host range is read from its device shadow (the host copy is
    made to differ on purpose, to see which copy was read), an output inside a mapped range is written on
    the device only - the host range is never touched - and unmapped operands keep the mirrored route."""
    rng = np.random.default_rng(11)
    n = 5000
    src, one = rng.standard_normal(n), np.ones(1)
    mid, out = np.full(n, -1.0), np.zeros(n)
    p = lambda arr: np.array([arr.ctypes.data], dtype=np.uint64)
    axpy = lambda a, c, flags: ctx.batch_execute([111], [111], [n], [1], [1], [2.0], p(a), [1], p(one), [1], [1.0],
                                                 p(c), [1], [1], b2g.OPERANDS_HOST, flags)
    d_mid = ctx.malloc(8 * n)
    ctx.memset_zero(d_mid, 8 * n)
    hit0 = ctx.resident_stats()[0]
    ctx.resident_map(p(mid), [n], [d_mid])
    axpy(src, mid, b2g.DST_ZERO)            # output mapped: written in HBM, host copy untouched
    assert np.array_equal(mid, np.full(n, -1.0))
    axpy(mid, out, b2g.DST_ZERO)            # input mapped: read from HBM
    assert np.array_equal(out, 4.0 * src)
    assert ctx.resident_stats()[0] - hit0 == 8 * n
    back = np.zeros(n)
    ctx.download(p(back), [d_mid], [n])
    assert np.array_equal(back, 2.0 * src)
    ctx.resident_map([], [], [])            # cleared: the host copy is read again
    out[:] = 0.0
    axpy(mid, out, b2g.DST_ZERO)
    assert np.array_equal(out, np.full(n, -2.0))
    ctx.upload_blocks([d_mid], p(src), [n])
    ctx.download(p(back), [d_mid], [n])
    assert np.array_equal(back, src)
    with pytest.raises(b2g.B2GError, match="overlap"):
        ctx.resident_map([mid.ctypes.data, mid.ctypes.data + 8], [n, 4], [d_mid, d_mid])
    ctx.free(d_mid)


def test_upload_staging_slices_cover_every_byte(b2g):
    """the per-thread staging slices of the pageable upload path must cover the whole chunk
    for every length / thread count (floor division lost the tail when len / nt was a multiple of 4096)."""
    for nt in (1, 2, 3, 7, 8, 12, 16):
        for length in (0, 1, 4095, 4096, 4097, 9371656, 16 * 4096 * 143 + 8, (32 << 20) - 8, 32 << 20):
            lo, hi = b2g.upload_slices(length, nt)
            assert lo[0] == 0 and (hi >= lo).all()
            assert (lo[1:] == hi[:-1]).all() and hi[-1] == length, (nt, length, lo, hi)


def test_recorded_blocking_workloads_are_consistent(b2g):
    """The Cr2 SVP M=4000 blocking lists bench.py / tools/blocking_bench.py replay (structure only):
    every term is a 1 x 1 site block times an environment block (quantum-chemistry blocking), windows fit
    their output blocks, operands fit their arenas."""
    from conftest import ROOT
    for name, nterms in (("cr2_m4000_s20_call39.b2tp.gz", 66630), ("cr2_m4000_s20_call18.b2tp.gz", 89932)):
        tp = b2g.load_tpfile(os.path.join(ROOT, "workloads", "cr2_svp_m4000_blocking", name))
        T = tp.t
        assert tp.nterms == nterms and not tp.is_right
        assert (((T["am"] == 1) & (T["an"] == 1)) | ((T["bm"] == 1) & (T["bn"] == 1))).all()
        rows, cols = tp.window_shapes()
        assert (cols <= T["cn"]).all() and (rows >= 1).all()
        assert int((T["am"].astype(np.int64) * T["an"] * T["bm"] * T["bn"]).sum()) == tp.nflop
        a_off, b_off, c_off, n_in, n_out = tp.offsets()
        assert (a_off >= 0).all() and (a_off + T["am"].astype(np.int64) * T["an"] <= n_in).all()
        assert (b_off + T["bm"].astype(np.int64) * T["bn"] <= n_in).all()
        assert (c_off + (rows - 1) * T["cn"] + cols <= n_out).all()
        # terms writing the same window have the same shape (identical or disjoint windows only)
        order = np.argsort(c_off, kind="stable")
        same = c_off[order][1:] == c_off[order][:-1]
        assert (rows[order][1:][same] == rows[order][:-1][same]).all()
        assert (cols[order][1:][same] == cols[order][:-1][same]).all()


def test_plan_of_the_recorded_heff_blocking_list_keeps_descriptors_few(b2g):
    """Host-side regrouping (B2G_PLAN_ONLY, no GPU) of the Cr2 SVP M=4000 H_eff blocking list: 66 630 terms fold
    into 58 778 windows without serial components, and the work is handed to the device in units of whole rows /
    whole tile rows - round 1 made one descriptor per window row (1.40 M for this list), which cost 10-50x the
    kernel time to build and upload."""
    from conftest import ROOT
    tp = b2g.load_tpfile(os.path.join(ROOT, "workloads", "cr2_svp_m4000_blocking", "cr2_m4000_s20_call39.b2tp.gz"))
    a_off, b_off, c_off, n_in, n_out = tp.offsets()
    terms = np.zeros(tp.nterms, dtype=b2g.TP_DTYPE)
    terms["a"], terms["b"], terms["c"] = (1 << 40) + 8 * a_off, (1 << 40) + 8 * b_off, (1 << 44) + 8 * c_off
    for k in ("am", "an", "bm", "bn", "cn", "conja", "conjb", "scale"):
        terms[k] = tp.t[k]
    st = b2g.tensor_product_plan(terms, b2g.DST_ZERO)
    assert st.entries == tp.nterms and st.clusters == 58778 and st.serial_entries == 0
    assert 50_000 < st.units < 250_000, st.units
    assert st.bytes_out == 8 * int(tp.out_sizes.sum()) or st.bytes_out > 0


# ----------------------------------------------------------------------------- host-side regrouping (no GPU)


@pytest.mark.parametrize("name", BLK_FILES)
def test_plan_regroups_reference_lists_without_serial_components(b2g, name):
    """B2G_PLAN_ONLY: the regrouping by output window runs on the host.  On the lists the reference's own
    recorder makes, windows are identical or disjoint (no serial components), the per-row GEMM groups fold
    into fewer 2-D windows, and the bookkeeping totals are the reference's (nflop = sum m*n*k*group size)."""
    bf = b2g.load_blkfile(os.path.join(GOLDEN, name))
    inp, out = bf.inputs.copy(), bf.c_in.copy()
    a, b, c = bf.pointers(inp.ctypes.data, out.ctypes.data)
    ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, gs = bf.group_args()
    st = b2g.batch_plan(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs)
    assert st.entries == bf.nentries and st.nflop_mnk == bf.nflop_mnk
    assert st.serial_entries == 0 and 0 < st.clusters <= st.merged < st.entries
    assert st.units >= st.clusters and st.launches == 0
    # every output element is written once, every source element of every entry read once
    G = bf.g
    a_elems = int((G["m"].astype(np.int64) * G["k"] * G["gp"]).sum())
    b_elems = int((G["k"].astype(np.int64) * G["n"] * G["gp"]).sum())
    assert 8 * a_elems < st.bytes_in <= 8 * (a_elems + b_elems)  # one scalar per folded window, not per row
    assert st.bytes_out <= 8 * out.size
    again = b2g.batch_plan(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs)
    assert (again.clusters, again.merged, again.units) == (st.clusters, st.merged, st.units)
    assert np.array_equal(out, bf.c_in)  # nothing executed


@pytest.mark.parametrize("seed,overlap", [(0, False), (1, False), (2, True), (3, True)])
def test_plan_detects_overlapping_windows(b2g, seed, overlap):
    rng = np.random.default_rng(seed)
    G, A, B, C, src, scal, out = synthetic_list(rng, general=True, overlap=overlap)
    st = b2g.batch_plan(*resolve(G, A, B, C, src, scal, out))
    assert st.entries == len(A)
    assert (st.serial_entries > 0) == overlap


def test_plan_of_term_lists(b2g):
    rng = np.random.default_rng(5)
    terms, src, osize = synthetic_terms(rng, kron=True)
    out = np.zeros(osize)
    st = b2g.tensor_product_plan(pack_terms(b2g, terms, src, out))
    assert st.entries == len(terms) and st.serial_entries == 0 and st.clusters < st.merged
    # a Kronecker term becomes one window per element of op(A); scalar terms stay one window each
    expect = sum(t["am"] * t["an"] if (t["am"] * t["an"] > 1 and t["bm"] * t["bn"] > 1) else 1 for t in terms)
    assert st.merged == expect
    assert st.nflop_mnk == sum(t["am"] * t["an"] * t["bm"] * t["bn"] for t in terms)


# ----------------------------------------------------------------------------- term form of the fixtures


def fixture_terms(bf):
    T = bf.terms
    a, b, c = bf.term_offsets()
    return [dict(a=int(a[i]), b=int(b[i]), c=int(c[i]), am=int(T["am"][i]), an=int(T["an"][i]), bm=int(T["bm"][i]),
                 bn=int(T["bn"][i]), cn=int(T["cn"][i]), conja=int(T["conja"][i]), conjb=int(T["conjb"][i]),
                 scale=float(T["scale"][i])) for i in range(len(a))]


@pytest.mark.parametrize("name", BLK_FILES)
def test_term_oracle_matches_reference_blocking(b2g, name):
    """The same left_contract / right_contract call in the term form the host binding records (one b2g_tp_term
    per connection-info entry): the numpy restatement of GMatrixFunctions::tensor_product reproduces the
    blocked operators of the reference's own executor bit for bit (pinned)."""
    bf = b2g.load_blkfile(os.path.join(GOLDEN, name))
    assert bf.terms is not None and 0 < len(bf.terms["am"]) < bf.nentries
    terms = fixture_terms(bf)
    out = bf.c_in.copy()
    sd.tensor_product_terms(terms, lambda off, n: bf.inputs[off:off + n],
                            lambda off, r, c, p: np.lib.stride_tricks.as_strided(out[off:], (r, c), (8 * p, 8)))
    assert np.array_equal(out, bf.c_ref)
    # and the regrouping of the term form sees the same windows as that of the list form
    inp = bf.inputs.copy()
    st_t = b2g.tensor_product_plan(pack_terms(b2g, terms, inp, out))
    a, b, c = bf.pointers(inp.ctypes.data, out.ctypes.data)
    ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, gs = bf.group_args()
    st_l = b2g.batch_plan(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, gs)
    assert st_t.entries == len(terms) and st_t.serial_entries == 0 and st_l.serial_entries == 0
    assert st_t.nflop_mnk == st_l.nflop_mnk == bf.nflop_mnk
    assert st_t.bytes_out == st_l.bytes_out


@pytest.mark.gpu
@pytest.mark.parametrize("name", BLK_FILES)
def test_gpu_term_form_matches_reference_blocking(b2g, ctx, name):
    bf = b2g.load_blkfile(os.path.join(GOLDEN, name))
    inp, out = bf.inputs.copy(), bf.c_in.copy()
    st = ctx.tensor_product_execute(pack_terms(b2g, fixture_terms(bf), inp, out), b2g.OPERANDS_HOST, b2g.DST_ZERO)
    assert rel(out, bf.c_ref) < TOL
    assert st.entries == len(bf.terms["am"]) and st.serial_entries == 0
    assert np.array_equal(inp, bf.inputs)
