"""oracle/seqdump.py — TEST INFRASTRUCTURE, NOT PRODUCT.

Reader for the `.b2seq` files written by oracle/ref_harness.cpp (the
unmodified reference stopped right after EffectiveHamiltonian::precompute(),
/root/reference/src/dmrg/effective_hamiltonian.hpp:226) and ctypes access to
the plain-C restatement in oracle/replay.c.

File layout (little endian):
    8   bytes  magic  b"B2SEQ\\0\\0\\2"
    16  u64    npairs, n_arenas, csize, vsize, max_work, nflop (reference
               units: m*n*k), has_data, site, bond_dim, n_sites, ndav_ref,
               has_eigs, 4 reserved
    8   f64    e_ref (Davidson eigenvalue without const_e), const_e,
               t_ref_matvec (s), davidson conv_thrd, 4 reserved
    16  i32[npairs]  ta0 tb0 m0 n0 k0 lda0 ldb0 ldc0 ta1 tb1 m1 n1 k1 lda1 ldb1 ldc1
    4   f64[npairs]  alpha0 beta0 alpha1 beta1
    7   i64[npairs]  a0_off  b0_arena b0_off  a1_arena a1_off  c1_off  w_off
        u64[n_arenas] arena sizes (doubles)
    if has_data: arenas (f64, concatenated), c[csize], v_ref[vsize],
                 diag[csize], ket0[csize]

Pair i means (batch_gemm.hpp:564-575, 1634-1643; all matrices row-major):
    W           = alpha0 * op(c[a0_off:])        * op(arena[b0_arena][b0_off:])   (beta0 = 0)
    v[c1_off:] += alpha1 * op(arena[a1_arena][a1_off:]) * W                         (beta1 = 1)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
I32_NAMES = ["ta0", "tb0", "m0", "n0", "k0", "lda0", "ldb0", "ldc0",
             "ta1", "tb1", "m1", "n1", "k1", "lda1", "ldb1", "ldc1"]
F64_NAMES = ["alpha0", "beta0", "alpha1", "beta1"]
I64_NAMES = ["a0_off", "b0_arena", "b0_off", "a1_arena", "a1_off", "c1_off", "w_off"]


@dataclass
class SeqDump:
    npairs: int
    csize: int
    vsize: int
    max_work: int
    nflop_mnk: int
    site: int
    bond_dim: int
    n_sites: int
    ndav_ref: int
    has_eigs: bool
    e_ref: float
    const_e: float
    t_ref_matvec: float
    conv_thrd: float
    arena_sizes: np.ndarray
    p: dict = field(default_factory=dict)      # per-pair arrays by name
    arenas: np.ndarray | None = None            # all operand doubles, concatenated
    c: np.ndarray | None = None
    v_ref: np.ndarray | None = None
    diag: np.ndarray | None = None
    ket0: np.ndarray | None = None

    @property
    def arena_starts(self) -> np.ndarray:
        s = np.zeros(len(self.arena_sizes) + 1, dtype=np.int64)
        np.cumsum(self.arena_sizes, out=s[1:])
        return s

    @property
    def flops(self) -> float:
        """2*m*n*k FLOPs of one matvec (the reference counts m*n*k)."""
        return 2.0 * self.nflop_mnk

    def operand_offsets(self):
        """Offsets (doubles) of the b0 / a1 operator blocks inside `arenas`."""
        st = self.arena_starts
        return st[self.p["b0_arena"]] + self.p["b0_off"], st[self.p["a1_arena"]] + self.p["a1_off"]


def load(path: str) -> SeqDump:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"B2SEQ\0\0\2":
        raise ValueError(f"{path}: bad magic {raw[:8]!r}")
    pos = 8
    hdr = np.frombuffer(raw, dtype="<u8", count=16, offset=pos); pos += 128
    dh = np.frombuffer(raw, dtype="<f8", count=8, offset=pos); pos += 64
    n, na = int(hdr[0]), int(hdr[1])
    p = {}
    for nm in I32_NAMES:
        p[nm] = np.frombuffer(raw, dtype="<i4", count=n, offset=pos).copy(); pos += 4 * n
    for nm in F64_NAMES:
        p[nm] = np.frombuffer(raw, dtype="<f8", count=n, offset=pos).copy(); pos += 8 * n
    for nm in I64_NAMES:
        p[nm] = np.frombuffer(raw, dtype="<i8", count=n, offset=pos).copy(); pos += 8 * n
    asz = np.frombuffer(raw, dtype="<u8", count=na, offset=pos).astype(np.int64); pos += 8 * na
    d = SeqDump(npairs=n, csize=int(hdr[2]), vsize=int(hdr[3]), max_work=int(hdr[4]),
                nflop_mnk=int(hdr[5]), site=int(hdr[7]), bond_dim=int(hdr[8]),
                n_sites=int(hdr[9]), ndav_ref=int(hdr[10]), has_eigs=bool(hdr[11]),
                e_ref=float(dh[0]), const_e=float(dh[1]), t_ref_matvec=float(dh[2]),
                conv_thrd=float(dh[3]), arena_sizes=asz, p=p)
    if int(hdr[6]):
        tot = int(asz.sum())
        d.arenas = np.frombuffer(raw, dtype="<f8", count=tot, offset=pos).copy(); pos += 8 * tot
        d.c = np.frombuffer(raw, dtype="<f8", count=d.csize, offset=pos).copy(); pos += 8 * d.csize
        d.v_ref = np.frombuffer(raw, dtype="<f8", count=d.vsize, offset=pos).copy(); pos += 8 * d.vsize
        d.diag = np.frombuffer(raw, dtype="<f8", count=d.csize, offset=pos).copy(); pos += 8 * d.csize
        d.ket0 = np.frombuffer(raw, dtype="<f8", count=d.csize, offset=pos).copy(); pos += 8 * d.csize
    return d


def synth_operands(d: SeqDump, seed: int = 0) -> None:
    """Fill a structure-only dump with seeded synthetic operator data."""
    rng = np.random.default_rng(seed)
    d.arenas = rng.standard_normal(int(d.arena_sizes.sum()))
    d.c = rng.standard_normal(d.csize)


# ---------------------------------------------------------------- C oracle
_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "_ref", "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle _ref/liboracle.so` "
                                    "(or __graft_entry__.build())")
        _lib = ctypes.CDLL(path)
    return _lib


def _ptr(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def replay(d: SeqDump, c: np.ndarray | None = None, scale: float = 1.0,
           nthreads: int = 1, arenas: np.ndarray | None = None) -> np.ndarray:
    """sigma = H.c through oracle/replay.c (b2o_seq_matvec)."""
    c = np.ascontiguousarray(d.c if c is None else c, dtype=np.float64)
    ar = np.ascontiguousarray(d.arenas if arenas is None else arenas, dtype=np.float64)
    b0o, a1o = d.operand_offsets()
    base = ar.ctypes.data
    b0 = (base + 8 * b0o).astype(np.uint64)
    a1 = (base + 8 * a1o).astype(np.uint64)
    v = np.zeros(d.vsize, dtype=np.float64)
    P = d.p
    i32, f64, i64 = ctypes.c_int32, ctypes.c_double, ctypes.c_int64
    vpp = ctypes.POINTER(ctypes.c_void_p)
    L = lib()
    L.b2o_seq_matvec.restype = None
    L.b2o_seq_matvec(
        i64(d.npairs),
        _ptr(P["ta0"], i32), _ptr(P["tb0"], i32), _ptr(P["m0"], i32), _ptr(P["n0"], i32),
        _ptr(P["k0"], i32), _ptr(P["lda0"], i32), _ptr(P["ldb0"], i32), _ptr(P["ldc0"], i32),
        _ptr(P["alpha0"], f64), _ptr(P["beta0"], f64), _ptr(P["a0_off"], i64),
        b0.ctypes.data_as(vpp),
        _ptr(P["ta1"], i32), _ptr(P["tb1"], i32), _ptr(P["m1"], i32), _ptr(P["n1"], i32),
        _ptr(P["k1"], i32), _ptr(P["lda1"], i32), _ptr(P["ldb1"], i32), _ptr(P["ldc1"], i32),
        _ptr(P["alpha1"], f64), _ptr(P["beta1"], f64), a1.ctypes.data_as(vpp),
        _ptr(P["c1_off"], i64), i64(max(d.max_work, 1)), _ptr(c, f64), _ptr(v, f64),
        i64(d.vsize), f64(scale), ctypes.c_int(nthreads))
    return v


def _openblas_dgemm():
    """Address of dgemm_ in the OpenBLAS copied next to the reference binaries
    (oracle/_ref) or the scipy-bundled one (same library)."""
    import glob
    cands = glob.glob(os.path.join(_HERE, "_ref", "libscipy_openblas*.so"))
    if not cands:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))
    L = ctypes.CDLL(cands[0], mode=ctypes.RTLD_GLOBAL)
    try:
        L.scipy_openblas_set_num_threads(1)   # one BLAS thread per OpenMP thread, like the reference runs
    except AttributeError:
        pass
    return ctypes.cast(L.scipy_dgemm_, ctypes.c_void_p), L


def replay_blas(d: SeqDump, c: np.ndarray | None = None, scale: float = 1.0, nthreads: int = 1,
                arenas: np.ndarray | None = None, v: np.ndarray | None = None) -> np.ndarray:
    """sigma = H.c with every GEMM done by OpenBLAS dgemm_, OpenMP over pairs with private
    sigma replicas - the reference's Tasked executor (batch_gemm.hpp:1613-1688). CPU baseline."""
    c = np.ascontiguousarray(d.c if c is None else c, dtype=np.float64)
    ar = np.ascontiguousarray(d.arenas if arenas is None else arenas, dtype=np.float64)
    b0o, a1o = d.operand_offsets()
    base = ar.ctypes.data
    b0 = (base + 8 * b0o).astype(np.uint64)
    a1 = (base + 8 * a1o).astype(np.uint64)
    if v is None:
        v = np.zeros(d.vsize, dtype=np.float64)
    P = d.p
    i32, f64, i64 = ctypes.c_int32, ctypes.c_double, ctypes.c_int64
    vpp = ctypes.POINTER(ctypes.c_void_p)
    fn, keep = _openblas_dgemm()
    L = lib()
    L.b2o_seq_matvec_blas.restype = None
    L.b2o_seq_matvec_blas(
        fn, i64(d.npairs),
        _ptr(P["ta0"], i32), _ptr(P["tb0"], i32), _ptr(P["m0"], i32), _ptr(P["n0"], i32),
        _ptr(P["k0"], i32), _ptr(P["lda0"], i32), _ptr(P["ldb0"], i32), _ptr(P["ldc0"], i32),
        _ptr(P["alpha0"], f64), _ptr(P["beta0"], f64), _ptr(P["a0_off"], i64),
        b0.ctypes.data_as(vpp),
        _ptr(P["ta1"], i32), _ptr(P["tb1"], i32), _ptr(P["m1"], i32), _ptr(P["n1"], i32),
        _ptr(P["k1"], i32), _ptr(P["lda1"], i32), _ptr(P["ldb1"], i32), _ptr(P["ldc1"], i32),
        _ptr(P["alpha1"], f64), _ptr(P["beta1"], f64), a1.ctypes.data_as(vpp),
        _ptr(P["c1_off"], i64), i64(max(d.max_work, 1)), _ptr(c, f64), _ptr(v, f64),
        i64(d.vsize), f64(scale), ctypes.c_int(nthreads))
    return v


def replay_numpy(d: SeqDump, c: np.ndarray | None = None, scale: float = 1.0,
                 arenas: np.ndarray | None = None) -> np.ndarray:
    """Independent pure-numpy restatement (small cases only)."""
    c = d.c if c is None else c
    ar = d.arenas if arenas is None else arenas
    b0o, a1o = d.operand_offsets()
    P = d.p
    v = np.zeros(d.vsize)

    def view(buf, off, rows, cols, ld):
        return np.lib.stride_tricks.as_strided(buf[off:], shape=(rows, cols), strides=(8 * ld, 8))

    for i in range(d.npairs):
        m0, n0, k0 = int(P["m0"][i]), int(P["n0"][i]), int(P["k0"][i])
        A = view(c, int(P["a0_off"][i]), k0, m0, int(P["lda0"][i])).T if P["ta0"][i] else \
            view(c, int(P["a0_off"][i]), m0, k0, int(P["lda0"][i]))
        B = view(ar, int(b0o[i]), n0, k0, int(P["ldb0"][i])).T if P["tb0"][i] else \
            view(ar, int(b0o[i]), k0, n0, int(P["ldb0"][i]))
        W = P["alpha0"][i] * (A @ B)
        m1, n1, k1 = int(P["m1"][i]), int(P["n1"][i]), int(P["k1"][i])
        assert (k1, n1) == (m0, n0) and not P["tb1"][i] and P["ldb1"][i] == n0
        A1 = view(ar, int(a1o[i]), k1, m1, int(P["lda1"][i])).T if P["ta1"][i] else \
            view(ar, int(a1o[i]), m1, k1, int(P["lda1"][i]))
        C = view(v, int(P["c1_off"][i]), m1, n1, int(P["ldc1"][i]))
        C += (P["alpha1"][i] * scale) * (A1 @ W)
    return v


def batch_perform(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size) -> None:
    """Sequential execution of a grouped GEMM list on host pointers (integer addresses) through
    the plain-C restatement b2o_dgemm_batch (oracle/replay.c; cblas_xgemm_batch semantics,
    batch_gemm.hpp:81-111): entry by entry in list order, which is what BatchGEMMSeq::simple_perform
    produces and what auto_perform reproduces up to the order of additions."""
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    f64 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
    ta, tb, m, n, k, lda, ldb, ldc, gs = map(i32, (np.asarray(ta) == 112, np.asarray(tb) == 112, m, n, k, lda, ldb,
                                                   ldc, group_size))
    alpha, beta, a, b, c = f64(alpha), f64(beta), u64(a), u64(b), u64(c)
    L = lib()
    L.b2o_dgemm_batch.restype = None
    L.b2o_dgemm_batch.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 14
    L.b2o_dgemm_batch(len(gs), ta.ctypes.data, tb.ctypes.data, m.ctypes.data, n.ctypes.data, k.ctypes.data,
                      alpha.ctypes.data, a.ctypes.data, lda.ctypes.data, b.ctypes.data, ldb.ctypes.data,
                      beta.ctypes.data, c.ctypes.data, ldc.ctypes.data, gs.ctypes.data)


def tensor_product_terms(terms, read, write) -> None:
    """numpy restatement of a list of eager GMatrixFunctions<double>::tensor_product calls
    (matrix_functions.hpp:1269-1397): for every term, in list order,
        C[(i*bm' + k), (j*bn' + l)] += scale * op(A)(i, j) * op(B)(k, l)
    terms: iterable of dicts (a, b, c = element offsets; am an bm bn cn conja conjb scale);
    read(off, n) -> view of n source doubles; write(off, rows, cols, pitch) -> 2-D output view."""
    for t in terms:
        A = read(t["a"], t["am"] * t["an"]).reshape(t["am"], t["an"])
        B = read(t["b"], t["bm"] * t["bn"]).reshape(t["bm"], t["bn"])
        if t["conja"]:
            A = A.T
        if t["conjb"]:
            B = B.T
        K = np.kron(A, B)
        W = write(t["c"], K.shape[0], K.shape[1], t["cn"])
        W += t["scale"] * K
