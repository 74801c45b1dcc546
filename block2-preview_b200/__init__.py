"""block2-preview_b200 — B200-native executor for block2's Davidson H.C hot path.

Python host side of the C ABI in include/b2g.h (ctypes; the library is
block2-preview_b200/libb2g.so, built by `make lib` / __graft_entry__.build()).
There is no CPU fallback: every compute entry point raises when the CUDA
library or a GPU is missing.

Names follow the reference's operator surface:
    SeqPlan            <-> BatchGEMMSeq<double> after EffectiveHamiltonian::precompute()
    SeqPlan.__call__   <-> BatchGEMMSeq::operator()(c, v, scale)   (core/batch_gemm.hpp:1570)
    SeqPlan.davidson   <-> EffectiveHamiltonian::eigs -> IterativeMatrixFunctions::davidson
    dgemm_batch        <-> cblas_xgemm_batch / BatchGEMM::perform  (core/batch_gemm.hpp:81-111)
    Context.batch_execute <-> BatchGEMMSeq::auto_perform on the blocking list of
                           TensorFunctions::left_contract / right_contract (core/tensor_functions.hpp:2842, 2941)

The directory name is not an importable identifier; load it with `import b2gpkg`
(repo root), which registers this package as `block2_preview_b200`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

from .blkfile import BlkFile, load_blkfile  # noqa: F401
from .seqfile import SeqFile, load_seqfile  # noqa: F401
from .tpfile import TPFile, load_tpfile  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2g.so")

B2G_NOTRANS, B2G_TRANS = 111, 112
OPERANDS_HOST, OPERANDS_DEVICE = 0, 1

EXPORTS = [
    "b2g_last_error", "b2g_device_count", "b2g_context_create", "b2g_context_destroy",
    "b2g_context_launches", "b2g_context_stream", "b2g_context_synchronize",
    "b2g_plan_create", "b2g_plan_destroy", "b2g_plan_get_stats", "b2g_seq_matvec",
    "b2g_seq_matvec_dev", "b2g_plan_profile", "b2g_pairs_execute", "b2g_dgemm_batch", "b2g_batch_execute", "b2g_tensor_product_execute", "b2g_resident_map", "b2g_resident_stats", "b2g_download",
    "b2g_upload_blocks", "b2g_host_register", "b2g_host_unregister", "b2g_mem_info", "b2g_debug_upload_slices", "b2g_davidson", "b2g_comm_unique_id",
    "b2g_comm_init", "b2g_comm_destroy", "b2g_allreduce_sum", "b2g_malloc", "b2g_free",
    "b2g_memcpy_h2d", "b2g_memcpy_d2h", "b2g_memset_zero",
    "b2g_prof_enabled", "b2g_prof_record", "b2g_prof_dump", "b2g_debug_tiled_plan", "b2g_mem_trim", "b2g_syevd",
]


class B2GError(RuntimeError):
    pass


class _Batch(ctypes.Structure):
    _fields_ = [("count", c_int64),
                ("ta", POINTER(c_int32)), ("tb", POINTER(c_int32)),
                ("m", POINTER(c_int32)), ("n", POINTER(c_int32)), ("k", POINTER(c_int32)),
                ("lda", POINTER(c_int32)), ("ldb", POINTER(c_int32)), ("ldc", POINTER(c_int32)),
                ("alpha", POINTER(c_double)), ("beta", POINTER(c_double)),
                ("a", POINTER(c_void_p)), ("b", POINTER(c_void_p)), ("c", POINTER(c_void_p))]


class PlanStats(ctypes.Structure):
    _fields_ = [("pairs", c_int64), ("csize", c_int64), ("vsize", c_int64), ("nflop_mnk", c_int64),
                ("operand_doubles", c_int64), ("arenas", c_int64), ("launches", c_int64),
                ("n_small", c_int64), ("n_large", c_int64), ("upload_seconds", c_double),
                ("mirrored_doubles", c_int64), ("workspace_doubles", c_int64)]


class BlockingStats(ctypes.Structure):
    _fields_ = [("entries", c_int64), ("merged", c_int64), ("clusters", c_int64), ("units", c_int64),
                ("serial_entries", c_int64), ("nflop_mnk", c_int64), ("bytes_in", c_int64), ("bytes_out", c_int64),
                ("launches", c_int64), ("kernel_ms", c_double), ("upload_seconds", c_double),
                ("download_seconds", c_double), ("plan_seconds", c_double)]


DST_ZERO = 1
PLAN_ONLY = 8


class TPTerm(ctypes.Structure):
    """b2g_tp_term: one GMatrixFunctions::tensor_product call of the blocking step."""
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("c", c_void_p), ("am", c_int32), ("an", c_int32),
                ("bm", c_int32), ("bn", c_int32), ("cn", c_int32), ("conja", c_int32), ("conjb", c_int32),
                ("reserved", c_int32), ("scale", c_double)]


TP_DTYPE = np.dtype([("a", np.uint64), ("b", np.uint64), ("c", np.uint64), ("am", np.int32), ("an", np.int32),
                     ("bm", np.int32), ("bn", np.int32), ("cn", np.int32), ("conja", np.int32), ("conjb", np.int32),
                     ("reserved", np.int32), ("scale", np.float64)], align=True)


class KernelStat(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("flops", c_double), ("ms", c_double), ("units", c_int64)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load libb2g.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2GError(f"{LIB_PATH} not found: build it with `make lib` or __graft_entry__.build(); "
                           "there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.b2g_last_error.restype = c_char_p
        L.b2g_context_launches.restype = c_int64
        L.b2g_context_stream.restype = c_void_p
        L.b2g_context_launches.argtypes = [c_void_p]
        L.b2g_context_stream.argtypes = [c_void_p]
        L.b2g_context_create.argtypes = [c_int, POINTER(c_void_p)]
        L.b2g_context_destroy.argtypes = [c_void_p]
        L.b2g_context_synchronize.argtypes = [c_void_p]
        L.b2g_plan_create.argtypes = [c_void_p, POINTER(_Batch), POINTER(_Batch), c_int64, c_int64, c_int64,
                                      c_int, POINTER(c_void_p)]
        L.b2g_plan_destroy.argtypes = [c_void_p]
        L.b2g_plan_get_stats.argtypes = [c_void_p, POINTER(PlanStats)]
        L.b2g_seq_matvec.argtypes = [c_void_p, c_void_p, c_void_p, c_double]
        L.b2g_seq_matvec_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_double]
        L.b2g_plan_profile.argtypes = [c_void_p, c_void_p, c_void_p, c_double, POINTER(KernelStat), c_int,
                                       POINTER(c_int)]
        L.b2g_pairs_execute.argtypes = [c_void_p, POINTER(_Batch), POINTER(_Batch), c_int64, POINTER(PlanStats)]
        L.b2g_dgemm_batch.argtypes = [c_void_p, c_int64] + [c_void_p] * 13
        L.b2g_batch_execute.argtypes = [c_void_p, c_int64] + [c_void_p] * 14 + [c_int, c_int, POINTER(BlockingStats)]
        L.b2g_tensor_product_execute.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_int, POINTER(BlockingStats)]
        L.b2g_resident_map.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        L.b2g_resident_stats.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
        L.b2g_download.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        L.b2g_upload_blocks.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        L.b2g_host_register.argtypes = [c_void_p, c_void_p, c_size_t]
        L.b2g_host_unregister.argtypes = [c_void_p, c_void_p]
        L.b2g_mem_info.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64)]
        L.b2g_debug_upload_slices.argtypes = [c_int64, c_int, c_void_p, c_void_p]
        L.b2g_davidson.argtypes = [c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_int, c_int,
                                   POINTER(c_double), POINTER(c_int)]
        L.b2g_comm_unique_id.argtypes = [c_void_p]
        L.b2g_comm_init.argtypes = [c_void_p, c_int, c_int, c_void_p]
        L.b2g_comm_destroy.argtypes = [c_void_p]
        L.b2g_allreduce_sum.argtypes = [c_void_p, c_void_p, c_int64]
        L.b2g_malloc.argtypes = [c_void_p, c_size_t, POINTER(c_void_p)]
        L.b2g_free.argtypes = [c_void_p, c_void_p]
        L.b2g_memcpy_h2d.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
        L.b2g_memcpy_d2h.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
        L.b2g_memset_zero.argtypes = [c_void_p, c_void_p, c_size_t]
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise B2GError(f"{what} failed (code {rc}): {lib().b2g_last_error().decode()}")


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptrs(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(arr: np.ndarray, ct):
    return arr.ctypes.data_as(POINTER(ct))


def _make_batches(batch0: dict, batch1: dict):
    structs, keep = [], []
    for bt in (batch0, batch1):
        arrs = {k: _i32(bt[k]) for k in ("ta", "tb", "m", "n", "k", "lda", "ldb", "ldc")}
        arrs.update({k: _f64(bt[k]) for k in ("alpha", "beta")})
        arrs.update({k: _ptrs(bt[k]) for k in ("a", "b", "c")})
        keep.append(arrs)
        s = _Batch()
        s.count = len(arrs["m"])
        for k in ("ta", "tb", "m", "n", "k", "lda", "ldb", "ldc"):
            setattr(s, k, _p(arrs[k], c_int32))
        s.alpha, s.beta = _p(arrs["alpha"], c_double), _p(arrs["beta"], c_double)
        for k in ("a", "b", "c"):
            setattr(s, k, ctypes.cast(arrs[k].ctypes.data, POINTER(c_void_p)))
        structs.append(s)
    return structs, keep


class Context:
    """One GPU, one stream (one process per GPU under torchrun)."""

    def __init__(self, device: int = 0):
        self._h = c_void_p()
        _check(lib().b2g_context_create(device, byref(self._h)), "b2g_context_create")
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().b2g_context_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(lib().b2g_context_launches(self._h))

    @property
    def stream(self) -> int:
        return int(lib().b2g_context_stream(self._h) or 0)

    def synchronize(self) -> None:
        _check(lib().b2g_context_synchronize(self._h), "b2g_context_synchronize")

    # --- multi-GPU (NCCL) ---
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(lib().b2g_comm_unique_id(buf), "b2g_comm_unique_id")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes) -> None:
        assert len(uid) == 128
        _check(lib().b2g_comm_init(self._h, nranks, rank, ctypes.c_char_p(uid)), "b2g_comm_init")

    def allreduce_sum(self, dev_ptr: int, count: int) -> None:
        _check(lib().b2g_allreduce_sum(self._h, c_void_p(dev_ptr), count), "b2g_allreduce_sum")

    def pairs_execute(self, batch0: dict, batch1: dict, max_work: int) -> PlanStats:
        """Run a chained-pair list with ALL operands in host memory once (tensor_rotate lists):
        W_i = alpha0*op(A0_i)*op(B0_i); C1_i += alpha1*op(A1_i)*W_i, results added into the host C1."""
        structs, keep = _make_batches(batch0, batch1)
        st = PlanStats()
        _check(lib().b2g_pairs_execute(self._h, byref(structs[0]), byref(structs[1]), max_work, byref(st)),
               "b2g_pairs_execute")
        return st

    def dgemm_batch(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size) -> None:
        """Grouped GEMM list on device pointers (cblas_dgemm_batch signature)."""
        ta, tb, m, n, k, lda, ldb, ldc, gs = map(_i32, (ta, tb, m, n, k, lda, ldb, ldc, group_size))
        alpha, beta = _f64(alpha), _f64(beta)
        a, b, c = _ptrs(a), _ptrs(b), _ptrs(c)
        _check(lib().b2g_dgemm_batch(self._h, len(gs), ta.ctypes.data, tb.ctypes.data, m.ctypes.data,
                                     n.ctypes.data, k.ctypes.data, alpha.ctypes.data, a.ctypes.data,
                                     lda.ctypes.data, b.ctypes.data, ldb.ctypes.data, beta.ctypes.data,
                                     c.ctypes.data, ldc.ctypes.data, gs.ctypes.data), "b2g_dgemm_batch")


def _batch_execute(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size, operand_space, flags):
    ta, tb, m, n, k, lda, ldb, ldc, gs = map(_i32, (ta, tb, m, n, k, lda, ldb, ldc, group_size))
    alpha, beta = _f64(alpha), _f64(beta)
    a, b, c = _ptrs(a), _ptrs(b), _ptrs(c)
    st = BlockingStats()
    _check(lib().b2g_batch_execute(ctx._h, len(gs), ta.ctypes.data, tb.ctypes.data, m.ctypes.data, n.ctypes.data,
                                   k.ctypes.data, alpha.ctypes.data, a.ctypes.data, lda.ctypes.data, b.ctypes.data,
                                   ldb.ctypes.data, beta.ctypes.data, c.ctypes.data, ldc.ctypes.data,
                                   gs.ctypes.data, operand_space, flags, byref(st)), "b2g_batch_execute")
    return st


def _ctx_batch_execute(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size,
                       operand_space: int = OPERANDS_HOST, flags: int = 0) -> BlockingStats:
    """Run a conflict-carrying single-batch GEMM list (the blocking list of left_contract /
    right_contract, cblas_dgemm_batch group signature) once; contributions to one output element
    are applied in list order.  Host operands: results land in the host blocks."""
    return _batch_execute(self, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size, operand_space,
                          flags)


Context.batch_execute = _ctx_batch_execute


class _NoContext:
    """Stand-in for B2G_PLAN_ONLY calls, which need no device."""
    _h = None


def batch_plan(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size, flags: int = 0) -> BlockingStats:
    """Host-side regrouping of a blocking list only (B2G_PLAN_ONLY): entries, folded windows, clusters,
    warp units, serial components and algorithmic bytes, without a GPU."""
    return _batch_execute(_NoContext, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, group_size,
                          OPERANDS_HOST, flags | PLAN_ONLY)


def tensor_product_plan(terms: np.ndarray, flags: int = 0) -> BlockingStats:
    terms = np.ascontiguousarray(terms, dtype=TP_DTYPE)
    st = BlockingStats()
    _check(lib().b2g_tensor_product_execute(None, len(terms), terms.ctypes.data, OPERANDS_HOST, flags | PLAN_ONLY,
                                            byref(st)), "b2g_tensor_product_execute")
    return st


def _ctx_tensor_product_execute(self, terms: np.ndarray, operand_space: int = OPERANDS_HOST,
                                flags: int = 0) -> BlockingStats:
    """terms: structured array of TP_DTYPE (b2g_tp_term): C window += scale * op(A) (x) op(B) each,
    same-window terms applied in list order."""
    terms = np.ascontiguousarray(terms, dtype=TP_DTYPE)
    assert TP_DTYPE.itemsize == ctypes.sizeof(TPTerm)
    st = BlockingStats()
    _check(lib().b2g_tensor_product_execute(self._h, len(terms), terms.ctypes.data, operand_space, flags, byref(st)),
           "b2g_tensor_product_execute")
    return st


Context.tensor_product_execute = _ctx_tensor_product_execute


def _ctx_resident_map(self, host_ptrs, doubles, dev_ptrs) -> None:
    """Device-resident operands: host ranges whose content lives at the given device addresses.  Host-pointer
    entry points read such inputs in place and write such outputs in place (not copied back) until the table
    is replaced; an empty table clears it."""
    hp, nd, dp = _ptrs(host_ptrs), np.ascontiguousarray(doubles, dtype=np.int64), _ptrs(dev_ptrs)
    assert len(hp) == len(nd) == len(dp)
    _check(lib().b2g_resident_map(self._h, len(hp), hp.ctypes.data, nd.ctypes.data, dp.ctypes.data),
           "b2g_resident_map")


def _ctx_resident_stats(self):
    """(host->device bytes avoided by resident inputs, bytes mirrored from the host) so far."""
    hit, mir = c_int64(), c_int64()
    _check(lib().b2g_resident_stats(self._h, byref(hit), byref(mir)), "b2g_resident_stats")
    return hit.value, mir.value


def _ctx_malloc(self, nbytes: int) -> int:
    p = c_void_p()
    _check(lib().b2g_malloc(self._h, nbytes, byref(p)), "b2g_malloc")
    return int(p.value)


def _ctx_free(self, dev: int) -> None:
    _check(lib().b2g_free(self._h, c_void_p(dev)), "b2g_free")


def _ctx_memset_zero(self, dev: int, nbytes: int) -> None:
    _check(lib().b2g_memset_zero(self._h, c_void_p(dev), nbytes), "b2g_memset_zero")


def _ctx_download(self, host_ptrs, dev_ptrs, doubles) -> None:
    hp, dp, nd = _ptrs(host_ptrs), _ptrs(dev_ptrs), np.ascontiguousarray(doubles, dtype=np.int64)
    _check(lib().b2g_download(self._h, len(hp), hp.ctypes.data, dp.ctypes.data, nd.ctypes.data), "b2g_download")


def _ctx_upload_blocks(self, dev_ptrs, host_ptrs, doubles) -> None:
    hp, dp, nd = _ptrs(host_ptrs), _ptrs(dev_ptrs), np.ascontiguousarray(doubles, dtype=np.int64)
    _check(lib().b2g_upload_blocks(self._h, len(hp), dp.ctypes.data, hp.ctypes.data, nd.ctypes.data),
           "b2g_upload_blocks")


def _ctx_host_register(self, arr: np.ndarray) -> None:
    _check(lib().b2g_host_register(self._h, c_void_p(arr.ctypes.data), arr.nbytes), "b2g_host_register")


def _ctx_host_unregister(self, arr: np.ndarray) -> None:
    _check(lib().b2g_host_unregister(self._h, c_void_p(arr.ctypes.data)), "b2g_host_unregister")


def _ctx_mem_info(self):
    f, t = c_int64(), c_int64()
    _check(lib().b2g_mem_info(self._h, byref(f), byref(t)), "b2g_mem_info")
    return f.value, t.value


Context.resident_map = _ctx_resident_map
Context.resident_stats = _ctx_resident_stats
Context.malloc = _ctx_malloc
Context.free = _ctx_free
Context.memset_zero = _ctx_memset_zero
Context.download = _ctx_download
Context.upload_blocks = _ctx_upload_blocks
Context.host_register = _ctx_host_register
Context.host_unregister = _ctx_host_unregister
Context.mem_info = _ctx_mem_info


def upload_slices(length: int, nt: int):
    """Test hook: the byte ranges the staging threads of the pageable upload path copy for one chunk."""
    lo, hi = np.zeros(nt, dtype=np.int64), np.zeros(nt, dtype=np.int64)
    _check(lib().b2g_debug_upload_slices(length, nt, lo.ctypes.data, hi.ctypes.data), "b2g_debug_upload_slices")
    return lo, hi


class SeqPlan:
    """Device form of one recorded H.C pair list (BatchGEMMSeq after precompute())."""

    def __init__(self, ctx: Context, batch0: dict, batch1: dict, max_work: int, csize: int, vsize: int,
                 operand_space: int = OPERANDS_HOST, keepalive=None):
        """batch0/batch1: dicts with the BatchGEMM<double> arrays
        ta tb m n k lda ldb ldc alpha beta a b c (a/b/c as integer addresses)."""
        self.ctx = ctx
        self._keep = keepalive
        self._h = c_void_p()
        structs, keep = _make_batches(batch0, batch1)
        _check(lib().b2g_plan_create(ctx._h, byref(structs[0]), byref(structs[1]), max_work, csize, vsize,
                                     operand_space, byref(self._h)), "b2g_plan_create")
        self.csize, self.vsize = csize, vsize

    @classmethod
    def from_seqfile(cls, ctx: Context, sf: SeqFile, operands, operand_space: int = OPERANDS_HOST) -> "SeqPlan":
        """Build from a .b2seq pair list. `operands`: base address (int) of the concatenated
        operator arenas, host (numpy array accepted) or device according to operand_space."""
        keep = operands
        base = operands.ctypes.data if isinstance(operands, np.ndarray) else int(operands)
        b0, b1 = sf.as_batches(base)
        return cls(ctx, b0, b1, sf.max_work, sf.csize, sf.vsize, operand_space, keepalive=keep)

    def close(self) -> None:
        if self._h:
            lib().b2g_plan_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stats(self) -> PlanStats:
        st = PlanStats()
        _check(lib().b2g_plan_get_stats(self._h, byref(st)), "b2g_plan_get_stats")
        return st

    def __call__(self, c: np.ndarray, v: np.ndarray, scale: float = 1.0) -> None:
        """v += scale * H.c with host buffers (BatchGEMMSeq::operator())."""
        assert c.dtype == np.float64 and v.dtype == np.float64 and c.flags.c_contiguous and v.flags.c_contiguous
        assert c.size == self.csize and v.size == self.vsize
        _check(lib().b2g_seq_matvec(self._h, c.ctypes.data, v.ctypes.data, scale), "b2g_seq_matvec")

    def matvec_dev(self, c_ptr: int, v_ptr: int, scale: float = 1.0) -> None:
        """Device-resident c and sigma; asynchronous on the context stream."""
        _check(lib().b2g_seq_matvec_dev(self._h, c_void_p(c_ptr), c_void_p(v_ptr), scale), "b2g_seq_matvec_dev")

    def profile(self, c_ptr: int, v_ptr: int, scale: float = 1.0):
        """One matvec with CUDA events between the launches: [(name, flops, ms, units), ...]."""
        buf = (KernelStat * 64)()
        n = c_int()
        _check(lib().b2g_plan_profile(self._h, c_void_p(c_ptr), c_void_p(v_ptr), scale, buf, 64, byref(n)),
               "b2g_plan_profile")
        return [(buf[i].name.decode(), buf[i].flops, buf[i].ms, buf[i].units) for i in range(n.value)]

    def davidson(self, diag: np.ndarray, ket: np.ndarray, conv_thrd: float = 5e-6, rel_conv_thrd: float = 0.0,
                 max_iter: int = 5000, soft_max_iter: int = -1, deflation_min_size: int = 2,
                 deflation_max_size: int = 50):
        """Lowest eigenpair; ket is the initial guess and is overwritten. Returns (energy, ndav)."""
        diag, e, nd = _f64(diag), c_double(), c_int()
        assert ket.dtype == np.float64 and ket.flags.c_contiguous and ket.size == self.csize
        _check(lib().b2g_davidson(self._h, diag.ctypes.data, ket.ctypes.data, conv_thrd, rel_conv_thrd, max_iter,
                                  soft_max_iter, deflation_min_size, deflation_max_size, byref(e), byref(nd)),
               "b2g_davidson")
        return e.value, nd.value
