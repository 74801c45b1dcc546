"""Reader for `.b2blk` files: the blocking list one TensorFunctions::left_contract /
right_contract call records (block2 src/core/tensor_functions.hpp:2842-2885, 2941-2984) in the
cblas_dgemm_batch group form of BatchGEMM<double> (src/core/batch_gemm.hpp:237-247), written by
the reference-side recorder harness (`blkdump` mode).  Format only — no arithmetic lives here.

Little endian: 8-byte magic b"B2BLK\\0\\0\\1"; u64[8] (ngroups, nentries, n_in, n_out, nflop_mnk,
is_right, call, reserved); i32[ngroups] x 9 (ta tb m n k lda ldb ldc gp); f64[ngroups] x 2
(alpha beta); i64[nentries] x 6 (a_arena a_off b_arena b_off c_arena c_off); u64[n_in] input arena
sizes; u64[n_out] output arena sizes; f64 input arenas; f64 output arenas on entry; f64 output
arenas after the reference's own BatchGEMMSeq::auto_perform.  Optional second section, the same
call in term form (b2g_tp_term, include/b2g.h): 8-byte magic b"B2TERMS\1"; u64 nterms;
i32[nterms] x 7 (am an bm bn cn conja conjb); f64[nterms] scale; i64[nterms] x 6 (a_arena a_off
b_arena b_off c_arena c_off) over the same arenas.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_I32 = ["ta", "tb", "m", "n", "k", "lda", "ldb", "ldc", "gp"]
_F64 = ["alpha", "beta"]
_I64 = ["a_arena", "a_off", "b_arena", "b_off", "c_arena", "c_off"]


@dataclass
class BlkFile:
    ngroups: int
    nentries: int
    nflop_mnk: int
    is_right: bool
    in_sizes: np.ndarray
    out_sizes: np.ndarray
    g: dict = field(default_factory=dict)   # per-group arrays
    e: dict = field(default_factory=dict)   # per-entry arrays
    inputs: np.ndarray | None = None        # concatenated input arenas
    c_in: np.ndarray | None = None          # concatenated output arenas on entry
    c_ref: np.ndarray | None = None         # ... after the reference executor
    terms: dict | None = None               # term form of the same call (per-term arrays), if recorded

    def term_offsets(self):
        """Element offsets (a, b into the concatenated inputs; c into the concatenated outputs) of the terms."""
        si = np.zeros(len(self.in_sizes) + 1, dtype=np.int64)
        np.cumsum(self.in_sizes, out=si[1:])
        so = np.zeros(len(self.out_sizes) + 1, dtype=np.int64)
        np.cumsum(self.out_sizes, out=so[1:])
        T = self.terms
        return si[T["a_arena"]] + T["a_off"], si[T["b_arena"]] + T["b_off"], so[T["c_arena"]] + T["c_off"]

    def pointers(self, in_base: int, out_base: int):
        """Entry pointers (a, b, c) as integer addresses for arenas laid out back to back at
        in_base / out_base (bytes)."""
        si = np.zeros(len(self.in_sizes) + 1, dtype=np.int64)
        np.cumsum(self.in_sizes, out=si[1:])
        so = np.zeros(len(self.out_sizes) + 1, dtype=np.int64)
        np.cumsum(self.out_sizes, out=so[1:])
        a = (in_base + 8 * (si[self.e["a_arena"]] + self.e["a_off"])).astype(np.uint64)
        b = (in_base + 8 * (si[self.e["b_arena"]] + self.e["b_off"])).astype(np.uint64)
        c = (out_base + 8 * (so[self.e["c_arena"]] + self.e["c_off"])).astype(np.uint64)
        return a, b, c

    def group_args(self):
        """(ta, tb, m, n, k, alpha, lda, ldb, beta, ldc, group_size) with CBLAS transpose codes."""
        G = self.g
        return (np.where(G["ta"] != 0, 112, 111).astype(np.int32), np.where(G["tb"] != 0, 112, 111).astype(np.int32),
                G["m"], G["n"], G["k"], G["alpha"], G["lda"], G["ldb"], G["beta"], G["ldc"], G["gp"])


def load_blkfile(path: str) -> BlkFile:
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as f:
            raw = np.frombuffer(f.read(), dtype=np.uint8)
    else:
        raw = np.fromfile(path, dtype=np.uint8)
    if bytes(raw[:8]) != b"B2BLK\0\0\1":
        raise ValueError(f"{path}: not a .b2blk file")
    pos = 8

    def take(dtype, count):
        nonlocal pos
        out = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += np.dtype(dtype).itemsize * count
        return out

    hdr = take(np.uint64, 8)
    ng, ne, n_in, n_out = (int(x) for x in hdr[:4])
    bf = BlkFile(ngroups=ng, nentries=ne, nflop_mnk=int(hdr[4]), is_right=bool(hdr[5]),
                 in_sizes=np.zeros(0, np.int64), out_sizes=np.zeros(0, np.int64))
    for name in _I32:
        bf.g[name] = take(np.int32, ng).copy()
    for name in _F64:
        bf.g[name] = take(np.float64, ng).copy()
    for name in _I64:
        bf.e[name] = take(np.int64, ne).copy()
    bf.in_sizes = take(np.uint64, n_in).astype(np.int64)
    bf.out_sizes = take(np.uint64, n_out).astype(np.int64)
    tin, tout = int(bf.in_sizes.sum()), int(bf.out_sizes.sum())
    bf.inputs = take(np.float64, tin).copy()
    bf.c_in = take(np.float64, tout).copy()
    bf.c_ref = take(np.float64, tout).copy()
    if pos != raw.size:
        if bytes(raw[pos:pos + 8]) != b"B2TERMS\1":
            raise ValueError(f"{path}: {raw.size - pos} trailing bytes")
        pos += 8
        nt = int(take(np.uint64, 1)[0])
        bf.terms = {}
        for name in ("am", "an", "bm", "bn", "cn", "conja", "conjb"):
            bf.terms[name] = take(np.int32, nt).copy()
        bf.terms["scale"] = take(np.float64, nt).copy()
        for name in ("a_arena", "a_off", "b_arena", "b_off", "c_arena", "c_off"):
            bf.terms[name] = take(np.int64, nt).copy()
        if pos != raw.size:
            raise ValueError(f"{path}: {raw.size - pos} trailing bytes")
    return bf
