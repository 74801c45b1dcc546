"""Reader for `.b2tp` workload files: the blocking step (TensorFunctions::left_contract /
right_contract, block2 src/core/tensor_functions.hpp:2842-2885, 2941-2984) of one site as a list of
b2g_tp_term descriptors (include/b2g.h) with arena-relative operands.  Shapes only — no operator
values; format only — no arithmetic lives here.

Little endian: 8-byte magic b"B2TP\\0\\0\\0\\1"; u64[8] (nterms, n_in, n_out, is_right, call,
nflop, 2 reserved); i32[nterms] x 7 (am an bm bn cn conja conjb); f64[nterms] scale;
i64[nterms] x 6 (a_arena a_off b_arena b_off c_arena c_off); u64[n_in], u64[n_out] arena sizes
in doubles.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_I32 = ["am", "an", "bm", "bn", "cn", "conja", "conjb"]
_I64 = ["a_arena", "a_off", "b_arena", "b_off", "c_arena", "c_off"]


@dataclass
class TPFile:
    nterms: int
    is_right: bool
    nflop: int
    in_sizes: np.ndarray
    out_sizes: np.ndarray
    t: dict = field(default_factory=dict)

    def offsets(self):
        """Element offsets of (a, b, c) with the arenas laid out back to back (16-byte aligned starts)."""
        def starts(sz):
            st = np.zeros(len(sz) + 1, dtype=np.int64)
            np.cumsum((sz + 1) // 2 * 2, out=st[1:])
            return st
        si, so = starts(self.in_sizes), starts(self.out_sizes)
        T = self.t
        return (si[T["a_arena"]] + T["a_off"], si[T["b_arena"]] + T["b_off"], so[T["c_arena"]] + T["c_off"],
                int(si[-1]), int(so[-1]))

    def window_shapes(self):
        T = self.t
        rows = np.where(T["conja"] != 0, T["an"], T["am"]).astype(np.int64) * np.where(T["conjb"] != 0, T["bn"], T["bm"])
        cols = np.where(T["conja"] != 0, T["am"], T["an"]).astype(np.int64) * np.where(T["conjb"] != 0, T["bm"], T["bn"])
        return rows, cols


def load_tpfile(path: str) -> TPFile:
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as f:
            raw = np.frombuffer(f.read(), dtype=np.uint8)
    else:
        raw = np.fromfile(path, dtype=np.uint8)
    if bytes(raw[:8]) != b"B2TP\0\0\0\1":
        raise ValueError(f"{path}: not a .b2tp file")
    pos = 8

    def take(dtype, count):
        nonlocal pos
        out = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += np.dtype(dtype).itemsize * count
        return out

    hdr = take(np.uint64, 8)
    nt, n_in, n_out = int(hdr[0]), int(hdr[1]), int(hdr[2])
    tp = TPFile(nterms=nt, is_right=bool(hdr[3]), nflop=int(hdr[5]), in_sizes=np.zeros(0, np.int64),
                out_sizes=np.zeros(0, np.int64))
    for name in _I32:
        tp.t[name] = take(np.int32, nt).copy()
    tp.t["scale"] = take(np.float64, nt).copy()
    for name in _I64:
        tp.t[name] = take(np.int64, nt).copy()
    tp.in_sizes = take(np.uint64, n_in).astype(np.int64)
    tp.out_sizes = take(np.uint64, n_out).astype(np.int64)
    if pos != raw.size:
        raise ValueError(f"{path}: {raw.size - pos} trailing bytes")
    return tp
