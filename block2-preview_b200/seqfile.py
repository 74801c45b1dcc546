"""Reader for `.b2seq` workload files: the flat GEMM-pair list that
EffectiveHamiltonian::precompute() records (block2 src/dmrg/effective_hamiltonian.hpp:226),
serialised by the host driver.  Format only — no arithmetic lives here.

Little endian: 8-byte magic b"B2SEQ\\0\\0\\2"; 16 u64 (npairs, n_arenas, csize, vsize,
max_work, nflop_mnk, has_data, site, bond_dim, n_sites, ndav_ref, has_eigs, 4 reserved);
8 f64 (e_ref, const_e, t_ref_matvec, conv_thrd, 4 reserved); 16 i32[npairs]
(ta0 tb0 m0 n0 k0 lda0 ldb0 ldc0 ta1 tb1 m1 n1 k1 lda1 ldb1 ldc1); 4 f64[npairs]
(alpha0 beta0 alpha1 beta1); 7 i64[npairs] (a0_off b0_arena b0_off a1_arena a1_off c1_off
w_off); u64[n_arenas] arena sizes; optional data (arenas, c, v_ref, diag, ket0).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_I32 = ["ta0", "tb0", "m0", "n0", "k0", "lda0", "ldb0", "ldc0", "ta1", "tb1", "m1", "n1", "k1", "lda1", "ldb1", "ldc1"]
_F64 = ["alpha0", "beta0", "alpha1", "beta1"]
_I64 = ["a0_off", "b0_arena", "b0_off", "a1_arena", "a1_off", "c1_off", "w_off"]


@dataclass
class SeqFile:
    npairs: int
    csize: int
    vsize: int
    max_work: int
    nflop_mnk: int
    site: int
    bond_dim: int
    n_sites: int
    ndav_ref: int
    e_ref: float
    const_e: float
    conv_thrd: float
    arena_sizes: np.ndarray
    p: dict = field(default_factory=dict)
    arenas: np.ndarray | None = None
    c: np.ndarray | None = None
    v_ref: np.ndarray | None = None
    diag: np.ndarray | None = None
    ket0: np.ndarray | None = None

    @property
    def operand_doubles(self) -> int:
        return int(self.arena_sizes.sum())

    @property
    def flops(self) -> float:
        """Conventional 2*m*n*k FLOPs of one matvec."""
        return 2.0 * self.nflop_mnk

    def operand_offsets(self):
        st = np.zeros(len(self.arena_sizes) + 1, dtype=np.int64)
        np.cumsum(self.arena_sizes, out=st[1:])
        return st[self.p["b0_arena"]] + self.p["b0_off"], st[self.p["a1_arena"]] + self.p["a1_off"]

    def pair_flops(self) -> np.ndarray:
        """2*m*n*k FLOPs of every pair (both GEMMs)."""
        P = self.p
        return 2.0 * (P["m0"].astype(np.float64) * P["n0"] * P["k0"] + P["m1"].astype(np.float64) * P["n1"] * P["k1"])

    def subset(self, mask: np.ndarray) -> "SeqFile":
        """The pair list restricted to `mask`, with unreferenced operator arenas dropped
        (one rank's slice of the MPO terms, or a bounded CPU sample)."""
        P = {k: v[mask] for k, v in self.p.items()}
        used = np.union1d(P["b0_arena"], P["a1_arena"]).astype(np.int64)
        remap = np.full(len(self.arena_sizes), -1, dtype=np.int64)
        remap[used] = np.arange(len(used))
        P["b0_arena"], P["a1_arena"] = remap[P["b0_arena"]], remap[P["a1_arena"]]
        nf = int((P["m0"].astype(np.int64) * P["n0"] * P["k0"] + P["m1"].astype(np.int64) * P["n1"] * P["k1"]).sum())
        mw = int((P["m0"].astype(np.int64) * P["n0"]).max()) if len(P["m0"]) else 0
        out = SeqFile(npairs=int(mask.sum()), csize=self.csize, vsize=self.vsize, max_work=mw, nflop_mnk=nf,
                      site=self.site, bond_dim=self.bond_dim, n_sites=self.n_sites, ndav_ref=0, e_ref=0.0,
                      const_e=self.const_e, conv_thrd=self.conv_thrd, arena_sizes=self.arena_sizes[used], p=P)
        if self.arenas is not None:
            st = np.zeros(len(self.arena_sizes) + 1, dtype=np.int64)
            np.cumsum(self.arena_sizes, out=st[1:])
            out.arenas = np.concatenate([self.arenas[st[a]:st[a + 1]] for a in used]) if len(used) else np.zeros(0)
            out.c = self.c
        return out

    def shard(self, rank: int, nranks: int) -> "SeqFile":
        """One rank's slice of the H.C terms.  The reference splits the MPO terms by operator
        ownership (ParallelRuleQC, block2 src/dmrg/qc_parallel_rule.hpp:44-80: owner = index % size);
        on a recorded list the operator of a term is its left-block operand, so a pair goes to
        rank (a1_arena % nranks).  sum_ranks sigma_r == sigma."""
        return self.subset((self.p["a1_arena"] % nranks) == rank)

    def as_batches(self, operand_base: int):
        """The two BatchGEMM<double> array sets exactly as the reference holds them:
        wavefunction / work operands null-based, operator operands real addresses."""
        P = self.p
        b0o, a1o = self.operand_offsets()
        w = (8 * P["w_off"]).astype(np.uint64)
        b0 = dict(ta=np.where(P["ta0"] != 0, 112, 111), tb=np.where(P["tb0"] != 0, 112, 111), m=P["m0"], n=P["n0"],
                  k=P["k0"], lda=P["lda0"], ldb=P["ldb0"], ldc=P["ldc0"], alpha=P["alpha0"], beta=P["beta0"],
                  a=(8 * P["a0_off"]).astype(np.uint64), b=(operand_base + 8 * b0o).astype(np.uint64), c=w)
        b1 = dict(ta=np.where(P["ta1"] != 0, 112, 111), tb=np.where(P["tb1"] != 0, 112, 111), m=P["m1"], n=P["n1"],
                  k=P["k1"], lda=P["lda1"], ldb=P["ldb1"], ldc=P["ldc1"], alpha=P["alpha1"], beta=P["beta1"],
                  a=(operand_base + 8 * a1o).astype(np.uint64), b=w, c=(8 * P["c1_off"]).astype(np.uint64))
        return b0, b1


def load_seqfile(path: str, with_data: bool = True) -> SeqFile:
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as f:
            raw = np.frombuffer(f.read(), dtype=np.uint8)
    else:
        raw = np.memmap(path, dtype=np.uint8, mode="r")
    if bytes(raw[:8]) != b"B2SEQ\0\0\2":
        raise ValueError(f"{path}: not a .b2seq file")
    pos = 8

    def take(dtype, count):
        nonlocal pos
        nbytes = np.dtype(dtype).itemsize * count
        out = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += nbytes
        return out

    hdr = take("<u8", 16)
    dh = take("<f8", 8)
    n, na = int(hdr[0]), int(hdr[1])
    p = {}
    for nm in _I32:
        p[nm] = np.array(take("<i4", n))
    for nm in _F64:
        p[nm] = np.array(take("<f8", n))
    for nm in _I64:
        p[nm] = np.array(take("<i8", n))
    asz = np.array(take("<u8", na)).astype(np.int64)
    sf = SeqFile(npairs=n, csize=int(hdr[2]), vsize=int(hdr[3]), max_work=int(hdr[4]), nflop_mnk=int(hdr[5]),
                 site=int(hdr[7]), bond_dim=int(hdr[8]), n_sites=int(hdr[9]), ndav_ref=int(hdr[10]),
                 e_ref=float(dh[0]), const_e=float(dh[1]), conv_thrd=float(dh[3]), arena_sizes=asz, p=p)
    if int(hdr[6]) and with_data:
        sf.arenas = np.array(take("<f8", int(asz.sum())))
        sf.c = np.array(take("<f8", sf.csize))
        sf.v_ref = np.array(take("<f8", sf.vsize))
        sf.diag = np.array(take("<f8", sf.csize))
        sf.ket0 = np.array(take("<f8", sf.csize))
    return sf
