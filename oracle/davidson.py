"""oracle/davidson.py — TEST INFRASTRUCTURE, NOT PRODUCT.

numpy restatement of the reference's ground-state Davidson
(IterativeMatrixFunctions<double>::davidson with k = 1, DavidsonTypes::Normal,
/root/reference/src/core/iterative_matrix_functions.hpp:864-1173; Olsen
preconditioner :93-108).  Pinned against the reference's own eigenvalue and
iteration count stored in tests/golden/*.b2seq (tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np


def olsen_precondition(q: np.ndarray, c: np.ndarray, ld: float, aa: np.ndarray) -> None:
    """q = Kinv q - (c, Kinv q) / (c, Kinv c) Kinv c   (:93-108)."""
    t = c.copy()
    mask = np.abs(ld - aa) > 1e-12
    t[mask] /= (ld - aa[mask])
    q[mask] /= (ld - aa[mask])
    q += t * (-(c @ q) / (c @ t))


def davidson(op, aa: np.ndarray, v0: np.ndarray, conv_thrd: float = 5e-6, rel_conv_thrd: float = 0.0,
             max_iter: int = 5000, soft_max_iter: int = -1, deflation_min_size: int = 2,
             deflation_max_size: int = 50):
    """Returns (eigenvalue, ndav, eigenvector). `op(x)` returns H.x as a new array."""
    n, k = v0.size, 1
    deflation_min_size = max(deflation_min_size, k)
    deflation_max_size = max(deflation_max_size, k + k // 2)
    bs = np.zeros((deflation_max_size, n))
    sigmas = np.zeros((deflation_max_size, n))
    bs[0] = v0 / np.linalg.norm(v0)                      # :945-955
    m, msig, xiter, ck = k, 0, 0, 0
    ld = np.zeros(1)
    while xiter < max_iter and (soft_max_iter == -1 or xiter < soft_max_iter):
        xiter += 1
        while msig < m:                                  # :971-977
            sigmas[msig] = op(bs[msig])
            msig += 1
        alpha = np.zeros((m, m))
        for i in range(m):                               # :999-1000, lower triangle only
            for j in range(i + 1):
                alpha[i, j] = bs[i] @ sigmas[j]
        sym = alpha + alpha.T - np.diag(np.diag(alpha))
        ld, vec = np.linalg.eigh(sym)                    # dsyev; eigenvector j -> row j (:1004)
        rot = vec.T
        sigmas[:m] = rot @ sigmas[:m]                    # :1005-1026
        bs[:m] = rot @ bs[:m]
        q = sigmas[0] - ld[0] * bs[0]                    # :1072-1073 (ck = 0, Normal ordering)
        qq = q @ q
        olsen_precondition(q, bs[0], ld[0], aa)
        if abs(qq) < conv_thrd + abs(ld[0]) ** 2 * rel_conv_thrd ** 2 and m >= k:   # :1098-1100
            ck += 1
            break
        if m >= deflation_max_size:                      # :1104-1107
            m = msig = deflation_min_size
        for j in range(m):                               # :1139-1146 (sequential, q updated in place)
            q -= bs[j] * (bs[j] @ q)
        q /= np.linalg.norm(q)
        bs[m] = q
        m += 1
        if xiter == soft_max_iter:
            break
    return float(ld[0]), xiter, bs[0].copy()
