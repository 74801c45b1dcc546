#!/usr/bin/env python
"""BASELINE.json config 5: synthetic random-integral 60-orbital FCIDUMP (SURVEY.md 8d recipe):
NORB=60, NELEC=60, MS2=0, ORBSYM all 1, t_ij and v_ijkl ~ N(0,1) * exp(-|i-j|/4) (v: both index pairs), symmetrised
(t symmetric, v 8-fold), numpy seed 0.  Writes the standard FCIDUMP text format."""
import sys

import numpy as np

n = 60
out = sys.argv[1] if len(sys.argv) > 1 else "workloads/_gen/config5/RANDOM60.FCIDUMP"
rng = np.random.default_rng(0)
idx = np.arange(n)
decay = np.exp(-np.abs(idx[:, None] - idx[None, :]) / 4.0)
t = rng.standard_normal((n, n)) * decay
t = 0.5 * (t + t.T)
v = rng.standard_normal((n, n, n, n)) * decay[:, :, None, None] * decay[None, None, :, :]
# 8-fold symmetry of real (ij|kl): i<->j, k<->l, (ij)<->(kl)
v = v + v.transpose(1, 0, 2, 3)
v = v + v.transpose(0, 1, 3, 2)
v = v + v.transpose(2, 3, 0, 1)
v *= 0.125
with open(out, "w") as f:
    f.write(" &FCI NORB=%d,NELEC=%d,MS2=0,\n  ORBSYM=%s\n  ISYM=1,\n &END\n" % (n, n, ",".join(["1"] * n) + ","))
    for i in range(n):
        for j in range(i + 1):
            for k in range(n):
                for l in range(k + 1):
                    if i * (i + 1) // 2 + j >= k * (k + 1) // 2 + l and abs(v[i, j, k, l]) > 1e-10:
                        f.write("%20.16E %4d %4d %4d %4d\n" % (v[i, j, k, l], i + 1, j + 1, k + 1, l + 1))
    for i in range(n):
        for j in range(i + 1):
            if abs(t[i, j]) > 1e-10:
                f.write("%20.16E %4d %4d %4d %4d\n" % (t[i, j], i + 1, j + 1, 0, 0))
    f.write("%20.16E %4d %4d %4d %4d\n" % (0.0, 0, 0, 0, 0))
print(out)
