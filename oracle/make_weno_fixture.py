"""Subsamples the reference's WENO5 Shu-Osher data (data/weno5_shuosher.mat: x, rho, u, p at
t = 1.8, 25002 points; only ever plotted by examples/1D/shu-osher.jl:70-74) into a small fixture.
Needs /root/reference (this container only); the fixture it writes is committed.

    python oracle/make_weno_fixture.py
"""
import os

import numpy as np
from scipy.io import loadmat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = loadmat("/root/reference/data/weno5_shuosher.mat")
x, rho, u, p = (np.asarray(m[k], dtype=np.float64).reshape(-1) for k in ("x", "rho", "u", "p"))
idx = np.linspace(0, len(x) - 1, 1251).round().astype(int)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "weno5_shuosher_sub.npz"), x=x[idx], rho=rho[idx], u=u[idx], p=p[idx])
print("wrote", len(idx), "points; x in", x[0], x[-1])
