"""Generates tests/golden/*.npz with the CPU oracle (oracle/p2de_oracle.cpp).

The reference's own tests pin no numbers (test/test_smoke.jl only checks that nothing throws)
and Julia is not installed here, so these vectors pin the ORACLE (and through it the CUDA path)
against regressions; they are not outputs of the Julia code ("parity unpinned").

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from p2de_b200 import (ESLimitedLowOrderPos, GaussCollocation, LaxFriedrichsOnProjectedVal,  # noqa: E402
                       NodewiseScaledExtrapolation, PositivityAndRelaxedCellEntropyBound, SubcellLimiter,
                       TVDAndCellEntropyBound, ZhangShuLimiter)

GAUSS = dict(basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation(),
             rhs=ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal()))

CASES = {
    # test/test_smoke.jl:44-67: vortex, K=(5,5), T=2e-2, CFL=1, dt0=1e-2 -> 2 steps
    **{f"smoke_vortex_N{N}_{name}": (lambda N=N, lim=lim: P.vortex(N=N, K=(5, 5), limiter=lim), 2)
       for N in (1, 2, 3, 4) for name, lim in (("subcell", SubcellLimiter()), ("zhangshu", ZhangShuLimiter()))},
    "dmr_N3_subcell": (lambda: P.dmr(N=3, K=(16, 4)), 5),
    "sedov_N2_zhangshu": (lambda: P.sedov(N=2, K=(8, 8), limiter=ZhangShuLimiter()), 5),
    # row 8f-1: the configuration of examples/2D/kelvin-helmholtz.jl:44-55
    "kh_N3_gauss_nodewise_subcell": (lambda: P.kelvin_helmholtz(N=3, K=(6, 6), **GAUSS), 3),
    # row 8f-2: TVD and cell-entropy bounds on plateau-free data (tests/test_gpu_bounds.py explains why not the vortex)
    "wave_N3_tvd_cellentropy": (lambda: P.wave2d(N=3, limiter=SubcellLimiter(bound=TVDAndCellEntropyBound())), 3),
    "wave_N2_relaxed_cellentropy": (lambda: P.wave2d(N=2, limiter=SubcellLimiter(bound=PositivityAndRelaxedCellEntropyBound(beta=0.5))), 3),
}
ORACLE_ONLY = set()     # cases without a GPU kernel (none at present)


def run_case(factory, nsteps):
    param, rd, md, dd, bc, U0 = P.setup(factory())
    orc = Oracle(param, dd, bc, threads=1)
    orc.set_state(U0)
    tp = param.timestepping_param
    dt1 = orc.rhs(tp.t0, tp.CFL * tp.dt0, 1)
    out = {"U0": U0, "rhsU_stage1": orc.field("rhsU"), "dt_stage1": np.array(dt1)}
    if param.rhs_limiter.code == 2:
        out["L_local_stage1"] = orc.field("L_local")[0]
        if param.rhs_limiter.bound.code >= 5 and param.equation.dim == 2:
            # TVD bounds: coefficients of faces with f_bar_H - f_bar_L = rounding noise are 0 or 1 by the sign of that
            # noise (tests/test_gpu_bounds.py: significant_faces); the mask of the faces that are compared
            n = param.N + 1
            sig = []
            for ax in "xy":
                fH = orc.field(f"f_bar_H_{ax}").reshape(-1, n * n + n, 4)
                fL = orc.field(f"f_bar_L_{ax}").reshape(-1, n * n + n, 4)
                sig.append(np.abs(fH - fL).max(-1) > 1e-10 * np.abs(fH).max())
            out["L_local_sig_stage1"] = np.stack(sig, axis=1)
    else:
        out["L_stage1"] = orc.field("L")[0]
    orc.set_state(U0)
    t, dth = tp.t0, []
    for _ in range(nsteps):
        if t >= tp.T:
            break
        dt = orc.ssp33_step(t)
        t += dt
        dth.append(dt)
    out["U_final"] = orc.get_state()
    out["dthist"] = np.array(dth)
    return out


if __name__ == "__main__":
    outdir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, (factory, nsteps) in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **run_case(factory, nsteps))
        print("wrote", name)
