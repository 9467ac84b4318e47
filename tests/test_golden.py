"""Golden vectors (tests/golden/*.npz, made by oracle/make_golden.py): pin the oracle on CPU and
the CUDA path on the GPU against regressions."""
import glob
import os

import numpy as np
import pytest

import problems as P
from oracle.make_golden import CASES, ORACLE_ONLY, run_case

GOLD = [p for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
        if os.path.basename(p)[:-4] in CASES]      # (weno5_shuosher_sub.npz is reference data, see test_oracle_1d.py)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(path):
    name = os.path.basename(path)[:-4]
    g = np.load(path)
    out = run_case(*CASES[name])
    for key in g.files:
        if g[key].dtype == bool:
            assert np.array_equal(out[key], g[key]), key
        else:
            assert rel(out[key], g[key]) < 1e-13, key


GPU_GOLD = [p for p in GOLD if os.path.basename(p)[:-4] not in ORACLE_ONLY]


@pytest.mark.gpu
@pytest.mark.parametrize("path", GPU_GOLD, ids=[os.path.basename(p)[:-4] for p in GPU_GOLD])
def test_gpu_reproduces_golden(path):
    from p2de_b200 import TimeParam
    from p2de_b200.api import State, rhs
    from p2de_b200.types import Solver
    name = os.path.basename(path)[:-4]
    g = np.load(path)
    factory, nsteps = CASES[name]
    param, rd, md, dd, bc, U0 = P.setup(factory())
    assert np.array_equal(U0, g["U0"])
    solver = Solver(param=param, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc)
    st.set_state(U0)
    tp = param.timestepping_param
    dt1 = rhs(st, solver, None, TimeParam(t=tp.t0, dt=tp.CFL * tp.dt0, nstage=1))
    tol = 1e-11 if "gauss" in name else 1e-12     # entropy-variable round trip (pow / exp / log), test_gpu_gauss.py
    assert abs(dt1 - float(g["dt_stage1"])) <= 10 * tol * dt1
    assert rel(st.preallocation.rhsU, g["rhsU_stage1"]) < tol
    if "L_local_stage1" in g.files:
        sig = g["L_local_sig_stage1"] if "L_local_sig_stage1" in g.files else 1.0
        assert np.abs((st.preallocation.L_local[0] - g["L_local_stage1"]) * sig).max() < 10 * tol
    else:
        assert np.abs(st.preallocation.L[0] - g["L_stage1"]).max() < 1e-12
    st.set_state(U0)
    t = tp.t0
    for i in range(len(g["dthist"])):
        dt = st.ssp33_step(t)
        t += dt
        assert abs(dt - g["dthist"][i]) <= 1e-10 * dt
    assert rel(st.preallocation.Uq, g["U_final"]) < 1e-8


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_dense_restatement_reproduces_golden(path):
    """The committed vectors against the SECOND restatement (tests/dense_rhs.py, written independently of the oracle that generated
    them): what the GPU tests are held to is what two differently written readings of the reference agree on."""
    from dense_rhs import dense_limited_rhs, dense_ssp33_step, s_modified
    name = os.path.basename(path)[:-4]
    g = np.load(path)
    factory, nsteps = CASES[name]
    param, rd, md, dd, bc, U0 = P.setup(factory())
    assert np.array_equal(U0, g["U0"])
    tp = param.timestepping_param
    smin = float(s_modified(param.equation.gamma, U0).min())
    d = dense_limited_rhs(param, dd, bc, U0, tp.t0, tp.CFL * tp.dt0, 1, smin=smin)
    assert rel(d["rhsU"], g["rhsU_stage1"]) < 1e-12
    assert abs(d["dt"] - float(g["dt_stage1"])) <= 1e-13 * d["dt"]
    if "L_stage1" in g.files:
        assert np.abs(d["L"] - g["L_stage1"]).max() < 1e-12
    else:
        K, n = dd.sizes.K, param.N + 1
        Lg = g["L_local_stage1"].reshape(K, 2, n * (n + 1))
        assert np.abs(d["Lx"] - Lg[:, 0].reshape(K, n, n + 1)).max() < 1e-12
        assert np.abs(d["Ly"] - Lg[:, 1].reshape(K, n + 1, n)).max() < 1e-12
    U, t = U0, tp.t0
    for i in range(len(g["dthist"])):
        U, dt = dense_ssp33_step(param, dd, bc, U, t, smin=smin)
        assert abs(dt - g["dthist"][i]) <= 1e-12 * dt
        t += dt
    assert rel(U, g["U_final"]) < 1e-11
