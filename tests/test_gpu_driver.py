"""GPU tests of the time-loop surface around the hot path: SSP33! as one library call with the
DataHistory snapshots kept on the device (SSPRK33.jl:28-55), calculate_error reduced on the device (postprocess.jl:1-46),
and the NaN behaviour of the CFL dt (Julia's `min` propagates NaN, `while t < T` then ends)."""
import math

import numpy as np
import pytest

import problems as P
from p2de_b200 import ZhangShuLimiter
from p2de_b200 import types as T
from test_gpu_parity import make_pair, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("problem", [P.vortex(N=3, K=(8, 8), CFL=0.5, T=0.06, dt0=5e-3), P.sod(N=3, K=40, T=0.02),
                                     P.dmr(N=3, K=(64, 8), T=2e-3)], ids=["vortex", "sod-1d", "dmr-wide"])
def test_ssp33_loop_with_snapshot_ring(problem):
    from p2de_b200.api import SSP33
    import dataclasses
    param0 = problem[0]
    param0 = dataclasses.replace(param0, postprocessing_param=dataclasses.replace(param0.postprocessing_param, output_interval=3))
    param, solver, st, orc, U0 = make_pair((param0, problem[1], problem[2]), keep_diagnostics=False)
    tp = param.timestepping_param
    # the oracle's loop with the reference's push rule (SSPRK33.jl:41-55)
    t, i, th, Uh, dth = tp.t0, 1, [], [], []
    while t < tp.T:
        dt = orc.ssp33_step(t)
        t += dt; i += 1; dth.append(dt)
        if i % 3 == 0 or abs(t - tp.T) < 1e-10:
            th.append(t); Uh.append(orc.get_state())
    hist = SSP33(st, solver, None)
    assert len(hist.dthist) == len(dth) and np.allclose(hist.dthist, dth, rtol=1e-9, atol=0)
    assert len(hist.thist) == len(th) and np.allclose(hist.thist, th, rtol=1e-10, atol=0)
    assert abs(hist.thist[-1] - tp.T) < 1e-10
    for Ug, Uo in zip(hist.Uhist, Uh):
        assert rel(Ug, Uo) < 1e-7
    assert rel(st.preallocation.Uq, orc.get_state()) < 1e-7
    st.close()


def test_snapshot_ring_keeps_the_newest():
    param, solver, st, orc, U0 = make_pair(P.vortex(N=2, K=(6, 6), CFL=0.5, T=10.0, dt0=5e-3), keep_diagnostics=False)
    st.snapshot_ring(2, 2)                       # every 2nd step, room for two
    t, dth = st.ssp33_run(0.0, max_steps=9)      # step counter i = 2..10 -> pushes at i = 2, 4, 6, 8, 10
    assert len(dth) == 9
    snaps = st.snapshots()
    assert [s[1] for s in snaps] == [8, 10]
    assert abs(snaps[-1][0] - t) < 1e-14 and np.array_equal(snaps[-1][2], st.preallocation.Uq)
    st.close()


@pytest.mark.parametrize("problem", [P.vortex(N=3, K=(8, 8), CFL=0.5, T=0.03), P.sod(N=3, K=40, T=0.01)], ids=["vortex", "sod-1d"])
def test_calculate_error_on_device(problem):
    from p2de_b200.api import calculate_error
    from p2de_b200 import primitive_to_conservative
    param, solver, st, orc, U0 = make_pair(problem, keep_diagnostics=False)
    t = param.timestepping_param.t0
    for _ in range(3):
        t += st.ssp33_step(t)
    if param.equation.dim == 2:
        exact = lambda eq, x, y, tt: P.vortex_exact(eq, x, y, tt)
    else:
        exact = lambda eq, x, tt: (1.0 + 0.1 * np.sin(x), 0.2 + 0 * x, 1.0 + 0 * x)
    err = calculate_error(st, solver, exact)
    # the reference's loop in numpy
    md, dd = solver.md, solver.discrete_data
    Uq = st.preallocation.Uq
    args = (md.xq,) if md.yq is None else (md.xq, md.yq)
    ex = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), md.xq.shape) for c in
                   primitive_to_conservative(param.equation, exact(param.equation, *args, param.timestepping_param.T))], axis=-1)
    wJ = (dd.ops.wq * dd.geom.Jq.reshape(-1)[0])[None, :, None]
    diff = np.abs(ex - Uq)
    L1 = L2 = Linf = 0.0
    for c in range(param.equation.Nc):
        if np.abs(ex[..., c]).max() > 1e-14:
            L1 += (wJ[..., 0] * diff[..., c]).sum() / (wJ[..., 0] * np.abs(ex[..., c])).sum()
            L2 += math.sqrt((wJ[..., 0] * diff[..., c] ** 2).sum()) / math.sqrt((wJ[..., 0] * ex[..., c] ** 2).sum())
            Linf += diff[..., c].max() / np.abs(ex[..., c]).max()
    assert abs(err.L1err - L1) <= 1e-12 * L1 and abs(err.L2err - L2) <= 1e-12 * L2 and abs(err.Linferr - Linf) <= 1e-14 * Linf
    st.close()


@pytest.mark.parametrize("problem", [P.vortex(N=3, K=(16, 4)), P.vortex(N=3, K=(5, 5), limiter=ZhangShuLimiter()), P.sod(N=3, K=20)],
                         ids=["subcell", "zhangshu", "sod-1d"])
def test_nan_state_gives_nan_dt_like_the_reference(problem):
    """A NaN wavespeed makes the CFL candidate NaN; Julia's `min` propagates it (low_order_graph_viscosity.jl:230-242),
    dt becomes NaN and the time loop ends.  The oracle does the same."""
    param, solver, st, orc, U0 = make_pair(problem, keep_diagnostics=False)
    U = U0.copy()
    U[3, 1, 0] = float("nan")
    st.set_state(U); orc.set_state(U)
    assert math.isnan(orc.ssp33_step(0.0))
    assert math.isnan(st.ssp33_step(0.0))
    t, dth = st.ssp33_run(0.0, max_steps=5)
    assert len(dth) == 1 and math.isnan(t)        # `while t < T` is false for NaN
    st.close()
