"""Row 8f-1 on the ORACLE first: 2D Gauss collocation with NodewiseScaledExtrapolation
(src/dg/filter.jl, rhs.jl:113-133, flux_differencing.jl:288-328, low_order_graph_viscosity.jl:249-327),
the configuration of every shipped 2D example (e.g. examples/2D/kelvin-helmholtz.jl:44-55).  The
reference pins no numbers for it (and the path is broken at HEAD, oracle deviation D4), so the
restatement is held to the invariants of the scheme.  GPU parity for the same row: tests/test_gpu_gauss.py."""
import numpy as np
import pytest

import problems as P
from oracle.oracle import Oracle
from p2de_b200 import (ESLimitedLowOrderPos, GaussCollocation, LaxFriedrichsOnProjectedVal,
                       NodewiseScaledExtrapolation, SubcellLimiter, ZhangShuLimiter, primitive_to_conservative)

RHS = ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal())
GAUSS = dict(basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation(), rhs=RHS)


def make(problem, threads=2):
    param, rd, md, dd, bc, U0 = P.setup(problem)
    orc = Oracle(param, dd, bc, threads=threads)
    orc.set_state(U0)
    return param, md, dd, orc, U0


def total(dd, U):
    wJ = dd.ops.wq[None, :] * dd.geom.Jq
    return (wJ[..., None] * U).sum(axis=(0, 1))


@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_free_stream_preserved_on_gauss_nodes(limiter):
    param, ic, bc = P.vortex(N=3, K=(4, 3), limiter=limiter, **GAUSS)
    const = lambda prm, x, y: primitive_to_conservative(prm.equation, (1.2 + 0 * x, 0.3 + 0 * x, -0.4 + 0 * x, 0.9 + 0 * x))
    param, md, dd, orc, U0 = make((param, const, bc))
    orc.rhs(0.0, 1e-2, 1)
    for f in ("rhsU", "rhsH", "rhsL"):
        assert np.abs(orc.field(f)).max() < 1e-11
    assert (orc.field("theta_local").reshape(3, -1)[0] == 1.0).all()


def test_kelvin_helmholtz_example_configuration_is_conservative_and_positive():
    """examples/2D/kelvin-helmholtz.jl:44-55: N=3 Gauss, Nodewise, Subcell(PositivityBound), periodic."""
    param, md, dd, orc, U0 = make(P.kelvin_helmholtz(N=3, K=(8, 8), limiter=SubcellLimiter(), **GAUSS))
    c0 = total(dd, U0)
    t = 0.0
    for _ in range(6):
        t += orc.ssp33_step(t)
    U = orc.get_state()
    assert np.abs(total(dd, U) - c0).max() < 1e-12 * np.abs(c0).max()
    rho, rhoe = U[..., 0], U[..., 3] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) / U[..., 0]
    assert rho.min() > 0 and rhoe.min() > 0
    th = orc.field("theta_local")
    assert th.min() >= 0.0 and th.max() == 1.0


def test_theta_is_active_on_an_underresolved_vortex_and_one_on_a_resolved_one():
    """The bounds of filter.jl:84-98 are relative to the unlimited extrapolation: the strong test vortex
    (near-vacuum core: rho varies 14x across one 64x64 element) trips them at 8x8 and not at 100x100."""
    param, md, dd, orc, U0 = make(P.vortex(N=3, K=(8, 8), **GAUSS))
    orc.rhs(0.0, 5e-3, 1)
    th = orc.field("theta_local").reshape(3, -1)[0]
    assert (th < 1.0).any() and th.min() >= 0.0
    param, md, dd, orc, U0 = make(P.vortex(N=3, K=(100, 100), **GAUSS), threads=4)
    orc.rhs(0.0, 5e-3, 1)
    assert (orc.field("theta_local").reshape(3, -1)[0] == 1.0).all()


def test_vortex_converges_under_refinement():
    errs = []
    for K1 in (16, 32):
        param, md, dd, orc, U0 = make(P.vortex(N=2, K=(K1, K1), T=0.1, CFL=0.5, **GAUSS), threads=4)
        t = 0.0
        while t < 0.1 - 1e-14:
            t += orc.ssp33_step(t)
        ex = np.stack(primitive_to_conservative(param.equation, P.vortex_exact(param.equation, md.xq, md.yq, t)), axis=-1)
        wJ = dd.ops.wq[None, :] * dd.geom.Jq
        errs.append(float(np.sqrt((wJ[..., None] * (orc.get_state() - ex) ** 2).sum())))
    assert np.log2(errs[0] / errs[1]) > 2.0, errs
