"""Host-side mirror of src/dg/init.jl: textbook properties of the operators it builds."""
import numpy as np
import pytest

import problems as P
from p2de_b200 import (GaussCollocation, LobattoCollocation, gauss_lobatto_quad, gauss_quad, initialize_data,
                       make_periodic, structured_mapP)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 6])
def test_quadrature_rules(N):
    for rule, exact in ((gauss_lobatto_quad, 2 * N - 1), (gauss_quad, 2 * N + 1)):
        x, w = rule(N)
        assert len(x) == N + 1 and np.all(np.diff(x) > 0)
        for p in range(exact + 1):
            assert abs((w * x ** p).sum() - (2.0 / (p + 1) if p % 2 == 0 else 0.0)) < 1e-14
    assert gauss_lobatto_quad(N)[0][0] == -1.0 and gauss_lobatto_quad(N)[0][-1] == 1.0


@pytest.mark.parametrize("basis", [LobattoCollocation(), GaussCollocation()], ids=["lgl", "gauss"])
@pytest.mark.parametrize("N", [1, 2, 3, 4])
def test_hybridized_sbp_operators(N, basis):
    """init.jl:148-155: Srsh_db = [Q - Q^T, E^T B; -B E, 0]; skew; Q + Q^T = E^T B E; sizes."""
    param, _, _ = P.vortex(N=N, K=(2, 2), basis=basis)
    rd, md, dd = initialize_data(param)
    sz, ops = dd.sizes, dd.ops
    n = N + 1
    assert (sz.Nq, sz.Nfp, sz.Nh) == (n * n, 4 * n, n * n + 4 * n)
    for d in range(2):
        S = ops.Srsh_db[d]
        assert np.abs(S + S.T).max() < 1e-15
        assert np.abs(S[sz.Nq:, sz.Nq:]).max() == 0
        Q = np.diag(ops.wq) @ rd.Drst[d]
        B = np.diag(ops.Brs[d])
        assert np.abs(Q + Q.T - ops.Vf.T @ B @ ops.Vf).max() < 1e-13      # SBP property
        assert np.abs(S[:sz.Nq, :sz.Nq] - (Q - Q.T)).max() < 1e-14
        assert np.abs(S[:sz.Nq, sz.Nq:] - ops.Vf.T @ B).max() < 1e-15
        assert np.abs(rd.Drst[d].sum(1)).max() < 1e-13                    # derivative of a constant
        S0 = ops.Srs0[d]
        assert np.abs(S0 + S0.T).max() == 0 and np.count_nonzero(S0) == 2 * n * (n - 1)
    pairs = len(ops.Srsh_nnz)
    assert pairs == {True: n * n * (n - 1) + 4 * n * n, False: n * n * (n - 1) + 4 * n}[basis.code == 1]
    assert len(ops.Srs0_nnz) == 2 * n * (n - 1)
    # face maps: left, right, bottom, top; corners belong to two faces, vertical face first
    assert list(ops.fq2q[:n]) == [1 + j * n for j in range(n)]
    assert list(ops.fq2q[3 * n:]) == [1 + i + (n - 1) * n for i in range(n)]
    assert ops.q2fq[0] == [1, 2 * n + 1]


def test_mapP_is_an_involution_and_periodic_wrap():
    param, _, _ = P.vortex(N=2, K=(4, 3))
    rd, md, dd = initialize_data(param)
    Nfp = dd.sizes.Nfp
    assert ((md.mapP == md.mapM).sum()) == 2 * (4 + 3) * 3              # boundary face nodes map to themselves
    mp = make_periodic(md).mapP
    flat = mp.reshape(-1) - 1
    assert np.array_equal(flat[flat], np.arange(flat.size))             # P(P(x)) = x
    assert not (mp == md.mapM).any()
    # matching face nodes coincide geometrically (modulo the period)
    xf, yf = md.xf.reshape(-1), md.yf.reshape(-1)
    assert np.abs(((xf[flat] - xf) + 5.0) % 10.0 - 5.0).max() < 1e-12
    assert np.abs(((yf[flat] - yf) + 5.0) % 10.0 - 5.0).max() < 1e-12
    assert np.array_equal(structured_mapP(2, 3, 4, 3, (True, True)), mp)
