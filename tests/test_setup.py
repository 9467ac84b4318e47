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


# ---- orderings the reference hard-codes downstream of StartUpDG: if the setup restated in p2de_b200/init.py (which the
#      oracle AND the CUDA path consume) ordered nodes or faces differently, every GPU-vs-oracle test would still agree.
#      These pin the conventions one by one against the reference's own index arithmetic.
@pytest.mark.parametrize("basis", [LobattoCollocation(), GaussCollocation()], ids=["lgl", "gauss"])
@pytest.mark.parametrize("N", [1, 2, 3, 4])
def test_node_and_face_orderings_the_reference_hard_codes(N, basis):
    param, _, _ = P.vortex(N=N, K=(3, 2), basis=basis)
    rd, md, dd = initialize_data(param)
    n = N + 1
    r1 = rd.r1D
    # volume nodes: `reshape(view(rhsxyH, :, k), N1D, N1D)[si, sj]` with f_bar_x accumulated along si (subcell.jl:178-196):
    # the first index runs along x (r), i.e. r is the fastest-varying coordinate
    rq, sq = rd.rq.reshape(n, n), rd.sq.reshape(n, n)          # C order: [j, i]
    assert np.allclose(rq, np.tile(r1, (n, 1))) and np.allclose(sq, np.tile(r1[:, None], (1, n)))
    # face nodes (subcell_face_idx_to_quad_face_index_x / _y, limiter_utils.jl:97-123): faces 1..N1D are si == 1 (r = -1)
    # indexed by sj, N1D+1..2N1D are si == N1D+1 (r = +1), then sj == 1 (s = -1) indexed by si, then s = +1
    assert np.all(rd.rf[:n] == -1) and np.allclose(rd.sf[:n], r1)
    assert np.all(rd.rf[n:2 * n] == 1) and np.allclose(rd.sf[n:2 * n], r1)
    assert np.all(rd.sf[2 * n:3 * n] == -1) and np.allclose(rd.rf[2 * n:3 * n], r1)
    assert np.all(rd.sf[3 * n:] == 1) and np.allclose(rd.rf[3 * n:], r1)
    # apply_LF_dissipation_to_BF (rhs_utils.jl:84-91): faces i <= 2 N1D carry the x-component, the others the y-component,
    # i.e. Br is nonzero exactly on the first 2 N1D face nodes (negative on the left face), Bs on the last 2 N1D
    Br, Bs = dd.ops.Brs
    assert np.all(Br[:n] < 0) and np.all(Br[n:2 * n] > 0) and np.all(Br[2 * n:] == 0)
    assert np.all(Bs[:2 * n] == 0) and np.all(Bs[2 * n:3 * n] < 0) and np.all(Bs[3 * n:] > 0)
    # fq2q (init.jl:209-213): the volume node a face node belongs to: the line end in the reshape convention above
    fq2q = np.asarray(dd.ops.fq2q) - 1
    for a in range(n):
        assert fq2q[a] == 0 + a * n and fq2q[n + a] == (n - 1) + a * n
        assert fq2q[2 * n + a] == a and fq2q[3 * n + a] == a + (n - 1) * n
    # Vf extrapolates along the face node's own grid line only (a 0/1 row on Lobatto nodes)
    Vf = dd.ops.Vf
    for f in range(4 * n):
        line = [fq2q[f] % n + j * n for j in range(n)] if f >= 2 * n else [(fq2q[f] // n) * n + i for i in range(n)]
        off = np.setdiff1d(np.arange(n * n), line)
        assert np.abs(Vf[f, off]).max() < 1e-14 and abs(Vf[f].sum() - 1) < 1e-13
        if basis.code == 0:
            assert Vf[f, fq2q[f]] == 1.0 and np.count_nonzero(Vf[f]) == 1


@pytest.mark.parametrize("N", [1, 3])
def test_mapP_linear_index_and_partner_faces(N):
    """subcell_index_P_x / _y (limiter_utils.jl:125-181): mapP[iface, k] is a 1-based linear index into [Nfp, K]
    (iP = mod1(p, Nfp), kP = div(p - 1, Nfp) + 1); across an x-face the partner is on the opposite x-face at the same sj,
    across a y-face on the opposite y-face at the same si; element k = ix + Kx * iy (x fastest)."""
    param, _, _ = P.vortex(N=N, K=(4, 3))
    rd, md, dd = initialize_data(param)
    bc = make_periodic(md)
    n, Nfp = N + 1, 4 * (N + 1)
    Kx, Ky = 4, 3
    mapP = np.asarray(bc.mapP) - 1                  # [K, Nfp] 0-based linear index
    for k in range(Kx * Ky):
        ix, iy = k % Kx, k // Kx
        # element centres follow k = ix + Kx * iy
        assert abs(md.xq[k].mean() - (param.xL[0] + (ix + 0.5) * (param.xR[0] - param.xL[0]) / Kx)) < 1e-12
        assert abs(md.yq[k].mean() - (param.xL[1] + (iy + 0.5) * (param.xR[1] - param.xL[1]) / Ky)) < 1e-12
        for f in range(Nfp):
            kP, fP = mapP[k, f] // Nfp, mapP[k, f] % Nfp
            face, a = f // n, f % n
            assert fP % n == a and fP // n == face ^ 1           # opposite face of the pair, same position along it
            exp = {0: ((ix - 1) % Kx, iy), 1: ((ix + 1) % Kx, iy), 2: (ix, (iy - 1) % Ky), 3: (ix, (iy + 1) % Ky)}[face]
            assert kP == exp[0] + Kx * exp[1]
