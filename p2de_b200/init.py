"""Host-side mirror of the reference's one-time setup, `src/dg/init.jl`.

The reference builds its reference element and mesh through StartUpDG/NodesAndModes
(not vendored, versions unpinned: Project.toml:25-41).  Only Line and Quad elements with
LGL or Gauss collocation on uniform Cartesian meshes are ever built (init.jl:66-103,137),
and all of those are textbook tensor-product objects, constructed here directly:

    initialize_reference_data   init.jl:66-131  -> reference_element_1d / build_operators
    initialize_operators        init.jl:133-228 -> build_operators
    low_order_operators         init.jl:276-340 -> _low_order_1d / build_operators
    initialize_uniform_mesh_data, geometric_factors  init.jl:230-274 -> uniform_mesh
    init_U!                     init.jl:342-361 -> sample_initial_condition

Orderings (fixed by the reference's own index arithmetic): volume node
`i + j*N1D` (r fastest); faces left, right, bottom, top with N1D nodes each ordered by the
free coordinate; elements x-fastest; `mapP` 1-based linear index into [Nfp, K].

All arrays here are numpy, C-ordered with the ELEMENT index first, i.e. `A[k, i]` is the
reference's `A[i, k]`; the raw memory is identical to Julia's column-major `[i, k]`.
Operator matrices are math-indexed `M[i, j]` and transposed to column-major at the ABI.
"""
from __future__ import annotations

import math
from typing import Callable, Tuple

import numpy as np
from numpy.polynomial import legendre as _leg

from .types import (BASIS_GAUSS, BCData, Discretization, GeomData, MeshData, Operators, Param,
                    SizeData, Solver, num_elements)


# ----------------------------------------------------------------------------- 1D rules
def gauss_quad(N: int) -> Tuple[np.ndarray, np.ndarray]:
    """gauss_quad(0, 0, N): N+1 Gauss-Legendre nodes and weights (init.jl:68,86,303)."""
    x, w = _leg.leggauss(N + 1)
    x = 0.5 * (x - x[::-1])          # enforce exact symmetry
    w = 0.5 * (w + w[::-1])
    return x, w


def gauss_lobatto_quad(N: int) -> Tuple[np.ndarray, np.ndarray]:
    """gauss_lobatto_quad(0, 0, N): N+1 Legendre-Gauss-Lobatto nodes/weights (init.jl:76,308)."""
    if N < 1:
        raise ValueError("LGL needs N >= 1")
    cN = np.zeros(N + 1)
    cN[N] = 1.0
    if N == 1:
        x = np.array([-1.0, 1.0])
    else:
        d1 = _leg.legder(cN)
        d2 = _leg.legder(d1)
        xi = np.sort(_leg.legroots(d1).real)
        for _ in range(4):            # Newton polish of the roots of P_N'
            xi = xi - _leg.legval(xi, d1) / _leg.legval(xi, d2)
        x = np.concatenate([[-1.0], xi, [1.0]])
        x = 0.5 * (x - x[::-1])
    w = 2.0 / (N * (N + 1) * _leg.legval(x, cN) ** 2)
    w = 0.5 * (w + w[::-1])
    return x, w


def _bary_weights(x: np.ndarray) -> np.ndarray:
    n = len(x)
    b = np.ones(n)
    for j in range(n):
        for m in range(n):
            if m != j:
                b[j] /= (x[j] - x[m])
    return b


def lagrange_diff_matrix(x: np.ndarray) -> np.ndarray:
    """D[i, j] = l_j'(x_i) for the Lagrange basis on the nodes x."""
    n = len(x)
    b = _bary_weights(x)
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[i, j] = (b[j] / b[i]) / (x[i] - x[j])
        D[i, i] = -np.sum(D[i, np.arange(n) != i])
    return D


def lagrange_eval(x: np.ndarray, xe: float) -> np.ndarray:
    """[l_j(xe)]_j ; exact unit vector when xe coincides with a node."""
    n = len(x)
    out = np.zeros(n)
    hit = np.where(x == xe)[0]
    if len(hit):
        out[hit[0]] = 1.0
        return out
    for j in range(n):
        v = 1.0
        for m in range(n):
            if m != j:
                v *= (xe - x[m]) / (x[j] - x[m])
        out[j] = v
    return out


def orthonormal_legendre(x: np.ndarray, N: int) -> np.ndarray:
    """V[i, n] = P_n(x_i) * sqrt((2n+1)/2)  (jacobiP(x, 0, 0, n))."""
    V = np.zeros((len(x), N + 1))
    for n in range(N + 1):
        c = np.zeros(n + 1)
        c[n] = 1.0
        V[:, n] = _leg.legval(x, c) * math.sqrt((2 * n + 1) / 2.0)
    return V


def _droptol(A: np.ndarray, tol: float) -> np.ndarray:
    """droptol!(sparse(A), tol): entries with |a| <= tol become structural zeros."""
    A = np.array(A, dtype=np.float64, copy=True)
    A[np.abs(A) <= tol] = 0.0
    return A


def _low_order_1d(N: int):
    """construct_low_order_operators_1D (init.jl:276-295)."""
    n = N + 1
    Q = np.zeros((n, n))
    for i in range(n - 1):
        Q[i, i + 1] = 0.5
        Q[i + 1, i] = -0.5
    Q[0, 0] = -0.5
    Q[n - 1, n - 1] = 0.5
    S = 0.5 * (Q - Q.T)
    Vf_low = np.zeros((2, n))
    Vf_low[0, 0] = 1.0
    Vf_low[1, n - 1] = 1.0
    return Q, S, Vf_low


class RefElemData:
    """The pieces of StartUpDG.RefElemData that init.jl:141 destructures."""

    def __init__(self, dim, N, basis):
        self.dim, self.N, self.basis = dim, N, basis
        gauss = basis == BASIS_GAUSS
        r1, w1 = gauss_quad(N) if gauss else gauss_lobatto_quad(N)
        self.r1D, self.w1D = r1, w1
        n = N + 1
        D1 = lagrange_diff_matrix(r1)
        vl, vr = lagrange_eval(r1, -1.0), lagrange_eval(r1, 1.0)
        V1 = orthonormal_legendre(r1, N)
        if dim == 1:
            self.rq, self.wq = r1.copy(), w1.copy()
            self.Drst = (D1,)
            self.Vf = np.stack([vl, vr])
            self.wf = np.ones(2)
            self.nrstJ = (np.array([-1.0, 1.0]),)
            self.VDM = V1
            self.rf = np.array([-1.0, 1.0])
        else:
            I = np.eye(n)
            self.rq = np.tile(r1, n)            # r fastest
            self.sq = np.repeat(r1, n)
            self.wq = np.kron(w1, w1)
            self.Drst = (np.kron(I, D1), np.kron(D1, I))
            Vf = np.zeros((4 * n, n * n))
            for a in range(n):                  # a = free-coordinate index on the face
                for m in range(n):
                    Vf[a, m + a * n] = vl[m]              # left   (r=-1), node j=a
                    Vf[n + a, m + a * n] = vr[m]          # right  (r=+1)
                    Vf[2 * n + a, a + m * n] = vl[m]      # bottom (s=-1), node i=a
                    Vf[3 * n + a, a + m * n] = vr[m]      # top    (s=+1)
            self.Vf = Vf
            self.wf = np.tile(w1, 4)
            z, o = np.zeros(n), np.ones(n)
            self.nrstJ = (np.concatenate([-o, o, z, z]), np.concatenate([z, z, -o, o]))
            self.rf = np.concatenate([-o, o, r1, r1])
            self.sf = np.concatenate([r1, r1, -o, o])
            self.VDM = np.kron(V1, V1)           # mode a + b*(N+1): deg a in r, b in s
        self.M = np.diag(self.wq)
        self.Vq = np.eye(len(self.wq))
        self.Pq = np.eye(len(self.wq))


# ----------------------------------------------------------------------------- mesh
def uniform_mesh(param: Param, rd: RefElemData) -> MeshData:
    """initialize_uniform_mesh_data + MeshData (init.jl:230-252), non-periodic mapP."""
    dim = rd.dim
    n = param.N + 1
    if dim == 1:
        Kx, Ky = int(param.K), 1
        xL, xR = float(param.xL), float(param.xR)
        hx = (xR - xL) / Kx
        K = Kx
        ix = np.arange(K)
        xq = xL + hx * (ix[:, None] + 0.5 * (rd.rq[None, :] + 1.0))
        xf = xL + hx * (ix[:, None] + 0.5 * (rd.rf[None, :] + 1.0))
        yq = yf = None
        Nfp = 2
        J = np.full((K, n), hx / 2.0)
        rxJ, sxJ, ryJ, syJ = 1.0, 0.0, 0.0, 0.0
    else:
        Kx, Ky = int(param.K[0]), int(param.K[1])
        hx = (param.xR[0] - param.xL[0]) / Kx
        hy = (param.xR[1] - param.xL[1]) / Ky
        K = Kx * Ky
        k = np.arange(K)
        ix, iy = k % Kx, k // Kx
        xq = param.xL[0] + hx * (ix[:, None] + 0.5 * (rd.rq[None, :] + 1.0))
        yq = param.xL[1] + hy * (iy[:, None] + 0.5 * (rd.sq[None, :] + 1.0))
        xf = param.xL[0] + hx * (ix[:, None] + 0.5 * (rd.rf[None, :] + 1.0))
        yf = param.xL[1] + hy * (iy[:, None] + 0.5 * (rd.sf[None, :] + 1.0))
        Nfp = 4 * n
        J = np.full((K, n * n), hx * hy / 4.0)
        rxJ, sxJ, ryJ, syJ = hy / 2.0, 0.0, 0.0, hx / 2.0
    mapM = (np.arange(K)[:, None] * Nfp + np.arange(Nfp)[None, :] + 1).astype(np.int64)
    mapP = structured_mapP(dim, n, Kx, Ky, (False, False))
    mapB = mapM[mapP == mapM]
    return MeshData(K=K, xq=xq, yq=yq, xf=xf, yf=yf, mapM=mapM, mapP=mapP, mapB=mapB, J=J,
                    rxJ=rxJ, sxJ=sxJ, ryJ=ryJ, syJ=syJ, Kxy=(Kx, Ky), is_periodic=(False, False))


def structured_mapP(dim, n, Kx, Ky, periodic) -> np.ndarray:
    """mapP[k, f] (1-based into [Nfp, K]) of the uniform mesh; self on non-periodic boundaries."""
    if dim == 1:
        K, Nfp = Kx, 2
        k = np.arange(K)
        mapP = np.zeros((K, Nfp), dtype=np.int64)
        kl, kr = k - 1, k + 1
        self_l, self_r = kl < 0, kr >= K
        if periodic[0]:
            kl, kr = kl % K, kr % K
            self_l[:] = False
            self_r[:] = False
        mapP[:, 0] = np.where(self_l, k * Nfp + 0, kl * Nfp + 1) + 1
        mapP[:, 1] = np.where(self_r, k * Nfp + 1, kr * Nfp + 0) + 1
        return mapP
    K, Nfp = Kx * Ky, 4 * n
    k = np.arange(K)
    ix, iy = k % Kx, k // Kx
    mapP = np.zeros((K, Nfp), dtype=np.int64)
    a = np.arange(n)[None, :]
    # (face offset of mine, face offset of neighbour, dix, diy, periodic flag, extent)
    for fo, fno, dix, diy in ((0, n, -1, 0), (n, 0, 1, 0), (2 * n, 3 * n, 0, -1), (3 * n, 2 * n, 0, 1)):
        jx, jy = ix + dix, iy + diy
        out = (jx < 0) | (jx >= Kx) | (jy < 0) | (jy >= Ky)
        wrap = periodic[0] if dix != 0 else periodic[1]
        if wrap:
            jx, jy = jx % Kx, jy % Ky
            out = np.zeros_like(out)
        kn = np.where(out, k, jx + jy * Kx)
        fn = np.where(out[:, None], fo + a, fno + a)
        mapP[:, fo:fo + n] = kn[:, None] * Nfp + fn + 1
    return mapP


def make_periodic(md: MeshData, periodic=None) -> MeshData:
    """StartUpDG.make_periodic(md): wrap mapP in every direction (test/test_smoke.jl:27)."""
    import dataclasses
    dim = 1 if md.yq is None else 2
    if periodic is None:
        periodic = (True, True)
    n = md.xq.shape[1] if dim == 1 else int(round(math.sqrt(md.xq.shape[1])))
    Kx, Ky = md.Kxy
    mapP = structured_mapP(dim, n, Kx, Ky, periodic)
    return dataclasses.replace(md, mapP=mapP, mapB=md.mapM[mapP == md.mapM], is_periodic=tuple(periodic))


# ----------------------------------------------------------------------------- operators
def build_operators(param: Param, rd: RefElemData, md: MeshData) -> Discretization:
    """initialize_operators (init.jl:133-228) + low_order_operators (init.jl:297-340)."""
    ZEROTOL = param.global_constants.ZEROTOL
    N, dim = param.N, rd.dim
    n = N + 1
    wq, wf, M, Pq, Vq, Vf = rd.wq, rd.wf, rd.M, rd.Pq, rd.Vq, rd.Vf
    Nq, Nfp = len(wq), Vf.shape[0]
    Nh = Nq + Nfp
    Vh = np.vstack([Vq, Vf])
    Qrs = tuple(_droptol(Pq.T @ M @ D @ Pq, ZEROTOL) for D in rd.Drst)
    Ef = Vf @ Pq
    Brs_full = tuple(_droptol(np.diag(wf * nJ), ZEROTOL) for nJ in rd.nrstJ)
    Srsh_db = []
    for Q, B in zip(Qrs, Brs_full):
        Qh = _droptol(0.5 * np.block([[Q - Q.T, Ef.T @ B], [-B @ Ef, B]]), ZEROTOL)
        Srsh_db.append(2.0 * (0.5 * (Qh - Qh.T)))
    Minv = 1.0 / np.diag(M)
    Vh_d = _droptol(Vh, ZEROTOL)
    MinvVhT = Minv[:, None] * Vh_d.T
    MinvVfT = Minv[:, None] * Vf.T

    # low-order operators
    Q01D, S01D, Vf_low1 = _low_order_1d(N)
    if dim == 1:
        Srs0 = (_droptol(S01D, ZEROTOL),)
        Vf_low = Vf_low1
    else:
        M1D = np.diag(rd.w1D)
        Qr0 = _droptol(np.kron(M1D, Q01D), ZEROTOL)
        Qs0 = _droptol(np.kron(Q01D, M1D), ZEROTOL)
        Srs0 = (_droptol(0.5 * (Qr0 - Qr0.T), ZEROTOL), _droptol(0.5 * (Qs0 - Qs0.T), ZEROTOL))
        if rd.basis == BASIS_GAUSS:     # low_order_extrapolation(::GaussQuadrature), init.jl:324-336
            Js = np.concatenate([np.arange(0, (n - 1) * n + 1, n), np.arange(n - 1, n * n, n),
                                 np.arange(0, n), np.arange((n - 1) * n, n * n)])
            Vf_low = np.zeros((Nfp, Nq))
            Vf_low[np.arange(Nfp), Js] = 1.0
        else:
            Vf_low = _droptol(Vf, ZEROTOL)
    Vf_low = Vf_low.copy()
    Vf_low[np.abs(Vf_low - 1.0) < ZEROTOL] = 1.0            # init.jl:181-185

    Srsh_nnz = [(i + 1, j + 1) for j in range(Nh) for i in range(j + 1, Nh)
                if sum(abs(S[i, j]) for S in Srsh_db) != 0]
    Srs0_nnz = [(i + 1, j + 1) for j in range(Nq) for i in range(j + 1, Nq)
                if sum(abs(S[i, j]) for S in Srs0) != 0]
    fq2q = np.array([int(np.argmax(Vf_low[i] == 1.0)) + 1 for i in range(Nfp)], dtype=np.int64)
    q2fq = [[f + 1 for f in range(Nfp) if Vf_low[f, i] == 1.0] for i in range(Nq)]

    K = num_elements(param)
    Jq = md.J                                                # Jq = Vq * J with Vq = I (init.jl:259,270)
    GJ = (md.rxJ,) if dim == 1 else (md.rxJ, md.sxJ, md.ryJ, md.syJ)
    GJh = tuple(np.full((K, Nh), g) for g in GJ) if K * Nh <= (1 << 22) else tuple(
        np.broadcast_to(np.float64(g), (K, Nh)) for g in GJ)
    sizes = SizeData(K=K, N1D=n, Nd=dim, Nc=dim + 2, Np=rd.VDM.shape[1], Nq=Nq, Nfp=Nfp, Nh=Nh, Ns=3)
    geom = GeomData(J=md.J, Jq=Jq, GJh=GJh)
    ops = Operators(Srsh_db=tuple(Srsh_db), Srs0=Srs0, Srsh_nnz=Srsh_nnz, Srs0_nnz=Srs0_nnz,
                    Brs=tuple(np.diag(B).copy() for B in Brs_full), Vh=Vh, MinvVhT=MinvVhT,
                    VDM_inv=np.linalg.inv(rd.VDM), Vq=Vq, Vf=Vf, Vf_low=Vf_low, Pq=Pq,
                    MinvVfT=MinvVfT, wq=wq.copy(), q2fq=q2fq, fq2q=fq2q)
    return Discretization(sizes=sizes, geom=geom, ops=ops)


def light_mesh(param: Param, rd: RefElemData) -> MeshData:
    """Uniform-mesh metadata WITHOUT per-element arrays (J is a zero-memory broadcast view and the
    coordinate / map arrays are None): for meshes of millions of elements that are handed to the
    library in structured form (`p2de_bcdata.mapP == NULL`, `p2de_geometry.uniform`)."""
    dim, n = rd.dim, param.N + 1
    if dim == 1:
        Kx, Ky = int(param.K), 1
        hx = (float(param.xR) - float(param.xL)) / Kx
        J, GJ = hx / 2.0, (1.0, 0.0, 0.0, 0.0)
    else:
        Kx, Ky = int(param.K[0]), int(param.K[1])
        hx, hy = (param.xR[0] - param.xL[0]) / Kx, (param.xR[1] - param.xL[1]) / Ky
        J, GJ = hx * hy / 4.0, (hy / 2.0, 0.0, 0.0, hx / 2.0)
    K = Kx * Ky
    return MeshData(K=K, xq=None, yq=None, xf=None, yf=None, mapM=None, mapP=None, mapB=None,
                    J=np.broadcast_to(np.float64(J), (K, n ** dim)), rxJ=GJ[0], sxJ=GJ[1], ryJ=GJ[2], syJ=GJ[3],
                    Kxy=(Kx, Ky), is_periodic=(False, False))


def element_nodes(param: Param, rd: RefElemData, k: np.ndarray):
    """(xq, yq) [len(k), Nq] of the elements `k` of the uniform mesh (same arithmetic as uniform_mesh)."""
    Kx, Ky = int(param.K[0]), int(param.K[1])
    hx, hy = (param.xR[0] - param.xL[0]) / Kx, (param.xR[1] - param.xL[1]) / Ky
    ix, iy = k % Kx, k // Kx
    xq = param.xL[0] + hx * (ix[:, None] + 0.5 * (rd.rq[None, :] + 1.0))
    yq = param.xL[1] + hy * (iy[:, None] + 0.5 * (rd.sq[None, :] + 1.0))
    return xq, yq


def initialize_data(param: Param, light: bool = False):
    """initialize_data / initialize_reference_data (init.jl:62-103)."""
    rd = RefElemData(param.equation.dim, param.N, param.approximation_basis.code)
    md = light_mesh(param, rd) if light else uniform_mesh(param, rd)
    return rd, md, build_operators(param, rd, md)


def sample_initial_condition(param: Param, md: MeshData, initial_condition: Callable) -> np.ndarray:
    """init_U! (init.jl:342-361): Uq[k, i, :] = initial_condition(param, xq[, yq])."""
    Nc = param.equation.Nc
    K, Nq = md.xq.shape
    args = (md.xq,) if md.yq is None else (md.xq, md.yq)
    try:                                    # vectorised callback (arrays in, [..., Nc] or tuple out)
        out = initial_condition(param, *args)
        out = np.stack([np.broadcast_to(np.asarray(o, dtype=np.float64), md.xq.shape) for o in out], axis=-1) \
            if isinstance(out, (tuple, list)) else np.asarray(out, dtype=np.float64)
        if out.shape == (K, Nq, Nc):
            return np.ascontiguousarray(out)
    except Exception:
        pass
    Uq = np.zeros((K, Nq, Nc))
    for k in range(K):
        for i in range(Nq):
            Uq[k, i] = initial_condition(param, *(a[k, i] for a in args))
    return Uq


def primitive_to_conservative(equation, U):
    """primitive_to_conservative (src/math/compressible_Navier_Stokes.jl:1-16); array-friendly."""
    g = equation.gamma
    if equation.dim == 1:
        rho, u, p = U
        return (rho, rho * u, p / (g - 1) + 0.5 * rho * u ** 2)
    rho, u, v, p = U
    return (rho, rho * u, rho * v, p / (g - 1) + 0.5 * rho * (u ** 2 + v ** 2))
