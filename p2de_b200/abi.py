"""ctypes mirror of include/p2de_b200.h and the packing of the reference's structs into it.

This is what the Julia glue does with `Ref{p2de_config}` + `pointer(A)` under `GC.@preserve`
(julia/P2DEB200.jl); here the same bytes are produced from the numpy mirror in init.py.
Matrices are math-indexed `A[i, j]` on the Python side and handed over column-major, i.e.
exactly Julia's memory for `A[i, j]`.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from .types import BCData, Discretization, Param

ABI_VERSION = 1
c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dim", C.c_int32), ("N", C.c_int32), ("basis", C.c_int32),
        ("K", C.c_int64), ("Kx", C.c_int32), ("Ky", C.c_int32),
        ("Nq", C.c_int32), ("Nfp", C.c_int32), ("Nh", C.c_int32), ("Np", C.c_int32),
        ("rhs_type", C.c_int32), ("vol_flux", C.c_int32), ("surf_flux_low", C.c_int32),
        ("surf_flux_high", C.c_int32), ("proj_limiter", C.c_int32), ("limiter", C.c_int32),
        ("bound", C.c_int32), ("shockcapture", C.c_int32), ("keep_diagnostics", C.c_int32),
        ("device", C.c_int32), ("lgl_projection_roundtrip", C.c_int32), ("_reserved", C.c_int32),
        ("hennemann_a", C.c_double), ("hennemann_c", C.c_double), ("bound_beta", C.c_double),
        ("gamma", C.c_double), ("POSTOL", C.c_double), ("ZEROTOL", C.c_double),
        ("zeta", C.c_double), ("eta", C.c_double),
        ("CFL", C.c_double), ("dt0", C.c_double), ("t0", C.c_double), ("T", C.c_double),
    ]


class OperatorsC(C.Structure):
    _fields_ = [
        ("Srsh_db", c_double_p * 2), ("Srs0", c_double_p * 2), ("Brs", c_double_p * 2),
        ("Vf", c_double_p), ("Vf_low", c_double_p), ("MinvVhT", c_double_p), ("MinvVfT", c_double_p),
        ("VDM_inv", c_double_p), ("wq", c_double_p), ("fq2q", c_int64_p),
    ]


class GeometryC(C.Structure):
    _fields_ = [
        ("J", c_double_p), ("Jq", c_double_p), ("GJh", c_double_p * 4),
        ("uniform", C.c_int32), ("_pad", C.c_int32),
        ("J_const", C.c_double), ("GJ_const", C.c_double * 4),
    ]


class BCDataC(C.Structure):
    _fields_ = [
        ("mapP", c_int64_p), ("periodic_x", C.c_int32), ("periodic_y", C.c_int32),
        ("nI", C.c_int64), ("mapI", c_int64_p), ("Ival", c_double_p),
        ("nO", C.c_int64), ("mapO", c_int64_p),
    ]


def _dp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_double_p) if a is not None else c_double_p()


def _ip(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_int64_p) if a is not None else c_int64_p()


def _colmajor(A: np.ndarray) -> np.ndarray:
    """math-indexed A[i, j] -> flat buffer in Julia (column-major) order."""
    return np.ascontiguousarray(np.asarray(A, dtype=np.float64).T).reshape(-1)


class PackedProblem:
    """Owns every buffer the four ABI structs point into (keeps them alive)."""

    def __init__(self, param: Param, dd: Discretization, bc: Optional[BCData], *, Kx_Ky=None,
                 structured_bc=None, uniform_geometry=True, keep_diagnostics=False, device=-1,
                 lgl_projection_roundtrip=False):
        sz, ops, geom = dd.sizes, dd.ops, dd.geom
        eq = param.equation
        self.keep: Dict[str, Any] = {}
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.dim, cfg.N, cfg.basis = eq.dim, param.N, param.approximation_basis.code
        cfg.K = sz.K
        if Kx_Ky is None:
            Kx_Ky = (int(param.K), 1) if eq.dim == 1 else (int(param.K[0]), int(param.K[1]))
        cfg.Kx, cfg.Ky = Kx_Ky
        cfg.Nq, cfg.Nfp, cfg.Nh, cfg.Np = sz.Nq, sz.Nfp, sz.Nh, sz.Np
        rhs = param.rhs
        cfg.rhs_type = rhs.code
        from . import types as T
        if rhs.code == T.RHS_LOW_ORDER_POSITIVITY:
            cfg.vol_flux, cfg.surf_flux_low, cfg.surf_flux_high = 0, rhs.surface_flux.code, T.SURFFLUX_LF_PROJECTED
        elif rhs.code == T.RHS_FLUX_DIFF:
            cfg.vol_flux, cfg.surf_flux_low, cfg.surf_flux_high = rhs.volume_flux.code, T.SURFFLUX_LF_NODAL, rhs.surface_flux.code
        else:
            cfg.vol_flux = rhs.high_order_volume_flux.code
            cfg.surf_flux_low = rhs.low_order_surface_flux.code
            cfg.surf_flux_high = rhs.high_order_surface_flux.code
        cfg.proj_limiter = param.entropyproj_limiter.code
        lim = param.rhs_limiter
        cfg.limiter = lim.code
        cfg.bound = lim.bound.code if lim.code != T.LIMITER_NONE else 0
        cfg.bound_beta = float(getattr(getattr(lim, "bound", None), "beta", 0.0))
        sc = getattr(lim, "shockcapture", T.NoShockCapture())
        cfg.shockcapture = sc.code
        cfg.hennemann_a, cfg.hennemann_c = float(getattr(sc, "a", 0.5)), float(getattr(sc, "c", 1.8))
        cfg.keep_diagnostics = int(keep_diagnostics)
        cfg.device = device
        cfg.lgl_projection_roundtrip = int(lgl_projection_roundtrip)
        cfg.gamma = eq.gamma
        cfg.POSTOL, cfg.ZEROTOL = param.global_constants.POSTOL, param.global_constants.ZEROTOL
        cfg.zeta, cfg.eta = param.limiting_param.zeta, param.limiting_param.eta
        tp = param.timestepping_param
        cfg.CFL, cfg.dt0, cfg.t0, cfg.T = tp.CFL, tp.dt0, tp.t0, tp.T
        self.cfg = cfg

        o = OperatorsC()
        k = self.keep
        for d in range(eq.dim):
            k[f"Srsh{d}"] = _colmajor(ops.Srsh_db[d]); o.Srsh_db[d] = _dp(k[f"Srsh{d}"])
            k[f"Srs0{d}"] = _colmajor(ops.Srs0[d]); o.Srs0[d] = _dp(k[f"Srs0{d}"])
            k[f"Brs{d}"] = np.ascontiguousarray(ops.Brs[d], dtype=np.float64); o.Brs[d] = _dp(k[f"Brs{d}"])
        for name in ("Vf", "Vf_low", "MinvVhT", "MinvVfT", "VDM_inv"):
            k[name] = _colmajor(getattr(ops, name)); setattr(o, name, _dp(k[name]))
        k["wq"] = np.ascontiguousarray(ops.wq, dtype=np.float64); o.wq = _dp(k["wq"])
        k["fq2q"] = np.ascontiguousarray(ops.fq2q, dtype=np.int64); o.fq2q = _ip(k["fq2q"])
        self.ops = o

        g = GeometryC()
        if uniform_geometry:
            g.uniform = 1
            g.J_const = float(geom.Jq.reshape(-1)[0])
            for a in range(len(geom.GJh)):
                g.GJ_const[a] = float(np.asarray(geom.GJh[a]).reshape(-1)[0])
        else:
            g.uniform = 0
            k["J"] = np.ascontiguousarray(geom.J, dtype=np.float64); g.J = _dp(k["J"])
            k["Jq"] = np.ascontiguousarray(geom.Jq, dtype=np.float64); g.Jq = _dp(k["Jq"])
            for a in range(len(geom.GJh)):
                k[f"GJh{a}"] = np.ascontiguousarray(geom.GJh[a], dtype=np.float64); g.GJh[a] = _dp(k[f"GJh{a}"])
        self.geom = g

        b = BCDataC()
        if structured_bc is not None:           # (periodic_x, periodic_y): no mapP array at all
            b.periodic_x, b.periodic_y = int(structured_bc[0]), int(structured_bc[1])
        else:
            k["mapP"] = np.ascontiguousarray(bc.mapP, dtype=np.int64); b.mapP = _ip(k["mapP"])
        if bc is not None:
            k["mapI"], k["mapO"], k["Ival"] = bc.mapI, bc.mapO, np.ascontiguousarray(bc.Ival, dtype=np.float64)
            b.nI, b.nO = len(bc.mapI), len(bc.mapO)
            b.mapI, b.mapO, b.Ival = _ip(bc.mapI), _ip(bc.mapO), _dp(k["Ival"])
        self.bc = b
