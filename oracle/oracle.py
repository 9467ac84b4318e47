"""ctypes wrapper of the CPU oracle (oracle/p2de_oracle.cpp).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from p2de_b200.abi import BCDataC, Config, GeometryC, OperatorsC, PackedProblem  # noqa: E402

_LIBS = {}
_DEFAULT = "ref"
_SO = {"ref": "libp2de_oracle.so", "native": "libp2de_oracle_native.so", "fma": "libp2de_oracle_fma.so",
       "series": "libp2de_oracle_series.so"}


def build(force: bool = False, variant: str = "ref") -> str:
    """Builds one variant of the oracle (oracle/Makefile):
    "ref"    portable -O2, -ffp-contract=off: THE checker (Julia does not contract a*b+c);
    "fma"    same source with -ffp-contract=fast -mfma: a legal re-association of the same formulas, used by tests to
             measure how far the reference formulation itself moves under rounding (tolerance probe);
    "series" the portable build with logmean's log branch evaluated by a cancellation-free series where it converges:
             second tolerance probe (how much of a difference is the rounding noise of -da / (log aL - log aR) itself);
    "native" -O3 -march=native, -ffp-contract=off: bench.py's CPU arm, compiled on the machine that runs it."""
    so = os.path.join(_HERE, _SO[variant])
    src = os.path.join(_HERE, "p2de_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "p2de_b200.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", _SO[variant]] + (["-B"] if force else []))
    return so


def use_native_build():
    """bench.py's CPU arm: make the -O3 -march=native build the default variant.  Falls back to the portable build if
    it cannot be compiled here; returns the flags actually in use."""
    global _DEFAULT
    try:
        build(variant="native")
        _DEFAULT = "native"
        return "-O3 -march=native -ffp-contract=off -fopenmp"
    except Exception:
        _DEFAULT = "ref"
        return "-O2 -ffp-contract=off -fopenmp"


def lib(variant: str = None):
    variant = variant or _DEFAULT
    if variant not in _LIBS:
        L = C.CDLL(build(variant=variant))
        L.oracle_phase_times.restype = C.c_int64
        L.oracle_phase_times.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32]
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(Config), C.POINTER(OperatorsC), C.POINTER(GeometryC), C.POINTER(BCDataC)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_rhs.restype = C.c_double
        L.oracle_rhs.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32]
        L.oracle_ssp33_step.restype = C.c_double
        L.oracle_ssp33_step.argtypes = [C.c_void_p, C.c_double]
        L.oracle_get_field.restype = C.c_int64
        L.oracle_get_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
        L.oracle_reduce.restype = C.c_double
        L.oracle_reduce.argtypes = [C.c_void_p, C.c_int32]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_set_threads.argtypes = [C.c_int32]
        L.oracle_max_threads.restype = C.c_int32
        L.oracle_logmean.restype = C.c_double
        L.oracle_logmean.argtypes = [C.c_double, C.c_double]
        for f in ("oracle_v_ufun_2d", "oracle_u_vfun_2d", "oracle_fluxes_2d"):
            getattr(L, f).argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_fS_2d.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_limiting_param_2d.restype = C.c_double
        L.oracle_limiting_param_2d.argtypes = [C.c_double, C.c_void_p, C.c_void_p] + [C.c_double] * 4
        _LIBS[variant] = L
    return _LIBS[variant]


class Oracle:
    """One reference-faithful CPU solver+state (mirrors State/Solver of the reference)."""

    def __init__(self, param, discrete_data, bcdata, *, structured_bc=None, threads=None, variant=None):
        self.L = lib(variant)
        self.sizes = discrete_data.sizes
        self.packed = PackedProblem(param, discrete_data, bcdata, structured_bc=structured_bc)
        if threads is not None:
            self.L.oracle_set_threads(int(threads))
        self.h = self.L.oracle_create(C.byref(self.packed.cfg), C.byref(self.packed.ops),
                                      C.byref(self.packed.geom), C.byref(self.packed.bc))
        if not self.h:
            raise RuntimeError("oracle_create: " + self.L.oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    def set_state(self, Uq):
        Uq = np.ascontiguousarray(Uq, dtype=np.float64)
        assert Uq.shape == (self.sizes.K, self.sizes.Nq, self.sizes.Nc)
        self.L.oracle_set_state(self.h, Uq.ctypes.data)

    def get_state(self):
        out = np.empty((self.sizes.K, self.sizes.Nq, self.sizes.Nc))
        self.L.oracle_get_state(self.h, out.ctypes.data)
        return out

    def rhs(self, t, dt, nstage):
        return self.L.oracle_rhs(self.h, t, dt, nstage)

    def ssp33_step(self, t):
        return self.L.oracle_ssp33_step(self.h, t)

    def reduce(self, what):
        return self.L.oracle_reduce(self.h, what)

    def phase_times(self, reset=False):
        """Seconds per phase under the reference's TimerOutputs labels, in first-use order."""
        n = self.L.oracle_phase_times(self.h, None, 0, 0)
        buf = C.create_string_buffer(int(n))
        self.L.oracle_phase_times(self.h, buf, n, int(reset))
        out = {}
        for line in buf.value.decode().splitlines():
            k, v = line.split("\t")
            out[k] = float(v)
        return out

    def field(self, name, shape=None):
        n = self.L.oracle_get_field(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n)
        self.L.oracle_get_field(self.h, name.encode(), out.ctypes.data, n)
        s = self.sizes
        default = {
            "Uq": (s.K, s.Nq, s.Nc), "rhsU": (s.K, s.Nq, s.Nc), "rhsH": (s.K, s.Nq, s.Nc), "rhsL": (s.K, s.Nq, s.Nc),
            "resW": (s.K, s.Nq, s.Nc), "vq": (s.K, s.Nq, s.Nc), "u_tilde": (s.K, s.Nh, s.Nc),
            "rhsxyH": (s.K, s.Nq, s.Nd, s.Nc), "rhsxyL": (s.K, s.Nq, s.Nd, s.Nc),
            "BF_H": (s.K, s.Nfp, s.Nd, s.Nc), "BF_L": (s.K, s.Nfp, s.Nd, s.Nc),
            "L": (s.Ns, s.K), "L_local": (s.Ns, s.K, s.Nd, s.Nq + s.N1D),
        }
        shape = shape or default.get(name)
        return out.reshape(shape) if shape else out


def run_ssp33(oracle: Oracle, t0: float, T: float, max_steps: int = 10 ** 9):
    """The `while t < T` loop of SSP33! (timestepping/SSPRK33.jl:28-44) on the oracle."""
    t, dthist = t0, []
    while t < T and len(dthist) < max_steps:
        dt = oracle.ssp33_step(t)
        t += dt
        dthist.append(dt)
    return t, dthist
