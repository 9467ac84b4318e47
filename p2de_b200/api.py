"""The reference's operator surface for the hot path, backed by libp2de_b200.so.

    initialize_DG(param, ic, bc) -> (solver, state, state_param)     src/dg/init.jl:44-60
    rhs(state, solver, state_param, time_param) -> dt   (`rhs!`)      src/dg/rhs/rhs.jl:5-13
    apply_rhs_limiter(limiter, state, solver, state_param, time_param) (`apply_rhs_limiter!`)
                                                                      src/dg/limiter/limiter.jl:8-56
    SSP33(state, solver, state_param) -> DataHistory   (`SSP33!`)     src/timestepping/SSPRK33.jl:1-62
    check_conservation(state, solver)                                 src/dg/utils.jl:1-12
    calculate_error(state, solver, exact_sol)                         src/dg/postprocess.jl:1-46

Julia's `!` cannot be part of a Python name; argument order, meaning and return values are the
reference's.  `state.preallocation.<field>` reads the observable arrays of
src/common/types/State.jl:1-26 back from the device (numpy, element index first).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Optional

import numpy as np

from . import lib as _lib
from . import types as T
from .abi import PackedProblem
from .init import initialize_data, primitive_to_conservative, sample_initial_condition
from .types import (BCData, DataHistory, ErrorData, Param, Solver, StateParam, TimeParam)


class Preallocation:
    """Device-backed view of the reference's `Preallocation` (State.jl:1-26)."""

    def __init__(self, state: "State"):
        self._s = state

    def _field(self, code, shape):
        s = self._s
        out = np.empty(shape, dtype=np.float64)
        _lib.check(s.L, s.h, s.L.p2de_get_field(s.h, code, out.ctypes.data, out.size))
        return out

    @property
    def Uq(self):
        sz = self._s.sizes
        return self._field(T.FIELD_UQ, (sz.K, sz.Nq, sz.Nc))

    @Uq.setter
    def Uq(self, value):
        self._s.set_state(value)

    @property
    def resW(self):
        sz = self._s.sizes
        return self._field(T.FIELD_RESW, (sz.K, sz.Nq, sz.Nc))

    @property
    def rhsU(self):
        sz = self._s.sizes
        return self._field(T.FIELD_RHSU, (sz.K, sz.Nq, sz.Nc))

    @property
    def rhsH(self):
        sz = self._s.sizes
        return self._field(T.FIELD_RHSH, (sz.K, sz.Nq, sz.Nc))

    @property
    def rhsL(self):
        sz = self._s.sizes
        return self._field(T.FIELD_RHSL, (sz.K, sz.Nq, sz.Nc))

    @property
    def L(self):          # reference: L[k, nstage]  -> here [nstage, k]
        sz = self._s.sizes
        return self._field(T.FIELD_L, (sz.Ns, sz.K))

    @property
    def L_local(self):    # reference: L_local[idx, d, k, nstage] -> here [nstage, k, d, idx]
        sz = self._s.sizes
        return self._field(T.FIELD_L_LOCAL, (sz.Ns, sz.K, sz.Nd, sz.Nq + sz.N1D))   # 1D: entries > Nq+1 unused (= 1)

    @property
    def theta(self):
        sz = self._s.sizes
        return self._field(T.FIELD_THETA, (sz.Ns, sz.K))

    @property
    def theta_local(self):
        sz = self._s.sizes
        return self._field(T.FIELD_THETA_LOCAL, (sz.Ns, sz.K, sz.Nfp))


class State:
    """`State(preallocation, cache)` of the reference; the caches live on the device."""

    def __init__(self, solver: Solver, bcdata: BCData, *, keep_diagnostics=False, device=-1,
                 structured_bc=None, Kx_Ky=None, lgl_projection_roundtrip=False):
        self.L = _lib.load()
        self.sizes = solver.discrete_data.sizes
        self.packed = PackedProblem(solver.param, solver.discrete_data, bcdata, keep_diagnostics=keep_diagnostics,
                                    device=device, structured_bc=structured_bc, Kx_Ky=Kx_Ky,
                                    lgl_projection_roundtrip=lgl_projection_roundtrip)
        h = C.c_void_p()
        rc = self.L.p2de_create(C.byref(self.packed.cfg), C.byref(self.packed.ops), C.byref(self.packed.geom),
                                C.byref(self.packed.bc), C.byref(h))
        _lib.check(self.L, None, rc)
        self.h = h
        self.preallocation = Preallocation(self)

    def close(self):
        if getattr(self, "h", None):
            self.L.p2de_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, Uq):
        Uq = np.ascontiguousarray(Uq, dtype=np.float64)
        sz = self.sizes
        if Uq.shape != (sz.K, sz.Nq, sz.Nc):
            raise ValueError(f"Uq must have shape {(sz.K, sz.Nq, sz.Nc)}, got {Uq.shape}")
        _lib.check(self.L, self.h, self.L.p2de_set_state(self.h, Uq.ctypes.data))

    def set_stream(self, cuda_stream_ptr: int):
        _lib.check(self.L, self.h, self.L.p2de_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _lib.check(self.L, self.h, self.L.p2de_synchronize(self.h))

    def ssp33_step(self, t: float) -> float:
        dt = C.c_double()
        _lib.check(self.L, self.h, self.L.p2de_ssp33_step(self.h, t, C.byref(dt)))
        return dt.value

    def ssp33_step_async(self, t: float):
        _lib.check(self.L, self.h, self.L.p2de_ssp33_step_async(self.h, t))

    def last_dt(self) -> float:
        dt = C.c_double()
        _lib.check(self.L, self.h, self.L.p2de_last_dt(self.h, C.byref(dt)))
        return dt.value

    def reduce(self, what: int) -> float:
        out = C.c_double()
        _lib.check(self.L, self.h, self.L.p2de_reduce(self.h, what, C.byref(out)))
        return out.value

    def profile(self, enable: bool):
        _lib.check(self.L, self.h, self.L.p2de_profile(self.h, int(enable)))

    def profile_get(self, kernel_id: int):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.L, self.h, self.L.p2de_profile_get(self.h, kernel_id, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_state_async_ptr(self, host_ptr: int):
        _lib.check(self.L, self.h, self.L.p2de_set_state_async(self.h, C.c_void_p(host_ptr)))

    def get_state_async_ptr(self, host_ptr: int):
        _lib.check(self.L, self.h, self.L.p2de_get_state_async(self.h, C.c_void_p(host_ptr)))

    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte ncclUniqueId (call on one rank, broadcast to the others)."""
        L = _lib.load()
        buf = (C.c_uint8 * 128)()
        _lib.check(L, None, L.p2de_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank: int, nranks: int, unique_id: bytes):
        """Join the y-stripe communicator: this handle owns stripe `rank` of `nranks`."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _lib.check(self.L, self.h, self.L.p2de_comm_init(self.h, rank, nranks, buf))

    def ssp33_run(self, t: float, max_steps: int = 2 ** 62):
        """The `while t < T` loop of SSP33! inside the library (p2de_ssp33_run): returns (t, dthist)."""
        tt, n = C.c_double(t), C.c_int64()
        cap = min(max_steps, 1 << 22)
        dth = np.empty(cap, dtype=np.float64)
        _lib.check(self.L, self.h, self.L.p2de_ssp33_run(self.h, C.byref(tt), cap, C.byref(n), dth.ctypes.data))
        return tt.value, dth[:n.value].copy()

    def snapshot_ring(self, slots: int, output_interval: int):
        _lib.check(self.L, self.h, self.L.p2de_snapshot_ring(self.h, int(slots), int(output_interval)))

    def snapshots(self):
        """[(t, step, Uq)] of the snapshots still in the ring, oldest first."""
        n = int(self.L.p2de_snapshot_count(self.h))
        out, sz = [], self.sizes
        for idx in range(n):
            U = np.empty((sz.K, sz.Nq, sz.Nc), dtype=np.float64)
            t, step = C.c_double(), C.c_int64()
            rc = self.L.p2de_snapshot_get(self.h, idx, U.ctypes.data, C.byref(t), C.byref(step))
            if rc == 0:
                out.append((t.value, int(step.value), U))
        return out

    def error_sums(self, exact: np.ndarray) -> np.ndarray:
        """p2de_calculate_error: [6, Nc] = sum wJ|ex-U|, sum wJ|ex-U|^2, max|ex-U|, sum wJ|ex|, sum wJ ex^2, max|ex|."""
        sz = self.sizes
        ex = np.ascontiguousarray(exact, dtype=np.float64)
        if ex.shape != (sz.K, sz.Nq, sz.Nc):
            raise ValueError(f"exact must have shape {(sz.K, sz.Nq, sz.Nc)}, got {ex.shape}")
        out = np.empty((6, sz.Nc), dtype=np.float64)
        _lib.check(self.L, self.h, self.L.p2de_calculate_error(self.h, ex.ctypes.data, out.ctypes.data))
        return out

    DBG_NAMES = ("cta_general", "cta_interior", "cta_defer", "elem_logs", "elem", "lines", "lines_not_easy", "limiter_slow")

    def debug_counters(self, enable: bool = True) -> dict:
        """p2de_debug_counters: current values (since counting was enabled), then start (enable) / stop counting."""
        buf = (C.c_uint64 * len(self.DBG_NAMES))()
        _lib.check(self.L, self.h, self.L.p2de_debug_counters(self.h, int(enable), buf))
        return dict(zip(self.DBG_NAMES, (int(v) for v in buf)))

    def kernel_launch_count(self) -> int:
        return int(self.L.p2de_kernel_launch_count(self.h))


def initialize_DG(param: Param, initial_condition: Callable, initial_boundary_conditions: Callable, **state_kw):
    """initialize_DG (init.jl:44-60): reference data, operators, BC callback, device state, init_U!."""
    rd, md, discrete_data = initialize_data(param)
    bcdata = initial_boundary_conditions(param, md)
    solver = Solver(param=param, rd=rd, md=md, discrete_data=discrete_data)
    state = State(solver, bcdata, **state_kw)
    state.set_state(sample_initial_condition(param, md, initial_condition))
    return solver, state, StateParam(bcdata=bcdata)


def rhs(state: State, solver: Solver, state_param: StateParam, time_param: TimeParam) -> float:
    """`rhs!` (rhs.jl:5-13): fills rhsU (and L / L_local[..., nstage]); returns dt."""
    dt = C.c_double()
    _lib.check(state.L, state.h, state.L.p2de_rhs(state.h, time_param.t, time_param.dt, time_param.nstage, C.byref(dt)))
    return dt.value


def apply_rhs_limiter(limiter, state: State, solver: Solver, state_param: StateParam, time_param: TimeParam) -> None:
    """`apply_rhs_limiter!` (limiter.jl:8-56).  On the device the limiter is fused behind the two
    RHS evaluations, so this re-evaluates the stage with the limiter's `time_param.dt`; the
    observable result (rhsU, L, L_local[..., nstage]) is the reference's."""
    if limiter is not solver.param.rhs_limiter and limiter != solver.param.rhs_limiter:
        raise ValueError("limiter differs from solver.param.rhs_limiter")
    rhs(state, solver, state_param, time_param)


def check_conservation(state: State, solver: Solver) -> float:
    """check_conservation (dg/utils.jl:1-12), reduced on the device."""
    return state.reduce(T.REDUCE_CONSERVATION)


def SSP33(state: State, solver: Solver, state_param: StateParam, verbose: bool = False, max_snapshots: int = 64) -> DataHistory:
    """`SSP33!` (SSPRK33.jl:1-62): the whole time loop runs inside the library (p2de_ssp33_run), three fused stages per
    step on the device; the states the reference pushes to Uhist every `output_interval` steps and at the final time are
    kept in a device-side ring (the newest `max_snapshots`) and read back afterwards.  Lhist / thetahist carry the
    values of the last step (State.jl:21-24 keeps one L per stage, not a history)."""
    tp = solver.param.timestepping_param
    output_interval = solver.param.postprocessing_param.output_interval
    hist = DataHistory()
    state.snapshot_ring(max_snapshots, output_interval)
    t, dth = state.ssp33_run(tp.t0)
    hist.dthist.extend(float(d) for d in dth)
    for ts, step, U in state.snapshots():
        hist.thist.append(ts)
        hist.Uhist.append(U)
        hist.Lhist.append(state.preallocation.L)
        hist.thetahist.append(state.preallocation.theta)
        if verbose:
            print(f"Snapshot at time {ts}, step {step}, final time {tp.T}")
    if verbose:
        print("total_conservation =", check_conservation(state, solver))
    state.snapshot_ring(0, 0)
    return hist


def calculate_error(state: State, solver: Solver, exact_sol: Callable, verbose: bool = False) -> ErrorData:
    """calculate_error (postprocess.jl:1-46); `exact_sol(equation, x[, y], t)` returns primitives.  The caller's callback
    is evaluated on the host (it is the caller's code), the weighted sums and maxima over the mesh are taken on the device
    (p2de_calculate_error)."""
    param, md = solver.param, solver.md
    Tend = param.timestepping_param.T
    args = (md.xq,) if md.yq is None else (md.xq, md.yq)
    ex = np.stack([np.broadcast_to(np.asarray(c, dtype=np.float64), md.xq.shape) for c in
                   primitive_to_conservative(param.equation, exact_sol(param.equation, *args, Tend))], axis=-1)
    L1err, L2err, Linferr, L1ex, L2ex, Linfex = state.error_sums(ex)
    L1 = L2 = Linf = 0.0
    for c in range(param.equation.Nc):
        if Linfex[c] > 1e-14:
            L1 += L1err[c] / L1ex[c]
            L2 += math.sqrt(L2err[c]) / math.sqrt(L2ex[c])
            Linf += Linferr[c] / Linfex[c]
    if verbose:
        print(f"N = {param.N}, K = {param.K}\nL1 error is {L1}\nL2 error is {L2}\nLinf error is {Linf}")
    return ErrorData(L1, L2, Linf)
