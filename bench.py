#!/usr/bin/env python
"""bench.py — DOF-updates/s (FP64, per RK stage) of the per-stage DG RHS + limiter hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one iteration of the SSP33! time loop (3 stages: rhs! + limiter + SSP combine) over
the whole synthetic mesh.  `value` = 3 * K_elements * Nq * n_steps / time, state resident in HBM.
`e2e` = the same through the C ABI with HOST buffers: every step copies the state host->device
(pinned memory), runs the 3 stages and copies the state back.
`--impl reference` times the CPU oracle (a C++/OpenMP restatement of the reference's Julia
algorithm, SURVEY.md §8c: Julia is not installed on this image) on a bounded sample.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# host cores available to this process, taken before any OpenMP runtime binds the main thread (OMP_PROC_BIND below)
NCORES = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

METRIC = "DOF-updates/sec (FP64, per RK stage)"
UNIT = "DOF-updates/s"

# name -> (problem factory name in tests/problems.py, N, (Kx, Ky) per GPU, limiter)
WORKLOADS = {
    "S-DMR": dict(problem="dmr", N=3, K=(4096, 1024), note="2D double-Mach-reflection data, N=3 LGL, 4096x1024 quads (4.19M elements, 67.1M nodes), subcell positivity limiter; inflow/outflow BCs through the reference's BCData"),
    "S-DMR-small": dict(problem="dmr", N=3, K=(512, 128), note="S-DMR at 512x128"),
    "S-DMR-mid": dict(problem="dmr", N=3, K=(1024, 512), note="S-DMR at 1024x512 (profiling size)"),
    "S-KH": dict(problem="kelvin_helmholtz", N=4, K=(4096, 512), note="Kelvin-Helmholtz, N=4 LGL, 4096x512 per GPU, periodic"),
    "S-VORTEX": dict(problem="vortex", N=3, K=(2048, 2048), note="isentropic vortex (test/test_smoke.jl data), N=3 LGL, 2048x2048, periodic: smooth, the limiter never engages "
                     "(SURVEY.md 8d)"),
    # data-independence checks of the headline (same mesh and scheme as S-DMR, different data):
    "S-WAVE": dict(problem="wave2d", N=3, K=(4096, 1024), note="plateau-free smooth periodic data (tests/problems.py: wave2d), N=3 LGL, 4096x1024: "
                   "no constant elements, most node pairs off logmean's series branch (logs evaluated everywhere)"),
    # (a "developed" double Mach reflection is not available: with the copy-out top boundary the reference's BC surface offers --
    #  it has neither the wall nor the time-dependent exact-shock condition -- the flow where the oblique shock meets the top
    #  boundary accelerates without bound, in the oracle as well: max wavespeed 12 -> 100 within 400 steps on 128x32.  The
    #  long-time shock-dominated case is therefore the Sedov blast, which the reference ships and which stays stable.)
    "S-SEDOV-developed": dict(problem="sedov", N=3, K=(4096, 1024), developed=dict(K=(512, 128), steps=3000, tile=(8, 8)),
                              note="Sedov blast (examples/2D/sedov.jl data: background pressure 1e-5) advanced 3000 SSP-RK3 steps on a 512x128 periodic mesh "
                                   "and tiled 8x8 seamlessly: expanding strong shocks into near-vacuum, the limiter's exact evaluation active along every front"),
    # the configuration of examples/2D/kelvin-helmholtz.jl:44-55: Gauss collocation,
    # NodewiseScaledExtrapolation, LaxFriedrichsOnProjectedVal, subcell positivity limiter
    "S-KH-gauss": dict(problem="kelvin_helmholtz", N=3, K=(2048, 512), gauss=True,
                       note="Kelvin-Helmholtz, N=3 Gauss + NodewiseScaledExtrapolation, 2048x512 per GPU, periodic (generic kernels)"),
}
CPU_SAMPLE_K = (1024, 256)   # bounded sample of the same workload for the CPU arm: 4.2 M nodes, ~0.5 s per step on 16 cores


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_dof(N, limiter_code):
    """SURVEY.md §8d / DESIGN.md §5: Uq read 32 + resW 32 + Uq write 32 + L_local write."""
    n = N + 1
    if limiter_code == 2:
        return 96.0 + 8.0 * 2 * n * (n + 1) / (n * n)
    return 96.0 + 8.0 / (n * n)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_problem(workload, K):
    import problems as P
    from p2de_b200 import initialize_data
    w = WORKLOADS[workload]
    kw = {}
    if w.get("gauss"):
        from p2de_b200 import (ESLimitedLowOrderPos, GaussCollocation, LaxFriedrichsOnProjectedVal,
                               NodewiseScaledExtrapolation)
        kw = dict(basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation(),
                  rhs=ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal()))
    problem = getattr(P, w["problem"])(N=w["N"], K=K, **kw)
    param, ic, bcf = problem
    return param, ic, bcf


def boundary_data_light(param, workload):
    """BCData for the structured (mapP == NULL) path without building per-element arrays."""
    import problems as P
    from p2de_b200 import BCData, primitive_to_conservative
    w = WORKLOADS[workload]
    if w["problem"] != "dmr":
        return BCData(np.zeros((0, 0), dtype=np.int64), [], [], []), (True, True)
    n = param.N + 1
    Nfp = 4 * n
    Kx, Ky = param.K
    a = np.arange(n)

    def idx(k, face):
        return (k[:, None] * Nfp + face * n + a[None, :] + 1).reshape(-1)
    rows, cols = np.arange(Ky), np.arange(Kx)
    left, right = idx(rows * Kx, 0), idx(rows * Kx + Kx - 1, 1)
    bottom, top = idx(cols, 2), idx((Ky - 1) * Kx + cols, 3)
    Ival = np.tile(np.array(primitive_to_conservative(param.equation, P.DMR_POST)), (len(left), 1))
    return BCData(np.zeros((0, 0), dtype=np.int64), left, np.concatenate([right, bottom, top]), Ival), (False, False)


def initial_state(param, rd, ic, out):
    """Fill out[K, Nq, Nc] chunk by chunk (the full coordinate arrays are never materialised)."""
    from p2de_b200 import element_nodes
    K = out.shape[0]
    chunk = 1 << 18
    for s in range(0, K, chunk):
        k = np.arange(s, min(K, s + chunk))
        xq, yq = element_nodes(param, rd, k)
        out[s:s + len(k)] = np.stack([np.broadcast_to(c, xq.shape) for c in ic(param, xq, yq)], axis=-1)


def source_hash():
    """sha256 over the kernel sources (every file of csrc/): printed in the bench line so that records can be matched to code."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "p2de_b200", "csrc")
    for f in sorted(x for x in os.listdir(d) if x.endswith((".cu", ".cuh"))):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


def kernel_sass_hashes(lib, names):
    """{mangled name: sha256[:16] of its SASS} for the named kernels of a built library (cuobjdump -sass -fun), or None.

    profiles/r2_traffic.json is an ncu capture of three kernels; it is quoted only while the machine code of exactly those
    kernels in the library this run loaded is the one a build of the captured sources gives (the hash over all of csrc/
    also changes when an unrelated kernel does).  (tools/kernel_sass_hash.py compares whole libraries the same way; its hashes are taken
    over the full dump's slightly different framing of a function and are not interchangeable with these.)"""
    import hashlib
    import re
    import shutil
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        return None
    res = {}
    for want in names:
        try:
            out = subprocess.run([exe, "-sass", "-fun", want, lib], capture_output=True, text=True, timeout=60).stdout
        except Exception:
            return None
        name, body = None, []
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if name == want:
                    break
                name, body = m.group(1), []
            elif line.strip().startswith(".......") and name == want:
                break
            elif name == want:
                body.append(line.rstrip())
        if name != want or not body:
            return None
        res[want] = hashlib.sha256("\n".join(body).encode()).hexdigest()[:16]
    return res


def developed_state(workload, host_np, device):
    """Initial state of S-DMR-developed: a small double-Mach-reflection run on the GPU (through the same library), tiled."""
    import problems as P
    from p2de_b200.api import State
    from p2de_b200.types import Solver
    dv = WORKLOADS[workload]["developed"]
    w = WORKLOADS[workload]
    param, rd, md, dd, bc, U0 = P.setup(getattr(P, w["problem"])(N=w["N"], K=dv["K"], T=1e9))
    st = State(Solver(param=param, rd=rd, md=md, discrete_data=dd), bc, device=device)
    st.set_state(U0)
    t = 0.0
    for _ in range(dv["steps"]):
        t += st.ssp33_step(t)
    U = st.preallocation.Uq
    st.close()
    if not (np.isfinite(t) and np.isfinite(U).all() and (U[..., 0] > 0).all()):
        raise RuntimeError(f"{workload}: the small run did not stay finite / positive (t = {t})")
    kx, ky = dv["K"]; tx, ty = dv["tile"]
    Us = U.reshape(ky, kx, U.shape[1], U.shape[2])
    Kx = kx * tx
    for jy in range(ty):                       # fill row-block by row-block (no 2 GB temporary)
        rowblock = np.tile(Us, (1, tx, 1, 1)).reshape(ky * Kx, U.shape[1], U.shape[2])
        host_np[jy * ky * Kx:(jy + 1) * ky * Kx] = rowblock
    return {"small_mesh": list(dv["K"]), "steps": dv["steps"], "t_end": t, "tile": list(dv["tile"])}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from p2de_b200 import initialize_data
    from p2de_b200.api import State
    from p2de_b200.types import Solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(args, args.workload, world, rank, local, headline=True)
    if out is not None and rank == 0:
        if world == 1 and not args.no_extra and not args.no_e2e and args.workload == "S-DMR":
            # the headline data is the reference's prescribed initial condition, which is mostly plateau: the same mesh
            # and scheme on plateau-free and on developed-shock data, with the data-dependent shortcuts counted
            out["extra"] = {}
            for wl in ("S-WAVE", "S-SEDOV-developed"):
                try:
                    ex = measure(args, wl, world, rank, local, headline=False)
                    out["extra"][wl] = ex
                except Exception as e:      # an extra line must never cost the headline
                    out["extra"][wl] = {"error": repr(e)}
            vals = [v["value"] for v in out["extra"].values() if "value" in v] + [out["value"]]
            out["extra"]["worst_case_value"] = min(vals)
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.workload, steps=6, warmup=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def multi_gpu_check(world, rank, local):
    """Outside the timed region: y-stripes over `world` ranks against the whole mesh on rank 0's GPU (tests/multigpu_check.py)."""
    import multigpu_check as mc
    os.environ["P2DE_OVERLAP_MIN_ROWS"] = "3"       # the check's stripes are 3 rows tall: run the overlapped exchange on them
    res = mc.run_cases(rank, world, local, quick=True)
    if res is None:
        return None
    # the periodic-y case is held to 1e-13 relative instead: on ONE GPU the first and last element row reach their neighbours
    # through the wrap and run the kernel's general instantiation, in the stripes they are rows next to a halo row and run
    # the INTERIOR one (same source, two instantiations: last-bit differences)
    try:
        strict = [r for r in res if not r["case"].startswith("kh-periodic")]
        return {"world": world, "bitwise_equal": all(r["bitwise_equal"] for r in strict), "ok": all(r["ok"] for r in res),
                "bitwise_cases": [r["case"] for r in strict],
                "max_abs_diff_periodic": max([r["max_abs_diff"] for r in res if r not in strict], default=None), "cases": res}
    except Exception as e:      # (a summary must never cost the measured line)
        return {"world": world, "error": repr(e), "cases": res}


def measure(args, workload, world, rank, local, headline):
    import torch
    import torch.distributed as dist
    from p2de_b200 import initialize_data
    from p2de_b200.api import State
    from p2de_b200.types import Solver
    w = WORKLOADS[workload]
    K = w["K"]
    # weak scaling: every rank owns a K[0] x K[1] stripe of element rows of a K[0] x (K[1]*world)
    # mesh on a domain stretched in y accordingly (y-stripes)
    import dataclasses
    from p2de_b200.partition import local_bcdata, local_param
    gparam, ic, _ = build_problem(workload, K)
    Ly = gparam.xR[1] - gparam.xL[1]
    gparam = dataclasses.replace(gparam, K=(K[0], K[1] * world), xR=(gparam.xR[0], gparam.xL[1] + Ly * world))
    gbc, periodic = boundary_data_light(gparam, workload)
    param = local_param(gparam, rank, world)
    bc = local_bcdata(gparam, gbc, rank, world)
    rd, md, dd = initialize_data(param, light=True)
    solver = Solver(param=param, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc, device=local, structured_bc=periodic)
    stream = torch.cuda.current_stream()
    st.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(State.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        st.comm_init(rank, world, bytes(uid.cpu().tolist()))
    sz = dd.sizes
    host = torch.empty((sz.K, sz.Nq, sz.Nc), dtype=torch.float64, pin_memory=True)
    dev_info = None
    if w.get("developed"):
        dev_info = developed_state(workload, host.numpy(), local)
    else:
        initial_state(param, rd, ic, host.numpy())
    st.set_state_async_ptr(host.data_ptr())
    st.synchronize()
    t0 = param.timestepping_param.t0
    nbytes = host.numel() * 8

    trace = os.environ.get("P2DE_BENCH_TRACE") == "1"

    def tr(msg):
        if trace:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def barrier():
        # (device first: a host-side barrier collective must not be launched into a GPU that still runs the step's own
        #  collectives on another communicator)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    t = t0
    tr("state set")
    for _ in range(args.warmup):
        t += st.ssp33_step(t)
    tr("warm-up done")
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    st.profile(True)
    launches0 = st.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        # one more untimed step: its halo exchange aligns the ranks on the DEVICE timeline (after a host barrier the ranks'
        # first launches are milliseconds apart, and the early rank's clock would run while it waits for the late one)
        st.ssp33_step_async(t)
        st.profile(True)
        launches0 = st.kernel_launch_count()
    ev0.record(stream)
    tt = t
    for _ in range(args.steps):
        st.ssp33_step_async(tt)
        tt += 0.0          # t only enters through the (T - t) cap; T is far away for the bench configs
    ev1.record(stream)
    tr("timed steps enqueued")
    barrier()
    tr("timed steps done")
    ms = ev0.elapsed_time(ev1)
    launches = st.kernel_launch_count() - launches0
    stage_ms, n_stage = st.profile_get(0)
    upd_ms, n_upd = st.profile_get(1)
    proj_ms, n_proj = st.profile_get(2)       # Gauss: entropy-projection kernel (0 launches otherwise)
    gap_ms, n_gap = st.profile_get(100)       # device time between consecutive hot-path launches (exchanges, collectives, bubbles)
    st.profile(False)
    clocks = sampler.stop()
    per_rank = None
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        # every rank's own kernel time and its waiting time: the ranks run in lock step (three exchanges and one
        # all-reduce per step), so the slowest GPU of the box sets the pace and the others' waiting shows up between launches
        mine = torch.tensor([stage_ms / (3.0 * args.steps), gap_ms / args.steps],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"stage_kernels_ms_per_stage": [round(float(a[0]), 4) for a in allr],
                    "between_launches_ms_per_step": [round(float(a[1]), 4) for a in allr]}
    dof_per_stage = sz.K * sz.Nq * world
    value = 3.0 * dof_per_stage * args.steps / (ms * 1e-3)
    # data-dependent shortcuts taken, counted by the kernels themselves during one more (untimed) step
    counters = None
    if param.rhs_limiter.code == 2 and not w.get("gauss"):
        st.debug_counters(True)
        st.ssp33_step(t)
        c = st.debug_counters(False)
        if c["elem"] and c["lines"]:
            counters = {"elements_with_logs_frac": c["elem_logs"] / c["elem"], "lines_not_all_easy_frac": c["lines_not_easy"] / c["lines"],
                        "limiter_slow_calls_per_element": c["limiter_slow"] / c["elem"],
                        "interior_cta_frac": c["cta_interior"] / max(c["cta_interior"] + c["cta_general"], 1), "raw": c}
    peak, peak_src = peaks()
    A = algorithmic_bytes_per_dof(param.N, param.rhs_limiter.code)
    # all kernels of one RK stage, device time (per RK stage, not per launch: with several ranks a stage is two row-range
    # launches of the same kernel, boundary rows first)
    n_rk = 3 * args.steps
    stage_total_ms = (stage_ms + upd_ms + proj_ms) / n_rk
    achieved = A * sz.K * sz.Nq / (stage_total_ms * 1e-3) / 1e9 if n_stage else None
    if not headline:
        mr, mre = st.reduce(1), st.reduce(2)
        st.close()
        del host
        return {"value": value, "unit": UNIT, "ms_per_step": ms / args.steps, "stage_kernel_ms": stage_ms / n_rk,
                "roofline_frac": (achieved / peak) if achieved else None, "gpu_launches": int(launches), "counters": counters,
                "description": w["note"], "developed": dev_info, "min_rho": mr, "min_rhoe": mre}

    # ---- end to end through the C ABI with host buffers: every step = host->device copy of that step's
    # input state (pinned), the 3 stages, device->host copy of the result.  Steps are independent jobs,
    # so three handles on three streams are used in turn: the copies of one job overlap the compute and
    # the copies of the others (PCIe is full duplex); `serial_value` is the same with one handle on one stream.
    def e2e_loop(states, streams, hosts, n):
        for i in range(n):
            j = i % len(states)
            with torch.cuda.stream(streams[j]):
                states[j].set_state_async_ptr(hosts[j].data_ptr())
                states[j].ssp33_step_async(t0)
                states[j].get_state_async_ptr(hosts[j].data_ptr())

    def timed_e2e(states, streams, hosts, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for sx in streams[1:]:
            sx.wait_stream(streams[0])
        e0.record(streams[0])
        for sx in streams[1:]:
            sx.wait_stream(streams[0])
        e2e_loop(states, streams, hosts, n)
        for sx in streams[1:]:
            streams[0].wait_stream(sx)
        e1.record(streams[0])
        barrier()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            tm = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            t_ms = float(tm.item())
        return t_ms

    e2e_steps = max(3, min(args.steps, 9))
    if args.no_e2e:      # kernel A/B runs only (tools/ab_variants.sh): not a bench line the driver reads
        return {"value": value, "ms_per_step": ms / args.steps, "stage_kernel_ms": stage_ms / n_rk,
                "update_kernel_ms": upd_ms / max(n_upd, 1), "n_stage": n_stage, "n_upd": n_upd, "counters": counters,
                "between_launches_ms_per_step": gap_ms / args.steps, "per_rank": per_rank,
                "gpu_launches": int(launches), "clocks": clocks, "lib": os.environ.get("P2DE_B200_LIB", "in-tree")}
    tr("e2e serial")
    e2e_loop([st], [stream], [host], 2)
    if not w.get("developed"):
        initial_state(param, rd, ic, host.numpy())
    serial_ms = timed_e2e([st], [stream], [host], e2e_steps)
    tr("e2e serial done")
    # two more handles + streams + pinned buffers for the pipelined variant (one job = ~2 copy times + 1 compute
    # time on its own stream, so three jobs in flight keep both copy directions busy)
    states, streams, hosts = [st], [stream], [host]
    # (single GPU only: with several ranks the three handles would be three NCCL communicators working concurrently on
    #  different streams, whose kernels the GPUs may start in different orders -- a known way to deadlock NCCL)
    for _ in range(2 if world == 1 else 0):
        sx = torch.cuda.Stream()
        s2 = State(solver, bc, device=local, structured_bc=periodic)
        s2.set_stream(sx.cuda_stream)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.tensor(list(State.comm_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, 0)
            s2.comm_init(rank, world, bytes(uid.cpu().tolist()))
        h2 = torch.empty_like(host).pin_memory()
        h2.copy_(host)
        states.append(s2); streams.append(sx); hosts.append(h2)
    if world == 1:
        e2e_loop(states, streams, hosts, 3)
        torch.cuda.synchronize()
        for hx in hosts[:-1]:
            hx.copy_(hosts[-1])
        e2e_ms = timed_e2e(states, streams, hosts, e2e_steps)
        e2e_value = 3.0 * dof_per_stage * e2e_steps / (e2e_ms * 1e-3)
    else:
        e2e_ms, e2e_value = None, None
    tr("e2e done")
    serial_value = 3.0 * dof_per_stage * e2e_steps / (serial_ms * 1e-3)
    for s2 in states[1:]:
        s2.close()

    # DRAM traffic and FP64 instruction counts come from one ncu --set full capture (profiles/r2_traffic.json, written by
    # tools/summarize_ncu.py); they are only quoted for the kernel sources they were captured from (source_hash)
    traffic, fp64, traffic_note = None, None, "no ncu capture of these kernel sources in profiles/ (traffic: null)"
    try:
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tall = json.load(f)
            tj = tall.get(workload, {})
            want = tall.get("kernel_sass", {})
            from p2de_b200.lib import SO_PATH
            have = kernel_sass_hashes(SO_PATH, sorted(want)) if want else None
            if tall.get("source_hash") == source_hash() or (have is not None and have == want):
                traffic, fp64 = tj.get("per_stage_total_bytes"), tj.get("fp64")
                traffic_note = ("DRAM bytes per stage (dram__bytes_read.sum + dram__bytes_write.sum of the three stage kernels of one step / 3) from one "
                                "ncu --set full capture (profiles/r2_traffic.json); quoted because "
                                + ("the kernel sources are the captured ones (source hash checked)" if tall.get("source_hash") == source_hash() else
                                   "the SASS of stage_subcell_s1/s2/s3<4,8> in the library this run loaded is identical to a build of the captured sources "
                                   "(cuobjdump hashes checked against the file's kernel_sass; other kernels of csrc/ have changed since)"))
            else:
                traffic_note = ("profiles/r2_traffic.json was captured from other kernel code (source hash and kernel SASS hashes both differ"
                                + (", cuobjdump unavailable" if (want and have is None) else "") + "): not quoted")
    except Exception as e:      # evidence look-up only: it must never cost the measured line
        traffic, fp64, traffic_note = None, None, f"profiles/r2_traffic.json could not be matched to the loaded library ({e!r}): not quoted"
    fp64_peak = None
    ppath = os.path.join(ROOT, "profiles", "r1_fp64_peak.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            fp64_peak = json.load(f)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "description": w["note"], "N": param.N, "elements_per_gpu": list(K),
                   "dof_updates_per_stage": dof_per_stage, "stages_per_step": 3,
                   "l2": "state arrays (2.1 GB each at S-DMR) are far larger than the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world} (y-stripes)" if world > 1 else "single GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                     "traffic_note": traffic_note,
                     "kernel": "all kernels of one RK stage (subcell family: one stage_subcell_s1/s2/s3 launch; generic path: + projection / update kernels)", "algorithmic_bytes_per_dof_update": A,
                     "stage_kernel_ms": stage_ms / n_rk, "stage_kernel_launches_per_stage": n_stage / n_rk,
                     "update_kernel_ms": upd_ms / max(n_upd, 1),
                     "stage_kernel_share": stage_ms / max(stage_ms + upd_ms + proj_ms, 1e-30),
                     "projection_kernel_ms": (proj_ms / n_proj) if n_proj else None},
        # second view of the same stage: the stage kernel is bound by the FP64 CUDA-core pipe, not by HBM.
        # flops per DOF-update counted by ncu (DFMA = 2), peak = DFMA micro-benchmark on this GPU type
        # (tools/fp64_peak.cu, profiles/r1_fp64_peak.json)
        "roofline_fp64": ({"bound": "fp64", "unit": "TFLOP/s",
                           "achieved": fp64["flops_per_dof_update"] * sz.K * sz.Nq / (stage_ms / n_rk * 1e-3) / 1e12,
                           "peak": fp64_peak["dfma_tflops"],
                           "frac": fp64["flops_per_dof_update"] * sz.K * sz.Nq / (stage_ms / n_rk * 1e-3) / 1e12 / fp64_peak["dfma_tflops"],
                           "pipe_frac": fp64["lane_ops_per_dof_update"] * sz.K * sz.Nq / (stage_ms / n_rk * 1e-3)
                                        / (fp64_peak["dadd_per_clk_sm"] * fp64_peak["sms"] * fp64_peak["clock_mhz"] * 1e6),
                           "kernel": "stage_subcell_s1/s2/s3", "flops_per_dof_update": fp64["flops_per_dof_update"],
                           "source": "ncu thread-instruction counts (profiles/r2_traffic.json) and tools/fp64_peak.cu (profiles/r1_fp64_peak.json)"}
                          if (fp64 and fp64_peak and n_stage) else None),
        # headline e2e = ONE handle on ONE stream: host->device copy of the state, the step, device->host copy, strictly in
        # sequence like a time loop that needs step n's output before step n+1 (PCIe-bound: 4.3 GB per step).  The pipelined
        # figure (three independent jobs in flight, copies overlapping compute) is kept beside it.
        "e2e": {"value": serial_value, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "steps": e2e_steps, "ms_per_step": serial_ms / e2e_steps,
                "mode": "serial: one handle, one stream, copy in -> 3 stages -> copy out per step",
                "pipelined_value": e2e_value, "pipelined_ms_per_step": (e2e_ms / e2e_steps) if e2e_ms else None,
                "pipelining": "3 handles on 3 streams used in turn: the H2D / D2H copies of one step overlap the compute and copies of the others"},
        "gpu_launches": int(launches), "clocks": clocks, "counters": counters,
        "between_launches_ms_per_step": gap_ms / args.steps, "per_rank": per_rank,
        "source_hash": source_hash(),
    }
    if world > 1 and not args.no_multi_gpu_check:
        out["multi_gpu_check"] = multi_gpu_check(world, rank, local)
    st.close()
    return out


def cpu_baseline(workload, steps, warmup):
    """The oracle (C++/OpenMP restatement of the reference's CPU algorithm) on all host cores, on a bounded sample of
    the same workload: built here with -O3 -march=native -ffp-contract=off, threads pinned (OMP_PROC_BIND=close, set in
    main() before the OpenMP runtime loads), per-phase wall time under the reference's TimerOutputs labels."""
    import problems as P
    from oracle import oracle as O
    flags = O.use_native_build()
    K = CPU_SAMPLE_K
    param, ic, bcf = build_problem(workload, K)
    param_, rd, md, dd, bc, U0 = P.setup((param, ic, bcf))
    cores = NCORES
    orc = O.Oracle(param, dd, bc, threads=cores)
    orc.set_state(U0)
    t = param.timestepping_param.t0
    for _ in range(warmup):
        t += orc.ssp33_step(t)
    orc.phase_times(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        t += orc.ssp33_step(t)
    el = time.perf_counter() - t0
    phases = orc.phase_times()
    sz = dd.sizes
    v = 3.0 * sz.K * sz.Nq * steps / el
    return {"value": v, "unit": UNIT, "cores": int(O.lib().oracle_max_threads()), "kind": "port",
            "sample": f"{workload} data on a {K[0]}x{K[1]} mesh (N={param.N}, {sz.K * sz.Nq} nodes), {steps} SSP-RK3 steps after {warmup} warm-up; "
                      "C++/OpenMP restatement of the reference's CPU algorithm (Julia is not installed on this image)",
            "same_config": False, "build": flags, "omp_proc_bind": os.environ.get("OMP_PROC_BIND"),
            "ms_per_step": el / steps * 1e3,
            # seconds per SSP-RK3 step under the reference's TimerOutputs labels (rhs.jl:6-51, limiter.jl:9-51, SSPRK33.jl:29)
            "phase_ms_per_step": {k: round(1e3 * x / steps, 3) for k, x in phases.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cb = cpu_baseline(args.workload, steps=args.steps, warmup=args.warmup)
    w = WORKLOADS[args.workload]
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "description": w["note"], "N": w["N"], "elements_per_gpu": list(w["K"]),
                      "sample_elements": list(CPU_SAMPLE_K)},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S-DMR", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident timing only (kernel A/B runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the S-WAVE / S-DMR-developed lines under `extra`")
    ap.add_argument("--no-multi-gpu-check", action="store_true", help="skip the stripes-vs-single-GPU check at N > 1")
    args = ap.parse_args()
    os.environ.setdefault("OMP_PROC_BIND", "close")     # CPU arm: pinned threads (read when libgomp loads)
    os.environ.setdefault("OMP_PLACES", "cores")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
