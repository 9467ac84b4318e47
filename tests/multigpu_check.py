"""Run under torchrun (one rank per GPU): each rank advances its y-stripe with halo exchange over
NCCL; rank 0 also advances the whole mesh on its own GPU.  The stripes must reproduce the
single-GPU result BITWISE (every element sees the same neighbour data and runs the same code).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("P2DE_OVERLAP_MIN_ROWS", "3")   # exercise the overlapped exchange on these small stripes too (read at comm_init)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P  # noqa: E402
from p2de_b200 import (ESLimitedLowOrderPos, GaussCollocation, LaxFriedrichsOnProjectedVal,  # noqa: E402
                       NodewiseScaledExtrapolation, SubcellLimiter, ZhangShuLimiter)
from p2de_b200.api import State  # noqa: E402
from p2de_b200.partition import local_bcdata, local_param, stripe_rows  # noqa: E402
from p2de_b200.types import Solver  # noqa: E402
from p2de_b200 import initialize_data  # noqa: E402


def default_cases(quick=False):
    cases = [("dmr-subcell", P.dmr(N=3, K=(16, 12)), (False, False)),
             ("vortex-periodic-subcell", P.vortex(N=2, K=(6, 8), CFL=0.5, T=10.0), (True, True)),
             ("kh-periodic-zhangshu", P.kelvin_helmholtz(N=3, K=(8, 8), limiter=ZhangShuLimiter()), (True, True)),
             # wide enough for batches strictly inside the mesh: on one GPU the rows next to the cut run the kernel's
             # compile-time INTERIOR version, in the stripes its general version (same arithmetic, different instantiation)
             ("dmr-wide-subcell", P.dmr(N=3, K=(64, 8)), (False, False)),
             # the shipped-examples configuration (Gauss + NodewiseScaledExtrapolation): projected face states and the
             # un-symmetrised coefficients travel with the halo rows
             ("kh-gauss-nodewise-subcell", P.kelvin_helmholtz(N=2, K=(8, 8), basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation(),
                                                           rhs=ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal())), (True, True))]
    if quick:      # bench.py --gpus N: the headline configuration on a mesh with 3 rows per rank, plus one periodic case
        cases = [("dmr-wide-subcell", None, (False, False)), ("kh-periodic-subcell", None, (True, True))]
    return cases


def run_cases(rank, world, local, quick=False, nsteps=4):
    """Each rank advances its y-stripe with halo exchange over NCCL; rank 0 also advances the whole mesh on its own GPU.
    Needs an initialised torch.distributed NCCL group.  Returns one record per case on rank 0 (None elsewhere)."""
    out = []
    for name, problem, periodic in default_cases(quick):
        if problem is None:
            problem = P.dmr(N=3, K=(64, 3 * world)) if name.startswith("dmr") else P.kelvin_helmholtz(N=3, K=(64, 3 * world))
        param, rd, md, dd, bc, U0 = P.setup(problem)
        Kx, Ky = param.K
        lp = local_param(param, rank, world)
        lrd, lmd, ldd = initialize_data(lp, light=True)
        lbc = local_bcdata(param, bc, rank, world)
        iy0, iy1 = stripe_rows(Ky, rank, world)
        st = State(Solver(param=lp, rd=lrd, md=lmd, discrete_data=ldd), lbc, device=local, structured_bc=periodic)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(State.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        st.comm_init(rank, world, bytes(uid.cpu().tolist()))
        st.set_state(U0[iy0 * Kx:iy1 * Kx])
        t, dts = param.timestepping_param.t0, []
        for _ in range(nsteps):
            dt = st.ssp33_step(t); t += dt; dts.append(dt)
        mine = torch.from_numpy(st.preallocation.Uq).cuda()
        parts = [torch.empty((stripe_rows(Ky, r, world)[1] - stripe_rows(Ky, r, world)[0]) * Kx, *mine.shape[1:],
                             dtype=torch.float64, device="cuda") for r in range(world)]
        if len({p.shape for p in parts}) == 1:
            dist.all_gather(parts, mine)
        else:
            for r in range(world):
                if r == rank:
                    parts[r] = mine
                dist.broadcast(parts[r], r)
        if rank == 0:
            full = State(Solver(param=param, rd=rd, md=md, discrete_data=dd), bc, device=local)
            full.set_state(U0)
            t2, dts2 = param.timestepping_param.t0, []
            for _ in range(nsteps):
                dt = full.ssp33_step(t2); t2 += dt; dts2.append(dt)
            ref = full.preallocation.Uq
            got = torch.cat(parts).cpu().numpy()
            bitwise = bool(np.array_equal(got, ref) and dts == dts2)
            maxdiff = float(np.abs(got - ref).max())
            # the rows next to a cut run the INTERIOR instantiation on one GPU and the general one in the stripes (two
            # instantiations of the same source): the last bits may differ there
            ok = bitwise or (maxdiff <= 1e-13 * float(np.abs(ref).max()) and np.allclose(dts, dts2, rtol=1e-14, atol=0))
            out.append({"case": name, "mesh": [int(Kx), int(Ky)], "world": world, "steps": nsteps, "bitwise_equal": bitwise,
                        "max_abs_diff": maxdiff, "ok": bool(ok)})
            full.close()
        dist.barrier()
        st.close()          # ncclCommDestroy: one library communicator per case, not several alive at once
    return out if rank == 0 else None


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_cases(rank, world, local)
    ok = True
    if rank == 0:
        for r in res:
            print(f"[multigpu_check] {r['case']}: world={world} bitwise_equal={r['bitwise_equal']} max|diff|={r['max_abs_diff']:.3e} ok={r['ok']}")
            ok = ok and r["ok"]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
