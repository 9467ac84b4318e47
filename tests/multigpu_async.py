"""Run under torchrun on >= 2 GPUs: the overlapped / peer-to-peer halo exchange with steps enqueued back to back
(p2de_ssp33_step_async, no host synchronisation between steps) against the single-GPU result.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_async.py
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
from p2de_b200.api import State  # noqa: E402
from p2de_b200.partition import local_bcdata, local_param, stripe_rows  # noqa: E402
from p2de_b200.types import Solver  # noqa: E402
from p2de_b200 import initialize_data  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(os.environ.get("NSTEPS", "12"))
    profile = os.environ.get("PROFILE", "0") == "1"
    ok = True
    for name, problem, periodic in [("dmr", P.dmr(N=3, K=(64, 4 * world), T=1e9), (False, False)),
                                    ("kh", P.kelvin_helmholtz(N=3, K=(64, 3 * world), T=1e9), (True, True))]:
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
        t0 = param.timestepping_param.t0
        print(f"[{rank}] {name}: enqueue {nsteps} steps", flush=True)
        if profile:
            st.profile(True)
        for _ in range(nsteps):
            st.ssp33_step_async(t0)
        st.synchronize()
        print(f"[{rank}] {name}: done", flush=True)
        mine = torch.from_numpy(st.preallocation.Uq).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            full = State(Solver(param=param, rd=rd, md=md, discrete_data=dd), bc, device=local)
            full.set_state(U0)
            for _ in range(nsteps):
                full.ssp33_step_async(t0)
            full.synchronize()
            ref = full.preallocation.Uq
            got = torch.cat(parts).cpu().numpy()
            same = bool(np.array_equal(got, ref))
            print(f"[multigpu_async] {name}: world={world} steps={nsteps} bitwise_equal={same} max|diff|={np.abs(got - ref).max():.3e}", flush=True)
            ok = ok and (same or np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max())   # periodic y: the wrap rows run the general instantiation on one GPU
            full.close()
        dist.barrier()
        st.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
