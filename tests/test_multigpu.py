"""N>1 path: host-side partition logic on CPU (gloo, world_size 2) and, when the box has >= 2 GPUs,
the real NCCL halo exchange checked bitwise against the single-GPU result (tests/multigpu_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import problems as P
from p2de_b200 import initialize_data, structured_mapP
from p2de_b200.partition import halo_rows, local_bcdata, local_param, stripe_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stripes_tile_the_mesh_and_bc_lists_are_renumbered():
    param, rd, md, dd, bc, U0 = P.setup(P.dmr(N=2, K=(6, 7)))
    Kx, Ky = param.K
    Nfp = dd.sizes.Nfp
    rows = [stripe_rows(Ky, r, 3) for r in range(3)]
    assert rows[0][0] == 0 and rows[-1][1] == Ky and all(rows[i][1] == rows[i + 1][0] for i in range(2))
    nI = nO = 0
    for r in range(3):
        lp = local_param(param, r, 3)
        lb = local_bcdata(param, bc, r, 3)
        iy0, iy1 = rows[r]
        assert lp.K == (Kx, iy1 - iy0)
        assert abs(lp.xL[1] - (param.xL[1] + iy0 * (param.xR[1] - param.xL[1]) / Ky)) < 1e-15
        # local index + offset == a global index of the same list, values follow
        glob = lb.mapI + iy0 * Kx * Nfp
        assert set(glob) <= set(bc.mapI) and set(lb.mapO + iy0 * Kx * Nfp) <= set(bc.mapO)
        for g, v in zip(glob, lb.Ival):
            assert np.array_equal(v, bc.Ival[list(bc.mapI).index(g)])
        nI += len(lb.mapI); nO += len(lb.mapO)
        # stripe coordinates are the global ones
        lrd, lmd, ldd = initialize_data(lp)
        assert np.allclose(lmd.yq, md.yq[iy0 * Kx:iy1 * Kx], rtol=0, atol=1e-14)
    assert nI == len(bc.mapI) and nO == len(bc.mapO)


def _gloo_worker(rank, world, port, q, periodic=True):
    """Two ranks exchange their boundary element rows the way the library does over NCCL and
    check that the halo row equals the neighbour's owned row of the global array; the CFL dt is
    min-all-reduced."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Kx, Ky, Nq = 5, 6, 9
    rng = np.random.default_rng(7)
    U = rng.standard_normal((Kx * Ky, Nq, 4))                  # same global array on every rank
    iy0, iy1 = stripe_rows(Ky, rank, world)
    mine = U[iy0 * Kx:iy1 * Kx]
    bottom, top = halo_rows(mine, Kx)
    lo, hi = (rank - 1) % world, (rank + 1) % world             # p2de_comm_init: rank_lo, rank_hi
    has_lo, has_hi = periodic or rank > 0, periodic or rank < world - 1
    ghost_lo, ghost_hi = torch.empty(Kx, Nq, 4, dtype=torch.float64), torch.empty(Kx, Nq, 4, dtype=torch.float64)
    reqs = []                                                   # same issue order as exchange_rows (capi.cu)
    if has_hi: reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(top)), hi))
    if has_lo: reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(bottom)), lo))
    if has_lo: reqs.append(dist.irecv(ghost_lo, lo))
    if has_hi: reqs.append(dist.irecv(ghost_hi, hi))
    for r in reqs:
        r.wait()
    ok = True
    if has_lo: ok &= np.array_equal(ghost_lo.numpy(), U[((iy0 - 1) % Ky) * Kx:((iy0 - 1) % Ky + 1) * Kx])
    if has_hi: ok &= np.array_equal(ghost_hi.numpy(), U[(iy1 % Ky) * Kx:(iy1 % Ky + 1) * Kx])
    dt = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MIN)
    ok &= float(dt) == 0.1
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.parametrize("world,periodic", [(2, True), (2, False), (4, True), (4, False)],
                         ids=["world2-periodic", "world2-open", "world4-periodic", "world4-open"])
def test_halo_exchange_order_and_dt_allreduce_gloo(world, periodic):
    """world 2, periodic: both neighbours are the same rank; world 4: distinct neighbours, and with open ends the first and
    last stripe have one neighbour only (the cases of has_lo / has_hi in p2de_comm_init)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 7 * world + int(periodic)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q, periodic)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res == [(r, True) for r in range(world)]


@pytest.mark.gpu
def test_two_gpu_stripes_match_single_gpu_bitwise():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("bitwise_equal=True") == 3
