"""Element-row (y-stripe) partition of a uniform quad mesh over the GPUs of one box.

With elements numbered x-fastest (`k = ix + iy*Kx`, src/dg/postprocess.jl:170-171) a stripe of
element rows is one contiguous k-range, so a rank's share of every `[.., K]` array of the
reference is a contiguous slice and only faces 3/4 (bottom/top) cross ranks.
Host-side logic only (numpy); the exchange itself is done by the library over NCCL
(p2de_comm_init) or, in the CPU tests, by torch.distributed/gloo.
"""
from __future__ import annotations

import dataclasses
from typing import Tuple

import numpy as np

from .types import BCData, Param


def stripe_rows(Ky: int, rank: int, nranks: int) -> Tuple[int, int]:
    """[iy0, iy1) of the element rows owned by `rank` (rows split as evenly as possible)."""
    base, rem = divmod(Ky, nranks)
    iy0 = rank * base + min(rank, rem)
    return iy0, iy0 + base + (1 if rank < rem else 0)


def local_param(param: Param, rank: int, nranks: int) -> Param:
    """Param of the stripe: same N/options, K = (Kx, rows of this rank), y-extent of the stripe."""
    Kx, Ky = int(param.K[0]), int(param.K[1])
    iy0, iy1 = stripe_rows(Ky, rank, nranks)
    hy = (param.xR[1] - param.xL[1]) / Ky
    return dataclasses.replace(param, K=(Kx, iy1 - iy0), xL=(param.xL[0], param.xL[1] + iy0 * hy),
                               xR=(param.xR[0], param.xL[1] + iy1 * hy))


def local_bcdata(param: Param, bc: BCData, rank: int, nranks: int) -> BCData:
    """Restrict global mapI/mapO/Ival (1-based linear indices into [Nfp, K]) to the stripe and
    renumber them locally.  mapP is dropped: stripes use the structured path of the library."""
    n = param.N + 1
    Nfp = 4 * n
    Kx, Ky = int(param.K[0]), int(param.K[1])
    iy0, iy1 = stripe_rows(Ky, rank, nranks)
    k0, k1 = iy0 * Kx, iy1 * Kx

    def pick(idx):
        k = (idx - 1) // Nfp
        keep = (k >= k0) & (k < k1)
        return keep, idx[keep] - k0 * Nfp

    keepI, mapI = pick(bc.mapI)
    _, mapO = pick(bc.mapO)
    return BCData(np.zeros((0, 0), dtype=np.int64), mapI, mapO, bc.Ival[keepI] if len(bc.mapI) else bc.Ival)


def halo_rows(U_owned: np.ndarray, Kx: int):
    """(bottom owned row, top owned row) of a stripe array whose first axis is the element index."""
    return U_owned[:Kx], U_owned[-Kx:]
