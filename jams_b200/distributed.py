"""Multi-GPU plumbing: one process per GPU, x-slab decomposition, halo handles exchanged over
``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).  The data path itself never goes through a
collective: boundary planes are stored straight into the neighbours' ghost planes by the stage kernels
(P2P stores over NVLink) and ordered with epoch flags (jb_halo_connect, include/jams_b200.h).  Only the
monitor reductions use an all-reduce.
"""
from __future__ import annotations

import numpy as np


def slab_range(nx_global: int, rank: int, n_ranks: int):
    """Contiguous x-range [x0, x0+nx) of the global site order owned by ``rank``.
    Equal slabs are required (every rank writes into its neighbours' boxes with its own geometry)."""
    if nx_global % n_ranks != 0:
        raise RuntimeError(f"lattice size along x ({nx_global}) must be divisible by the number of ranks ({n_ranks})")
    nx = nx_global // n_ranks
    return rank * nx, nx


def ring_neighbours(rank: int, n_ranks: int, periodic_x: bool):
    """(lo, hi) neighbour ranks of a slab, None across an open boundary"""
    if n_ranks == 1:
        return None, None
    lo = rank - 1 if rank > 0 else (n_ranks - 1 if periodic_x else None)
    hi = rank + 1 if rank < n_ranks - 1 else (0 if periodic_x else None)
    return lo, hi


class TorchComm:
    """rank/world + the three collectives the host layer needs, on an initialised torch.distributed group"""

    def __init__(self, periodic_x=True, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        self.periodic_x = bool(periodic_x)
        self.device = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def all_gather_bytes(self, blob: bytes):
        t = self.torch.frombuffer(bytearray(blob), dtype=self.torch.uint8).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world_size)]
        self.dist.all_gather(out, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    def allreduce_sum(self, arr):
        t = self.torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy().reshape(np.shape(arr))

    def allreduce_max(self, value: float) -> float:
        t = self.torch.tensor([float(value)], dtype=self.torch.float64).to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self, ctx=None):
        if ctx is not None:
            ctx.synchronize()
        self.dist.barrier()

    def connect_halos(self, ctx):
        """exchange halo handles and map the ring neighbours' boxes"""
        blobs = self.all_gather_bytes(ctx.halo_export_handle())
        lo, hi = ring_neighbours(self.rank, self.world_size, self.periodic_x)
        ctx.halo_connect(blobs[lo] if lo is not None else None, blobs[hi] if hi is not None else None)
        self.barrier(ctx)
