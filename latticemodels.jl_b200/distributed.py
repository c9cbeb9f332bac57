"""Multi-GPU plumbing: one process per GPU, Psi columns sharded, H replicated
(SURVEY.md section 8e).  ``torch.distributed`` only moves the 128-byte NCCL id; the per-frame
all-reduce of [rho | J] happens inside the library on its own communicator."""
from __future__ import annotations

import os

import numpy as np

from .context import Context, shard_range, unique_id  # noqa: F401


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def attach_communicator(ctx: Context, peer_slot_doubles: int = 4 << 20):
    """Create the library's NCCL communicator over an initialised torch.distributed group."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ctx
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = unique_id() if rank == 0 else bytes(128)
    dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    ctx.comm_init(bytes(t.cpu().tolist()), rank, world)
    if peer_slot_doubles and 2 <= world <= 8 and dist.get_backend() == "nccl":
        # NVLink peer-memory exchange for the per-frame reduction: all-gather the IPC handles
        mine = torch.tensor(list(ctx.peer_handle(peer_slot_doubles)), dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.peer_attach(b"".join(bytes(h.cpu().tolist()) for h in allh))
    return ctx


def allreduce_host(partial: np.ndarray) -> np.ndarray:
    """Sum per-rank partial observables on the host through torch.distributed (any backend).
    Used by the CPU (gloo) tests of the sharding logic and as the reduction for ranks that run
    without a library communicator."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return partial
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
