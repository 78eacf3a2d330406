"""Multi-GPU plumbing: one process per GPU, torch.distributed only for set-up.

The 2^n state is sharded by its top log2(world) index bits (reference qubits 0..g-1,
SURVEY.md section 8e).  torch.distributed (gloo or nccl) is used exactly once, to
all-gather the CUDA-IPC handles of the shards; after that every exchange is a direct
load/store on the partner's mapped buffer over NVLink inside libqipb200's own kernels.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .state import Context, State


def init_sharded_state(n: int, dtype=np.complex128, ctx: Optional[Context] = None) -> State:
    """Create this rank's shard and map all peers (torch.distributed must be initialised)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = ctx or Context(int(os.environ.get("LOCAL_RANK", rank)))
    st = State(n, dtype, ctx, rank=rank, world_size=world)
    if world > 1:
        amp, flag = st.ipc_export()
        gathered = [None] * world
        dist.all_gather_object(gathered, (amp, flag))
        st.ipc_import(b"".join(g[0] for g in gathered), b"".join(g[1] for g in gathered))
        dist.barrier()
    return st


def gather_state(st: State) -> Optional[np.ndarray]:
    """Download every shard (canonical layout) and concatenate on rank 0 (tests only)."""
    import torch.distributed as dist

    local = st.download()
    if st.world_size == 1:
        return local
    parts = [None] * st.world_size if dist.get_rank() == 0 else None
    dist.gather_object(local, parts, dst=0)
    if dist.get_rank() == 0:
        return np.concatenate(parts)
    return None
