"""Process-group helpers with the reference's names (distributed/utils.py:17-92).

The reference spawns one process per GPU with ``mp.spawn`` and a tcp:// init method; here ranks come from the environment
(``torchrun`` / ``python -m torch.distributed.run``: RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT), which is what the
bench driver uses.  Backend "nccl" on GPUs; "gloo" on CPU for the host-logic tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_master() -> bool:
    return get_rank() == 0


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def distributed_init(backend: str = None) -> int:
    """Initialise from the torchrun environment; returns the local device index.  Mirrors distributed_init (:65-92) incl. the
    dummy all_reduce that warms the communicator up (:85-86)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world <= 1:
        return local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, init_method="env://")
    t = torch.zeros(1, device=f"cuda:{local}" if backend == "nccl" else "cpu")
    dist.all_reduce(t)
    return local
