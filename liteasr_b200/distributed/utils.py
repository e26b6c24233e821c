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


def _spawned(local_rank: int, func, cfg, world: int, port: int) -> None:
    """distributed_func (distributed/utils.py:101-116): one process per GPU, rank = local rank on this single node."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(local_rank), LOCAL_RANK=str(local_rank),
                      WORLD_SIZE=str(world))
    d = getattr(cfg, "distributed", None)
    if d is not None:
        for k, v in (("device_id", local_rank), ("rank", local_rank)):
            try:
                setattr(d, k, v)
            except Exception:  # noqa: BLE001  (read-only config objects)
                pass
    distributed_init("nccl" if torch.cuda.is_available() else "gloo")
    try:
        func(cfg)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def call_func(func, cfg, nprocs: int = None):
    """Launcher with the reference's name and behaviour (distributed/utils.py:119-139): no CUDA -> warn and return; one GPU (or
    ``cfg.distributed.world_size == 1``) -> call ``func(cfg)`` in-process; otherwise ``mp.spawn`` one process per local GPU,
    each of which initialises NCCL (single node: 127.0.0.1 rendezvous) and runs ``func(cfg)``.  ``func`` and ``cfg`` must be
    picklable, as with the reference."""
    import logging
    import socket

    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        logging.getLogger(__name__).warning("CUDA is NOT available!")
        return None
    world = getattr(getattr(cfg, "distributed", None), "world_size", None)
    n = nprocs if nprocs is not None else (int(world) if world else torch.cuda.device_count())
    if torch.cuda.device_count() == 1 or n == 1:
        logging.getLogger(__name__).info("using only one single GPU, not apply DDP training")
        return func(cfg)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(fn=_spawned, args=(func, cfg, n, port), nprocs=n, join=True)
    return None
