"""Data parallelism for the hot path (the reference's ``liteasr/distributed``): one process per GPU, NCCL over NVLink."""
from .flat_ddp import FlatDDP  # noqa: F401
from .utils import barrier, distributed_init, get_rank, get_world_size, is_master  # noqa: F401
