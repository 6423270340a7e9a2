"""Small helpers of the reference's ``hydragen/utils.py`` that the hot path and its tests use."""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def rdiff(a, b, eps: float = 1e-8):
    """Relative difference, the parity metric of every reference test (hydragen/utils.py:13-15)."""
    diff = (a - b).abs()
    return 2 * diff / (a.abs() + b.abs() + eps)


dtype_map = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def get_rank() -> int:
    """hydragen/utils.py:87-93: single node, LOCAL_RANK is the rank (RANK honoured when set)."""
    return int(os.environ.get("RANK", os.environ.get("LOCAL_RANK", "0")))


def get_world_size() -> int:
    """hydragen/utils.py:96-99."""
    return int(os.environ.get("WORLD_SIZE", os.environ.get("LOCAL_WORLD_SIZE", "1")))


def is_local() -> bool:
    return get_rank() == 0


def local_print(*args, **kwargs):
    """hydragen/utils.py:105-107."""
    if is_local():
        print(*args, **kwargs)


def maybe_init_dist(backend: str | None = None):
    """hydragen/utils.py:118-133: bring up one process per GPU (NCCL) when launched by torchrun;
    returns the rank, or None when running single-process.  ``backend='gloo'`` is the CPU test path."""
    rank, world = get_rank(), get_world_size()
    if world < 2:
        return None
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank
