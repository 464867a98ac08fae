"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for rendezvous only.

The endpoint index shards by endpoint range: rank g owns endpoints
[g·E/G, (g+1)·E/G).  The data-path collectives (presence masks, per-request
(score, endpoint) pairs) run inside libfi_epp over its own NCCL communicator; this
module only distributes the communicator id, picks shard ranges, and reduces timings.
Works on the gloo backend too (CPU tests of the host logic, world_size 2).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(num_endpoints: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, count) of the pool owned by `rank`: contiguous, balanced, covers the pool exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(num_endpoints, world)
    begin = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return begin, count


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist

    if dist.is_initialized():
        return
    rank, world, local = env_rank_world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(local)
    dist.init_process_group(backend=backend, rank=rank, world_size=world)


def _device():
    import torch
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Rank `src` supplies `payload`; every rank returns it (used for the NCCL unique id)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert payload is not None
        return payload
    t = torch.zeros(nbytes, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def all_gather_array(a: np.ndarray) -> np.ndarray:
    """[world, ...] stack of the ranks' equally-shaped arrays (host logic / tests)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return a[None]
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw.copy()).to(_device())
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in outs])


def setup_sharded_picker(make_cfg, num_endpoints: int):
    """Create this rank's EndpointPicker for its endpoint-range shard and join the
    library's communicator.  make_cfg(begin, count, device) -> fi_epp_config."""
    import torch.distributed as dist

    from .picker import EndpointPicker

    rank, world, local = env_rank_world()
    begin, count = shard_range(num_endpoints, rank, world)
    picker = EndpointPicker(make_cfg(begin, count, local))
    if world > 1:
        uid = EndpointPicker.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 128, src=0)
        picker.comm_init(uid, rank, world)
        if dist.is_initialized():
            dist.barrier()
    return picker
