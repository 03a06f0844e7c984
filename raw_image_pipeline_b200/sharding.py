"""Frame sharding across the GPUs of one box: frames are independent, so the path
shards with no data-path collective -- one process per GPU, each owning a contiguous chunk of the
batch (or one camera stream).  The only communication is the bench's barrier and max-over-ranks
reduction of the step time, done with torch.distributed (NCCL on the GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank`; sizes differ by at most one and the
    chunks tile [0, n_items) in rank order (contiguous chunks keep pinned staging buffers local)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def stream_owner(stream_index: int, world: int) -> int:
    """Camera stream -> rank.  A stream never moves between GPUs, so the CCC Kalman state
    (2 floats per stream) stays on one device."""
    return stream_index % world


def max_over_ranks(value: float, device=None) -> float:
    """max of a per-rank scalar (step time) over the job; identity when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: int, seconds_max: float, world: int) -> float:
    """Whole-job units/s under weak scaling: every rank processed `units_per_rank` in at most
    `seconds_max` (the max over ranks)."""
    return world * units_per_rank / seconds_max
