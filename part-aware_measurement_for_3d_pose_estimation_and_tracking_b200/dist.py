"""Multi-GPU plumbing: independent sequences are partitioned across ranks (one process per GPU);
the algorithm has no exchange step, so the only collectives are a sum-reduction of run counters
and an optional gather of the fixed-stride per-frame result blocks.
Backend: NCCL over NVLink on the B200 box, gloo in the CPU tests."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_sequences(total: int, rank: int, world: int) -> List[int]:
    """Sequence ``s`` runs on rank ``s mod world`` (round-robin; a sequence never spans GPUs
    because its frames are serially dependent through the track state)."""
    return list(range(rank, total, world))


def init(backend: str | None = None):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_counters(counters):
    """In-place SUM all-reduce of a 1-D int64/float64 tensor of run counters (reports, frames,
    PCP hits/totals, MPJPE sum/count)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(out: Dict[str, "object"], dst: int = 0):
    """Gather the per-rank result blocks ``count/ids/joints`` (equal shapes on every rank) to
    ``dst``; returns a dict of tensors with a leading rank axis on ``dst`` and ``None`` elsewhere."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return {k: v[None] for k, v in out.items() if v is not None}
    world, rank = dist.get_world_size(), dist.get_rank()
    res = {}
    for k, v in out.items():
        if v is None:
            continue
        bufs = [torch.empty_like(v) for _ in range(world)] if rank == dst else None
        dist.gather(v, bufs, dst=dst)
        if rank == dst:
            res[k] = torch.stack(bufs)
    return res if rank == dst else None


def all_gather_results(out: Dict[str, "object"], gathered: Dict[str, "object"]):
    """One collective per result tensor: the fixed-stride blocks ``count/ids/joints`` of every rank into
    ``gathered[k]`` of shape ``(world,) + out[k].shape`` on every rank (NCCL all-gather over NVLink on the B200
    box).  Device resident on both sides, asynchronous on the current stream."""
    import torch.distributed as dist
    for k, buf in gathered.items():
        dist.all_gather_into_tensor(buf.view(-1), out[k].reshape(-1))
    return gathered
