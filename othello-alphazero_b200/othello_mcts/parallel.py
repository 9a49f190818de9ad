"""Game-level data parallelism over the GPUs of one box.

Self-play games share nothing but read-only weights (the reference itself plays them one after another,
train.py:394-398), so the path shards by game id with NO collective on the hot path: rank r owns a
contiguous block of game ids, its own node pools and a full copy of the folded weights.  The engine's
random streams are keyed by game id, so a game's record does not depend on which rank played it.  The
only communication is the final gather of records / training samples (and counters) to rank 0.
"""
from __future__ import annotations

import numpy as np
import torch


def partition_games(total_games: int, world_size: int, rank: int) -> tuple[int, int]:
    """(first_game_id, num_games) of `rank`: contiguous blocks, sizes differ by at most one."""
    if total_games < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad partition request: total_games={total_games}, world_size={world_size}, rank={rank}")
    base, extra = divmod(total_games, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owner_of(game_id: int, total_games: int, world_size: int) -> int:
    """Rank that owns `game_id` under partition_games."""
    base, extra = divmod(total_games, world_size)
    split = extra * (base + 1)
    if game_id < split:
        return game_id // (base + 1)
    return extra + (game_id - split) // max(base, 1)


def _gather_tensor(t: torch.Tensor, dst: int, group) -> torch.Tensor | None:
    """Concatenate per-rank tensors (dim 0 may differ) on `dst`.  Works on gloo (CPU) and nccl (CUDA)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    home = t.device
    if dist.get_backend(group) == "nccl" and not t.is_cuda:      # NCCL moves device memory: stage host tensors on this rank's GPU
        t = t.cuda()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes) if sizes else 0
    padded = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst, group=group)
        return torch.cat([b[:k] for b, k in zip(bufs, sizes)], dim=0).to(home)
    dist.gather(padded, None, dst=dst, group=group)
    return None


def gather_samples(samples: dict, dst: int = 0, group=None) -> dict | None:
    """Training tensors of every rank ({"features","policies","values"}, rank order = game-id order) on `dst`.
    Without an initialised process group this is the identity."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return samples
    out = {k: _gather_tensor(v, dst, group) for k, v in samples.items()}
    return out if dist.get_rank(group) == dst else None


def gather_records(records: dict, dst: int = 0, group=None) -> dict | None:
    """Compact game records of every rank (SelfPlayEngine.records()) on `dst`, in game-id order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return records
    plies = _gather_tensor(torch.from_numpy(np.ascontiguousarray(records["plies"])), dst, group)
    raw = np.ascontiguousarray(records["records"])
    as_bytes = torch.from_numpy(raw.view(np.uint8).reshape(raw.shape[0], -1))
    rec = _gather_tensor(as_bytes, dst, group)
    if dist.get_rank(group) != dst:
        return None
    return {"plies": plies.numpy(), "records": rec.numpy().view(raw.dtype).reshape(-1, raw.shape[1])}


def reduce_counters(counters: dict, group=None) -> dict:
    """Sum the additive engine counters over ranks (every rank gets the totals)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return dict(counters)
    keys = [k for k, v in counters.items() if isinstance(v, int)]
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    out = dict(counters)
    out.update({k: int(v) for k, v in zip(keys, t.tolist())})
    return out
