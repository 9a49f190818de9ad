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


def samples_from_records(records: dict, history_size: int, cuda_device: int = 0, max_plies: int | None = None) -> dict:
    """Expand compact game records (SelfPlayEngine.records() / gather_records(): 176 bytes per ply) into the training
    tensors `_self_play` returns (train.py:432-450: 8 symmetries per ply, outcome from the final discs) on THIS rank's
    GPU.  A multi-GPU run gathers the records and calls this on the destination instead of shipping the expanded
    samples (4 612 bytes each, about 2.2 GB per 1 000 games)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    plies = np.ascontiguousarray(records["plies"], dtype=np.int32)
    rec = np.ascontiguousarray(records["records"])
    n_games = int(plies.shape[0])
    mp = int(rec.shape[1]) if max_plies is None else int(max_plies)
    n = C.c_size_t()
    _lib.check(lib.oz_training_samples_from_records(cuda_device, rec.ctypes.data_as(C.c_void_p), plies.ctypes.data_as(C.c_void_p), n_games, mp,
                                                    history_size, None, None, None, 0, C.byref(n)))
    planes = 1 + 2 * history_size
    f = np.zeros((n.value, planes, 8, 8), np.float32)
    p = np.zeros((n.value, 65), np.float32)
    v = np.zeros((n.value,), np.float32)
    if n.value:
        _lib.check(lib.oz_training_samples_from_records(cuda_device, rec.ctypes.data_as(C.c_void_p), plies.ctypes.data_as(C.c_void_p), n_games, mp,
                                                        history_size, f.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                                        v.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
    return {"features": torch.from_numpy(f), "policies": torch.from_numpy(p), "values": torch.from_numpy(v)}


def broadcast_weights(net, engines, src: int = 0, group=None) -> None:
    """New weights to every rank's engine(s) between training iterations: rank `src` folds `net`
    (BatchNorm into bf16 convolution weights, weights.fold_alphazero_net) and ONE broadcast moves the folded blob
    (5.4 MB for 128x10b, half of the fp32 state_dict); every rank uploads it into `engines` (a SelfPlayEngine / MCTS or
    a list of them).  On other ranks `net` may be None.  Without a process group: a plain local upload."""
    import ctypes as C
    import torch.distributed as dist
    from . import weights
    from ._lib import OzResnetShape, check
    engines = engines if isinstance(engines, (list, tuple)) else [engines]
    active = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if active else src
    shape_t = torch.zeros(5, dtype=torch.int64)
    blob_t = None
    if rank == src:
        shape, blob = weights.fold_cached(net)
        shape_t = torch.tensor([shape.in_planes, shape.channels, shape.blocks, shape.value_hidden, len(blob)], dtype=torch.int64)
        blob_t = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if active:
        on_gpu = dist.get_backend(group) == "nccl"
        if on_gpu:
            shape_t = shape_t.cuda()
        dist.broadcast(shape_t, src=src, group=group)
        if blob_t is None:
            blob_t = torch.empty(int(shape_t[4].item()), dtype=torch.uint8)
        if on_gpu:
            blob_t = blob_t.cuda()
        dist.broadcast(blob_t, src=src, group=group)
        shape_t, blob_t = shape_t.cpu(), blob_t.cpu()
    shape = OzResnetShape(*[int(x) for x in shape_t[:4].tolist()])
    blob = blob_t.numpy().tobytes()
    for eng in engines:
        eng.load_folded(shape, blob)


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
