"""CPU, world_size 2 over gloo: the multi-GPU host logic (game partition by id, final gather of records and
training samples, counter reduction).  The hot path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from othello_mcts import parallel

REC = np.dtype([("p1", "<u8"), ("p2", "<u8"), ("player", "<i4"), ("action", "<i4"), ("n_children", "<i4"),
                ("visit_counts", "<i4", (36,)), ("_pad", "<i4")])


def test_partition_covers_every_game_once():
    for total in (0, 1, 7, 1000, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, n = parallel.partition_games(total, world, r)
                seen += list(range(first, first + n))
                for g in range(first, first + n):
                    assert parallel.owner_of(g, total, world) == r
            assert seen == list(range(total))
    with pytest.raises(ValueError):
        parallel.partition_games(10, 2, 2)


def _fake_rank_output(first, n):
    """Deterministic per-game stand-ins for what a rank's engine returns, keyed by GAME ID only."""
    plies = np.array([40 + (g * 7) % 21 for g in range(first, first + n)], np.int32)
    rec = np.zeros((n, 128), REC)
    for i, g in enumerate(range(first, first + n)):
        rec[i, : plies[i]]["action"] = (g + np.arange(plies[i])) % 65
        rec[i, : plies[i]]["p1"] = np.uint64(g) << np.uint64(8)
    rows = int(plies.sum()) * 8
    gid = np.repeat(np.arange(first, first + n), plies * 8).astype(np.float32)
    samples = {"features": torch.from_numpy(gid)[:, None, None, None].expand(rows, 3, 8, 8).contiguous(),
               "policies": torch.from_numpy(gid)[:, None].expand(rows, 65).contiguous(), "values": torch.from_numpy(gid)}
    return {"plies": plies, "records": rec}, samples


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, n = parallel.partition_games(total, world, rank)
        records, samples = _fake_rank_output(first, n)
        got_r = parallel.gather_records(records)
        got_s = parallel.gather_samples(samples)
        ctr = parallel.reduce_counters({"simulations": 800 * n, "plies": n, "eval_ms": 1.5})
        if rank == 0:
            want_r, want_s = _fake_rank_output(0, total)          # what a single rank owning every game returns
            ok = (np.array_equal(got_r["plies"], want_r["plies"]) and got_r["records"].tobytes() == want_r["records"].tobytes()
                  and all(torch.equal(got_s[k], want_s[k]) for k in want_s)
                  and ctr["simulations"] == 800 * total and ctr["plies"] == total)
            q.put(bool(ok))
        else:
            assert got_r is None and got_s is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 16])
def test_gather_is_invariant_to_world_size(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


class _FakeEngine:
    def __init__(self):
        self.loaded = None

    def load_folded(self, shape, blob):
        self.loaded = ((shape.in_planes, shape.channels, shape.blocks, shape.value_hidden), bytes(blob))


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from othello_mcts import weights
        from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm
        net = None
        if rank == 0:                                      # only the training rank owns the new weights
            torch.manual_seed(3)
            net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32).eval()
            randomize_batchnorm(net, 4)
        eng = _FakeEngine()
        parallel.broadcast_weights(net, eng, src=0)
        if rank == 0:
            shape, blob = weights.fold_alphazero_net(net)
            q.put(("want", (shape.in_planes, shape.channels, shape.blocks, shape.value_hidden), blob))
        q.put(("got", rank) + eng.loaded)
    finally:
        dist.destroy_process_group()


def test_broadcast_weights_delivers_the_folded_blob_to_every_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    items = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = [x for x in items if x[0] == "want"][0]
    got = [x for x in items if x[0] == "got"]
    assert len(got) == 2 and all(g[2] == want[1] and g[3] == want[2] for g in got)
