#!/usr/bin/env python
"""Games partitioned over the GPUs of one box, final NCCL gather, checked against one GPU owning every game.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_selfplay.py [--games 64]

Every rank plays its contiguous block of game ids to the end (no collective on the hot path), then
`othello_mcts.parallel.gather_records / gather_samples / reduce_counters` bring records, training tensors and counters
to rank 0 over NCCL.  Rank 0 then plays all games alone and compares: records, samples and simulation totals must be
identical (random streams are keyed by game id).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import othello_mcts as om  # noqa: E402
from othello_mcts import parallel  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=64)
    ap.add_argument("--sims", type=int, default=128)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=64).eval()
    kw = dict(history_size=4, num_simulations=a.sims, num_threads=2, batch_size=8, seed=77, cuda_device=local)
    first, n = parallel.partition_games(a.games, world, rank)
    eng = om.SelfPlayEngine(n, first_game_id=first, **kw)
    t0 = time.perf_counter()
    rec = eng.run(net)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    samples = eng.training_samples()
    all_rec = parallel.gather_records(rec)              # compact: 176 bytes per ply
    all_samples = parallel.gather_samples(samples)      # expanded: 4 612 bytes per sample (kept for comparison)
    ctr = parallel.reduce_counters(eng.counters())
    parallel.broadcast_weights(net if rank == 0 else None, eng, src=0)      # next iteration's weights: one NCCL broadcast of the folded blob
    if rank == 0:
        expanded = parallel.samples_from_records(all_rec, history_size=4, cuda_device=local)   # records expanded on the destination
        solo = om.SelfPlayEngine(a.games, first_game_id=0, **kw)
        want = solo.run(net)
        want_s = solo.training_samples()
        ok_rec = bool(np.array_equal(all_rec["plies"], want["plies"]) and all_rec["records"].tobytes() == want["records"].tobytes())
        ok_samples = all(torch.equal(all_samples[k], want_s[k]) and torch.equal(expanded[k], want_s[k]) for k in want_s)
        ok_ctr = ctr["simulations"] == solo.counters()["simulations"] and ctr["nodes_expanded"] == solo.counters()["nodes_expanded"]
        print(json.dumps({"world_size": world, "games": a.games, "games_per_rank": n, "rank0_play_s": dt, "samples": int(want_s["values"].shape[0]),
                          "gathered_record_bytes": int(all_rec["records"].nbytes), "gathered_sample_bytes": int(sum(t.numel() * 4 for t in all_samples.values())),
                          "records_identical": ok_rec, "samples_identical": bool(ok_samples), "counters_identical": bool(ok_ctr)}))
        assert ok_rec and ok_samples and ok_ctr
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
