#!/usr/bin/env python
"""A whole self-play iteration on one GPU (BASELINE.json config: 128x10b, N games, 800 sims/move), start to finish:
every game is played to its end, records are kept, the training tensors `_self_play` would return (train.py:432-450)
are emitted, and the whole-run rates are printed as one JSON line.

  python tools/selfplay_run.py [--games 1000] [--slots 1000] [--sims 800] [--channels 128 --blocks 9]

--slots < --games uses the rolling game pool (finished games leave, their slots restart with unplayed ids)."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import othello_mcts as om  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1000)
    ap.add_argument("--slots", type=int, default=1000)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--samples", action="store_true", help="also materialise the training tensors on the host")
    a = ap.parse_args()
    torch.manual_seed(0)
    net = AlphaZeroNet(17, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval()
    eng = om.SelfPlayEngine(a.slots, history_size=8, num_simulations=a.sims, num_threads=a.threads, batch_size=a.batch, seed=a.seed)
    eng.prepare(net)
    eng.play_ply()                                   # warm-up (weights upload, first launches), then start over
    rolling = a.slots < a.games
    if rolling:
        eng.set_game_pool(a.games)
    else:
        eng.reset()
    torch.cuda.synchronize()
    c0 = eng.counters()
    t0 = time.perf_counter()
    plies = 0
    active_trace = []
    while True:
        _, _, unfinished = eng.play_ply()
        plies += 1
        active_trace.append(unfinished)
        if unfinished == 0:
            break
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = eng.counters()
    if rolling:
        rec = eng.finished_records()
        game_plies = rec["plies"]
    else:
        rec = eng.records()
        game_plies = rec["plies"]
    out = {"games": int(len(game_plies)), "slots": a.slots, "sims_per_move": a.sims, "network": f"{a.channels}x{a.blocks + 1}b",
           "wall_s": dt, "engine_plies": plies, "mean_game_plies": float(np.mean(game_plies)), "min_game_plies": int(np.min(game_plies)),
           "max_game_plies": int(np.max(game_plies)), "simulations": c1["simulations"] - c0["simulations"],
           "sims_per_s": (c1["simulations"] - c0["simulations"]) / dt, "evaluations": c1["evaluations"] - c0["evaluations"],
           "slot_utilisation": (c1["simulations"] - c0["simulations"]) / max(1, c1["evaluations"] - c0["evaluations"]),
           "games_per_hour": len(game_plies) / dt * 3600.0, "nodes_expanded": c1["nodes_expanded"] - c0["nodes_expanded"],
           "gpu_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "active_after_ply": active_trace[::8]}
    free, total = torch.cuda.mem_get_info()
    out["gpu_mem_used_gb"] = (total - free) / 1e9
    if a.samples:
        t1 = time.perf_counter()
        s = eng.training_samples()
        out["training_samples"] = int(s["values"].shape[0])
        out["samples_emit_s"] = time.perf_counter() - t1
        out["value_mean"] = float(s["values"].mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
