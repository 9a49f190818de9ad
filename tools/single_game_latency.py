#!/usr/bin/env python
"""Latency of the reference-shaped single-game path: `MCTS.search(net)` as `othello-train` / `othello-play` call it
(train.py:415-416, player.py:251: one game, 800 simulations, T = 2 x B = 16 leaves per network call).

  python tools/single_game_latency.py [--sims 800] [--searches 20] [--graphs 1]

Prints ms per search and simulations/s (reference accounting).  The reference's README quotes 28 k simulations/s for
this call on an RTX 4090."""
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
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--searches", type=int, default=20)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    a = ap.parse_args()
    torch.manual_seed(0)
    net = AlphaZeroNet(17, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval().cuda()
    m = om.MCTS(history_size=8, torch_device="cuda", num_simulations=a.sims, num_threads=a.threads, batch_size=a.batch)
    m.set_graphs(a.graphs)
    inflight = a.threads * a.batch
    sims_per_search = -(-a.sims // inflight) * inflight
    for _ in range(3):                                   # warm-up: weights upload, eager search, graph capture
        m.search(net)
    times, applies = [], []
    for _ in range(a.searches):
        if m.position().is_terminal():
            m.reset_position()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.search(net)                                    # blocking, as in the reference
        times.append(time.perf_counter() - t0)
        counts = m.visit_counts()
        act = m.position().legal_actions()[int(np.argmax(counts))]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        m.apply_action(act)                              # re-root (mcts.cpp:114-165), blocking
        torch.cuda.synchronize()
        applies.append(time.perf_counter() - t1)
    ms = 1e3 * float(np.median(times))
    print(json.dumps({"network": f"{a.channels}x{a.blocks + 1}b", "sims_per_search": sims_per_search, "leaves_per_network_call": inflight,
                      "graphs": a.graphs, "ms_per_search_median": ms, "ms_per_search_min": 1e3 * min(times), "ms_per_search_max": 1e3 * max(times),
                      "sims_per_s": sims_per_search / (ms * 1e-3), "ms_per_round": ms / (sims_per_search / inflight),
                      "ms_per_apply_action_median": 1e3 * float(np.median(applies))}))


if __name__ == "__main__":
    main()
