#!/usr/bin/env python
"""The reference's own `_self_play` (train.py:404-452, byte-for-byte copy under oracle/_ref/python) driving the drop-in: one
game at a time, 800 simulations per move, exactly what `othello-train` does between two training epochs.  Reports seconds
per game and where a ply's wall time goes (every MCTS method the loop calls is timed from outside; the loop is untouched).

  python tools/unmodified_self_play.py [--games 3] [--binding ctypes|cpp]
"""
import argparse
import json
import sys
import time
from argparse import Namespace
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=3)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--binding", default="ctypes", choices=["ctypes", "cpp"])
    a = ap.parse_args()
    sys.path.insert(0, str(ROOT / "oracle" / "_ref" / "python"))
    sys.path.insert(0, str(ROOT / "othello-alphazero_b200" / "cpp_binding") if a.binding == "cpp" else str(ROOT / "othello-alphazero_b200"))
    import numpy as np
    import torch
    import othello_mcts
    from othello_alphazero import train as ref_train
    from othello_alphazero.neural_net import AlphaZeroNet

    torch.manual_seed(0)
    np.random.seed(0)
    net = AlphaZeroNet(in_channels=17, num_squares=64, num_actions=65, conv_channels=128, num_residual_blocks=9,
                       value_head_hidden_channels=128).to("cuda").eval()
    mcts = ref_train.MCTS(history_size=8, torch_device="cuda", torch_pin_memory=False, num_simulations=a.sims, num_threads=2,
                          batch_size=16, c_puct_base=20000.0, c_puct_init=2.5, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
    spent, calls = defaultdict(float), defaultdict(int)

    def timed(name):
        inner = getattr(mcts, name)

        def f(*args, **kw):
            t0 = time.perf_counter()
            r = inner(*args, **kw)
            spent[name] += time.perf_counter() - t0
            calls[name] += 1
            return r
        return f
    if a.binding == "ctypes":                      # the compiled binding's methods are read-only attributes
        for name in ("search", "visit_counts", "self_play_data", "apply_action", "position", "reset_position"):
            setattr(mcts, name, timed(name))
    ref_train._self_play(mcts, net, Namespace(self_play_temperature=1.0))     # warm-up game (weights upload, graph capture)
    spent.clear(); calls.clear()
    plies = 0
    t0 = time.perf_counter()
    for _ in range(a.games):
        data = ref_train._self_play(mcts, net, Namespace(self_play_temperature=1.0))
        plies += len(data["features"]) // 8
    wall = time.perf_counter() - t0
    rec = {"what": "unmodified reference _self_play (train.py:404-452) on the drop-in, 128x10b, one game at a time", "binding": a.binding,
           "module": othello_mcts.__file__, "games": a.games, "sims_per_move": a.sims, "plies": plies, "s_per_game": wall / a.games,
           "ms_per_ply": 1e3 * wall / plies, "games_per_hour": 3600 * a.games / wall,
           "ms_per_ply_by_call": {k: 1e3 * v / plies for k, v in spent.items()}, "calls_per_ply": {k: calls[k] / plies for k in calls}}
    rec["ms_per_ply_python_loop_and_numpy"] = rec["ms_per_ply"] - sum(rec["ms_per_ply_by_call"].values())
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
