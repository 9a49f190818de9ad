#!/usr/bin/env python
"""An evaluation match on one GPU: two random-init networks of the same shape play `--pairs` pairs of games (both
colours) in lockstep, as `evaluation.play_games` (evaluation.py:15-90) would one game at a time.  Prints one JSON line.

  python tools/match_run.py [--pairs 500] [--sims 800] [--channels 128 --blocks 9]"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import torch  # noqa: E402

import othello_mcts as om  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=500)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    nets = []
    for s in (0, 1):
        torch.manual_seed(s)
        nets.append(AlphaZeroNet(17, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval())
    m = om.MatchEngine(a.pairs, history_size=8, num_simulations=a.sims, num_threads=a.threads, batch_size=a.batch, seed=a.seed)
    ea, eb = m.engines()
    ea.prepare(nets[0]); eb.prepare(nets[1])
    m.play_ply(); m.play_ply()                          # warm-up, then start over
    torch.cuda.synchronize()
    c0 = [e.counters() for e in (ea, eb)]
    t0 = time.perf_counter()
    out = m.play(nets[0], nets[1])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = [e.counters() for e in (ea, eb)]
    sims = sum(y["simulations"] - x["simulations"] for x, y in zip(c0, c1))
    evals = sum(y["evaluations"] - x["evaluations"] for x, y in zip(c0, c1))
    res = out["result"]
    print(json.dumps({"games": int(len(res)), "pairs": a.pairs, "sims_per_move": a.sims, "network": f"{a.channels}x{a.blocks + 1}b",
                      "wall_s": dt, "lockstep_plies": out["plies"], "simulations": sims, "sims_per_s": sims / dt, "evaluations": evals,
                      "games_per_hour": len(res) / dt * 3600.0, "black_wins": int((res == 1).sum()), "white_wins": int((res == 2).sum()),
                      "draws": int((res == 0).sum()), "score_a": out["score_a"], "gpu_mem_gb_used": (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9}))


if __name__ == "__main__":
    main()
