"""Generate tests/golden/golden_stats.json: DISTRIBUTIONS produced by the unmodified reference (oracle/_ref/othello_mcts,
the reference's pybind11 module compiled by oracle/build_oracle.py) for the parts whose random streams cannot be
replayed bit for bit (private mt19937 per search, search_thread.cpp:22; np.random in train.py:421-430).

  * root visit counts of 2 048 searches of one position (200 sims, T=1 x B=16, eps 0.25, alpha 0.5);
  * 1 024 self-play games driven exactly like train.py:404-452 (`_self_play`: sample from the visit counts for the
    first 12 plies, then a random argmax) at 64 sims/move: first move, game length in plies, final disc difference.

The evaluator is the exact synthetic net of SURVEY.md section 8(c) (D4-equivariant, exactly representable values), so the
GPU engine's built-in 'synth' evaluator sees the same numbers.  Run in the build container (about 6 minutes):
  python tests/golden/generate_golden_stats.py
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import othello_mcts as ref  # noqa: E402  (the reference extension)


def synth_net(x: torch.Tensor):
    occ = x[:, 1] + x[:, 2]
    box = torch.nn.functional.conv2d(occ[:, None], torch.ones(1, 1, 3, 3), padding=1)[:, 0]
    w = torch.cat([1.0 + box.flatten(1), torch.ones(x.shape[0], 1)], dim=1)
    value = (1.0 - 2.0 * x[:, 0, 0, 0]) * (x[:, 1].sum(dim=(1, 2)) - x[:, 2].sum(dim=(1, 2))) / 64.0
    return {"policy": (w / w.sum(dim=1, keepdim=True)).clone(), "value": value.clone()}


def root_visits(n=2048):
    kw = dict(history_size=4, num_simulations=200, num_threads=1, batch_size=16, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
    m = ref.MCTS(torch_device="cpu", **kw)
    rows = []
    for _ in range(n):
        m.reset_position()
        for a in (19, 18):
            m.apply_action(a)
        m.search(synth_net)
        rows.append(m.visit_counts())
    return {"kwargs": kw, "opening": [19, 18], "visit_counts": rows}


def self_play(n=1024):
    kw = dict(history_size=4, num_simulations=64, num_threads=1, batch_size=8, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
    m = ref.MCTS(torch_device="cpu", **kw)
    np.random.seed(2026)
    first, plies, diff = [], [], []
    for _ in range(n):
        m.reset_position()                                   # train.py:410
        t = 0
        while True:
            position = m.position()
            if position.is_terminal():
                break
            m.search(synth_net)
            vc = np.array(m.visit_counts(), dtype=np.float32)
            if t < 12:                                       # train.py:422-426 (temperature 1)
                p = vc / vc.sum()
                action = int(np.random.choice(position.legal_actions(), p=p))
            else:                                            # train.py:427-430
                best = np.where(vc == vc.max())[0]
                action = position.legal_actions()[np.random.choice(best)]
            if t == 0:
                first.append(action)
            m.apply_action(action)
            t += 1
        plies.append(t)
        diff.append(bin(position.player1_discs()).count("1") - bin(position.player2_discs()).count("1"))
    return {"kwargs": kw, "temperature_plies": 12, "first_move": first, "plies": plies, "disc_diff": diff}


def main():
    torch.set_num_threads(1)
    out = {"source": "unmodified reference module (oracle/_ref/othello_mcts) + exact synthetic net", "root": root_visits(), "games": self_play()}
    path = Path(__file__).with_name("golden_stats.json")
    path.write_text(json.dumps(out, separators=(",", ":")))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
