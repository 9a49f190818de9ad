"""Generate tests/golden/golden.json from the UNMODIFIED reference (needs /root/reference).

Everything here is produced by the reference's own code: the rules/encoding through
oracle/_ref/libref_shim.so (the reference headers behind a C ABI) and the search through the reference
pybind11 module built by oracle/build_oracle.py.  The oracle is not involved.  Run in the build
container:  python tests/golden/generate_golden.py
"""
import ctypes as C
import json
import random
import struct
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import build_oracle  # noqa: E402
import pyoracle as po  # noqa: E402  (only for the Pos struct and the shim loader)

build_oracle.build_all()
R = po.ref_shim()
import othello_mcts as ref  # noqa: E402  (oracle/_ref/othello_mcts = the reference extension)


def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", x))[0]


def pos_tuple(p):
    return [p.player, p.p1, p.p2, p.legal, p.next_legal]


def gen_rules(n_games=120, seed=7):
    rng = random.Random(seed)
    rows = []
    for _ in range(n_games):
        p = po.Pos()
        R.ref_initial_position(C.byref(p))
        while p.player != 0:
            acts = (C.c_int * 40)()
            n = R.ref_legal_actions(C.byref(p), acts)
            a = rng.choice(list(acts[:n]))
            q = po.Pos()
            R.ref_apply_action(C.byref(p), a, C.byref(q))
            me, opp = (p.p1, p.p2) if p.player == 1 else (p.p2, p.p1)
            fl = R.ref_get_flips(1 << (63 - a), me, opp) if a != 64 else 0
            rows.append({"pos": pos_tuple(p), "actions": list(acts[:n]), "action": a, "flips": fl,
                         "legal_of_me": R.ref_get_legal_moves(me, opp), "next": pos_tuple(q)})
            p = q
    return rows


def gen_features(seed=11):
    rng = random.Random(seed)
    out = []
    for H in (1, 4, 8):
        for _ in range(6):
            # a random game prefix gives a real history chain
            hist = []
            p = po.Pos()
            R.ref_initial_position(C.byref(p))
            plies = rng.randint(0, 30)
            hist.append(pos_tuple(p))
            for _ in range(plies):
                if p.player == 0:
                    break
                acts = (C.c_int * 40)()
                n = R.ref_legal_actions(C.byref(p), acts)
                q = po.Pos()
                R.ref_apply_action(C.byref(p), rng.choice(list(acts[:n])), C.byref(q))
                p = q
                if p.player == 0:
                    break
                hist.append(pos_tuple(p))
            hist = hist[::-1][:H]                      # newest first
            arr = (po.Pos * len(hist))()
            for i, h in enumerate(hist):
                arr[i].player, arr[i].p1, arr[i].p2, arr[i].legal, arr[i].next_legal = h
            for t in range(8):
                f = np.zeros((1 + 2 * H) * 64, np.float32)
                R.ref_positions_to_features(arr, len(hist), f.ctypes.data_as(C.POINTER(C.c_float)), H, t)
                planes = f.reshape(1 + 2 * H, 64)
                assert set(np.unique(planes)) <= {0.0, 1.0}
                masks = [int(sum(1 << (63 - i) for i in range(64) if pl[i] != 0)) for pl in planes]
                out.append({"H": H, "t": t, "history": hist, "plane_masks": masks})
    return out


def synth_net(x: torch.Tensor):
    """SURVEY.md 8(c) exact synthetic net in plain torch ops (every value exactly representable)."""
    occ = x[:, 1] + x[:, 2]
    box = torch.nn.functional.conv2d(occ[:, None], torch.ones(1, 1, 3, 3), padding=1)[:, 0]
    w = torch.cat([1.0 + box.flatten(1), torch.ones(x.shape[0], 1)], dim=1)
    policy = w / w.sum(dim=1, keepdim=True)
    value = (1.0 - 2.0 * x[:, 0, 0, 0]) * (x[:, 1].sum(dim=(1, 2)) - x[:, 2].sum(dim=(1, 2))) / 64.0
    return {"policy": policy.clone(), "value": value.clone()}


def gen_search():
    games = []
    for (H, N, T, B) in [(4, 200, 1, 16), (8, 100, 1, 1), (2, 300, 1, 64), (8, 800, 1, 32)]:
        m = ref.MCTS(history_size=H, num_simulations=N, num_threads=T, batch_size=B, dirichlet_epsilon=0.0)
        plies = []
        while not m.position().is_terminal():
            m.search(synth_net)
            vc = m.visit_counts()
            q = [f32_bits(v) for v in m.mean_action_values()]
            spd = m.self_play_data()
            pol = [[f32_bits(float(v)) for v in spd["policy"][t].tolist()] for t in (0, 3)]
            a = m.position().legal_actions()[int(np.argmax(vc))]   # first argmax
            plies.append({"visit_counts": vc, "q_bits": q, "action": a, "policy_bits_t0_t3": pol})
            m.apply_action(a)
            if len(games) == 3 and len(plies) >= 6:                 # the 800-sim game: first plies only
                break
        pos = m.position()
        games.append({"H": H, "N": N, "T": T, "B": B, "plies": plies,
                      "final": [pos.player(), pos.player1_discs(), pos.player2_discs()]})
    return games


def main():
    start = po.Pos()
    R.ref_initial_position(C.byref(start))
    golden = {
        "source": "unmodified reference (position.h, transformation.h, mcts.cpp, search_thread.cpp)",
        "perft": [int(R.ref_perft(C.byref(start), d)) for d in range(1, 11)],
        "d4": [[R.ref_transform_action(a, t) for a in range(65)] for t in range(8)],
        "rules": gen_rules(),
        "features": gen_features(),
        "search": gen_search(),
    }
    out = Path(__file__).with_name("golden.json")
    out.write_text(json.dumps(golden, separators=(",", ":")))
    print("wrote", out, out.stat().st_size, "bytes;", len(golden["rules"]), "rule rows")


if __name__ == "__main__":
    main()
