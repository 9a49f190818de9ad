"""GPU: the reference's OWN Python callers, unmodified, running against the drop-in `othello_mcts`.

oracle/build_oracle.py places the reference's Python package (python/othello_alphazero/*.py, byte for byte) under
oracle/_ref/python/ (git-ignored, shipped to the GPU box).  With the drop-in first on sys.path, `from othello_mcts import
MCTS` inside train.py:11 / player.py:12 resolves to this repo's module, so these tests execute
  * `_self_play(mcts, neural_net, args)`           train.py:404-452
  * `AlphaZeroPlayer(...)`, `.get_action()`, `play_game(...)`   player.py:33-96, 177-259
exactly as `othello-train` / `othello-play` would, with the reference's own AlphaZeroNet class as the network (the
engine recognises it and evaluates it on the native tcgen05 path).
"""
import json
import sys
from argparse import Namespace
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
REF_PY = ROOT / "oracle" / "_ref" / "python"


@pytest.fixture(scope="module")
def ref_pkg():
    if not (REF_PY / "othello_alphazero" / "train.py").exists():
        pytest.skip("oracle/_ref/python not built (python oracle/build_oracle.py needs /root/reference)")
    import othello_mcts                      # the drop-in (conftest puts othello-alphazero_b200/ first on sys.path)
    assert "othello-alphazero_b200" in othello_mcts.__file__
    if str(REF_PY) not in sys.path:
        sys.path.append(str(REF_PY))
    from othello_alphazero import neural_net, player, train
    assert train.MCTS is othello_mcts.MCTS and player.MCTS is othello_mcts.MCTS
    return Namespace(train=train, player=player, neural_net=neural_net, om=othello_mcts)


NET_KW = dict(in_channels=9, num_squares=64, num_actions=65, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32)


def _planes_to_mask(plane: torch.Tensor) -> int:
    return int(sum(1 << (63 - i) for i, v in enumerate(plane.flatten().tolist()) if v != 0.0))


def test_unmodified_self_play_runs_on_the_dropin(ref_pkg):
    """train.py:404-452 verbatim.  The returned samples must describe one legal game from the initial position."""
    om = ref_pkg.om
    torch.manual_seed(3)
    net = ref_pkg.neural_net.AlphaZeroNet(**NET_KW).to("cuda").eval()
    config = {"history_size": 4, "torch_device": "cuda", "torch_pin_memory": False, "num_simulations": 64, "num_threads": 2,
              "batch_size": 8, "c_puct_base": 20000.0, "c_puct_init": 2.5, "dirichlet_epsilon": 0.25, "dirichlet_alpha": 0.5}
    mcts = ref_pkg.train.MCTS(**config)                                  # train.py:250: MCTS(**config["mcts"]), ten keys
    np.random.seed(0)
    data = ref_pkg.train._self_play(mcts, net, Namespace(self_play_temperature=1.0))
    f, p, v = data["features"], data["policies"], data["values"]
    assert len(f) == len(p) == len(v) and len(f) % 8 == 0 and len(f) >= 8 * 50
    assert mcts.counters()["evaluations"] > 0 and mcts._evaluator == 3   # the native ResNet evaluated the leaves
    plies = len(f) // 8
    pos = om.Position.initial_position()
    for k in range(plies):
        x, pi = f[8 * k], p[8 * k]                                       # symmetry 0 = identity
        assert x.shape == (9, 8, 8) and x.dtype == torch.float32 and pi.shape == (65,)
        assert float(x[0, 0, 0]) == pos.player() - 1                     # plane 0: side to move
        assert _planes_to_mask(x[1]) == pos.player1_discs() and _planes_to_mask(x[2]) == pos.player2_discs()
        assert abs(float(pi.sum()) - 1.0) < 1e-5
        legal = pos.legal_actions()
        assert all(float(pi[a]) == 0.0 for a in range(65) if a not in legal)
        for t in range(1, 8):                                            # the other 7 symmetries hold the same mass
            assert abs(float(p[8 * k + t].sum()) - 1.0) < 1e-5 and float(f[8 * k + t].sum()) == float(x.sum())
        # the action played leads to the next recorded position
        if k + 1 < plies:
            nxt = f[8 * (k + 1)]
            want = (_planes_to_mask(nxt[1]), _planes_to_mask(nxt[2]))
            moved = [a for a in legal if (pos.apply_action(a).player1_discs(), pos.apply_action(a).player2_discs()) == want]
            assert len(moved) >= 1 and any(float(pi[a]) > 0.0 for a in moved), k
            pos = pos.apply_action(moved[0])
    # the last recorded position is one move from the end; values alternate sign every ply (train.py:447-450)
    finals = [pos.apply_action(a) for a in pos.legal_actions()]
    assert any(q.is_terminal() for q in finals)
    z = float(v[0])
    assert z in (-1.0, 0.0, 1.0)
    for k in range(plies):
        assert all(float(v[8 * k + t]) == (z if k % 2 == 0 else -z) for t in range(8))


def test_unmodified_alphazero_player_and_play_game(ref_pkg, tmp_path):
    """player.py:177-259 (AlphaZeroPlayer: checkpoint loading, MCTS with dirichlet_epsilon = 0, get_action) and
    play_game (player.py:33-96) verbatim, two players with different checkpoints."""
    dirs = []
    for i in range(2):
        torch.manual_seed(10 + i)
        net = ref_pkg.neural_net.AlphaZeroNet(**NET_KW)
        d = tmp_path / f"ckpt{i}"
        d.mkdir()
        (d / "config.json").write_text(json.dumps({"neural_net": NET_KW}))
        torch.save(net.state_dict(), d / "neural_net.pth")                # train.py:316-320
        dirs.append(d)
    mk = lambda d: ref_pkg.player.AlphaZeroPlayer(device="cuda", pin_memory=False, num_simulations=48, num_threads=1, batch_size=8,
                                                  c_puct_base=20000.0, c_puct_init=2.5, checkpoint_dir=d, compile_neural_net=False,
                                                  compile_neural_net_backend="inductor", compile_neural_net_mode="default", quiet=True)
    p1, p2 = mk(dirs[0]), mk(dirs[1])
    assert p1.mcts.dirichlet_epsilon() == 0.0 and p1.mcts.history_size() == 4
    # get_action = search + most visited action (player.py:248-259)
    np.random.seed(1)
    a = p1.get_action()
    vc = p1.mcts.visit_counts()
    assert sum(vc) == 48 - 8 and a in [p1.mcts.position().legal_actions()[i] for i, c in enumerate(vc) if c == max(vc)]
    result = ref_pkg.player.play_game(p1, p2, quiet=True)
    assert result in (0, 1, 2)
    assert p1.mcts.position().is_terminal() and p2.mcts.position().is_terminal()
    assert p1.mcts.position().player1_discs() == p2.mcts.position().player1_discs()
    b, w = bin(p1.mcts.position().player1_discs()).count("1"), bin(p1.mcts.position().player2_discs()).count("1")
    assert result == (1 if b > w else 2 if w > b else 0)
    assert p1.mcts.counters()["evaluations"] > 0 and p2.mcts.counters()["evaluations"] > 0
    assert p1.mcts._evaluator == 3 and p2.mcts._evaluator == 3            # both ran on the native evaluator


@pytest.mark.parametrize("binding", ["ctypes", "pybind11"])
def test_unmodified_cli_entry_points_train_resume_play(ref_pkg, tmp_path, binding):
    """The north-star sentence "so `othello-train` and the play CLI drive it unchanged", executed: the reference's
    `othello_alphazero.train.main()` (one tiny iteration from scratch: self-play through `othello_mcts.MCTS`, training,
    checkpoint), again with `--from-checkpoint` (train.py:316-350: `MCTS(**config["mcts"])` from the saved config), then
    `othello_alphazero.play.main()` with an AlphaZero player loaded from that checkpoint against the greedy player -- in
    fresh interpreters with `othello_mcts` resolved to the drop-in (ctypes package / compiled pybind11 module)."""
    import os
    import subprocess
    pkg_root = ROOT / "othello-alphazero_b200"
    if binding == "pybind11":
        if not list((pkg_root / "cpp_binding" / "othello_mcts").glob("_othello_mcts_impl*.so")):
            pytest.skip("pybind11 binding not built")
        pkg_root = pkg_root / "cpp_binding"
    env = dict(os.environ, PYTHONPATH=f"{pkg_root}:{REF_PY}")
    out = tmp_path / "ckpt"
    common = ["--device", "cuda", "--iterations", "1", "--self-play-games-per-iteration", "2", "--history-size", "4",
              "--neural-net-conv-channels", "128", "--neural-net-residual-blocks", "2", "--neural-net-value-head-hidden-channels", "32",
              "--mcts-simulations", "32", "--mcts-threads", "1", "--mcts-batch-size", "8", "--training-batch-size", "64",
              "--training-dataloader-workers", "0", "--output-dir", str(out)]

    def run(module, argv):
        code = (f"import sys, othello_mcts; assert r'{pkg_root}' in othello_mcts.__file__, othello_mcts.__file__; "
                f"from othello_alphazero import {module}; sys.argv = ['othello-{module}'] + {argv!r}; {module}.main()")
        return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)

    r = run("train", common)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("config.json", "neural_net.pth", "optimizer.pth", "lr_scheduler.pth", "stats.json", "iteration.txt"):
        assert (out / "001" / name).exists(), name
    stats = json.loads((out / "001" / "stats.json").read_text())
    assert all(np.isfinite(stats[k]) for k in ("total_loss", "policy_loss", "value_loss", "l2_loss"))
    cfg = json.loads((out / "001" / "config.json").read_text())
    assert cfg["mcts"]["num_simulations"] == 32 and cfg["neural_net"]["in_channels"] == 9

    r = run("train", common + ["--from-checkpoint", str(out / "001")])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (out / "002" / "neural_net.pth").exists()

    r = run("play", ["--player1", "alphazero", "--player2", "greedy", "--alphazero-device", "cuda", "--alphazero-checkpoint", str(out / "002"),
                     "--alphazero-simulations", "32", "--alphazero-threads", "1", "--alphazero-batch-size", "8"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Game over" in r.stdout and ("Black wins" in r.stdout or "White wins" in r.stdout or "Draw" in r.stdout)
    assert "  a b c d e f g h" in r.stdout and "Action-value:" in r.stdout          # Position.__str__ and mean_action_values() in use
