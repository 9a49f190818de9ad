"""The compiled pybind11 binding (csrc/othello_mcts_pybind.cpp -> cpp_binding/othello_mcts/_othello_mcts_impl*.so): the
reference's own binding file (cpp/src/lib/othello_mcts.cpp:49-151) with its bodies replaced by calls into the C ABI.
It must behave like the reference module and like the ctypes package next door: same surface, exception types and
message texts (CPU tests), and bit-identical searches against the reference goldens and against the ctypes binding --
including the AlphaZeroNet path, whose weights this binding folds in C++ from the LibTorch tensors (GPU tests).
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import f32_bits

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "othello-alphazero_b200" / "cpp_binding" / "othello_mcts"


@pytest.fixture(scope="module")
def cpp():
    if not list(PKG.glob("_othello_mcts_impl*.so")):
        pytest.skip("pybind11 binding not built (python othello-alphazero_b200/build_ext.py)")
    spec = importlib.util.spec_from_file_location("othello_mcts_cpp", PKG / "__init__.py", submodule_search_locations=[str(PKG)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["othello_mcts_cpp"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_exports_and_abi(cpp):
    from othello_mcts import _lib
    assert set(cpp.__all__) == {"MCTS", "Position", "get_flips", "get_legal_moves"}      # cpp/src/othello_mcts/__init__.py:6
    assert cpp._othello_mcts_impl.abi_version() == _lib.load().oz_abi_version()
    for name in ("reset_position", "position", "search", "visit_counts", "mean_action_values", "self_play_data",
                 "apply_action", "history_size", "set_history_size", "torch_device", "set_torch_device",
                 "num_simulations", "set_num_simulations", "num_threads", "set_num_threads", "batch_size",
                 "set_batch_size", "c_puct_base", "set_c_puct_base", "c_puct_init", "set_c_puct_init",
                 "dirichlet_epsilon", "set_dirichlet_epsilon", "dirichlet_alpha", "set_dirichlet_alpha"):
        assert callable(getattr(cpp.MCTS, name)), name
    import re
    doc = cpp.MCTS.__init__.__doc__                        # keyword names, order and defaults of othello_mcts.cpp:90-112
    for name, default in (("history_size", "4"), ("torch_device", "'cpu'"), ("torch_pin_memory", "False"), ("num_simulations", "800"),
                          ("num_threads", "2"), ("batch_size", "16"), ("c_puct_base", "20000.0"), ("c_puct_init", "2.5"),
                          ("dirichlet_epsilon", "0.25"), ("dirichlet_alpha", "0.5")):
        assert re.search(rf"\b{name}: [^,=]+ = {re.escape(default)}[,)]", doc), (name, doc)
    order = [doc.index(k) for k in ("history_size", "torch_device", "torch_pin_memory", "num_simulations", "num_threads",
                                    "batch_size", "c_puct_base", "c_puct_init", "dirichlet_epsilon", "dirichlet_alpha")]
    assert order == sorted(order)


def test_position_matches_golden_and_ctypes(cpp, golden):
    import othello_mcts as ct
    p = cpp.Position.initial_position()
    assert (p.player(), p.player1_discs(), p.player2_discs(), p.legal_moves()) == (1, 0x0000000810000000, 0x0000001008000000, 0x0000102004080000)
    assert p.legal_actions() == [19, 26, 37, 44] and str(p) == str(ct.Position.initial_position())
    assert cpp.get_legal_moves(0, 0) == 0
    pos, prev_next = None, None
    for row in golden["rules"][:2500]:
        if pos is None or row["pos"] != prev_next:
            pos = cpp.Position.initial_position()
        assert [pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()] == row["pos"][:4]
        assert pos.legal_actions() == row["actions"]
        me, opp = (row["pos"][1], row["pos"][2]) if row["pos"][0] == 1 else (row["pos"][2], row["pos"][1])
        assert cpp.get_legal_moves(me, opp) == row["legal_of_me"]
        if row["action"] != 64:
            assert cpp.get_flips(1 << (63 - row["action"]), me, opp) == row["flips"]
        pos = pos.apply_action(row["action"])
        prev_next = row["next"]
        assert [pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()] == row["next"][:4]
        if pos.is_terminal():
            assert pos.legal_actions() == []
            pos = None


def test_position_errors_match_reference_texts(cpp):
    p = cpp.Position.initial_position()
    assert [p[27], p[28], p[0]] == [2, 1, 0] and p.is_legal_move(19) and not p.is_legal_move(0)
    assert str(p).splitlines()[3] == "3 · · · × · · · ·"
    with pytest.raises(IndexError, match=r"Expected 0 <= index < 64, but got 64\."):
        p[64]
    with pytest.raises(IndexError, match=r"Expected 0 <= index < 64, but got -1\."):
        p.is_legal_move(-1)
    with pytest.raises(ValueError, match=r"^0 is not a legal action\.$"):
        p.apply_action(0)
    with pytest.raises(IndexError, match=r"Expected 0 <= action < 65, but got 65\."):
        p.apply_action(65)
    with pytest.raises(ValueError, match=r"Pass is not allowed when there are legal moves\."):
        p.apply_pass()
    with pytest.raises(ValueError, match=r"Expected a single bit in move_mask, but got 0b0{62}11\."):
        p.apply_move(3)
    q = p.apply_move(1 << (63 - 26))
    assert q.player() == 2 and q.legal_moves() == 0x0000280020000000 and p.player() == 1
    with pytest.raises(TypeError):
        cpp.Position()


@pytest.mark.parametrize("kwargs,msg", [
    (dict(history_size=0), "Expected history_size >= 1, but got 0."),
    (dict(num_threads=-2), "Expected num_threads >= 1, but got -2."),
    (dict(c_puct_base=0.0), "Expected c_puct_base > 0.0, but got 0.000000."),
    (dict(dirichlet_epsilon=1.5), "Expected 0.0 <= dirichlet_epsilon <= 1.0, but got 1.500000."),
])
def test_mcts_validation_messages(cpp, kwargs, msg):
    with pytest.raises(ValueError) as exc:
        cpp.MCTS(**kwargs)
    assert str(exc.value) == msg


def test_no_gpu_fails_loudly(cpp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cpp.MCTS()


@pytest.mark.gpu
@pytest.mark.parametrize("game", [0, 2])
def test_cpp_binding_matches_reference_golden(cpp, golden, game):
    g = golden["search"][game]
    m = cpp.MCTS(history_size=g["H"], torch_device="cuda", num_simulations=g["N"], num_threads=g["T"], batch_size=g["B"],
                 dirichlet_epsilon=0.0, seed=5, game_id=game)
    for i, ply in enumerate(g["plies"]):
        m.search("synth")
        assert m.visit_counts() == ply["visit_counts"], i
        assert [f32_bits(v) for v in m.mean_action_values()] == ply["q_bits"], i
        spd = m.self_play_data()
        assert len(spd["features"]) == 8 and spd["features"][0].shape == (1 + 2 * g["H"], 8, 8) and not spd["features"][0].is_cuda
        for k, t in enumerate((0, 3)):
            assert [f32_bits(v) for v in spd["policy"][t].tolist()] == ply["policy_bits_t0_t3"][k], (i, t)
        m.apply_action(ply["action"])
    p = m.position()
    assert [p.player(), p.player1_discs(), p.player2_discs()] == g["final"] and p.is_terminal()
    with pytest.raises(ValueError, match="Self-play data cannot be generated from a terminal position."):
        m.self_play_data()


@pytest.mark.gpu
def test_cpp_binding_callable_and_native_network_equal_ctypes(cpp):
    """Three routes through the same engine must give the same trees: the ctypes MCTS on the native evaluator (weights
    folded by weights.py), the pybind11 MCTS on the native evaluator (weights folded in C++ from the LibTorch tensors),
    and the pybind11 MCTS driven through the reference's callable contract by a callable that runs the same folded net."""
    import ctypes as C
    import torch
    import othello_mcts as ct
    from othello_mcts import _lib, weights
    from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm
    torch.manual_seed(4)
    net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=64).eval().cuda()
    randomize_batchnorm(net, 5)
    lib = _lib.load()
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))

    def via_handle(x):
        assert x.is_cuda and x.dtype == torch.float32 and x.shape == (16, 9, 8, 8)
        n = x.shape[0]
        pol = torch.empty((n, 65), device="cuda", dtype=torch.float32)
        val = torch.empty((n,), device="cuda", dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.oz_resnet_forward_device(h, C.c_void_p(x.data_ptr()), n, C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), stream))
        return {"policy": pol, "value": val}

    kw = dict(history_size=4, torch_device="cuda", num_simulations=96, num_threads=1, batch_size=16, dirichlet_epsilon=0.25, seed=3)
    a, b, c = ct.MCTS(**kw), cpp.MCTS(**kw), cpp.MCTS(**kw)
    try:
        for ply in range(8):
            a.search(net)
            b.search(net)
            c.search(via_handle)
            assert a.visit_counts() == b.visit_counts() == c.visit_counts(), ply
            assert a.mean_action_values() == b.mean_action_values() == c.mean_action_values(), ply
            act = a.position().legal_actions()[int(np.argmax(a.visit_counts()))]
            for m in (a, b, c):
                m.apply_action(act)
        assert b.evaluator() == 3 and c.evaluator() == 0 and b.counters()["evaluations"] > 0
        # an in-place weight update is picked up (re-fold trigger = parameter versions)
        with torch.no_grad():
            net.policy_head.linear.bias.add_(0.5 * torch.randn_like(net.policy_head.linear.bias))
        a.search(net)
        b.search(net)
        assert a.visit_counts() == b.visit_counts()
        b.set_num_simulations(32)
        assert b.num_simulations() == 32
        with pytest.raises(ValueError, match=r"Expected batch_size >= 1, but got 0\."):
            b.set_batch_size(0)
        with pytest.raises(IndexError, match=r"Expected 0 <= action < 65, but got 70\."):
            b.apply_action(70)
    finally:
        lib.oz_resnet_destroy(h)


@pytest.mark.gpu
def test_unmodified_reference_self_play_on_the_compiled_binding(cpp, tmp_path):
    """`othello_mcts` resolved to the COMPILED binding (cpp_binding/ first on sys.path, in a fresh interpreter), the
    reference's Python package unmodified from oracle/_ref/python: train.py:404-452 `_self_play` with the reference's
    AlphaZeroNet runs on the native evaluator and returns a well-formed game."""
    import subprocess
    ref_py = ROOT / "oracle" / "_ref" / "python"
    if not (ref_py / "othello_alphazero" / "train.py").exists():
        pytest.skip("oracle/_ref/python not built")
    code = r'''
import sys
from argparse import Namespace
import numpy as np, torch
import othello_mcts
assert othello_mcts.__file__.replace("\\", "/").endswith("cpp_binding/othello_mcts/__init__.py"), othello_mcts.__file__
from othello_alphazero import train
from othello_alphazero.neural_net import AlphaZeroNet
assert train.MCTS is othello_mcts.MCTS
torch.manual_seed(7)
net = AlphaZeroNet(in_channels=9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).to("cuda").eval()
mcts = train.MCTS(history_size=4, torch_device="cuda", torch_pin_memory=False, num_simulations=64, num_threads=2, batch_size=8,
                  c_puct_base=20000.0, c_puct_init=2.5, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
np.random.seed(1)
d = train._self_play(mcts, net, Namespace(self_play_temperature=1.0))
n = len(d["features"])
assert n == len(d["policies"]) == len(d["values"]) and n % 8 == 0 and n >= 8 * 50
assert d["features"][0].shape == (9, 8, 8) and abs(float(d["policies"][0].sum()) - 1.0) < 1e-5
assert mcts.evaluator() == 3 and mcts.counters()["evaluations"] > 0
assert mcts.position().is_terminal()
print("OK", n)
'''
    env = dict(**__import__("os").environ, PYTHONPATH=f"{ROOT / 'othello-alphazero_b200' / 'cpp_binding'}:{ref_py}")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
