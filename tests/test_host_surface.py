"""CPU: the drop-in boundary without a GPU -- the library loads and exports every symbol the header
declares, the host-side Position matches the golden vectors, and argument validation raises what the
reference raises (same exception types and texts).  No kernel is launched here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from othello_mcts import _lib
    header = (ROOT / "include" / "othello_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(oz_[a-z0-9_]+)\s*\(", header))
    declared -= {"oz_status"}
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"declared in include/othello_b200.h but not exported: {missing}"
    unbound = sorted(declared - set(_lib.SIGNATURES))
    assert not unbound, f"no ctypes signature for: {unbound}"
    assert _lib.load().oz_abi_version() == 4


def test_struct_sizes_match_abi():
    from othello_mcts import _lib
    assert C.sizeof(_lib.OzPosition) == 40
    assert C.sizeof(_lib.OzPlyRecord) == 176
    assert C.sizeof(_lib.OzConfig) == 80
    assert C.sizeof(_lib.OzCounters) == 120


def test_position_matches_golden(golden):
    from othello_mcts import Position, get_flips, get_legal_moves
    p = Position.initial_position()
    assert (p.player(), p.player1_discs(), p.player2_discs(), p.legal_moves()) == (
        1, 0x0000000810000000, 0x0000001008000000, 0x0000102004080000)
    assert p.legal_actions() == [19, 26, 37, 44]
    assert get_legal_moves(0, 0) == 0
    # replay golden rows through the host-side value type
    pos, prev_next = None, None
    for row in golden["rules"]:
        if pos is None or row["pos"] != prev_next:
            pos = Position.initial_position()
            assert row["pos"][1:3] == [pos.player1_discs(), pos.player2_discs()]
        assert [pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()] == row["pos"][:4]
        assert pos.legal_actions() == row["actions"]
        me, opp = (row["pos"][1], row["pos"][2]) if row["pos"][0] == 1 else (row["pos"][2], row["pos"][1])
        assert get_legal_moves(me, opp) == row["legal_of_me"]
        if row["action"] != 64:
            assert get_flips(1 << (63 - row["action"]), me, opp) == row["flips"]
        pos = pos.apply_action(row["action"])
        prev_next = row["next"]
        assert [pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()] == row["next"][:4]
        if pos.is_terminal():
            assert pos.legal_actions() == []
            pos = None


def test_position_random_playouts_vs_oracle(oracle):
    """The host-side Position value type (position.h:39-149, 264-427) against the oracle on 150 seeded random playouts
    (about 9 000 positions with passes, double passes and early ends): every field, the action list, the flips of the move
    played and the successor -- bit for bit."""
    from helpers import playout_positions
    from othello_mcts import Position, get_flips, get_legal_moves
    rows = playout_positions(oracle, 150, seed=77)
    assert len(rows) > 8000 and any(a == 64 for _, _, a, _ in rows)
    pos = None
    for p, acts, a, q in rows:
        if pos is None or (pos.player(), pos.player1_discs(), pos.player2_discs()) != (p.player, p.p1, p.p2):
            pos = Position.initial_position()                      # a new playout starts
        assert (pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()) == (p.player, p.p1, p.p2, p.legal)
        assert pos.legal_actions() == acts and not pos.is_terminal()
        me, opp = (p.p1, p.p2) if p.player == 1 else (p.p2, p.p1)
        assert get_legal_moves(me, opp) == p.legal
        if a != 64:
            flips = get_flips(1 << (63 - a), me, opp)
            assert flips != 0 and (flips & ~opp) == 0              # a legal move flips opponent discs only
            assert (q.p1 if p.player == 1 else q.p2) == me | flips | (1 << (63 - a))
        pos = pos.apply_action(a)
        assert (pos.player(), pos.player1_discs(), pos.player2_discs(), pos.legal_moves()) == (q.player, q.p1, q.p2, q.legal)
        assert pos.is_terminal() == (q.player == 0)
        if pos.is_terminal():
            assert pos.legal_actions() == []
            pos = None


def test_position_api_and_errors():
    from othello_mcts import Position
    p = Position.initial_position()
    assert [p[27], p[28], p[0]] == [2, 1, 0]
    assert p.is_legal_move(19) and not p.is_legal_move(0)
    assert str(p).splitlines()[0] == "  a b c d e f g h"
    assert str(p).splitlines()[3] == "3 · · · × · · · ·"
    assert str(p).splitlines()[4] == "4 · · × ○ ● · · ·"
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
    with pytest.raises(ValueError, match=r"0b0{63}1 is not a legal move\."):
        p.apply_move(1)
    q = p.apply_move(1 << (63 - 26))
    assert q.player() == 2 and q.legal_moves() == 0x0000280020000000
    assert p.player() == 1                      # value semantics: p is unchanged
    with pytest.raises(TypeError):
        Position()


@pytest.mark.parametrize("kwargs,msg", [
    (dict(history_size=0), "Expected history_size >= 1, but got 0."),
    (dict(num_simulations=0), "Expected num_simulations >= 1, but got 0."),
    (dict(num_threads=-2), "Expected num_threads >= 1, but got -2."),
    (dict(batch_size=0), "Expected batch_size >= 1, but got 0."),
    (dict(c_puct_base=0.0), "Expected c_puct_base > 0.0, but got 0.000000."),
    (dict(c_puct_init=-1.0), "Expected c_puct_init >= 0.0, but got -1.000000."),
    (dict(dirichlet_epsilon=1.5), "Expected 0.0 <= dirichlet_epsilon <= 1.0, but got 1.500000."),
    (dict(dirichlet_alpha=-0.5), "Expected dirichlet_alpha >= 0.0, but got -0.500000."),
])
def test_mcts_validation_messages(kwargs, msg):
    """mcts.cpp:167-241: validation runs before any device work, so it is checkable without a GPU."""
    from othello_mcts import MCTS
    with pytest.raises(ValueError) as exc:
        MCTS(**kwargs)
    assert str(exc.value) == msg


def test_mcts_signature_matches_reference():
    """othello_mcts.cpp:90-112: keyword names, order and defaults."""
    import inspect
    from othello_mcts import MCTS
    params = list(inspect.signature(MCTS.__init__).parameters.values())[1:11]
    assert [(p.name, p.default) for p in params] == [
        ("history_size", 4), ("torch_device", "cpu"), ("torch_pin_memory", False), ("num_simulations", 800),
        ("num_threads", 2), ("batch_size", 16), ("c_puct_base", 20000.0), ("c_puct_init", 2.5),
        ("dirichlet_epsilon", 0.25), ("dirichlet_alpha", 0.5)]
    for name in ("reset_position", "position", "search", "visit_counts", "mean_action_values", "self_play_data",
                 "apply_action", "history_size", "set_history_size", "torch_device", "set_torch_device",
                 "num_simulations", "set_num_simulations", "num_threads", "set_num_threads", "batch_size",
                 "set_batch_size", "c_puct_base", "set_c_puct_base", "c_puct_init", "set_c_puct_init",
                 "dirichlet_epsilon", "set_dirichlet_epsilon", "dirichlet_alpha", "set_dirichlet_alpha"):
        assert callable(getattr(MCTS, name))


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from othello_mcts import MCTS, movegen_batch
    with pytest.raises(RuntimeError, match="no CUDA device"):
        MCTS()
    with pytest.raises(RuntimeError):
        movegen_batch(np.array([1], np.uint64), np.array([2], np.uint64))
    from othello_mcts.neural_net import AlphaZeroNet
    from othello_mcts.training import NativeTrainer
    with pytest.raises(RuntimeError, match="no CPU path"):
        NativeTrainer(AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32), max_batch=32)
    from othello_mcts import MatchEngine
    with pytest.raises(RuntimeError, match="no CUDA device"):
        MatchEngine(2)


def test_trainer_state_vectors_follow_the_module():
    """The trainer's flat parameter / buffer vectors are the module's parameters() order and its BatchNorm buffers in
    module order: the element counts the C ABI computes from the shape must equal what PyTorch reports (checked here
    without a device through the same formulas; oz_trainer_create is exercised by the GPU tests)."""
    import torch
    from othello_mcts.neural_net import AlphaZeroNet
    for blocks, hv in ((1, 32), (9, 128)):
        net = AlphaZeroNet(17, conv_channels=128, num_residual_blocks=blocks, value_head_hidden_channels=hv)
        C, L = 128, 1 + 2 * blocks
        n_params = C * 17 * 9 + 3 * C + (L - 1) * (C * C * 9 + 3 * C) + (2 * C + 2 + 2 + 2 + 65 * 128 + 65) + (C + 1 + 1 + 1 + hv * 64 + hv + hv + 1)
        assert n_params == sum(p.numel() for p in net.parameters())
        n_buffers = 2 * C * L + 4 + 2
        assert n_buffers == sum(2 * m.num_features for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d))
        names = [n for n, _ in net.named_parameters()]
        assert names[:4] == ["conv_block.conv.weight", "conv_block.conv.bias", "conv_block.norm.weight", "conv_block.norm.bias"]
        assert names[-14:-8] == ["policy_head.conv.weight", "policy_head.conv.bias", "policy_head.norm.weight", "policy_head.norm.bias",
                                 "policy_head.linear.weight", "policy_head.linear.bias"]
        assert names[-8:] == ["value_head.conv.weight", "value_head.conv.bias", "value_head.norm.weight", "value_head.norm.bias",
                              "value_head.linear1.weight", "value_head.linear1.bias", "value_head.linear2.weight", "value_head.linear2.bias"]


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = ROOT / "othello-alphazero_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert "pyoracle" not in text and "oz_oracle" not in text and "liboz_oracle" not in text, path


def test_weight_folding_matches_batchnorm():
    """Folding conv+BN (eval) into one affine conv reproduces the module output in fp32."""
    import torch
    from othello_mcts import weights
    from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm
    torch.manual_seed(0)
    net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32).eval()
    randomize_batchnorm(net, 1)
    w, b = weights._fold(net.conv_block.conv, net.conv_block.norm)
    x = torch.randn(3, 9, 8, 8)
    want = net.conv_block.norm(net.conv_block.conv(x))
    got = torch.nn.functional.conv2d(x, w, b, padding=1)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    shape, blob = weights.fold_alphazero_net(net)
    assert (shape.in_planes, shape.channels, shape.blocks, shape.value_hidden) == (9, 128, 1, 32)
    from othello_mcts import _lib
    assert len(blob) == _lib.load().oz_resnet_blob_bytes(C.byref(shape))
    # first conv section: tap-major bf16 [9][128][32]; tap 4 (centre), cout 5, cin 3
    arr = np.frombuffer(blob[: 9 * 128 * 32 * 2], dtype=np.int16).reshape(9, 128, 32)
    want_bf16 = w[5, 3, 1, 1].to(torch.bfloat16).view(torch.int16).item()
    assert arr[4, 5, 3] == want_bf16 and (arr[:, :, 9:] == 0).all()
    assert weights.natively_supported(net) and weights.is_alphazero_net(net)
    net.train()
    assert not weights.natively_supported(net)
