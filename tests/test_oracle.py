"""CPU: pin the oracle (oracle/oz_oracle.c) against the golden vectors generated from the unmodified
reference, and -- where oracle/_ref holds the compiled reference -- against the reference itself."""
import ctypes as C
import random

import numpy as np
import pytest

from helpers import f32_bits, planes_to_masks, playout_positions


def _pos(po, t):
    p = po.Pos()
    p.player, p.p1, p.p2, p.legal, p.next_legal = t
    return p


def test_rules_golden(oracle, golden):
    L = oracle.lib()
    for row in golden["rules"]:
        p = _pos(oracle, row["pos"])
        me, opp = (p.p1, p.p2) if p.player == 1 else (p.p2, p.p1)
        assert L.ozo_get_legal_moves(me, opp) == row["legal_of_me"]
        acts = (C.c_int * 40)()
        n = L.ozo_legal_actions(C.byref(p), acts)
        assert list(acts[:n]) == row["actions"]
        if row["action"] != 64:
            assert L.ozo_get_flips(1 << (63 - row["action"]), me, opp) == row["flips"]
        q = oracle.Pos()
        L.ozo_apply_action(C.byref(p), row["action"], C.byref(q))
        assert list(q.key()) == row["next"]


def test_known_answers(oracle):
    """SURVEY.md section 8(c) vectors."""
    L = oracle.lib()
    p = oracle.Pos()
    L.ozo_initial_position(C.byref(p))
    assert (p.p1, p.p2, p.legal) == (0x0000000810000000, 0x0000001008000000, 0x0000102004080000)
    assert L.ozo_get_legal_moves(0, 0) == 0
    assert L.ozo_get_flips(1 << (63 - 26), p.p1, p.p2) == 0x0000001000000000
    q = oracle.Pos()
    L.ozo_apply_action(C.byref(p), 26, C.byref(q))
    assert q.legal == 0x0000280020000000
    assert [L.ozo_transform_action(1, t) for t in range(8)] == [1, 6, 15, 55, 62, 57, 48, 8]
    assert [L.ozo_transform_action(10, t) for t in range(8)] == [10, 13, 22, 46, 53, 50, 41, 17]
    assert [L.ozo_transform_action(19, t) for t in range(8)] == [19, 20, 29, 37, 44, 43, 34, 26]


def test_d4_table_golden(oracle, golden):
    L = oracle.lib()
    assert [[L.ozo_transform_action(a, t) for a in range(65)] for t in range(8)] == golden["d4"]


def test_perft_golden(oracle, golden):
    L = oracle.lib()
    p = oracle.Pos()
    L.ozo_initial_position(C.byref(p))
    for d, want in enumerate(golden["perft"][:8], start=1):
        assert L.ozo_perft(C.byref(p), d) == want


def test_features_golden(oracle, golden):
    L = oracle.lib()
    for row in golden["features"]:
        H, t, hist = row["H"], row["t"], row["history"]
        arr = (oracle.Pos * len(hist))()
        for i, h in enumerate(hist):
            arr[i].player, arr[i].p1, arr[i].p2, arr[i].legal, arr[i].next_legal = h
        f = np.zeros(((1 + 2 * H), 64), np.float32)
        L.ozo_positions_to_features(arr, len(hist), oracle.fptr(f), H, t)
        assert planes_to_masks(f) == row["plane_masks"]


@pytest.mark.parametrize("game", range(4))
def test_search_golden(oracle, golden, game):
    """Deterministic regime (T=1, eps=0, D4-equivariant exact net): visit counts, Q bit patterns,
    self-play policies and the chosen actions of whole games equal the reference's."""
    g = golden["search"][game]
    m = oracle.OracleMCTS(history_size=g["H"], num_simulations=g["N"], num_threads=g["T"], batch_size=g["B"],
                          dirichlet_epsilon=0.0, evaluator=oracle.EVAL_SYNTH, seed=123, game_id=game)
    for ply in g["plies"]:
        m.search()
        assert m.visit_counts() == ply["visit_counts"]
        assert [f32_bits(v) for v in m.mean_action_values()] == ply["q_bits"]
        _, pol = m.self_play_data()
        for k, t in enumerate((0, 3)):
            assert [f32_bits(v) for v in pol[t]] == ply["policy_bits_t0_t3"][k]
        m.apply_action(ply["action"])
    if len(g["plies"]) > 10:
        p = m.position()
        assert [p.player, p.p1, p.p2] == g["final"]


def test_rules_vs_reference(oracle, ref_shim):
    L, R = oracle.lib(), ref_shim
    for p, acts, a, q in playout_positions(oracle, 300, seed=5):
        r = oracle.Pos()
        R.ref_apply_action(C.byref(p), a, C.byref(r))
        assert r.key() == q.key()
        me, opp = (p.p1, p.p2) if p.player == 1 else (p.p2, p.p1)
        assert R.ref_get_legal_moves(me, opp) == L.ozo_get_legal_moves(me, opp)
        if a != 64:
            assert R.ref_get_flips(1 << (63 - a), me, opp) == L.ozo_get_flips(1 << (63 - a), me, opp)
    rng = random.Random(3)
    for _ in range(20000):       # arbitrary (also unreachable) disc sets
        a, b = rng.getrandbits(64), rng.getrandbits(64)
        b &= ~a
        assert R.ref_get_legal_moves(a, b) == L.ozo_get_legal_moves(a, b)
        mv = 1 << rng.randrange(64)
        assert R.ref_get_flips(mv, a, b) == L.ozo_get_flips(mv, a, b)


def test_search_vs_reference_module(oracle, ref_module):
    """Fresh run of the real reference module against the oracle (not only the stored goldens)."""
    import torch

    def net(x):
        pol, val = oracle.synth_net_numpy(x.numpy())
        return {"policy": torch.from_numpy(pol).clone(), "value": torch.from_numpy(val).clone()}

    H, N, B = 3, 150, 8
    m = ref_module.MCTS(history_size=H, num_simulations=N, num_threads=1, batch_size=B, dirichlet_epsilon=0.0)
    o = oracle.OracleMCTS(history_size=H, num_simulations=N, num_threads=1, batch_size=B, dirichlet_epsilon=0.0,
                          evaluator=oracle.EVAL_SYNTH)
    for ply in range(20):
        m.search(net)
        o.search()
        vc = m.visit_counts()
        assert vc == o.visit_counts()
        assert [f32_bits(v) for v in m.mean_action_values()] == [f32_bits(v) for v in o.mean_action_values()]
        a = m.position().legal_actions()[int(np.argmax(vc))]
        m.apply_action(a)
        o.apply_action(a)


def test_choose_action_distribution(oracle):
    """The engine's action rule: proportional sampling while ply < temperature_plies, then argmax."""
    m = oracle.OracleMCTS(history_size=2, num_simulations=64, num_threads=1, batch_size=8, dirichlet_epsilon=0.0)
    m.search()
    counts = np.array(m.visit_counts(), float)
    acts = m.legal_actions()
    hist = {a: 0 for a in acts}
    for gid in range(4000):
        m.set_rng(99, gid)
        hist[m.choose_action(ply=0)] += 1
    freq = np.array([hist[a] for a in acts]) / 4000.0
    assert np.abs(freq - counts / counts.sum()).max() < 0.03
    best = {acts[i] for i in np.flatnonzero(counts == counts.max())}
    assert all(m.choose_action(ply=50) in best for _ in range(5))
