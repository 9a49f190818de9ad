"""GPU: bitboard rule kernels, feature encoding and perft against the oracle and the golden vectors
(bit-exact).  Everything goes through the C ABI (othello_mcts -> libothello_b200.so)."""
import ctypes as C

import numpy as np
import pytest

from helpers import planes_to_masks, playout_positions, pos_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def playouts(oracle):
    return playout_positions(oracle, 1500, seed=2024)      # ~90k positions


def test_movegen_flips_apply_vs_oracle(oracle, playouts):
    import othello_mcts as om
    from othello_mcts.batch import POSITION_DTYPE
    L = oracle.lib()
    pos = [r[0] for r in playouts]
    acts = np.array([r[2] for r in playouts], np.int32)
    me = np.array([p.p1 if p.player == 1 else p.p2 for p in pos], np.uint64)
    opp = np.array([p.p2 if p.player == 1 else p.p1 for p in pos], np.uint64)
    legal = om.movegen_batch(me, opp)
    want = np.array([L.ozo_get_legal_moves(int(a), int(b)) for a, b in zip(me, opp)], np.uint64)
    assert (legal == want).all()
    moves = np.array([(1 << (63 - a)) if a != 64 else 0 for a in acts], np.uint64)
    fl = om.flips_batch(moves, me, opp)
    wantf = np.array([L.ozo_get_flips(int(m), int(a), int(b)) for m, a, b in zip(moves, me, opp)], np.uint64)
    assert (fl == wantf).all()
    out = om.apply_action_batch(pos_array(pos, POSITION_DTYPE), acts)
    nxt = pos_array([r[3] for r in playouts], POSITION_DTYPE)
    for f in ("player", "p1", "p2", "legal", "next_legal"):
        assert (out[f] == nxt[f]).all(), f


def test_movegen_random_bitboards_vs_oracle(oracle):
    """Arbitrary disc sets (also unreachable ones) and the edge cases: empty, full, disjointness."""
    import othello_mcts as om
    L = oracle.lib()
    rng = np.random.default_rng(7)
    a = rng.integers(0, 2**64, 200_000, dtype=np.uint64)
    b = rng.integers(0, 2**64, 200_000, dtype=np.uint64) & ~a
    a[:4] = [0, 2**64 - 1, 0, 0x8100000000000081]
    b[:4] = [0, 0, 2**64 - 1, 0x7EFFFFFFFFFFFF7E]
    legal = om.movegen_batch(a, b)
    idx = rng.integers(0, len(a), 20_000)
    for i in idx.tolist() + [0, 1, 2, 3]:
        assert int(legal[i]) == L.ozo_get_legal_moves(int(a[i]), int(b[i]))
    mv = np.uint64(1) << rng.integers(0, 64, len(a)).astype(np.uint64)
    fl = om.flips_batch(mv, a, b)
    for i in idx.tolist():
        assert int(fl[i]) == L.ozo_get_flips(int(mv[i]), int(a[i]), int(b[i]))
    assert om.movegen_batch(np.zeros(0, np.uint64), np.zeros(0, np.uint64)).size == 0     # empty input


def test_rules_golden_on_gpu(golden):
    import othello_mcts as om
    from othello_mcts.batch import POSITION_DTYPE
    rows = golden["rules"]
    pos = np.zeros(len(rows), POSITION_DTYPE)
    for i, r in enumerate(rows):
        pos[i]["player"], pos[i]["p1"], pos[i]["p2"], pos[i]["legal"], pos[i]["next_legal"] = r["pos"]
    out = om.apply_action_batch(pos, np.array([r["action"] for r in rows], np.int32))
    for i, r in enumerate(rows):
        assert [int(out[i][f]) for f in ("player", "p1", "p2", "legal", "next_legal")] == r["next"]


def test_features_golden_and_oracle(oracle, golden, playouts):
    import othello_mcts as om
    from othello_mcts.batch import POSITION_DTYPE
    for H in (1, 4, 8):
        rows = [r for r in golden["features"] if r["H"] == H]
        hist = np.zeros((len(rows), H), POSITION_DTYPE)
        for i, r in enumerate(rows):
            for k, h in enumerate(r["history"]):
                hist[i, k]["player"], hist[i, k]["p1"], hist[i, k]["p2"], hist[i, k]["legal"], hist[i, k]["next_legal"] = h
        f = om.features_batch(hist, [len(r["history"]) for r in rows], [r["t"] for r in rows], H)
        for i, r in enumerate(rows):
            assert planes_to_masks(f[i].reshape(1 + 2 * H, 64)) == r["plane_masks"]
    # bigger sweep against the oracle: sliding windows over playouts, random symmetry
    L = oracle.lib()
    H = 8
    rng = np.random.default_rng(1)
    n = 4000
    starts = rng.integers(0, len(playouts) - 1, n)
    hist = np.zeros((n, H), POSITION_DTYPE)
    nh = np.zeros(n, np.int32)
    tf = rng.integers(0, 8, n).astype(np.int32)
    want = np.zeros((n, (1 + 2 * H) * 64), np.float32)
    for i, s in enumerate(starts.tolist()):
        chain = [playouts[s][0]]
        k = s
        while len(chain) < H and k > 0 and playouts[k - 1][3].key() == playouts[k][0].key():
            k -= 1
            chain.append(playouts[k][0])
        nh[i] = len(chain)
        arr = (oracle.Pos * len(chain))(*chain)
        for j, p in enumerate(chain):
            hist[i, j]["player"], hist[i, j]["p1"], hist[i, j]["p2"] = p.player, p.p1, p.p2
        L.ozo_positions_to_features(arr, len(chain), oracle.fptr(want[i]), H, int(tf[i]))
    got = om.features_batch(hist, nh, tf, H).reshape(n, -1)
    assert (got == want).all()


def test_perft_golden(golden):
    import othello_mcts as om
    for d, want in enumerate(golden["perft"], start=1):
        assert om.perft(d) == want, d
        assert om.perft(d, mode=1) == want, d
    assert om.perft(0) == 1
    assert om.perft(11) == 212_258_800          # SURVEY.md section 8(c), verified against the reference there


def test_perft_from_midgame_vs_oracle(oracle, playouts):
    """Non-initial roots, including positions where the side to move must pass or the game ends."""
    import othello_mcts as om
    L = oracle.lib()
    picks = [playouts[i][0] for i in (10, 500, 3000, 20000)]
    picks += [r[0] for r in playouts if r[0].legal == 0][:3]          # forced passes
    picks += [r[0] for r in playouts[-60:] if r[0].player != 0][-3:]  # near the end of a game
    for p in picks:
        pos = om.Position.initial_position()
        pos._p.player, pos._p.p1, pos._p.p2, pos._p.legal, pos._p.next_legal = p.player, p.p1, p.p2, p.legal, p.next_legal
        for d in (1, 2, 5, 6):
            assert om.perft(d, pos) == L.ozo_perft(C.byref(p), d), (p.key(), d)
            assert om.perft(d, pos, mode=1) == L.ozo_perft(C.byref(p), d), (p.key(), d)
