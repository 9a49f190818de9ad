"""GPU: the device search engine against the golden traces of the unmodified reference and against the
oracle -- visit counts, Q bit patterns, chosen actions, self-play data: all bit-exact."""
import numpy as np
import pytest

from helpers import f32_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("game", range(4))
def test_mcts_matches_reference_golden(golden, game):
    """The reference's own results (deterministic regime) reproduced by the CUDA engine through the
    MCTS drop-in class with the built-in exact net."""
    from othello_mcts import MCTS
    g = golden["search"][game]
    m = MCTS(history_size=g["H"], torch_device="cuda", num_simulations=g["N"], num_threads=g["T"],
             batch_size=g["B"], dirichlet_epsilon=0.0, seed=5, game_id=game)
    for i, ply in enumerate(g["plies"]):
        m.search("synth")
        assert m.visit_counts() == ply["visit_counts"], i
        assert [f32_bits(v) for v in m.mean_action_values()] == ply["q_bits"], i
        spd = m.self_play_data()
        for k, t in enumerate((0, 3)):
            assert [f32_bits(v) for v in spd["policy"][t].tolist()] == ply["policy_bits_t0_t3"][k], (i, t)
        m.apply_action(ply["action"])
    if len(g["plies"]) > 10:
        p = m.position()
        assert [p.player(), p.player1_discs(), p.player2_discs()] == g["final"]
        assert p.is_terminal()


def test_callable_evaluator_path_matches_golden(golden):
    """The reference's evaluator contract: a Python callable fed device tensors (othello_mcts.cpp:36-45)."""
    import torch
    from othello_mcts import MCTS

    def net(x):
        assert x.is_cuda and x.dtype == torch.float32 and x.shape[1:] == (9, 8, 8)
        occ = x[:, 1] + x[:, 2]
        box = torch.nn.functional.conv2d(occ[:, None], torch.ones(1, 1, 3, 3, device=x.device), padding=1)[:, 0]
        w = torch.cat([1.0 + box.flatten(1), torch.ones(x.shape[0], 1, device=x.device)], dim=1)
        value = (1.0 - 2.0 * x[:, 0, 0, 0]) * (x[:, 1].sum(dim=(1, 2)) - x[:, 2].sum(dim=(1, 2))) / 64.0
        return {"policy": w / w.sum(dim=1, keepdim=True), "value": value}

    g = golden["search"][0]
    m = MCTS(history_size=g["H"], torch_device="cuda", num_simulations=g["N"], num_threads=1, batch_size=g["B"],
             dirichlet_epsilon=0.0)
    for ply in g["plies"][:12]:
        m.search(net)
        assert m.visit_counts() == ply["visit_counts"]
        assert [f32_bits(v) for v in m.mean_action_values()] == ply["q_bits"]
        m.apply_action(ply["action"])


@pytest.mark.parametrize("cfg", [
    dict(H=8, N=800, T=2, B=16, eps=0.25, ev="hash"),     # the reference's default search shape
    dict(H=4, N=96, T=1, B=8, eps=0.25, ev="hash"),
    dict(H=1, N=130, T=3, B=7, eps=1.0, ev="synth"),      # pure-noise priors, odd batch
    dict(H=2, N=64, T=1, B=1, eps=0.0, ev="hash"),
])
def test_noisy_search_vs_oracle(oracle, cfg):
    """Dirichlet noise (injected draws), random symmetries (shared Philox streams) and a
    non-equivariant pseudo-random net: whole games, every ply compared bit for bit."""
    from othello_mcts import MCTS
    table = np.random.default_rng(3).gamma(0.5, 1.0, 4099).astype(np.float32)
    table[::97] = 0.0
    ev = {"hash": oracle.EVAL_HASH, "synth": oracle.EVAL_SYNTH}[cfg["ev"]]
    seed, gid = 0xC0FFEE1234, 17
    o = oracle.OracleMCTS(history_size=cfg["H"], num_simulations=cfg["N"], num_threads=cfg["T"], batch_size=cfg["B"],
                          dirichlet_epsilon=cfg["eps"], evaluator=ev, seed=seed, game_id=gid, gamma_table=table)
    m = MCTS(history_size=cfg["H"], torch_device="cuda", num_simulations=cfg["N"], num_threads=cfg["T"],
             batch_size=cfg["B"], dirichlet_epsilon=cfg["eps"], seed=seed, game_id=gid)
    m.set_gamma_table(table)
    max_plies = 70 if cfg["N"] < 500 else 14
    for ply in range(max_plies):
        if m.position().is_terminal():
            break
        o.search()
        m.search(cfg["ev"])
        assert m.visit_counts() == o.visit_counts(), ply
        assert [f32_bits(v) for v in m.mean_action_values()] == [f32_bits(v) for v in o.mean_action_values()], ply
        if ply % 5 == 0:
            f, p = o.self_play_data()
            spd = m.self_play_data()
            assert all((spd["features"][t].numpy() == f[t]).all() for t in range(8))
            assert all((spd["policy"][t].numpy() == p[t]).all() for t in range(8))
        a = o.choose_action(ply, 12)
        o.apply_action(a)
        m.apply_action(a)
        op = o.position()
        mp = m.position()
        assert (mp.player(), mp.player1_discs(), mp.player2_discs(), mp.legal_moves()) == (op.player, op.p1, op.p2, op.legal)


def test_lockstep_games_vs_oracle(oracle):
    """Many games in lockstep on the device, each compared with its own oracle instance, including
    the on-device action choice, the compact records and the training tensors."""
    from othello_mcts import SelfPlayEngine
    G, H, N, T, B = 24, 4, 48, 2, 4
    seed, first = 77, 1000
    table = np.random.default_rng(9).gamma(0.5, 1.0, 1021).astype(np.float32)
    e = SelfPlayEngine(G, history_size=H, num_simulations=N, num_threads=T, batch_size=B, dirichlet_epsilon=0.25,
                       seed=seed, first_game_id=first, temperature_plies=6)
    e.set_gamma_table(table)
    e.prepare("hash")
    os_ = [oracle.OracleMCTS(history_size=H, num_simulations=N, num_threads=T, batch_size=B, dirichlet_epsilon=0.25,
                             evaluator=oracle.EVAL_HASH, seed=seed, game_id=first + g, gamma_table=table) for g in range(G)]
    done = [False] * G
    samples = [([], []) for _ in range(G)]
    for ply in range(130):
        want_actions = np.full(G, -1, np.int32)
        want_counts = np.zeros((G, 36), np.int32)
        for g, o in enumerate(os_):
            if done[g]:
                continue
            o.search()
            vc = o.visit_counts()
            want_counts[g, : len(vc)] = vc
            want_actions[g] = o.choose_action(ply, 6)
            f, p = o.self_play_data()
            samples[g][0].append(f)
            samples[g][1].append(p)
        actions, counts, unfinished = e.play_ply(want_counts=True)
        live = [g for g in range(G) if not done[g]]
        assert (counts[live] == want_counts[live]).all(), ply
        assert (actions == want_actions).all(), ply
        for g in live:
            os_[g].apply_action(int(want_actions[g]))
            done[g] = os_[g].position().player == 0
        assert unfinished == G - sum(done)
        if unfinished == 0:
            break
    assert all(done)
    pos = e.positions()
    for g, o in enumerate(os_):
        op = o.position()
        assert (pos[g].player(), pos[g].player1_discs(), pos[g].player2_discs()) == (0, op.p1, op.p2)
    rec = e.records()
    ts = e.training_samples()
    row = 0
    for g, o in enumerate(os_):
        n = int(rec["plies"][g])
        assert n == len(samples[g][0])
        op = o.position()
        d = bin(op.p1).count("1") - bin(op.p2).count("1")
        z = 1.0 if d > 0 else (-1.0 if d < 0 else 0.0)
        for k in range(n):
            assert (ts["features"][row: row + 8].numpy() == samples[g][0][k]).all(), (g, k)
            assert (ts["policies"][row: row + 8].numpy() == samples[g][1][k]).all(), (g, k)
            assert (ts["values"][row: row + 8].numpy() == (z if k % 2 == 0 else -z)).all()
            row += 8
    assert row == len(ts["values"])
    c = e.counters()
    assert c["games_finished"] == G and c["simulations"] > 0 and c["kernel_launches"] > 0


def test_partition_invariance():
    """Results are keyed by game id, not by where a game runs: 16 games in one engine equal the
    same games split over two engines (the multi-GPU partition of SURVEY.md section 8e)."""
    from othello_mcts import SelfPlayEngine
    kw = dict(history_size=2, num_simulations=32, num_threads=1, batch_size=8, seed=5, temperature_plies=4)
    whole = SelfPlayEngine(16, first_game_id=0, **kw)
    a, b = SelfPlayEngine(8, first_game_id=0, **kw), SelfPlayEngine(8, first_game_id=8, **kw)
    for e in (whole, a, b):
        e.prepare("hash")
    for _ in range(10):
        w, _, _ = whole.play_ply()
        x, _, _ = a.play_ply()
        y, _, _ = b.play_ply()
        assert (w == np.concatenate([x, y])).all()


def test_tree_reuse_setters_and_errors(oracle):
    from othello_mcts import MCTS
    m = MCTS(history_size=4, torch_device="cuda", num_simulations=100, num_threads=1, batch_size=16, dirichlet_epsilon=0.0)
    o = oracle.OracleMCTS(history_size=4, num_simulations=100, num_threads=1, batch_size=16, dirichlet_epsilon=0.0)
    assert m.visit_counts() == [] and m.mean_action_values() == []
    with pytest.raises(ValueError, match="The root node has not been expanded yet."):
        m.self_play_data()
    m.apply_action(19)                 # apply on an unexpanded root (mcts.cpp:140-148)
    o.apply_action(19)
    m.search("synth"); o.search()
    assert m.visit_counts() == o.visit_counts()
    assert sum(m.visit_counts()) == 112 - 16          # Q1: the first batch only expands the root
    with pytest.raises(ValueError, match=r"^0 is not a legal action\.$"):
        m.apply_action(0)
    with pytest.raises(IndexError, match=r"Expected 0 <= action < 65, but got 70\."):
        m.apply_action(70)
    with pytest.raises(ValueError, match="Pass is not allowed when there are legal moves."):
        m.apply_action(64)
    # setters between searches keep the tree (reference: plain member updates)
    m.set_num_simulations(40); o2 = 40
    m.set_batch_size(4)
    assert (m.num_simulations(), m.batch_size(), m.num_threads()) == (40, 4, 1)
    with pytest.raises(ValueError, match=r"Expected batch_size >= 1, but got 0\."):
        m.set_batch_size(0)
    a = m.position().legal_actions()[int(np.argmax(m.visit_counts()))]
    before = m.visit_counts()[int(np.argmax(m.visit_counts()))]
    m.apply_action(a)
    m.search("synth")
    assert sum(m.visit_counts()) >= before - 1          # Q2: the subtree and its counts survive re-rooting
    m.reset_position()
    assert m.position().legal_actions() == [19, 26, 37, 44] and m.visit_counts() == []


def test_rolling_pool_is_invariant_to_slot_count():
    """Rolling game pool: 10 games through 3 slots (finished games leave, slots restart with the next unplayed id)
    give, game id by game id, the records of 10 games played side by side -- RNG streams are keyed by game id, not
    by slot -- and the training tensors of the store match the side-by-side ones."""
    import torch
    from othello_mcts import SelfPlayEngine
    kw = dict(history_size=4, num_simulations=48, num_threads=1, batch_size=8, dirichlet_epsilon=0.25, seed=77,
              first_game_id=100, temperature_plies=6)
    wide = SelfPlayEngine(10, **kw)
    wide.run("hash")
    ref = wide.records()
    ref_samples = wide.training_samples()
    roll = SelfPlayEngine(3, **kw)
    got = roll.run("hash", total_games=10)
    assert roll.finished_games() == 10
    assert sorted(got["game_ids"].tolist()) == list(range(100, 110))
    offsets, rows = {}, 0
    for g in range(10):
        offsets[100 + g] = rows
        rows += int(ref["plies"][g]) * 8
    roll_samples = roll.training_samples()
    assert roll_samples["features"].shape[0] == rows
    r0 = 0
    for row, gid in enumerate(got["game_ids"].tolist()):
        g = gid - 100
        n = int(ref["plies"][g])
        assert int(got["plies"][row]) == n
        for name in ("p1", "p2", "player", "action", "n_children", "visit_counts"):      # every field but the padding
            assert np.array_equal(got["records"][row, :n][name], ref["records"][g, :n][name]), (gid, name)
        a, b = offsets[gid], offsets[gid] + n * 8
        for k in ("features", "policies", "values"):
            assert torch.equal(roll_samples[k][r0:r0 + n * 8], ref_samples[k][a:b]), (gid, k)
        r0 += n * 8
    c = roll.counters()
    assert c["games_finished"] == 10


def test_arena_growth_is_transparent():
    """Trees live in fixed arenas (the reference's are heap-allocated, search_node.h:29).  Before every search the
    engine checks the worst case (one new node with 33 children per simulation) and reallocates bigger arenas when
    needed; a run that starts with absurdly small arenas must give exactly the records of a run that never grows."""
    from othello_mcts import SelfPlayEngine
    kw = dict(history_size=4, num_simulations=64, num_threads=2, batch_size=8, dirichlet_epsilon=0.25, seed=9, temperature_plies=8)
    small = SelfPlayEngine(6, max_nodes_per_game=64, **kw)
    big = SelfPlayEngine(6, **kw)
    a = small.run("hash")
    b = big.run("hash")
    assert (a["plies"] == b["plies"]).all() and a["plies"].min() > 8
    for name in ("p1", "p2", "player", "action", "n_children", "visit_counts"):
        assert np.array_equal(a["records"][name], b["records"][name]), name


@pytest.mark.parametrize("evaluator", ["hash", "resnet"])
def test_graph_replay_equals_eager_launches(evaluator):
    """Launch-bound batches replay one captured search round (the round index lives on the device).  Same trees as
    launching every kernel eagerly, search after search, across re-roots and an arena growth."""
    import torch
    from othello_mcts import MCTS
    from othello_mcts.neural_net import AlphaZeroNet
    net = "hash"
    if evaluator == "resnet":
        torch.manual_seed(1)
        net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).eval().cuda()
    kw = dict(history_size=4, torch_device="cuda", num_simulations=96, num_threads=2, batch_size=8, dirichlet_epsilon=0.25, seed=4)
    eager, graph = MCTS(**kw), MCTS(**kw)
    eager.set_graphs(0)
    graph.set_graphs(2)
    launches0 = graph.counters()["kernel_launches"]
    for ply in range(14):
        eager.search(net)
        graph.search(net)
        assert eager.visit_counts() == graph.visit_counts(), ply
        assert eager.mean_action_values() == graph.mean_action_values(), ply
        if ply == 6:
            graph.set_num_simulations(128)             # parameter change: the graph is rebuilt
            eager.set_num_simulations(128)
        a = graph.position().legal_actions()[int(np.argmax(graph.visit_counts()))]
        eager.apply_action(a)
        graph.apply_action(a)
    assert graph.counters()["kernel_launches"] > launches0


@pytest.mark.parametrize("evaluator,groups", [("hash", 2), ("hash", 3), ("resnet", 2), ("resnet", 4), ("resnet", 8)])
def test_pipelined_search_equals_single_stream(evaluator, groups):
    """The pipelined search splits the games into groups on their own streams (one group's tree kernels run underneath
    another group's convolutions).  Every game's search is independent of the others, so whole games must come out
    identical to the single-stream engine: same records, bit for bit, also with profiling events on."""
    import torch
    from othello_mcts import SelfPlayEngine
    from othello_mcts.neural_net import AlphaZeroNet
    net = "hash"
    if evaluator == "resnet":
        torch.manual_seed(2)
        net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).eval().cuda()
    kw = dict(history_size=4, num_simulations=64, num_threads=2, batch_size=8, dirichlet_epsilon=0.25, seed=11, temperature_plies=8)
    one, piped = SelfPlayEngine(11, **kw), SelfPlayEngine(11, **kw)
    one.set_pipeline(1)
    piped.set_pipeline(groups, pdl=(groups == 4))
    piped.set_profiling(groups == 2)
    a = one.run(net)
    b = piped.run(net)
    assert (a["plies"] == b["plies"]).all() and a["plies"].min() > 8
    for name in ("p1", "p2", "player", "action", "n_children", "visit_counts"):
        assert np.array_equal(a["records"][name], b["records"][name]), name
    ca, cb = one.counters(), piped.counters()
    for name in ("simulations", "evaluations", "nodes_expanded", "searches", "plies"):
        assert ca[name] == cb[name], name
    if groups == 2:
        assert cb["eval_ms"] > 0 and cb["select_ms"] > 0


def test_automatic_graph_branches_equal_single_chain():
    """Graph-replayed searches of 8 to 48 games on a 128-channel network split the games into parallel branches of the
    search's graph by themselves (each branch its own cluster-tower launches).  Whole games must come out identical to the
    same engine with the split switched off."""
    import torch
    from othello_mcts import SelfPlayEngine
    from othello_mcts.neural_net import AlphaZeroNet
    torch.manual_seed(3)
    net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).eval().cuda()
    kw = dict(history_size=4, num_simulations=64, num_threads=2, batch_size=16, dirichlet_epsilon=0.25, seed=5, temperature_plies=8)
    one, auto = SelfPlayEngine(12, **kw), SelfPlayEngine(12, **kw)      # 12 games x 32 leaves = 12 clusters of 32 boards
    one.set_pipeline(1)
    a = one.run(net)
    b = auto.run(net)
    assert (a["plies"] == b["plies"]).all() and a["plies"].min() > 8
    for name in ("p1", "p2", "player", "action", "n_children", "visit_counts"):
        assert np.array_equal(a["records"][name], b["records"][name]), name
    ca, cb = one.counters(), auto.counters()
    for name in ("simulations", "evaluations", "nodes_expanded", "searches", "plies"):
        assert ca[name] == cb[name], name


def test_samples_from_travelled_records_equal_engine_samples():
    """parallel.samples_from_records (oz_training_samples_from_records): compact records expanded away from the engine
    that played them -- what a multi-GPU run does on the destination rank -- equal the engine's own training tensors."""
    import torch
    from othello_mcts import SelfPlayEngine, parallel
    eng = SelfPlayEngine(6, history_size=3, num_simulations=32, num_threads=1, batch_size=8, seed=17, temperature_plies=6)
    rec = eng.run("hash")
    want = eng.training_samples()
    got = parallel.samples_from_records(rec, history_size=3)
    assert want["values"].shape[0] == int(rec["plies"].sum()) * 8 > 0
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_match_engine_vs_oracle_players(oracle):
    """Lockstep match play against the ORACLE: every game is replayed by play_game's loop (player.py:33-96) with two
    AlphaZeroPlayer-shaped players (player.py:177-259) built on the oracle's MCTS restatement -- own tree each,
    dirichlet_epsilon = 0, most visited action, BOTH players apply every action -- where the two "networks" are the
    built-in evaluators 'hash' (player A) and 'synth' (player B).  Final discs, results and every action must agree."""
    from othello_mcts import MatchEngine
    po = oracle
    P = 4
    kw = dict(history_size=4, num_simulations=64, num_threads=2, batch_size=8, c_puct_base=20000.0, c_puct_init=2.5)
    match = MatchEngine(P, seed=33, **kw)
    out = match.play("hash", "synth")
    assert out["plies"] >= 50 and set(out["result"]) <= {0, 1, 2}
    final = match.engines()[0].positions_array()
    ev = {"a": po.EVAL_HASH, "b": po.EVAL_SYNTH}
    for g in range(2 * P):
        black, white = ("a", "b") if g < P else ("b", "a")                       # evaluation.py:63-83: both colours
        players = [po.OracleMCTS(evaluator=ev[black], seed=33, game_id=g, dirichlet_epsilon=0.0, dirichlet_alpha=0.5, **kw),
                   po.OracleMCTS(evaluator=ev[white], seed=33, game_id=g, dirichlet_epsilon=0.0, dirichlet_alpha=0.5, **kw)]
        ply = 0
        while True:                                                                # play_game (player.py:44-76)
            pos = players[0].position()
            if pos.player == 0:
                break
            mover = players[pos.player - 1]
            mover.search()                                                         # get_action: search, most visited (player.py:248-259)
            a = mover.choose_action(ply, 0)
            players[0].apply_action(a)
            players[1].apply_action(a)
            ply += 1
        assert (pos.p1, pos.p2) == (int(final["p1"][g]), int(final["p2"][g])), g
        nb, nw = bin(pos.p1).count("1"), bin(pos.p2).count("1")
        assert out["result"][g] == (1 if nb > nw else 2 if nw > nb else 0)
        assert (out["black_discs"][g], out["white_discs"][g]) == (nb, nw)
    rows = MatchEngine.game_results(out, "a", "b")
    assert len(rows) == 2 * P and rows[0]["player1"] == "a" and rows[1]["player1"] == "b"


def test_match_engine_equals_game_by_game_play():
    """Lockstep match play (two engines, each searching the games where its network is to move) must give exactly the
    games the reference's loop gives when it plays them one at a time: play_game (player.py:33-96) with two
    AlphaZeroPlayer-shaped players (own tree each, epsilon = 0, most visited action, both apply every action)."""
    import torch
    from othello_mcts import MatchEngine, SelfPlayEngine
    from othello_mcts.neural_net import AlphaZeroNet
    torch.manual_seed(5)
    net_a = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32).eval().cuda()
    net_b = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32).eval().cuda()
    P = 3
    kw = dict(history_size=4, num_simulations=48, num_threads=1, batch_size=8, seed=21)
    match = MatchEngine(P, **kw)
    out = match.play(net_a, net_b)
    assert out["plies"] >= 50 and set(out["result"]) <= {0, 1, 2}
    assert (out["black_discs"] + out["white_discs"] <= 64).all()
    rows = MatchEngine.game_results(out, "a", "b")
    assert len(rows) == 2 * P and rows[0]["player1"] == "a" and rows[1]["player1"] == "b"
    final = match.engines()[0].positions_array()

    for g in range(2 * P):
        black_net, white_net = (net_a, net_b) if g < P else (net_b, net_a)
        pk = dict(kw, dirichlet_epsilon=0.0, temperature_plies=0, first_game_id=g)
        players = [SelfPlayEngine(1, **pk), SelfPlayEngine(1, **pk)]          # index 0 = Black's player, 1 = White's
        while True:
            pos = players[0].positions_array()[0]
            if pos["player"] == 0:
                break
            mover = int(pos["player"]) - 1
            players[mover].search(black_net if mover == 0 else white_net)
            a = players[mover].choose_actions()
            players[0].apply_actions(a)
            players[1].apply_actions(a)
        assert pos["p1"] == final["p1"][g] and pos["p2"] == final["p2"][g], g


def test_full_size_search_invariants():
    """BASELINE.json config 3 at full size (128x10b random-init, 1 000 lockstep games, 800 sims/move, T x B = 2 x 16, random
    symmetries, Dirichlet noise on), checked through properties that do not depend on the size: the reference's visit
    accounting, value bounds, counters, and invariance of every game to how the games are partitioned."""
    import torch
    from othello_mcts import SelfPlayEngine
    from othello_mcts.neural_net import AlphaZeroNet
    torch.manual_seed(0)
    net = AlphaZeroNet(17, conv_channels=128, num_residual_blocks=9, value_head_hidden_channels=128).eval()
    kw = dict(history_size=8, num_simulations=800, num_threads=2, batch_size=16, c_puct_base=20000.0, c_puct_init=2.5,
              dirichlet_epsilon=0.25, dirichlet_alpha=0.5, seed=2026)
    G = 1000
    eng = SelfPlayEngine(G, **kw)
    eng.prepare(net)
    eng.search(net)
    n, counts, q, root_n = eng.root_stats()
    assert (n == 4).all()                                              # four legal moves at the start
    assert (root_n == 800).all()                                       # root.N += 1 per simulation (search_thread.cpp:78)
    assert (counts.sum(axis=1) == 800 - 32).all()                      # Q1: the first T*B selections hit the childless root
    assert (counts[:, 4:] == 0).all() and (counts[:, :4] >= 0).all()
    assert np.isfinite(q).all() and (np.abs(q) <= 1.0).all()           # Q is a mean of values in [-1, 1]
    c = eng.counters()
    assert c["simulations"] == G * 800 and c["evaluations"] == G * 800
    assert 0 < c["nodes_expanded"] <= G * 800 * 33 and c["path_nodes"] > 0 and c["children_read"] > 0
    assert len({tuple(r) for r in counts[:, :4]}) > G // 2             # noise + random symmetries: games differ

    # the same games 500..999 in an engine of their own (as another GPU would own them): identical, bit for bit
    part = SelfPlayEngine(G // 2, first_game_id=G // 2, **kw)
    part.search(net)
    n2, counts2, q2, root_n2 = part.root_stats()
    assert np.array_equal(counts[G // 2:], counts2) and np.array_equal(q[G // 2:].view(np.uint32), q2.view(np.uint32))

    # Q2: the chosen child's visits carry over into the next search's root count
    actions, _, unfinished = eng.play_ply(want_counts=False)          # searches again (tree reuse is off at ply 0: same root), then moves
    assert unfinished == G
    _, counts_b, _, _ = eng.root_stats()                                # the new roots' children after the re-root
    eng.search(net)
    _, counts_c, _, root_n_c = eng.root_stats()
    assert (root_n_c >= 800).all() and (counts_c.sum(axis=1) >= counts_b.sum(axis=1) + 800 - 32 - 1).all()
    pos = eng.positions_array()
    assert (pos["player"] == 2).all() and (np.bitwise_count(pos["p1"]) + np.bitwise_count(pos["p2"]) == 5).all()


def test_exploration_rate_table_grows_with_repeated_searches(oracle):
    """The reference evaluates logf((1 + N + c_base) / c_base) for any visit count and allows any number of searches from
    the same root (search_thread.cpp:198-203); the engine's table indexed by N is extended on demand.  300 searches of 16
    simulations push the root past the initial 4 096 entries; visit counts and Q stay bit-identical to the oracle."""
    from othello_mcts import MCTS
    po = oracle
    kw = dict(history_size=2, num_simulations=16, num_threads=1, batch_size=16, dirichlet_epsilon=0.0)
    m = MCTS(torch_device="cuda", seed=1, game_id=0, **kw)
    o = po.OracleMCTS(evaluator=po.EVAL_HASH, seed=1, game_id=0, **kw)
    for i in range(300):
        m.search("hash")
        o.search()
        if i % 50 == 49 or i == 299:
            assert m.visit_counts() == o.visit_counts(), i
            assert [f32_bits(v) for v in m.mean_action_values()] == [f32_bits(v) for v in o.mean_action_values()], i
    assert sum(m.visit_counts()) == 300 * 16 - 16 and sum(m.visit_counts()) > 4096


def test_self_play_temperature_changes_the_sampling_weights():
    """train.py:423: visit_counts ** (1 / temperature).  A low temperature concentrates the sampled first moves on the most
    visited child; temperature 1 spreads them in proportion to the counts; a high temperature flattens them."""
    from othello_mcts import SelfPlayEngine
    kw = dict(history_size=2, num_simulations=64, num_threads=1, batch_size=8, dirichlet_epsilon=0.25, seed=9)
    share_of_best, spread = {}, {}
    for temp in (0.02, 1.0, 20.0):
        eng = SelfPlayEngine(512, self_play_temperature=temp, **kw)
        eng.search("hash")
        n, counts, _, _ = eng.root_stats()
        acts = eng.choose_actions()
        legal = [19, 26, 37, 44]
        idx = np.array([legal.index(a) for a in acts])
        share_of_best[temp] = float((counts[np.arange(len(idx)), idx] == counts[:, :4].max(axis=1)).mean())   # a most visited child (ties count)
        spread[temp] = np.bincount(idx, minlength=4) / len(idx)
    assert share_of_best[0.02] > 0.95 > share_of_best[1.0] > share_of_best[20.0]     # (14 / 15) ** 50 = 0.03: near-argmax
    assert spread[20.0].min() > 0.15                               # nearly uniform over the four opening moves
    with pytest.raises(ValueError, match=r"Expected self_play_temperature > 0.0"):
        SelfPlayEngine(2, self_play_temperature=0.0, **kw)
