"""GPU: statistical parity of the parts whose random streams cannot be pinned bit for bit.

The reference seeds a private mt19937 from std::random_device per search (search_thread.cpp:22), so its Dirichlet
noise (gamma_distribution(alpha, 1), :230-249) and symmetry draws cannot be replayed.  The bit-exact tests inject
the noise; here the engine's OWN on-device sampler (Philox + Marsaglia-Tsang) is compared with the unmodified
reference module (oracle/_ref, CPU) as distributions, and against the analytic moments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _exact_net(x):
    """D4-equivariant synthetic net of SURVEY.md section 8c (same formula as the engine's built-in 'synth')."""
    import torch
    occ = x[:, 1] + x[:, 2]
    box = torch.nn.functional.conv2d(occ[:, None], torch.ones(1, 1, 3, 3, device=x.device), padding=1)[:, 0]
    w = torch.cat([1.0 + box.flatten(1), torch.ones(x.shape[0], 1, device=x.device)], dim=1)
    value = (1.0 - 2.0 * x[:, 0, 0, 0]) * (x[:, 1].sum(dim=(1, 2)) - x[:, 2].sum(dim=(1, 2))) / 64.0
    return {"policy": w / w.sum(dim=1, keepdim=True), "value": value}


def test_root_visit_distribution_matches_reference(ref_module):
    """Same position, same search shape, eps = 0.25, alpha = 0.5: the per-child root visit counts of 400 engine
    searches and 400 reference searches are two samples of one distribution (two-sample KS per child)."""
    from scipy.stats import ks_2samp
    from othello_mcts import SelfPlayEngine
    kw = dict(history_size=4, num_simulations=200, num_threads=1, batch_size=16, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
    n = 400
    opening = [19, 18]                                 # two plies in: an asymmetric position with several children
    eng = SelfPlayEngine(n, seed=20261018, **kw)
    for a in opening:
        eng.apply_actions(np.full(n, a, np.int32))
    eng.search("synth")
    nch, counts, _, _ = eng.root_stats()
    k = int(nch[0])
    assert (nch == k).all() and k >= 3
    ours = counts[:, :k]

    m = ref_module.MCTS(torch_device="cpu", **kw)
    theirs = np.zeros((n, k), np.int64)
    for i in range(n):
        m.reset_position()
        for a in opening:
            m.apply_action(a)
        m.search(_exact_net)
        theirs[i] = m.visit_counts()
    assert ours.sum(axis=1).tolist() == theirs.sum(axis=1).tolist()          # same simulation accounting (quirk Q1)
    assert ours.std(axis=0).max() > 1.0                                      # the noise does move the search
    for c in range(k):
        p = ks_2samp(ours[:, c], theirs[:, c]).pvalue
        assert p > 1e-4, f"child {c}: KS p = {p:.2e}; means {ours[:, c].mean():.2f} vs {theirs[:, c].mean():.2f}"


def test_noise_free_search_is_identical_to_reference(ref_module):
    """Control for the test above: with eps = 0 both sides are deterministic and must agree exactly."""
    from othello_mcts import MCTS
    kw = dict(history_size=4, num_simulations=200, num_threads=1, batch_size=16, dirichlet_epsilon=0.0)
    a, b = MCTS(torch_device="cuda", **kw), ref_module.MCTS(torch_device="cpu", **kw)
    for act in (19, 18):
        a.apply_action(act)
        b.apply_action(act)
    a.search("synth")
    b.search(_exact_net)
    assert a.visit_counts() == b.visit_counts()
    assert a.mean_action_values() == b.mean_action_values()


def test_symmetry_draws_are_uniform():
    """search_thread.cpp:24,92: one uniform draw from {0..7} per non-terminal leaf.  The engine's draws are visible
    through the features it hands to an external evaluator: plane 1 of a leaf is the Black discs under symmetry t."""
    import torch
    from scipy.stats import chisquare
    from othello_mcts import MCTS, Position
    from othello_mcts.batch import features_batch, POSITION_DTYPE
    # a position with no symmetry at all: after 19, 18, 17 the 8 images of the root are pairwise different
    pos = Position.initial_position()
    for a in (19, 18, 17):
        pos = pos.apply_action(a)
    hist = np.zeros((8, 1), POSITION_DTYPE)
    for t in range(8):
        hist[t, 0] = (pos.player(), 0, pos.player1_discs(), pos.player2_discs(), pos.legal_moves(), 0)
    images = features_batch(hist, np.ones(8, np.int32), np.arange(8, dtype=np.int32), 1)[:, 1].reshape(8, 64)
    assert len({im.tobytes() for im in images}) == 8
    seen = []

    def spy(x):
        # first round after a re-root: every slot holds the (childless) root itself
        b = x[:, 1].reshape(x.shape[0], 64).cpu().numpy()
        for row in b:
            hit = [t for t in range(8) if (row == images[t]).all()]
            if len(hit) == 1:
                seen.append(hit[0])
        n = x.shape[0]
        return {"policy": torch.full((n, 65), 1.0 / 65, device=x.device), "value": torch.zeros(n, device=x.device)}

    for seed in range(40):
        m = MCTS(history_size=1, torch_device="cuda", num_simulations=16, num_threads=1, batch_size=16, seed=seed)
        for a in (19, 18, 17):
            m.apply_action(a)
        m.search(spy)
    assert len(seen) == 40 * 16
    freq = np.bincount(seen, minlength=8)
    assert chisquare(freq).pvalue > 1e-4, freq


def test_self_play_loop_as_train_py_drives_it():
    """The call sequence of train.py:404-452 (_self_play) against the drop-in MCTS with a real AlphaZeroNet on the
    native evaluator: reset_position, position().is_terminal(), search(net), visit_counts(), self_play_data(),
    apply_action() until the game ends; sample shapes / dtypes / normalisation as the training loop expects."""
    import torch
    from othello_mcts import MCTS
    from othello_mcts.neural_net import AlphaZeroNet
    torch.manual_seed(0)
    net = AlphaZeroNet(9, conv_channels=128, num_residual_blocks=1, value_head_hidden_channels=32).eval().cuda()
    mcts = MCTS(history_size=4, torch_device="cuda", torch_pin_memory=True, num_simulations=32, num_threads=2, batch_size=8,
                c_puct_base=20000.0, c_puct_init=2.5, dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
    rng = np.random.default_rng(0)
    features, policies = [], []
    mcts.reset_position()
    plies = 0
    with torch.no_grad():
        while not mcts.position().is_terminal():
            mcts.search(net)
            visit_counts = np.array(mcts.visit_counts(), dtype=np.float64)
            actions = mcts.position().legal_actions()
            assert len(visit_counts) == len(actions) and visit_counts.sum() > 0
            if plies < 12:
                action = int(rng.choice(actions, p=visit_counts / visit_counts.sum()))
            else:
                best = np.flatnonzero(visit_counts == visit_counts.max())
                action = int(actions[int(rng.choice(best))])
            data = mcts.self_play_data()
            assert len(data["features"]) == 8 and len(data["policy"]) == 8
            for f, p in zip(data["features"], data["policy"]):
                assert f.shape == (9, 8, 8) and f.dtype == torch.float32 and not f.is_cuda
                assert p.shape == (65,) and abs(float(p.sum()) - 1.0) < 1e-5
            features.extend(data["features"])
            policies.extend(data["policy"])
            mcts.apply_action(action)
            plies += 1
    final = mcts.position()
    assert final.is_terminal() and final.legal_actions() == [] and 9 <= plies <= 128
    assert len(features) == 8 * plies
    with pytest.raises(ValueError):
        mcts.self_play_data()                          # mcts.cpp:64-68: terminal position


def _golden_stats():
    import json
    from pathlib import Path
    return json.loads((Path(__file__).resolve().parent / "golden" / "golden_stats.json").read_text())


def test_root_visit_distribution_2048_searches_vs_reference_golden():
    """SURVEY.md section 4 (>= 1 k searches): 2 048 engine searches of one position with the engine's own on-device
    noise (Philox + Marsaglia-Tsang gamma) and symmetry draws against 2 048 searches of the unmodified reference
    (tests/golden/golden_stats.json, generate_golden_stats.py): two-sample KS on every root child's visit count."""
    from scipy.stats import ks_2samp
    from othello_mcts import SelfPlayEngine
    ref = _golden_stats()["root"]
    theirs = np.array(ref["visit_counts"])
    n, k = theirs.shape
    eng = SelfPlayEngine(n, seed=77, **ref["kwargs"])
    for a in ref["opening"]:
        eng.apply_actions(np.full(n, a, np.int32))
    eng.search("synth")
    nch, counts, _, _ = eng.root_stats()
    assert (nch == k).all()
    ours = counts[:, :k]
    assert ours.sum(axis=1).tolist() == theirs.sum(axis=1).tolist()
    for c in range(k):
        p = ks_2samp(ours[:, c], theirs[:, c]).pvalue
        assert p > 1e-4, f"child {c}: KS p = {p:.2e}; means {ours[:, c].mean():.2f} vs {theirs[:, c].mean():.2f}"
        assert abs(ours[:, c].mean() - theirs[:, c].mean()) < 4.0 * theirs[:, c].std() / np.sqrt(n) * np.sqrt(2) + 0.5


def test_self_play_histograms_vs_reference_golden():
    """1 024 whole self-play games of the lockstep engine (device-side action choice: proportional to the visit counts
    for 12 plies, then a random argmax; on-device noise and symmetries) against 1 024 games the unmodified reference
    played through train.py's `_self_play` loop (golden_stats.json): first-move histogram (chi-square test of
    homogeneity), game length in plies and final disc difference (two-sample KS)."""
    from scipy.stats import chi2_contingency, ks_2samp
    from othello_mcts import SelfPlayEngine
    ref = _golden_stats()["games"]
    n = len(ref["plies"])
    eng = SelfPlayEngine(n, seed=4242, temperature_plies=ref["temperature_plies"], **ref["kwargs"])
    rec = eng.run("synth")
    assert (rec["plies"] > 0).all()
    first = rec["records"]["action"][:, 0]
    pos = eng.positions_array()
    assert (pos["player"] == 0).all()
    diff = np.bitwise_count(pos["p1"]).astype(np.int64) - np.bitwise_count(pos["p2"]).astype(np.int64)
    moves = sorted(set(ref["first_move"]) | set(first.tolist()))
    assert moves == [19, 26, 37, 44]                                   # the four legal opening moves
    table = np.array([[int((first == m).sum()) for m in moves], [ref["first_move"].count(m) for m in moves]])
    assert chi2_contingency(table).pvalue > 1e-4, table
    p_len = ks_2samp(rec["plies"], np.array(ref["plies"])).pvalue
    p_diff = ks_2samp(diff, np.array(ref["disc_diff"])).pvalue
    assert p_len > 1e-4 and p_diff > 1e-4, (p_len, p_diff, rec["plies"].mean(), np.mean(ref["plies"]), diff.mean(), np.mean(ref["disc_diff"]))
