"""Lockstep AlphaZero-vs-AlphaZero match play on one GPU.

The reference evaluates two checkpoints by playing them against each other one game at a time:
`evaluation.play_games` (evaluation.py:15-90) creates two `AlphaZeroPlayer`s (player.py:177-259: each owns an
`MCTS` with dirichlet_epsilon = 0 and its own network), and `play_game` (player.py:33-96) alternates
`get_action` (search, then the most visited action, ties broken at random) with `apply_action` on BOTH players,
once with either colour.  Here all games of a match run in lockstep: one device engine per network holds a tree
for every game (as each reference player does), and at every ply each engine searches exactly the games in which
its player is to move (`oz_engine_search_range`) -- a pass is a ply (the reference searches forced passes too), so
Black is to move at even plies and White at odd plies in every unfinished game, and the games an engine has to
search always form one contiguous half.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check
from .engine import SelfPlayEngine


class MatchEngine:
    """`num_pairs` pairs of games between network A and network B.  Game g < num_pairs: A is Black (player 1),
    B is White; game num_pairs + g: the same pair with colours swapped (evaluation.py:63-83 plays both)."""

    def __init__(self, num_pairs: int, history_size: int = 4, num_simulations: int = 800, num_threads: int = 2,
                 batch_size: int = 16, c_puct_base: float = 20000.0, c_puct_init: float = 2.5, *, seed: int = 0,
                 first_game_id: int = 0, sym_mode: int = _lib.SYM_RANDOM, cuda_device: int | None = None) -> None:
        if num_pairs < 1:
            raise ValueError(f"Expected num_pairs >= 1, but got {num_pairs}.")
        self.num_pairs = num_pairs
        self.num_games = 2 * num_pairs
        kw = dict(history_size=history_size, num_simulations=num_simulations, num_threads=num_threads, batch_size=batch_size,
                  c_puct_base=c_puct_base, c_puct_init=c_puct_init, dirichlet_epsilon=0.0, dirichlet_alpha=0.5,   # player.py:208-219
                  seed=seed, first_game_id=first_game_id, sym_mode=sym_mode, temperature_plies=0, cuda_device=cuda_device)
        self._a = SelfPlayEngine(self.num_games, **kw)
        self._b = SelfPlayEngine(self.num_games, **kw)
        self.plies = 0

    def reset(self) -> None:
        self._a.reset()
        self._b.reset()
        self.plies = 0

    def engines(self) -> tuple[SelfPlayEngine, SelfPlayEngine]:
        return self._a, self._b

    def play_ply(self) -> int:
        """One ply of every unfinished game; returns the number of games still running."""
        P = self.num_pairs
        black_to_move = self.plies % 2 == 0
        # A is Black in games [0, P) and White in [P, 2P)
        first_a, first_b = (0, P) if black_to_move else (P, 0)
        check(self._a._lib.oz_engine_search_range(self._a._h, first_a, P))
        check(self._b._lib.oz_engine_search_range(self._b._h, first_b, P))
        acts_a, acts_b = self._a.choose_actions(), self._b.choose_actions()      # most visited, random tie-break (player.py:253-258)
        acts = np.concatenate([acts_a[:P], acts_b[P:]]) if black_to_move else np.concatenate([acts_b[:P], acts_a[P:]])
        self._a.apply_actions(acts)                                              # play_game applies to both players (player.py:74-76)
        self._b.apply_actions(acts)
        self.plies += 1
        return int((acts >= 0).sum())

    def play(self, net_a, net_b, max_plies: int = 200) -> dict:
        """Play all games to the end.  `net_a` / `net_b`: AlphaZeroNet (native evaluator) or 'synth' / 'hash'.
        Returns {"result": [2P] 1 Black wins / 2 White wins / 0 draw (play_game's return value), "black_discs",
        "white_discs", "a_is_black": [2P] bool, "score_a": wins + draws/2 of network A, "plies": lockstep plies}."""
        self.reset()
        self._a.prepare(net_a)
        self._b.prepare(net_b)
        for _ in range(max_plies):
            if self.play_ply() == 0:
                break
        pos = self._a.positions_array()
        if (pos["player"] != 0).any():
            raise RuntimeError(f"{int((pos['player'] != 0).sum())} games unfinished after {max_plies} plies")
        black = np.bitwise_count(pos["p1"]).astype(np.int32)
        white = np.bitwise_count(pos["p2"]).astype(np.int32)
        result = np.where(black > white, 1, np.where(white > black, 2, 0)).astype(np.int32)   # player.py:83-96
        a_is_black = np.arange(self.num_games) < self.num_pairs
        a_wins = ((result == 1) & a_is_black) | ((result == 2) & ~a_is_black)
        return {"result": result, "black_discs": black, "white_discs": white, "a_is_black": a_is_black,
                "score_a": float(a_wins.sum() + 0.5 * (result == 0).sum()), "plies": self.plies}

    @staticmethod
    def game_results(outcome: dict, player_a: str, player_b: str) -> list[dict]:
        """The rows `evaluation.play_games` appends to its JSON file (evaluation.py:66-83), pair by pair."""
        P = len(outcome["result"]) // 2
        rows = []
        for g in range(P):
            rows.append({"player1": player_a, "player2": player_b, "result": int(outcome["result"][g])})
            rows.append({"player1": player_b, "player2": player_a, "result": int(outcome["result"][P + g])})
        return rows
