"""B200-native drop-in for the reference's `othello_mcts` package.

Same exports as the reference (cpp/src/othello_mcts/__init__.py:1-6: ``MCTS, Position, get_flips,
get_legal_moves``; torch is imported first there as well), plus the additive batched surface:
``SelfPlayEngine`` (lockstep games on one GPU), ``MatchEngine`` (lockstep match play between two networks), ``perft``
and the batched rule calls.
Everything computes in libothello_b200.so (sm_100a kernels behind include/othello_b200.h).
"""
import torch  # noqa: F401  (kept first, like the reference package)

from .engine import MCTS, SelfPlayEngine
from .match import MatchEngine
from .position import Position, get_flips, get_legal_moves
from .batch import apply_action_batch, features_batch, flips_batch, movegen_batch, perft

__all__ = ["MCTS", "Position", "get_flips", "get_legal_moves", "SelfPlayEngine", "MatchEngine", "perft", "movegen_batch",
           "flips_batch", "apply_action_batch", "features_batch"]
