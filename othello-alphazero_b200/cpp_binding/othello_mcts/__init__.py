"""`othello_mcts` as the reference ships it -- a compiled pybind11 extension re-exported by a three-line package
(reference: cpp/src/othello_mcts/__init__.py:1-6) -- but bound to the B200 engine: `_othello_mcts_impl` is
csrc/othello_mcts_pybind.cpp (the reference's binding file with its bodies replaced by calls into the C ABI of
include/othello_b200.h), built by build_ext.py.  Put this directory's parent on sys.path instead of
othello-alphazero_b200/ to use it; the ctypes package next door offers the same surface plus the batched classes.
"""
import torch  # noqa: F401  (libtorch must be loaded first, as in the reference package)

from ._othello_mcts_impl import MCTS, Position, get_flips, get_legal_moves  # noqa: F401

__all__ = ["MCTS", "Position", "get_flips", "get_legal_moves"]
