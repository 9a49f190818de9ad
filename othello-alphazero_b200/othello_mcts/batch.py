"""Batched rule kernels and perft (additive surface; include/othello_b200.h "rules")."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import OzPosition, check
from .position import Position

POSITION_DTYPE = np.dtype([("player", "<i4"), ("_pad", "<i4"), ("p1", "<u8"), ("p2", "<u8"), ("legal", "<u8"),
                           ("next_legal", "<u8")])
assert POSITION_DTYPE.itemsize == C.sizeof(OzPosition)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def movegen_batch(player, opponent) -> np.ndarray:
    """legal[i] = get_legal_moves(player[i], opponent[i]) on the GPU (position.h:202-229)."""
    a = np.ascontiguousarray(player, dtype=np.uint64)
    b = np.ascontiguousarray(opponent, dtype=np.uint64)
    out = np.zeros(a.shape, np.uint64)
    check(_lib.load().oz_movegen_batch(_p(a), _p(b), _p(out), a.size))
    return out


def flips_batch(move, player, opponent) -> np.ndarray:
    """flips[i] = get_flips(move[i], player[i], opponent[i]) on the GPU (position.h:231-262)."""
    m = np.ascontiguousarray(move, dtype=np.uint64)
    a = np.ascontiguousarray(player, dtype=np.uint64)
    b = np.ascontiguousarray(opponent, dtype=np.uint64)
    out = np.zeros(a.shape, np.uint64)
    check(_lib.load().oz_flips_batch(_p(m), _p(a), _p(b), _p(out), a.size))
    return out


def apply_action_batch(positions: np.ndarray, actions) -> np.ndarray:
    """positions: structured array of POSITION_DTYPE; out[i] = positions[i].apply_action(actions[i])."""
    pos = np.ascontiguousarray(positions, dtype=POSITION_DTYPE)
    act = np.ascontiguousarray(actions, dtype=np.int32)
    out = np.zeros(pos.shape, POSITION_DTYPE)
    check(_lib.load().oz_apply_action_batch(_p(pos), _p(act), _p(out), pos.size))
    return out


def features_batch(history: np.ndarray, n_hist, transformation, history_size: int) -> np.ndarray:
    """history: [n, history_size] POSITION_DTYPE (index 0 = current); returns [n, 1+2H, 8, 8] float32
    (transformation.h:83-116)."""
    hist = np.ascontiguousarray(history, dtype=POSITION_DTYPE)
    n = hist.shape[0]
    nh = np.ascontiguousarray(n_hist, dtype=np.int32)
    tf = np.ascontiguousarray(transformation, dtype=np.int32)
    out = np.zeros((n, 1 + 2 * history_size, 8, 8), np.float32)
    check(_lib.load().oz_features_batch(_p(hist), _p(nh), _p(tf), history_size, _p(out), n))
    return out


def perft(depth: int, position: Position | None = None, mode: int = 0, return_ms: bool = False):
    """Leaf count `depth` plies below `position` (default: the initial position) on the GPU."""
    pos = position if position is not None else Position.initial_position()
    nodes = C.c_uint64()
    ms = C.c_double()
    check(_lib.load().oz_perft(C.byref(pos._p), int(depth), int(mode), C.byref(nodes), C.byref(ms)))
    return (nodes.value, ms.value) if return_ms else nodes.value
