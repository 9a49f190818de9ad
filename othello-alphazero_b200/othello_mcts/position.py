"""`Position`, `get_legal_moves`, `get_flips`: the rules surface of the reference module
(reference: cpp/src/lib/othello_mcts.cpp:53-87 binding cpp/src/include/position.h)."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import OzPosition, check

_MASK64 = (1 << 64) - 1


def _u64(x: int, name: str) -> int:
    if not isinstance(x, int) or isinstance(x, bool):
        raise TypeError(f"{name} must be an int")
    if not (0 <= x <= _MASK64):
        raise TypeError(f"{name} must fit in an unsigned 64-bit integer")
    return x


def get_legal_moves(player_discs: int, opponent_discs: int) -> int:
    """Bitboard of legal moves (position.h:202-229)."""
    return _lib.load().oz_get_legal_moves(_u64(player_discs, "player_discs"), _u64(opponent_discs, "opponent_discs"))


def get_flips(move_mask: int, player_discs: int, opponent_discs: int) -> int:
    """Bitboard of discs flipped by a move (position.h:231-262)."""
    return _lib.load().oz_get_flips(_u64(move_mask, "move_mask"), _u64(player_discs, "player_discs"),
                                    _u64(opponent_discs, "opponent_discs"))


class Position:
    """Immutable Othello position (position.h:39-149).  Every apply_* returns a new object."""

    __slots__ = ("_p",)

    def __init__(self, *_args, **_kwargs):
        raise TypeError("Position: No constructor defined!")  # same as the pybind11 class

    @classmethod
    def _wrap(cls, raw: OzPosition) -> "Position":
        self = object.__new__(cls)
        self._p = raw
        return self

    @staticmethod
    def initial_position() -> "Position":
        raw = OzPosition()
        check(_lib.load().oz_position_initial(C.byref(raw)))
        return Position._wrap(raw)

    def player(self) -> int:
        return self._p.player

    def player1_discs(self) -> int:
        return self._p.p1

    def player2_discs(self) -> int:
        return self._p.p2

    def __getitem__(self, index: int) -> int:
        out = C.c_int()
        check(_lib.load().oz_position_at(C.byref(self._p), int(index), C.byref(out)))
        return out.value

    def legal_moves(self) -> int:
        return self._p.legal

    def is_legal_move(self, index: int) -> bool:
        out = C.c_int()
        check(_lib.load().oz_position_is_legal_move(C.byref(self._p), int(index), C.byref(out)))
        return bool(out.value)

    def legal_actions(self) -> list[int]:
        buf = (C.c_int * 40)()
        n = C.c_int()
        check(_lib.load().oz_position_legal_actions(C.byref(self._p), buf, C.byref(n)))
        return list(buf[: n.value])

    def apply_move(self, move_mask: int) -> "Position":
        raw = OzPosition()
        check(_lib.load().oz_position_apply_move(C.byref(self._p), _u64(move_mask, "move_mask"), C.byref(raw)))
        return Position._wrap(raw)

    def apply_pass(self) -> "Position":
        raw = OzPosition()
        check(_lib.load().oz_position_apply_pass(C.byref(self._p), C.byref(raw)))
        return Position._wrap(raw)

    def apply_action(self, action: int) -> "Position":
        raw = OzPosition()
        check(_lib.load().oz_position_apply_action(C.byref(self._p), int(action), C.byref(raw)))
        return Position._wrap(raw)

    def is_terminal(self) -> bool:
        return self._p.player == 0

    def __str__(self) -> str:
        buf = C.create_string_buffer(512)
        check(_lib.load().oz_position_to_string(C.byref(self._p), buf, 512))
        return buf.value.decode("utf-8")
