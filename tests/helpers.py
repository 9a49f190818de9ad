"""Shared test helpers: random-playout position sets (via the oracle) and bit tricks."""
import ctypes as C
import random
import struct

import numpy as np


def f32_bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def bits_to_f32(b: int) -> float:
    return struct.unpack("<f", struct.pack("<I", b))[0]


def playout_positions(po, n_games: int, seed: int):
    """Positions met in random playouts: list of (Pos, legal action list, chosen action, next Pos)."""
    L = po.lib()
    rng = random.Random(seed)
    rows = []
    for _ in range(n_games):
        p = po.Pos()
        L.ozo_initial_position(C.byref(p))
        while p.player != 0:
            acts = (C.c_int * 40)()
            n = L.ozo_legal_actions(C.byref(p), acts)
            a = rng.choice(list(acts[:n]))
            q = po.Pos()
            L.ozo_apply_action(C.byref(p), a, C.byref(q))
            rows.append((p, list(acts[:n]), a, q))
            p = q
    return rows


def pos_array(rows_pos, dtype):
    arr = np.zeros(len(rows_pos), dtype)
    for i, p in enumerate(rows_pos):
        arr[i]["player"], arr[i]["p1"], arr[i]["p2"], arr[i]["legal"], arr[i]["next_legal"] = (
            p.player, p.p1, p.p2, p.legal, p.next_legal)
    return arr


def planes_to_masks(planes: np.ndarray):
    """[P,8,8] or [P,64] 0/1 planes -> list of 64-bit masks (square i <-> bit 63-i)."""
    pl = np.asarray(planes).reshape(planes.shape[0], 64)
    return [int(sum(1 << (63 - i) for i in range(64) if v[i] != 0)) for v in pl]
