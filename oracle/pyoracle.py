"""TEST INFRASTRUCTURE ONLY: ctypes view of oracle/liboz_oracle.so and oracle/_ref/libref_shim.so.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
only.  The product package (othello_mcts) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
if str(HERE) not in sys.path:
    sys.path.insert(0, str(HERE))
import build_oracle  # noqa: E402

u64, i32, f32p = C.c_uint64, C.c_int, C.POINTER(C.c_float)


class Pos(C.Structure):
    _fields_ = [("player", C.c_int32), ("p1", u64), ("p2", u64), ("legal", u64), ("next_legal", u64)]

    def key(self):
        return (self.player, self.p1, self.p2, self.legal, self.next_legal)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(build_oracle.build_oracle()))
        PP = C.POINTER(Pos)
        L.ozo_get_legal_moves.restype = u64; L.ozo_get_legal_moves.argtypes = [u64, u64]
        L.ozo_get_flips.restype = u64; L.ozo_get_flips.argtypes = [u64, u64, u64]
        L.ozo_initial_position.argtypes = [PP]
        L.ozo_apply_move.argtypes = [PP, u64, PP]
        L.ozo_apply_pass.argtypes = [PP, PP]
        L.ozo_apply_action.argtypes = [PP, i32, PP]
        L.ozo_legal_actions.restype = i32; L.ozo_legal_actions.argtypes = [PP, C.POINTER(i32)]
        L.ozo_perft.restype = u64; L.ozo_perft.argtypes = [PP, i32]
        L.ozo_transform_action.restype = i32; L.ozo_transform_action.argtypes = [i32, i32]
        L.ozo_positions_to_features.argtypes = [PP, i32, f32p, i32, i32]
        L.ozo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.ozo_mcts_create.restype = C.c_void_p
        L.ozo_mcts_create.argtypes = [i32, i32, i32, i32, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ozo_mcts_destroy.argtypes = [C.c_void_p]
        L.ozo_mcts_set_rng.argtypes = [C.c_void_p, u64, C.c_uint32, f32p, i32, i32]
        L.ozo_mcts_set_evaluator.argtypes = [C.c_void_p, i32, C.c_void_p, C.c_void_p]
        L.ozo_mcts_reset_position.argtypes = [C.c_void_p]
        L.ozo_mcts_position.argtypes = [C.c_void_p, PP]
        L.ozo_mcts_search.argtypes = [C.c_void_p]
        L.ozo_mcts_visit_counts.restype = i32; L.ozo_mcts_visit_counts.argtypes = [C.c_void_p, C.POINTER(i32)]
        L.ozo_mcts_mean_action_values.restype = i32; L.ozo_mcts_mean_action_values.argtypes = [C.c_void_p, f32p]
        L.ozo_mcts_root_visit_count.restype = i32; L.ozo_mcts_root_visit_count.argtypes = [C.c_void_p]
        L.ozo_mcts_self_play_data.restype = i32; L.ozo_mcts_self_play_data.argtypes = [C.c_void_p, f32p, f32p]
        L.ozo_mcts_apply_action.restype = i32; L.ozo_mcts_apply_action.argtypes = [C.c_void_p, i32]
        L.ozo_mcts_choose_action.restype = i32; L.ozo_mcts_choose_action.argtypes = [C.c_void_p, i32, i32]
        L.ozo_mcts_num_nodes.restype = C.c_long; L.ozo_mcts_num_nodes.argtypes = [C.c_void_p]
        L.ozo_eval_synth.argtypes = [f32p, i32, i32, f32p, f32p]
        L.ozo_eval_hash.argtypes = [f32p, i32, i32, f32p, f32p]
        _lib = L
    return _lib


def ref_shim() -> C.CDLL | None:
    """The unmodified reference rules/encoding behind a C ABI (None if never built)."""
    global _ref
    if _ref is None:
        path = build_oracle.build_ref_shim()
        if path is None or not Path(path).exists():
            return None
        R = C.CDLL(str(path))
        PP = C.POINTER(Pos)
        R.ref_get_legal_moves.restype = u64; R.ref_get_legal_moves.argtypes = [u64, u64]
        R.ref_get_flips.restype = u64; R.ref_get_flips.argtypes = [u64, u64, u64]
        R.ref_initial_position.argtypes = [PP]
        R.ref_apply_move.argtypes = [PP, u64, PP]
        R.ref_apply_pass.argtypes = [PP, PP]
        R.ref_apply_action.argtypes = [PP, i32, PP]
        R.ref_legal_actions.restype = i32; R.ref_legal_actions.argtypes = [PP, C.POINTER(i32)]
        R.ref_perft.restype = u64; R.ref_perft.argtypes = [PP, i32]
        R.ref_transform_action.restype = i32; R.ref_transform_action.argtypes = [i32, i32]
        R.ref_positions_to_features.argtypes = [PP, i32, f32p, i32, i32]
        _ref = R
    return _ref


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(f32p)


EVAL_CALLBACK, EVAL_SYNTH, EVAL_HASH = 0, 1, 2
SYM_PHILOX, SYM_IDENTITY = 0, 1
_EVAL_CB = C.CFUNCTYPE(None, C.c_void_p, f32p, i32, i32, f32p, f32p)


class OracleMCTS:
    """Python handle on the C restatement of MCTS (same method names as the reference)."""

    def __init__(self, history_size=4, num_simulations=800, num_threads=2, batch_size=16,
                 c_puct_base=20000.0, c_puct_init=2.5, dirichlet_epsilon=0.25, dirichlet_alpha=0.5,
                 evaluator=EVAL_SYNTH, seed=0, game_id=0, gamma_table=None, sym_mode=SYM_PHILOX):
        self.L = lib()
        self.H = history_size
        self.h = self.L.ozo_mcts_create(history_size, num_simulations, num_threads, batch_size,
                                        c_puct_base, c_puct_init, dirichlet_epsilon, dirichlet_alpha)
        self._table = None
        self._cb = None
        self.set_rng(seed, game_id, gamma_table, sym_mode)
        if callable(evaluator):
            self.set_callback(evaluator)
        else:
            self.L.ozo_mcts_set_evaluator(self.h, evaluator, None, None)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ozo_mcts_destroy(self.h)
            self.h = None

    def set_rng(self, seed, game_id, gamma_table=None, sym_mode=SYM_PHILOX):
        if gamma_table is not None:
            self._table = np.ascontiguousarray(gamma_table, dtype=np.float32)
            self.L.ozo_mcts_set_rng(self.h, seed, game_id, fptr(self._table), len(self._table), sym_mode)
        else:
            self._table = None
            self.L.ozo_mcts_set_rng(self.h, seed, game_id, None, 0, sym_mode)

    def set_callback(self, fn):
        """fn(features[B,P,8,8] float32 ndarray) -> (policy[B,65], value[B])."""
        def tramp(_ctx, feat, batch, planes, pol, val):
            x = np.ctypeslib.as_array(feat, shape=(batch, planes, 8, 8))
            p, v = fn(x)
            np.ctypeslib.as_array(pol, shape=(batch, 65))[:] = np.asarray(p, dtype=np.float32)
            np.ctypeslib.as_array(val, shape=(batch,))[:] = np.asarray(v, dtype=np.float32)
        self._cb = _EVAL_CB(tramp)
        self.L.ozo_mcts_set_evaluator(self.h, EVAL_CALLBACK, C.cast(self._cb, C.c_void_p), None)

    def reset_position(self):
        self.L.ozo_mcts_reset_position(self.h)

    def position(self) -> Pos:
        p = Pos()
        self.L.ozo_mcts_position(self.h, C.byref(p))
        return p

    def search(self):
        self.L.ozo_mcts_search(self.h)

    def visit_counts(self):
        buf = (i32 * 40)()
        n = self.L.ozo_mcts_visit_counts(self.h, buf)
        return list(buf[:n])

    def mean_action_values(self):
        buf = np.zeros(40, np.float32)
        n = self.L.ozo_mcts_mean_action_values(self.h, fptr(buf))
        return buf[:n].copy()

    def root_visit_count(self):
        return self.L.ozo_mcts_root_visit_count(self.h)

    def self_play_data(self):
        planes = 1 + 2 * self.H
        f = np.zeros((8, planes, 8, 8), np.float32)
        p = np.zeros((8, 65), np.float32)
        rc = self.L.ozo_mcts_self_play_data(self.h, fptr(f), fptr(p))
        if rc:
            raise ValueError("terminal position" if rc == -1 else "root not expanded")
        return f, p

    def apply_action(self, a):
        rc = self.L.ozo_mcts_apply_action(self.h, a)
        if rc == -1:
            raise IndexError(a)
        if rc == -2:
            raise ValueError(a)

    def choose_action(self, ply, temperature_plies=12):
        return self.L.ozo_mcts_choose_action(self.h, ply, temperature_plies)

    def legal_actions(self):
        p = self.position()
        buf = (i32 * 40)()
        n = self.L.ozo_legal_actions(C.byref(p), buf)
        return list(buf[:n])


def synth_net_numpy(x: np.ndarray):
    """The exact synthetic net via the C evaluator (for feeding other implementations)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b, planes = x.shape[0], x.shape[1]
    pol = np.zeros((b, 65), np.float32)
    val = np.zeros((b,), np.float32)
    lib().ozo_eval_synth(fptr(x), b, planes, fptr(pol), fptr(val))
    return pol, val


def hash_net_numpy(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    b, planes = x.shape[0], x.shape[1]
    pol = np.zeros((b, 65), np.float32)
    val = np.zeros((b,), np.float32)
    lib().ozo_eval_hash(fptr(x), b, planes, fptr(pol), fptr(val))
    return pol, val
