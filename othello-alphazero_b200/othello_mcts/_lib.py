"""ctypes binding of libothello_b200.so (include/othello_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there
is no Python or CPU fallback for the kernels behind it.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libothello_b200.so"

OZ_OK, OZ_ERR_INVALID_ARGUMENT, OZ_ERR_OUT_OF_RANGE, OZ_ERR_CUDA, OZ_ERR_CAPACITY, OZ_ERR_STATE = range(6)
EVAL_EXTERNAL, EVAL_SYNTH, EVAL_HASH, EVAL_RESNET = range(4)
SYM_RANDOM, SYM_IDENTITY = 0, 1
MAX_CHILDREN = 36

u64, i32, i64, f32 = C.c_uint64, C.c_int32, C.c_int64, C.c_float
P = C.POINTER


class OzPosition(C.Structure):
    _fields_ = [("player", i32), ("p1", u64), ("p2", u64), ("legal", u64), ("next_legal", u64)]


class OzConfig(C.Structure):
    _fields_ = [("num_games", i32), ("history_size", i32), ("num_simulations", i32), ("num_threads", i32),
                ("batch_size", i32), ("c_puct_base", f32), ("c_puct_init", f32), ("dirichlet_epsilon", f32),
                ("dirichlet_alpha", f32), ("seed", u64), ("first_game_id", C.c_uint32), ("device", i32),
                ("evaluator", i32), ("sym_mode", i32), ("temperature_plies", i32), ("max_nodes_per_game", i32),
                ("self_play_temperature", C.c_float), ("reserved", i32)]


class OzResnetShape(C.Structure):
    _fields_ = [("in_planes", i32), ("channels", i32), ("blocks", i32), ("value_hidden", i32)]


class OzPlyRecord(C.Structure):
    _fields_ = [("p1", u64), ("p2", u64), ("player", i32), ("action", i32), ("n_children", i32),
                ("visit_counts", i32 * MAX_CHILDREN)]


class OzCounters(C.Structure):
    _fields_ = [("simulations", u64), ("evaluations", u64), ("nodes_expanded", u64), ("searches", u64),
                ("plies", u64), ("games_finished", u64), ("kernel_launches", u64), ("select_ms", C.c_double),
                ("eval_ms", C.c_double), ("expand_ms", C.c_double), ("reroot_ms", C.c_double), ("conv_ms", C.c_double),
                ("conv_launches", u64), ("path_nodes", u64), ("children_read", u64)]


class OzTrainConfig(C.Structure):
    _fields_ = [("lr", f32), ("momentum", f32), ("l2_weight", f32), ("bn_momentum", f32), ("bn_eps", f32)]


# name -> (restype, argtypes); every symbol include/othello_b200.h declares
SIGNATURES = {
    "oz_last_error": (C.c_char_p, []),
    "oz_abi_version": (C.c_int, []),
    "oz_get_legal_moves": (u64, [u64, u64]),
    "oz_get_flips": (u64, [u64, u64, u64]),
    "oz_position_initial": (C.c_int, [P(OzPosition)]),
    "oz_position_at": (C.c_int, [P(OzPosition), C.c_int, P(C.c_int)]),
    "oz_position_is_legal_move": (C.c_int, [P(OzPosition), C.c_int, P(C.c_int)]),
    "oz_position_legal_actions": (C.c_int, [P(OzPosition), P(C.c_int), P(C.c_int)]),
    "oz_position_apply_move": (C.c_int, [P(OzPosition), u64, P(OzPosition)]),
    "oz_position_apply_pass": (C.c_int, [P(OzPosition), P(OzPosition)]),
    "oz_position_apply_action": (C.c_int, [P(OzPosition), C.c_int, P(OzPosition)]),
    "oz_position_to_string": (C.c_int, [P(OzPosition), C.c_char_p, C.c_size_t]),
    "oz_transform_action": (C.c_int, [C.c_int, C.c_int]),
    "oz_movegen_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oz_flips_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oz_apply_action_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "oz_movegen_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "oz_perft": (C.c_int, [P(OzPosition), C.c_int, C.c_int, P(u64), P(C.c_double)]),
    "oz_debug_int_peak": (C.c_int, [P(C.c_double), P(C.c_double)]),
    "oz_features_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "oz_config_default": (None, [P(OzConfig)]),
    "oz_engine_create": (C.c_int, [P(OzConfig), P(C.c_void_p)]),
    "oz_engine_destroy": (None, [C.c_void_p]),
    "oz_engine_set_history_size": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_set_num_simulations": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_set_num_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_set_batch_size": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_set_c_puct_base": (C.c_int, [C.c_void_p, f32]),
    "oz_engine_set_c_puct_init": (C.c_int, [C.c_void_p, f32]),
    "oz_engine_set_dirichlet_epsilon": (C.c_int, [C.c_void_p, f32]),
    "oz_engine_set_dirichlet_alpha": (C.c_int, [C.c_void_p, f32]),
    "oz_engine_set_evaluator": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_get_config": (C.c_int, [C.c_void_p, P(OzConfig)]),
    "oz_engine_set_gamma_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "oz_engine_reset": (C.c_int, [C.c_void_p]),
    "oz_engine_positions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oz_resnet_blob_bytes": (C.c_size_t, [P(OzResnetShape)]),
    "oz_resnet_blob_offsets": (C.c_int, [P(OzResnetShape), P(i64), C.c_int, P(C.c_int)]),
    "oz_engine_load_resnet": (C.c_int, [C.c_void_p, P(OzResnetShape), C.c_void_p, C.c_size_t]),
    "oz_resnet_create": (C.c_int, [C.c_int, P(C.c_void_p)]),
    "oz_resnet_destroy": (None, [C.c_void_p]),
    "oz_resnet_load": (C.c_int, [C.c_void_p, P(OzResnetShape), C.c_void_p, C.c_size_t]),
    "oz_resnet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "oz_resnet_forward_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_debug_conv3x3": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_int, C.c_size_t, P(C.c_double)]),
    "oz_debug_set_tower": (C.c_int, [C.c_int]),
    "oz_debug_set_heads_boards": (C.c_int, [C.c_int]),
    "oz_debug_tower_timeline": (C.c_int, [C.c_void_p, P(C.c_int)]),
    "oz_debug_wgrad3x3": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, P(C.c_double)]),
    "oz_engine_search": (C.c_int, [C.c_void_p]),
    "oz_engine_search_range": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "oz_engine_search_begin": (C.c_int, [C.c_void_p]),
    "oz_engine_rounds": (C.c_int, [C.c_void_p]),
    "oz_engine_select_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_engine_expand_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_engine_search_end": (C.c_int, [C.c_void_p]),
    "oz_engine_root_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_engine_self_play_data": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "oz_engine_apply_actions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oz_engine_choose_actions": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oz_engine_play_ply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, P(i32)]),
    "oz_engine_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "oz_engine_training_samples": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "oz_training_samples_from_records": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_size_t, P(C.c_size_t)]),
    "oz_engine_counters": (C.c_int, [C.c_void_p, P(OzCounters)]),
    "oz_engine_set_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_set_pipeline": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "oz_engine_set_game_pool": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_finished_games": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "oz_engine_finished_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, P(C.c_int32)]),
    "oz_engine_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_engine_stream": (C.c_void_p, [C.c_void_p]),
    "oz_train_config_default": (None, [P(OzTrainConfig)]),
    "oz_trainer_create": (C.c_int, [C.c_int, P(OzResnetShape), C.c_int, P(C.c_void_p)]),
    "oz_trainer_destroy": (None, [C.c_void_p]),
    "oz_trainer_num_params": (C.c_size_t, [C.c_void_p]),
    "oz_trainer_num_buffers": (C.c_size_t, [C.c_void_p]),
    "oz_trainer_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_trainer_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oz_trainer_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, P(OzTrainConfig), C.c_int, C.c_void_p]),
    "oz_trainer_step_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, P(OzTrainConfig), C.c_int, C.c_void_p]),
    "oz_trainer_outputs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "oz_trainer_set_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "oz_trainer_stream": (C.c_void_p, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (python build_ext.py)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a library has not been built "
                "(run `python othello-alphazero_b200/build_ext.py`); there is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    """Raise what the reference's pybind11 layer would raise (othello_mcts.cpp, default translation)."""
    if status == OZ_OK:
        return
    msg = load().oz_last_error().decode("utf-8", "replace")
    if status == OZ_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if status == OZ_ERR_OUT_OF_RANGE:
        raise IndexError(msg)
    raise RuntimeError(msg)
