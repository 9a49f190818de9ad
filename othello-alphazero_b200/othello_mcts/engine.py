"""Device engine handles: `MCTS` (the reference's single-game surface) and `SelfPlayEngine`
(additive: many lockstep games per GPU).  Both sit on include/othello_b200.h via ctypes.

reference: cpp/src/lib/othello_mcts.cpp:89-150 (MCTS binding), cpp/src/include/mcts.h,
python/othello_alphazero/train.py:404-452 (_self_play).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, weights
from ._lib import (EVAL_EXTERNAL, EVAL_HASH, EVAL_RESNET, EVAL_SYNTH, MAX_CHILDREN, OzConfig, OzCounters, OzPlyRecord,
                   OzPosition, check)
from .position import Position

_BUILTIN = {"synth": EVAL_SYNTH, "hash": EVAL_HASH}
_RECORD_DTYPE = np.dtype([("p1", "<u8"), ("p2", "<u8"), ("player", "<i4"), ("action", "<i4"), ("n_children", "<i4"),
                          ("visit_counts", "<i4", (MAX_CHILDREN,)), ("_pad", "<i4")])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _Engine:
    """Shared plumbing over an oz_engine handle."""

    def __init__(self, num_games, history_size, torch_device, torch_pin_memory, num_simulations, num_threads,
                 batch_size, c_puct_base, c_puct_init, dirichlet_epsilon, dirichlet_alpha, seed=0, first_game_id=0,
                 sym_mode=_lib.SYM_RANDOM, temperature_plies=12, max_nodes_per_game=0, cuda_device=None,
                 self_play_temperature=1.0):
        self._lib = _lib.load()
        self._torch_device = str(torch_device)
        self._torch_pin_memory = bool(torch_pin_memory)
        cfg = OzConfig()
        self._lib.oz_config_default(C.byref(cfg))
        cfg.num_games = num_games
        cfg.history_size, cfg.num_simulations = history_size, num_simulations
        cfg.num_threads, cfg.batch_size = num_threads, batch_size
        cfg.c_puct_base, cfg.c_puct_init = c_puct_base, c_puct_init
        cfg.dirichlet_epsilon, cfg.dirichlet_alpha = dirichlet_epsilon, dirichlet_alpha
        cfg.seed, cfg.first_game_id = seed, first_game_id
        cfg.sym_mode, cfg.temperature_plies, cfg.max_nodes_per_game = sym_mode, temperature_plies, max_nodes_per_game
        cfg.self_play_temperature = self_play_temperature
        cfg.evaluator = EVAL_EXTERNAL
        if cuda_device is None:
            dev = torch.device(self._torch_device)
            cuda_device = dev.index if (dev.type == "cuda" and dev.index is not None) else 0
        cfg.device = cuda_device
        self._cuda_device = cuda_device
        self._h = C.c_void_p()
        check(self._lib.oz_engine_create(C.byref(cfg), C.byref(self._h)))
        self._evaluator = EVAL_EXTERNAL
        self._net_fingerprint = None
        self._gamma = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.oz_engine_destroy(h)
            self._h = None

    # -- configuration ------------------------------------------------------------------
    def _cfg(self) -> OzConfig:
        cfg = OzConfig()
        check(self._lib.oz_engine_get_config(self._h, C.byref(cfg)))
        return cfg

    def _set_evaluator(self, kind: int) -> None:
        if kind != self._evaluator:
            check(self._lib.oz_engine_set_evaluator(self._h, kind))
            self._evaluator = kind

    def set_gamma_table(self, table) -> None:
        """Test hook: inject the Dirichlet draws (see include/othello_b200.h)."""
        if table is None:
            check(self._lib.oz_engine_set_gamma_table(self._h, None, 0))
            self._gamma = None
        else:
            self._gamma = np.ascontiguousarray(table, dtype=np.float32)
            check(self._lib.oz_engine_set_gamma_table(self._h, _ptr(self._gamma), len(self._gamma)))

    def load_network(self, net) -> None:
        """Fold and upload an AlphaZeroNet if its weights changed since the last upload."""
        fp = weights.weights_fingerprint(net)
        if fp != self._net_fingerprint:
            shape, blob = weights.fold_cached(net, fp)
            check(self._lib.oz_engine_load_resnet(self._h, C.byref(shape), blob, len(blob)))
            self._net_fingerprint = fp

    def load_folded(self, shape, blob: bytes) -> None:
        """Upload an already folded weight blob (weights.fold_alphazero_net, parallel.broadcast_weights) and select the
        native evaluator; a later load_network(net) folds again only if `net` differs from what it last uploaded."""
        check(self._lib.oz_engine_load_resnet(self._h, C.byref(shape), blob, len(blob)))
        self._net_fingerprint = ("folded", hash(blob))
        self._set_evaluator(EVAL_RESNET)

    # -- search -------------------------------------------------------------------------
    def _search(self, neural_net) -> None:
        if isinstance(neural_net, str):
            self._set_evaluator(_BUILTIN[neural_net])
            check(self._lib.oz_engine_search(self._h))
            return
        if weights.is_alphazero_net(neural_net) and weights.natively_supported(neural_net) and \
                weights.shape_of(neural_net).in_planes == 1 + 2 * self._cfg().history_size:
            self.load_network(neural_net)
            self._set_evaluator(EVAL_RESNET)
            check(self._lib.oz_engine_search(self._h))
            return
        self._search_with_callable(neural_net)

    def _search_with_callable(self, neural_net) -> None:
        """The evaluator contract of othello_mcts.cpp:36-45: features [B,1+2H,8,8] float32 on
        torch_device -> {"policy": [B,65], "value": [B]}.  The tree stays on the GPU."""
        self._set_evaluator(EVAL_EXTERNAL)
        cfg = self._cfg()
        n = cfg.num_games * cfg.num_threads * cfg.batch_size
        planes = 1 + 2 * cfg.history_size
        cuda = torch.device("cuda", self._cuda_device)
        target = torch.device(self._torch_device)
        with torch.cuda.device(cuda), torch.no_grad():
            feats = torch.zeros((n, planes, 8, 8), dtype=torch.float32, device=cuda)
            stream = C.c_void_p(torch.cuda.current_stream(cuda).cuda_stream)
            check(self._lib.oz_engine_search_begin(self._h))
            try:
                for _ in range(self._lib.oz_engine_rounds(self._h)):
                    check(self._lib.oz_engine_select_device(self._h, C.c_void_p(feats.data_ptr()), stream))
                    x = feats if target.type == "cuda" else feats.to(target)
                    out = neural_net(x)
                    policy = out["policy"].detach().to(cuda, torch.float32).contiguous()
                    value = out["value"].detach().to(cuda, torch.float32).contiguous()
                    if policy.shape != (n, 65) or value.shape != (n,):
                        raise ValueError(f"evaluator returned policy {tuple(policy.shape)} / value {tuple(value.shape)}, "
                                         f"expected ({n}, 65) / ({n},)")
                    check(self._lib.oz_engine_expand_device(self._h, C.c_void_p(policy.data_ptr()),
                                                            C.c_void_p(value.data_ptr()), stream))
            finally:
                check(self._lib.oz_engine_search_end(self._h))

    # -- read-out -----------------------------------------------------------------------
    def _positions(self) -> list[Position]:
        g = self._cfg().num_games
        raw = (OzPosition * g)()
        check(self._lib.oz_engine_positions(self._h, raw))
        out = []
        for r in raw:
            c = OzPosition()
            C.memmove(C.byref(c), C.byref(r), C.sizeof(OzPosition))
            out.append(Position._wrap(c))
        return out

    def _root_stats(self):
        g = self._cfg().num_games
        n = np.zeros(g, np.int32)
        counts = np.zeros((g, MAX_CHILDREN), np.int32)
        q = np.zeros((g, MAX_CHILDREN), np.float32)
        rootn = np.zeros(g, np.int32)
        check(self._lib.oz_engine_root_stats(self._h, _ptr(n), _ptr(counts), _ptr(q), _ptr(rootn)))
        return n, counts, q, rootn

    def counters(self) -> dict:
        c = OzCounters()
        check(self._lib.oz_engine_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in OzCounters._fields_}

    def set_profiling(self, on: bool) -> None:
        check(self._lib.oz_engine_set_profiling(self._h, int(on)))

    def set_graphs(self, mode: int) -> None:
        """CUDA-graph replay of search rounds: 0 never, 1 (default) for launch-bound leaf batches, 2 always."""
        check(self._lib.oz_engine_set_graphs(self._h, int(mode)))

    def set_pipeline(self, groups: int, pdl: bool = True) -> None:
        """Pipelined search: split the games into `groups` groups on their own streams so one group's tree kernels run
        underneath another group's convolutions (graph-replayed searches: parallel branches of the graph).  0 = automatic,
        1 = off, 2..8.  Results do not depend on it."""
        check(self._lib.oz_engine_set_pipeline(self._h, int(groups), int(pdl)))


class MCTS(_Engine):
    """Drop-in for the reference's `othello_mcts.MCTS` (othello_mcts.cpp:89-150): one game, same
    constructor keywords, defaults, validation messages and method names.  `search` accepts the same
    callable; an (eval-mode) AlphaZeroNet on a CUDA torch_device runs on the native evaluator."""

    def __init__(self, history_size: int = 4, torch_device: str = "cpu", torch_pin_memory: bool = False,
                 num_simulations: int = 800, num_threads: int = 2, batch_size: int = 16, c_puct_base: float = 20000.0,
                 c_puct_init: float = 2.5, dirichlet_epsilon: float = 0.25, dirichlet_alpha: float = 0.5, *,
                 seed: int = 0, game_id: int = 0, sym_mode: int = _lib.SYM_RANDOM) -> None:
        super().__init__(1, history_size, torch_device, torch_pin_memory, num_simulations, num_threads, batch_size,
                         c_puct_base, c_puct_init, dirichlet_epsilon, dirichlet_alpha, seed=seed, first_game_id=game_id,
                         sym_mode=sym_mode)

    def reset_position(self) -> None:
        check(self._lib.oz_engine_reset(self._h))

    def position(self) -> Position:
        return self._positions()[0]

    def search(self, neural_net) -> None:
        self._search(neural_net)

    def visit_counts(self) -> list[int]:
        n, counts, _, _ = self._root_stats()
        return [int(x) for x in counts[0, : n[0]]]

    def mean_action_values(self) -> list[float]:
        n, _, q, _ = self._root_stats()
        return [float(x) for x in q[0, : n[0]]]

    def self_play_data(self) -> dict:
        planes = 1 + 2 * self._cfg().history_size
        f = np.zeros((8, planes, 8, 8), np.float32)
        p = np.zeros((8, 65), np.float32)
        check(self._lib.oz_engine_self_play_data(self._h, 0, _ptr(f), _ptr(p)))
        # lists of 8 CPU tensors, as mcts.h:23-31 / train.py:433-434
        return {"features": [torch.from_numpy(f[t].copy()) for t in range(8)],
                "policy": [torch.from_numpy(p[t].copy()) for t in range(8)]}

    def apply_action(self, action: int) -> None:
        a = np.array([int(action)], np.int32)
        check(self._lib.oz_engine_apply_actions(self._h, _ptr(a)))

    # getter / setter pairs (othello_mcts.cpp:131-150)
    def history_size(self) -> int: return self._cfg().history_size
    def set_history_size(self, value: int) -> None: check(self._lib.oz_engine_set_history_size(self._h, int(value)))
    def torch_device(self) -> str: return self._torch_device
    def set_torch_device(self, value: str) -> None: self._torch_device = str(value)
    def torch_pin_memory(self) -> bool: return self._torch_pin_memory
    def set_torch_pin_memory(self, value: bool) -> None: self._torch_pin_memory = bool(value)
    def num_simulations(self) -> int: return self._cfg().num_simulations
    def set_num_simulations(self, value: int) -> None: check(self._lib.oz_engine_set_num_simulations(self._h, int(value)))
    def num_threads(self) -> int: return self._cfg().num_threads
    def set_num_threads(self, value: int) -> None: check(self._lib.oz_engine_set_num_threads(self._h, int(value)))
    def batch_size(self) -> int: return self._cfg().batch_size
    def set_batch_size(self, value: int) -> None: check(self._lib.oz_engine_set_batch_size(self._h, int(value)))
    def c_puct_base(self) -> float: return self._cfg().c_puct_base
    def set_c_puct_base(self, value: float) -> None: check(self._lib.oz_engine_set_c_puct_base(self._h, float(value)))
    def c_puct_init(self) -> float: return self._cfg().c_puct_init
    def set_c_puct_init(self, value: float) -> None: check(self._lib.oz_engine_set_c_puct_init(self._h, float(value)))
    def dirichlet_epsilon(self) -> float: return self._cfg().dirichlet_epsilon
    def set_dirichlet_epsilon(self, value: float) -> None: check(self._lib.oz_engine_set_dirichlet_epsilon(self._h, float(value)))
    def dirichlet_alpha(self) -> float: return self._cfg().dirichlet_alpha
    def set_dirichlet_alpha(self, value: float) -> None: check(self._lib.oz_engine_set_dirichlet_alpha(self._h, float(value)))


class SelfPlayEngine(_Engine):
    """Additive surface: `num_games` self-play games in lockstep on one GPU.
    Game g uses RNG streams keyed by `first_game_id + g`, so results do not depend on how games are
    partitioned over GPUs."""

    def __init__(self, num_games: int, history_size: int = 4, torch_device: str = "cuda", torch_pin_memory: bool = False,
                 num_simulations: int = 800, num_threads: int = 2, batch_size: int = 16, c_puct_base: float = 20000.0,
                 c_puct_init: float = 2.5, dirichlet_epsilon: float = 0.25, dirichlet_alpha: float = 0.5, *, seed: int = 0,
                 first_game_id: int = 0, sym_mode: int = _lib.SYM_RANDOM, temperature_plies: int = 12,
                 max_nodes_per_game: int = 0, cuda_device: int | None = None, self_play_temperature: float = 1.0) -> None:
        # self_play_temperature: train.py:423 (--self-play-temperature): sampling weights visit_counts ** (1 / T) while
        # ply < temperature_plies
        super().__init__(num_games, history_size, torch_device, torch_pin_memory, num_simulations, num_threads, batch_size,
                         c_puct_base, c_puct_init, dirichlet_epsilon, dirichlet_alpha, seed=seed, first_game_id=first_game_id,
                         sym_mode=sym_mode, temperature_plies=temperature_plies, max_nodes_per_game=max_nodes_per_game,
                         cuda_device=cuda_device, self_play_temperature=self_play_temperature)
        self.num_games = num_games

    def reset(self) -> None:
        check(self._lib.oz_engine_reset(self._h))

    def positions(self) -> list[Position]:
        return self._positions()

    def positions_array(self) -> np.ndarray:
        """Root positions of all games as one structured array (batch.POSITION_DTYPE: player, p1, p2, legal, next_legal)."""
        from .batch import POSITION_DTYPE
        out = np.zeros(self.num_games, POSITION_DTYPE)
        check(self._lib.oz_engine_positions(self._h, _ptr(out)))
        return out

    def search(self, neural_net) -> None:
        self._search(neural_net)

    def root_stats(self):
        """(n_children [G], visit_counts [G,36], mean_action_values [G,36], root_visit_count [G])."""
        return self._root_stats()

    def choose_actions(self) -> np.ndarray:
        a = np.zeros(self.num_games, np.int32)
        check(self._lib.oz_engine_choose_actions(self._h, _ptr(a)))
        return a

    def apply_actions(self, actions) -> None:
        a = np.ascontiguousarray(actions, dtype=np.int32)
        check(self._lib.oz_engine_apply_actions(self._h, _ptr(a)))

    def prepare(self, neural_net) -> None:
        """Select the evaluator for play_ply: 'synth' / 'hash' or an AlphaZeroNet."""
        if isinstance(neural_net, str):
            self._set_evaluator(_BUILTIN[neural_net])
        else:
            self.load_network(neural_net)
            self._set_evaluator(EVAL_RESNET)

    def play_ply(self, want_counts: bool = False):
        """One ply for every unfinished game (search, choose, record, apply).  Returns
        (actions [G], visit_counts [G,36] or None, games still running)."""
        actions = np.zeros(self.num_games, np.int32)
        counts = np.zeros((self.num_games, MAX_CHILDREN), np.int32) if want_counts else None
        unfinished = C.c_int32()
        check(self._lib.oz_engine_play_ply(self._h, _ptr(actions), _ptr(counts) if want_counts else None,
                                           C.byref(unfinished)))
        return actions, counts, unfinished.value

    def set_game_pool(self, total_games: int) -> None:
        """Rolling game pool: the `num_games` slots play `total_games` games (ids first_game_id ..); a finished game
        moves to the store of completed games and its slot restarts with the next unplayed id.  0 switches it off.
        Resets every game."""
        check(self._lib.oz_engine_set_game_pool(self._h, int(total_games)))

    def finished_games(self) -> int:
        n = C.c_int32()
        check(self._lib.oz_engine_finished_games(self._h, C.byref(n)))
        return n.value

    def finished_records(self, max_plies: int = 128) -> dict:
        """Completed games of the rolling pool, in completion order: {"game_ids", "plies", "records"}."""
        n = self.finished_games()
        ids = np.zeros(n, np.uint32)
        plies = np.zeros(n, np.int32)
        recs = (OzPlyRecord * max(1, n * max_plies))()
        cnt = C.c_int32()
        check(self._lib.oz_engine_finished_records(self._h, _ptr(ids), _ptr(plies), recs, max_plies, n, C.byref(cnt)))
        arr = np.frombuffer(recs, dtype=_RECORD_DTYPE)[: n * max_plies].reshape(n, max_plies)
        return {"game_ids": ids, "plies": plies, "records": arr}

    def run(self, neural_net, total_games: int | None = None, max_plies: int = 100000) -> dict:
        """Play to the end.  Without `total_games`: the `num_games` lockstep games, returns records().  With
        `total_games` >= num_games: rolling pool, returns finished_records() (completion order, with game ids)."""
        if total_games is not None:
            self.set_game_pool(total_games)
        self.prepare(neural_net)
        for _ in range(max_plies):
            _, _, unfinished = self.play_ply()
            if unfinished == 0:
                break
        return self.finished_records() if total_games is not None else self.records()

    def records(self, max_plies: int = 128) -> dict:
        plies = np.zeros(self.num_games, np.int32)
        recs = (OzPlyRecord * (self.num_games * max_plies))()
        check(self._lib.oz_engine_records(self._h, _ptr(plies), recs, max_plies))
        arr = np.frombuffer(recs, dtype=_RECORD_DTYPE).reshape(self.num_games, max_plies)
        return {"plies": plies, "records": arr}

    def training_samples(self) -> dict:
        """The tensors `_self_play` returns (train.py:432-450) for all finished games, 8 symmetries per ply."""
        n = C.c_size_t()
        check(self._lib.oz_engine_training_samples(self._h, None, None, None, 0, C.byref(n)))
        planes = 1 + 2 * self._cfg().history_size
        f = np.zeros((n.value, planes, 8, 8), np.float32)
        p = np.zeros((n.value, 65), np.float32)
        v = np.zeros((n.value,), np.float32)
        if n.value:
            check(self._lib.oz_engine_training_samples(self._h, _ptr(f), _ptr(p), _ptr(v), n.value, C.byref(n)))
        return {"features": torch.from_numpy(f), "policies": torch.from_numpy(p), "values": torch.from_numpy(v)}
