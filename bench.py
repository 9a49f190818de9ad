#!/usr/bin/env python
"""Headline benchmark: MCTS simulations/s of lockstep Othello self-play on the 128x10b network,
800 simulations per move (BASELINE.json metric), on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our engine (one process per GPU)
  python bench.py --impl reference ...                           the reference's C++ MCTS + LibTorch on host cores

A "step" is one self-play ply for every lockstep game on the GPU: one 800-simulation search per game
(select -> encode -> ResNet -> expand/backup rounds), action choice, record, re-root.  `value` counts
simulations the way the reference does (ceil(N/(T*B))*T*B per search, search_thread.cpp:48-50), with
trees, weights and all inputs resident in HBM.  `e2e` drives the same plies through the
reference-shaped API (search / visit counts to the host / host-side action choice / apply_action from
host buffers), copies inside the timed region.  Timing: CUDA events on the engine stream, barrier +
synchronize on both sides, max over ranks.  Rank 0 prints one JSON line.

Beside the headline (BASELINE.json configs[2], 1 000 lockstep games per GPU = weak scaling) the same line carries two
sub-records so that the driver's own runs hold BASELINE configs 4 and 5 (skip them with --no-extra-configs):
  "config4": 8 192 lockstep games partitioned over the N GPUs (8192/N per GPU; N = 1 is the strong-scaling denominator),
  "config5": 256x20b, 1 600 sims/move, leaf batch 64 ... 4096 per network call and GPU, on the N GPUs,
  "config3_leaf_batch_256": the headline network with 8 games per GPU, i.e. configs[2]'s "leaf batch 256" read literally.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT / "othello-alphazero_b200", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

NET = dict(in_channels=17, conv_channels=128, num_residual_blocks=9, value_head_hidden_channels=128)  # 128x10b, H = 8
MCTS_KW = dict(history_size=8, num_simulations=800, num_threads=2, batch_size=16, c_puct_base=20000.0, c_puct_init=2.5,
               dirichlet_epsilon=0.25, dirichlet_alpha=0.5)


# dram__bytes_read.sum + dram__bytes_write.sum of a conv launch per board, from one ncu --set full capture of a 128 -> 128
# layer over 32768 boards (profiles/r01_conv_final_kernel.json, same figures as the earlier r01_conv_pair_kernel.json): plain layer
# 0.537 + 0.496 GB, residual layer 1.074 + 0.496 GB.
NCU_TRAFFIC_BYTES_PER_BOARD = (0.537230e9 + 0.498725e9 + 1.074230e9 + 0.494879e9) / 2 / 32768


def flops_per_eval(cin=17, c=128, blocks=9, hv=128):
    """Dense 2*MAC convention of SURVEY.md section 8(d) (= torch.utils.flop_counter)."""
    conv = 2 * 576 * c * (cin + 2 * blocks * c)
    heads = 2 * (c * 2 * 64 + 128 * 65) + 2 * (c * 64 + 64 * hv + hv)
    return conv + heads, conv


# The kernel legitimately skips the tile-taps that only see the zero rows -1 / 8 (6 of 72 per board); the dense FLOP
# convention of SURVEY.md 8(d) still counts them, so the issued tensor work is this fraction of the dense figure.
ISSUED_MMA_FRACTION = 66.0 / 72.0
# ALU-pipe warp instructions per perft leaf of dfs_kernel<3> at depth 11, from the ncu capture profiles/r02_perft_dfs_ncu.csv:
# smsp__inst_executed_pipe_alu.sum = 7 129 723 840 for 212 258 800 leaves (8 146 791 858 warp instructions in all; the walk
# diverges, 38 % of the lane slots carry a thread, so the pipe is counted in warp instructions, as ncu does).
PERFT_ALU_WARP_INST_PER_NODE = 7129723840 / 212258800


def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        # "under load" = samples whose power draw is at least half of the highest one seen: the sampler also runs while
        # the ranks wait for each other at the barriers around the timed region
        def _f(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        pmax = max((_f(r[3]) for r in rows), default=0.0)
        loaded = [r for r in rows if _f(r[3]) >= 0.5 * pmax] or rows
        sm = sorted(float(r[1]) for r in loaded)
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(rows), "samples_under_load": len(sm), "power_w_max": pmax}


# ----------------------------------------------------------------------------------------------
def make_net(net_kw=None, reference_class=False):
    """Random-init network (torch.manual_seed(0), default PyTorch init, eval mode).  Our arm builds the repo's class
    (pinned bit for bit to the reference class by tests/test_net_golden.py); the reference arm builds the reference's own
    class from oracle/_ref/python when build() has placed it there."""
    import torch
    AlphaZeroNet = None
    if reference_class and (ROOT / "oracle" / "_ref" / "python" / "othello_alphazero" / "neural_net.py").exists():
        sys.path.insert(0, str(ROOT / "oracle" / "_ref" / "python"))
        try:
            from othello_alphazero.neural_net import AlphaZeroNet
        except Exception:
            AlphaZeroNet = None
    if AlphaZeroNet is None:
        from othello_mcts.neural_net import AlphaZeroNet
    torch.manual_seed(0)
    return AlphaZeroNet(**(net_kw or NET)).eval()


def workload_config(args, world: int) -> dict:
    """The workload both arms are quoted on (BASELINE.json configs[2], per GPU)."""
    G, inflight = args.games, args.threads * args.batch
    sims_per_search = -(-args.sims // inflight) * inflight
    return {"workload": f"{args.channels}x{args.blocks + 1}b random-init ResNet self-play, {G} lockstep games per GPU, {args.sims} sims/move "
                        f"(T={args.threads} x B={args.batch} leaves in flight per game, leaf batch {G * inflight} per network call), "
                        "random-symmetry eval, Dirichlet eps 0.25 alpha 0.5",
            "games_per_gpu": G, "sims_per_move": sims_per_search, "leaf_batch": G * inflight,
            "l2": "per-step working set (activations 16 KB/board/layer x leaf batch) far exceeds the 126 MB L2",
            "parallelism": f"games partitioned by id over {world} GPU(s), no collective on the hot path"}


def reference_search_rate(steps: int, warmup: int, threads: int):
    """BASELINE.json config 1: the reference's C++ MCTS (or the oracle port) + LibTorch CPU evaluation of
    the random-init 128x10b net, 800 sims/move, T=2 x B=16, on the host cores.  Each step = one search."""
    import numpy as np
    import torch
    torch.set_num_threads(threads)
    net = make_net(reference_class=True)
    kind = "port"
    mcts = None
    ref_dir = ROOT / "oracle" / "_ref"
    if (ref_dir / "othello_mcts" / "__init__.py").exists():
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ref_othello_mcts", ref_dir / "othello_mcts" / "__init__.py",
                                                          submodule_search_locations=[str(ref_dir / "othello_mcts")])
            mod = importlib.util.module_from_spec(spec)
            sys.modules["ref_othello_mcts"] = mod
            spec.loader.exec_module(mod)
            mcts = mod.MCTS(torch_device="cpu", torch_pin_memory=False, **MCTS_KW)
            kind = "reference"
        except Exception:
            mcts = None
    with torch.no_grad():
        if mcts is not None:
            def search():
                mcts.search(net)
            def advance():
                vc = np.array(mcts.visit_counts(), dtype=np.float64)
                acts = mcts.position().legal_actions()
                mcts.apply_action(int(np.random.choice(acts, p=vc / vc.sum())))
                if mcts.position().is_terminal():
                    mcts.reset_position()
        else:
            import pyoracle as po
            def cb(x):
                out = net(torch.from_numpy(x.copy()))
                return out["policy"].numpy(), out["value"].numpy()
            kw = {k: v for k, v in MCTS_KW.items()}
            om = po.OracleMCTS(evaluator=cb, seed=0, game_id=0, **kw)
            def search():
                om.search()
            def advance():
                om.apply_action(om.choose_action(0, 1 << 30))
                if om.position().player == 0:
                    om.reset_position()
        np.random.seed(0)
        for _ in range(warmup):
            search(); advance()
        t = 0.0
        for _ in range(steps):
            t0 = time.perf_counter(); search(); t += time.perf_counter() - t0
            advance()
    sims = steps * 800
    return sims / t, t, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(args.warmup, 3)           # same --steps / --warmup handling as our arm
    rate, t, kind = reference_search_rate(steps, warmup, cores)
    line = {"metric": "mcts_simulations_per_second", "value": rate, "unit": "sims/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, max(1, args.gpus)),
            "reference_sample": "the reference plays the games one at a time (train.py:394-398): each step = one 800-simulation "
                                "search of one game, C++ MCTS (T=2 x B=16) + LibTorch CPU on all host cores",
            "cpu_baseline": {"value": rate, "unit": "sims/s", "cores": cores, "kind": kind,
                             "sample": f"{steps} searches x 800 sims, one game, after {warmup} warm-up searches"},
            "e2e": {"value": rate, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def perft_cpu_baseline(cores: int) -> dict | None:
    """perft through the REFERENCE's own Position code (position.h:202-262 behind
    oracle/_ref/libref_shim.so; the oracle's C port if that was never built) on the host: depth 10 on one thread, depth 11
    over a depth-4 frontier (244 positions) on `cores` threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    try:
        import pyoracle as po
        lib, pre, kind = po.ref_shim(), "ref_", "reference"
        if lib is None:
            lib, pre, kind = po.lib(), "ozo_", "port"
        fn = lambda name: getattr(lib, pre + name)
        start = po.Pos()
        fn("initial_position")(C.byref(start))
        t0 = time.perf_counter()
        n1 = int(fn("perft")(C.byref(start), 10))
        t1 = time.perf_counter() - t0
        frontier = [start]
        for _ in range(4):
            nxt = []
            for p in frontier:
                acts = (C.c_int * 40)()
                for a in list(acts[: fn("legal_actions")(C.byref(p), acts)]):
                    q = po.Pos()
                    fn("apply_action")(C.byref(p), a, C.byref(q))
                    nxt.append(q)
            frontier = nxt
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as pool:
            nN = sum(pool.map(lambda q: int(fn("perft")(C.byref(q), 7)), frontier))
        tN = time.perf_counter() - t0
        return {"kind": kind, "unit": "nodes/s", "one_thread": {"depth": 10, "nodes": n1, "nodes_per_s": n1 / t1, "seconds": t1},
                "all_threads": {"depth": 11, "nodes": nN, "threads": cores, "nodes_per_s": nN / tN, "seconds": tN,
                                "split": f"{len(frontier)} depth-4 frontier positions"},
                "value": nN / tN, "cores": cores,
                "sample": "reference Position code (position.h) compiled -O3: perft(10) on 1 thread, perft(11) over the depth-4 frontier on all threads"}
    except Exception as exc:  # the baseline is reported beside the GPU number, never a reason to lose the line
        return {"error": repr(exc)}


def run_ours(args):
    import numpy as np
    import torch
    import othello_mcts as om
    from othello_mcts import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keeps NCCL's banner off stdout: one JSON line only
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    G = args.games
    kw = dict(MCTS_KW)
    kw.update(num_simulations=args.sims, num_threads=args.threads, batch_size=args.batch)
    inflight = args.threads * args.batch
    sims_per_search = -(-args.sims // inflight) * inflight
    net_kw = dict(NET, conv_channels=args.channels, num_residual_blocks=args.blocks, value_head_hidden_channels=args.hidden)
    net = make_net(net_kw)
    lib = _lib.load()
    steps, warmup = max(1, args.steps), max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(eng, fn, n_steps):
        """CUDA events on the engine's stream around n_steps calls, barrier + synchronize on both sides, max over ranks."""
        stream = torch.cuda.ExternalStream(lib.oz_engine_stream(eng._h), device=torch.device("cuda", local))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n_steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        timed.local_ms = ms                                       # this rank's own time (per-rank breakdown below)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def allgather(*vals):
        """[rank][i]: every rank's values (one small all_gather outside the timed region)."""
        t = torch.tensor([float(v) for v in vals], device="cuda", dtype=torch.float64)
        if dist is None:
            return [t.tolist()]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.tolist() for o in out]

    def allsum(*vals):
        t = torch.tensor([float(v) for v in vals], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t)
        return [float(x) for x in t.tolist()]

    # games are partitioned contiguously by id: rank r owns [r*G, (r+1)*G); no collective on the hot path
    eng = om.SelfPlayEngine(G, seed=args.seed, first_game_id=rank * G, cuda_device=local, **kw)
    eng.prepare(net)
    eng.set_profiling(not args.no_profiling)
    eng.set_pipeline(args.pipeline, bool(args.pipeline_pdl))

    def dev_step():
        eng.play_ply(want_counts=False)

    rng = np.random.default_rng(args.seed + rank)
    h2d = d2h = 0
    sq = np.arange(64, dtype=np.uint64)

    def e2e_step():
        """search -> visit counts to the host -> action choice on the host (train.py:421-430, tau = 1: proportional to
        the visit counts) -> apply_action from a host array.  The choice is vectorised over games with NumPy."""
        nonlocal h2d, d2h
        eng.search(net)
        n, counts, _, _ = eng._root_stats()                       # D2H: visit counts, child counts, Q, root N
        pos = eng.positions_array()                               # D2H: root positions
        c = counts.astype(np.float64)
        tot_c = c.sum(axis=1)
        pick = (rng.random(G) * tot_c)[:, None]
        j = np.minimum((np.cumsum(c, axis=1) <= pick).sum(axis=1), np.maximum(n - 1, 0))     # sampled child index
        bits = ((pos["legal"][:, None] >> (np.uint64(63) - sq)[None, :]) & np.uint64(1)).astype(np.int32)   # [G,64] by action
        rk = np.cumsum(bits, axis=1) * bits                       # 1-based rank of every legal action
        acts = np.argmax(rk == (j + 1)[:, None], axis=1).astype(np.int32)
        acts[pos["legal"] == 0] = 64                              # pass
        acts[(n == 0) | (pos["player"] == 0)] = -1                # finished games
        eng.apply_actions(acts)                                   # H2D: actions (+ D2H of positions for validation)
        h2d = G * 4
        d2h = G * (4 + 36 * 4 + 36 * 4 + 4) + 2 * G * 40

    for _ in range(warmup):
        dev_step()
    c0 = eng.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(eng, dev_step, steps)
    ms_local = timed.local_ms
    clocks = sampler.stop() if rank == 0 else None
    c1 = eng.counters()
    sims = (c1["simulations"] - c0["simulations"])
    evals = c1["evaluations"] - c0["evaluations"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    eval_ms = c1["eval_ms"] - c0["eval_ms"]
    conv_ms = c1["conv_ms"] - c0["conv_ms"]
    conv_launches = c1["conv_launches"] - c0["conv_launches"]
    tot = allsum(sims, evals, launches)                           # the only collective: final counters
    # every rank's own step time and phases: `ms` is the slowest rank's, the phase counters below are rank 0's, so at
    # N > 1 the difference between the two (GPUs under the power cap settle at different clocks) must not be read as host time
    per_rank = allgather(ms_local, eval_ms, conv_ms, c1["select_ms"] - c0["select_ms"], c1["expand_ms"] - c0["expand_ms"])
    value = tot[0] / (ms * 1e-3)

    # Same device-resident step measured in the middle game: every game is fast-forwarded to ply `--midgame-ply` with
    # full-width searches on the built-in hash evaluator (tree kernels only), then the network is switched back.
    mid = None
    if args.midgame_ply > 0:
        eng.reset()
        eng.prepare("hash")
        for _ in range(args.midgame_ply):
            eng.play_ply(want_counts=False)
        eng.prepare(net)
        dev_step()
        m0 = eng.counters()
        ms_mid = timed(eng, dev_step, 2)
        m1 = eng.counters()
        mid_sims = allsum(m1["simulations"] - m0["simulations"])[0]
        mid = {"ply": args.midgame_ply + 1, "steps": 2, "value": mid_sims / (ms_mid * 1e-3), "unit": "sims/s",
               "ms_per_step": ms_mid / 2, "select_ms_per_step": (m1["select_ms"] - m0["select_ms"]) / 2,
               "expand_ms_per_step": (m1["expand_ms"] - m0["expand_ms"]) / 2}
        eng.reset()
        eng.prepare(net)
        for _ in range(2):
            dev_step()

    # end to end through the reference-shaped calls with host buffers, over ALL --steps plies
    e2e_step()                                                    # host-path warm-up (NumPy buffers, pinned staging)
    ms_e2e = timed(eng, e2e_step, steps)
    active_sims = steps * G * sims_per_search * world
    e2e_value = active_sims / (ms_e2e * 1e-3)

    groups = max(1, args.pipeline)
    sel_ms, exp_ms = (c1["select_ms"] - c0["select_ms"]) / steps, (c1["expand_ms"] - c0["expand_ms"]) / steps
    if groups > 1:     # phases of different game groups overlap: network / conv = time at least one group spent there
        phases = {"pipeline_groups": groups, "network (union over groups)": eval_ms / steps, "of_which_conv": conv_ms / steps,
                  "select_encode (sum over groups, runs under the other group's network)": sel_ms,
                  "expand_backup (sum over groups, runs under the other group's network)": exp_ms,
                  "step time with no network running": (ms - eval_ms) / steps}
    else:
        phases = {"pipeline_groups": 1, "select_encode": sel_ms, "network": eval_ms / steps, "of_which_conv": conv_ms / steps,
                  "expand_backup": exp_ms,
                  "rest_of_ply (choose, record, re-root, syncs)": (ms_local - eval_ms) / steps - sel_ms - exp_ms}
        if world > 1:
            phases["note"] = ("rank 0's own phases and own step time; ms_per_step above is the slowest rank's; per_rank lists "
                              "every rank's step time and phases (CUDA events on that rank's engine stream)")
    per_rank_rec = {"ms_per_step": [r[0] / steps for r in per_rank], "network": [r[1] / steps for r in per_rank],
                    "of_which_conv": [r[2] / steps for r in per_rank], "select_encode": [r[3] / steps for r in per_rank],
                    "expand_backup": [r[4] / steps for r in per_rank],
                    "rest_of_ply": [(r[0] - r[1] - r[3] - r[4]) / steps for r in per_rank]}
    # Tree traffic: algorithmic bytes of select / encode / expand from the engine's counters, per simulation:
    # select reads a (N, W, Q, P, meta, child) row = 24 B per child per level below the root (the root's children once per
    # round), then writes the virtual loss along the path (N, W, Q and the path entry = 16 B per node); expand reads the leaf (20 B), writes 44 B
    # per new node (+ 4 B prior gathered from the 260 B policy row), and backs up along the path (read N, W; write W, Q;
    # path entry = 20 B per node); encode gathers 16 B per history position and writes the 4 096 B bf16 NHWC input.
    d_sims = max(1, c1["simulations"] - c0["simulations"])
    d_path, d_child, d_nodes = c1["path_nodes"] - c0["path_nodes"], c1["children_read"] - c0["children_read"], c1["nodes_expanded"] - c0["nodes_expanded"]
    planes_bytes = 64 * 32 * 2                                   # the network's bf16 NHWC input, written by encode_kernel
    tree_bytes = 24 * d_child + 16 * d_path + 20 * d_path + 48 * d_nodes + d_sims * (20 + 4 + 16 * MCTS_KW["history_size"] + planes_bytes)
    tree_ms = (c1["select_ms"] - c0["select_ms"]) + (c1["expand_ms"] - c0["expand_ms"])

    del eng                                                       # free the arenas before the sub-records allocate theirs
    extra = {}
    if not args.no_extra_configs:
        extra["config4"] = run_config4(args, om, net, kw, rank, world, local, timed, allsum)
        extra["config5"] = run_config5(args, om, rank, world, local, timed, allsum)
        extra["config3_leaf_batch_256"] = run_config3_leaf256(args, om, net, kw, rank, world, local, timed, allsum)

    if rank == 0:
        f_all, f_conv = flops_per_eval(c=args.channels, blocks=args.blocks, hv=args.hidden)
        pk = peaks()
        nn_share = eval_ms / ms if ms > 0 else None
        # dominant kernel: the 3x3 convolution.  Algorithmic FLOPs of this rank's conv launches in the timed region
        # (dense 2*MAC, zero-padding taps counted: SURVEY.md 8d) / their CUDA-event time on the engine stream.
        achieved = (evals * f_conv) / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
        frac = (achieved / pk["bf16_sustained"]) if achieved else None
        # best of 3: the event window of oz_perft spans the host round trips between its breadth-first levels
        perft_nodes, perft_ms = min((om.perft(11, return_ms=True) for _ in range(3)), key=lambda r: r[1])
        perft_rate = perft_nodes / (perft_ms * 1e-3) if perft_ms > 0 else None
        int_ops, int_ms = C.c_double(), C.c_double()
        _lib.check(lib.oz_debug_int_peak(C.byref(int_ops), C.byref(int_ms)))
        cores = os.cpu_count() or 1
        line = {
            "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": steps, "ms_per_step": ms_e2e / steps,
                    "path": "SelfPlayEngine.search(net) -> root_stats (host) -> numpy action choice -> apply_actions (host buffers)"},
            "gpu_launches": int(tot[2]),
            "nn_evals_per_second": tot[1] / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": frac, "traffic": NCU_TRAFFIC_BYTES_PER_BOARD * G * inflight if args.channels == 128 else None,
                         "kernel": "conv3x3_pair_kernel (tcgen05.mma cta_group::2 implicit GEMM)",
                         "flop_per_launch": (evals * f_conv / conv_launches) if conv_launches else None,
                         "us_per_launch": (1e3 * conv_ms / conv_launches) if conv_launches else None,
                         "launches_timed": int(conv_launches), "conv_share_of_step": conv_ms / ms if ms > 0 else None,
                         "nn_share_of_step": nn_share, "frac_of_burst_peak": (achieved / pk["bf16_burst"]) if achieved else None,
                         "issued_mma_fraction": ISSUED_MMA_FRACTION,
                         "frac_issued_work": frac * ISSUED_MMA_FRACTION if frac else None,
                         "issued_note": "the dense convention (SURVEY 8d) counts the 6 of 72 tile-taps per board the kernel skips because they "
                                        "only see the zero rows -1 / 8; frac_issued_work = frac x 66/72 is the tensor work actually issued",
                         "whole_step_frac": (tot[1] / world * f_all) / (ms * 1e-3) / 1e12 / pk["bf16_sustained"] if ms > 0 else None,
                         "traffic_source": "constant from the stored ncu --set full capture profiles/r01_conv_final_kernel.json "
                                           "(dram read+write bytes per launch, mean of a plain and a residual 128->128 layer, scaled to this "
                                           "leaf batch), not measured in this run; algorithmic: 32 KB / 48 KB per board",
                         "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)"},
            "perft": {"depth": 11, "nodes": int(perft_nodes), "nodes_per_s": perft_rate, "device_ms": perft_ms,
                      "roofline": {"bound": "integer ALU pipe issue (LOP3 / SHF)", "unit": "G warp-inst/s",
                                   "achieved": perft_rate * PERFT_ALU_WARP_INST_PER_NODE / 1e9 if perft_rate else None,
                                   "peak": int_ops.value / 32 / 1e9,
                                   "frac": perft_rate * PERFT_ALU_WARP_INST_PER_NODE / (int_ops.value / 32) if perft_rate and int_ops.value else None,
                                   "alu_warp_inst_per_node": PERFT_ALU_WARP_INST_PER_NODE,
                                   "peak_source": "measured in this run: oz_debug_int_peak (8 independent LOP3 + SHF chains per thread, all lanes "
                                                  f"active, {int_ms.value:.3f} ms) = 16 lanes per clock and SM sub-partition; instructions per node from the "
                                                  "ncu capture profiles/r02_perft_dfs_ncu.csv (ncu itself: ALU pipe 95.8 % of peak inside dfs_kernel); the "
                                                  "event window of oz_perft also holds the breadth-first frontier levels in front of dfs_kernel"},
                      "cpu_baseline": None if (world > 1 or args.no_cpu_baseline) else perft_cpu_baseline(cores)},
            "phases_ms_per_step": phases,
            "per_rank": per_rank_rec,
            "tree": {"bound": "latency (HBM pointer chasing, sequential within a game)", "bytes_per_sim": tree_bytes / d_sims,
                     "mean_path_nodes": d_path / d_sims, "mean_children_per_level": d_child / max(1, d_path),
                     "nodes_expanded_per_sim": d_nodes / d_sims, "achieved": (tree_bytes / (tree_ms * 1e-3) / 1e9) if tree_ms > 0 else None,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": (tree_bytes / (tree_ms * 1e-3) / 1e9 / pk["hbm"]) if tree_ms > 0 else None,
                     "kernels": "select_kernel + encode_kernel + expand_kernel, CUDA-event time on the engine stream (rank 0)"},
            "midgame": mid,
            "clocks": clocks,
        }
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            rate, t, kind = reference_search_rate(args.cpu_steps, 1, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "sims/s", "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_steps} searches x 800 sims of one game (T=2 x B=16, LibTorch CPU 128x10b), {t:.1f} s"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_config4(args, om, net, kw, rank, world, local, timed, allsum) -> dict:
    """BASELINE.json configs[3]: 128x10b self-play, 8 192 lockstep games partitioned over the N GPUs (no collective on
    the hot path).  Total work is fixed, so the N = 1 run of the same record is the strong-scaling denominator."""
    total = args.config4_games
    G = total // world
    steps, warmup = max(1, min(args.steps, args.config4_steps)), 2
    try:
        eng = om.SelfPlayEngine(G, seed=args.seed, first_game_id=rank * G, cuda_device=local, **kw)
        eng.prepare(net)
        step = lambda: eng.play_ply(want_counts=False)
        for _ in range(warmup):
            step()
        c0 = eng.counters()
        ms = timed(eng, step, steps)
        c1 = eng.counters()
        sims = allsum(c1["simulations"] - c0["simulations"])[0]
        del eng
        inflight = kw["num_threads"] * kw["batch_size"]
        return {"workload": f"BASELINE configs[3]: 128x10b self-play, {G * world} lockstep games partitioned by id over {world} GPU(s) "
                            f"({G} per GPU, leaf batch {G * inflight} per network call), {kw['num_simulations']} sims/move",
                "metric": "mcts_simulations_per_second", "value": sims / (ms * 1e-3), "unit": "sims/s", "n_gpus": world,
                "games_total": G * world, "games_per_gpu": G, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
                "scaling": "strong", "collectives_on_hot_path": 0}
    except Exception as exc:
        return {"error": repr(exc)}


def run_config3_leaf256(args, om, net, kw, rank, world, local, timed, allsum) -> dict:
    """BASELINE.json configs[2] read literally -- "leaf batch 256" per network call: 8 lockstep games per GPU x (T x B = 32)
    leaves.  The headline runs the same search with 1 000 games per GPU (32 000 boards per call) because a B200 needs about
    10^4 boards per launch to fill its SMs; this is the measured point at a 256-board call."""
    try:
        inflight = kw["num_threads"] * kw["batch_size"]
        G = max(1, 256 // inflight)
        eng = om.SelfPlayEngine(G, seed=args.seed, first_game_id=rank * G, cuda_device=local, **kw)
        eng.prepare(net)
        step = lambda: eng.play_ply(want_counts=False)
        steps, warmup = 10, 3
        for _ in range(warmup):
            step()
        c0 = eng.counters()
        ms = timed(eng, step, steps)
        c1 = eng.counters()
        sims, evals = allsum(c1["simulations"] - c0["simulations"], c1["evaluations"] - c0["evaluations"])
        del eng
        f_all, _ = flops_per_eval(c=args.channels, blocks=args.blocks, hv=args.hidden)
        tf = evals / world * f_all / (ms * 1e-3) / 1e12
        return {"workload": f"{args.channels}x{args.blocks + 1}b self-play, {G} lockstep games per GPU = leaf batch {G * inflight} per network call, "
                            f"{kw['num_simulations']} sims/move, {world} GPU(s)",
                "metric": "mcts_simulations_per_second", "value": sims / (ms * 1e-3), "unit": "sims/s", "n_gpus": world, "games_per_gpu": G,
                "leaf_batch_per_gpu": G * inflight, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "tflops_per_gpu": tf,
                "frac_of_sustained_peak": tf / peaks()["bf16_sustained"]}
    except Exception as exc:
        return {"error": repr(exc)}


def run_config5(args, om, rank, world, local, timed, allsum) -> dict:
    """BASELINE.json configs[4]: 256x20b (AlphaZeroNet(17, 64, 65, 256, 19, 256)), 1 600 sims/move, leaf batch per network
    call swept 64 ... 4096 on every one of the N GPUs (games per GPU = leaf batch / (T x B))."""
    try:
        net_kw = dict(NET, conv_channels=256, num_residual_blocks=19, value_head_hidden_channels=256)
        net = make_net(net_kw)
        kw = dict(MCTS_KW, num_simulations=1600)
        inflight = kw["num_threads"] * kw["batch_size"]
        f_all, _ = flops_per_eval(c=256, blocks=19, hv=256)
        pk = peaks()
        steps, warmup = max(1, args.config5_steps), 3
        points = []
        for leaf in args.config5_leaf:
            G = max(1, leaf // inflight)
            eng = om.SelfPlayEngine(G, seed=args.seed, first_game_id=rank * G, cuda_device=local, **kw)
            eng.prepare(net)
            step = lambda: eng.play_ply(want_counts=False)
            for _ in range(warmup):
                step()
            c0 = eng.counters()
            ms = timed(eng, step, steps)
            c1 = eng.counters()
            sims, evals = allsum(c1["simulations"] - c0["simulations"], c1["evaluations"] - c0["evaluations"])
            del eng
            rate = sims / (ms * 1e-3)
            tf = evals / world * f_all / (ms * 1e-3) / 1e12
            points.append({"leaf_batch_per_gpu": G * inflight, "games_per_gpu": G, "value": rate, "unit": "sims/s", "steps": steps,
                           "ms_per_step": ms / steps, "tflops_per_gpu": tf, "frac_of_sustained_peak": tf / pk["bf16_sustained"]})
        return {"workload": "BASELINE configs[4]: 256x20b random-init ResNet self-play, 1600 sims/move, leaf batch per network call and GPU "
                            f"swept over {list(args.config5_leaf)}, {world} GPU(s), games partitioned by id",
                "metric": "mcts_simulations_per_second", "n_gpus": world, "flop_per_eval": f_all, "warmup": warmup, "points": points}
    except Exception as exc:
        return {"error": repr(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=1000, help="lockstep games per GPU")
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--channels", type=int, default=128, help="conv_channels (BASELINE config 5: 256)")
    ap.add_argument("--blocks", type=int, default=9, help="residual blocks (config 5: 19)")
    ap.add_argument("--hidden", type=int, default=128, help="value head hidden units (config 5: 256)")
    ap.add_argument("--no-profiling", action="store_true", help="no per-phase CUDA events inside the searches (roofline fields stay empty)")
    ap.add_argument("--pipeline", type=int, default=0, help="game groups of the pipelined search (branches of a graph-replayed search): 0 auto, 1 off, 2..8")
    ap.add_argument("--pipeline-pdl", type=int, default=1, help="keep programmatic dependent launch while pipelined")
    ap.add_argument("--cpu-steps", type=int, default=36, help="searches of the CPU baseline sample (about 10-15 s on 16 host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--midgame-ply", type=int, default=24, help="also time 2 plies after fast-forwarding every game this far (0 = skip)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config4 / config5 sub-records")
    ap.add_argument("--config4-games", type=int, default=8192, help="total lockstep games of the config4 sub-record")
    ap.add_argument("--config4-steps", type=int, default=4)
    ap.add_argument("--config5-steps", type=int, default=10)
    ap.add_argument("--config5-leaf", type=int, nargs="*", default=[64, 128, 256, 512, 1024, 2048, 4096])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
