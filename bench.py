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
def make_net(net_kw=None):
    import torch
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
    net = make_net()
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
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    rate, t, kind = reference_search_rate(steps, warmup, cores)
    line = {"metric": "mcts_simulations_per_second", "value": rate, "unit": "sims/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": dict(workload_config(args, max(1, args.gpus)),
                           reference_sample="the reference plays the games one at a time (train.py:394-398): each step = one 800-simulation "
                                            "search of one game, C++ MCTS (T=2 x B=16) + LibTorch CPU on all host cores"),
            "cpu_baseline": {"value": rate, "unit": "sims/s", "cores": cores, "kind": kind,
                             "sample": f"{steps} searches x 800 sims, one game, after {warmup} warm-up searches"},
            "e2e": {"value": rate, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


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
    net_name = f"{args.channels}x{args.blocks + 1}b"
    net = make_net(net_kw)
    # games are partitioned contiguously by id: rank r owns [r*G, (r+1)*G); no collective on the hot path
    eng = om.SelfPlayEngine(G, seed=args.seed, first_game_id=rank * G, cuda_device=local, **kw)
    eng.prepare(net)
    eng.set_profiling(not args.no_profiling)
    eng.set_pipeline(args.pipeline, bool(args.pipeline_pdl))
    lib = _lib.load()
    stream = torch.cuda.ExternalStream(lib.oz_engine_stream(eng._h), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

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
        rank = np.cumsum(bits, axis=1) * bits                     # 1-based rank of every legal action
        acts = np.argmax(rank == (j + 1)[:, None], axis=1).astype(np.int32)
        acts[pos["legal"] == 0] = 64                              # pass
        acts[(n == 0) | (pos["player"] == 0)] = -1                # finished games
        eng.apply_actions(acts)                                   # H2D: actions (+ D2H of positions for validation)
        h2d = G * 4
        d2h = G * (4 + 36 * 4 + 36 * 4 + 4) + 2 * G * 40

    for _ in range(max(args.warmup, 3)):
        dev_step()
    c0 = eng.counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    c1 = eng.counters()
    sims = (c1["simulations"] - c0["simulations"])
    evals = c1["evaluations"] - c0["evaluations"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    eval_ms = c1["eval_ms"] - c0["eval_ms"]
    conv_ms = c1["conv_ms"] - c0["conv_ms"]
    conv_launches = c1["conv_launches"] - c0["conv_launches"]
    tot = torch.tensor([float(sims), float(evals), float(launches)], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot)                                      # the only collective: final counters
    value = float(tot[0].item()) / (ms * 1e-3)

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
        ms_mid = timed(dev_step, 2)
        m1 = eng.counters()
        mid_sims = torch.tensor([float(m1["simulations"] - m0["simulations"])], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(mid_sims)
        mid = {"ply": args.midgame_ply + 1, "steps": 2, "value": float(mid_sims.item()) / (ms_mid * 1e-3), "unit": "sims/s",
               "ms_per_step": ms_mid / 2, "select_ms_per_step": (m1["select_ms"] - m0["select_ms"]) / 2,
               "expand_ms_per_step": (m1["expand_ms"] - m0["expand_ms"]) / 2}
        eng.reset()
        eng.prepare(net)
        for _ in range(2):
            dev_step()

    ms_e2e = timed(e2e_step, max(1, min(args.steps, 3)))
    e2e_steps = max(1, min(args.steps, 3))
    active_sims = e2e_steps * G * sims_per_search * world
    e2e_value = active_sims / (ms_e2e * 1e-3)

    groups = max(1, args.pipeline)
    sel_ms, exp_ms = (c1["select_ms"] - c0["select_ms"]) / args.steps, (c1["expand_ms"] - c0["expand_ms"]) / args.steps
    if groups > 1:     # phases of different game groups overlap: network / conv = time at least one group spent there
        phases = {"pipeline_groups": groups, "network (union over groups)": eval_ms / args.steps, "of_which_conv": conv_ms / args.steps,
                  "select_encode (sum over groups, runs under the other group's network)": sel_ms,
                  "expand_backup (sum over groups, runs under the other group's network)": exp_ms,
                  "step time with no network running": (ms - eval_ms) / args.steps}
    else:
        phases = {"pipeline_groups": 1, "select_encode": sel_ms, "network": eval_ms / args.steps, "of_which_conv": conv_ms / args.steps,
                  "expand_backup": exp_ms,
                  "rest_of_ply (choose, record, re-root, syncs)": (ms - eval_ms) / args.steps - sel_ms - exp_ms}
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
    if rank == 0:
        f_all, f_conv = flops_per_eval(c=args.channels, blocks=args.blocks, hv=args.hidden)
        pk = peaks()
        nn_share = eval_ms / ms if ms > 0 else None
        # dominant kernel: the 3x3 convolution.  Algorithmic FLOPs of this rank's conv launches in the timed region
        # (dense 2*MAC, zero-padding taps counted: SURVEY.md 8d) / their CUDA-event time on the engine stream.
        achieved = (evals * f_conv) / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
        # best of 3: the event window of oz_perft spans the host round trips between its breadth-first levels
        perft_nodes, perft_ms = min((om.perft(11, return_ms=True) for _ in range(3)), key=lambda r: r[1])
        line = {
            "metric": "mcts_simulations_per_second", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "sims/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "path": "SelfPlayEngine.search(net) -> root_stats (host) -> numpy action choice -> apply_actions (host buffers)"},
            "gpu_launches": int(tot[2].item()),
            "nn_evals_per_second": float(tot[1].item()) / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / pk["bf16_sustained"]) if achieved else None, "traffic": NCU_TRAFFIC_BYTES_PER_BOARD * G * inflight if args.channels == 128 else None,
                         "kernel": "conv3x3_pair_kernel (tcgen05.mma cta_group::2 implicit GEMM)",
                         "flop_per_launch": (evals * f_conv / conv_launches) if conv_launches else None,
                         "us_per_launch": (1e3 * conv_ms / conv_launches) if conv_launches else None,
                         "launches_timed": int(conv_launches), "conv_share_of_step": conv_ms / ms if ms > 0 else None,
                         "nn_share_of_step": nn_share, "frac_of_burst_peak": (achieved / pk["bf16_burst"]) if achieved else None,
                         "traffic_note": "dram read+write bytes per launch (mean of a plain and a residual 128->128 layer, ncu --set full capture in "
                                         "profiles/ scaled to this leaf batch); algorithmic: 32 KB / 48 KB per board",
                         "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)"},
            "perft": {"depth": 11, "nodes": int(perft_nodes), "nodes_per_s": perft_nodes / (perft_ms * 1e-3) if perft_ms > 0 else None,
                      "device_ms": perft_ms},
            "phases_ms_per_step": phases,
            "tree": {"bound": "latency (HBM pointer chasing, sequential within a game)", "bytes_per_sim": tree_bytes / d_sims,
                     "mean_path_nodes": d_path / d_sims, "mean_children_per_level": d_child / max(1, d_path),
                     "nodes_expanded_per_sim": d_nodes / d_sims, "achieved": (tree_bytes / (tree_ms * 1e-3) / 1e9) if tree_ms > 0 else None,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": (tree_bytes / (tree_ms * 1e-3) / 1e9 / pk["hbm"]) if tree_ms > 0 else None,
                     "kernels": "select_kernel + encode_kernel + expand_kernel, CUDA-event time on the engine stream (rank 0)"},
            "midgame": mid,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, t, kind = reference_search_rate(args.cpu_steps, 1, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "sims/s", "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_steps} searches x 800 sims of one game (T=2 x B=16, LibTorch CPU 128x10b), {t:.1f} s"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--pipeline", type=int, default=0, help="game groups of the pipelined search: 0 auto, 1 off, 2..4")
    ap.add_argument("--pipeline-pdl", type=int, default=1, help="keep programmatic dependent launch while pipelined")
    ap.add_argument("--cpu-steps", type=int, default=36, help="searches of the CPU baseline sample (about 10-15 s on 16 host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--midgame-ply", type=int, default=24, help="also time 2 plies after fast-forwarding every game this far (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
