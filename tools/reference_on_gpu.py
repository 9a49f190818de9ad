#!/usr/bin/env python
"""the UNMODIFIED reference on the same B200 -- its C++ MCTS (oracle/_ref/othello_mcts) with
`torch_device="cuda"`, i.e. CPU tree search + LibTorch CUDA evaluation of the reference's own AlphaZeroNet in batches of
16 (train.py:252-258; README.md:25 quotes 28 k sims/s for this arrangement on an RTX 4090) -- eager and, when inductor
works on the box, under torch.compile, beside this repo's single-game drop-in `MCTS.search(net)` on the same network
and search shape (128x10b, 800 sims/move, T=2 x B=16).  One JSON line.

  python tools/reference_on_gpu.py [--searches 20]
"""
import argparse
import importlib.util
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))
KW = dict(history_size=8, num_simulations=800, num_threads=2, batch_size=16, c_puct_base=20000.0, c_puct_init=2.5,
          dirichlet_epsilon=0.25, dirichlet_alpha=0.5)
NET = dict(in_channels=17, num_squares=64, num_actions=65, conv_channels=128, num_residual_blocks=9, value_head_hidden_channels=128)


def load_reference():
    ref_dir = ROOT / "oracle" / "_ref"
    spec = importlib.util.spec_from_file_location("ref_othello_mcts", ref_dir / "othello_mcts" / "__init__.py",
                                                  submodule_search_locations=[str(ref_dir / "othello_mcts")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_othello_mcts"] = mod
    spec.loader.exec_module(mod)
    sys.path.insert(0, str(ref_dir / "python"))
    from othello_alphazero.neural_net import AlphaZeroNet
    return mod, AlphaZeroNet


def time_searches(mcts, net, n, warmup=3):
    def advance():
        vc = np.array(mcts.visit_counts(), dtype=np.float64)
        acts = mcts.position().legal_actions()
        mcts.apply_action(int(np.random.choice(acts, p=vc / vc.sum())))
        if mcts.position().is_terminal():
            mcts.reset_position()
    np.random.seed(0)
    for _ in range(warmup):
        mcts.search(net); advance()
    torch.cuda.synchronize()
    t = 0.0
    for _ in range(n):
        t0 = time.perf_counter(); mcts.search(net); torch.cuda.synchronize(); t += time.perf_counter() - t0
        advance()
    return {"searches": n, "ms_per_search": 1e3 * t / n, "sims_per_s": 800 * n / t}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--searches", type=int, default=20)
    a = ap.parse_args()
    out = {"what": "single game, 128x10b random init, 800 sims/move, T=2 x B=16, one B200", "gpu": torch.cuda.get_device_name(0)}
    ref, RefNet = load_reference()
    torch.manual_seed(0)
    net = RefNet(**NET).to("cuda").eval().requires_grad_(False)
    with torch.no_grad():
        m = ref.MCTS(torch_device="cuda", torch_pin_memory=True, **KW)
        out["reference_cpp_mcts_libtorch_cuda_eager"] = time_searches(m, net, a.searches)
        try:
            cnet = torch.compile(net, fullgraph=True, dynamic=False)
            cnet(torch.zeros((16, 17, 8, 8), device="cuda"))
            m2 = ref.MCTS(torch_device="cuda", torch_pin_memory=True, **KW)
            out["reference_cpp_mcts_libtorch_cuda_torch_compile"] = time_searches(m2, cnet, a.searches)
        except Exception as exc:
            out["reference_cpp_mcts_libtorch_cuda_torch_compile"] = {"unavailable": repr(exc)[:300]}
        import othello_mcts as om
        d = om.MCTS(torch_device="cuda", **KW)
        out["dropin_single_game_native"] = time_searches(d, net, a.searches)
        d2 = om.MCTS(torch_device="cuda", **KW)
        out["dropin_single_game_callable_eager_net"] = time_searches(d2, lambda x: net(x), a.searches)
    out["dropin_vs_reference_eager"] = out["dropin_single_game_native"]["sims_per_s"] / out["reference_cpp_mcts_libtorch_cuda_eager"]["sims_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
