#!/usr/bin/env python
"""Time the native ResNet evaluator alone (device-resident inputs, CUDA events on the launch stream).

  python tools/nn_forward.py [--boards 32768] [--iters 10] [--channels 128] [--blocks 9]

Prints evals/s and the achieved dense TFLOP/s (SURVEY.md section 8(d) FLOP count).  Also the command
profiled with ncu for profiles/ ("Evidence")."""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import torch  # noqa: E402

from othello_mcts import _lib, weights  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--boards", type=int, default=32768)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--planes", type=int, default=17)
    a = ap.parse_args()
    lib = _lib.load()
    torch.manual_seed(0)
    net = AlphaZeroNet(a.planes, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval()
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
    x = (torch.rand(a.boards, a.planes, 8, 8, device="cuda") < 0.3).float()
    pol = torch.empty(a.boards, 65, device="cuda")
    val = torch.empty(a.boards, device="cuda")
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)

    def fwd():
        _lib.check(lib.oz_resnet_forward_device(h, C.c_void_p(x.data_ptr()), a.boards, C.c_void_p(pol.data_ptr()),
                                                C.c_void_p(val.data_ptr()), sp))
    for _ in range(a.warmup):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters):
        fwd()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    c, cin, r, hv = a.channels, a.planes, a.blocks, a.hidden
    flop = 2 * 576 * c * (cin + 2 * r * c) + 2 * (c * 2 * 64 + 128 * 65) + 2 * (c * 64 + 64 * hv + hv)
    print(json.dumps({"boards": a.boards, "ms_per_forward": ms, "evals_per_s": a.boards / (ms * 1e-3),
                      "tflops_dense": a.boards * flop / (ms * 1e-3) / 1e12, "flop_per_eval": flop,
                      "policy_sum": float(pol.sum(dim=1).mean())}))
    lib.oz_resnet_destroy(h)


if __name__ == "__main__":
    main()
