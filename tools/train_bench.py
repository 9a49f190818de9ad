#!/usr/bin/env python
"""Training-step throughput: the native step (`NativeTrainer`, oz_trainer_step_device) beside the reference's own loop body
(PyTorch, train.py:486-502) on the same GPU, same network (128x10b), same batch, device-resident inputs.

  python tools/train_bench.py [--batch 256] [--steps 50] [--channels 128 --blocks 9]

Prints one JSON line: steps/s and samples/s of both, and what a 480 000-sample epoch (README.md:23) would take."""
import argparse
import copy
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import torch  # noqa: E402
from torch import nn  # noqa: E402

from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402
from othello_mcts.training import NativeTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--skip-pytorch", action="store_true", help="time only the native step (for ncu captures)")
    ap.add_argument("--graphs", type=int, default=1, help="CUDA-graph replay of the native step")
    ap.add_argument("--tf32", type=int, default=1, help="PyTorch arm: allow TF32 convolutions (the PyTorch default)")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = bool(a.tf32)
    torch.manual_seed(0)
    net = AlphaZeroNet(17, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).cuda()
    n = a.batch
    f = (torch.rand(n, 17, 8, 8, device="cuda") < 0.3).float()
    p = torch.softmax(torch.randn(n, 65, device="cuda"), dim=1)
    v = torch.randint(-1, 2, (n,), device="cuda").float()
    lr, mom, l2 = 0.02, 0.9, 1e-4

    ref = copy.deepcopy(net).train()
    opt = torch.optim.SGD(ref.parameters(), lr=lr, momentum=mom)

    def ref_step():
        opt.zero_grad()
        out = ref(f)
        loss = -(p * out["policy"].log()).sum(dim=1).mean() + nn.functional.mse_loss(out["value"], v) + l2 * sum(
            q.square().sum() for q in ref.parameters())
        loss.backward()
        opt.step()
        return loss

    tr = NativeTrainer(net, lr=lr, momentum=mom, l2_weight=l2, max_batch=n)

    def time_it(fn, steps):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    import ctypes as C
    from othello_mcts._lib import check
    lib, h = tr._lib, tr._h

    def native_step():
        check(lib.oz_trainer_step_device(h, C.c_void_p(f.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr()), n, C.byref(tr.cfg), 1, None))

    tr.set_graphs(bool(a.graphs))
    t_ref = float("nan") if a.skip_pytorch else time_it(lambda: ref_step(), a.steps)
    t_nat = time_it(native_step, a.steps)
    flop = 3 * n * (2 * 576 * a.channels * (17 + 2 * a.blocks * a.channels))       # forward + input gradients + weight gradients of the 3x3 convolutions
    print(json.dumps({"network": f"{a.channels}x{a.blocks + 1}b", "batch": n, "steps": a.steps, "graphs": bool(a.graphs),
                      "native_ms_per_step": 1e3 * t_nat, "native_samples_per_s": n / t_nat, "native_conv_tflops": flop / t_nat / 1e12,
                      "pytorch_ms_per_step": 1e3 * t_ref, "pytorch_samples_per_s": n / t_ref, "pytorch_tf32": bool(a.tf32),
                      "speedup": t_ref / t_nat, "epoch_480k_native_s": 480000 / n * t_nat, "epoch_480k_pytorch_s": 480000 / n * t_ref}))


if __name__ == "__main__":
    main()
