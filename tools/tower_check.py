#!/usr/bin/env python
"""Cluster tower (tower_cluster_kernel) against the per-layer convolution launches: bit equality of policy / value on
the same folded network, error against the fp32 PyTorch network, and the time of one forward on each path.

  python tools/tower_check.py [--iters 300]
"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from othello_mcts import _lib, weights  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    a = ap.parse_args()
    lib = _lib.load()
    ok = True
    mc = C.c_int()
    _lib.check(lib.oz_debug_tower_timeline(None, C.byref(mc)))
    print(json.dumps({"tower_clusters_resident_at_once": mc.value}), flush=True)
    for blocks, hidden in ((1, 32), (9, 128), (19, 128)):
        torch.manual_seed(blocks)
        net = AlphaZeroNet(17, conv_channels=128, num_residual_blocks=blocks, value_head_hidden_channels=hidden).eval()
        randomize_batchnorm(net, blocks + 1)
        shape, blob = weights.fold_alphazero_net(net)
        h = C.c_void_p()
        _lib.check(lib.oz_resnet_create(0, C.byref(h)))
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        for n in (1, 16, 32, 33, 64, 100, 256, 512, 1024):
            g = torch.Generator().manual_seed(n)
            x = (torch.rand(n, 17, 8, 8, generator=g) < 0.3).to(torch.float32)
            with torch.no_grad():
                ref = net(x)
            xd = x.cuda()
            out = {}
            rec = {"blocks": blocks, "boards": n}
            for mode in (0, 1):
                _lib.check(lib.oz_debug_set_tower(mode))
                pol = torch.zeros((n, 65), device="cuda")
                val = torch.zeros((n,), device="cuda")
                stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                args = (h, C.c_void_p(xd.data_ptr()), n, C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), stream)
                _lib.check(lib.oz_resnet_forward_device(*args))
                torch.cuda.synchronize()
                out[mode] = (pol.cpu().numpy().copy(), val.cpu().numpy().copy())
                for _ in range(20):
                    _lib.check(lib.oz_resnet_forward_device(*args))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    _lib.check(lib.oz_resnet_forward_device(*args))
                e1.record()
                torch.cuda.synchronize()
                rec["us_tower" if mode else "us_layers"] = round(1e3 * e0.elapsed_time(e1) / a.iters, 2)
            _lib.check(lib.oz_debug_set_tower(-1))
            same = bool(np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]))
            rec["bit_identical"] = same
            rec["tower_dp"] = float(np.abs(out[1][0] - ref["policy"].numpy()).max())
            rec["tower_dv"] = float(np.abs(out[1][1] - ref["value"].numpy()).max())
            rec["layers_dp"] = float(np.abs(out[0][0] - ref["policy"].numpy()).max())
            rec["max_policy_diff_between_paths"] = float(np.abs(out[0][0] - out[1][0]).max())
            ok = ok and same
            print(json.dumps(rec), flush=True)
        lib.oz_resnet_destroy(h)
    print("ALL BIT-IDENTICAL" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
