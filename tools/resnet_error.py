#!/usr/bin/env python
"""Measured error of the native evaluator (bf16 tcgen05 convolutions, fp32 accumulation) against the fp32 network, at
the full depth of both BASELINE networks: 128x10b and 256x20b.  The fp32 side is othello_mcts.neural_net.AlphaZeroNet on
the CPU, which tests/test_net_golden.py pins bit for bit to the reference class (neural_net.py:138-172); the 8 golden
inputs are compared with the reference class's own recorded outputs as well.  Writes one JSON document to stdout.

  python tools/resnet_error.py [--boards 256] > profiles/r02_resnet_error.json
"""
import argparse
import ctypes as C
import json
import struct
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))
from othello_mcts import _lib, weights  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm  # noqa: E402


def native_forward(net, x):
    lib = _lib.load()
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    try:
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        n = x.shape[0]
        pol, val = np.zeros((n, 65), np.float32), np.zeros(n, np.float32)
        xs = np.ascontiguousarray(x.numpy())
        _lib.check(lib.oz_resnet_forward(h, xs.ctypes.data_as(C.c_void_p), n, pol.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p)))
    finally:
        lib.oz_resnet_destroy(h)
    return pol, val


def planes(n, p, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, p, 8, 8, generator=g) < 0.3).to(torch.float32)
    x[:, 0] = (torch.rand(n, 1, 1, generator=g) < 0.5).to(torch.float32)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--boards", type=int, default=256)
    args = ap.parse_args()
    gold = json.loads((ROOT / "tests" / "golden" / "golden_net.json").read_text())
    out = {"what": "native evaluator (bf16 operands / activations, fp32 accumulation) vs fp32 torch CPU of the same weights",
           "tolerances_in_tests": {"policy": 2e-3, "value_128x10b": 5e-3, "value_256x20b": 1e-2}, "nets": {}}
    for name, e in gold["nets"].items():
        torch.manual_seed(gold["seed"])
        net = AlphaZeroNet(**e["kwargs"]).eval()
        randomize_batchnorm(net, gold["seed"] + 1)
        rows = {}
        # (1) the 8 golden inputs against the reference class's recorded fp32 outputs
        x = planes(gold["n_inputs"], e["kwargs"]["in_channels"], gold["seed"] + 2)
        pol, val = native_forward(net, x)
        rp = np.array([struct.unpack("<f", struct.pack("<I", b))[0] for b in e["policy_bits"]], np.float32).reshape(-1, 65)
        rv = np.array([struct.unpack("<f", struct.pack("<I", b))[0] for b in e["value_bits"]], np.float32)
        rows["golden_inputs_vs_reference_class"] = {"boards": int(x.shape[0]), "policy_max_abs": float(np.abs(pol - rp).max()),
                                                     "value_max_abs": float(np.abs(val - rv).max())}
        # (2) a larger sample against the pinned fp32 mirror
        x = planes(args.boards, e["kwargs"]["in_channels"], 99)
        with torch.no_grad():
            ref = net(x)
        pol, val = native_forward(net, x)
        dp, dv = np.abs(pol - ref["policy"].numpy()), np.abs(val - ref["value"].numpy())
        agree = float((pol.argmax(1) == ref["policy"].numpy().argmax(1)).mean())
        rows["random_planes_vs_fp32"] = {"boards": args.boards, "policy_max_abs": float(dp.max()), "policy_mean_abs": float(dp.mean()),
                                         "value_max_abs": float(dv.max()), "value_mean_abs": float(dv.mean()),
                                         "policy_argmax_agreement": agree,
                                         "fp32_value_range": [float(ref["value"].min()), float(ref["value"].max())],
                                         "fp32_policy_max": float(ref["policy"].max())}
        out["nets"][name] = rows
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
