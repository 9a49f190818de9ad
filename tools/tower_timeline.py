#!/usr/bin/env python
"""Per-layer clock stamps of tower_cluster_kernel (cluster rank 8).  OZ_TOWER_DBG must be set in the environment."""
import ctypes as C, sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))
import numpy as np, torch
from othello_mcts import _lib, weights
from othello_mcts.neural_net import AlphaZeroNet
lib = _lib.load()
net = AlphaZeroNet(17, conv_channels=128, num_residual_blocks=9, value_head_hidden_channels=128).eval()
shape, blob = weights.fold_alphazero_net(net)
h = C.c_void_p(); _lib.check(lib.oz_resnet_create(0, C.byref(h))); _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
n = 32
x = (torch.rand(n, 17, 8, 8) < 0.3).float().cuda(); pol = torch.zeros(n, 65, device="cuda"); val = torch.zeros(n, device="cuda")
args = (h, C.c_void_p(x.data_ptr()), n, C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(0))
for _ in range(5): _lib.check(lib.oz_resnet_forward_device(*args))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): lib.oz_resnet_forward_device(*args)
e1.record(); torch.cuda.synchronize()
print("flags", os.environ.get("OZ_TOWER_DBG"), "us per forward", 1e3 * e0.elapsed_time(e1) / 200)
buf = np.zeros((64, 8), np.int64)
mc = C.c_int()
_lib.check(lib.oz_debug_tower_timeline(buf.ctypes.data_as(C.c_void_p), C.byref(mc)))
print("clusters the device runs at once:", mc.value)
t0 = buf[0, 0]
print("layer: [first_mma, issue_done, tmem_full_seen, nbr_done_seen(after ch0 math), ch0_arrived, ch1_arrived] clocks since layer-0 start")
for l in range(19):
    r = buf[l] - t0
    nxt = (buf[l + 1, 0] - buf[l, 4]) if l < 18 else None
    print(l, r[:6].tolist(), "| issue", r[1] - r[0], "tmem_full-issue_done", r[2] - r[1], "ld+math+nbr", r[3] - r[2], "ch0 store+fence+arrive", r[4] - r[3],
          "ch1", r[5] - r[4], "| next: rows0 seen after ch0 arrive", (buf[l + 1, 6] - buf[l, 4]) if l < 18 else None, "halo0 after rows0", (buf[l + 1, 7] - buf[l + 1, 6]) if l < 18 else None,
          "peer0+w after halo0", (buf[l + 1, 0] - buf[l + 1, 7]) if l < 18 else None, "| next layer first MMA after ch0 arrive", nxt, "| layer period", (buf[l + 1, 0] - buf[l, 0]) if l < 18 else None)
