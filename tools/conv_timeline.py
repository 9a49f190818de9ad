#!/usr/bin/env python
"""%globaltimer stamps of CTA 0 of every per-layer convolution launch of one forward (measurement build only:
OZ_NVCC_FLAGS=-DOZ_CONV_TIMELINE python othello-alphazero_b200/build_ext.py --force).

  python tools/conv_timeline.py --channels 256 --blocks 19 --hidden 256 --boards 64
"""
import argparse, ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))
import numpy as np, torch
from othello_mcts import _lib, weights
from othello_mcts.neural_net import AlphaZeroNet

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=128); ap.add_argument("--blocks", type=int, default=9)
ap.add_argument("--hidden", type=int, default=128); ap.add_argument("--boards", type=int, default=32)
a = ap.parse_args()
lib = _lib.load()
raw = C.CDLL(str(ROOT / "othello-alphazero_b200/othello_mcts/libothello_b200.so"))
_lib.check(lib.oz_debug_set_tower(0))
net = AlphaZeroNet(17, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval()
shape, blob = weights.fold_alphazero_net(net)
h = C.c_void_p(); _lib.check(lib.oz_resnet_create(0, C.byref(h))); _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
n = a.boards
x = (torch.rand(n, 17, 8, 8) < 0.3).float().cuda(); pol = torch.zeros(n, 65, device="cuda"); val = torch.zeros(n, device="cuda")
args = (h, C.c_void_p(x.data_ptr()), n, C.c_void_p(pol.data_ptr()), C.c_void_p(val.data_ptr()), C.c_void_p(0))
for _ in range(10): _lib.check(lib.oz_resnet_forward_device(*args))
torch.cuda.synchronize()
buf = np.zeros((128, 8), np.uint64)
assert raw.oz_debug_conv_timeline(buf.ctypes.data_as(C.c_void_p)) == 0
L = 1 + 2 * a.blocks
b = buf[:L].astype(np.int64)
print(f"{a.channels}x{a.blocks + 1}b, {n} boards; ns relative to the launch's own start; columns: setup_done griddep_wait_returned first_act_full last_mma_issued epilogue_sees_acc epilogue_stored kernel_end | start - previous kernel_end | period")
for l in range(L):
    r = b[l] - b[l, 0]
    gap = (b[l, 0] - b[l - 1, 7]) if l else 0
    per = (b[l, 7] - b[l - 1, 7]) if l else 0
    print(l, r[1:8].tolist(), "| start-prev_end", int(gap), "| wait_returned-prev_end", int(b[l, 2] - b[l - 1, 7]) if l else 0, "| period", int(per))
