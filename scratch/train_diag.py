import sys, copy
sys.path.insert(0, "othello-alphazero_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
from othello_mcts.neural_net import AlphaZeroNet
from othello_mcts.training import NativeTrainer
from test_gpu_training import _reference_step, _batch, _param_slices
torch.manual_seed(3)
planes, lr, mom, l2, n = 9, 0.02, 0.9, 1e-4, 64
net = AlphaZeroNet(planes, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() == 1: p.add_(0.1 * torch.randn_like(p))
ref = copy.deepcopy(net)
opt = torch.optim.SGD(ref.parameters(), lr=lr, momentum=mom)
tr = NativeTrainer(net, lr=lr, momentum=mom, l2_weight=l2, max_batch=n)
for step in range(3):
    f, p, v = _batch(n, planes, 100 + step)
    want, wg, wo = _reference_step(ref, opt, f.cuda(), p.cuda(), v.cuda(), l2)
    got = tr.step(f, p, v)
    print(step, got, want)
    g = tr.gradients()
    for name, sl in _param_slices(ref):
        a, b = g[sl].astype(np.float64), wg[sl].astype(np.float64)
        cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        print(f"  {name:45s} n={b.size:7d} |ref|={np.linalg.norm(b):.3e} ratio={np.linalg.norm(a)/(np.linalg.norm(b)+1e-30):.3f} cos={cos:.4f}")
