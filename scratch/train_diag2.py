import sys, copy
sys.path.insert(0, "othello-alphazero_b200"); sys.path.insert(0, "tests")
import numpy as np, torch
from torch import nn
from othello_mcts.neural_net import AlphaZeroNet
from othello_mcts.training import NativeTrainer
from test_gpu_training import _batch, _param_slices
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
planes, lr, mom, l2, n = 9, 0.02, 0.9, 1e-4, 64
net = AlphaZeroNet(planes, conv_channels=128, num_residual_blocks=2, value_head_hidden_channels=32).cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() == 1: p.add_(0.1 * torch.randn_like(p))
def grads_of(model, f, p, v, autocast):
    model.train(); model.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = model(f)
    pol, val = out["policy"].float(), out["value"].float()
    loss = -(p * pol.log()).sum(dim=1).mean() + nn.functional.mse_loss(val, v) + l2 * sum(q.square().sum() for q in model.parameters())
    loss.backward()
    return torch.nn.utils.parameters_to_vector([q.grad for q in model.parameters()]).detach().cpu().numpy().astype(np.float64)
f, p, v = _batch(n, planes, 100)
g32 = grads_of(copy.deepcopy(net), f.cuda(), p.cuda(), v.cuda(), False)
g16 = grads_of(copy.deepcopy(net), f.cuda(), p.cuda(), v.cuda(), True)
tr = NativeTrainer(net, lr=lr, momentum=mom, l2_weight=l2, max_batch=n)
tr.step(f, p, v, apply_update=False)
go = tr.gradients().astype(np.float64)
cos = lambda a, b: float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
print(f"{'tensor':45s} ours-vs-fp32  autocast-vs-fp32")
for name, sl in _param_slices(net):
    print(f"{name:45s} {cos(go[sl], g32[sl]):.4f}        {cos(g16[sl], g32[sl]):.4f}")
print("all", cos(go, g32), cos(g16, g32))
