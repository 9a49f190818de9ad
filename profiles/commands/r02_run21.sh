O=gpurun_out/r02u; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_reference_callers.py -m gpu -x -q > $O/pytest_callers.log 2>&1; echo "pytest callers rc=$?" >> $O/summary.txt
python - > $O/cosines.txt 2>&1 <<'PY'
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'othello-alphazero_b200')
import numpy as np, torch, copy
import test_gpu_training as T
from othello_mcts.neural_net import AlphaZeroNet
from othello_mcts.training import NativeTrainer
for (channels,blocks,hidden,n) in [(128,2,32,64),(128,3,128,96),(256,1,64,32)]:
    for rep in range(4):
        torch.manual_seed(3)
        ref = AlphaZeroNet(9, conv_channels=channels, num_residual_blocks=blocks, value_head_hidden_channels=hidden).cuda()
        with torch.no_grad():
            for p in ref.parameters():
                if p.dim()==1: p.add_(0.1*torch.randn_like(p))
        opt = torch.optim.SGD(ref.parameters(), lr=0.02, momentum=0.9)
        tr = NativeTrainer(ref, lr=0.02, momentum=0.9, l2_weight=1e-4, max_batch=n)
        cs=[]
        for step in range(3):
            tr.load_from(ref, momentum_buffers=T._momentum_vector(opt, ref))
            f,p,v = T._batch(n, 9, 100+step)
            want, wg, wo = T._reference_step(ref, opt, f.cuda(), p.cuda(), v.cuda(), 1e-4)
            tr.step(f,p,v)
            g = tr.gradients().astype(np.float64); b = wg.astype(np.float64)
            cs.append(float(g@b/(np.linalg.norm(g)*np.linalg.norm(b))))
        print(channels,blocks,hidden,n,rep,[round(c,5) for c in cs])
PY
