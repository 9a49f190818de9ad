set -x
O=gpurun_out/r01p; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/summary.txt
timeout 120 python tools/nn_forward.py --boards 32000 --iters 10 > $O/nn.log 2>&1
timeout 600 python bench.py --midgame-ply 0 > $O/bench.log 2> $O/bench.err; echo "bench rc=$?" >> $O/summary.txt
python - > $O/latency.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "othello-alphazero_b200")
import torch
from othello_mcts import MCTS
from othello_mcts.neural_net import AlphaZeroNet
torch.manual_seed(0)
net = AlphaZeroNet(17, conv_channels=128, num_residual_blocks=9, value_head_hidden_channels=128).eval().cuda()
m = MCTS(history_size=8, torch_device="cuda", num_simulations=800, num_threads=2, batch_size=16)
for _ in range(3):
    m.search(net); m.apply_action(m.position().legal_actions()[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    m.search(net); m.apply_action(m.position().legal_actions()[0])
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print({"single_game_search_ms": dt * 1e3, "sims_per_s": 800 / dt})
PY
