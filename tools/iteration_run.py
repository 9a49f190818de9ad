#!/usr/bin/env python
"""One or more whole AlphaZero iterations on one GPU, the schedule of the reference's `_run_iteration` (train.py:386-401):
self-play of `--games` games with the current network (lockstep, `SelfPlayEngine`), the 8-fold augmented training tensors
(`training_samples`, train.py:432-450), one epoch of SGD over them in shuffled batches (`NativeTrainer`, train.py:455-521),
weights back into the module.  Prints one JSON line per iteration with the time of every phase.

  python tools/iteration_run.py [--iterations 2] [--games 1000] [--sims 800] [--batch 256] [--channels 128 --blocks 9]"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import othello_mcts as om  # noqa: E402
from othello_mcts.neural_net import AlphaZeroNet  # noqa: E402
from othello_mcts.training import NativeTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--games", type=int, default=1000)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--mcts-batch", type=int, default=16)
    ap.add_argument("--batch", type=int, default=256, help="training batch size (README.md:89)")
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--history", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    torch.manual_seed(a.seed)
    net = AlphaZeroNet(1 + 2 * a.history, conv_channels=a.channels, num_residual_blocks=a.blocks, value_head_hidden_channels=a.hidden).eval()
    eng = om.SelfPlayEngine(a.games, history_size=a.history, num_simulations=a.sims, num_threads=a.threads, batch_size=a.mcts_batch, seed=a.seed)
    trainer = NativeTrainer(net, lr=a.lr, momentum=0.9, l2_weight=1e-4, max_batch=a.batch)
    rng = np.random.default_rng(a.seed)
    for it in range(a.iterations):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.reset()
        rec = eng.run(net)                                   # the current weights are folded and uploaded when they changed
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        data = eng.training_samples()                        # {"features", "policies", "values"} for every ply x 8 symmetries
        t2 = time.perf_counter()
        n = int(data["values"].shape[0])
        f, p, v = (data[k].cuda() for k in ("features", "policies", "values"))
        perm = torch.from_numpy(rng.permutation(n)).cuda()
        steps = n // a.batch                                 # drop_last=True (train.py:479)
        sums = {"total_loss": 0.0, "policy_loss": 0.0, "value_loss": 0.0, "l2_loss": 0.0}
        first = last = None
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        for s in range(steps):
            idx = perm[s * a.batch:(s + 1) * a.batch]
            want = s % 100 == 0 or s == steps - 1           # reading the losses back synchronises: only every 100th step
            if want:
                losses = trainer.step(f[idx], p[idx], v[idx])
                first = first or losses
                last = losses
            else:
                trainer.step_async(f[idx], p[idx], v[idx])
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        trainer.store_to(net)
        net.eval()
        t5 = time.perf_counter()
        print(json.dumps({"iteration": it, "games": a.games, "sims_per_move": a.sims, "network": f"{a.channels}x{a.blocks + 1}b",
                          "self_play_s": t1 - t0, "mean_game_plies": float(np.mean(rec["plies"])), "samples": n, "samples_emit_s": t2 - t1,
                          "upload_shuffle_s": t3 - t2, "train_steps": steps, "train_batch": a.batch, "train_s": t4 - t3,
                          "weights_back_s": t5 - t4, "iteration_s": t5 - t0, "first_losses": first, "last_losses": last}), flush=True)


if __name__ == "__main__":
    main()
