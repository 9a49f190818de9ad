"""Generate tests/golden/golden_net.json from the UNMODIFIED reference network class
(/root/reference/python/othello_alphazero/neural_net.py:138-172; needs /root/reference).

The repo's `othello_mcts.neural_net.AlphaZeroNet` is a mirror of that class used as the fp32 oracle of the native
evaluator on the GPU box (where /root/reference does not exist).  These vectors pin the mirror to the reference:
for 128x10b `(17, 64, 65, 128, 9, 128)` and 256x20b `(17, 64, 65, 256, 19, 256)` at a fixed seed
  * the state_dict key list and a SHA-256 over every parameter / buffer (default init after torch.manual_seed, then
    non-trivial BatchNorm statistics from a seeded generator) -- exact;
  * fp32 policy / value of the reference class (eval mode, CPU) for 8 fixed inputs, as float32 bit patterns.
Run in the build container:  python tests/golden/generate_golden_net.py
"""
import hashlib
import json
import struct
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, "/root/reference/python")
from othello_alphazero.neural_net import AlphaZeroNet as RefNet  # noqa: E402  (the reference class itself)

SHAPES = {
    "128x10b": dict(in_channels=17, num_squares=64, num_actions=65, conv_channels=128, num_residual_blocks=9,
                    value_head_hidden_channels=128),
    "256x20b": dict(in_channels=17, num_squares=64, num_actions=65, conv_channels=256, num_residual_blocks=19,
                    value_head_hidden_channels=256),
}
SEED = 20261018
N_INPUTS = 8


def randomize_batchnorm(net, seed):
    """Non-trivial eval-mode statistics (same recipe as othello_mcts.neural_net.randomize_batchnorm; restated here so
    the generator depends on nothing but torch and the reference class)."""
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def fixed_inputs(n, planes, seed):
    """0/1 planes shaped like real inputs: a constant side-to-move plane and sparse disc planes."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, planes, 8, 8, generator=g) < 0.3).to(torch.float32)
    x[:, 0] = (torch.rand(n, 1, 1, generator=g) < 0.5).to(torch.float32)
    return x


def state_sha256(net):
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def bits(t):
    return [struct.unpack("<I", struct.pack("<f", float(v)))[0] for v in t.flatten().tolist()]


def main():
    torch.set_num_threads(1)                       # one thread: the summation order inside a convolution is fixed
    out = {"source": "unmodified reference class python/othello_alphazero/neural_net.py (AlphaZeroNet), eval mode, CPU fp32, "
                     "torch " + torch.__version__, "seed": SEED, "n_inputs": N_INPUTS, "nets": {}}
    for name, kw in SHAPES.items():
        torch.manual_seed(SEED)
        net = RefNet(**kw).eval()
        randomize_batchnorm(net, SEED + 1)
        x = fixed_inputs(N_INPUTS, kw["in_channels"], SEED + 2)
        with torch.no_grad():
            y = net(x)
        out["nets"][name] = {"kwargs": kw, "state_keys": list(net.state_dict().keys()), "state_sha256": state_sha256(net),
                             "num_parameters": sum(p.numel() for p in net.parameters()),
                             "policy_bits": bits(y["policy"]), "value_bits": bits(y["value"])}
        print(name, out["nets"][name]["num_parameters"], "parameters", out["nets"][name]["state_sha256"][:16])
    path = Path(__file__).with_name("golden_net.json")
    path.write_text(json.dumps(out, separators=(",", ":")))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
