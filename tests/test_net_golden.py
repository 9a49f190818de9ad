"""The repo's fp32 network class (othello_mcts.neural_net.AlphaZeroNet: the oracle of the native evaluator on the GPU box,
where /root/reference does not exist) pinned to the UNMODIFIED reference class python/othello_alphazero/neural_net.py:138-172
through tests/golden/golden_net.json (generate_golden_net.py imports the reference class itself):

  * CPU: same state_dict keys, bit-identical parameters after the same seed (SHA-256), outputs for 8 fixed inputs within
    1e-6 of the reference's (bit-identical on the CPU the vectors were generated on; other CPUs may pick another oneDNN
    convolution kernel with a different summation order) -- at 128x10b AND at the full 256x20b depth;
  * CPU, when /root/reference is present: the live reference class against the mirror, bit for bit, incl. a state_dict
    round trip (reference checkpoints load into the mirror and vice versa);
  * GPU: the native bf16 evaluator against the reference's fp32 outputs at full depth, stated tolerance.
"""
import ctypes as C
import hashlib
import json
import struct
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = json.loads((ROOT / "tests" / "golden" / "golden_net.json").read_text())

# Tolerances of the native evaluator (bf16 operands, fp32 accumulation, bf16 activations between layers) against the
# reference's fp32 outputs; the measured errors are recorded in profiles/r02_resnet_error.json (tools/resnet_error.py).
POLICY_TOL = 2e-3
VALUE_TOL = {"128x10b": 5e-3, "256x20b": 1e-2}


def _from_bits(bits):
    return np.array([struct.unpack("<f", struct.pack("<I", b))[0] for b in bits], np.float32)


def _mirror(name):
    from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm
    e = GOLD["nets"][name]
    torch.manual_seed(GOLD["seed"])
    net = AlphaZeroNet(**e["kwargs"]).eval()
    randomize_batchnorm(net, GOLD["seed"] + 1)
    return net, e


def _inputs(planes):
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    g = torch.Generator().manual_seed(GOLD["seed"] + 2)
    x = (torch.rand(GOLD["n_inputs"], planes, 8, 8, generator=g) < 0.3).to(torch.float32)
    x[:, 0] = (torch.rand(GOLD["n_inputs"], 1, 1, generator=g) < 0.5).to(torch.float32)
    return x


def _sha(net):
    h = hashlib.sha256()
    for k, v in net.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", ["128x10b", "256x20b"])
def test_mirror_class_reproduces_reference_vectors(name):
    net, e = _mirror(name)
    assert list(net.state_dict().keys()) == e["state_keys"]
    assert sum(p.numel() for p in net.parameters()) == e["num_parameters"]
    assert _sha(net) == e["state_sha256"], "parameter initialisation differs from the reference class"
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        with torch.no_grad():
            y = net(_inputs(e["kwargs"]["in_channels"]))
    finally:
        torch.set_num_threads(threads)
    pol, val = _from_bits(e["policy_bits"]).reshape(-1, 65), _from_bits(e["value_bits"])
    assert np.abs(y["policy"].numpy() - pol).max() <= 1e-6
    assert np.abs(y["value"].numpy() - val).max() <= 1e-6


@pytest.mark.parametrize("name", ["128x10b", "256x20b"])
def test_live_reference_class_equals_mirror(name):
    ref_py = Path("/root/reference/python")
    if not (ref_py / "othello_alphazero" / "neural_net.py").exists():
        ref_py = ROOT / "oracle" / "_ref" / "python"
    if not (ref_py / "othello_alphazero" / "neural_net.py").exists():
        pytest.skip("reference python package not available")
    sys.path.insert(0, str(ref_py))
    from othello_alphazero.neural_net import AlphaZeroNet as RefNet
    from othello_mcts.neural_net import randomize_batchnorm
    mirror, e = _mirror(name)
    torch.manual_seed(GOLD["seed"])
    ref = RefNet(**e["kwargs"]).eval()
    randomize_batchnorm(ref, GOLD["seed"] + 1)
    assert _sha(ref) == _sha(mirror) == e["state_sha256"]
    x = _inputs(e["kwargs"]["in_channels"])
    with torch.no_grad():
        a, b = ref(x), mirror(x)
        assert torch.equal(a["policy"], b["policy"]) and torch.equal(a["value"], b["value"])
        # checkpoints travel both ways
        ref2 = RefNet(**e["kwargs"]).eval()
        ref2.load_state_dict(mirror.state_dict())
        c = ref2(x)
    assert torch.equal(a["policy"], c["policy"]) and torch.equal(a["value"], c["value"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["128x10b", "256x20b"])
def test_native_evaluator_vs_reference_fp32_full_depth(name):
    """oz_resnet_forward (tcgen05 convolutions + fused heads) at the FULL depth of both BASELINE networks against the
    reference class's own fp32 outputs (golden).  Tolerance: policy max |dp| <= 2e-3; value max |dv| <= 5e-3 (10b) /
    1e-2 (20b: 39 bf16 activation roundings in sequence).  Measured on B200 (profiles/r02_resnet_error.json): policy
    2.5e-5 / 7.8e-5, value 1.9e-4 / 1.6e-3 on these inputs; 3.2e-5 / 1.1e-4 and 3.8e-4 / 2.2e-3 over 256 boards."""
    from othello_mcts import _lib, weights
    lib = _lib.load()
    net, e = _mirror(name)
    x = _inputs(e["kwargs"]["in_channels"])
    n = x.shape[0]
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    try:
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        pol, val = np.zeros((n, 65), np.float32), np.zeros(n, np.float32)
        xs = x.numpy()
        _lib.check(lib.oz_resnet_forward(h, xs.ctypes.data_as(C.c_void_p), n, pol.ctypes.data_as(C.c_void_p),
                                         val.ctypes.data_as(C.c_void_p)))
    finally:
        lib.oz_resnet_destroy(h)
    rp, rv = _from_bits(e["policy_bits"]).reshape(-1, 65), _from_bits(e["value_bits"])
    dp, dv = np.abs(pol - rp).max(), np.abs(val - rv).max()
    assert np.allclose(pol.sum(axis=1), 1.0, atol=1e-4)
    assert dp <= POLICY_TOL and dv <= VALUE_TOL[name], f"{name}: policy max |dp| {dp:.5f}, value max |dv| {dv:.5f}"
