"""GPU: the native ResNet evaluator (tcgen05 implicit-GEMM convolutions + fused heads) against a plain
PyTorch fp32 reference of the same network.  Floating point, so a stated tolerance instead of bit
equality:
  * one conv layer (bf16 operands, fp32 accumulate, bf16 output): |d| <= 2^-7 * |ref| + 2e-2
  * whole network vs fp32 torch: policy max |dp| <= 5e-3, value max |dv| <= 2e-2 (measured
    errors at full depth are in profiles/r02_resnet_error.json: 3.2e-5 / 3.8e-4 at 128x10b, 1.1e-4 / 2.2e-3 at 256x20b)
  * the search built on it: same argmax action as the fp32 network in the large majority of roots
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


def _from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("cin,cout,boards,skip,relu", [
    (32, 128, 16, False, True),      # first layer shape (planes padded to 32), one unit
    (128, 128, 16, True, True),      # residual layer, one unit
    (128, 128, 50, True, True),      # ragged board count (padded to 64 inside), several passes per CTA
    (128, 128, 16 * 160, False, False),   # more passes than SMs: persistent loop + barrier phase wrap
    (256, 256, 32, True, True),      # 256-channel tower: two cout tiles, 8 K blocks
    (128, 128, 16 * 18, True, True),      # 36 pair-passes at two rows per pass: the one-row-per-pass kernel, 72 passes on 74 CTA pairs
    (128, 128, 16 * 20, True, True),      # 40 pair-passes: just above the switch, the two-rows-per-pass kernel
    (256, 256, 16 * 8, True, True),       # 256 channels, 32 pair-passes at two rows per pass: one-row kernel with streamed weights
    (256, 256, 16 * 12, False, True),     # ... and 48: two-row kernel
    (128, 128, 16 * 320, True, True),     # 640 pair-passes on 74 pairs: the dynamic pass scheduler (>= 8 passes per pair), ring wraps
    (256, 256, 16 * 160, False, True),    # 640 pair-passes with two cout tiles and streamed weights under the dynamic scheduler
])
def test_conv3x3_vs_torch(cin, cout, boards, skip, relu):
    from othello_mcts import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(cin * 1000 + boards)
    x = torch.randn(boards, 8, 8, cin, generator=g).to(torch.bfloat16).to(torch.float32)       # NHWC
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)).to(torch.bfloat16).to(torch.float32)
    bias = torch.randn(cout, generator=g) * 0.1
    sk = torch.randn(boards, 8, 8, cout, generator=g).to(torch.bfloat16).to(torch.float32) if skip else None
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=1).permute(0, 2, 3, 1)
    if skip:
        ref = ref + sk
    if relu:
        ref = torch.relu(ref)
    w_taps = w.permute(2, 3, 0, 1).reshape(9, cout, cin)
    xin, win = _bf16_bits(x), _bf16_bits(w_taps)
    skb = _bf16_bits(sk) if skip else None
    out = np.zeros((boards, 8, 8, cout), np.uint16)
    b32 = bias.numpy().astype(np.float32)
    ms = C.c_double()
    _lib.check(lib.oz_debug_conv3x3(xin.ctypes.data_as(C.c_void_p), cin, win.ctypes.data_as(C.c_void_p),
                                    b32.ctypes.data_as(C.c_void_p), skb.ctypes.data_as(C.c_void_p) if skip else None,
                                    int(relu), out.ctypes.data_as(C.c_void_p), cout, boards, C.byref(ms)))
    got = _from_bits(out)
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 2e-2
    bad = int((err > tol).sum())
    assert bad == 0, f"{bad} of {err.numel()} outside tolerance, max err {err.max():.4f}, kernel {ms.value:.3f} ms"


def _net(blocks=9, channels=128, planes=17, hidden=128, seed=0):
    from othello_mcts.neural_net import AlphaZeroNet, randomize_batchnorm
    torch.manual_seed(seed)
    net = AlphaZeroNet(planes, conv_channels=channels, num_residual_blocks=blocks, value_head_hidden_channels=hidden).eval()
    randomize_batchnorm(net, seed + 1)
    return net


def _random_feature_planes(n, planes, seed):
    """0/1 planes shaped like real inputs (side-to-move plane + disc planes)."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, planes, 8, 8, generator=g) < 0.3).to(torch.float32)
    x[:, 0] = (torch.rand(n, 1, 1, generator=g) < 0.5).to(torch.float32)
    return x


@pytest.mark.parametrize("blocks,channels,planes,hidden,n", [
    (1, 128, 9, 32, 40),
    (9, 128, 17, 128, 300),        # 128x10b, the benchmark network
    (2, 256, 17, 256, 48),         # 256-channel variant
])
def test_resnet_forward_vs_torch_fp32(blocks, channels, planes, hidden, n):
    from othello_mcts import _lib, weights
    lib = _lib.load()
    net = _net(blocks, channels, planes, hidden)
    x = _random_feature_planes(n, planes, 5)
    with torch.no_grad():
        ref = net(x)
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    try:
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        pol = np.zeros((n, 65), np.float32)
        val = np.zeros(n, np.float32)
        xs = x.numpy()
        _lib.check(lib.oz_resnet_forward(h, xs.ctypes.data_as(C.c_void_p), n, pol.ctypes.data_as(C.c_void_p),
                                         val.ctypes.data_as(C.c_void_p)))
    finally:
        lib.oz_resnet_destroy(h)
    dp = np.abs(pol - ref["policy"].numpy()).max()
    dv = np.abs(val - ref["value"].numpy()).max()
    assert np.allclose(pol.sum(axis=1), 1.0, atol=1e-4)
    assert dp <= 5e-3 and dv <= 2e-2, f"policy max abs diff {dp:.4f}, value max abs diff {dv:.4f}"


@pytest.mark.parametrize("blocks,hidden", [(1, 32), (9, 128)])
def test_cluster_tower_equals_per_layer_launches(blocks, hidden):
    """Launch-bound batches of a 128-channel network run every convolution layer in ONE launch of 16-CTA clusters with the
    activations in distributed shared memory (tower_cluster_kernel, csrc/oz_resnet.cu).  It issues the same MMAs in the same
    order as the per-layer kernels, so policy and value must be BIT-identical -- for ragged board counts, for several
    clusters, and for more clusters than the device runs at once -- and both stay within the evaluator's tolerance of the
    fp32 PyTorch network (neural_net.py:138-172)."""
    from othello_mcts import _lib, weights
    lib = _lib.load()
    resident = C.c_int()
    _lib.check(lib.oz_debug_tower_timeline(None, C.byref(resident)))
    assert resident.value >= 1, "the cluster tower cannot be launched on this device"
    net = _net(blocks, 128, 17, hidden, seed=blocks)
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    try:
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        for n in (1, 16, 32, 33, 100, 32 * resident.value, 32 * resident.value + 5, 64 * resident.value):
            x = _random_feature_planes(n, 17, 100 + n)
            with torch.no_grad():
                ref = net(x)
            xs = x.numpy()
            out = {}
            for mode in (0, 1):
                _lib.check(lib.oz_debug_set_tower(mode))
                pol = np.zeros((n, 65), np.float32)
                val = np.zeros(n, np.float32)
                _lib.check(lib.oz_resnet_forward(h, xs.ctypes.data_as(C.c_void_p), n, pol.ctypes.data_as(C.c_void_p),
                                                 val.ctypes.data_as(C.c_void_p)))
                out[mode] = (pol, val)
            assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]), f"tower != per-layer path at {n} boards"
            dp = np.abs(out[1][0] - ref["policy"].numpy()).max()
            dv = np.abs(out[1][1] - ref["value"].numpy()).max()
            assert dp <= 5e-3 and dv <= 2e-2, f"{n} boards: policy max abs diff {dp:.4f}, value max abs diff {dv:.4f}"
    finally:
        _lib.check(lib.oz_debug_set_tower(-1))
        lib.oz_resnet_destroy(h)


@pytest.mark.parametrize("blocks,channels,hidden", [(2, 128, 128), (1, 256, 256), (1, 128, 32)])
def test_heads_one_board_per_warp_equals_four(blocks, channels, hidden):
    """heads_kernel (neural_net.py:68-128) runs one board per warp for launch-bound batches and four beyond; every output is
    the same sum in the same order, so policy and value must be BIT-identical whichever instantiation runs -- with the 1x1
    head convolutions fused into the last conv layer (128 channels) and computed by the kernel itself (256 channels)."""
    from othello_mcts import _lib, weights
    lib = _lib.load()
    net = _net(blocks, channels, 17, hidden, seed=7)
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    try:
        _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))
        for n in (1, 31, 32, 257, 2400):
            xs = _random_feature_planes(n, 17, 300 + n).numpy()
            out = {}
            for boards in (1, 4):
                _lib.check(lib.oz_debug_set_heads_boards(boards))
                pol = np.zeros((n, 65), np.float32)
                val = np.zeros(n, np.float32)
                _lib.check(lib.oz_resnet_forward(h, xs.ctypes.data_as(C.c_void_p), n, pol.ctypes.data_as(C.c_void_p),
                                                 val.ctypes.data_as(C.c_void_p)))
                out[boards] = (pol, val)
            assert np.array_equal(out[1][0], out[4][0]) and np.array_equal(out[1][1], out[4][1]), f"heads<1> != heads<4> at {n} boards"
            assert np.allclose(out[1][0].sum(axis=1), 1.0, atol=1e-4)
    finally:
        _lib.check(lib.oz_debug_set_heads_boards(0))
        lib.oz_resnet_destroy(h)


def test_single_game_search_same_tree_with_and_without_tower():
    """The reference-shaped single-game MCTS.search (32 leaves per network call, graph replay of the whole search with the
    kernels chained by programmatic dependent launch) must build the same tree on the cluster tower and on the per-layer
    launches, ply after ply."""
    from othello_mcts import MCTS, _lib
    lib = _lib.load()
    net = _net(2, 128, 9, 64).cuda()
    kw = dict(history_size=4, torch_device="cuda", num_simulations=160, num_threads=2, batch_size=16, dirichlet_epsilon=0.25, seed=11)
    try:
        _lib.check(lib.oz_debug_set_tower(1))
        a = MCTS(**kw)
        _lib.check(lib.oz_debug_set_tower(0))
        b = MCTS(**kw)
        for ply in range(12):
            _lib.check(lib.oz_debug_set_tower(1))
            a.search(net)
            _lib.check(lib.oz_debug_set_tower(0))
            b.search(net)
            assert a.visit_counts() == b.visit_counts() and a.mean_action_values() == b.mean_action_values(), ply
            act = a.position().legal_actions()[int(np.argmax(a.visit_counts()))]
            a.apply_action(act)
            b.apply_action(act)
    finally:
        _lib.check(lib.oz_debug_set_tower(-1))


def test_search_with_native_network():
    """MCTS.search(AlphaZeroNet) takes the fused native path (tree kernels + in-engine evaluator).  Driving
    the same search through the callable contract, with the callable running the SAME folded network via
    the stand-alone oz_resnet handle, must give identical trees: the two paths differ only in plumbing."""
    from othello_mcts import MCTS, _lib, weights
    lib = _lib.load()
    net = _net(2, 128, 9, 64).cuda()
    shape, blob = weights.fold_alphazero_net(net)
    h = C.c_void_p()
    _lib.check(lib.oz_resnet_create(0, C.byref(h)))
    _lib.check(lib.oz_resnet_load(h, C.byref(shape), blob, len(blob)))

    def via_handle(x):
        n = x.shape[0]
        pol = torch.empty((n, 65), device="cuda", dtype=torch.float32)
        val = torch.empty((n,), device="cuda", dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.oz_resnet_forward_device(h, C.c_void_p(x.data_ptr()), n, C.c_void_p(pol.data_ptr()),
                                                C.c_void_p(val.data_ptr()), stream))
        return {"policy": pol, "value": val}

    kw = dict(history_size=4, torch_device="cuda", num_simulations=128, num_threads=1, batch_size=16,
              dirichlet_epsilon=0.25, seed=3)
    native, callable_path = MCTS(**kw), MCTS(**kw)
    try:
        for ply in range(10):
            native.search(net)
            callable_path.search(via_handle)
            assert native.visit_counts() == callable_path.visit_counts(), ply
            assert native.mean_action_values() == callable_path.mean_action_values(), ply
            if ply == 0:
                assert sum(native.visit_counts()) == 128 - 16
            act = native.position().legal_actions()[int(np.argmax(native.visit_counts()))]
            native.apply_action(act)
            callable_path.apply_action(act)
    finally:
        lib.oz_resnet_destroy(h)
    assert native.counters()["evaluations"] > 0


def test_second_device_in_one_process():
    """cudaFuncSetAttribute is per device: an engine on cuda:1 after one on cuda:0 in the same process must get the
    convolution's dynamic shared-memory limit too (and identical results).  Needs two GPUs; skipped otherwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from othello_mcts import MCTS
    net = _net(2, 128, 9, 64)
    kw = dict(history_size=4, num_simulations=64, num_threads=1, batch_size=16, dirichlet_epsilon=0.25, seed=3)
    a, b = MCTS(torch_device="cuda:0", **kw), MCTS(torch_device="cuda:1", **kw)
    for _ in range(3):
        a.search(net)
        b.search(net)
        assert a.visit_counts() == b.visit_counts() and a.mean_action_values() == b.mean_action_values()
        act = a.position().legal_actions()[int(np.argmax(a.visit_counts()))]
        a.apply_action(act)
        b.apply_action(act)
