"""Fold an AlphaZeroNet (reference: python/othello_alphazero/neural_net.py:138-172) into the
weight blob the native evaluator reads (layout: csrc/oz_resnet.cuh).

Eval-mode BatchNorm is an affine map, so it folds into the preceding convolution:
``W' = W * s``, ``b' = (b - mean) * s + beta`` with ``s = gamma / sqrt(var + eps)``.  The
3x3 kernels are stored tap-major ``[ky*3+kx][cout][cin]`` in bf16 (first layer: planes
zero-padded to 32); biases and the small head tensors stay fp32.  The network parameters
remain owned by the Python module; this only reads them.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import OzResnetShape, check


def unwrap(module):
    """torch.compile wraps the network in an OptimizedModule (train.py:307-308)."""
    return getattr(module, "_orig_mod", module)


def is_alphazero_net(obj) -> bool:
    m = unwrap(obj)
    return all(hasattr(m, a) for a in ("conv_block", "residual_blocks", "policy_head", "value_head")) and isinstance(
        m, torch.nn.Module)


def shape_of(net) -> OzResnetShape:
    net = unwrap(net)
    conv = net.conv_block.conv
    return OzResnetShape(in_planes=conv.in_channels, channels=conv.out_channels, blocks=len(net.residual_blocks),
                         value_hidden=net.value_head.linear1.out_features)


def natively_supported(net) -> bool:
    """Shapes the sm_100a evaluator covers; anything else goes through the callable path."""
    net = unwrap(net)
    try:
        s = shape_of(net)
        ph, vh = net.policy_head, net.value_head
        return (s.channels in (128, 256, 384) and 1 <= s.in_planes <= 32 and 1 <= s.value_hidden <= 512
                and ph.conv.out_channels == 2 and ph.linear.in_features == 128 and ph.linear.out_features == 65
                and vh.conv.out_channels == 1 and vh.linear1.in_features == 64 and not net.training)
    except AttributeError:
        return False


def _fold(conv, bn):
    w = conv.weight.detach().to("cpu", torch.float64)
    b = conv.bias.detach().to("cpu", torch.float64) if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    s = bn.weight.detach().to("cpu", torch.float64) / torch.sqrt(bn.running_var.detach().to("cpu", torch.float64) + bn.eps)
    w = w * s.view(-1, 1, 1, 1)
    b = (b - bn.running_mean.detach().to("cpu", torch.float64)) * s + bn.bias.detach().to("cpu", torch.float64)
    return w.to(torch.float32), b.to(torch.float32)


def _conv3x3_bytes(w: torch.Tensor, cin_pad: int) -> bytes:
    cout, cin = w.shape[0], w.shape[1]
    t = torch.zeros(9, cout, cin_pad, dtype=torch.float32)
    t[:, :, :cin] = w.permute(2, 3, 0, 1).reshape(9, cout, cin)      # [ky][kx][cout][cin] -> tap-major
    return t.to(torch.bfloat16).view(torch.int16).numpy().tobytes()


def _f32(t: torch.Tensor) -> bytes:
    return t.detach().to("cpu", torch.float32).contiguous().numpy().tobytes()


def fold_alphazero_net(net) -> tuple[OzResnetShape, bytes]:
    """Returns (shape, blob) for oz_engine_load_resnet / oz_resnet_load."""
    net = unwrap(net)
    lib = _lib.load()
    shape = shape_of(net)
    total = lib.oz_resnet_blob_bytes(C.byref(shape))
    offs = (C.c_int64 * 256)()
    n = C.c_int()
    check(lib.oz_resnet_blob_offsets(C.byref(shape), offs, 256, C.byref(n)))
    sections: list[bytes] = []
    w, b = _fold(net.conv_block.conv, net.conv_block.norm)
    sections += [_conv3x3_bytes(w, 32), _f32(b)]
    for blk in net.residual_blocks:
        for conv, bn in ((blk.conv1, blk.norm1), (blk.conv2, blk.norm2)):
            w, b = _fold(conv, bn)
            sections += [_conv3x3_bytes(w, shape.channels), _f32(b)]
    ph, vh = net.policy_head, net.value_head
    w, b = _fold(ph.conv, ph.norm)
    sections += [_f32(w.reshape(2, shape.channels)), _f32(b), _f32(ph.linear.weight), _f32(ph.linear.bias)]
    w, b = _fold(vh.conv, vh.norm)
    sections += [_f32(w.reshape(shape.channels)), _f32(b), _f32(vh.linear1.weight), _f32(vh.linear1.bias),
                 _f32(vh.linear2.weight.reshape(-1)), _f32(vh.linear2.bias)]
    assert len(sections) == n.value, (len(sections), n.value)
    blob = bytearray(total)
    for off, sec in zip(offs[: n.value], sections):
        blob[off: off + len(sec)] = sec
    return shape, bytes(blob)


def _walk_tensors(module, out: list) -> None:
    for p in module._parameters.values():
        if p is not None:
            out.append((p.data_ptr(), p._version))
    for b in module._buffers.values():
        if b is not None:
            out.append((b.data_ptr(), b._version))
    for child in module._modules.values():
        if child is not None:
            _walk_tensors(child, out)


def weights_fingerprint(net) -> tuple:
    """Changes whenever a parameter or buffer is updated in place or replaced (re-fold trigger).  Runs before EVERY search
    (a single-game search is 5.9 ms), so it walks the module tree directly: `net.parameters()` / `net.buffers()` build
    names and de-duplication sets on the way and cost 0.2-0.4 ms for 128x10b, this 0.1 ms."""
    out: list = []
    _walk_tensors(unwrap(net), out)
    return tuple(out)


_fold_cache: tuple | None = None


def fold_cached(net, fingerprint=None) -> tuple[OzResnetShape, bytes]:
    """fold_alphazero_net, remembering the last result: several engines that load the same unchanged network (match
    play, one engine per leaf-batch size in a sweep) fold it once."""
    global _fold_cache
    fp = fingerprint if fingerprint is not None else weights_fingerprint(net)
    if _fold_cache is None or _fold_cache[0] != fp:
        shape, blob = fold_alphazero_net(net)
        _fold_cache = (fp, shape, blob)
    return _fold_cache[1], _fold_cache[2]
