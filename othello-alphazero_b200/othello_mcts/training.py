"""Native training step for `AlphaZeroNet`.

reference: python/othello_alphazero/train.py:455-521 (`_train`): per batch `optimizer.zero_grad()`, forward in train
mode, `policy_loss + value_loss + l2_loss`, `backward()`, `optimizer.step()` with `torch.optim.SGD(lr, momentum)`.
`NativeTrainer` does the body of that loop in one C-ABI call (`oz_trainer_step`, include/othello_b200.h): the 3x3
convolutions (forward and input gradients) on the engine's tcgen05 kernel in bf16, everything else in fp32, master
weights in fp32.  Parameters travel as one flat vector in `net.parameters()` order, so moving them between the torch
module and the trainer is `parameters_to_vector` / `vector_to_parameters`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, weights
from ._lib import OzTrainConfig, check


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _batch_norms(net):
    return [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]


class NativeTrainer:
    """SGD training of an `AlphaZeroNet` on the GPU.  `step()` = one iteration of the loop of train.py:486-502."""

    def __init__(self, net, lr: float = 0.001, momentum: float = 0.9, l2_weight: float = 1e-4, max_batch: int = 256,
                 cuda_device: int = 0) -> None:
        net = getattr(net, "_orig_mod", net)
        if not weights.is_alphazero_net(net):
            raise TypeError("NativeTrainer needs an AlphaZeroNet")
        self._lib = _lib.load()
        self.shape = weights.shape_of(net)
        self.cfg = OzTrainConfig()
        self._lib.oz_train_config_default(C.byref(self.cfg))
        self.cfg.lr, self.cfg.momentum, self.cfg.l2_weight = lr, momentum, l2_weight
        bns = _batch_norms(net)
        if bns:
            self.cfg.bn_momentum, self.cfg.bn_eps = bns[0].momentum, bns[0].eps
        self.max_batch = (max_batch + 31) // 32 * 32
        self._inflight = []
        self._steps_since_store = 0      # train-mode forwards since load_from / store_to (BatchNorm num_batches_tracked)
        self._h = C.c_void_p()
        check(self._lib.oz_trainer_create(cuda_device, C.byref(self.shape), self.max_batch, C.byref(self._h)))
        self.num_params = self._lib.oz_trainer_num_params(self._h)
        self.num_buffers = self._lib.oz_trainer_num_buffers(self._h)
        self.load_from(net)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.oz_trainer_destroy(h)                 # synchronises the trainer's stream first
            self._h = None
        self._inflight = []

    # -- state exchange with the torch module ------------------------------------------------------------------
    def load_from(self, net, momentum_buffers=None) -> None:
        """Copy parameters and BatchNorm running statistics from `net` (and optionally SGD momentum buffers, a flat
        vector in parameter order, e.g. from optimizer.state_dict())."""
        net = getattr(net, "_orig_mod", net)
        p = torch.nn.utils.parameters_to_vector([q.detach().float().cpu() for q in net.parameters()]).numpy()
        if p.size != self.num_params:
            raise ValueError(f"network has {p.size} parameters, the trainer expects {self.num_params}")
        b = np.concatenate([np.concatenate([m.running_mean.detach().float().cpu().numpy(), m.running_var.detach().float().cpu().numpy()])
                            for m in _batch_norms(net)]).astype(np.float32)
        if b.size != self.num_buffers:
            raise ValueError(f"network has {b.size} BatchNorm buffer entries, the trainer expects {self.num_buffers}")
        p = np.ascontiguousarray(p, np.float32)
        mom = None if momentum_buffers is None else np.ascontiguousarray(momentum_buffers, np.float32)
        check(self._lib.oz_trainer_set_state(self._h, _ptr(p), _ptr(b), None if mom is None else _ptr(mom)))
        self._steps_since_store = 0

    def store_to(self, net) -> None:
        """Write parameters and running statistics back into `net` (e.g. before self-play or a checkpoint)."""
        net = getattr(net, "_orig_mod", net)
        p, b = np.zeros(self.num_params, np.float32), np.zeros(self.num_buffers, np.float32)
        check(self._lib.oz_trainer_get_state(self._h, _ptr(p), _ptr(b), None, None))
        with torch.no_grad():
            # in-place copies (p.copy_), not vector_to_parameters: that rebinds .data without bumping _version, and the
            # engine's re-fold trigger (weights.weights_fingerprint) keys on (data_ptr, _version)
            flat, o = torch.from_numpy(p), 0
            for q in net.parameters():
                k = q.numel()
                q.copy_(flat[o:o + k].view_as(q))
                o += k
            o = 0
            for m in _batch_norms(net):
                c = m.num_features
                m.running_mean.copy_(torch.from_numpy(b[o:o + c])); m.running_var.copy_(torch.from_numpy(b[o + c:o + 2 * c]))
                m.num_batches_tracked += self._steps_since_store      # what BatchNorm2d.forward counts in train mode
                o += 2 * c
        self._steps_since_store = 0

    def gradients(self) -> np.ndarray:
        """Gradients of the last step (flat, parameter order), L2 term included."""
        g = np.zeros(self.num_params, np.float32)
        check(self._lib.oz_trainer_get_state(self._h, None, None, None, _ptr(g)))
        return g

    def momentum_buffers(self) -> np.ndarray:
        m = np.zeros(self.num_params, np.float32)
        check(self._lib.oz_trainer_get_state(self._h, None, None, _ptr(m), None))
        return m

    # -- the step ----------------------------------------------------------------------------------------------
    def step(self, features, policy, value, apply_update: bool = True) -> dict:
        """One batch: features [n,1+2H,8,8], target policies [n,65], target values [n] (CPU tensors / arrays, as the
        DataLoader of train.py:472-489 yields them; CUDA tensors are used in place).  n must be a multiple of 32.
        Returns the four losses train.py:504-507 logs."""
        losses = np.zeros(4, np.float32)
        n = int(features.shape[0])
        self._steps_since_store += 1
        if isinstance(features, torch.Tensor) and features.is_cuda:
            f, p, v = (t.detach().to(torch.float32).contiguous() for t in (features, policy, value))
            torch.cuda.current_stream(f.device).synchronize()    # the tensors are complete; the call below blocks until the losses are back
            check(self._lib.oz_trainer_step_device(self._h, C.c_void_p(f.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr()), n,
                                                   C.byref(self.cfg), int(apply_update), _ptr(losses)))
        else:
            f = np.ascontiguousarray(np.asarray(features, dtype=np.float32))
            p = np.ascontiguousarray(np.asarray(policy, dtype=np.float32))
            v = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
            check(self._lib.oz_trainer_step(self._h, _ptr(f), _ptr(p), _ptr(v), n, C.byref(self.cfg), int(apply_update), _ptr(losses)))
        return {"total_loss": float(losses[0]), "policy_loss": float(losses[1]), "value_loss": float(losses[2]), "l2_loss": float(losses[3])}

    def step_async(self, features, policy, value) -> None:
        """`step` for CUDA tensors without reading the losses back (no host synchronisation: steps queue up on the
        trainer's stream).  The trainer keeps the tensors alive until the step that reads them has finished, so the
        caller may drop them."""
        n = int(features.shape[0])
        self._steps_since_store += 1
        f, p, v = (t.detach().to(torch.float32).contiguous() for t in (features, policy, value))
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(f.device))
        ts = torch.cuda.ExternalStream(self._lib.oz_trainer_stream(self._h), device=f.device)
        ts.wait_event(ev)                                    # the trainer's stream reads the tensors after they are produced
        check(self._lib.oz_trainer_step_device(self._h, C.c_void_p(f.data_ptr()), C.c_void_p(p.data_ptr()), C.c_void_p(v.data_ptr()), n,
                                               C.byref(self.cfg), 1, None))
        # the inputs must outlive the queued copy: hold them until the step that consumed them has finished
        done = torch.cuda.Event()
        done.record(ts)
        self._inflight.append((done, f, p, v))
        while self._inflight and self._inflight[0][0].query():
            self._inflight.pop(0)

    def set_graphs(self, on: bool) -> None:
        """CUDA-graph replay of the step (default on; identical results)."""
        check(self._lib.oz_trainer_set_graphs(self._h, int(on)))

    def outputs(self, n: int):
        """(policy [n,65], value [n]) of the last step's forward pass (train mode)."""
        p, v = np.zeros((n, 65), np.float32), np.zeros(n, np.float32)
        check(self._lib.oz_trainer_outputs(self._h, n, _ptr(p), _ptr(v)))
        return p, v
