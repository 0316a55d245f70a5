"""The optimizers and the learning-rate scheduler of ``Model.initialize_optimizer`` (ctunet/pytorch/Model.py:510-546)
as ONE device-side update over the flat gradient buffer.

  Adam / AdamW (amsgrad=True)     Model.py:514-527
  RMSprop (momentum)              Model.py:528-534
  SGD (momentum)                  Model.py:535-541
  ReduceLROnPlateau()             Model.py:544-546, stepped every ITERATION on the training loss (Model.py:369-371)

``torch.optim`` runs one (fused) or a dozen (foreach) library kernels per step and its scheduler reads the loss on the
host -- a synchronisation per iteration.  Here the whole update is one launch of ``ctu_optim_step`` (every parameter,
addressed through a chunk table built once) and a one-thread ``ctu_optim_post`` that advances the step counter and
applies torch's ``ReduceLROnPlateau.step(loss)`` rule to the learning rate, which lives in device memory.  Nothing
touches the host, so the optimizer and the scheduler sit inside the captured training step.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from . import _lib
from ._lib import call, stream_ptr

KINDS = {"adam": 0, "adamw": 1, "rmsprop": 2, "sgd": 3}


class FlatOptimizer:
    """State: ``lr`` / scheduler record = device double[10], ``step`` = device int64, moment buffers = flat fp32 tensors
    aligned with ``flat_grad``.  ``params[i]`` owns ``flat_grad[offsets[i] : offsets[i] + numel]``."""

    def __init__(self, params, flat_grad: torch.Tensor, kind: str = "adam", lr: float = 1e-4, weight_decay: float = 0.0,
                 momentum: float = 0.0, betas=(0.9, 0.999), eps: float = 1e-8, alpha: float = 0.99, amsgrad: bool = True,
                 plateau: bool = False, plateau_cfg=None):
        if kind not in KINDS:
            raise ValueError("optimizer %r (adam, adamw, rmsprop, sgd)" % (kind,))
        lib = _lib.load()
        self.kind, self.params, self.flat_grad = kind, list(params), flat_grad
        dev = flat_grad.device
        if not flat_grad.is_cuda:
            raise RuntimeError("FlatOptimizer runs on CUDA only (no CPU fallback)")
        self.hyper = dict(beta1=float(betas[0]), beta2=float(betas[1]), eps=float(eps), weight_decay=float(weight_decay),
                          momentum=float(momentum), alpha=float(alpha), amsgrad=int(bool(amsgrad)))
        n = sum(p.numel() for p in self.params)
        if n > flat_grad.numel():
            raise ValueError("flat gradient buffer is smaller than the parameters")
        z = lambda: torch.zeros(n, dtype=torch.float32, device=dev)
        k = KINDS[kind]
        self.state0 = z() if (k != 3 or momentum != 0) else None
        self.state1 = z() if (k <= 1 or (k == 2 and momentum > 0)) else None
        self.state2 = z() if (k <= 1 and amsgrad) else None
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        # torch.optim.lr_scheduler.ReduceLROnPlateau defaults (what Model.py:545 constructs)
        cfg = dict(factor=0.1, patience=10, threshold=1e-4, min_lr=0.0, cooldown=0, eps=1e-8)
        cfg.update(plateau_cfg or {})
        self.plateau = bool(plateau)
        self.sched = torch.tensor([lr, float("inf"), 0, 0, cfg["factor"], cfg["patience"], cfg["threshold"], cfg["min_lr"],
                                   cfg["cooldown"], cfg["eps"]], dtype=torch.float64, device=dev)
        # chunk table
        ce = lib.ctu_optim_chunk_elems()
        rec = lib.ctu_optim_chunk_bytes()
        if rec != 24:
            raise RuntimeError("unexpected chunk record size %d" % rec)
        buf = bytearray()
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise TypeError("parameters must be contiguous float32 tensors on %s" % dev)
            base, m = p.data_ptr(), p.numel()
            for s in range(0, m, ce):
                buf += struct.pack("<QqiI", base + 4 * s, off + s, min(ce, m - s), 0)
            off += m
        self.n_chunks = len(buf) // rec
        self.chunks = torch.from_numpy(np.frombuffer(bytes(buf), dtype=np.uint8).copy()).to(dev)
        self._param_ptrs = [p.data_ptr() for p in self.params]

    @property
    def lr(self) -> float:
        """Current learning rate (host read: synchronises)."""
        return float(self.sched[0])

    def step_peer(self, sync, n_tail: int, grad_scale: float = 1.0):
        """All-reduce + update in one launch over the peer-mapped gradient buffers of ``sync`` (parallel.PeerGradSync): the
        gradient of every element is the rank-ordered mean over all ranks; the first ``n_tail`` floats of the loss tail are
        averaged into ``sync.tail_avg``, whose last entry (the total loss) then feeds the plateau scheduler."""
        if [p.data_ptr() for p in self.params] != self._param_ptrs:
            raise RuntimeError("a parameter's storage moved after the optimizer was built (re-create the TrainStep)")
        h = self.hyper
        st = stream_ptr()
        ptr = lambda t: t.data_ptr() if t is not None else None
        call("ctu_optim_step_peer", KINDS[self.kind], self.chunks.data_ptr(), self.n_chunks, sync.h_grads, sync.h_flags,
             sync.world, sync.rank, ptr(self.state0), ptr(self.state1), ptr(self.state2), self.sched.data_ptr(),
             self.step_count.data_ptr(), h["beta1"], h["beta2"], h["eps"], h["weight_decay"], h["momentum"], h["alpha"],
             h["amsgrad"], float(grad_scale), sync.tail_avg.data_ptr(), sync.n_grad, int(n_tail), st)
        use = int(self.plateau)
        loss = sync.tail_avg[n_tail - 1:n_tail]
        call("ctu_optim_post", self.step_count.data_ptr(), self.sched.data_ptr(), loss.data_ptr() if use else None, use, st)

    def step(self, loss: torch.Tensor = None, grad_scale: float = 1.0):
        """Enqueue the update; ``loss`` (device float scalar) feeds the plateau scheduler AFTER the update, the order of
        Model.py:367-371.  No host synchronisation."""
        if [p.data_ptr() for p in self.params] != self._param_ptrs:
            raise RuntimeError("a parameter's storage moved after the optimizer was built (re-create the TrainStep)")
        h = self.hyper
        st = stream_ptr()
        ptr = lambda t: t.data_ptr() if t is not None else None
        call("ctu_optim_step", KINDS[self.kind], self.chunks.data_ptr(), self.n_chunks, self.flat_grad.data_ptr(),
             ptr(self.state0), ptr(self.state1), ptr(self.state2), self.sched.data_ptr(), self.step_count.data_ptr(),
             h["beta1"], h["beta2"], h["eps"], h["weight_decay"], h["momentum"], h["alpha"], h["amsgrad"], float(grad_scale), st)
        use = int(self.plateau and loss is not None)
        call("ctu_optim_post", self.step_count.data_ptr(), self.sched.data_ptr(), loss.data_ptr() if use else None, use, st)
