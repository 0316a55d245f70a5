"""One training iteration of the reference's step driver on the B200 path.

Restates the 'train' branch of ``Model.forward_pass`` (ctunet/pytorch/Model.py:342-374):
H2D copy, ``input.requires_grad_()``, forward, ``comp_losses_metrics``, ``backward``,
``optimizer.step()``, ``param.grad = None`` -- with the optimizer of ``Model.initialize_optimizer``
(Model.py:510-520: Adam, amsgrad=True).  The reference reads every loss component back with
``float(...)`` (five host syncs per batch, ProblemHandler.py:253-302); here the step enqueues
everything and returns device tensors, the caller decides when to read them (one sync).

``graph=True``: the step is a fixed sequence of ~300 launches with no host
synchronisation, so after ``GRAPH_WARMUP`` eager iterations it is captured ONCE in a CUDA graph and every
later call is a copy of the batch into the graph's static input buffers plus one ``cudaGraphLaunch`` -- the
Python / ctypes dispatch cost (longer than the GPU work at 128^3) disappears from the step.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .losses import dice_ce
from .parallel import GradSync

GRAPH_WARMUP = 2      # eager iterations before the capture (allocator pools, lazy optimizer state, cached constants)


class TrainStep:
    def __init__(self, model, handler: str = "double", dice_lambda: float = 1.0, ce_lambda: float = 1.0,
                 lr: float = 1e-4, weight_decay: float = 0.0, optimizer: str = "adam",
                 grad_sync: Optional[GradSync] = None, input_requires_grad: bool = True, graph: bool = False):
        if handler not in ("double", "single"):
            raise ValueError("handler: 'double' (FlapRecWithShapePriorDoubleOut) or 'single' (ProblemHandler)")
        self.model = model
        self.handler = handler
        self.dice_lambda, self.ce_lambda = float(dice_lambda), float(ce_lambda)
        self.grad_sync = grad_sync
        self.input_requires_grad = input_requires_grad       # Model.py:351-352
        model._grad_sink = grad_sync
        params = list(model.parameters())
        self.graph = bool(graph)
        self._graph = None
        self._graph_opt = None
        self._static = None
        self._static_out = None
        self._calls = 0
        self.launches_per_step = None                         # C-ABI calls recorded in the captured step
        # fused=True: one multi-tensor kernel per step instead of the ~12 foreach launches (0.25 ms at the END of the step,
        # where nothing overlaps them); same update rule, fp32 math
        cap = dict(capturable=True, fused=True) if self.graph else dict(fused=True)
        if optimizer == "adam":                               # Model.py:514-520
            self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, amsgrad=True, **cap)
        elif optimizer == "adamw":                            # Model.py:521-527
            self.optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, amsgrad=True, **cap)
        elif optimizer == "sgd":                              # Model.py:535-541
            self.optimizer = torch.optim.SGD(params, lr=lr, momentum=0.99, weight_decay=weight_decay, fused=True)
        else:
            raise ValueError("optimizer %r" % optimizer)

    def loss(self, out, target):
        """Weighted sum in the reference's order; returns (total, stacked components)."""
        ce_on, terms = self.ce_lambda != 0, []
        if self.handler == "double":                          # ProblemHandler.py:228-298
            (sk_p, fl_p), (sk_t, fl_t) = out, target
            ce_s, d_s = dice_ce(sk_p, sk_t, True, ce_on)
            ce_f, d_f = dice_ce(fl_p, fl_t, True, ce_on)
            if ce_on:
                terms += [self.ce_lambda * ce_s, self.ce_lambda * ce_f]
            if self.dice_lambda != 0:
                terms += [self.dice_lambda * d_s, self.dice_lambda * d_f]
        else:                                                 # ProblemHandler.py:59-91
            ce, d = dice_ce(out, target, False, ce_on)
            if ce_on:
                terms.append(self.ce_lambda * ce)
            if self.dice_lambda != 0:
                terms.append(self.dice_lambda * d)
        total = sum(terms)
        return total, torch.stack([t.detach() for t in terms] + [total.detach()])

    def __call__(self, image: torch.Tensor, target):
        """Enqueues one full iteration; returns the device tensor [components..., total] (no host sync).
        In graph mode the returned tensor is the graph's static output: read it before the next call."""
        if not self.graph:
            return self._eager(image, target)
        self._calls += 1
        if self._graph is None and self._calls <= GRAPH_WARMUP:
            return self._eager(image, target)
        flat = [image] + (list(target) if isinstance(target, (tuple, list)) else [target])
        if self._graph is None:
            self._static = [torch.empty_like(t) for t in flat]
        elif any(a.shape != b.shape or a.dtype != b.dtype for a, b in zip(flat, self._static)):
            raise RuntimeError("graph mode: the batch shape changed after the step was captured")
        for dst, src in zip(self._static, flat):
            if dst.data_ptr() != src.data_ptr():              # a pipeline may fill static_inputs() in place
                dst.copy_(src, non_blocking=True)
        if self._graph is None:
            st_img = self._static[0]
            st_tgt = tuple(self._static[1:]) if isinstance(target, (tuple, list)) else self._static[1]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launches
            if self.grad_sync is None:
                with torch.cuda.graph(g):                     # records the launches; nothing executes here
                    self._static_out = self._eager(st_img, st_tgt)
            else:
                # data parallel: graph 1 = forward + loss + backward (gradients land in the flat buffer), ONE eager
                # NCCL all-reduce of that buffer, graph 2 = optimizer step on views of the flat buffer
                if not self.grad_sync.deferred:
                    raise RuntimeError("graph mode needs GradSync(deferred=True)")
                with torch.cuda.graph(g):
                    self._static_out = self._forward_backward(st_img, st_tgt)
                self.grad_sync.check_complete()
                self.grad_sync.begin_step()
                self.grad_sync.publish()
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    self.optimizer.step()
                self._graph_opt = g2
                for p in self.model.parameters():
                    p.grad = None
            self.launches_per_step = _lib.launches - l0
            self._graph = g
        return self._replay()

    def static_inputs(self):
        """(image, target) buffers the captured graph reads, or None before the capture: a data pipeline that writes
        the next batch straight into them (and passes them to ``__call__``) saves the per-step device copy."""
        if self._graph is None:
            return None
        tgt = tuple(self._static[1:]) if len(self._static) > 2 else self._static[1]
        return self._static[0], tgt

    def _replay(self):
        self._graph.replay()
        if self.grad_sync is not None:
            self.grad_sync.reduce_all()
            self._graph_opt.replay()
        return self._static_out

    def step_from_masks(self, broken: torch.Tensor, full: torch.Tensor, flap: torch.Tensor, atlas=None):
        """One iteration from the uint8 masks of a batch ([B,D,H,W] each, on the device): the float image (+ atlas
        channel) and the two one-hot targets of datasets.py:195-235 are produced by ``ctu_encode_flaprec_u8`` --
        straight into the captured graph's static inputs once the step is captured -- so a training loop ships
        3 bytes per voxel over PCIe instead of 24."""
        from .utilities import encode_flaprec_batch
        if self.handler != "double":
            raise ValueError("step_from_masks feeds the double-output handler (full skull + flap targets)")
        if self._graph is not None:
            encode_flaprec_batch(broken, full, flap, atlas, out=(self._static[0], (self._static[1], self._static[2])))
            return self._replay()
        image, target = encode_flaprec_batch(broken, full, flap, atlas)
        return self(image, target)

    def _forward_backward(self, image: torch.Tensor, target):
        self.model.train()
        if self.input_requires_grad:
            image = image.detach().requires_grad_()
        out = self.model(image)
        total, comps = self.loss(out, target)
        total.backward()
        return comps

    def _eager(self, image: torch.Tensor, target):
        self.model.train()
        if self.input_requires_grad:
            image = image.detach().requires_grad_()
        out = self.model(image)
        total, comps = self.loss(out, target)
        total.backward()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        self.optimizer.step()
        for p in self.model.parameters():                     # Model.py:373-374
            p.grad = None
        return comps


class LossReadback:
    """Deferred read-back of the loss components.

    The reference reads every component with ``float(...)`` as soon as it exists (ProblemHandler.py:253-302): five host
    syncs per batch, each idling the GPU until the host has enqueued the next kernels.  Here every step's component
    tensor is copied asynchronously into pinned host memory on the step's stream; ``push`` returns the values of the
    PREVIOUS step (already complete, or waited for while the current step runs), ``drain`` the last one.  The
    ``losses_and_metrics`` lists (Model.py:363, ProblemHandler.py:71-102) receive the same floats, one step later."""

    def __init__(self, n_values: int, depth: int = 2):
        self.host = [torch.empty(n_values, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0
        self.bytes_per_step = 4 * n_values

    def push(self, comps: torch.Tensor):
        slot = self.count % len(self.host)
        self.host[slot].copy_(comps, non_blocking=True)
        self.events[slot].record()
        self.count += 1
        lag = len(self.host) - 1                # depth 2: the previous step; depth 3: two steps back, ...
        return self._read(self.count - 1 - lag) if self.count > lag else None

    def drain(self):
        """Values of every step not yet returned by ``push`` (oldest first); the last entry is the final step's."""
        lag = len(self.host) - 1
        out = [self._read(i) for i in range(max(0, self.count - lag), self.count)]
        return out[-1] if out else None

    def _read(self, index: int):
        slot = index % len(self.host)
        self.events[slot].synchronize()
        return self.host[slot].tolist()
