"""Data-parallel gradient synchronisation: one process per GPU, NCCL all-reduce over NVLink.

The reference's only multi-GPU mechanism is single-process ``nn.DataParallel``
(ctunet/pytorch/Model.py:481-486): scatter the batch, replicate the module, gather outputs, reduce
gradients onto GPU 0; BatchNorm statistics stay per replica.  The B200 equivalent is one process per
GPU with gradient averaging, which is mathematically the same step when the per-rank batches are
equal: the Dice term is a mean of per-sample terms (utilities.py:50) and the CE
term a mean over all voxels (ProblemHandler.py:251).  BatchNorm stays per-rank, as in the reference.

All live gradients of a step land in ONE flat fp32 buffer (UNetSP: 4.6 MB) split into a few buckets
in gradient-production order (head first, first encoder block last).  The engine writes each
parameter gradient straight into its slice; when the last gradient of a bucket has been written the
bucket is all-reduced (AVG) on a side stream while the remaining backward kernels keep running.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradSync:
    TAIL = 8      # floats appended to the flat buffer: the step's loss components, averaged together with the gradients

    def __init__(self, module: torch.nn.Module, process_group=None, n_buckets: int = 3, skip_prefixes=("cblock.",),
                 deferred: bool = False):
        """``skip_prefixes``: parameters that never receive a gradient (the generic UNet's discarded
        center block, models.py:241) -- they are left with ``grad = None`` exactly like the reference."""
        self.module = module
        self.group = process_group
        # deferred: no bucket is reduced while the backward pass runs; finish() reduces the whole flat buffer in one
        # call.  Used by the CUDA-graph step (trainer.py): forward+backward and the optimizer are two captured graphs
        # with ONE eager NCCL all-reduce (4.6 MB for UNetSP) between them.
        self.deferred = bool(deferred)
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if type(module).__name__ in ("recAE_v2_fixed", "UNet4_2IC"):
            skip_prefixes = ()
        live = [(n, p) for n, p in named if not n.startswith(tuple(skip_prefixes))] if skip_prefixes else named
        live = list(reversed(live))                       # production order of the backward pass
        total = sum(p.numel() for _, p in live)
        ref = live[0][1]
        # [gradients | TAIL]: the reference computes its loss on the GATHERED batch (nn.DataParallel, Model.py:486), i.e. the
        # mean over ranks of the per-rank losses -- the tail carries the loss components through the same all-reduce, so
        # logging and ReduceLROnPlateau (Model.py:369-371) see identical values on every rank.
        self.n_grad = total
        self.flat = self._alloc_flat(total + self.TAIL, ref.device)
        self.tail = self.flat[total:]
        self.slices: Dict[int, torch.Tensor] = {}
        self.bucket_of: Dict[int, int] = {}
        self.names = [n for n, _ in live]
        per = (total + n_buckets - 1) // n_buckets
        self.bucket_ranges: List[List[int]] = []
        off, b_start, b = 0, 0, 0
        for i, (n, p) in enumerate(live):
            self.slices[id(p)] = self.flat[off:off + p.numel()].view_as(p)
            self.bucket_of[id(p)] = b
            off += p.numel()
            last = i == len(live) - 1
            if off - b_start >= per or last:
                self.bucket_ranges.append([b_start, off])
                b_start = off
                b += 1
        self.bucket_ranges[-1][1] = total + self.TAIL     # the loss tail travels with the last bucket
        self.n_in_bucket = [0] * len(self.bucket_ranges)
        for k, bi in self.bucket_of.items():
            self.n_in_bucket[bi] += 1
        self.params = [p for _, p in live]
        self.cuda = self.flat.is_cuda
        self.comm_stream = torch.cuda.Stream(device=self.flat.device) if self.cuda else None
        self._remaining: List[int] = []
        self._handles = []
        self.begin_step()

    def _alloc_flat(self, n: int, device) -> torch.Tensor:
        return torch.zeros(n, dtype=torch.float32, device=device)

    # -- engine-facing ---------------------------------------------------------------------------
    def buffer_for(self, param) -> Optional[torch.Tensor]:
        """Slice of the flat buffer the gradient kernels write into (None: not a synced parameter)."""
        return self.slices.get(id(param))

    def delivered(self, param) -> None:
        """The gradient of ``param`` is fully written (on the current stream)."""
        bi = self.bucket_of.get(id(param))
        if bi is None:
            return
        if self.cuda:                                      # gradients of one bucket may come from several streams
            self._producers[bi].add(torch.cuda.current_stream())
        self._remaining[bi] -= 1
        if self._remaining[bi] == 0 and not self.deferred:
            self._launch(bi)

    # -- step protocol ---------------------------------------------------------------------------
    def begin_step(self) -> None:
        self._remaining = list(self.n_in_bucket)
        self._handles = []
        self._producers = [set() for _ in self.n_in_bucket]

    def _launch(self, bi: int) -> None:
        lo, hi = self.bucket_ranges[bi]
        view = self.flat[lo:hi]
        if self.world == 1:
            return
        if self.cuda:
            for st in self._producers[bi] | {torch.cuda.current_stream()}:
                self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo (CPU tests): no AVG, no streams
            h = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._handles.append((h, view))

    def check_complete(self) -> None:
        if any(r != 0 for r in self._remaining):
            missing = [n for n, p in zip(self.names, self.params) if self._remaining[self.bucket_of[id(p)]] != 0]
            raise RuntimeError("gradient sync: some gradients were never produced, e.g. %s" % missing[:3])

    def reduce_all(self) -> None:
        """One all-reduce (average) of the whole flat buffer on the current stream (deferred mode)."""
        if self.world > 1:
            if self.cuda:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.div_(self.world)

    def publish(self) -> None:
        """``param.grad`` = view of the flat buffer."""
        for p in self.params:
            p.grad = self.slices[id(p)]

    def finish(self, publish: bool = True) -> None:
        """Make the averaged gradients visible to the optimizer (current stream) and, with ``publish``, expose them as
        ``param.grad`` views of the flat buffer (the ``FlatOptimizer`` reads the buffer itself)."""
        self.check_complete()
        if self.deferred:
            self.reduce_all()
        else:
            if self.cuda:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
            for h, view in self._handles:
                h.wait()
                view.div_(self.world)
        if publish:
            self.publish()
        self.begin_step()


class _RawCudaBuffer:
    """A caller-owned device allocation seen through ``__cuda_array_interface__`` (zero-copy ``torch.as_tensor``)."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerGradSync(GradSync):
    """Gradient exchange over NVLink peer memory, fused with the optimizer (csrc/peer.cu): every rank's flat gradient
    buffer sits in an IPC-shared allocation that all ranks of the node map; the optimizer kernel reads element f from ALL
    ranks, averages in rank order and applies the update in one pass.  No NCCL call inside the step -- the whole iteration
    (forward, loss, backward, exchange, optimizer, scheduler) is ONE CUDA graph, and nothing on the host sits between the
    ranks.  ``torch.distributed`` is only used to hand the IPC handles around at construction."""

    def __init__(self, module: torch.nn.Module, process_group=None, skip_prefixes=("cblock.",)):
        from . import _lib
        self._lib = _lib
        self._base = None
        super().__init__(module, process_group, n_buckets=1, skip_prefixes=skip_prefixes, deferred=True)
        if not self.cuda:
            raise RuntimeError("PeerGradSync needs CUDA devices")
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        lib = _lib.load()
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(self._handle), group=process_group)
        else:
            handles = [bytes(self._handle)]
        self._opened = []
        bases = []
        for r, h in enumerate(handles):
            if r == self.rank:
                bases.append(self._base)
                continue
            import ctypes
            out = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            rc = lib.ctu_peer_open(buf, ctypes.byref(out))
            if rc != 0:
                raise RuntimeError("ctu_peer_open(rank %d) failed (%d): %s -- peer access over NVLink is required" % (r, rc, _lib.last_error()))
            self._opened.append(out.value)
            bases.append(out.value)
        fb = lib.ctu_peer_flag_bytes()
        self.h_flags = _lib.ptr_array(bases)
        self.h_grads = _lib.ptr_array([b + fb for b in bases])
        self.tail_avg = torch.zeros(self.TAIL, dtype=torch.float32, device=self.flat.device)
        if self.world > 1:
            dist.barrier(group=process_group)           # every rank has mapped every buffer before the first step

    def _alloc_flat(self, n: int, device) -> torch.Tensor:
        import ctypes
        lib = self._lib.load()
        fb = lib.ctu_peer_flag_bytes()
        ptr = ctypes.c_void_p()
        self._handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            rc = lib.ctu_peer_alloc(fb + 4 * n, ctypes.byref(ptr), self._handle)
        if rc != 0:
            raise RuntimeError("ctu_peer_alloc failed (%d): %s" % (rc, self._lib.last_error()))
        self._base = ptr.value
        self._raw = _RawCudaBuffer(self._base + fb, n)
        return torch.as_tensor(self._raw, device=device)

    # the exchange is inside the optimizer kernel: nothing to do between backward and the update
    def reduce_all(self) -> None:
        pass

    def wait_released(self) -> None:
        """Enqueue: every peer has finished reading this rank's previous gradients (before the buffer is written again)."""
        self._lib.call("ctu_peer_wait_done", self.h_grads, self.h_flags, self.world, self.rank, self._lib.stream_ptr())

    def signal(self) -> None:
        """Enqueue: this rank's gradients (and loss tail) of the current step are complete."""
        self._lib.call("ctu_peer_signal", self.h_grads, self.h_flags, self.world, self.rank, self._lib.stream_ptr())

    def error(self) -> int:
        return int(self._lib.load().ctu_peer_error(self._base))

    def close(self) -> None:
        lib = self._lib.load()
        torch.cuda.synchronize()
        for p in self._opened:
            lib.ctu_peer_close(p)
        self._opened = []


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced partition of ``n_items`` independent units (patches / volumes) over ranks:
    inference and preprocessing shard with no data-path collective."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
