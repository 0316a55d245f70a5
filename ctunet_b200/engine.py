"""Host-side executor of the U-Net hot path: a tape of fused stages over blocked activations.

Every stage is one or a few calls into the C ABI (include/ctunet_b200.h); PyTorch only owns the
memory and the stream.  A forward pass in training mode records a tape of backward closures; the
whole network is exposed to autograd as ONE ``torch.autograd.Function`` (see models.py), so a
training step is a fixed sequence of kernel launches with no host synchronisation and can be
captured in a CUDA graph.

Reference behaviour restated here (file:line relative to the reference root):
  * conv -> BatchNorm3d(train: batch statistics) -> ReLU stages   ctunet/pytorch/models.py:25-46
  * MaxPool3d(2,2) after every down block                           models.py:190-191, 233
  * channel concat of (up-block output, skip) folded into the consumer   models.py:249, 528-534
  * reentrant-checkpoint double update of BatchNorm running stats   models.py:232
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import CTU_BF16, CTU_F32, call, int_array, ptr_array, stream_ptr

BN_MOMENTUM = 0.1
BN_EPS = 1e-5

# "auto": tcgen05 implicit GEMM where the kernel covers the shape, CUDA-core direct kernel elsewhere;
# "direct": CUDA-core kernels only (always the case in fp32 accumulate-check mode).
CONV_PATH = "auto"


class Act:
    """A channel-blocked activation [N][Cb][D][H][W][8] (bf16 in product mode, fp32 in check mode)."""
    __slots__ = ("buf", "c", "n", "d", "h", "w", "sums")

    def __init__(self, buf, c, n, d, h, w):
        self.buf, self.c, self.n, self.d, self.h, self.w = buf, c, n, d, h, w
        self.sums = None        # per-channel sum / sum of squares written by the producing conv's epilogue

    @property
    def cb(self):
        return (self.c + 7) // 8

    @property
    def spatial(self):
        return self.d * self.h * self.w

    @property
    def ptr(self):
        return self.buf.data_ptr()


class Engine:
    def __init__(self, device, compute_dtype: str, record: bool):
        if compute_dtype not in ("bf16", "fp32"):
            raise ValueError("compute dtype must be 'bf16' or 'fp32'")
        _lib.load()
        self.device = device
        self.tdtype = torch.bfloat16 if compute_dtype == "bf16" else torch.float32
        self.dtype = CTU_BF16 if compute_dtype == "bf16" else CTU_F32
        self.record = record
        self.tape: List[Callable[[], None]] = []
        self.agrads: Dict[int, Act] = {}            # id(Act) -> gradient Act
        self.pgrads: Dict[int, torch.Tensor] = {}   # id(param) -> gradient tensor
        self.want_input_grad = False
        self.grad_sink = None                       # parallel.GradSync: gradients land in its flat buffer
        self.use_tc = (CONV_PATH == "auto" and compute_dtype == "bf16" and bool(_lib.load().ctu_has_tensor_path()))

    # ------------------------------------------------------------------ helpers
    def new_act(self, c, n, d, h, w) -> Act:
        cb = (c + 7) // 8
        return Act(torch.empty((n, cb, d, h, w, 8), dtype=self.tdtype, device=self.device), c, n, d, h, w)

    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    def _grad_buffer(self, param) -> torch.Tensor:
        """Where a parameter-gradient kernel writes: a slice of the data-parallel flat buffer, or a new tensor."""
        if self.grad_sink is not None:
            buf = self.grad_sink.buffer_for(param)
            if buf is not None:
                return buf
        return torch.empty_like(param)

    def _add_pgrad(self, param, g):
        k = id(param)
        if k in self.pgrads:
            raise RuntimeError("internal: parameter received two gradients")
        self.pgrads[k] = g
        if self.grad_sink is not None:
            self.grad_sink.delivered(param)

    @staticmethod
    def _src_args(srcs: Sequence[Act]):
        return ptr_array([s.ptr for s in srcs]), int_array([s.c for s in srcs]), len(srcs)

    def _tc_ok(self, k, srcs, cout):
        return self.use_tc and tc_supported(k, [s.c for s in srcs], cout, srcs[0].d, srcs[0].h, srcs[0].w)

    def _conv_launch(self, srcs, wp, bias, y: Act, cout, k, sums):
        """One forward-style convolution launch (also used for the data gradient): tcgen05 implicit GEMM when
        the kernel covers the shape, CUDA-core direct kernel otherwise."""
        s0 = srcs[0]
        pa, ca, ns = self._src_args(srcs)
        tc = self._tc_ok(k, srcs, cout)
        wptr = wp.data_ptr()
        if tc:
            lib = _lib.load()
            wimg = torch.empty(lib.ctu_conv_tc_wimg_bytes(k, s0.c, cout), dtype=torch.uint8, device=self.device)
            call("ctu_conv_tc_pack_weight", wp.data_ptr(), wimg.data_ptr(), k, s0.c, cout, stream_ptr())
            wptr = wimg.data_ptr()
        call("ctu_conv3d_fprop", self.dtype, pa, ca, ns, wptr, bias.data_ptr() if bias is not None else None,
             y.ptr, sums.data_ptr() if sums is not None else None, cout, k, s0.n, s0.d, s0.h, s0.w, int(tc),
             stream_ptr())

    # ------------------------------------------------------------------ layout
    def pack(self, x: torch.Tensor) -> Act:
        n, c, d, h, w = x.shape
        a = self.new_act(c, n, d, h, w)
        call("ctu_pack_ncdhw", x.data_ptr(), a.ptr, self.dtype, n, c, d * h * w, stream_ptr())
        return a

    def unpack(self, a: Act) -> torch.Tensor:
        out = self.f32(a.n, a.c, a.d, a.h, a.w)
        call("ctu_unpack_ncdhw", a.ptr, out.data_ptr(), self.dtype, a.n, a.c, a.spatial, stream_ptr())
        return out

    # ------------------------------------------------------------------ Conv3d
    def conv(self, srcs: Sequence[Act], weight, bias, k: int, need_src_grad: Sequence[bool],
             bn_stats: bool = False) -> Act:
        """``bn_stats``: also produce the batch statistics of the output (``y.sums``) for the BatchNorm that follows."""
        cout = weight.shape[0]
        s0 = srcs[0]
        pa, ca, ns = self._src_args(srcs)
        lib = _lib.load()
        wp = self.f32(lib.ctu_conv_wpack_floats(cout, k, ns, ca))
        call("ctu_conv_pack_weight", weight.data_ptr(), wp.data_ptr(), cout, k, ns, ca, stream_ptr())
        y = self.new_act(cout, s0.n, s0.d, s0.h, s0.w)
        if bn_stats:
            y.sums = self.f64(2 * y.cb * 8)
        self._conv_launch(srcs, wp, bias, y, cout, k, y.sums)
        if self.record:
            srcs = list(srcs)
            need = list(need_src_grad)

            def bwd():
                dy = self.agrads.pop(id(y))
                pa, ca, ns = self._src_args(srcs)
                if weight.requires_grad:
                    dwp = torch.empty_like(wp)
                    db = self._grad_buffer(bias) if (bias is not None and bias.requires_grad) else None
                    call("ctu_conv3d_wgrad", self.dtype, pa, ca, ns, dy.ptr, dwp.data_ptr(),
                         db.data_ptr() if db is not None else None, cout, k, s0.n, s0.d, s0.h, s0.w,
                         int(self.use_tc and len(srcs) == 1 and bool(
                             lib.ctu_conv_tc_wgrad_supported(k, s0.c, cout, s0.d, s0.h, s0.w))), stream_ptr())
                    dw = self._grad_buffer(weight)
                    call("ctu_conv_unpack_wgrad", dwp.data_ptr(), dw.data_ptr(), cout, k, ns, ca, stream_ptr())
                    self._add_pgrad(weight, dw)
                    if db is not None:
                        self._add_pgrad(bias, db)
                for i, s in enumerate(srcs):
                    if not need[i]:
                        continue
                    wpd = self.f32(lib.ctu_conv_wpack_dgrad_floats(cout, k, s.c))
                    call("ctu_conv_pack_weight_dgrad", weight.data_ptr(), wpd.data_ptr(), cout, k, ns, ca, i, stream_ptr())
                    dx = self.new_act(s.c, s.n, s.d, s.h, s.w)
                    self._conv_launch([dy], wpd, None, dx, s.c, k, None)
                    self._set_agrad(s, dx)

            self.tape.append(bwd)
        return y

    def _set_agrad(self, act: Act, g: Act):
        if id(act) in self.agrads:
            raise RuntimeError("internal: activation received two gradients")
        self.agrads[id(act)] = g

    # ------------------------------------------------------------------ ConvTranspose3d k2 s2
    def convt(self, srcs: Sequence[Act], weight, bias, need_src_grad: Sequence[bool]) -> Act:
        cout = weight.shape[1]
        s0 = srcs[0]
        pa, ca, ns = self._src_args(srcs)
        lib = _lib.load()
        wp = self.f32(lib.ctu_convt_wpack_floats(cout, ns, ca))
        call("ctu_convt_pack_weight", weight.data_ptr(), wp.data_ptr(), cout, ns, ca, stream_ptr())
        y = self.new_act(cout, s0.n, 2 * s0.d, 2 * s0.h, 2 * s0.w)
        call("ctu_convt2_fprop", self.dtype, pa, ca, ns, wp.data_ptr(), bias.data_ptr() if bias is not None else None,
             y.ptr, cout, s0.n, s0.d, s0.h, s0.w, stream_ptr())
        if self.record:
            srcs = list(srcs)
            need = list(need_src_grad)

            def bwd():
                dy = self.agrads.pop(id(y))
                pa, ca, ns = self._src_args(srcs)
                if weight.requires_grad:
                    dwp = torch.empty_like(wp)
                    db = self._grad_buffer(bias) if (bias is not None and bias.requires_grad) else None
                    call("ctu_convt2_wgrad", self.dtype, pa, ca, ns, dy.ptr, dwp.data_ptr(),
                         db.data_ptr() if db is not None else None, cout, s0.n, s0.d, s0.h, s0.w, stream_ptr())
                    dw = self._grad_buffer(weight)
                    call("ctu_convt_unpack_wgrad", dwp.data_ptr(), dw.data_ptr(), cout, ns, ca, stream_ptr())
                    self._add_pgrad(weight, dw)
                    if db is not None:
                        self._add_pgrad(bias, db)
                for i, s in enumerate(srcs):
                    if not need[i]:
                        continue
                    wpd = self.f32(lib.ctu_convt_wpack_dgrad_floats(cout, s.c))
                    call("ctu_convt_pack_weight_dgrad", weight.data_ptr(), wpd.data_ptr(), cout, ns, ca, i, stream_ptr())
                    dx = self.new_act(s.c, s.n, s.d, s.h, s.w)
                    call("ctu_convt2_dgrad", self.dtype, dy.ptr, wpd.data_ptr(), dx.ptr, cout, s.c,
                         s.n, s.d, s.h, s.w, stream_ptr())
                    self._set_agrad(s, dx)

            self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ BatchNorm3d + ReLU (+ MaxPool)
    def bn_relu(self, y: Act, bn, training: bool, extra_updates: int = 0, pool: bool = False):
        """Returns ``a`` or ``(a, pooled)``.  ``extra_updates``: additional running-stat updates applied
        when the backward pass runs (the reentrant-checkpoint recomputation of the reference)."""
        c = y.c
        cpad = y.cb * 8
        count = float(y.n * y.spatial)
        ss = self.f32(4 * cpad)
        sums = None
        st = stream_ptr()
        if training:
            sums = y.sums
            if sums is None:
                sums = self.f64(2 * cpad)
                call("ctu_bn_stats", self.dtype, y.ptr, c, y.n, y.spatial, sums.data_ptr(), st)
            track = bn.track_running_stats and bn.running_mean is not None
            mom = BN_MOMENTUM if bn.momentum is None else float(bn.momentum)
            call("ctu_bn_finalize", sums.data_ptr(), count, bn.weight.data_ptr(), bn.bias.data_ptr(),
                 bn.running_mean.data_ptr() if track else None, bn.running_var.data_ptr() if track else None,
                 bn.num_batches_tracked.data_ptr() if track else None, mom, float(bn.eps), c, 1, 1, ss.data_ptr(), st)
        else:
            call("ctu_bn_finalize", None, count, bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                 bn.running_var.data_ptr(), None, 0.0, float(bn.eps), c, 0, 0, ss.data_ptr(), st)
        a = self.new_act(c, y.n, y.d, y.h, y.w)
        pooled = self.new_act(c, y.n, y.d // 2, y.h // 2, y.w // 2) if pool else None
        call("ctu_bn_relu_fwd", self.dtype, y.ptr, ss.data_ptr(), a.ptr, pooled.ptr if pool else None,
             c, y.n, y.d, y.h, y.w, st)
        if self.record:
            if not training:
                raise RuntimeError("backward through eval-mode BatchNorm is not supported by the fused path")

            def bwd():
                dA = self.agrads.pop(id(a), None)
                dP = self.agrads.pop(id(pooled), None) if pool else None
                if dA is None and dP is None:
                    raise RuntimeError("internal: BatchNorm stage received no gradient")
                st = stream_ptr()
                if extra_updates > 0 and bn.track_running_stats and bn.running_mean is not None:
                    mom = BN_MOMENTUM if bn.momentum is None else float(bn.momentum)
                    call("ctu_bn_running_update", sums.data_ptr(), count, bn.running_mean.data_ptr(),
                         bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), mom, c, extra_updates, st)
                sums2 = self.f64(2 * cpad)
                pA = dA.ptr if dA is not None else None
                pP = dP.ptr if dP is not None else None
                call("ctu_bn_relu_bwd_reduce", self.dtype, y.ptr, ss.data_ptr(), pA, pP, sums2.data_ptr(),
                     c, y.n, y.d, y.h, y.w, st)
                dy = self.new_act(c, y.n, y.d, y.h, y.w)
                dg, db = self._grad_buffer(bn.weight), self._grad_buffer(bn.bias)
                call("ctu_bn_relu_bwd_apply", self.dtype, y.ptr, ss.data_ptr(), bn.weight.data_ptr(), pA, pP,
                     sums2.data_ptr(), count, dy.ptr, dg.data_ptr(), db.data_ptr(), c, y.n, y.d, y.h, y.w, st)
                self._add_pgrad(bn.weight, dg)
                self._add_pgrad(bn.bias, db)
                self._set_agrad(y, dy)

            self.tape.append(bwd)
        return (a, pooled) if pool else a

    # ------------------------------------------------------------------ head
    def head(self, srcs: Sequence[Act], weight, bias, flags: int):
        """last_conv (1x1x1 + bias) + output non-linearities.  Returns fp32 NCDHW tensor(s)."""
        cout = weight.shape[0]
        s0 = srcs[0]
        pa, ca, ns = self._src_args(srcs)
        sp = bool(flags & (_lib.HEAD_SP | _lib.HEAD_SP_SOFTMAX))
        if sp:
            out0 = self.f32(s0.n, 2, s0.d, s0.h, s0.w)
            out1 = self.f32(s0.n, 2, s0.d, s0.h, s0.w)
        else:
            out0, out1 = self.f32(s0.n, cout, s0.d, s0.h, s0.w), None
        call("ctu_head_fwd", self.dtype, pa, ca, ns, weight.data_ptr(), bias.data_ptr(), cout, flags,
             out0.data_ptr(), out1.data_ptr() if sp else None, s0.n, s0.spatial, stream_ptr())
        if self.record:
            srcs = list(srcs)

            def bwd(g0, g1):
                pa, ca, ns = self._src_args(srcs)
                dsrcs = [self.new_act(s.c, s.n, s.d, s.h, s.w) for s in srcs]
                dw, db = self._grad_buffer(weight), self._grad_buffer(bias)
                call("ctu_head_bwd", self.dtype, pa, ca, ns, weight.data_ptr(), bias.data_ptr(), cout, flags,
                     g0.data_ptr() if g0 is not None else None, g1.data_ptr() if g1 is not None else None,
                     ptr_array([d.ptr for d in dsrcs]), dw.data_ptr(), db.data_ptr(), s0.n, s0.spatial, stream_ptr())
                self._add_pgrad(weight, dw)
                self._add_pgrad(bias, db)
                for s, d in zip(srcs, dsrcs):
                    self._set_agrad(s, d)

            self.head_bwd = bwd
        return (out0, out1) if sp else out0

    # ------------------------------------------------------------------ backward driver
    def backward(self, g0, g1):
        self.head_bwd(g0, g1)
        for fn in reversed(self.tape):
            fn()
        self.tape = []


def tc_supported(k, src_channels, cout, d, h, w) -> bool:
    """Shapes covered by the tcgen05 implicit-GEMM convolution (the predicate lives in conv_tc.cu)."""
    if len(src_channels) != 1:
        return False
    return bool(_lib.load().ctu_conv_tc_supported(k, src_channels[0], cout, d, h, w))
