"""GPU CT preprocessing and sliding-window inference (BASELINE.json configs 4 and 5).

The reference has NO implementation of HU windowing, thresholding, resampling or patch-wise
inference (they live in the author's separate ``headctools`` project); the
semantics implemented here are the ones frozen in ``oracle/unet_oracle.py`` (clamp + rescale,
``>=`` threshold, PyTorch 'nearest' / 'trilinear, align_corners=False' resampling, non-overlapping
patch grid) -- parity with the reference is therefore unpinned by construction.
"""
from __future__ import annotations

import torch

from ._lib import call, stream_ptr
from .utilities import hard_segm_from_tensor


def _vol(t, dtype, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s runs on CUDA tensors only (no CPU fallback)" % what)
    if t.dim() != 3:
        raise ValueError("%s expects a [D, H, W] volume" % what)
    if t.dtype != dtype:
        raise TypeError("%s expects %s, got %s" % (what, dtype, t.dtype))
    return t.contiguous()


def hu_window(hu: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    """int16 HU -> float32 in [0, 1]: (clamp(hu, lo, hi) - lo) / (hi - lo)."""
    hu = _vol(hu, torch.int16, "hu_window")
    out = torch.empty(hu.shape, dtype=torch.float32, device=hu.device)
    call("ctu_hu_window", hu.data_ptr(), out.data_ptr(), hu.numel(), float(lo), float(hi), stream_ptr())
    return out


def hu_threshold(hu: torch.Tensor, thr: int) -> torch.Tensor:
    """int16 HU -> uint8 bone mask (hu >= thr)."""
    hu = _vol(hu, torch.int16, "hu_threshold")
    out = torch.empty(hu.shape, dtype=torch.uint8, device=hu.device)
    call("ctu_hu_threshold", hu.data_ptr(), out.data_ptr(), hu.numel(), int(thr), stream_ptr())
    return out


def resample_nearest(vol: torch.Tensor, out_size) -> torch.Tensor:
    """PyTorch 'nearest' semantics: src = min(floor(dst * in/out), in - 1) per axis (float32 scale)."""
    if vol.dtype not in (torch.float32, torch.uint8):
        raise TypeError("resample_nearest expects float32 or uint8")
    vol = _vol(vol, vol.dtype, "resample_nearest")
    out = torch.empty(tuple(out_size), dtype=vol.dtype, device=vol.device)
    fn = "ctu_resample_nearest_f32" if vol.dtype == torch.float32 else "ctu_resample_nearest_u8"
    call(fn, vol.data_ptr(), out.data_ptr(), *vol.shape, *out.shape, stream_ptr())
    return out


def nearest_source_index(out_size: int, in_size: int, device="cuda") -> torch.Tensor:
    idx = torch.empty(out_size, dtype=torch.int32, device=device)
    call("ctu_resample_nearest_index", idx.data_ptr(), out_size, in_size, stream_ptr())
    return idx


def resample_trilinear(vol: torch.Tensor, out_size) -> torch.Tensor:
    """F.interpolate(mode='trilinear', align_corners=False) semantics, float32."""
    vol = _vol(vol, torch.float32, "resample_trilinear")
    out = torch.empty(tuple(out_size), dtype=torch.float32, device=vol.device)
    call("ctu_resample_trilinear_f32", vol.data_ptr(), out.data_ptr(), *vol.shape, *out.shape, stream_ptr())
    return out


def _captured_patch_labels(model, xb: torch.Tensor):
    """Eval-mode forward + argmax of one patch batch, replayed from a CUDA graph captured per (model, batch shape): an
    inference pass is ~110 launches with no host synchronisation, and at small patch batches the Python / ctypes dispatch
    (about 2 ms) is longer than the GPU work."""
    graphs = model.__dict__.setdefault("_infer_graphs", {})       # lives and dies with the model
    key = (tuple(xb.shape), str(xb.device), getattr(model, "compute_dtype", None))
    ent = graphs.get(key)
    weights = list(model.parameters()) + list(model.buffers())
    if ent is not None and (len(ent[3]) != len(weights) or any(a is not b for a, b in zip(ent[3], weights))):
        ent = None            # a parameter / buffer OBJECT was replaced: the capture reads the old storage
    if ent is None:
        static_in = torch.empty_like(xb)
        static_in.copy_(xb)

        def run():
            o = model(static_in)
            o = o if isinstance(o, tuple) else (o,)
            return [hard_segm_from_tensor(ok) for ok in o]

        run()                                   # warm-up: weight-preparation plan, allocator pools
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = run()
        ent = (g, static_in, outs, weights)
        graphs[key] = ent
        if len(graphs) > 8:                     # a handful of shapes at most: drop the oldest capture
            graphs.pop(next(iter(graphs)))
    g, static_in, outs, _ = ent
    static_in.copy_(xb)
    g.replay()
    return outs


@torch.no_grad()
def sliding_window_argmax(model, vol: torch.Tensor, patch: int = 128, batch: int = 4, graph: bool = True):
    """Eval-mode model over the non-overlapping ``patch``^3 grid of ``vol`` [Cin, D, H, W]; returns the
    stitched float32 label volume(s) (one per model output), as ``hard_segm_from_tensor`` would label
    each patch (utilities.py:103-124).  ``graph``: replay every full patch batch from a captured CUDA graph."""
    if vol.dim() != 4:
        raise ValueError("expected [Cin, D, H, W]")
    c, d, h, w = vol.shape
    if d % patch or h % patch or w % patch:
        raise ValueError("volume %s must be a multiple of the patch size %d" % ((d, h, w), patch))
    was_training = model.training
    model.eval()
    origins = [(z, y, x) for z in range(0, d, patch) for y in range(0, h, patch) for x in range(0, w, patch)]
    outs = None
    for i in range(0, len(origins), batch):
        chunk = origins[i:i + batch]
        xb = torch.stack([vol[:, z:z + patch, y:y + patch, x:x + patch] for z, y, x in chunk]).contiguous()
        if graph and vol.is_cuda:
            labs = _captured_patch_labels(model, xb)
        else:
            o = model(xb)
            o = o if isinstance(o, tuple) else (o,)
            labs = [hard_segm_from_tensor(ok) for ok in o]
        if outs is None:
            outs = [torch.empty((d, h, w), dtype=torch.float32, device=vol.device) for _ in labs]
        for k, lab in enumerate(labs):
            for j, (z, y, x) in enumerate(chunk):
                outs[k][z:z + patch, y:y + patch, x:x + patch] = lab[j]
    model.train(was_training)
    return outs
