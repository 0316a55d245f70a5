"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the product's fused up-sampling stage.

The reference up block starts with ``ConvTranspose3d(in_c, in_c, 2, 2)`` followed by ``Conv3d(in_c, out_c, k, 1, k//2)``
(ctunet/pytorch/models.py:37-38; legacy models.py:427-430).  Both are linear, so their composition is ONE 3x3x3
convolution on the LOW-resolution grid that produces the 8 output phases (q_d, q_h, q_w) of every low-res voxel:

    y[2u + q][co] = sum_{delta in {-1,0,1}^3} sum_ci Wc[(q, co)][ci][delta] * x[u + delta][ci]
    Wc[(q, co)][ci][delta] = sum_{k : floor((q + k - pad) / 2) = delta} sum_cm W3[co][cm][k] * WT[ci][cm][(q + k - pad) mod 2]

per dimension, with the transposed convolution's bias carried by an extra all-ones input channel (zero outside the
volume like every other channel, which reproduces the zero padding of the 8x larger intermediate exactly).  The product
never materialises that intermediate (ctunet_b200/csrc/fuse.cu builds Wc on the GPU; engine.py runs the convolution).
This module states the same algebra in PyTorch so that (a) it can be checked against conv_transpose3d -> conv3d on
the CPU and (b) the CUDA composition / decomposition kernels can be checked against it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def phase_taps(k: int):
    """[(q, kk, delta, p)] for one dimension: output phase q, high-res tap kk, low-res offset delta, convT phase p."""
    pad = k // 2
    out = []
    for q in range(2):
        for kk in range(k):
            o = q + kk - pad
            out.append((q, kk, o // 2, o % 2))          # python floor division / non-negative modulo
    return out


def compose(wt: torch.Tensor, bt, w3: torch.Tensor, k: int) -> torch.Tensor:
    """wt [cin][cm][2][2][2] (ConvTranspose3d), bt [cm] or None, w3 [cout][cm][k][k][k]  ->
    Wn [8 * cop][cin + 1][3][3][3] with cop = 8 * ceil(cout / 8); output channel (q * cop + co), q = qd*4 + qh*2 + qw;
    input channel `cin` is the all-ones channel carrying the transposed convolution's bias."""
    cin, cm = wt.shape[0], wt.shape[1]
    cout = w3.shape[0]
    cop = (cout + 7) // 8 * 8
    wta = wt if bt is None else torch.cat([wt, bt.view(1, cm, 1, 1, 1).expand(1, cm, 2, 2, 2)], 0)
    if bt is None:
        wta = torch.cat([wt, torch.zeros(1, cm, 2, 2, 2, dtype=wt.dtype)], 0)
    wn = torch.zeros(8, cop, cin + 1, 3, 3, 3, dtype=w3.dtype)
    taps = phase_taps(k)
    for qd, kd, dd, pd in taps:
        for qh, kh, dh, ph in taps:
            for qw, kw, dw, pw in taps:
                q = qd * 4 + qh * 2 + qw
                # [cout][cm] x [cin+1][cm] -> [cout][cin+1]
                wn[q, :cout, :, dd + 1, dh + 1, dw + 1] += w3[:, :, kd, kh, kw] @ wta[:, :, pd, ph, pw].t()
    return wn.view(8 * cop, cin + 1, 3, 3, 3)


def depth_to_space(y: torch.Tensor, cout: int) -> torch.Tensor:
    """[B][8*cop][d][h][w] phase-major -> [B][cout][2d][2h][2w]."""
    b, c8, d, h, w = y.shape
    cop = c8 // 8
    y = y.view(b, 2, 2, 2, cop, d, h, w)[:, :, :, :, :cout]
    return y.permute(0, 4, 5, 1, 6, 2, 7, 3).reshape(b, cout, 2 * d, 2 * h, 2 * w)


def space_to_depth(g: torch.Tensor, cop: int) -> torch.Tensor:
    """[B][cout][2d][2h][2w] -> [B][8*cop][d][h][w] phase-major (pad channels zero)."""
    b, c, D, H, W = g.shape
    g = g.view(b, c, D // 2, 2, H // 2, 2, W // 2, 2).permute(0, 3, 5, 7, 1, 2, 4, 6)
    out = torch.zeros(b, 2, 2, 2, cop, D // 2, H // 2, W // 2, dtype=g.dtype)
    out[:, :, :, :, :c] = g
    return out.reshape(b, 8 * cop, D // 2, H // 2, W // 2)


def fused_up_conv(x: torch.Tensor, wt, bt, w3, b3, k: int) -> torch.Tensor:
    """conv3d(conv_transpose3d(x, wt, bt, stride=2), w3, b3, padding=k//2) through the composed low-res convolution."""
    cout = w3.shape[0]
    wn = compose(wt, bt, w3, k)
    ones = torch.ones(x.shape[0], 1, *x.shape[2:], dtype=x.dtype)
    y = F.conv3d(torch.cat([x, ones], 1), wn, None, padding=1)
    y = depth_to_space(y, cout)
    if b3 is not None:
        y = y + b3.view(1, -1, 1, 1, 1)
    return y
