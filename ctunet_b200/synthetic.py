"""Synthetic skull-CT training data generated on the GPU: an ellipsoidal bone
shell phantom, a seeded virtual craniectomy through the product's own masking kernels
(transforms.py:241-300 semantics), the atlas stand-in channel (datasets.py:30-47) and one-hot float32
targets (datasets.py:209-214).  Used by bench.py and the examples; there are no real datasets here."""
from __future__ import annotations

import random

import numpy as np
import torch

from .utilities import random_blank_patch


def skull_phantom(size: int, seed: int, device) -> torch.Tensor:
    """uint8 [S,S,S] binary skull: ellipsoidal shell, thickness ~4.5 voxels at 128^3."""
    g = torch.Generator().manual_seed(seed)
    jit = (0.03 * torch.rand(3, generator=g)).tolist()
    lin = torch.linspace(-1, 1, size, device=device)
    zz, yy, xx = torch.meshgrid(lin, lin, lin, indexing="ij")
    r = torch.sqrt((zz / (0.80 + jit[0])) ** 2 + (yy / (0.88 + jit[1])) ** 2 + (xx / (0.72 + jit[2])) ** 2)
    thick = max(3.0, 4.5 * size / 128.0) / (size / 2.0)
    return ((r >= 1.0 - thick) & (r <= 1.0)).to(torch.uint8)


def make_training_batch(batch: int, in_channels: int, size: int, seed: int, device="cuda"):
    """(image [B,Cin,S,S,S] float32 {0,1}, (skull_onehot, flap_onehot) each [B,2,S,S,S] float32)."""
    random.seed(seed)
    np.random.seed(seed)
    atlas = skull_phantom(size, 999, device).float()
    imgs, sks, fls = [], [], []
    for b in range(batch):
        full = skull_phantom(size, seed * 131 + b, device)
        broken, flap = random_blank_patch(full, 1, True)
        chans = [broken.float()]
        if in_channels > 1:
            chans.append(atlas)
        imgs.append(torch.stack(chans, 0))
        sks.append(torch.stack((1 - full, full), 0).float())
        fls.append(torch.stack((1 - flap, flap), 0).float())
    return torch.stack(imgs).contiguous(), (torch.stack(sks).contiguous(), torch.stack(fls).contiguous())
