"""GPU versions of the hot-path helpers of ``ctunet.utilities`` / ``ctunet.pytorch.transforms``.

  hard_segm_from_tensor   ctunet/utilities.py:103-124
  shape_3d                ctunet/utilities.py:127-178
  random_blank_patch      ctunet/pytorch/transforms.py:241-300
  SkullRandomHole         ctunet/pytorch/transforms.py:52-94
  encode_flaprec_batch    ctunet/pytorch/datasets.py:195-235, :30-47 (one-hot targets, atlas channel)
  dice_coeff, hausdorff   ctunet/utilities.py:53-70 (reporting metrics; monai restated -> parity unpinned)
  SaltAndPepper           ctunet/pytorch/transforms.py:13-49

The 'flap' shape of shape_3d calls the un-vendored, unpinned ``raster_geometry`` package in the reference
(utilities.py:145-166); it is not installed here, so ``cylinder`` / ``cube`` are RESTATED from the package's published
algorithm (oracle/unet_oracle.py: ``rg_cylinder``, ``rg_cube``) -- sphere and box are pinned to the reference,
the flap shape is parity unpinned.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from ._lib import call, stream_ptr
from ._lib import load as _load
from .losses import dice_loss  # noqa: F401  (re-export: utils.dice_loss)

_SHAPES = {"circle": 0, "sphere": 0, "square": 1, "box": 1, "cube": 1, "flap": 2, "autoimplant": 2}


def _need_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s runs on CUDA tensors only (no CPU fallback)" % what)


def hard_segm_from_tensor(prob_map: torch.Tensor, keep_dims: bool = False) -> torch.Tensor:
    """Channel argmax as float32; ties go to the lowest index (utilities.py:118-124)."""
    _need_cuda(prob_map, "hard_segm_from_tensor")
    x = prob_map.float().contiguous()
    five = x.dim() == 5
    if not five:
        x = x.unsqueeze(0)
    b, c = x.shape[0], x.shape[1]
    spatial = x[0, 0].numel()
    out = torch.empty((b,) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    call("ctu_argmax_channels", x.data_ptr(), out.data_ptr(), b, c, spatial, stream_ptr())
    if five:
        return out.unsqueeze(1) if keep_dims else out
    out = out[0]
    return out.unsqueeze(0) if keep_dims else out


def blank_patch(image: torch.Tensor, center, size, shape: str, c_diam=None):
    """``masked = image AND outside``, ``extracted = image AND inside`` (uint8), the arithmetic of
    transforms.py:286-296 for a given centre / radius / shape.  ``center`` is a host triple or a device int32[3].
    The 'flap' shape draws its cylinder radius ``c_diam`` from the numpy RNG as utilities.py:146-148 unless given."""
    _need_cuda(image, "blank_patch")
    if shape not in _SHAPES:
        raise ValueError("shape %r is not supported (sphere, box, flap)" % (shape,))
    if _SHAPES[shape] == 2 and c_diam is None:
        c_diam = np.random.uniform(0.25, 1) * size / 4
    img = image.to(torch.uint8).contiguous()
    if img.dim() != 3:
        raise ValueError("expected a [D, H, W] volume")
    if not isinstance(center, torch.Tensor):
        center = torch.tensor([int(c) for c in center], dtype=torch.int32, device=img.device)
    masked, extracted = torch.empty_like(img), torch.empty_like(img)
    d, h, w = img.shape
    call("ctu_flap_mask_u8", img.data_ptr(), masked.data_ptr(), extracted.data_ptr(), d, h, w, center.data_ptr(),
         float(size), _SHAPES[shape], float(c_diam or 0.0), stream_ptr())
    return masked, extracted


def shape_3d(center, size, image_size, shape="flap", device="cuda", c_diam=None) -> torch.Tensor:
    """utilities.py:127-178: 0 inside the shape (bounds inclusive), 1 outside; float64 for sphere / box, uint8 for
    the flap shape, as the reference returns them."""
    ones = torch.ones(tuple(image_size), dtype=torch.uint8, device=device)
    masked, _ = blank_patch(ones, center, size, shape, c_diam)
    return masked if _SHAPES[shape] == 2 else masked.to(torch.float64)


def count_nonzero(image: torch.Tensor) -> int:
    img = image.to(torch.uint8).contiguous()
    cnt = torch.empty(1, dtype=torch.int64, device=img.device)
    call("ctu_count_nonzero_u8", img.data_ptr(), img.numel(), cnt.data_ptr(), stream_ptr())
    return int(cnt.item())


def kth_nonzero(image: torch.Tensor, k: int) -> torch.Tensor:
    """Device int32[3] = ``np.argwhere(image > 0)[k]`` (C order), transforms.py:249-252."""
    img = image.to(torch.uint8).contiguous()
    d, h, w = img.shape
    scratch = torch.empty((img.numel() + 4095) // 4096 + 1, dtype=torch.int64, device=img.device)
    coords = torch.empty(3, dtype=torch.int32, device=img.device)
    call("ctu_kth_nonzero_u8", img.data_ptr(), d, h, w, int(k), scratch.data_ptr(), coords.data_ptr(), stream_ptr())
    return coords


def radius_bounds(image_size):
    """transforms.py:266-267"""
    min_radius = (int(np.min(image_size)) // 5) - 1
    max_radius = np.max([min_radius, np.max(image_size) // 3.5])
    return min_radius, max_radius


def random_blank_patch(image: torch.Tensor, prob=1, return_extracted=False, p_type="random",
                       valid_shapes=("sphere", "box", "flap")):
    """transforms.py:241-300 on a CUDA volume.  Consumes the host RNGs in the reference's order
    (random.uniform, np.random.choice, np.random.randint, np.random.randint, and np.random.uniform for the flap
    shape) so a seeded run picks the same voxel index, radius, shape index and cylinder radius as the reference would
    for the same nonzero count."""
    _need_cuda(image, "random_blank_patch")
    img = image.to(torch.uint8).contiguous()
    r = random.uniform(0, 1)
    if prob >= r:
        n = count_nonzero(img)
        if n:
            center = kth_nonzero(img, int(np.random.choice(n)))
            min_radius, max_radius = radius_bounds(tuple(img.shape))
            size = np.random.randint(min_radius, max_radius)
            if p_type not in valid_shapes:
                p_type = valid_shapes[np.random.randint(0, len(valid_shapes))]
            masked, extracted = blank_patch(img, center, size, p_type)
            return (masked, extracted) if return_extracted else masked
    return (img, torch.zeros_like(img)) if return_extracted else img


class SkullRandomHole(object):
    """transforms.py:52-94 for CUDA tensors: ``sample['image']`` is [D,H,W] or [B,D,H,W]."""

    def __init__(self, p=1, double_output=False):
        self.p = p
        self.double_output = double_output

    def __call__(self, sample):
        img = sample["image"]
        _need_cuda(img, "SkullRandomHole")
        is_batch = img.dim() == 4
        vols = img if is_batch else img.unsqueeze(0)
        full = vols.to(torch.uint8)
        brk, flap = [], []
        for i in range(vols.shape[0]):
            m, e = random_blank_patch(full[i], self.p, True)
            brk.append(m)
            flap.append(e)
        brk, flap = torch.stack(brk), torch.stack(flap)
        if not is_batch:
            brk, flap = brk[0], flap[0]
        if self.double_output:
            return {"image": brk, "target": (full, flap)}
        return {"image": brk, "target": flap}


def encode_flaprec_batch(broken: torch.Tensor, full: torch.Tensor, flap: torch.Tensor, atlas=None, out=None):
    """What ``FlapRecWShapePrior2OTrainDataset.__getitem__`` + the default collate hand the model
    (datasets.py:195-235), computed on the device from uint8 masks [B,D,H,W]:
    image [B,Cin,D,H,W] float32 (channel 0 = the broken skull, channel 1 = ``atlas`` [D,H,W] float32 as
    ``load_atlas_and_append_at_axis`` appends it, datasets.py:30-47) and the two targets
    ``one_hot(label, 2).movedim(-1, 1).float()`` [B,2,D,H,W] (datasets.py:209-214).  The host then ships 3 bytes per
    voxel instead of 24.  ``out = (image, (skull_target, flap_target))`` writes into existing tensors;
    ``out = (image, None)`` produces the image only (the fused head + loss kernels of the training step read the uint8
    label masks themselves)."""
    for t in (broken, full, flap):
        _need_cuda(t, "encode_flaprec_batch")
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape != broken.shape:
            raise TypeError("encode_flaprec_batch expects three uint8 [B, D, H, W] masks of one shape")
    b = broken.shape[0]
    vol = tuple(broken.shape[1:])
    spatial = broken[0].numel()
    cin = 1 if atlas is None else 2
    if atlas is not None and (atlas.dtype != torch.float32 or tuple(atlas.shape) != vol or not atlas.is_cuda):
        raise TypeError("atlas: float32 CUDA volume of the mask shape")
    if out is None:
        image = torch.empty((b, cin) + vol, dtype=torch.float32, device=broken.device)
        sk = torch.empty((b, 2) + vol, dtype=torch.float32, device=broken.device)
        fl = torch.empty_like(sk)
    elif out[1] is None:
        image, sk, fl = out[0], None, None
        if image.dtype != torch.float32 or tuple(image.shape) != (b, cin) + vol or not image.is_contiguous() or not image.is_cuda:
            raise TypeError("encode_flaprec_batch: out image must be contiguous float32 CUDA [B, C, D, H, W]")
    else:
        image, (sk, fl) = out
        for t, c in ((image, cin), (sk, 2), (fl, 2)):
            if t.dtype != torch.float32 or tuple(t.shape) != (b, c) + vol or not t.is_contiguous() or not t.is_cuda:
                raise TypeError("encode_flaprec_batch: out tensors must be contiguous float32 CUDA [B, C, D, H, W]")
    call("ctu_encode_flaprec_u8", broken.contiguous().data_ptr(), full.contiguous().data_ptr(),
         flap.contiguous().data_ptr(), atlas.contiguous().data_ptr() if atlas is not None else None, image.data_ptr(),
         sk.data_ptr() if sk is not None else None, fl.data_ptr() if fl is not None else None, b, cin, spatial, stream_ptr())
    return image, ((sk, fl) if sk is not None else None)


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """uint8 {0,1} mask [..., S] -> bit-packed uint8 [..., S / 8], voxel v = bit ``v & 7`` of byte ``v >> 3`` (numpy
    ``packbits(bitorder="little")``): the wire format of ``encode_flaprec_bits`` / ``TrainStep.step_from_bits``.  Works on
    host or device tensors (a data pipeline would emit this format directly)."""
    flat = (mask.reshape(mask.shape[0], -1) != 0).to(torch.uint8)
    if flat.shape[1] % 8:
        raise ValueError("the volume size must be a multiple of 8 voxels")
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=mask.device)
    return (flat.view(flat.shape[0], -1, 8) * w).sum(-1, dtype=torch.int32).to(torch.uint8).contiguous()


def encode_flaprec_bits(broken_bits, full_bits, flap_bits, vol_shape, atlas=None, out=None):
    """``encode_flaprec_batch`` from bit-packed masks ([B, D*H*W/8] uint8 each, ``pack_mask_bits``): returns / fills
    ``(image float32 [B,Cin,D,H,W], (full_mask, flap_mask) uint8 [B,D,H,W])`` -- the label masks are what the fused head +
    loss kernels of ``trainer.TrainStep`` take as targets."""
    for t in (broken_bits, full_bits, flap_bits):
        _need_cuda(t, "encode_flaprec_bits")
        if t.dtype != torch.uint8 or t.dim() != 2 or t.shape != broken_bits.shape or not t.is_contiguous():
            raise TypeError("encode_flaprec_bits expects three contiguous uint8 [B, D*H*W/8] bit volumes of one shape")
    b = broken_bits.shape[0]
    vol = tuple(vol_shape)
    spatial = vol[0] * vol[1] * vol[2]
    if broken_bits.shape[1] * 8 != spatial:
        raise ValueError("bit volumes of %d bytes do not match the volume shape %s" % (broken_bits.shape[1], vol))
    cin = 1 if atlas is None else 2
    if atlas is not None and (atlas.dtype != torch.float32 or tuple(atlas.shape) != vol or not atlas.is_cuda):
        raise TypeError("atlas: float32 CUDA volume of the mask shape")
    if out is None:
        image = torch.empty((b, cin) + vol, dtype=torch.float32, device=broken_bits.device)
        fm = torch.empty((b,) + vol, dtype=torch.uint8, device=broken_bits.device)
        lm = torch.empty_like(fm)
    else:
        image, (fm, lm) = out
        for t, shp, dt in ((image, (b, cin) + vol, torch.float32), (fm, (b,) + vol, torch.uint8), (lm, (b,) + vol, torch.uint8)):
            if t.dtype != dt or tuple(t.shape) != shp or not t.is_contiguous() or not t.is_cuda:
                raise TypeError("encode_flaprec_bits: out = (float32 image [B,C,D,H,W], (uint8 [B,D,H,W], uint8 [B,D,H,W]))")
    call("ctu_encode_flaprec_bits", broken_bits.data_ptr(), full_bits.data_ptr(), flap_bits.data_ptr(),
         atlas.contiguous().data_ptr() if atlas is not None else None, image.data_ptr(), fm.data_ptr(), lm.data_ptr(), b, cin,
         spatial, stream_ptr())
    return image, (fm, lm)


# ---------------------------------------------------------------------------------------------- reporting metrics
def _metric_pair(pred, target, what):
    _need_cuda(pred, what)
    _need_cuda(target, what)
    if pred.dim() != 5 or pred.shape != target.shape or not 2 <= pred.shape[1] <= 4:
        raise ValueError("%s expects prediction and one-hot target of one [B, C, D, H, W] shape, 2 <= C <= 4" % what)
    return pred.detach().float().contiguous(), target.detach().float().contiguous()


def dice_coeff(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """``utils.dice_coeff`` (utilities.py:53-60): mean over (sample, foreground class) of the hard-label Dice coefficient
    ``2|P n T| / (|P| + |T|)`` with P = ``argmax(pred, 1) == class`` -- ``monai.metrics.compute_meandice(one_hot(argmax),
    target, include_background=False)`` restated (NaN where the target class is empty, which then propagates through the
    mean exactly as in the reference).  Returns a 0-dim float32 CUDA tensor; no host synchronisation.
    One divergence, documented: when NO voxel of the whole batch is predicted foreground the reference's
    ``one_hot(argmax)`` has a single channel and monai raises on the shape mismatch; here the class counts as empty."""
    p, t = _metric_pair(pred, target, "dice_coeff")
    b, c = p.shape[0], p.shape[1]
    counts = torch.empty(3 * b * (c - 1), dtype=torch.float64, device=p.device)
    out = torch.empty(1, dtype=torch.float32, device=p.device)
    call("ctu_dice_coeff", p.data_ptr(), t.data_ptr(), b, c, p[0, 0].numel(), counts.data_ptr(), out.data_ptr(), stream_ptr())
    return out[0]


def hausdorff(result_b: torch.Tensor, reference_b: torch.Tensor) -> torch.Tensor:
    """``utils.hausdorff`` (utilities.py:63-70): mean over (sample, foreground class) of the symmetric Hausdorff distance
    between the surfaces of ``argmax(result_b, 1) == class`` and ``reference_b[:, class] == 1``
    (``monai.metrics.compute_hausdorff_distance`` restated: surface = mask minus its 6-neighbourhood erosion, Euclidean
    distance in voxels); NaN / inf (an empty surface) become ``max(reference_b.shape)`` as utilities.py:64,69 do.
    Exact integer squared distance transform on the device; returns a 0-dim float64 CUDA tensor, no host sync."""
    p, t = _metric_pair(result_b, reference_b, "hausdorff")
    b, c, d, h, w = p.shape
    lib = _load()
    nbytes = lib.ctu_hausdorff_workspace_bytes(b, c, d, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
    out = torch.empty(1, dtype=torch.float64, device=p.device)
    call("ctu_hausdorff", p.data_ptr(), t.data_ptr(), b, c, d, h, w, float(max(reference_b.shape)), ws.data_ptr(), nbytes,
         out.data_ptr(), stream_ptr())
    return out[0]


# ---------------------------------------------------------------------------------------------- SaltAndPepper
class SaltAndPepper(object):
    """transforms.py:13-49 for CUDA tensors ([D,H,W] or [B,D,H,W], any dtype; returns float32 like ``torch.FloatTensor``).

    Kept from the reference, in its order: ``self.noise_density`` is OVERWRITTEN by ``np.random.uniform(0, noise_density)``
    on every call and key (transforms.py:31 -- the noise decays towards zero over training), then per
    image one ``random.uniform(0, 1)`` gate against ``p``.  The two per-voxel uniform fields come from a Philox
    counter-based generator on the device (seeded from the numpy RNG, one draw per noisy image) instead of 2 x D*H*W
    float64 draws from numpy's MT19937 stream on the host: same distribution, different stream.  ``fields=(u_black, u_white)``
    (float64 CUDA tensors of the batch shape) replaces the generator -- with the reference's own draws the output is
    bit-identical to the reference's."""

    def __init__(self, p=1, noise_density=0.2, salt_ratio=0.1, keyws=("image", "target"), apply_to=(True, False)):
        self.p = p
        self.noise_density = noise_density
        self.salt_ratio = salt_ratio
        self.keyws = keyws
        self.apply_to = apply_to

    def __call__(self, sample, fields=None):
        for i, keyw in enumerate(self.keyws):
            if not self.apply_to[i]:
                continue
            img = sample[keyw]
            _need_cuda(img, "SaltAndPepper")
            is_batch = img.dim() == 4
            vols = (img if is_batch else img.unsqueeze(0)).to(torch.uint8).contiguous()   # .astype(np.uint8), :29-30
            out = vols.clone()
            self.noise_density = np.random.uniform(0, self.noise_density)                # :31
            for j in range(vols.shape[0]):
                r = random.uniform(0, 1)                                                  # :33
                if self.p >= r:
                    ub = uw = None
                    seed = 0
                    if fields is not None:
                        ub = (fields[0] if is_batch else fields[0].unsqueeze(0))[j].contiguous()
                        uw = (fields[1] if is_batch else fields[1].unsqueeze(0))[j].contiguous()
                        if ub.dtype != torch.float64 or uw.dtype != torch.float64 or ub.shape != vols[j].shape:
                            raise TypeError("fields: two float64 CUDA tensors of the image shape")
                    else:
                        seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
                    call("ctu_salt_pepper_u8", vols[j].data_ptr(), out[j].data_ptr(), vols[j].numel(),
                         float(self.noise_density), float(self.salt_ratio), ub.data_ptr() if ub is not None else None,
                         uw.data_ptr() if uw is not None else None, seed, 0, stream_ptr())
            out = out.float()
            sample[keyw] = out if is_batch else out[0]
        return sample
