"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the 'flap' shape of the virtual craniectomy.

Runs the UNMODIFIED reference code (ctunet/utilities.py:145-166 ``shape_3d(shape="flap")`` and
ctunet/pytorch/transforms.py:241-300 ``random_blank_patch(p_type="flap")``) from /root/reference with the two
functions it imports from the un-vendored, unpinned ``raster_geometry`` package (utilities.py:18) bound to the
restatements in oracle/unet_oracle.py (``rg_cylinder``, ``rg_cube``).  Everything the reference itself does --
the relative positions, the two cylinder edges, the union, the RNG order -- is therefore pinned; the voxelisation
inside raster_geometry is not (PARITY UNPINNED for that part).

    python oracle/make_golden_flap.py      (needs /root/reference; writes tests/golden/flap_shape_golden.pt)
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O                      # noqa: E402
from oracle.reference_loader import load_reference       # noqa: E402


def main():
    _, _, UT, TR = load_reference()
    UT.cylinder = lambda shape, height, radius, axis, position: O.rg_cylinder(shape, height, radius, axis, position)
    UT.cube = lambda shape, side, position: O.rg_cube(shape, side, position)
    gold = {"cases": []}
    for center, size, image_size, seed in [((8, 8, 8), 6, (16, 16, 16), 0), ((5, 20, 9), 9, (12, 32, 24), 1),
                                           ((30, 40, 33), 24, (64, 64, 64), 2), ((0, 3, 60), 11, (24, 40, 64), 3)]:
        np.random.seed(seed)
        shp = UT.shape_3d(np.asarray(center), size, image_size, shape="flap")
        np.random.seed(seed)
        c_diam = np.random.uniform(0.25, 1) * size / 4
        gold["cases"].append({"center": list(center), "size": size, "image_size": list(image_size), "seed": seed,
                              "c_diam": float(c_diam), "dtype": str(shp.dtype), "zeros": int((shp == 0).sum()),
                              "packed": torch.from_numpy(np.packbits(shp.astype(np.uint8)))})
        print(center, size, image_size, "zeros", gold["cases"][-1]["zeros"], shp.dtype)
    rng = np.random.RandomState(7)
    img = (rng.rand(24, 32, 40) > 0.6).astype(np.uint8)
    random.seed(4); np.random.seed(4)
    masked, extracted = TR.random_blank_patch(img.copy(), 1, True, p_type="flap")
    gold["random_blank_patch"] = {"img": torch.from_numpy(img), "seed": 4, "masked": torch.from_numpy(masked),
                                  "extracted": torch.from_numpy(extracted)}
    random.seed(9); np.random.seed(9)
    masked, extracted = TR.random_blank_patch(img.copy(), 1, True)       # p_type="random": the shape index is drawn too
    gold["random_blank_patch_any"] = {"seed": 9, "masked": torch.from_numpy(masked), "extracted": torch.from_numpy(extracted)}
    out = os.path.join(ROOT, "tests", "golden", "flap_shape_golden.pt")
    torch.save(gold, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
