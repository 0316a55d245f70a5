"""BASELINE.json configs 4 and 5 on one GPU (not bench.py lines; reported in DESIGN.md):

  config 5: preprocessing of a synthetic 512 x 512 x 256 int16 head CT -- HU window, bone threshold, trilinear / nearest
            resample to the training grid, virtual-craniectomy masking -- with every kernel's achieved GB/s
  config 4: the same volume resampled to 512 x 512 x 256 -> 4 x 4 x 2 = 32 non-overlapping 128^3 patches through the
            eval-mode UNetSP + argmax (sliding_window_argmax), voxels/s

    python scripts/bench_infer.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ctunet_b200 as C
from ctunet_b200 import preprocess as P

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e-3


D, H, W = 256, 512, 512
g = torch.Generator(device=dev).manual_seed(0)
lin = [torch.linspace(-1, 1, n, device=dev) for n in (D, H, W)]
zz, yy, xx = torch.meshgrid(*lin, indexing="ij")
r = torch.sqrt((zz / 0.8) ** 2 + (yy / 0.88) ** 2 + (xx / 0.72) ** 2)
hu = torch.full((D, H, W), -1000, dtype=torch.int16, device=dev)
hu[r <= 1.0] = 40
hu[(r >= 0.93) & (r <= 1.0)] = 1000
hu = (hu.float() + 30 * torch.randn(hu.shape, device=dev, generator=g)).clamp(-1024, 3071).to(torch.int16)
del zz, yy, xx, r
vox = D * H * W
print("volume %dx%dx%d int16 (%.0f MB)" % (D, H, W, vox * 2 / 1e6))

rows = []
t = timeit(lambda: P.hu_window(hu, -100.0, 1500.0))
rows.append(("hu_window int16 -> f32", t, vox * (2 + 4)))
t = timeit(lambda: P.hu_threshold(hu, 300))
rows.append(("hu_threshold int16 -> u8", t, vox * (2 + 1)))
win = P.hu_window(hu, -100.0, 1500.0)
bone = P.hu_threshold(hu, 300)
out = (128, 256, 256)
ovox = out[0] * out[1] * out[2]
t = timeit(lambda: P.resample_trilinear(win, out))
rows.append(("resample_trilinear f32 512x512x256 -> 256x256x128", t, vox * 4 + ovox * 4))
t = timeit(lambda: P.resample_nearest(bone, out))
rows.append(("resample_nearest u8 512x512x256 -> 256x256x128", t, ovox * 1 + ovox * 1))
small = P.resample_nearest(bone, out)
center = C.kth_nonzero(bone, 12345)
t = timeit(lambda: C.blank_patch(bone, center, 80, "sphere"))
rows.append(("flap mask (sphere) u8 512x512x256", t, vox * 3))
t = timeit(lambda: C.blank_patch(bone, center, 80, "flap", 12.0))
rows.append(("flap mask (flap shape) u8 512x512x256", t, vox * 3))
t = timeit(lambda: C.kth_nonzero(bone, 12345))
rows.append(("k-th non-zero voxel (count + select)", t, vox * 1))


def pipeline():
    b = P.hu_threshold(hu, 300)
    w = P.hu_window(hu, -100.0, 1500.0)
    P.resample_trilinear(w, out)
    s = P.resample_nearest(b, out)
    c = C.kth_nonzero(s, 1000)
    C.blank_patch(s, c, 40, "sphere")


t_pipe = timeit(pipeline)
for name, t, by in rows:
    print("  %-52s %8.1f us  %7.0f GB/s" % (name, t * 1e6, by / t / 1e9))
print("config 5 pipeline (threshold + window + trilinear + nearest + pick + mask): %.2f ms per volume = %.2f Gvox/s (input voxels)"
      % (t_pipe * 1e3, vox / t_pipe / 1e9))

# ---- config 4: sliding-window inference over the full-resolution volume
torch.manual_seed(0)
net = C.UNetSP().to(dev).eval()
atlas = (torch.rand(D, H, W, device=dev, generator=g) > 0.8).float()
volume = torch.stack((bone.float(), atlas))            # [2, D, H, W]
for batch in (1, 4, 8):
    t = timeit(lambda: P.sliding_window_argmax(net, volume, patch=128, batch=batch), reps=3, warm=1)
    print("config 4 sliding-window UNetSP eval + argmax, 32 patches of 128^3, batch %d: %.1f ms per volume = %.2f Gvox/s"
          % (batch, t * 1e3, vox / t / 1e9))
