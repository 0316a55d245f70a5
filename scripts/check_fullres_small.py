"""examples/UNetSPDO/FlapRecSP2O_512.ini: UNetSPSmall (5 levels, i_size 4) in eval mode on a full-resolution 512 x 512 x 224 volume
(GPU): runs, is finite, and agrees with the fp32 accumulate-check mode on the labels.  python scripts/check_fullres_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ctunet_b200 as C

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.rand(1, 2, 224, 512, 512, device=dev, generator=g) > 0.8).float()
labels, outs = {}, {}
for mode in ("bf16", "fp32"):
    C.set_compute_dtype(mode)
    torch.manual_seed(0)
    net = C.UNetSPSmall().to(dev).eval()
    with torch.no_grad():
        for _ in range(2):
            sk, fl = net(x)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        sk, fl = net(x)
        lab = C.hard_segm_from_tensor(fl)
        e.record()
        torch.cuda.synchronize()
    assert torch.isfinite(sk).all() and torch.isfinite(fl).all()
    labels[mode] = torch.stack((C.hard_segm_from_tensor(sk), lab))
    outs[mode] = (sk.clone(), fl.clone())
    print("%s: UNetSPSmall eval 1x2x224x512x512 -> %s in %.1f ms (%.2f Gvox/s), flap voxels %d, peak mem %.1f GB"
          % (mode, tuple(fl.shape), s.elapsed_time(e), 224 * 512 * 512 / s.elapsed_time(e) / 1e6, int(lab.sum()),
             torch.cuda.max_memory_allocated() / 1e9))
    del net
C.set_compute_dtype("bf16")
agree = (labels["bf16"] == labels["fp32"]).float().mean().item()
err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs["bf16"], outs["fp32"]))
print("bf16 vs fp32 check mode: label agreement %.6f (skull + flap), max relative output error %.2e" % (agree, err))
assert agree > 0.99 and err < 2e-2
