"""Every model preset of ctunet.pytorch.models through a few captured training steps at full size (GPU): finite, decreasing
loss and ms/step.  python scripts/preset_sweep.py [size] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ctunet_b200 as C
from ctunet_b200.synthetic import make_training_batch
from ctunet_b200.trainer import TrainStep

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
PRESETS = [("UNetSP", "double", 2), ("UNetDO", "double", 1), ("UNetSPSmall", "double", 2), ("UNet4b2i3o", "single3", 2),
           ("UNet5b2i3o", "single3", 2), ("UNet4b1i3o", "single3", 1), ("UNet", "single", 1), ("UNet4_2IC", "single", 2),
           ("recAE_v2_fixed", "single", 1)]
for name, handler, cin in PRESETS:
    torch.manual_seed(0)
    net = getattr(C, name)().to(dev)
    img, (sk_t, fl_t) = make_training_batch(batch, cin, size, seed=3, device=dev)
    if handler == "single3":          # 3-channel sigmoid output against a 3-class one-hot target (background, flap, rest)
        tgt = torch.stack((1 - sk_t[:, 1], fl_t[:, 1], sk_t[:, 1] - fl_t[:, 1]), 1).contiguous()
        step = TrainStep(net, "single", 1.0, 1.0, lr=1e-3, graph=True)
    else:
        tgt = (sk_t, fl_t) if handler == "double" else sk_t
        step = TrainStep(net, handler, 1.0, 1.0, lr=1e-3, graph=True)
    losses = []
    for it in range(8):
        losses.append(step(img, tgt).tolist()[-1])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for it in range(5):
        step(img, tgt)
    e.record()
    torch.cuda.synchronize()
    ok = all(l == l and abs(l) < 1e6 for l in losses) and losses[-1] < losses[0]
    print("%-16s %s  loss %.4f -> %.4f  %.2f ms/step (batch %d, %d^3)" % (name, "ok  " if ok else "FAIL", losses[0], losses[-1],
                                                                       s.elapsed_time(e) / 5, batch, size))
    del step, net
    torch.cuda.empty_cache()
