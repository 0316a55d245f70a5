"""ms per captured training step, nothing else (A/B runs of scheduling / kernel knobs):
    python scripts/quick_step.py [--model UNetSP] [--steps 60]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="UNetSP")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import bench as B
    import ctunet_b200 as C
    from ctunet_b200 import _lib
    from ctunet_b200.trainer import TrainStep

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.load()
    C.set_compute_dtype("bf16")
    torch.manual_seed(0)
    net = getattr(C, a.model)().to(dev)
    step = TrainStep(net, B.HANDLER[a.model], 1.0, 1.0, lr=1e-4, scheduler=True, graph=True)
    cin = B.in_channels(a.model)
    hb = B.synthetic_batch(a.batch, cin, a.size, seed=1234)
    img, sk_t, fl_t = (t.to(dev) for t in (hb[0],) + hb[1])
    masks = [t.to(torch.uint8).contiguous() for t in (sk_t[:, 1], fl_t[:, 1])]
    target = tuple(masks) if B.HANDLER[a.model] == "double" else sk_t
    for _ in range(5):
        out = step(img, target)
    s_img, s_tgt = step.static_inputs()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.steps):
            out = step(s_img, s_tgt)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / a.steps)
    comps = step.last_components() if hasattr(step, "last_components") else None
    print("%s %s: %.4f ms/step  %s" % (a.tag, a.model, best, comps if comps is not None else ""), flush=True)


if __name__ == "__main__":
    main()
