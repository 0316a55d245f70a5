"""Device timeline of the CAPTURED training step (CUDA-graph replay) from CUPTI via torch.profiler: every kernel with its
stream, start and duration -- the real concurrency of the main / weight-gradient / weight-preparation branches.

    python scripts/trace_step.py [--model UNetSP] [--out gpurun_out/trace.txt]
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.txt"))
    a = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
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
        step(img, target)
    s_img, s_tgt = step.static_inputs()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step(s_img, s_tgt)
        torch.cuda.synchronize()
    import json
    tmp = a.out + ".chrome.json"
    prof.export_chrome_trace(tmp)
    tr = json.load(open(tmp))
    os.remove(tmp)
    evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "ts" in e]
    evs.sort(key=lambda e: e["ts"])
    n = len(evs) // 3          # three identical replays: keep the last
    evs = evs[2 * n:]
    t0 = evs[0]["ts"]
    streams = {}
    lines = []
    end_max = 0.0
    busy = {}
    for e in evs:
        sid = streams.setdefault(e.get("args", {}).get("stream", 0), len(streams))
        s = e["ts"] - t0
        d = e.get("dur", 0.0)
        end_max = max(end_max, s + d)
        busy[sid] = busy.get(sid, 0.0) + d
        lines.append("%d %9.1f %8.1f  %s" % (sid, s, d, e["name"][:120]))
    hdr = "# %d kernels, span %.1f us, streams %d, busy per stream %s" % (
        len(evs), end_max, len(streams), {k: round(v, 1) for k, v in sorted(busy.items())})
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write(hdr + "\n" + "\n".join(lines) + "\n")
    print(hdr)


if __name__ == "__main__":
    main()
