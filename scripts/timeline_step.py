"""Timeline of one eager, multi-stream UNetSP training step: a CUDA-event pair around every C-ABI call on the stream it is
launched on, all time-stamped against one base event.  A spin kernel holds the main stream until the host has enqueued
the whole step, so the picture is the device-side schedule (what the captured graph replays), not the host's.

    python scripts/timeline_step.py [--model UNetSP] [--out gpurun_out/timeline.txt]

Columns: stream (0 = main), start us, end us, duration us, call.  Event records cost ~2-3 us each and serialise with the
kernels of their stream: read the structure (who waits for whom, idle spans of the main stream), not absolute times.
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.txt"))
    a = ap.parse_args()
    import torch
    import bench as B
    import ctunet_b200 as C
    from ctunet_b200 import _lib
    from ctunet_b200.trainer import TrainStep
    import ctunet_b200.engine as E
    import ctunet_b200.losses as LS
    import ctunet_b200.optim as OP
    import ctunet_b200.trainer as TR
    import ctunet_b200.utilities as UT

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.load()
    C.set_compute_dtype("bf16")
    torch.manual_seed(0)
    net = getattr(C, a.model)().to(dev)
    step = TrainStep(net, B.HANDLER[a.model], 1.0, 1.0, lr=1e-4, scheduler=True, graph=False)
    cin = B.in_channels(a.model)
    hb = B.synthetic_batch(a.batch, cin, a.size, seed=1234)
    img, sk_t, fl_t = (t.to(dev) for t in (hb[0],) + hb[1])
    masks = [t.to(torch.uint8).contiguous() for t in (sk_t[:, 1], fl_t[:, 1])]
    target = tuple(masks) if B.HANDLER[a.model] == "double" else sk_t
    for _ in range(3):
        step(img, target)
    torch.cuda.synchronize()

    records = []
    orig = _lib.call
    main_stream = torch.cuda.current_stream().cuda_stream
    ids = {main_stream: 0}

    def timed(name, *args):
        st = torch.cuda.current_stream()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record(st)
        orig(name, *args)
        e.record(st)
        sid = ids.setdefault(st.cuda_stream, len(ids))
        records.append((sid, name, args, s, e))

    mods = (E, LS, OP, TR, UT)
    for m in mods:
        m.call = timed
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    torch.cuda._sleep(int(0.06 * 1.9e9))
    t0 = torch.cuda.Event(enable_timing=True)
    t0.record()
    step(img, target)
    t1 = torch.cuda.Event(enable_timing=True)
    t1.record()
    torch.cuda.synchronize()
    for m in mods:
        m.call = orig
    lines = []
    zero = base.elapsed_time(t0)
    rows = []
    for sid, name, args, s, e in records:
        key = B.describe(name, args, 2)[0]
        rows.append((base.elapsed_time(s) - zero, base.elapsed_time(e) - zero, sid, key))
    rows.sort()
    lines.append("# step span %.1f us, %d calls, streams %d" % (1e3 * t0.elapsed_time(t1), len(rows), len(ids)))
    busy = {}
    last_end = {}
    for s, e, sid, key in rows:
        gap = s - last_end.get(sid, s)
        last_end[sid] = e
        busy[sid] = busy.get(sid, 0.0) + (e - s)
        lines.append("%d %9.1f %9.1f %7.1f gap %6.1f  %s" % (sid, 1e3 * s, 1e3 * e, 1e3 * (e - s), 1e3 * gap, key))
    for sid, b in sorted(busy.items()):
        lines.append("# stream %d busy %.1f us" % (sid, 1e3 * b))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print(lines[0])
    for l in lines[-len(busy):]:
        print(l)


if __name__ == "__main__":
    main()
