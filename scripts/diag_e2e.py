"""End-to-end loop diagnostics on N GPUs (torchrun): where the gap between the resident step and the host-fed step comes
from.  Per rank, max over ranks: H2D of one batch of uint8 masks alone; the captured step alone; the e2e loop with the
prefetch one / two batches ahead and with / without the loss read-back."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ctunet_b200 as C
from ctunet_b200.parallel import PeerGradSync
from ctunet_b200.synthetic import make_training_batch
from ctunet_b200.trainer import LossReadback, TrainStep


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        dist.init_process_group("nccl", device_id=dev)
    steps = 30
    img, (sk_t, fl_t) = make_training_batch(4, 2, 128, seed=100 + rank, device=dev)
    masks = [t.to(torch.uint8).contiguous() for t in (img[:, 0], sk_t[:, 1], fl_t[:, 1])]
    atlas = img[0, 1].contiguous()
    host = [m.cpu().pin_memory() for m in masks]
    frac = float(os.environ.get("H2D_FRACTION", "1"))
    nb = max(1, int(masks[0].shape[0] * frac * 1024)) if frac < 1 else None      # copy only the first nb KiB-ish rows
    torch.manual_seed(0)
    net = C.UNetSP().to(dev)
    step = TrainStep(net, "double", 1.0, 1.0, lr=1e-4, scheduler=True, grad_sync=PeerGradSync(net) if world > 1 else None, graph=True)
    for _ in range(4):
        step.step_from_masks(*masks, atlas)

    def timed(fn, n=steps, after=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        if after:
            after()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    out = {}
    out["resident"] = timed(lambda: step.step_from_masks(*masks, atlas))
    cs = torch.cuda.Stream(device=dev)
    dbuf = [[torch.empty_like(m) for m in masks] for _ in range(3)]

    def cp(d_, s_):
        if frac < 1:
            k = max(1, int(d_.numel() * frac))
            d_.view(-1)[:k].copy_(s_.view(-1)[:k], non_blocking=True)
        else:
            d_.copy_(s_, non_blocking=True)

    def h2d():
        with torch.cuda.stream(cs):
            for d_, s_ in zip(dbuf[0], host):
                cp(d_, s_)
        torch.cuda.current_stream().wait_stream(cs)
    out["h2d_alone"] = timed(h2d)

    def overlapped():            # H2D on the side stream, NO dependency with the step: pure interference
        with torch.cuda.stream(cs):
            for d_, s_ in zip(dbuf[1], host):
                cp(d_, s_)
        step.step_from_masks(*masks, atlas)
    out["step+independent_h2d"] = timed(overlapped)
    torch.cuda.synchronize()

    def make(depth, readback):
        st = {"i": 0, "ev": {}}
        rb = LossReadback(len(step.keys), depth=2)

        def prefetch(i):
            slot = i % 3
            with torch.cuda.stream(cs):
                for d_, s_ in zip(dbuf[slot], host):
                    cp(d_, s_)
                ev = torch.cuda.Event()
                ev.record(cs)
            st["ev"][i] = ev

        def one():
            i = st["i"]
            for j in range(i, i + depth + 1):
                if j not in st["ev"]:
                    if j >= 3:      # slot reuse: the consumer of batch j-3 must be enqueued and done
                        cs.wait_stream(torch.cuda.current_stream())
                    prefetch(j)
            torch.cuda.current_stream().wait_event(st["ev"].pop(i))
            comps = step.step_from_masks(*dbuf[i % 3], atlas)
            st["i"] += 1
            if readback:
                rb.push(comps)
        return one, rb

    for depth in (1,):
        for readback in (True,):
            one, rb = make(depth, readback)
            for _ in range(3):
                one()
            out["e2e_depth%d_%s" % (depth, "rb" if readback else "norb")] = timed(one, after=rb.drain if readback else None)
    if rank == 0:
        print("world %d: " % world + "  ".join("%s %.3f" % kv for kv in out.items()))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
