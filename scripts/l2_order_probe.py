"""Does the traversal ORDER of a bandwidth-bound consumer matter on B200's 126 MB L2?

A producer streams a 134 MB tensor (four 33.5 MB slabs, written slab 0..3, the size of one level-0 activation of the
benchmarked step); a consumer then reads it slab by slab either in the same order (0..3: the slabs it wants first are the
ones evicted longest ago) or in reverse (3..0: most recently written first).  Reports the consumer's time for both.

    python scripts/l2_order_probe.py
"""
import torch


def run(order, slabs, kind, reps=20):
    dev = torch.device("cuda")
    n = 33554432 // 2  # bf16 elements per slab: 33.5 MB
    src = torch.randn(slabs, n, device=dev, dtype=torch.bfloat16)
    mid = torch.empty_like(src)
    dst = torch.empty_like(src)
    acc = torch.zeros(slabs, device=dev, dtype=torch.float32)
    flush = torch.empty(512 * 1024 * 1024, device=dev, dtype=torch.uint8)
    ts = []
    for _ in range(reps):
        flush.zero_()
        for i in range(slabs):          # producer: reads src, writes mid, slab 0..slabs-1
            torch.add(src[i], 1.0, out=mid[i])
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in order:
            if kind == "copy":          # consumer reads mid, writes dst (BatchNorm forward / backward-apply shape)
                torch.add(mid[i], 1.0, out=dst[i])
            else:                       # consumer only reads (reduction)
                acc[i] = mid[i].float().sum() if False else torch.sum(mid[i], dtype=torch.float32)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]


def main():
    for slabs in (4, 8):
        for kind in ("copy", "reduce"):
            f = run(list(range(slabs)), slabs, kind)
            r = run(list(range(slabs - 1, -1, -1)), slabs, kind)
            mb = slabs * 33.5
            print("slabs %d (%.0f MB) %-6s consumer: same order %.1f us, reverse order %.1f us (%.2fx)" % (slabs, mb, kind, f, r, f / r))


if __name__ == "__main__":
    main()
