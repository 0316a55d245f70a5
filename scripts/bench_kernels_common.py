"""Helpers shared by the per-kernel timing scripts (GPU)."""
import torch

dev = torch.device("cuda:0")


def act(n, c, d, h, w):
    cb = (c + 7) // 8
    return (torch.randn(n, cb, d, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)


def timeit(fn, reps=20):
    """Device time per call in microseconds: the calls are captured in a CUDA graph (no host launch gaps)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3
