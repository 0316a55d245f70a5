"""Time individual C-ABI kernels on the shapes of the UNetSP step (GPU).  python scripts/bench_kernels.py [wgrad|fprop]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctunet_b200 import _lib
from ctunet_b200._lib import call, int_array, ptr_array, stream_ptr

lib = _lib.load()
dev = torch.device("cuda:0")


def act(n, c, d, h, w):
    cb = (c + 7) // 8
    return (torch.randn(n, cb, d, h, w, 8, device=dev) * 0.5).to(torch.bfloat16)


def timeit(fn, reps=20):
    """Device time per call: the calls are captured in a CUDA graph (no host launch overhead between kernels)."""
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


def wgrad(chans, cout, n, d, h, w, k=3, phase=0):
    srcs = [act(n, c, d, h, w) for c in chans]
    dy = act(n, cout, d, h, w)
    ca = int_array(chans)
    dwp = torch.empty(lib.ctu_conv_wpack_floats(cout, k, len(chans), ca), device=dev)
    pa = ptr_array([s.data_ptr() for s in srcs])
    fn = lambda: call("ctu_conv3d_wgrad", 1, pa, ca, len(chans), dy.data_ptr(), dwp.data_ptr(), None, phase, cout, k, n, d, h, w, 1,
                      stream_ptr())
    return timeit(fn)


def fprop(chans, cout, n, d, h, w, k=3, stat_cout=0):
    srcs = [act(n, c, d, h, w) for c in chans]
    y = act(n, cout, d, h, w)
    ca = int_array(chans)
    wp = torch.randn(lib.ctu_conv_wpack_floats(cout, k, len(chans), ca), device=dev) * 0.05
    wimg = torch.empty(lib.ctu_conv_tc_wimg_bytes(k, len(chans), ca, cout), dtype=torch.uint8, device=dev)
    call("ctu_conv_tc_pack_weight", wp.data_ptr(), wimg.data_ptr(), k, len(chans), ca, cout, stream_ptr())
    sums = torch.zeros(2 * 512, dtype=torch.float64, device=dev)
    pa = ptr_array([s.data_ptr() for s in srcs])
    fn = lambda: call("ctu_conv3d_fprop", 1, pa, ca, len(chans), wimg.data_ptr(), None, y.data_ptr(), sums.data_ptr(), stat_cout,
                      cout, k, n, d, h, w, 1, stream_ptr())
    return timeit(fn)


SHAPES = [
    ("L0 7->7 @128^3", [7], 7, 4, 128, 128, 128),
    ("L0 2->7 @128^3", [2], 7, 4, 128, 128, 128),
    ("L0 up-fused [14,14,1]->64 @64^3", [14, 14, 1], 64, 4, 64, 64, 64),
    ("L1 14->14 @64^3", [14], 14, 4, 64, 64, 64),
    ("L1 up-fused [28,28,1]->128 @32^3", [28, 28, 1], 128, 4, 32, 32, 32),
    ("L2 28->28 @32^3", [28], 28, 4, 32, 32, 32),
    ("L2 up-fused [56,56,1]->256 @16^3", [56, 56, 1], 256, 4, 16, 16, 16),
]
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which == "k5":      # the legacy 5^3 family at the benchmarked sizes (recAE_v2_fixed / UNet4_2IC level 0 and 1)
    for name, chans, cout, n, d, h, w in [("5^3 8->8 @128^3", [8], 8, 4, 128, 128, 128), ("5^3 1->8 @128^3", [1], 8, 4, 128, 128, 128),
                                          ("5^3 16->16 @64^3", [16], 16, 4, 64, 64, 64)]:
        t = fprop(chans, cout, n, d, h, w, 5)
        fl = 2.0 * n * d * h * w * sum(chans) * cout * 125
        print("%-20s fprop %8.1f us = %6.1f TFLOP/s   wgrad %8.1f us" % (name, t, fl / t / 1e6, wgrad(chans, cout, n, d, h, w, 5)), flush=True)
    sys.exit(0)
for name, chans, cout, n, d, h, w in SHAPES:
    line = "%-36s" % name
    if which in ("wgrad", "both"):
        line += " wgrad %8.1f us" % wgrad(chans, cout, n, d, h, w, 3, cout // 8 - 1 if len(chans) == 3 else 0)
    if which in ("fprop", "both"):
        sc = cout // 8 if len(chans) == 3 else 0
        line += " fprop %8.1f us" % fprop(chans, cout, n, d, h, w, 3, sc)
        if len(chans) == 3:      # data gradients of the fused stage: cout channels -> each source
            line += " dgrad " + " ".join("%6.1f" % fprop([cout], c, n, d, h, w) for c in chans[:2])
        else:
            line += " dgrad %8.1f us" % fprop([cout], chans[0], n, d, h, w)
    print(line, flush=True)


def bn_bwd(c, n, d, h, w, pool=False, pm=False):
    y = act(n, c * (8 if pm else 1) if pm else c, d // (2 if pm else 1), h // (2 if pm else 1), w // (2 if pm else 1)) if pm else act(n, c, d, h, w)
    dA = act(n, c, d, h, w)
    dP = act(n, c, d // 2, h // 2, w // 2) if pool else None
    dy = torch.empty_like(y)
    cpad = (c + 7) // 8 * 8
    ss = torch.rand(4 * cpad, device=dev)
    gamma = torch.rand(cpad, device=dev)
    sums2 = torch.zeros(2 * cpad, dtype=torch.float64, device=dev)
    dg, db = torch.empty(cpad, device=dev), torch.empty(cpad, device=dev)
    pA, pP = dA.data_ptr(), (dP.data_ptr() if pool else None)
    r = timeit(lambda: call("ctu_bn_relu_bwd_reduce", 1, y.data_ptr(), ss.data_ptr(), pA, pP, sums2.data_ptr(), c, n, d, h, w,
                            int(pm), stream_ptr()))
    a = timeit(lambda: call("ctu_bn_relu_bwd_apply", 1, y.data_ptr(), ss.data_ptr(), gamma.data_ptr(), pA, pP, sums2.data_ptr(),
                            float(n * d * h * w), dy.data_ptr(), dg.data_ptr(), db.data_ptr(), c, n, d, h, w, int(pm),
                            stream_ptr()))
    return r, a


if which in ("bn", "both"):
    for name, kw in (("plain", {}), ("pool", {"pool": True}), ("phase-major", {"pm": True})):
        r, a = bn_bwd(7, 4, 128, 128, 128, **kw)
        print("BN bwd c7 @4x128^3 %-12s reduce %7.1f us  apply %7.1f us" % (name, r, a), flush=True)
