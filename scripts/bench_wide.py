"""Weight-streaming kernel (conv_wide.cu, variant 2) against the resident-weights kernel (conv_tc.cu, variant 1) and the
CUDA-core kernel (variant 0) on single-source layers (GPU).  python scripts/bench_wide.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctunet_b200 import _lib
from ctunet_b200._lib import call, int_array, ptr_array, stream_ptr
from bench_kernels_common import act, timeit   # noqa: E402  (same directory)

lib = _lib.load()
dev = torch.device("cuda:0")


def fprop(variant, cin, cout, n, d, h, w, k):
    x, y = act(n, cin, d, h, w), act(n, cout, d, h, w)
    ca = int_array([cin])
    wp = torch.randn(lib.ctu_conv_wpack_floats(cout, k, 1, ca), device=dev) * 0.05
    if variant == 0:
        wk = wp
    elif variant == 1:
        if not lib.ctu_conv_tc_supported(k, 1, ca, cout, d, h, w):
            return None
        wk = torch.empty(lib.ctu_conv_tc_wimg_bytes(k, 1, ca, cout), dtype=torch.uint8, device=dev)
        call("ctu_conv_tc_pack_weight", wp.data_ptr(), wk.data_ptr(), k, 1, ca, cout, stream_ptr())
    else:
        if not lib.ctu_conv_wide_supported(k, cin, cout, n, d, h, w):
            return None
        wk = torch.empty(lib.ctu_conv_wide_wimg_bytes(k, cin, cout, n, d, h, w), dtype=torch.uint8, device=dev)
        call("ctu_conv_wide_pack_weight", wp.data_ptr(), wk.data_ptr(), k, cin, cout, n, d, h, w, stream_ptr())
    pa = ptr_array([x.data_ptr()])
    fn = lambda: call("ctu_conv3d_fprop", 1, pa, ca, 1, wk.data_ptr(), None, y.data_ptr(), None, 0, cout, k, n, d, h, w,
                      variant, stream_ptr())
    return timeit(fn)


SHAPES = [
    (3, 28, 28, 4, 32, 32, 32), (3, 14, 28, 4, 32, 32, 32), (3, 28, 14, 4, 32, 32, 32),
    (3, 28, 56, 4, 16, 16, 16), (3, 56, 56, 4, 16, 16, 16), (3, 56, 28, 4, 16, 16, 16),
    (3, 56, 112, 4, 8, 8, 8), (3, 112, 112, 4, 8, 8, 8),
    (3, 256, 56, 4, 16, 16, 16), (3, 128, 28, 4, 32, 32, 32), (3, 64, 14, 4, 64, 64, 64),
    (5, 28, 28, 4, 32, 32, 32), (5, 32, 32, 4, 32, 32, 32), (5, 16, 32, 4, 32, 32, 32),
    (5, 56, 56, 4, 16, 16, 16), (5, 64, 64, 4, 16, 16, 16), (5, 128, 64, 4, 16, 16, 16), (5, 64, 128, 4, 16, 16, 16),
    (5, 112, 56, 4, 16, 16, 16), (5, 32, 64, 4, 16, 16, 16), (5, 28, 56, 4, 16, 16, 16),
    (5, 64, 128, 4, 8, 8, 8), (5, 128, 128, 4, 8, 8, 8), (5, 56, 112, 4, 8, 8, 8), (5, 112, 112, 4, 8, 8, 8),
    (5, 16, 16, 4, 64, 64, 64), (3, 14, 14, 4, 64, 64, 64),
]
print("%-28s %10s %10s %10s   TFLOP/s(best)" % ("layer", "cuda-core", "resident", "streamed"))
for k, cin, cout, n, d, h, w in SHAPES:
    ts = [fprop(v, cin, cout, n, d, h, w, k) if not (v == 0 and cin * cout * k ** 3 * n * d * h * w > 3e12) else None
          for v in (0, 1, 2)]
    fl = 2.0 * n * d * h * w * cin * cout * k ** 3
    best = min(t for t in ts if t is not None)
    print("k%d %3d->%-3d @%dx%d^3 %-6s %10s %10s %10s   %.0f" % (
        k, cin, cout, n, d, "", *["%.1f" % t if t is not None else "-" for t in ts], fl / best / 1e6))


def wgrad(variant, cin, cout, n, d, h, w, k):
    x, dy = act(n, cin, d, h, w), act(n, cout, d, h, w)
    ca = int_array([cin])
    if variant == 1 and not lib.ctu_conv_tc_wgrad_supported(k, 1, ca, cout, d, h, w):
        return None
    if variant == 2 and not lib.ctu_conv_wide_wgrad_supported(k, cin, cout, d, h, w):
        return None
    dwp = torch.empty(lib.ctu_conv_wpack_floats(cout, k, 1, ca), device=dev)
    pa = ptr_array([x.data_ptr()])
    fn = lambda: call("ctu_conv3d_wgrad", 1, pa, ca, 1, dy.data_ptr(), dwp.data_ptr(), None, 0, cout, k, n, d, h, w, variant,
                      stream_ptr())
    return timeit(fn)


print("\nweight gradient: %-12s %10s %10s %10s" % ("layer", "cuda-core", "16-wide", "tap-stat."))
for k, cin, cout, n, d, h, w in [(3, 28, 56, 4, 16, 16, 16), (3, 56, 56, 4, 16, 16, 16), (3, 56, 112, 4, 8, 8, 8),
                                 (5, 32, 64, 4, 16, 16, 16), (5, 64, 64, 4, 16, 16, 16), (5, 128, 64, 4, 16, 16, 16),
                                 (5, 56, 56, 4, 16, 16, 16), (5, 112, 56, 4, 16, 16, 16),
                                 (5, 64, 128, 4, 8, 8, 8), (5, 128, 128, 4, 8, 8, 8),
                                 (3, 28, 28, 4, 32, 32, 32), (5, 32, 32, 4, 32, 32, 32), (5, 16, 32, 4, 32, 32, 32)]:
    ts = [wgrad(v, cin, cout, n, d, h, w, k) for v in (0, 1, 2)]
    print("k%d %3d->%-3d @%dx%d^3 %-6s %10s %10s %10s" % (k, cin, cout, n, d, "", *["%.1f" % t if t is not None else "-" for t in ts]))
