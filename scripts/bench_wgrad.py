"""Times ctu_conv3d_wgrad (tensor path) alone on the layer shapes of the benchmarked step (CUDA events, 20 launches after
3 warm-ups, L2 flushed by the 270 MB of operands).  Env knobs CTU_WGRAD_V1 / CTU_WG2_* select the kernel variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ctunet_b200 import _lib
from ctunet_b200._lib import call, int_array, ptr_array, stream_ptr

CASES = [(7, 7, 4, 128), (2, 7, 4, 128), (14, 14, 4, 64), (7, 14, 4, 64), (14, 28, 4, 32), (28, 28, 4, 32)]


def main():
    lib = _lib.load()
    out = []
    for cin, cout, n, s in CASES:
        cb, cob = (cin + 7) // 8, (cout + 7) // 8
        x = torch.randn(n, cb, s, s, s, 8, device="cuda").to(torch.bfloat16)
        dy = torch.randn(n, cob, s, s, s, 8, device="cuda").to(torch.bfloat16)
        ca = int_array([cin])
        dwp = torch.empty(lib.ctu_conv_wpack_floats(cout, 3, 1, ca), device="cuda")
        def run():
            call("ctu_conv3d_wgrad", 1, ptr_array([x.data_ptr()]), ca, 1, dy.data_ptr(), dwp.data_ptr(), None, 0, cout, 3, n, s, s, s,
                 1, stream_ptr())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            run()
        b.record()
        torch.cuda.synchronize()
        out.append("%d->%d@%d^3 %.1f us" % (cin, cout, s, a.elapsed_time(b) / 20 * 1e3))
    print(" | ".join(out))


if __name__ == "__main__":
    main()
