"""One training step of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).

    python scripts/profile_step.py [model] [batch] [size] [warmup]

Prints the device time of the profiled step so that a plain run (no ncu) doubles as a sanity check.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench as B
import ctunet_b200 as C
from ctunet_b200.trainer import TrainStep

model = sys.argv[1] if len(sys.argv) > 1 else "UNetSP"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
size = int(sys.argv[3]) if len(sys.argv) > 3 else 128
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 2
handler = B.HANDLER[model]
cin = B.in_channels(model)

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
C.set_compute_dtype("bf16")
torch.manual_seed(0)
net = getattr(C, model)().to(dev)
# the bench's step (bench.py run_b200_arm), eager so that every kernel is a separate launch for ncu
step = TrainStep(net, handler, 1.0, 1.0, lr=1e-4, scheduler=True)
hb = B.synthetic_batch(batch, cin, size, seed=1234)
img, sk_t, fl_t = (t.to(dev) for t in (hb[0],) + hb[1])
target = tuple(t.to(torch.uint8).contiguous() for t in (sk_t[:, 1], fl_t[:, 1])) if handler == "double" else sk_t
for _ in range(warm):
    step(img, target)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
s.record()
comps = step(img, target)
e.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("step ms %.3f  loss %s" % (s.elapsed_time(e), [round(v, 5) for v in comps.tolist()]))
