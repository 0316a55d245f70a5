import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_upfuse as T
import ctunet_b200.engine as E
case = T.STAGES_TC[int(sys.argv[1]) if len(sys.argv) > 1 else 3]
res = {}
orig = E.Engine._conv_bwd
for mode in ("sparse", "dense", "sparse-sync"):
    E.WGRAD_ASYNC = mode != "sparse-sync"
    if mode == "dense":
        def patched(self, *a, **kw):
            kw["phase_cout"] = 0
            return orig(self, *a, **kw)
        E.Engine._conv_bwd = patched
    else:
        E.Engine._conv_bwd = orig
    ar, ao, gx, dxs, rg, gg = T._run_stage("bf16", case, force=False)
    print(mode, {k: round(float((gg[k] - rg[k]).norm() / rg[k].norm()), 4) for k in rg})
