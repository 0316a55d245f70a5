import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctunet_b200 import _lib
from ctunet_b200._lib import call, int_array, ptr_array, stream_ptr
lib = _lib.load(); dev = torch.device("cuda:0")
def act(n, c, d, h, w):
    cb = (c + 7) // 8
    return torch.randn(n, cb, d, h, w, 8, device=dev).to(torch.bfloat16)
chans = [int(a) for a in sys.argv[1].split(",")]; cnat = int(sys.argv[2])
cop = (cnat + 7) // 8 * 8; cout = 8 * cop; P = cop // 8
n, d, h, w = 1, 3, 16, 16
srcs = [act(n, c, d, h, w) for c in chans]; dy = act(n, cout, d, h, w)
ca = int_array(chans); pa = ptr_array([s.data_ptr() for s in srcs]); ns = len(chans)
outs = []
for phase in (0, cnat):
    dwp = torch.empty(lib.ctu_conv_wpack_floats(cout, 3, ns, ca), device=dev)
    call("ctu_conv3d_wgrad", 1, pa, ca, ns, dy.data_ptr(), dwp.data_ptr(), None, phase, cout, 3, n, d, h, w, 1, stream_ptr())
    dw = torch.zeros(cout, sum(chans), 27, device=dev)
    call("ctu_conv_unpack_wgrad", dwp.data_ptr(), dw.data_ptr(), cout, 3, ns, ca, stream_ptr())
    torch.cuda.synchronize(); outs.append(dw.cpu().view(8, cop, sum(chans), 3, 3, 3))
a, b = outs
for q in range(8):
    qd, qh, qw = q >> 2, (q >> 1) & 1, q & 1
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                valid = kd in (qd, qd + 1) and kh in (qh, qh + 1) and kw in (qw, qw + 1)
                e = (a[q, :, :, kd, kh, kw] - b[q, :, :, kd, kh, kw]).abs().max().item()
                ref = a[q, :, :, kd, kh, kw].abs().max().item()
                if valid and e > 1e-2 * ref:
                    print("q", (qd, qh, qw), "tap", (kd, kh, kw), "err", round(e, 3), "ref", round(ref, 3), "sparse max", round(b[q, :, :, kd, kh, kw].abs().max().item(), 3))
print("done", "dense norm", float(a.norm()), "sparse norm", float(b.norm()), "nan a", bool(torch.isnan(a).any()))
cin = sum(chans) - 1
wt = torch.randn(cin, cin, 2, 2, 2, device=dev); bt = torch.randn(cin, device=dev); w3 = torch.randn(cnat, cin, 3, 3, 3, device=dev)
res = []
for dw in outs:
    dwn = dw.reshape(cout, cin + 1, 27).contiguous().to(dev)
    dwt = torch.empty_like(wt); dbt = torch.empty_like(bt); dw3 = torch.empty_like(w3)
    call("ctu_upfuse_decompose", dwn.data_ptr(), None, wt.data_ptr(), bt.data_ptr(), w3.data_ptr(), dwt.data_ptr(), dbt.data_ptr(), dw3.data_ptr(), None, cin, cnat, 3, stream_ptr())
    torch.cuda.synchronize(); res.append((dwt.cpu(), dbt.cpu(), dw3.cpu()))
for nm, x, y in zip(("dwt", "dbt", "dw3"), res[0], res[1]):
    print(nm, "rel diff", float((x - y).norm() / x.norm()), "norms", float(x.norm()), float(y.norm()), "nan", bool(torch.isnan(x).any()), bool(torch.isnan(y).any()))
# which invalid entries are non-zero in the sparse result?
nz = 0
for q in range(8):
    qd, qh, qw = q >> 2, (q >> 1) & 1, q & 1
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                valid = kd in (qd, qd + 1) and kh in (qh, qh + 1) and kw in (qw, qw + 1)
                m = b[q, :, :, kd, kh, kw].abs().max().item()
                if not valid and m > 0: nz += 1
print("non-zero invalid (q,tap) blocks in sparse result:", nz, "has nan:", bool(torch.isnan(b).any()))
