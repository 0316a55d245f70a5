"""Diagnostic (GPU): per-parameter gradient error of the product path against the CPU oracle."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctunet_b200 as C
from oracle import unet_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "UNetSP"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 32
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kind = sys.argv[4] if len(sys.argv) > 4 else "binary"
cfg = O.PRESETS[name]
handler = "double" if cfg.head != "plain" else "single"
g = torch.Generator().manual_seed(7)
x = (torch.rand(batch, cfg.input_channels, size, size, size, generator=g) > 0.7).float()
if kind == "cont":
    x = torch.rand(batch, cfg.input_channels, size, size, size, generator=g)
g = torch.Generator().manual_seed(11)
sk = (torch.rand(batch, size, size, size, generator=g) > 0.6).long()
fl = ((torch.rand(batch, size, size, size, generator=g) > 0.8) & (sk > 0)).long()
oh = lambda t: torch.nn.functional.one_hot(t, 2).permute(0, 4, 1, 2, 3).float().contiguous()
sk_t, fl_t = oh(sk), oh(fl)

sd = O.build_state_dict(cfg, seed=0)
pn = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
for k in pn:
    sd[k].requires_grad_()
out = O.unet_forward(sd, x.clone().requires_grad_(), cfg, training=True)
loss, _ = (O.loss_double_output(out, (sk_t, fl_t), 1., 1.) if handler == "double" else O.loss_single_output(out, sk_t, 1., 1.))
loss.backward()
for mode in ("fp32", "bf16"):
    C.set_compute_dtype(mode)
    torch.manual_seed(0)
    net = getattr(C, name)().cuda().train()
    step = C.trainer.TrainStep(net, handler) if hasattr(C, "trainer") else None
    o = net(x.cuda().requires_grad_())
    fake = types.SimpleNamespace(params=dict(dice_lambda=1., ce_lambda=1., save_dice_plots=False, save_hd_plots=False), losses_and_metrics={}, pt_loss=None)
    if handler == "double":
        C.FlapRecWithShapePriorDoubleOut.comp_losses_metrics(fake, o, (sk_t.cuda(), fl_t.cuda()), 0, 1, verbose=False)
    else:
        C.ProblemHandler.comp_losses_metrics(fake, o, sk_t.cuda(), 0, 1, verbose=False)
    fake.pt_loss.backward()
    print("==", name, mode, "loss", float(fake.pt_loss), "oracle", float(loss))
    oo = o if isinstance(o, tuple) else (o,)
    rr = out if isinstance(out, tuple) else (out,)
    for a, b in zip(oo, rr):
        print("   out relerr(max)", float((a.detach().cpu() - b.detach()).abs().max() / b.detach().abs().max()))
    for k, p in net.named_parameters():
        if sd[k].grad is None:
            continue
        gr, gg = sd[k].grad.double(), p.grad.cpu().double()
        cos = float((gg * gr).sum() / (gg.norm() * gr.norm()))
        print("   %-28s |g| %.3e  normwise rel %.3e  max rel %.3e cos %.4f" % (k, gr.norm(), (gg - gr).norm() / gr.norm(), (gg - gr).abs().max() / gr.abs().max(), cos))
