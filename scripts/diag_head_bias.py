"""Diagnostic: last_conv.bias gradient of the bf16 product step vs the bf16-storage oracle under engine variants."""
import os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctunet_b200 as C
import ctunet_b200.engine as E
from ctunet_b200._lib import CTU_BF16
from oracle import unet_oracle as O

size, batch = int(sys.argv[1]), int(sys.argv[2])
cfg = O.PRESETS["UNetSP"]
g = torch.Generator().manual_seed(99)
x = torch.rand(batch, 2, size, size, size, generator=g)
_, target = O.make_training_batch(batch, 2, size, seed=77)
sd = O.build_state_dict(cfg, seed=0)
pn = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
for k in pn:
    sd[k].requires_grad_()
out = O.unet_forward(sd, x, cfg, training=True, act_round=torch.bfloat16)
out[0].retain_grad(); out[1].retain_grad()
loss, _ = O.loss_double_output(out, target, 1.0, 1.0)
loss.backward()
ref = {k: sd[k].grad for k in pn}

def run(tag):
    torch.manual_seed(0)
    net = C.UNetSP().to("cuda").train()
    fake = types.SimpleNamespace(params=dict(dice_lambda=1.0, ce_lambda=1.0, save_dice_plots=False, save_hd_plots=False),
                                 losses_and_metrics={}, pt_loss=None)
    o = net(x.cuda().requires_grad_())
    o[0].retain_grad(); o[1].retain_grad()
    C.FlapRecWithShapePriorDoubleOut.comp_losses_metrics(fake, o, tuple(t.cuda() for t in target), 0, 1, verbose=False)
    fake.pt_loss.backward()
    torch.cuda.synchronize()
    named = dict(net.named_parameters())
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    print("%-28s bias %.5f  weight %.5f  bn5.bias %.5f  dpred0 %.2e dpred1 %.2e out0 %.2e" % (
        tag, rel(named["last_conv.bias"].grad, ref["last_conv.bias"]), rel(named["last_conv.weight"].grad, ref["last_conv.weight"]),
        rel(named["u_blocks.3.block.5.bias"].grad, ref["u_blocks.3.block.5.bias"]),
        rel(o[0].grad, out[0].grad), rel(o[1].grad, out[1].grad), rel(o[0].detach(), out[0].detach())))
    print("   bias grad product", named["last_conv.bias"].grad.tolist(), "oracle", ref["last_conv.bias"].tolist())

run("default")
E.HEAD_FROM_OUTPUTS = False
run("HEAD_FROM_OUTPUTS off")
E.HEAD_FROM_OUTPUTS = True
# the head's parameter gradients recomputed with torch from the product's own tensors
