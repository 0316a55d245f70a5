"""Per-tensor gradient accuracy of the bf16 product path against three CPU statements of the same step
(profiles/r2_grad_parity.txt):

  fp32      the reference arithmetic (oracle, fp32 activations)
  bf16      the oracle with the product's bf16 ACTIVATION STORAGE points (act_round=bf16), fp32 arithmetic
  bf16'     the same bf16-storage oracle on an input perturbed by 3e-7 relative noise (one fp32 ulp or two): how far two
            equally valid evaluation orders of the SAME bf16 network are apart -- the floor any implementation sits on

Usage: python scripts/grad_noise_study.py [size] [batch]   (GPU box; imports oracle/ as the checker)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctunet_b200 as C
    from oracle import unet_oracle as O
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    name = "UNetSP"
    cfg = O.PRESETS[name]
    g = torch.Generator().manual_seed(99)
    x = torch.rand(batch, cfg.input_channels, size, size, size, generator=g)
    xp = x * (1 + 3e-7 * torch.randn(x.shape, generator=g))
    _, target = O.make_training_batch(batch, cfg.input_channels, size, seed=77)

    def oracle(inp, act_round, grad_round=False, weight_round=False):
        sd = O.build_state_dict(cfg, seed=0)
        pn = [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        for k in pn:
            sd[k].requires_grad_()
        out = O.unet_forward(sd, inp, cfg, training=True, act_round=act_round, grad_round=grad_round, weight_round=weight_round)
        loss, _ = O.loss_double_output(out, target, 1.0, 1.0)
        loss.backward()
        return {k: sd[k].grad for k in pn}, float(loss)

    g32, l32 = oracle(x, None)
    g16, l16 = oracle(x, torch.bfloat16)
    g16p, l16p = oracle(xp, torch.bfloat16)
    g16g, _ = oracle(x, torch.bfloat16, True)
    g16gp, _ = oracle(xp, torch.bfloat16, True)
    g16w, _ = oracle(x, torch.bfloat16, True, True)
    torch.manual_seed(0)
    net = C.UNetSP().to("cuda").train()
    fake = types.SimpleNamespace(params=dict(dice_lambda=1.0, ce_lambda=1.0, save_dice_plots=False, save_hd_plots=False),
                                 losses_and_metrics={}, pt_loss=None)
    out = net(x.cuda().requires_grad_())
    C.FlapRecWithShapePriorDoubleOut.comp_losses_metrics(fake, out, tuple(t.cuda() for t in target), 0, 1, verbose=False)
    fake.pt_loss.backward()
    gp = {k: (p.grad.cpu() if p.grad is not None else None) for k, p in net.named_parameters()}
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    print("UNetSP batch %d x %d^3, continuous inputs; losses: fp32 %.6f  bf16-oracle %.6f  bf16-oracle' %.6f  product %.6f"
          % (batch, size, l32, l16, l16p, float(fake.pt_loss)))
    print("(bf16g = bf16 storage of activations AND of their gradients, as the product stores them)")
    print("%-28s %10s %12s %12s %12s %12s %12s %12s" % ("tensor", "|g| fp32", "prod~bf16", "bf16'~bf16", "bf16~fp32", "prod~fp32",
                                                      "prod~bf16g", "bf16g'~bf16g") + " %12s %12s" % ("bf16gw~bf16g", "prod~bf16gw"))
    print("(bf16gw = additionally the convolution weights rounded to bf16 for the products, as the tensor-core kernels do)")
    tot = {k: [0.0, 0.0] for k in ("pb", "bb", "bf", "pf", "pg", "gg", "wg", "pw")}
    for k, v in g32.items():
        if v is None:
            continue
        n = float(v.norm())
        if n < 1e-6:
            continue
        print("%-28s %10.3e %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f" % (
            k, n, rel(gp[k], g16[k]), rel(g16p[k], g16[k]), rel(g16[k], v), rel(gp[k], v), rel(gp[k], g16g[k]),
            rel(g16gp[k], g16g[k]), rel(g16w[k], g16g[k]), rel(gp[k], g16w[k])))
        for key, a, b in (("pb", gp[k], g16[k]), ("bb", g16p[k], g16[k]), ("bf", g16[k], v), ("pf", gp[k], v),
                          ("pg", gp[k], g16g[k]), ("gg", g16gp[k], g16g[k]), ("wg", g16w[k], g16g[k]), ("pw", gp[k], g16w[k])):
            tot[key][0] += float((a.double() - b.double()).norm() ** 2)
            tot[key][1] += float(b.double().norm() ** 2)
    print("%-28s %10s %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f %12.4f" % (
        "WHOLE GRADIENT (normwise)", "", *[(tot[k][0] / tot[k][1]) ** 0.5 for k in ("pb", "bb", "bf", "pf", "pg", "gg", "wg", "pw")]))


if __name__ == "__main__":
    main()
