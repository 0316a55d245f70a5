"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt from the UNMODIFIED reference.

Run inside the build container (needs /root/reference):

    python oracle/make_golden.py

The reference ships no tests, fixtures or golden vectors of its own, so
parity is pinned by executing the reference's own classes / functions on seeded synthetic inputs
and freezing what they return.  The fixtures are small (weights are NOT stored: every model is
re-created from ``torch.manual_seed(seed)`` and checked against the stored checksums).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CLASSES = ["UNet", "UNetSP", "UNetSPSmall", "UNetDO", "UNet4_2IC", "recAE_v2_fixed"]


def synth_input(cin, size, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(1, cin, size, size, size, generator=g) > 0.7).float()


def synth_targets(batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    sk = (torch.rand(batch, size, size, size, generator=g) > 0.6).long()
    fl = ((torch.rand(batch, size, size, size, generator=g) > 0.8) & (sk > 0)).long()
    oh = lambda t: torch.nn.functional.one_hot(t, 2).permute(0, 4, 1, 2, 3).float().contiguous()
    return oh(sk), oh(fl)


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    MD, PH, UT, TR = load_reference()
    gold = {"torch_version": str(torch.__version__), "classes": {}}   # plain data only: the tests load it with weights_only=True

    # ---- (1) construction + eval forward known answers ----
    for name in CLASSES:
        torch.manual_seed(0)
        net = getattr(MD, name)()
        sd = net.state_dict()
        params = list(net.parameters())
        cin = params[0].shape[1]
        entry = {
            "n_params": sum(p.numel() for p in params),
            "n_state_entries": len(sd),
            "abs_sum": float(sum(p.detach().double().abs().sum() for p in params)),
            "first3": params[0].detach().flatten()[:3].clone(),
            "keys": list(sd.keys()),
            "shapes": [tuple(v.shape) for v in sd.values()],
            "cin": cin,
        }
        net.eval()
        size = 32
        x = synth_input(cin, size, 1)
        with torch.no_grad():
            out = net(x)
        outs = out if isinstance(out, tuple) else (out,)
        entry["eval32"] = {
            "x_sum": float(x.sum()),
            "out_sums": [float(o.double().sum()) for o in outs],
            "argmax_ones": [int(torch.argmax(o, 1).sum()) for o in outs],
            "out_slices": [o[:, :, 12:20, 12:20, 12:20].clone() for o in outs],
            "hard_segm_slice": [UT.hard_segm_from_tensor(o)[:, 12:20, 12:20, 12:20].clone() for o in outs],
        }
        gold["classes"][name] = entry
        print(name, entry["n_params"], entry["abs_sum"], entry["eval32"]["out_sums"])

    # ---- (2) one full training step through the reference's own loss handlers ----
    steps = {}
    for name, handler, size, batch in [("UNetSP", "double", 16, 2), ("UNetDO", "double", 32, 1),
                                       ("UNetSPSmall", "double", 32, 2), ("UNet4_2IC", "single", 16, 2),
                                       ("recAE_v2_fixed", "single", 16, 2)]:
        torch.manual_seed(0)
        net = getattr(MD, name)()
        net.train()
        cin = next(net.parameters()).shape[1]
        g = torch.Generator().manual_seed(7)
        x = (torch.rand(batch, cin, size, size, size, generator=g) > 0.7).float()
        sk_t, fl_t = synth_targets(batch, size, 11)
        fake = types.SimpleNamespace(params=dict(dice_lambda=1.0, ce_lambda=1.0, save_dice_plots=False,
                                                 save_hd_plots=False), losses_and_metrics={}, pt_loss=None)
        x.requires_grad_()                      # Model.py:351-352
        out = net(x)
        if handler == "double":
            PH.FlapRecWithShapePriorDoubleOut.comp_losses_metrics(fake, out, (sk_t, fl_t), 0, 1)
        else:
            PH.ProblemHandler.comp_losses_metrics(fake, out, sk_t, 0, 1)
        fake.pt_loss.backward()
        outs = out if isinstance(out, tuple) else (out,)
        rec = {
            "size": size, "batch": batch, "handler": handler,
            "loss": float(fake.pt_loss), "components": {k: v[0] for k, v in fake.losses_and_metrics.items()},
            "out_sums": [float(o.detach().double().sum()) for o in outs],
            "grad_none": [n for n, p in net.named_parameters() if p.grad is None],
            "grad_abs_sum": {n: float(p.grad.double().abs().sum()) for n, p in net.named_parameters()
                             if p.grad is not None},
            "grad_head": {n: p.grad.flatten()[:8].clone() for n, p in net.named_parameters() if p.grad is not None},
            "x_grad_abs_sum": float(x.grad.double().abs().sum()),
            "bn_after": {k: v.clone() for k, v in net.state_dict().items()
                         if k.endswith(("running_mean", "running_var", "num_batches_tracked"))},
        }
        steps[name] = rec
        print("step", name, rec["loss"], rec["components"], len(rec["grad_none"]))
    gold["train_step"] = steps

    # ---- (3) loss / utility known answers ----
    g = torch.Generator().manual_seed(3)
    p = torch.rand(2, 2, 6, 6, 6, generator=g)
    t = (torch.rand(2, 2, 6, 6, 6, generator=g) > 0.5).float()
    gold["dice"] = {"p": p, "t": t, "value": float(UT.dice_loss()(p, t))}
    sph = UT.shape_3d((8, 8, 8), 4, (16, 16, 16), shape="sphere")
    box = UT.shape_3d((8, 8, 8), 4, (16, 16, 16), shape="box")
    sph2 = UT.shape_3d((3, 10, 5), 6, (12, 16, 14), shape="sphere")
    box2 = UT.shape_3d((3, 10, 5), 6, (12, 16, 14), shape="box")
    gold["shape_3d"] = {"sphere_zeros": int((sph == 0).sum()), "box_zeros": int((box == 0).sum()),
                        "dtype": str(sph.dtype),
                        "sphere2": torch.from_numpy(np.packbits(sph2.astype(np.uint8))),
                        "box2": torch.from_numpy(np.packbits(box2.astype(np.uint8)))}
    rng = np.random.RandomState(5)
    img = (rng.rand(16, 20, 24) > 0.7).astype(np.uint8)
    import random
    random.seed(1); np.random.seed(1)
    masked, extracted = TR.random_blank_patch(img.copy(), 1, True, p_type="sphere")
    random.seed(1); np.random.seed(1)
    # replay the reference's draws (transforms.py:243, 252, 268) to record centre and radius
    random.uniform(0, 1)
    pixels = np.argwhere(img > 0)
    center = pixels[np.random.choice(pixels.shape[0])]
    min_r = (np.min(img.shape) // 5) - 1
    max_r = np.max([min_r, np.max(img.shape) // 3.5])
    size_r = np.random.randint(min_r, max_r)
    gold["blank_patch"] = {"img": torch.from_numpy(img), "center": [int(c) for c in center], "size": int(size_r),
                           "masked": torch.from_numpy(masked), "extracted": torch.from_numpy(extracted),
                           "radius_bounds": [int(min_r), int(max_r)],
                           "n_nonzero": int(pixels.shape[0])}
    x5 = torch.rand(2, 2, 4, 4, 4, generator=g)
    x5[0, :, 0, 0, 0] = 0.5  # a tie -> lowest index
    gold["hard_segm"] = {"x": x5, "y": UT.hard_segm_from_tensor(x5), "y4": UT.hard_segm_from_tensor(x5[0])}
    torch.save(gold, os.path.join(OUT, "reference_golden.pt"))
    print("wrote", os.path.join(OUT, "reference_golden.pt"),
          os.path.getsize(os.path.join(OUT, "reference_golden.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
