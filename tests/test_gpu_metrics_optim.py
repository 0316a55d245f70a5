"""GPU: the reporting metrics of the loss handlers (dice_coeff / hausdorff, utilities.py:53-70), the device-side optimizers
and ReduceLROnPlateau (Model.py:369-371, 510-546), SaltAndPepper (transforms.py:13-49) and the val / test branches of the
step driver (Model.py:376-380), each against the CPU oracle / torch.optim."""
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _onehot(lab, c=2):
    return torch.nn.functional.one_hot(lab.long(), c).movedim(-1, 1).float().contiguous()


def _blobs(shape, seed, thr=0.0):
    """Smooth random fields -> blobby masks (surfaces with structure, unlike white noise)."""
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(shape[0], 1, *shape[1:], generator=g)
    k = torch.ones(1, 1, 5, 5, 5) / 125
    for _ in range(2):
        f = torch.nn.functional.conv3d(f, k, padding=2)
    return f[:, 0] * 30 > thr


@pytest.mark.parametrize("shape,c", [((2, 32, 32, 32), 2), ((1, 24, 40, 56), 2), ((2, 16, 16, 16), 3), ((4, 64, 64, 64), 2)])
def test_dice_coeff_and_hausdorff_match_oracle(shape, c):
    from ctunet_b200.utilities import dice_coeff, hausdorff
    from oracle import unet_oracle as O
    g = torch.Generator().manual_seed(sum(shape) + c)
    if c == 2:
        pm, tm = _blobs(shape, 1), _blobs(shape, 2)
        pred = torch.stack((torch.rand(shape, generator=g) * 0.5, torch.where(pm, 0.8, 0.1) * torch.ones(shape)), 1)
        target = _onehot(tm)
    else:
        pred = torch.rand(shape[0], c, *shape[1:], generator=g)
        target = _onehot((torch.rand(shape, generator=g) * c).long().clamp(max=c - 1), c)
    d_ref, h_ref = O.dice_coeff(pred, target), O.hausdorff(pred, target)
    d, h = dice_coeff(pred.to(DEV), target.to(DEV)), hausdorff(pred.to(DEV), target.to(DEV))
    assert d.dtype == torch.float32 and d.dim() == 0 and h.dim() == 0
    assert float(d) == pytest.approx(float(d_ref), rel=1e-6)
    assert float(h) == float(h_ref)                  # integer squared distances, one sqrt: bit-exact


def test_metrics_edge_cases():
    """Empty target (NaN Dice, inf_alt Hausdorff), empty prediction, a single voxel, surfaces touching the border."""
    from ctunet_b200.utilities import dice_coeff, hausdorff
    from oracle import unet_oracle as O
    shp = (1, 12, 10, 14)
    empty = torch.zeros(shp, dtype=torch.bool)
    one = empty.clone()
    one[0, 3, 4, 5] = True
    full = torch.ones(shp, dtype=torch.bool)
    corner = empty.clone()
    corner[0, :4, :3, :5] = True
    for pm, tm in ((one, empty), (empty, one), (one, one), (full, corner), (corner, one), (full, full)):
        pred = torch.stack((torch.full(shp, 0.5), torch.where(pm, 0.9, 0.1)), 1)
        target = _onehot(tm)
        d_ref, h_ref = O.dice_coeff(pred, target), O.hausdorff(pred, target)
        d, h = dice_coeff(pred.to(DEV), target.to(DEV)), hausdorff(pred.to(DEV), target.to(DEV))
        assert (torch.isnan(d_ref) and torch.isnan(d.cpu())) or float(d) == pytest.approx(float(d_ref), rel=1e-6)
        assert float(h) == float(h_ref), (float(h), float(h_ref))
    # batch of two where only one sample has an empty surface: mean of (value, inf_alt)
    pred = torch.stack((torch.full((2,) + shp[1:], 0.5), torch.cat((torch.where(corner, 0.9, 0.1), torch.where(empty, 0.9, 0.1)))), 1)
    target = _onehot(torch.cat((one, one)))
    assert float(hausdorff(pred.to(DEV), target.to(DEV))) == float(O.hausdorff(pred, target))


def test_handlers_append_metrics_like_the_reference():
    """All six stock .ini set b_save_dice_plots (and the UNetSPDO ones b_save_hd_plots): the installed handlers must produce
    the reference's keys (ProblemHandler.py:277-295) instead of raising."""
    import ctunet_b200 as C
    from oracle import unet_oracle as O
    g = torch.Generator().manual_seed(3)
    shp = (2, 16, 16, 16)
    sk_p, fl_p = torch.rand(2, 2, *shp[1:], generator=g), torch.rand(2, 2, *shp[1:], generator=g)
    sk_t, fl_t = _onehot(_blobs(shp, 5)), _onehot(_blobs(shp, 6, 0.5))
    fake = types.SimpleNamespace(params=dict(dice_lambda=1.0, ce_lambda=0.0, save_dice_plots=True, save_hd_plots=True),
                                 losses_and_metrics={}, pt_loss=None)
    pg = (sk_p.to(DEV).requires_grad_(), fl_p.to(DEV).requires_grad_())
    C.FlapRecWithShapePriorDoubleOut.comp_losses_metrics(fake, pg, (sk_t.to(DEV), fl_t.to(DEV)), 0, 1, verbose=False)
    lm = fake.losses_and_metrics
    assert list(lm) == ["dice_loss_sk", "dice_loss_fl", "dice_coef_sk", "dice_coef_fl", "hd_coef_sk", "hd_coef_fl", "epoch_loss"]
    sm = lambda t: torch.softmax(t, 1)
    assert float(lm["dice_coef_sk"][0]) == pytest.approx(float(O.dice_coeff(sm(sk_p), sk_t)), rel=1e-6)
    assert float(lm["dice_coef_fl"][0]) == pytest.approx(float(O.dice_coeff(sm(fl_p), fl_t)), rel=1e-6)
    assert float(lm["hd_coef_sk"][0]) == float(O.hausdorff(sm(sk_p), sk_t))
    assert float(lm["hd_coef_fl"][0]) == float(O.hausdorff(sm(fl_p), fl_t))
    # Model.update_plots_tensorboard_avg (Model.py:396-401): sum(list) / len(list), then float()
    for k, v in lm.items():
        float(sum(v) / len(v))
    fake.pt_loss.backward()
    # single-output handler: dice_coef only (ProblemHandler.py:84-88)
    fake = types.SimpleNamespace(params=dict(dice_lambda=1.0, ce_lambda=1.0, save_dice_plots=True), losses_and_metrics={}, pt_loss=None)
    C.ProblemHandler.comp_losses_metrics(fake, sm(sk_p).to(DEV), sk_t.to(DEV), 0, 1, verbose=False)
    assert list(fake.losses_and_metrics) == ["ce", "dice_loss", "dice_coef", "epoch_loss"]
    assert float(fake.losses_and_metrics["dice_coef"][0]) == pytest.approx(float(O.dice_coeff(sm(sk_p), sk_t)), rel=1e-6)


# ------------------------------------------------------------------------------------------------ optimizers
def _torch_opt(kind, params, lr, wd, mom):
    if kind == "adam":
        return torch.optim.Adam(params, lr=lr, weight_decay=wd, amsgrad=True)
    if kind == "adamw":
        return torch.optim.AdamW(params, lr=lr, weight_decay=wd, amsgrad=True)
    if kind == "rmsprop":
        return torch.optim.RMSprop(params, lr=lr, weight_decay=wd, momentum=mom)
    return torch.optim.SGD(params, lr=lr, momentum=mom, weight_decay=wd)


@pytest.mark.parametrize("kind,wd,mom", [("adam", 0.0, 0.0), ("adam", 1e-2, 0.0), ("adamw", 1e-2, 0.0), ("rmsprop", 0.0, 0.99),
                                         ("rmsprop", 1e-3, 0.0), ("sgd", 0.0, 0.99), ("sgd", 1e-3, 0.0)])
def test_flat_optimizer_matches_torch_optim(kind, wd, mom):
    """Five steps of the one-launch update against torch.optim on the CPU (single-tensor fp32 implementation) with the
    arguments of Model.py:514-541; the count of entries that are not BIT-identical is reported."""
    from ctunet_b200.optim import FlatOptimizer
    g = torch.Generator().manual_seed(11)
    shapes = [(7, 2, 3, 3, 3), (7,), (14, 7, 3, 3, 3), (3, 14, 1, 1, 1), (1500,)]
    ref = [torch.randn(s, generator=g).requires_grad_() for s in shapes]
    mine = [r.detach().clone().to(DEV) for r in ref]
    n = sum(r.numel() for r in ref)
    flat = torch.zeros(n + 8, device=DEV)
    opt_r = _torch_opt(kind, ref, 1e-3, wd, mom)
    opt = FlatOptimizer(mine, flat, kind=kind, lr=1e-3, weight_decay=wd, momentum=mom)
    for it in range(5):
        grads = [torch.randn(s, generator=g) * (0.1 + it) for s in shapes]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        opt_r.step()
        flat[:n].copy_(torch.cat([gr.flatten() for gr in grads]))
        opt.step()
    torch.cuda.synchronize()
    assert int(opt.step_count) == 5
    differing = 0
    for r, m in zip(ref, mine):
        d = (m.cpu() - r.detach()).abs()
        differing += int((d != 0).sum())
        assert float(d.max()) <= 4e-7 * max(1.0, float(r.detach().abs().max())), (kind, float(d.max()))
    print("%s: %d of %d entries not bit-identical to torch.optim after 5 steps" % (kind, differing, n))


def test_plateau_scheduler_matches_torch_reduce_lr_on_plateau():
    """60 iterations of a loss sequence with plateaus, improvements below the relative threshold and a NaN-free noise
    floor: the device-side rule reproduces torch's LR trajectory exactly (Model.py:369-371 steps it every iteration)."""
    from ctunet_b200.optim import FlatOptimizer
    rng = np.random.RandomState(0)
    losses = np.concatenate([np.linspace(2.0, 1.0, 8), 1.0 - 1e-6 * np.arange(15), 0.9 + 0.01 * rng.rand(20), np.full(17, 0.5)])
    p = torch.zeros(4, requires_grad=True)
    opt_r = torch.optim.Adam([p], lr=1e-4, amsgrad=True)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt_r)
    pm = torch.zeros(4, device=DEV)
    flat = torch.zeros(12, device=DEV)
    opt = FlatOptimizer([pm], flat, kind="adam", lr=1e-4, plateau=True)
    ref_lr, got_lr = [], []
    for v in losses:
        lt = torch.tensor([v], dtype=torch.float32)
        p.grad = torch.ones(4)
        opt_r.step()
        sched.step(lt[0])
        ref_lr.append(opt_r.param_groups[0]["lr"])
        opt.step(loss=lt.to(DEV))
        got_lr.append(opt.lr)
    assert got_lr == ref_lr
    assert len(set(ref_lr)) >= 3                      # the sequence really drives several reductions


def test_train_step_scheduler_and_optimizers_run_captured():
    """TrainStep(scheduler=...) inside the captured step: the LR trajectory equals torch's ReduceLROnPlateau fed the
    losses the step itself reported; RMSprop / SGD / AdamW steps run and lower the loss."""
    import ctunet_b200 as C
    from ctunet_b200.synthetic import make_training_batch
    from ctunet_b200.trainer import TrainStep
    img, (sk_t, fl_t) = make_training_batch(2, 2, 32, seed=7, device=DEV)
    torch.manual_seed(0)
    net = C.UNetSP().to(DEV)
    step = TrainStep(net, "double", 1.0, 1.0, lr=3e-2, scheduler=dict(patience=2), graph=True)
    p = torch.zeros(1, requires_grad=True)
    opt_r = torch.optim.Adam([p], lr=3e-2)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt_r, patience=2)
    lrs, ref = [], []
    for it in range(24):
        comps = step(img, (sk_t, fl_t))
        total = comps[-1:].clone().cpu()
        sched.step(total[0])
        ref.append(opt_r.param_groups[0]["lr"])
        lrs.append(step.lr)
    assert step._graph is not None
    assert lrs == ref
    for kind in ("rmsprop", "sgd", "adamw"):
        torch.manual_seed(0)
        net = C.UNetSP().to(DEV)
        st = TrainStep(net, "double", 1.0, 1.0, lr=1e-4 if kind != "sgd" else 1e-3, optimizer=kind, momentum=0.9, graph=True)
        vals = [float(st(img, (sk_t, fl_t))[-1]) for _ in range(6)]
        assert vals[-1] < vals[0], (kind, vals)


# ------------------------------------------------------------------------------------------------ SaltAndPepper
def test_salt_and_pepper_logic_is_bit_exact_and_density_decays():
    from ctunet_b200.utilities import SaltAndPepper
    from oracle import unet_oracle as O
    img = (torch.rand(2, 12, 16, 20, generator=torch.Generator().manual_seed(0)) > 0.5).float()
    rs = np.random.RandomState(3)
    ub, uw = rs.uniform(0, 1, img.shape), rs.uniform(0, 1, img.shape)
    np.random.seed(8)
    random.seed(8)
    sp = SaltAndPepper(p=0.6, noise_density=0.4)
    got = sp({"image": img.to(DEV), "target": img.to(DEV)}, fields=(torch.from_numpy(ub).to(DEV), torch.from_numpy(uw).to(DEV)))
    # the reference's draw order (transforms.py:31-34): density first, then one gate per image
    np.random.seed(8)
    random.seed(8)
    nd = np.random.uniform(0, 0.4)
    exp = img.numpy().astype(np.uint8)
    for i in range(2):
        if 0.6 >= random.uniform(0, 1):
            exp[i] = O.salt_and_pepper(exp[i], nd, 0.1, ub[i], uw[i])
    assert sp.noise_density == nd
    assert got["image"].dtype == torch.float32 and torch.equal(got["image"].cpu(), torch.from_numpy(exp).float())
    assert torch.equal(got["target"].cpu(), img)                    # apply_to = (True, False)
    # generator mode: flip rates follow the thresholds; successive calls decay the density (the reference's quirk)
    sp = SaltAndPepper(p=1, noise_density=0.3)
    vol = torch.zeros(64, 64, 64, device=DEV)
    out = sp({"image": vol, "target": vol})["image"]
    d1 = sp.noise_density
    assert abs(float(out.mean()) - d1 * 0.1) < 3e-3                 # salt on an empty volume: density * salt_ratio
    vol1 = torch.ones(64, 64, 64, device=DEV)
    out = sp({"image": vol1, "target": vol1})["image"]
    d2 = sp.noise_density
    assert d2 <= d1
    assert abs((1 - float(out.mean())) - d2 * 0.9 * (1 - d2 * 0.1)) < 3e-3
    a = sp({"image": vol1, "target": vol1})["image"]
    b = sp({"image": vol1, "target": vol1})["image"]
    assert not torch.equal(a, b) or sp.noise_density < 1e-4         # fresh Philox key per call


# ------------------------------------------------------------------------------------------------ val / test branches
@pytest.mark.parametrize("graph", [False, True])
def test_eval_step_matches_oracle(graph):
    import ctunet_b200 as C
    from ctunet_b200.trainer import EvalStep
    from oracle import unet_oracle as O
    cfg = O.PRESETS["UNetSP"]
    x, (sk_t, fl_t) = O.make_training_batch(2, 2, 32, seed=21)
    sd = O.build_state_dict(cfg, seed=0)
    with torch.no_grad():
        ref = O.unet_forward(sd, x, cfg, training=False)
        _, comps = O.loss_double_output(ref, (sk_t, fl_t), 1.0, 1.0)
    C.set_compute_dtype("fp32")
    torch.manual_seed(0)
    net = C.UNetSP().to(DEV).train()
    C.set_compute_dtype("bf16")
    ev = EvalStep(net, "double", 1.0, 1.0, metrics=True, graph=graph)
    xg, tg = x.to(DEV), (sk_t.to(DEV), fl_t.to(DEV))
    for _ in range(2):
        vals = dict(zip(ev.keys, ev(xg, tg).tolist()))
    assert net.training                                            # the driver restores the mode
    assert int(net.state_dict()["d_blocks.0.block.1.num_batches_tracked"]) == 0      # eval: no buffer moves
    for k, v in comps.items():
        assert vals[k] == pytest.approx(float(v), rel=1e-4), k
    assert vals["dice_coef_sk"] == pytest.approx(float(O.dice_coeff(ref[0], sk_t)), abs=2e-3)
    assert vals["hd_coef_fl"] == pytest.approx(float(O.hausdorff(ref[1], fl_t)), abs=1.5)
    labs = ev.labels(xg)
    for lab, r in zip(labs, ref):
        assert lab.dtype == torch.float32 and lab.shape == (2, 32, 32, 32)
        agree = (lab.cpu() == O.hard_segm_from_tensor(r)).float().mean().item()
        assert agree > 0.9999


# ------------------------------------------------------------------------------------------------ fused head + loss
@pytest.mark.parametrize("name,handler,mode,lams", [("UNetSP", "double", "fp32", (1.0, 1.0)), ("UNetSP", "double", "bf16", (1.0, 0.5)),
                                                    ("UNetSPSmall", "double", "fp32", (1.0, 1.0)), ("UNetDO", "double", "fp32", (1.0, 0.0)),
                                                    ("recAE_v2_fixed", "single", "fp32", (1.0, 1.0)),
                                                    ("UNet4_2IC", "single", "bf16", (0.0, 2.0))])
def test_fused_head_loss_equals_separate_kernels(name, handler, mode, lams):
    """trainer.FUSED_HEAD_LOSS: head + Dice/CE in one pass (csrc/head.cu head_loss_*) against the separate head, loss and
    head-backward kernels on the same network: identical loss components, identical flat gradient buffer."""
    import ctunet_b200 as C
    import ctunet_b200.trainer as TR
    from oracle import unet_oracle as O
    cfg = O.PRESETS[name]
    size = 32
    x, (sk_t, fl_t) = O.make_training_batch(2, cfg.input_channels, size, seed=13)
    target = (sk_t.to(DEV), fl_t.to(DEV)) if handler == "double" else sk_t.to(DEV)
    res = []
    for fused in (False, True):
        C.set_compute_dtype(mode)
        torch.manual_seed(0)
        net = getattr(C, name)().to(DEV)
        C.set_compute_dtype("bf16")
        step = TR.TrainStep(net, handler, dice_lambda=lams[0], ce_lambda=lams[1], lr=1e-4)
        old = TR.FUSED_HEAD_LOSS
        TR.FUSED_HEAD_LOSS = fused
        try:
            comps = step._forward_backward(x.to(DEV), target).clone()
        finally:
            TR.FUSED_HEAD_LOSS = old
        torch.cuda.synchronize()
        res.append((comps.cpu(), step.grads.flat[:step.grads.n_grad].clone().cpu(), step.keys))
    (c0, g0, k0), (c1, g1, k1) = res
    assert k0 == k1 and len(c0) == len(k0)
    assert torch.allclose(c0, c1, rtol=1e-6, atol=1e-7), (c0, c1)
    scale = float(g0.abs().max())
    assert scale > 0
    # same arithmetic per voxel; only the summation order of the parameter-gradient partials differs
    assert float((g0 - g1).abs().max()) <= (2e-5 if mode == "fp32" else 2e-3) * scale
    rel = float((g0 - g1).norm() / g0.norm())
    assert rel < (1e-5 if mode == "fp32" else 2e-3), rel


# ------------------------------------------------------------------------------------------------ peer-memory exchange
@pytest.mark.parametrize("graph", [False, True])
def test_peer_exchange_single_rank_equals_plain_step(graph):
    """parallel.PeerGradSync (csrc/peer.cu: all-reduce + optimizer in one kernel over IPC-mapped gradient buffers) with a
    world of one rank -- the flag protocol, the rank-ordered mean, the averaged loss tail and the fused update -- against
    the plain single-GPU step.  (Two and eight ranks: scripts/check_peer_ddp.py under torchrun on the GPU box.)"""
    import ctunet_b200 as C
    from ctunet_b200.parallel import PeerGradSync
    from ctunet_b200.synthetic import make_training_batch
    from ctunet_b200.trainer import TrainStep
    img, (sk_t, fl_t) = make_training_batch(2, 2, 32, seed=5, device=DEV)
    nets, steps = [], []
    for peer in (False, True):
        C.set_compute_dtype("fp32")
        torch.manual_seed(0)
        net = C.UNetSP().to(DEV)
        C.set_compute_dtype("bf16")
        nets.append(net)
        steps.append(TrainStep(net, "double", 1.0, 1.0, lr=1e-3, scheduler=True, grad_sync=PeerGradSync(net) if peer else None,
                               graph=graph))
    assert steps[1].peer and not steps[1].split_graph
    for it in range(5):
        a = steps[0](img, (sk_t, fl_t)).tolist()
        b = steps[1](img, (sk_t, fl_t)).tolist()
        assert a == pytest.approx(b, rel=2e-4, abs=2e-4), it
    assert steps[1].grads.error() == 0
    assert int(steps[1].optimizer.step_count) == 5
    for (k, p), (_, q) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        assert float((p - q).abs().max()) <= 2e-3 * 5 + 1e-6, k          # Adam: rounding-level gradient noise -> <= lr per step
    steps[1].grads.close()


@pytest.mark.parametrize("name,handler", [("UNetSP", "double"), ("recAE_v2_fixed", "single")])
def test_uint8_mask_targets_equal_one_hot_float_targets(name, handler):
    """The fused head + loss kernels accept the label masks themselves (uint8 [B,D,H,W], datasets.py:209-214 before the
    one-hot encoding): bit-identical loss components and gradients to the float one-hot targets."""
    import ctunet_b200 as C
    import ctunet_b200.trainer as TR
    from oracle import unet_oracle as O
    cfg = O.PRESETS[name]
    x, (sk_t, fl_t) = O.make_training_batch(2, cfg.input_channels, 32, seed=17)
    onehot = (sk_t.to(DEV), fl_t.to(DEV)) if handler == "double" else sk_t.to(DEV)
    masks = tuple(t[:, 1].to(torch.uint8).contiguous().to(DEV) for t in (sk_t, fl_t))
    masks = masks if handler == "double" else masks[0]
    res = []
    for tgt in (onehot, masks):
        torch.manual_seed(0)
        net = getattr(C, name)().to(DEV)
        step = TR.TrainStep(net, handler, 1.0, 1.0, lr=1e-4)
        comps = step._forward_backward(x.to(DEV), tgt).clone()
        torch.cuda.synchronize()
        res.append((comps.cpu(), step.grads.flat[:step.grads.n_grad].clone().cpu()))
    assert torch.equal(res[0][0], res[1][0])
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-6 * float(res[0][1].abs().max())   # atomics order only
