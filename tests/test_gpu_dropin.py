"""GPU: the drop-in under the UNMODIFIED reference trainer.  ``ctunet.pytorch.Model.forward_pass`` (Model.py:324-380) from
``oracle/_ref`` (the pip-installed reference, oracle/build_ref.py) runs with ``ctunet_b200.install()`` applied: its
``eval(model_class)`` / ``eval(problem_handler)`` resolve to the B200 modules, its own torch.optim.Adam(amsgrad) and
ReduceLROnPlateau drive them.  Also: nn.DataParallel replicas (Model.py:486) and weight-plan invalidation."""
import pytest
import torch

from oracle.reference_loader import reference_available

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _restore_after(fn):
    """install() rebinds names inside the live reference modules: undo it so other tests see the stock reference."""
    import ctunet_b200
    from oracle.reference_loader import load_reference
    MM = load_reference(with_trainer=True)[4]
    anomaly = torch.is_anomaly_enabled()
    try:
        return fn(MM)
    finally:
        ctunet_b200.uninstall()
        torch.autograd.set_detect_anomaly(anomaly)
        torch.set_grad_enabled(True)


@pytest.mark.skipif(not reference_available(), reason="reference not present (oracle/_ref is built by oracle/build_ref.py)")
@pytest.mark.parametrize("example", ["autoimplant_FlapRecSP2O", "FlapRecSP2O_128", "AutoImplant2020_wShapePrior"])
def test_reference_forward_pass_with_dropin_installed(example):
    """Three training batches + one validation batch through the reference's own forward_pass with a stock example's
    hyper-parameters (metrics on, scheduler on) -- first-batch losses equal the fp32 oracle's, the loss goes down, BN buffers
    move twice per step (the reentrant-checkpoint quirk), the checkpoint is loadable by the stock reference class."""
    import ctunet_b200
    from oracle import unet_oracle as O
    from oracle.ref_harness import ListLoader, example_params, make_trainer

    def run(MM):
        params = example_params(example)
        assert params["save_dice_plots"] is True
        params.setdefault("save_hd_plots", False)             # quirk 7: the key is read unconditionally
        params["learning_rate"] = 1e-3
        params["resume_model"] = ""                           # (the stock files point at the author's private checkpoints)
        name = params["model_class"]
        cfg = O.PRESETS[name]
        double = cfg.head != "plain"
        x, (sk_t, fl_t) = O.make_training_batch(2, cfg.input_channels, 32, seed=31)
        sample = {"image": x, "target": [sk_t, fl_t] if double else sk_t}
        ctunet_b200.install(MM)
        ctunet_b200.set_compute_dtype("bf16")
        torch.manual_seed(0)
        m = make_trainer(params, DEV)
        m.initialize_models()
        assert type(m.models["main"]).__module__ == "ctunet_b200.models" or isinstance(m.models["main"], torch.nn.DataParallel)
        m.initialize_optimizer()
        m.forward_pass("train", ListLoader([sample] * 3))
        train_lm = {k: list(v) for k, v in m.losses_and_metrics.items()}
        m.forward_pass("val", ListLoader([sample]))
        return m, train_lm, x, sk_t, fl_t, cfg, double

    m, lm, x, sk_t, fl_t, cfg, double = _restore_after(run)
    sd = O.build_state_dict(cfg, seed=0)
    out = O.unet_forward(sd, x.clone().requires_grad_(), cfg, training=True)
    _, comps = (O.loss_double_output(out, (sk_t, fl_t), m.params["dice_lambda"], m.params["ce_lambda"]) if double else
                O.loss_single_output(out, sk_t, m.params["dice_lambda"], m.params["ce_lambda"]))
    for k, v in comps.items():
        assert abs(lm[k][0] - float(v)) < 5e-3, (k, lm[k][0], float(v))
    assert len(lm["epoch_loss"]) == 3 and lm["epoch_loss"][2] < lm["epoch_loss"][0]
    metric_keys = [k for k in lm if k.startswith(("dice_coef", "hd_coef"))]
    assert metric_keys and all(len(lm[k]) == 3 for k in metric_keys)
    assert len(m.losses_and_metrics["epoch_loss"]) == 4                     # + the validation batch
    net = m.models["main"]
    net = net.module if isinstance(net, torch.nn.DataParallel) else net
    first_bn = "d_blocks.0.block.1" if cfg.family == "generic" else "dblock1.1"
    assert int(net.state_dict()[first_bn + ".num_batches_tracked"]) == 6    # 3 steps x 2 (eval adds none)
    if "scheduler" in m.params:                                             # Model.py:544-546: created iff the key exists
        assert type(m.params["scheduler"]).__name__ == "ReduceLROnPlateau"
    # the reference's own class loads what the drop-in trained (bare state_dict, Model.py:282)
    from oracle.reference_loader import load_reference
    ref_cls = getattr(load_reference()[0], m.params["model_class"])
    ref_cls().load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})


@pytest.mark.skipif(not reference_available(), reason="reference not present (oracle/_ref is built by oracle/build_ref.py)")
@pytest.mark.parametrize("example", ["autoimplant_FlapRecSP2O", "AutoImplant2020_wShapePrior"])
def test_graph_dropin_matches_eager_dropin(example):
    """``install(graph=True)``: the autograd node replays captured forward / backward graphs from the third iteration on.
    Six training batches through the reference's own forward_pass give the loss trajectory, parameters and BatchNorm buffers of
    the eager drop-in (same kernels, same order; fp32 atomics in the weight gradients are the only non-determinism)."""
    import ctunet_b200
    from oracle import unet_oracle as O
    from oracle.ref_harness import ListLoader, example_params, make_trainer

    def run_with(graph):
        def run(MM):
            params = example_params(example)
            params["save_dice_plots"] = False
            params["save_hd_plots"] = False
            params["learning_rate"] = 1e-3
            params["resume_model"] = ""
            cfg = O.PRESETS[params["model_class"]]
            double = cfg.head != "plain"
            samples = []
            for i in range(6):
                x, (sk_t, fl_t) = O.make_training_batch(2, cfg.input_channels, 32, seed=40 + i)
                samples.append({"image": x, "target": [sk_t, fl_t] if double else sk_t})
            ctunet_b200.install(MM, graph=graph)
            ctunet_b200.set_compute_dtype("bf16")
            torch.manual_seed(0)
            m = make_trainer(params, DEV)
            m.initialize_models()
            m.initialize_optimizer()
            torch.autograd.set_detect_anomaly(False)
            m.forward_pass("train", ListLoader(samples))
            losses = list(m.losses_and_metrics["epoch_loss"])
            net = m.models["main"]
            net = net.module if isinstance(net, torch.nn.DataParallel) else net
            used = bool(getattr(net, "_dropin_graphs", None)) and any(st.bwd is not None for st in net._dropin_graphs.values())
            return losses, {k: v.detach().float().cpu().clone() for k, v in net.state_dict().items()}, used
        return _restore_after(run)

    l_eager, sd_eager, used_eager = run_with(False)
    l_graph, sd_graph, used_graph = run_with(True)
    assert used_graph and not used_eager
    assert len(l_graph) == 6
    for a, b in zip(l_eager, l_graph):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (l_eager, l_graph)
    # Adam normalises every gradient entry by its own running magnitude: an entry whose gradient is noise (the order of the
    # fp32 atomics differs from run to run) still moves by lr per step in a direction that noise decides, so single entries
    # may differ by a good part of the 6 x lr they can travel -- the model as a whole may not
    def normwise(pred):
        keys = [k for k in sd_eager if pred(k)]
        num = sum(float((sd_graph[k] - sd_eager[k]).pow(2).sum()) for k in keys)
        den = sum(float(sd_eager[k].pow(2).sum()) for k in keys)
        return (num / den) ** 0.5

    is_buf = lambda k: k.endswith(("running_mean", "running_var"))
    e_par = normwise(lambda k: not is_buf(k) and not k.endswith("num_batches_tracked"))
    e_buf = normwise(is_buf)
    print("graph vs eager drop-in after 6 steps: parameters %.2e, BatchNorm buffers %.2e (normwise)" % (e_par, e_buf))
    assert e_par <= 1e-2 and e_buf <= 3e-2, (e_par, e_buf)
    for k, v in sd_eager.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd_graph[k]) == int(v) and int(v) in (6, 12), k          # 6 steps x 2 (reentrant-checkpoint quirk; the dead center block: x 1)

def test_data_parallel_replica_trains():
    """What nn.DataParallel does per step (Model.py:486): ``replicate`` the module, run the replica.  A replica has no
    parameters of its own (``parameters()`` is empty); the gradients must still reach the real parameters through the
    broadcast copies -- identical to the plain module's gradients."""
    import ctunet_b200 as C
    from torch.nn.parallel import replicate
    from ctunet_b200.synthetic import make_training_batch
    C.set_compute_dtype("fp32")
    torch.manual_seed(0)
    net = C.UNetSP().to(DEV).train()
    C.set_compute_dtype("bf16")
    img, (sk_t, fl_t) = make_training_batch(2, 2, 16, seed=3, device=DEV)

    def grads(module):
        for p in net.parameters():
            p.grad = None
        sk, fl = module(img.clone().requires_grad_())
        (sk * sk_t).sum().add((fl * fl_t).sum()).backward()
        return {n: (p.grad.clone() if p.grad is not None else None) for n, p in net.named_parameters()}

    plain = grads(net)
    rep = replicate(net, [torch.cuda.current_device()])[0]
    assert getattr(rep, "_is_replica", False) and not list(rep.parameters())
    via_replica = grads(rep)
    assert any(v is not None for v in via_replica.values())
    for n, g in plain.items():
        if g is None:
            # Broadcast's backward hands an unused input zeros rather than None (so does the reference under DataParallel:
            # the dead center block then sees an all-zero gradient and Adam leaves it where it is)
            assert via_replica[n] is None or not bool(via_replica[n].any()), n
        else:
            assert via_replica[n] is not None, n
            assert torch.allclose(via_replica[n], g, rtol=1e-4, atol=1e-6 + 1e-4 * float(g.abs().max())), n
    dp = torch.nn.DataParallel(net, device_ids=[torch.cuda.current_device()])
    sk, fl = dp(img)
    assert sk.shape == (2, 2, 16, 16, 16)


def test_weight_plan_follows_replaced_parameters():
    """The cached weight-preparation plan (and the captured inference graph) must not keep serving a parameter OBJECT that
    was replaced after the first forward (load_state_dict(assign=True), a swapped head for fine-tuning)."""
    import ctunet_b200 as C
    from ctunet_b200 import preprocess as P
    torch.manual_seed(0)
    net = C.UNetSP().to(DEV).eval()
    x = (torch.rand(1, 2, 32, 32, 32, generator=torch.Generator().manual_seed(2)) > 0.7).float().to(DEV)
    with torch.no_grad():
        a0 = net(x)[0].clone()
        net(x)                                                          # second pass: served from the recorded plan
        torch.manual_seed(1)
        other = C.UNetSP().to(DEV).eval()
        b_ref = other(x)[0].clone()
        net.load_state_dict(other.state_dict(), assign=True)           # every parameter / buffer object is replaced
        b = net(x)[0]
    assert not torch.allclose(a0, b_ref)
    assert torch.equal(b, b_ref)
    vol = x[0]
    l0 = P.sliding_window_argmax(net, vol, patch=32, batch=1)[0].clone()
    torch.manual_seed(0)
    again = C.UNetSP().to(DEV).eval()
    net.load_state_dict(again.state_dict(), assign=True)
    l1 = P.sliding_window_argmax(net, vol, patch=32, batch=1)[0]
    l_ref = P.sliding_window_argmax(again, vol, patch=32, batch=1)[0]
    assert torch.equal(l1, l_ref)
