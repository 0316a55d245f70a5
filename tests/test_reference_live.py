"""CPU: the restated oracle and the test harness against the LIVE reference (``/root/reference`` in the build container,
``oracle/_ref`` -- the pip-installed copy, oracle/build_ref.py -- anywhere else).  Skipped when neither exists."""
import pytest
import torch

from oracle import unet_oracle as O
from oracle.reference_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present (run oracle/build_ref.py)")


def _batch(cin, size, seed, batch=1):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(batch, cin, size, size, size, generator=g) > 0.7).float()
    sk = (torch.rand(batch, size, size, size, generator=g) > 0.6).long()
    fl = ((torch.rand(batch, size, size, size, generator=g) > 0.8) & (sk > 0)).long()
    oh = lambda t: torch.nn.functional.one_hot(t, 2).permute(0, 4, 1, 2, 3).float().contiguous()
    return x, oh(sk), oh(fl)


def test_metrics_oracle_equals_reference_functions():
    from oracle.reference_loader import load_reference
    UT = load_reference()[2]
    for seed, shape in ((0, (2, 2, 12, 16, 20)), (1, (1, 2, 24, 8, 8))):
        g = torch.Generator().manual_seed(seed)
        p = torch.rand(*shape, generator=g)
        t = torch.nn.functional.one_hot((torch.rand(shape[0], *shape[2:], generator=g) > 0.5).long(), 2).movedim(4, 1).float()
        assert torch.equal(O.dice_coeff(p, t), UT.dice_coeff(p, t))
        assert torch.equal(O.hausdorff(p, t), UT.hausdorff(p, t))
    # an empty target class: NaN Dice (propagates through the mean), inf_alt Hausdorff
    t0 = torch.zeros(1, 2, 8, 8, 8)
    t0[:, 0] = 1
    p = torch.rand(1, 2, 8, 8, 8, generator=torch.Generator().manual_seed(3))
    assert torch.isnan(UT.dice_coeff(p, t0)) and torch.isnan(O.dice_coeff(p, t0))
    assert float(UT.hausdorff(p, t0)) == 8.0 == float(O.hausdorff(p, t0))


def test_salt_and_pepper_oracle_equals_reference_class():
    import random
    import numpy as np
    from oracle.reference_loader import load_reference
    TR = load_reference()[3]
    img = (torch.rand(2, 6, 7, 8, generator=torch.Generator().manual_seed(0)) > 0.5).float()
    np.random.seed(5)
    random.seed(5)
    sp = TR.SaltAndPepper(p=0.7, noise_density=0.3)
    ref = sp({"image": img.clone(), "target": img.clone()})["image"]
    # replay the draws in the reference's order (transforms.py:31-41)
    np.random.seed(5)
    random.seed(5)
    nd = np.random.uniform(0, 0.3)
    out = img.numpy().astype(np.uint8)
    for i in range(2):
        if 0.7 >= random.uniform(0, 1):
            ub = np.random.uniform(0, 1, out[i].shape)
            uw = np.random.uniform(0, 1, out[i].shape)
            out[i] = O.salt_and_pepper(out[i], nd, 0.1, ub, uw)
    assert sp.noise_density == nd                      # the self-decaying density quirk (transforms.py:31)
    assert torch.equal(ref, torch.from_numpy(out).float())


def test_reference_forward_pass_runs_through_the_harness_on_cpu():
    """The harness used by the GPU drop-in test and the bench reference arm: the reference's own forward_pass,
    model, handler, Adam(amsgrad) and ReduceLROnPlateau on CPU -- and the oracle reproduces its first loss."""
    from oracle.ref_harness import ListLoader, make_trainer
    x, sk, fl = _batch(2, 16, 3, batch=2)      # batch 2: the (dead) center block sees 1^3 voxels per sample
    params = dict(model_class="UNetSP", problem_handler="FlapRecWithShapePriorDoubleOut", optimizer="adam",
                  learning_rate=1e-3, momentum=0.99, weight_decay=0.0, dice_lambda=1.0, ce_lambda=1.0,
                  save_dice_plots=True, save_hd_plots=True, scheduler=True)
    torch.manual_seed(0)
    m = make_trainer(params, "cpu")
    anomaly = torch.is_anomaly_enabled()
    torch.autograd.set_detect_anomaly(False)
    try:
        m.initialize_models()
        m.initialize_optimizer()
        m.forward_pass("train", ListLoader([{"image": x, "target": [sk, fl]}] * 2))
        m.forward_pass("val", ListLoader([{"image": x, "target": [sk, fl]}]))
    finally:
        torch.autograd.set_detect_anomaly(anomaly)
        torch.set_grad_enabled(True)
    lm = m.losses_and_metrics
    assert set(lm) == {"ce_sk", "ce_fl", "dice_loss_sk", "dice_loss_fl", "dice_coef_sk", "dice_coef_fl", "hd_coef_sk",
                       "hd_coef_fl", "epoch_loss"}
    assert len(lm["epoch_loss"]) == 3
    sd = O.build_state_dict(O.PRESETS["UNetSP"], seed=0)
    out = O.unet_forward(sd, x, O.PRESETS["UNetSP"], training=True)
    loss, _ = O.loss_double_output(out, (sk, fl), 1.0, 1.0)
    assert float(loss) == pytest.approx(lm["epoch_loss"][0], rel=1e-6)
    assert type(m.params["scheduler"]).__name__ == "ReduceLROnPlateau"


def test_install_rebinds_the_names_the_reference_evals():
    import ctunet_b200
    from oracle.reference_loader import load_reference
    MD, PH, UT, TR, MM = load_reference(with_trainer=True)
    ref_hard = UT.hard_segm_from_tensor
    try:
        done = ctunet_b200.install(MM)
        assert MM.UNetSP is ctunet_b200.UNetSP and eval("UNetSP", vars(MM)) is ctunet_b200.UNetSP
        assert "utils.dice_coeff" in done and MM.utils.dice_coeff is ctunet_b200.utilities.dice_coeff
        h = eval("FlapRecWithShapePriorDoubleOut", vars(MM))()
        assert h.comp_losses_metrics.__module__ == "ctunet_b200.losses"
        assert MM.utils.hard_segm_from_tensor is ref_hard                        # deliberately NOT rebound
    finally:
        ctunet_b200.uninstall()
    assert MM.UNetSP is MD.UNetSP and UT.dice_coeff.__module__ == "ctunet.utilities"
    assert eval("FlapRecWithShapePriorDoubleOut", vars(MM))().comp_losses_metrics.__module__ == "ctunet.pytorch.ProblemHandler"


def test_stock_example_inis_parse_and_name_installed_classes():
    import ctunet_b200
    from oracle.ref_harness import EXAMPLES, example_params
    for name in EXAMPLES:
        p = example_params(name)
        assert p["model_class"] in ctunet_b200.MODEL_CLASSES
        assert hasattr(ctunet_b200, p["problem_handler"])
        assert p.get("save_dice_plots") is True          # every stock example switches the metrics on
