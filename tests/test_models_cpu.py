"""CPU: host-side logic of the drop-in surface -- constructors, state_dict layout, same-seed
initialisation, loud failure without CUDA, and the name-rebinding plug-in used by the reference's
``eval(model_class)`` lookup (ctunet/pytorch/Model.py:101,485,488)."""
import types

import pytest
import torch

CLASSES = ["UNet", "UNetSP", "UNetSPSmall", "UNetDO", "UNet4_2IC", "recAE_v2_fixed"]


@pytest.mark.parametrize("name", CLASSES)
def test_state_dict_layout_and_same_seed_init(golden, name):
    import ctunet_b200 as C
    g = golden["classes"][name]
    torch.manual_seed(0)
    net = getattr(C, name)()
    sd = net.state_dict()
    assert list(sd.keys()) == g["keys"]
    assert [tuple(v.shape) for v in sd.values()] == g["shapes"]
    assert len(sd) == g["n_state_entries"]
    params = list(net.parameters())
    assert sum(p.numel() for p in params) == g["n_params"]
    assert abs(float(sum(p.detach().double().abs().sum() for p in params)) - g["abs_sum"]) < 1e-9
    assert torch.equal(params[0].detach().flatten()[:3], g["first3"])
    assert all(v.dtype == torch.float32 for k, v in sd.items() if not k.endswith("num_batches_tracked"))
    assert sd[[k for k in sd if k.endswith("num_batches_tracked")][0]].dtype == torch.int64


def test_reference_attributes_present():
    import ctunet_b200 as C
    net = C.UNet()
    for attr in ["chk", "skip", "apply_softmax", "apply_sigmoid", "fc_layer", "cat", "mp", "d_blocks", "cblock",
                 "u_blocks", "last_conv"]:
        assert hasattr(net, attr), attr
    assert net.chk is True and net.apply_sigmoid is True and net.apply_softmax is False
    assert len(net.d_blocks) == 4 and len(net.u_blocks) == 4
    leg = C.recAE_v2_fixed()
    for attr in ["chk", "mp", "dblock1", "dblock4", "cblock_center", "ublock1", "ublock4", "last_conv"]:
        assert hasattr(leg, attr), attr
    custom = C.UNet(input_channels=3, out_channels=4, n_blocks=2, i_size=5)
    assert custom.last_conv.weight.shape == (4, 10, 1, 1, 1)


def test_unsupported_variants_raise():
    import ctunet_b200 as C
    with pytest.raises(NotImplementedError):
        C.UNet(residual=True)
    with pytest.raises(NotImplementedError):
        C.UNet(fc_layer=[8, 4])
    with pytest.raises(NotImplementedError):
        C.UNet(cat=False)
    with pytest.raises(NotImplementedError):
        C.UNet(kern_sz_conv=3, padding=0)
    with pytest.raises(NotImplementedError):
        C.UNetBlock(2, 4).forward(torch.zeros(1))


def test_no_cpu_fallback():
    import ctunet_b200 as C
    net = C.UNetSP()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 2, 16, 16, 16))
    with pytest.raises(RuntimeError):
        C.hard_segm_from_tensor(torch.zeros(1, 2, 4, 4, 4))
    with pytest.raises(RuntimeError):
        C.dice_loss()(torch.zeros(1, 2, 4), torch.zeros(1, 2, 4))
    with pytest.raises(RuntimeError):
        C.blank_patch(torch.zeros(4, 4, 4, dtype=torch.uint8), (1, 1, 1), 1, "sphere")


def test_product_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ctunet_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "/root/reference" not in src, fn


def test_install_rebinds_reference_names():
    """``install`` makes eval('UNetSP')() inside the trainer module resolve to the B200 class and swaps
    the loss half of the handlers, leaving the reference's dataset/writer halves alone."""
    import ctunet_b200 as C

    class RefHandler:                      # stands in for ctunet.pytorch.ProblemHandler.ProblemHandler
        @staticmethod
        def comp_losses_metrics(model, prediction, target, idx, n_imgs):
            raise AssertionError("reference loss must have been replaced")

        def write_predictions(self, *a):
            return "reference writer kept"

    class RefDouble(RefHandler):
        pass

    trainer = types.ModuleType("fake_trainer")
    trainer.ProblemHandler = RefHandler
    trainer.FlapRecWithShapePriorDoubleOut = RefDouble
    trainer.UNetSP = object
    done = C.install(trainer)
    C.dropin._SAVED.clear()                # (a throw-away module: nothing to restore)
    assert eval("UNetSP", vars(trainer)) is C.UNetSP
    assert eval("recAE_v2_fixed", vars(trainer)) is C.recAE_v2_fixed
    assert "UNetSP" in done and "FlapRecWithShapePriorDoubleOut.comp_losses_metrics" in done
    assert RefDouble().write_predictions() == "reference writer kept"
    assert RefDouble.comp_losses_metrics is not RefHandler.comp_losses_metrics
    assert RefDouble.comp_losses_metrics.__doc__.startswith("ProblemHandler.py:213")


def test_install_into_real_reference_when_present():
    from oracle.reference_loader import reference_available, load_reference
    if not reference_available():
        pytest.skip("reference tree not mounted (GPU box)")
    import importlib
    import ctunet_b200 as C
    load_reference()
    trainer = importlib.import_module("ctunet.pytorch.Model")
    ref_cls = trainer.UNetSP
    ph = trainer.FlapRecWithShapePriorDoubleOut.comp_losses_metrics
    try:
        C.install()
        assert type(eval("UNetSP", vars(trainer))()).__module__ == "ctunet_b200.models"
        assert trainer.FlapRecWithShapePriorDoubleOut.comp_losses_metrics is not ph
    finally:
        assert C.uninstall() > 0
        torch.autograd.set_detect_anomaly(False)
    assert trainer.UNetSP is ref_cls and trainer.FlapRecWithShapePriorDoubleOut.comp_losses_metrics is ph
    assert trainer.utils.dice_coeff.__module__ == "ctunet.utilities"


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: the installed reference from oracle/_ref or /root/reference
    when present, else the oracle port -- the one place besides tests / smoke that may run oracle/) needs no GPU and prints
    one JSON line with the keys the driver reads; both variants are exercised."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from oracle.reference_loader import reference_available
    kinds = []
    for hide in (False, True):
        env = dict(os.environ)
        if hide:                 # point the loader at nothing: the port must take over
            env["CTUNET_REFERENCE_ROOT"] = "/nonexistent"
            env["CTUNET_REFERENCE_INSTALL"] = "/nonexistent"
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--size", "32", "--steps",
                              "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = out.stdout.strip().splitlines()
        assert len(lines) == 1, lines                     # the reference's own prints must not reach stdout
        line = json.loads(lines[-1])
        assert line["impl"] == "reference" and line["unit"] == "voxels/s" and line["value"] > 0
        assert line["higher_is_better"] is True and line["gpu_launches"] == 0
        assert line["cpu_baseline"]["cores"] >= 1
        assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
        kinds.append(line["cpu_baseline"]["kind"])
    assert kinds[1] == "port" and kinds[0] == ("reference" if reference_available() else "port")
