"""Plugs the B200 path into the reference's trainer.

The reference resolves its model and loss plug-ins by name with ``eval`` inside
``ctunet/pytorch/Model.py`` (``eval(self.params["model_class"])()`` at Model.py:485,488 and
``eval(self.params["problem_handler"])()`` at Model.py:101), against names star-imported from
``ctunet.pytorch.models`` (Model.py:18) and ``ctunet.pytorch.ProblemHandler`` (Model.py:17).
``install()`` rebinds those names in the trainer module's globals, so an unmodified ``.ini`` /
``run.py`` (examples/UNetSPDO/run.py) builds the B200 modules and losses instead.
"""
from __future__ import annotations

from . import losses, models

MODEL_CLASSES = ["UNet", "UNet4b2i3o", "UNet5b2i3o", "UNet4b1i3o", "UNetSP", "UNetSPSmall", "UNetDO",
                 "recAE_v2_fixed", "UNet4_2IC", "UNetBlock", "CenterBlock", "ResidualBlock",
                 "down_block_cr", "up_block_cr"]
HANDLER_CLASSES = ["FlapRecWithShapePriorDoubleOut", "FlapRecDoubleOut"]


def install(trainer_module=None, replace_losses: bool = True):
    """Rebind the reference's model (and optionally loss-handler) names to the B200 classes.

    ``trainer_module`` defaults to ``ctunet.pytorch.Model`` (must be importable).  The reference's
    handler classes are kept (they own the dataset / NIfTI-writer halves, which are out of scope);
    only their ``comp_losses_metrics`` static method is replaced by the fused one.
    Returns the list of rebound names."""
    if trainer_module is None:
        import importlib
        trainer_module = importlib.import_module("ctunet.pytorch.Model")
    done = []
    for name in MODEL_CLASSES:
        setattr(trainer_module, name, getattr(models, name))
        done.append(name)
    if replace_losses:
        for name in HANDLER_CLASSES + ["ProblemHandler", "FlapRec", "FlapRecWithShapePrior"]:
            ref_cls = getattr(trainer_module, name, None)
            ours = getattr(losses, name)
            if ref_cls is None:
                setattr(trainer_module, name, ours)
            elif "comp_losses_metrics" in vars(ours):
                ref_cls.comp_losses_metrics = staticmethod(vars(ours)["comp_losses_metrics"].__func__)
            done.append(name + ".comp_losses_metrics")
    return done
