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


UTILITY_FUNCTIONS = ["dice_loss", "dice_coeff", "hausdorff"]
_MISSING = object()
_SAVED = []          # (owner, attribute name, previous value in the owner's __dict__ or _MISSING), in rebinding order


def _rebind(owner, name, value):
    _SAVED.append((owner, name, vars(owner).get(name, _MISSING)))
    setattr(owner, name, value)


def uninstall() -> int:
    """Undo every ``install()`` since the last ``uninstall()``: the reference's own classes, handlers and utility functions
    are back.  Returns the number of restored names."""
    n = len(_SAVED)
    while _SAVED:
        owner, name, old = _SAVED.pop()
        if old is _MISSING:
            if name in vars(owner):
                delattr(owner, name)
        else:
            setattr(owner, name, old)
    return n


def install(trainer_module=None, replace_losses: bool = True, data_parallel: str = "keep", graph: bool = False):
    """Rebind the reference's model (and optionally loss-handler) names to the B200 classes.

    ``trainer_module`` defaults to ``ctunet.pytorch.Model`` (must be importable).  The reference's handler classes are
    kept (they own the dataset / NIfTI-writer halves, which are out of scope); only their ``comp_losses_metrics`` static
    method is replaced by the fused one -- including the ``dice_coef_*`` / ``hd_coef_*`` metrics every stock ``.ini``
    switches on.  ``ctunet.utilities.dice_loss / dice_coeff / hausdorff`` are rebound as well (CUDA tensors only).
    NOT rebound, on purpose: ``utils.hard_segm_from_tensor`` (the reference calls it on ``.cpu()`` tensors inside its
    file writers, ProblemHandler.py:340-341) and the dataset transforms (they run on host tensors in DataLoader workers);
    their device versions live in ``ctunet_b200.utilities`` for pipelines that keep the data on the GPU.

    ``data_parallel``: 'keep' leaves ``Model.new_model`` alone -- with several visible GPUs the reference wraps the model
    in ``nn.DataParallel`` (Model.py:481-486), which these modules support (replicas read the broadcast weights);
    'single' patches ``new_model`` to build the bare module on the current device, for one-process-per-GPU launches
    (torchrun + ``parallel.GradSync``).  ``graph``: the installed modules replay captured CUDA graphs for the forward and
    the backward pass of a training iteration (``models._GraphNetFn``; the third iteration of a batch shape captures) -- the
    eager launch stream is host-bound behind the reference's own step driver.  Returns the list of rebound names."""
    import importlib
    if trainer_module is None:
        trainer_module = importlib.import_module("ctunet.pytorch.Model")
    if data_parallel not in ("keep", "single"):
        raise ValueError("data_parallel: 'keep' or 'single'")
    _rebind(models, "DROPIN_GRAPH", bool(graph))
    done = []
    for name in MODEL_CLASSES:
        _rebind(trainer_module, name, getattr(models, name))
        done.append(name)
    if replace_losses:
        for name in HANDLER_CLASSES + ["ProblemHandler", "FlapRec", "FlapRecWithShapePrior"]:
            ref_cls = getattr(trainer_module, name, None)
            ours = getattr(losses, name)
            if ref_cls is None:
                _rebind(trainer_module, name, ours)
            elif "comp_losses_metrics" in vars(ours):
                _rebind(ref_cls, "comp_losses_metrics", staticmethod(vars(ours)["comp_losses_metrics"].__func__))
            done.append(name + ".comp_losses_metrics")
        from . import utilities as ours_utils
        ref_utils = getattr(trainer_module, "utils", None)
        if ref_utils is not None:
            for name in UTILITY_FUNCTIONS:
                _rebind(ref_utils, name, getattr(ours_utils, name))
                done.append("utils." + name)
    if data_parallel == "single" and hasattr(trainer_module, "Model"):
        def new_model(self):                                  # Model.py:474-491 without the nn.DataParallel wrapper
            model = eval(self.params["model_class"], vars(trainer_module))()
            model.to(self.params["device"])
            return model
        _rebind(trainer_module.Model, "new_model", new_model)
        done.append("Model.new_model")
    return done
