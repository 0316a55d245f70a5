"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference trainer (``ctunet.pytorch.Model.Model``) without its
file-system half.

``Model.__init__`` (ctunet/pytorch/Model.py:24-145) resolves output folders, builds DataLoaders over NIfTI files
(SimpleITK) and a TensorBoard writer, then trains.  The hot path -- ``initialize_models`` / ``initialize_optimizer`` /
``forward_pass`` (Model.py:324-380, 474-546) -- only needs the attributes set below, so the object is created with
``__new__`` and handed an in-memory loader of ``{'image', 'target'}`` samples.  Every method that then runs is the
reference's own code, unmodified: with ``ctunet_b200.install()`` applied it exercises the drop-in (the installed model
classes and loss handlers are looked up by the reference's ``eval``), without it the reference's stock CPU / cuDNN path.
"""
import os

import torch

from .reference_loader import load_reference

EXAMPLES = {     # the six stock example configurations (examples/**.ini) -> (relative path)
    "FlapRecSP2O": "examples/UNetSPDO/FlapRecSP2O.ini",
    "FlapRecSP2O_128": "examples/UNetSPDO/FlapRecSP2O_128.ini",
    "FlapRecSP2O_512": "examples/UNetSPDO/FlapRecSP2O_512.ini",
    "autoimplant_FlapRecSP2O": "examples/autoimplant2020/UNetSPDO/FlapRecSP2O.ini",
    "AutoImplant2020_woShapePrior": "examples/autoimplant2020/UNet/AutoImplant2020_woShapePrior.ini",
    "AutoImplant2020_wShapePrior": "examples/autoimplant2020/UNetSP/AutoImplant2020_wShapePrior.ini",
}


def example_params(name: str) -> dict:
    """The reference's own ini parser (utilities.set_cfg_params, utilities.py:215-256) on a stock example file."""
    from .reference_loader import reference_root
    UT = load_reference()[2]
    root = reference_root()
    path = os.path.join(root, EXAMPLES[name])
    defaults = {"resume_model": "", "force_resumed": False}
    return UT.set_cfg_params(path, defaults)


def make_trainer(params: dict, device: str):
    """A ``ctunet.pytorch.Model.Model`` ready for ``initialize_models()`` / ``initialize_optimizer()`` /
    ``forward_pass(phase, loader)``; ``params`` as ``load_params`` returns them (model_class, problem_handler, optimizer,
    learning_rate, momentum, weight_decay, dice_lambda, ce_lambda, save_dice_plots, save_hd_plots[, scheduler])."""
    MM = load_reference(with_trainer=True)[4]
    m = MM.Model.__new__(MM.Model)
    m.params = dict(params)
    m.params["device"] = torch.device(device)
    m.params.setdefault("resume_model", "")
    m.params.setdefault("name", "harness")
    # Model.py:101-103 -- the handler instance is resolved by eval in the trainer module's namespace
    m.problem_handler = eval(m.params["problem_handler"], vars(MM))()
    m.write_predictions = m.problem_handler.write_predictions
    m.comp_losses_metrics = m.problem_handler.comp_losses_metrics
    m.models = {"main": None, "acnn": None}
    m.out_paths = None
    m.current_epoch = m.current_train_iteration = 0
    m.best_model = {"epoch": 1, "value": None}
    m.pt_loss = []
    m.losses_and_metrics = {}
    return m


class ListLoader(list):
    """An in-memory stand-in for the DataLoader: a list of ``{'image': Tensor, 'target': Tensor | [Tensor, Tensor]}``
    batches (host tensors, moved by ``forward_pass`` itself, Model.py:343-349).  Targets must be LISTS for the
    double-output handlers, as the default collate delivers them (``type(sample['target']) == list``, Model.py:345)."""
