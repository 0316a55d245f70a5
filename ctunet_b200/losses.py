"""Fused soft-Dice + CrossEntropy loss and the reference's ``comp_losses_metrics`` plug-ins.

Reference: ``utils.dice_loss`` (ctunet/utilities.py:35-50) and
``ProblemHandler.comp_losses_metrics`` / ``FlapRecWithShapePriorDoubleOut.comp_losses_metrics``
(ctunet/pytorch/ProblemHandler.py:44-102, 213-309).  The handlers keep the reference's duck-typed
contract -- they read ``model.params[...]``, leave a scalar tensor in ``model.pt_loss`` on which
``.backward()`` works (Model.py:366) and append Python floats per key to
``model.losses_and_metrics`` -- but read all components back with ONE host synchronisation per
batch instead of one per component.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import call, stream_ptr


def _check(pred, target):
    if not (pred.is_cuda and target.is_cuda):
        raise RuntimeError("ctunet_b200 losses run on CUDA only (no CPU fallback)")
    if pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise TypeError("prediction and target must be float32")
    if pred.shape != target.shape or pred.dim() < 3:
        raise ValueError("prediction %s and one-hot target %s must have the same [B, C, ...] shape"
                         % (tuple(pred.shape), tuple(target.shape)))
    if not 1 <= pred.shape[1] <= 4:
        raise NotImplementedError("the fused loss covers 1..4 channels")


class _DiceCE(torch.autograd.Function):
    """Returns a float32[2] tensor: (CrossEntropy mean, soft-Dice loss)."""

    @staticmethod
    def forward(ctx, pred, target, softmax_for_dice: bool, want_ce: bool):
        _check(pred, target)
        pred, target = pred.contiguous(), target.contiguous()
        b, c = pred.shape[0], pred.shape[1]
        spatial = pred[0, 0].numel()
        sums = torch.empty(4 * b, dtype=torch.float64, device=pred.device)
        out = torch.empty(2, dtype=torch.float32, device=pred.device)
        call("ctu_dice_ce_fwd", pred.data_ptr(), target.data_ptr(), b, c, spatial, int(softmax_for_dice),
             int(want_ce), sums.data_ptr(), out.data_ptr(), stream_ptr())
        ctx.save_for_backward(pred, target, sums)
        ctx.cfg = (b, c, spatial, int(softmax_for_dice), int(want_ce))
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target, sums = ctx.saved_tensors
        b, c, spatial, sm, ce = ctx.cfg
        g = g.contiguous().float()
        dpred = torch.empty_like(pred)
        call("ctu_dice_ce_bwd", pred.data_ptr(), target.data_ptr(), b, c, spatial, sm, ce, sums.data_ptr(),
             g.data_ptr(), dpred.data_ptr(), stream_ptr())
        return dpred, None, None, None


def dice_ce(pred, target, softmax_for_dice: bool, want_ce: bool = True):
    """(ce, dice) as 0-dim tensors.  ``softmax_for_dice``: Dice is taken on softmax(pred, dim=1)
    (ProblemHandler.py:234-235); CE always treats ``pred`` as logits (ProblemHandler.py:69, 251)."""
    out = _DiceCE.apply(pred, target, softmax_for_dice, want_ce)
    return out[0], out[1]


class dice_loss(nn.Module):
    """Drop-in for ``ctunet.utilities.dice_loss`` (utilities.py:35-50): ``dice_loss()(probs, onehot)``."""

    def forward(self, output, masks):
        if not masks.is_contiguous() or not output.is_contiguous():
            # the reference uses .view(b_size, -1), which raises on non-contiguous inputs
            raise RuntimeError("view size is not compatible with input tensor's size and stride")
        b = masks.size(0)
        return _DiceCE.apply(output.reshape(b, 1, -1), masks.reshape(b, 1, -1), False, False)[1]


def _append(lm, key, value):
    lm.setdefault(key, []).append(value)


class ProblemHandler:
    """ctunet/pytorch/ProblemHandler.py:21-102 (the loss half; datasets and NIfTI writers are out of scope)."""

    def __init__(self, train_dataset_class=None, test_dataset_class=None):
        self.train_dataset_class = train_dataset_class
        self.test_dataset_class = test_dataset_class

    def write_predictions(self, predictions, input_filepaths, output_folder_name, input_imgs):
        raise NotImplementedError("NIfTI prediction writers need SimpleITK (file IO is outside the hot path); "
                                  "ctunet_b200.install() keeps the reference's own writers and only replaces the labelling "
                                  "(utils.hard_segm_from_tensor)")

    @staticmethod
    def comp_losses_metrics(model, prediction, target, idx, n_imgs, verbose=True):
        """ProblemHandler.py:44-102: ce_lambda * CE(prediction, argmax(target)) + dice_lambda * Dice."""
        ce_l, dice_l = model.params["ce_lambda"], model.params["dice_lambda"]
        if target.dim() != 5:
            raise NotImplementedError("the fused loss expects a one-hot float target [B, C, D, H, W]")
        ce, dice = dice_ce(prediction, target, softmax_for_dice=False, want_ce=ce_l != 0)
        terms, keys = [], []
        if ce_l != 0:
            terms.append(ce_l * ce)
            keys.append("ce")
        if dice_l != 0:
            terms.append(dice_l * dice)
            keys.append("dice_loss")
        model.pt_loss = sum(terms)
        vals = torch.stack([t.detach() for t in terms] + [model.pt_loss.detach()]).tolist()   # one sync
        for k, v in zip(keys, vals):
            _append(model.losses_and_metrics, k, v)
        if model.params["save_dice_plots"] is True:                                           # ProblemHandler.py:84-88
            from .utilities import dice_coeff
            _append(model.losses_and_metrics, "dice_coef", dice_coeff(prediction, target))    # a tensor, as the reference
        _append(model.losses_and_metrics, "epoch_loss", vals[-1])
        if verbose:
            print("    Batch {}/{} ({:.0f}%)\tLoss: {:.6f}".format(idx + 1, n_imgs, 100.0 * (idx + 1) / n_imgs, vals[-1]))


class FlapRec(ProblemHandler):
    """ProblemHandler.py:166-173"""


class FlapRecWithShapePrior(ProblemHandler):
    """ProblemHandler.py:176-188"""


class FlapRecWithShapePriorDoubleOut(ProblemHandler):
    """ProblemHandler.py:191-309"""

    def __init__(self, with_sp=True):
        super().__init__()
        self.with_sp = with_sp

    @staticmethod
    def comp_losses_metrics(model, prediction, target, idx, n_imgs, verbose=True):
        """ProblemHandler.py:213-309: for the (full skull, flap) pair, CE on the raw predictions and
        soft-Dice on their softmax, weighted by ce_lambda / dice_lambda and summed in the reference's
        order (ce_sk, ce_fl, dice_sk, dice_fl)."""
        ce_l, dice_l = model.params["ce_lambda"], model.params["dice_lambda"]
        sk_p, fl_p = prediction
        sk_t, fl_t = target
        ce_s, dice_s = dice_ce(sk_p, sk_t, softmax_for_dice=True, want_ce=ce_l != 0)
        ce_f, dice_f = dice_ce(fl_p, fl_t, softmax_for_dice=True, want_ce=ce_l != 0)
        terms, keys = [], []
        if ce_l != 0:
            terms += [ce_l * ce_s, ce_l * ce_f]
            keys += ["ce_sk", "ce_fl"]
        if dice_l != 0:
            terms += [dice_l * dice_s, dice_l * dice_f]
            keys += ["dice_loss_sk", "dice_loss_fl"]
        model.pt_loss = sum(terms)
        vals = torch.stack([t.detach() for t in terms] + [model.pt_loss.detach()]).tolist()   # one sync
        lm = model.losses_and_metrics
        for k, v in zip(keys, vals):
            _append(lm, k, v)
        # Metrics (ProblemHandler.py:277-295).  The reference evaluates them on softmax(prediction); the hard labels they
        # are built from (argmax over the channel axis) are the same for the raw predictions.  Appended as 0-dim device
        # tensors like the reference does (Model.update_plots_tensorboard_avg averages and float()s them once per epoch),
        # so they cost no host synchronisation here.  `save_hd_plots` is read unconditionally, like ProblemHandler.py:287.
        if model.params["save_dice_plots"] is True:
            from .utilities import dice_coeff
            _append(lm, "dice_coef_sk", dice_coeff(sk_p, sk_t))
            _append(lm, "dice_coef_fl", dice_coeff(fl_p, fl_t))
        if model.params["save_hd_plots"] is True:
            from .utilities import hausdorff
            _append(lm, "hd_coef_sk", hausdorff(sk_p, sk_t))
            _append(lm, "hd_coef_fl", hausdorff(fl_p, fl_t))
        _append(lm, "epoch_loss", vals[-1])
        if verbose:
            print("    Batch {}/{} ({:.0f}%)\tLoss: {:.6f}".format(idx + 1, n_imgs, 100.0 * (idx + 1) / n_imgs, vals[-1]))


class FlapRecDoubleOut(FlapRecWithShapePriorDoubleOut):
    """ProblemHandler.py:357-359"""

    def __init__(self):
        super().__init__(with_sp=False)
