"""TEST INFRASTRUCTURE ONLY -- `monai.metrics.compute_meandice` / `compute_hausdorff_distance` RESTATED.

PARITY UNPINNED: the reference imports monai (ctunet/utilities.py:19) without declaring or pinning it
(setup.py:6-8) and it is not installable here.  Restated from MONAI's published algorithm (0.5 - 0.8 API):

compute_meandice(y_pred, y, include_background=True) -> [B, C']:
    drop channel 0 unless include_background; per (batch, channel):
    f = 2 * sum(y * y_pred) / (sum(y) + sum(y_pred)), NaN where sum(y) == 0.
compute_hausdorff_distance(y_pred, y, include_background=False, distance_metric="euclidean",
                           percentile=None, directed=False) -> [B, C-1]:
    per (batch, class): edges = mask XOR binary_erosion(mask) (6-neighbourhood, outside = background) of both
    masks (cropped to the bounding box of their union -- which does not change the edge set), directed distance
    = max over the edge voxels of A of the Euclidean distance transform to the edge voxels of B; the metric is
    the max of the two directions.  An empty edge set gives NaN / inf (the reference maps both to
    ``max(reference.shape)``, utilities.py:63-70).
"""
import numpy as np
import torch


def _drop_background(y_pred, y, include_background):
    if not include_background:
        y_pred, y = y_pred[:, 1:], y[:, 1:]
    return y_pred, y


def compute_meandice(y_pred, y, include_background=True):
    y_pred, y = _drop_background(y_pred, y, include_background)
    y = y.float()
    y_pred = y_pred.float()
    if y.shape != y_pred.shape:
        raise ValueError("y_pred and y should have same shapes.")
    axes = list(range(2, y_pred.dim()))
    intersection = torch.sum(y * y_pred, dim=axes)
    y_o = torch.sum(y, axes)
    y_pred_o = torch.sum(y_pred, dim=axes)
    denominator = y_o + y_pred_o
    return torch.where(y_o > 0, (2.0 * intersection) / denominator,
                       torch.tensor(float("nan"), device=y_o.device))


def get_mask_edges(seg_pred, seg_gt):
    from scipy.ndimage import binary_erosion
    seg_pred, seg_gt = np.asarray(seg_pred) != 0, np.asarray(seg_gt) != 0
    if not np.any(seg_pred | seg_gt):
        return np.zeros_like(seg_pred), np.zeros_like(seg_gt)
    nz = np.argwhere(seg_pred | seg_gt)
    lo, hi = nz.min(0), nz.max(0) + 1
    sl = tuple(slice(a, b) for a, b in zip(lo, hi))
    seg_pred, seg_gt = seg_pred[sl], seg_gt[sl]
    return binary_erosion(seg_pred) ^ seg_pred, binary_erosion(seg_gt) ^ seg_gt


def _directed(edges_a, edges_b):
    from scipy.ndimage import distance_transform_edt
    if not np.any(edges_b):
        dis = np.inf * np.ones_like(edges_b, dtype=np.float64)
    else:
        if not np.any(edges_a):
            return np.inf            # max over dis[edges_b] of an all-inf map
        dis = distance_transform_edt(~edges_b)
    sd = np.asarray(dis[edges_a])
    if sd.shape == (0,):
        return np.nan
    return sd.max()


def compute_hausdorff_distance(y_pred, y, include_background=False, distance_metric="euclidean", percentile=None,
                               directed=False):
    if distance_metric != "euclidean" or percentile is not None:
        raise NotImplementedError("restated for the reference's call (utilities.py:64-68) only")
    y_pred, y = _drop_background(y_pred, y, include_background)
    if y.shape != y_pred.shape:
        raise ValueError("y_pred and y should have same shapes.")
    b, c = y_pred.shape[:2]
    hd = np.empty((b, c))
    yp, yt = y_pred.detach().cpu().numpy(), y.detach().cpu().numpy()
    for bi in range(b):
        for ci in range(c):
            ep, eg = get_mask_edges(yp[bi, ci], yt[bi, ci])
            d = _directed(ep, eg)
            if not directed:
                d2 = _directed(eg, ep)
                d = np.nan if (np.isnan(d) or np.isnan(d2)) else max(d, d2)
            hd[bi, ci] = d
    return torch.from_numpy(hd)
