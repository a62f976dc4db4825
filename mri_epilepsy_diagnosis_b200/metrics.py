"""Validation overlap metrics on the device (the counting part): `compute_dice_coefficient`
(segmentation/metrics.py:312-329) and `get_iou_score` (segmentation/routine.py:198-204) on GPU label volumes, as
`validate_dsc_asd` (routine.py:216-237) calls them after `logits.argmax(dim=1)`.  One kernel pass counts everything both need;
5 integers come back to the host.  The surface distances of the same loop (`compute_surface_distances`) are not built."""
from __future__ import annotations

import numpy as np
import torch

from ._cabi import check, lib, need_cuda, stream


def _u8(t):
    need_cuda(t, "metrics")
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    elif t.dtype != torch.uint8:
        t = t.to(torch.uint8)                     # `.astype(np.uint8)` at routine.py:226-229 (wraps like numpy)
    return t.contiguous()


def overlap_counts(prediction, ground_truth):
    """(gt.sum(), pred.sum(), (gt & pred).sum(), #(pred>0 and gt>0), #(pred>0 or gt>0)) as Python ints"""
    p, g = _u8(prediction), _u8(ground_truth)
    if p.shape != g.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(p.shape)} {tuple(g.shape)}")
    out = torch.empty(5, dtype=torch.int64, device=p.device)
    check(lib().b200_overlap_counts(p.data_ptr(), g.data_ptr(), p.numel(), out.data_ptr(), stream()))
    return tuple(int(v) for v in out.tolist())


def compute_dice_coefficient(mask_gt, mask_pred):
    """segmentation/metrics.py:312-329: NaN when both masks are empty"""
    sg, sp, sand, _, _ = overlap_counts(mask_pred, mask_gt)
    volume_sum = np.uint64(sg) + np.uint64(sp)
    if volume_sum == 0:
        return np.nan
    return 2 * np.uint64(sand) / volume_sum


def get_iou_score(prediction, ground_truth):
    """segmentation/routine.py:198-204: float32 sums of the logical masks (exact below 2^24 voxels), then the reference's
    own expression `float(intersection) / union` on them"""
    _, _, _, inter, union = overlap_counts(prediction, ground_truth)
    intersection, union_ = 0, 0
    intersection += np.float32(inter)
    union_ += np.float32(union)
    return float(intersection) / union_


def calculate_overlap(surface, prediction):
    """the (dsc, iou) half of calculate_metrics (routine.py:206-214) from ONE counting pass"""
    sg, sp, sand, inter, union = overlap_counts(prediction, surface)
    vs = np.uint64(sg) + np.uint64(sp)
    dsc = np.nan if vs == 0 else 2 * np.uint64(sand) / vs
    return dsc, float(np.float32(inter)) / np.float32(union)
