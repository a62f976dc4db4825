"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the counting metrics of the validation loop,
`compute_dice_coefficient` (segmentation/metrics.py:312-329) and `get_iou_score` (segmentation/routine.py:198-204).
Pinned by tests/golden/overlap_metrics.npz (oracle/make_golden.py `metrics_cases`: the reference's own functions)."""
import numpy as np


def compute_dice_coefficient(mask_gt, mask_pred):       # metrics.py:312-329 (np.NaN there; removed in numpy 2)
    volume_sum = mask_gt.sum() + mask_pred.sum()
    if volume_sum == 0:
        return np.nan
    volume_intersect = (mask_gt & mask_pred).sum()
    return 2 * volume_intersect / volume_sum


def get_iou_score(prediction, ground_truth):            # routine.py:198-204
    intersection, union = 0, 0
    intersection += np.logical_and(prediction > 0, ground_truth > 0).astype(np.float32).sum()
    union += np.logical_or(prediction > 0, ground_truth > 0).astype(np.float32).sum()
    return float(intersection) / union
