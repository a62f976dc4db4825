"""FCD mask generation -- CPU restatement of detection/model_utils.py:118-216 (TEST INFRASTRUCTURE).

`FCDMaskGenerator.get_mask` = per-patch classification over the sliding-window patches (:136-180), a 4-neighbour vote
(:182-193) and painting the patch labels back into a volume (:195-216).  The restatement keeps the reference's observable
behaviour, including two quirks that a reader would not guess:
  * `_postprocess` indexes `patch_map_tensor` with INT64 0/1 arrays (`change_to_pos`, `change_to_neg`), i.e. fancy indexing
    along axis 0: whole slabs 0 / 1 of the (4, Y//h, Z) map are overwritten, the vote itself never reaches the map;
  * `_masking` paints rows with the slice `-j:-j-h:-1`, which is EMPTY for the first strip (j = 0) and one voxel off the
    strip's rows for the others.
The classifier is a callable `labels = classify(patches)` on a (P,2,h,w) float64 array so that the same logic can be checked
with any model (the reference calls its global `model` patch by patch).
"""
from __future__ import annotations

import numpy as np

from . import patches as P

SLOT_OF = {0: 0, 1: 3, 2: 1, 3: 2}      # emission order patch_1, patch_2, patch_3, patch_4 -> patch_map_tensor index (:160-178)


def plan_slots(plan, X, w):
    """patch-map index (0..3) of every plan row: patch_1 -> 0, patch_3 -> 1, patch_4 -> 2, patch_2 -> 3 (model_utils.py:160-178)."""
    mid = X // 2 - w
    c0 = plan[:, P.C0]
    slot = np.where(c0 == mid, 1, np.where(c0 == X - mid - w, 2, np.where(c0 < mid, 0, 3)))
    return slot.astype(np.int64)


def predictions_per_batches(img, gmpm, classify, h=16, w=32):
    """model_utils.py:136-180 -> int64 (4, Y//h, Z)."""
    X, Y, Z = gmpm.shape
    plan = P.patch_plan(gmpm, None, h, w)
    labels = np.asarray(classify(P.gather_patches(img, plan, h, w))).astype(np.int64)
    pm = np.zeros((4, Y // h, Z), dtype=np.int64)
    pm[plan_slots(plan, X, w), plan[:, P.ROW0] // h, plan[:, P.SLICE]] = labels
    return pm


def postprocess(patch_map):
    """model_utils.py:182-193, quirk included (integer arrays used as indices along axis 0)."""
    from scipy.signal import convolve
    pm = patch_map.copy()
    k = 0.25 * np.array([[[0, 1, 0], [1, 0, 1], [0, 1, 0]]])
    res = convolve(pm, k, mode="same")
    change_to_pos = (res == 1.0).astype("int64")
    change_to_neg = (res == 0.0).astype("int64")
    pm[change_to_pos] = 1
    pm[change_to_neg] = 0
    return pm


def masking(img, gmpm, patch_map, h=16, w=32):
    """model_utils.py:195-216 (same slices, same order: later assignments overwrite earlier ones)."""
    X, Y, Z = gmpm.shape
    final = np.zeros_like(img)
    mid = X // 2 - w
    for i in range(Z):
        S = np.rot90(gmpm[:, :, i])
        for j in range(0, Y, h):
            strip = S[j:j + h]
            if strip.sum() == 0.0:
                continue
            start = int((strip.sum(0) > 0).argmax())
            if start < mid:
                final[start:start + w, -j:-j - h:-1, i] = patch_map[0, j // h, i]
                final[-start - w:-start, -j:-j - h:-1, i] = patch_map[3, j // h, i]
            final[mid:mid + w, -j:-j - h:-1, i] = patch_map[1, j // h, i]
            final[-mid - w:-mid, -j:-j - h:-1, i] = patch_map[2, j // h, i]
    return final


def get_mask(img, gmpm, classify, h=16, w=32):
    """model_utils.py:218-222."""
    pm = predictions_per_batches(img, gmpm, classify, h, w)
    pm = postprocess(pm)
    return masking(img, gmpm, pm, h, w).astype("int64")


def get_iou(pred_mask, true_mask):
    """model_utils.py:224-228."""
    assert pred_mask.shape == true_mask.shape, "Wrong shape of masks"
    return np.logical_and(pred_mask, true_mask).sum() / np.logical_or(pred_mask, true_mask).sum()


def mean_threshold_classifier(thr=0.5):
    """A deterministic stand-in for the (unshipped) best_model.pth: label 1 when the mean of channel 0 exceeds `thr`."""
    def classify(patches):
        return (patches[:, 0].mean(axis=(1, 2)) > thr).astype(np.int64)
    return classify
