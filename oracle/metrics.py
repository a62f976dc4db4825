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


def compute_surface_distances(mask_gt, mask_pred, spacing_mm, neighbour_code_to_normals):
    """segmentation/metrics.py:25-178 restated (same scipy calls the reference makes: ndimage.correlate with the 2x2x2 bit-weight
    kernel, ndimage.distance_transform_edt with `sampling`), with `np.inf` for the `np.Inf` numpy 2 removed.  The normal table
    (metrics.py:333-600) is passed in.  Pinned by tests/golden/surface_distances.npz (made by the reference's own function)."""
    from scipy import ndimage
    area = np.zeros([256])
    for code in range(256):
        normals = np.array(neighbour_code_to_normals[code])
        s = 0
        for i in range(normals.shape[0]):
            n = np.zeros([3])
            n[0] = normals[i, 0] * spacing_mm[1] * spacing_mm[2]
            n[1] = normals[i, 1] * spacing_mm[0] * spacing_mm[2]
            n[2] = normals[i, 2] * spacing_mm[0] * spacing_mm[1]
            s += np.linalg.norm(n)
        area[code] = s
    mask_all = mask_gt | mask_pred
    empty = {"distances_gt_to_pred": np.array([]), "distances_pred_to_gt": np.array([]), "surfel_areas_gt": np.array([]), "surfel_areas_pred": np.array([])}
    lo, hi = np.zeros(3, np.int64), np.zeros(3, np.int64)
    for ax in range(3):
        proj = np.max(mask_all, axis=tuple(a for a in range(3) if a != ax))
        nz = np.nonzero(proj)[0]
        if len(nz) == 0:
            return empty
        lo[ax], hi[ax] = nz.min(), nz.max()
    crop = []
    for m in (mask_gt, mask_pred):
        c = np.zeros((hi - lo) + 2, np.uint8)
        c[0:-1, 0:-1, 0:-1] = m[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1]
        crop.append(c)
    kernel = np.array([[[128, 64], [32, 16]], [[8, 4], [2, 1]]])
    codes = [ndimage.correlate(c.astype(np.uint8), kernel, mode="constant", cval=0) for c in crop]
    borders = [(c != 0) & (c != 255) for c in codes]
    dist = [ndimage.distance_transform_edt(~b, sampling=spacing_mm) if b.any() else np.inf * np.ones(b.shape) for b in borders]
    out = {}
    for name, own, other in (("gt", 0, 1), ("pred", 1, 0)):
        d, a = dist[other][borders[own]], area[codes[own]][borders[own]]
        if d.shape != (0,):
            srt = np.array(sorted(zip(d, a)))
            d, a = srt[:, 0], srt[:, 1]
        out["distances_gt_to_pred" if name == "gt" else "distances_pred_to_gt"] = d
        out["surfel_areas_" + name] = a
    return out
