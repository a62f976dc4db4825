"""Validation metrics on the device, as `validate_dsc_asd` (segmentation/routine.py:216-237) calls them
after `logits.argmax(dim=1)`:

  compute_dice_coefficient, get_iou_score        segmentation/metrics.py:312-329, routine.py:198-204 -- one counting pass, 5 integers
  compute_surface_distances                      segmentation/metrics.py:25-178 -- neighbour codes + EXACT Euclidean distance
                                                 transform + surfel lists on the GPU (the reference spends ~7.5 s per volume in scipy)
  compute_average_surface_distance, compute_robust_hausdorff, compute_surface_overlap_at_tolerance,
  compute_surface_dice_at_tolerance              metrics.py:181-309 -- numpy arithmetic on the (small) sorted surfel lists
  calculate_metrics                              routine.py:206-214

Same names, arguments, return types (numpy float64 arrays / floats) and empty-mask behaviour as the reference.  The surfel-area
table is derived per call from data/neighbour_code_normals.npy -- the 256-entry marching-cubes normal table of Google's
surface-distance library (Apache-2.0) that the reference file vendors at metrics.py:333-600, shipped as data -- with the
reference's own arithmetic (metrics.py:58-71)."""
from __future__ import annotations

import ctypes as C
import os

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


# ----------------------------------------------------------------------------------------------- surface distances
_NORMALS = None


def _normals():
    global _NORMALS
    if _NORMALS is None:
        _NORMALS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "neighbour_code_normals.npy"))   # (256, 4, 3) + count in [:, :, 0]
    return _NORMALS


def neighbour_code_to_surface_area(spacing_mm):
    """metrics.py:58-71, operation by operation (float64)"""
    tab = _normals()
    out = np.zeros([256])
    for code in range(256):
        cnt = int(tab["count"][code])
        normals = tab["normals"][code][:cnt]
        sum_area = 0
        for normal_idx in range(cnt):
            n = np.zeros([3])
            n[0] = normals[normal_idx, 0] * spacing_mm[1] * spacing_mm[2]
            n[1] = normals[normal_idx, 1] * spacing_mm[0] * spacing_mm[2]
            n[2] = normals[normal_idx, 2] * spacing_mm[0] * spacing_mm[1]
            sum_area += np.linalg.norm(n)
        out[code] = sum_area
    return out


def _surface_lists(gt, pred, spacing_mm):
    """device part: (dist2, code) of every surfel of gt measured against pred's surface and vice versa (unordered), + border counts"""
    D, H, W = gt.shape
    dev = gt.device
    corners = (D + 1) * (H + 1) * (W + 1)
    L = lib()
    unit = all(float(v) == 1.0 for v in spacing_mm)
    codes, counts = [], torch.zeros(4, dtype=torch.int32, device=dev)
    for i, m in enumerate((gt, pred)):
        code = torch.empty(corners, dtype=torch.uint8, device=dev)
        check(L.b200_surface_codes(m.data_ptr(), D, H, W, code.data_ptr(), counts.data_ptr() + 4 * i, stream()))
        codes.append(code)
    n_gt, n_pred = (int(v) for v in counts[:2].tolist())
    dt = torch.int32 if unit else torch.float64
    sp = None if unit else (C.c_double * 3)(*[float(v) for v in spacing_mm])
    out = []
    for i, (src, other, n_src, n_other) in enumerate(((codes[0], codes[1], n_gt, n_pred), (codes[1], codes[0], n_pred, n_gt))):
        if n_src == 0:
            out.append((None, None))
            continue
        if n_other == 0:
            out.append(("inf", src))
            continue
        d2 = torch.empty(corners, dtype=dt, device=dev)
        scratch = torch.empty(corners, dtype=dt, device=dev)
        check(L.b200_surface_edt(other.data_ptr(), D + 1, H + 1, W + 1, sp, d2.data_ptr(), scratch.data_ptr(), stream()))
        o_d2 = torch.empty(n_src, dtype=dt, device=dev)
        o_code = torch.empty(n_src, dtype=torch.uint8, device=dev)
        check(L.b200_surface_collect(src.data_ptr(), d2.data_ptr(), int(not unit), corners, o_d2.data_ptr(), o_code.data_ptr(),
                                     counts.data_ptr() + 4 * (2 + i), stream()))
        out.append((o_d2, o_code))
    return out, (n_gt, n_pred), unit


def compute_surface_distances(mask_gt, mask_pred, spacing_mm):
    """segmentation/metrics.py:25-178.  mask_gt / mask_pred: 3-D arrays or CUDA tensors (bool or 0/1 uint8, as validate_dsc_asd passes
    them; any non-zero voxel is inside).  Returns the reference's dict of four numpy float64 arrays, sorted by (distance, area).
    Distances are bit-exact for unit spacing (exact integer squared distances, one fp64 sqrt); for other spacings they are the
    fp64 evaluation of sqrt(sum((delta * spacing)^2)) at the exact nearest surface point."""
    dev = mask_gt.device if torch.is_tensor(mask_gt) and mask_gt.is_cuda else (mask_pred.device if torch.is_tensor(mask_pred) and mask_pred.is_cuda else torch.device("cuda"))
    to_dev = lambda m: _u8(m if torch.is_tensor(m) else torch.from_numpy(np.ascontiguousarray(np.asarray(m) != 0)).to(dev))
    gt, pred = to_dev(mask_gt), to_dev(mask_pred)
    if gt.dim() != 3 or gt.shape != pred.shape:
        raise ValueError(f"masks must be 3-D and of the same shape, got {tuple(gt.shape)} and {tuple(pred.shape)}")
    empty = {"distances_gt_to_pred": np.array([]), "distances_pred_to_gt": np.array([]), "surfel_areas_gt": np.array([]), "surfel_areas_pred": np.array([])}
    areas = neighbour_code_to_surface_area(spacing_mm)
    lists, (n_gt, n_pred), unit = _surface_lists(gt, pred, spacing_mm)
    if n_gt == 0 and n_pred == 0:
        return empty
    uniq = np.unique(areas)                                        # ascending distinct areas: rank = order-preserving 8-bit code
    rank = torch.from_numpy(np.searchsorted(uniq, areas).astype(np.int64)).to(dev)
    uniq_dev = torch.from_numpy(uniq).to(dev)
    res = []
    for d2, code in lists:
        if d2 is None:
            res.append((np.array([]), np.array([])))
            continue
        if isinstance(d2, str):                                    # the other mask is empty: every distance is inf (metrics.py:148-149)
            a = uniq_dev[rank[code[(code != 0) & (code != 255)].long()]].sort().values
            res.append((np.full(a.numel(), np.inf), a.cpu().numpy()))
            continue
        r = rank[code.long()]
        if unit:                                                   # one integer key: (squared distance, area rank) -> lexicographic order
            key = (d2.long() << 8) | r
            key = key.sort().values
            dist = (key >> 8).double().sqrt()
            area = uniq_dev[key & 255]
        else:
            dist = d2.sqrt()
            o1 = torch.sort(r, stable=True).indices
            o2 = torch.sort(dist[o1], stable=True).indices
            order = o1[o2]
            dist, area = dist[order], uniq_dev[r[order]]
        res.append((dist.cpu().numpy(), area.cpu().numpy()))
    return {"distances_gt_to_pred": res[0][0], "distances_pred_to_gt": res[1][0], "surfel_areas_gt": res[0][1], "surfel_areas_pred": res[1][1]}


def compute_average_surface_distance(surface_distances):
    """metrics.py:181-207"""
    d_gp, d_pg = surface_distances["distances_gt_to_pred"], surface_distances["distances_pred_to_gt"]
    a_g, a_p = surface_distances["surfel_areas_gt"], surface_distances["surfel_areas_pred"]
    return (np.sum(d_gp * a_g) / np.sum(a_g), np.sum(d_pg * a_p) / np.sum(a_p))


def compute_robust_hausdorff(surface_distances, percent):
    """metrics.py:210-250"""
    out = []
    for d, a in ((surface_distances["distances_gt_to_pred"], surface_distances["surfel_areas_gt"]),
                 (surface_distances["distances_pred_to_gt"], surface_distances["surfel_areas_pred"])):
        if len(d) > 0:
            cum = np.cumsum(a) / np.sum(a)
            idx = np.searchsorted(cum, percent / 100.0)
            out.append(d[min(idx, len(d) - 1)])
        else:
            out.append(np.inf)
    return max(out)


def compute_surface_overlap_at_tolerance(surface_distances, tolerance_mm):
    """metrics.py:253-281"""
    d_gp, d_pg = surface_distances["distances_gt_to_pred"], surface_distances["distances_pred_to_gt"]
    a_g, a_p = surface_distances["surfel_areas_gt"], surface_distances["surfel_areas_pred"]
    return (np.sum(a_g[d_gp <= tolerance_mm]) / np.sum(a_g), np.sum(a_p[d_pg <= tolerance_mm]) / np.sum(a_p))


def compute_surface_dice_at_tolerance(surface_distances, tolerance_mm):
    """metrics.py:284-309"""
    d_gp, d_pg = surface_distances["distances_gt_to_pred"], surface_distances["distances_pred_to_gt"]
    a_g, a_p = surface_distances["surfel_areas_gt"], surface_distances["surfel_areas_pred"]
    return (np.sum(a_g[d_gp <= tolerance_mm]) + np.sum(a_p[d_pg <= tolerance_mm])) / (np.sum(a_g) + np.sum(a_p))


def calculate_metrics(surface, prediction):
    """segmentation/routine.py:206-214 -> (dsc, asd_mean, asd_std, iou); the two "asd" values are the reference's names for the
    (gt -> pred, pred -> gt) average surface distances"""
    dsc, iou = calculate_overlap(surface, prediction)
    asd_mean, asd_std = compute_average_surface_distance(compute_surface_distances(surface, prediction, spacing_mm=(1, 1, 1)))
    return dsc, asd_mean, asd_std, iou
