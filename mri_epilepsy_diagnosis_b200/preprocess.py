"""Intensity preprocessing on the device: the collate-time histogram standardisation of
classification/train_ENC_CLF.ipynb [cell 9] -- `normalize`, `default_collate`, `_get_percentiles`, `_standardize_cutoff` -- and
the crop of utils/data.py:16-30 (`reshape_image`).  Same names, arguments and error behaviour; tensors live on the GPU and the
percentiles come from an exact radix select (csrc/preprocess.cuh), so a batch never goes back to the host.

The reference spends 16 s per iteration here (train_ENC_CLF.ipynb [cell 23]) against < 1 s of compute."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi
from ._cabi import HistStdDesc, check, lib, need_cuda, stream

DEFAULT_CUTOFF = 0.01, 0.99                       # train_ENC_CLF.ipynb [cell 9]
STANDARD_RANGE = 0, 100
RANGE_TO_USE = (0, 1, 2, 4, 5, 6, 7, 8, 10, 11, 12)


def _standardize_cutoff(cutoff):
    cutoff = np.asarray(cutoff, dtype=np.float64).copy()
    cutoff[0] = max(0., cutoff[0])
    cutoff[1] = min(1., cutoff[1])
    cutoff[0] = np.min([cutoff[0], 0.09])
    cutoff[1] = np.max([cutoff[1], 0.91])
    return cutoff


def _get_percentiles(percentiles_cutoff):
    quartiles = np.arange(25, 100, 25).tolist()
    deciles = np.arange(10, 100, 10).tolist()
    all_percentiles = list(percentiles_cutoff) + quartiles + deciles
    return np.array(sorted(set(all_percentiles)))


def _desc(landmarks, cutoff, epsilon):
    cutoff_ = DEFAULT_CUTOFF if cutoff is None else cutoff
    percentiles = _get_percentiles(100 * np.array(_standardize_cutoff(cutoff_)))
    mapping = np.asarray(landmarks, dtype=np.float64).reshape(-1)
    if len(percentiles) > 16 or len(mapping) != len(percentiles):
        raise ValueError(f"normalize: {len(mapping)} landmarks for {len(percentiles)} percentiles (the library takes at most 16)")
    if max(RANGE_TO_USE) >= len(percentiles):
        raise IndexError(f"normalize: range_to_use needs 13 percentiles, got {len(percentiles)}")       # the reference's mapping[range_to_use] raises
    d = HistStdDesc()
    # np.percentile divides by a.dtype.type(100): float64 percentiles / float32(100) stays float64
    q = np.true_divide(percentiles.astype(np.float64), np.float32(100))
    for i, v in enumerate(q):
        d.q[i] = float(v)
        d.landmarks[i] = float(mapping[i])
    for i, r in enumerate(RANGE_TO_USE):
        d.range_idx[i] = r
    d.nq, d.nrange, d.eps = len(q), len(RANGE_TO_USE), float(epsilon)
    return d


def _run(tensor, landmarks, mask, cutoff, epsilon, want_out, want_pct):
    need_cuda(tensor, "normalize")
    x = tensor.detach()
    if x.dtype != torch.float32:
        x = x.float()                               # `data.reshape(-1).astype(np.float32)`
    x = x.contiguous()
    n = x.numel()
    m = None
    if mask is not None:
        m = torch.as_tensor(mask, device=x.device).reshape(-1)
        if m.numel() != n:
            raise IndexError(f"boolean index did not match: mask has {m.numel()} elements, data {n}")
        m = (m != 0).to(torch.uint8).contiguous()
    d = _desc(landmarks, cutoff, epsilon)
    out = torch.empty_like(x) if want_out else None
    pct = torch.empty(d.nq, dtype=torch.float64, device=x.device) if want_pct else None
    nws = lib().b200_histstd_workspace_bytes()
    ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
    check(lib().b200_histstd_normalize(C.byref(d), x.data_ptr(), m.data_ptr() if m is not None else None, n, out.data_ptr() if want_out else None,
                                       pct.data_ptr() if want_pct else None, ws.data_ptr(), nws, stream()))
    return out, pct


def normalize(tensor, landmarks, mask=None, cutoff=None, epsilon=1e-5):
    """train_ENC_CLF.ipynb [cell 9] `normalize`: float32 tensor of the input's shape, on the input's device."""
    out, _ = _run(tensor, landmarks, mask, cutoff, epsilon, True, False)
    return out.view(tensor.shape)


def percentile_values(tensor, mask=None, cutoff=None):
    """the 13 landmark percentiles `np.percentile(data[mask], percentiles)` of the same cell (float64, on the device)"""
    _, pct = _run(tensor, np.zeros(13), mask, cutoff, 1e-5, False, True)
    return pct


def default_collate(batch, landmarks):
    """[cell 9] `default_collate`: batch = list of (X, y, domain); X tensors on the GPU.  `landmarks` replaces the
    `np.load('fcd_train_data_landmarks.npy')` the reference performs on every call."""
    X = torch.stack([normalize(item[0], landmarks) for item in batch])
    y = torch.tensor([item[1] for item in batch], dtype=torch.long, device=X.device)
    domain = torch.tensor([item[2] for item in batch], dtype=torch.long, device=X.device)
    return X, y, domain


def reshape_image(img, coord_min, img_shape):
    """utils/data.py:16-30: crop `img_shape` starting at `coord_min`, AssertionError when the volume is too small."""
    img_shape = tuple(int(s) for s in img_shape)
    img = img[coord_min[0]:coord_min[0] + img_shape[0], coord_min[1]:coord_min[1] + img_shape[1], coord_min[2]:coord_min[2] + img_shape[2]]
    if tuple(img.shape[:3]) != img_shape:
        raise AssertionError(f"Current image shape: {tuple(img.shape[:3])}, desired image shape: {img_shape}")
    return img.reshape((1,) + img_shape)
