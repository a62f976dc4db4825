"""Sliding-grid patch inference and random-patch sampling (row f-3) -- CPU restatement (TEST INFRASTRUCTURE).

PARITY UNPINNED.  The reference does this through the third-party package `torchio` (`torchio.inference.GridSampler`,
`torchio.inference.GridAggregator`, `torchio.Queue(sampler_class=torchio.sampler.ImageSampler)`), which is neither installed here
nor present under /root/reference, and whose version is pinned nowhere.  The call sites fix the API generation:
  * segmentation/pretraining_3d_unet.ipynb [cell 26, 35]: `GridSampler(sample, patch_size=(64,)*3, patch_overlap=(4,)*3)`,
    `GridAggregator(sample, patch_overlap)`, `aggregator.add_batch(labels, locations)`, `get_output_tensor()` used as `predicted[0]`;
  * segmentation/routine.py:150-178: `torchio.Queue(..., sampler_class=torchio.sampler.ImageSampler, ...)`
-- `sampler_class=` / `ImageSampler` and a sampler that takes the sample exist only in torchio <= 0.16 (spring 2020; 0.17 replaced
them by `sampler=UniformSampler(...)`), so the algorithm restated here is the published one of those releases, itself adapted from
NiftyNet: window starts every `patch - 2*overlap` voxels plus a last window flush with the volume end (plus a middle window when
only two starts result), and an aggregator that crops `overlap` voxels from EVERY side of each predicted window before writing it,
later windows overwriting earlier ones; voxels no cropped window reaches (the outer `overlap` shell) stay 0.
The anchors are the reference's call sites above and the properties the tests check (every interior voxel written, exact
copy semantics, last-writer-wins order); tests/golden/grid_kat6_UNPINNED.npz is a regression vector made by THIS file.
"""
from __future__ import annotations

import numpy as np


def enumerate_step_points(starting, ending, win_size, step_size):
    """torchio <= 0.16 `GridSampler._enumerate_step_points` (NiftyNet `enumerate_step_points`)."""
    starting, ending = max(int(starting), 0), max(int(ending), 0)
    win_size, step_size = max(int(win_size), 1), max(int(step_size), 1)
    if starting > ending:
        starting, ending = ending, starting
    points = []
    while starting + win_size <= ending:
        points.append(starting)
        starting += step_size
    points.append(max(ending - win_size, 0))
    points = np.unique(points).flatten()
    if len(points) == 2:                       # too few samples: one more in the middle
        points = np.append(points, np.round(np.mean(points)))
    _, uniq = np.unique(points, return_index=True)
    return points[np.sort(uniq)]


def grid_spatial_coordinates(shape, window_shape, border):
    """torchio <= 0.16 `GridSampler._grid_spatial_coordinates` -> int32 (L, 6) rows (i0, j0, k0, i1, j1, k1), in ITS order
    (np.meshgrid default 'xy' indexing, then reshape((3, -1)).T)."""
    num_dims = len(shape)
    grid_size = [max(w - 2 * b, 0) for w, b in zip(window_shape, border)]
    steps = [enumerate_step_points(0, shape[i], window_shape[i], grid_size[i]) for i in range(num_dims)]
    starting = np.asanyarray(np.meshgrid(*steps)).reshape((num_dims, -1)).T
    coords = np.zeros((starting.shape[0], num_dims * 2), dtype=np.int32)
    coords[:, :num_dims] = starting
    for idx in range(num_dims):
        coords[:, num_dims + idx] = starting[:, idx] + window_shape[idx]
    assert np.all(np.max(coords, axis=0)[num_dims:] <= np.asarray(shape)[:num_dims]), "window larger than the volume"
    return coords


def extract_patches(volume, locations):
    """GridSampler.__getitem__ for every location: volume (C, D, H, W) -> (L, C, pd, ph, pw)."""
    return np.stack([volume[:, a:d, b:e, c:f] for a, b, c, d, e, f in locations])


def crop_batch(windows, location, border):
    """torchio <= 0.16 `GridAggregator.crop_batch`."""
    if not border or not any(border):
        return windows, location
    location = location.astype(int)
    spatial_shape = np.array(windows.shape[2:])
    for idx in range(3):
        location[:, idx] = location[:, idx] + border[idx]
        location[:, idx + 3] = location[:, idx + 3] - border[idx]
    if np.any(location < 0):
        return windows, location
    cropped_shape = np.max(location[:, 3:6] - location[:, 0:3], axis=0)
    left = np.floor((spatial_shape - cropped_shape) / 2).astype(int)
    i0, j0, k0 = left
    i1, j1, k1 = left + cropped_shape
    return windows[:, :, i0:i1, j0:j1, k0:k1], location


def aggregate(output, windows, locations, border):
    """`GridAggregator.add_batch`: sequential overwrite of the cropped windows (batch, 1, pd, ph, pw) into `output` (D, H, W)."""
    cropped, _ = crop_batch(windows, np.copy(locations), border)
    _, locs = crop_batch(np.ones_like(windows), np.copy(locations), border)
    for window, loc in zip(cropped, locs):
        i0, j0, k0, i1, j1, k1 = loc
        output[i0:i1, j0:j1, k0:k1] = window.squeeze(0) if window.shape[0] == 1 else window
    return output


def random_indices(shape, patch_size, randint):
    """torchio <= 0.16 `ImageSampler.get_random_indices`: per dimension `randint(max_start)` with max_start = size - patch
    EXCLUSIVE (the last valid start is never drawn; 0 when the patch spans the dimension).  `randint(n)` is the caller's
    source of uniform integers in [0, n) -- torch.randint(n, size=(1,)).item() in torchio."""
    ini = []
    for s, p in zip(shape, patch_size):
        if s - p < 0:
            raise ValueError(f"Patch size {tuple(patch_size)} must not be larger than image size {tuple(shape)}")
        ini.append(0 if s - p == 0 else int(randint(int(s - p))))
    return np.array(ini), np.array(ini) + np.array(patch_size)
