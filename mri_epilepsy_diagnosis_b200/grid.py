"""Sliding-grid patch inference and random-patch training queue on the GPU (row f-3).

Drop-in for the three torchio classes the reference uses around its 3-D U-Net:

  GridSampler(sample, patch_size, patch_overlap)      segmentation/pretraining_3d_unet.ipynb [cell 26, 35]
  GridAggregator(sample, patch_overlap)               .add_batch(labels, locations) / .get_output_tensor()
  Queue(subjects, max_length, samples_per_volume, patch_size, shuffle_subjects, shuffle_patches)
                                                      segmentation/routine.py:150-178 (torchio.Queue + torchio.sampler.ImageSampler)

torchio is third-party and absent (version unpinned); the window placement, border cropping, overwrite order and random-start
rule follow its published <= 0.16 algorithm (the API generation the call sites use), see oracle/grid.py.  Volumes stay on the
device: the window gather and the label scatter are one kernel launch each (b200_grid_gather / b200_grid_aggregate) instead of a
Python loop over windows with a host round trip per batch.

A `sample` is what the reference's loops index: `{'MRI': {'data': tensor (C,D,H,W)}, 'LABEL': {'data': ...}}`; a bare (C,D,H,W) /
(D,H,W) tensor is accepted as `{'MRI': {'data': t}}`.
"""
from __future__ import annotations

import numpy as np
import torch

from ._cabi import check, lib, need_cuda, stream

DATA, LOCATION = "data", "location"


def _tuple3(v):
    if isinstance(v, (int, np.integer)):
        return (int(v),) * 3
    v = tuple(int(a) for a in v)
    if len(v) != 3:
        raise ValueError(f"expected an int or 3 values, got {v}")
    return v


def _as_sample(sample):
    if isinstance(sample, dict):
        return sample
    t = sample if torch.is_tensor(sample) else torch.as_tensor(np.asarray(sample))
    return {"MRI": {DATA: t if t.dim() == 4 else t[None]}}


def _images(sample):
    return {k: v for k, v in sample.items() if isinstance(v, dict) and DATA in v}


def _spatial_shape(sample):
    shapes = {tuple(v[DATA].shape[-3:]) for v in _images(sample).values()}
    if len(shapes) != 1:
        raise ValueError(f"images of a sample must share their spatial shape, got {shapes}")
    return shapes.pop()


# ----------------------------------------------------------------------------------------------- window placement (host, integers)
def _step_points(ending, win_size, step_size):
    ending, win_size, step_size = max(int(ending), 0), max(int(win_size), 1), max(int(step_size), 1)
    pts, s = [], 0
    while s + win_size <= ending:
        pts.append(s)
        s += step_size
    pts.append(max(ending - win_size, 0))
    pts = sorted(set(pts))
    if len(pts) == 2:                                   # two windows only: a third one half-way (np.round: half to even)
        pts.append(int(np.round((pts[0] + pts[1]) / 2.0)))
    seen, out = set(), []
    for p in pts:                                       # first occurrence order
        if p not in seen:
            seen.add(p)
            out.append(p)
    return out


def grid_locations(shape, patch_size, patch_overlap):
    """int32 (L, 6) window corners (i0, j0, k0, i1, j1, k1) in the reference library's order: starts every `patch - 2*overlap`
    voxels + one window flush with the end (+ a middle one when that leaves just two); first axis varies second-fastest,
    second axis slowest, third axis fastest (numpy.meshgrid's default 'xy' indexing followed by reshape((3, -1)).T)."""
    shape, patch, border = _tuple3(shape), _tuple3(patch_size), _tuple3(patch_overlap)
    if any(p > s for p, s in zip(patch, shape)):
        raise AssertionError(f"window size {patch} larger than the volume {shape}")
    pts = [_step_points(shape[i], patch[i], max(patch[i] - 2 * border[i], 0)) for i in range(3)]
    rows = [(a, b, c) for b in pts[1] for a in pts[0] for c in pts[2]]
    ini = np.asarray(rows, dtype=np.int32).reshape(-1, 3)
    return np.hstack([ini, ini + np.asarray(patch, dtype=np.int32)]).astype(np.int32)


def _gather(vol, loc_dev, patch):
    """vol (C,D,H,W) CUDA tensor -> (L,C,pd,ph,pw), one launch"""
    need_cuda(vol, "grid_gather")
    vol = vol.contiguous()
    C, D, H, W = vol.shape
    L = loc_dev.shape[0]
    out = torch.empty((L, C) + tuple(patch), dtype=vol.dtype, device=vol.device)
    check(lib().b200_grid_gather(vol.element_size(), vol.data_ptr(), loc_dev.data_ptr(), L, C, D, H, W, patch[0], patch[1], patch[2],
                                 out.data_ptr(), stream()))
    return out


class GridSampler:
    """All windows of one volume.  `sampler[i]` is the i-th cropped sample (+ 'location'); `batches(n)` yields what a
    DataLoader(grid_sampler, batch_size=n) would collate: {'MRI': {'data': (B,C,p,p,p)}, ..., 'location': (B,6) int64}."""

    def __init__(self, sample, patch_size, patch_overlap, device="cuda"):
        self.sample = _as_sample(sample)
        self.patch_size, self.patch_overlap = _tuple3(patch_size), _tuple3(patch_overlap)
        self.locations = grid_locations(_spatial_shape(self.sample), self.patch_size, self.patch_overlap)
        self.device = torch.device(device)
        self._loc_dev = torch.from_numpy(self.locations).to(self.device)
        self._cache = {}

    def __len__(self):
        return len(self.locations)

    def _patches(self, name):
        if name not in self._cache:
            vol = self.sample[name][DATA]
            vol = vol.to(self.device) if torch.is_tensor(vol) else torch.as_tensor(np.asarray(vol), device=self.device)
            self._cache[name] = _gather(vol if vol.dim() == 4 else vol[None], self._loc_dev, self.patch_size)
        return self._cache[name]

    def __getitem__(self, index):
        if not -len(self) <= index < len(self):
            raise IndexError(index)
        out = {k: {**v, DATA: self._patches(k)[index]} for k, v in _images(self.sample).items()}
        out[LOCATION] = self.locations[index]
        return out

    def batches(self, batch_size):
        loc64 = torch.from_numpy(self.locations.astype(np.int64))
        for a in range(0, len(self), batch_size):
            b = {k: {DATA: self._patches(k)[a:a + batch_size]} for k in _images(self.sample)}
            b[LOCATION] = loc64[a:a + batch_size]
            yield b


class GridAggregator:
    """Scatter of per-window label maps back into the volume: `overlap` voxels are cropped from every side of each window,
    later windows overwrite earlier ones (in GridSampler order), voxels no cropped window reaches stay 0."""

    def __init__(self, sample, patch_overlap, device="cuda"):
        self.shape = _tuple3(sample) if isinstance(sample, (tuple, list)) else _spatial_shape(_as_sample(sample))
        self.patch_overlap = _tuple3(patch_overlap)
        self.device = torch.device(device)
        self._labels, self._locs = [], []
        self._output = None

    def add_batch(self, windows, locations):
        """windows (B, 1, pd, ph, pw) label maps (any integer / float dtype), locations (B, 6)"""
        if windows.dim() != 5 or windows.shape[1] != 1:
            raise ValueError(f"add_batch expects (batch, 1, d, h, w) windows, got {tuple(windows.shape)}")
        locations = torch.as_tensor(np.asarray(locations) if not torch.is_tensor(locations) else locations)
        if locations.shape != (windows.shape[0], 6):
            raise ValueError("one (i0, j0, k0, i1, j1, k1) row per window expected")
        self._labels.append(windows.to(self.device))
        self._locs.append(locations.to(torch.int32))
        self._output = None

    def get_output_tensor(self):
        """(1, D, H, W) tensor of the windows' dtype"""
        if self._output is None:
            if not self._labels:
                return torch.zeros((1,) + self.shape, dtype=torch.uint8, device=self.device)
            labels = torch.cat(self._labels).contiguous()
            locs = torch.cat(self._locs).to(self.device).contiguous()
            need_cuda(labels, "grid_aggregate")
            pd, ph, pw = labels.shape[2:]
            ext = (locs[:, 3:] - locs[:, :3]).cpu()
            if not bool((ext == torch.tensor([pd, ph, pw], dtype=torch.int32)).all()):
                raise ValueError("locations do not match the window size")
            out = torch.zeros((1,) + self.shape, dtype=labels.dtype, device=self.device)
            b = self.patch_overlap
            check(lib().b200_grid_aggregate(labels.element_size(), labels.data_ptr(), locs.data_ptr(), labels.shape[0], self.shape[0], self.shape[1],
                                            self.shape[2], pd, ph, pw, b[0], b[1], b[2], out.data_ptr(), stream()))
            self._output = out
        return self._output


def sliding_window_labels(model, sample, patch_size=64, patch_overlap=4, batch_size=16, image="MRI", channels_dimension=1):
    """The inference loop of pretraining_3d_unet.ipynb [cell 26]: windows -> model -> argmax(dim=1, keepdim=True) -> aggregate.
    Returns the (1, D, H, W) uint8 label volume (on the device)."""
    sampler = GridSampler(sample, patch_size, patch_overlap)
    aggregator = GridAggregator(sampler.sample, patch_overlap)
    with torch.no_grad():
        for batch in sampler.batches(batch_size):
            logits = model(batch[image][DATA])
            labels = logits.argmax(dim=channels_dimension, keepdim=True)
            aggregator.add_batch(labels.to(torch.uint8), batch[LOCATION])
    return aggregator.get_output_tensor()


# ----------------------------------------------------------------------------------------------- random-patch training queue
def random_patch_locations(shape, patch_size, n, generator=None):
    """n windows with starts drawn like ImageSampler.get_random_indices: per dimension torch.randint(size - patch) -- the upper
    bound is EXCLUSIVE, so the last valid start is never drawn; 0 when the patch spans the dimension.  Draw order: window by
    window, dimension by dimension (one scalar draw each, like the reference library)."""
    shape, patch = _tuple3(shape), _tuple3(patch_size)
    if any(p > s for p, s in zip(patch, shape)):
        raise ValueError(f"Patch size {patch} must not be larger than image size {shape}")
    ini = np.zeros((n, 3), dtype=np.int32)
    for i in range(n):
        for d in range(3):
            m = shape[d] - patch[d]
            ini[i, d] = 0 if m == 0 else int(torch.randint(m, size=(1,), generator=generator).item())
    return np.hstack([ini, ini + np.asarray(patch, dtype=np.int32)]).astype(np.int32)


class Queue:
    """torchio.Queue(subjects_dataset, max_length, samples_per_volume, patch_size, sampler_class=ImageSampler, shuffle_subjects,
    shuffle_patches) for device-resident subjects: an iterable over single-patch samples (`len(queue)` = subjects x
    samples_per_volume, what the reference's DataLoader sees); `batches(n)` collates n of them.  The patch list is refilled,
    `max_length // samples_per_volume` subjects at a time, when it runs empty (the reference library's policy); all patches of
    a refill are gathered by ONE kernel launch per image instead of one crop per patch in loader worker processes."""

    def __init__(self, subjects_dataset, max_length, samples_per_volume, patch_size, shuffle_subjects=True, shuffle_patches=True,
                 device="cuda", generator=None):
        self.subjects = [_as_sample(s) for s in subjects_dataset]
        self.max_length, self.samples_per_volume = int(max_length), int(samples_per_volume)
        self.patch_size = _tuple3(patch_size)
        self.shuffle_subjects, self.shuffle_patches = bool(shuffle_subjects), bool(shuffle_patches)
        self.device, self.generator = torch.device(device), generator

    def __len__(self):
        return len(self.subjects) * self.samples_per_volume

    def _subject_order(self):
        n = len(self.subjects)
        return torch.randperm(n, generator=self.generator).tolist() if self.shuffle_subjects else list(range(n))

    def __iter__(self):
        order = self._subject_order()
        per_fill = max(1, self.max_length // self.samples_per_volume)
        for a in range(0, len(order), per_fill):
            patches = []
            for si in order[a:a + per_fill]:
                s = self.subjects[si]
                locs = random_patch_locations(_spatial_shape(s), self.patch_size, self.samples_per_volume, self.generator)
                loc_dev = torch.from_numpy(locs).to(self.device)
                got = {}
                for k, v in _images(s).items():
                    vol = v[DATA].to(self.device)
                    got[k] = _gather(vol if vol.dim() == 4 else vol[None], loc_dev, self.patch_size)
                for i in range(self.samples_per_volume):
                    item = {k: {DATA: g[i]} for k, g in got.items()}
                    item["index_ini"] = locs[i, :3]
                    patches.append(item)
            if self.shuffle_patches:
                perm = torch.randperm(len(patches), generator=self.generator).tolist()
                patches = [patches[i] for i in perm]
            yield from patches

    def batches(self, batch_size):
        buf = []
        for item in self:
            buf.append(item)
            if len(buf) == batch_size:
                yield self._collate(buf)
                buf = []
        if buf:
            yield self._collate(buf)

    @staticmethod
    def _collate(items):
        out = {k: {DATA: torch.stack([it[k][DATA] for it in items])} for k in items[0] if k != "index_ini"}
        out["index_ini"] = torch.as_tensor(np.stack([it["index_ini"] for it in items]).astype(np.int64))
        return out
