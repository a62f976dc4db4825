"""Sliding-window mirrored patch extraction on the GPU -- drop-in for detection/patch_utils.py.

Same names, argument meaning, output order/dtype and error behaviour as the reference:
  get_only_patches(target_np, gmpm, h=16, w=32)                      patch_utils.py:142-191
  get_all_patches_and_labels(target_np, gmpm, mask_np, h=16, w=32)   patch_utils.py:17-140
  get_image_patches(input_img, input_mask=None, h=16, w=32, gmpm=...) patch_utils.py:193-205 (min-max normalisation + either of the above)
Inputs may be numpy arrays (copied to the GPU) or CUDA tensors of shape (X,Y,Z); outputs are CUDA tensors
((P,2,h,w) float64, labels bool).  `patch_plan` exposes the integer decisions (bit-exact with the reference).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._cabi import PatchDesc, check, lib, stream


def _dev(a, dtype, device):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    return a.to(device=device, dtype=dtype).contiguous()


def patch_plan(gmpm, mask=None, h=16, w=32, upsample=None, device="cuda"):
    """-> int32 CUDA tensor (P,5): slice, row0, c0, c1, label -- in the reference's emission order."""
    gm = _dev(gmpm, torch.float64, device)
    if gm.dim() != 3:
        raise ValueError("gmpm must be a 3-D volume")
    X, Y, Z = gm.shape
    if upsample is None:
        upsample = mask is not None
    m = None
    if mask is not None:
        m = _dev(mask, torch.bool, device).to(torch.uint8)
        if tuple(m.shape) != (X, Y, Z):
            raise ValueError("mask and template shapes differ")
    pd = PatchDesc(X, Y, Z, h, w, int(mask is not None), int(bool(upsample)))
    L = lib()
    rows = int(L.b200_patch_max_rows(C.byref(pd)))
    plan = torch.empty((max(rows, 1), 5), dtype=torch.int32, device=gm.device)
    meta = torch.zeros(2, dtype=torch.int32, device=gm.device)           # count, status
    nws = L.b200_patch_workspace_bytes(C.byref(pd))
    ws = torch.empty(nws, dtype=torch.uint8, device=gm.device)
    check(L.b200_patch_plan(C.byref(pd), gm.data_ptr(), None if m is None else m.data_ptr(), plan.data_ptr(), meta.data_ptr(),
                            meta.data_ptr() + 4, ws.data_ptr(), nws, stream()))
    count, status = (int(v) for v in meta.tolist())
    if status & 1:
        raise AssertionError("start_idx != 0")                            # patch_utils.py:160
    if status & 2:
        # the reference fails in np.concatenate when a ragged (< h rows) strip has to be emitted
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    return plan[:count]


def gather(target, plan, h=16, w=32, dtype=torch.float64):
    t = _dev(target, torch.float64, plan.device)
    X, Y, Z = t.shape
    P = plan.shape[0]
    out = torch.empty((P, 2, h, w), dtype=dtype, device=plan.device)
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("gather: dtype must be float64 or float32")
    pd = PatchDesc(X, Y, Z, h, w, 0, 0)
    check(lib().b200_patch_gather(C.byref(pd), t.data_ptr(), plan.contiguous().data_ptr(), P, int(dtype == torch.float32), out.data_ptr(), stream()))
    return out


def get_only_patches(target_np, gmpm, h=16, w=32, device="cuda"):
    plan = patch_plan(gmpm, None, h, w, device=device)
    return gather(target_np, plan, h, w)


def get_all_patches_and_labels(target_np, gmpm, mask_np, h=16, w=32, device="cuda"):
    plan = patch_plan(gmpm, mask_np, h, w, device=device)
    return gather(target_np, plan, h, w), plan[:, 4].bool()


def load_nifti(path):
    """float64 array of a NIfTI-1 file (.nii / .nii.gz), what `nib.load(path).get_fdata()` returns for the files the reference reads:
    single-file NIfTI-1, little endian, scl_slope / scl_inter applied when set.  (nibabel is not a dependency of this package.)"""
    import gzip
    import struct
    with (gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")) as f:
        raw = f.read()
    if struct.unpack("<i", raw[0:4])[0] != 348:
        raise ValueError(f"{path}: not a little-endian NIfTI-1 file")
    dim = struct.unpack("<8h", raw[40:56])
    datatype = struct.unpack("<h", raw[70:72])[0]
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    slope, inter = struct.unpack("<2f", raw[112:120])
    dtypes = {2: "<u1", 4: "<i2", 8: "<i4", 16: "<f4", 64: "<f8", 256: "<i1", 512: "<u2", 768: "<u4"}
    if datatype not in dtypes:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    shape = dim[1:1 + dim[0]]
    data = np.frombuffer(raw, dtype=dtypes[datatype], count=int(np.prod(shape)), offset=vox_offset).reshape(shape, order="F").astype(np.float64)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope) and np.isfinite(inter):
            data = data * np.float64(slope) + np.float64(inter)
    return data


def minmax_normalize(target, device="cuda"):
    """(t - t.min()) / (t.max() - t.min()) in float64 on the device -- patch_utils.py:196, bit-exact with numpy"""
    t = _dev(target if not isinstance(target, str) else load_nifti(target), torch.float64, device)
    out = torch.empty_like(t)
    nws = lib().b200_minmax_workspace_bytes()
    ws = torch.empty(nws, dtype=torch.uint8, device=t.device)
    check(lib().b200_minmax_normalize(t.data_ptr(), t.numel(), out.data_ptr(), ws.data_ptr(), nws, stream()))
    return out


def get_image_patches(input_img_name, input_mask_name=None, h=16, w=32, gmpm=None, device="cuda"):
    """detection/patch_utils.py:193-205.  `input_img_name` / `input_mask_name`: NIfTI paths (as in the reference) or arrays / tensors
    already in memory; `gmpm`: the grey-matter template the reference reads from a notebook global.  Returns (patches (P,2,h,w)
    float64, labels (P,) bool) as CUDA tensors."""
    if gmpm is None:
        raise ValueError("get_image_patches: pass the grey-matter template as gmpm= (a module-level global in the reference)")
    target = minmax_normalize(input_img_name, device)
    if input_mask_name is not None:
        m = load_nifti(input_mask_name) if isinstance(input_mask_name, str) else input_mask_name
        m = (m > 0) if not torch.is_tensor(m) else (m > 0)
        return get_all_patches_and_labels(target, gmpm, m, h=h, w=w, device=device)
    patches = get_only_patches(target, gmpm, h=h, w=w, device=device)
    return patches, torch.zeros(patches.shape[0], dtype=torch.bool, device=patches.device)
