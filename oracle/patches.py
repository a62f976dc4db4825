"""Sliding-window mirrored patch extraction -- CPU restatement (TEST INFRASTRUCTURE).

Follows detection/patch_utils.py: `get_only_patches` :142-191 and
`get_all_patches_and_labels` :17-140.  The reference grows its output with O(P^2)
np.concatenate calls; here the same decisions are taken first as an integer *plan*
(one row per emitted patch, in the reference's emission order) and the float data is
then gathered once.  Index decisions are exact; the gather is a pure copy, so the
result is bit-identical to the reference's.

Geometry (patch_utils.py:145-147): for axial index i the working slice is
`rot90(vol[:, :, i])`, i.e. S[r, c] = vol[c, Y-1-r, i] with shape (Y, X).
A patch is (2, h, w): channel 0 = S[row0:row0+h, c0:c0+w]; channel 1 is read right to
left starting at column c1 (the left-right mirror), i.e. S[row0:row0+h, c1-k], k=0..w-1.
"""
from __future__ import annotations

import numpy as np

# plan columns
SLICE, ROW0, C0, C1, LABEL = range(5)


def _strip_decision(gm_strip):
    """patch_utils.py:153-160: skip all-zero strips; first column with a positive column sum."""
    if gm_strip.sum() == 0.0:
        return None
    start = int((gm_strip.sum(0) > 0).argmax())
    if start == 0:
        raise AssertionError("start_idx != 0")                      # :160
    return start


def _emit(plan, i, row0, c0, c1, label, nrows, h):
    if nrows != h:
        # the reference would fail in np.concatenate with mismatching (2, nrows, w) vs (2, h, w)
        raise ValueError(f"truncated strip ({nrows} rows) would be emitted at slice {i}, row {row0}")
    plan.append((i, row0, c0, c1, int(label)))


def patch_plan(gmpm, mask=None, h=16, w=32, upsample=None):
    """Integer plan of the patches the reference emits, in its order.

    mask=None            -> get_only_patches (:142-191), labels all 0
    mask given           -> first pass of get_all_patches_and_labels (:20-76) plus, when
                            `upsample` is not False, the positive-only passes at row
                            offsets k=1..h-1 (:79-137).
    """
    X, Y, Z = gmpm.shape
    mid = X // 2 - w                                                 # :158
    plan = []
    if upsample is None:
        upsample = mask is not None

    def strip(i, row0, positives_only):
        S = np.rot90(gmpm[:, :, i])[row0:row0 + h]
        start = _strip_decision(S)
        if start is None:
            return
        nrows = S.shape[0]
        M = None if mask is None else np.rot90(mask[:, :, i])[row0:row0 + h]

        def lab(c0):
            if M is None:
                return False
            return bool(M[:, c0:c0 + w].sum() > 0)                  # :167,:172,:183,:188

        cands = []
        if start < mid:                                              # :173 side patches
            cands.append((start, X - 1 - start))                    # patch_1 :163-166
            cands.append((X - start - w, start + w - 1))            # patch_2 :168-171
        cands.append((mid, X - 1 - mid))                            # patch_3 :179-182
        cands.append((X - mid - w, mid + w - 1))                    # patch_4 :184-187
        for c0, c1 in cands:
            l = lab(c0)
            if positives_only and not l:
                continue
            _emit(plan, i, row0, c0, c1, l, nrows, h)

    for i in range(Z):
        for j in range(0, Y, h):
            strip(i, j, False)
    if upsample:
        for k in range(1, h):                                        # :79
            for i in range(Z):
                for j in range(0, Y - h, h):                         # :85
                    strip(i, k + j, True)
    return np.asarray(plan, dtype=np.int64).reshape(-1, 5)


def gather_patches(target, plan, h=16, w=32):
    """Copy the planned windows out of `target` (X,Y,Z).  Returns (P,2,h,w), target's dtype."""
    X, Y, Z = target.shape
    P = plan.shape[0]
    out = np.empty((P, 2, h, w), dtype=target.dtype)
    if P == 0:
        return out
    rr = np.arange(h)[None, :, None]
    cc = np.arange(w)[None, None, :]
    i = plan[:, SLICE][:, None, None]
    y = Y - 1 - (plan[:, ROW0][:, None, None] + rr)
    out[:, 0] = target[plan[:, C0][:, None, None] + cc, y, i]
    out[:, 1] = target[plan[:, C1][:, None, None] - cc, y, i]
    return out


def get_only_patches(target_np, gmpm, h=16, w=32):
    """patch_utils.py:142-191."""
    return gather_patches(target_np, patch_plan(gmpm, None, h, w), h, w)


def get_all_patches_and_labels(target_np, gmpm, mask_np, h=16, w=32):
    """patch_utils.py:17-140 -> (patches (P,2,h,w) float64, labels (P,) bool)."""
    plan = patch_plan(gmpm, mask_np, h, w)
    return gather_patches(target_np, plan, h, w), plan[:, LABEL].astype(bool)


def minmax_normalise(vol):
    """patch_utils.py:196."""
    return (vol - vol.min()) / (vol.max() - vol.min())


def read_nifti1_f32(path):
    """Minimal NIfTI-1 reader for the shipped GM template
    (detection/MNI152_T1_1mm_brain_gray.nii.gz): gzip, 352-byte header, float32, F-order."""
    import gzip
    import struct
    raw = gzip.open(path, "rb").read()
    dim = struct.unpack("<8h", raw[40:56])
    datatype = struct.unpack("<h", raw[70:72])[0]
    vox_offset = int(struct.unpack("<f", raw[108:112])[0])
    assert datatype == 16, datatype
    shape = dim[1:1 + dim[0]]
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype="<f4", count=n, offset=vox_offset)
    return data.reshape(shape, order="F")
