"""FCD mask generation on the GPU -- mirror of `FCDMaskGenerator` (detection/model_utils.py:118-228).

Same method names and results as the reference class, with the patch classifier run on ALL patches of the volume in a few
large batches (the reference classifies 4 752 patches one at a time, model_utils.py:130-134):

    gen = FCDMaskGenerator(model, gmpm)            # model: a converted PatchModel (or any (P,2,h,w) -> (P,2) logits callable)
    mask = gen.get_mask(img_np)                     # int64 (X,Y,Z) CUDA tensor

Everything stays on the device and in the library: sliding-window plan + gather (patch kernels), batched PatchModel inference
(tcgen05 conv kernels), argmax, then three small integer kernels (b200_fcd_scatter_labels / _vote / _paint) for the (4, Y//h, Z)
patch map, the 4-neighbour vote and the paint-back -- no cuDNN dispatch, no eager indexing.

Two behaviours of the reference are reproduced on purpose (they are what `get_mask` returns today; see oracle/detect.py):
`_postprocess` uses int64 0/1 arrays as INDICES (model_utils.py:190-193), so slabs 0 / 1 of the patch map are overwritten
instead of the voted elements, and `_masking` paints `-j:-j-h:-1`, which is empty for the first strip.  Pass `fixed=True` to
`_postprocess` / `get_mask` for the evidently intended boolean-mask vote.
"""
from __future__ import annotations

import torch

from . import patches as P
from ._cabi import check, lib, stream


class FCDMaskGenerator:
    def __init__(self, model, gmpm, h=16, w=32, batch=2048, device="cuda"):
        self.model, self.h, self.w, self.batch = model, h, w, batch
        self.gmpm = P._dev(gmpm, torch.float64, device)
        self.plan = P.patch_plan(self.gmpm, None, h, w, device=device).contiguous()     # (P,5) int32: slice, row0, c0, c1, label

    # ---- model_utils.py:136-180, batched
    def _get_predictions_per_batches(self, img):
        X, Y, Z = self.gmpm.shape
        patches = P.gather(img, self.plan, self.h, self.w, dtype=torch.float32)
        labels = []
        with torch.no_grad():
            for s in range(0, patches.shape[0], self.batch):
                labels.append(torch.argmax(self.model(patches[s:s + self.batch]), dim=1))
        labels = (torch.cat(labels) if labels else torch.zeros(0, dtype=torch.int64, device=self.gmpm.device)).to(torch.int64).contiguous()
        pm = torch.zeros((4, Y // self.h, Z), dtype=torch.int64, device=self.gmpm.device)
        check(lib().b200_fcd_scatter_labels(self.plan.data_ptr(), self.plan.shape[0], labels.data_ptr(), X, Y, Z, self.h, self.w, pm.data_ptr(), stream()))
        return pm

    # ---- model_utils.py:182-193
    def _postprocess(self, img, patch_map_tensor, fixed=False):
        """4-neighbour vote.  fixed=False reproduces the reference: it indexes the map with the int64 0/1 arrays (:190-193), so slabs
        0 / 1 are overwritten instead of the voted elements; fixed=True is the boolean-mask vote.  Returns a new tensor."""
        pm = patch_map_tensor.to(torch.int64).contiguous()
        out = torch.empty_like(pm)
        flags = torch.zeros(4, dtype=torch.int32, device=pm.device)
        check(lib().b200_fcd_vote(pm.data_ptr(), pm.shape[1], pm.shape[2], int(bool(fixed)), out.data_ptr(), flags.data_ptr(), stream()))
        return out

    # ---- model_utils.py:195-216 (same boxes, same assignment order)
    def _masking(self, img, patch_map_tensor):
        X, Y, Z = self.gmpm.shape
        pm = patch_map_tensor.to(torch.int64).contiguous()
        final = torch.zeros((X, Y, Z), dtype=torch.int64, device=self.gmpm.device)
        check(lib().b200_fcd_paint(self.plan.data_ptr(), self.plan.shape[0], pm.data_ptr(), X, Y, Z, self.h, self.w, final.data_ptr(), stream()))
        return final

    def get_mask(self, img, fixed=False):
        """model_utils.py:218-222 -> int64 (X,Y,Z)."""
        img = P._dev(img, torch.float64, self.gmpm.device)
        pm = self._get_predictions_per_batches(img)
        pm = self._postprocess(img, pm, fixed=fixed)
        return self._masking(img, pm).to(torch.int64)

    @staticmethod
    def get_iou(pred_mask, true_mask):
        """model_utils.py:224-228."""
        assert pred_mask.shape == true_mask.shape, "Wrong shape of masks"
        p, t = pred_mask.bool(), true_mask.bool()
        return float((p & t).sum()) / float((p | t).sum())
