"""FCD mask generation on the GPU -- mirror of `FCDMaskGenerator` (detection/model_utils.py:118-228).

Same method names and results as the reference class, with the patch classifier run on ALL patches of the volume in a few
large batches (the reference classifies 4 752 patches one at a time, model_utils.py:130-134):

    gen = FCDMaskGenerator(model, gmpm)            # model: a converted PatchModel (or any (P,2,h,w) -> (P,2) logits callable)
    mask = gen.get_mask(img_np)                     # int64 (X,Y,Z) CUDA tensor

Everything stays on the device: sliding-window plan + gather (libb200nn patch kernels), batched PatchModel inference (tcgen05
conv kernels), argmax, scatter into the (4, Y//h, Z) patch map, the 4-neighbour vote and the paint-back.

Two behaviours of the reference are reproduced on purpose (they are what `get_mask` returns today; see oracle/detect.py):
`_postprocess` uses int64 0/1 arrays as INDICES (model_utils.py:190-193), so slabs 0 / 1 of the patch map are overwritten
instead of the voted elements, and `_masking` paints `-j:-j-h:-1`, which is empty for the first strip.  Pass `fixed=True` to
`_postprocess` / `get_mask` for the evidently intended boolean-mask vote.
"""
from __future__ import annotations

import torch

from . import patches as P


class FCDMaskGenerator:
    def __init__(self, model, gmpm, h=16, w=32, batch=2048, device="cuda"):
        self.model, self.h, self.w, self.batch = model, h, w, batch
        self.gmpm = P._dev(gmpm, torch.float64, device)
        self.plan = P.patch_plan(self.gmpm, None, h, w, device=device)          # (P,5) int32: slice, row0, c0, c1, label
        X = self.gmpm.shape[0]
        mid = X // 2 - w
        c0 = self.plan[:, 2].long()
        # patch_1 -> 0, patch_3 -> 1, patch_4 -> 2, patch_2 -> 3  (model_utils.py:160-178)
        self.slot = torch.where(c0 == mid, 1, torch.where(c0 == X - mid - w, 2, torch.where(c0 < mid, 0, 3)))

    # ---- model_utils.py:136-180, batched
    def _get_predictions_per_batches(self, img):
        X, Y, Z = self.gmpm.shape
        patches = P.gather(img, self.plan, self.h, self.w, dtype=torch.float32)
        labels = []
        with torch.no_grad():
            for s in range(0, patches.shape[0], self.batch):
                labels.append(torch.argmax(self.model(patches[s:s + self.batch]), dim=1))
        labels = torch.cat(labels) if labels else torch.zeros(0, dtype=torch.int64, device=self.gmpm.device)
        pm = torch.zeros((4, Y // self.h, Z), dtype=torch.int64, device=self.gmpm.device)
        pm[self.slot, (self.plan[:, 1] // self.h).long(), self.plan[:, 0].long()] = labels
        return pm

    # ---- model_utils.py:182-193
    def _postprocess(self, img, patch_map_tensor, fixed=False):
        pm = patch_map_tensor
        k = 0.25 * torch.tensor([[0.0, 1.0, 0.0], [1.0, 0.0, 1.0], [0.0, 1.0, 0.0]], dtype=torch.float64, device=pm.device)
        # scipy.signal.convolve(mode='same') with a symmetric (1,3,3) kernel == zero-padded 2-D correlation per slab; exact in fp64
        res = torch.nn.functional.conv2d(pm.double()[:, None], k[None, None], padding=1)[:, 0]
        pos, neg = res == 1.0, res == 0.0
        if fixed:
            pm[pos] = 1
            pm[neg] = 0
            return pm
        # the reference indexes with the int64 0/1 arrays: pm[v] = 1 for every value v that occurs in change_to_pos, then
        # pm[v] = 0 for every value v in change_to_neg
        for arr, val in ((pos, 1), (neg, 0)):
            if bool((~arr).any()):
                pm[0] = val
            if bool(arr.any()):
                pm[1] = val
        return pm

    # ---- model_utils.py:195-216 (same boxes, same assignment order)
    def _masking(self, img, patch_map_tensor):
        X, Y, Z = self.gmpm.shape
        h, w = self.h, self.w
        dev = self.gmpm.device
        final = torch.zeros((X, Y, Z), dtype=torch.float64, device=dev)
        i = self.plan[:, 0].long(); j = self.plan[:, 1].long(); c0 = self.plan[:, 2].long()
        val = patch_map_tensor[self.slot, j // h, i].double()
        kx = torch.arange(w, device=dev)[None, :, None]
        ky = torch.arange(h, device=dev)[None, None, :]
        for slot in (0, 3, 1, 2):                              # the reference's order inside a strip; later boxes overwrite earlier ones
            sel = (self.slot == slot) & (j > 0)               # `-0:-h:-1` is an empty slice: the first strip is never painted
            if not bool(sel.any()):
                continue
            xs = c0[sel][:, None, None] + kx
            ys = (Y - j[sel])[:, None, None] - ky              # -j, -j-1, ..., -j-h+1
            zs = i[sel][:, None, None].expand(-1, w, h)
            ok = (xs < X) & (ys >= 0)
            xs, ys = xs.expand(-1, w, h), ys.expand(-1, w, h)
            v = val[sel][:, None, None].expand(-1, w, h)
            final[xs[ok.expand(-1, w, h)], ys[ok.expand(-1, w, h)], zs[ok.expand(-1, w, h)]] = v[ok.expand(-1, w, h)]
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
