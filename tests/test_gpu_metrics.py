"""Validation overlap metrics on the device (counting part; b200_overlap_counts) against the golden
vectors of the reference's own compute_dice_coefficient / get_iou_score and against the CPU oracle.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import mri_epilepsy_diagnosis_b200 as pkg
    pkg._cabi.lib()
    return pkg


def test_overlap_against_reference_golden(B, golden):
    g = golden("overlap_metrics")
    for name in ("blobs", "labels5", "disjoint", "pred_empty"):
        pred, gt = torch.from_numpy(g[f"{name}_pred"]).cuda(), torch.from_numpy(g[f"{name}_gt"]).cuda()
        assert B.metrics.compute_dice_coefficient(gt, pred) == float(g[f"{name}_dsc"]), name
        iou = B.metrics.get_iou_score(pred, gt)
        assert iou == g[f"{name}_iou"] and np.asarray(iou).dtype == g[f"{name}_iou"].dtype, name
        dsc2, iou2 = B.metrics.calculate_overlap(gt, pred)
        assert dsc2 == float(g[f"{name}_dsc"]) and iou2 == iou


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (16, 16, 16), (33, 47, 29), (128, 128, 128), (192, 224, 192)], ids=lambda s: "x".join(map(str, s)))
def test_overlap_against_oracle(B, shape):
    from oracle import metrics as M
    rng = np.random.default_rng(sum(shape))
    pred = (rng.random(shape) > 0.8).astype(np.uint8) * rng.integers(1, 4, shape).astype(np.uint8)
    gt = (rng.random(shape) > 0.7).astype(np.uint8)
    pg, gg = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    c = B.metrics.overlap_counts(pg, gg)
    assert c == (int(gt.sum()), int(pred.sum()), int((gt & pred).sum()), int(((pred > 0) & (gt > 0)).sum()), int(((pred > 0) | (gt > 0)).sum()))
    assert B.metrics.compute_dice_coefficient(gg, pg) == M.compute_dice_coefficient(gt, pred)
    assert B.metrics.get_iou_score(pg, gg) == M.get_iou_score(pred, gt)
    # unaligned views (the 16-byte path must fall back), bool and int64 labels (`argmax` output) as the loop produces them
    if pred.size > 40:
        a, b = pg.reshape(-1)[3:-5], gg.reshape(-1)[3:-5]
        assert B.metrics.get_iou_score(a, b) == M.get_iou_score(pred.reshape(-1)[3:-5], gt.reshape(-1)[3:-5])
    assert B.metrics.get_iou_score(pg.long(), gg.bool()) == M.get_iou_score(pred, gt)


def test_overlap_edge_cases(B):
    z = torch.zeros(8, 8, 8, dtype=torch.uint8).cuda()
    assert np.isnan(B.metrics.compute_dice_coefficient(z, z))                 # both masks empty (metrics.py:325-326)
    with pytest.raises(ValueError):
        B.metrics.get_iou_score(z, z[:4])
    with pytest.raises(RuntimeError):
        B.metrics.get_iou_score(z.cpu(), z.cpu())                              # no CPU path
    # after inference: argmax of logits -> labels, exactly like validate_dsc_asd (routine.py:222-229)
    logits = torch.randn(1, 2, 16, 16, 16, generator=torch.Generator().manual_seed(0)).cuda()
    target = (torch.rand(1, 1, 16, 16, 16, generator=torch.Generator().manual_seed(1)) > 0.5).float().cuda()
    labels = logits.argmax(dim=1)
    from oracle import metrics as M
    p, t = labels[0].cpu().numpy().astype(np.uint8), target.cpu().numpy().astype(np.uint8)[0][0]
    assert B.metrics.calculate_overlap(target[0][0], labels[0]) == (M.compute_dice_coefficient(t, p), M.get_iou_score(p, t))
