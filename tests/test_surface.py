"""Row f-2, the surface half: `compute_surface_distances` and the statistics on it (segmentation/metrics.py:25-309) --
oracle restatement and CUDA path against vectors made by the REFERENCE's own functions (tests/golden/surface_distances.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics as OM

KEYS = ("distances_gt_to_pred", "distances_pred_to_gt", "surfel_areas_gt", "surfel_areas_pred")
CASES = ["ellipsoids", "ellipsoids_aniso", "shell", "faces", "blobs", "blobs_aniso", "single_voxel", "identical"]
SPACING = {"": (1, 1, 1), "_aniso": (1.0, 0.8, 2.5)}


def _table():
    from mri_epilepsy_diagnosis_b200 import metrics
    tab = metrics._normals()
    return [tab["normals"][c][:int(tab["count"][c])].tolist() for c in range(256)]


def _spacing(case):
    return SPACING["_aniso" if case.endswith("_aniso") else ""]


@pytest.mark.parametrize("case", CASES)
def test_oracle_against_reference_vectors(golden, case):
    g = golden("surface_distances")
    sd = OM.compute_surface_distances(g[case + ":gt"], g[case + ":pred"], _spacing(case), _table())
    for k in KEYS:
        assert np.array_equal(sd[k], g[f"{case}:{k}"]), k


def test_area_table_host_arithmetic(golden):
    from mri_epilepsy_diagnosis_b200 import metrics
    g = golden("surface_distances")
    assert np.array_equal(metrics.neighbour_code_to_surface_area((1, 1, 1)), g["area_table_111"])
    a = metrics.neighbour_code_to_surface_area((1, 1, 1))
    assert a[0] == 0 and a[255] == 0 and abs(a[1] - np.sqrt(3) / 8) < 1e-15 and abs(a[15] - 1.0) < 1e-15      # one corner / a full face


def test_oracle_empty_masks():
    z = np.zeros((6, 7, 8), np.uint8)
    one = z.copy(); one[2:4, 3:5, 1:6] = 1
    assert all(len(v) == 0 for v in OM.compute_surface_distances(z, z, (1, 1, 1), _table()).values())
    sd = OM.compute_surface_distances(one, z, (1, 1, 1), _table())
    assert len(sd["distances_pred_to_gt"]) == 0 and len(sd["distances_gt_to_pred"]) > 0 and np.isinf(sd["distances_gt_to_pred"]).all()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def M():
    from mri_epilepsy_diagnosis_b200 import metrics
    return metrics


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_surface_distances_against_reference_vectors(M, golden, case):
    g = golden("surface_distances")
    gt, pred = torch.from_numpy(g[case + ":gt"]).cuda(), torch.from_numpy(g[case + ":pred"]).cuda()
    sd = M.compute_surface_distances(gt, pred, _spacing(case))
    for k in KEYS:
        want = g[f"{case}:{k}"]
        assert sd[k].dtype == np.float64 and sd[k].shape == want.shape, k
        if case.endswith("_aniso"):
            # fp64 products with non-unit spacing: last-bit freedom in a distance, which may also swap two surfels whose distances
            # tie in exact arithmetic -- distances element-wise, areas as a multiset (the statistics below are order-independent)
            if k.startswith("distances"):
                assert np.allclose(sd[k], want, rtol=1e-14, atol=0), k
            else:
                assert np.array_equal(np.sort(sd[k]), np.sort(want)), k
        else:
            assert np.array_equal(sd[k], want), k                                   # exact integer squared distances + one fp64 sqrt
    tol = dict(rtol=1e-13, atol=0) if case.endswith("_aniso") else dict(rtol=0, atol=0)
    assert np.allclose(np.array(M.compute_average_surface_distance(sd)), g[case + ":asd"], **tol)
    assert np.allclose(M.compute_robust_hausdorff(sd, 95), g[case + ":hd95"], **tol)
    assert np.allclose(np.array(M.compute_surface_overlap_at_tolerance(sd, 1.0)), g[case + ":overlap1"], **tol)
    assert np.allclose(M.compute_surface_dice_at_tolerance(sd, 1.0), g[case + ":sdice1"], **tol)


@pytest.mark.gpu
def test_surface_distances_accept_numpy_bool_and_handle_empty_masks(M, golden):
    g = golden("surface_distances")
    sd = M.compute_surface_distances(g["shell:gt"].astype(bool), g["shell:pred"].astype(bool), (1, 1, 1))
    assert np.array_equal(sd["distances_gt_to_pred"], g["shell:distances_gt_to_pred"])
    z = np.zeros((6, 7, 8), np.uint8)
    one = z.copy(); one[2:4, 3:5, 1:6] = 1
    assert all(len(v) == 0 and v.dtype == np.float64 for v in M.compute_surface_distances(z, z, (1, 1, 1)).values())
    want = OM.compute_surface_distances(one, z, (1, 1, 1), _table())
    got = M.compute_surface_distances(one, z, (1, 1, 1))
    for k in KEYS:
        assert np.array_equal(got[k], want[k]), k
    got = M.compute_surface_distances(z, one, (1, 1, 1))
    assert len(got["distances_gt_to_pred"]) == 0 and np.isinf(got["distances_pred_to_gt"]).all()
    with pytest.raises(ValueError):
        M.compute_surface_distances(z, z[:5], (1, 1, 1))


@pytest.mark.gpu
def test_calculate_metrics_on_a_full_mni_volume_against_live_oracle(M):
    """config 3's grid (192 x 224 x 192): what validate_dsc_asd computes per volume -- Dice, both average surface distances, IoU --
    against the oracle (scipy's exact EDT on the host, a few seconds) on smooth random label volumes."""
    from scipy import ndimage
    from oracle import metrics as OMx
    rng = np.random.default_rng(3)
    f = ndimage.gaussian_filter(rng.random((192, 224, 192)).astype(np.float32), 6.0)
    f = (f - f.min()) / (f.max() - f.min())
    gt, pred = (f > 0.55).astype(np.uint8), (ndimage.shift(f, (1.5, -2.0, 1.0), order=1) > 0.56).astype(np.uint8)
    dsc, asd_mean, asd_std, iou = M.calculate_metrics(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda())
    sd = OMx.compute_surface_distances(gt, pred, (1, 1, 1), _table())
    want_asd = (np.sum(sd["distances_gt_to_pred"] * sd["surfel_areas_gt"]) / np.sum(sd["surfel_areas_gt"]),
                np.sum(sd["distances_pred_to_gt"] * sd["surfel_areas_pred"]) / np.sum(sd["surfel_areas_pred"]))
    assert (asd_mean, asd_std) == want_asd                                             # bit-exact: same sorted lists, same numpy sums
    assert dsc == OMx.compute_dice_coefficient(gt, pred) and iou == OMx.get_iou_score(pred, gt)
    assert len(sd["distances_gt_to_pred"]) > 50000
