"""Row f-3: sliding-grid patch inference (GridSampler / GridAggregator) and the random-patch Queue.

torchio is third-party and absent -> 'parity unpinned' (see oracle/grid.py): the CPU tests pin the host integer logic to the oracle
restatement and to the regression vector grid_kat6_UNPINNED.npz, and check the properties the algorithm must have; the GPU tests
check the two kernels bit for bit against the oracle's numpy slicing / sequential overwrites."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import grid as OG


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def G():
    from mri_epilepsy_diagnosis_b200 import grid
    return grid


SHAPES = [((192, 224, 192), 64, 4), ((64, 64, 64), 64, 4), ((100, 70, 130), 64, 4), ((128, 128, 128), 64, 8), ((65, 64, 130), 64, 4),
          ((200, 64, 96), (64, 64, 32), (4, 0, 2)), ((182, 218, 182), 64, 4), ((70, 70, 70), 64, 0), ((96, 96, 96), 32, 4), ((121, 64, 64), 64, 4)]


@pytest.mark.parametrize("shape,patch,ov", SHAPES)
def test_window_placement_matches_oracle(G, shape, patch, ov):
    p3 = (patch,) * 3 if isinstance(patch, int) else patch
    o3 = (ov,) * 3 if isinstance(ov, int) else ov
    want = OG.grid_spatial_coordinates(shape, p3, o3)
    got = G.grid_locations(shape, patch, ov)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    assert (got[:, :3] >= 0).all() and (got[:, 3:] <= np.array(shape)).all() and ((got[:, 3:] - got[:, :3]) == np.array(p3)).all()
    # the cropped windows cover every voxel at least `overlap` away from the volume faces (what the aggregator can write)
    cover = OG.aggregate(np.zeros(shape, np.uint8), np.ones((len(want), 1) + p3, np.uint8), want, o3)
    inner = tuple(slice(o, s - o) for o, s in zip(o3, shape))
    assert cover[inner].all() and int(cover.sum()) == int(np.prod([s - 2 * o for s, o in zip(shape, o3)]))


def test_regression_vector(G, golden):
    g = golden("grid_kat6_UNPINNED")
    loc = G.grid_locations((192, 224, 192), 64, 4)
    assert np.array_equal(loc, g["locations"]) and len(loc) == int(g["n"]) == 64
    assert np.array_equal(G.grid_locations((100, 70, 130), 64, 4), g["small_locations"])
    rng = np.random.default_rng(6)
    labels = rng.integers(0, 2, (len(loc), 1, 64, 64, 64)).astype(np.uint8)
    out = OG.aggregate(np.zeros((192, 224, 192), np.uint8), labels, loc, (4, 4, 4))
    assert sha16(out) == str(g["agg_sha"]) and int(out.sum()) == int(g["agg_sum"])
    assert int(g["written"]) == 184 * 216 * 184


def test_too_small_volume_raises(G):
    with pytest.raises(AssertionError):
        G.grid_locations((32, 64, 64), 64, 4)
    with pytest.raises(ValueError):
        G.random_patch_locations((32, 64, 64), 64, 1)


def test_random_patch_locations_follow_the_reference_draw_order(G):
    """one scalar torch.randint(size - patch) per dimension, window by window; upper bound exclusive; 0 when the patch spans the axis"""
    gen = torch.Generator().manual_seed(11)
    loc = G.random_patch_locations((100, 64, 130), 64, 50, generator=gen)
    gen2 = torch.Generator().manual_seed(11)
    want = []
    for _ in range(50):
        ini, fin = OG.random_indices((100, 64, 130), (64, 64, 64), lambda n: torch.randint(n, size=(1,), generator=gen2).item())
        want.append(np.concatenate([ini, fin]))
    assert np.array_equal(loc, np.array(want))
    assert (loc[:, 1] == 0).all() and loc[:, 0].max() <= 35 and loc[:, 2].max() <= 65          # size - patch is never drawn


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8, torch.float64])
def test_gather_is_exact(G, dtype):
    rng = np.random.default_rng(1)
    vol = torch.from_numpy(rng.integers(0, 200, (2, 70, 100, 130)).astype(np.float32)).to(dtype)
    s = G.GridSampler({"MRI": {"data": vol}}, 64, 4)
    want = OG.extract_patches(vol.float().numpy(), s.locations)
    got = torch.stack([s[i]["MRI"]["data"] for i in range(len(s))]).float().cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    batches = list(s.batches(5))
    assert sum(b["MRI"]["data"].shape[0] for b in batches) == len(s) and batches[0]["location"].dtype == torch.int64
    assert np.array_equal(torch.cat([b["location"] for b in batches]).numpy(), s.locations)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,patch,ov", [((192, 224, 192), 64, 4), ((100, 70, 130), 64, 4), ((65, 64, 130), 64, 4), ((200, 64, 96), (64, 64, 32), (4, 0, 2))])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
def test_aggregator_equals_sequential_overwrites(G, golden, shape, patch, ov, dtype):
    p3 = (patch,) * 3 if isinstance(patch, int) else patch
    o3 = (ov,) * 3 if isinstance(ov, int) else ov
    loc = G.grid_locations(shape, patch, ov)
    rng = np.random.default_rng(6)
    labels = rng.integers(0, 2, (len(loc), 1) + p3).astype(np.uint8)
    want = OG.aggregate(np.zeros(shape, np.uint8), labels, loc, o3)
    agg = G.GridAggregator(shape, ov)
    lab = torch.from_numpy(labels).to(dtype).cuda()
    for a in range(0, len(loc), 7):                                      # ragged batches, like a DataLoader
        agg.add_batch(lab[a:a + 7], torch.from_numpy(loc[a:a + 7].astype(np.int64)))
    out = agg.get_output_tensor()
    assert tuple(out.shape) == (1,) + tuple(shape) and out.dtype == dtype
    assert np.array_equal(out[0].cpu().numpy().astype(np.uint8), want)
    if shape == (192, 224, 192) and dtype == torch.uint8:
        assert sha16(out[0].cpu().numpy()) == str(golden("grid_kat6_UNPINNED")["agg_sha"])


@pytest.mark.gpu
def test_sliding_window_inference_loop(G):
    """pretraining_3d_unet.ipynb [cell 26] end to end with a real (tiny, random-weight) network of the library: the label volume
    equals the oracle's loop fed with the SAME per-window label maps, and every interior voxel is labelled."""
    import mri_epilepsy_diagnosis_b200 as B
    torch.manual_seed(0)
    net = B.convert(B.zoo.Unet(c=1, n=16, norm="in", num_classes=2).cuda().eval(), dtype=torch.bfloat16)
    vol = torch.randn(1, 80, 64, 112, generator=torch.Generator().manual_seed(3)).cuda()
    sample = {"MRI": {"data": vol}}
    pred = G.sliding_window_labels(net, sample, patch_size=64, patch_overlap=4, batch_size=4)
    assert tuple(pred.shape) == (1, 80, 64, 112) and pred.dtype == torch.uint8
    sampler = G.GridSampler(sample, 64, 4)
    wins = []
    with torch.no_grad():
        for b in sampler.batches(4):
            wins.append(net(b["MRI"]["data"]).argmax(dim=1, keepdim=True).to(torch.uint8).cpu().numpy())
    want = OG.aggregate(np.zeros((80, 64, 112), np.uint8), np.concatenate(wins), sampler.locations, (4, 4, 4))
    assert np.array_equal(pred[0].cpu().numpy(), want)
    assert 0 < int(pred.sum()) < pred.numel()


@pytest.mark.gpu
def test_random_patch_queue(G):
    """segmentation/routine.py:150-178: Queue(ImageSampler) -> DataLoader: sample count, patch contents (= slices of the subject at the
    drawn starts, image and label cut at the SAME place), collated batch layout `prepare_batch` expects."""
    rng = np.random.default_rng(2)
    subjects = []
    for i in range(3):
        img = torch.from_numpy(rng.normal(size=(1, 70 + i, 80, 90)).astype(np.float32)).cuda()
        subjects.append({"MRI": {"data": img}, "LABEL": {"data": (img > 0.5).to(torch.uint8)}})
    q = G.Queue(subjects, max_length=8, samples_per_volume=4, patch_size=64, shuffle_subjects=True, shuffle_patches=True,
                generator=torch.Generator().manual_seed(4))
    assert len(q) == 12
    items = list(q)
    assert len(items) == 12
    for it in items:
        x, y, ini = it["MRI"]["data"], it["LABEL"]["data"], it["index_ini"]
        assert tuple(x.shape) == (1, 64, 64, 64) and y.dtype == torch.uint8
        assert torch.equal(y, (x > 0.5).to(torch.uint8))                                     # same crop for image and label
        owner = [s for s in subjects if torch.equal(s["MRI"]["data"][:, ini[0]:ini[0] + 64, ini[1]:ini[1] + 64, ini[2]:ini[2] + 64], x)]
        assert len(owner) == 1
    batches = list(G.Queue(subjects, 8, 4, 64, False, False, generator=torch.Generator().manual_seed(4)).batches(5))
    assert [b["MRI"]["data"].shape[0] for b in batches] == [5, 5, 2] and tuple(batches[0]["LABEL"]["data"].shape) == (5, 1, 64, 64, 64)
