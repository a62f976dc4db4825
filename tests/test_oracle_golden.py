"""CPU: the oracle restatement reproduces what the real reference modules produced
(tests/golden/*.npz, made by oracle/make_golden.py in the build container)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err, thin
from oracle import graphs, patches, weights

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different op grouping only


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}


@pytest.mark.parametrize("norm,train", [("bn", True), ("bn", False), ("in", True), ("gn", False)])
def test_unet3d(golden, norm, train):
    g = golden(f"unet3d_{norm}_{'train' if train else 'eval'}")
    sd = _leaf(weights.unet3d_state(1, 16, 2, norm, seed=1))
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 1, 32, 32, 32, generator=gen)
    t = (torch.rand(2, 1, 32, 32, 32, generator=gen) > 0.5).float()
    with torch.set_grad_enabled(train):
        logits = graphs.unet3d(sd, x, norm=norm, dropout=0.5, training=train)
    assert rel_err(logits, g["logits"]) < TOL
    assert sha16(logits.detach().argmax(1).numpy().astype(np.uint8)) == str(g["argmax_sha"])
    if train:
        loss = graphs.dice_loss_mean(logits, t)
        assert abs(float(loss) - float(g["loss"])) < 1e-6
        loss.backward()
        for k in g.files:
            if k.startswith("grad:"):
                assert rel_err(thin(sd[k[5:]].grad), g[k]) < 5e-4, k
        none = sorted(k for k, v in sd.items() if v.requires_grad and v.grad is None)
        assert none == list(g["none_grads"])          # dead conv2/bn2 branch (unet3d.py:43-46)
        if norm == "bn":
            assert rel_err(sd["convd1.bn2.running_mean"], g["rm:convd1.bn2"]) < TOL
            assert rel_err(sd["convu1.bn3.running_var"], g["rv:convu1.bn3"]) < TOL
            assert int(sd["convd1.bn2.num_batches_tracked"]) == 1


def test_autoencoder(golden):
    g = golden("ae_d4_train")
    sd = _leaf(weights.ae_state(depth=4, c_base=16, seed=3))
    x = weights.synthetic_t1w((2, 1, 32, 32, 32), seed=4)
    rec = graphs.autoencoder(sd, x, 4, graphs.AE_DOWN, graphs.AE_UP, training=True)
    assert rel_err(rec, g["rec"]) < TOL
    loss = torch.nn.functional.mse_loss(rec, x)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(sd[k[5:]].grad, g[k]) < 5e-4, k
    g2 = golden("ae_d4_eval_odd")
    xo = weights.synthetic_t1w((1, 1, 36, 28, 20), seed=5)
    with torch.no_grad():
        rec = graphs.autoencoder(sd, xo, 4, graphs.AE_DOWN, graphs.AE_UP, training=False)
    assert rel_err(rec, g2["rec"]) < TOL


def _shipped(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


def test_fader_kat1(golden):
    """KAT-1: shipped *_93_6_4.pth, eval, x=randn(2,1,192^3) seed 0."""
    g = golden("fader_kat1_eval")
    enc, clf, disc = _shipped("encoder_93_6_4.pth"), _shipped("clf_93_6_4.pth"), _shipped("disc_93_6_4.pth")
    x = torch.randn(2, 1, 192, 192, 192, generator=torch.Generator().manual_seed(0))
    assert np.allclose(x[0, 0, 0, 0, :3].numpy(), g["x_head"])
    with torch.no_grad():
        lat, sizes = graphs.encoder(enc, x, 3, graphs.FADER_DOWN)
        pc = graphs.fader_head(clf, lat, graphs.FADER_HEAD, pfx="clf")
        pd = graphs.fader_head(disc, lat, graphs.FADER_HEAD, pfx="disc")
    assert tuple(lat.shape) == (2, 32, 3, 3, 3) and sizes[0] == (192, 192, 192)
    assert rel_err(lat, g["latent"]) < TOL and rel_err(pc, g["clf"]) < TOL and rel_err(pd, g["disc"]) < TOL
    assert abs(float(lat.sum()) - 230.012459) < 1e-2 and pd.argmax(1).tolist() == [3, 3]


def test_fader_train(golden):
    g = golden("fader_encoder_train96")
    enc = _leaf(_shipped("encoder_93_6_4.pth"))
    xs = torch.randn(4, 1, 96, 96, 96, generator=torch.Generator().manual_seed(7))
    lat, _ = graphs.encoder(enc, xs, 3, graphs.FADER_DOWN, training=True)
    assert rel_err(lat, g["latent"]) < TOL
    loss = (lat * torch.linspace(-1, 1, lat.numel()).view_as(lat)).sum() / lat.numel()
    loss.backward()
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(enc[k[5:]].grad, g[k]) < 5e-4, k
    assert rel_err(enc["encode.0.block.5_batch_norm.running_mean"], g["rm0"]) < TOL
    assert rel_err(enc["encode.2.block.5_batch_norm.running_var"], g["rv2"]) < TOL


def test_fader_heads_train(golden):
    g = golden("fader_heads_train")
    clf, disc = _leaf(_shipped("clf_93_6_4.pth")), _shipped("disc_93_6_4.pth")
    gen = torch.Generator().manual_seed(7)
    torch.randn(4, 1, 96, 96, 96, generator=gen)            # same generator stream as make_golden
    lat = torch.randn(6, 32, 3, 3, 3, generator=gen, requires_grad=True)
    y = torch.tensor([0, 1, 1, 0, 1, 0]); dom = torch.tensor([3, 0, 17, 5, 9, 9])
    torch.manual_seed(5)
    pc = graphs.fader_head(clf, lat, graphs.FADER_HEAD, training=True, pfx="clf")
    pd = graphs.fader_head(disc, lat, graphs.FADER_HEAD, training=False, pfx="disc")
    assert rel_err(pc, g["pc"]) < TOL and rel_err(pd, g["pd"]) < TOL
    ce = torch.nn.functional.cross_entropy(pc, y, weight=torch.tensor([1.0, 2.0]))
    adv = graphs.adv_loss(dom, pd, 18)
    assert abs(float(ce) - float(g["ce"])) < 1e-5 and abs(float(adv) - float(g["adv"])) < 1e-5
    (ce + 0.05 * adv).backward()
    assert rel_err(lat.grad, g["dlat"]) < 5e-4
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(clf[k[5:]].grad, g[k]) < 5e-4, k


def test_fepegar_unet_shipped_checkpoint(golden):
    """unet.UNet restatement: strict key/shape coverage of the shipped checkpoint + KAT-2
    regression ('parity unpinned': third-party source absent from the reference)."""
    g = golden("fepegar_kat2_eval_UNPINNED")
    sd = _shipped("whole_im_train_seg_parc_epoch_7.pth")
    synth = weights.fepegar_unet_state(8)
    assert set(sd) == set(synth) and len(sd) == 154
    assert all(sd[k].shape == synth[k].shape for k in sd)
    assert sum(v.numel() for k, v in sd.items() if ".block." not in k and "running" not in k and "num_batches" not in k) == 246412
    x = torch.randn(1, 1, 64, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        logits = graphs.fepegar_unet(sd, x)
    am = logits.argmax(1).numpy().astype(np.uint8)
    assert int(am.sum()) == int(g["fg"]) == 9359                    # KAT-2 foreground voxels
    assert sha16(am) == str(g["argmax_sha"]) == "44fea58a16f0af67"  # KAT-2 hash from SURVEY section 4
    assert abs(float(logits.double().sum()) - 47153.5106) < 0.5


def test_patches_kat4(golden):
    g = golden("patches_kat4")
    gm = patches.read_nifti1_f32(os.path.join(GOLDEN, "MNI152_T1_1mm_brain_gray.nii.gz"))
    assert gm.shape == (182, 218, 182) and sha16(gm.astype(np.float32)) == str(g["template_sha"]) == "9334d4cb5ce117c5"
    gm = gm.astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    plan = patches.patch_plan(gm, None, 16, 32)
    assert np.array_equal(plan, g["plan"].astype(np.int64))
    out = patches.gather_patches(img, plan)
    assert out.shape == tuple(g["shape"]) == (4752, 2, 16, 32)
    assert sha16(out) == str(g["sha"]) == "eefcd430a1f24abe"
    assert np.array_equal(out[:4], g["first"]) and np.array_equal(out[-4:], g["last"])
    # channel 1 is the left-right mirror of channel 0 (patch_utils.py:163-166): size-independent property
    S = np.rot90(img[:, :, plan[0, 0]])
    assert np.array_equal(out[0, 1], S[plan[0, 1]:plan[0, 1] + 16, ::-1][:, 181 - plan[0, 3]:181 - plan[0, 3] + 32])


def test_patches_labelled(golden):
    g = golden("patches_labelled")
    gm = patches.read_nifti1_f32(os.path.join(GOLDEN, "MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    xx, yy, zz = np.meshgrid(np.arange(182), np.arange(218), np.arange(182), indexing="ij")
    c, r = g["mask_center"], g["mask_radii"]
    mask = ((xx - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((zz - c[2]) / r[2]) ** 2 < 1
    p, l = patches.get_all_patches_and_labels(img, gm, mask, 16, 32)
    assert p.shape == tuple(g["shape"]) and sha16(p) == str(g["sha"])
    assert np.array_equal(l, g["labels"])


def test_patches_edge_cases():
    gm = np.zeros((182, 218, 182))
    assert patches.get_only_patches(np.ones_like(gm), gm).shape == (0, 2, 16, 32)     # empty template
    gm[0, 100, 5] = 1.0                                                              # start_idx == 0
    with pytest.raises(AssertionError):
        patches.patch_plan(gm)
    gm[:] = 0; gm[40, 5, 7] = 0.5                                                    # rot90 row 212 -> ragged last strip
    with pytest.raises(ValueError):
        patches.patch_plan(gm)


def test_patch_model(golden):
    g = golden("patch_model")
    k4 = golden("patches_kat4")
    gm = patches.read_nifti1_f32(os.path.join(GOLDEN, "MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    xb = torch.from_numpy(patches.gather_patches(img, k4["plan"].astype(np.int64)[:96])).float()
    sd = _leaf(weights.patch_model_state(seed=9))
    with torch.no_grad():
        ev = graphs.patch_model(sd, xb, training=False)
    assert rel_err(ev, g["eval_logits"]) < TOL
    torch.manual_seed(3)
    tr = graphs.patch_model(sd, xb, training=True)
    assert rel_err(tr, g["train_logits"]) < TOL
    loss = torch.nn.functional.cross_entropy(tr, torch.arange(96) % 2)
    loss.backward()
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(sd[k[5:]].grad, g[k]) < 5e-4, k


def test_op_pins(golden):
    g = golden("op_pins")
    x = torch.randn(2, 4, 16, 16, 16, generator=torch.Generator().manual_seed(0))
    _, idx = torch.nn.functional.max_pool3d(x, 2, 2, return_indices=True)
    assert int(idx.sum()) == 8382113 and sha16(idx.numpy()) == "0329242df874dbc0"     # KAT-3
    assert np.array_equal(g["tie_idx"].reshape(-1), np.array([0, 2, 8, 10, 32, 34, 40, 42]))
    assert np.allclose(g["tri_false"], [0, .25, .75, 1.25, 1.75, 2.25, 2.75, 3])
    assert np.allclose(g["nearest_3to7"], [0, 0, 0, 1, 1, 2, 2])


def _kat5_image():
    img = np.random.default_rng(1).random((182, 218, 182))
    img[40:150, 60:150, 50:120] *= 1.35
    return img


def test_fcd_mask_kat5(golden):
    """oracle/detect.py against vectors produced by the REFERENCE FCDMaskGenerator (detection/model_utils.py:118-228; generated by
    oracle/make_golden.py detect): patch map, the post-processing (int-array-as-index quirk included), the painted mask."""
    from oracle import detect
    g = golden("fcd_mask_kat5")
    gm = patches.read_nifti1_f32(os.path.join(GOLDEN, "MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = _kat5_image()
    thr = float(g["thr"])
    classify = lambda p: (torch.from_numpy(p[:, 0]).double().mean(dim=(1, 2)) > thr).numpy().astype(np.int64)
    pm = detect.predictions_per_batches(img, gm, classify)
    assert np.array_equal(pm, g["patch_map"].astype(np.int64)) and pm.shape == (4, 13, 182)
    post = detect.postprocess(pm)
    assert np.array_equal(post, g["post"].astype(np.int64))
    assert not post[0].any() and not post[1].any() and np.array_equal(post[2:], pm[2:])      # the quirk: slabs 0/1 are overwritten
    mask = detect.masking(img, gm, post)
    assert float(mask.sum()) == float(g["mask_sum"]) and sha16(mask.astype(np.int8)) == str(g["mask_sha"])
    unvoted = detect.masking(img, gm, pm)
    assert float(unvoted.sum()) == float(g["mask_unvoted_sum"]) and sha16(unvoted.astype(np.int8)) == str(g["mask_unvoted_sha"])
    assert not unvoted[:, 218 - 15:, :].any()                  # the first strip (j = 0) is never painted (`-0:-16:-1` is empty)
    assert abs(detect.get_iou(unvoted, img > 1.0) - float(g["iou"])) < 1e-12


def test_histstd_cell9(golden):
    """oracle/preprocess.py against vectors produced by the NOTEBOOK's own `normalize` (classification/train_ENC_CLF.ipynb
    [cell 9], generated by oracle/make_golden.py histstd) with the shipped 13 landmarks: bit-exact float32 volumes and float64
    percentiles, incl. a zero background, heavy ties, a constant image (every diff_perc < epsilon), `mask=` and `cutoff=`."""
    from oracle import preprocess
    g = golden("histstd_cell9")
    lm = g["landmarks"]
    assert lm.shape == (13,) and lm.dtype == np.float64
    for name in ("brain", "dense", "const", "steps"):
        x = g[f"{name}_x"]
        assert np.array_equal(preprocess.percentile_values(x), g[f"{name}_pct"]), name
        assert np.array_equal(preprocess.normalize(x, lm), g[f"{name}_y"]), name
    assert np.array_equal(preprocess.normalize(g["brain_x"], lm, mask=g["brain_x"] > 0), g["brain_masked_y"])
    assert np.array_equal(preprocess.normalize(g["dense_x"], lm, cutoff=(0.05, 0.95)), g["dense_cut_y"])
    assert np.all(np.diff(g["steps_y"].reshape(16, -1)[:, 0]) >= 0)             # the map is monotone
    v = np.arange(5 * 6 * 7, dtype=np.float32).reshape(5, 6, 7)
    assert np.array_equal(preprocess.reshape_image(v, (1, 2, 3), (3, 3, 3)), v[1:4, 2:5, 3:6].reshape(1, 3, 3, 3))
    with pytest.raises(AssertionError):
        preprocess.reshape_image(v, (3, 2, 3), (3, 3, 3))


def test_overlap_metrics(golden):
    """oracle/metrics.py against the reference's own compute_dice_coefficient (segmentation/metrics.py:312-329) and get_iou_score
    (segmentation/routine.py:198-204), vectors from oracle/make_golden.py metrics: bit-equal values AND result types."""
    from oracle import metrics as M
    g = golden("overlap_metrics")
    for name in ("blobs", "labels5", "disjoint", "pred_empty"):
        pred, gt = g[f"{name}_pred"], g[f"{name}_gt"]
        assert M.compute_dice_coefficient(gt, pred) == float(g[f"{name}_dsc"]), name
        iou = M.get_iou_score(pred, gt)
        assert iou == g[f"{name}_iou"] and np.asarray(iou).dtype == g[f"{name}_iou"].dtype, name
    z = np.zeros((4, 4, 4), np.uint8)
    assert np.isnan(M.compute_dice_coefficient(z, z))                # metrics.py:325-326 (np.NaN there)


def test_modified_3dunet(golden):
    """segmentation/models/modified_3dunet.py: oracle restatement vs vectors of the REAL class (eval; train step with Dropout3d p=0)."""
    g = golden("modified3dunet")
    import mri_epilepsy_diagnosis_b200 as pkg
    template = pkg.zoo.Modified3DUNet(1, 2, 8).state_dict()
    assert list(template.keys()) == list(g["keys"])            # the zoo mirror has the reference's state_dict keys, in order
    sd = _leaf(weights.seeded_like(template, seed=31))
    gen = torch.Generator().manual_seed(32)
    x = torch.randn(2, 1, 32, 32, 32, generator=gen)
    t = (torch.rand(2, 1, 32, 32, 32, generator=gen) > 0.5).float()
    with torch.no_grad():
        ev = graphs.modified_3dunet(sd, x)
    assert rel_err(ev, g["eval_logits"]) < TOL and sha16(ev.argmax(1).numpy().astype(np.uint8)) == str(g["argmax_sha"])
    logits = graphs.modified_3dunet(sd, x, training=True, p_drop=0.0)
    assert rel_err(logits, g["train_logits"]) < TOL
    loss = graphs.dice_loss_mean(logits, t)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(thin(sd[k[5:]].grad, 4096), g[k]) < 5e-4, k


@pytest.mark.parametrize("name", ["voxresnet_b3", "voxresnet_b4", "cnn_b3", "dilated_cnn"])
def test_cnn_model_family(golden, name):
    """classification/models/cnn_model.py (VoxResNet incl. the n_blocks=4 activation quirk, CNN, DilatedCNN)."""
    import mri_epilepsy_diagnosis_b200 as pkg
    g = golden(name)
    ctor, shape, fn, seed = CNN_CASES[name]
    template = ctor(pkg.zoo).state_dict()
    assert list(template.keys()) == list(g["keys"])
    sd = _leaf(weights.seeded_like(template, seed=40 + seed))
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(50 + seed))
    with torch.no_grad():
        ev = fn({k: v.detach().clone() for k, v in sd.items()}, x, False)
    assert rel_err(ev, g["eval_out"]) < TOL
    tr = fn(sd, x, True)
    assert rel_err(tr, g["train_out"]) < TOL
    y = torch.arange(shape[0]) % 2
    loss = torch.nn.functional.nll_loss(torch.log(tr), y) if name == "dilated_cnn" else torch.nn.functional.cross_entropy(tr, y)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for k in g.files:
        if k.startswith("grad:"):
            assert rel_err(thin(sd[k[5:]].grad, 4096), g[k]) < 2e-3, k
        if k.startswith("buf:"):
            assert rel_err(sd[k[4:]], g[k]) < TOL, k


CNN_CASES = {
    "voxresnet_b3": (lambda z: z.VoxResNet((32, 32, 32), 2, 16, 2, 3), (3, 1, 32, 32, 32), lambda sd, x, tr: graphs.voxresnet(sd, x, 3, 2, tr), 0),
    "voxresnet_b4": (lambda z: z.VoxResNet((32, 32, 32), 2, 8, 1, 4), (2, 1, 32, 32, 32), lambda sd, x, tr: graphs.voxresnet(sd, x, 4, 1, tr), 1),
    "cnn_b3": (lambda z: z.CNN((32, 40, 24), 16, 3), (4, 1, 32, 40, 24), lambda sd, x, tr: graphs.cnn(sd, x, 3, 1, tr), 2),
    "dilated_cnn": (lambda z: z.DilatedCNN((180, 180, 180), 16), (2, 1, 180, 180, 180), lambda sd, x, tr: graphs.dilated_cnn(sd, x, tr), 3),
}
