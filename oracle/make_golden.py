#!/usr/bin/env python
"""Generate tests/golden/* by running the REAL reference modules (TEST INFRASTRUCTURE).

Run in the build container, where /root/reference exists:

    python oracle/make_golden.py

Everything here calls the reference's own classes/functions (through oracle.refload)
on seeded inputs and deterministic weights from oracle.weights, and freezes the
results.  The oracle restatement (oracle.graphs / oracle.patches) is then checked
against these files by the CPU test-suite, and the CUDA path by the GPU suite.

Data fixtures (not source) copied next to the vectors: the three shipped fader
checkpoints, one shipped segmentation checkpoint and the GM template.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import graphs, patches, refload, weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"  wrote {name}.npz ({os.path.getsize(path) / 1024:.0f} KiB)")


def grads_of(model):
    return {k: p.grad for k, p in model.named_parameters()}


def thin(t, cap=32768):
    """Large tensors are stored as a strided sample of their flattening (tests use the same rule)."""
    t = t.detach() if torch.is_tensor(t) else torch.as_tensor(t)
    if t.numel() <= cap:
        return t
    flat = t.reshape(-1)
    return flat[:: flat.numel() // cap]


def unet3d_cases():
    for norm, train in (("bn", True), ("bn", False), ("in", True), ("gn", False)):
        torch.manual_seed(0)
        sd = weights.unet3d_state(1, 16, 2, norm, seed=1)
        net = refload.make_unet3d(c=1, n=16, dropout=0.5, norm=norm, num_classes=2)
        net.load_state_dict(sd, strict=True)
        net.train(train)
        g = torch.Generator().manual_seed(2)
        x = torch.randn(2, 1, 32, 32, 32, generator=g)
        t = (torch.rand(2, 1, 32, 32, 32, generator=g) > 0.5).float()
        out = {}
        if train:
            logits = net(x)
            # loss exactly as segmentation/routine.py:272-274
            routine = refload.seg_routine_module()
            loss = routine.get_dice_loss(torch.softmax(logits, dim=1), t).mean()
            loss.backward()
            gr = grads_of(net)
            out["loss"] = loss.detach()
            for k in ("convd1.conv1.weight", "convd1.conv3.weight", "convd2.conv1.weight", "convd5.conv3.weight",
                      "convu1.conv3.weight", "convu1.conv2.weight", "seg1.weight", "seg1.bias", "convu3.conv1.weight"):
                out["grad:" + k] = thin(gr[k])
            if norm == "bn":
                out["grad:convd1.bn1.weight"] = gr["convd1.bn1.weight"]
                out["grad:convu1.bn3.bias"] = gr["convu1.bn3.bias"]
                out["rm:convd1.bn2"] = net.convd1.bn2.running_mean
                out["rv:convu1.bn3"] = net.convu1.bn3.running_var
            out["none_grads"] = np.array(sorted(k for k, v in gr.items() if v is None))
            out["grad_norms"] = np.array([float(v.norm()) if v is not None else -1.0 for v in gr.values()])
            out["grad_keys"] = np.array(list(gr.keys()))
        else:
            with torch.no_grad():
                logits = net(x)
        out["logits"] = logits.detach()
        out["argmax_sha"] = np.array(sha16(logits.detach().argmax(1).numpy().astype(np.uint8)))
        save(f"unet3d_{norm}_{'train' if train else 'eval'}", **out)


def ae_cases():
    ae_mod = refload.ae_module()
    down = dict(graphs.AE_DOWN)
    up = dict(graphs.AE_UP)
    kw = dict(c_in=1, is_skip=False, deapth=4, c_base=16, inc_size=2, reduce_size=False,
              down_block_kwargs=down, up_block_kwargs=up)
    torch.manual_seed(0)
    net = ae_mod.AE(**kw)
    net.load_state_dict(weights.ae_state(depth=4, c_base=16, seed=3), strict=True)
    net.train()
    x = weights.synthetic_t1w((2, 1, 32, 32, 32), seed=4)
    rec = net(x)
    loss = torch.nn.MSELoss()(rec, x)
    loss.backward()
    gr = grads_of(net)
    save("ae_d4_train", rec=rec.detach(), loss=loss.detach(),
         **{"grad:" + k: gr[k] for k in ("enc.encode.0.block.1_convx.weight", "enc.encode.0.block.3_convz.weight",
                                         "enc.encode.3.block.2_convy.weight", "dec.decode.3.block.4_convz.weight",
                                         "dec.decode.0.block.2_convx.bias", "dec.vox.weight",
                                         "enc.encode.1.block.5_batch_norm.weight")},
         grad_norms=np.array([float(v.norm()) for v in gr.values()]), grad_keys=np.array(list(gr.keys())))
    # odd input size: exercises the nearest-to-size branch AE_model.py:116-119
    net.eval()
    xo = weights.synthetic_t1w((1, 1, 36, 28, 20), seed=5)
    with torch.no_grad():
        save("ae_d4_eval_odd", rec=net(xo))


def fader_cases():
    ae_mod = refload.ae_module()
    down = dict(graphs.FADER_DOWN)
    up = dict(up="upsample", scale=4, scale_mode="nearest", conv_k=3, conv_pad=1, conv_s=1, batch_norm=False, act="l_relu")
    enc = ae_mod.AE(c_in=1, is_skip=False, deapth=3, c_base=8, inc_size=2, reduce_size=False,
                    down_block_kwargs=down, up_block_kwargs=up).enc
    head_kw = dict(c_in=32, c_out=64, conv_k=3, conv_s=1, conv_pad=0, l_in=64, l_out=32, batch_norm=True, act="relu", p_drop=0.5)
    disc = ae_mod.Discriminator(n_domains=18, **head_kw)
    clf = ae_mod.Classificator(n_class=2, **head_kw)
    for m, f in ((enc, "encoder"), (clf, "clf"), (disc, "disc")):
        m.load_state_dict(torch.load(refload.path(f"classification/{f}_93_6_4.pth"), map_location="cpu", weights_only=True), strict=True)
        m.eval()
    # KAT-1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1, 192, 192, 192, generator=g)
    with torch.no_grad():
        lat, _ = enc(x)
        save("fader_kat1_eval", latent=lat, clf=clf(lat), disc=disc(lat), x_head=x[0, 0, 0, 0, :3])
    # one adversarial encoder+clf step at a small size (train_ENC_CLF.ipynb [cell 16] second half)
    g = torch.Generator().manual_seed(7)
    xs = torch.randn(4, 1, 96, 96, 96, generator=g)
    y = torch.tensor([0, 1, 1, 0])
    dom = torch.tensor([3, 0, 17, 5])
    # 96^3 -> latent 1^3 would break the p0 k3 heads; use the encoder only + a surrogate loss on the latent
    enc.train()
    lat, _ = enc(xs)
    loss = (lat * torch.linspace(-1, 1, lat.numel()).view_as(lat)).sum() / lat.numel()
    loss.backward()
    gr = grads_of(enc)
    save("fader_encoder_train96", latent=lat.detach(), loss=loss.detach(),
         **{"grad:" + k: v for k, v in gr.items()},
         rm0=enc.encode[0].block["5_batch_norm"].running_mean, rv2=enc.encode[2].block["5_batch_norm"].running_var)
    # heads in train mode on a synthetic latent (B,32,3,3,3), with the notebook's losses
    clf.train(); disc.eval()
    torch.manual_seed(11)
    lat = torch.randn(6, 32, 3, 3, 3, generator=g, requires_grad=True)
    y = torch.tensor([0, 1, 1, 0, 1, 0]); dom = torch.tensor([3, 0, 17, 5, 9, 9])
    torch.manual_seed(5)                       # dropout mask
    pc = clf(lat); pd = disc(lat)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0]))(pc, y)
    adv = graphs.adv_loss(dom, pd, 18)
    (ce + 0.05 * adv).backward()
    save("fader_heads_train", pc=pc.detach(), pd=pd.detach(), ce=ce.detach(), adv=adv.detach(), dlat=lat.grad,
         **{"grad:" + k: v for k, v in grads_of(clf).items()})


def _all_grads(net, cap=4096):
    return {"grad:" + k: thin(p.grad, cap) for k, p in net.named_parameters() if p.grad is not None}


def modified_unet_cases():
    """segmentation/models/modified_3dunet.py through the REAL class: eval, and one training step's gradients with the
    Dropout3d probability set to 0 on the instance (the CUDA path cannot replay the CPU RNG stream; source untouched)."""
    mod = refload.modified_unet_module()
    torch.manual_seed(0)
    net = mod.Modified3DUNet(1, 2, 8)
    sd = weights.seeded_like(net.state_dict(), seed=31)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 1, 32, 32, 32, generator=g)
    t = (torch.rand(2, 1, 32, 32, 32, generator=g) > 0.5).float()
    net.eval()
    with torch.no_grad():
        ev = net(x)
    assert torch.equal(graphs.modified_3dunet(sd, x), ev), "oracle modified_3dunet != reference (eval)"
    net.train()
    net.dropout3d.p = 0.0
    logits = net(x)
    loss = refload.seg_routine_module().get_dice_loss(torch.softmax(logits, dim=1), t).mean()
    loss.backward()
    save("modified3dunet", keys=np.array(list(sd.keys())), eval_logits=ev, train_logits=logits.detach(), loss=loss.detach(),
         argmax_sha=np.array(sha16(ev.argmax(1).numpy().astype(np.uint8))), **_all_grads(net))


def cnn_model_cases():
    """classification/models/cnn_model.py (VoxResNet incl. the n_blocks=4 `activation_6` quirk, CNN, DilatedCNN) through the
    REAL classes: eval probabilities/logits, train-mode logits, CrossEntropy gradients and BatchNorm running statistics."""
    mod = refload.cnn_module()
    cases = {
        "voxresnet_b3": (lambda: mod.VoxResNet((32, 32, 32), 2, 16, 2, 3), (3, 1, 32, 32, 32), lambda sd, x, tr: graphs.voxresnet(sd, x, 3, 2, tr)),
        "voxresnet_b4": (lambda: mod.VoxResNet((32, 32, 32), 2, 8, 1, 4), (2, 1, 32, 32, 32), lambda sd, x, tr: graphs.voxresnet(sd, x, 4, 1, tr)),
        "cnn_b3": (lambda: mod.CNN((32, 40, 24), 16, 3), (4, 1, 32, 40, 24), lambda sd, x, tr: graphs.cnn(sd, x, 3, 1, tr)),
        "dilated_cnn": (lambda: mod.DilatedCNN((180, 180, 180), 16), (2, 1, 180, 180, 180), lambda sd, x, tr: graphs.dilated_cnn(sd, x, tr)),
    }
    for i, (name, (ctor, shape, oracle)) in enumerate(cases.items()):
        torch.manual_seed(0)
        net = ctor()
        sd = weights.seeded_like(net.state_dict(), seed=40 + i)
        net.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(50 + i)
        x = torch.randn(*shape, generator=g)
        y = torch.arange(shape[0]) % 2
        net.eval()
        with torch.no_grad():
            ev = net(x)
        assert torch.equal(oracle({k: v.clone() for k, v in sd.items()}, x, False), ev), f"oracle {name} != reference (eval)"
        net.train()
        tr = net(x)
        assert torch.equal(oracle({k: v.clone() for k, v in sd.items()}, x, True), tr.detach()), f"oracle {name} != reference (train)"
        loss = (torch.nn.NLLLoss()(torch.log(tr), y) if name == "dilated_cnn" else torch.nn.CrossEntropyLoss()(tr, y))
        loss.backward()
        bn_keys = [k for k in net.state_dict() if k.endswith("running_mean") or k.endswith("running_var")]
        first_bn, last_bn = bn_keys[0], bn_keys[-1]
        save(name, keys=np.array(list(sd.keys())), eval_out=ev, train_out=tr.detach(), loss=loss.detach(),
             **{"buf:" + k: net.state_dict()[k] for k in (first_bn, last_bn)}, **_all_grads(net))


def patch_cases():
    pu = refload.patch_utils_module()
    gm = patches.read_nifti1_f32(refload.path("detection/MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    t0 = time.time()
    ref = pu.get_only_patches(img, gm, 16, 32)
    print(f"  reference get_only_patches: {time.time() - t0:.1f}s, {ref.shape}")
    plan = patches.patch_plan(gm, None, 16, 32)
    mine = patches.gather_patches(img, plan)
    assert ref.shape == mine.shape and np.array_equal(ref, mine), "oracle patches != reference"
    save("patches_kat4", shape=np.array(ref.shape), total=np.array(ref.sum()), sha=np.array(sha16(ref)),
         plan=plan.astype(np.int16), template_sha=np.array(sha16(gm.astype(np.float32))), first=ref[:4], last=ref[-4:])
    # labelled variant with a synthetic lesion mask (ellipsoid), incl. the k=1..15 positive-only passes
    xx, yy, zz = np.meshgrid(np.arange(182), np.arange(218), np.arange(182), indexing="ij")
    mask = ((xx - 60) / 9.0) ** 2 + ((yy - 120) / 11.0) ** 2 + ((zz - 90) / 7.0) ** 2 < 1
    t0 = time.time()
    rp, rl = pu.get_all_patches_and_labels(img, gm, mask, 16, 32)
    print(f"  reference get_all_patches_and_labels: {time.time() - t0:.1f}s, {rp.shape}, positives {int(rl.sum())}")
    plan2 = patches.patch_plan(gm, mask, 16, 32)
    mp = patches.gather_patches(img, plan2)
    assert np.array_equal(rp, mp) and np.array_equal(rl, plan2[:, patches.LABEL].astype(bool))
    save("patches_labelled", shape=np.array(rp.shape), total=np.array(rp.sum()), sha=np.array(sha16(rp)),
         labels=rl, plan=plan2.astype(np.int16), mask_center=np.array([60, 120, 90]), mask_radii=np.array([9.0, 11.0, 7.0]))
    # patch classifier (2-D) on the first 96 patches, eval and train
    PatchModel, _ = refload.patch_model_classes()
    torch.manual_seed(0)
    net = PatchModel()
    net.load_state_dict(weights.patch_model_state(seed=9), strict=True)
    xb = torch.from_numpy(ref[:96]).float()
    net.eval()
    with torch.no_grad():
        ev = net(xb)
    net.train()
    torch.manual_seed(3)
    tr = net(xb)
    yb = (torch.arange(96) % 2)
    loss = torch.nn.CrossEntropyLoss()(tr, yb)
    loss.backward()
    gr = grads_of(net)
    save("patch_model", eval_logits=ev, train_logits=tr.detach(), loss=loss.detach(),
         **{"grad:" + k: gr[k] for k in ("conv_blocks.0.conv.weight", "conv_blocks.4.conv.weight", "conv_blocks.2.bn.weight", "fc2.weight")})


def patch_model_nodrop_case():
    """PatchModel in TRAIN mode with the Dropout probability set to 0 on the instance (a GPU run cannot replay the CPU mask):
    logits, CrossEntropy loss, every parameter gradient and the BatchNorm running statistics, on 96 real patches."""
    pu = refload.patch_utils_module()
    gm = patches.read_nifti1_f32(refload.path("detection/MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    plan = patches.patch_plan(gm, None, 16, 32)[:96]
    xb = torch.from_numpy(patches.gather_patches(img, plan)).float()
    PatchModel, _ = refload.patch_model_classes()
    torch.manual_seed(0)
    net = PatchModel()
    net.load_state_dict(weights.patch_model_state(seed=9), strict=True)
    net.train()
    net.dropout.p = 0.0
    tr = net(xb)
    loss = torch.nn.CrossEntropyLoss()(tr, torch.arange(96) % 2)
    loss.backward()
    save("patch_model_nodrop", train_logits=tr.detach(), loss=loss.detach(), x_sha=np.array(sha16(xb.numpy())),
         rm0=net.conv_blocks[0].bn.running_mean, rv4=net.conv_blocks[4].bn.running_var, **_all_grads(net, cap=8192))


def rel_err_t(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30))


def grid_case():
    """Row f-3 (sliding-grid inference).  torchio is third-party and absent: this vector comes from the oracle restatement of its
    published <= 0.16 algorithm (oracle/grid.py) -- a REGRESSION anchor, 'parity unpinned' -- on the MNI-sized volume of config 3
    with the notebook's arguments (64^3 windows, overlap 4)."""
    from oracle import grid
    shape = (192, 224, 192)
    loc = grid.grid_spatial_coordinates(shape, (64, 64, 64), (4, 4, 4))
    rng = np.random.default_rng(6)
    labels = rng.integers(0, 2, (len(loc), 1, 64, 64, 64)).astype(np.uint8)
    out = grid.aggregate(np.zeros(shape, np.uint8), labels, loc, (4, 4, 4))
    vol = rng.integers(0, 255, (1,) + shape).astype(np.uint8)
    patches_ = grid.extract_patches(vol, loc)
    save("grid_kat6_UNPINNED", locations=loc, n=np.array(len(loc)), agg_sha=np.array(sha16(out)), agg_sum=np.array(int(out.sum())),
         written=np.array(int((grid.aggregate(np.zeros(shape, np.uint8), np.ones_like(labels), loc, (4, 4, 4)) > 0).sum())),
         patches_sha=np.array(sha16(patches_)), small_locations=grid.grid_spatial_coordinates((100, 70, 130), (64, 64, 64), (4, 4, 4)))


def surface_cases():
    """Row f-2, surface half: `compute_surface_distances` and the statistics built on it through the REFERENCE's own functions
    (segmentation/metrics.py:25-309) on synthetic label volumes (ellipsoids, a hollow shell, touching the volume faces, random blobs,
    disjoint far-apart objects, one empty mask).  Also writes the package's data asset: the 256-entry normal table that
    metrics.py:333-600 vendors from Google's surface-distance library (Apache-2.0), as a structured .npy (data, not source)."""
    from oracle import metrics as M
    ref = refload._load("_ref_metrics", "segmentation/metrics.py")
    table = ref.neighbour_code_to_normals
    asset = np.zeros(256, dtype=[("count", np.int32), ("normals", np.float64, (4, 3))])
    for code, normals in enumerate(table):
        arr = np.array(normals, dtype=np.float64).reshape(-1, 3)
        asset["count"][code] = len(arr)
        asset["normals"][code][:len(arr)] = arr
    dst = os.path.join(ROOT, "mri_epilepsy_diagnosis_b200", "data", "neighbour_code_normals.npy")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.save(dst, asset)
    print("  wrote", os.path.relpath(dst, ROOT))
    rng = np.random.default_rng(12)
    zz, yy, xx = np.meshgrid(np.arange(48), np.arange(56), np.arange(40), indexing="ij")
    ell = lambda c, r: (((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 < 1)
    blobs = lambda thr: ndimage_blobs(rng, (48, 56, 40), thr)
    cases = {
        "ellipsoids": (ell((24, 25, 20), (10, 15, 9)), ell((25, 27, 21), (9.5, 14, 10))),
        "shell": (ell((24, 28, 20), (14, 16, 12)) & ~ell((24, 28, 20), (8, 9, 6)), ell((24, 28, 20), (13, 16, 12))),
        "faces": (ell((2, 28, 20), (10, 12, 8)), ell((44, 30, 36), (12, 10, 9))),          # cut by the volume faces, far apart
        "blobs": (blobs(0.62), blobs(0.6)),
        "single_voxel": (np.pad(np.ones((1, 1, 1), bool), ((5, 42), (7, 48), (9, 30))), ell((24, 25, 20), (6, 6, 6))),
        "identical": (ell((24, 25, 20), (10, 15, 9)), ell((24, 25, 20), (10, 15, 9))),
    }
    out = {}
    for name, (gt, pred) in cases.items():
        gt8, pr8 = gt.astype(np.uint8), pred.astype(np.uint8)                                 # validate_dsc_asd passes uint8 0/1 volumes
        for sp_name, spacing in (("", (1, 1, 1)), ("_aniso", (1.0, 0.8, 2.5))):
            if sp_name and name not in ("ellipsoids", "blobs"):
                continue
            sd = ref.compute_surface_distances(gt8, pr8, spacing)
            mine = M.compute_surface_distances(gt8, pr8, spacing, table)
            for k in sd:
                assert np.array_equal(sd[k], mine[k]), (name, k)
            key = name + sp_name
            out[key + ":gt"], out[key + ":pred"] = gt8, pr8
            for k, v in sd.items():
                out[f"{key}:{k}"] = v
            out[key + ":asd"] = np.array(ref.compute_average_surface_distance(sd))
            out[key + ":hd95"] = np.array(ref.compute_robust_hausdorff(sd, 95))
            out[key + ":overlap1"] = np.array(ref.compute_surface_overlap_at_tolerance(sd, 1.0))
            out[key + ":sdice1"] = np.array(ref.compute_surface_dice_at_tolerance(sd, 1.0))
            print(f"  {key}: {len(sd['distances_gt_to_pred'])} / {len(sd['distances_pred_to_gt'])} surfels, asd {out[key + ':asd']}, hd95 {float(out[key + ':hd95']):.4f}")
    out["area_table_111"] = M_area(ref, (1, 1, 1))
    save("surface_distances", **out)


def M_area(ref, spacing):
    """the reference's surfel-area table for `spacing`, by running its own code on a one-voxel mask and reading the table back
    through the areas it reports is not possible for all 256 codes -- so evaluate metrics.py:58-71's expression on its table"""
    area = np.zeros([256])
    for code in range(256):
        normals = np.array(ref.neighbour_code_to_normals[code])
        s = 0
        for i in range(normals.shape[0]):
            n = np.zeros([3])
            n[0] = normals[i, 0] * spacing[1] * spacing[2]
            n[1] = normals[i, 1] * spacing[0] * spacing[2]
            n[2] = normals[i, 2] * spacing[0] * spacing[1]
            s += np.linalg.norm(n)
        area[code] = s
    return area


def ndimage_blobs(rng, shape, thr):
    from scipy import ndimage
    f = ndimage.gaussian_filter(rng.random(shape), 2.5)
    f = (f - f.min()) / (f.max() - f.min())
    return f > thr


def detect_cases():
    """FCD mask generation (detection/model_utils.py:118-228) run through the REFERENCE class with a deterministic stand-in
    classifier (best_model.pth is not shipped): pins the patch map, the post-processing quirk and the painted mask."""
    from oracle import detect
    gm = patches.read_nifti1_f32(refload.path("detection/MNI152_T1_1mm_brain_gray.nii.gz")).astype(np.float64)
    img = np.random.default_rng(1).random((182, 218, 182))
    img[40:150, 60:150, 50:120] *= 1.35                     # a brighter block (both hemispheres) so that the labels are not spatially uniform
    thr = 0.6

    def model(patch_torch):                                   # what the reference's `model(patch_torch)` must look like
        m = patch_torch[:, 0].double().mean(dim=(1, 2))
        return torch.stack([thr - m, m - thr], dim=1)

    Gen = refload.fcd_mask_generator_class(model)
    gen = object.__new__(Gen)
    gen.model, gen.h, gen.w, gen.gmpm = model, 16, 32, gm
    t0 = time.time()
    pm_ref = gen._get_predictions_per_batches(img)
    post_ref = gen._postprocess(img, pm_ref.copy())
    mask_ref = gen._masking(img, post_ref)
    mask_unvoted_ref = gen._masking(img, pm_ref)              # the painting rule on a non-trivial map
    print(f"  reference FCDMaskGenerator: {time.time() - t0:.1f}s, positives in map {int(pm_ref.sum())}, painted {int(mask_ref.sum())} / {int(mask_unvoted_ref.sum())}")
    classify = lambda p: (torch.from_numpy(p[:, 0]).double().mean(dim=(1, 2)) > thr).numpy().astype(np.int64)
    pm = detect.predictions_per_batches(img, gm, classify)
    assert np.array_equal(pm, pm_ref), "oracle patch map != reference"
    assert np.array_equal(detect.postprocess(pm), post_ref), "oracle postprocess != reference"
    assert np.array_equal(detect.masking(img, gm, post_ref), mask_ref) and np.array_equal(detect.masking(img, gm, pm_ref), mask_unvoted_ref)
    save("fcd_mask_kat5", patch_map=pm_ref.astype(np.int8), post=post_ref.astype(np.int8), mask_sum=np.array(mask_ref.sum()),
         mask_sha=np.array(sha16(mask_ref.astype(np.int8))), mask_unvoted_sum=np.array(mask_unvoted_ref.sum()),
         mask_unvoted_sha=np.array(sha16(mask_unvoted_ref.astype(np.int8))), thr=np.array(thr),
         iou=np.array(gen.get_iou(mask_unvoted_ref, img > 1.0)))


def op_pins():
    """SURVEY section 8 a-9 pins, produced by the same torch calls the reference modules make."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, 16, generator=g)
    y, idx = torch.nn.MaxPool3d(2, 2, return_indices=True)(x)
    ties = torch.zeros(1, 1, 4, 4, 4)
    _, tidx = torch.nn.MaxPool3d(2, 2, return_indices=True)(ties)
    odd = torch.randn(1, 2, 5, 7, 9, generator=g)
    yo, io = torch.nn.MaxPool3d(2, 2, return_indices=True)(odd)
    y42, i42 = torch.nn.MaxPool3d(4, 2, return_indices=True)(x)
    line = torch.arange(4.0).view(1, 1, 1, 1, 4).expand(1, 1, 2, 2, 4)
    save("op_pins", pool_idx=idx, pool_idx_sum=np.array(int(idx.sum())), pool_sha=np.array(sha16(idx.numpy())),
         tie_idx=tidx, odd_idx=io, odd_val=yo, k4s2_idx=i42,
         tri_false=torch.nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)(line)[0, 0, 0, 0],
         tri_true=torch.nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)(line)[0, 0, 0, 0],
         nearest_3to7=torch.nn.functional.interpolate(torch.arange(3.0).view(1, 1, 1, 1, 3), size=(1, 1, 7))[0, 0, 0, 0])


def fepegar_case():
    """unet.UNet is third-party and absent: these vectors come from the oracle restatement
    itself (regression anchor, 'parity unpinned'); strict key coverage of every shipped
    checkpoint is what is pinned."""
    import glob
    names = sorted(glob.glob(refload.path("segmentation/weights/*.pth")))
    want = set(weights.fepegar_unet_state(8).keys())
    for p in names:
        sd = torch.load(p, map_location="cpu", weights_only=True)
        assert set(sd.keys()) == want, p
        ref_shapes = {k: tuple(v.shape) for k, v in weights.fepegar_unet_state(8).items()}
        assert all(tuple(sd[k].shape) == ref_shapes[k] for k in want), p
    sd = torch.load(refload.path("segmentation/weights/whole_im_train_seg_parc_epoch_7.pth"), map_location="cpu", weights_only=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, 64, 64, 64, generator=g)
    with torch.no_grad():
        logits = graphs.fepegar_unet(sd, x)
    am = logits.argmax(1).numpy().astype(np.uint8)
    save("fepegar_kat2_eval_UNPINNED", logits_sum=np.array(float(logits.double().sum())), logits_absmean=np.array(float(logits.abs().mean())),
         fg=np.array(int(am.sum())), argmax_sha=np.array(sha16(am)), logits_slice=logits[0, :, 32], n_checkpoints=np.array(len(names)))


def copy_fixtures():
    for rel in ("classification/encoder_93_6_4.pth", "classification/clf_93_6_4.pth", "classification/disc_93_6_4.pth",
                "segmentation/weights/whole_im_train_seg_parc_epoch_7.pth", "detection/MNI152_T1_1mm_brain_gray.nii.gz"):
        dst = os.path.join(OUT, os.path.basename(rel))
        shutil.copyfile(refload.path(rel), dst)
        os.chmod(dst, 0o644)
        print("  copied", rel)


def histstd_cases():
    """Histogram standardisation (classification/train_ENC_CLF.ipynb [cell 9] `normalize`) run through the NOTEBOOK's own
    function with the shipped landmarks (classification/fcd_train_data_landmarks.npy): pins percentiles and mapped volumes."""
    from oracle import preprocess
    ref = refload.histstd_functions()
    landmarks = np.load(refload.path("classification/fcd_train_data_landmarks.npy"))
    assert landmarks.shape == (13,)
    rng = np.random.default_rng(5)
    vols = {}
    brain = rng.gamma(2.0, 120.0, (24, 28, 20)).astype(np.float32)            # skewed intensities ...
    brain[:6] = 0; brain[:, :5] = 0; brain[:, :, 15:] = 0                      # ... inside a zero background (55 % of the voxels)
    vols["brain"] = brain
    vols["dense"] = rng.normal(300.0, 80.0, (17, 19, 23)).astype(np.float32)   # odd sizes, negative values possible
    vols["const"] = np.full((8, 8, 8), 3.5, np.float32)                        # every landmark equal: diff_perc < epsilon everywhere
    vols["steps"] = np.repeat(np.arange(16, dtype=np.float32), 100).reshape(16, 10, 10) * 7.0      # heavy ties
    out = {"landmarks": landmarks}
    for name, v in vols.items():
        want = ref["normalize"](torch.from_numpy(v), landmarks).numpy()
        pv = np.percentile(v.reshape(-1), ref["_get_percentiles"](100 * np.array(ref["_standardize_cutoff"](ref["DEFAULT_CUTOFF"]))))
        assert np.array_equal(preprocess.normalize(v, landmarks), want), f"oracle normalize != notebook ({name})"
        assert np.array_equal(preprocess.percentile_values(v), pv)
        out[f"{name}_x"], out[f"{name}_y"], out[f"{name}_pct"] = v, want, pv
        print(f"  {name}: percentiles {pv[[0, 6, 12]]}, out range [{want.min():.3f}, {want.max():.3f}]")
    m = vols["brain"] > 0                                                       # the `mask` argument (foreground only)
    want = ref["normalize"](torch.from_numpy(vols["brain"]), landmarks, mask=m).numpy()
    assert np.array_equal(preprocess.normalize(vols["brain"], landmarks, mask=m), want)
    out["brain_masked_y"] = want
    want = ref["normalize"](torch.from_numpy(vols["dense"]), landmarks, cutoff=(0.05, 0.95)).numpy()     # _standardize_cutoff clamps to (0.05, 0.95) -> 5 / 95
    assert np.array_equal(preprocess.normalize(vols["dense"], landmarks, cutoff=(0.05, 0.95)), want)
    out["dense_cut_y"] = want
    save("histstd_cell9", **out)


def metrics_cases():
    """Dice / IoU of the validation loop through the REFERENCE's own functions (segmentation/metrics.py:312-329,
    segmentation/routine.py:198-204) on uint8 label volumes, incl. empty masks and labels > 1."""
    from oracle import metrics as M
    ref_metrics = refload._load("_ref_metrics", "segmentation/metrics.py")
    routine = refload.seg_routine_module()
    rng = np.random.default_rng(8)
    out = {}
    cases = {"blobs": ((rng.random((24, 30, 22)) > 0.7).astype(np.uint8), (rng.random((24, 30, 22)) > 0.6).astype(np.uint8)),
             "labels5": (rng.integers(0, 5, (17, 19, 13)).astype(np.uint8), rng.integers(0, 5, (17, 19, 13)).astype(np.uint8)),
             "disjoint": (np.pad(np.ones((4, 4, 4), np.uint8), ((0, 8), (0, 8), (0, 8))), np.pad(np.ones((4, 4, 4), np.uint8), ((8, 0), (8, 0), (8, 0)))),
             "pred_empty": (np.zeros((9, 9, 9), np.uint8), (rng.random((9, 9, 9)) > 0.5).astype(np.uint8))}
    for name, (pred, gt) in cases.items():
        dsc, iou = ref_metrics.compute_dice_coefficient(gt, pred), routine.get_iou_score(pred, gt)
        assert M.compute_dice_coefficient(gt, pred) == dsc and M.get_iou_score(pred, gt) == iou and type(M.get_iou_score(pred, gt)) is type(iou)
        out[f"{name}_pred"], out[f"{name}_gt"], out[f"{name}_dsc"], out[f"{name}_iou"] = pred, gt, np.array(dsc), np.array(iou)
        print(f"  {name}: dice {dsc:.6f}  iou {iou:.6f} ({type(iou).__name__})")
    save("overlap_metrics", **out)


if __name__ == "__main__":
    assert refload.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["fixtures", "ops", "unet3d", "ae", "fader", "fepegar", "patches", "detect", "histstd", "metrics", "modified_unet", "cnn_model", "patch_nodrop", "grid", "surface"]
    table = dict(fixtures=copy_fixtures, ops=op_pins, unet3d=unet3d_cases, ae=ae_cases, fader=fader_cases,
                 fepegar=fepegar_case, patches=patch_cases, modified_unet=modified_unet_cases, cnn_model=cnn_model_cases, patch_nodrop=patch_model_nodrop_case, grid=grid_case, surface=surface_cases, detect=detect_cases, histstd=histstd_cases, metrics=metrics_cases)
    for w in which:
        print(w)
        table[w]()
