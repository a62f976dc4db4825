#!/usr/bin/env python
"""Throughput of every BASELINE.json config on ONE GPU (CUDA events, eager launches), next to bench.py's headline config 2.
   python tools/bench_configs.py [cfg ...]     cfg in: c1 c3 c4 c5 fp8 fp16 inf
   c1  conv3d autoencoder fwd+bwd, batch 2 x 128^3 (bf16 body, MSE loss)                      voxels/s
   c3  unet3d train step on one 192x224x192 volume (per-GPU share of the DP config)          voxels/s
   c4  fader encoder+classifier+discriminator step (train_ENC_CLF.ipynb [cell 16], n_d = 1), batch 8 x 192^3 per GPU
   c5  sliding-window patch gather on the MNI152 template + batched PatchModel inference     patches/s
   fp8/fp16  unet.UNet(first=8/16) train step, batch 4 x 128^3                               voxels/s
   inf unet3d inference (eval, no grad), batch 4 x 128^3                                      voxels/s
"""
import gzip
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__

pkg = __graft_entry__.build()
from mri_epilepsy_diagnosis_b200 import functional as BF  # noqa: E402

dev = torch.device("cuda", 0)
BF16 = torch.bfloat16


PROFILE = os.environ.get("B200_PROFILE_ONE_STEP") == "1"      # under ncu --profile-from-start off: capture exactly one step


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if PROFILE:
        torch.cuda.cudart().cudaProfilerStart()
        fn()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return 1.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, units, unit):
    print(f"{name:72s} {ms:9.2f} ms/step  {units / ms * 1e3 / 1e6:10.2f} M{unit}/s", flush=True)


def unet_step(net, x, t, opt):
    def step():
        opt.zero_grad()
        loss = BF.softmax_dice_loss(net(x), t)
        loss.backward()
        opt.step()
    return step


cfgs = sys.argv[1:] or ["c1", "c3", "c4", "c5", "fp8", "fp16", "inf", "pre", "val"]
torch.manual_seed(0)
if "c1" in cfgs:
    net = pkg.convert(pkg.zoo.config1_autoencoder(depth=6, c_base=16).to(dev).train(), dtype=BF16)
    x = torch.randn(2, 1, 128, 128, 128, device=dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)

    def step():
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(net(x).float(), x)
        loss.backward()
        opt.step()
    report("c1 AE depth 6 c_base 16, batch 2 x 128^3, fwd+bwd+Adam (bf16)", timeit(step), x.numel(), "voxel")
    del net, opt
if "c3" in cfgs:
    net = pkg.convert(pkg.zoo.Unet(c=1, n=16, dropout=0.5, norm="bn", num_classes=2).to(dev).train(), dtype=BF16)
    x = torch.randn(1, 1, 192, 224, 192, device=dev)
    t = (torch.rand(1, 1, 192, 224, 192, device=dev) > 0.5).float()
    opt = torch.optim.AdamW(net.parameters())
    report("c3 unet3d train step, 1 x 192x224x192 (per-GPU share)", timeit(unet_step(net, x, t, opt)), x.numel(), "voxel")
    del net, opt
for name, first in (("fp8", 8), ("fp16", 16)):
    if name in cfgs:
        net = pkg.convert(pkg.zoo.FepegarUNet(out_channels_first_layer=first).to(dev).train(), dtype=BF16)
        x = torch.randn(4, 1, 128, 128, 128, device=dev)
        t = (torch.rand(4, 1, 128, 128, 128, device=dev) > 0.5).float()
        opt = torch.optim.AdamW(net.parameters())
        report(f"unet.UNet(first={first}) train step, batch 4 x 128^3", timeit(unet_step(net, x, t, opt)), x.numel(), "voxel")
        del net, opt
if "inf" in cfgs:
    net = pkg.convert(pkg.zoo.Unet(c=1, n=16, dropout=0.5, norm="bn", num_classes=2).to(dev).eval(), dtype=BF16)
    x = torch.randn(4, 1, 128, 128, 128, device=dev)
    with torch.no_grad():
        report("unet3d inference (eval), batch 4 x 128^3 + argmax", timeit(lambda: net(x).argmax(dim=1)), x.numel(), "voxel")
    del net
if "c4" in cfgs:
    enc = pkg.convert(pkg.zoo.fader_encoder().to(dev), dtype=BF16)
    clf = pkg.convert(pkg.zoo.Classificator(n_class=2, **pkg.zoo.FADER_HEAD).to(dev), dtype=BF16)
    disc = pkg.convert(pkg.zoo.Discriminator(n_domains=18, **pkg.zoo.FADER_HEAD).to(dev), dtype=BF16)
    B = 8
    x = torch.randn(B, 1, 192, 192, 192, device=dev)
    y = torch.randint(0, 2, (B,), device=dev)
    dom = torch.randint(0, 18, (B,), device=dev)
    opt_e = torch.optim.Adam(list(enc.parameters()) + list(clf.parameters()), lr=7e-4, weight_decay=1e-4)
    opt_d = torch.optim.Adam(disc.parameters(), lr=5e-4, weight_decay=1e-4)
    ce_y = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0], device=dev))
    ce_d = torch.nn.CrossEntropyLoss()

    def fader_step(lam=0.1):
        # discriminator step on a detached latent (encoder in eval mode, frozen), then the encoder+classifier step with the
        # adversarial term -- classification/train_ENC_CLF.ipynb [cell 14, 16]
        enc.eval(); disc.train()
        with torch.no_grad():
            lat = enc(x)[0]
        opt_d.zero_grad()
        ce_d(disc(lat), dom).backward()
        opt_d.step()
        enc.train(); clf.train(); disc.eval()
        for p in disc.parameters():
            p.requires_grad = False
        opt_e.zero_grad()
        lat = enc(x)[0]
        logp = torch.log_softmax(disc(lat), dim=1)
        adv = -(torch.ones_like(logp) / 18.0 * logp).sum(1).mean()           # push the discriminator towards the uniform posterior
        (ce_y(clf(lat), y) + lam * adv).backward()
        opt_e.step()
        for p in disc.parameters():
            p.requires_grad = True
    report("c4 fader enc+clf+disc step (n_d=1), batch 8 x 192^3 (bf16)", timeit(fader_step, iters=3), x.numel(), "voxel")
    del enc, clf, disc
if "c5" in cfgs:
    raw = gzip.open(os.path.join(ROOT, "tests", "golden", "MNI152_T1_1mm_brain_gray.nii.gz")).read()
    gm = np.frombuffer(raw, dtype="<f4", offset=352, count=182 * 218 * 182).reshape((182, 218, 182), order="F").astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    gm_d = torch.as_tensor(gm, device=dev)
    img_d = torch.as_tensor(img, device=dev)
    patches = pkg.patches.get_only_patches(img_d, gm_d, 16, 32)
    n = patches.shape[0]
    report(f"c5 patch gather on MNI152 1mm ({n} patches, device-resident volumes)", timeit(lambda: pkg.patches.get_only_patches(img_d, gm_d, 16, 32)), n, "patch")
    t0 = time.perf_counter()
    pkg.patches.get_only_patches(img, gm, 16, 32)
    torch.cuda.synchronize()
    print(f"   (same call from host numpy volumes incl. H2D: {1e3 * (time.perf_counter() - t0):.1f} ms)")
    pm = pkg.convert(pkg.zoo.PatchModel().to(dev).eval(), dtype=BF16)
    pf = patches.float()
    with torch.no_grad():
        report(f"c5 PatchModel inference, all {n} patches in one batch + argmax", timeit(lambda: pm(pf).argmax(dim=1)), n, "patch")

if "pre" in cfgs:
    # f-1: the collate-time histogram standardisation of classification/train_ENC_CLF.ipynb [cell 9] on one 192^3 volume
    # (the loaders' size): device kernels vs the numpy primitives the notebook's function spends its time in (oracle/ is test
    # infrastructure and is not imported by tools)
    lm = np.load(os.path.join(ROOT, "tests", "golden", "histstd_cell9.npz"))["landmarks"]
    g = torch.Generator(device="cuda").manual_seed(0)
    vol = torch.empty(192, 192, 192, device=dev).exponential_(0.01, generator=g)
    vol[:40] = 0
    ms = timeit(lambda: pkg.preprocess.normalize(vol, lm), iters=20, warmup=3)
    nbytes = vol.numel() * 4 * 5
    report("f-1 histogram standardisation, one 192^3 volume (3 select passes + map)", ms, vol.numel(), "voxel")
    print(f"   ({nbytes / ms / 1e6:.0f} GB/s of the 5 x 4 B per voxel the passes move = {100 * nbytes / ms / 1e6 / 6544.7:.0f} % of the HBM peak)")
    host = vol.cpu().numpy()
    t0 = time.perf_counter()
    flat = host.reshape(-1)
    pv = np.percentile(flat, [1, 10, 20, 25, 30, 40, 50, 60, 70, 75, 80, 90, 99])
    bins = np.digitize(flat, pv[[1, 2, 4, 5, 6, 7, 8, 10, 11]])
    _ = (np.linspace(0.5, 1.5, 10)[bins] * flat + np.linspace(0.0, 9.0, 10)[bins]).astype(np.float32)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    print(f"   (the same work in numpy on the host -- np.percentile + np.digitize + float64 affine map: {cpu_ms:.0f} ms per volume = {cpu_ms / ms:.0f}x)")

if "val" in cfgs:
    # f-2 (counting part): Dice + IoU of one predicted label volume against the ground truth, 192x224x192 (config 3's volume)
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = (torch.rand(192, 224, 192, device=dev, generator=g) > 0.8).to(torch.uint8)
    gt = (torch.rand(192, 224, 192, device=dev, generator=g) > 0.7).to(torch.uint8)
    out = torch.empty(5, dtype=torch.int64, device=dev)
    lib = pkg._cabi.lib()
    ms = timeit(lambda: lib.b200_overlap_counts(pred.data_ptr(), gt.data_ptr(), pred.numel(), out.data_ptr(), pkg._cabi.stream()), iters=50, warmup=5)
    report("f-2 overlap counts (Dice + IoU), one 192x224x192 label pair", ms, pred.numel(), "voxel")
    print(f"   ({2 * pred.numel() / ms / 1e6:.0f} GB/s of the 2 B per voxel read = {100 * 2 * pred.numel() / ms / 1e6 / 6544.7:.0f} % of the HBM peak; the pair is 16.5 MB, L2-resident)")
    p, t = pred.cpu().numpy(), gt.cpu().numpy()
    t0 = time.perf_counter()
    _ = 2 * (t & p).sum() / (t.sum() + p.sum())
    _ = float(np.logical_and(p > 0, t > 0).astype(np.float32).sum()) / np.logical_or(p > 0, t > 0).astype(np.float32).sum()
    print(f"   (the same two expressions in numpy on the host: {1e3 * (time.perf_counter() - t0):.1f} ms; host API incl. the 40-byte read-back: "
          f"{timeit(lambda: pkg.metrics.calculate_overlap(gt, pred), iters=20, warmup=3):.3f} ms)")
