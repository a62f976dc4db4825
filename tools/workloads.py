"""The other BASELINE.json configurations and the 'next' rows (SURVEY section 8 d/f) as timed workloads on ONE GPU.

bench.py calls `run_all()` at N = 1 and prints the result under "other_configs", so that every configuration has a number from the
same driver-run process as the headline; `python tools/workloads.py [names]` prints them alone.  CUDA events, >= 3 warm-up
iterations, inputs larger than L2 (or stated otherwise).  HBM-bound workloads report the bytes their layers have to move once
(`layer_io_bytes`: sum over leaf modules of input + output bytes in the dtype they run in, x3 for forward + backward) against the
measured copy bandwidth -- a whole-step figure, not one kernel's.
"""
from __future__ import annotations

import gzip
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import numpy as np
import torch


def _timeit(fn, iters=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _layer_io_bytes(net, *inputs):
    """forward bytes every leaf module reads + writes once (activations only), measured with hooks on one eager forward"""
    total, hooks = [0], []

    def hook(m, args, out):
        for t in list(args) + (list(out) if isinstance(out, (tuple, list)) else [out]):
            if torch.is_tensor(t) and t.dim() >= 4:
                total[0] += t.numel() * t.element_size()
    for m in net.modules():
        if next(m.children(), None) is None:
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        net(*inputs)
    for h in hooks:
        h.remove()
    return total[0]


def make_config1(pkg, dev, world=1):
    """BASELINE config 1: conv3d autoencoder (train_AE.ipynb [cell 8]: depth 6, c_base 16) fwd + MSE + bwd + Adam, batch 2 x 128^3.
    Returns the pieces bench.py and the timing wrapper below share: step(x_dev) -> loss tensor, the pinned host batch, voxel and
    layer-I/O byte counts."""
    torch.manual_seed(0)
    net = pkg.convert(pkg.zoo.config1_autoencoder(depth=6, c_base=16).to(dev).train(), dtype=torch.bfloat16)
    xh = torch.randn(2, 1, 128, 128, 128).pin_memory()
    x = xh.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=True, fused=True)
    io = _layer_io_bytes(net, x)
    if world > 1:
        pkg.dp.broadcast_parameters(net)
    if os.environ.get("B200_WORKLOAD_EAGER") == "1":           # for ncu launch lists: plain eager launches
        def step(a):
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(net(a).float(), a)
            loss.backward()
            opt.step()
            return loss.detach()
    else:
        g = pkg.graphed.GraphedTrainStep(net, lambda y, t: torch.nn.functional.mse_loss(y.float(), t), opt, x, x)
        step = lambda a: g(a, a)
        step.graph = g
    return {"step": step, "x_host": (xh,), "x_dev": (x,), "voxels": x.numel(), "layer_io_bytes_fwd": io, "bytes_per_step": 3 * io,
            "workload": "AE depth 6 c_base 16, batch 2 x 128^3 per GPU, fwd + MSE + bwd + Adam, bf16 body, cuda-graph replay",
            "keep": (net, opt)}


def config1_autoencoder(pkg, dev, hbm_gbs):
    w = make_config1(pkg, dev)
    ms = _timeit(lambda: w["step"](*w["x_dev"]))
    gbs = w["bytes_per_step"] / ms / 1e6
    return {"workload": w["workload"], "ms_per_step": ms, "value": w["voxels"] / ms * 1e3, "unit": "voxels/s",
            "layer_io_bytes_fwd": w["layer_io_bytes_fwd"], "hbm_gbs_achieved": gbs, "hbm_frac": gbs / hbm_gbs}


def config3_full_volume(pkg, dev, hbm_gbs):
    """BASELINE config 3, the per-GPU share: unet3d train step on one 192 x 224 x 192 volume"""
    from mri_epilepsy_diagnosis_b200 import functional as BF
    torch.manual_seed(0)
    net = pkg.convert(pkg.zoo.Unet(c=1, n=16, dropout=0.5, norm="bn", num_classes=2).to(dev).train(), dtype=torch.bfloat16)
    x = torch.randn(1, 1, 192, 224, 192, device=dev)
    t = (torch.rand(1, 1, 192, 224, 192, device=dev) > 0.5).float()
    opt = torch.optim.AdamW(net.parameters(), capturable=True, fused=True)
    step = pkg.graphed.GraphedTrainStep(net, BF.softmax_dice_loss, opt, x, t)
    ms = _timeit(lambda: step(x, t))
    return {"workload": "unet3d.Unet(c=1,n=16,bn) train step, 1 x 192x224x192, bf16, cuda-graph replay", "ms_per_step": ms,
            "value": x.numel() / ms * 1e3, "unit": "voxels/s", "model_tflops": 518186.0 * x.numel() / ms / 1e9}


def make_config4(pkg, dev, world=1):
    """BASELINE config 4, the per-GPU share: fader encoder + classifier + discriminator step (train_ENC_CLF.ipynb [cell 14, 16], n_d = 1),
    batch 8 x 192^3 per GPU (global batch 64 on 8 GPUs), shapes of the shipped *_93_6_4.pth.  Data parallel: dp.attach on both optimizers."""
    torch.manual_seed(0)
    BF16 = torch.bfloat16
    enc = pkg.convert(pkg.zoo.fader_encoder().to(dev), dtype=BF16)
    clf = pkg.convert(pkg.zoo.Classificator(n_class=2, **pkg.zoo.FADER_HEAD).to(dev), dtype=BF16)
    disc = pkg.convert(pkg.zoo.Discriminator(n_domains=18, **pkg.zoo.FADER_HEAD).to(dev), dtype=BF16)
    B = 8
    xh = torch.randn(B, 1, 192, 192, 192).pin_memory()
    x = xh.to(dev)
    y = torch.randint(0, 2, (B,), device=dev)
    dom = torch.randint(0, 18, (B,), device=dev)
    graphed = world == 1 and os.environ.get("B200_WORKLOAD_EAGER") != "1"       # one GPU: the whole two-optimizer step is captured
    opt_e = torch.optim.Adam(list(enc.parameters()) + list(clf.parameters()), lr=7e-4, weight_decay=1e-4, capturable=graphed)
    opt_d = torch.optim.Adam(disc.parameters(), lr=5e-4, weight_decay=1e-4, capturable=graphed)
    ce_y = torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 2.0], device=dev))
    ce_d = torch.nn.CrossEntropyLoss()
    io = _layer_io_bytes(enc, x)
    buckets = []
    if world > 1:
        for m in (enc, clf, disc):
            pkg.dp.broadcast_parameters(m)
        buckets = [pkg.dp.GradientBucket(list(enc.parameters()) + list(clf.parameters()), opt_e, buckets=1),
                   pkg.dp.GradientBucket(list(disc.parameters()), opt_d, buckets=1)]

    def step(xb, lam=0.1):
        enc.eval(); disc.train()
        with torch.no_grad():
            lat = enc(xb)[0]
        opt_d.zero_grad()
        ce_d(disc(lat), dom).backward()
        opt_d.step()
        enc.train(); clf.train(); disc.eval()
        for p in disc.parameters():
            p.requires_grad = False
        opt_e.zero_grad()
        lat = enc(xb)[0]
        logp = torch.log_softmax(disc(lat), dim=1)
        adv = -(torch.ones_like(logp) / 18.0 * logp).sum(1).mean()
        loss = ce_y(clf(lat), y) + lam * adv
        loss.backward()
        opt_e.step()
        for p in disc.parameters():
            p.requires_grad = True
        return loss.detach()
    launch = "eager launches"
    run = step
    per_replay = None
    if graphed:
        # the reference's loop body (train_ENC_CLF.ipynb [cell 16]: discriminator step, then encoder + classifier step) captured once:
        # ~800 launches of 5-600 us each are launch-bound when enqueued from Python
        try:
            static_x = x.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step(static_x)
                l0 = pkg.launch_count()
                step(static_x)
                per_replay = pkg.launch_count() - l0          # library kernels of one step = of one replay
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step(static_x)

            def run(xb, _g=graph, _x=static_x, _l=static_loss):
                if xb.data_ptr() != _x.data_ptr():
                    _x.copy_(xb, non_blocking=True)
                _g.replay()
                return _l
            launch = "cuda-graph replay"
        except Exception as e:          # a step that cannot be captured is reported and run eagerly
            launch = f"eager launches (capture failed: {type(e).__name__})"
            per_replay = None
            torch.cuda.synchronize(dev)
    # 2 encoder forwards + 1 encoder backward (~2 forwards' worth of bytes) per step
    return {"step": run, "x_host": (xh,), "x_dev": (x,), "voxels": x.numel(), "layer_io_bytes_fwd": io, "bytes_per_step": 4 * io,
            "workload": f"fader enc+clf+disc step (n_d=1), batch 8 x 192^3 per GPU, bf16 body, {launch}", "launches_per_step": per_replay,
            "keep": (enc, clf, disc, opt_e, opt_d, buckets)}


def config4_fader(pkg, dev, hbm_gbs):
    w = make_config4(pkg, dev)
    ms = _timeit(lambda: w["step"](*w["x_dev"]), iters=4)
    gbs = w["bytes_per_step"] / ms / 1e6
    return {"workload": w["workload"], "ms_per_step": ms, "value": w["voxels"] / ms * 1e3, "unit": "voxels/s",
            "layer_io_bytes_fwd": w["layer_io_bytes_fwd"], "hbm_gbs_achieved": gbs, "hbm_frac": gbs / hbm_gbs}


def config5_detection(pkg, dev, hbm_gbs):
    """BASELINE config 5: sliding-window patch extraction over the MNI152 1 mm template + batched PatchModel inference + vote + paint"""
    raw = gzip.open(os.path.join(ROOT, "tests", "golden", "MNI152_T1_1mm_brain_gray.nii.gz")).read()
    gm = np.frombuffer(raw, dtype="<f4", offset=352, count=182 * 218 * 182).reshape((182, 218, 182), order="F").astype(np.float64)
    img = np.random.default_rng(0).random((182, 218, 182))
    gm_d, img_d = torch.as_tensor(gm, device=dev), torch.as_tensor(img, device=dev)
    n = pkg.patches.get_only_patches(img_d, gm_d, 16, 32).shape[0]
    ms_gather = _timeit(lambda: pkg.patches.get_only_patches(img_d, gm_d, 16, 32), iters=10)
    pm = pkg.convert(pkg.zoo.PatchModel().to(dev).eval(), dtype=torch.bfloat16)
    gen = pkg.detect.FCDMaskGenerator(pm, gm_d, batch=4752)
    ms_mask = _timeit(lambda: gen.get_mask(img_d), iters=5)
    return {"workload": f"get_only_patches on MNI152 1 mm ({n} patches) / FCDMaskGenerator.get_mask (plan + gather + PatchModel + vote + paint)",
            "patches": n, "gather_ms": ms_gather, "gather_patches_per_s": n / ms_gather * 1e3, "get_mask_ms": ms_mask, "volumes_per_s": 1e3 / ms_mask,
            "gather_hbm_frac": (2 * 7.22e6 * 8 + n * 2 * 16 * 32 * 8) / ms_gather / 1e6 / hbm_gbs}


def next_rows(pkg, dev, hbm_gbs):
    """rows f-1 / f-2 / f-3 at the size of one MNI-registered volume"""
    from scipy import ndimage
    out = {}
    lm = np.load(os.path.join(ROOT, "tests", "golden", "histstd_cell9.npz"))["landmarks"]
    vol = torch.empty(192, 192, 192, device=dev).exponential_(0.01)
    vol[:40] = 0
    out["f1_histogram_standardisation_192cube_ms"] = _timeit(lambda: pkg.preprocess.normalize(vol, lm), iters=20)
    rng = np.random.default_rng(3)
    f = ndimage.gaussian_filter(rng.random((192, 224, 192)).astype(np.float32), 6.0)
    f = (f - f.min()) / (f.max() - f.min())
    gt = torch.from_numpy((f > 0.55).astype(np.uint8)).to(dev)
    pred = torch.from_numpy((np.roll(f, 2, 1) > 0.56).astype(np.uint8)).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        res = pkg.metrics.calculate_metrics(gt, pred)
    out["f2_calculate_metrics_192x224x192_ms"] = 1e3 * (time.perf_counter() - t0) / 3          # incl. the sort and the D2H of the surfel lists
    out["f2_values"] = [float(v) for v in res]
    net = pkg.convert(pkg.zoo.Unet(c=1, n=16, norm="bn", num_classes=2).to(dev).eval(), dtype=torch.bfloat16)
    sample = {"MRI": {"data": torch.randn(1, 192, 224, 192, device=dev)}}
    out["f3_sliding_window_64cube_overlap4_192x224x192_ms"] = _timeit(lambda: pkg.grid.sliding_window_labels(net, sample, 64, 4, batch_size=16), iters=3, warmup=2)
    return out


ALL = {"config1": config1_autoencoder, "config3": config3_full_volume, "config4": config4_fader, "config5": config5_detection, "next_rows": next_rows}


def run_all(pkg, dev, hbm_gbs, names=None):
    out = {}
    for name in names or ALL:
        try:
            out[name] = ALL[name](pkg, dev, hbm_gbs)
        except Exception as e:          # a workload that cannot run is reported, not hidden
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    import __graft_entry__
    pkg = __graft_entry__.build()
    torch.cuda.set_device(0)
    print(json.dumps(run_all(pkg, torch.device("cuda", 0), 6544.7, sys.argv[1:] or None), indent=1))
