#!/usr/bin/env python
"""Headline benchmark: 3-D U-Net training throughput (voxels/s) -- BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model unet3d|fepegar16|fepegar8]

A step = one pass of the hot path over one batch of synthetic input, exactly the body of the reference's
training loop (segmentation/routine.py:266-281): zero_grad -> model(x) -> softmax -> dice loss -> mean -> backward ->
AdamW.step.  Workload at every N: `unet3d.Unet(c=1, n=16, norm='bn', num_classes=2)`, batch 4 x 128^3 per GPU, bf16
activations (weak scaling: per-GPU batch fixed; gradients all-reduced over NCCL).

One JSON line on stdout (rank 0):
  value          voxels/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e            the same step called with HOST (pinned) inputs: H2D copy of the batch and D2H read of the loss inside
  roofline       tensor-pipe roofline of the dominant kernel (conv_umma_kernel): algorithmic conv FLOPs of its launches
                 in the timed region / their CUDA-event durations, against MEASURED_PEAKS.json (sustained bf16)
  cpu_baseline   the oracle (CPU restatement of the reference path, oracle/graphs.py) timed on this box's host cores
  --impl reference  times that CPU path alone (the reference's own implementation of the path is its CPU PyTorch path)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "3D U-Net train voxels/s"
FWD_BWD_FLOP_PER_VOXEL = 518186.0      # unet3d.Unet(c=1,n=16,num_classes=2): SURVEY section 8(a-1) / BASELINE.md section 3


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": float(p.get("bf16_tflops_sustained") or p["bf16_tflops"]), "hbm_gbs": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:
        return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}     # B200_PROFILING.md fallback (sustained)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(n, size, seed):
    """size: edge of a cubic volume or (D, H, W)"""
    dhw = (size,) * 3 if isinstance(size, int) else tuple(size)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 1, *dhw, generator=g)
    t = (torch.rand(n, 1, *dhw, generator=g) > 0.5).float()
    return x, t


def dice_loss_mean(logits, targets, eps=1e-9):
    """segmentation/routine.py:272-274 + :239-253 (device-side loss, stays PyTorch: SURVEY a-8)."""
    p0 = torch.softmax(logits, dim=1)
    p1, g0 = 1 - p0, targets
    g1 = 1 - g0
    tp = (p0 * g0).sum(dim=(2, 3, 4)); fp = (p0 * g1).sum(dim=(2, 3, 4)); fn = (p1 * g0).sum(dim=(2, 3, 4))
    return (1 - 2 * tp / (2 * tp + fp + fn + eps)).mean()


def build_model(pkg, name, norm="bn", literal=False):
    if name == "unet3d":
        return (pkg.zoo.Unet(c=1, n=16, dropout=0.5, norm=norm, num_classes=2, literal=literal),
                f"unet3d.Unet(c=1,n=16,norm={norm},num_classes=2)" + (" [literal op sequence]" if literal else ""))
    if name == "fepegar16":
        return pkg.zoo.FepegarUNet(out_channels_first_layer=16), "unet.UNet(first=16)"
    if name == "fepegar8":
        return pkg.zoo.FepegarUNet(out_channels_first_layer=8), "unet.UNet(first=8)"
    raise SystemExit(f"unknown model {name}")


def _oracle_state(model_name, norm):
    from oracle import graphs, weights
    if model_name == "unet3d":
        sd = weights.unet3d_state(1, 16, 2, norm, seed=0)
        fwd = lambda s, x: graphs.unet3d(s, x, norm, 0.5, True)
    else:
        sd = weights.fepegar_unet_state(16 if model_name == "fepegar16" else 8, seed=0, duplicate_keys=False)
        fwd = lambda s, x: graphs.fepegar_unet(s, x, True)
    return sd, fwd


def cpu_reference_step(model_name, size, steps, warmup, threads, norm="bn", batch=1, budget_s=None):
    """The reference's CPU path for the step (oracle restatement), fp32, on `threads` host threads.  With `budget_s` the
    number of timed steps is cut so that the run ends in about that many seconds; returns (voxels per step, times of the timed steps,
    warm-up steps actually run)."""
    from oracle import graphs
    torch.set_num_threads(threads)
    sd, fwd = _oracle_state(model_name, norm)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad])
    x, t = synthetic_batch(batch, size, 0)
    times, began = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = graphs.dice_loss_mean(fwd(sd, x), t)
        loss.backward()
        opt.step()
        float(loss)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            if budget_s is not None and time.perf_counter() - began + dt > budget_s:
                break
    return x.numel(), times, warmup


def torch_gpu_baseline(model_name, batch, size, steps, warmup, norm, dev):
    """Stock PyTorch on the SAME B200 for the same step -- what the unmodified reference modules execute on a GPU: ATen/cuDNN
    convolutions, eager BatchNorm/ReLU/cat/interpolate, autograd, torch.optim.AdamW.  The graph is the oracle's functional
    restatement of unet3d.py (torch.nn.functional calls, identical operator sequence incl. the dead branch), moved to CUDA.
    Variants: fp32 (PyTorch defaults, i.e. TF32 convolutions allowed) and bf16 autocast, NCDHW and channels_last_3d, eager and
    whole-step CUDA graph; cudnn.benchmark on.  Nothing of this repo's library runs here."""
    from oracle import graphs
    torch.backends.cudnn.benchmark = True
    x, t = synthetic_batch(batch, size, 0)
    voxels = x.numel()
    out = {}

    def variant(name, autocast, channels_last, graphed):
        sd0, fwd = _oracle_state(model_name, norm)
        sd = {}
        for k, v in sd0.items():
            v = v.to(dev)
            if channels_last and v.dim() == 5:
                v = v.contiguous(memory_format=torch.channels_last_3d)
            sd[k] = v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v
        opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad], capturable=graphed)
        xd, td = x.to(dev), t.to(dev)
        if channels_last:
            xd = xd.contiguous(memory_format=torch.channels_last_3d)

        def body():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                logits = fwd(sd, xd)
            loss = graphs.dice_loss_mean(logits.float(), td)
            loss.backward()
            opt.step()
            return loss
        run = None
        if graphed:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    opt.zero_grad(set_to_none=True)
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                body()
            run = graph.replay
        else:
            def run():
                opt.zero_grad(set_to_none=True)
                body()
        for _ in range(max(3, warmup)):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": ms, "value": voxels / (ms / 1e3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}

    for name, ac, cl, gr in (("bf16_autocast_channels_last_3d_graph", True, True, True), ("bf16_autocast_channels_last_3d_eager", True, True, False),
                             ("bf16_autocast_ncdhw_graph", True, False, True), ("fp32_tf32_ncdhw_graph", False, False, True)):
        try:
            torch.cuda.reset_peak_memory_stats()
            variant(name, ac, cl, gr)
        except Exception as e:      # a variant PyTorch cannot run (e.g. capture failure) is reported, not hidden
            out[name] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    ok = {k: v for k, v in out.items() if "value" in v}
    best = max(ok.items(), key=lambda kv: kv[1]["value"]) if ok else (None, {})
    return {"what": "stock torch " + torch.__version__ + " / cuDNN " + str(torch.backends.cudnn.version()) + " on this GPU, same step, same batch",
            "best_variant": best[0], "value": best[1].get("value"), "ms_per_step": best[1].get("ms_per_step"), "unit": "voxels/s", "variants": out}


def other_config(pkg, args, rank, world, local, base):
    """BASELINE configs 1 and 4 through the same contract as the headline: W warm-up steps, K timed steps between barriers, max over
    ranks, one JSON line; HBM roofline = layer-I/O bytes of the step (tools/workloads.py) / time against the measured copy bandwidth."""
    from tools import workloads
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    w = (workloads.make_config1 if args.config == 1 else workloads.make_config4)(pkg, dev, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms
    l0 = pkg.launch_count()
    graph = getattr(w["step"], "graph", None)
    if graph is not None:      # kernels of one replay = kernels enqueued by one eager execution of the captured body
        graph._body(eager=True)
        per_step = pkg.launch_count() - l0
    for _ in range(max(3, args.warmup)):
        w["step"](*w["x_dev"])
    if w.get("launches_per_step"):          # a workload that captured its own step reports the launches of one replay
        per_step = w["launches_per_step"]
    elif graph is None:
        per_step = (pkg.launch_count() - l0) // max(3, args.warmup)
    copy = torch.cuda.Stream()

    def fetch():               # H2D of the next batch on a copy stream while the current step computes
        with torch.cuda.stream(copy):
            bufs = [h.to(dev, non_blocking=True) for h in w["x_host"]]
            ev = torch.cuda.Event()
            ev.record(copy)
        return bufs, ev

    def e2e_run(n):
        nxt = fetch()
        for i in range(n):
            bufs, ev = nxt
            if i + 1 < n:
                nxt = fetch()
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for b in bufs:
                b.record_stream(cur)
            float(w["step"](*bufs))
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    t0 = time.time()
    ms = timed(lambda: w["step"](*w["x_dev"]), args.steps)
    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1) / args.steps
    clocks = sampler.summary(t0, time.time()) if sampler else None
    pk = peaks()
    if rank == 0:
        gbs = w["bytes_per_step"] / ms / 1e6
        print(json.dumps({**base, "metric": "conv3d autoencoder train voxels/s" if args.config == 1 else "fader encoder+classifier+discriminator step voxels/s",
                          "value": world * w["voxels"] / (ms / 1e3), "ms_per_step": ms, "dtype": "bf16",
                          "config": {"workload": w["workload"], "baseline_config": args.config, "parallelism": f"dp{world}",
                                     "l2": "every 128^3 / 192^3 activation tensor exceeds the 126 MB L2"},
                          "e2e": {"value": world * w["voxels"] / (ms_e2e / 1e3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                                  "h2d_bytes_per_step": sum(h.numel() * h.element_size() for h in w["x_host"]), "d2h_bytes_per_step": 4},
                          "gpu_launches": int(per_step * args.steps), "clocks": clocks,
                          "roofline": {"bound": "hbm", "kernel": "whole step (axis_gather / axis_wgrad dominate, profiles/r2_config*_kernel_table.md)",
                                       "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
                                       "note": "achieved = bytes every layer of the step has to read + write once (forward layer I/O measured with hooks, "
                                               "x3 for forward + backward; x4 for the fader step's two encoder forwards) / step time"}}), flush=True)
    if dist is not None:
        import gc
        w = None
        dist.barrier(); torch.cuda.synchronize(); gc.collect()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


def main():
    if os.environ.get("B200_BENCH_WATCHDOG"):            # debugging aid: dump every thread's stack if the run has not finished in time
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["B200_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--model", default="unet3d")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--volume", default=None, help="D,H,W of a non-cubic volume (overrides --size), e.g. 192,224,192")
    ap.add_argument("--layer-table", default=None, help="write the per-layer (pass, kernel, shape) timing table of the roofline leg to this file")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4],
                    help="BASELINE.json config: 2 = unet3d batch 4 x 128^3 (default, the headline), 3 = one 192x224x192 volume per GPU (data parallel), "
                         "1 = conv3d autoencoder batch 2 x 128^3, 4 = fader encoder+classifier+discriminator step, batch 8 x 192^3 per GPU (HBM-bound: "
                         "their roofline block is bytes, not FLOPs)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock-PyTorch-on-this-GPU leg")
    ap.add_argument("--no-dropin-leg", action="store_true", help="skip the timing of the drop-in path (convert() of a stock model + eager loop)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short timings of BASELINE configs 1/3/4/5 and rows f-1..f-3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--no-literal-leg", action="store_true", help="skip the extra timing of the literal operator sequence")
    ap.add_argument("--literal-graph", action="store_true",
                    help="unet3d only: execute unet3d.py's operator sequence one to one (dead branch materialised, conv2 after the "
                         "upsample) instead of the equivalent rewritten graph")
    ap.add_argument("--norm", default="bn", choices=["bn", "in", "gn"], help="unet3d normalisation (BASELINE config 2 names bn and in)")
    ap.add_argument("--torch-loss", action="store_true", help="compute the Dice loss with torch ops instead of the fused kernel")
    ap.add_argument("--eager", action="store_true", help="do not capture the step in a CUDA graph (launch-bound at this size)")
    ap.add_argument("--profile-json", default=None, help="write the per-layer conv timing table here")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    if args.config == 3 and args.volume is None:
        args.volume, args.batch = "192,224,192", 1
    vol = tuple(int(v) for v in args.volume.split(",")) if args.volume else (args.size,) * 3
    vol_s = f"{vol[0]}^3" if vol[0] == vol[1] == vol[2] else "x".join(str(v) for v in vol)
    workload = f"{args.model} train step (fwd+dice+bwd+AdamW), batch {args.batch} x {vol_s} per GPU, bf16"
    base = {"metric": METRIC, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic (seeded randn volumes, random-init weights)"}

    # ------------------------------------------------------------------ reference arm: the CPU path on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        # the config's own batch (e.g. 4 x 128^3: about 5-6 s per step on 16 cores); the number of timed steps is cut to a
        # ~150 s budget and the line reports the steps / warm-up ACTUALLY run
        nvox, times, wu = cpu_reference_step(args.model, vol, max(1, args.steps), min(args.warmup, 1), cores, args.norm, batch=args.batch, budget_s=150.0)
        ms = 1e3 * sum(times) / len(times)
        val = nvox / (ms / 1e3)
        print(json.dumps({**base, "impl": "reference", "steps": len(times), "warmup": wu, "steps_requested": args.steps, "warmup_requested": args.warmup,
                          "value": val, "ms_per_step": ms, "dtype": "f32",
                          "config": {"workload": workload, "sample": f"the config's batch, {args.batch} x {vol_s} per step, computed in fp32 on the host cores",
                                     "timing": "time.perf_counter"},
                          "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": cores, "kind": "port",
                                           "sample": f"oracle/graphs.py restatement, {len(times)} timed steps of {args.batch} x {vol_s} after {wu} warm-up"},
                          "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    if args.impl == "torch_gpu":
        if rank != 0:
            return
        assert torch.cuda.is_available()
        torch.cuda.set_device(local)
        tg = torch_gpu_baseline(args.model, args.batch, vol, args.steps, args.warmup, args.norm, torch.device("cuda", local))
        print(json.dumps({**base, "impl": "torch_gpu", "value": tg["value"], "ms_per_step": tg["ms_per_step"], "dtype": "bf16",
                          "config": {"workload": workload}, "torch_gpu_baseline": tg}), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import __graft_entry__
    pkg = __graft_entry__.build()
    if args.config in (1, 4):
        return other_config(pkg, args, rank, world, local, base)
    from mri_epilepsy_diagnosis_b200 import functional as BF
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net, model_desc = build_model(pkg, args.model, args.norm, args.literal_graph)
    sync = (None, world) if (args.sync_bn and world > 1) else None
    net = pkg.convert(net.to(dev).train(), dtype=torch.bfloat16, sync=sync)
    opt = torch.optim.AdamW(net.parameters(), capturable=not args.eager, fused=True)
    xh, th = synthetic_batch(args.batch, vol, seed=rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.to(dev), th.to(dev)
    voxels = xh.numel()
    if world > 1:
        pkg.dp.broadcast_parameters(net)
    # softmax -> get_dice_loss -> .mean() (segmentation/routine.py:272-274) as the library's fused loss kernel
    loss_fn = dice_loss_mean if args.torch_loss else BF.softmax_dice_loss
    if args.eager:
        bucket = pkg.dp.attach(net, opt) if world > 1 else None

        def step(x, t):
            opt.zero_grad()
            loss = loss_fn(net(x), t)
            loss.backward()
            opt.step()
            return loss
    else:
        # the loop body of segmentation/routine.py:266-281 captured once in a CUDA graph and replayed (graphed.py)
        step = pkg.graphed.GraphedTrainStep(net, loss_fn, opt, xd, td)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / n
        if dist is not None:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms, t0, time.time()

    launches_a = pkg.launch_count()
    if not args.eager:       # kernels of one step = kernels enqueued by one eager execution of the same body
        step._body(eager=True)
    launches_per_step = pkg.launch_count() - launches_a
    for _ in range(max(3, args.warmup)):
        step(xd, td)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    launches0 = pkg.launch_count()
    ms, t0, t1 = timed(lambda: step(xd, td), args.steps)
    launches = (pkg.launch_count() - launches0) if args.eager else launches_per_step * args.steps

    # end to end: host (pinned) inputs in, host scalar out, every step.  Graph mode: the H2D copy of step i+1's batch runs on a
    # copy stream while step i computes (graphed.prefetch / step_prefetched) -- K copies for K steps inside the timed region.
    def e2e_step():
        x = xh.to(dev, non_blocking=True)
        t = th.to(dev, non_blocking=True)
        return float(step(x, t))

    def e2e_run(n):
        if args.eager:
            for _ in range(n):
                e2e_step()
            return
        step.prefetch(xh, th)
        for i in range(n):
            if i + 1 < n:
                step.prefetch(xh, th)
            float(step.step_prefetched())
    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    # clocks: every nvidia-smi sample (100 ms period) that fell inside the two timed regions (device-resident and end-to-end)
    clocks = sampler.summary(t0, time.time()) if sampler else None
    if dist is not None:
        tms = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms)

    # roofline leg: per-launch CUDA events around every conv kernel over another K (eager) steps
    eager_step = step if args.eager else (lambda x, t: step._body(eager=True))
    BF.PROFILE = []
    torch.cuda.synchronize()
    for _ in range(args.steps):
        eager_step(xd, td)
    torch.cuda.synchronize()
    rows, BF.PROFILE = BF.PROFILE, None
    agg = {}
    for which, algo, flops, shape, a, b in rows:
        k = (which, algo, shape)
        r = agg.setdefault(k, [0.0, 0.0, 0])
        r[0] += flops; r[1] += a.elapsed_time(b) * 1e-3; r[2] += 1
    # kernel behind every (pass, algo): the tcgen05 kernels are the dense contractions of the path
    KERNEL = {(0, 2): "row_fwd_kernel", (1, 2): "row_fwd_kernel", (2, 2): "row_wgrad_kernel", (0, 1): "conv_umma_kernel", (1, 1): "conv_umma_kernel",
              (2, 1): "conv_wgrad_umma_kernel"}
    if args.layer_table and rank == 0:          # per-layer evidence: every (pass, kernel, shape) of the step with its time and FLOP rate
        names = {0: "fwd", 1: "dgrad", 2: "wgrad"}
        with open(args.layer_table, "w") as f:
            f.write("| pass | kernel | Ci, Co, D, H, W, k | launches / step | us / launch | TFLOP/s |\n|---|---|---|---:|---:|---:|\n")
            for (which, algo, shape), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| {names[which]} | {KERNEL.get((which, algo), 'simt / special')} | {shape} | {v[2] / args.steps:g} | "
                        f"{1e6 * v[1] / v[2]:.1f} | {v[0] / v[1] / 1e12:.0f} |\n")
    tc = {k: v for k, v in agg.items() if k[1] != 0}
    tc_flops, tc_s = sum(v[0] for v in tc.values()), sum(v[1] for v in tc.values())
    conv_s = sum(v[1] for v in agg.values())
    pk = peaks()
    # dominant kernel = the tcgen05 kernel with the largest share of the step; `achieved` = algorithmic FLOPs of ALL its launches /
    # the sum of their durations (= FLOPs per launch / average launch duration).  Its heaviest layer shape is reported as detail
    # (fwd and dgrad of a "same" convolution with Ci == Co are the same kernel on the same geometry).
    per_kernel, sig = {}, {}
    for (which, algo, shape), v in tc.items():
        kn = KERNEL[(which, algo)]
        r = per_kernel.setdefault(kn, [0.0, 0.0, 0])
        r[0] += v[0]; r[1] += v[1]; r[2] += v[2]
        key = (kn, ((min(shape[0], shape[1]), max(shape[0], shape[1])) + tuple(shape[2:])) if which != 2 else tuple(shape))
        r = sig.setdefault(key, [0.0, 0.0, 0])
        r[0] += v[0]; r[1] += v[1]; r[2] += v[2]
    dom_kernel, dom = max(per_kernel.items(), key=lambda kv: kv[1][1]) if per_kernel else ("none", [0.0, 1e-9, 0])
    top = max(((k, v) for k, v in sig.items() if k[0] == dom_kernel), key=lambda kv: kv[1][1], default=None)
    achieved = dom[0] / dom[1] / 1e12
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass
    top_launch = None
    if top is not None:
        (_, ts), tv = top
        top_launch = {"layer": f"Conv3d {ts[0]}<->{ts[1]} k{ts[5]} on {args.batch} x {ts[2]}x{ts[3]}x{ts[4]} bf16", "launches_per_step": tv[2] / max(1, args.steps),
                      "ms_per_launch": 1e3 * tv[1] / max(1, tv[2]), "achieved": tv[0] / tv[1] / 1e12, "frac": tv[0] / tv[1] / 1e12 / pk["bf16_tflops"],
                      "algorithmic_bytes": 2.0 * args.batch * ts[2] * ts[3] * ts[4] * (ts[0] + ts[1]),
                      "traffic": traffic_tab.get(f"{dom_kernel}|" + "|".join(str(int(x)) for x in ts) + f"|n{args.batch}")}
    roofline = {"bound": "tensor", "kernel": dom_kernel, "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"],
                "traffic": traffic_tab.get(dom_kernel) if (args.model == "unet3d" and args.batch == 4 and vol == (128, 128, 128)) else None,
                "traffic_unit": "bytes per launch, dram read + write averaged over the kernel's launches of one step (ncu; profiles/traffic.json)",
                "peak_source": pk["source"] + " (sustained bf16)",
                "launches_per_step": dom[2] / max(1, args.steps), "ms_per_launch_avg": 1e3 * dom[1] / max(1, dom[2]),
                "share_of_step": (dom[1] / args.steps) / (ms / 1e3), "top_launch": top_launch,
                "all_tcgen05_convs": {"achieved": tc_flops / tc_s / 1e12 if tc_s > 0 else 0.0, "frac": (tc_flops / tc_s / 1e12 / pk["bf16_tflops"]) if tc_s > 0 else 0.0,
                                      "share_of_step": (tc_s / args.steps) / (ms / 1e3), "launches_per_step": sum(v[2] for v in tc.values()) / max(1, args.steps)},
                "all_conv_share_of_step": (conv_s / args.steps) / (ms / 1e3),
                "note": "achieved = algorithmic FLOPs (2*N*Do*Ho*Wo*Co*Ci*taps per launch) of the dominant kernel's launches / their CUDA-event "
                        "durations; events on the launching stream around each launch (eager replays of the same step)"}
    if args.profile_json and rank == 0:
        names = {0: "fwd", 1: "dgrad", 2: "wgrad"}
        table = [{"pass": names[k[0]], "algo": {0: "direct", 1: "umma", 2: "row"}[k[1]], "Ci": k[2][0], "Co": k[2][1], "out": list(k[2][2:5]), "kd": k[2][5],
                  "calls": v[2], "ms_per_call": 1e3 * v[1] / v[2], "tflops": v[0] / v[1] / 1e12} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
        with open(args.profile_json, "w") as f:
            json.dump({"ms_per_step": ms, "rows": table}, f, indent=1)

    # transparency leg (single GPU, default workload): the same step with unet3d.py's operator sequence executed one to one
    # (zoo.Unet(literal=True): dead branch materialised, conv2 on the upsampled tensor) -- see DESIGN.md section 4
    literal = None
    if world == 1 and args.model == "unet3d" and not args.literal_graph and not args.eager and not args.no_literal_leg:
        step = None
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        net2, _ = build_model(pkg, args.model, args.norm, True)
        net2 = pkg.convert(net2.to(dev).train(), dtype=torch.bfloat16)
        opt2 = torch.optim.AdamW(net2.parameters(), capturable=True, fused=True)
        step2 = pkg.graphed.GraphedTrainStep(net2, loss_fn, opt2, xd, td)
        for _ in range(3):
            step2(xd, td)
        ms2, _, _ = timed(lambda: step2(xd, td), args.steps)
        literal = {"ms_per_step": ms2, "value": voxels / (ms2 / 1e3), "unit": "voxels/s",
                   "what": "same step, reference operator sequence executed one to one (no graph-level rewrites)"}
        del step2, net2, opt2
    def shutdown():
        """CUDA graphs that hold captured NCCL kernels must be gone before the communicator is torn down (destroy_process_group
        otherwise never returns); a watchdog ends the process if the teardown stalls anyway -- every result is printed by then."""
        if dist is None:
            return
        import gc
        dist.barrier()
        torch.cuda.synchronize()
        gc.collect()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)
    if rank != 0:
        step = eager_step = net = opt = None
        shutdown()
        return
    out = {**base, "value": world * voxels / (ms / 1e3), "ms_per_step": ms, "dtype": "bf16",
           "config": {"workload": workload, "network": model_desc, "volumes_per_step": args.batch * world, "volume": list(vol),
                      "optimizer": "torch.optim.AdamW(fused=True)" + ("" if args.eager else " inside the captured step"),
                      "parallelism": f"dp{world}" + ("+syncbn" if sync else ""),
                      "launch": "eager" if args.eager else "cuda-graph replay of the step" + (" (bucketed NCCL all-reduce captured inside, overlapping the backward pass)" if world > 1 else ""),
                      "l2": "inputs and every activation tensor (>= 268 MB each at 16ch x 128^3 x 4) exceed the 126 MB L2; no flush needed"},
           "e2e": {"value": world * voxels / (ms_e2e / 1e3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": xh.numel() * 4 + th.numel() * 4, "d2h_bytes_per_step": 4},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "literal_graph": literal,
           "model_tflops": FWD_BWD_FLOP_PER_VOXEL * voxels / (ms / 1e3) / 1e12 if args.model == "unet3d" else None}
    if args.gpus == 1 and not args.no_torch_baseline:
        # the bar SURVEY section 2 sets for every kernel: the PyTorch/cuDNN path the unmodified reference modules hit on the same B200
        step = eager_step = net = opt = None           # release this arm's graph, activations and optimizer state first
        torch.cuda.empty_cache()
        tg = torch_gpu_baseline(args.model, args.batch, vol, args.steps, args.warmup, args.norm, dev)
        if tg.get("value"):
            tg["ours_over_torch_gpu"] = out["value"] / tg["value"]
        out["torch_gpu_baseline"] = tg
    if args.gpus == 1 and args.model == "unet3d" and not args.no_dropin_leg:
        # the drop-in path as a reference user runs it: a STOCK torch.nn model with unet3d.py's graph (tests/refshaped.StockUnet3d:
        # F.upsample, F.dropout3d, torch.cat, in-place ReLU) -> nn.convert() -> the eager loop body of segmentation/routine.py:266-281
        # (zero_grad, forward, torch softmax + Dice, backward, optimizer.step) with dp.attach hooked on; no graph capture, no zoo rewrite
        step = eager_step = net = opt = None
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import refshaped
        torch.manual_seed(0)
        stock = refshaped.StockUnet3d(c=1, n=16, dropout=0.5, norm=args.norm, num_classes=2)
        dnet = pkg.convert(stock.to(dev).train(), dtype=torch.bfloat16)
        dopt = torch.optim.AdamW(dnet.parameters())
        dbucket = pkg.dp.attach(dnet, dopt)

        def dropin_step():
            dopt.zero_grad()
            loss = dice_loss_mean(dnet(xd), td)
            loss.backward()
            dopt.step()
            return loss
        for _ in range(3):
            dropin_step()
        msd, _, _ = timed(dropin_step, args.steps)
        dbucket.remove()
        out["dropin_eager"] = {"ms_per_step": msd, "value": voxels / (msd / 1e3), "unit": "voxels/s",
                               "what": "nn.convert() of a stock torch.nn unet3d-shaped model, eager routine.py loop body (torch softmax + Dice, torch AdamW), "
                                       "no CUDA graph and no graph-level rewrites: the reference's op sequence incl. the dead branch"}
        # the same converted stock model, loop body captured: ONE extra line for the user (graphed.GraphedTrainStep around the model,
        # the loss and a capturable optimizer); still no zoo rewrite -- the stock graph incl. the dead branch and F.dropout3d is replayed
        try:
            dbucket = None
            torch.manual_seed(0)
            dopt2 = torch.optim.AdamW(dnet.parameters(), capturable=True, fused=True)
            gstep = pkg.graphed.GraphedTrainStep(dnet, dice_loss_mean, dopt2, xd, td)
            for _ in range(3):
                gstep(xd, td)
            msg, _, _ = timed(lambda: gstep(xd, td), args.steps)
            out["dropin_graphed"] = {"ms_per_step": msg, "value": voxels / (msg / 1e3), "unit": "voxels/s",
                                     "what": "the same converted stock model inside graphed.GraphedTrainStep (CUDA-graph replay of the eager loop body)"}
        except Exception as e:      # a stock op that cannot be captured is reported, not hidden
            out["dropin_graphed"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        gstep = dopt2 = None
        dnet = dopt = stock = None
        torch.cuda.empty_cache()
    if args.gpus == 1 and not args.no_other_configs and args.model == "unet3d" and args.config == 2:
        # every other BASELINE configuration + the 'next' rows, timed in this same process (tools/workloads.py)
        from tools import workloads
        step = eager_step = net = opt = None
        torch.cuda.empty_cache()
        out["other_configs"] = workloads.run_all(pkg, dev, pk["hbm_gbs"])
    if args.gpus == 1 and not args.no_cpu_baseline:
        nvox, times, _ = cpu_reference_step(args.model, vol, 2, 1, cores, args.norm)
        v = nvox / (sum(times) / len(times))
        out["cpu_baseline"] = {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                               "sample": f"oracle/graphs.py (CPU restatement of the reference step), fp32, 2 timed steps of 1 x {vol_s} after 1 warm-up"}
    print(json.dumps(out), flush=True)
    step = eager_step = net = opt = None
    shutdown()


if __name__ == "__main__":
    main()
