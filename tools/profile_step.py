#!/usr/bin/env python
"""One train step of the bench workload between cudaProfilerStart/Stop, for ncu launch lists:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
       python tools/profile_step.py [--model unet3d] [--batch 4] [--size 128]
Run it WITHOUT ncu first (it must exit 0); it then also prints a torch-side per-kernel table (CUDA events are not used
here: this script is for launch lists, never for bench values)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import __graft_entry__
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="unet3d")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
pkg = __graft_entry__.build()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net, desc = bench.build_model(pkg, a.model)
net = pkg.convert(net.to(dev).train(), dtype=torch.bfloat16)
opt = torch.optim.AdamW(net.parameters(), fused=True)
x, t = bench.synthetic_batch(a.batch, a.size, 0)
x, t = x.to(dev), t.to(dev)


def step():
    opt.zero_grad()
    with pkg.nn.defer_batch_counters():
        loss = pkg.functional.softmax_dice_loss(net(x), t)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", a.steps, "step(s) of", desc)
