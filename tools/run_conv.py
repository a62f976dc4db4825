#!/usr/bin/env python
"""Run one convolution shape through the drop-in module (for ncu captures and micro-timing).
   python tools/run_conv.py --ci 16 --co 16 --size 128 --n 4 --iters 5 [--bwd]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__

ap = argparse.ArgumentParser()
ap.add_argument("--ci", type=int, default=16)
ap.add_argument("--co", type=int, default=16)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
pkg = __graft_entry__.build()
torch.manual_seed(0)
m = pkg.nn.Conv3d(a.ci, a.co, a.k, 1, a.k // 2, bias=False).cuda()
m.compute_dtype = torch.bfloat16
x = torch.randn(a.n, a.ci, a.size, a.size, a.size, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(a.bwd)
for _ in range(2):
    y = m(x)
    if a.bwd:
        y.backward(y.detach())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    y = m(x)
    if a.bwd:
        y.backward(y.detach())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
flops = 2.0 * a.n * a.size ** 3 * a.ci * a.co * a.k ** 3 * (3 if a.bwd else 1)
print(f"conv {a.ci}->{a.co} k{a.k} {a.n}x{a.size}^3 {'fwd+bwd' if a.bwd else 'fwd'}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
