#!/usr/bin/env python
"""Micro-timing of the HBM-bound operators at the bench shapes (CUDA events, inputs larger than L2).
   python tools/op_bench.py [op ...]     ops: upsample norm pool stem head cat"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
pkg = __graft_entry__.build()
from mri_epilepsy_diagnosis_b200 import functional as BF, _cabi as cabi, nn as bnn
dev = "cuda"
HBM = 6544.7

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def report(name, ms, nbytes):
    print(f"{name:52s} {ms*1e3:8.1f} us   {nbytes/ms/1e6:7.0f} GB/s  ({100*nbytes/ms/1e6/HBM:4.1f}% of HBM peak)", flush=True)

def cl(n, c, s, dtype=torch.bfloat16):
    return torch.randn(n, c, s, s, s, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)

ops = sys.argv[1:] or ["upsample", "norm", "pool", "stem", "head", "cat"]
if "upsample" in ops:
    for (c, s, dt) in ((16, 64, torch.bfloat16), (32, 64, torch.bfloat16), (32, 32, torch.bfloat16), (64, 32, torch.bfloat16), (2, 64, torch.float32)):
        x = cl(4, c, s, dt).requires_grad_(True)
        y = BF.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        eb = x.element_size()
        report(f"upsample x2 fwd C={c} {s}^3->{2*s}^3 {dt}", timeit(lambda: BF.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)), x.numel()*eb*9)
        g = torch.randn_like(y)
        report(f"upsample x2 bwd C={c}", timeit(lambda: torch.autograd.grad(y, x, g, retain_graph=True)), x.numel()*eb*9)
if "norm" in ops:
    for (c, s) in ((16, 128), (32, 128), (64, 64)):
        x = cl(4, c, s).requires_grad_(True)
        bn = bnn.BatchNorm3d(c).cuda().train()
        y = bn(x, act=cabi.ACT_RELU)
        nb = x.numel() * 2
        report(f"bn fwd (stats+apply+relu) C={c} {s}^3", timeit(lambda: bn(x, act=cabi.ACT_RELU)), nb*3)
        g = torch.randn_like(y)
        report(f"bn bwd (partial+apply) C={c} {s}^3", timeit(lambda: torch.autograd.grad(y, x, g, retain_graph=True)), nb*7)
        r = cl(4, c, s)
        y2 = bn(x, act=cabi.ACT_RELU, residual=r)
        report(f"bn fwd +residual C={c} {s}^3", timeit(lambda: bn(x, act=cabi.ACT_RELU, residual=r)), nb*4)
if "pool" in ops:
    x = cl(4, 16, 128).requires_grad_(True)
    y = BF.max_pool(x, 2, 2)
    report("maxpool fwd C=16 128^3", timeit(lambda: BF.max_pool(x, 2, 2)), x.numel()*2*1.25)
    g = torch.randn_like(y)
    report("maxpool bwd C=16 128^3", timeit(lambda: torch.autograd.grad(y, x, g, retain_graph=True)), x.numel()*2*1.25)
if "stem" in ops:
    x = torch.randn(4, 1, 128, 128, 128, device=dev)
    m = bnn.Conv3d(1, 16, 3, 1, 1, bias=False).cuda(); m.compute_dtype = torch.bfloat16
    y = m(x)
    report("stem fwd 1->16 128^3 (x fp32)", timeit(lambda: m(x)), x.numel()*4 + y.numel()*2)
    g = torch.randn_like(y)
    report("stem wgrad 1->16 128^3", timeit(lambda: torch.autograd.grad(y, m.weight, g, retain_graph=True)), x.numel()*4 + y.numel()*2)
if "head" in ops:
    x = cl(4, 32, 128).requires_grad_(True)
    m = bnn.Conv3d(32, 2, 1).cuda(); m.compute_dtype = torch.bfloat16; m.out_dtype = torch.float32
    y = m(x)
    report("head fwd 32->2 128^3", timeit(lambda: m(x)), x.numel()*2 + y.numel()*4)
    g = torch.randn_like(y)
    report("head dgrad+wgrad 32->2 128^3", timeit(lambda: torch.autograd.grad(y, (x, m.weight), g, retain_graph=True)), 3*x.numel()*2 + 2*y.numel()*4)
if "cat" in ops:
    a, b = cl(4, 16, 128), cl(4, 16, 128)
    report("torch.cat 16+16 128^3", timeit(lambda: torch.cat([a, b], 1)), a.numel()*2*4)
    report("torch add bf16 16ch 128^3", timeit(lambda: a + b), a.numel()*2*3)
    report("copy_ (clone) 16ch 128^3", timeit(lambda: a.clone()), a.numel()*2*2)
