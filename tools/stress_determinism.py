#!/usr/bin/env python
"""Repeat conv forward/backward on the tcgen05 test shapes and compare every run bit for bit with the first one
(all kernels reduce in a fixed order, so any difference is a race).   python tools/stress_determinism.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__
__graft_entry__.build()
import mri_epilepsy_diagnosis_b200 as B

CASES = [(1, 128, 64, (4, 16, 8)), (2, 64, 128, (5, 8, 8)), (1, 256, 128, (4, 8, 16)), (1, 256, 256, (3, 16, 16)), (2, 128, 256, (8, 8, 8)),
         (2, 128, 128, (16, 16, 16)), (1, 96, 32, (5, 16, 16)), (2, 16, 16, (10, 12, 16)), (1, 64, 64, (6, 17, 9)), (2, 32, 32, (9, 20, 11)),
         (1, 32, 32, (4, 9, 128)), (2, 16, 16, (6, 10, 128)), (1, 16, 32, (4, 9, 128))]          # W = 128: the kx-folded mode of row_fwd_kernel
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
bad = 0
for (N, Ci, Co, size) in CASES:
    g = torch.Generator().manual_seed(Ci)
    mod = B.nn.Conv3d(Ci, Co, 3, 1, 1, bias=False).cuda()
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(N, Ci, *size, generator=g).cuda().bfloat16()
    gy = torch.randn(N, Co, *size, generator=g).cuda().bfloat16()
    ref = None
    for it in range(iters):
        xg = x.clone().requires_grad_(True)
        mod.weight.grad = None
        y = mod(xg)
        y.backward(gy)
        torch.cuda.synchronize()
        cur = (y.detach().clone(), xg.grad.clone(), mod.weight.grad.clone())
        if ref is None:
            ref = cur
        else:
            for name, a, b in zip(("fwd", "dgrad", "wgrad"), ref, cur):
                if not torch.equal(a, b):
                    bad += 1
                    d = (a.float() - b.float()).abs()
                    print(f"MISMATCH case={(N, Ci, Co, size)} iter={it} {name}: {int((d > 0).sum())} elements, max {float(d.max()):.3e}", flush=True)
    print("case", (N, Ci, Co, size), "done", flush=True)
print("mismatches:", bad)
