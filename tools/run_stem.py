"""Forward + backward of the 1 -> 16 stem on 4 x 128^3: timing (CUDA events) and an ncu target for the stem kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__

pkg = __graft_entry__.build()
dev = torch.device("cuda", 0)
conv = pkg.nn.Conv3d(1, 16, 3, 1, 1, bias=False).to(dev)
conv.compute_dtype = torch.bfloat16
x = torch.randn(4, 1, 128, 128, 128, device=dev)
y = conv(x)
gy = torch.ones_like(y)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def bwd():
    conv.weight.grad = None
    y.backward(gy, retain_graph=True)


print(f"stem fwd {timed(lambda: conv(x)):.1f} us, wgrad (+ autograd overhead) {timed(bwd):.1f} us, B200_STEM_MMA={os.environ.get('B200_STEM_MMA', '1')}")
