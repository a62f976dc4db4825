"""Micro-benchmark of the BatchNorm3d + ReLU forward / backward kernels on the headline layer shapes (bf16, channels-last).
Prints ms and the HBM GB/s of the algorithmic traffic (forward: stats read + apply read/write = 3 units; backward: 2 + 3 units)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import __graft_entry__
    pkg = __graft_entry__.build()
    from mri_epilepsy_diagnosis_b200 import _cabi as cabi
    dev = torch.device("cuda", 0)
    for C, S in ((16, 128), (32, 128), (64, 64)):
        bn = pkg.nn.BatchNorm3d(C).to(dev).train()
        x = torch.randn(4, C, S, S, S, device=dev).bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        gy = torch.randn_like(x)
        unit = x.numel() * 2

        def fwd():
            return bn(x, act=cabi.ACT_RELU)
        y = fwd()

        def bwd():
            x.grad = None
            y.backward(gy, retain_graph=True)
        for name, fn, units in (("fwd", fwd, 3), ("bwd", bwd, 5)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"U={os.environ.get('B200_NORM_U', 'default')} C={C} S={S} {name}: {ms:.3f} ms, {units * unit / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
