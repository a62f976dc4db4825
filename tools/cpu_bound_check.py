#!/usr/bin/env python
"""Is the bench step GPU-bound or launch-bound?  Compares host enqueue time per step with the device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__, bench
pkg = __graft_entry__.build()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net, _ = bench.build_model(pkg, "unet3d")
net = pkg.convert(net.to(dev).train(), dtype=torch.bfloat16)
opt = torch.optim.AdamW(net.parameters())
x, t = bench.synthetic_batch(4, 128, 0)
x, t = x.to(dev), t.to(dev)
def step():
    opt.zero_grad()
    loss = bench.dice_loss_mean(net(x), t)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): step()
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/n:.2f} ms/step, device {e0.elapsed_time(e1)/n:.2f} ms/step, wall {1e3*(t2-t0)/n:.2f} ms/step, launches/step {pkg.launch_count()/13:.0f}")
