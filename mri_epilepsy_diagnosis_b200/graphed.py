"""Whole-step CUDA-graph capture for the reference's training-loop body (segmentation/routine.py:266-281:
zero_grad -> model(x) -> loss -> backward -> optimizer.step).

The 3-D U-Net step issues ~450 kernels of 10-700 us each; at batch 4 x 128^3 the host needs about as long to enqueue them
through autograd as the B200 needs to run them, so the step is launch-bound.  Capturing the step once and replaying it
removes the host from the loop.  Every library call only enqueues kernels on the current stream and takes caller-owned
buffers, so the C ABI is capture-safe as it stands (tensor maps are built on the host and passed by value).

Single GPU: the graph holds zero_grad + forward + loss + backward + optimizer.step (the optimizer must be constructed with
`capturable=True`).  Data parallel: the graph holds zero_grad + forward + loss + backward; gradients accumulate directly into
one flat fp32 buffer (each `p.grad` is a view of it), which is all-reduced over NCCL after the replay, followed by the
optimizer step (eager or its own graph) -- one collective per step, no per-parameter copies.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import functional as BF
from . import nn as bnn


class GraphedTrainStep:
    """step = GraphedTrainStep(model, loss_fn, optimizer, x_example, t_example); loss = step(x, t)

    `x`/`t` may live on the host (pinned) or the device; they are copied into the static input buffers of the graph.
    Returns the (static) loss tensor of the step; read it with `.item()` / `float()` when needed.

    Construction runs a probe step and `warmup` eager steps on `x_example` (CUDA-graph capture needs warmed-up allocations
    and lazily created optimizer state).  Those steps are UNDONE before the capture: parameters, buffers (BatchNorm running
    statistics, `num_batches_tracked`), optimizer state and the CPU/CUDA RNG streams are restored in place, so that N calls
    perform exactly the N optimizer steps of the reference loop.  Optimizer state that did not exist before construction is
    zeroed rather than deleted (the captured `optimizer.step()` needs the tensors): identical to a fresh Adam/AdamW/SGD(momentum,
    dampening=0) state; an optimizer whose fresh state is not all zeros must be stepped once by the caller before construction.
    """

    def __init__(self, model, loss_fn, optimizer, x_example, t_example, process_group=None, warmup=3):
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = next(model.parameters()).device
        self.x = torch.empty(x_example.shape, dtype=x_example.dtype, device=dev)
        self.t = torch.empty(t_example.shape, dtype=t_example.dtype, device=dev)
        self.x.copy_(x_example)
        self.t.copy_(t_example)
        snap = self._snapshot(dev)
        # Probe step (eager): parameters that receive no gradient (unet3d's dead conv2/bn2 branch, unet3d.py:43-46; frozen
        # fader sub-networks) must keep grad=None so that the optimizer skips them exactly like in the reference's loop.
        optimizer.zero_grad(set_to_none=True)
        self.loss_fn(model(self.x), self.t).backward()
        self.params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
        optimizer.zero_grad(set_to_none=True)
        # gradients live in one flat buffer: zeroed by one memset inside the graph, all-reduced by one collective
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.capture_opt = self.world == 1
        if self.capture_opt:
            for g in optimizer.param_groups:
                if "capturable" in g and not g["capturable"]:
                    raise RuntimeError("GraphedTrainStep: construct the optimizer with capturable=True so optimizer.step() can be captured")

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body(eager=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap, dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body(eager=False)
        self.opt_graph = None

    def _snapshot(self, dev):
        state = {}
        for p, st in self.opt.state.items():
            state[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        return {"tensors": [(t, t.detach().clone()) for t in list(self.model.parameters()) + list(self.model.buffers())],
                "opt": state, "cpu_rng": torch.get_rng_state(), "cuda_rng": torch.cuda.get_rng_state(dev)}

    def _restore(self, snap, dev):
        with torch.no_grad():
            for t, saved in snap["tensors"]:
                t.copy_(saved)
            for p, st in self.opt.state.items():
                before = snap["opt"].get(p)
                for k, v in st.items():
                    if not torch.is_tensor(v):
                        if before is not None and k in before:
                            st[k] = before[k]
                    elif before is not None and torch.is_tensor(before.get(k)):
                        v.copy_(before[k])
                    else:
                        v.zero_()
        torch.set_rng_state(snap["cpu_rng"])
        torch.cuda.set_rng_state(snap["cuda_rng"], dev)

    def _fwd_bwd(self):
        self.flat.zero_()
        with bnn.defer_batch_counters():
            loss = self.loss_fn(self.model(self.x), self.t)
        with BF.deferred_wgrad():          # wgrad kernels accumulate into the flat buffer on a side stream; one join here
            loss.backward()
        return loss.detach()

    def _reduce_and_step(self):
        if self.world > 1:
            dist.all_reduce(self.flat, group=self.pg)
            self.flat.div_(self.world)
        self.opt.step()

    def _body(self, eager):
        loss = self._fwd_bwd()
        if eager or self.capture_opt:
            if eager:
                self._reduce_and_step()
            else:
                self.opt.step()
        return loss

    def __call__(self, x, t):
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        if not self.capture_opt:
            self._reduce_and_step()
        return self.loss

    # ---- host-fed steps with the next batch's H2D copy overlapped with the current step (what a pin_memory DataLoader with
    # non_blocking copies gives the reference loop): `prefetch(x, t)` starts the copy of a PINNED host batch into a staging
    # buffer on a copy stream; `step_prefetched()` waits for it, moves it into the graph's static inputs (device-to-device)
    # and replays.  Every step still performs exactly one H2D copy of its own inputs.
    def prefetch(self, x_host, t_host):
        if not hasattr(self, "_stage"):
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.t)) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.x.device)
            self._ready = [None, None]
            self._consumed = [None, None]
            self._slot = 0
            self._pending = []
        s = self._slot
        self._slot ^= 1
        if self._consumed[s] is not None:                 # the step that read this staging slot must have copied it out
            self._copy_stream.wait_event(self._consumed[s])
        with torch.cuda.stream(self._copy_stream):
            self._stage[s][0].copy_(x_host, non_blocking=True)
            self._stage[s][1].copy_(t_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._ready[s] = ev
        self._pending.append(s)

    def step_prefetched(self):
        s = self._pending.pop(0)
        cur = torch.cuda.current_stream(self.x.device)
        cur.wait_event(self._ready[s])
        self.x.copy_(self._stage[s][0], non_blocking=True)
        self.t.copy_(self._stage[s][1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._consumed[s] = ev
        self.graph.replay()
        if not self.capture_opt:
            self._reduce_and_step()
        return self.loss
