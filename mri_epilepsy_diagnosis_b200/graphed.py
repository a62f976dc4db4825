"""Whole-step CUDA-graph capture for the reference's training-loop body (segmentation/routine.py:266-281:
zero_grad -> model(x) -> loss -> backward -> optimizer.step).

The 3-D U-Net step issues ~450 kernels of 10-700 us each; at batch 4 x 128^3 the host needs about as long to enqueue them
through autograd as the B200 needs to run them, so the step is launch-bound.  Capturing the step once and replaying it
removes the host from the loop.  Every library call only enqueues kernels on the current stream and takes caller-owned
buffers, so the C ABI is capture-safe as it stands (tensor maps are built on the host and passed by value).

Single GPU: the graph holds zero_grad + forward + loss + backward + optimizer.step (the optimizer must be constructed with
`capturable=True`).  Data parallel: gradients accumulate directly into one flat fp32 buffer (each `p.grad` is a view of it),
cut into a few contiguous buckets; the NCCL all-reduce (AVG) of a bucket is captured INSIDE the graph on a communication
stream, forked as soon as the backward pass has produced the bucket's last gradient (late layers first), so it overlaps the
rest of the backward pass; the optimizer step is captured too when the optimizer is capturable.  No per-parameter copies.
"""
from __future__ import annotations

import contextlib
import os

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

    def __init__(self, model, loss_fn, optimizer, x_example, t_example, process_group=None, warmup=3, buckets=3):
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self._bucketing = False
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
        capturable = all(g.get("capturable", False) for g in optimizer.param_groups)
        self.capture_opt = self.world == 1 or capturable
        if self.world == 1 and not capturable:
            raise RuntimeError("GraphedTrainStep: construct the optimizer with capturable=True so optimizer.step() can be captured")
        if self.world > 1:
            self._make_buckets(max(1, int(buckets)), dev)

        self.pack_plan = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with BF.PackPlan.recording() as packs:       # which (layer, pass) weight copies one step derives
                self._body(eager=True)
            for _ in range(max(1, warmup) - 1):
                self._body(eager=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap, dev)
        if os.environ.get("B200_PACK_BATCHED", "1") != "0":
            # every packed weight copy of the step from ONE launch at its start instead of one launch per layer and pass
            plan = BF.PackPlan(packs)
            torch.cuda.synchronize(dev)
            self.pack_plan = plan if plan.table is not None else None
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

    # ---- gradient buckets (data parallel): contiguous slices of the flat buffer in parameter order; the backward pass finishes them
    # from the last one to the first
    def _make_buckets(self, n, dev):
        total = self.flat.numel()
        per = -(-total // n)
        self._bucket_of, self._bucket_range, self._members = {}, [], []
        off = lo = 0
        cur = []
        for p in self.params:
            cur.append(p)
            off += p.numel()
            if off - lo >= per:
                self._bucket_range.append((lo, off)); self._members.append(cur)
                lo, cur = off, []
        if cur:
            self._bucket_range.append((lo, off)); self._members.append(cur)
        for b, ps in enumerate(self._members):
            for p in ps:
                self._bucket_of[p] = b
        self.comm = torch.cuda.Stream(device=dev)
        # one notification per parameter and backward pass: autograd runs the post-accumulate-grad hook of a leaf once all of its
        # gradient contributions have been produced -- also when they were None because the deferred-wgrad path accumulated them
        # itself on the side stream (the hook then fires after that accumulation has been enqueued)
        self._hooks = [p.register_post_accumulate_grad_hook(self._notify) for p in self.params]

    def _notify(self, p):
        if not self._bucketing:
            return
        b = self._bucket_of.get(p)
        if b is None or self._launched[b]:
            return
        self._remaining[b] -= 1
        # bucket 0 holds the first layers: the backward pass finishes them last, nothing is left to overlap with, and it is
        # reduced after the side stream has been joined (see _body)
        if self._remaining[b] <= 0 and b > 0:
            self._launch_bucket(b)

    def _launch_bucket(self, b):
        dev = self.flat.device
        cur = torch.cuda.current_stream(dev)
        self.comm.wait_stream(cur)
        self.comm.wait_stream(BF._side_stream(dev))            # deferred weight gradients are accumulated on the side stream
        lo, hi = self._bucket_range[b]
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG, group=self.pg)
        self._launched[b] = True

    def _fwd_bwd(self, bucketing=False):
        self.flat.zero_()
        plan = self.pack_plan
        if plan is not None:
            plan.run()
        with (plan.serving() if plan is not None else contextlib.nullcontext()):
            with bnn.defer_batch_counters():
                loss = self.loss_fn(self.model(self.x), self.t)
            if bucketing:
                self._remaining = [len(ps) for ps in self._members]
                self._launched = [False] * len(self._members)
                self._bucketing = True
            try:
                with BF.deferred_wgrad():     # wgrad kernels accumulate into the flat buffer on a side stream; one join here
                    loss.backward()
            finally:
                self._bucketing = False
        return loss.detach()

    def _reduce_and_step(self):
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.pg)
        self.opt.step()

    def _body(self, eager):
        if eager or self.world == 1:
            loss = self._fwd_bwd()
            if eager:
                self._reduce_and_step()
            else:
                self.opt.step()
            return loss
        # data parallel, captured: bucketed all-reduce overlapped with the backward pass, then (capturable optimizers) the step
        loss = self._fwd_bwd(bucketing=True)
        for b in reversed(range(len(self._members))):
            if not self._launched[b]:
                self._launch_bucket(b)
        torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)
        if self.capture_opt:
            self.opt.step()
        return loss

    def __call__(self, x, t):
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        BF.invalidate_packed_weights()            # the replay changed the parameters without any Python running (inference caches)
        if not self.capture_opt:
            self.opt.step()
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
        BF.invalidate_packed_weights()            # the replay changed the parameters without any Python running (inference caches)
        if not self.capture_opt:
            self.opt.step()
        return self.loss
