"""Data parallelism for the reference's unmodified training loops.

One process per GPU.  The loops take `model` and `optimizer` as arguments
(segmentation/routine.py:261, classification/routine.py:15), so data parallelism is attached from outside:

  * every parameter gets a post-accumulate-grad hook that copies its gradient into one flat fp32 bucket;
    when the last expected gradient of the step has landed the bucket is all-reduced (NCCL over NVLink on
    GPUs, gloo in the CPU tests) on a side stream, overlapping the rest of backward;
  * an optimizer step-pre-hook waits for the collective, divides by the world size and hands the averaged
    gradients back.  Parameters that received no gradient on ANY rank this step (unet3d's dead conv2/bn2
    branch, unet3d.py:43-46; frozen fader sub-networks, train_ENC_CLF.ipynb [cell 16]) keep grad=None, so the
    optimizer skips them exactly as on one GPU (no weight decay, no Adam state); the ranks agree on that set
    with one small MAX all-reduce of a seen-mask.  A parameter seen on some ranks only contributes zeros from
    the others;
  * a second backward before optimizer.step() (gradient accumulation) marks the bucket dirty: its collective
    is re-issued from the accumulated p.grad values at step time instead of racing the one in flight;
  * BatchNorm statistics are all-reduced when the model was converted with `sync=` (nn.convert).

`optimizer.step()` at routine.py:278 therefore runs unmodified.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradientBucket:
    def __init__(self, params, optimizer, process_group=None, buckets=2):
        self.params = [p for p in params if p.requires_grad]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        dev = self.params[0].device
        # reverse order: late layers finish backward first, so they sit in the bucket that is reduced first
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        per = max(1, -(-total // max(1, buckets)))
        self.flat, self.slots, self.members = [], {}, []
        cur, size = [], 0
        for p in order:
            cur.append(p)
            size += p.numel()
            if size >= per:
                self._close(cur, dev)
                cur, size = [], 0
        if cur:
            self._close(cur, dev)
        self.seen = [set() for _ in self.flat]
        self.dirty = [False] * len(self.flat)
        self.works = [None] * len(self.flat)
        self.index = {p: i for i, p in enumerate(self.params)}
        self.comm_stream = torch.cuda.Stream() if dev.type == "cuda" else None
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.handles.append(optimizer.register_step_pre_hook(self._before_step))

    def _close(self, plist, dev):
        buf = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=dev)
        b = len(self.flat)
        off = 0
        for p in plist:
            self.slots[p] = (b, off, p.numel())
            off += p.numel()
        self.flat.append(buf)
        self.members.append(list(plist))

    def _on_grad(self, p):
        b, off, n = self.slots[p]
        if p in self.seen[b]:
            # gradient accumulation: p.grad now holds the SUM of several backward passes while the bucket (and possibly a
            # collective in flight) holds the first one -- re-reduce this bucket from p.grad at step time
            self.dirty[b] = True
            return
        self.flat[b][off:off + n].copy_(p.grad.reshape(-1))
        self.seen[b].add(p)
        # fire once every member that CAN still get a gradient has one; members without grads are resolved at step time
        if len(self.seen[b]) == len(self.members[b]):
            self._launch(b)

    def _launch(self, b):
        if self.works[b] is not None or self.world == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.works[b] = dist.all_reduce(self.flat[b], group=self.pg, async_op=True)
        else:
            self.works[b] = dist.all_reduce(self.flat[b], group=self.pg, async_op=True)

    def _seen_anywhere(self):
        """per-parameter flag: some rank produced a gradient this step (one small MAX all-reduce)"""
        mask = torch.zeros(len(self.params), dtype=torch.int32, device=self.flat[0].device)
        idx = [self.index[p] for s in self.seen for p in s]
        if idx:
            mask[torch.tensor(idx, device=mask.device)] = 1
        if self.world > 1:
            dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=self.pg)
        return mask.cpu().tolist()

    def _before_step(self, optimizer, args, kwargs):
        anywhere = self._seen_anywhere()
        for b, buf in enumerate(self.flat):
            if self.dirty[b]:
                if self.works[b] is not None:           # drain the collective issued after the first backward
                    self.works[b].wait()
                    if self.comm_stream is not None:
                        torch.cuda.current_stream().wait_stream(self.comm_stream)
                    self.works[b] = None
                for p in self.seen[b]:
                    _, off, n = self.slots[p]
                    buf[off:off + n].copy_(p.grad.reshape(-1))
                self.dirty[b] = False
            # parameters without a local gradient this step contribute zeros
            for p in self.members[b]:
                if p not in self.seen[b]:
                    _, off, n = self.slots[p]
                    buf[off:off + n].zero_()
            self._launch(b)
            if self.works[b] is not None:
                self.works[b].wait()
                if self.comm_stream is not None:
                    torch.cuda.current_stream().wait_stream(self.comm_stream)
            if self.world > 1:
                buf.div_(self.world)
            for p in self.members[b]:
                _, off, n = self.slots[p]
                g = buf[off:off + n].view_as(p)
                if not anywhere[self.index[p]]:
                    continue                            # no gradient on any rank: grad stays None, the optimizer skips it
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            self.seen[b].clear()
            self.works[b] = None

    def remove(self):
        for h in self.handles:
            h.remove()


def attach(model, optimizer, process_group=None, buckets=2):
    """Attach gradient all-reduce to (model, optimizer); returns the GradientBucket (call .remove() to detach)."""
    return GradientBucket(list(model.parameters()), optimizer, process_group, buckets)


def broadcast_parameters(model, src=0, process_group=None):
    """Make every rank start from rank `src`'s weights and buffers."""
    if not dist.is_initialized():
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src, group=process_group)


def shard(batch_indices, rank=None, world=None):
    """Contiguous shard of a list of volumes / patches for this rank (config 3/4/5 partitioning)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    n = len(batch_indices)
    per = -(-n // world)
    return batch_indices[rank * per:min(n, (rank + 1) * per)]
