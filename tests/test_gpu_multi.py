"""Multi-GPU parity, needs >= 2 GPUs on the box (skipped otherwise; run with `gpurun --gpus 2`):
N ranks x local batch == one device x the global batch, for
  * the captured data-parallel step (graphed.GraphedTrainStep: bucketed NCCL all-reduce inside the CUDA graph, overlapped with the
    backward pass; deferred weight gradients; captured optimizer) on an InstanceNorm U-Net (no cross-sample statistics), and
  * DP + SyncBN (nn.convert(sync=...)) on the BatchNorm U-Net: batch statistics all-reduced in both directions.
fp32 ("tf32-off") kernels so that the comparison is tight; the reference loop is the eager single-GPU loop of routine.py:266-281."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(B, norm, sync=None, dtype=torch.float32):
    from oracle import weights
    net = B.zoo.Unet(c=1, n=16, dropout=0.0, norm=norm, num_classes=2)
    net.load_state_dict(weights.unet3d_state(1, 16, 2, norm, seed=5), strict=True)
    return B.convert(net.cuda().train(), dtype=dtype, sync=sync)


def _batches(steps):
    g = torch.Generator().manual_seed(8)
    return [(torch.randn(4, 1, 32, 32, 32, generator=g), (torch.rand(4, 1, 32, 32, 32, generator=g) > 0.5).float()) for _ in range(steps)]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mri_epilepsy_diagnosis_b200 as B
    out = {}
    data = _batches(3)
    lo, hi = rank * 2, rank * 2 + 2
    # ---- captured DP step, InstanceNorm
    net = _make(B, "in")
    opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    x0, t0 = data[0]
    step = B.graphed.GraphedTrainStep(net, B.functional.softmax_dice_loss, opt, x0[lo:hi].cuda(), t0[lo:hi].cuda(), warmup=2, buckets=3)
    out["graph_buckets"] = len(step._members)
    losses = []
    for x, t in data:
        losses.append(float(step(x[lo:hi].cuda(), t[lo:hi].cuda())))
    out["dp_in_losses"] = losses
    out["dp_in_params"] = {k: v.detach().cpu() for k, v in net.named_parameters() if v.grad is not None}
    # ---- eager DP + SyncBN
    net = _make(B, "bn", sync=(None, world))
    opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
    bucket = B.dp.attach(net, opt)
    losses = []
    for x, t in data:
        opt.zero_grad()
        loss = B.functional.softmax_dice_loss(net(x[lo:hi].cuda()), t[lo:hi].cuda())
        loss.backward()
        opt.step()
        losses.append(float(loss))
    bucket.remove()
    out["dp_bn_losses"] = losses
    out["dp_bn_params"] = {k: v.detach().cpu() for k, v in net.named_parameters() if v.grad is not None}
    out["dp_bn_buffers"] = {k: v.detach().cpu() for k, v in net.named_buffers() if k in ("convd1.bn1.running_mean", "convu1.bn3.running_var", "convd1.bn2.running_var")}
    out["dead_grad_none"] = net.convd1.conv2.weight.grad is None
    # ---- bf16 SyncBN: the conv-epilogue partial statistics are all-reduced (no extra pass over the activations)
    net = _make(B, "bn", sync=(None, world), dtype=torch.bfloat16)
    x, t = data[0]
    loss = B.functional.softmax_dice_loss(net(x[lo:hi].cuda()), t[lo:hi].cuda())
    loss.backward()
    out["bf16_sync_loss"] = float(loss)
    out["bf16_sync_buffers"] = {k: v.detach().float().cpu() for k, v in net.named_buffers() if k in ("convd1.bn1.running_mean", "convd1.bn2.running_var", "convu1.bn3.running_var", "convd3.bn3.running_mean")}
    out["bf16_sync_grad"] = net.convu1.conv3.weight.grad.detach().cpu()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_one_device_on_the_global_batch():
    import torch.multiprocessing as mp
    import mri_epilepsy_diagnosis_b200 as B
    from conftest import rel_err
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=600) for _ in procs)
    [p.join(60) for p in procs]
    data = _batches(3)
    torch.cuda.set_device(0)
    for norm, key in (("in", "dp_in"), ("bn", "dp_bn")):
        net = _make(B, norm)
        opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
        want = []
        for x, t in data:
            opt.zero_grad()
            loss = B.functional.softmax_dice_loss(net(x.cuda()), t.cuda())
            loss.backward()
            opt.step()
            want.append(float(loss))
        for r in (0, 1):
            got = res[r][key + "_losses"]
            if norm == "bn":
                # SyncBN: every rank normalises with the GLOBAL statistics, its loss is the mean over its own half of the batch
                assert abs(0.5 * (res[0][key + "_losses"][0] + res[1][key + "_losses"][0]) - want[0]) < 2e-5
            else:
                assert abs(0.5 * (res[0][key + "_losses"][0] + res[1][key + "_losses"][0]) - want[0]) < 2e-5
            assert len(got) == 3
        for i in range(3):                                  # the trajectory stays together: same weights after every step on both ranks and on one GPU
            assert abs(0.5 * (res[0][key + "_losses"][i] + res[1][key + "_losses"][i]) - want[i]) < 1e-4, (norm, i)
        ref = dict(net.named_parameters())
        report = {k: (float((v - res[1][key + "_params"][k]).abs().max()), rel_err(v, ref[k])) for k, v in res[0][key + "_params"].items()}
        bad = {k: v for k, v in report.items() if v[0] != 0.0 or v[1] >= 2e-4}
        print(f"[multi] {norm}: {len(report)} parameters compared; not bit-identical across ranks or off the one-GPU run: {bad}")
        for k, v in res[0][key + "_params"].items():
            assert torch.equal(v, res[1][key + "_params"][k]), (norm, k, report)            # ranks hold bit-identical weights
            assert rel_err(v, ref[k]) < 2e-4, (norm, k, report)
        if norm == "bn":
            bufs = dict(net.named_buffers())
            for k, v in res[0]["dp_bn_buffers"].items():
                assert rel_err(v, bufs[k]) < 1e-4, k
            assert res[0]["dead_grad_none"] and res[1]["dead_grad_none"]
    # bf16 DP + SyncBN (statistics from the all-reduced conv-epilogue partials) against one GPU on the global batch, bf16 tolerances
    net = _make(B, "bn", dtype=torch.bfloat16)
    x, t = data[0]
    loss = B.functional.softmax_dice_loss(net(x.cuda()), t.cuda())
    loss.backward()
    assert abs(0.5 * (res[0]["bf16_sync_loss"] + res[1]["bf16_sync_loss"]) - float(loss)) < 3e-3
    bufs = dict(net.named_buffers())
    for k, v in res[0]["bf16_sync_buffers"].items():
        assert torch.equal(v, res[1]["bf16_sync_buffers"][k]), k                   # both ranks hold the same global statistics
        assert rel_err(v, bufs[k].float()) < 2e-2, k
    g = 0.5 * (res[0]["bf16_sync_grad"] + res[1]["bf16_sync_grad"])
    assert rel_err(g, net.convu1.conv3.weight.grad) < 6e-2
    assert res[0]["graph_buckets"] >= 2
