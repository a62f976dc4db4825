#!/usr/bin/env python
"""GPU-side parity report (diagnostic, not a test): the bf16 CUDA path of unet3d.Unet against
  (a) the fp32 oracle (the reference's CPU path) and
  (b) the oracle in bf16-storage mode (oracle/graphs.py: same fp32 arithmetic, values rounded where the device stores bf16),
for logits, loss and EVERY parameter gradient.  Writes gpurun_out/parity_report.json and prints a table.

    python tools/parity_report.py [--sizes 32,128] [--norms bn,in]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="32,128")
    ap.add_argument("--norms", default="bn,in")
    args = ap.parse_args()
    import __graft_entry__
    pkg = __graft_entry__.build()
    from oracle import graphs, weights
    from conftest import tensor_core_convs
    report = []
    for size in [int(s) for s in args.sizes.split(",")]:
        n = 2 if size <= 32 else 1
        for norm in args.norms.split(","):
            for literal in (False, True):
                sd = weights.unet3d_state(1, 16, 2, norm, seed=1)
                g = torch.Generator().manual_seed(2)
                x = torch.randn(n, 1, size, size, size, generator=g)
                t = (torch.rand(n, 1, size, size, size, generator=g) > 0.5).float()
                net = pkg.zoo.Unet(c=1, n=16, dropout=0.5, norm=norm, num_classes=2, literal=literal)
                net.load_state_dict(sd, strict=True)
                net = pkg.convert(net.cuda().train(), dtype=torch.bfloat16)
                tc = tensor_core_convs(net, x.cuda())
                acts, hooks = {}, []
                if literal:        # the literal graph calls every conv / norm through its module: record what each one stores
                    for name, m in net.named_modules():
                        if isinstance(m, (torch.nn.modules.conv._ConvNd, torch.nn.modules.batchnorm._BatchNorm, torch.nn.InstanceNorm3d)):
                            def hook(mod, args, out, name=name):
                                o = out[0] if isinstance(out, tuple) else out
                                if o is not None:
                                    acts[name] = o.detach().float().cpu()
                            hooks.append(m.register_forward_hook(hook))
                logits = net(x.cuda())
                for h in hooks:
                    h.remove()
                loss = pkg.functional.softmax_dice_loss(logits, t.cuda())
                loss.backward()
                torch.cuda.synchronize()
                grads = {k: p.grad.detach().float().cpu() for k, p in net.named_parameters() if p.grad is not None}

                taps = {}

                def oracle(storage):
                    osd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
                    if storage:
                        with graphs.bf16_storage(lambda pfx, w: pfx in tc), graphs.record_taps() as tp:
                            lg = graphs.unet3d(osd, x, norm, 0.5, True, commute_up=not literal)
                            taps["storage"] = tp
                            ls = graphs.dice_loss_mean(lg, t)
                            ls.backward()
                    else:
                        with graphs.record_taps() as tp:
                            lg = graphs.unet3d(osd, x, norm, 0.5, True)
                            taps["fp32"] = tp
                        ls = graphs.dice_loss_mean(lg, t)
                        ls.backward()
                    return lg.detach(), float(ls), {k: v.grad for k, v in osd.items() if getattr(v, "grad", None) is not None}
                l32, loss32, g32 = oracle(False)
                l16, loss16, g16 = oracle(True)
                row = {"size": size, "batch": n, "norm": norm, "literal": literal, "tensor_core_convs": sorted(tc),
                       "logits_vs_fp32": rel(logits, l32), "logits_vs_storage": rel(logits, l16), "storage_vs_fp32": rel(l16, l32),
                       "loss": float(loss), "loss_fp32": loss32, "loss_storage": loss16, "grads": {}}
                for k in sorted(grads):
                    row["grads"][k] = {"vs_fp32": rel(grads[k], g32[k]), "vs_storage": rel(grads[k], g16[k]), "storage_vs_fp32": rel(g16[k], g32[k]),
                                       "norm": float(g32[k].norm())}
                row["layers"] = {}
                for k, v in acts.items():
                    if k in taps["storage"] and taps["storage"][k].shape == v.shape:
                        row["layers"][k] = {"vs_fp32": rel(v, taps["fp32"][k]), "vs_storage": rel(v, taps["storage"][k]),
                                            "mismatch_frac": float((v != taps["storage"][k]).float().mean())}
                        print(f"    [layer] {k:22s} vs_fp32 {row['layers'][k]['vs_fp32']:.2e} vs_storage {row['layers'][k]['vs_storage']:.2e} "
                              f"differing elements {row['layers'][k]['mismatch_frac']:.4f}")
                worst = max(row["grads"].items(), key=lambda kv: kv[1]["vs_storage"])
                print(f"[parity] {size}^3 x{n} norm={norm} literal={literal}: logits vs fp32 {row['logits_vs_fp32']:.2e} vs storage {row['logits_vs_storage']:.2e} "
                      f"(storage vs fp32 {row['storage_vs_fp32']:.2e}); worst grad vs storage {worst[0]} {worst[1]['vs_storage']:.2e} "
                      f"(vs fp32 {worst[1]['vs_fp32']:.2e})", flush=True)
                for k, v in row["grads"].items():
                    print(f"    {k:28s} |g|={v['norm']:.3e} vs_fp32 {v['vs_fp32']:.2e} vs_storage {v['vs_storage']:.2e} storage_vs_fp32 {v['storage_vs_fp32']:.2e}")
                report.append(row)
                del net
                torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
