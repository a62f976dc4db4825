"""Direct CUDA-core kernels for the HBM-bound convolutions (conv_small.cuh): Cin=1 3x3x3 stems and 1x1x1 heads with a few
output channels, against the CPU oracle (torch fp32 conv on the same bf16-rounded operands).
Tolerances: bf16 stored outputs rel <= 1e-2; fp32 outputs / fp32 parameter gradients rel <= 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import mri_epilepsy_diagnosis_b200 as pkg
    pkg._cabi.lib()
    return pkg


@pytest.mark.parametrize("co", [8, 16, 32])
@pytest.mark.parametrize("xdtype", [torch.float32, torch.bfloat16], ids=["x_fp32", "x_bf16"])
@pytest.mark.parametrize("size", [(5, 7, 9), (4, 16, 33), (3, 6, 128), (3, 11, 200)], ids=["odd", "w33", "w128", "w200"])
def test_stem_fwd_wgrad(B, co, xdtype, size):
    """Cin = 1 stems: FFMA forward, warp-MMA weight gradient (Co = 16; hi/lo bf16 split of the fp32 volume) or the FFMA one."""
    g = torch.Generator().manual_seed(co + size[2])
    ref = torch.nn.Conv3d(1, co, 3, 1, 1, bias=True)
    mod = B.nn.Conv3d(1, co, 3, 1, 1, bias=True).cuda()
    mod.load_state_dict(ref.state_dict())
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(2, 1, *size, generator=g)
    if xdtype == torch.bfloat16:
        x = x.bfloat16().float()
    yr = ref(x)
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    y = mod(x.cuda().to(xdtype))
    assert y.dtype == torch.bfloat16
    y.backward(gy.cuda().bfloat16())
    assert rel_err(y.float(), yr) < 1e-2, "forward"
    assert rel_err(mod.weight.grad, ref.weight.grad) < 1e-4, "wgrad"       # fp32 accumulation of exact bf16/fp32 products
    assert rel_err(mod.bias.grad, ref.bias.grad) < 1e-4, "bias grad"
    mod.zero_grad()
    mod(x.cuda().to(xdtype)).backward(gy.cuda().bfloat16())
    first = mod.weight.grad.clone()
    mod.zero_grad()
    mod(x.cuda().to(xdtype)).backward(gy.cuda().bfloat16())
    assert torch.equal(first, mod.weight.grad), "wgrad must be run-to-run deterministic"


def test_stem_mma_forward_opt_in():
    """The warp-MMA stem FORWARD kernel is opt-in (B200_STEM_MMA_FWD=1, read once per process): run it in a child process against
    torch's fp32 convolution -- fp32 and bf16 volumes, ragged rows, two x chunks."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
import mri_epilepsy_diagnosis_b200 as B
for xdtype, size in ((torch.float32, (3, 11, 200)), (torch.bfloat16, (4, 16, 33)), (torch.float32, (2, 8, 128))):
    g = torch.Generator().manual_seed(size[2])
    ref = torch.nn.Conv3d(1, 16, 3, 1, 1, bias=True)
    mod = B.nn.Conv3d(1, 16, 3, 1, 1, bias=True).cuda()
    mod.load_state_dict(ref.state_dict())
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(2, 1, *size, generator=g)
    if xdtype == torch.bfloat16:
        x = x.bfloat16().float()
    yr = ref(x)
    y = mod(x.cuda().to(xdtype)).float().cpu()
    err = float((y - yr).norm() / yr.norm())
    exact = float((y == yr.bfloat16().float()).float().mean())
    assert err < 3e-3 and exact > 0.98, (err, exact)
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200_STEM_MMA_FWD="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("ci", [8, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("co,out_dtype", [(2, torch.float32), (1, torch.float32), (4, torch.float32), (5, torch.bfloat16), (3, torch.bfloat16)])
def test_head_fwd_dgrad_wgrad(B, ci, co, out_dtype):
    g = torch.Generator().manual_seed(ci * 7 + co)
    ref = torch.nn.Conv3d(ci, co, 1, bias=True)
    mod = B.nn.Conv3d(ci, co, 1, bias=True).cuda()
    mod.load_state_dict(ref.state_dict())
    mod.compute_dtype, mod.out_dtype = torch.bfloat16, out_dtype
    x = torch.randn(2, ci, 5, 9, 11, generator=g).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    if out_dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    yr.backward(gy)
    xg = x.cuda().bfloat16().requires_grad_(True)
    y = mod(xg)
    assert y.dtype == out_dtype
    y.backward(gy.cuda().to(out_dtype))
    assert rel_err(y.float(), yr) < (1e-4 if out_dtype == torch.float32 else 1e-2), "forward"
    assert rel_err(xg.grad.float(), xr.grad) < 1e-2, "dgrad (bf16 stored)"
    assert rel_err(mod.weight.grad, ref.weight.grad) < 1e-4, "wgrad"
    assert rel_err(mod.bias.grad, ref.bias.grad) < 1e-4, "bias grad"


def test_head_argmax_bit_exact_on_ties(B):
    """segmentation/routine.py:226 argmax(dim=1): equal logits must stay equal (first channel wins) -- identical rows of the
    head weight give bit-identical fp32 logits from the fused dot product."""
    mod = B.nn.Conv3d(32, 2, 1, bias=True).cuda()
    with torch.no_grad():
        mod.weight[1].copy_(mod.weight[0])
        mod.bias[1].copy_(mod.bias[0])
    mod.compute_dtype, mod.out_dtype = torch.bfloat16, torch.float32
    x = torch.randn(1, 32, 4, 8, 16, generator=torch.Generator().manual_seed(0)).cuda().bfloat16()
    y = mod(x)
    assert torch.equal(y[:, 0], y[:, 1])
    assert int(y.argmax(dim=1).sum()) == 0


AXIS_CASES = [
    # name, N, Ci, Co, (D,H,W), kernel, stride, padding
    ("x_k6s2_1_8", 2, 1, 8, (24, 10, 12), (6, 1, 1), (2, 1, 1), (2, 0, 0)),        # fader encoder block 0 (AE_model.py:9-14)
    ("y_k6s2_8_8", 2, 8, 8, (12, 20, 12), (1, 6, 1), (1, 2, 1), (0, 2, 0)),
    ("z_k6s2_8_8", 2, 8, 8, (6, 10, 26), (1, 1, 6), (1, 1, 2), (0, 0, 2)),
    ("x_k6s2_8_16", 1, 8, 16, (20, 9, 7), (6, 1, 1), (2, 1, 1), (2, 0, 0)),
    ("x_k3_1_16", 2, 1, 16, (9, 8, 16), (3, 1, 1), 1, (1, 0, 0)),                  # autoencoder stem (train_AE.ipynb [cell 8])
    ("y_k3_16_1", 1, 16, 1, (5, 12, 9), (1, 3, 1), 1, (0, 1, 0)),                  # reconstruction tail
    ("z_k3_1_1", 2, 1, 1, (6, 7, 19), (1, 1, 3), 1, (0, 0, 1)),
    ("x_k3p0_16_16", 1, 16, 16, (9, 5, 6), (3, 1, 1), 1, 0),
]


@pytest.mark.parametrize("case", AXIS_CASES, ids=[c[0] for c in AXIS_CASES])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_axis_conv_fwd_dgrad_wgrad(B, case, dtype):
    """Separable 1-D convolutions with few channels (conv_axis.cuh) against torch fp32 on the same (rounded) operands."""
    name, N, Ci, Co, size, k, s, p = case
    g = torch.Generator().manual_seed(len(name) * 3 + Ci + Co)
    ref = torch.nn.Conv3d(Ci, Co, k, s, p, bias=True)
    mod = B.nn.Conv3d(Ci, Co, k, s, p, bias=True).cuda()
    mod.load_state_dict(ref.state_dict())
    mod.compute_dtype = dtype
    x = torch.randn(N, Ci, *size, generator=g)
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g)
    if dtype == torch.bfloat16:
        gy = gy.bfloat16().float()
    yr.backward(gy)
    xg = x.cuda().to(dtype).requires_grad_(True)
    yg = mod(xg)
    assert tuple(yg.shape) == tuple(yr.shape)
    yg.backward(gy.cuda().to(yg.dtype))
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(yg.float(), yr) < tol, "forward"
    assert rel_err(xg.grad.float(), xr.grad) < tol, "dgrad"
    assert rel_err(mod.weight.grad, ref.weight.grad) < (1e-4 if dtype == torch.float32 else 2e-3), "wgrad"
    assert rel_err(mod.bias.grad, ref.bias.grad) < (1e-4 if dtype == torch.float32 else 2e-3), "bias grad"
