"""tcgen05/TMEM/TMA implicit-GEMM convolution (conv_umma.cuh) against the CPU oracle and against the SIMT kernels.

bf16 operands, fp32 accumulation: tolerance rel <= 1e-2 against the fp32 oracle on the same bf16-rounded inputs
(observed ~3e-3: the only differences are the bf16 rounding of the weights and of the stored output)."""
import ctypes

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


CASES = [
    # name, N, Ci, Co, (D,H,W), kernel, padding, bias
    ("c16_16", 2, 16, 16, (10, 12, 16), 3, 1, False),
    ("c16_32", 1, 16, 32, (8, 16, 24), 3, 1, True),
    ("c32_32", 2, 32, 32, (9, 20, 11), 3, 1, False),
    ("c32_16", 1, 32, 16, (6, 16, 8), 3, 1, True),
    ("c48_16", 1, 48, 16, (8, 8, 16), 3, 1, True),
    ("c64_64", 1, 64, 64, (6, 17, 9), 3, 1, False),
    ("c96_32", 1, 96, 32, (5, 16, 16), 3, 1, True),
    ("c128_64", 1, 128, 64, (4, 16, 8), 3, 1, False),
    ("c64_128", 2, 64, 128, (5, 8, 8), 3, 1, False),
    ("c256_128", 1, 256, 128, (4, 8, 16), 3, 1, False),
    ("c256_256", 1, 256, 256, (3, 16, 16), 3, 1, False),
    # deep U-Net levels: fewer tile groups than SMs -> split-K over channel chunks and filter planes, fp32 partials + reduce
    ("deep128_256", 2, 128, 256, (8, 8, 8), 3, 1, True),
    ("deep256_256", 2, 256, 256, (8, 8, 8), 3, 1, False),
    ("deep256_128", 1, 256, 128, (8, 8, 8), 3, 1, True),
    ("deep128_128", 2, 128, 128, (16, 16, 16), 3, 1, False),
    ("deep192_64", 1, 192, 64, (16, 16, 16), 3, 1, True),
    ("pw64_32", 1, 64, 32, (8, 16, 8), 1, 0, False),
    ("pw256_128", 1, 256, 128, (4, 16, 16), 1, 0, False),
    ("valid3", 1, 32, 64, (7, 14, 30), 3, 0, True),
    ("k133", 2, 16, 32, (6, 20, 12), (1, 3, 3), (0, 1, 1), True),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_umma_conv_fwd_dgrad(B, case):
    name, N, Ci, Co, size, k, p, bias = case
    g = torch.Generator().manual_seed(len(name) * 7 + Ci)
    ref = torch.nn.Conv3d(Ci, Co, k, 1, p, bias=bias)
    with torch.no_grad():
        ref.weight.copy_((torch.randn(ref.weight.shape, generator=g) * (2.0 / (Ci * np.prod(ref.kernel_size))) ** 0.5).bfloat16().float())
    mod = B.nn.Conv3d(Ci, Co, k, 1, p, bias=bias).cuda()
    mod.load_state_dict(ref.state_dict())
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(N, Ci, *size, generator=g).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    xg = x.cuda().bfloat16().requires_grad_(True)
    cd, _ = mod._cfg().desc(xg, mod.weight, torch.bfloat16)
    assert B._cabi.lib().b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_FWD) in (B._cabi.ALGO_UMMA, B._cabi.ALGO_ROW), "case is meant to hit a tcgen05 path"
    assert B._cabi.lib().b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_DGRAD) in (B._cabi.ALGO_UMMA, B._cabi.ALGO_ROW)
    assert B._cabi.lib().b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_WGRAD) in (B._cabi.ALGO_UMMA, B._cabi.ALGO_ROW)
    yg = mod(xg)
    yg.backward(gy.cuda().bfloat16())
    torch.cuda.synchronize()
    assert rel_err(yg.float(), yr) < 1e-2, "forward"
    assert rel_err(xg.grad.float(), xr.grad) < 1e-2, "dgrad"
    assert rel_err(mod.weight.grad, ref.weight.grad) < 1e-2, "wgrad"
    # and against the fp32-FMA kernels on identical operands: only output rounding separates them
    mod.allow_umma = False
    ys = mod(xg.detach())
    assert rel_err(yg.float(), ys.float()) < 6e-3


def test_umma_conv2d_patch_model_shapes(B):
    """detection PatchModel blocks (model_utils.py:44-52): 2-D 3x3 valid convs on (N, C, 14..6, 30..22)."""
    g = torch.Generator().manual_seed(3)
    for Ci, Co, hw in ((16, 32, (14, 30)), (32, 64, (12, 28)), (64, 128, (10, 26)), (128, 256, (8, 24))):
        ref = torch.nn.Conv2d(Ci, Co, 3)
        mod = B.nn.Conv2d(Ci, Co, 3).cuda()
        mod.load_state_dict(ref.state_dict())
        mod.compute_dtype = torch.bfloat16
        x = torch.randn(24, Ci, *hw, generator=g).bfloat16().float()
        y = mod(x.cuda().bfloat16())
        assert rel_err(y.float(), ref(x)) < 1e-2, (Ci, Co)


def test_umma_large_volume_matches_simt(B):
    """BASELINE config-2 layer shape (32->32 at 128^3 is too slow for the oracle): tcgen05 vs SIMT on 1x32ch 48x64x40."""
    g = torch.Generator().manual_seed(1)
    mod = B.nn.Conv3d(32, 32, 3, 1, 1, bias=False).cuda()
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(1, 32, 48, 64, 40, generator=g).cuda().bfloat16()
    y = mod(x)
    mod.allow_umma = False
    assert rel_err(y.float(), mod(x).float()) < 6e-3
    # linearity (size-independent property): conv(a*x) == a*conv(x) exactly for a power of two
    mod.allow_umma = True
    assert torch.equal(mod(x * 2), y * 2)
    # wgrad: tcgen05 vs fp32-FMA kernels on identical bf16 operands, and run-to-run determinism
    gy = torch.randn(y.shape, generator=g).cuda().bfloat16()
    grads = []
    for umma in (True, True, False):
        mod.allow_umma = umma
        mod.zero_grad()
        mod(x).backward(gy)
        grads.append(mod.weight.grad.clone())
    assert torch.equal(grads[0], grads[1])
    assert rel_err(grads[0], grads[2]) < 2e-3
