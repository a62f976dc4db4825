"""Second-generation "row-slab" tcgen05 kernels (conv_row.cuh) against the CPU oracle (torch fp32 conv on the same
bf16-rounded operands).  Tolerance: rel <= 1e-2 (bf16 operands, fp32 accumulation in TMEM)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import mri_epilepsy_diagnosis_b200 as pkg
    pkg._cabi.lib()
    return pkg


WGRAD_CASES = [
    # name, N, Ci, Co, (D,H,W), kernel, padding, bias
    ("c16_16", 2, 16, 16, (10, 12, 16), 3, 1, False),
    ("c16_16_w128", 1, 16, 16, (5, 9, 128), 3, 1, True),
    ("c32_32", 2, 32, 32, (9, 20, 32), 3, 1, False),
    ("c32_32_w64", 1, 32, 32, (7, 13, 64), 3, 1, True),
    ("c16_32", 1, 16, 32, (8, 16, 48), 3, 1, True),
    ("c32_16", 1, 32, 16, (6, 16, 16), 3, 1, True),
    ("c48_16", 1, 48, 16, (8, 8, 16), 3, 1, True),
    ("c64_64", 1, 64, 64, (6, 17, 32), 3, 1, False),
    ("c96_32", 1, 96, 32, (5, 16, 16), 3, 1, True),
    ("c128_64", 1, 128, 64, (4, 16, 16), 3, 1, False),
    ("c64_128", 2, 64, 128, (5, 8, 16), 3, 1, False),
    ("c256_256", 1, 256, 256, (4, 16, 16), 3, 1, False),
    ("w192", 1, 16, 16, (3, 5, 192), 3, 1, False),
    ("w224", 1, 32, 16, (2, 4, 224), 3, 1, False),
    ("pw32_16", 1, 32, 16, (8, 16, 32), 1, 0, False),
    ("pw64_32", 2, 64, 32, (8, 16, 16), 1, 0, True),
    ("pw256_128", 1, 256, 128, (4, 16, 16), 1, 0, False),
    ("k311", 1, 16, 16, (9, 12, 32), (3, 1, 1), (1, 0, 0), True),
    ("k133", 2, 16, 32, (6, 20, 16), (1, 3, 3), (0, 1, 1), True),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_row_wgrad(B, case):
    name, N, Ci, Co, size, k, p, bias = case
    g = torch.Generator().manual_seed(len(name) * 11 + Co)
    ref = torch.nn.Conv3d(Ci, Co, k, 1, p, bias=bias)
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
    assert B._cabi.lib().b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_WGRAD) == B._cabi.ALGO_ROW, "case is meant to hit the row-slab wgrad"
    mod(xg).backward(gy.cuda().bfloat16())
    torch.cuda.synchronize()
    assert rel_err(mod.weight.grad, ref.weight.grad) < 1e-2, "wgrad"
    if bias:
        assert rel_err(mod.bias.grad, ref.bias.grad) < 1e-2, "bias grad"


def test_row_wgrad_deterministic_and_exact_on_integers(B):
    """Small-integer operands make every product and partial sum exactly representable: the tcgen05 result must equal
    the fp32 oracle bit for bit, and two runs must agree bit for bit (fixed-order reduction of the per-CTA partials)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randint(-3, 4, (2, 32, 12, 24, 64), generator=g).float()
    gy = torch.randint(-2, 3, (2, 32, 12, 24, 64), generator=g).float()
    ref = torch.nn.Conv3d(32, 32, 3, 1, 1, bias=False)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(gy)
    mod = B.nn.Conv3d(32, 32, 3, 1, 1, bias=False).cuda()
    mod.compute_dtype = torch.bfloat16
    grads = []
    for _ in range(2):
        mod.zero_grad()
        mod(x.cuda().bfloat16()).backward(gy.cuda().bfloat16())
        grads.append(mod.weight.grad.clone())
    assert torch.equal(grads[0], grads[1])
    assert torch.equal(grads[0].cpu(), ref.weight.grad)
