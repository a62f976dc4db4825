"""Row-slab tcgen05 forward / dgrad kernel (conv_rowf.cuh) against the CPU oracle (torch fp32 conv on the same bf16-rounded
operands).  Tolerance: rel <= 1e-2 (bf16 operands and bf16 stored output, fp32 accumulation in TMEM)."""
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


CASES = [
    # name, N, Ci, Co, (D,H,W), kernel, padding, bias
    ("c16_16_w128", 1, 16, 16, (5, 21, 128), 3, 1, False),       # per-row tiling, ragged last row block
    ("c32_32_w128", 1, 32, 32, (6, 7, 128), 3, 1, True),
    ("c16_32_w64", 2, 16, 32, (9, 20, 64), 3, 1, False),         # linear tiling across rows (66-slot pitch)
    ("c32_16_w64", 1, 32, 16, (7, 13, 64), 3, 1, True),
    ("c32_32_w32", 2, 32, 32, (12, 30, 32), 3, 1, False),
    ("c16_16_w48", 1, 16, 16, (8, 16, 48), 3, 1, True),
    ("c16_16_w24", 2, 16, 16, (10, 12, 24), 3, 1, False),
    ("c16_16_w192", 1, 16, 16, (3, 9, 192), 3, 1, False),         # config-3 width
    ("c32_16_w224", 1, 32, 16, (4, 6, 224), 3, 1, True),
    ("c16_16_w254", 1, 16, 16, (3, 6, 254), 3, 1, False),
    ("pw32_16", 1, 32, 16, (8, 16, 32), 1, 0, False),
    ("pw32_16_w64", 2, 32, 16, (6, 64, 64), 1, 0, False),        # tall row block, partial last tile reads past the plane image
    ("pw64_32_w128", 2, 64, 32, (3, 16, 128), 1, 0, True),
    ("pw64_64_w40", 1, 64, 64, (8, 13, 40), 1, 0, False),
    ("k133", 2, 16, 32, (7, 20, 16), (1, 3, 3), (0, 1, 1), True),
    ("k311", 1, 16, 16, (12, 12, 32), (3, 1, 1), (1, 0, 0), True),
    # streamed weights (all 27 taps do not fit in shared memory), plane ring of 3 with early release
    ("s64_64_w64", 1, 64, 64, (7, 15, 64), 3, 1, False),
    ("s64_32_w64", 1, 64, 32, (8, 9, 64), 3, 1, True),
    ("s32_64_w32", 2, 32, 64, (6, 20, 32), 3, 1, False),
    ("s64_128_w16", 2, 64, 128, (9, 16, 16), 3, 1, True),
    ("s16_256_w24", 1, 16, 256, (9, 20, 24), 3, 1, False),
    ("s64_64_d1", 4, 64, 64, (1, 33, 32), 3, 1, False),
    ("s64_64_d2", 2, 64, 64, (2, 27, 40), 3, 1, False),
    ("d1", 4, 32, 32, (1, 33, 64), 3, 1, False),                  # single plane: both z neighbours are padding
    ("d2", 2, 16, 16, (2, 17, 128), 3, 1, False),
    # kx-folded mode (see row_fwd_fold_geom): one MMA per (kz, ky) with N = 3*Cout, shifted sum in the epilogue; W = 128 cases first
    ("f32_32_w128", 2, 32, 32, (3, 9, 128), 3, 1, True),
    ("f32_16_w128", 2, 32, 16, (3, 8, 128), 3, 1, False),
    ("f64_16_w128", 1, 64, 16, (4, 9, 128), 3, 1, False),
    ("f64_32_w128", 1, 64, 32, (4, 9, 128), 3, 1, True),
    ("f133_w128", 2, 16, 16, (6, 10, 128), (1, 3, 3), (0, 1, 1), True),
    ("f16_16_w128_tall", 1, 16, 16, (3, 37, 128), 3, 1, False),
    ("f16_32_w128", 1, 16, 32, (4, 9, 128), 3, 1, True),          # folded since round 2 (two 16-channel epilogue passes per tile)
    # folded with 2 / 4 whole rows per tile (W = 64 / 32; c16_32_w64, c32_16_w64, c32_32_w32 above fold too), ragged last tiles
    ("f32_32_w64", 2, 32, 32, (5, 13, 64), 3, 1, True),
    ("f16_16_w32", 1, 16, 16, (6, 23, 32), 3, 1, False),
    ("f133_w64", 2, 32, 16, (6, 10, 64), (1, 3, 3), (0, 1, 1), True),
    # folded with Cout = 64 (N = 192, Cin <= 32): the two epilogue warp sets split the channels of every tile
    ("f32_64_w64", 2, 32, 64, (6, 13, 64), 3, 1, True),
    ("f16_64_w128", 2, 16, 64, (3, 9, 128), 3, 1, False),
    ("f32_64_w32", 2, 32, 64, (5, 22, 32), 3, 1, True),
    # folded with streamed weights (Cin = 64: the nine folded blocks do not fit next to a plane ring)
    ("fs64_32_w64", 2, 64, 32, (6, 18, 64), 3, 1, True),
    ("fs64_16_w32", 1, 64, 16, (9, 21, 32), 3, 1, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_row_fwd_dgrad(B, case):
    name, N, Ci, Co, size, k, p, bias = case
    g = torch.Generator().manual_seed(len(name) * 13 + Ci)
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
    lib = B._cabi.lib()
    assert lib.b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_FWD) == B._cabi.ALGO_ROW, "case is meant to hit the row-slab forward kernel"
    # dgrad gathers Co channels: more than one swizzle atom per voxel (Co > 64) stays on the first-generation tile kernel
    assert lib.b200_conv_algo(ctypes.byref(cd), B._cabi.PASS_DGRAD) == (B._cabi.ALGO_ROW if Co <= 64 else B._cabi.ALGO_UMMA)
    yg = mod(xg)
    yg.backward(gy.cuda().bfloat16())
    torch.cuda.synchronize()
    assert rel_err(yg.float(), yr) < 1e-2, "forward"
    assert rel_err(xg.grad.float(), xr.grad) < 1e-2, "dgrad"
    assert rel_err(mod.weight.grad, ref.weight.grad) < 1e-2, "wgrad"


def test_row_fwd_exact_on_integers_and_deterministic(B):
    """Small-integer operands: every product and partial sum is exact in fp32 and the result fits bf16, so the tcgen05
    output must equal the oracle bit for bit (this pins the tap -> slot-shift mapping at every border)."""
    g = torch.Generator().manual_seed(9)
    for (Ci, Co, size) in ((16, 16, (4, 10, 128)), (32, 32, (5, 9, 64)), (16, 32, (3, 18, 40))):
        x = torch.randint(-2, 3, (2, Ci) + size, generator=g).float()
        ref = torch.nn.Conv3d(Ci, Co, 3, 1, 1, bias=False)
        with torch.no_grad():
            ref.weight.copy_(torch.randint(-1, 2, ref.weight.shape, generator=g).float())
        mod = B.nn.Conv3d(Ci, Co, 3, 1, 1, bias=False).cuda()
        mod.load_state_dict(ref.state_dict())
        mod.compute_dtype = torch.bfloat16
        y1 = mod(x.cuda().bfloat16())
        y2 = mod(x.cuda().bfloat16())
        yr = ref(x)
        assert float(yr.abs().max()) < 256          # exactly representable in bf16
        assert torch.equal(y1, y2)
        assert torch.equal(y1.float().cpu(), yr), (Ci, Co, size)


def test_row_fwd_large_volume_properties(B):
    """BASELINE config-2 layer shapes are too slow for the CPU oracle at full size: check the row-slab kernel against the
    first-generation tile kernel's maths through size-independent properties on a 1 x 32ch x 64 x 128 x 128 volume."""
    g = torch.Generator().manual_seed(1)
    mod = B.nn.Conv3d(32, 32, 3, 1, 1, bias=False).cuda()
    mod.compute_dtype = torch.bfloat16
    x = torch.randn(1, 32, 64, 128, 128, generator=g).cuda().bfloat16()
    y = mod(x)
    # linearity: conv(2x) == 2 conv(x) exactly
    assert torch.equal(mod(x * 2), y * 2)
    # translation along z by one plane (interior planes): conv(shift(x)) == shift(conv(x))
    xs = torch.roll(x, 1, dims=2)
    ys = mod(xs)
    assert torch.equal(ys[:, :, 3:-3], torch.roll(y, 1, dims=2)[:, :, 3:-3])
    # against the fp32-FMA kernels on identical operands: only output rounding separates them
    mod.allow_umma = False
    assert rel_err(y.float(), mod(x).float()) < 6e-3


@pytest.mark.parametrize("W,Ci", [(128, 16), (128, 32), (64, 16)], ids=["folded_16", "folded_32", "folded_w64"])
def test_fused_statistics_and_dual_conv(B, W, Ci):
    """conv + BatchNorm partial sums in the epilogue (b200_conv_fwd_stats) and the dead/live pair of unet3d.py:43-46 in one launch
    (b200_conv_fwd_stats_tail), in the kx-folded mode with one and with two rows per tile: statistics against the stored outputs."""
    F_ = B.functional
    torch.manual_seed(1000 + W + Ci)                     # the two modules below draw their weights from the global generator
    g = torch.Generator().manual_seed(W + Ci)
    x = torch.randn(2, Ci, 5, 12, W, generator=g).cuda().bfloat16()
    conv = B.nn.Conv3d(Ci, 16, 3, 1, 1, bias=False).cuda()
    conv.compute_dtype = torch.bfloat16
    y, part = F_.conv(x, conv.weight, None, conv._cfg(), torch.bfloat16, want_stats=True)
    assert part is not None
    y_plain = conv(x)
    assert torch.equal(y, y_plain)
    s = part.double().sum(0)                                                   # sums of the fp32 accumulators, before the bf16 rounding of y
    yd = y.double().permute(1, 0, 2, 3, 4).reshape(16, -1)
    # the per-channel sums of ~15 k zero-mean values cancel to ~1 % of sum|y|, so bf16 rounding of y (2^-9 each) shows up as a
    # few 1e-3 of the sum: 5e-3 bounds it, an indexing error would give O(1)
    assert rel_err(s[0], yd.sum(1)) < 5e-3 and rel_err(s[1], (yd * yd).sum(1)) < 2e-3
    dead = B.nn.Conv3d(Ci, 16, 3, 1, 1, bias=False).cuda()
    assert F_.dual_conv_supported(x, dead.weight, conv.weight, conv._cfg(), torch.bfloat16)
    y3, part2 = F_.dual_conv(x, dead.weight, conv.weight, conv._cfg(), torch.bfloat16)
    # the live half is the plain convolution: bit-equal when both launches use the same mode, else equal up to the summation order
    assert torch.equal(y3, y_plain) or rel_err(y3.float(), y_plain.float()) < 2e-3
    dead.compute_dtype = torch.bfloat16
    yd2 = dead(x).double().permute(1, 0, 2, 3, 4).reshape(16, -1)
    s2 = part2.double().sum(0)
    assert rel_err(s2[0, :16], yd2.sum(1)) < 5e-3 and rel_err(s2[1, :16], (yd2 * yd2).sum(1)) < 2e-3
    assert rel_err(s2[0, 16:], yd.sum(1)) < 5e-3 and rel_err(s2[1, 16:], (yd * yd).sum(1)) < 2e-3


def test_fused_statistics_folded_64_channels(B):
    """Cout = 64 in the kx-folded mode (each epilogue warp set owns 32 channels of every tile): fused BatchNorm sums and the dead/live
    pair with 32 + 32 channels (the 64^3 level of unet3d.py:43-46), statistics against the stored outputs."""
    F_ = B.functional
    torch.manual_seed(77)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(2, 32, 5, 12, 64, generator=g).cuda().bfloat16()
    conv = B.nn.Conv3d(32, 64, 3, 1, 1, bias=False).cuda()
    conv.compute_dtype = torch.bfloat16
    y, part = F_.conv(x, conv.weight, None, conv._cfg(), torch.bfloat16, want_stats=True)
    assert part is not None and torch.equal(y, conv(x))
    ref = torch.nn.functional.conv3d(x.float().cpu(), conv.weight.detach().bfloat16().float().cpu(), padding=1)
    assert rel_err(y.float(), ref) < 1e-2
    s = part.double().sum(0)
    yd = y.double().permute(1, 0, 2, 3, 4).reshape(64, -1)
    assert rel_err(s[0], yd.sum(1)) < 5e-3 and rel_err(s[1], (yd * yd).sum(1)) < 2e-3
    live, dead = B.nn.Conv3d(32, 32, 3, 1, 1, bias=False).cuda(), B.nn.Conv3d(32, 32, 3, 1, 1, bias=False).cuda()
    live.compute_dtype = dead.compute_dtype = torch.bfloat16
    assert F_.dual_conv_supported(x, dead.weight, live.weight, live._cfg(), torch.bfloat16)
    y3, part2 = F_.dual_conv(x, dead.weight, live.weight, live._cfg(), torch.bfloat16)
    yl, ydd = live(x), dead(x)
    assert torch.equal(y3, yl) or rel_err(y3.float(), yl.float()) < 2e-3
    s2 = part2.double().sum(0)
    for half, yy in ((slice(0, 32), ydd), (slice(32, 64), yl)):
        yv = yy.double().permute(1, 0, 2, 3, 4).reshape(32, -1)
        assert rel_err(s2[0, half], yv.sum(1)) < 5e-3 and rel_err(s2[1, half], (yv * yv).sum(1)) < 2e-3
