"""Deterministic synthetic state dicts with the reference models' exact keys/shapes
(TEST INFRASTRUCTURE).  Values come from a seeded CPU torch.Generator, so the build
container and the GPU box produce identical tensors without shipping checkpoints.

Key lists follow: unet3d.py:20-126 (162 keys at c=1,n=16), the shipped
segmentation/weights/*.pth (unet.UNet, 154 keys at F=8), AE_model.py:4-312 and
detection/model_utils.py:19-52.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


class _Gen:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = OrderedDict()

    def conv(self, pfx, cout, cin, k, bias=True):
        k = (k,) * 3 if isinstance(k, int) else tuple(k)
        fan_in = cin * math.prod(k)
        self.sd[pfx + ".weight"] = torch.randn((cout, cin) + k, generator=self.g) * math.sqrt(2.0 / fan_in)
        if bias:
            self.sd[pfx + ".bias"] = torch.randn(cout, generator=self.g) * 0.1

    def linear(self, pfx, cout, cin):
        self.sd[pfx + ".weight"] = torch.randn(cout, cin, generator=self.g) * math.sqrt(1.0 / cin)
        self.sd[pfx + ".bias"] = torch.randn(cout, generator=self.g) * 0.1

    def bn(self, pfx, c, buffers=True):
        self.sd[pfx + ".weight"] = torch.rand(c, generator=self.g) + 0.5
        self.sd[pfx + ".bias"] = torch.randn(c, generator=self.g) * 0.1
        if buffers:
            self.sd[pfx + ".running_mean"] = torch.randn(c, generator=self.g) * 0.1
            self.sd[pfx + ".running_var"] = torch.rand(c, generator=self.g) + 0.5
            self.sd[pfx + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    def prelu(self, pfx):
        self.sd[pfx + ".weight"] = torch.rand(1, generator=self.g) * 0.4 + 0.05


def unet3d_state(c=1, n=16, num_classes=2, norm="bn", seed=0):
    g = _Gen(seed)

    def nrm(pfx, ch):
        if norm == "bn":
            g.bn(pfx, ch)
        elif norm == "gn":
            g.bn(pfx, ch, buffers=False)
        # 'in': InstanceNorm3d(affine=False) has no state

    chans = [c, n, 2 * n, 4 * n, 8 * n, 16 * n]
    for i in range(1, 6):
        p = f"convd{i}"
        for j, cin in ((1, chans[i - 1]), (2, chans[i]), (3, chans[i])):
            g.conv(f"{p}.conv{j}", chans[i], cin, 3, bias=False)
            nrm(f"{p}.bn{j}", chans[i])
    for lvl, planes, first in ((4, 16 * n, True), (3, 8 * n, False), (2, 4 * n, False), (1, 2 * n, False)):
        p = f"convu{lvl}"
        if not first:
            g.conv(p + ".conv1", planes, 2 * planes, 3, bias=False)
            nrm(p + ".bn1", planes)
        g.conv(p + ".conv2", planes // 2, planes, 1, bias=False)
        nrm(p + ".bn2", planes // 2)
        g.conv(p + ".conv3", planes, planes, 3, bias=False)
        nrm(p + ".bn3", planes)
    g.conv("seg3", num_classes, 8 * n, 1)
    g.conv("seg2", num_classes, 4 * n, 1)
    g.conv("seg1", num_classes, 2 * n, 1)
    return g.sd


def fepegar_unet_state(first=8, in_channels=1, out_classes=2, num_encoding_blocks=3, seed=0, duplicate_keys=True):
    """Keys of `unet.UNet` as found in segmentation/weights/*.pth.  Every ConvolutionalBlock
    registers its layers twice (conv_layer/norm_layer/activation_layer and block.{0,1,2});
    with duplicate_keys=True both names are emitted (sharing storage, like the module)."""
    g = _Gen(seed)

    def block(pfx, cin, cout, norm=True, act=True):
        g.conv(pfx + ".conv_layer", cout, cin, 3)
        names = [("conv_layer", ("weight", "bias"))]
        if norm:
            g.bn(pfx + ".norm_layer", cout)
            names.append(("norm_layer", ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")))
        if act:
            g.prelu(pfx + ".activation_layer")
            names.append(("activation_layer", ("weight",)))
        if duplicate_keys:
            for idx, (nm, fields) in enumerate(names):
                for f in fields:
                    g.sd[f"{pfx}.block.{idx}.{f}"] = g.sd[f"{pfx}.{nm}.{f}"]

    F = first
    cin = in_channels
    skips = []
    for i in range(num_encoding_blocks - 1):
        c1 = F * 2 ** i
        p = f"encoder.encoding_blocks.{i}"
        block(p + ".conv1", cin, c1, norm=(i > 0))
        block(p + ".conv2", c1, 2 * c1)
        cin = 2 * c1
        skips.append(cin)
    block("bottom_block.conv1", cin, cin)
    block("bottom_block.conv2", cin, 2 * cin)
    cin = 2 * cin
    for i in range(num_encoding_blocks - 1):
        skip = skips[-1 - i]
        p = f"decoder.decoding_blocks.{i}"
        block(p + ".conv1", cin + skip, skip)
        block(p + ".conv2", skip, skip)
        cin = skip
    g.conv("classifier.conv_layer", out_classes, cin, 1)
    if duplicate_keys:
        g.sd["classifier.block.0.weight"] = g.sd["classifier.conv_layer.weight"]
        g.sd["classifier.block.0.bias"] = g.sd["classifier.conv_layer.bias"]
    return g.sd


def _sep3(g, pfx, names, cin, cout, k):
    g.conv(f"{pfx}.{names[0]}", cout, cin, (k, 1, 1))
    g.conv(f"{pfx}.{names[1]}", cout, cout, (1, k, 1))
    g.conv(f"{pfx}.{names[2]}", cout, cout, (1, 1, k))


def encoder_state(channels=(1, 8, 16, 32), k=6, batch_norm=True, seed=0, pfx="encode", g=None):
    g = g or _Gen(seed)
    for i in range(len(channels) - 1):
        _sep3(g, f"{pfx}.{i}.block", ("1_convx", "2_convy", "3_convz"), channels[i], channels[i + 1], k)
        if batch_norm:
            g.bn(f"{pfx}.{i}.block.5_batch_norm", channels[i + 1])
    return g.sd


def ae_state(depth=6, c_base=16, inc=2, c_in=1, k=3, seed=0):
    chans = [c_in] + [c_base * inc ** i for i in range(depth)]
    g = _Gen(seed)
    encoder_state(chans, k, True, pfx="enc.encode", g=g)
    rev = chans[::-1]
    for i in range(depth):
        _sep3(g, f"dec.decode.{i}.block", ("2_convx", "3_convy", "4_convz"), rev[i], rev[i + 1], k)
        g.bn(f"dec.decode.{i}.block.5_batch_norm", rev[i + 1])
    g.conv("dec.vox", 1, 1, 3)
    return g.sd


def fader_head_state(pfx="clf", c_in=32, c_out=64, k=3, l_in=64, l_out=32, n_out=2, seed=0):
    g = _Gen(seed)
    _sep3(g, pfx, ("1_convx", "2_convy", "3_convz"), c_in, c_out, k)
    g.linear(pfx + ".5_l1", l_out, l_in)
    g.bn(pfx + ".6_batch_norm", l_out)
    g.linear(pfx + ".9_l_f", n_out, l_out)
    return g.sd


def patch_model_state(seed=0):
    g = _Gen(seed)
    ch = (2, 16, 32, 64, 128, 256)
    for i in range(5):
        k = (3, 3)
        fan_in = ch[i] * 9
        g.sd[f"conv_blocks.{i}.conv.weight"] = torch.randn((ch[i + 1], ch[i]) + k, generator=g.g) * math.sqrt(2.0 / fan_in)
        g.sd[f"conv_blocks.{i}.conv.bias"] = torch.randn(ch[i + 1], generator=g.g) * 0.1
        g.bn(f"conv_blocks.{i}.bn", ch[i + 1])
    g.linear("fc1", 256, 3 * 11 * 256)
    g.linear("fc2", 2, 256)
    return g.sd


def clone_state(sd):
    return OrderedDict((k, v.clone()) for k, v in sd.items())


def synthetic_t1w(shape, seed=0):
    """Synthetic z-normalised T1w-like volume: smooth ellipsoid 'brain' + noise, background
    exactly 0 (what torchio ZNormalization + CropOrPad leave) -- SURVEY section 8(d) config 1."""
    n, c, d, h, w = shape
    g = torch.Generator().manual_seed(seed)
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, d), torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    r = (zz / 0.8) ** 2 + (yy / 0.9) ** 2 + (xx / 0.7) ** 2
    brain = (r < 1).float()
    vol = brain * (1.0 + 0.5 * torch.cos(6 * zz) * torch.sin(5 * yy) + 0.3 * xx)
    out = vol[None, None].repeat(n, c, 1, 1, 1) + 0.1 * torch.randn(shape, generator=g) * brain
    m = out[brain[None, None].expand_as(out) > 0]
    out = (out - m.mean()) / m.std() * brain
    return out


def seeded_like(template, seed):
    """Deterministic values for ANY state dict: `template` maps key -> tensor (only shape/dtype are used), e.g. the
    `state_dict()` of a freshly constructed reference module or of its zoo mirror (identical keys and order, so both sides
    regenerate the same tensors from the seed instead of shipping checkpoints).  Conv/Linear weights ~ N(0, 2/fan_in),
    norm weights U(0.5, 1.5), biases N(0, 0.1), running_mean N(0, 0.1), running_var U(0.5, 1.5), counters 0."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, v in template.items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shape, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shape, generator=g) * 0.1
        elif k.endswith(".bias"):
            sd[k] = torch.randn(shape, generator=g) * 0.1
        elif len(shape) > 1:
            sd[k] = torch.randn(shape, generator=g) * math.sqrt(2.0 / math.prod(shape[1:]))
        else:
            sd[k] = torch.rand(shape, generator=g) + 0.5
    return sd
