"""Host-side mirrors of the reference's model interfaces, assembled from the drop-in modules.

The reference checkout is not importable everywhere (and its `unet.UNet` dependency is
third-party and absent), so the graphs the hot path is benchmarked and tested on are
restated here with IDENTICAL constructor arguments, attribute names and state_dict keys:

  Unet            segmentation/models/unet3d.py:82-126   (ConvD :20-47, ConvU :50-79)
  FepegarUNet     `unet.UNet` as called at segmentation/routine.py:346-356; keys of segmentation/weights/*.pth
  AE/Encoder/...  classification/models/AE_model.py:4-312
  PatchModel      detection/model_utils.py:19-52

Every shipped checkpoint loads into these with strict=True.  Users of the real reference
files get the same kernels through `nn.convert(model)` / `nn.patch()`.
"""
from __future__ import annotations

import torch
import torch.nn as tnn

from . import _cabi as cabi
from . import functional as BF
from . import nn as bnn


# ------------------------------------------------------------------ unet3d.Unet
def _norm3d(planes, norm):
    """unet3d.py:8-17"""
    if norm == "bn":
        return bnn.BatchNorm3d(planes)
    if norm == "gn":
        return bnn.GroupNorm(4, planes)
    if norm == "in":
        return bnn.InstanceNorm3d(planes)
    raise ValueError("normalization type {} is not supported".format(norm))


def _c3(cin, cout, k):
    return bnn.Conv3d(cin, cout, k, 1, k // 2, bias=False)


def _conv_norm(conv, norm, x, act=cabi.ACT_NONE, residual=None, stats_only=False):
    """act(norm(conv(x)) [+ residual]) as two kernels + one finalize: the convolution's epilogue also accumulates the batch
    statistics of its own output (BatchNorm, training), the apply pass folds the residual add and the activation."""
    part = None
    if isinstance(norm, bnn.BatchNorm3d) and norm.training:          # (SyncBN all-reduces these partials, functional._NormFn)
        y, part = conv(x, want_stats=True)
    else:
        y = conv(x)
    return norm(y, residual=residual, act=act, stats_partial=part, stats_only=stats_only)


class ConvD(tnn.Module):
    """Encoder stage (unet3d.py:20-47).  conv2/bn2 feed a branch whose result the reference discards
    (:43-46); it is still executed so BatchNorm running statistics and the RNG stream match."""

    def __init__(self, inplanes, planes, dropout=0.0, norm="gn", first=False):
        super().__init__()
        self.first, self.dropout = first, dropout
        self.maxpool = bnn.MaxPool3d(2, 2)
        self.relu = bnn.ReLU(inplace=True)
        for i, cin in ((1, inplanes), (2, planes), (3, planes)):
            setattr(self, f"conv{i}", _c3(cin, planes, 3))
            setattr(self, f"bn{i}", _norm3d(planes, norm))

    literal = False       # True: execute unet3d.py:42-47 op by op (dead branch materialised), kernel-level fusions only

    def forward(self, x, pooled=False):
        if not self.first and not pooled:
            x = self.maxpool(x)
        x = _conv_norm(self.conv1, self.bn1, x)
        if self.literal:
            y = _conv_norm(self.conv2, self.bn2, x, act=cabi.ACT_RELU)                       # :43
            if self.dropout > 0:
                y = torch.nn.functional.dropout3d(y, self.dropout)                           # :44-45
            del y                                                                            # overwritten at :46
            return _conv_norm(self.conv3, self.bn3, x, act=cabi.ACT_RELU, residual=x)        # :46-47
        # unet3d.py:43-45: y = relu(bn2(conv2(x))); y = dropout3d(y) is overwritten at :46.  Its only observable effects are
        # bn2's running statistics (BatchNorm, training) and the RNG draw of dropout3d; exactly those are performed.
        bn_train = isinstance(self.bn2, bnn.BatchNorm3d) and self.bn2.training
        if self.dropout > 0:                       # F.dropout3d is always in training mode there: one Bernoulli draw per (n, c)
            x.new_empty((x.shape[0], self.conv2.out_channels, 1, 1, 1)).bernoulli_(1 - self.dropout)
        if bn_train and self.conv2.bias is None and self.conv3.bias is None and x.dtype == torch.bfloat16 and \
                BF.dual_conv_supported(x, self.conv2.weight, self.conv3.weight, self.conv3._cfg(), x.dtype):
            # conv2 and conv3 read the same x: ONE convolution with concatenated output channels gives both statistics; only
            # conv3's half is written (b200_conv_fwd_stats_tail)
            cdead = self.conv2.out_channels
            y3, part = BF.dual_conv(x, self.conv2.weight, self.conv3.weight, self.conv3._cfg(), x.dtype)
            self.bn2(y3, stats_partial=part[:, :, :cdead].contiguous(), stats_only=True)
            return self.bn3(y3, residual=x, act=cabi.ACT_RELU, stats_partial=part[:, :, cdead:].contiguous())
        if isinstance(self.bn2, bnn.BatchNorm3d) and self.bn2.training:
            _conv_norm(self.conv2, self.bn2, x, stats_only=True)
        return _conv_norm(self.conv3, self.bn3, x, act=cabi.ACT_RELU, residual=x)


class ConvU(tnn.Module):
    """Decoder stage (unet3d.py:50-79)."""

    def __init__(self, planes, norm="gn", first=False):
        super().__init__()
        self.first = first
        if not first:
            self.conv1, self.bn1 = _c3(2 * planes, planes, 3), _norm3d(planes, norm)
        self.conv2, self.bn2 = _c3(planes, planes // 2, 1), _norm3d(planes // 2, norm)
        self.conv3, self.bn3 = _c3(planes, planes, 3), _norm3d(planes, norm)
        self.relu = bnn.ReLU(inplace=True)

    literal = False       # True: conv2 runs on the upsampled tensor as written at unet3d.py:73-74

    def forward(self, x, prev, lazy_prev_grad=False):
        if not self.first:
            x = _conv_norm(self.conv1, self.bn1, x, act=cabi.ACT_RELU)
        if self.literal:
            y = BF.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
            y = _conv_norm(self.conv2, self.bn2, y, act=cabi.ACT_RELU)
            return _conv_norm(self.conv3, self.bn3, BF.concat(prev, y), act=cabi.ACT_RELU)
        # unet3d.py:73-74 computes conv2(upsample(x)).  conv2 is 1x1x1 and trilinear interpolation is a per-channel convex
        # combination of voxels, so the two commute exactly: upsample(conv2(x)) is the same function with the convolution done on
        # 1/8 of the voxels and the interpolation on half of the channels (results differ by bf16 rounding only).
        y = BF.interpolate(self.conv2(x), scale_factor=2, mode="trilinear", align_corners=False)
        y = self.bn2(y, act=cabi.ACT_RELU)
        y = BF.concat(prev, y, lazy_grad_a=lazy_prev_grad)
        return _conv_norm(self.conv3, self.bn3, y, act=cabi.ACT_RELU)


class Unet(tnn.Module):
    """unet3d.py:82-126.  `literal=True` (not a reference argument) makes every stage execute the reference's operator sequence
    one to one -- dead branch materialised, 1x1x1 conv on the upsampled tensor -- keeping only the kernel-level fusions
    (statistics in the conv epilogue, activation / residual in the apply pass); the default graph computes the same function
    with the rewrites described in DESIGN.md section 4."""

    def __init__(self, c=4, n=16, dropout=0.5, norm="gn", num_classes=5, literal=False):
        super().__init__()
        self.upsample = bnn.Upsample(scale_factor=2, mode="trilinear", align_corners=False)   # intent of unet3d.py:85
        widths = [c, n, 2 * n, 4 * n, 8 * n, 16 * n]
        for i in range(1, 6):
            setattr(self, f"convd{i}", ConvD(widths[i - 1], widths[i], dropout, norm, first=(i == 1)))
        self.convu4 = ConvU(16 * n, norm, True)
        self.convu3, self.convu2, self.convu1 = ConvU(8 * n, norm), ConvU(4 * n, norm), ConvU(2 * n, norm)
        self.seg3, self.seg2, self.seg1 = (bnn.Conv3d(w * n, num_classes, 1) for w in (8, 4, 2))
        for m in self.modules():
            if isinstance(m, (ConvD, ConvU)):
                m.literal = bool(literal)
        for m in self.modules():                                                              # unet3d.py:103-108
            if isinstance(m, tnn.Conv3d):
                tnn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (tnn.BatchNorm3d, tnn.GroupNorm)):
                tnn.init.constant_(m.weight, 1)
                tnn.init.constant_(m.bias, 0)

    def forward(self, x):
        if self.convd1.literal or not x.is_cuda:
            x1 = self.convd1(x)
            x2 = self.convd2(x1)
            x3 = self.convd3(x2)
            x4 = self.convd4(x3)
            x5 = self.convd5(x4)
            y4 = self.convu4(x5, x4)
            y3 = self.convu3(y4, x3)
            y2 = self.convu2(y3, x2)
            y1 = self.convu1(y2, x1)
        else:
            # every encoder output has two consumers (the next level through the max-pool and the decoder's concat): pool_skip
            # sums their gradients inside the pooling backward kernel, fed by the concat gradient's channel window in place
            p1, x1 = BF.pool_skip(self.convd1(x))
            p2, x2 = BF.pool_skip(self.convd2(p1, pooled=True))
            p3, x3 = BF.pool_skip(self.convd3(p2, pooled=True))
            p4, x4 = BF.pool_skip(self.convd4(p3, pooled=True))
            x5 = self.convd5(p4, pooled=True)
            y4 = self.convu4(x5, x4, lazy_prev_grad=True)
            y3 = self.convu3(y4, x3, lazy_prev_grad=True)
            y2 = self.convu2(y3, x2, lazy_prev_grad=True)
            y1 = self.convu1(y2, x1, lazy_prev_grad=True)
        s3 = self.seg3(y3)
        s2 = self.seg2(y2) + self.upsample(s3)
        return self.seg1(y1) + self.upsample(s2)


# ------------------------------------------------------------------ third-party unet.UNet
class _FpBlock(tnn.Module):
    """conv -> [BatchNorm3d] -> PReLU, registered twice (named attributes and `block` Sequential) like the original,
    so the checkpoint's duplicated keys (`conv_layer.*` and `block.0.*`) both exist and share storage."""

    def __init__(self, cin, cout, k=3, norm=True, act=True):
        super().__init__()
        layers = []
        self.conv_layer = bnn.Conv3d(cin, cout, k, padding=(k + 1) // 2 - 1)
        layers.append(self.conv_layer)
        self.norm_layer = None
        if norm:
            self.norm_layer = bnn.BatchNorm3d(cout)
            layers.append(self.norm_layer)
        self.activation_layer = None
        if act:
            self.activation_layer = bnn.PReLU()
            layers.append(self.activation_layer)
        self.block = tnn.Sequential(*layers)

    def forward(self, x):
        return self.block(x)


class _FpStage(tnn.Module):
    def __init__(self, cin, c1, c2, first_norm=True):
        super().__init__()
        self.conv1 = _FpBlock(cin, c1, norm=first_norm)
        self.conv2 = _FpBlock(c1, c2)

    def forward(self, x):
        return self.conv2(self.conv1(x))


class FepegarUNet(tnn.Module):
    """`unet.UNet(in_channels=1, out_classes=2, dimensions=3, num_encoding_blocks=3, out_channels_first_layer=F,
    normalization='batch', upsampling_type='linear', padding=True, activation='PReLU')` -- segmentation/routine.py:346-356."""

    def __init__(self, in_channels=1, out_classes=2, num_encoding_blocks=3, out_channels_first_layer=16):
        super().__init__()
        F = out_channels_first_layer
        self.encoder = tnn.Module()
        self.encoder.encoding_blocks = tnn.ModuleList()
        cin, skips = in_channels, []
        for i in range(num_encoding_blocks - 1):
            c1 = F * 2 ** i
            self.encoder.encoding_blocks.append(_FpStage(cin, c1, 2 * c1, first_norm=(i > 0)))
            cin = 2 * c1
            skips.append(cin)
        self.bottom_block = _FpStage(cin, cin, 2 * cin)
        cin *= 2
        self.decoder = tnn.Module()
        self.decoder.decoding_blocks = tnn.ModuleList()
        for skip in reversed(skips):
            self.decoder.decoding_blocks.append(_FpStage(cin + skip, skip, skip))
            cin = skip
        self.classifier = _FpBlock(cin, out_classes, k=1, norm=False, act=False)
        self.pool = bnn.MaxPool3d(2)

    def forward(self, x):
        skips = []
        for stage in self.encoder.encoding_blocks:
            x = stage(x)
            skips.append(x)
            x = self.pool(x)
        x = self.bottom_block(x)
        for stage in self.decoder.decoding_blocks:
            x = BF.upsample_concat(skips.pop(), x, 2, "trilinear", False)     # cat((skip, up(x)), 1), one pass
            x = stage(x)
        return self.classifier(x)


# ------------------------------------------------------------------ AE family
def _act(name):
    return bnn.LeakyReLU() if name == "l_relu" else bnn.ReLU()


def _gain(name):
    return tnn.init.calculate_gain("leaky_relu", 0.01) if name == "l_relu" else tnn.init.calculate_gain("relu")


def _sep_convs(names, cin, cout, k, s, p):
    shapes = (((k, 1, 1), (s, 1, 1), (p, 0, 0)), ((1, k, 1), (1, s, 1), (0, p, 0)), ((1, 1, k), (1, 1, s), (0, 0, p)))
    chans = ((cin, cout), (cout, cout), (cout, cout))
    return {n: bnn.Conv3d(ci, co, kernel_size=ks, stride=st, padding=pd) for n, (ci, co), (ks, st, pd) in zip(names, chans, shapes)}


def _xavier(block, gain):
    for m in block.values():                                                  # AE_model.py:39-43
        if hasattr(m, "weight") and m.weight is not None and m.weight.dim() > 1:
            tnn.init.xavier_uniform_(m.weight.data, gain=gain)
            tnn.init.constant_(m.bias.data, 0)


class DownBlock(tnn.Module):
    """AE_model.py:4-53; modules run in sorted-key order (conv x/y/z, pool, BN, act)."""

    def __init__(self, c_in, c_out, skip=False, **kw):
        super().__init__()
        self.skip = skip
        self.block = tnn.ModuleDict(_sep_convs(("1_convx", "2_convy", "3_convz"), c_in, c_out, kw["conv_k"], kw["conv_s"], kw["conv_pad"]))
        self.block["4_pooling"] = bnn.MaxPool3d(kernel_size=kw["maxpool_k"], stride=kw["maxpool_s"])
        if kw["batch_norm"]:
            self.block["5_batch_norm"] = bnn.BatchNorm3d(c_out)
        self.block["6_act"] = _act(kw["act"])
        self.init_gain = _gain(kw["act"])
        _xavier(self.block, self.init_gain)

    def forward(self, x):
        before = tuple(x.shape[2:])
        for _, m in sorted(self.block.items()):
            x = m(x)
        return x, before


class UpBlock(tnn.Module):
    """AE_model.py:56-120."""

    def __init__(self, c_in, c_out, skip=False, **kw):
        super().__init__()
        self.skip = skip
        self.block = tnn.ModuleDict()
        if kw["up"] == "transpose_conv":
            self.block["1_upsample"] = bnn.ConvTranspose3d(c_in, c_out, kernel_size=kw["scale"], stride=kw["scale"], padding=kw["t_conv_pad"])
        else:
            self.block["1_upsample"] = bnn.Upsample(scale_factor=kw["scale"], mode=kw["scale_mode"])
        self.block.update(_sep_convs(("2_convx", "3_convy", "4_convz"), c_in, c_out, kw["conv_k"], kw["conv_s"], kw["conv_pad"]))
        if kw["batch_norm"]:
            self.block["5_batch_norm"] = bnn.BatchNorm3d(c_out)
        self.block["6_act"] = _act(kw["act"])
        self.init_gain = _gain(kw["act"])
        _xavier(self.block, self.init_gain)

    def forward(self, x, shape_before_pool=None, x_before_pool=None):
        for key, m in sorted(self.block.items()):
            x = m(x)
            if key == "1_upsample" and any(a > b for a, b in zip(shape_before_pool, x.shape[2:])):
                x = BF.interpolate(x, size=tuple(shape_before_pool))          # AE_model.py:116-119 (nearest)
        return x


class Encoder(tnn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.encode = tnn.ModuleList()
        if kw["reduce_size"]:
            self.encode.append(bnn.Conv3d(1, 1, kernel_size=4, stride=4, padding=0))
        for i in range(kw["deapth"]):
            self.encode.append(DownBlock(kw["chanels"][i], kw["chanels"][i + 1], kw["skip_map"][i], **kw["down_block_kwargs"]))

    def forward(self, x):
        sizes = []
        for m in self.encode:
            if isinstance(m, DownBlock):
                x, s = m(x)
            else:                                   # reduce_size stem; the reference would fail to unpack here
                s = tuple(x.shape[2:])
                x = m(x)
            sizes.append(s)
        return x, sizes


class Decoder(tnn.Module):
    def __init__(self, **kw):
        super().__init__()
        self.decode = tnn.ModuleList()
        for i in range(kw["deapth"]):
            self.decode.append(UpBlock(kw["chanels"][i], kw["chanels"][i + 1], kw["skip_map"][i], **kw["up_block_kwargs"]))
        if kw["reduce_size"]:
            self.decode.append(bnn.ConvTranspose3d(1, 1, kernel_size=4, stride=4, padding=0))
        self.vox = bnn.Conv3d(1, 1, kernel_size=3, stride=1, padding=1)

    def forward(self, x, size_list):
        size_list.reverse()
        for i, m in enumerate(self.decode):
            x = m(x, size_list[i]) if isinstance(m, UpBlock) else m(x)
        return self.vox(x)


class AE(tnn.Module):
    def __init__(self, **kw):
        super().__init__()
        depth = kw["deapth"]
        skip_map = kw["skip_map"] if kw["is_skip"] else [False] * depth
        chans = [kw["c_in"]] + [kw["c_base"] * kw["inc_size"] ** i for i in range(depth)]
        self.enc = Encoder(deapth=depth, chanels=chans, skip_map=skip_map, reduce_size=kw["reduce_size"],
                           down_block_kwargs=kw["down_block_kwargs"])
        self.dec = Decoder(deapth=depth, chanels=chans[::-1], skip_map=skip_map[::-1], reduce_size=kw["reduce_size"],
                           up_block_kwargs=kw["up_block_kwargs"])

    def forward(self, x):
        z, sizes = self.enc(x)
        return self.dec(z, sizes)


class _FaderHead(tnn.Module):
    _attr = "clf"
    _out_key = "n_class"

    def __init__(self, **kw):
        super().__init__()
        d = tnn.ModuleDict(_sep_convs(("1_convx", "2_convy", "3_convz"), kw["c_in"], kw["c_out"], kw["conv_k"], kw["conv_s"], kw["conv_pad"]))
        d["4_flat"] = tnn.Flatten()
        d["5_l1"] = tnn.Linear(kw["l_in"], kw["l_out"])
        if kw["batch_norm"]:
            d["6_batch_norm"] = tnn.BatchNorm1d(kw["l_out"])
        d["7_act"] = tnn.LeakyReLU() if kw["act"] == "l_relu" else tnn.ReLU()      # 2-D (B, l_out) tensors: stays PyTorch (K14)
        d["8_drop"] = tnn.Dropout(kw["p_drop"])
        d["9_l_f"] = tnn.Linear(kw["l_out"], kw[self._out_key])
        setattr(self, self._attr, d)
        self.init_gain = _gain(kw["act"])
        _xavier(d, self.init_gain)

    def forward(self, x):
        for key, m in sorted(getattr(self, self._attr).items()):
            x = m(x)
            if key == "3_convz":
                x = x.float()        # the fully-connected tail runs in fp32 PyTorch
        return x


class Classificator(_FaderHead):
    """AE_model.py:264-312"""
    _attr, _out_key = "clf", "n_class"


class Discriminator(_FaderHead):
    """AE_model.py:213-262"""
    _attr, _out_key = "disc", "n_domains"


# ------------------------------------------------------------------ detection PatchModel (2-D)
class ConvolutionBlock(tnn.Module):
    def __init__(self, in_c, out_c, pad=0):
        super().__init__()
        self.conv = bnn.Conv2d(in_c, out_c, kernel_size=3, padding=pad)
        self.bn = bnn.BatchNorm2d(out_c)
        self.relu = bnn.ReLU()

    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))


class PatchModel(tnn.Module):
    """detection/model_utils.py:19-42"""

    def __init__(self):
        super().__init__()
        widths = (2, 16, 32, 64, 128, 256)
        self.conv_blocks = tnn.Sequential(*[ConvolutionBlock(a, b) for a, b in zip(widths, widths[1:])], bnn.MaxPool2d(2))
        self.flatten = tnn.Flatten()
        self.dropout = tnn.Dropout(p=0.4)
        self.fc1 = tnn.Linear(3 * 11 * 256, 256)
        self.fc2 = tnn.Linear(256, 2)

    def forward(self, x):
        x = self.conv_blocks(x)
        x = self.flatten(x.float())          # logical NCHW flatten order, as the reference's fc1 expects
        x = self.dropout(x)
        return self.fc2(torch.relu(self.fc1(x)))


# kwargs used by the reference notebooks
FADER_DOWN = dict(conv_k=6, conv_pad=2, conv_s=2, maxpool_k=2, maxpool_s=2, batch_norm=True, act="l_relu")        # train_ENC_CLF.ipynb [cell 17]
FADER_UP = dict(up="upsample", scale=4, scale_mode="nearest", conv_k=3, conv_pad=1, conv_s=1, batch_norm=False, act="l_relu")
FADER_HEAD = dict(c_in=32, c_out=64, conv_k=3, conv_s=1, conv_pad=0, l_in=64, l_out=32, batch_norm=True, act="relu", p_drop=0.5)
AE_DOWN = dict(conv_k=3, conv_pad=1, conv_s=1, maxpool_k=2, maxpool_s=2, batch_norm=True, act="relu")              # train_AE.ipynb [cell 8]
AE_UP = dict(up="upsample", scale=2, scale_mode="nearest", conv_k=3, conv_pad=1, conv_s=1, batch_norm=True, act="relu")


def fader_encoder():
    return AE(c_in=1, is_skip=False, deapth=3, c_base=8, inc_size=2, reduce_size=False, down_block_kwargs=FADER_DOWN, up_block_kwargs=FADER_UP).enc


def config1_autoencoder(depth=6, c_base=16):
    return AE(c_in=1, is_skip=False, deapth=depth, c_base=c_base, inc_size=2, reduce_size=False, down_block_kwargs=AE_DOWN, up_block_kwargs=AE_UP)


# ------------------------------------------------------------------ fused execution of nn.Sequential chains
def _act_code(m):
    if isinstance(m, tnn.LeakyReLU):
        return cabi.ACT_LEAKY if abs(m.negative_slope - 0.01) < 1e-12 else None
    return cabi.ACT_RELU if isinstance(m, tnn.ReLU) else None


_NORMS = (bnn.BatchNorm3d, bnn.BatchNorm2d, bnn.InstanceNorm3d, bnn.GroupNorm)


def run_fused(layers, x):
    """Execute a chain of drop-in modules with the kernel-level fusions the library offers, WITHOUT changing the function:
    conv -> BatchNorm (training) takes the statistics from the conv epilogue, norm -> ReLU/LeakyReLU(0.01) is one apply pass.
    Modules that are not ours (Linear, Dropout, Flatten, ...) are called as they are, on a contiguous fp32 tensor once the
    data leaves the convolutional body."""
    layers = list(layers)
    i = 0
    while i < len(layers):
        m = layers[i]
        nxt = layers[i + 1] if i + 1 < len(layers) else None
        nxt2 = layers[i + 2] if i + 2 < len(layers) else None
        if isinstance(m, bnn._ConvMixin) and isinstance(nxt, _NORMS):
            act = _act_code(nxt2)
            x = _conv_norm(m, nxt, x, act=cabi.ACT_NONE if act is None else act)
            i += 2 if act is None else 3
        elif isinstance(m, _NORMS) and _act_code(nxt) is not None:
            x = m(x, act=_act_code(nxt))
            i += 2
        elif isinstance(m, (tnn.Linear, tnn.Flatten, tnn.BatchNorm1d)) and x.dim() > 2:
            x = m(x.float().contiguous())       # logical NCDHW order, as the reference's `view(N, -1)` expects
            i += 1
        else:
            x = m(x)
            i += 1
    return x


# ------------------------------------------------------------------ segmentation/models/modified_3dunet.py
class Modified3DUNet(tnn.Module):
    """segmentation/models/modified_3dunet.py:4-189 (Isensee-style U-Net: stride-2 context convolutions, InstanceNorm,
    LeakyReLU, nearest x2 localisation path, two deep-supervision heads).  Same constructor, attribute names and state_dict
    keys; the forward is the reference's operator sequence with norm+LeakyReLU pairs executed as one pass."""

    def __init__(self, in_channels, n_classes, base_n_filter=8):
        super().__init__()
        self.in_channels, self.n_classes, self.base_n_filter = in_channels, n_classes, base_n_filter
        f = base_n_filter
        c3 = lambda ci, co, s=1: bnn.Conv3d(ci, co, kernel_size=3, stride=s, padding=1, bias=False)
        c1 = lambda ci, co: bnn.Conv3d(ci, co, kernel_size=1, stride=1, padding=0, bias=False)
        self.lrelu = bnn.LeakyReLU()
        self.dropout3d = tnn.Dropout3d(p=0.6)
        self.upsacle = bnn.Upsample(scale_factor=2, mode="nearest")          # (sic) attribute name of the reference
        self.softmax = tnn.Softmax(dim=1)
        # context pathway, level 1 (:17-20)
        self.conv3d_c1_1, self.conv3d_c1_2 = c3(in_channels, f), c3(f, f)
        self.lrelu_conv_c1 = tnn.Sequential(bnn.LeakyReLU(), c3(f, f))
        self.inorm3d_c1 = bnn.InstanceNorm3d(f)
        # context pathway, levels 2..5 (:23-40): stride-2 conv, a (norm, lrelu, conv) block applied twice, a norm
        for lvl in range(2, 6):
            w = f * 2 ** (lvl - 1)
            setattr(self, f"conv3d_c{lvl}", c3(w // 2, w, 2))
            setattr(self, f"norm_lrelu_conv_c{lvl}", tnn.Sequential(bnn.InstanceNorm3d(w), bnn.LeakyReLU(), c3(w, w)))
            if lvl < 5:
                setattr(self, f"inorm3d_c{lvl}", bnn.InstanceNorm3d(w))
        up = lambda ci, co: tnn.Sequential(bnn.InstanceNorm3d(ci), bnn.LeakyReLU(), bnn.Upsample(scale_factor=2, mode="nearest"), c3(ci, co),
                                           bnn.InstanceNorm3d(co), bnn.LeakyReLU())
        cnl = lambda ci, co: tnn.Sequential(c3(ci, co), bnn.InstanceNorm3d(co), bnn.LeakyReLU())
        self.norm_lrelu_upscale_conv_norm_lrelu_l0 = up(16 * f, 8 * f)                    # :41
        self.conv3d_l0, self.inorm3d_l0 = c1(8 * f, 8 * f), bnn.InstanceNorm3d(8 * f)     # :43-44
        for lvl, w in ((1, 16 * f), (2, 8 * f), (3, 4 * f)):                             # :47-59
            setattr(self, f"conv_norm_lrelu_l{lvl}", cnl(w, w))
            setattr(self, f"conv3d_l{lvl}", c1(w, w // 2))
            setattr(self, f"norm_lrelu_upscale_conv_norm_lrelu_l{lvl}", up(w // 2, w // 4))
        self.conv_norm_lrelu_l4 = cnl(2 * f, 2 * f)                                       # :62
        self.conv3d_l4 = c1(2 * f, n_classes)
        self.ds2_1x1_conv3d, self.ds3_1x1_conv3d = c1(8 * f, n_classes), c1(4 * f, n_classes)

    def forward(self, x):
        out = self.conv3d_c1_1(x)                                                         # :103-112
        res = out
        out = self.conv3d_c1_2(self.lrelu(out))
        out = run_fused(self.lrelu_conv_c1, self.dropout3d(out))
        out = out + res
        context = [self.lrelu(out)]
        out = self.inorm3d_c1(out, act=cabi.ACT_LEAKY)
        for lvl in range(2, 6):                                                           # :115-157
            out = getattr(self, f"conv3d_c{lvl}")(out)
            res = out
            block = getattr(self, f"norm_lrelu_conv_c{lvl}")
            out = run_fused(block, self.dropout3d(run_fused(block, out)))
            out = out + res
            if lvl < 5:
                out = getattr(self, f"inorm3d_c{lvl}")(out, act=cabi.ACT_LEAKY)
                context.append(out)
        out = run_fused(self.norm_lrelu_upscale_conv_norm_lrelu_l0, out)                  # :158
        out = self.inorm3d_l0(self.conv3d_l0(out), act=cabi.ACT_LEAKY)                    # :160-162
        ds = {}
        for lvl in (1, 2, 3):                                                             # :165-183
            out = run_fused(getattr(self, f"conv_norm_lrelu_l{lvl}"), BF.concat(out, context[4 - lvl]))
            ds[lvl] = out
            out = run_fused(getattr(self, f"norm_lrelu_upscale_conv_norm_lrelu_l{lvl}"), getattr(self, f"conv3d_l{lvl}")(out))
        out = run_fused(self.conv_norm_lrelu_l4, BF.concat(out, context[0]))              # :186-188
        pred = self.conv3d_l4(out)
        s = self.upsacle(self.ds2_1x1_conv3d(ds[2])) + self.ds3_1x1_conv3d(ds[3])         # :190-194
        return pred + self.upsacle(s)                                                     # :196


# ------------------------------------------------------------------ classification/models/cnn_model.py
class BasicBlock(tnn.Module):
    """cnn_model.py:17-40: relu(bn2(conv2(relu(bn1(conv1(x))))) + x); the residual add and the final ReLU ride in bn2's apply pass."""

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = bnn.Conv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = bnn.BatchNorm3d(planes)
        self.relu = bnn.ReLU(inplace=True)
        self.conv2 = bnn.Conv3d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = bnn.BatchNorm3d(planes)
        self.stride = stride

    def forward(self, x):
        out = _conv_norm(self.conv1, self.bn1, x, act=cabi.ACT_RELU)
        return _conv_norm(self.conv2, self.bn2, out, act=cabi.ACT_RELU, residual=x)


class _SequentialModel(tnn.Module):
    def forward(self, x):
        return run_fused(self.model.children(), x)


class VoxResNet(_SequentialModel):
    """cnn_model.py:43-101.  The reference adds `activation_6` twice (:85 inside the n_blocks >= 4 stage and :95 after
    fully_conn_1); nn.Sequential.add_module keeps the FIRST position of a repeated name, so with n_blocks=4 there is no ReLU
    after fully_conn_1 -- reproduced here by issuing the same add_module sequence."""

    def __init__(self, input_shape=(128, 128, 128), num_classes=2, n_filters=32, stride=2, n_blocks=3, n_flatten_units=None, dropout=0, n_fc_units=128):
        super().__init__()
        n = n_filters
        seq = self.model = tnn.Sequential()
        add = seq.add_module
        add("conv3d_1", bnn.Conv3d(1, n, kernel_size=3, padding=1, stride=stride))
        add("batch_norm_1", bnn.BatchNorm3d(n)); add("activation_1", bnn.ReLU(inplace=True))
        add("conv3d_2", bnn.Conv3d(n, n, kernel_size=3, padding=1))
        add("batch_norm_2", bnn.BatchNorm3d(n)); add("activation_2", bnn.ReLU(inplace=True))
        widths = [(n, 2 * n), (2 * n, 2 * n), (2 * n, 4 * n), (4 * n, 4 * n)]
        for stage in range(1, 5):
            if stage > 1 and n_blocks < stage:
                break
            ci, co = widths[stage - 1]
            add(f"conv3d_{stage + 2}", bnn.Conv3d(ci, co, kernel_size=3, padding=1, stride=2))
            add(f"block_{2 * stage - 1}", BasicBlock(co, co)); add(f"block_{2 * stage}", BasicBlock(co, co))
            add(f"batch_norm_{stage + 2}", bnn.BatchNorm3d(co)); add(f"activation_{stage + 2}", bnn.ReLU(inplace=True))
        if n_flatten_units is None:
            import numpy as np
            n_flatten_units = int(4 * n * np.prod(np.array(input_shape) // (2 ** n_blocks * stride)))
        add("flatten_1", tnn.Flatten())
        add("fully_conn_1", tnn.Linear(n_flatten_units, n_fc_units))
        add("activation_6", tnn.ReLU(inplace=True))
        add("dropout_1", tnn.Dropout(dropout))
        add("fully_conn_2", tnn.Linear(n_fc_units, num_classes))


class CNN(_SequentialModel):
    """cnn_model.py:104-175: n_blocks x [conv-bn-relu, conv-bn-relu, MaxPool3d(2)], Flatten, Linear, BatchNorm1d, ReLU."""

    def __init__(self, input_shape=(64, 76, 48), n_filters=16, n_blocks=3, stride=1, n_fc_units=128):
        super().__init__()
        seq = self.model = tnn.Sequential()
        cin, idx = 1, 1
        for b in range(n_blocks):
            co = n_filters * 2 ** b
            for j in range(2):
                seq.add_module(f"conv3d_{idx}", bnn.Conv3d(cin, co, kernel_size=3, stride=stride if idx == 1 else 1, padding=1))
                seq.add_module(f"batch_norm_{idx}", bnn.BatchNorm3d(co))
                seq.add_module(f"activation_{idx}", bnn.ReLU(inplace=True))
                cin, idx = co, idx + 1
            seq.add_module(f"max_pool3d_{b + 1}", bnn.MaxPool3d(kernel_size=2))
        seq.add_module("flatten_1", tnn.Flatten())
        div = 2 ** n_blocks * stride
        seq.add_module("fully_conn_1", tnn.Linear(cin * (input_shape[0] // div) * (input_shape[1] // div) * (input_shape[2] // div), n_fc_units))
        seq.add_module("batch_norm_9", tnn.BatchNorm1d(n_fc_units))
        seq.add_module("activation_9", tnn.ReLU(inplace=True))


class DilatedCNN(_SequentialModel):
    """cnn_model.py:207-257: dilation-3 convolutions (stride 2 / no padding and stride 1 / padding 3), MaxPool3d(4, 2)."""

    def __init__(self, input_shape=(180, 180, 180), n_channels=32):
        super().__init__()
        c = n_channels
        seq = self.model = tnn.Sequential()
        #        (cin, cout, stride, padding, pool after)
        spec = [(1, c, 2, 0, False), (c, c, 1, 3, True), (c, 2 * c, 2, 0, False), (2 * c, 2 * c, 1, 3, True), (2 * c, 4 * c, 1, 3, False),
                (4 * c, 4 * c, 1, 0, False)]
        pools = 0
        for i, (ci, co, s, p, pool) in enumerate(spec, 1):
            seq.add_module(f"conv3d_{i}", bnn.Conv3d(ci, co, kernel_size=3, stride=s, dilation=3, padding=p))
            seq.add_module(f"batch_norm_{i}", bnn.BatchNorm3d(co))
            seq.add_module(f"activation_{i}", bnn.LeakyReLU())
            if pool:
                pools += 1
                seq.add_module(f"max_pool3d_{pools}", bnn.MaxPool3d(kernel_size=4, stride=2))
        seq.add_module("flatten_1", tnn.Flatten())
        seq.add_module("fully_conn_1", tnn.Linear(4 * c * ((input_shape[0] - 61) // 16 - 5) ** 3, 256))
        seq.add_module("activation_7", tnn.LeakyReLU())
        seq.add_module("fully_conn_2", tnn.Linear(256, 128))
        seq.add_module("activation_8", tnn.LeakyReLU())
        seq.add_module("fully_conn_3", tnn.Linear(128, 2))
        seq.add_module("softmax", tnn.Softmax(dim=-1))
