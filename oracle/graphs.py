"""Functional CPU restatement of the reference model graphs (TEST INFRASTRUCTURE).

Each graph is a plain function `f(sd, x, ...)` over a state dict `sd` whose keys and
shapes are exactly those of the corresponding reference nn.Module, so the shipped
.pth files and `module.state_dict()` of the real reference feed it directly.
Everything runs in fp32 on the CPU through torch.nn.functional, i.e. through the same
ATen CPU kernels the reference's own CPU path uses; autograd works through it, so
gradients w.r.t. `sd` entries are the reference's gradients.

Op pins: cross-correlation conv with zero padding, BN eps 1e-5 /
momentum 0.1 / biased var for normalisation and unbiased for running_var,
InstanceNorm affine=False, MaxPool floor mode, trilinear align_corners=False.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- bf16-storage mode
# The CUDA path keeps activations and activation gradients in HBM as bf16 and feeds the tensor-core convolutions bf16
# copies of the weights; all arithmetic in between is fp32.  `with bf16_storage(...)` makes the graphs below round at the
# same places -- every tensor an operator STORES (conv output, norm(+residual)+activation output, upsample output) and every
# gradient tensor an operator's backward stores -- while every operator still executes through the same fp32 ATen CPU kernels.
# It is the reference's algorithm with the storage precision of the device path, used to show that the distance between the
# bf16 CUDA results and the fp32 reference IS storage rounding (tests/test_gpu_models.py); with the mode off (default) the
# graphs are the plain fp32 reference path.
_STORE = {"on": False, "round_weight": None}


class bf16_storage:
    """round_weight(prefix, weight) -> bool: whether the convolution `prefix` consumes a bf16 copy of its weight (true for
    the layers the device runs on tensor cores; the Cin=1 stems, the few-channel heads and the separable convs read fp32)."""

    def __init__(self, round_weight=None):
        self.round_weight = round_weight or (lambda pfx, w: w.shape[0] >= 8 and w.shape[1] >= 8)

    def __enter__(self):
        self.prev = dict(_STORE)
        _STORE.update(on=True, round_weight=self.round_weight)
        return self

    def __exit__(self, *exc):
        _STORE.update(self.prev)
        return False


def _r(t):
    return t.to(torch.bfloat16).to(t.dtype)


# optional recording of every stored intermediate (diagnostics: tools/parity_report.py compares them layer by layer)
_TAPS = None


class record_taps:
    def __enter__(self):
        global _TAPS
        self.prev, _TAPS = _TAPS, {}
        return _TAPS

    def __exit__(self, *exc):
        global _TAPS
        _TAPS = self.prev
        return False


def _tap(name, t):
    if _TAPS is not None:
        _TAPS[name] = t.detach()
    return t


class _StoreFn(torch.autograd.Function):
    """value stored as bf16 (forward) / its gradient stored as bf16 (backward)"""

    @staticmethod
    def forward(ctx, t, fwd, bwd):
        ctx.bwd = bwd
        return _r(t) if fwd else t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return (_r(g) if ctx.bwd else g), None, None


def _q(t):
    """an operator's output as the device stores it (bf16 value; the gradient arriving here is a stored bf16 tensor too)"""
    return _StoreFn.apply(t, True, True) if _STORE["on"] else t


def _gq(t):
    """an operator's input: the gradient its backward writes is stored as bf16"""
    return _StoreFn.apply(t, False, True) if _STORE["on"] and t.requires_grad else t


def _wq(pfx, w):
    """bf16 copy of a weight for the tensor-core layers (the fp32 master weight receives the unrounded fp32 gradient)"""
    if _STORE["on"] and _STORE["round_weight"](pfx, w):
        return _StoreFn.apply(w, True, False)
    return w


# --------------------------------------------------------------------------- helpers
def _bn(sd, pfx, x, training):
    """nn.BatchNorm{1,2,3}d forward incl. running-stat side effects."""
    rm, rv = sd[pfx + ".running_mean"], sd[pfx + ".running_var"]
    if training:
        key = pfx + ".num_batches_tracked"
        if key in sd:
            sd[key] += 1
    return F.batch_norm(x, rm, rv, sd[pfx + ".weight"], sd[pfx + ".bias"],
                        training, BN_MOMENTUM, BN_EPS)


def _norm(sd, pfx, x, kind, training):
    """unet3d.py:8-17 `normalization(planes, norm)`."""
    x = _gq(x)          # (storage mode: the caller stores the result after the fused residual add / activation)
    if kind == "bn":
        return _bn(sd, pfx, x, training)
    if kind == "in":   # nn.InstanceNorm3d(planes): affine=False, no running stats
        return F.instance_norm(x, eps=BN_EPS)
    if kind == "gn":   # nn.GroupNorm(4, planes)
        return F.group_norm(x, 4, sd[pfx + ".weight"], sd[pfx + ".bias"], BN_EPS)
    raise ValueError(kind)


def _conv(sd, pfx, x, stride=1, padding=0, dilation=1, store=True):
    """store=False: the layer emits fp32 (the <= 4-channel segmentation heads)"""
    y = F.conv3d(_gq(x), _wq(pfx, sd[pfx + ".weight"]), sd.get(pfx + ".bias"), stride, padding, dilation)
    return _tap(pfx, _q(y) if store else y)


# ----------------------------------------------------------------- unet3d.Unet (a-1)
def _convd(sd, pfx, x, first, norm, dropout, training):
    """unet3d.py:39-47 ConvD.forward, including the dead conv2/bn2 branch (:43-45)
    whose only effects are BN running-stat updates and RNG consumption."""
    if not first:
        x = F.max_pool3d(x, 2, 2)                                       # :41
    x = _tap(pfx + ".bn1", _q(_norm(sd, pfx + ".bn1", _conv(sd, pfx + ".conv1", x, 1, 1), norm, training)))   # :42
    dead = F.relu(_norm(sd, pfx + ".bn2", _conv(sd, pfx + ".conv2", x, 1, 1), norm, training))  # :43
    if dropout > 0:
        dead = F.dropout3d(dead, dropout)                               # :44-45 (always "training")
    del dead
    y = _norm(sd, pfx + ".bn3", _conv(sd, pfx + ".conv3", x, 1, 1), norm, training)   # :46
    return _tap(pfx + ".bn3", _q(F.relu(_gq(x) + y)))                   # :47


def _up2(t):
    return F.interpolate(_gq(t), scale_factor=2, mode="trilinear", align_corners=False)


def _convu(sd, pfx, x, prev, first, norm, training, commute_up=False):
    """unet3d.py:68-79 ConvU.forward.  commute_up=True evaluates :73-74 as upsample(conv2(x)) instead of conv2(upsample(x)):
    conv2 is 1x1x1 and the interpolation is a per-channel convex combination of voxels, so both orders are the same function
    in exact arithmetic; it only matters to the bf16-storage mode, which must round where the device graph (zoo.ConvU) stores."""
    if not first:
        x = _tap(pfx + ".bn1", _q(F.relu(_norm(sd, pfx + ".bn1", _conv(sd, pfx + ".conv1", x, 1, 1), norm, training))))  # :71
    if commute_up:
        y = _q(_up2(_conv(sd, pfx + ".conv2", x, 1, 0)))
    else:
        y = _conv(sd, pfx + ".conv2", _q(_up2(x)), 1, 0)                                   # :73
    y = _tap(pfx + ".bn2", _q(F.relu(_norm(sd, pfx + ".bn2", y, norm, training))))         # :74
    y = torch.cat([_gq(prev), _gq(y)], 1)                                                  # :76
    return _tap(pfx + ".bn3", _q(F.relu(_norm(sd, pfx + ".bn3", _conv(sd, pfx + ".conv3", y, 1, 1), norm, training))))  # :77


def unet3d(sd, x, norm="bn", dropout=0.5, training=False, commute_up=False):
    """unet3d.py:110-126 Unet.forward.  `self.upsample` (:85, broken as written) is the
    upstream BraTS2017 nn.Upsample(scale_factor=2, trilinear, align_corners=False)."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="trilinear", align_corners=False)    # fp32 heads: never stored as bf16
    x1 = _convd(sd, "convd1", x, True, norm, dropout, training)
    x2 = _convd(sd, "convd2", x1, False, norm, dropout, training)
    x3 = _convd(sd, "convd3", x2, False, norm, dropout, training)
    x4 = _convd(sd, "convd4", x3, False, norm, dropout, training)
    x5 = _convd(sd, "convd5", x4, False, norm, dropout, training)
    y4 = _convu(sd, "convu4", x5, x4, True, norm, training, commute_up)
    y3 = _convu(sd, "convu3", y4, x3, False, norm, training, commute_up)
    y2 = _convu(sd, "convu2", y3, x2, False, norm, training, commute_up)
    y1 = _convu(sd, "convu1", y2, x1, False, norm, training, commute_up)
    s3 = _conv(sd, "seg3", y3, store=False)                              # :122
    s2 = _conv(sd, "seg2", y2, store=False) + up(s3)                     # :123
    return _conv(sd, "seg1", y1, store=False) + up(s2)                   # :124


# ------------------------------------------- third-party unet.UNet (a-2, restated)
def _fp_block(sd, pfx, x, training):
    """One fepegar ConvolutionalBlock: conv(3^3, pad 1, bias) -> [BatchNorm3d] -> PReLU.
    The norm is absent on the very first conv (no `norm_layer.*` keys in the checkpoint)."""
    x = _conv(sd, pfx + ".conv_layer", x, 1, 1)
    if pfx + ".norm_layer.weight" in sd:
        x = _bn(sd, pfx + ".norm_layer", x, training)
        if training and pfx + ".block.1.num_batches_tracked" in sd and \
                sd[pfx + ".block.1.num_batches_tracked"] is not sd[pfx + ".norm_layer.num_batches_tracked"]:
            sd[pfx + ".block.1.num_batches_tracked"] += 1    # same module registered twice
    return F.prelu(x, sd[pfx + ".activation_layer.weight"])


def fepegar_unet(sd, x, training=False, num_encoding_blocks=3):
    """`unet.UNet(in=1, out_classes=2, dimensions=3, num_encoding_blocks=3, normalization='batch',
    upsampling_type='linear', padding=True, activation='PReLU')` -- call site
    segmentation/routine.py:346-356.  Encoder has num_encoding_blocks-1 pooled blocks,
    then the bottom block, then as many decoding blocks; skips are the pre-pool tensors;
    decoder concatenates (skip, upsampled) and upsamples trilinearly, align_corners=False."""
    skips = []
    for i in range(num_encoding_blocks - 1):
        p = f"encoder.encoding_blocks.{i}"
        x = _fp_block(sd, p + ".conv1", x, training)
        x = _fp_block(sd, p + ".conv2", x, training)
        skips.append(x)
        x = F.max_pool3d(x, 2)
    x = _fp_block(sd, "bottom_block.conv1", x, training)
    x = _fp_block(sd, "bottom_block.conv2", x, training)
    for i in range(num_encoding_blocks - 1):
        p = f"decoder.decoding_blocks.{i}"
        x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)
        x = torch.cat((skips[-1 - i], x), dim=1)
        x = _fp_block(sd, p + ".conv1", x, training)
        x = _fp_block(sd, p + ".conv2", x, training)
    return _conv(sd, "classifier.conv_layer", x)


# ------------------------------------------------ AE family (a-3, a-4) AE_model.py
def _act(x, act):
    # AE_model.py:31-36: 'l_relu' -> nn.LeakyReLU() (slope 0.01), anything else -> nn.ReLU()
    return F.leaky_relu(x, 0.01) if act == "l_relu" else F.relu(x)


def _sep3(sd, pfx, names, x, k, s, p):
    """The three separable convs (k,1,1) (1,k,1) (1,1,k) -- AE_model.py:9-26."""
    x = _conv(sd, f"{pfx}.{names[0]}", x, (s, 1, 1), (p, 0, 0))
    x = _conv(sd, f"{pfx}.{names[1]}", x, (1, s, 1), (0, p, 0))
    return _conv(sd, f"{pfx}.{names[2]}", x, (1, 1, s), (0, 0, p))


def down_block(sd, pfx, x, kw, training):
    """AE_model.py:45-53: modules run in sorted-key order: convx, convy, convz,
    MaxPool3d, BatchNorm3d (after the pool), activation."""
    shape_before = tuple(x.shape[2:])
    x = _sep3(sd, pfx + ".block", ("1_convx", "2_convy", "3_convz"), x, kw["conv_k"], kw["conv_s"], kw["conv_pad"])
    x = F.max_pool3d(x, kw["maxpool_k"], kw["maxpool_s"])
    if kw["batch_norm"]:
        x = _bn(sd, pfx + ".block.5_batch_norm", x, training)
    return _act(x, kw["act"]), shape_before


def up_block(sd, pfx, x, shape_before, kw, training):
    """AE_model.py:109-120 (upsample variant; `up='transpose_conv'` uses ConvTranspose3d :62-68)."""
    if kw["up"] == "transpose_conv":
        x = F.conv_transpose3d(x, sd[pfx + ".block.1_upsample.weight"], sd[pfx + ".block.1_upsample.bias"],
                               kw["scale"], kw["t_conv_pad"])
    else:
        x = F.interpolate(x, scale_factor=kw["scale"], mode=kw["scale_mode"])
    if any(a > b for a, b in zip(shape_before, x.shape[2:])):          # :116-119 odd sizes
        x = F.interpolate(x, tuple(shape_before))
    x = _sep3(sd, pfx + ".block", ("2_convx", "3_convy", "4_convz"), x, kw["conv_k"], kw["conv_s"], kw["conv_pad"])
    if kw["batch_norm"]:
        x = _bn(sd, pfx + ".block.5_batch_norm", x, training)
    return _act(x, kw["act"])


def encoder(sd, x, depth, down_kw, training=False, reduce_size=False, pfx="encode"):
    """AE_model.py:137-144 Encoder.forward -> (latent, size_list)."""
    sizes, i0 = [], 0
    if reduce_size:                                                       # :127-128
        x = _conv(sd, f"{pfx}.0", x, 4, 0)
        sizes.append(None)
        i0 = 1
    for i in range(depth):
        x, s = down_block(sd, f"{pfx}.{i + i0}", x, down_kw, training)
        sizes.append(s)
    return x, sizes


def autoencoder(sd, x, depth, down_kw, up_kw, training=False):
    """AE_model.py:207-210 AE.forward (reduce_size=False as in train_AE.ipynb [cell 8])."""
    z, sizes = encoder(sd, x, depth, down_kw, training, pfx="enc.encode")
    sizes = sizes[::-1]                                                   # :166
    for i in range(depth):
        z = up_block(sd, f"dec.decode.{i}", z, sizes[i], up_kw, training)
    return _conv(sd, "dec.vox", z, 1, 1)                                  # :160-164,169


def fader_head(sd, x, kw, training=False, pfx="clf"):
    """AE_model.py:258-262 / :308-312 Discriminator / Classificator forward (sorted keys):
    convx, convy, convz, Flatten, Linear, [BatchNorm1d], act, Dropout, Linear."""
    x = _sep3(sd, pfx, ("1_convx", "2_convy", "3_convz"), x, kw["conv_k"], kw["conv_s"], kw["conv_pad"])
    x = F.linear(x.flatten(1), sd[pfx + ".5_l1.weight"], sd[pfx + ".5_l1.bias"])
    if kw["batch_norm"]:
        x = _bn(sd, pfx + ".6_batch_norm", x, training)
    x = _act(x, kw["act"])
    x = F.dropout(x, kw.get("p_drop", 0.5), training)
    return F.linear(x, sd[pfx + ".9_l_f.weight"], sd[pfx + ".9_l_f.bias"])


# --------------------------------------------- detection PatchModel (a-6), 2-D
def patch_model(sd, x, training=False):
    """detection/model_utils.py:19-52: 5x [Conv2d 3x3 p0 -> BatchNorm2d -> ReLU], MaxPool2d(2),
    Flatten, Dropout(0.4), Linear 8448->256, ReLU, Linear 256->2."""
    for i in range(5):
        p = f"conv_blocks.{i}"
        x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"])
        x = F.relu(_bn(sd, p + ".bn", x, training))
    x = F.max_pool2d(x, 2).flatten(1)
    x = F.dropout(x, 0.4, training)
    x = torch.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    return F.linear(x, sd["fc2.weight"], sd["fc2.bias"])


# ------------------------------------- segmentation/models/modified_3dunet.py (a-1 family)
def _in_lrelu(x):
    return _q(F.leaky_relu(F.instance_norm(_gq(x), eps=BN_EPS), 0.01))


def modified_3dunet(sd, x, training=False, p_drop=0.6):
    """modified_3dunet.py:101-196 Modified3DUNet.forward.  nn.InstanceNorm3d(affine=False) and nn.LeakyReLU() defaults;
    every conv is bias-free; `self.dropout3d` (p=0.6) acts only in training mode."""
    c = lambda pfx, t, stride=1, pad=1, store=True: _conv(sd, pfx, t, stride, pad, 1, store)
    drop = lambda t: F.dropout3d(t, p_drop, True) if (training and p_drop > 0) else t
    up = lambda t: _q(F.interpolate(_gq(t), scale_factor=2, mode="nearest"))
    out = c("conv3d_c1_1", x)                                                 # :103
    res = out
    out = c("conv3d_c1_2", _q(F.leaky_relu(_gq(out), 0.01)))                  # :105-106
    out = c("lrelu_conv_c1.1", _q(F.leaky_relu(_gq(drop(out)), 0.01)))        # :107-108
    out = _q(_gq(out) + _gq(res))                                             # :110
    ctx = [_q(F.leaky_relu(_gq(out), 0.01))]                                  # :111
    out = _in_lrelu(out)                                                      # :112-113
    for lvl in range(2, 6):                                                   # :116-157
        out = c(f"conv3d_c{lvl}", out, 2)
        res = out
        blk = f"norm_lrelu_conv_c{lvl}.2"
        out = c(blk, _in_lrelu(out))
        out = c(blk, _in_lrelu(drop(out)))
        out = _q(_gq(out) + _gq(res))
        if lvl < 5:
            out = _in_lrelu(out)
            ctx.append(out)

    def up_block(pfx, t):                                                     # norm, lrelu, nearest x2, conv, norm, lrelu (:91-99)
        return _in_lrelu(c(pfx + ".3", up(_in_lrelu(t))))
    out = up_block("norm_lrelu_upscale_conv_norm_lrelu_l0", out)              # :158
    out = _in_lrelu(c("conv3d_l0", out, 1, 0))                                # :160-162
    ds = {}
    for lvl in (1, 2, 3):                                                     # :165-183
        out = _in_lrelu(c(f"conv_norm_lrelu_l{lvl}.0", torch.cat([_gq(out), _gq(ctx[4 - lvl])], 1)))
        ds[lvl] = out
        out = up_block(f"norm_lrelu_upscale_conv_norm_lrelu_l{lvl}", c(f"conv3d_l{lvl}", out, 1, 0))
    out = _in_lrelu(c("conv_norm_lrelu_l4.0", torch.cat([_gq(out), _gq(ctx[0])], 1)))      # :186-187
    pred = c("conv3d_l4", out, 1, 0, store=False)                             # :188
    up32 = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")         # fp32 heads
    s = up32(c("ds2_1x1_conv3d", ds[2], 1, 0, store=False)) + c("ds3_1x1_conv3d", ds[3], 1, 0, store=False)   # :190-194
    return pred + up32(s)                                                     # :196


# ------------------------------------------- classification/models/cnn_model.py (a-5)
def _cbr(sd, conv, bn, x, training, stride=1, padding=1, dilation=1, act="relu", residual=None):
    y = _bn(sd, bn, _gq(_conv(sd, conv, x, stride, padding, dilation)), training)
    if residual is not None:
        y = y + _gq(residual)
    return _q(F.leaky_relu(y, 0.01) if act == "l_relu" else F.relu(y) if act == "relu" else y)


def _basic_block(sd, pfx, x, training):
    """cnn_model.py:27-40"""
    out = _cbr(sd, pfx + ".conv1", pfx + ".bn1", x, training)
    return _cbr(sd, pfx + ".conv2", pfx + ".bn2", out, training, residual=x)


def voxresnet(sd, x, n_blocks=3, stride=2, training=False, dropout=0.0):
    """cnn_model.py:43-101 VoxResNet.forward = the nn.Sequential in registration order.  `activation_6` is registered twice
    (:85, :95): for n_blocks >= 4 the name keeps its first position, so NO ReLU follows fully_conn_1 in that case."""
    m = "model."
    x = _cbr(sd, m + "conv3d_1", m + "batch_norm_1", x, training, stride)
    x = _cbr(sd, m + "conv3d_2", m + "batch_norm_2", x, training)
    for stage in range(1, min(n_blocks, 4) + 1 if n_blocks >= 1 else 2):
        x = _conv(sd, f"{m}conv3d_{stage + 2}", x, 2, 1)
        x = _basic_block(sd, f"{m}block_{2 * stage - 1}", x, training)
        x = _basic_block(sd, f"{m}block_{2 * stage}", x, training)
        x = _q(F.relu(_bn(sd, f"{m}batch_norm_{stage + 2}", _gq(x), training)))
    x = F.linear(x.flatten(1), sd[m + "fully_conn_1.weight"], sd[m + "fully_conn_1.bias"])
    if n_blocks < 4:
        x = F.relu(x)
    x = F.dropout(x, dropout, training)
    return F.linear(x, sd[m + "fully_conn_2.weight"], sd[m + "fully_conn_2.bias"])


def cnn(sd, x, n_blocks=3, stride=1, training=False):
    """cnn_model.py:104-175 CNN.forward."""
    m, idx = "model.", 1
    for b in range(n_blocks):
        for _ in range(2):
            x = _cbr(sd, f"{m}conv3d_{idx}", f"{m}batch_norm_{idx}", x, training, stride if idx == 1 else 1)
            idx += 1
        x = F.max_pool3d(x, 2)
    x = F.linear(x.flatten(1), sd[m + "fully_conn_1.weight"], sd[m + "fully_conn_1.bias"])
    return F.relu(_bn(sd, m + "batch_norm_9", x, training))


def dilated_cnn(sd, x, training=False):
    """cnn_model.py:207-257 DilatedCNN.forward (dilation 3 everywhere; softmax output)."""
    m = "model."
    spec = [(2, 0, False), (1, 3, True), (2, 0, False), (1, 3, True), (1, 3, False), (1, 0, False)]
    for i, (s, p, pool) in enumerate(spec, 1):
        x = _cbr(sd, f"{m}conv3d_{i}", f"{m}batch_norm_{i}", x, training, s, p, 3, act="l_relu")
        if pool:
            x = F.max_pool3d(x, 4, 2)
    x = x.flatten(1)
    for j in (1, 2):
        x = F.leaky_relu(F.linear(x, sd[f"{m}fully_conn_{j}.weight"], sd[f"{m}fully_conn_{j}.bias"]), 0.01)
    return F.softmax(F.linear(x, sd[m + "fully_conn_3.weight"], sd[m + "fully_conn_3.bias"]), dim=-1)


# ------------------------------------------------------- losses on the device path
def dice_loss_mean(logits, targets, eps=1e-9):
    """segmentation/routine.py:272-274 with get_dice_score :239-250.  Note the broadcast:
    probabilities (B,2,...) against targets (B,1,...) -> per-(batch,channel) dice, then mean."""
    p0 = F.softmax(logits, dim=1)
    g0 = targets
    p1, g1 = 1 - p0, 1 - g0
    tp = (p0 * g0).sum(dim=(2, 3, 4))
    fp = (p0 * g1).sum(dim=(2, 3, 4))
    fn = (p1 * g0).sum(dim=(2, 3, 4))
    return (1 - 2 * tp / (2 * tp + fp + fn + eps)).mean()


def adv_loss(domain, pred_logits, n_domains):
    """classification/train_ENC_CLF.ipynb [cell 14]: -mean((1-onehot) * log_softmax)."""
    onehot = torch.zeros((domain.shape[0], n_domains), dtype=torch.int32)
    onehot.scatter_(1, domain.view(-1, 1), 1)
    return -torch.mean((1 - onehot) * F.log_softmax(pred_logits, dim=1))


# kwargs of the shipped fader checkpoints -- classification/train_ENC_CLF.ipynb [cell 17]
FADER_DOWN = dict(conv_k=6, conv_pad=2, conv_s=2, maxpool_k=2, maxpool_s=2, batch_norm=True, act="l_relu")
FADER_HEAD = dict(conv_k=3, conv_s=1, conv_pad=0, batch_norm=True, act="relu", p_drop=0.5)
# kwargs of config 1 -- classification/train_AE.ipynb [cell 8]
AE_DOWN = dict(conv_k=3, conv_pad=1, conv_s=1, maxpool_k=2, maxpool_s=2, batch_norm=True, act="relu")
AE_UP = dict(up="upsample", scale=2, scale_mode="nearest", conv_k=3, conv_pad=1, conv_s=1, batch_norm=True, act="relu")
