"""Drop-in torch.nn modules for the reference's operator boundary.

The reference binds its operators at import with `import torch.nn as nn`
(segmentation/models/unet3d.py:1, classification/models/AE_model.py:2, ...), so the
boundary is the torch.nn class set.  Every class here SUBCLASSES the torch.nn class it
replaces and overrides only `forward`: constructor signatures, parameter/buffer names and
shapes, `state_dict()` keys and `isinstance` checks (unet3d.py:103-108, AE_model.py:39-43) are
therefore identical, and the shipped .pth files load with strict=True.

Three ways in:
  * `convert(model, dtype=torch.bfloat16)` re-classes the torch.nn instances of an already built
    reference model in place (parameters untouched) and routes the `F.*` calls made inside the
    reference's forward()s (unet3d.py:73, AE_model.py:119) through a scoped functional proxy;
  * `with patch():` swaps the symbols on torch.nn so an unmodified reference file picks them up
    while it constructs its model;
  * use the classes directly.

Activations flow as channels-last (logical NCDHW) tensors of `compute_dtype`; parameters stay
fp32 in PyTorch layout, so torch.optim and torch.save work unchanged.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch
import torch.nn as tnn
import torch.nn.functional as TF

from . import _cabi as cabi
from . import functional as BF

__all__ = ["Conv3d", "Conv2d", "ConvTranspose3d", "BatchNorm3d", "BatchNorm2d", "InstanceNorm3d", "GroupNorm", "ReLU", "LeakyReLU",
           "PReLU", "MaxPool3d", "MaxPool2d", "Upsample", "convert", "patch", "functional_proxy"]


class _B200Mixin:
    """Per-module knobs set by convert(); class attributes are the defaults for direct construction."""
    compute_dtype = None          # None: follow the input tensor's dtype
    out_dtype = None              # convs only: override the output dtype (fp32 logits from a bf16 body)
    allow_umma = True             # False: force the fp32-FMA kernels (tf32-off mode)
    sync = None                   # (process_group, world_size) for SyncBN

    def _prep(self, x):
        BF.need_cuda(x, type(self).__name__)
        dt = self.compute_dtype
        if dt is not None and x.dtype != dt and x.is_floating_point():
            # a stem conv reads the loader's fp32 batch directly (no extra cast pass); everything else is cast
            if not (isinstance(self, tnn.modules.conv._ConvNd) and x.dtype == torch.float32 and x.shape[1] < 8):
                x = x.to(dt)
        return x


class _ConvMixin(_B200Mixin):
    _transposed_conv = False

    def _cfg(self):
        cfg = self.__dict__.get("_b200_cfg")
        if cfg is None or cfg.allow_umma != self.allow_umma:
            cfg = BF.ConvConfig(self.stride, self.padding, self.dilation, self._transposed_conv, self.allow_umma)
            self.__dict__["_b200_cfg"] = cfg
        return cfg

    def forward(self, x, output_size=None, want_stats=False):
        """want_stats=True (fused graphs only): returns (y, partial) -- see functional.conv."""
        if self.groups != 1:
            raise RuntimeError("b200nn convolutions support groups=1 (all the reference uses)")
        if self.padding_mode != "zeros" or isinstance(self.padding, str):
            raise RuntimeError("b200nn convolutions support explicit zero padding only")
        if self._transposed_conv and (output_size is not None or any(self.output_padding)):
            raise RuntimeError("b200nn.ConvTranspose3d: output_padding/output_size are not supported")
        x = self._prep(x)
        out_dtype = self.out_dtype or (self.compute_dtype if self.compute_dtype is not None else x.dtype)
        weight = self.weight
        pad_c = self._tensor_core_channel_pad(x)
        if pad_c:
            # e.g. unet.UNet(first=8)'s 8 -> 16 convolution: 8 zero input channels (and zero weight columns) make the layer a
            # 16-channel tcgen05 problem instead of a CUDA-core one; autograd slices both gradients back
            x = BF.pad_channels(x, pad_c)
            weight = torch.cat([weight, weight.new_zeros((weight.shape[0], pad_c) + tuple(weight.shape[2:]))], 1)
        return BF.conv(x, weight, self.bias, self._cfg(), out_dtype, want_stats)

    def _tensor_core_channel_pad(self, x):
        ci = x.shape[1]
        if not self.allow_umma or self._transposed_conv or x.dtype != torch.bfloat16 or ci < 8 or ci % 16 == 0 or ci % 8 != 0:
            return 0
        if self.out_channels % 16 != 0 or x.numel() // ci < 4096:
            return 0
        k, s, p, d = self.kernel_size, self.stride, self.padding, self.dilation
        if any(v != 1 for v in s) or any(v != 1 for v in d) or any(kk not in (1, 3) for kk in k) or any(pp != kk // 2 for pp, kk in zip(p, k)):
            return 0
        return 16 - ci % 16


class Conv3d(_ConvMixin, tnn.Conv3d):
    pass


class Conv2d(_ConvMixin, tnn.Conv2d):
    pass


class ConvTranspose3d(_ConvMixin, tnn.ConvTranspose3d):
    _transposed_conv = True


# When a list, training-mode BatchNorm layers append their `num_batches_tracked` to it instead of launching one add each;
# the owner (graphed.GraphedTrainStep) bumps them all with a single multi-tensor add after the forward pass.
_deferred_counters = None


class defer_batch_counters:
    def __enter__(self):
        global _deferred_counters
        self._prev = _deferred_counters
        _deferred_counters = self.counters = []
        return self

    def __exit__(self, *exc):
        global _deferred_counters
        _deferred_counters = self._prev
        if self.counters and exc[0] is None:
            torch._foreach_add_(self.counters, 1)
        return False


class _BatchNormMixin(_B200Mixin):
    fused_act = cabi.ACT_NONE     # set by fused graphs; plain drop-in use keeps NONE
    fused_slope = 0.01

    def forward(self, x, residual=None, act=None, stats_partial=None, stats_only=False):
        """Plain call = torch.nn.BatchNorm semantics.  Fused graphs (zoo) may pass `residual` / `act` (B200_ACT_*) to fold the
        residual add and ReLU into the apply pass, `stats_partial` from a conv epilogue, and `stats_only=True` when the
        normalised output is never used (only the running-statistics side effect is performed)."""
        x = self._prep(x)
        self._check_input_dim(x)
        # torch.nn.modules.batchnorm._BatchNorm.forward semantics
        if self.momentum is None:
            factor = 0.0
        else:
            factor = self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            if _deferred_counters is not None and self.momentum is not None:
                _deferred_counters.append(self.num_batches_tracked)      # bumped by one _foreach_add_ at the end of the step
            else:
                self.num_batches_tracked.add_(1)
            if self.momentum is None:
                factor = 1.0 / float(self.num_batches_tracked)
        use_batch = self.training or (self.running_mean is None and self.running_var is None)
        rm = self.running_mean if (not self.training or self.track_running_stats) else None
        rv = self.running_var if (not self.training or self.track_running_stats) else None
        sync = self.sync if self.training else None
        if stats_only:
            if use_batch and rm is not None and (sync is None or stats_partial is not None):
                BF.batchnorm_update_running(x, rm, rv, factor, self.eps, stats_partial, sync=sync)
                return None
            if not use_batch or rm is None:
                return None             # eval mode / no running statistics: a discarded BatchNorm output has no side effect at all
        return BF.norm(x, self.weight, self.bias, kind=cabi.NORM_BATCH, running_mean=rm, running_var=rv, use_batch_stats=use_batch,
                       momentum=factor, eps=self.eps, act=self.fused_act if act is None else act, slope=self.fused_slope, residual=residual,
                       sync=sync, stats_partial=stats_partial)


class BatchNorm3d(_BatchNormMixin, tnn.BatchNorm3d):
    pass


class BatchNorm2d(_BatchNormMixin, tnn.BatchNorm2d):
    pass


class InstanceNorm3d(_B200Mixin, tnn.InstanceNorm3d):
    def forward(self, x, residual=None, act=None, stats_partial=None, stats_only=False):
        if stats_only:
            return None                 # no running statistics: a discarded output has no side effect
        x = self._prep(x)
        if self.track_running_stats:
            raise RuntimeError("b200nn.InstanceNorm3d: track_running_stats=True is not supported (the reference never sets it)")
        return BF.norm(x, self.weight, self.bias, kind=cabi.NORM_INSTANCE, use_batch_stats=True, eps=self.eps,
                       act=cabi.ACT_NONE if act is None else act, residual=residual)


class GroupNorm(_B200Mixin, tnn.GroupNorm):
    def forward(self, x, residual=None, act=None, stats_partial=None, stats_only=False):
        if stats_only:
            return None
        x = self._prep(x)
        return BF.norm(x, self.weight, self.bias, kind=cabi.NORM_GROUP, groups=self.num_groups, use_batch_stats=True, eps=self.eps,
                       act=cabi.ACT_NONE if act is None else act, residual=residual)


class ReLU(_B200Mixin, tnn.ReLU):
    def forward(self, x):
        return BF.relu(x, self.inplace)


class LeakyReLU(_B200Mixin, tnn.LeakyReLU):
    def forward(self, x):
        return BF.leaky_relu(x, self.negative_slope, self.inplace)


class PReLU(_B200Mixin, tnn.PReLU):
    def forward(self, x):
        return BF.prelu(self._prep(x), self.weight)


class _MaxPoolMixin(_B200Mixin):
    def forward(self, x):
        pad = self.padding if isinstance(self.padding, int) else max(self.padding)
        dil = self.dilation if isinstance(self.dilation, int) else max(self.dilation)
        if pad != 0 or dil != 1 or self.ceil_mode:
            raise RuntimeError("b200nn.MaxPool supports padding=0, dilation=1, ceil_mode=False (all the reference uses)")
        return BF.max_pool(self._prep(x), self.kernel_size, self.stride, self.return_indices)


class MaxPool3d(_MaxPoolMixin, tnn.MaxPool3d):
    pass


class MaxPool2d(_MaxPoolMixin, tnn.MaxPool2d):
    pass


class Upsample(_B200Mixin, tnn.Upsample):
    def forward(self, x):
        # fp32 tensors with <= 4 channels are the fp32 segmentation logits of the deep-supervision heads (unet3d.py:122-124):
        # they stay fp32 through the interpolation (cf. `fp32_heads` in convert) instead of being rounded to the body's dtype
        if not (x.dtype == torch.float32 and x.dim() in (4, 5) and x.shape[1] <= 4):
            x = self._prep(x)
        else:
            BF.need_cuda(x, "Upsample")
        return BF.interpolate(x, self.size, self.scale_factor, self.mode, self.align_corners)


# torch.nn class -> replacement (exact type match only: subclasses defined by user code are left alone)
_SWAP = {tnn.Conv3d: Conv3d, tnn.Conv2d: Conv2d, tnn.ConvTranspose3d: ConvTranspose3d, tnn.BatchNorm3d: BatchNorm3d,
         tnn.BatchNorm2d: BatchNorm2d, tnn.InstanceNorm3d: InstanceNorm3d, tnn.GroupNorm: GroupNorm, tnn.ReLU: ReLU,
         tnn.LeakyReLU: LeakyReLU, tnn.PReLU: PReLU, tnn.MaxPool3d: MaxPool3d, tnn.MaxPool2d: MaxPool2d, tnn.Upsample: Upsample}


# --------------------------------------------------------------------------- functional proxy
# torch's own functions, captured before patch() can rebind them; used only for tensors that are not on a GPU
# (such tensors are outside this library's path -- e.g. CPU-side metric code sharing the patched namespace).
_T_INTERPOLATE, _T_RELU, _T_LEAKY, _T_POOL3D = TF.interpolate, TF.relu, TF.leaky_relu, TF.max_pool3d


def _f_interpolate(input, size=None, scale_factor=None, mode="nearest", align_corners=None, **kw):
    if isinstance(input, torch.Tensor) and input.is_cuda and input.dim() in (4, 5) and input.dtype in (torch.float32, torch.bfloat16):
        return BF.interpolate(input, size, scale_factor, mode, align_corners)
    return _T_INTERPOLATE(input, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners, **kw)


def _f_relu(input, inplace=False):
    return BF.relu(input, inplace) if input.is_cuda else _T_RELU(input, inplace)


def _f_leaky_relu(input, negative_slope=0.01, inplace=False):
    return BF.leaky_relu(input, negative_slope, inplace) if input.is_cuda else _T_LEAKY(input, negative_slope, inplace)


def _f_max_pool3d(input, kernel_size, stride=None, padding=0, dilation=1, ceil_mode=False, return_indices=False):
    if input.is_cuda and padding == 0 and dilation == 1 and not ceil_mode:
        return BF.max_pool(input, kernel_size, stride, return_indices)
    return _T_POOL3D(input, kernel_size, stride, padding, dilation, ceil_mode, return_indices)


class _FunctionalProxy(types.ModuleType):
    """Stands in for `torch.nn.functional` inside ONE reference module's namespace: interpolate / upsample /
    relu / leaky_relu / max_pool3d go to the sm_100a kernels for CUDA tensors, everything else is torch's."""

    _OVERRIDES = {"interpolate": _f_interpolate, "upsample": _f_interpolate, "relu": _f_relu, "leaky_relu": _f_leaky_relu,
                  "max_pool3d": _f_max_pool3d}

    def __getattr__(self, name):
        ov = _FunctionalProxy._OVERRIDES.get(name)
        return ov if ov is not None else getattr(TF, name)


functional_proxy = _FunctionalProxy("mri_epilepsy_diagnosis_b200.functional_proxy")


# --------------------------------------------------------------------------- convert / patch
def _leave_body_hook(module, args):
    """Forward pre-hook for modules that are NOT ours but receive the convolutional body's output (Linear, Flatten -- incl.
    user-defined ones doing `input.view(N, -1)`, cnn_model.py:8-10 -- and BatchNorm1d): hand them what stock PyTorch would have
    produced, a contiguous (logical NCDHW order) fp32 tensor instead of a channels-last bf16 one."""
    out = []
    for a in args:
        if isinstance(a, torch.Tensor) and a.is_cuda and a.is_floating_point():
            if a.dim() in (4, 5) and not a.is_contiguous():
                a = a.contiguous()
            if a.dtype == torch.bfloat16:
                a = a.float()
        out.append(a)
    return tuple(out)


def _is_foreign_sink(m):
    if isinstance(m, _B200Mixin) or next(m.children(), None) is not None:
        return False
    if isinstance(m, (tnn.Linear, tnn.Flatten, tnn.BatchNorm1d)):
        return True
    return not type(m).__module__.startswith(("torch.", "mri_epilepsy_diagnosis_b200"))      # user-defined leaf (e.g. a Flatten)


def convert(model, dtype=torch.bfloat16, allow_umma=True, fp32_heads=True, sync=None, rebind_functional=True):
    """Re-class every torch.nn operator instance of `model` in place (parameters, buffers and hooks untouched).

    dtype          activation dtype of the body (torch.bfloat16 or torch.float32)
    allow_umma     False forces the fp32-FMA convolution kernels (tf32-off mode, tolerance 1e-4)
    fp32_heads     convolutions with <= 4 output channels (segmentation logits) emit fp32
    sync           (process_group, world_size): BatchNorm statistics are all-reduced (SyncBN)
    rebind_functional  route `F.interpolate/F.upsample/...` calls made inside the model's own forward()s
                   through `functional_proxy` by rebinding the name `F` in the defining modules' namespaces
    """
    if dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("convert: dtype must be torch.float32 or torch.bfloat16")
    for m in model.modules():
        repl = _SWAP.get(type(m))
        if repl is not None:
            m.__class__ = repl
        if isinstance(m, _B200Mixin):
            m.compute_dtype = dtype
            m.allow_umma = bool(allow_umma) and dtype == torch.bfloat16
            if isinstance(m, _BatchNormMixin):
                m.sync = sync
            if isinstance(m, _ConvMixin):
                co = m.out_channels
                m.out_dtype = torch.float32 if (fp32_heads and co <= 4) else None
        if _is_foreign_sink(m) and not getattr(m, "_b200_sink_hook", False):
            m.register_forward_pre_hook(_leave_body_hook)
            m._b200_sink_hook = True
        if rebind_functional:
            ns = sys.modules.get(type(m).__module__)
            if ns is not None and not type(m).__module__.startswith(("torch.", "mri_epilepsy_diagnosis_b200")):
                for name in ("F", "functional"):
                    if getattr(ns, name, None) is TF:
                        setattr(ns, name, functional_proxy)
    return model


@contextlib.contextmanager
def patch(functional=True):
    """Swap the operator classes on `torch.nn` (and, optionally, interpolate/upsample on torch.nn.functional)
    while a reference file constructs its model: `with patch(): net = unet3d.Unet(...)`."""
    saved = {name: getattr(tnn, name) for name in ("Conv3d", "Conv2d", "ConvTranspose3d", "BatchNorm3d", "BatchNorm2d", "InstanceNorm3d",
                                                    "GroupNorm", "ReLU", "LeakyReLU", "PReLU", "MaxPool3d", "MaxPool2d", "Upsample")}
    g = globals()
    for name in saved:
        setattr(tnn, name, g[name])
    saved_f = {}
    if functional:
        for name in ("interpolate", "upsample"):
            saved_f[name] = getattr(TF, name)
            setattr(TF, name, _f_interpolate)
    try:
        yield
    finally:
        for name, cls in saved.items():
            setattr(tnn, name, cls)
        for name, fn in saved_f.items():
            setattr(TF, name, fn)
