"""ctypes binding of libb200nn.so (include/b200nn.h).  There is no fallback: if the shared
library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200nn.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
NORM_BATCH, NORM_INSTANCE, NORM_GROUP = 0, 1, 2
UP_NEAREST, UP_TRILINEAR, UP_TRILINEAR_ALIGNED = 0, 1, 2
PASS_FWD, PASS_DGRAD, PASS_WGRAD = 0, 1, 2
ALGO_SIMT, ALGO_UMMA, ALGO_ROW = 0, 1, 2

i32, i64, f32, vp, sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class ConvDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("x_dtype", "y_dtype", "N", "Ci", "Di", "Hi", "Wi", "Co", "Do", "Ho", "Wo", "kd", "kh", "kw",
                                   "sd", "sh", "sw", "pd", "ph", "pw", "dd", "dh", "dw", "transposed", "allow_umma")]


class PackEntry(C.Structure):
    _fields_ = [("src0", vp), ("src1", vp), ("idx", vp), ("dst", vp), ("count", i64), ("first_block", i64), ("n0", i32), ("dst_bf16", i32)]


class NormDesc(C.Structure):
    _fields_ = [("dtype", i32), ("N", i32), ("C", i32), ("S", i64), ("kind", i32), ("G", i32), ("eps", f32), ("momentum", f32),
                ("act", i32), ("slope", f32)]


class PoolDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("dtype", "N", "C", "Di", "Hi", "Wi", "Do", "Ho", "Wo", "kd", "kh", "kw", "sd", "sh", "sw")]


class UpDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("dtype", "mode", "N", "C", "Di", "Hi", "Wi", "Do", "Ho", "Wo", "Ctot", "c_off")]


class DiceDesc(C.Structure):
    _fields_ = [("dtype", i32), ("N", i32), ("C", i32), ("S", i64), ("eps", f32)]


class HistStdDesc(C.Structure):
    _fields_ = [("q", C.c_double * 16), ("landmarks", C.c_double * 16), ("range_idx", i32 * 16), ("nq", i32), ("nrange", i32), ("eps", C.c_double)]


class PatchDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("X", "Y", "Z", "h", "w", "with_mask", "upsample_passes")]


P = C.POINTER
_SIGS = {
    "b200_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_launch_count": (C.c_uint64, []),
    "b200_conv_algo": (C.c_int, [P(ConvDesc), C.c_int]),
    "b200_conv_packed_bytes": (sz, [P(ConvDesc), C.c_int]),
    "b200_conv_pack_weights": (C.c_int, [P(ConvDesc), C.c_int, vp, vp, vp]),
    "b200_pack_batched": (C.c_int, [vp, C.c_int, i64, vp]),
    "b200_conv_workspace_bytes": (sz, [P(ConvDesc), C.c_int]),
    "b200_conv_fwd": (C.c_int, [P(ConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    "b200_conv_stats_chunks": (C.c_int, [P(ConvDesc)]),
    "b200_conv_fwd_stats": (C.c_int, [P(ConvDesc), vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_conv_fwd_stats_tail": (C.c_int, [P(ConvDesc), vp, vp, vp, vp, C.c_int, vp, vp, sz, vp]),
    "b200_conv_dgrad": (C.c_int, [P(ConvDesc), vp, vp, vp, vp, sz, vp]),
    "b200_conv_wgrad": (C.c_int, [P(ConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    "b200_norm_workspace_bytes": (sz, [P(NormDesc)]),
    "b200_norm_stats": (C.c_int, [P(NormDesc), vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_norm_stats_from_partial": (C.c_int, [P(NormDesc), vp, C.c_int, vp, vp, vp, vp, vp]),
    "b200_norm_stats_from_running": (C.c_int, [P(NormDesc), vp, vp, vp, vp, vp]),
    "b200_syncbn_pack": (C.c_int, [i32, f32, vp, vp, vp, vp]),
    "b200_syncbn_finalize": (C.c_int, [i32, f32, f32, C.c_double, vp, vp, vp, vp, vp, vp]),
    "b200_norm_apply": (C.c_int, [P(NormDesc), vp, vp, vp, vp, vp, vp, vp, vp]),
    "b200_norm_bwd": (C.c_int, [P(NormDesc), C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_norm_bwd_reduce": (C.c_int, [P(NormDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_norm_bwd_apply": (C.c_int, [P(NormDesc), C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_act_fwd": (C.c_int, [C.c_int, C.c_int, f32, i64, vp, vp, vp]),
    "b200_act_bwd": (C.c_int, [C.c_int, C.c_int, f32, i64, vp, vp, vp, vp]),
    "b200_prelu_fwd": (C.c_int, [C.c_int, i64, vp, vp, vp, vp]),
    "b200_prelu_workspace_bytes": (sz, [i64]),
    "b200_prelu_bwd": (C.c_int, [C.c_int, i64, vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_add_act_fwd": (C.c_int, [C.c_int, C.c_int, f32, i64, vp, vp, vp, vp]),
    "b200_softmax_dice_workspace_bytes": (sz, [P(DiceDesc)]),
    "b200_softmax_dice_fwd": (C.c_int, [P(DiceDesc), vp, vp, vp, vp, vp, sz, vp]),
    "b200_softmax_dice_bwd": (C.c_int, [P(DiceDesc), vp, vp, vp, vp, vp, vp]),
    "b200_maxpool_fwd": (C.c_int, [P(PoolDesc), vp, vp, vp, vp, vp]),
    "b200_maxpool_bwd": (C.c_int, [P(PoolDesc), vp, vp, vp, vp]),
    "b200_maxpool_bwd_add": (C.c_int, [P(PoolDesc), vp, vp, vp, C.c_int, vp, vp]),
    "b200_upsample_fwd": (C.c_int, [P(UpDesc), vp, vp, vp]),
    "b200_upsample_bwd": (C.c_int, [P(UpDesc), vp, vp, vp]),
    "b200_copy_channels": (C.c_int, [C.c_int, i64, i32, vp, i32, i32, vp, i32, i32, vp]),
    "b200_to_channels_last": (C.c_int, [C.c_int, C.c_int, i32, i32, i64, vp, vp, vp]),
    "b200_from_channels_last": (C.c_int, [C.c_int, C.c_int, i32, i32, i64, vp, vp, vp]),
    "b200_patch_workspace_bytes": (sz, [P(PatchDesc)]),
    "b200_patch_max_rows": (i64, [P(PatchDesc)]),
    "b200_patch_plan": (C.c_int, [P(PatchDesc), vp, vp, vp, vp, vp, vp, sz, vp]),
    "b200_patch_gather": (C.c_int, [P(PatchDesc), vp, vp, i64, C.c_int, vp, vp]),
    "b200_grid_gather": (C.c_int, [C.c_int, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "b200_grid_aggregate": (C.c_int, [C.c_int, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "b200_overlap_counts": (C.c_int, [vp, vp, i64, vp, vp]),
    "b200_minmax_workspace_bytes": (sz, []),
    "b200_minmax_normalize": (C.c_int, [vp, i64, vp, vp, sz, vp]),
    "b200_fcd_scatter_labels": (C.c_int, [vp, i64, vp, i32, i32, i32, i32, i32, vp, vp]),
    "b200_fcd_vote": (C.c_int, [vp, i32, i32, C.c_int, vp, vp, vp]),
    "b200_fcd_paint": (C.c_int, [vp, i64, vp, i32, i32, i32, i32, i32, vp, vp]),
    "b200_surface_codes": (C.c_int, [vp, i32, i32, i32, vp, vp, vp]),
    "b200_surface_edt": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, vp]),
    "b200_surface_collect": (C.c_int, [vp, vp, C.c_int, i64, vp, vp, vp, vp]),
    "b200_histstd_workspace_bytes": (sz, []),
    "b200_histstd_normalize": (C.c_int, [P(HistStdDesc), vp, vp, i64, vp, vp, vp, sz, vp]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """Load libb200nn.so (once).  Raises RuntimeError -- never falls back -- when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU or cuDNN fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("b200nn: " + lib().b200_last_error().decode())


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise RuntimeError(f"b200nn supports float32 and bfloat16 activations, got {dt}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"b200nn.{what}: expected a CUDA tensor (there is no CPU path), got device {t.device}")
    if t.device.index != torch.cuda.current_device():
        # kernels are enqueued on the CURRENT device's stream; a tensor of another device would be touched from the wrong context
        raise RuntimeError(f"b200nn.{what}: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}; "
                           f"call torch.cuda.set_device({t.device.index}) (one process per GPU) or wrap the call in torch.cuda.device(...)")
