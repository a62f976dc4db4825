"""mri_epilepsy_diagnosis_b200 -- B200-native (sm_100a) implementation of the 3-D-convolutional hot path of
kondratevakate/mri-epilepsy-diagnosis, behind the reference's own operator boundary (torch.nn).

    from mri_epilepsy_diagnosis_b200 import nn as b200nn
    model = b200nn.convert(reference_model.cuda(), dtype=torch.bfloat16)     # routine.py loops run as-is

Everything computes in libb200nn.so (hand-written CUDA, C ABI in include/b200nn.h); there is no CPU path,
no cuDNN dispatch and no Triton.  Importing the package does not load the library; the first operator call does,
and raises RuntimeError if it is missing.
"""
from . import _cabi, functional, nn, patches, zoo, dp, graphed, detect, preprocess, metrics, grid  # noqa: F401
from .nn import convert, patch  # noqa: F401

__version__ = "0.1.0"


def library_path():
    return _cabi.LIB_PATH


def launch_count():
    """Kernels launched by libb200nn.so in this process (bench.py's gpu_launches)."""
    return int(_cabi.lib().b200_launch_count())
