import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def thin(t, cap=32768):
    """Same sub-sampling rule as oracle/make_golden.py:thin."""
    if t.numel() <= cap:
        return t
    flat = t.reshape(-1)
    return flat[:: flat.numel() // cap]


def rel_err(a, b):
    import torch
    a = torch.as_tensor(np.asarray(a)).double() if not hasattr(a, "double") else a.detach().cpu().double()
    b = torch.as_tensor(np.asarray(b)).double() if not hasattr(b, "double") else b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def tensor_core_convs(net, x):
    """Names of the convolution modules of `net` whose forward, on input `x`, runs on the tcgen05 kernels -- i.e. the layers that
    consume a bf16 copy of their weight (the oracle's bf16-storage mode rounds exactly those weights).  One eval-mode dry run with
    forward pre-hooks (eval: every convolution goes through its own module, nothing is updated); asks the library
    (b200_conv_algo), does not guess from the channel counts."""
    import ctypes
    import torch
    from mri_epilepsy_diagnosis_b200 import _cabi as cabi
    from mri_epilepsy_diagnosis_b200.nn import _ConvMixin
    found, hooks = set(), []

    def make(name):
        def pre(mod, args):
            xx = mod._prep(args[0])
            cd, _ = mod._cfg().desc(xx, mod.weight, mod.out_dtype or xx.dtype)
            fwd = cabi.lib().b200_conv_algo(ctypes.byref(cd), cabi.PASS_FWD)
            if fwd != cabi.ALGO_SIMT:
                found.add(name)
        return pre
    for name, m in net.named_modules():
        if isinstance(m, _ConvMixin):
            hooks.append(m.register_forward_pre_hook(make(name)))
    was = net.training
    net.eval()
    with torch.no_grad():
        net(x)
    net.train(was)
    for h in hooks:
        h.remove()
    return found
