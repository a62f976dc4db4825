"""Import the REAL reference modules from /root/reference (TEST INFRASTRUCTURE).

Only usable where the read-only reference checkout exists (the build container); the
GPU box does not have it, so nothing in `-m gpu` tests, smoke() or bench.py calls
this.  It is how `oracle/make_golden.py` produces tests/golden/*, and how
tests/test_oracle_vs_reference.py re-checks the restatement live when it can.

Missing third-party imports of the reference files (nibabel, nilearn, nipype,
matplotlib, IPython, torchio, unet, comet_ml) are satisfied with empty stub modules
(SURVEY.md section 8c); no reference source is copied or modified.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "segmentation", "models"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name) if name[0].islower() else type(name, (), {})

    def __call__(self, *a, **k):
        return None


def _stub(*names):
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def unet3d_module():
    return _load("_ref_unet3d", "segmentation/models/unet3d.py")


def make_unet3d(**kw):
    """Construct unet3d.Unet despite the broken `F.interpolate(scale_factor=...)` at :85,
    by letting that one no-input call return the upstream nn.Upsample."""
    import torch.nn as nn
    import torch.nn.functional as F
    m = unet3d_module()
    real = F.interpolate

    def shim(*a, **k):
        if not a and "input" not in k:
            return nn.Upsample(scale_factor=k["scale_factor"], mode=k["mode"], align_corners=k["align_corners"])
        return real(*a, **k)

    m.F.interpolate = shim
    try:
        net = m.Unet(**kw)
    finally:
        m.F.interpolate = real
    return net


def ae_module():
    return _load("_ref_ae_model", "classification/models/AE_model.py")


def cnn_module():
    return _load("_ref_cnn_model", "classification/models/cnn_model.py")


def modified_unet_module():
    return _load("_ref_modified_3dunet", "segmentation/models/modified_3dunet.py")


def patch_utils_module():
    _stub("nibabel", "nilearn", "nilearn.plotting", "nilearn.datasets", "nilearn.image",
          "matplotlib", "matplotlib.pyplot", "nipype", "nipype.interfaces", "nipype.interfaces.fsl")
    return _load("_ref_patch_utils", "detection/patch_utils.py")


def patch_model_classes():
    """detection/model_utils.py is not valid Python (:10, :12, :230); take the two class
    definitions (:19-52) out of the file text and exec them in a clean namespace."""
    src = open(os.path.join(REF, "detection/model_utils.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("class PatchModel"))
    end = next(i for i, l in enumerate(src) if l.startswith("def train"))
    code = "\n".join(src[start:end])
    ast.parse(code)
    import torch
    import torch.nn as nn
    ns = {"torch": torch, "nn": nn}
    exec(compile(code, "model_utils.py[19:52]", "exec"), ns)
    return ns["PatchModel"], ns["ConvolutionBlock"]


def fcd_mask_generator_class(model):
    """`class FCDMaskGenerator` (detection/model_utils.py:118-228) out of the syntactically invalid file, exec'd with the
    names it expects as globals (`model` -- :132 uses a GLOBAL, not self.model --, np, torch, scipy.signal.convolve).  The text
    is used as it is except that `.cuda()` calls are dropped (no GPU in the build container); instances are created with
    object.__new__ because __init__ reads best_model.pth and the template through nibabel (:120-127)."""
    src = open(os.path.join(REF, "detection/model_utils.py")).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("class FCDMaskGenerator"))
    end = next(i for i, l in enumerate(src) if "def save_nii_mask" in l)
    code = "\n".join(src[start:end]).replace(".cuda()", "")
    ast.parse(code)
    import numpy as np
    import torch
    from scipy.signal import convolve
    ns = {"torch": torch, "np": np, "convolve": convolve, "model": model, "nib": _Stub("nibabel"), "PatchModel": None}
    exec(compile(code, "model_utils.py[118:228]", "exec"), ns)
    return ns["FCDMaskGenerator"]


def histstd_functions():
    """`normalize`, `_get_percentiles`, `_standardize_cutoff` of classification/train_ENC_CLF.ipynb [cell 9], exec'd as written
    up to `def default_collate` with the names the notebook had imported; `np.bool` (removed in numpy 1.24, used at the
    `mask is None` branch) is restored as an alias of `bool` on a numpy PROXY namespace, not on numpy itself."""
    import json
    import types
    from typing import Tuple
    import numpy as np
    import torch
    nb = json.load(open(os.path.join(REF, "classification/train_ENC_CLF.ipynb")))
    src = "".join(nb["cells"][9]["source"])
    assert "def normalize(" in src and "def default_collate" in src
    code = src[:src.index("def default_collate")]
    npx = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    npx.bool = bool
    ns = {"np": npx, "torch": torch, "Tuple": Tuple}
    exec(compile(code, "train_ENC_CLF.ipynb[cell 9]", "exec"), ns)
    return ns


def seg_routine_module():
    """segmentation/routine.py with its third-party imports stubbed."""
    _stub("IPython", "IPython.display", "matplotlib", "matplotlib.pyplot", "torchio", "torchio.transforms",
          "unet", "comet_ml", "torchvision", "torchvision.models", "torchvision.models.vgg", "metrics")
    import sys as _s
    tio = _s.modules["torchio"]
    for k in ("AFFINE", "DATA", "PATH", "TYPE", "STEM"):
        setattr(tio, k, k.lower())
    sys.path.insert(0, os.path.join(REF, "segmentation"))
    try:
        return _load("_ref_seg_routine", "segmentation/routine.py")
    finally:
        sys.path.pop(0)


def clf_routine_module():
    _stub("IPython", "IPython.display", "matplotlib", "matplotlib.pyplot", "comet_ml")
    return _load("_ref_clf_routine", "classification/routine.py")


def path(rel):
    return os.path.join(REF, rel)
