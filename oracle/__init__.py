"""CPU oracle for the 3-D-convolutional hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (plain PyTorch fp32 on the host + numpy) of what
the reference `kondratevakate/mri-epilepsy-diagnosis` computes on the path named by
BASELINE.json:north_star.  Every function cites the reference file:line it follows.

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
    `--impl reference` legs may import anything from here;
  * nothing under `mri_epilepsy_diagnosis_b200/` imports it -- the product path is the
    CUDA library and fails loudly when that library is missing.

Parity status: PINNED.  The oracle is checked against outputs of the reference's own
Python modules, imported from /root/reference in the build container by
`oracle/make_golden.py` (committed) and frozen under `tests/golden/`.  The one piece
that cannot be pinned that way is the third-party `unet.UNet` (fepegar/unet, PyPI
`unet`, version unpinned, not vendored in the reference): `graphs.fepegar_unet` is
restated from the call site segmentation/routine.py:346-356 and the key/shape list of
the shipped checkpoints, and is anchored on strict loading of every
segmentation/weights/*.pth -- its numerics are "parity unpinned" beyond that.
"""
