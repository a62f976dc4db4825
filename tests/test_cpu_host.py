"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, host logic,
state-dict compatibility with the shipped checkpoints, oracle isolation, data-parallel plumbing (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def _lib_path():
    return os.path.join(ROOT, "mri_epilepsy_diagnosis_b200", "libb200nn.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib_path()):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    return _lib_path()


def test_library_exports_every_header_symbol(built):
    header = open(os.path.join(ROOT, "include", "b200nn.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 30
    handle = ctypes.CDLL(built)
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    from mri_epilepsy_diagnosis_b200 import _cabi
    assert sorted(_cabi.EXPORTS) == names, "ctypes table and header disagree"
    handle.b200_version.restype = ctypes.c_int
    assert handle.b200_version() == 100


def test_library_is_sm100a_with_blackwell_instructions(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_desc_validation_without_gpu(built):
    """Descriptor validation is host code: it must reject bad shapes before any launch."""
    from mri_epilepsy_diagnosis_b200 import _cabi
    L = _cabi.lib()
    d = _cabi.ConvDesc(0, 0, 1, 4, 8, 8, 8, 4, 9, 8, 8, 3, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1)
    assert L.b200_conv_fwd(ctypes.byref(d), 1, 1, None, 1, None, 0, None) != 0
    assert b"output size" in L.b200_last_error()
    nd = _cabi.NormDesc(0, 2, 6, 10, 2, 4, 1e-5, 0.1, 0, 0.0)        # GroupNorm with C % G != 0
    assert L.b200_norm_apply(ctypes.byref(nd), 16, 16, 16, None, None, None, 16, None) != 0


def test_modules_refuse_cpu_tensors(built):
    from mri_epilepsy_diagnosis_b200 import nn as bnn
    for m, x in ((bnn.Conv3d(1, 2, 3), torch.zeros(1, 1, 4, 4, 4)), (bnn.BatchNorm3d(2), torch.zeros(1, 2, 4, 4, 4)),
                 (bnn.MaxPool3d(2), torch.zeros(1, 2, 4, 4, 4)), (bnn.ReLU(), torch.zeros(4))):
        with pytest.raises(RuntimeError, match="CUDA"):
            m(x)


def test_missing_library_fails_loudly(built, monkeypatch):
    from mri_epilepsy_diagnosis_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libb200nn.so")
    with pytest.raises(RuntimeError, match="no CPU or cuDNN fallback"):
        _cabi.lib()


def test_shipped_checkpoints_load_strict():
    from mri_epilepsy_diagnosis_b200 import zoo
    ld = lambda n: torch.load(os.path.join(GOLDEN, n), map_location="cpu", weights_only=True)
    zoo.FepegarUNet(out_channels_first_layer=8).load_state_dict(ld("whole_im_train_seg_parc_epoch_7.pth"), strict=True)
    zoo.fader_encoder().load_state_dict(ld("encoder_93_6_4.pth"), strict=True)
    zoo.Classificator(n_class=2, **zoo.FADER_HEAD).load_state_dict(ld("clf_93_6_4.pth"), strict=True)
    zoo.Discriminator(n_domains=18, **zoo.FADER_HEAD).load_state_dict(ld("disc_93_6_4.pth"), strict=True)


def test_zoo_state_dicts_match_reference_key_lists():
    from mri_epilepsy_diagnosis_b200 import zoo
    from oracle import weights
    net = zoo.Unet(c=1, n=16, norm="bn", num_classes=2)
    assert list(net.state_dict().keys()).sort() == list(weights.unet3d_state().keys()).sort()
    assert len(net.state_dict()) == 162 and sum(p.numel() for p in net.parameters()) == 9449590
    assert sum(p.numel() for p in zoo.config1_autoencoder().parameters()) == 3675623
    assert sum(p.numel() for p in zoo.PatchModel().parameters()) == 2556914
    assert sum(p.numel() for p in zoo.FepegarUNet(out_channels_first_layer=8).parameters()) == 246412
    assert sum(p.numel() for p in zoo.FepegarUNet(out_channels_first_layer=16).parameters()) == 983564
    assert len(zoo.Unet(c=1, n=16, norm="in", num_classes=2).state_dict()) == len(weights.unet3d_state(norm="in"))


def test_convert_keeps_parameters_and_keys():
    import torch.nn as nn
    from mri_epilepsy_diagnosis_b200 import nn as bnn
    ref = nn.Sequential(nn.Conv3d(1, 8, 3), nn.BatchNorm3d(8), nn.ReLU(inplace=True), nn.MaxPool3d(2), nn.ConvTranspose3d(8, 2, 2, 2),
                        nn.GroupNorm(2, 2), nn.PReLU(), nn.Upsample(scale_factor=2), nn.Conv2d(1, 1, 1))
    before = {k: v.data_ptr() for k, v in ref.state_dict().items()}
    net = bnn.convert(ref, dtype=torch.bfloat16)
    assert {k: v.data_ptr() for k, v in net.state_dict().items()} == before
    assert all(type(m).__module__.startswith("mri_epilepsy") for m in net)
    assert net[0].compute_dtype == torch.bfloat16 and net[2].inplace and isinstance(net[0], nn.Conv3d)
    assert net[4].out_dtype == torch.float32 and net[0].out_dtype is None          # <=4 output channels -> fp32 head


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import it (tier rule 3)."""
    pkg = os.path.join(ROOT, "mri_epilepsy_diagnosis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_conv_output_shape_logic():
    from mri_epilepsy_diagnosis_b200.functional import ConvConfig
    w = torch.zeros(8, 1, 6, 1, 1)
    x = torch.zeros(2, 1, 192, 192, 192, device="meta")
    cd, shape = ConvConfig((2, 1, 1), (2, 0, 0), 1).desc(x, w, torch.float32)
    assert shape == (2, 8, 96, 192, 192) and (cd.kd, cd.sd, cd.pd) == (6, 2, 2)          # AE_model.py:9-14 with k6 s2 p2
    cd, shape = ConvConfig(2, 0, 1, transposed=True).desc(torch.zeros(1, 4, 5, 6, 7, device="meta"), torch.zeros(4, 3, 2, 2, 2), torch.float32)
    assert shape == (1, 3, 10, 12, 14)
    cd, shape = ConvConfig(1, 0, 1).desc(torch.zeros(4, 2, 16, 32, device="meta"), torch.zeros(16, 2, 3, 3), torch.float32)
    assert shape == (4, 16, 14, 30) and cd.Di == 1 and cd.kd == 1


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_epilepsy_diagnosis_b200 import dp
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    dead = torch.nn.Linear(4, 4)                       # parameters that never get a gradient (unet3d dead branch)
    params = list(net.parameters()) + list(dead.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    dp.broadcast_parameters(net)
    bucket = dp.GradientBucket(params, opt, buckets=2)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 3, generator=g)
    mine = dp.shard(list(range(8)), rank, world)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(X[mine]), Y[mine]).backward()
        opt.step()
    # gradient accumulation: two backward passes before one step (the second must not race the first pass's collective)
    opt.zero_grad()
    for half in (mine[:2], mine[2:]):
        torch.nn.functional.mse_loss(net(X[half]), Y[half]).backward()
    opt.step()
    # parameters without a gradient on any rank keep grad=None: the optimizer skips them as on one GPU
    q.put((rank, [p.detach().tolist() for p in net.parameters()], [p.grad is None for p in dead.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_bucket_matches_single_process_gloo():
    """world_size=2 over gloo == one process on the global batch (equal shards, mean loss)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    [p.join(60) for p in procs]
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 3, generator=g)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(X), Y).backward()
        opt.step()
    opt.zero_grad()
    for half in ([0, 1, 4, 5], [2, 3, 6, 7]):           # first / second halves of both ranks' shards
        torch.nn.functional.mse_loss(net(X[half]), Y[half]).backward()
    opt.step()
    for r in res:
        for a, b in zip(r[1], net.parameters()):
            assert torch.allclose(torch.tensor(a), b.detach(), atol=1e-6)
        assert all(r[2])


def test_ctypes_descriptors_match_the_header_layout(tmp_path):
    """Every POD descriptor of include/b200nn.h against its ctypes mirror in _cabi.py: same size and the same offset for every field
    (a C program compiled with gcc prints sizeof / offsetof; a silent mismatch would hand the kernels garbage geometry)."""
    import ctypes
    import shutil
    import subprocess
    from mri_epilepsy_diagnosis_b200 import _cabi as cabi
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"b200_conv_desc": cabi.ConvDesc, "b200_pack_entry": cabi.PackEntry, "b200_norm_desc": cabi.NormDesc, "b200_dice_desc": cabi.DiceDesc,
             "b200_pool_desc": cabi.PoolDesc, "b200_up_desc": cabi.UpDesc, "b200_patch_desc": cabi.PatchDesc, "b200_histstd_desc": cabi.HistStdDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200nn.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = pairs[cname]
        expect = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == expect, f"{cname}.{field}: header {value}, ctypes {expect}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
