"""The drop-in boundary, checked without a GPU: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/mrb_b200.h declares; `_C` exposes the reference's 14 names; `layers`
exposes the reference's __all__ with the reference's constructor signatures; CPU tensors are refused
loudly (no fallback)."""
import ctypes
import inspect
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REF_C_NAMES = [  # reference csrc/vision.cpp:10-24
    "nms", "roi_align_forward", "roi_align_backward", "roi_pool_forward", "roi_pool_backward",
    "sigmoid_focalloss_forward", "sigmoid_focalloss_backward", "deform_conv_forward",
    "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
    "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward"]

REF_LAYERS_ALL = [  # reference layers/__init__.py:23-46
    "nms", "roi_align", "ROIAlign", "roi_pool", "ROIPool", "smooth_l1_loss", "Conv2d", "DFConv2d",
    "ConvTranspose2d", "interpolate", "BatchNorm2d", "FrozenBatchNorm2d", "SigmoidFocalLoss", "deform_conv",
    "modulated_deform_conv", "DeformConv", "ModulatedDeformConv", "ModulatedDeformConvPack",
    "deform_roi_pooling", "DeformRoIPooling", "DeformRoIPoolingPack", "ModulatedDeformRoIPoolingPack"]


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mrb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/mrb_b200.h but not exported: %s" % missing
    lib.mrb_error_string.restype = ctypes.c_char_p
    assert lib.mrb_error_string(-2).decode().startswith("mrb:")
    assert lib.mrb_version() == 100


def test_library_is_sm100a_and_has_no_torch_dependency(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    ldd = subprocess.run(["ldd", built_lib], capture_output=True, text=True).stdout
    libs = [l.split()[0] for l in ldd.splitlines() if l.strip()]
    assert not [l for l in libs if "torch" in l or "c10" in l or "cudnn" in l or "cublas" in l], libs


def test_workspace_queries_run_on_host(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.mrb_nms_workspace_bytes.restype = ctypes.c_size_t
    n = 2000
    cb = (n + 63) // 64
    assert lib.mrb_nms_workspace_bytes(ctypes.c_int(n)) >= n * 16 + n * 4 + n * 4 + n + n * cb * 8
    assert lib.mrb_nms_workspace_bytes(ctypes.c_int(0)) > 0


def test_C_surface(built_lib):
    from maskrcnn_benchmark import _C
    for name in REF_C_NAMES:
        assert callable(getattr(_C, name)), name
    # positional arity of the reference signatures (csrc/*.h)
    arity = {"nms": 3, "roi_align_forward": 6, "roi_align_backward": 10, "roi_pool_forward": 5,
             "roi_pool_backward": 11, "sigmoid_focalloss_forward": 5, "sigmoid_focalloss_backward": 6,
             "deform_conv_forward": 17, "deform_conv_backward_input": 18, "deform_conv_backward_parameters": 18,
             "modulated_deform_conv_forward": 19, "modulated_deform_conv_backward": 24,
             "deform_psroi_pooling_forward": 13, "deform_psroi_pooling_backward": 15}
    for name, n in arity.items():
        assert len(inspect.signature(getattr(_C, name)).parameters) == n, name


def test_C_refuses_cpu_tensors(built_lib):
    from maskrcnn_benchmark import _C
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.nms(torch.rand(4, 4), torch.rand(4), 0.5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.roi_align_forward(torch.rand(1, 2, 8, 8), torch.zeros(1, 5), 1.0, 2, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C.sigmoid_focalloss_forward(torch.rand(4, 8), torch.zeros(4, dtype=torch.int32), 8, 2.0, 0.25)


def test_layers_surface(built_lib):
    from maskrcnn_benchmark import layers
    assert sorted(layers.__all__) == sorted(REF_LAYERS_ALL)
    for n in REF_LAYERS_ALL:
        assert hasattr(layers, n), n
    assert repr(layers.ROIAlign((7, 7), 0.25, 2)) == "ROIAlign(output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)"
    assert repr(layers.ROIPool((7, 7), 0.0625)) == "ROIPool(output_size=(7, 7), spatial_scale=0.0625)"
    assert repr(layers.SigmoidFocalLoss(2.0, 0.25)) == "SigmoidFocalLoss(gamma=2.0, alpha=0.25)"
    bn = layers.FrozenBatchNorm2d(4)
    assert sorted(dict(bn.named_buffers())) == ["bias", "running_mean", "running_var", "weight"]
    assert not list(bn.parameters())
    df = layers.DFConv2d(8, 16, with_modulated_dcn=True)
    assert sorted(k for k, _ in df.named_parameters()) == ["conv.weight", "offset.bias", "offset.weight"]
    assert df.offset.out_channels == 27
    assert layers.DFConv2d(8, 16, with_modulated_dcn=False).offset.out_channels == 18
    pk = layers.ModulatedDeformConvPack(8, 8, 3, padding=1)
    assert "conv_offset_mask.weight" in dict(pk.named_parameters())


def test_layers_cpu_semantics_that_need_no_kernel():
    """Pieces of `layers` that are plain tensor algebra in the reference too."""
    from maskrcnn_benchmark import layers
    # FrozenBatchNorm2d: x * w * rsqrt(var) + (b - mean * w * rsqrt(var)), NO eps (batch_norm.py:27-31)
    bn = layers.FrozenBatchNorm2d(3)
    g = torch.Generator().manual_seed(0)
    bn.weight.copy_(torch.rand(3, generator=g) + 0.5)
    bn.bias.copy_(torch.randn(3, generator=g))
    bn.running_mean.copy_(torch.randn(3, generator=g))
    bn.running_var.copy_(torch.rand(3, generator=g) * 1.5 + 0.5)
    x = torch.randn(2, 3, 5, 7, generator=g)
    want = (x - bn.running_mean[None, :, None, None]) / bn.running_var.sqrt()[None, :, None, None] \
        * bn.weight[None, :, None, None] + bn.bias[None, :, None, None]
    torch.testing.assert_close(bn(x), want, rtol=1e-5, atol=1e-5)
    s, b = bn.scale_shift()
    torch.testing.assert_close(x * s[None, :, None, None] + b[None, :, None, None], bn(x))
    # smooth_l1_loss with beta knee
    a, t = torch.tensor([0.0, 0.05, 1.0]), torch.zeros(3)
    got = layers.smooth_l1_loss(a, t, beta=1. / 9, size_average=False)
    want = 0.5 * 0.05 ** 2 * 9 + (1.0 - 0.5 / 9)
    assert abs(float(got) - want) < 1e-6
    # empty-batch shims
    conv = layers.Conv2d(4, 6, 3, stride=2, padding=1)
    assert conv(torch.zeros(0, 4, 9, 9)).shape == (0, 6, 5, 5)
    assert layers.interpolate(torch.zeros(0, 4, 5, 5), scale_factor=2).shape == (0, 4, 10, 10)
    dt = layers.ConvTranspose2d(4, 2, 2, 2, 0)
    assert dt(torch.zeros(0, 4, 7, 7)).shape == (0, 2, 14, 14)
    # SigmoidFocalLoss on CPU tensors follows the reference's python path
    logits, targets = torch.randn(6, 5, generator=g), torch.tensor([0, 1, 5, -1, 2, 0], dtype=torch.int32)
    crit = layers.SigmoidFocalLoss(2.0, 0.25)
    assert torch.isfinite(crit(logits, targets))


def test_reference_modeling_imports_over_our_layers(built_lib, tmp_path):
    """Boundary acceptance (authoring container only: needs the reference checkout): the reference's
    own modeling/backbone + poolers import UNMODIFIED on top of our `layers` / `_C` and build the
    R-50-FPN backbone.  Runs in a subprocess so that MRB_REFERENCE_ROOT path stitching is isolated."""
    ref = os.environ.get("MRB_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "maskrcnn_benchmark")):
        pytest.skip("reference checkout absent")
    shims = os.path.join(ROOT, "tests", "_shims")
    code = r'''
import sys, os
sys.path.insert(0, os.path.join(%r, "maskrcnn-benchmark_b200")); sys.path.insert(0, %r)
import mrb_test_compat  # test-only: torch._six / np.float aliases the reference still uses
import torch
import maskrcnn_benchmark, maskrcnn_benchmark.layers as L
assert "maskrcnn-benchmark_b200" in L.__file__
from maskrcnn_benchmark.config import cfg
cfg.merge_from_file(os.path.join(%r, "configs", "e2e_mask_rcnn_R_50_FPN_1x.yaml"))
from maskrcnn_benchmark.modeling.backbone import build_backbone
from maskrcnn_benchmark.modeling.poolers import Pooler
import maskrcnn_benchmark.modeling.backbone.resnet as R
assert "reference" in R.__file__ or %r in R.__file__
bb = build_backbone(cfg)
n = sum(p.numel() for p in bb.parameters())
assert type(bb.body.stem.bn1).__module__ == "maskrcnn_benchmark.layers.batch_norm"
assert "maskrcnn-benchmark_b200" in sys.modules["maskrcnn_benchmark.layers.batch_norm"].__file__
pool = Pooler((7, 7), (0.25, 0.125, 0.0625, 0.03125), 2)
assert type(pool.poolers[0]).__module__ == "maskrcnn_benchmark.layers.roi_align"
from maskrcnn_benchmark.modeling.detector import build_detection_model
m = build_detection_model(cfg)
keys = list(m.state_dict().keys())
print("OK", n, len(keys))
''' % (ROOT, shims, ref, ref)
    env = dict(os.environ, MRB_REFERENCE_ROOT=ref)
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().startswith("OK")


def test_flat_sgd_matches_torch_sgd_and_bumps_versions():
    """mrb_b200.optim.FlatSGD == torch.optim.SGD (momentum, weight decay) on channels_last conv weights, and the
    update is visible through p._version (the bf16 operand caches of the backend key on it)."""
    import copy
    from mrb_b200.optim import FlatSGD
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.Conv2d(8, 8, 1))
    for p in m.parameters():
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    m2 = copy.deepcopy(m)
    o1 = FlatSGD(m.named_parameters(), lr=0.1, momentum=0.9, weight_decay=0.01, bias_lr_factor=1.0, weight_decay_bias=0.01)
    o2 = torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9, weight_decay=0.01)
    versions = [p._version for p in m.parameters()]
    for it in range(3):
        x = torch.randn(2, 4, 8, 8)
        for mm, oo in ((m, o1), (m2, o2)):
            oo.zero_grad()
            mm(x).square().mean().backward()
            oo.step()
    for a, b in zip(m.parameters(), m2.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
    assert all(p._version > v for p, v in zip(m.parameters(), versions))


def test_glue_wrappers_refuse_cpu_tensors(built_lib):
    """the detection-glue / loss wrappers of mrb_b200.ops have no CPU path: CPU tensors raise instead of falling back"""
    import pytest
    import torch
    from mrb_b200 import ops
    b = torch.zeros(2, 10, 4)
    with pytest.raises(RuntimeError):
        ops.rpn_collect(torch.zeros(20, 4), torch.zeros(20), torch.zeros(20, dtype=torch.int64), torch.zeros(2, dtype=torch.int32),
                        [10], 2, 10, 10, True)
    with pytest.raises(RuntimeError):
        ops.roi_assign_sample(b, torch.ones(2, 10, dtype=torch.bool), torch.rand(2, 10), torch.zeros(2, 1, 4),
                              torch.ones(2, 1, dtype=torch.int64), torch.ones(2, dtype=torch.int32), 8, 0.25, 0.5, 0.5, (10, 10, 5, 5))
    with pytest.raises(RuntimeError):
        ops.rpn_anchor_match(torch.zeros(16, 4), torch.zeros(2, 1, 4), torch.ones(2, dtype=torch.int32), torch.ones(2), torch.ones(2),
                             0.7, 0.3, 0.0)
    with pytest.raises(RuntimeError):
        ops.box_head_loss(torch.zeros(4, 408), torch.zeros(4, dtype=torch.int64), torch.zeros(4, 4), 81)
    with pytest.raises(RuntimeError):
        ops.rpn_topk_decode(torch.zeros(1, 2, 2, 16), 3, torch.zeros(12, 4), 4, torch.ones(1), torch.ones(1), torch.zeros(1, 4, 4),
                            torch.zeros(1, 4))
    with pytest.raises(RuntimeError):
        ops.box_postprocess(torch.zeros(20, 408), 81, b, torch.ones(2, 10, dtype=torch.bool), torch.ones(2), torch.ones(2), 0.05,
                            (10, 10, 5, 5), 0.5, 100)
    with pytest.raises(RuntimeError):
        ops.mask_targets_rect(torch.zeros(3, 4), torch.zeros(3, 5), 28)
