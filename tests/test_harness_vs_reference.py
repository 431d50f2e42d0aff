"""The harness model (mrb_b200.model) against the UNMODIFIED reference model, same weights, same
images, CPU, eval mode.  The reference runs in a subprocess with its own modeling/, layers/ and (via
oracle/_ref) its own csrc CPU kernels; the harness runs on the CPU checker backend.  This pins the
harness's host logic (anchors, decode, clip, NMS order, level mapping, pooling, post-processing) and
proves state_dict compatibility (strict load of the reference's state_dict).  Needs /root/reference:
skipped elsewhere; tests/golden/harness_tiny.pt carries the same check to the GPU box."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MRB_REFERENCE", "/root/reference")


def tiny_cfg():
    from mrb_b200.model import RCNNConfig
    return RCNNConfig(mask_on=False, stem_out=8, width_per_group=8, res2_out=32, fpn_out=32, mlp_head_dim=64,
                      pre_nms_top_n_test=100, post_nms_top_n_test=100, fpn_post_nms_top_n_test=100)


@pytest.fixture(scope="module")
def ref_dump(tmp_path_factory, built_lib):
    if not os.path.isdir(os.path.join(REF, "maskrcnn_benchmark")):
        pytest.skip("reference checkout absent")
    import oracle
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built")
    out = str(tmp_path_factory.mktemp("ref") / "ref_tiny.pt")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "run_reference_model.py"), out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    return torch.load(out)


def test_state_dict_keys_and_outputs_match_reference(ref_dump):
    from oracle.cpu_backend import CpuCheckerBackend
    from mrb_b200.model import GeneralizedRCNN
    d = ref_dump
    model = GeneralizedRCNN(tiny_cfg(), CpuCheckerBackend()).eval()
    missing, unexpected = model.load_state_dict(d["state_dict"], strict=True)
    assert not missing and not unexpected
    be = model.be
    with torch.no_grad():
        feats = model.backbone.run(be, d["images"])
        for a, b in zip(feats, d["feats"]):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
        (boxes, scores, valid), _ = model.rpn.run(be, feats, d["image_sizes"], None, False)
        for i, (rb, rs) in enumerate(d["proposals"]):
            n = int(valid[i].sum())
            assert n == len(rb)
            torch.testing.assert_close(boxes[i][:n], rb, rtol=1e-4, atol=1e-3)
            torch.testing.assert_close(scores[i][:n], rs, rtol=1e-4, atol=1e-5)
        dets = model(d["images"], d["image_sizes"])
    for mine, (rb, rs, rl) in zip(dets, d["dets"]):
        assert len(mine["boxes"]) == len(rb)
        o1, o2 = torch.argsort(mine["scores"], descending=True), torch.argsort(rs, descending=True)
        torch.testing.assert_close(mine["scores"][o1], rs[o2], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(mine["boxes"][o1], rb[o2], rtol=1e-4, atol=1e-2)
        assert torch.equal(mine["labels"][o1], rl[o2])


def test_full_size_state_dict_keys_match_reference(built_lib):
    """e2e_mask_rcnn_R_50_FPN_1x: identical key set and shapes (checkpoint compatibility)."""
    if not os.path.isdir(os.path.join(REF, "maskrcnn_benchmark")):
        pytest.skip("reference checkout absent")
    code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests", "_shims")); sys.path.insert(0, os.path.join(%r, "maskrcnn-benchmark_b200"))
import mrb_test_compat, torch
os.environ["MRB_REFERENCE_ROOT"] = %r
from maskrcnn_benchmark.config import cfg
from maskrcnn_benchmark.modeling.detector import build_detection_model
cfg.merge_from_file(os.path.join(%r, "configs", "e2e_mask_rcnn_R_50_FPN_1x.yaml"))
cfg.merge_from_list(["MODEL.DEVICE", "cpu"])
ref = {k: tuple(v.shape) for k, v in build_detection_model(cfg).state_dict().items()}
from mrb_b200.model import RCNNConfig, GeneralizedRCNN
from mrb_b200.model.backend import Backend
mine = {k: tuple(v.shape) for k, v in GeneralizedRCNN(RCNNConfig(), Backend()).state_dict().items()}
assert ref == mine, (sorted(set(ref) ^ set(mine))[:10], [k for k in ref if k in mine and ref[k] != mine[k]][:10])
print("OK", len(ref))
''' % (ROOT, ROOT, ROOT, REF, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, MRB_REFERENCE_ROOT=REF))
    assert r.returncode == 0, r.stdout + r.stderr
