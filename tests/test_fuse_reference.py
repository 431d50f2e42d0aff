"""mrb_b200.fuse.fuse_model over the UNMODIFIED reference GeneralizedRCNN (its modeling/, structures/, losses,
samplers; this repository's layers), CPU checker backend: the fused graph reproduces the reference's own train-mode
forward -- all five losses -- and every parameter gradient of e2e_mask_rcnn_R_50_FPN_1x (width-reduced), with the
same sampled proposals.  This is the train-step pin of the fusion pass against the reference.  Needs a reference
checkout (baseline/_ref mirror or /root/reference): skipped elsewhere."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_ref():
    sys.path.insert(0, os.path.join(ROOT, "maskrcnn-benchmark_b200"))
    from mrb_b200 import refenv
    return refenv.find_reference_root() is not None


@pytest.mark.parametrize("cfgname", ["e2e_mask_rcnn_R_50_FPN_1x.yaml", "e2e_mask_rcnn_X_101_32x8d_FPN_1x.yaml",
                                     "e2e_faster_rcnn_R_50_FPN_1x.yaml"])
def test_fused_reference_graph_equals_reference_train_step(built_lib, oracle_mod, cfgname):
    if not _have_ref():
        pytest.skip("reference checkout absent")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refgraph", "run_cpu.py"), cfgname],
                       capture_output=True, text=True, timeout=900, cwd=os.path.join(ROOT, "tests", "refgraph"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    f = out["report"]["fused"]
    assert f.get("stem") == 1 and f.get("fpn") == 1 and f.get("rpn_head") == 1 and f.get("box_head") == 1
    assert f.get("mask_head", 0) == (0 if "faster" in cfgname else 1)
    assert sum(v for k, v in f.items() if k.startswith("bottleneck")) == (33 if "X_101" in cfgname else 16)
    assert not out["report"]["skipped"]
    assert out["n_grads"] > 60 and out["worst_rel_grad"] < 1e-3


def test_glue_bindings_fall_back_to_the_reference_on_cpu(built_lib, oracle_mod):
    """fuse_model's detection-glue bindings (RPNPostProcessor, box PostProcessor, RPN loss targets, box subsample, mask target
    preparation, project_masks_on_boxes) have no CPU path: on CPU tensors every one of them must hand over to the reference's own
    code, so the train step and the detections stay exactly the unfused reference's."""
    if not _have_ref():
        pytest.skip("reference checkout absent")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refgraph", "run_cpu.py"), "e2e_mask_rcnn_R_50_FPN_1x.yaml",
                        "--glue-bindings"], capture_output=True, text=True, timeout=900, cwd=os.path.join(ROOT, "tests", "refgraph"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    f = out["report"]["fused"]
    for k, v in (("rpn_postprocessor", 2), ("box_postprocessor", 1), ("rpn_loss_targets", 1), ("box_subsample", 1),
                 ("mask_prepare_targets", 1), ("mask_targets", 1)):
        assert f.get(k) == v, f
    assert out["n_grads"] > 60 and out["worst_rel_grad"] < 1e-3
