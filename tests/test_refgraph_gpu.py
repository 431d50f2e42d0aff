"""Boundary acceptance on the B200: the UNMODIFIED reference GeneralizedRCNN (RPN, ROI heads, losses, samplers from
the reference tree mirrored under baseline/_ref) runs its train step over this repository's `layers` / `_C`, unfused
(each layers.Conv2d -> tcgen05 engine) and after mrb_b200.fuse.fuse_model(), and matches the fp32 CPU checker (the same
reference graph on ATen + oracle kernels) on FPN features, all five losses and the parameter gradients."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def test_reference_generalized_rcnn_over_layers_on_gpu(built_lib, oracle_mod):
    sys.path.insert(0, os.path.join(ROOT, "maskrcnn-benchmark_b200"))
    from mrb_b200 import refenv
    if refenv.find_reference_root() is None:
        pytest.skip("reference mirror absent (baseline/_ref is created by build() where /root/reference exists)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refgraph", "run_gpu.py")],
                       capture_output=True, text=True, timeout=850, cwd=os.path.join(ROOT, "tests", "refgraph"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "refgraph_gpu.json"), "w"), indent=1)
    # every conv of the unfused reference graph reached the engine (the 7x7 stem included), none fell back to ATen
    assert out["engine_calls_unfused"] >= 60 and out["aten_fallbacks_unfused"] == 0, out
    f = out["report"]["fused"]
    assert f.get("stem") == 1 and f.get("fpn") == 1 and f.get("rpn_head") == 1 and f.get("box_head") == 1 and f.get("mask_head") == 1
    assert f.get("bottleneck[fn]") == 16 and not out["report"]["skipped"]
    assert out["libmrb_launches"]["fused"] > 100
    for tag in ("unfused", "fused"):
        # bf16 operands / activations vs the fp32 checker
        assert max(out["feat_rel_err"][tag]) < 3e-2, out["feat_rel_err"]
        lc, lg = out["losses"]["checker"], out["losses"][tag]
        for k in lc:
            assert abs(lc[k] - lg[k]) <= 5e-2 * max(1.0, abs(lc[k])), (tag, k, lc[k], lg[k])
        assert out["grad_rel_err"][tag]["median"] < 5e-2, out["grad_rel_err"][tag]
