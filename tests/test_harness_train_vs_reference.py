"""Train-mode pin of the harness model against the UNMODIFIED reference model (CPU, same weights, deterministic
samplers): all five losses to 2e-4, every parameter gradient to 2e-3 of its scale.  Complements
tests/test_harness_vs_reference.py (eval mode) and tests/test_fuse_reference.py (the fused reference graph)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_harness_train_step_equals_reference(built_lib, oracle_mod):
    sys.path.insert(0, os.path.join(ROOT, "maskrcnn-benchmark_b200"))
    from mrb_b200 import refenv
    if refenv.find_reference_root() is None:
        pytest.skip("reference checkout absent")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "refgraph", "run_harness_train_cpu.py")],
                       capture_output=True, text=True, timeout=900, cwd=os.path.join(ROOT, "tests", "refgraph"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_grads"] > 60 and out["worst_rel_grad"][0] < 2e-3
    assert set(out["losses_harness"]) == {"loss_objectness", "loss_rpn_box_reg", "loss_classifier", "loss_box_reg", "loss_mask"}
