"""Generate tests/golden/*.npz from the REFERENCE itself (run in the authoring container only).

 1. nms_reference_tests.npz : the two known-answer vectors of the reference's tests/test_nms.py
    (:11-58 and :60-217), captured by running those test functions with a recording `nms`
    (implemented by oracle/_ref, i.e. the reference's own nms_cpu.cpp) and a recording
    numpy.testing.assert_array_equal -- so inputs AND expected outputs are the reference's.
 2. nms_ref_random.npz      : reference nms_cpu outputs on seeded clustered boxes.
 3. roi_align_ref.npz       : reference ROIAlign_forward_cpu outputs on seeded inputs (small case in
    full; BASELINE config 1 -- 1x256x200x336, 100 boxes, 7x7, S=2, scale 0.25 -- subsampled).
Inputs that are too large to commit are regenerated from the seed by tests/_inputs.py.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("MRB_REFERENCE", "/root/reference")

import oracle  # noqa: E402
import _inputs  # noqa: E402

oracle.build()
ref = oracle.ref()
assert ref is not None, "oracle/_ref missing: needs the reference tree"


def capture_reference_nms_tests():
    calls, expected = [], []

    def rec_nms(boxes, scores, thr):
        out = ref.nms(boxes, scores, float(thr))
        calls.append((boxes.numpy().copy(), scores.numpy().copy(), float(thr)))
        return out

    pkg = types.ModuleType("maskrcnn_benchmark")
    layers = types.ModuleType("maskrcnn_benchmark.layers")
    layers.nms = rec_nms
    pkg.layers = layers
    saved = {k: sys.modules.get(k) for k in ("maskrcnn_benchmark", "maskrcnn_benchmark.layers")}
    sys.modules["maskrcnn_benchmark"] = pkg
    sys.modules["maskrcnn_benchmark.layers"] = layers
    orig = np.testing.assert_array_equal

    def rec_assert(a, b, *args, **kw):
        expected.append(np.asarray(b).astype(np.int64).copy())
        return orig(a, b, *args, **kw)

    np.testing.assert_array_equal = rec_assert
    try:
        spec = importlib.util.spec_from_file_location("ref_test_nms", os.path.join(REF, "tests", "test_nms.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        t = mod.TestNMS()
        t.test_nms_cpu()
        t.test_nms1_cpu()
    finally:
        np.testing.assert_array_equal = orig
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert len(calls) == len(expected) == 6
    out = {"n": np.array(len(calls))}
    for i, ((b, s, thr), e) in enumerate(zip(calls, expected)):
        out["boxes%d" % i], out["scores%d" % i], out["thr%d" % i], out["keep%d" % i] = b, s, np.float32(thr), e
    np.savez_compressed(os.path.join(HERE, "nms_reference_tests.npz"), **out)


def ref_random_nms():
    out = {}
    cases = [(819, 0.7, 0), (1000, 0.7, 1), (2000, 0.7, 2), (2000, 0.5, 3), (6000, 0.7, 4), (257, 0.3, 5)]
    out["cases"] = np.array(cases, dtype=np.float64)
    for i, (n, thr, seed) in enumerate(cases):
        boxes, scores = _inputs.nms_boxes(n, seed)
        keep = ref.nms(boxes, scores, float(thr))
        out["keep%d" % i] = keep.numpy()
    np.savez_compressed(os.path.join(HERE, "nms_ref_random.npz"), **out)


def ref_roi_align():
    out = {}
    # small, stored in full
    feat, rois = _inputs.roi_align_small()
    for tag, (p, s) in {"7x7s2": (7, 2), "14x14s2": (14, 2), "7x7s0": (7, 0), "3x5s1": ((3, 5), 1)}.items():
        ph, pw = (p, p) if isinstance(p, int) else p
        out["small_" + tag] = ref.roi_align_forward(feat, rois, 0.25, ph, pw, s).numpy()
    # BASELINE config 1, subsampled
    feat, rois = _inputs.roi_align_config1()
    y = ref.roi_align_forward(feat, rois, 0.25, 7, 7, 2).numpy().reshape(-1)
    out["config1_stride"] = np.array(97)
    out["config1_samples"] = y[::97].copy()
    out["config1_sum"] = np.array(y.astype(np.float64).sum())
    np.savez_compressed(os.path.join(HERE, "roi_align_ref.npz"), **out)


if __name__ == "__main__":
    capture_reference_nms_tests()
    ref_random_nms()
    ref_roi_align()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
