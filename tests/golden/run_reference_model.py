"""Run the UNMODIFIED reference (its own modeling/, structures/, layers/ and -- through oracle/_ref --
its own CPU csrc kernels) on a tiny R-50-FPN-shaped config, CPU, eval mode, and dump weights +
intermediate tensors.  Authoring container only (needs /root/reference).  Used by
tests/test_harness_vs_reference.py (live) and to produce tests/golden/harness_tiny.pt."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MRB_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "_shims"))

import mrb_test_compat  # noqa: E402,F401
import torch  # noqa: E402

import oracle  # noqa: E402

TINY = ["MODEL.RESNETS.STEM_OUT_CHANNELS", 8, "MODEL.RESNETS.WIDTH_PER_GROUP", 8, "MODEL.RESNETS.RES2_OUT_CHANNELS", 32,
        "MODEL.RESNETS.BACKBONE_OUT_CHANNELS", 32, "MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM", 64,
        "MODEL.RPN.PRE_NMS_TOP_N_TEST", 100, "MODEL.RPN.POST_NMS_TOP_N_TEST", 100, "MODEL.RPN.FPN_POST_NMS_TOP_N_TEST", 100,
        "MODEL.DEVICE", "cpu"]


def main(out_path):
    pkg = types.ModuleType("maskrcnn_benchmark")
    pkg.__path__ = [os.path.join(REF, "maskrcnn_benchmark")]
    sys.modules["maskrcnn_benchmark"] = pkg
    refc = oracle.ref()
    assert refc is not None
    pkg._C = refc
    sys.modules["maskrcnn_benchmark._C"] = refc
    from maskrcnn_benchmark.config import cfg
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    from maskrcnn_benchmark.structures.image_list import to_image_list
    cfg.merge_from_file(os.path.join(REF, "configs", "e2e_faster_rcnn_R_50_FPN_1x.yaml"))
    cfg.merge_from_list(TINY)
    cfg.freeze()
    torch.manual_seed(0)
    model = build_detection_model(cfg).eval()
    # make the frozen BN non-trivial and scores non-degenerate
    g = torch.Generator().manual_seed(1)
    for n, b in model.named_buffers():
        if n.endswith("running_var"):
            b.copy_(torch.rand(b.shape, generator=g) * 1.5 + 0.5)
        elif n.endswith("running_mean") or (n.endswith(".bias") and "bn" in n):
            b.copy_(torch.randn(b.shape, generator=g) * 0.1)
        elif n.endswith(".weight") and ("bn" in n or "downsample.1" in n):
            b.copy_((torch.rand(b.shape, generator=g) * 0.4 + 0.3))
    for n, p in model.named_parameters():
        if "cls_score" in n or "bbox_pred" in n or "cls_logits" in n:
            p.data.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.5))
    images = [torch.randn(3, 128, 150, generator=g) * 40, torch.randn(3, 120, 160, generator=g) * 40]
    il = to_image_list(images, 32)
    with torch.no_grad():
        feats = model.backbone(il.tensors)
        proposals, _ = model.rpn(il, feats, None)
        dets = model(il)
    out = {
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
        "images": il.tensors.clone(), "image_sizes": [tuple(s) for s in il.image_sizes],
        "feats": [f.clone() for f in feats],
        "proposals": [(p.bbox.clone(), p.get_field("objectness").clone()) for p in proposals],
        "dets": [(d.bbox.clone(), d.get_field("scores").clone(), d.get_field("labels").clone()) for d in dets],
    }
    torch.save(out, out_path)
    print("saved", out_path, [f.shape for f in feats], [len(p) for p in proposals], [len(d) for d in dets])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "harness_tiny.pt"))
