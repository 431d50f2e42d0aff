"""Shared by tests/test_fuse_reference.py (CPU) and tests/test_refgraph_gpu.py (B200): build the UNMODIFIED
reference `build_detection_model(cfg)` on top of this repository's `maskrcnn_benchmark.layers` / `_C`
(mrb_b200.refenv), at a width-reduced e2e_mask_rcnn_R_50_FPN_1x, with deterministic inputs and sampling.

Test infrastructure: CPU tensors reaching `_C` are served by the oracle (the product `_C` has no CPU path and
raises); CUDA tensors always go to libmrb_b200.so."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

TINY = ["MODEL.RESNETS.STEM_OUT_CHANNELS", 8, "MODEL.RESNETS.WIDTH_PER_GROUP", 8, "MODEL.RESNETS.RES2_OUT_CHANNELS", 32,
        "MODEL.RESNETS.BACKBONE_OUT_CHANNELS", 32, "MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM", 64,
        "MODEL.ROI_MASK_HEAD.CONV_LAYERS", (16, 16, 16, 16),
        "MODEL.RPN.PRE_NMS_TOP_N_TRAIN", 200, "MODEL.RPN.POST_NMS_TOP_N_TRAIN", 200, "MODEL.RPN.FPN_POST_NMS_TOP_N_TRAIN", 200,
        "MODEL.RPN.PRE_NMS_TOP_N_TEST", 100, "MODEL.RPN.POST_NMS_TOP_N_TEST", 100, "MODEL.RPN.FPN_POST_NMS_TOP_N_TEST", 100,
        "MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE", 64]
IMAGE_HW = [(128, 150), (120, 160)]


def activate():
    from mrb_b200 import refenv
    root = refenv.activate()
    if root is None:
        raise RuntimeError("no reference checkout (MRB_REFERENCE_ROOT, baseline/_ref, /root/reference)")
    return root


def route_cpu_C_to_oracle():
    """CPU tensors -> oracle kernels, CUDA tensors -> the product.  (Checker plumbing, tests only.)"""
    import oracle
    from maskrcnn_benchmark import _C
    real = {n: getattr(_C, n) for n in ("nms", "roi_align_forward", "roi_align_backward")}

    def nms(d, s, t):
        return real["nms"](d, s, t) if d.is_cuda else oracle.nms(d.contiguous(), s.contiguous(), t)

    def raf(x, r, sc, ph, pw, s):
        return real["roi_align_forward"](x, r, sc, ph, pw, s) if x.is_cuda else \
            oracle.roi_align_forward(x.contiguous(), r.contiguous(), sc, ph, pw, s)

    def rab(g, r, sc, ph, pw, b, c, h, w, s):
        return real["roi_align_backward"](g, r, sc, ph, pw, b, c, h, w, s) if g.is_cuda else \
            oracle.roi_align_backward(g.contiguous(), r.contiguous(), sc, ph, pw, b, c, h, w, s)
    _C.nms, _C.roi_align_forward, _C.roi_align_backward = nms, raf, rab


def deterministic_randperm():
    """The reference samples with torch.randperm on the tensors' device (balanced_positive_negative_sampler.py:
    49-50); CPU and CUDA generators differ, so draw every permutation from one CPU generator."""
    orig = torch.randperm
    state = {"g": torch.Generator().manual_seed(1234)}

    def randperm(n, *a, device=None, **kw):
        return orig(n, generator=state["g"]).to(device if device is not None else "cpu")

    torch.randperm = randperm

    def reseed(seed=1234):
        state["g"] = torch.Generator().manual_seed(seed)
    return reseed


def build(config="e2e_mask_rcnn_R_50_FPN_1x.yaml", opts=(), device="cpu", tiny=True):
    from mrb_b200 import refenv
    from maskrcnn_benchmark.config import cfg as _cfg
    from maskrcnn_benchmark.modeling.detector import build_detection_model
    cfg = _cfg.clone()
    cfg.merge_from_file(refenv.config_path(config))
    cfg.merge_from_list((TINY if tiny else []) + list(opts) + ["MODEL.DEVICE", "cpu"])
    cfg.freeze()
    torch.manual_seed(0)
    model = build_detection_model(cfg)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, b in model.named_buffers():      # non-trivial frozen BN statistics
            if n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) * 1.5 + 0.5)
            elif n.endswith("running_mean") or (n.endswith(".bias") and ("bn" in n or "downsample.1" in n)):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif n.endswith(".weight") and ("bn" in n or "downsample.1" in n):
                b.copy_(torch.rand(b.shape, generator=g) * 0.4 + 0.3)
    return model.to(device), cfg


def inputs(device="cpu"):
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    from maskrcnn_benchmark.structures.image_list import to_image_list
    from maskrcnn_benchmark.structures.segmentation_mask import SegmentationMask
    g = torch.Generator().manual_seed(2)
    images = [torch.randn(3, h, w, generator=g) * 40 for h, w in IMAGE_HW]
    il = to_image_list(images, 32).to(device)
    targets = []
    for (h, w) in IMAGE_HW:
        bx = torch.tensor([[10., 12., 70., 90.], [60., 30., 140., 100.], [5., 60., 50., 110.]])
        t = BoxList(bx, (w, h), mode="xyxy")
        t.add_field("labels", torch.tensor([3, 17, 60]))
        polys = [[[float(b[0]), float(b[1]), float(b[2]), float(b[1]), float(b[2]), float(b[3]), float(b[0]), float(b[3])]]
                 for b in bx]
        t.add_field("masks", SegmentationMask(polys, (w, h), mode="poly"))
        targets.append(t.to(device))
    return il, targets


def train_step(model, il, targets, reseed):
    reseed()
    torch.manual_seed(7)
    model.zero_grad()
    losses = model(il, targets)
    sum(losses.values()).backward()
    return ({k: float(v.detach()) for k, v in losses.items()},
            {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.grad is not None})
