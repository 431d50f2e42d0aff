"""Fusion pass over the reference's UNMODIFIED module graph.

`fuse_model(model)` walks an instantiated reference model (maskrcnn_benchmark.modeling.detector.
build_detection_model(cfg), running on top of this repository's `layers` / `_C`) and rebinds the `forward`
of the modules whose bodies are chains the conv engine fuses into one launch each:

  reference code (file:line)                                           becomes
  -------------------------------------------------------------------  ------------------------------------------
  BaseStem.forward           conv1 -> bn1 -> relu_ -> max_pool2d       stem conv (4x1 over s2d view) + BN + ReLU,
    (backbone/resnet.py:359-366)                                        NHWC max-pool kernel
  Bottleneck.forward         (conv -> FrozenBN -> relu_) x2,           4 fused launches forward; hand-scheduled
    (backbone/resnet.py:324-344)  conv -> FrozenBN, += identity, relu_  backward with ReLU masks / BN scale /
                                                                        residual join in dgrad epilogues
    ... with groups=32 3x3 (X-101, resnet.py:302-311)                   grouped tcgen05 conv (block-diagonal super-groups)
    ... with DFConv2d 3x3 (configs/dcn, resnet.py:286-300)              offset conv + sampler + tcgen05 GEMM (mrb_b200.dcn)
  FPN.forward                inner 1x1 + F.interpolate + add, 3x3,     lateral conv with the 2x-upsampled top-down
    (backbone/fpn.py:43-76)    LastLevelMaxPool                         map read in the epilogue; 3x3; pool
  RPNHead.forward            relu(conv3x3), cls_logits, bbox_pred      3x3+bias+ReLU; ONE 1x1 with Cout = A + 4A
    (rpn/rpn.py:98-105)
  Pooler.forward             LevelMapper + 4 ROIAlign + index_put      one multi-level ROIAlign launch
    (poolers.py:91-121)
  FPN2MLPFeatureExtractor    pooler, relu(fc6), relu(fc7)              pool + two GEMMs with bias+ReLU epilogues
    (box_head/roi_box_feature_extractors.py:74-81)
  FPNPredictor               cls_score, bbox_pred                      one GEMM, fp32 logits
    (box_head/roi_box_predictors.py:55-62)
  MaskRCNNFPNFeatureExtractor  pooler, relu(conv3x3) x4                pool (NHWC) + fused convs
    (mask_head/roi_mask_feature_extractors.py:59-65)
  MaskRCNNC4Predictor        relu(conv5_mask), mask_fcn_logits         deconv as two strided 1x1 convs, 1x1 fp32
    (mask_head/roi_mask_predictors.py:29-32)

  RPNPostProcessor.forward   per (level, image): top-k, decode, clip,      per level top-k + decode launch, ONE batched NMS,
    (rpn/inference.py:76-181)  NMS (sync), BoxList ops; top-k over levels    ONE selection launch, one sync (detect_glue.cu)
  PostProcessor.forward      per (image, class): nonzero, decode, NMS      softmax+decode+threshold launch, ONE batched NMS
    (box_head/inference.py:45-149)  (sync), cat, kthvalue on the host         over images x classes, top-k launch, one sync
  project_masks_on_boxes     host loop: crop, resize, rasterise, upload    one launch on the polygon vertices
    (mask_head/loss.py:11-42)
  RPNLossComputation.__call__  per image: IoU 268k x G, Matcher, dense     anchor labelling (2 launches) + sampling/encode
    target side (rpn/loss.py:56-110)  encode, nonzero, randperm            (1 cluster launch) for the batch; loss arithmetic kept
  FastRCNNLossComputation.subsample  per image: IoU, Matcher, encode,      one launch for the batch, one sync
    (box_head/loss.py:82-118)          sampler, nonzero                    (the last five only with backend.fused_glue)

Nothing else changes: module classes, parameters, buffers and state_dict keys are the reference's; the anchor
generator, matcher/sampler, box coder and all losses are the reference's own Python.
Modules are recognised by structure (attribute names / layer types), not by import, so the pass needs no reference
checkout; anything that does not match exactly is left alone (it still reaches the engine per conv through
layers.Conv2d, unfused).  Activations between fused modules travel as bf16 NHWC (`channels_last`); tensors
handed back to unfused reference code (RPN logits, box/mask logits) are fp32.

Backends: B200Backend (product).  The CPU checker backend of oracle/ implements the same interface, which is how
tests pin this pass against the unfused reference forward on CPU."""
import types

import torch
from torch import nn

from mrb_b200.model.backbone import Bottleneck as _HBottleneck
from mrb_b200.model.backbone import FrozenAffine


# ---------------------------------------------------------------------------------- recognition
def _is_fbn(m):
    return type(m).__name__ == "FrozenBatchNorm2d" and hasattr(m, "scale_shift")


def _conv_ok(m, k, stride=None, pad=None, bias=None):
    if not isinstance(m, nn.Conv2d) or isinstance(m, nn.ConvTranspose2d):
        return False
    if m.groups != 1 or tuple(m.dilation) != (1, 1) or m.padding_mode != "zeros" or isinstance(m.padding, str):
        return False
    if tuple(m.kernel_size) != (k, k) or m.stride[0] != m.stride[1] or m.padding[0] != m.padding[1]:
        return False
    if stride is not None and m.stride[0] not in (stride if isinstance(stride, tuple) else (stride,)):
        return False
    if pad is not None and m.padding[0] != pad:
        return False
    if bias is not None and (m.bias is not None) != bias:
        return False
    return m.in_channels % 8 == 0


def _act(be, x):
    """Activation entering a fused module: the backend's dtype / memory format (bf16 NHWC on the B200)."""
    if x.dtype != be.act_dtype:
        x = x.to(be.act_dtype)
    if getattr(be, "channels_last", False) and x.dim() == 4 and not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    return x


def _bind(mod, fn):
    mod.forward = types.MethodType(fn, mod)
    mod._mrb_fused = True


# ---------------------------------------------------------------------------------- backbone
def _fuse_stem(mod, be, rep):
    c, b = getattr(mod, "conv1", None), getattr(mod, "bn1", None)
    if not (isinstance(c, nn.Conv2d) and _is_fbn(b)) or tuple(c.weight.shape[1:]) != (3, 7, 7) or c.bias is not None \
            or tuple(c.stride) != (2, 2) or tuple(c.padding) != (3, 3) or c.weight.shape[0] % 8:
        return False
    if any(p.requires_grad for p in mod.parameters()):
        rep["skipped"].append("stem is trainable (the fused stem has no backward)")
        return False
    aff = FrozenAffine(b)

    def forward(self, x):
        a, s = aff.get()
        y = be.stem(x, self.conv1.weight, a, s)
        return be.max_pool(y, 3, 2, 1)
    _bind(mod, forward)
    return True


def _bottleneck_kind(mod, be):
    """'fn' (whole block as one autograd node), 'conv' (per-conv fusion) or None."""
    need = ("conv1", "bn1", "conv2", "bn2", "conv3", "bn3")
    if not all(hasattr(mod, n) for n in need) or not hasattr(mod, "downsample"):
        return None
    if not all(_is_fbn(getattr(mod, n)) for n in ("bn1", "bn2", "bn3")):
        return None
    if not (_conv_ok(mod.conv1, 1, (1, 2), 0, False) and _conv_ok(mod.conv3, 1, 1, 0, False)):
        return None
    ds = mod.downsample
    if ds is not None and not (isinstance(ds, nn.Sequential) and len(ds) == 2 and _conv_ok(ds[0], 1, (1, 2), 0, False)
                               and _is_fbn(ds[1])):
        return None
    c2 = mod.conv2
    if _conv_ok(c2, 3, 1, 1, False):
        return "fn" if hasattr(be, "bottleneck") else "conv"
    if _conv_ok(c2, 3, 2, 1, False) and getattr(be, "stride2_3x3", True):
        return "conv"
    if hasattr(be, "bottleneck_general") and be.bottleneck_general_ok(mod):
        return "general"
    return None


def _stride_of(m):
    s = m.stride
    return int(s if isinstance(s, int) else s[0])


def _fuse_bottleneck(mod, be, kind):
    mod.strides = (_stride_of(mod.conv1), _stride_of(mod.conv2), _stride_of(mod.downsample[0]) if mod.downsample is not None else 1)
    mod._aff = [FrozenAffine(b) for b in (mod.bn1, mod.bn2, mod.bn3)]
    mod._aff_d = FrozenAffine(mod.downsample[1]) if mod.downsample is not None else None
    mod._g_premasked = False                      # decided by _wire_resnet once the consumers are known
    mod._mrb_premasks_input = kind == "fn"        # _BottleneckFn returns grad_x already multiplied by [x > 0]
    if kind == "general":
        def forward(self, x):
            return be.bottleneck_general(self, _act(be, x))
    else:
        def forward(self, x):
            return _HBottleneck.run(self, be, _act(be, x))
    _bind(mod, forward)


def _fuse_fpn(mod, be, rep):
    inner, layer = getattr(mod, "inner_blocks", None), getattr(mod, "layer_blocks", None)
    if not (isinstance(inner, list) and isinstance(layer, list) and inner and len(inner) == len(layer)):
        return False
    top = getattr(mod, "top_blocks", None)
    if top is not None and type(top).__name__ != "LastLevelMaxPool":
        return False
    for a, b in zip(inner, layer):
        if not (_conv_ok(getattr(mod, a, None), 1, 1, 0, True) and _conv_ok(getattr(mod, b, None), 3, 1, 1, True)):
            return False
    mod._mrb_premask_inputs = False               # set by _wire_resnet when the producer is a fused ResNet

    def forward(self, x):
        if len(x) != len(self.inner_blocks):
            raise RuntimeError("fused FPN: expected %d feature maps, got %d" % (len(self.inner_blocks), len(x)))
        pm = self._mrb_premask_inputs
        last, results = None, []
        for feat, ib, lb in zip(list(x)[::-1], self.inner_blocks[::-1], self.layer_blocks[::-1]):
            ib, lb = getattr(self, ib), getattr(self, lb)
            last = be.lateral_topdown(_act(be, feat), ib.weight, ib.bias, last, premask_x=pm)
            results.insert(0, be.conv(last, lb.weight, bias=lb.bias, pad=1))
        if self.top_blocks is not None:
            results.append(be.max_pool(results[-1], 1, 2, 0))          # LastLevelMaxPool (fpn.py:77-79)
        return tuple(results)
    _bind(mod, forward)
    return True


def _wire_resnet(model, be):
    """Decide, per fused bottleneck, whether every consumer of its output hands back a gradient already masked by
    [out > 0] (then the block skips its own mask pass), and whether a fused FPN may mask the gradients it returns."""
    for parent in model.modules():
        body, fpn = getattr(parent, "body", None), getattr(parent, "fpn", None)
        if body is None or not hasattr(body, "stages") or not hasattr(body, "return_features"):
            continue
        stages = [list(getattr(body, s)) for s in body.stages]
        returned = [bool(body.return_features[s]) for s in body.stages]
        fpn_fused = fpn is not None and getattr(fpn, "_mrb_fused", False) and hasattr(fpn, "_mrb_premask_inputs") and \
            sum(returned) == len(fpn.inner_blocks)
        all_fused = all(getattr(b, "_mrb_fused", False) for st in stages for b in st)
        if fpn_fused:
            # every FPN input is the ReLU output of a fused bottleneck: masking grad by [x > 0] is exact
            fpn._mrb_premask_inputs = all_fused
        for si, st in enumerate(stages):
            for bi, blk in enumerate(st):
                if not getattr(blk, "_mrb_fused", False):
                    continue
                if bi + 1 < len(st):
                    ok = getattr(st[bi + 1], "_mrb_premasks_input", False)
                else:
                    ok = True
                    if si + 1 < len(stages):
                        ok = ok and getattr(stages[si + 1][0], "_mrb_premasks_input", False)
                    if returned[si]:
                        ok = ok and fpn_fused and all_fused
                    elif si + 1 == len(stages):
                        ok = False             # last stage output leaves the backbone to an unknown consumer
                blk._g_premasked = bool(ok)


# ---------------------------------------------------------------------------------- RPN head
def _fuse_rpn_head(mod, be):
    c, cl, bp = getattr(mod, "conv", None), getattr(mod, "cls_logits", None), getattr(mod, "bbox_pred", None)
    if not (_conv_ok(c, 3, 1, 1, True) and _conv_ok(cl, 1, 1, 0, True) and _conv_ok(bp, 1, 1, 0, True)):
        return False
    if bp.out_channels != 4 * cl.out_channels or len(list(mod.children())) != 3:
        return False

    def forward(self, x):
        a = self.cls_logits.out_channels
        w = torch.cat([self.cls_logits.weight, self.bbox_pred.weight], 0)
        b = torch.cat([self.cls_logits.bias, self.bbox_pred.bias], 0)
        logits, bbox_reg = [], []
        for f in x:
            t = be.conv(_act(be, f), self.conv.weight, bias=self.conv.bias, pad=1, relu=True, gy_premasked=True)
            o = be.conv(t, w, bias=b, out_fp32=True, premask_x=True).float()
            # plain NCHW-contiguous fp32 for the reference's permute_and_flatten / .view chains (rpn/utils.py:11-15)
            logits.append(o[:, :a].contiguous())
            bbox_reg.append(o[:, a:].contiguous())
        return logits, bbox_reg
    _bind(mod, forward)
    return True


# ---------------------------------------------------------------------------------- ROI heads
def _pooler_ok(p):
    if type(p).__name__ != "Pooler" or not hasattr(p, "poolers") or not hasattr(p, "map_levels"):
        return False
    ml = p.map_levels
    scales = [float(r.spatial_scale) for r in p.poolers]
    if len(scales) != 4 or scales != [0.25, 0.125, 0.0625, 0.03125]:
        return False
    if (float(ml.k_min), float(ml.k_max), float(ml.s0), float(ml.lvl0), float(ml.eps)) != (2.0, 5.0, 224.0, 4.0, 1e-6):
        return False
    sr = {int(r.sampling_ratio) for r in p.poolers}
    osz = {tuple(r.output_size) for r in p.poolers}
    return len(sr) == 1 and len(osz) == 1 and len({o for s in osz for o in s}) == 1


def _rois_of(boxes):
    """convert_to_roi_format (poolers.py:72-89)"""
    bb = torch.cat([b.bbox for b in boxes], 0).float()
    ids = torch.cat([torch.full((len(b), 1), float(i), dtype=bb.dtype, device=bb.device) for i, b in enumerate(boxes)], 0)
    return torch.cat([ids, bb], 1)


def _pool(be, pooler, x, boxes, nhwc):
    n = len(pooler.poolers)
    feats = [_act(be, f) for f in list(x)[:n]]
    r0 = pooler.poolers[0]
    return be.roi_align_fpn(feats, _rois_of(boxes), [float(r.spatial_scale) for r in pooler.poolers],
                            int(r0.output_size[0]), int(r0.sampling_ratio), nhwc)


def _fuse_pooler(mod, be):
    def forward(self, x, boxes):
        dt = x[0].dtype
        return _pool(be, self, x, boxes, False).to(dt)
    _bind(mod, forward)


def _fuse_box_head(fe, pr, be):
    if not (hasattr(fe, "pooler") and _pooler_ok(fe.pooler) and isinstance(getattr(fe, "fc6", None), nn.Linear)
            and isinstance(getattr(fe, "fc7", None), nn.Linear) and fe.fc6.bias is not None and fe.fc7.bias is not None):
        return False
    if not (isinstance(getattr(pr, "cls_score", None), nn.Linear) and isinstance(getattr(pr, "bbox_pred", None), nn.Linear)
            and not hasattr(pr, "avgpool") and pr.cls_score.bias is not None and pr.bbox_pred.bias is not None):
        return False
    if fe.fc6.in_features % 8 or fe.fc6.out_features % 8 or fe.fc7.out_features % 8:
        return False

    def fe_forward(self, x, proposals):
        x = _pool(be, self.pooler, x, proposals, False)
        x = x.flatten(1)
        x = be.linear(x, self.fc6.weight, self.fc6.bias, relu=True, gy_premasked=True)
        return be.linear(x, self.fc7.weight, self.fc7.bias, relu=True, premask_x=True, gy_premasked=True)

    def pr_forward(self, x):
        if x.dim() == 4:
            x = x.flatten(1)
        nc = self.cls_score.out_features
        w = torch.cat([self.cls_score.weight, self.bbox_pred.weight], 0)
        b = torch.cat([self.cls_score.bias, self.bbox_pred.bias], 0)
        o = be.linear(_act(be, x), w, b, relu=False, out_fp32=True, premask_x=True).float()
        return o[:, :nc], o[:, nc:]
    _bind(fe, fe_forward)
    _bind(pr, pr_forward)
    return True


def _fuse_mask_head(fe, pr, be):
    if not (hasattr(fe, "pooler") and _pooler_ok(fe.pooler) and isinstance(getattr(fe, "blocks", None), list) and fe.blocks):
        return False
    if not all(_conv_ok(getattr(fe, n, None), 3, 1, 1, True) for n in fe.blocks):
        return False
    dc, lg = getattr(pr, "conv5_mask", None), getattr(pr, "mask_fcn_logits", None)
    if not (isinstance(dc, nn.ConvTranspose2d) and tuple(dc.kernel_size) == (2, 2) and tuple(dc.stride) == (2, 2)
            and tuple(dc.padding) == (0, 0) and dc.bias is not None and dc.in_channels % 8 == 0 and dc.out_channels % 8 == 0
            and _conv_ok(lg, 1, 1, 0, True)):
        return False

    def fe_forward(self, x, proposals):
        x = _pool(be, self.pooler, x, proposals, True)
        for i, name in enumerate(self.blocks):
            c = getattr(self, name)
            x = be.conv(x, c.weight, bias=c.bias, pad=1, relu=True, premask_x=i > 0, gy_premasked=True)
        return x

    def pr_forward(self, x):
        x = be.deconv2x2(_act(be, x), self.conv5_mask.weight, self.conv5_mask.bias, relu=True, premask_x=True, gy_premasked=True)
        lgc = self.mask_fcn_logits
        return be.conv(x, lgc.weight, bias=lgc.bias, out_fp32=True, premask_x=True).float()
    _bind(fe, fe_forward)
    _bind(pr, pr_forward)
    return True


# ---------------------------------------------------------------------------------- the pass
# ---------------------------------------------------------------------------------- detection glue (opt-in with the backend)
def _fuse_rpn_postprocessor(mod, be):
    """RPNPostProcessor.forward (rpn/inference.py:125-181): per level one top-k + one decode launch, ONE batched NMS launch
    sequence for all (image, level) problems and ONE selection launch (csrc/detect_glue.cu), then a single host
    synchronisation to size the returned BoxLists -- instead of a Python loop over levels x images with an NMS-sizing
    synchronisation and a dozen BoxList operations each.  Same proposals (ties between equal scores aside)."""
    need = ("pre_nms_top_n", "post_nms_top_n", "nms_thresh", "min_size", "box_coder", "fpn_post_nms_top_n", "fpn_post_nms_per_batch")
    if type(mod).__name__ != "RPNPostProcessor" or not all(hasattr(mod, n) for n in need):
        return False
    if not (hasattr(mod.box_coder, "weights") and hasattr(mod.box_coder, "bbox_xform_clip")):
        return False
    orig = mod.forward
    cache = {}

    def forward(self, anchors, objectness, box_regression, targets=None):
        from mrb_b200 import ops
        n, L = len(anchors), len(objectness)
        dev = objectness[0].device
        per_batch = bool(self.training and self.fpn_post_nms_per_batch)
        if self.min_size != 0 or dev.type != "cuda" or not (1 < L <= 8) or (per_batch and n > 8) or \
                (self.training and targets is None):
            return orig(anchors, objectness, box_regression, targets)
        with torch.no_grad():
            sizes = tuple(tuple(a[0].size) for a in anchors)                  # (width, height) of every image
            key = (sizes, str(dev))
            if key not in cache:
                cache[key] = (torch.tensor([float(s[0]) for s in sizes], device=dev), torch.tensor([float(s[1]) for s in sizes], device=dev))
            widths, heights = cache[key]
            lgs, dls, ks = [], [], []
            for o, r in zip(objectness, box_regression):
                _, a, h, w = o.shape
                lgs.append(o.detach().float().permute(0, 2, 3, 1).reshape(n, -1))                             # rpn/utils.py:11-15
                dls.append(r.detach().float().view(n, a, 4, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, 4))
                ks.append(min(int(self.pre_nms_top_n), a * h * w))
            boxes = torch.empty((n * sum(ks), 4), dtype=torch.float32, device=dev)
            scores = torch.empty((n * sum(ks),), dtype=torch.float32, device=dev)
            off = 0
            for l in range(L):
                k = ks[l]
                idx = lgs[l].topk(k, dim=1, sorted=True)[1]                                                   # inference.py:91-95
                ops.rpn_decode(lgs[l], dls[l], anchors[0][l].bbox, idx, widths, heights, boxes[off:off + n * k],
                               scores[off:off + n * k], self.box_coder.weights, self.box_coder.bbox_xform_clip)
                off += n * k
            # rows of every problem are in descending-score order (sorted top-k): the NMS skips its rank sort
            keep, counts = ops.nms_batched(boxes, scores, [k for k in ks for _ in range(n)], float(self.nms_thresh), presorted=True)
            gb = gc = None
            gs = [0] * n
            if self.training and targets is not None:                                                          # add_gt_proposals
                gs = [len(t) for t in targets]
                gmax = max(1, max(gs))
                gb = torch.zeros((n, gmax, 4), dtype=torch.float32, device=dev)
                for i, t in enumerate(targets):
                    if gs[i]:
                        gb[i, :gs[i]] = t.convert("xyxy").bbox
                gc = torch.zeros((n,), dtype=torch.int32, device=dev)
                for i in range(n):                 # scalar fills: a pageable host -> device copy would block until the stream drains
                    gc[i:i + 1].fill_(gs[i])
            b, s, v = ops.rpn_collect(boxes, scores, keep, counts, ks, n, int(self.post_nms_top_n), int(self.fpn_post_nms_top_n),
                                      per_batch, gb, gc)
            wcols = b.shape[1] - (gb.shape[1] if gb is not None else 0)
            nv = v[:, :wcols].sum(1).tolist()                     # the one host synchronisation of the proposal stage
            box_cls = type(anchors[0][0])
            out = []
            for i in range(n):
                bb, ss = b[i, :nv[i]], s[i, :nv[i]]
                if gs[i]:
                    bb = torch.cat([bb, b[i, wcols:wcols + gs[i]]])
                    ss = torch.cat([ss, s[i, wcols:wcols + gs[i]]])
                bl = box_cls(bb, anchors[i][0].size, mode="xyxy")
                bl.add_field("objectness", ss)
                out.append(bl)
            return out
    _bind(mod, forward)
    return True


def _fuse_box_postprocessor(mod, be):
    """PostProcessor.forward (roi_heads/box_head/inference.py:45-149): softmax + per-class decode + clip + threshold in one
    launch, ONE batched NMS over the images x classes problems, one top-detections launch, one host synchronisation to size
    the returned BoxLists -- instead of a Python loop over images x 80 classes with a `nonzero` and an NMS-sizing
    synchronisation each.  Same detections (exact ties at the detections_per_img cut aside: the reference keeps all of them)."""
    need = ("score_thresh", "nms", "detections_per_img", "box_coder", "cls_agnostic_bbox_reg")
    if type(mod).__name__ != "PostProcessor" or not all(hasattr(mod, n) for n in need):
        return False
    if not (hasattr(mod.box_coder, "weights") and hasattr(mod.box_coder, "bbox_xform_clip")):
        return False
    orig = mod.forward
    cache = {}

    def forward(self, x, boxes):
        from mrb_b200 import ops
        class_logits, box_regression = x
        dev = class_logits.device
        n = len(boxes)
        counts = [len(b) for b in boxes]
        p = max(counts) if counts else 0
        c = class_logits.shape[1]
        if dev.type != "cuda" or self.cls_agnostic_bbox_reg or getattr(self, "bbox_aug_enabled", False) or p == 0 or \
                self.detections_per_img <= 0 or box_regression.shape[1] != 4 * c or n * (c - 1) * p >= 2 ** 24:
            return orig(x, boxes)
        with torch.no_grad():
            packed = torch.cat([class_logits.float(), box_regression.float()], 1)                # [R, 5C]
            if all(k == p for k in counts):
                outputs = packed
                props = torch.stack([b.convert("xyxy").bbox for b in boxes]).float()
                valid = torch.ones((n, p), dtype=torch.bool, device=dev)
            else:
                outputs = packed.new_zeros((n * p, packed.shape[1]))
                props = packed.new_zeros((n, p, 4))
                valid = torch.zeros((n, p), dtype=torch.bool, device=dev)
                off = 0
                for i, (b, k) in enumerate(zip(boxes, counts)):
                    outputs[i * p:i * p + k] = packed[off:off + k]
                    props[i, :k] = b.convert("xyxy").bbox
                    valid[i, :k] = True
                    off += k
            sizes = tuple(tuple(b.size) for b in boxes)
            key = (sizes, str(dev))
            if key not in cache:
                cache[key] = (torch.tensor([float(s[0]) for s in sizes], device=dev), torch.tensor([float(s[1]) for s in sizes], device=dev))
            widths, heights = cache[key]
            bb, ss, ll, cc = ops.box_postprocess(outputs.contiguous(), c, props, valid, widths, heights, float(self.score_thresh),
                                                 self.box_coder.weights, float(self.nms), int(self.detections_per_img),
                                                 self.box_coder.bbox_xform_clip)
            ks = cc.tolist()                                # the one host synchronisation of the post-processing
            box_cls = type(boxes[0])
            out = []
            for i in range(n):
                bl = box_cls(bb[i, :ks[i]], boxes[i].size, mode="xyxy")
                bl.add_field("scores", ss[i, :ks[i]])
                bl.add_field("labels", ll[i, :ks[i]])
                out.append(bl)
            return out
    _bind(mod, forward)
    return True


class _FusedRPNLoss:
    """RPNLossComputation.__call__ (rpn/loss.py:92-131) with the target side as three launches: anchor labelling
    (mrb_rpn_anchor_match: IoU, Matcher with low-quality matches, between-threshold discards) + the anchors' own visibility
    field, BalancedPositiveNegativeSampler + BoxCoder.encode of the sampled positives (mrb_rpn_sample; the sample is "the n
    smallest of iid keys", the distribution of the reference's randperm[:n]).  The loss arithmetic on the ~2 x 384 sampled rows
    is the reference's (same ops, same normalisation), gathered from its own concat_box_prediction_layers layout."""

    def __init__(self, orig):
        import sys
        self.orig = orig
        self.concat = getattr(sys.modules.get(type(orig).__module__), "concat_box_prediction_layers")
        for n in ("proposal_matcher", "fg_bg_sampler", "box_coder", "copied_fields", "generate_labels_func", "discard_cases"):
            setattr(self, n, getattr(orig, n))

    def __getattr__(self, name):              # anything else (match_targets_to_anchors, prepare_targets ...) is the reference's
        return getattr(self.orig, name)

    def __call__(self, anchors, objectness, box_regression, targets):
        import torch.nn.functional as F
        from mrb_b200 import ops
        o = self.orig
        dev = objectness[0].device
        n = len(anchors)
        gs = [len(t) for t in targets]
        if dev.type != "cuda" or min(gs) == 0 or n > 64:
            return o(anchors, objectness, box_regression, targets)
        with torch.no_grad():
            anc = torch.cat([a.bbox for a in anchors[0]], 0).float()               # the grid is the same for every image
            vis = torch.stack([torch.cat([a.get_field("visibility") for a in per], 0) for per in anchors]).bool()
            gmax = max(gs)
            gb = torch.zeros((n, gmax, 4), dtype=torch.float32, device=dev)
            gc = torch.zeros((n,), dtype=torch.int32, device=dev)
            for i, t in enumerate(targets):
                gb[i, :gs[i]] = t.convert("xyxy").bbox
                gc[i:i + 1].fill_(gs[i])
            m = self.proposal_matcher
            wh = torch.zeros((n,), dtype=torch.float32, device=dev)     # image sizes: unused, the visibility test is off (-1)
            labels, matched = ops.rpn_anchor_match(anc, gb, gc, wh, wh, float(m.high_threshold), float(m.low_threshold), -1.0)
            labels = torch.where(vis, labels, torch.full((), -1.0, device=dev))    # discard_cases: not_visibility
            keys = torch.rand(labels.shape, device=dev)
            sp = self.fg_bg_sampler
            pos_idx, pos_ok, reg_t, sel_idx, sel_lab, sel_w = ops.rpn_sample(
                labels, matched, keys, anc, gb, int(sp.batch_size_per_image), float(sp.positive_fraction), self.box_coder.weights)
        obj, reg = self.concat(objectness, box_regression)                        # [N * A, 1], [N * A, 4], image-major
        a_tot = anc.shape[0]
        obj = obj.reshape(n, a_tot)
        reg = reg.reshape(n, a_tot, 4)
        num = sel_w.sum().clamp(min=1)
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        diff = torch.abs(torch.where(pos_ok[..., None], torch.gather(reg, 1, pos_idx[..., None].expand(-1, -1, 4)).float() - reg_t, zero))
        beta = 1.0 / 9
        l1 = torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)
        box_loss = torch.where(pos_ok[..., None], l1, zero).sum() / num           # smooth_l1_loss(size_average=False) / #sampled
        bce = F.binary_cross_entropy_with_logits(torch.gather(obj, 1, sel_idx).float(), sel_lab, reduction="none")
        objectness_loss = torch.where(sel_w > 0, bce, zero).sum() / num           # mean over the sampled anchors
        return objectness_loss, box_loss


def _fuse_rpn_loss(parent, be):
    ev = getattr(parent, "loss_evaluator", None)
    if type(ev).__name__ != "RPNLossComputation":
        return False
    need = ("proposal_matcher", "fg_bg_sampler", "box_coder", "discard_cases", "generate_labels_func")
    if not all(hasattr(ev, n) for n in need):
        return False
    m = ev.proposal_matcher
    if not getattr(m, "allow_low_quality_matches", False) or sorted(ev.discard_cases) != ["between_thresholds", "not_visibility"] or \
            getattr(ev.generate_labels_func, "__name__", "") != "generate_rpn_labels":
        return False
    parent.loss_evaluator = _FusedRPNLoss(ev)
    return True


def _fuse_box_subsample(ev, be):
    """FastRCNNLossComputation.subsample (roi_heads/box_head/loss.py:82-118): boxlist_iou + Matcher + labels + regression
    targets + BalancedPositiveNegativeSampler for the batch as ONE launch (mrb_roi_assign_sample), one host synchronisation to
    size the returned BoxLists (the reference synchronises in every `nonzero`).  Returned BoxLists carry the same fields
    (objectness, labels, regression_targets) for the sampled proposals in proposal order."""
    import types
    need = ("proposal_matcher", "fg_bg_sampler", "box_coder", "cls_agnostic_bbox_reg", "subsample")
    if type(ev).__name__ != "FastRCNNLossComputation" or not all(hasattr(ev, n) for n in need):
        return False
    if getattr(ev.proposal_matcher, "allow_low_quality_matches", True) or getattr(ev, "_mrb_fused", False):
        return False
    orig = ev.subsample
    ev._mrb_fused = True

    def subsample(self, proposals, targets):
        from mrb_b200 import ops
        n = len(proposals)
        dev = proposals[0].bbox.device
        counts = [len(p) for p in proposals]
        gs = [len(t) for t in targets]
        pm = max(counts)
        if dev.type != "cuda" or min(gs) == 0 or pm == 0 or pm > 8192 or not all(t.has_field("labels") for t in targets):
            return orig(proposals, targets)
        with torch.no_grad():
            boxes = torch.zeros((n, pm, 4), dtype=torch.float32, device=dev)
            valid = torch.zeros((n, pm), dtype=torch.bool, device=dev)
            gmax = max(gs)
            gb = torch.zeros((n, gmax, 4), dtype=torch.float32, device=dev)
            gl = torch.zeros((n, gmax), dtype=torch.int64, device=dev)
            gc = torch.zeros((n,), dtype=torch.int32, device=dev)
            for i, (p, t) in enumerate(zip(proposals, targets)):
                boxes[i, :counts[i]] = p.convert("xyxy").bbox
                valid[i, :counts[i]] = True
                gb[i, :gs[i]] = t.convert("xyxy").bbox
                gl[i, :gs[i]] = t.get_field("labels")
                gc[i:i + 1].fill_(gs[i])
            keys = torch.rand((n, pm), device=dev)
            m, sp = self.proposal_matcher, self.fg_bg_sampler
            s = int(sp.batch_size_per_image)
            out = ops.roi_assign_sample(boxes, valid, keys, gb, gl, gc, s, float(sp.positive_fraction), float(m.high_threshold),
                                        float(m.low_threshold), self.box_coder.weights, 0, with_index=True)
            ks = (out["labels"] >= 0).sum(1).tolist()               # the one host synchronisation of the sampling
            rois = out["rois"].view(n, s, 5)
            res = []
            for i, p in enumerate(proposals):
                k = ks[i]
                bl = type(p)(rois[i, :k, 1:], p.size, mode="xyxy")
                idx = out["index"][i, :k]
                for f in p.fields():
                    bl.add_field(f, p.get_field(f)[idx])
                bl.add_field("labels", out["labels"][i, :k])
                bl.add_field("regression_targets", out["reg_targets"][i, :k])
                res.append(bl)
        self._proposals = res
        return res
    ev.subsample = types.MethodType(subsample, ev)
    return True


def _fuse_mask_prepare_targets(ev, be):
    """MaskRCNNLossComputation.prepare_targets (roi_heads/mask_head/loss.py:68-98): the reference indexes the targets' BoxList and
    SegmentationMask by the matched indices (Python loops over the positives, twice) before the host-side rasterisation.  Here the
    matching is three small tensor ops, the polygons of ALL instances of the image are packed once (cached per SegmentationMask
    object) and the rasteriser takes the matched instance index of every positive proposal on the device: no per-proposal
    Python, one `nonzero` as in the reference."""
    import types
    import weakref
    if type(ev).__name__ != "MaskRCNNLossComputation" or not all(hasattr(ev, n) for n in ("proposal_matcher", "discretization_size")) \
            or getattr(ev, "_mrb_fused", False):
        return False
    pm = ev.proposal_matcher
    if getattr(pm, "allow_low_quality_matches", True) or float(pm.low_threshold) != float(pm.high_threshold):
        return False
    orig = ev.prepare_targets
    cache = {}

    def polyset_of(seg, dev):
        from mrb_b200 import ops
        key = id(seg)
        hit = cache.get(key)
        if hit is not None and hit[0]() is seg and hit[1].xy.device == dev:
            return hit[1]
        ps = ops.PolygonSet([[p.tolist() for p in pi.polygons] for pi in seg.instances.polygons], dev)
        try:
            cache[key] = (weakref.ref(seg, lambda _r, k=key: cache.pop(k, None)), ps)
        except TypeError:
            pass
        return ps

    def prepare_targets(self, proposals, targets):
        from mrb_b200 import ops
        from mrb_b200.model import box_ops
        labels, masks = [], []
        for p, t in zip(proposals, targets):
            dev = p.bbox.device
            seg = t.get_field("masks") if t.has_field("masks") else None
            if dev.type != "cuda" or len(t) == 0 or getattr(seg, "mode", None) != "poly" or getattr(seg.instances, "polygons", None) is None \
                    or t.mode != "xyxy" or p.mode != "xyxy":
                return orig(proposals, targets)
            with torch.no_grad():
                m = self.proposal_matcher
                q = box_ops.box_iou(t.bbox.float(), p.bbox.float())                 # boxlist_iou, [G, P]
                vals, midx = q.max(dim=0)                                           # matcher.py:60-81, no low-quality pass
                lab = t.get_field("labels").to(torch.int64)[midx]
                lab = torch.where(vals < float(m.low_threshold), torch.zeros_like(lab), lab)
                pos = torch.nonzero(lab > 0).squeeze(1)                             # as the reference (mask_head/loss.py:86)
                bx = p.bbox.float()[pos]
                w, h = seg.size
                x1 = bx[:, 0].clamp(min=0, max=w - 1)                               # PolygonInstance.crop: clamped, at least 1 x 1
                y1 = bx[:, 1].clamp(min=0, max=h - 1)
                x2 = torch.maximum(bx[:, 2].clamp(min=0, max=w), x1 + 1)
                y2 = torch.maximum(bx[:, 3].clamp(min=0, max=h), y1 + 1)
                mk = ops.mask_targets_polygons(polyset_of(seg, dev), torch.stack([x1, y1, x2, y2], 1), midx[pos].to(torch.int32),
                                               int(self.discretization_size)) if pos.numel() else \
                    torch.empty(0, dtype=torch.float32, device=dev)
            labels.append(lab)
            masks.append(mk)
        return labels, masks
    ev.prepare_targets = types.MethodType(prepare_targets, ev)
    ev._mrb_fused = True
    return True


def _fuse_mask_targets(be, rep):
    """project_masks_on_boxes (roi_heads/mask_head/loss.py:11-42): the per-proposal crop / resize / rasterise loop on the
    HOST (flagged as a bottleneck at loss.py:31-32) becomes one launch on the polygons' vertices (csrc/mask_targets.cu).
    Module-level function: rebinding it affects every MaskRCNNLossComputation of the process (opt-in, idempotent).
    Cell-centre even-odd rule; pycocotools' rleFrPoly samples a 5x upsampled boundary instead and can differ on cells the
    polygon boundary passes through."""
    import sys
    m = sys.modules.get("maskrcnn_benchmark.modeling.roi_heads.mask_head.loss")
    if m is None or not hasattr(m, "project_masks_on_boxes"):
        return False
    orig = m.project_masks_on_boxes
    if getattr(orig, "_mrb_fused", False):
        return True

    def project_masks_on_boxes(segmentation_masks, proposals, discretization_size):
        from mrb_b200 import ops
        dev = proposals.bbox.device
        inst = getattr(segmentation_masks, "instances", None)
        polys = getattr(inst, "polygons", None)
        if dev.type != "cuda" or getattr(segmentation_masks, "mode", None) != "poly" or polys is None:
            return orig(segmentation_masks, proposals, discretization_size)
        r = len(polys)
        if r == 0:
            return torch.empty(0, dtype=torch.float32, device=dev)
        pset = ops.PolygonSet([[p.tolist() for p in pi.polygons] for pi in polys], dev)
        bx = proposals.convert("xyxy").bbox.float()
        w, h = segmentation_masks.size
        # PolygonInstance.crop (structures/segmentation_mask.py:270-292): the box is clamped to the image, at least 1 x 1
        x1 = bx[:, 0].clamp(min=0, max=w - 1)
        y1 = bx[:, 1].clamp(min=0, max=h - 1)
        x2 = torch.maximum(bx[:, 2].clamp(min=0, max=w), x1 + 1)
        y2 = torch.maximum(bx[:, 3].clamp(min=0, max=h), y1 + 1)
        rois = torch.stack([x1, y1, x2, y2], 1)
        return ops.mask_targets_polygons(pset, rois, torch.arange(r, device=dev, dtype=torch.int32), int(discretization_size))
    project_masks_on_boxes._mrb_fused = True
    project_masks_on_boxes._mrb_orig = orig
    m.project_masks_on_boxes = project_masks_on_boxes
    rep["fused"]["mask_targets"] = 1
    return True


def fuse_model(model, backend=None, channels_last_weights=True, sampling=True):
    """Rebind the forwards listed in the module docstring, in place.  Returns a report
    {"fused": {kind: count}, "skipped": [reasons]}.  Idempotent.
    sampling=False leaves the two stages that DRAW RANDOM SAMPLES (RPN loss targets, box-head subsample) to the reference's
    own Python: their fused forms sample with iid keys instead of torch.randperm -- the same distribution, not the same
    sample -- so comparisons that pin the reference's random stream (tests/refgraph) keep the reference's sampler."""
    if backend is None:
        from mrb_b200 import engine
        backend = engine.default_backend()
    be = backend
    rep = {"fused": {}, "skipped": []}

    def bump(k):
        rep["fused"][k] = rep["fused"].get(k, 0) + 1

    if channels_last_weights and getattr(be, "channels_last", False):
        # conv weights in KRSC memory (torch.channels_last): same logical tensors / state_dict; the tcgen05 weight
        # gradient, the bf16 operand copies and the optimizer state then share one layout
        for p in model.parameters():
            if p.dim() == 4 and not p.is_contiguous(memory_format=torch.channels_last):
                p.data = p.data.contiguous(memory_format=torch.channels_last)
    for name, mod in list(model.named_modules()):
        if getattr(mod, "_mrb_fused", False):
            continue
        cls = type(mod).__name__
        if "Stem" in cls and hasattr(mod, "conv1") and hasattr(mod, "bn1"):
            if _fuse_stem(mod, be, rep):
                bump("stem")
            continue
        kind = _bottleneck_kind(mod, be)
        if kind is not None:
            _fuse_bottleneck(mod, be, kind)
            bump("bottleneck[%s]" % kind)
            continue
        if hasattr(mod, "conv1") and hasattr(mod, "bn3") and hasattr(mod, "downsample"):
            rep["skipped"].append("%s: bottleneck variant not fused (%s)" % (name, type(mod.conv2).__name__))
            continue
        if cls == "FPN":
            if _fuse_fpn(mod, be, rep):
                bump("fpn")
            else:
                rep["skipped"].append("%s: FPN variant not fused" % name)
            continue
        if hasattr(mod, "cls_logits") and hasattr(mod, "bbox_pred") and hasattr(mod, "conv"):
            if _fuse_rpn_head(mod, be):
                bump("rpn_head")
            continue
        fe, pr = getattr(mod, "feature_extractor", None), getattr(mod, "predictor", None)
        if fe is not None and pr is not None:
            if hasattr(pr, "cls_score") and _fuse_box_head(fe, pr, be):
                bump("box_head")
            elif hasattr(pr, "mask_fcn_logits") and _fuse_mask_head(fe, pr, be):
                bump("mask_head")
            else:
                rep["skipped"].append("%s: ROI head variant not fused (%s / %s)" % (name, type(fe).__name__, type(pr).__name__))
            continue
    for name, mod in model.named_modules():
        if not getattr(mod, "_mrb_fused", False) and _pooler_ok(mod):
            _fuse_pooler(mod, be)
            bump("pooler")
    if getattr(be, "fused_glue", False):
        for name, mod in model.named_modules():
            if not getattr(mod, "_mrb_fused", False) and _fuse_rpn_postprocessor(mod, be):
                bump("rpn_postprocessor")
            elif not getattr(mod, "_mrb_fused", False) and _fuse_box_postprocessor(mod, be):
                bump("box_postprocessor")
            if type(getattr(mod, "loss_evaluator", None)).__name__ == "MaskRCNNLossComputation" and \
                    _fuse_mask_prepare_targets(mod.loss_evaluator, be):
                bump("mask_prepare_targets")
            if not sampling:
                continue
            if type(getattr(mod, "loss_evaluator", None)).__name__ == "RPNLossComputation" and _fuse_rpn_loss(mod, be):
                bump("rpn_loss_targets")
            elif type(getattr(mod, "loss_evaluator", None)).__name__ == "FastRCNNLossComputation" and \
                    _fuse_box_subsample(mod.loss_evaluator, be):
                bump("box_subsample")
        _fuse_mask_targets(be, rep)
    _wire_resnet(model, be)
    model._mrb_backend = be
    return rep
