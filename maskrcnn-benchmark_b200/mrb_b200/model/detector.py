"""GeneralizedRCNN for the R-50-FPN configs (reference modeling/detector/generalized_rcnn.py:16-65)."""
import torch
from torch import nn

from .backbone import Backbone
from .config import RCNNConfig
from .roi_heads import BoxHead, MaskHead, _to_rois
from .rpn import RPN


class _Heads(nn.Module):
    pass


class GeneralizedRCNN(nn.Module):
    def __init__(self, cfg, backend):
        super().__init__()
        self.cfg = cfg
        self.be = backend
        self.backbone = Backbone(cfg)
        self.rpn = RPN(cfg, self.backbone.out_channels)
        self.roi_heads = _Heads()
        self.roi_heads.box = BoxHead(cfg, self.backbone.out_channels)
        if cfg.mask_on:
            self.roi_heads.mask = MaskHead(cfg, self.backbone.out_channels)

    def pad_images(self, images):
        """to_image_list(size_divisible=32): zero-pad a list of [3,H,W] to a common divisible size
        (structures/image_list.py:29-72)."""
        d = self.cfg.size_divisibility
        sizes = [tuple(im.shape[-2:]) for im in images]
        h = -(-max(s[0] for s in sizes) // d) * d
        w = -(-max(s[1] for s in sizes) // d) * d
        batch = images[0].new_zeros((len(images), 3, h, w))
        for i, im in enumerate(images):
            batch[i, :, :im.shape[1], :im.shape[2]] = im
        return batch, sizes

    def _mask_loss(self, be, feats, rois, labels, gidx, targets, sm=None, gtp=None):
        mask = self.roi_heads.mask
        n, s = labels.shape
        res = self.cfg.mask_resolution
        fused = sm is not None and "mask_rois" in sm
        gt_all = None if fused else torch.stack([t["boxes"][gidx[i]] for i, t in enumerate(targets)])          # [n, s, 4]
        # polygon segmentations ("polygons": per instance, a list of flat [x0, y0, x1, y1, ...] sequences -- the reference's
        # SegmentationMask(mode="poly")): rasterised on the GPU, one launch (ops.mask_targets_polygons); without them the
        # instance mask of a ground-truth box is its rectangle
        polyset = inst_all = None
        if all("polygons" in t for t in targets):
            from mrb_b200 import ops
            polyset = targets[0].get("_polyset")
            if polyset is None:
                polyset = ops.PolygonSet([p for t in targets for p in t["polygons"]], labels.device)
            off, acc = [], 0
            for t in targets:
                off.append(acc)
                acc += len(t["polygons"])
            if fused:
                inst_off = targets[0].get("_inst_off")
                if inst_off is None:
                    inst_off = torch.tensor(off, device=gidx.device, dtype=gidx.dtype)[:, None]
            else:
                inst_all = gidx + torch.tensor(off, device=gidx.device, dtype=gidx.dtype)[:, None]   # [n, s] global instance index
        m = self.cfg.mask_rois_per_image
        if fused:
            # the positives-first list came out of the assign-and-sample launch
            rois_sel, lab_sel, wsel, mg = sm["mask_rois"], sm["mask_labels"], sm["mask_weight"], sm["mask_gt_index"]
            if polyset is not None:
                tgt = ops.mask_targets_polygons(polyset, rois_sel[:, 1:], (mg + inst_off).reshape(-1), res)
            else:
                gt_sel = torch.gather(gtp[0], 1, mg[..., None].expand(-1, -1, 4)).reshape(-1, 4)
                from mrb_b200 import ops as _ops
                tgt = _ops.mask_targets_rect(gt_sel, rois_sel, res)
            if getattr(be, "fused_losses", False):
                from mrb_b200 import ops as _ops
                return _ops.mask_head_loss(mask.run(be, feats, rois_sel, padded_logits=True), lab_sel, tgt, wsel)
            sel_logits = mask.run(be, feats, rois_sel, select=lab_sel)
            bce = torch.nn.functional.binary_cross_entropy_with_logits(sel_logits.float(), tgt,
                                                                       reduction="none").mean((1, 2))
            return torch.where(wsel > 0, bce, torch.zeros((), dtype=bce.dtype, device=bce.device)).sum() / wsel.sum().clamp(min=1)
        if m > 0:
            posm = labels > 0
            order = torch.sort((~posm).to(torch.int8), dim=1, stable=True)[1][:, :m]        # positives first
            wsel = torch.gather(posm, 1, order).reshape(-1).float()
            rois_sel = torch.gather(rois.view(n, s, 5), 1, order[..., None].expand(-1, -1, 5)).reshape(-1, 5)
            lab_sel = torch.gather(labels, 1, order).reshape(-1).clamp(min=0)
            gt_sel = torch.gather(gt_all, 1, order[..., None].expand(-1, -1, 4)).reshape(-1, 4)
            sel_logits = mask.run(be, feats, rois_sel, select=lab_sel)
            if polyset is not None:
                tgt = ops.mask_targets_polygons(polyset, rois_sel[:, 1:], torch.gather(inst_all, 1, order).reshape(-1), res)
            else:
                tgt = mask.mask_targets(gt_sel, rois_sel[:, 1:], res)
            bce = torch.nn.functional.binary_cross_entropy_with_logits(sel_logits.float(), tgt,
                                                                       reduction="none").mean((1, 2))
            return torch.where(wsel > 0, bce, torch.zeros((), dtype=bce.dtype, device=bce.device)).sum() / wsel.sum().clamp(min=1)
        lab = labels.reshape(-1)
        pos = (lab > 0).nonzero().squeeze(1)             # keep_only_positive_boxes (mask_head.py:11-32)
        rois_pos = rois[pos]
        logits = mask.run(be, feats, rois_pos)
        if polyset is not None:
            tgt = ops.mask_targets_polygons(polyset, rois_pos[:, 1:], inst_all.reshape(-1)[pos], res)
        else:
            tgt = mask.mask_targets(gt_all.reshape(n * s, 4)[pos], rois_pos[:, 1:], res)
        return mask.loss(logits, lab[pos], tgt)

    def forward(self, images, image_sizes, targets=None, generator=None):
        """images: [N,3,H,W] fp32 already padded; image_sizes: [(h, w)]; targets: list of dicts with
        'boxes' [G,4] xyxy fp32 and 'labels' [G] int64 (masks == box rectangles)."""
        be = self.be
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        gtp = sm = None
        if self.training and getattr(be, "fused_glue", False):
            from mrb_b200 import ops
            gtp = ops.pad_targets(targets, images.device)      # ground truth as fixed-shape tensors for the glue kernels
        feats = self.backbone.run(be, images)
        if self.training and hasattr(be, "heads_boundary"):
            feats = be.heads_boundary(feats)        # data-parallel runs: marks where the heads' gradients are complete
        proposals, losses = self.rpn.run(be, feats, image_sizes, targets, self.training, generator, gtp=gtp)
        box = self.roi_heads.box
        if self.training:
            if gtp is not None and proposals[0].shape[1] <= 8192:
                mm = self.cfg.mask_rois_per_image if self.cfg.mask_on else 0
                sm = box.subsample_fused(proposals, gtp, generator, mm)
                rois, labels, reg_t, gidx = sm["rois"], sm["labels"], sm["reg_targets"], sm["gt_index"]
            else:
                boxes, labels, reg_t, gidx = box.subsample(proposals, targets, generator, be)
                rois = _to_rois(boxes)
            if self.cfg.mask_on:
                # The mask branch needs only the sampled boxes/labels and the FPN features.  With a backend that offers
                # stream lanes it is issued on its own stream, BEFORE the box branch: autograd replays every node on the
                # stream of its forward, so forward and backward of the two heads overlap (the box head's FC layers
                # and losses leave most of the 148 SMs idle).
                lane = getattr(be, "lanes", None)
                st = lane[-1] if (lane and getattr(self.cfg, "parallel_heads", False)) else None
                if st is not None:
                    cur = torch.cuda.current_stream()
                    st.wait_stream(cur)
                    for t in list(feats) + [rois, labels, gidx] + (list(sm.values()) + list(gtp) if sm is not None else []):
                        t.record_stream(st)
                    with torch.cuda.stream(st):
                        loss_mask = self._mask_loss(be, feats, rois, labels, gidx, targets, sm, gtp)
                    loss_mask.record_stream(cur)
                else:
                    loss_mask = None
            x = box.features(be, feats, rois)
            if getattr(be, "fused_losses", False):
                from mrb_b200 import ops
                lc, lb = ops.box_head_loss(box.predict_packed(be, x), labels, reg_t, self.cfg.num_classes)
            else:
                cls, reg = box.predict(be, x)
                lc, lb = box.loss(cls, reg, labels, reg_t)
            losses.update({"loss_classifier": lc, "loss_box_reg": lb})
            if self.cfg.mask_on:
                if st is not None:
                    torch.cuda.current_stream().wait_stream(st)
                else:
                    loss_mask = self._mask_loss(be, feats, rois, labels, gidx, targets, sm, gtp)
                losses["loss_mask"] = loss_mask
            return losses
        boxes, _, valid = proposals
        rois = _to_rois(boxes)
        x = box.features(be, feats, rois)
        if getattr(be, "fused_glue", False) and self.cfg.detections_per_img > 0 and boxes.shape[1] * (self.cfg.num_classes - 1) < 2 ** 20:
            # fixed-shape, sync-free post-processing (padded detections): the whole eval forward is CUDA-graph capturable
            widths, heights = self.rpn._sizes(image_sizes, boxes.device)
            return box.postprocess_fused(box.predict_packed(be, x), proposals, widths, heights)
        cls, reg = box.predict(be, x)
        return box.postprocess(be, cls, reg, proposals, image_sizes)


def build_model(cfg=None, backend=None, device="cuda"):
    cfg = cfg or RCNNConfig()
    if backend is None:
        from .backend import B200Backend
        backend = B200Backend()
    model = GeneralizedRCNN(cfg, backend).to(device)
    if getattr(backend, "channels_last", False):
        # keep the fp32 master weights of every conv in KRSC memory (torch.channels_last): same logical shape and
        # state_dict, but the tcgen05 wgrad output, the optimizer state and the bf16 operand copies then all share
        # one layout (no per-step transposes, foreach optimizer fast path)
        for p in model.parameters():
            if p.dim() == 4:
                p.data = p.data.contiguous(memory_format=torch.channels_last)
    return model
