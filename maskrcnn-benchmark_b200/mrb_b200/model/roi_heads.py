"""Box and mask heads (reference modeling/roi_heads/box_head/{box_head,roi_box_feature_extractors,
roi_box_predictors,loss,inference}.py, mask_head/{mask_head,roi_mask_feature_extractors,
roi_mask_predictors,loss}.py, modeling/poolers.py).  Fixed-shape, mask-based formulation of the
sampling so that the box head runs without host synchronisation; the mask head gathers its
positive ROIs (one `nonzero`, as the reference does in keep_only_positive_boxes)."""
import torch
import torch.nn.functional as F
from torch import nn

from maskrcnn_benchmark.layers import Conv2d, ConvTranspose2d

from . import box_ops


def _to_rois(boxes):
    """[N, P, 4] -> [N*P, 5] with the image index in column 0 (poolers.py:72-89)."""
    n, p, _ = boxes.shape
    ids = torch.arange(n, device=boxes.device, dtype=boxes.dtype)[:, None, None].expand(n, p, 1)
    return torch.cat([ids, boxes], 2).reshape(n * p, 5)


class BoxHead(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg
        r = cfg.box_resolution
        fe = nn.Module()
        fe.fc6 = nn.Linear(in_channels * r * r, cfg.mlp_head_dim)
        fe.fc7 = nn.Linear(cfg.mlp_head_dim, cfg.mlp_head_dim)
        for l in (fe.fc6, fe.fc7):
            nn.init.kaiming_uniform_(l.weight, a=1)        # make_layers.py:80-92
            nn.init.constant_(l.bias, 0)
        self.feature_extractor = fe
        pr = nn.Module()
        pr.cls_score = nn.Linear(cfg.mlp_head_dim, cfg.num_classes)
        pr.bbox_pred = nn.Linear(cfg.mlp_head_dim, cfg.num_classes * 4)
        nn.init.normal_(pr.cls_score.weight, std=0.01)     # roi_box_predictors.py:49-53
        nn.init.normal_(pr.bbox_pred.weight, std=0.001)
        for l in (pr.cls_score, pr.bbox_pred):
            nn.init.constant_(l.bias, 0)
        self.predictor = pr
        self.box_coder = box_ops.BoxCoder(cfg.bbox_reg_weights)
        self.matcher = box_ops.Matcher(cfg.roi_fg_iou, cfg.roi_bg_iou, allow_low_quality_matches=False)

    @torch.no_grad()
    def subsample(self, proposals, targets, generator=None, be=None, keys=None):
        """box_head/loss.py:41-118: match, label, sample 512 per image (25 % positive).
        -> boxes [N, S, 4], labels [N, S] (-1 = padding), reg targets [N, S, 4], matched gt [N, S]"""
        cfg = self.cfg
        boxes, _, valid = proposals
        n = boxes.shape[0]
        S = cfg.roi_batch_size
        def image(i):
            t = targets[i]
            q = box_ops.box_iou(t["boxes"], boxes[i])
            midx = self.matcher(q)
            lab = t["labels"][midx.clamp(min=0)].to(torch.int64)
            lab = torch.where(midx == box_ops.Matcher.BELOW_LOW, torch.zeros_like(lab), lab)
            lab = torch.where(midx == box_ops.Matcher.BETWEEN, -torch.ones_like(lab), lab)
            lab = torch.where(valid[i], lab, -torch.ones_like(lab))
            pos, neg = box_ops.sample_pos_neg(lab, S, cfg.roi_positive_fraction, generator,
                                              key=None if keys is None else keys[i])
            sel = pos | neg
            # fixed-size gather: selected rows first, in index order (== nonzero(pos | neg))
            order = torch.sort((~sel).to(torch.int8), stable=True)[1][:S]
            ok = sel[order]
            b = boxes[i][order]
            g = midx.clamp(min=0)[order]
            return b, torch.where(ok, lab[order], -torch.ones_like(lab[order])), self.box_coder.encode(t["boxes"][g], b), g

        # images are independent: one stream lane each when the backend offers them
        fork = getattr(be, "fork", None) if be is not None else None
        with torch.no_grad():
            if fork is not None:
                hs = [fork((boxes, valid, targets[i]["boxes"], targets[i]["labels"]), lambda i=i: image(i), lane=i) for i in range(n)]
                outs = [be.join(h) for h in hs]
            else:
                outs = [image(i) for i in range(n)]
        ob, ol, ot, og = zip(*outs)
        return torch.stack(ob), torch.stack(ol), torch.stack(ot), torch.stack(og)

    @torch.no_grad()
    def subsample_fused(self, proposals, gtp, generator=None, mask_rois_per_image=0, keys=None):
        """subsample (+ the mask branch's positives-first list) as one launch: mrb_roi_assign_sample (csrc/detect_glue.cu).
        -> dict(rois [N*S, 5], labels [N, S], reg_targets [N, S, 4], gt_index [N, S], mask_*)."""
        from mrb_b200 import ops
        cfg = self.cfg
        boxes, _, valid = proposals
        if keys is None:
            keys = torch.rand(boxes.shape[:2], device=boxes.device, generator=generator)
        return ops.roi_assign_sample(boxes, valid, keys, gtp[0], gtp[1], gtp[2], cfg.roi_batch_size, cfg.roi_positive_fraction,
                                     cfg.roi_fg_iou, cfg.roi_bg_iou, cfg.bbox_reg_weights,
                                     min(int(mask_rois_per_image), cfg.roi_batch_size))

    def features(self, be, feats, rois):
        cfg = self.cfg
        x = be.roi_align_fpn(feats[:4], rois, cfg.pooler_scales, cfg.box_resolution, cfg.box_sampling_ratio, nhwc=False)
        x = x.flatten(1)
        fe = self.feature_extractor
        x = be.linear(x, fe.fc6.weight, fe.fc6.bias, relu=True, gy_premasked=True)
        return be.linear(x, fe.fc7.weight, fe.fc7.bias, relu=True, premask_x=True, gy_premasked=True)

    def predict(self, be, x):
        pr = self.predictor
        w = torch.cat([pr.cls_score.weight, pr.bbox_pred.weight], 0)
        b = torch.cat([pr.cls_score.bias, pr.bbox_pred.bias], 0)
        o = be.linear(x, w, b, relu=False, out_fp32=True, premask_x=True)
        nc = self.cfg.num_classes
        return o[:, :nc], o[:, nc:]

    def predict_packed(self, be, x):
        """cls_score and bbox_pred as one GEMM, the padded [R, ld >= 5 C] fp32 result as it comes out of the engine"""
        pr = self.predictor
        w = torch.cat([pr.cls_score.weight, pr.bbox_pred.weight], 0)
        b = torch.cat([pr.cls_score.bias, pr.bbox_pred.bias], 0)
        return be.linear(x, w, b, relu=False, out_fp32=True, premask_x=True, keep_padded=True)

    def loss(self, class_logits, box_regression, labels, reg_targets):
        """box_head/loss.py:120-167"""
        labels = labels.reshape(-1)
        reg_targets = reg_targets.reshape(-1, 4)
        cls_loss = F.cross_entropy(class_logits.float(), labels, ignore_index=-1)
        pos = labels > 0
        idx = (4 * labels.clamp(min=0))[:, None] + torch.arange(4, device=labels.device)[None, :]
        pred = torch.gather(box_regression.float(), 1, idx)
        # The reference INDEXES the positive rows (loss.py:150-160); this fixed-shape form must select, not multiply:
        # a degenerate proposal (exp(dw) underflows with random-init RPN deltas -> width 0) encodes to an infinite
        # target, harmless in a row that is never positive but inf * 0 = NaN under a multiplicative mask.
        zero = torch.zeros((), dtype=pred.dtype, device=pred.device)
        diff = torch.abs(torch.where(pos[:, None], pred - reg_targets, zero))
        l1 = torch.where(diff < 1.0, 0.5 * diff * diff, diff - 0.5)
        box_loss = torch.where(pos[:, None], l1, zero).sum() / (labels >= 0).sum().clamp(min=1)
        return cls_loss, box_loss

    @torch.no_grad()
    def postprocess_fused(self, outputs, proposals, widths, heights):
        """postprocess on the packed predictor output, three launches + one batched NMS, fixed shapes and no host
        synchronisation (ops.box_postprocess): a list of {"boxes" [D, 4], "scores" [D], "labels" [D], "count" []} per image,
        rows beyond `count` zero."""
        from mrb_b200 import ops
        cfg = self.cfg
        boxes, _, valid = proposals
        b, s, l, c = ops.box_postprocess(outputs, cfg.num_classes, boxes, valid, widths, heights, cfg.score_thresh,
                                         cfg.bbox_reg_weights, cfg.roi_nms, cfg.detections_per_img)
        return [{"boxes": b[i], "scores": s[i], "labels": l[i], "count": c[i]} for i in range(b.shape[0])]

    @torch.no_grad()
    def postprocess(self, be, class_logits, box_regression, proposals, image_sizes):
        """box_head/inference.py:45-149: softmax, decode, clip, per-class NMS, top-100."""
        cfg = self.cfg
        boxes, _, valid = proposals
        n, p, _ = boxes.shape
        prob = F.softmax(class_logits.float(), -1).view(n, p, -1)
        dec = self.box_coder.decode(box_regression.float(), boxes.reshape(-1, 4)).view(n, p, -1)
        results = []
        for i in range(n):
            h, w = image_sizes[i]
            bx = box_ops.clip_boxes(dec[i].reshape(p, -1, 4), w, h)
            sc = prob[i]
            ok = valid[i][:, None] & (sc > cfg.score_thresh)
            ok[:, 0] = False
            det_b, det_s, det_l, sizes = [], [], [], []
            for j in range(1, cfg.num_classes):
                inds = ok[:, j].nonzero().squeeze(1)
                det_b.append(bx[inds, j])
                det_s.append(sc[inds, j])
                det_l.append(torch.full_like(inds, j))
                sizes.append(int(inds.numel()))
            cb, cs, cl = torch.cat(det_b), torch.cat(det_s), torch.cat(det_l)
            if cb.numel():
                keep, counts = be.nms_batched(cb.contiguous(), cs.contiguous(), sizes, cfg.roi_nms)
                counts = counts.tolist()
                off, sel = 0, []
                for sz, c in zip(sizes, counts):
                    sel.append(keep[off:off + c] + off)
                    off += sz
                sel = torch.cat(sel)
                cb, cs, cl = cb[sel], cs[sel], cl[sel]
            if cs.numel() > cfg.detections_per_img > 0:
                thr = torch.kthvalue(cs.cpu(), cs.numel() - cfg.detections_per_img + 1)[0].item()
                k = cs >= thr
                cb, cs, cl = cb[k], cs[k], cl[k]
            results.append({"boxes": cb, "scores": cs, "labels": cl})
        return results


class MaskHead(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg
        fe = nn.Module()
        self.blocks = []
        cin = in_channels
        for i, c in enumerate(cfg.mask_conv_layers, 1):
            conv = Conv2d(cin, c, 3, 1, 1)
            nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")   # make_layers.py:44-77
            nn.init.constant_(conv.bias, 0)
            setattr(fe, "mask_fcn%d" % i, conv)
            self.blocks.append("mask_fcn%d" % i)
            cin = c
        self.feature_extractor = fe
        pr = nn.Module()
        pr.conv5_mask = ConvTranspose2d(cin, cin, 2, 2, 0)
        pr.mask_fcn_logits = Conv2d(cin, cfg.num_classes, 1, 1, 0)
        for name, p in pr.named_parameters():                                            # roi_mask_predictors.py:21-27
            if "bias" in name:
                nn.init.constant_(p, 0)
            elif "weight" in name:
                nn.init.kaiming_normal_(p, mode="fan_out", nonlinearity="relu")
        self.predictor = pr

    def run(self, be, feats, rois, select=None, padded_logits=False):
        """Mask logits [R, num_classes, M, M]; with `select` (class index per ROI, training) only the selected class
        plane of each ROI, [R, M, M] (mask_head/loss.py:120-126).  padded_logits: the logit conv's bf16 NHWC output with
        its channels rounded up to 8, for ops.mask_head_loss."""
        cfg = self.cfg
        x = be.roi_align_fpn(feats[:4], rois, cfg.pooler_scales, cfg.mask_resolution_pool, cfg.mask_sampling_ratio,
                             nhwc=True)
        fe, pr = self.feature_extractor, self.predictor
        for i, name in enumerate(self.blocks):
            c = getattr(fe, name)
            x = be.conv(x, c.weight, bias=c.bias, pad=1, relu=True, premask_x=i > 0, gy_premasked=True)
        x = be.deconv2x2(x, pr.conv5_mask.weight, pr.conv5_mask.bias, relu=True, premask_x=True, gy_premasked=True)
        if padded_logits:
            return be.conv(x, pr.mask_fcn_logits.weight, bias=pr.mask_fcn_logits.bias, out_fp32=False, premask_x=True,
                           keep_padded=True)
        if select is not None and hasattr(be, "conv_select"):
            return be.conv_select(x, pr.mask_fcn_logits.weight, pr.mask_fcn_logits.bias, select, premask_x=True)
        logits = be.conv(x, pr.mask_fcn_logits.weight, bias=pr.mask_fcn_logits.bias, out_fp32=True, premask_x=True)
        if select is not None:
            logits = logits[torch.arange(logits.shape[0], device=logits.device), select]
        return logits

    @staticmethod
    @torch.no_grad()
    def mask_targets(gt_boxes, proposals, m):
        """project_masks_on_boxes (mask_head/loss.py:11-42) for the benchmark's rectangle masks: the
        instance mask of a ground-truth box is its rectangle, so the target of a proposal is that
        rectangle cropped to the proposal and sampled at the m x m cell centres."""
        x1, y1, x2, y2 = proposals.unbind(1)
        sx = m / (x2 - x1).clamp(min=1e-6)
        sy = m / (y2 - y1).clamp(min=1e-6)
        c = torch.arange(m, device=proposals.device, dtype=torch.float32) + 0.5
        gx1, gx2 = (gt_boxes[:, 0] - x1) * sx, (gt_boxes[:, 2] - x1) * sx
        gy1, gy2 = (gt_boxes[:, 1] - y1) * sy, (gt_boxes[:, 3] - y1) * sy
        inx = (c[None, :] >= gx1[:, None]) & (c[None, :] <= gx2[:, None])
        iny = (c[None, :] >= gy1[:, None]) & (c[None, :] <= gy2[:, None])
        return (iny[:, :, None] & inx[:, None, :]).float()

    def loss(self, mask_logits, labels_pos, targets):
        """mask_head/loss.py:100-133"""
        if mask_logits.shape[0] == 0:
            return mask_logits.sum() * 0
        idx = torch.arange(mask_logits.shape[0], device=mask_logits.device)
        return F.binary_cross_entropy_with_logits(mask_logits[idx, labels_pos].float(), targets)
