"""RPN: head, anchors, proposal selection, loss (reference modeling/rpn/rpn.py:73-205,
rpn/inference.py:14-201, rpn/loss.py:23-157, rpn/anchor_generator.py:34-148).

Differences from the reference, all on the host-logic side and none in the arithmetic of a proposal:
the per-(image, level) Python loops are batched (one top-k / decode / clip per level for the whole
batch, ONE batched NMS launch sequence for all image x level problems) and kept fixed-shape with
validity masks, so the step has no host synchronisation here (the reference syncs in every NMS call,
csrc/cuda/nms.cu:100, and in every `nonzero`)."""
import torch
import torch.nn.functional as F
from torch import nn

from maskrcnn_benchmark.layers import Conv2d, smooth_l1_loss

from . import box_ops


class BufferList(nn.Module):
    """rpn/anchor_generator.py:13-31: keeps cell anchors in the state_dict under cell_anchors.<i>"""

    def __init__(self, buffers):
        super().__init__()
        for i, b in enumerate(buffers):
            self.register_buffer(str(i), b)

    def __iter__(self):
        return iter(self._buffers.values())

    def __len__(self):
        return len(self._buffers)


class AnchorGenerator(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.strides = cfg.anchor_strides
        self.straddle_thresh = cfg.straddle_thresh
        self.cell_anchors = BufferList([box_ops.cell_anchors(s, (sz,), cfg.aspect_ratios)
                                        for s, sz in zip(cfg.anchor_strides, cfg.anchor_sizes)])
        self._cache = {}

    def num_anchors_per_location(self):
        return [len(c) for c in self.cell_anchors]

    def grid(self, grid_sizes, device):
        key = (tuple(grid_sizes), str(device))
        if key not in self._cache:
            self._cache[key] = [box_ops.grid_anchors(c, s, gh, gw, device)
                                for (gh, gw), s, c in zip(grid_sizes, self.strides, self.cell_anchors)]
        return self._cache[key]

    def visibility(self, anchors, width, height):
        t = self.straddle_thresh
        if t < 0:
            return torch.ones(anchors.shape[0], dtype=torch.bool, device=anchors.device)
        return (anchors[:, 0] >= -t) & (anchors[:, 1] >= -t) & (anchors[:, 2] < width + t) & (anchors[:, 3] < height + t)


class RPNHead(nn.Module):
    """rpn.py:73-106"""

    def __init__(self, in_channels, num_anchors):
        super().__init__()
        self.conv = Conv2d(in_channels, in_channels, 3, 1, 1)
        self.cls_logits = Conv2d(in_channels, num_anchors, 1)
        self.bbox_pred = Conv2d(in_channels, num_anchors * 4, 1)
        for l in (self.conv, self.cls_logits, self.bbox_pred):
            nn.init.normal_(l.weight, std=0.01)
            nn.init.constant_(l.bias, 0)
        self.num_anchors = num_anchors

    def run(self, be, feats):
        # cls and bbox 1x1 heads run as ONE conv with Cout = A + 4A over the shared 3x3 output
        w = torch.cat([self.cls_logits.weight, self.bbox_pred.weight], 0)
        b = torch.cat([self.cls_logits.bias, self.bbox_pred.bias], 0)
        logits, deltas = [], []
        for f in feats:
            t = be.conv(f, self.conv.weight, bias=self.conv.bias, pad=1, relu=True, gy_premasked=True)
            o = be.conv(t, w, bias=b, out_fp32=True, premask_x=True)   # [N, 5A, H, W]
            o = o.permute(0, 2, 3, 1)                             # [N, H, W, 5A] (a view for channels_last)
            n, h, ww, _ = o.shape
            a = self.num_anchors
            logits.append(o[..., :a].reshape(n, h * ww * a))      # (h, w, a) order == permute_and_flatten
            deltas.append(o[..., a:].reshape(n, h * ww * a, 4))
        return logits, deltas

    def run_packed(self, be, feats):
        """The head outputs as they leave the engine: per level one NHWC fp32 tensor [N, H, W, ld] holding, per location,
        A logits, 4A deltas and zero padding up to ld = roundup8(5A) -- read in place by the decode / loss launches."""
        w = torch.cat([self.cls_logits.weight, self.bbox_pred.weight], 0)
        b = torch.cat([self.cls_logits.bias, self.bbox_pred.bias], 0)
        outs = []
        for f in feats:
            t = be.conv(f, self.conv.weight, bias=self.conv.bias, pad=1, relu=True, gy_premasked=True)
            o = be.conv(t, w, bias=b, out_fp32=True, premask_x=True, keep_padded=True)   # [N, ld, H, W] channels_last
            outs.append(o.permute(0, 2, 3, 1))
        return outs


class RPN(nn.Module):
    def __init__(self, cfg, in_channels):
        super().__init__()
        self.cfg = cfg
        self.anchor_generator = AnchorGenerator(cfg)
        self.head = RPNHead(in_channels, self.anchor_generator.num_anchors_per_location()[0])
        self.box_coder = box_ops.BoxCoder((1.0, 1.0, 1.0, 1.0))
        self.matcher = box_ops.Matcher(cfg.rpn_fg_iou, cfg.rpn_bg_iou, allow_low_quality_matches=True)
        self._size_cache = {}

    # ------------------------------------------------------------------ proposals (no_grad)
    def _sizes(self, image_sizes, dev):
        """(widths [N], heights [N]) fp32 on the device, cached: host -> device copies must not sit inside a captured step"""
        key = (tuple(tuple(s) for s in image_sizes), str(dev))
        if key not in self._size_cache:
            self._size_cache[key] = (torch.tensor([s[1] for s in image_sizes], device=dev, dtype=torch.float32),
                                     torch.tensor([s[0] for s in image_sizes], device=dev, dtype=torch.float32))
        return self._size_cache[key]

    @torch.no_grad()
    def _select_proposals_fused(self, be, anchors, logits, deltas, image_sizes, gtp, training, outs=None):
        """select_proposals as 5 top-k + 5 decode launches (one lane per level), the batched NMS and ONE collect launch
        (csrc/detect_glue.cu); same contract.  The top-k runs on the logits (sigmoid is monotone; the k sigmoids are taken
        in the decode kernel)."""
        from mrb_b200 import ops
        cfg = self.cfg
        pre_n = cfg.pre_nms_top_n_train if training else cfg.pre_nms_top_n_test
        post_n = cfg.post_nms_top_n_train if training else cfg.post_nms_top_n_test
        fpn_post_n = cfg.fpn_post_nms_top_n_train if training else cfg.fpn_post_nms_top_n_test
        apl = self.head.num_anchors
        if outs is not None:         # head outputs [N, H, W, ld] read in place (RPNHead.run_packed)
            outs = [o.detach() for o in outs]
            logits, deltas = outs, outs
            ks = [min(pre_n, o.shape[1] * o.shape[2] * apl) for o in outs]
        else:
            ks = [min(pre_n, lg.shape[1]) for lg in logits]
        n = logits[0].shape[0]
        dev = logits[0].device
        widths, heights = self._sizes(image_sizes, dev)
        tot = n * sum(ks)
        boxes = torch.empty((tot, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((tot,), dtype=torch.float32, device=dev)
        fork = getattr(be, "fork", None)
        handles, off = [], 0
        for li, (anc, lg, dl, k) in enumerate(zip(anchors, logits, deltas, ks)):
            bo, so = boxes[off:off + n * k], scores[off:off + n * k]

            def level(anc=anc, lg=lg, dl=dl, k=k, bo=bo, so=so):
                if outs is not None:
                    if getattr(be, "fused_topk", False) and k <= 8192 and lg.shape[1] * lg.shape[2] * apl <= 8 * 50000:
                        ops.rpn_topk_decode(lg, apl, anc, k, widths, heights, bo, so, self.box_coder.weights, self.box_coder.clip)
                        return bo
                    idx = lg[..., :apl].reshape(n, -1).topk(k, dim=1, sorted=True)[1]
                    ops.rpn_decode_packed(lg, apl, anc, idx, widths, heights, bo, so, self.box_coder.weights, self.box_coder.clip)
                    return bo
                idx = lg.topk(k, dim=1, sorted=True)[1]                                   # inference.py:91-95
                ops.rpn_decode(lg, dl, anc, idx, widths, heights, bo, so, self.box_coder.weights, self.box_coder.clip)
                return bo
            if fork is not None:
                handles.append(fork((anc, lg, dl, widths, heights, boxes, scores), level, lane=li))
            else:
                level()
            off += n * k
        for h in handles:
            be.join(h)
        sizes = [k for k in ks for _ in range(n)]
        # every problem's rows come out of a sorted top-k (descending logit == descending sigmoid, ties in anchor order)
        try:
            keep, counts = be.nms_batched(boxes, scores, sizes, cfg.rpn_nms_thresh, presorted=True)
        except TypeError:            # a backend without the presorted form
            keep, counts = be.nms_batched(boxes, scores, sizes, cfg.rpn_nms_thresh)
        per_batch = bool(training and cfg.fpn_post_nms_per_batch)
        add_gt = training and gtp is not None
        return ops.rpn_collect(boxes, scores, keep, counts, ks, n, post_n, fpn_post_n, per_batch,
                               gtp[0] if add_gt else None, gtp[2] if add_gt else None)

    @torch.no_grad()
    def select_proposals(self, be, anchors, logits, deltas, image_sizes, targets, training, gtp=None):
        """-> (boxes [N, P, 4], scores [N, P], valid [N, P] bool).  P is fixed; invalid rows hold
        zero boxes / -1 scores."""
        cfg = self.cfg
        n = logits[0].shape[0]
        dev = logits[0].device
        if getattr(be, "fused_glue", False) and (not training or targets is None or gtp is not None) and \
                not (training and cfg.fpn_post_nms_per_batch and n > 8):
            return self._select_proposals_fused(be, anchors, logits, deltas, image_sizes, gtp, training)
        pre_n = cfg.pre_nms_top_n_train if training else cfg.pre_nms_top_n_test
        post_n = cfg.post_nms_top_n_train if training else cfg.post_nms_top_n_test
        fpn_post_n = cfg.fpn_post_nms_top_n_train if training else cfg.fpn_post_nms_top_n_test
        widths, heights = self._sizes(image_sizes, dev)
        def level(anc, lg, dl):
            k = min(pre_n, lg.shape[1])
            sc, idx = lg.sigmoid().topk(k, dim=1, sorted=True)                       # inference.py:91-95
            d = torch.gather(dl, 1, idx[..., None].expand(-1, -1, 4))
            bx = self.box_coder.decode(d.reshape(-1, 4).float(), anc[idx.reshape(-1)]).view(n, k, 4)
            # clip_to_image(remove_empty=False); remove_small_boxes(min_size=0) keeps everything
            lim = torch.stack([widths, heights, widths, heights], 1)[:, None, :] - 1      # [N, 1, 4]
            return torch.minimum(bx.clamp(min=0), lim), sc

        # the per-level pipelines are independent of each other: with a backend that offers extra streams they run
        # side by side (their top-k kernels are single- or few-block launches that leave the chip empty)
        fork = getattr(be, "fork", None)
        handles = []
        for li, (anc, lg, dl) in enumerate(zip(anchors, logits, deltas)):
            if fork is not None:
                handles.append(fork((anc, lg, dl, widths, heights), lambda anc=anc, lg=lg, dl=dl: level(anc, lg, dl), lane=li))
            else:
                handles.append(level(anc, lg, dl))
        all_boxes, all_scores, ks = [], [], []
        for h in handles:
            bx, sc = be.join(h) if fork is not None else h
            all_boxes.append(bx)
            all_scores.append(sc)
            ks.append(bx.shape[1])
        # one batched NMS over the N x L problems, problem order = (level, image)
        boxes = torch.cat([b.reshape(-1, 4) for b in all_boxes]).contiguous()
        scores = torch.cat([s.reshape(-1) for s in all_scores]).contiguous()
        sizes = [k for k in ks for _ in range(n)]
        keep, counts = be.nms_batched(boxes, scores, sizes, cfg.rpn_nms_thresh)       # keep: relative, ascending
        # per problem: first min(count, post_n) kept rows (scores are sorted, so ascending index == by score)
        out_b, out_s, out_v = [], [], []
        off = 0
        for li, k in enumerate(ks):
            kk = keep[off:off + n * k].view(n, k)
            cnt = counts[li * n:(li + 1) * n].clamp(max=post_n)
            m = min(k, post_n)
            pos = torch.arange(m, device=dev)[None, :]
            valid = pos < cnt[:, None]
            idx = kk[:, :m].clamp(min=0, max=k - 1)
            out_b.append(torch.gather(all_boxes[li], 1, idx[..., None].expand(-1, -1, 4)))
            out_s.append(torch.where(valid, torch.gather(all_scores[li], 1, idx), all_scores[li].new_full((), -1.0)))
            out_v.append(valid)
            off += n * k
        b = torch.cat(out_b, 1)
        s = torch.cat(out_s, 1)
        v = torch.cat(out_v, 1)
        # select_over_all_levels (inference.py:154-181)
        if training and cfg.fpn_post_nms_per_batch:
            flat = s.reshape(-1)
            topn = min(fpn_post_n, flat.numel())
            _, sel = flat.topk(topn, sorted=True)
            m = torch.zeros_like(flat, dtype=torch.bool)
            m.index_fill_(0, sel, True)      # (m[sel] = True would stage a CPU scalar: not graph-capturable)
            v = v & m.view_as(v)
            # compact every image to a fixed width of `topn` columns, valid rows first (stable)
            order = torch.sort((~v).to(torch.int8), dim=1, stable=True)[1][:, :topn]
        else:
            topn = min(fpn_post_n, s.shape[1])
            order = s.topk(topn, dim=1, sorted=True)[1]
        b = torch.gather(b, 1, order[..., None].expand(-1, -1, 4))
        s = torch.gather(s, 1, order)
        v = torch.gather(v, 1, order)
        if training and targets is not None:
            # add_gt_proposals (inference.py:53-74)
            gmax = max(t["boxes"].shape[0] for t in targets)
            gb = b.new_zeros((n, gmax, 4))
            gv = torch.zeros((n, gmax), dtype=torch.bool, device=dev)
            for i, t in enumerate(targets):
                g = t["boxes"].shape[0]
                gb[i, :g] = t["boxes"]
                gv[i, :g] = True
            b = torch.cat([b, gb], 1)
            s = torch.cat([s, gv.float()], 1)
            v = torch.cat([v, gv], 1)
        b = torch.where(v[..., None], b, torch.zeros((), device=dev))
        return b, s, v

    # ------------------------------------------------------------------ loss
    @torch.no_grad()
    def loss_targets_fused(self, anchors_all, image_sizes, targets, gtp, generator=None, stacked=False, keys=None):
        """loss_targets with the IoU / Matcher / labelling of all anchors of the batch as one launch pair
        (mrb_rpn_anchor_match); the sampling and the encoding of the sampled rows as in loss_targets."""
        from mrb_b200 import ops
        cfg = self.cfg
        widths, heights = self._sizes(image_sizes, anchors_all.device)
        labels, matched = ops.rpn_anchor_match(anchors_all, gtp[0], gtp[2], widths, heights, cfg.rpn_fg_iou, cfg.rpn_bg_iou,
                                               float(self.anchor_generator.straddle_thresh))
        if keys is None and labels.shape[1] <= 8 * 40000:
            # sampling + encode of the batch as ONE cluster launch (the PyTorch formulation below thins, compacts and top-k's
            # per image: ~30 launches each; it stays as the checker of the kernel, tests/test_glue_gpu.py)
            keys = torch.rand(labels.shape, device=labels.device, generator=generator)
            st = ops.rpn_sample(labels, matched, keys, anchors_all, gtp[0], cfg.rpn_batch_size, cfg.rpn_positive_fraction,
                                self.box_coder.weights)
            return st if stacked else [tuple(t[i] for t in st) for i in range(labels.shape[0])]
        out = []
        for i, t in enumerate(targets):
            pos_idx, pos_ok, neg_idx, neg_ok = box_ops.sample_pos_neg_idx(labels[i], cfg.rpn_batch_size,
                                                                            cfg.rpn_positive_fraction, generator,
                                                                            key=None if keys is None else keys[i])
            gt = t["boxes"][matched[i][pos_idx].long()]
            reg_t = self.box_coder.encode(gt, anchors_all[pos_idx])
            sel = torch.cat([pos_idx, neg_idx])
            sel_lab = torch.cat([torch.ones_like(pos_ok, dtype=torch.float32), torch.zeros_like(neg_ok, dtype=torch.float32)])
            sel_w = torch.cat([pos_ok, neg_ok]).float()
            out.append((pos_idx, pos_ok, reg_t, sel, sel_lab, sel_w))
        if stacked:
            return tuple(torch.stack([o[j] for o in out]) for j in range(6))
        return out

    @torch.no_grad()
    def loss_targets(self, anchors_all, visibility, targets, generator=None):
        """Anchor labelling + sampling + regression targets (loss.py:40-131, no gradients, independent of the
        network outputs): matching and labelling run over all ~268k anchors, everything after the sampling touches
        just the fixed-size sampled rows (the reference computes targets densely and then indexes with the sampled
        positions: same values).  Per image: (pos_idx, pos_ok, reg_t, sel, sel_lab, sel_w)."""
        cfg = self.cfg
        out = []
        for i, t in enumerate(targets):
            q = box_ops.box_iou(t["boxes"], anchors_all)                           # loss.py:40-52
            midx = self.matcher(q)
            lab = (midx >= 0).float()
            lab = torch.where(midx == box_ops.Matcher.BELOW_LOW, torch.zeros_like(lab), lab)
            lab = torch.where(~visibility[i], -torch.ones_like(lab), lab)          # not_visibility
            lab = torch.where(midx == box_ops.Matcher.BETWEEN, -torch.ones_like(lab), lab)
            pos_idx, pos_ok, neg_idx, neg_ok = box_ops.sample_pos_neg_idx(lab, cfg.rpn_batch_size,
                                                                            cfg.rpn_positive_fraction, generator)
            gt = t["boxes"][midx[pos_idx].clamp(min=0)]
            reg_t = self.box_coder.encode(gt, anchors_all[pos_idx])
            sel = torch.cat([pos_idx, neg_idx])
            sel_lab = torch.cat([torch.ones_like(pos_ok, dtype=torch.float32), torch.zeros_like(neg_ok, dtype=torch.float32)])
            sel_w = torch.cat([pos_ok, neg_ok]).float()
            out.append((pos_idx, pos_ok, reg_t, sel, sel_lab, sel_w))
        return out

    def loss(self, anchors_all, visibility, logits, deltas, targets, generator=None, prepared=None):
        obj = torch.cat(logits, 1)            # [N, A_total]
        reg = torch.cat(deltas, 1)            # [N, A_total, 4]
        if prepared is None:
            prepared = self.loss_targets(anchors_all, visibility, targets, generator)
        box_sum = obj_sum = num_sampled = 0
        beta = 1.0 / 9
        for i, (pos_idx, pos_ok, reg_t, sel, sel_lab, sel_w) in enumerate(prepared):
            num_sampled = num_sampled + sel_w.sum()
            zero = torch.zeros((), dtype=torch.float32, device=reg_t.device)
            diff = torch.abs(torch.where(pos_ok[:, None], reg[i][pos_idx].float() - reg_t, zero))     # select, never inf * 0
            l1 = torch.where(diff < beta, 0.5 * diff * diff / beta, diff - 0.5 * beta)
            box_sum = box_sum + torch.where(pos_ok[:, None], l1, zero).sum()
            bce = F.binary_cross_entropy_with_logits(obj[i][sel].float(), sel_lab, reduction="none")
            obj_sum = obj_sum + torch.where(sel_w > 0, bce, zero).sum()
        num_sampled = num_sampled.clamp(min=1)
        return obj_sum / num_sampled, box_sum / num_sampled

    def run(self, be, feats, image_sizes, targets, training, generator=None, gtp=None):
        grid_sizes = [f.shape[-2:] for f in feats]
        anchors = self.anchor_generator.grid(grid_sizes, feats[0].device)
        prepared = None
        fused = getattr(be, "fused_glue", False) and gtp is not None
        if training:
            akey = (tuple(tuple(g) for g in grid_sizes), str(feats[0].device))
            if getattr(self, "_anchors_all", (None,))[0] != akey:
                self._anchors_all = (akey, torch.cat(anchors, 0))
            anchors_all = self._anchors_all[1]
            vis = None if fused else torch.stack([self.anchor_generator.visibility(anchors_all, w, h) for (h, w) in image_sizes])
            # the target assignment needs nothing from the network: with a backend that offers a second stream it runs
            # there, concurrently with the RPN head convolutions and the proposal selection (a few hundred tiny kernels
            # that would otherwise sit on the critical path between two tensor-core phases)
            fork = getattr(be, "fork", None)
            packed = fused and getattr(be, "fused_losses", False)
            if fused:
                tfn = lambda: self.loss_targets_fused(anchors_all, image_sizes, targets, gtp, generator, stacked=packed)
            else:
                tfn = lambda: self.loss_targets(anchors_all, vis, targets, generator)
            forked = fork is not None
            if forked:
                ins = (anchors_all,) + tuple(t["boxes"] for t in targets) + (tuple(gtp) if fused else (vis,))
                prepared = fork(ins, tfn)
            elif fused:
                prepared = tfn()
        if (not training) and getattr(be, "fused_glue", False) and getattr(be, "fused_losses", False):
            outs = self.head.run_packed(be, feats)
            return self._select_proposals_fused(be, anchors, None, None, image_sizes, None, False, outs=outs), {}
        if training and packed:
            # fused losses: the head outputs stay in the engine's layout, read in place by decode and loss launches
            from mrb_b200 import ops
            outs = self.head.run_packed(be, feats)
            proposals = self._select_proposals_fused(be, anchors, None, None, image_sizes, gtp, training, outs=outs)
            if forked:
                prepared = be.join(prepared)
            pos_idx, pos_ok, reg_t, sel, sel_lab, sel_w = prepared
            lo, lb = ops.rpn_loss(outs, self.head.num_anchors, sel, sel_lab, sel_w, pos_idx, pos_ok, reg_t)
            return proposals, {"loss_objectness": lo, "loss_rpn_box_reg": lb}
        logits, deltas = self.head.run(be, feats)
        proposals = self.select_proposals(be, anchors, [l.detach() for l in logits], [d.detach() for d in deltas],
                                          image_sizes, targets, training, gtp=gtp)
        losses = {}
        if training:
            if prepared is not None and forked:
                prepared = be.join(prepared)
            lo, lb = self.loss(anchors_all, vis, logits, deltas, targets, generator, prepared)
            losses = {"loss_objectness": lo, "loss_rpn_box_reg": lb}
        return proposals, losses
