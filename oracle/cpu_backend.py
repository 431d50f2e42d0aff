"""CPU *checker* backend for the harness model: plain PyTorch fp32 for dense math + the oracle for
ROIAlign / NMS (the reference's own compiled CPU kernels, oracle/_ref, when `use_ref` and they exist).
Test / baseline infrastructure only: used by tests/ and by bench.py's cpu_baseline and
`--impl reference` legs.  The product backend is mrb_b200.model.backend.B200Backend (no CPU path)."""
import torch
import torch.nn.functional as F
from torch.autograd import Function

import oracle
from mrb_b200.model.backend import Backend


class _OracleRoiAlign(Function):
    @staticmethod
    def forward(ctx, feat, rois, scale, p, s, ref=None):
        ctx.save_for_backward(rois)
        ctx.cfg = (scale, p, s, tuple(feat.shape))
        if ref is not None:  # the reference's own ROIAlign_forward_cpu
            return ref.roi_align_forward(feat.detach().contiguous(), rois, scale, p, p, s)
        return oracle.roi_align_forward(feat.detach().contiguous(), rois, scale, p, p, s)

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        scale, p, s, shape = ctx.cfg
        return oracle.roi_align_backward(g.contiguous(), rois, scale, p, p, *shape, s), None, None, None, None, None


class CpuCheckerBackend(Backend):
    name = "cpu-checker"
    act_dtype = torch.float32

    def __init__(self, use_ref=False):
        self.ref = oracle.ref() if use_ref else None

    def prepare_input(self, images):
        return images

    @staticmethod
    def _affine(y, scale, shift, bias, residual, relu):
        if scale is not None:
            y = y * scale[None, :, None, None]
        add = shift if shift is not None else bias
        if add is not None:
            y = y + add[None, :, None, None]
        if residual is not None:
            y = y + residual
        return F.relu(y) if relu else y

    def conv(self, x, weight, scale=None, shift=None, bias=None, residual=None, stride=1, pad=0, relu=False,
             out_fp32=False, premask_x=False, gy_premasked=False):
        return self._affine(F.conv2d(x, weight, None, stride, pad), scale, shift, bias, residual, relu)

    def stem(self, images, weight, scale, shift):
        return self._affine(F.conv2d(images, weight, None, 2, 3), scale, shift, None, None, True)

    # grouped (ResNeXt) bottleneck variant: fp32 ATen grouped conv (checker of B200Backend.bottleneck_general)
    def bottleneck_general_ok(self, mod):
        c2 = mod.conv2
        return isinstance(c2, torch.nn.Conv2d) and c2.groups > 1 and c2.bias is None and tuple(c2.kernel_size) == (3, 3)

    def bottleneck_general(self, blk, x):
        s1, s3, sd = blk.strides
        (a1, b1), (a2, b2), (a3, b3) = (a.get() for a in blk._aff)
        y = self.conv(x, blk.conv1.weight, a1, b1, stride=s1, relu=True)
        c2 = blk.conv2
        y = self._affine(F.conv2d(y, c2.weight, None, s3, 1, 1, c2.groups), a2, b2, None, None, True)
        if blk.downsample is not None:
            ad, bd = blk._aff_d.get()
            idn = self.conv(x, blk.downsample[0].weight, ad, bd, stride=sd)
        else:
            idn = x
        return self.conv(y, blk.conv3.weight, a3, b3, residual=idn, relu=True)

    def max_pool(self, x, k, s, p):
        return F.max_pool2d(x, k, s, p)

    def upsample2x(self, x):
        return F.interpolate(x, scale_factor=2, mode="nearest")

    def linear(self, x, weight, bias, relu=False, out_fp32=False, premask_x=False, gy_premasked=False):
        y = F.linear(x, weight, bias)
        return F.relu(y) if relu else y

    def deconv2x2(self, x, weight, bias, relu=False, premask_x=False, gy_premasked=False):
        y = F.conv_transpose2d(x, weight, bias, 2, 0)
        return F.relu(y) if relu else y

    def roi_align_fpn(self, feats, rois, scales, pooled, sampling_ratio, nhwc):
        # LevelMapper, modeling/poolers.py:31-42
        area = (rois[:, 3] - rois[:, 1] + 1) * (rois[:, 4] - rois[:, 2] + 1)
        lv = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-6)).clamp(2, 5).long() - 2
        out = feats[0].new_zeros((rois.shape[0], feats[0].shape[1], pooled, pooled))
        for l, (f, sc) in enumerate(zip(feats, scales)):
            idx = (lv == l).nonzero().squeeze(1)
            if idx.numel():
                out = out.index_put((idx,), _OracleRoiAlign.apply(f, rois[idx].contiguous(), sc, pooled, sampling_ratio, self.ref))
        return out

    def nms_batched(self, boxes, scores, sizes, thr):
        keep = torch.zeros(max(int(sum(sizes)), 1), dtype=torch.int64)
        counts = torch.zeros(max(len(sizes), 1), dtype=torch.int32)
        off = 0
        for p, n in enumerate(sizes):
            if n:
                b, sc = boxes[off:off + n].contiguous(), scores[off:off + n].contiguous()
                if self.ref is not None:  # the reference's own nms_cpu
                    k = self.ref.nms(b, sc, float(thr))
                else:
                    k = oracle.nms(b, sc, thr, order=torch.sort(sc, stable=True, descending=True)[1])
                keep[off:off + len(k)] = k
                counts[p] = len(k)
            off += n
        return keep[:off], counts[:len(sizes)]
