"""ResNet-50 + FPN trunk (reference modeling/backbone/resnet.py:81-152,239-398; fpn.py:7-79;
backbone.py:23-46).  Module/parameter names reproduce the reference's state_dict keys; the forward
pass routes every conv (+FrozenBN +ReLU +residual) through Backend.conv as ONE fused call."""
import torch
from torch import nn

from maskrcnn_benchmark.layers import Conv2d, DFConv2d, FrozenBatchNorm2d


def _kaiming_uniform(conv):
    nn.init.kaiming_uniform_(conv.weight, a=1)
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)


class FrozenAffine:
    """Cached (scale, shift) of a FrozenBatchNorm2d (layers/batch_norm.py:27-31), fp32, recomputed only
    if the buffers change (they never do during training: the statistics are frozen)."""

    def __init__(self, bn):
        self.bn = bn
        self._key = None
        self._val = None

    def get(self):
        bn = self.bn
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, bn.weight.device)
        if key != self._key:
            self._val = tuple(t.contiguous() for t in bn.scale_shift())
            self._key = key
        return self._val


class Bottleneck(nn.Module):
    """resnet.py:239-344, frozen BN; the 3x3 is dense, grouped (ResNeXt) or deformable (DFConv2d)."""

    def __init__(self, cin, mid, cout, stride, stride_in_1x1=True, groups=1, dcn=None):
        super().__init__()
        self.downsample = None
        if cin != cout:
            self.downsample = nn.Sequential(Conv2d(cin, cout, 1, stride=stride, bias=False), FrozenBatchNorm2d(cout))
            _kaiming_uniform(self.downsample[0])
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(cin, mid, 1, stride=s1, bias=False)
        self.bn1 = FrozenBatchNorm2d(mid)
        if dcn is not None:
            # resnet.py:286-300 (DFConv2d initialises its own offset conv; the deformable weight keeps its default init)
            self.conv2 = DFConv2d(mid, mid, with_modulated_dcn=bool(dcn), kernel_size=3, stride=s3, groups=groups, dilation=1,
                                  deformable_groups=1, bias=False)
        else:
            self.conv2 = Conv2d(mid, mid, 3, stride=s3, padding=1, bias=False, groups=groups)
        self.general = dcn is not None or groups > 1
        self.bn2 = FrozenBatchNorm2d(mid)
        self.conv3 = Conv2d(mid, cout, 1, bias=False)
        self.bn3 = FrozenBatchNorm2d(cout)
        for c in (self.conv1, self.conv3) + ((self.conv2,) if dcn is None else ()):
            _kaiming_uniform(c)
        self.strides = (s1, s3, stride)
        self._aff = [FrozenAffine(b) for b in (self.bn1, self.bn2, self.bn3)]
        self._aff_d = FrozenAffine(self.downsample[1]) if self.downsample is not None else None

    def run(self, be, x):
        if getattr(self, "general", False):
            return be.bottleneck_general(self, x)
        if hasattr(be, "bottleneck") and self.strides[1] == 1:
            # every consumer of a bottleneck output (next block, FPN lateral) pre-masks the gradient it returns
            return be.bottleneck(self, x, g_premasked=getattr(self, "_g_premasked", True))
        s1, s3, sd = self.strides
        (a1, b1), (a2, b2), (a3, b3) = (a.get() for a in self._aff)
        y = be.conv(x, self.conv1.weight, a1, b1, stride=s1, relu=True)
        y = be.conv(y, self.conv2.weight, a2, b2, stride=s3, pad=1, relu=True)
        if self.downsample is not None:
            ad, bd = self._aff_d.get()
            idn = be.conv(x, self.downsample[0].weight, ad, bd, stride=sd)
        else:
            idn = x
        return be.conv(y, self.conv3.weight, a3, b3, residual=idn, relu=True)


class Stem(nn.Module):
    """resnet.py:347-366: conv 7x7/2 -> FrozenBN -> ReLU -> maxpool 3x3/2."""

    def __init__(self, out_channels):
        super().__init__()
        self.conv1 = Conv2d(3, out_channels, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = FrozenBatchNorm2d(out_channels)
        _kaiming_uniform(self.conv1)
        self._aff = FrozenAffine(self.bn1)

    def run(self, be, images):
        a, b = self._aff.get()
        y = be.stem(images, self.conv1.weight, a, b)
        return be.max_pool(y, 3, 2, 1)


class ResNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.stem = Stem(cfg.stem_out)
        cin, mid, cout = cfg.stem_out, cfg.width_per_group * cfg.num_groups, cfg.res2_out
        self.stage_names = []
        for i, n in enumerate(cfg.stage_blocks):
            blocks = []
            dcn = (cfg.with_modulated_dcn if cfg.stage_with_dcn[i] else None)
            for j in range(n):
                stride = 2 if (i > 0 and j == 0) else 1
                blocks.append(Bottleneck(cin, mid, cout, stride, cfg.stride_in_1x1, cfg.num_groups, dcn))
                cin = cout
            name = "layer%d" % (i + 1)
            self.add_module(name, nn.Sequential(*blocks))
            self.stage_names.append(name)
            mid, cout = mid * 2, cout * 2
        self.freeze_at = cfg.freeze_at
        # resnet.py:134-143: freeze stem (+ layer1 ... layer{freeze_at-1})
        for idx in range(cfg.freeze_at):
            m = self.stem if idx == 0 else getattr(self, "layer%d" % idx)
            for p in m.parameters():
                p.requires_grad = False

    def run(self, be, images):
        x = self.stem.run(be, images)
        outs = []
        sb = getattr(be, "stage_boundary", None)
        for name in self.stage_names:
            if sb is not None and x.requires_grad:
                x = sb([x], "backbone.body.%s." % name)[0]     # backward reaching x <=> this stage's gradients are issued
            for blk in getattr(self, name):
                x = blk.run(be, x)
            outs.append(x)
        return outs


class FPN(nn.Module):
    """fpn.py:7-79 with LastLevelMaxPool; conv_with_kaiming_uniform() convs (bias, no norm/relu)."""

    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        self.inner_blocks, self.layer_blocks = [], []
        for idx, cin in enumerate(in_channels_list, 1):
            inner, layer = "fpn_inner%d" % idx, "fpn_layer%d" % idx
            ib = Conv2d(cin, out_channels, 1)
            lb = Conv2d(out_channels, out_channels, 3, 1, 1)
            for c in (ib, lb):
                _kaiming_uniform(c)
            self.add_module(inner, ib)
            self.add_module(layer, lb)
            self.inner_blocks.append(inner)
            self.layer_blocks.append(layer)

    def run(self, be, feats):
        last = None
        results = []
        for feat, inner, layer in zip(feats[::-1], self.inner_blocks[::-1], self.layer_blocks[::-1]):
            ib, lb = getattr(self, inner), getattr(self, layer)
            # lateral + top-down add, fused; `feat` is a bottleneck (ReLU) output: hand back a pre-masked gradient
            last = be.lateral_topdown(feat, ib.weight, ib.bias, last)
            results.insert(0, be.conv(last, lb.weight, bias=lb.bias, pad=1))
        results.append(be.max_pool(results[-1], 1, 2, 0))  # P6 (fpn.py:77-79)
        return results


class Backbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.body = ResNet(cfg)
        c2 = cfg.res2_out
        self.fpn = FPN([c2, c2 * 2, c2 * 4, c2 * 8], cfg.fpn_out)
        self.out_channels = cfg.fpn_out

    def run(self, be, images):
        feats = self.body.run(be, images)
        sb = getattr(be, "stage_boundary", None)
        if sb is not None and self.training:
            feats = sb(feats, "backbone.fpn.")       # backward reaching C2..C5 <=> the FPN's gradients are issued
        return self.fpn.run(be, feats)
