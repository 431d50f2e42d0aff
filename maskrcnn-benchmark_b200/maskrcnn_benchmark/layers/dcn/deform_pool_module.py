"""DeformRoIPooling modules (reference layers/dcn/deform_pool_module.py:6-150).  No model in
modeling/ instantiates them; they exist for API completeness."""
from torch import nn

from .deform_pool_func import deform_roi_pooling


class DeformRoIPooling(nn.Module):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None,
                 sample_per_part=4, trans_std=.0):
        super().__init__()
        self.spatial_scale = spatial_scale
        self.out_size = out_size
        self.out_channels = out_channels
        self.no_trans = no_trans
        self.group_size = group_size
        self.part_size = out_size if part_size is None else part_size
        self.sample_per_part = sample_per_part
        self.trans_std = trans_std

    def _pool(self, data, rois, offset, no_trans):
        return deform_roi_pooling(data, rois, offset, self.spatial_scale, self.out_size, self.out_channels,
                                  no_trans, self.group_size, self.part_size, self.sample_per_part,
                                  self.trans_std)

    def forward(self, data, rois, offset):
        if self.no_trans:
            offset = data.new_empty(0)
        return self._pool(data, rois, offset, self.no_trans)


def _offset_mlp(out_size, out_channels, hidden, out_per_bin):
    return nn.Sequential(
        nn.Linear(out_size * out_size * out_channels, hidden), nn.ReLU(inplace=True),
        nn.Linear(hidden, hidden), nn.ReLU(inplace=True),
        nn.Linear(hidden, out_size * out_size * out_per_bin))


class DeformRoIPoolingPack(DeformRoIPooling):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None,
                 sample_per_part=4, trans_std=.0, deform_fc_channels=1024):
        super().__init__(spatial_scale, out_size, out_channels, no_trans, group_size, part_size,
                         sample_per_part, trans_std)
        self.deform_fc_channels = deform_fc_channels
        if not no_trans:
            self.offset_fc = _offset_mlp(self.out_size, self.out_channels, deform_fc_channels, 2)
            self.offset_fc[-1].weight.data.zero_()
            self.offset_fc[-1].bias.data.zero_()

    def forward(self, data, rois):
        assert data.size(1) == self.out_channels
        if self.no_trans:
            return self._pool(data, rois, data.new_empty(0), True)
        n = rois.shape[0]
        x = self._pool(data, rois, data.new_empty(0), True)
        offset = self.offset_fc(x.view(n, -1)).view(n, 2, self.out_size, self.out_size)
        return self._pool(data, rois, offset, False)


class ModulatedDeformRoIPoolingPack(DeformRoIPooling):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None,
                 sample_per_part=4, trans_std=.0, deform_fc_channels=1024):
        super().__init__(spatial_scale, out_size, out_channels, no_trans, group_size, part_size,
                         sample_per_part, trans_std)
        self.deform_fc_channels = deform_fc_channels
        if not no_trans:
            self.offset_fc = _offset_mlp(self.out_size, self.out_channels, deform_fc_channels, 2)
            self.offset_fc[-1].weight.data.zero_()
            self.offset_fc[-1].bias.data.zero_()
            self.mask_fc = nn.Sequential(
                nn.Linear(self.out_size * self.out_size * self.out_channels, deform_fc_channels),
                nn.ReLU(inplace=True),
                nn.Linear(deform_fc_channels, self.out_size * self.out_size * 1),
                nn.Sigmoid())
            self.mask_fc[2].weight.data.zero_()
            self.mask_fc[2].bias.data.zero_()

    def forward(self, data, rois):
        assert data.size(1) == self.out_channels
        if self.no_trans:
            return self._pool(data, rois, data.new_empty(0), True)
        n = rois.shape[0]
        x = self._pool(data, rois, data.new_empty(0), True)
        offset = self.offset_fc(x.view(n, -1)).view(n, 2, self.out_size, self.out_size)
        mask = self.mask_fc(x.view(n, -1)).view(n, 1, self.out_size, self.out_size)
        return self._pool(data, rois, offset, False) * mask
