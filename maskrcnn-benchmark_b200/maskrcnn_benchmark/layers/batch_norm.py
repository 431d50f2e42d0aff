"""FrozenBatchNorm2d (reference layers/batch_norm.py:6-31): y = x*scale + shift with
scale = weight * rsqrt(running_var) -- NO epsilon -- and shift = bias - running_mean*scale.
Buffer names (weight, bias, running_mean, running_var) are part of the checkpoint format."""
import torch
from torch import nn


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def scale_shift(self):
        """Per-channel (scale, shift) in fp32: what the conv engine fuses into its epilogue."""
        scale = self.weight.float() * self.running_var.float().rsqrt()
        shift = self.bias.float() - self.running_mean.float() * scale
        return scale, shift

    def forward(self, x):
        if x.dtype == torch.float16:
            # batch_norm.py:21-25 casts the buffers themselves to half
            self.weight = self.weight.half()
            self.bias = self.bias.half()
            self.running_mean = self.running_mean.half()
            self.running_var = self.running_var.half()
        scale = self.weight * self.running_var.rsqrt()
        shift = self.bias - self.running_mean * scale
        return x * scale.reshape(1, -1, 1, 1) + shift.reshape(1, -1, 1, 1)
