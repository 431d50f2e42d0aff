"""SigmoidFocalLoss (reference layers/sigmoid_focal_loss.py:9-74).

The reference dispatched CUDA tensors to `_C` and CPU tensors to a pure-PyTorch formula
(sigmoid_focal_loss_cpu, :40-50).  That formula is kept under the same name -- it is plain
tensor algebra, device-agnostic, and is what the reference's module runs for CPU inputs -- but the
module only routes *CPU* tensors to it; CUDA tensors always run the sm_100a kernel."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from maskrcnn_benchmark import _C


class _SigmoidFocalLoss(Function):
    @staticmethod
    def forward(ctx, logits, targets, gamma, alpha):
        ctx.save_for_backward(logits, targets)
        ctx.cfg = (logits.shape[1], gamma, alpha)
        return _C.sigmoid_focalloss_forward(logits, targets, logits.shape[1], gamma, alpha)

    @staticmethod
    @once_differentiable
    def backward(ctx, d_loss):
        logits, targets = ctx.saved_tensors
        num_classes, gamma, alpha = ctx.cfg
        d_logits = _C.sigmoid_focalloss_backward(logits, targets, d_loss.contiguous(), num_classes, gamma, alpha)
        return d_logits, None, None, None, None


sigmoid_focal_loss_cuda = _SigmoidFocalLoss.apply


def sigmoid_focal_loss_cpu(logits, targets, gamma, alpha):
    """-(t==c)*alpha*(1-p)^g*log(p) - (t!=c & t>=0)*(1-alpha)*p^g*log(1-p), c in 1..C"""
    classes = torch.arange(1, logits.shape[1] + 1, dtype=targets.dtype, device=targets.device)[None, :]
    t = targets[:, None]
    p = torch.sigmoid(logits)
    pos = (t == classes).float()
    neg = ((t != classes) & (t >= 0)).float()
    return -pos * alpha * (1 - p) ** gamma * torch.log(p) - neg * (1 - alpha) * p ** gamma * torch.log(1 - p)


class SigmoidFocalLoss(nn.Module):
    def __init__(self, gamma, alpha):
        super().__init__()
        self.gamma = gamma
        self.alpha = alpha

    def forward(self, logits, targets):
        loss_func = sigmoid_focal_loss_cuda if logits.is_cuda else sigmoid_focal_loss_cpu
        return loss_func(logits, targets, self.gamma, self.alpha).sum()

    def __repr__(self):
        return "%s(gamma=%s, alpha=%s)" % (self.__class__.__name__, self.gamma, self.alpha)
