"""Mirror of the reference's utils/losses.py (IoULoss, :11-34)."""
import torch.nn as nn

from .. import ops


class IoULoss(nn.Module):
    def __init__(self, size_average=True):
        super().__init__()
        self.size_average = size_average

    def forward(self, input, target):
        return ops.iou_loss(input, target, self.size_average)
