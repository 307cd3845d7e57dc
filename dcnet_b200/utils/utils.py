"""Mirror of the hot-path helpers of the reference's utils/utils.py: bbox_iou (:76-104), xywh2xyxy (:34-40),
xyxy2xywh (:25-31), AverageMeter (:8-23)."""
import torch

from .. import ops


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def xyxy2xywh(x):
    return torch.stack([(x[:, 0] + x[:, 2]) / 2, (x[:, 1] + x[:, 3]) / 2, x[:, 2] - x[:, 0], x[:, 3] - x[:, 1]], 1)


def xywh2xyxy(x):
    return torch.stack([x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2, x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2], 1)


def bbox_iou(box1, box2, x1y1x2y2=True):
    """IoU of boxes [n,4] vs [n,4] (or [1,4] broadcast against [n,4], as build_target uses it, train_DCNet.py:303):
    clamp(.,0) intersection, +1e-16 in the denominator, no +1 pixel convention (utils/utils.py:76-104)."""
    return ops.bbox_iou(box1.cuda() if not box1.is_cuda else box1, box2.cuda() if not box2.is_cuda else box2, x1y1x2y2)
