"""Mirror of the pieces of the reference's model/darknet.py that sit on the hot path:
ConvBatchNormReLU (model/darknet.py:118-156) and the YOLOLayer inference decode (:245-296, :365-375).
The Darknet-53 conv stack itself is a generic cuDNN backbone and stays the reference's own module
(SURVEY.md section 2, row 5): grounding_model takes it by injection."""
import torch
import torch.nn as nn

from .. import ops


class ConvBatchNormReLU(nn.Sequential):
    """Same sub-module names / parameter shapes as the reference (conv.weight, bn.*).  1x1 instances on the hot
    path are executed by ops.conv_bn_act (see grounding_model); called as a module it is plain PyTorch."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, leaky=False, relu=True):
        super().__init__()
        self.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, bias=False))
        self.add_module("bn", nn.BatchNorm2d(out_channels, eps=1e-5, momentum=0.999, affine=True))
        self.slope = 0.1 if leaky else 0.0
        if leaky:
            self.add_module("relu", nn.LeakyReLU(0.1))
        elif relu:
            self.add_module("relu", nn.ReLU())

    def fused(self, x1, x2=None, u=None, cc=None, fa=None, l2norm=False, precision=ops.TENSOR_TF32, fa_neg=None):
        """x1 [B,K1,N] (+ x2 [B,K2,N]) through the sm_100a kernels; the weight columns beyond K1+K2 (split-weight fusion)
        are applied by the caller through u / cc."""
        w = self.conv.weight.view(self.conv.weight.shape[0], -1)
        bn = self.bn
        # num_batches_tracked: nn.BatchNorm2d.forward increments it every training step; the statistics kernel does it here so
        # that state dicts stay interchangeable with a reference-trained checkpoint (no extra launch)
        return ops.conv_bn_act(x1, w, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training, x2=x2, u=u, cc=cc, fa=fa,
                               momentum=bn.momentum, eps=bn.eps, slope=self.slope, l2norm=l2norm, precision=precision, fa_neg=fa_neg,
                               num_batches_tracked=bn.num_batches_tracked)


class YOLOLayer(nn.Module):
    """Inference decode of the COCO head (model/darknet.py:262-296, 365-375) as one kernel."""

    def __init__(self, anchors, num_classes, img_dim):
        super().__init__()
        self.anchors = anchors
        self.num_anchors = len(anchors)
        self.num_classes = num_classes
        self.bbox_attrs = 5 + num_classes
        self.image_dim = img_dim

    def forward(self, x, targets=None):
        if targets is not None:
            raise NotImplementedError("YOLOLayer training loss is never reached by DCNet (targets=None, model/darknet.py:409-418)")
        return ops.yolo_layer_decode(x, self.anchors, self.num_classes, self.image_dim)
