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

    def fused(self, x1, x2=None, u=None, cc=None, fa=None, l2norm=False, precision=ops.TENSOR_TF32, fa_neg=None, flang=None, coords=None,
              round_in=True, round_out=False, stage_out=False):
        """x1 [B,K1,N] (+ x2 [B,K2,N]) through the sm_100a kernels; the weight columns beyond K1+K2 (split-weight fusion) act
        on (flang [B,Ct], coords [8,N]) -- or the caller applies them itself and passes the results as u / cc."""
        w = self.conv.weight.view(self.conv.weight.shape[0], -1)
        bn = self.bn
        # num_batches_tracked: nn.BatchNorm2d.forward increments it every training step; the statistics kernel does it here so
        # that state dicts stay interchangeable with a reference-trained checkpoint (no extra launch)
        return ops.conv_bn_act(x1, w, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training, x2=x2, u=u, cc=cc, fa=fa,
                               momentum=bn.momentum, eps=bn.eps, slope=self.slope, l2norm=l2norm, precision=precision, fa_neg=fa_neg,
                               num_batches_tracked=bn.num_batches_tracked, flang=flang, coords=coords, round_in=round_in, round_out=round_out,
                               stage_out=stage_out)


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


class DarknetTap(nn.Module):
    """the backbone hand-off.  The reference's Darknet.forward (model/darknet.py:391-431) runs, on every
    call, the COCO detection heads behind the three taps -- per scale a 3x3 convolution, the 255-channel head convolution and the
    YOLOLayer decode (model/yolov3.cfg:583-607, :669-693, :756-780) -- and then discards them when obj_out=False (:430-431), which
    is how DCNet uses it (model/DCNet_model.py:234, :344): 10 of the 106 blocks, 7 convolutions.  DarknetTap wraps a
    reference-style Darknet instance (anything with .module_defs / .module_list / .obj_out; its parameters stay where they are, so
    state dicts and yolov3.weights loading are unchanged) and executes only the layers the returned taps depend on:

        net.visumodel = DarknetTap(net.visumodel)

    A block is live iff a tap, or a live route / shortcut / sequential successor, reads its output.  With obj_out=True every block is
    live and the YOLOLayer decodes run through ops.yolo_layer_decode (one kernel instead of ~12 elementwise kernels per scale)."""

    def __init__(self, darknet):
        super().__init__()
        self.darknet = darknet
        self.live = self._liveness(darknet.module_defs, bool(getattr(darknet, "obj_out", False)))

    @staticmethod
    def _inputs(i, d):
        t = d["type"]
        if t == "route":
            return [i + l if l < 0 else l for l in (int(v) for v in d["layers"].split(","))]
        if t == "shortcut":
            return [i - 1, i + int(d["from"]) if int(d["from"]) < 0 else int(d["from"])]
        return [i - 1] if i > 0 else []

    @classmethod
    def _liveness(cls, defs, obj_out):
        n = len(defs)
        live = [False] * n
        for i, d in enumerate(defs):
            if d["type"] == "yoloconvolutional" and i > 0:
                live[i - 1] = True                    # the tap is the tensor ENTERING the head convolution (:406-407)
            if obj_out and d["type"] == "yolo":
                live[i] = True
        for i in range(n - 1, -1, -1):                # consumers come after producers: one reverse sweep propagates everything
            if live[i]:
                for j in cls._inputs(i, defs[i]):
                    live[j] = True
        return live

    def skipped_blocks(self):
        return [i for i, l in enumerate(self.live) if not l]

    def forward(self, x, targets=None):
        dk = self.darknet
        if targets is not None:
            return dk(x, targets)                     # the COCO training loss is never reached by DCNet: the reference's own path
        taps, dets, outs = [], [], []
        for i, (d, m) in enumerate(zip(dk.module_defs, dk.module_list)):
            t = d["type"]
            if t == "yoloconvolutional":
                taps.append(x)
            if not self.live[i]:
                outs.append(None)
                continue
            if t in ("convolutional", "upsample", "maxpool", "yoloconvolutional"):
                x = m(x)
            elif t == "route":
                x = torch.cat([outs[j] for j in self._inputs(i, d)], 1)
            elif t == "shortcut":
                a, b = self._inputs(i, d)
                x = outs[a] + outs[b]
            elif t == "yolo":
                layer = m[0]
                x = ops.yolo_layer_decode(x.contiguous(), layer.anchors, layer.num_classes, layer.image_dim) if x.is_cuda else m(x)
                dets.append(x)
            outs.append(x)
        if getattr(dk, "obj_out", False):
            return taps, torch.cat(dets, 1), 0.0, 0.0   # (:428-429) precision / recall are only accumulated in training
        return taps
