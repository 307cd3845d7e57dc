"""The hot path on its own: everything between the Darknet feature maps and the scalar loss that
this package replaces (a2-a18), with the stays-PyTorch neighbours -- Darknet, text encoder, the 3x3 head
(fcn_emb[s][1:], fcn_out) and the location branch -- represented by tensors supplied by the caller:

    raw[s]      [B,C_s,h_s,w_s]   Darknet maps                       (input, needs grad -> backbone)
    flang, fa   [B,512]           sentence / attended phrase vectors  (inputs, need grad -> text encoder)
    context     [B,T,1024]        BiLSTM outputs
    head[s]     [B,15,N_s]        what fcn_out(fcn_emb[1:](y)) returns (input, needs grad -> head)
    loc[s]      [B,N_s]           location scores
    dy_head[s]  [B,512,N_s]       the gradient the head's backward sends into the fusion output y
    bbox        [B,4]

One step = forward of a2-a11, targets + the five losses (a12-a17), train-time decode + IoU (a18, a15), and the backward
of all of it.  bench.py captures step() in a CUDA graph; the tests compare it with oracle.hotpath_restated()."""
import torch
import torch.nn as nn

from . import losses as LS
from . import ops, parallel
from .model.DCNet_model import CROSS_NEG_N, NEG_N, TOP_K, grounding_model


class _NoBackbone(nn.Module):
    def forward(self, x):
        raise RuntimeError("HotPath has no backbone: feed raw feature maps to step()")


class HotPath(nn.Module):
    def __init__(self, size=256, vocab=1000, cross_gpu_negatives=False):
        super().__init__()
        self.size = size
        # BASELINE config 5: rank-loss / pixel-to-text negatives taken from the GLOBAL batch (partner Bg-1-g) through an
        # all-gather of the text vectors and target cells (dcnet_b200/parallel.py); off = the reference's local reversal
        self.cross_gpu_negatives = cross_gpu_negatives
        # the mirror model supplies the hot-path parameters (same names/shapes/init as the reference)
        self.net = grounding_model(corpus=list(range(vocab)), emb_size=512, visumodel=_NoBackbone(), size=size)
        self.grids = [size // 32, size // 16, size // 8]
        self.hot_parameters = [p for n, p in self.net.named_parameters()
                               if n.startswith(("mapping_visu", "corr_conv")) or ".0.conv" in n and n.startswith("fcn_emb")
                               or ".0.bn" in n and n.startswith("fcn_emb")]

    def draw_indices(self, B):
        """host side of the sampling blocks: exact reference RNG stream -> (negpos [P,30,10] int32, negidx [B,N0,5] int64) numpy"""
        N0 = self.grids[0] ** 2
        return ops.pyrandom_interframe(B // 2, TOP_K, N0, NEG_N), ops.pyrandom_crossmodal(B, N0, CROSS_NEG_N)

    def forward_losses(self, raw, flang, fa, context, head, loc, bbox, negpos=None, negidx=None):
        net = self.net
        LS.configure(size=self.size)
        hw = [(m.shape[2], m.shape[3]) for m in raw]
        fv = net.map_visual(raw)
        q_if, k_if, neg_if, idx_if, _ = net.interframe(fv[0], negpos)
        best_n, gi, gj, t5, _, _ = ops.build_target(bbox, self.size, LS.args.anchor_imsize, LS.anchors_full)
        fa_neg = partner3 = None
        if self.cross_gpu_negatives:
            fa_neg, partner3 = parallel.global_partners(fa, best_n, gi, gj)
        corr, sim, neg_sim = net.correspondence(fv, fa, fa_neg)
        coords = [ops.coord_map(h, w, fa.device).flatten(1) for (h, w) in hw]
        y = net.fuse(corr, flang, coords)
        oo_obj = [ops.only_obj(head[s], sim[s]) for s in range(3)]
        pred = [ops.modulate_conf(head[s], sim[s], loc[s]) for s in range(3)]
        q_cm, k_cm, neg_cm, word, _ = net.crossmodal(fv[0], context, negidx)
        loss, comp, cell = LS.fused_losses(pred, sim, neg_sim, loc, bbox, q_if, k_if, neg_if, q_cm, k_cm, neg_cm,
                                           target=(best_n, gi, gj, t5), partner3=partner3)
        boxes, iou, _, _, _ = LS.decode_boxes(pred, bbox, cell[:3])
        return dict(loss=loss, comp=comp, y=y, iou=iou, boxes=boxes, cell=cell, corr=corr, sim=sim, pred=pred,
                    obj=[o[1] for o in oo_obj], idx_if=idx_if, word=word)

    def step(self, raw, flang, fa, context, head, loc, dy_head, bbox, negpos=None, negidx=None):
        """forward + backward.  Returns a [6+B] tensor: (loss, yolo, rank, loc, interframe, cross, iou[0..B))."""
        out = self.forward_losses(raw, flang, fa, context, head, loc, bbox, negpos, negidx)
        torch.autograd.backward([out['loss']] + list(out['y']), [None] + list(dy_head))
        c = out['comp']
        return torch.cat([torch.stack([out['loss'], c['yolo'], c['rank'], c['loc'], c['interframe'], c['cross']]).detach(), out['iou']])
