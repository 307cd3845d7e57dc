"""The loss / target functions that live in the reference's training script (train_DCNet.py:45-332), with the same
names and positional signatures, backed by the sm_100a kernels.  `from dcnet_b200.losses import *` replaces the
in-script definitions; call configure(size=..., anchor_imsize=..., anchors_full=...) where the script sets its
globals `args` / `anchors_full` (train_DCNet.py:377, :404-406)."""
import types

import torch
import torch.nn.functional as F

from . import ops

ANCHORS_FULL_DEFAULT = [(373.0, 326.0), (156.0, 198.0), (116.0, 90.0), (59.0, 119.0), (62.0, 45.0),
                        (30.0, 61.0), (33.0, 23.0), (16.0, 30.0), (10.0, 13.0)]
args = types.SimpleNamespace(size=256, anchor_imsize=416)
anchors_full = list(ANCHORS_FULL_DEFAULT)

__all__ = ["yolo_loss", "offset_loss", "confidence_loss", "Interframe_contrastive_loss", "Crossmodal_constrastive_loss",
           "rank_loss", "loc_loss", "build_target", "configure", "fused_losses", "negative_sim_score", "decode_boxes"]


def configure(size=None, anchor_imsize=None, anchors_full=None):
    if size is not None:
        args.size = int(size)
    if anchor_imsize is not None:
        args.anchor_imsize = anchor_imsize
    if anchors_full is not None:
        globals()["anchors_full"] = [tuple(a) for a in anchors_full]


def _as_index(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.long)
    if len(x) and torch.is_tensor(x[0]):
        return torch.stack([v.reshape(()) for v in x]).to(device=device, dtype=torch.long)
    return torch.tensor(list(x), device=device, dtype=torch.long)


def _packed(lst, dim_fix=None):
    if torch.is_tensor(lst):
        return lst                      # already packed [R, ...] (dcnet_b200.model returns packed tensors)
    p = getattr(lst, "packed", None)
    return p if p is not None else torch.stack(list(lst))


class TargetList(list):
    """dense target list that also carries the compact per-sample form used by the fused kernels"""
    compact = None


def build_target(raw_coord, pred):
    """train_DCNet.py:265-332 -> (bbox_list, best_gi, best_gj, best_n_list, bbox_center_list)"""
    best_n, gi, gj, t5, gt, gtc = ops.build_target(raw_coord, args.size, args.anchor_imsize, anchors_full, dense=True)
    gt, gtc = TargetList(gt), TargetList(gtc)
    gt.compact = gtc.compact = (best_n, gi, gj, t5)
    return gt, list(gi.unbind(0)), list(gj.unbind(0)), [int(v) for v in best_n.tolist()], gtc


def _compact(target, gi, gj, best_n_list, device):
    c = getattr(target, "compact", None)
    if c is not None:
        return c
    best_n = _as_index(best_n_list, device)
    gi, gj = _as_index(gi, device), _as_index(gj, device)
    B = best_n.shape[0]
    rows = []
    for b in range(B):
        s, a = int(best_n[b]) // 3, int(best_n[b]) % 3
        t = target[s]
        rows.append(t[b, a, :, gj[b], gi[b]] if t.dim() == 5 else t[b, :, gj[b], gi[b]])
    return best_n, gi, gj, torch.stack(rows).float().contiguous()


def _zeros_like_scores(input):
    B = input[0].shape[0]
    return [torch.zeros(B, p.shape[-1] * p.shape[-2], device=p.device, dtype=p.dtype) for p in input]


def yolo_loss(input, target, gi, gj, best_n_list, w_coord=5., w_neg=1. / 5, size_average=True):
    """train_DCNet.py:45-72.  input: 3 x [B,3,5,g,g]."""
    best_n, gi, gj, t5 = _compact(target, gi, gj, best_n_list, input[0].device)
    z = _zeros_like_scores(input)
    return ops.ground_losses(input, z, z, z, best_n, gi, gj, t5, w_coord=w_coord)[0]


def rank_loss(sim_score, neg_sim_score, target, gi, gj, best_n_list, w_coord=5., w_neg=1. / 5, size_average=True, margin=0.1):
    """train_DCNet.py:173-203"""
    dev = sim_score[0].device
    best_n, gi, gj, t5 = _compact(target, gi, gj, best_n_list, dev)
    B = sim_score[0].shape[0]
    pred = [torch.zeros(B, 15, s.shape[-1] * s.shape[-2], device=dev) for s in sim_score]
    z = [torch.zeros(B, s.shape[-1] * s.shape[-2], device=dev) for s in sim_score]
    return ops.ground_losses(pred, sim_score, neg_sim_score, z, best_n, gi, gj, t5, margin=margin)[1]


def loc_loss(loc_score, sim_score, target):
    """train_DCNet.py:205-220 (sim_score is unused by the reference too)"""
    dev = loc_score[0].device
    c = getattr(target, "compact", None)
    B = loc_score[0].shape[0]
    if c is None:
        gc = torch.cat([t[:, 4].reshape(B, -1) for t in target], 1)
        lc = torch.cat([s.reshape(B, -1) for s in loc_score], 1)
        return F.cross_entropy(lc, gc.max(1)[1])
    best_n, gi, gj, t5 = c
    pred = [torch.zeros(B, 15, s.shape[-1] * s.shape[-2], device=dev) for s in loc_score]
    z = [torch.zeros(B, s.shape[-1] * s.shape[-2], device=dev) for s in loc_score]
    return ops.ground_losses(pred, z, z, loc_score, best_n, gi, gj, t5)[2]


def Interframe_contrastive_loss(q_list, k_list, neg_list, T=0.07):
    """train_DCNet.py:114-136; accepts the reference's python lists or the packed tensors."""
    q, k, neg = _packed(q_list), _packed(k_list), _packed(neg_list)          # [R,P,C], [R,P,C], [R,P,n,C]
    R, P, C = q.shape
    rows = ops.infonce_rows(q.reshape(R * P, C), k.reshape(R * P, C), neg.reshape(R * P, -1, C), T)
    return rows.mean()


def Crossmodal_constrastive_loss(q_list, k_list, neg_list, T=0.07):
    """train_DCNet.py:140-166 (top_k = 1 positives per pixel)"""
    q, k, neg = _packed(q_list), _packed(k_list), _packed(neg_list)          # [R,B,C], [R,B,1,C], [R,B,n,C]
    R, B, C = q.shape
    assert k.shape[2] == 1
    rows = ops.infonce_rows(q.reshape(R * B, C), k.reshape(R * B, C), neg.reshape(R * B, -1, C), T)
    return rows.mean()


def offset_loss(input, target, gi, gj, best_n_list, w_coord=5., w_neg=1. / 5, size_average=True):
    """train_DCNet.py:74-94 (unused by the reference's training loop; kept for API completeness, plain torch)."""
    batch = input[0].size(0)
    pb, gb = [], []
    for ii in range(batch):
        s, a = best_n_list[ii] // 3, best_n_list[ii] % 3
        p = input[s][ii, a, :, gj[ii], gi[ii]]
        pb.append(torch.cat([torch.sigmoid(p[0:2]), p[2:4]]))
        gb.append(target[s][ii, a, :4, gj[ii], gi[ii]])
    pb = torch.stack(pb).view(-1, 2, 4)
    gb = torch.stack(gb).view(-1, 2, 4)
    return sum(F.mse_loss(pb[:, 0, i] - pb[:, 1, i], gb[:, 0, i] - gb[:, 1, i]) for i in range(4)) * w_coord


def confidence_loss(input, target, gi, gj, best_n_list, w_coord=5., w_neg=1. / 5, size_average=True):
    """train_DCNet.py:96-108 (dead code in the reference: it reads an undefined `batch`; here batch = input[0].size(0))."""
    batch = input[0].size(0)
    pc = torch.cat([p[:, :, 4, :, :].contiguous().view(batch, -1) for p in input], 1)
    pc = pc.view(-1, 2, pc.shape[1])
    return F.mse_loss(pc[:, 0, :], pc[:, 1, :])


def negative_sim_score(flang_attn, corr_feat):
    """train_DCNet.py:623-627 (plain restatement for callers that do not use the fused model.last_neg_sim_score)."""
    fa = flang_attn.flip(0)
    return [(fa * c[:, :512]).sum(1) for c in corr_feat]


_LOSS_W = {}


def fused_losses(pred_anchor, sim_score, neg_sim_score, loc_score, bbox, q_if, k_if, neg_if, q_cm, k_cm, neg_cm, target=None, partner3=None,
                 l_if=None, l_cm=None):
    """train_DCNet.py:615-642 in one pass: targets + the three grounding losses from one kernel + the two InfoNCE losses.
    Returns (loss, dict of the five components, (best_n, gi, gj, t5))."""
    if target is None:
        best_n, gi, gj, t5, _, _ = ops.build_target(bbox, args.size, args.anchor_imsize, anchors_full, dense=False)
    else:
        best_n, gi, gj, t5 = target
    g = ops.ground_losses(pred_anchor, sim_score, neg_sim_score, loc_score, best_n, gi, gj, t5, partner3=partner3)
    # l_if / l_cm: the two InfoNCE losses when the caller has already computed them (on another stream, dcnet_b200/hotpath.py)
    if l_if is None:
        l_if = Interframe_contrastive_loss(q_if, k_if, neg_if)
    if l_cm is None:
        l_cm = Crossmodal_constrastive_loss(q_cm, k_cm, neg_cm)
    # train_DCNet.py:642: loss = yolo + 100 rank + loc + 100 interframe + cross.  One dot product instead of three selects and four
    # adds keeps a dozen two-microsecond kernels (forward and SelectBackward) off the path between the losses and the backward.
    key = str(g.device)
    if key not in _LOSS_W:
        _LOSS_W[key] = torch.tensor([1.0, 100.0, 1.0], device=g.device)
    loss = torch.dot(g, _LOSS_W[key]) + (100 * l_if + l_cm)
    gd = g.detach()
    return loss, dict(yolo=gd[0], rank=gd[1], loc=gd[2], interframe=l_if, cross=l_cm), (best_n, gi, gj, t5)


def decode_boxes(pred_anchor, bbox=None, cell=None):
    """a18: cell=(best_n,gi,gj) -> train-time decode at the GT cell (train_DCNet.py:656-677); cell=None -> arg-max decode
    (:766-816).  Returns (boxes xyxy [B,4], iou [B] or None, best_n, gi, gj)."""
    bn, gi, gj = cell if cell is not None else (None, None, None)
    flat = [p.reshape(p.shape[0], 15, -1) for p in pred_anchor]
    return ops.decode(flat, args.size, args.anchor_imsize, anchors_full, bn, gi, gj, bbox)
