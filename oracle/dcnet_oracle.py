"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DCNet dense-correspondence hot path.

A restatement, in plain PyTorch (CPU, fp32 or fp64), of the algorithm of
mengcaopku/DCNet between the Darknet feature maps and the scalar losses.  Every
function cites the reference file:line it follows (paths relative to the
reference checkout).  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import this module; the product path
(dcnet_b200/) never does.

PINNING: the reference ships no tests or golden vectors, so the
oracle is pinned against the *reference itself*, imported unmodified in the build
container (oracle/ref_loader.py): tests/test_oracle_vs_reference.py asserts equality
with model/DCNet_model.py::grounding_model.forward and the train_DCNet.py losses on
seeded inputs, and tests/golden/make_golden.py stores reference outputs as committed
fixtures (tests/golden/*.pt) that travel to the GPU box.

The blocks are written for any input size (sum of positions SN), so the 416x416
configurations -- which the unmodified reference cannot run because of its
hard-wired 1344 (model/DCNet_model.py:259,584) -- use the same code that is proven
equal to the reference at 256x256.
"""
import math
import random as _pyrandom

import numpy as np
import torch
import torch.nn.functional as F

TAU = 10.0          # model/DCNet_model.py:251
TOP_K = 30          # model/DCNet_model.py:391
NEG_N = 10          # model/DCNet_model.py:392
CROSS_NEG_N = 5     # model/DCNet_model.py:53
T_NCE = 0.07        # train_DCNet.py:114,140
BN_EPS = 1e-5       # model/darknet.py:145
BN_MOMENTUM = 0.999  # model/darknet.py:145
ANCHORS_FULL = [(373.0, 326.0), (156.0, 198.0), (116.0, 90.0), (59.0, 119.0), (62.0, 45.0),
                (30.0, 61.0), (33.0, 23.0), (16.0, 30.0), (10.0, 13.0)]  # train_DCNet.py:404-406 (reversed COCO)


# ----------------------------------------------------------------------------------------------
# a1  ConvBatchNormReLU (1x1)                                  model/darknet.py:118-156
# ----------------------------------------------------------------------------------------------
def conv1x1_bn_relu(x, weight, gamma, beta, running_mean=None, running_var=None, training=True,
                    leaky=False, update_running=False, relu_mask=None):
    """x [B,K,N] (flattened NCHW), weight [C,K].  Train mode: batch statistics over (B,N), biased
    variance for the normalisation; running stats (momentum 0.999, unbiased var) updated in place
    when update_running.  Returns y [B,C,N].
    relu_mask (test aid, bool [B,C,N]): use this activation pattern instead of (y > 0) -- lets a test compare gradients
    at the SAME ReLU pattern as the implementation under test (the derivative of the function it actually evaluated)."""
    z = torch.einsum('ck,bkn->bcn', weight, x)
    if training:
        mean = z.mean(dim=(0, 2))
        var = z.var(dim=(0, 2), unbiased=False)
        if update_running and running_mean is not None:
            n = z.shape[0] * z.shape[2]
            with torch.no_grad():
                running_mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
                running_var.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * n / max(n - 1, 1))
    else:
        mean, var = running_mean, running_var
    y = (z - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS) * gamma[None, :, None] + beta[None, :, None]
    if relu_mask is not None:
        assert not leaky
        return y * relu_mask.to(y.dtype)
    return F.leaky_relu(y, 0.1) if leaky else F.relu(y)


def l2norm_channels(x, eps=1e-12):
    """F.normalize(x, p=2, dim=1) on [B,C,N]                   model/DCNet_model.py:359,469"""
    return x / x.norm(dim=1, keepdim=True).clamp_min(eps)


# ----------------------------------------------------------------------------------------------
# a7  generate_coord                                           model/DCNet_model.py:23-39
# ----------------------------------------------------------------------------------------------
def coord_map(h, w, dtype=torch.float32, device=None):
    """[8,h,w].  NOTE the transposed convention of the reference: meshgrid('ij') of (arange(h), arange(w))
    puts the ROW index in the 'x' channels, normalised by the WIDTH."""
    r = torch.arange(h, dtype=dtype, device=device)[:, None].expand(h, w)
    c = torch.arange(w, dtype=dtype, device=device)[None, :].expand(h, w)
    x0 = (r * 2 - w) / w
    y0 = (c * 2 - h) / h
    x1 = ((r + 1) * 2 - w) / w
    y1 = ((c + 1) * 2 - h) / h
    return torch.stack([x0, y0, x1, y1, (x0 + x1) / 2, (y0 + y1) / 2,
                        torch.full((h, w), 1.0 / h, dtype=dtype, device=device),
                        torch.full((h, w), 1.0 / w, dtype=dtype, device=device)], 0)


# ----------------------------------------------------------------------------------------------
# a4  inter-frame patch correspondence                         model/DCNet_model.py:381-430
# ----------------------------------------------------------------------------------------------
def canonical_topk(flat, k):
    """Top-k of a 1-D tensor, descending, ties broken by LOWER index first (torch.topk leaves tie
    order unspecified; this is the documented tie rule of the CUDA path)."""
    order = np.lexsort((np.arange(flat.numel()), -flat.detach().cpu().double().numpy()))
    idx = torch.from_numpy(order[:k].copy()).long().to(flat.device)
    return flat[idx], idx


def interframe_sample(f1, f2, rng=_pyrandom, top_k=TOP_K, neg_n=NEG_N, topk_fn=None):
    """f1,f2 [P,C,N0] unit-norm scale-0 maps of frame 1 / frame 2.
    Returns (q [top_k,P,C], k [top_k,P,C], neg [top_k,P,neg_n,C], idx [P,top_k] int64, negidx [P,top_k,neg_n]).
    RNG order: pairs outer, ranks inner, one random.sample per (pair, rank) (Appendix A.3)."""
    P, C, N0 = f1.shape
    S = torch.bmm(f1.transpose(1, 2), f2).flatten(1)            # :390 row-major [N0*N0], row = frame-1 position
    qs, ks, negs, idxs, negidxs = [], [], [], [], []
    for p in range(P):
        if topk_fn is None:
            _, idx = S[p].topk(top_k, dim=0, largest=True, sorted=True)   # :395
        else:
            _, idx = topk_fn(S[p], top_k)
        row, col = idx // N0, idx % N0                           # :407,:409
        nidx = []
        for j in range(top_k):
            pool = list(range(N0))
            pool.remove(int(col[j]))                             # :411-412
            nidx.append(rng.sample(pool, neg_n))                 # :413
        nidx = torch.tensor(nidx, dtype=torch.long, device=f1.device)
        qs.append(f1[p][:, row].t())                             # [top_k,C]
        ks.append(f2[p][:, col].t())
        negs.append(f2[p][:, nidx.flatten()].t().reshape(top_k, neg_n, C))
        idxs.append(idx)
        negidxs.append(nidx)
    q = torch.stack(qs, 1)
    k = torch.stack(ks, 1)
    neg = torch.stack(negs, 1)
    return q, k, neg, torch.stack(idxs), torch.stack(negidxs)


# ----------------------------------------------------------------------------------------------
# a5  co-attention                                             model/DCNet_model.py:449-464
# ----------------------------------------------------------------------------------------------
def coattention(f1, f2, tau=TAU):
    """f1,f2 [P,C,N].  S[i,j]=<f1_i,f2_j>.  O1[:,i]=sum_j f2_j softmax_j(tau S[i,:]) (:455,:458),
    O2[:,j]=sum_i f1_i softmax_i(tau S[:,j]) (:456,:459).  Returns (O1,O2) [P,C,N]."""
    S = torch.bmm(f1.transpose(1, 2), f2)
    A2 = F.softmax(S.transpose(1, 2) * tau, dim=1)               # [j,i], normalised over j
    A1 = F.softmax(S * tau, dim=1)                               # [i,j], normalised over i
    return torch.bmm(f2, A2), torch.bmm(f1, A1)


def interleave_pairs(a1, a2):
    """[P,...] x2 -> [2P,...] in frame order (2p = frame 1, 2p+1 = frame 2)   :464"""
    return torch.stack([a1, a2], 1).reshape((-1,) + tuple(a1.shape[1:]))


# ----------------------------------------------------------------------------------------------
# a9  pixel-to-text similarity                                 model/DCNet_model.py:525-535, train_DCNet.py:623-627
# ----------------------------------------------------------------------------------------------
def pix2text(corr, flang_attn, fa_partner=None):
    """corr [B,C,N], flang_attn [B,C] -> (sim [B,N], neg_sim [B,N]); neg uses the batch-reversed text (train_DCNet.py:623-627),
    or the given partner text vectors [B,C] (global-batch reversal of BASELINE config 5: the partner lives on another rank)."""
    sim = (flang_attn[:, :, None] * corr).sum(1)
    neg = ((flang_attn.flip(0) if fa_partner is None else fa_partner)[:, :, None] * corr).sum(1)
    return sim, neg


# ----------------------------------------------------------------------------------------------
# a11 cross-modal block                                        model/DCNet_model.py:625-637, :41-112
# ----------------------------------------------------------------------------------------------
def crossmodal_features(fvisu0, context, fm_weight, fm_bias):
    """fvisu0 [B,C,N0] (channel-normalised scale-0 map), context [B,T,2C] BiLSTM outputs.
    Returns vit [B,C,N0] (normalised over the SPATIAL axis :629), lag [B,T,C] (nearest 0.5x down-sample
    = even channels :631, normalised over the WORD axis :632), M [B,T,N0] = softmax_words(Conv1d(bmm)) :634-635."""
    vit = F.normalize(fvisu0, dim=2)
    lag = F.normalize(context[:, :, 0::2], dim=1)
    M = torch.bmm(lag, vit)
    M = F.softmax(F.conv1d(M, fm_weight, fm_bias, stride=1, padding=1), dim=1)
    return vit, lag, M


def crossmodal_sample(vit, lag, M, rng=_pyrandom, neg_n=CROSS_NEG_N):
    """Returns (q [N0,B,C], k [N0,B,1,C], neg [N0,B,neg_n,C], word [B,N0] int64, negidx [B,N0,neg_n]).
    Reproduces the reference's RNG consumption: for every (image ii, pixel jj) it draws B samples
    (population N0-1 when index==ii else N0) and keeps only the LAST one (index=B-1), whose
    vectors are read from image B-1 (Appendix A.6 / :81-96)."""
    B, C, N0 = vit.shape
    word = M.argmax(dim=1)                                       # topk(1) over words :48  -> [B,N0]
    negidx = torch.empty(B, N0, neg_n, dtype=torch.long)
    word = word.to(M.device)
    for ii in range(B):
        for jj in range(N0):
            last = None
            for index in range(B):
                pool = list(range(N0))
                if index == ii:
                    pool.remove(jj)
                last = rng.sample(pool, neg_n)
            negidx[ii, jj] = torch.tensor(last)
    negidx = negidx.to(vit.device)
    q = vit.permute(2, 0, 1).contiguous()                        # [N0,B,C]  :70
    k = torch.gather(lag, 1, word[:, :, None].expand(B, N0, C)).permute(1, 0, 2).unsqueeze(2).contiguous()
    neg = vit[B - 1].t()[negidx.reshape(-1)].reshape(B, N0, neg_n, C).permute(1, 0, 2, 3).contiguous()
    return q, k, neg, word, negidx


# ----------------------------------------------------------------------------------------------
# a12/a13 contrastive losses                                   train_DCNet.py:114-166
# ----------------------------------------------------------------------------------------------
def interframe_contrastive_loss(q, k, neg, T=T_NCE):
    """q,k [R,P,C]; neg [R,P,n,C] (packed form of the reference's lists)."""
    qn = F.normalize(q, dim=2)
    kn = F.normalize(k, dim=2)
    nn_ = F.normalize(neg, dim=3)
    l_pos = (qn * kn).sum(2, keepdim=True)
    l_neg = torch.einsum('rpc,rpnc->rpn', qn, nn_)
    logits = torch.cat([l_pos, l_neg], 2) / T
    R, P, L = logits.shape
    ce = F.cross_entropy(logits.reshape(R * P, L), torch.zeros(R * P, dtype=torch.long, device=logits.device), reduction='none')
    return ce.reshape(R, P).mean(1).mean(0)


def crossmodal_contrastive_loss(q, k, neg, T=T_NCE):
    """q [R,B,C]; k [R,B,1,C]; neg [R,B,n,C]."""
    return interframe_contrastive_loss(q, k[:, :, 0, :], neg, T)


# ----------------------------------------------------------------------------------------------
# a15 boxes                                                    utils/utils.py:25-40,76-104
# ----------------------------------------------------------------------------------------------
def bbox_iou(b1, b2):
    ix1 = torch.max(b1[:, 0], b2[:, 0]); iy1 = torch.max(b1[:, 1], b2[:, 1])
    ix2 = torch.min(b1[:, 2], b2[:, 2]); iy2 = torch.min(b1[:, 3], b2[:, 3])
    inter = torch.clamp(ix2 - ix1, 0) * torch.clamp(iy2 - iy1, 0)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    return inter / (a1 + a2 - inter + 1e-16)


def xywh2xyxy(x):
    return torch.stack([x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2, x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2], 1)


def xyxy2xywh(x):
    return torch.stack([(x[:, 0] + x[:, 2]) / 2, (x[:, 1] + x[:, 3]) / 2, x[:, 2] - x[:, 0], x[:, 3] - x[:, 1]], 1)


# ----------------------------------------------------------------------------------------------
# a14 build_target                                             train_DCNet.py:265-332
# ----------------------------------------------------------------------------------------------
def scaled_anchors(scale, size, anchor_imsize=416, anchors_full=ANCHORS_FULL):
    grid = size // (32 // (2 ** scale))
    return [(a[0] / (anchor_imsize / grid), a[1] / (anchor_imsize / grid)) for a in anchors_full[3 * scale:3 * scale + 3]]


def build_target(bbox, size, anchor_imsize=416, anchors_full=ANCHORS_FULL):
    """bbox [B,4] xyxy fp32 (already clamped to [0,size-1]).  Returns
    (gt [3 x [B,3,5,g,g]], gi [B] int64, gj [B] int64, best_n [B] int64, gt_center [3 x [B,5,g,g]]).
    All arithmetic in fp32 like the reference (coordinates are fp32 tensors; the 9 anchor IoUs are
    computed by bbox_iou on FloatTensors :299-303; np.argmax = first maximum :305)."""
    B = bbox.shape[0]
    bbox = bbox.float()
    grids = [size // (32 // (2 ** s)) for s in range(3)]
    coords = []
    for s in range(3):
        c = torch.stack([(bbox[:, 0] + bbox[:, 2]) / (2 * size), (bbox[:, 1] + bbox[:, 3]) / (2 * size),
                         (bbox[:, 2] - bbox[:, 0]) / size, (bbox[:, 3] - bbox[:, 1]) / size], 1) * grids[s]
        coords.append(c)
    gt = [torch.zeros(B, 3, 5, g, g) for g in grids]
    gtc = [torch.zeros(B, 5, g, g) for g in grids]
    gi_o = torch.zeros(B, dtype=torch.long); gj_o = torch.zeros(B, dtype=torch.long); bn_o = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        ious = []
        for s in range(3):
            gw, gh = coords[s][b, 2], coords[s][b, 3]
            sa = scaled_anchors(s, size, anchor_imsize, anchors_full)
            gt_box = torch.tensor([[0.0, 0.0, float(gw), float(gh)]], dtype=torch.float32)
            an = torch.tensor([[0.0, 0.0, a[0], a[1]] for a in sa], dtype=torch.float32)  # float64 -> fp32 like FloatTensor(np)
            ious += [float(v) for v in bbox_iou(gt_box, an)]
        best_n = int(np.argmax(np.array(ious, dtype=np.float32)))
        s = best_n // 3
        sa = scaled_anchors(s, size, anchor_imsize, anchors_full)
        gi = coords[s][b, 0].long(); gj = coords[s][b, 1].long()
        tx = coords[s][b, 0] - gi.float(); ty = coords[s][b, 1] - gj.float()
        tw = torch.log(coords[s][b, 2] / sa[best_n % 3][0] + 1e-16)
        th = torch.log(coords[s][b, 3] / sa[best_n % 3][1] + 1e-16)
        v = torch.stack([tx, ty, tw, th, torch.ones(())])
        gt[s][b, best_n % 3, :, gj, gi] = v
        gtc[s][b, :, gj, gi] = v
        gi_o[b], gj_o[b], bn_o[b] = gi, gj, best_n
    return gt, gi_o, gj_o, bn_o, gtc


# ----------------------------------------------------------------------------------------------
# a16/a17 grounding losses                                     train_DCNet.py:45-72,173-220
# ----------------------------------------------------------------------------------------------
def yolo_loss(pred, gt, gi, gj, best_n, w_coord=5.0):
    """pred, gt: 3 x [B,3,5,g,g]."""
    B = pred[0].shape[0]
    pb, gb = [], []
    for b in range(B):
        s, a = int(best_n[b]) // 3, int(best_n[b]) % 3
        p = pred[s][b, a, :, gj[b], gi[b]]
        pb.append(torch.cat([torch.sigmoid(p[0:2]), p[2:4]]))
        gb.append(gt[s][b, a, :4, gj[b], gi[b]])
    pb, gb = torch.stack(pb), torch.stack(gb)
    l_box = sum(F.mse_loss(pb[:, i], gb[:, i]) for i in range(4))
    pc = torch.cat([p[:, :, 4].reshape(B, -1) for p in pred], 1)
    gc = torch.cat([g[:, :, 4].reshape(B, -1) for g in gt], 1)
    return l_box * w_coord + F.cross_entropy(pc, gc.max(1)[1])


def rank_loss(sim, neg_sim, gt_center, margin=0.1, gt_center_partner=None):
    """train_DCNet.py:173-203.  gt_center_partner: dense centre targets of each sample's partner when the partner is not the
    local reversal B-1-b (cross-GPU negatives: the same formula on the concatenated batch)."""
    B = sim[0].shape[0]
    pos = torch.cat([s.reshape(B, -1) for s in sim], 1)
    neg = torch.cat([s.reshape(B, -1) for s in neg_sim], 1)
    gc = torch.cat([g[:, 4].reshape(B, -1) for g in gt_center], 1)
    p = (pos * gc).sum(-1)
    n1 = (neg * gc).sum(-1)
    gp = gc.flip(0) if gt_center_partner is None else torch.cat([g[:, 4].reshape(B, -1) for g in gt_center_partner], 1)
    n2 = (pos * gp).sum(-1)
    return (torch.clamp(margin + n1 - p, 0) + torch.clamp(margin + n2 - p, 0)).sum() / (B * 2)


def loc_loss(loc, gt_center):
    B = loc[0].shape[0]
    lc = torch.cat([s.reshape(B, -1) for s in loc], 1)
    gc = torch.cat([g[:, 4].reshape(B, -1) for g in gt_center], 1)
    return F.cross_entropy(lc, gc.max(1)[1])


def iou_loss(x, target, size_average=True):
    """utils/losses.py:26-34"""
    s = torch.sigmoid(x)
    inter = (s * target).sum()
    union = (s + target - s * target).sum()
    out = x.shape[0] - inter / union
    return out / x.shape[0] if size_average else out


# ----------------------------------------------------------------------------------------------
# a10 objectness / confidence modulation                       model/DCNet_model.py:545-552,612-621
# ----------------------------------------------------------------------------------------------
def only_obj(outbox):
    """outbox [B,15,N] -> mean over the 3 anchors of the conf logit (channel 5a+4)  [B,N]"""
    B, _, N = outbox.shape
    return outbox.reshape(B, 3, 5, N)[:, :, 4].mean(1)


def modulate_conf(outbox, sim, loc):
    """conf logit (channel 5a+4) *= sim*loc                    :619"""
    B, _, N = outbox.shape
    o = outbox.reshape(B, 3, 5, N)
    conf = o[:, :, 4:5] * sim[:, None, None] * loc[:, None, None]
    return torch.cat([o[:, :, :4], conf], 2).reshape(B, 15, N)


# ----------------------------------------------------------------------------------------------
# a18 decode                                                   train_DCNet.py:656-690 (train), :766-816 (eval)
# ----------------------------------------------------------------------------------------------
def decode_at(pred, gi, gj, best_n, size, anchor_imsize=416, anchors_full=ANCHORS_FULL):
    """Train-time decode at the GT cell.  pred 3 x [B,3,5,g,g] -> boxes xyxy [B,4] (pixels)."""
    B = pred[0].shape[0]
    out = torch.zeros(B, 4)
    for b in range(B):
        s, a = int(best_n[b]) // 3, int(best_n[b]) % 3
        stride = 32 // (2 ** s)
        sa = scaled_anchors(s, size, anchor_imsize, anchors_full)
        p = pred[s][b, a, :, gj[b], gi[b]]
        out[b, 0] = torch.sigmoid(p[0]) + gi[b].float()
        out[b, 1] = torch.sigmoid(p[1]) + gj[b].float()
        out[b, 2] = torch.exp(p[2]) * sa[a][0]
        out[b, 3] = torch.exp(p[3]) * sa[a][1]
        out[b] = out[b] * stride
    return xywh2xyxy(out)


def decode_argmax(pred, size, anchor_imsize=416, anchors_full=ANCHORS_FULL):
    """Eval decode: argmax over the concatenated [3*SN] modulated conf logits, first maximum in
    (scale, anchor, gj, gi) order.  Returns (boxes xyxy [B,4], scale, anchor, gj, gi  [B] int64 each)."""
    B = pred[0].shape[0]
    conf = torch.cat([p[:, :, 4].reshape(B, -1) for p in pred], 1)
    mx, loc = conf.max(1)
    boxes = torch.zeros(B, 4)
    S, A, GJ, GI = [torch.zeros(B, dtype=torch.long) for _ in range(4)]
    g0 = size // 32
    for b in range(B):
        l = int(loc[b])
        s = 0 if l < 3 * g0 ** 2 else (1 if l < 3 * g0 ** 2 + 3 * (2 * g0) ** 2 else 2)
        g = size // (32 // (2 ** s)); stride = 32 // (2 ** s)
        sa = scaled_anchors(s, size, anchor_imsize, anchors_full)
        hit = (pred[s][b, :, 4] == mx[b]).nonzero()
        a, gj, gi = [int(v) for v in hit[0]]
        p = pred[s][b, a, :, gj, gi]
        boxes[b, 0] = torch.sigmoid(p[0]) + gi
        boxes[b, 1] = torch.sigmoid(p[1]) + gj
        boxes[b, 2] = torch.exp(p[2]) * sa[a][0]
        boxes[b, 3] = torch.exp(p[3]) * sa[a][1]
        boxes[b] = boxes[b] * stride
        S[b], A[b], GJ[b], GI[b] = s, a, gj, gi
    return xywh2xyxy(boxes), S, A, GJ, GI


# ----------------------------------------------------------------------------------------------
# a19 YOLOLayer (COCO head decode)                             model/darknet.py:262-296,365-375
# ----------------------------------------------------------------------------------------------
def yolo_layer_decode(x, anchors, num_classes=80, image_dim=256):
    """x [B, A*(5+nc), g, g] -> [B, A*g*g, 5+nc]"""
    B, _, g, _ = x.shape
    A = len(anchors)
    stride = image_dim / g
    p = x.view(B, A, 5 + num_classes, g, g).permute(0, 1, 3, 4, 2)
    gx = torch.arange(g, dtype=x.dtype).view(1, 1, 1, g)
    gy = torch.arange(g, dtype=x.dtype).view(1, 1, g, 1)
    aw = torch.tensor([a[0] / (416 / g) for a in anchors], dtype=x.dtype).view(1, A, 1, 1)
    ah = torch.tensor([a[1] / (416 / g) for a in anchors], dtype=x.dtype).view(1, A, 1, 1)
    bx = torch.stack([torch.sigmoid(p[..., 0]) + gx, torch.sigmoid(p[..., 1]) + gy,
                      torch.exp(p[..., 2]) * aw, torch.exp(p[..., 3]) * ah], -1)
    return torch.cat([bx.reshape(B, -1, 4) * stride, torch.sigmoid(p[..., 4]).reshape(B, -1, 1),
                      torch.sigmoid(p[..., 5:]).reshape(B, -1, num_classes)], -1)


# ----------------------------------------------------------------------------------------------
# location branch (stays PyTorch in the product; restated here for any SN) model/DCNet_model.py:556-610
# ----------------------------------------------------------------------------------------------
def location_branch(net, coords, obj_score, context, embedded, word_id):
    """coords: 3 x [8,N_s]; obj_score 3 x [B,N_s].  Returns loc_score_map [B,SN]."""
    B = obj_score[0].shape[0]
    _, flang_loc = net.loc_attn(context, embedded, word_id)
    flang_loc = F.normalize(flang_loc, p=2, dim=1)
    coord_map_ = torch.cat([c.t() for c in coords], 0)[None].expand(B, -1, -1)   # [B,SN,8]
    obj = F.normalize(torch.cat(obj_score, 1), p=2, dim=1)
    SN = obj.shape[1]
    emb = net.loc_embedding(coord_map_.reshape(-1, 8)).reshape(B, SN, -1)
    emb = F.normalize(emb, p=2, dim=2)
    rel = torch.bmm(emb, emb.transpose(1, 2)) * obj[:, None, :]
    rel = net.loc_text_embedding(rel.reshape(-1, SN)).reshape(B, SN, -1).permute(0, 2, 1)
    rel = F.normalize(rel, p=2, dim=1)
    m = (rel * flang_loc[:, :, None]).sum(1)
    mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
    return (m - mn) / (mx - mn + 1e-6)


# ----------------------------------------------------------------------------------------------
# whole forward, restated                                      model/DCNet_model.py:340-650
# ----------------------------------------------------------------------------------------------
def _cbr(mod, x, training, relu_mask=None):
    """Apply a reference-style ConvBatchNormReLU module `mod` (has .conv.weight and .bn.*) as a 1x1 op on [B,K,N]."""
    w = mod.conv.weight.reshape(mod.conv.weight.shape[0], -1)
    return conv1x1_bn_relu(x, w, mod.bn.weight, mod.bn.bias, mod.bn.running_mean, mod.bn.running_var,
                           training=training, update_running=False, relu_mask=relu_mask)


def forward_restated(net, raw_fvisu, word_id, rng=_pyrandom, topk_fn=None, return_internals=False):
    """Restated grounding_model.forward for ANY input size.  `net` is anything exposing the reference's
    sub-module names (the reference instance itself, or dcnet_b200's mirror).  raw_fvisu: 3 x [B,C_s,h_s,w_s].
    BN running statistics are NOT updated (pure function)."""
    training = net.training
    B = raw_fvisu[0].shape[0]
    P = B // 2
    hw = [(m.shape[2], m.shape[3]) for m in raw_fvisu]
    fv = []
    for s in range(3):                                                            # :356-359
        x = raw_fvisu[s].flatten(2)
        fv.append(l2norm_channels(_cbr(net.mapping_visu._modules[str(s)], x, training)))
    C = fv[0].shape[1]
    f1 = [f.reshape(P, 2, C, -1)[:, 0] for f in fv]                               # :365-374
    f2 = [f.reshape(P, 2, C, -1)[:, 1] for f in fv]
    q_if, k_if, neg_if, idx_if, nidx_if = interframe_sample(f1[0], f2[0], rng, topk_fn=topk_fn)   # :381-430
    corr = []
    for s in range(3):                                                            # :449-469
        o1, o2 = coattention(f1[s], f2[s], net.temperature)
        x = interleave_pairs(torch.cat([f1[s], o1], 1), torch.cat([f2[s], o2], 1))
        corr.append(l2norm_channels(_cbr(net.corr_conv._modules[str(s)][0], x, training)))
    max_len = int((word_id != 0).sum(1).max().item())                             # :472-487
    word_id = word_id[:, :max_len]
    raw_flang, context, embedded = net.textmodel(word_id)
    flang = F.normalize(net.mapping_lang(raw_flang), p=2, dim=1)
    coords = [coord_map(h, w, device=raw_fvisu[0].device).flatten(1) for (h, w) in hw]
    inter, outbox = [], []
    for s in range(3):                                                            # :489-506
        N = corr[s].shape[2]
        x = torch.cat([corr[s], flang[:, :, None].expand(B, C, N), coords[s][None].expand(B, 8, N)], 1)
        seq = net.fcn_emb._modules[str(s)]
        y = _cbr(seq[0], x, training).reshape(B, -1, hw[s][0], hw[s][1])
        for m in list(seq)[1:]:
            y = m(y)
        inter.append(y)
        outbox.append(net.fcn_out._modules[str(s)](y).flatten(2))
    _, fa = net.sub_attn(context, embedded, word_id)                              # :525-528
    fa = F.normalize(fa, p=2, dim=1)
    sim = [pix2text(corr[s], fa)[0] for s in range(3)]                            # :530-535
    oo = [only_obj(outbox[s]) for s in range(3)]                                  # :545-552
    obj = [oo[s] * sim[s] for s in range(3)]
    locmap = location_branch(net, coords, obj, context, embedded, word_id)        # :556-610
    loc, st = [], 0
    for s in range(3):
        N = corr[s].shape[2]
        loc.append(locmap[:, st:st + N]); st += N
    outbox = [modulate_conf(outbox[s], sim[s], loc[s]) for s in range(3)]         # :612-621
    fm = net.feature_map[0]
    vit, lag, M = crossmodal_features(fv[0], context, fm.weight, fm.bias)         # :625-635
    q_cm, k_cm, neg_cm, word_cm, nidx_cm = crossmodal_sample(vit, lag, M, rng)    # :637
    shp = lambda t, s: t.reshape(t.shape[:-1] + hw[s])
    out = dict(
        outbox=[shp(outbox[s], s) for s in range(3)], sim_score=[shp(sim[s], s) for s in range(3)],
        loc_score=[shp(loc[s], s) for s in range(3)], corr_feat=[shp(corr[s], s) for s in range(3)],
        flang_attn=fa[:, :, None, None], frame_feature=q_if, corrspendence_feature=k_if, neg_feature=neg_if,
        vit_posit=q_cm, lag_posit=k_cm, neg_cross=neg_cm, only_obj=[shp(oo[s], s) for s in range(3)])
    if return_internals:
        out.update(fvisu=fv, flang=flang, context=context, interframe_idx=idx_if, interframe_negidx=nidx_if,
                   cross_word=word_cm, cross_negidx=nidx_cm, cross_map=M, vit=vit, lag=lag, intmd=inter)
    return out


def losses_restated(out, bbox, size):
    """train_DCNet.py:615-642 on the dict returned by forward_restated (train mode)."""
    dev = out['outbox'][0].device
    gt, gi, gj, best_n, gtc = build_target(bbox.detach().cpu(), size)
    gt, gtc = [t.to(dev) for t in gt], [t.to(dev) for t in gtc]
    pred = [o.reshape(o.shape[0], 3, 5, o.shape[2], o.shape[3]) for o in out['outbox']]
    fa = out['flang_attn'][:, :, 0, 0]
    neg_sim = [pix2text(c.flatten(2), fa)[1].reshape(s.shape) for c, s in zip(out['corr_feat'], out['sim_score'])]
    l_yolo = yolo_loss(pred, gt, gi, gj, best_n)
    l_rank = rank_loss(out['sim_score'], neg_sim, gtc)
    l_if = interframe_contrastive_loss(out['frame_feature'], out['corrspendence_feature'], out['neg_feature'])
    l_cm = crossmodal_contrastive_loss(out['vit_posit'], out['lag_posit'], out['neg_cross'])
    l_loc = loc_loss(out['loc_score'], gtc)
    total = l_yolo + 100 * l_rank + l_loc + 100 * l_if + l_cm                    # :642
    return dict(loss=total, yolo=l_yolo, rank=l_rank, interframe=l_if, cross=l_cm, loc=l_loc,
                gi=gi, gj=gj, best_n=best_n)


# ----------------------------------------------------------------------------------------------
# the hot path on its own: neighbours supplied as tensors
# ----------------------------------------------------------------------------------------------
def hotpath_restated(net, raw, flang, fa, context, head, loc, dy_head, bbox, size, rng=_pyrandom, backward=True, relu_masks=None,
                     fa_partner=None, bbox_partner=None):
    """Mirror of dcnet_b200.hotpath.HotPath.step for the CPU baseline and the parity tests: a2-a18 of
    model/DCNet_model.py:356-637 + train_DCNet.py:615-690 with Darknet / text encoder / head / location branch replaced by
    the given tensors (head[s] [B,15,N_s], loc[s] [B,N_s], dy_head[s] [B,512,N_s] = gradient entering the fusion output).
    relu_masks (test aid): dict(map=3x, corr=3x, fuse=3x bool [B,512,N_s]) pins the ReLU pattern of the three 1x1 layers.
    fa_partner [B,512] / bbox_partner [B,4]: text vector and box of each sample's rank-loss partner when it is not the local
    reversal B-1-b -- the slice a rank sees of the reference loss on the concatenated global batch (BASELINE config 5)."""
    training = net.training
    rm = relu_masks or {}
    mk = lambda k, s: rm[k][s] if k in rm else None
    B = raw[0].shape[0]
    P = B // 2
    hw = [(m.shape[2], m.shape[3]) for m in raw]
    fv = [l2norm_channels(_cbr(net.mapping_visu._modules[str(s)], raw[s].flatten(2), training, mk('map', s))) for s in range(3)]
    C = fv[0].shape[1]
    f1 = [f.reshape(P, 2, C, -1)[:, 0] for f in fv]
    f2 = [f.reshape(P, 2, C, -1)[:, 1] for f in fv]
    q_if, k_if, neg_if, idx_if, _ = interframe_sample(f1[0], f2[0], rng, topk_fn=canonical_topk)
    corr, sim, neg_sim, y = [], [], [], []
    for s in range(3):
        o1, o2 = coattention(f1[s], f2[s], net.temperature)
        x = interleave_pairs(torch.cat([f1[s], o1], 1), torch.cat([f2[s], o2], 1))
        c = l2norm_channels(_cbr(net.corr_conv._modules[str(s)][0], x, training, mk('corr', s)))
        corr.append(c)
        sm, ng = pix2text(c, fa, fa_partner)
        sim.append(sm); neg_sim.append(ng)
        N = c.shape[2]
        coord = coord_map(hw[s][0], hw[s][1], device=c.device).flatten(1)
        xin = torch.cat([c, flang[:, :, None].expand(B, C, N), coord[None].expand(B, 8, N)], 1)
        y.append(_cbr(net.fcn_emb._modules[str(s)][0], xin, training, mk('fuse', s)))
    pred = [modulate_conf(head[s], sim[s], loc[s]) for s in range(3)]
    fm = net.feature_map[0]
    vit, lag, M = crossmodal_features(fv[0], context, fm.weight, fm.bias)
    q_cm, k_cm, neg_cm, word, _ = crossmodal_sample(vit, lag, M, rng)
    dev = raw[0].device
    gt, gi, gj, best_n, gtc = build_target(bbox.detach().cpu(), size)
    gt, gtc = [t.to(dev) for t in gt], [t.to(dev) for t in gtc]
    g = [size // 32, size // 16, size // 8]
    pred5 = [p.reshape(B, 3, 5, g[s], g[s]) for s, p in enumerate(pred)]
    shp = lambda t, s: t.reshape(B, g[s], g[s])
    gtc_partner = None
    if bbox_partner is not None:
        gtc_partner = [t.to(dev) for t in build_target(bbox_partner.detach().cpu(), size)[4]]
    comp = dict(yolo=yolo_loss(pred5, gt, gi, gj, best_n),
                rank=rank_loss([shp(t, s) for s, t in enumerate(sim)], [shp(t, s) for s, t in enumerate(neg_sim)], gtc,
                               gt_center_partner=gtc_partner),
                loc=loc_loss([shp(t, s) for s, t in enumerate(loc)], gtc),
                interframe=interframe_contrastive_loss(q_if, k_if, neg_if),
                cross=crossmodal_contrastive_loss(q_cm, k_cm, neg_cm))
    loss = comp['yolo'] + 100 * comp['rank'] + comp['loc'] + 100 * comp['interframe'] + comp['cross']
    boxes = decode_at([p.detach() for p in pred5], gi, gj, best_n, size)
    iou = bbox_iou(boxes, bbox.detach().cpu().float())
    if backward:
        torch.autograd.backward([loss] + y, [None] + list(dy_head))
    return dict(loss=loss, comp=comp, y=y, iou=iou, boxes=boxes, best_n=best_n, gi=gi, gj=gj, corr=corr, sim=sim, pred=pred,
                idx_if=idx_if, word=word)


# ----------------------------------------------------------------------------------------------
# test-time multi-frame forward, restated                      model/test_DCNet_model.py:247-483
# ----------------------------------------------------------------------------------------------
def forward_test_restated(net, raw_fvisu, word_id, n_frame=5):
    """raw_fvisu 3 x [b*n_frame,C_s,h,w]; word_id [b,T].  Centre frame vs every other frame (:303-320), corr_conv + L2 norm per
    partner (:276-280), mean over partners (:324-332), then the same fusion / similarity / location / modulation as the
    training model with batch b.  Returns dict(outbox, sim_score, loc_score, corr_feat, only_obj, flang_attn)."""
    training = net.training
    BF = raw_fvisu[0].shape[0]
    b = BF // n_frame
    centre = n_frame // 2
    hw = [(m.shape[2], m.shape[3]) for m in raw_fvisu]
    fv = [l2norm_channels(_cbr(net.mapping_visu._modules[str(s)], raw_fvisu[s].flatten(2), training)) for s in range(3)]
    C = fv[0].shape[1]
    corr = []
    for s in range(3):
        f = fv[s].reshape(b, n_frame, C, -1)
        f1 = f[:, centre]
        outs = []
        for o in range(n_frame):
            if o == centre:
                continue
            o1, _ = coattention(f1, f[:, o], net.temperature)
            outs.append(l2norm_channels(_cbr(net.corr_conv._modules[str(s)][0], torch.cat([f1, o1], 1), training)))
        corr.append(torch.stack(outs, 0).mean(0))
    max_len = int((word_id != 0).sum(1).max().item())
    word_id = word_id[:, :max_len]
    raw_flang, context, embedded = net.textmodel(word_id)
    flang = F.normalize(net.mapping_lang(raw_flang), p=2, dim=1)
    coords = [coord_map(h, w, device=raw_fvisu[0].device).flatten(1) for (h, w) in hw]
    outbox = []
    for s in range(3):
        N = corr[s].shape[2]
        x = torch.cat([corr[s], flang[:, :, None].expand(b, C, N), coords[s][None].expand(b, 8, N)], 1)
        seq = net.fcn_emb._modules[str(s)]
        y = _cbr(seq[0], x, training).reshape(b, -1, hw[s][0], hw[s][1])
        for m in list(seq)[1:]:
            y = m(y)
        outbox.append(net.fcn_out._modules[str(s)](y).flatten(2))
    _, fa = net.sub_attn(context, embedded, word_id)
    fa = F.normalize(fa, p=2, dim=1)
    sim = [pix2text(corr[s], fa)[0] for s in range(3)]
    oo = [only_obj(outbox[s]) for s in range(3)]
    obj = [oo[s] * sim[s] for s in range(3)]
    locmap = location_branch(net, coords, obj, context, embedded, word_id)
    loc, st = [], 0
    for s in range(3):
        N = corr[s].shape[2]
        loc.append(locmap[:, st:st + N]); st += N
    outbox = [modulate_conf(outbox[s], sim[s], loc[s]) for s in range(3)]
    shp = lambda t, s: t.reshape(t.shape[:-1] + hw[s])
    return dict(outbox=[shp(outbox[s], s) for s in range(3)], sim_score=[shp(sim[s], s) for s in range(3)],
                loc_score=[shp(loc[s], s) for s in range(3)], corr_feat=[shp(corr[s], s) for s in range(3)],
                only_obj=[shp(oo[s], s) for s in range(3)], flang_attn=fa[:, :, None, None])


# ----------------------------------------------------------------------------------------------
# 8f-3  test-time cache writer and offline re-scoring      test_DCNet.py:546-701, post_processing.py:205-270
# ----------------------------------------------------------------------------------------------
def letterbox_image_size(ratio, dw, dh, size):
    """(img_w, img_h) of the un-letterboxed image the boxes are clamped to: the crop [top:bottom, left:right] of the size x size
    input resized by 1/ratio, with python's round() exactly as test_DCNet.py:617-624."""
    top, bottom = round(float(dh) - 0.1), size - round(float(dh) + 0.1)
    left, right = round(float(dw) - 0.1), size - round(float(dw) + 0.1)
    ratio = float(ratio)
    return round((right - left) / ratio), round((bottom - top) / ratio)


def topk_pred_boxes(pred5, fvisu, topk, ratio, dw, dh, size, anchor_imsize=416, anchors_full=ANCHORS_FULL):
    """One image (batch 1, as test_DCNet.py:save_cache runs).  pred5: 3 x [1,3,5,g,g]; fvisu: 3 x [1,C,g,g].
    Returns (boxes [k,1,4], scores list of k floats, cells list of (scale, anchor, gj, gi), feats [k,1,C]) following
    test_DCNet.py:593-645 and get_topk_pred_bbox :657-701: torch.topk over the concatenated confidences; the scale from the flat
    index (:662-667); the cell = first position of that scale equal to the value (np.where(...)[0], :682-683); box decode, un-letterbox
    and clamp (:688-696).  Top-k ties are canonicalised to the lower flat index (PyTorch leaves the order unspecified)."""
    conf = [p[:, :, 4].contiguous().view(1, -1) for p in pred5]                        # :595-597
    flat = torch.cat(conf, 1)[0]
    order = sorted(range(flat.numel()), key=lambda i: (-float(flat[i]), i))[:topk]     # topk(k) with ties -> lower index
    img_w, img_h = letterbox_image_size(ratio, dw, dh, size)
    g0 = size // 32
    boxes, scores, cells, feats = [], [], [], []
    for loc in order:
        v = flat[loc]
        s = 0 if loc < 3 * g0 ** 2 else (1 if loc < 3 * g0 ** 2 + 3 * (2 * g0) ** 2 else 2)     # :662-667
        grid, stride = size // (32 // (2 ** s)), 32 // (2 ** s)
        sa = scaled_anchors(s, size, anchor_imsize, anchors_full)
        pc = conf[s].view(3, grid, grid).numpy()
        a_, gj_, gi_ = np.where(pc == v.numpy())                                        # :682
        a, gi, gj = int(a_[0]), int(gi_[0]), int(gj_[0])
        p = pred5[s][0, a, :, gj, gi]
        b = torch.zeros(1, 4)
        b[0, 0] = torch.sigmoid(p[0]) + gi                                              # :688-692
        b[0, 1] = torch.sigmoid(p[1]) + gj
        b[0, 2] = torch.exp(p[2]) * sa[a][0]
        b[0, 3] = torch.exp(p[3]) * sa[a][1]
        b[0, :] = b[0, :] * stride
        b = xywh2xyxy(b)
        b[:, 0], b[:, 2] = (b[:, 0] - dw) / ratio, (b[:, 2] - dw) / ratio                # :695-696
        b[:, 1], b[:, 3] = (b[:, 1] - dh) / ratio, (b[:, 3] - dh) / ratio
        b[:, :2] = torch.clamp(b[:, :2], min=0)                                          # :697-698
        b[:, 2] = torch.clamp(b[:, 2], max=img_w)
        b[:, 3] = torch.clamp(b[:, 3], max=img_h)
        boxes.append(b); scores.append(float(v)); cells.append((s, a, gj, gi))
        feats.append(fvisu[s][:, :, gj, gi])                                             # :645  [1,C]
    return torch.stack(boxes), scores, cells, torch.stack(feats)


def post_rescore(centre_feat, ref_feats, ref_scores, invalid=()):
    """post_processing.py:239-274.  centre_feat [k,1,C]; ref_feats: R x [k,1,C]; ref_scores: R x [k]; invalid: frame indices whose cache
    was missing.  Returns (fused [k], best index, matched reference box [k,R])."""
    k = centre_feat.shape[0]
    R = len(ref_feats)
    refer = torch.cat(ref_feats, dim=1)                                                  # :239  [k,R,C]
    centre = centre_feat.unsqueeze(1)
    scores = torch.stack([torch.as_tensor(s, dtype=torch.float) for s in ref_scores]).permute(1, 0)     # :241  [k,R]
    C = refer.shape[2]
    refer = refer.view(-1, C).permute(1, 0)                                              # :245-246  [C, k*R]
    centre = centre.view(-1, C)
    sim = torch.bmm(centre.unsqueeze(0), refer.unsqueeze(0)).reshape(k, k, R)            # :249-251
    best_sim, idx = sim.max(dim=1)                                                       # :252
    ref_score = scores.gather(0, idx)                                                    # :254
    w = F.softmax(best_sim, dim=1)                                                       # :257
    if len(invalid) > 0:
        w[:, list(invalid)] = 0                                                          # :259-260
    fused = torch.sum(w * ref_score, dim=1)                                              # :262
    (where,) = np.where(fused.numpy() == fused.max().numpy())                            # :265
    return fused, int(where[0]), idx
