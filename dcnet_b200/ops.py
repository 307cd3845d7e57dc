"""torch.autograd wrappers over the C ABI (include/dcnet_b200.h).  PyTorch is plumbing here: it owns device memory,
streams and the autograd tape; every forward/backward body is a call into libdcnet_sm100.so.  CPU tensors are rejected:
there is no fallback path."""
import random

import numpy as np
import torch

from . import _lib

F32 = torch.float32


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _c(t, dtype=F32, name="tensor"):
    """checked contiguous CUDA tensor of the expected dtype"""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dcnet_b200: %s is on %s -- the hot path is CUDA-only (no CPU fallback)" % (name, t.device))
    if t.dtype != dtype:
        raise RuntimeError("dcnet_b200: %s has dtype %s, expected %s" % (name, t.dtype, dtype))
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------------------------
# host RNG (exact CPython random.sample stream)
# ------------------------------------------------------------------------------------------------------------------
def _rng_call(fn, out, *args):
    st = random.getstate()
    arr = np.array(st[1], dtype=np.uint32)
    rc = fn(arr.ctypes.data, *args, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(_lib.last_error())
    random.setstate((st[0], tuple(arr.tolist()), st[2]))
    return out


def pyrandom_interframe(P, top_k, N0, neg_n):
    """positions [P,top_k,neg_n] int32 drawn exactly like model/DCNet_model.py:411-413 would advance `random`."""
    out = np.empty((P, top_k, neg_n), dtype=np.int32)
    return _rng_call(_lib.lib().dcnet_pyrandom_interframe, out, P, top_k, N0, neg_n)


def pyrandom_crossmodal(B, N0, neg_n):
    """pixel indices [B,N0,neg_n] int64 of image B-1, consuming the stream like model/DCNet_model.py:81-96."""
    out = np.empty((B, N0, neg_n), dtype=np.int64)
    return _rng_call(_lib.lib().dcnet_pyrandom_crossmodal, out, B, N0, neg_n)


# ------------------------------------------------------------------------------------------------------------------
# plain (non-differentiable) entry points
# ------------------------------------------------------------------------------------------------------------------
def sgemm(A, B, C, M, N, K, batch=1, kbatch=1, sA=(0, 0, 0, 0), sB=(0, 0, 0, 0), sC=(0, 0, 0), idxA=None, idxB=None, idxC=None,
          alpha=1.0, beta=0.0, colscale=None, s_colscale=0, atomic=0):
    _lib.call("dcnet_sgemm", _p(A), _p(B), _p(C), M, N, K, batch, kbatch, *sA, *sB, *sC, _p(idxA), _p(idxB), _p(idxC),
              alpha, beta, _p(colscale), s_colscale, atomic, _st())
    return C


def gemm_tf32(A, B, a_mn, b_mn, M, N, K, alpha=1.0, out=None, atomic=0):
    """direct tensor-core GEMM: A [batch, M, K] (a_mn=0) or [batch, K, M] (a_mn=1); B [batch, N, K] or [batch, K, N].
    atomic=1: the result is ADDED to `out` (TMA reduce-add epilogue, reduction split over CTAs when there are few tiles)."""
    A, B = _c(A, name="A"), _c(B, name="B")
    batch = A.shape[0]
    C = torch.empty(batch, M, N, device=A.device, dtype=F32) if out is None else out
    _lib.call("dcnet_gemm_tf32", _p(A), int(a_mn), A.shape[2], A.shape[1] * A.shape[2], _p(B), int(b_mn), B.shape[2], B.shape[1] * B.shape[2],
              _p(C), N, M * N, M, N, K, batch, alpha, int(atomic), _st())
    return C


# Operands of the tf32 contractions are rounded to the nearest tf32 where they are produced (include/dcnet_b200.h, DCNET_RN_TF32): the
# MMA truncates, and the truncation bias adds up along the ~10 chained contractions of the backward.  False = the round-1 behaviour
# (diagnostics only).
RN_TF32 = True
_RN_FLAG = 0x100
# co-attention backward on fp16 operands (kind::f16 at twice the tf32 rate, the same 11 significant bits; include/dcnet_b200.h,
# dcnet_coattn_bwd `staged`).  False = the tf32 contractions (comparison).
BWD_FP16 = True
# the fused co-attention forward keeps its softmax weights (fp16) for the backward instead of the backward recomputing S (False: recompute)
KEEP_E = True


def round_tf32(x, out=None):
    """x rounded to the nearest tf32 value (fp32 storage); a tensor the library did not produce itself (weights, Darknet maps)"""
    x = _c(x.detach(), name="x")
    y = torch.empty_like(x) if out is None else out
    _lib.call("dcnet_round_tf32", _p(x), _p(y), x.numel(), _st())
    return y


def cast_bf16(x):
    """fp32 -> bf16 (round to nearest even) with the library's own kernel"""
    x = _c(x, name="x")
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.call("dcnet_cast_bf16", _p(x), _p(y), x.numel(), _st())
    return y


def gemm_bf16(A, B, a_mn, b_mn, M, N, K, alpha=1.0):
    """tcgen05 kind::f16 GEMM on bf16 operands (same operand conventions as gemm_tf32), fp32 result."""
    A, B = _c(A, torch.bfloat16, "A"), _c(B, torch.bfloat16, "B")
    batch = A.shape[0]
    C = torch.empty(batch, M, N, device=A.device, dtype=F32)
    _lib.call("dcnet_gemm_bf16", _p(A), int(a_mn), A.shape[2], A.shape[1] * A.shape[2], _p(B), int(b_mn), B.shape[2], B.shape[1] * B.shape[2],
              _p(C), N, M * N, M, N, K, batch, alpha, 0, _st())
    return C


def cast_f16(x):
    """fp32 -> fp16 (round to nearest even) with the library's own kernel"""
    x = _c(x, name="x")
    y = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _lib.call("dcnet_cast_f16", _p(x), _p(y), x.numel(), _st())
    return y


def gemm_f16(A, B, a_mn, b_mn, M, N, K, alpha=1.0, out=None, atomic=0):
    """tcgen05 kind::f16 GEMM on fp16 operands (same operand conventions as gemm_tf32), fp32 result; atomic=1 adds into `out`."""
    A, B = _c(A, torch.float16, "A"), _c(B, torch.float16, "B")
    batch = A.shape[0]
    C = torch.empty(batch, M, N, device=A.device, dtype=F32) if out is None else out
    _lib.call("dcnet_gemm_f16", _p(A), int(a_mn), A.shape[2], A.shape[1] * A.shape[2], _p(B), int(b_mn), B.shape[2], B.shape[1] * B.shape[2],
              _p(C), N, M * N, M, N, K, batch, alpha, int(atomic), _st())
    return C


def pix2text(x, fa):
    """x [B,C,N], fa [B,C] -> sim [B,N]   (forward only; inference / clip path)"""
    x, fa = _c(x.detach(), name="x"), _c(fa.detach(), name="fa")
    B, C, N = x.shape
    sim = torch.empty(B, N, device=x.device, dtype=F32)
    _lib.call("dcnet_pix2text", _p(x), _p(fa), None, _p(sim), None, B, C, N, _st())
    return sim


def loc_rank8(E, obj, W, bias, bn_scale, bn_shift, flang, return_raw=False):
    """location scores without the [B,SN,SN] relation tensor (forward only; model/DCNet_model.py:556-603).
    E [SN,8] normalised coordinate embeddings, obj [B,SN] normalised objectness, W [C,SN] / bias [C] the Linear,
    bn_scale / bn_shift [C] the eval-mode BatchNorm1d affine, flang [B,C] -> score [B,SN] (min-max normalised per image)."""
    E, obj, W, flang = _c(E.detach(), name="E"), _c(obj.detach(), name="obj"), _c(W.detach(), name="W"), _c(flang.detach(), name="flang")
    bn_scale, bn_shift = _c(bn_scale.detach(), name="bn_scale"), _c(bn_shift.detach(), name="bn_shift")
    bias = None if bias is None else _c(bias.detach(), name="bias")
    B, SN = obj.shape
    C = W.shape[0]
    if E.shape != (SN, 8) or W.shape[1] != SN or flang.shape != (B, C):
        raise ValueError(f"loc_rank8: shapes E {tuple(E.shape)}, obj {tuple(obj.shape)}, W {tuple(W.shape)}, flang {tuple(flang.shape)}")
    G = torch.empty(B, C, 8, device=obj.device, dtype=F32)
    raw = torch.empty(B, SN, device=obj.device, dtype=F32)
    score = torch.empty(B, SN, device=obj.device, dtype=F32)
    _lib.call("dcnet_loc_rank8_fwd", _p(E), _p(obj), _p(W), SN, _p(bias), _p(bn_scale), _p(bn_shift), _p(flang),
              _p(G), _p(raw), _p(score), B, SN, C, _st())
    return (score, raw, G) if return_raw else score


class _LocRank8Train(torch.autograd.Function):
    """SURVEY 8(f) row 2, training: the location branch (model/DCNet_model.py:556-603) in its rank-8 form on this library's kernels --
    Linear(SN -> C) on bmm(E, E^T) * obj, BatchNorm1d with BATCH statistics (running statistics updated like nn.BatchNorm1d), ReLU,
    channel norm, dot with the phrase vector, per-image min-max -- and its backward (dcnet_loc_rank8_train_fwd / _bwd).  Neither the
    [B,SN,SN] relation tensor nor the [B*SN,C] activations exist."""

    @staticmethod
    def forward(ctx, E, obj, W, bias, gamma, beta, running_mean, running_var, nbt, momentum, eps, flang):
        E, obj, W, flang = _c(E, name="E"), _c(obj, name="obj"), _c(W, name="W"), _c(flang, name="flang")
        gamma, beta = _c(gamma, name="gamma"), _c(beta, name="beta")
        bias = None if bias is None else _c(bias, name="bias")
        B, SN = obj.shape
        C = W.shape[0]
        if E.shape != (SN, 8) or W.shape[1] != SN or flang.shape != (B, C) or C % 32 != 0:
            raise ValueError(f"loc_rank8_train: shapes E {tuple(E.shape)}, obj {tuple(obj.shape)}, W {tuple(W.shape)}, flang {tuple(flang.shape)}")
        dev = obj.device
        G = torch.empty(B, C, 8, device=dev, dtype=F32)
        mom = torch.empty(72, device=dev, dtype=F32)
        stats = torch.empty(4 * C, device=dev, dtype=F32)
        raw, inrm, score = (torch.empty(B, SN, device=dev, dtype=F32) for _ in range(3))
        _lib.call("dcnet_loc_rank8_train_fwd", _p(E), _p(obj), _p(W), SN, _p(bias), _p(gamma), _p(beta), float(eps), float(momentum),
                  _p(running_mean), _p(running_var), _p(nbt), _p(flang), _p(G), _p(mom), _p(stats), _p(raw), _p(inrm), _p(score), B, SN, C, _st())
        ctx.save_for_backward(E, obj, W, bias, flang, G, stats, raw, inrm)
        return score

    @staticmethod
    def backward(ctx, dscore):
        E, obj, W, bias, flang, G, stats, raw, inrm = ctx.saved_tensors
        B, SN = obj.shape
        C = W.shape[0]
        dev = obj.device
        dscore = _c(dscore, name="dscore")
        draw = torch.empty(B, SN, device=dev, dtype=F32)
        dG = torch.empty(B, C, 8, device=dev, dtype=F32)
        dE, dobj, dW = torch.empty_like(E), torch.empty_like(obj), torch.empty_like(W)
        dgamma, dbeta = torch.empty(C, device=dev, dtype=F32), torch.empty(C, device=dev, dtype=F32)
        dflang = torch.empty_like(flang)
        _lib.call("dcnet_loc_rank8_train_bwd", _p(E), _p(obj), _p(W), SN, _p(bias), _p(flang), _p(G), _p(stats), _p(raw), _p(inrm), _p(dscore),
                  _p(draw), _p(dG), _p(dE), _p(dobj), _p(dW), _p(dgamma), _p(dbeta), _p(dflang), B, SN, C, _st())
        dbias = torch.zeros_like(bias) if (bias is not None and ctx.needs_input_grad[3]) else None     # removed by the batch statistics
        return dE, dobj, dW, dbias, dgamma, dbeta, None, None, None, None, None, dflang


def loc_rank8_train(E, obj, W, bias, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, flang):
    """training-mode location scores [B,SN] with gradients to E [SN,8], obj [B,SN], the Linear (W [C,SN], bias [C]), the BatchNorm1d
    affine (gamma, beta [C]) and flang [B,C]; running_mean / running_var / num_batches_tracked are updated in place."""
    return _LocRank8Train.apply(E, obj, W, bias, gamma, beta, running_mean, running_var, num_batches_tracked, float(momentum), float(eps), flang)


def coord_map(h, w, device):
    out = torch.empty(8, h, w, device=device, dtype=F32)
    _lib.call("dcnet_coord_map", _p(out), h, w, _st())
    return out


def interframe_topk(fv0, top_k=30):
    """fv0 [2P,C,N0] -> idx [P,top_k] int64 (flat row*N0+col, descending, ties -> lower index)."""
    fv0 = _c(fv0, name="fv0")
    B, C, N0 = fv0.shape
    P = B // 2
    S0 = torch.empty(P, N0, N0, device=fv0.device, dtype=F32)
    idx = torch.empty(P, top_k, device=fv0.device, dtype=torch.long)
    _lib.call("dcnet_interframe_topk", _p(fv0), P, C, N0, top_k, _p(S0), _p(idx), _st())
    return idx, S0


def interframe_negidx(idx, negpos, N0):
    P, top_k = idx.shape
    neg_n = negpos.shape[-1]
    out = torch.empty(P, top_k, neg_n, device=idx.device, dtype=torch.long)
    _lib.call("dcnet_interframe_negidx", _p(idx), _p(_c(negpos, torch.int32, "negpos")), P, N0, top_k, neg_n, _p(out), _st())
    return out


def interframe_cols(idx, negpos, N0):
    """-> cols [top_k*P*(2+neg_n)] int64, rank-major: gather columns of q [top_k,P] | k [top_k,P] | negatives [top_k,P,neg_n]"""
    P, top_k = idx.shape
    neg_n = negpos.shape[-1]
    out = torch.empty(P * top_k * (2 + neg_n), device=idx.device, dtype=torch.long)
    _lib.call("dcnet_interframe_cols", _p(idx), _p(_c(negpos, torch.int32, "negpos")), P, N0, top_k, neg_n, _p(out), _st())
    return out


def crossmodal_words(lag, vit, fm_w, fm_b):
    lag, vit = _c(lag, name="lag"), _c(vit, name="vit")
    B, T, C = lag.shape
    N0 = vit.shape[2]
    if tuple(fm_w.shape) != (T, T, 3) or tuple(fm_b.shape) != (T,):
        raise RuntimeError("dcnet_b200: feature_map weight %s / bias %s do not match the sentence length T=%d (the reference's "
                           "Conv1d(20,20,3) raises on any other length, model/DCNet_model.py:288)" % (tuple(fm_w.shape), tuple(fm_b.shape), T))
    M = torch.empty(B, T, N0, device=lag.device, dtype=F32)
    word = torch.empty(B, N0, device=lag.device, dtype=torch.long)
    _lib.call("dcnet_crossmodal_words", _p(lag), _p(vit), _p(_c(fm_w.detach(), name="fm_w")), _p(_c(fm_b.detach(), name="fm_b")),
              fm_w.shape[0], fm_w.shape[1], _p(M), _p(word), B, T, C, N0, _st())
    return word, M


def _anchors_arr(anchors):
    a = np.ascontiguousarray(np.array(anchors, dtype=np.float32).reshape(-1))
    return a


def build_target(bbox, size, anchor_imsize, anchors_full, dense=False):
    """-> (best_n, gi, gj [B] int64, t5 [B,5], gt list|None, gt_center list|None)   (train_DCNet.py:265-332)"""
    bbox = _c(bbox.float(), name="bbox")
    B = bbox.shape[0]
    dev = bbox.device
    best_n = torch.empty(B, device=dev, dtype=torch.long)
    gi = torch.empty_like(best_n)
    gj = torch.empty_like(best_n)
    t5 = torch.empty(B, 5, device=dev, dtype=F32)
    gt = gtc = None
    ptrs = [None] * 6
    if dense:
        gs = [size // 32, size // 16, size // 8]
        gt = [torch.empty(B, 3, 5, g, g, device=dev, dtype=F32) for g in gs]
        gtc = [torch.empty(B, 5, g, g, device=dev, dtype=F32) for g in gs]
        ptrs = [_p(t) for t in gt + gtc]
    an = _anchors_arr(anchors_full)
    assert an.size == 18
    _lib.call("dcnet_build_target", _p(bbox), B, int(size), float(anchor_imsize), an.ctypes.data, _p(best_n), _p(gi), _p(gj), _p(t5),
              *ptrs, _st())
    return best_n, gi, gj, t5, gt, gtc


def decode(pred, size, anchor_imsize, anchors_full, best_n=None, gi=None, gj=None, target=None):
    """pred 3 x [B,15,N_s] (or [B,3,5,g,g]).  With (best_n,gi,gj): decode at that cell (train_DCNet.py:656-672); without:
    arg-max decode (:766-816).  -> boxes [B,4] xyxy, iou [B]|None, best_n, gi, gj."""
    pred = [_c(p.detach(), name="pred") for p in pred]
    B = pred[0].shape[0]
    dev = pred[0].device
    mode = 0 if best_n is not None else 1
    if mode == 1:
        best_n = torch.empty(B, device=dev, dtype=torch.long)
        gi = torch.empty_like(best_n)
        gj = torch.empty_like(best_n)
    boxes = torch.empty(B, 4, device=dev, dtype=F32)
    iou = torch.empty(B, device=dev, dtype=F32) if target is not None else None
    tgt = _c(target.float(), name="target") if target is not None else None
    an = _anchors_arr(anchors_full)
    _lib.call("dcnet_decode", _p(pred[0]), _p(pred[1]), _p(pred[2]), B, int(size) // 32, int(size), float(anchor_imsize), an.ctypes.data,
              mode, _p(best_n), _p(gi), _p(gj), _p(boxes), _p(tgt), _p(iou), _st())
    return boxes, iou, best_n, gi, gj


def bbox_iou(b1, b2, x1y1x2y2=True):
    """utils/utils.py:76-104; like the reference the two box lists broadcast against each other ([1,4] vs [n,4])."""
    if b1.dim() != 2 or b2.dim() != 2 or b1.shape[1] != 4 or b2.shape[1] != 4:
        raise ValueError("bbox_iou: boxes must be [n,4], got %s and %s" % (tuple(b1.shape), tuple(b2.shape)))
    if b1.shape[0] != b2.shape[0]:
        if b1.shape[0] != 1 and b2.shape[0] != 1:
            raise ValueError("bbox_iou: %d boxes against %d (row counts must match or one side must be a single box)" % (b1.shape[0], b2.shape[0]))
        n = max(b1.shape[0], b2.shape[0])
        b1, b2 = b1.expand(n, 4), b2.expand(n, 4)
    b1, b2 = _c(b1.float(), name="box1"), _c(b2.float(), name="box2")
    out = torch.empty(b1.shape[0], device=b1.device, dtype=F32)
    _lib.call("dcnet_bbox_iou", _p(b1), _p(b2), b1.shape[0], int(bool(x1y1x2y2)), _p(out), _st())
    return out


def topk_boxes(pred, feat, k, meta, size, anchor_imsize, anchors_full):
    """8f-3 (test_DCNet.py:593-645, :657-701).  pred 3 x [B,15,N_s] (or [B,3,5,g,g]), feat 3 x [B,C,N_s] (corr_feat), meta [B,5] =
    (ratio, dw, dh, img_w, img_h).  -> boxes [B,k,4] (original image, clamped), scores [B,k], cells [B,k,4] int64 = (scale, anchor,
    gj, gi), feats [B,k,C]."""
    pred = [_c(p.detach().reshape(p.shape[0], 15, -1), name="pred") for p in pred]
    feat = [_c(f.detach().reshape(f.shape[0], f.shape[1], -1), name="feat") for f in feat]
    meta = _c(meta.detach().float(), name="meta")
    B, C = feat[0].shape[0], feat[0].shape[1]
    dev = pred[0].device
    boxes = torch.empty(B, k, 4, device=dev, dtype=F32)
    scores = torch.empty(B, k, device=dev, dtype=F32)
    cells = torch.empty(B, k, 4, device=dev, dtype=torch.long)
    feats = torch.empty(B, k, C, device=dev, dtype=F32)
    an = _anchors_arr(anchors_full)
    _lib.call("dcnet_topk_boxes", *[_p(p) for p in pred], *[_p(f) for f in feat], B, int(size) // 32, int(size), C, int(k), float(anchor_imsize),
              an.ctypes.data, _p(meta), _p(boxes), _p(scores), _p(cells), _p(feats), _st())
    return boxes, scores, cells, feats


def post_rescore(centre, ref, ref_score, invalid=None):
    """8f-3 (post_processing.py:239-274).  centre [k,C], ref [k,R,C], ref_score [k,R], invalid [R] int32 (or None)
    -> fused [k], best [1] int64, match [k,R] int64."""
    centre, ref, ref_score = _c(centre.detach(), name="centre"), _c(ref.detach(), name="ref"), _c(ref_score.detach().float(), name="ref_score")
    k, R, C = ref.shape
    invalid = _c(invalid, torch.int32, "invalid") if invalid is not None else None
    fused = torch.empty(k, device=ref.device, dtype=F32)
    best = torch.empty(1, device=ref.device, dtype=torch.long)
    match = torch.empty(k, R, device=ref.device, dtype=torch.long)
    _lib.call("dcnet_post_rescore", _p(centre), _p(ref), _p(ref_score), _p(invalid), k, R, C, _p(fused), _p(best), _p(match), _st())
    return fused, best, match


def yolo_layer_decode(x, anchors, num_classes, image_dim):
    x = _c(x, name="x")
    B, _, g, _ = x.shape
    A = len(anchors)
    out = torch.empty(B, A * g * g, 5 + num_classes, device=x.device, dtype=F32)
    an = _anchors_arr(anchors)
    _lib.call("dcnet_yolo_layer_decode", _p(x), _p(out), B, A, num_classes, g, float(image_dim), an.ctypes.data, _st())
    return out


# ------------------------------------------------------------------------------------------------------------------
# differentiable ops
# ------------------------------------------------------------------------------------------------------------------
def _pad_n(t, Np):
    """[.., N] -> [.., Np] zero-padded copy (row pitch a multiple of 16 bytes so TMA can address it)"""
    return None if t is None else torch.nn.functional.pad(t, (0, Np - t.shape[-1]))


class _ConvBNAct(torch.autograd.Function):
    """a1/a2/a6/a8: y = [l2norm_c] act(BN(W[:, :K1] x1 + W[:, K1:K1+K2] x2 + u 1^T + cc)), plus the fused a9 dots.

    Maps whose row pitch is not a multiple of 16 bytes (N % 4 != 0: the 13x13 scale of 416x416 inputs) cannot be addressed by
    TMA.  Their three contractions run on zero-padded copies of the operands (pitch rounded up to 4 positions) so they stay on
    tcgen05: the padded positions contribute zeros to every reduction over N, the padded output columns are dropped, and the
    BatchNorm statistics are taken from the unpadded z."""

    @staticmethod
    def forward(ctx, x1, x2, weight, gamma, beta, u, cc, fa, fa_neg, running_mean, running_var, training, momentum, eps, slope, l2norm, precision, nbt=None,
                flang=None, coords=None, round_in=True, round_out=False, stage_out=False):
        """stage_out: also return the co-attention staging of y (fp16 copy + column norms, dcnet_bn_act_fwd_staged) as a last,
        non-differentiable output -- the fused co-attention forward of the next block then starts without a pass over y.
        round_in: x1 / x2 are not tf32-rounded yet (a producer of this library that was told to round hands them over rounded:
        round_in=False); round_out: y feeds another tf32 contraction, round it on the way out (see RN_TF32)."""
        x1 = _c(x1, name="x1")
        x2 = _c(x2, name="x2")
        weight = _c(weight, name="weight")
        u, cc, fa, fa_neg = _c(u, name="u"), _c(cc, name="cc"), _c(fa, name="fa"), _c(fa_neg, name="fa_neg")
        gamma, beta = _c(gamma, name="gamma"), _c(beta, name="beta")
        B, K1, N = x1.shape
        K2 = 0 if x2 is None else x2.shape[1]
        C, ldw = weight.shape
        dev = x1.device
        st = _st()
        flang, coords = _c(flang, name="flang"), _c(coords, name="coords")
        if flang is not None:
            # a8: the text / coordinate columns of the weight act on flang [B,Ct] and coords [8,N] (split-weight form of the
            # reference's cat([corr_feat, flang tile, coord]) -> 1x1 conv): u and cc come from this library's own kernels
            if u is not None or cc is not None:
                raise ValueError("conv_bn_act: give either (u, cc) or (flang, coords)")
            Ct = flang.shape[1]
            if ldw != K1 + K2 + Ct + (8 if coords is not None else 0) or (coords is not None and tuple(coords.shape) != (8, N)):
                raise ValueError("conv_bn_act: weight [%d,%d] does not split into %d + %d visual, %d text%s columns" % (
                    C, ldw, K1, K2, Ct, ", 8 coordinate" if coords is not None else ""))
            u = torch.empty(B, C, device=dev, dtype=F32)
            cc = torch.empty(C, N, device=dev, dtype=F32) if coords is not None else None
            _lib.call("dcnet_fuse_terms_fwd", _p(weight), ldw, K1 + K2, Ct, K1 + K2 + Ct, _p(flang), _p(coords), _p(u), _p(cc), B, C, N, st)
        z = torch.empty(B, C, N, device=dev, dtype=F32)
        mean = torch.empty(C, device=dev, dtype=F32)
        invstd = torch.empty(C, device=dev, dtype=F32)
        ctx_precision = precision
        if precision == EXACT_FWD_TF32_BWD:
            precision = EXACT_FP32
        wq, rounded = weight, False
        if RN_TF32 and precision == TENSOR_TF32:
            wq = round_tf32(weight)
            if round_in:
                x1 = round_tf32(x1)
                x2 = round_tf32(x2) if x2 is not None else None
            rounded = True
        elif not round_in:
            rounded = True
        weight_full, weight = weight, wq          # the contractions read wq; the fp32 text / coordinate kernels the parameter itself
        padded = precision == 1 and N % 4 != 0 and K1 % 32 == 0 and K2 % 32 == 0
        if padded:
            Np = (N + 3) // 4 * 4
            zp = torch.empty(B, C, Np, device=dev, dtype=F32)
            x1p, x2p, ccp = _pad_n(x1, Np), _pad_n(x2, Np), _pad_n(cc, Np)      # named: they must outlive the launch below
            _lib.call("dcnet_conv1x1_fwd", _p(x1p), K1, _p(x2p), K2, _p(weight), ldw, _p(u), _p(ccp), _p(zp), B, C, Np, None, precision, st)
            z = zp[..., :N].contiguous()
            if training:
                _lib.call("dcnet_bn_stats", _p(z), B, C, N, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var), _p(nbt), st)
        elif training and precision == 1:
            # tensor-core path: BatchNorm sums come out of the GEMM epilogue, z is not re-read
            sums = torch.empty(2 * C, device=dev, dtype=F32)
            _lib.call("dcnet_conv1x1_fwd", _p(x1), K1, _p(x2), K2, _p(weight), ldw, _p(u), _p(cc), _p(z), B, C, N, _p(sums), precision, st)
            _lib.call("dcnet_bn_finalize", _p(sums), B * N, C, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var), _p(nbt), st)
        else:
            _lib.call("dcnet_conv1x1_fwd", _p(x1), K1, _p(x2), K2, _p(weight), ldw, _p(u), _p(cc), _p(z), B, C, N, None, precision, st)
        if training and precision != 1 and not padded:
            _lib.call("dcnet_bn_stats", _p(z), B, C, N, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var), _p(nbt), st)
        elif not training:
            _lib.call("dcnet_bn_eval_stats", _p(running_mean), _p(running_var), C, eps, _p(mean), _p(invstd), st)
        y = torch.empty_like(z)
        sim = neg = None
        if fa is not None:
            sim = torch.empty(B, N, device=dev, dtype=F32)
            neg = torch.empty(B, N, device=dev, dtype=F32)
        rn_out = _RN_FLAG if (RN_TF32 and round_out and ctx_precision == TENSOR_TF32) else 0
        staged, nst = None, 0
        if stage_out:
            nst = _lib.lib().dcnet_coattn_stage_bytes(B, C, N)
            staged = torch.empty(nst, device=dev, dtype=torch.uint8)
        _lib.call("dcnet_bn_act_fwd_staged", _p(z), _p(mean), _p(invstd), _p(gamma), _p(beta), slope, int(l2norm) | rn_out, _p(y), _p(fa), _p(fa_neg),
                  _p(sim), _p(neg), B, C, N, _p(staged), nst, st)
        ctx.save_for_backward(x1, x2, weight_full, gamma, beta, fa, fa_neg, z, mean, invstd, flang, coords, wq)
        ctx.cfg = (training, slope, int(l2norm), u is not None, cc is not None, ctx_precision, rounded)
        outs = (y,) if fa is None else (y, sim, neg)
        if stage_out:
            # the staging buffer is an output without a gradient; autograd must not materialise a zero tensor of its size for it
            # (88 MB at 416x416: a 35 us fill on the critical path, profiles/r3x_timeline_c3.txt)
            ctx.mark_non_differentiable(staged)
            ctx.set_materialize_grads(False)
            outs = outs + (staged,)
        return outs[0] if len(outs) == 1 else outs

    @staticmethod
    def backward(ctx, dy, *more):
        dsim, dneg = (more[0], more[1]) if len(more) >= 2 else (None, None)
        if dy is None and not any(g is not None for g in more):
            return (None,) * 23
        x1, x2, weight_full, gamma, beta, fa, fa_neg, z, mean, invstd, flang, coords, weight = ctx.saved_tensors
        training, slope, l2norm, has_u, has_cc, precision, rounded = ctx.cfg
        terms = flang is not None          # u / cc were derived from (flang, coords) inside forward
        if precision == EXACT_FWD_TF32_BWD:
            precision = TENSOR_TF32      # no index depends on the gradients: the backward contractions run on tcgen05
        rn = RN_TF32 and precision == TENSOR_TF32
        if rn and weight is weight_full:
            weight = round_tf32(weight_full)        # exact-fp32 forward: the backward's operands are rounded here
        if rn and not rounded and ctx.needs_input_grad[2]:
            x1 = round_tf32(x1)
            x2 = round_tf32(x2) if x2 is not None else None
        B, K1, N = x1.shape
        K2 = 0 if x2 is None else x2.shape[1]
        C, ldw = weight.shape
        dev = x1.device
        st = _st()
        if dy is None:                       # set_materialize_grads(False): an unused output arrives as None
            dy = torch.zeros_like(z)
        dy = _c(dy, name="dy")
        if fa is not None:
            dsim = torch.zeros(B, N, device=dev, dtype=F32) if dsim is None else _c(dsim, name="dsim")
            dneg = torch.zeros(B, N, device=dev, dtype=F32) if dneg is None else _c(dneg, name="dneg")
        else:
            dsim = dneg = None
        dv = torch.empty_like(z)
        # every accumulator the reduce kernel adds into with atomics, zeroed by ONE fill
        want_dfa = fa is not None and ctx.needs_input_grad[7]
        want_dfa_neg = fa_neg is not None and ctx.needs_input_grad[8]
        acc = torch.zeros(2 * C + (B * C if want_dfa else 0) + (B * C if want_dfa_neg else 0), device=dev, dtype=F32)
        sums = acc[:2 * C].view(2, C)
        dfa = acc[2 * C:2 * C + B * C].view(B, C) if want_dfa else None
        dfa_neg = acc[acc.numel() - B * C:].view(B, C) if want_dfa_neg else None
        _lib.call("dcnet_bn_act_bwd_reduce", _p(z), _p(mean), _p(invstd), _p(gamma), _p(beta), slope, l2norm, _p(dy), _p(fa), _p(fa_neg),
                  _p(dsim), _p(dneg), _p(dv), _p(sums[0]), _p(sums[1]), _p(dfa), _p(dfa_neg), B, C, N, st)
        _lib.call("dcnet_bn_act_bwd_apply", _p(z), _p(mean), _p(invstd), _p(gamma), _p(dv), _p(sums[0]), _p(sums[1]),
                  int(training) | (_RN_FLAG if rn else 0), _p(dv), B, C, N, st)
        dz = dv
        need_w = ctx.needs_input_grad[2]
        if precision == 1 and N % 4 != 0 and K1 % 128 == 0 and K2 % 128 == 0:
            # pitch-padded copies (see the class docstring); zero padding leaves every sum over N unchanged
            Np = (N + 3) // 4 * 4
            dzp = _pad_n(dz, Np)
            dx1 = dx2 = None
            if ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[1]):
                dx1p = torch.empty(B, K1, Np, device=dev, dtype=F32)
                dx2p = torch.empty(B, K2, Np, device=dev, dtype=F32) if x2 is not None else None
                _lib.call("dcnet_conv1x1_bwd_data", _p(dzp), _p(weight), ldw, _p(dx1p), K1, _p(dx2p), K2, B, C, Np, precision, st)
                dx1 = dx1p[..., :N].contiguous() if ctx.needs_input_grad[0] else None
                dx2 = dx2p[..., :N].contiguous() if (x2 is not None and ctx.needs_input_grad[1]) else None
            dW = du = dcc = None
            covered = (K1 + K2 == ldw) or (terms and (coords is not None or K1 + K2 + flang.shape[1] == ldw))
            if need_w:        # the kernels overwrite every column they own: a fill only when some columns are nobody's
                dW = (torch.empty if covered else torch.zeros)(C, ldw, device=dev, dtype=F32)
            if has_u and (terms or ctx.needs_input_grad[5]):
                du = torch.empty(B, C, device=dev, dtype=F32)
            dccp = torch.empty(C, Np, device=dev, dtype=F32) if (has_cc and (terms or ctx.needs_input_grad[6])) else None
            if need_w or du is not None or dccp is not None:
                x1p = _pad_n(x1, Np) if need_w else None
                x2p = _pad_n(x2, Np) if (need_w and x2 is not None) else None
                _lib.call("dcnet_conv1x1_bwd_weight", _p(dzp), _p(x1p), K1, _p(x2p), K2, _p(dW), ldw, _p(du), _p(dccp), B, C, Np, precision, st)
            dcc = dccp[:, :N].contiguous() if dccp is not None else None
            return _ConvBNAct._finish(ctx, dx1, dx2, dW, sums, du, dcc, dfa, dfa_neg, weight_full, flang, coords, K1 + K2, B, C, N, st)
        dx1 = torch.empty_like(x1) if ctx.needs_input_grad[0] else None
        dx2 = torch.empty_like(x2) if (x2 is not None and ctx.needs_input_grad[1]) else None
        if dx1 is not None or dx2 is not None:
            mx = None
            if getattr(ctx, "want_dx2_absmax", False) and dx1 is not None and dx2 is not None and precision == TENSOR_TF32:
                # corr_conv inside _Correspondence: max |dx2[b]| out of the GEMM's epilogue (the co-attention backward's fp16 scale)
                mx = torch.empty(B, device=dev, dtype=torch.int32)
                ctx.dx2_absmax = mx
            _lib.call("dcnet_conv1x1_bwd_data_absmax", _p(dz), _p(weight), ldw, _p(dx1), K1, _p(dx2), K2, B, C, N, precision, _p(mx), st)
        dW = du = dcc = None
        covered = (K1 + K2 == ldw) or (terms and (coords is not None or K1 + K2 + flang.shape[1] == ldw))
        if need_w:
            dW = (torch.empty if covered else torch.zeros)(C, ldw, device=dev, dtype=F32)
        if has_u and (terms or ctx.needs_input_grad[5]):
            du = torch.empty(B, C, device=dev, dtype=F32)
        if has_cc and (terms or ctx.needs_input_grad[6]):
            dcc = torch.empty(C, N, device=dev, dtype=F32)
        if need_w or du is not None or dcc is not None:
            _lib.call("dcnet_conv1x1_bwd_weight", _p(dz), _p(x1) if need_w else None, K1, _p(x2) if need_w else None, K2,
                      _p(dW), ldw, _p(du), _p(dcc), B, C, N, precision, st)
        return _ConvBNAct._finish(ctx, dx1, dx2, dW, sums, du, dcc, dfa, dfa_neg, weight_full, flang, coords, K1 + K2, B, C, N, st)

    @staticmethod
    def _finish(ctx, dx1, dx2, dW, sums, du, dcc, dfa, dfa_neg, weight, flang, coords, kv, B, C, N, st):
        """gradients in the order of forward's inputs; with (flang, coords) the text / coordinate columns of dW and dflang"""
        dflang = None
        if flang is not None:
            Ct = flang.shape[1]
            dflang = torch.empty_like(flang) if ctx.needs_input_grad[18] else None
            if dflang is not None or dW is not None:
                _lib.call("dcnet_fuse_terms_bwd", _p(weight), weight.shape[1], kv, Ct, kv + Ct, _p(flang), _p(coords), _p(du), _p(dcc),
                          _p(dflang), _p(dW), B, C, N, st)
            du = dcc = None
        return (dx1, dx2, dW, sums[1], sums[0], du, dcc, dfa, dfa_neg, None, None, None, None, None, None, None, None, None, dflang, None, None, None, None)


EXACT_FP32, TENSOR_TF32, TENSOR_F16_FUSED, EXACT_FWD_TF32_BWD = 0, 1, 2, 3
TENSOR_BF16_FUSED = TENSOR_F16_FUSED      # round-1 name (the fused kernel ran on bf16 operands then)
FUSED_MIN_N = 128      # the fused fp16 co-attention forward is used from this many positions on (below: exact fp32, tiny)


def conv_bn_act(x1, weight, gamma, beta, running_mean, running_var, training, x2=None, u=None, cc=None, fa=None,
                momentum=0.999, eps=1e-5, slope=0.0, l2norm=False, precision=TENSOR_TF32, fa_neg=None, num_batches_tracked=None,
                flang=None, coords=None, round_in=True, round_out=False, stage_out=False):
    """x1 [B,K1,N] (+x2 [B,K2,N]); weight [C,ldw].  Returns y [B,C,N] or (y, sim, neg_sim) when fa [B,C] is given.
    precision: TENSOR_TF32 = tcgen05 GEMMs (<=1e-3 relative), EXACT_FP32 = CUDA-core fp32 (<=1e-5).
    (u [B,C], cc [C,N]): extra terms added to the conv output; or (flang [B,Ct], coords [8,N]): the fusion's text / coordinate
    inputs, whose weight columns follow the visual ones in `weight` (a8) -- u and cc are then computed and back-propagated here.
    round_in / round_out: tf32 rounding of the operands / of y (RN_TF32): pass round_in=False for inputs a kernel of this library
    already rounded (round_out=True of the producing layer), round_out=True when y feeds another tf32 contraction.
    stage_out: a last output `staged` (uint8 buffer) = the fused co-attention's staging of y, for correspondence(staged=)."""
    return _ConvBNAct.apply(x1, x2, weight, gamma, beta, u, cc, fa, fa_neg, running_mean, running_var, bool(training), float(momentum),
                            float(eps), float(slope), bool(l2norm), int(precision), num_batches_tracked if training else None,
                            flang, coords, bool(round_in), bool(round_out), bool(stage_out))


class _Conv3x3BNAct(torch.autograd.Function):
    """ConvBatchNormReLU(C, C, 3, 1, 1) of the grounding head (model/DCNet_model.py:316-337) on this library's
    kernels: implicit-GEMM 3x3 convolution on tcgen05 (csrc/conv3x3.cu: three column-shifted copies of the map, the nine taps inside the
    K loop, rows off the image zero-filled by TMA), BatchNorm statistics from its epilogue, the BN / activation kernels of the 1x1 layers.
    Widths that are not a multiple of 4 (13, 26 at 416x416) run at a padded width (zero pad columns = the image border)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, training, momentum, eps, slope, nbt, h, w, round_out):
        x = _c(x, name="x")
        weight, gamma, beta = _c(weight, name="weight"), _c(gamma, name="gamma"), _c(beta, name="beta")
        B, Cin, N = x.shape
        Cout = weight.shape[0]
        if N != h * w or tuple(weight.shape) != (Cout, Cin, 3, 3):
            raise ValueError("conv3x3_bn_act: x %s, weight %s, h*w = %d" % (tuple(x.shape), tuple(weight.shape), h * w))
        wp = (w + 3) // 4 * 4
        if not _lib.lib().dcnet_conv3x3_supported(Cin, Cout, h, wp):
            raise RuntimeError("conv3x3_bn_act: shape (Cin=%d, Cout=%d, %dx%d) is not supported by the tcgen05 kernel" % (Cin, Cout, h, w))
        dev, st = x.device, _st()
        rn = _RN_FLAG if RN_TF32 else 0
        wq = torch.empty(9, Cout, Cin, device=dev, dtype=F32)
        _lib.call("dcnet_conv3x3_pack_weight", _p(weight), _p(wq), Cout, Cin, rn, st)
        Np = h * wp
        xm, xp, x0 = (torch.empty(B, Cin, Np, device=dev, dtype=F32) for _ in range(3))
        _lib.call("dcnet_conv3x3_shift_padded", _p(x), _p(xm), _p(xp), _p(x0), B * Cin * h, w, wp, rn, st)
        mean = torch.empty(Cout, device=dev, dtype=F32)
        invstd = torch.empty(Cout, device=dev, dtype=F32)
        if wp == w:
            z = torch.empty(B, Cout, N, device=dev, dtype=F32)
            sums = torch.empty(2 * Cout, device=dev, dtype=F32) if training else None
            _lib.call("dcnet_conv3x3_fwd", _p(xm), _p(x0), _p(xp), _p(wq), _p(z), B, Cin, Cout, h, w, _p(sums), st)
            if training:
                _lib.call("dcnet_bn_finalize", _p(sums), B * N, Cout, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var), _p(nbt), st)
        else:
            zp = torch.empty(B, Cout, Np, device=dev, dtype=F32)
            _lib.call("dcnet_conv3x3_fwd", _p(xm), _p(x0), _p(xp), _p(wq), _p(zp), B, Cin, Cout, h, wp, None, st)
            z = torch.empty(B, Cout, N, device=dev, dtype=F32)
            _lib.call("dcnet_conv3x3_unpad", _p(zp), _p(z), B * Cout * h, w, wp, st)
            if training:
                _lib.call("dcnet_bn_stats", _p(z), B, Cout, N, eps, momentum, _p(mean), _p(invstd), _p(running_mean), _p(running_var), _p(nbt), st)
        if not training:
            _lib.call("dcnet_bn_eval_stats", _p(running_mean), _p(running_var), Cout, eps, _p(mean), _p(invstd), st)
        y = torch.empty_like(z)
        _lib.call("dcnet_bn_act_fwd", _p(z), _p(mean), _p(invstd), _p(gamma), _p(beta), slope, (_RN_FLAG if (RN_TF32 and round_out) else 0), _p(y),
                  None, None, None, None, B, Cout, N, st)
        ctx.save_for_backward(xm, x0, xp, wq, gamma, beta, z, mean, invstd)
        ctx.cfg = (training, slope, h, w, wp)
        return y

    @staticmethod
    def backward(ctx, dy):
        xm, x0, xp, wq, gamma, beta, z, mean, invstd = ctx.saved_tensors
        training, slope, h, w, wp = ctx.cfg
        B, Cout, N = z.shape
        Cin = wq.shape[2]
        Np = h * wp
        dev, st = z.device, _st()
        dy = _c(dy, name="dy")
        rn = _RN_FLAG if RN_TF32 else 0
        dv = torch.empty_like(z)
        sums = torch.zeros(2, Cout, device=dev, dtype=F32)
        _lib.call("dcnet_bn_act_bwd_reduce", _p(z), _p(mean), _p(invstd), _p(gamma), _p(beta), slope, 0, _p(dy), None, None, None, None, _p(dv),
                  _p(sums[0]), _p(sums[1]), None, None, B, Cout, N, st)
        _lib.call("dcnet_bn_act_bwd_apply", _p(z), _p(mean), _p(invstd), _p(gamma), _p(dv), _p(sums[0]), _p(sums[1]), int(training) | rn, _p(dv),
                  B, Cout, N, st)
        dz = dv
        dx = dW = None
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dz0 = dz
        if need_x or wp != w:
            dzm, dzp = (torch.empty(B, Cout, Np, device=dev, dtype=F32) for _ in range(2))
            dz0 = dz if wp == w else torch.empty(B, Cout, Np, device=dev, dtype=F32)
            _lib.call("dcnet_conv3x3_shift_padded", _p(dz), _p(dzm), _p(dzp), _p(dz0) if wp != w else None, B * Cout * h, w, wp, 0, st)   # dz is rounded already
        if need_x:
            if wp == w:
                dx = torch.empty(B, Cin, N, device=dev, dtype=F32)
                _lib.call("dcnet_conv3x3_bwd_data", _p(dzm), _p(dz0), _p(dzp), _p(wq), _p(dx), B, Cin, Cout, h, w, st)
            else:
                dxp = torch.empty(B, Cin, Np, device=dev, dtype=F32)
                _lib.call("dcnet_conv3x3_bwd_data", _p(dzm), _p(dz0), _p(dzp), _p(wq), _p(dxp), B, Cin, Cout, h, wp, st)
                dx = torch.empty(B, Cin, N, device=dev, dtype=F32)
                _lib.call("dcnet_conv3x3_unpad", _p(dxp), _p(dx), B * Cin * h, w, wp, st)
        if need_w:
            dWp = torch.empty(Cout, 9, Cin, device=dev, dtype=F32)
            dW = torch.empty(Cout, Cin, 3, 3, device=dev, dtype=F32)
            _lib.call("dcnet_conv3x3_bwd_weight", _p(dz0), _p(xm), _p(x0), _p(xp), _p(dWp), _p(dW), B, Cin, Cout, h, wp, st)
        return dx, dW, sums[1], sums[0], None, None, None, None, None, None, None, None, None, None


def conv3x3_supported(Cin, Cout, h, w):
    """shapes ops.conv3x3_bn_act runs (any width: widths that are not a multiple of 4 run at the next multiple)"""
    return bool(_lib.lib().dcnet_conv3x3_supported(int(Cin), int(Cout), int(h), (int(w) + 3) // 4 * 4))


def conv3x3_bn_act(x, weight, gamma, beta, running_mean, running_var, training, h, w, momentum=0.999, eps=1e-5, slope=0.0,
                   num_batches_tracked=None, round_out=False):
    """x [B,Cin,h*w], weight [Cout,Cin,3,3] (stride 1, padding 1, no bias) -> act(BN(conv3x3(x))) [B,Cout,h*w]; tcgen05 tf32."""
    return _Conv3x3BNAct.apply(x, weight, gamma, beta, running_mean, running_var, bool(training), float(momentum), float(eps), float(slope),
                               num_batches_tracked if training else None, int(h), int(w), bool(round_out))


class _Conv1x1Bias(torch.autograd.Function):
    """fcn_out[s][-1] = nn.Conv2d(256, 15, 1) with bias (model/DCNet_model.py:329-337): 15 output channels -- a memory-bound pass over
    the 256-channel map, exact fp32 on the CUDA cores (dcnet_conv1x1_* at precision 0); the bias enters as the per-image term u."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight, bias = _c(x, name="x"), _c(weight, name="weight"), _c(bias, name="bias")
        B, K, N = x.shape
        C = weight.shape[0]
        u = bias[None, :].expand(B, C).contiguous()
        z = torch.empty(B, C, N, device=x.device, dtype=F32)
        _lib.call("dcnet_conv1x1_fwd", _p(x), K, None, 0, _p(weight), K, _p(u), None, _p(z), B, C, N, None, EXACT_FP32, _st())
        ctx.save_for_backward(x, weight)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, weight = ctx.saved_tensors
        B, K, N = x.shape
        C = weight.shape[0]
        dz = _c(dz, name="dz")
        st = _st()
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _lib.call("dcnet_conv1x1_bwd_data", _p(dz), _p(weight), K, _p(dx), K, None, 0, B, C, N, EXACT_FP32, st)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dW = torch.empty(C, K, device=x.device, dtype=F32)
            du = torch.empty(B, C, device=x.device, dtype=F32)
            _lib.call("dcnet_conv1x1_bwd_weight", _p(dz), _p(x), K, None, 0, _p(dW), K, _p(du), None, B, C, N, EXACT_FP32, st)
            db = du.sum(0)
        return dx, dW, db


def conv1x1_bias(x, weight, bias):
    """x [B,K,N], weight [C,K], bias [C] -> [B,C,N]"""
    return _Conv1x1Bias.apply(x, weight, bias)


class _FuseTerms(torch.autograd.Function):
    """a8 as a node of its own: (u, cc) = (W_l flang, W_c coords) and their backward through dcnet_fuse_terms_*.  conv_bn_act does the
    same inline when it is given (flang, coords); this form lets a caller issue the terms early / on another stream (HotPath)."""

    @staticmethod
    def forward(ctx, weight, flang, coords, kv):
        weight, flang, coords = _c(weight, name="weight"), _c(flang, name="flang"), _c(coords, name="coords")
        C, ldw = weight.shape
        B, Ct = flang.shape
        N = coords.shape[1] if coords is not None else 1
        u = torch.empty(B, C, device=flang.device, dtype=F32)
        cc = torch.empty(C, N, device=flang.device, dtype=F32) if coords is not None else None
        _lib.call("dcnet_fuse_terms_fwd", _p(weight), ldw, kv, Ct, kv + Ct, _p(flang), _p(coords), _p(u), _p(cc), B, C, N, _st())
        ctx.save_for_backward(weight, flang, coords)
        ctx.kv = kv
        return (u, cc) if cc is not None else u

    @staticmethod
    def backward(ctx, du, dcc=None):
        weight, flang, coords = ctx.saved_tensors
        C, ldw = weight.shape
        B, Ct = flang.shape
        N = coords.shape[1] if coords is not None else 1
        du = _c(du, name="du")
        dcc = _c(dcc, name="dcc") if coords is not None else None
        dflang = torch.empty_like(flang) if ctx.needs_input_grad[1] else None
        dW = None
        if ctx.needs_input_grad[0]:
            # only the text / coordinate columns are this node's; the visual columns arrive from the conv node and autograd adds
            dW = torch.zeros(C, ldw, device=flang.device, dtype=F32)
        _lib.call("dcnet_fuse_terms_bwd", _p(weight), ldw, ctx.kv, Ct, ctx.kv + Ct, _p(flang), _p(coords), _p(du), _p(dcc), _p(dflang), _p(dW),
                  B, C, N, _st())
        return dW, dflang, None, None


def fuse_terms(weight, flang, coords, kv):
    """weight [C,ldw] = [W_v (kv columns) | W_l | W_c]; -> (u [B,C], cc [C,N]) (cc omitted when coords is None)"""
    return _FuseTerms.apply(weight, flang, coords, int(kv))


class _CoAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, frames, qa, kb, oidx, n_out, tau, precision, round_out=False, prestaged=None):
        """round_out: the attention maps feed a tf32 contraction (corr_conv): they leave rounded to the nearest tf32 (RN_TF32);
        prestaged: the staging of `frames` its producer already wrote (conv_bn_act(stage_out=True)) -- precision 2 only"""
        frames = _c(frames, name="frames")
        qa, kb, oidx = (_c(t, torch.int32, "index") for t in (qa, kb, oidx))
        F_, C, N = frames.shape
        nprob = qa.numel()
        out = torch.empty(n_out, C, N, device=frames.device, dtype=F32) if n_out == nprob else \
            torch.zeros(n_out, C, N, device=frames.device, dtype=F32)
        lse = torch.empty(nprob, N, device=frames.device, dtype=F32)
        staged = ekeep = rkeep = None
        ctx.precision = precision
        rn_out = _RN_FLAG if (RN_TF32 and round_out) else 0
        if precision == EXACT_FWD_TF32_BWD:
            # exact fp32 forward (lse included); the backward is the tcgen05 one with fused epilogues, which recomputes its own tf32
            # logits and re-normalises them (the saved lse is only a shift there, so the two precisions cannot disagree about P)
            precision = EXACT_FP32
            ctx.precision = TENSOR_F16_FUSED
        if precision == TENSOR_F16_FUSED and N < FUSED_MIN_N:
            # short key axes: with few keys the rounding of reduced-precision operands does not average out -- bf16 / tf32 gave 1.6e-3 / 1.15e-3 on the
            # worst of 112 problems at N = 64, bar 1e-3 -- and the contraction is tiny: exact fp32 forward; the backward is the
            # precision-2 one either way (it recomputes its own tf32 logits and uses the saved lse only as a shift)
            precision = EXACT_FP32
        if precision == TENSOR_F16_FUSED and C % 128 == 0 and C <= 512:
            # fused kernel: only the fp16 staging of the maps is needed; the backward's contractions read the same staging
            nbytes = _lib.lib().dcnet_coattn_stage_bytes(F_, C, N)
            if prestaged is not None:
                if prestaged.numel() < nbytes or prestaged.dtype != torch.uint8:
                    raise ValueError("coattention: prestaged buffer does not belong to frames of shape %s" % (tuple(frames.shape),))
                staged = prestaged
            else:
                staged = torch.empty(nbytes, device=frames.device, dtype=torch.uint8)
                _lib.call("dcnet_coattn_stage", _p(frames), F_, C, N, _p(staged), nbytes, _st())
            if BWD_FP16 and KEEP_E and ctx.needs_input_grad[0] and N % 4 == 0:
                # training: the kernel keeps its unnormalised weights E (fp16) and their row sums -- the backward then starts without
                # recomputing S = Fa^T Fb (HBM is not scarce: 468 MB at 416x416 for 32 problems)
                ekeep = torch.empty(_lib.lib().dcnet_coattn_keep_bytes(nprob, N), device=frames.device, dtype=torch.uint8)
                rkeep = torch.empty(nprob, N, device=frames.device, dtype=F32)
            _lib.call("dcnet_coattn_fused_fwd_keep", _p(staged), F_, _p(qa), _p(kb), _p(oidx), nprob, _p(out), n_out, _p(lse), C, N, tau, rn_out,
                      _p(ekeep), _p(rkeep), _st())
        else:
            nbytes = _lib.lib().dcnet_coattn_workspace_bytes(F_, nprob, C, N, precision)
            ws = torch.empty(nbytes, device=frames.device, dtype=torch.uint8)
            _lib.call("dcnet_coattn_fwd", _p(frames), F_, _p(qa), _p(kb), _p(oidx), nprob, _p(out), n_out, _p(lse), C, N, tau, precision | rn_out,
                      _p(ws), nbytes, _st())
        # the fp16 staging goes to the backward: its five contractions then run on fp16 operands (tf32's precision at twice the rate)
        ctx.save_for_backward(frames, qa, kb, oidx, out, lse, staged if BWD_FP16 else None, ekeep, rkeep)
        ctx.tau = tau
        return out

    @staticmethod
    def backward(ctx, dout, accumulate_into=None, dout_absmax=None):
        """accumulate_into (used by _Correspondence): a [F,C,N] gradient buffer that already holds the other consumers' contribution to
        d frames -- the kernels add into it (TMA reduce-add) instead of into a zero-filled tensor that autograd would add afterwards"""
        frames, qa, kb, oidx, out, lse, staged, ekeep, rkeep = ctx.saved_tensors
        F_, C, N = frames.shape
        nprob = qa.numel()
        dout = _c(dout, name="dout")
        dframes = torch.zeros_like(frames) if accumulate_into is None else accumulate_into
        nbytes = _lib.lib().dcnet_coattn_workspace_bytes(F_, nprob, C, N, ctx.precision)
        ws = torch.empty(nbytes, device=frames.device, dtype=torch.uint8)
        _lib.call("dcnet_coattn_bwd_ex", _p(frames), F_, _p(qa), _p(kb), _p(oidx), nprob, _p(out), out.shape[0], _p(lse), _p(dout),
                  _p(dout_absmax) if staged is not None else None, _p(dframes), C, N, ctx.tau, ctx.precision, _p(staged), _p(ekeep), _p(rkeep),
                  _p(ws), nbytes, _st())
        return dframes, None, None, None, None, None, None, None, None


def coattn_stage(frames):
    """fp16 staging + column norms of frames [F,C,N] for coattn_fused (forward only)."""
    frames = _c(frames.detach(), name="frames")
    F_, C, N = frames.shape
    nbytes = _lib.lib().dcnet_coattn_stage_bytes(F_, C, N)
    staged = torch.empty(nbytes, device=frames.device, dtype=torch.uint8)
    _lib.call("dcnet_coattn_stage", _p(frames), F_, C, N, _p(staged), nbytes, _st())
    return staged


def coattn_fused(staged, shape, qa, kb, oidx=None, n_out=None, tau=10.0, out=None, lse=None):
    """the fused tcgen05 kernel alone over a staged buffer (forward only) -> out [n_out,C,N], lse [nprob,N]"""
    F_, C, N = shape
    qa, kb = _c(qa, torch.int32, "index"), _c(kb, torch.int32, "index")
    nprob = qa.numel()
    oidx = _iota(nprob, qa.device) if oidx is None else _c(oidx, torch.int32, "index")
    n_out = nprob if n_out is None else n_out
    if out is None:
        out = torch.empty(n_out, C, N, device=staged.device, dtype=F32)
    if lse is None:
        lse = torch.empty(nprob, N, device=staged.device, dtype=F32)
    _lib.call("dcnet_coattn_fused_fwd", _p(staged), F_, _p(qa), _p(kb), _p(oidx), nprob, _p(out), n_out, _p(lse), C, N, float(tau), 0, _st())
    return out, lse


_IOTA = {}


def _iota(n, device):
    """cached arange(n) int32 (identity output map of the co-attention problems): no kernel launch per call"""
    key = (int(n), str(device))
    if key not in _IOTA:
        _IOTA[key] = torch.arange(n, device=device, dtype=torch.int32)
    return _IOTA[key]


class _FakeCtx:
    """stands in for an autograd ctx when one Function's forward / backward body is run inside another Function"""

    def __init__(self, needs_input_grad=()):
        self.saved_tensors = ()
        self.needs_input_grad = needs_input_grad

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


class _Correspondence(torch.autograd.Function):
    """a5 + a6 (+ a9) as ONE autograd node: attn = coattention(fv), y = corr_conv([fv | attn]) (+ pixel-to-text dots).  fv feeds both
    the co-attention and the conv; as two nodes autograd materialises two [B,C,N] gradients of fv (one of them into a zero-filled tensor)
    and adds them -- at 416x416 that is a 177 MB fill and a 531 MB add on the critical path of the finest scale.  Here the conv's data
    gradient is written first and the co-attention backward reduce-adds into the same buffer."""

    @staticmethod
    def forward(ctx, fv, qa, kb, tau, cprecision, weight, gamma, beta, fa, fa_neg, running_mean, running_var, training, momentum, eps, slope,
                precision, nbt, round_in, round_out, staged):
        c1, c2 = _FakeCtx((ctx.needs_input_grad[0],)), _FakeCtx()
        nprob = qa.numel()
        rn = RN_TF32 and precision == TENSOR_TF32
        if rn and round_in:
            # fv arrives unrounded (the scale-0 maps are exact fp32 for the index selections): every contraction below -- corr_conv,
            # and the five of the co-attention backward -- reads one rounded copy; the gradient is that of the identity
            fv = round_tf32(_c(fv, name="fv"))
            staged = None
        attn = _CoAttn.forward(c1, fv, qa, kb, _iota(nprob, qa.device), nprob, tau, cprecision, rn, staged)
        out = _ConvBNAct.forward(c2, fv, attn, weight, gamma, beta, None, None, fa, fa_neg, running_mean, running_var, training, momentum, eps,
                                 slope, True, precision, nbt, None, None, False, round_out, False)
        ctx.n1 = len(c1.saved_tensors)
        ctx.save_for_backward(*c1.saved_tensors, *c2.saved_tensors)
        ctx.c1_attrs = (c1.tau, c1.precision)
        ctx.c2_cfg = c2.cfg
        return out

    @staticmethod
    def backward(ctx, dy, dsim=None, dneg=None):
        saved = ctx.saved_tensors
        nig = ctx.needs_input_grad
        # conv node first: inputs (x1 = fv, x2 = attn, weight, gamma, beta, u, cc, fa, fa_neg, ...)
        c2 = _FakeCtx((True, True, nig[5], nig[6], nig[7], False, False, nig[8], nig[9]) + (False,) * 14)
        c2.saved_tensors = saved[ctx.n1:]
        c2.cfg = ctx.c2_cfg
        c2.want_dx2_absmax = BWD_FP16
        g2 = _ConvBNAct.backward(c2, dy, dsim, dneg)
        dfv, dattn, dW, dgamma, dbeta, dfa, dfa_neg = g2[0], g2[1], g2[2], g2[3], g2[4], g2[7], g2[8]
        c1 = _FakeCtx()
        c1.saved_tensors = saved[:ctx.n1]
        c1.tau, c1.precision = ctx.c1_attrs
        _CoAttn.backward(c1, dattn, accumulate_into=dfv, dout_absmax=getattr(c2, "dx2_absmax", None))
        return (dfv, None, None, None, None, dW, dgamma, dbeta, dfa, dfa_neg, None, None, None, None, None, None, None, None, None, None, None)


def correspondence(fv, qa, kb, weight, gamma, beta, running_mean, running_var, training, fa=None, fa_neg=None, tau=10.0, cprecision=TENSOR_F16_FUSED,
                   momentum=0.999, eps=1e-5, slope=0.0, precision=TENSOR_TF32, num_batches_tracked=None, round_in=True, round_out=False,
                   staged=None):
    """fv [B,C,N] (pairs = consecutive frames via qa / kb) -> corr_feat [B,Cout,N] (channel-normalised) or (corr_feat, sim, neg_sim) with fa.
    weight [Cout, 2C]: corr_conv applied to [fv | co-attention(fv)] (model/DCNet_model.py:449-469, :525-535).
    staged: the co-attention staging of fv written by its producer (conv_bn_act(stage_out=True)); None = staged here."""
    return _Correspondence.apply(fv, qa, kb, float(tau), int(cprecision), weight, gamma, beta, fa, fa_neg, running_mean, running_var, bool(training),
                                 float(momentum), float(eps), float(slope), int(precision), num_batches_tracked if training else None,
                                 bool(round_in), bool(round_out), staged)


def coattention(frames, qa, kb, oidx=None, n_out=None, tau=10.0, precision=1):
    """frames [F,C,N]; problem i: queries frame qa[i] attend to frame kb[i]; result row oidx[i] of out [n_out,C,N]."""
    if oidx is None:
        oidx = _iota(qa.numel(), qa.device)
    if n_out is None:
        n_out = qa.numel()
    return _CoAttn.apply(frames, qa, kb, oidx, int(n_out), float(tau), int(precision))


class _GatherCols(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, img, col):
        src = _c(src, name="src")
        img = _c(img, torch.int32, "img")
        col = _c(col, torch.long, "col")
        F_, C, N = src.shape
        n = img.numel()
        out = torch.empty(n, C, device=src.device, dtype=F32)
        _lib.call("dcnet_gather_cols", _p(src), _p(img), _p(col), n, _p(out), C, N, _st())
        ctx.save_for_backward(img, col)
        ctx.shape = (F_, C, N)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, col = ctx.saved_tensors
        F_, C, N = ctx.shape
        dout = _c(dout, name="dout")
        dsrc = torch.zeros(F_, C, N, device=dout.device, dtype=F32)
        _lib.call("dcnet_scatter_cols_add", _p(dout), _p(img), _p(col), img.numel(), _p(dsrc), C, N, _st())
        return dsrc, None, None


def gather_cols(src, img, col):
    """out[i,:] = src[img[i], :, col[i]]"""
    return _GatherCols.apply(src, img, col)


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, neg, T):
        q, k, neg = _c(q, name="q"), _c(k, name="k"), _c(neg, name="neg")
        G, C = q.shape
        n = neg.shape[1]
        out = torch.empty(G, device=q.device, dtype=F32)
        _lib.call("dcnet_infonce_fwd", _p(q), _p(k), _p(neg), G, n, C, T, _p(out), _st())
        ctx.save_for_backward(q, k, neg)
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, g):
        q, k, neg = ctx.saved_tensors
        G, C = q.shape
        n = neg.shape[1]
        g = _c(g, name="grad")
        dq, dk, dneg = torch.empty_like(q), torch.empty_like(k), torch.empty_like(neg)
        _lib.call("dcnet_infonce_bwd", _p(q), _p(k), _p(neg), G, n, C, ctx.T, _p(g), 1, _p(dq), _p(dk), _p(dneg), _st())
        return dq, dk, dneg, None


def infonce_rows(q, k, neg, T=0.07):
    """q,k [G,C]; neg [G,n,C] -> per-group CE loss [G]"""
    return _InfoNCE.apply(q, k, neg, float(T))


class _RowNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x, name="x")
        L = x.shape[-1]
        R = x.numel() // L
        y = torch.empty_like(x)
        nrm = torch.empty(R, device=x.device, dtype=F32)
        _lib.call("dcnet_rownorm_fwd", _p(x), _p(y), _p(nrm), R, L, _st())
        ctx.save_for_backward(y, nrm)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved_tensors
        dy = _c(dy, name="dy")
        L = y.shape[-1]
        dx = torch.empty_like(y)
        _lib.call("dcnet_rownorm_bwd", _p(y), _p(nrm), _p(dy), _p(dx), y.numel() // L, L, _st())
        return dx


def rownorm(x):
    """F.normalize over the innermost axis (model/DCNet_model.py:629 on the flattened spatial axis)"""
    return _RowNorm.apply(x)


class _LagNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, context):
        context = _c(context, name="context")
        B, T, C2 = context.shape
        C = C2 // 2
        lag = torch.empty(B, T, C, device=context.device, dtype=F32)
        nrm = torch.empty(B, C, device=context.device, dtype=F32)
        _lib.call("dcnet_lagnorm_fwd", _p(context), _p(lag), _p(nrm), B, T, C, _st())
        ctx.save_for_backward(lag, nrm)
        return lag

    @staticmethod
    def backward(ctx, dlag):
        lag, nrm = ctx.saved_tensors
        B, T, C = lag.shape
        dlag = _c(dlag, name="dlag")
        dctx = torch.empty(B, T, 2 * C, device=lag.device, dtype=F32)
        _lib.call("dcnet_lagnorm_bwd", _p(lag), _p(nrm), _p(dlag), _p(dctx), B, T, C, _st())
        return dctx


def lagnorm(context):
    """context [B,T,2C] -> normalize_words(context[:, :, 0::2])  (model/DCNet_model.py:631-632)"""
    return _LagNorm.apply(context)


class _OnlyObj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, sim):
        raw, sim = _c(raw, name="outbox"), _c(sim, name="sim")
        B, _, N = raw.shape
        oo = torch.empty(B, N, device=raw.device, dtype=F32)
        obj = torch.empty(B, N, device=raw.device, dtype=F32)
        _lib.call("dcnet_only_obj", _p(raw), _p(sim), _p(oo), _p(obj), B, N, _st())
        ctx.save_for_backward(sim, oo)
        return oo, obj

    @staticmethod
    def backward(ctx, doo, dobj):
        sim, oo = ctx.saved_tensors
        B, N = sim.shape
        doo = _c(doo, name="d only_obj") if doo is not None else None
        dobj = _c(dobj, name="d obj_score") if dobj is not None else None
        draw = torch.empty(B, 15, N, device=sim.device, dtype=F32)
        dsim = torch.empty_like(sim)
        _lib.call("dcnet_only_obj_bwd", _p(doo), _p(dobj), _p(sim), _p(oo), _p(draw), _p(dsim), B, N, _st())
        return draw, dsim


def only_obj(raw, sim):
    """raw [B,15,N], sim [B,N] -> (only_obj, obj_score) [B,N]   (model/DCNet_model.py:545-552)"""
    return _OnlyObj.apply(raw, sim)


class _Modulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, sim, loc):
        raw, sim, loc = _c(raw, name="outbox"), _c(sim, name="sim"), _c(loc, name="loc")
        B, _, N = raw.shape
        out = torch.empty_like(raw)
        _lib.call("dcnet_modulate_conf_fwd", _p(raw), _p(sim), _p(loc), _p(out), B, N, _st())
        ctx.save_for_backward(raw, sim, loc)
        return out

    @staticmethod
    def backward(ctx, dout):
        raw, sim, loc = ctx.saved_tensors
        B, _, N = raw.shape
        dout = _c(dout, name="dout")
        draw, dsim, dloc = torch.empty_like(raw), torch.empty_like(sim), torch.empty_like(loc)
        _lib.call("dcnet_modulate_conf_bwd", _p(raw), _p(sim), _p(loc), _p(dout), _p(draw), _p(dsim), _p(dloc), B, N, _st())
        return draw, dsim, dloc


def modulate_conf(raw, sim, loc):
    """conf logits (channels 5a+4) *= sim*loc   (model/DCNet_model.py:612-621)"""
    return _Modulate.apply(raw, sim, loc)


class _GroundLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, best_n, gi, gj, t5, partner3, w_coord, margin, *maps):
        maps = [_c(m, name="map") for m in maps]   # pred0..2, sim0..2, neg0..2, loc0..2
        B = maps[0].shape[0]
        g0 = int(round(maps[3].shape[1] ** 0.5))
        dev = maps[0].device
        losses = torch.empty(3, device=dev, dtype=F32)
        lse = torch.empty(2, B, device=dev, dtype=F32)
        _lib.call("dcnet_ground_loss_fwd", *[_p(m) for m in maps], _p(best_n), _p(gi), _p(gj), _p(t5), _p(partner3), B, g0, w_coord, margin,
                  _p(losses), _p(lse[0]), _p(lse[1]), _st())
        ctx.save_for_backward(best_n, gi, gj, t5, lse, *maps)
        ctx.partner3 = partner3
        ctx.cfg = (B, g0, w_coord, margin)
        return losses

    @staticmethod
    def backward(ctx, gl):
        best_n, gi, gj, t5, lse = ctx.saved_tensors[:5]
        maps = ctx.saved_tensors[5:]
        B, g0, w_coord, margin = ctx.cfg
        gl = _c(gl, name="grad")
        grads = [torch.empty_like(m) for m in maps]
        _lib.call("dcnet_ground_loss_bwd", *[_p(m) for m in maps], _p(best_n), _p(gi), _p(gj), _p(t5), _p(ctx.partner3), B, g0, w_coord, margin,
                  _p(lse[0]), _p(lse[1]), _p(gl), *[_p(g) for g in grads], _st())
        return (None, None, None, None, None, None, None, *grads)


def ground_losses(pred, sim, neg_sim, loc, best_n, gi, gj, t5, w_coord=5.0, margin=0.1, partner3=None):
    """pred 3 x [B,15,N_s]; sim/neg_sim/loc 3 x [B,N_s] -> tensor [yolo_loss, rank_loss, loc_loss] (train_DCNet.py:45-72,173-220).
    partner3 [3,B] int64 (best_n|gi|gj of each sample's rank-loss partner) overrides the local partner B-1-b (cross-GPU negatives)."""
    flat = [p.reshape(p.shape[0], 15, -1) for p in pred] + [s.reshape(s.shape[0], -1) for s in sim] + \
           [s.reshape(s.shape[0], -1) for s in neg_sim] + [s.reshape(s.shape[0], -1) for s in loc]
    return _GroundLoss.apply(best_n, gi, gj, t5, _c(partner3, torch.long, 'partner3'), float(w_coord), float(margin), *flat)


class _IoULoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, t, size_average):
        x, t = _c(x, name="input"), _c(t.to(x.dtype), name="target")
        acc = torch.zeros(2, device=x.device, dtype=F32)
        _lib.call("dcnet_iou_loss_sums", _p(x), _p(t), x.numel(), _p(acc), _st())
        ctx.save_for_backward(x, t, acc)
        ctx.scale = (1.0 / x.shape[0]) if size_average else 1.0
        return (x.shape[0] - acc[0] / acc[1]) * ctx.scale

    @staticmethod
    def backward(ctx, g):
        x, t, acc = ctx.saved_tensors
        dx = torch.empty_like(x)
        _lib.call("dcnet_iou_loss_bwd", _p(x), _p(t), x.numel(), _p(acc), _p(_c(g.reshape(1).float(), name="grad")), ctx.scale, _p(dx), _st())
        return dx, None, None


def iou_loss(x, target, size_average=True):
    return _IoULoss.apply(x, target, bool(size_average))
