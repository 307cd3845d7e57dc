"""the inference-side callers of the hot path -- the test-time cache writer of the reference's
test_DCNet.py (save_cache :546-654, get_topk_pred_bbox :657-701) and the offline re-scoring of post_processing.py (:181-274) --
on the sm_100a kernels (dcnet_topk_boxes, dcnet_post_rescore).  The on-disk format is the reference's: one .pth per frame holding

    pred_bbox_topk  tensor [k,1,4]   boxes in the original image (xyxy, clamped)
    pred_score_topk list of k floats their confidences
    visu_feat       tensor [k,1,512] correspondence feature of each box's cell

so caches written here are readable by the reference's post_processing.py and vice versa."""
import os

import numpy as np
import torch

from . import losses as LS
from . import ops


def letterbox_image_size(ratio, dw, dh, size):
    """(img_w, img_h) the predicted boxes are clamped to (test_DCNet.py:617-624: crop of the letterboxed input, resized by 1/ratio)"""
    top, bottom = round(float(dh) - 0.1), size - round(float(dh) + 0.1)
    left, right = round(float(dw) - 0.1), size - round(float(dw) + 0.1)
    ratio = float(ratio)
    return round((right - left) / ratio), round((bottom - top) / ratio)


def topk_boxes(pred_anchor, fvisu, topk, ratio, dw, dh, size=None):
    """pred_anchor: 3 x [B,15,g,g] (or [B,3,5,g,g]) as the test model returns them, fvisu: its 3 x [B,512,g,g] correspondence
    features; ratio / dw / dh: per-image letterbox parameters (floats or [B] tensors).  -> (boxes [B,k,4], scores [B,k],
    cells [B,k,4] = (scale, anchor, gj, gi), feats [B,k,512]), all on the device."""
    size = LS.args.size if size is None else size
    B = pred_anchor[0].shape[0]
    as_list = lambda v: [float(x) for x in (v.reshape(-1).tolist() if torch.is_tensor(v) else (v if isinstance(v, (list, tuple)) else [v] * B))]
    ratio, dw, dh = as_list(ratio), as_list(dw), as_list(dh)
    meta = torch.tensor([[ratio[b], dw[b], dh[b], *letterbox_image_size(ratio[b], dw[b], dh[b], size)] for b in range(B)],
                        dtype=torch.float32).to(pred_anchor[0].device, non_blocking=True)
    return ops.topk_boxes(pred_anchor, fvisu, topk, meta, size, LS.args.anchor_imsize, LS.anchors_full)


def cache_item(boxes, scores, feats):
    """one image's (boxes [k,4], scores [k], feats [k,C]) -> the reference's cache dict (test_DCNet.py:649-653)"""
    sc = scores.detach().cpu().numpy()
    return dict(pred_bbox_topk=boxes.detach().cpu().unsqueeze(1), pred_score_topk=[sc[i] for i in range(sc.shape[0])],
                visu_feat=feats.detach().cpu().unsqueeze(1))


def cache_file(cache_dir, img_path, batch_idx):
    """./cache/<savename>/<video>/<frame>_<batch_idx>.pth (test_DCNet.py:638-647, post_processing.py:181-187)"""
    vid_name, img_name = img_path.split('/')[-2], img_path.split('/')[-1]
    return os.path.join(cache_dir, vid_name, img_name.split('.JPEG')[0] + '_' + str(batch_idx) + '.pth')


def save_cache_item(cache_dir, img_path, batch_idx, item):
    path = cache_file(cache_dir, img_path, batch_idx)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(item, path)
    return path


def read_data(img_path, frm_idx, batch_idx, center_im=None, center_im_idx=None, cache_dir='./cache'):
    """post_processing.py:181-202: a reference frame whose cache is missing falls back to the centre frame's and is reported invalid"""
    save_file = cache_file(cache_dir, img_path, batch_idx)
    invalid = -1
    if not os.path.exists(save_file):
        save_file = cache_file(cache_dir, center_im, center_im_idx)
        invalid = frm_idx
    data = torch.load(save_file, weights_only=False)      # the score list holds numpy scalars, like the reference's files
    return data['pred_bbox_topk'], torch.tensor(np.asarray(data['pred_score_topk']), dtype=torch.float), data['visu_feat'], invalid


def rescore(pred_bbox_topk, visu_feat, ref_items, invalid=(), device="cuda"):
    """post_processing.py:239-274 on the GPU.  pred_bbox_topk [k,1,4] / visu_feat [k,1,C] of the centre frame; ref_items: R x
    (scores [k], feats [k,1,C]) of the reference frames (centre included, as upstream).  -> (box [1,4], fused scores [k], index)."""
    k = visu_feat.shape[0]
    R = len(ref_items)
    centre = visu_feat.reshape(k, -1).to(device)
    ref = torch.cat([f for _, f in ref_items], dim=1).to(device)                                       # [k,R,C]
    sc = torch.stack([torch.as_tensor(s, dtype=torch.float) for s, _ in ref_items]).permute(1, 0).contiguous().to(device)
    inv = None
    if len(invalid) > 0:
        inv = torch.zeros(R, dtype=torch.int32)
        inv[list(invalid)] = 1
        inv = inv.to(device)
    fused, best, _ = ops.post_rescore(centre, ref, sc, inv)
    idx = int(best.item())
    return pred_bbox_topk[idx], fused, idx
