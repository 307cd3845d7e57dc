"""-m gpu: SURVEY 8(f) row 3 -- the test-time cache writer (dcnet_topk_boxes) and the offline re-scoring (dcnet_post_rescore)
against the oracle restatement of test_DCNet.py:593-701 / post_processing.py:239-274; the cache files round-trip through the
reference's on-disk format."""
import pytest
import torch

from dcnet_b200 import cache, ops
from dcnet_b200 import losses as LS
from oracle import dcnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("size,k", [(256, 5), (416, 8)])
def test_topk_boxes_vs_oracle(size, k):
    LS.configure(size=size, anchor_imsize=416, anchors_full=O.ANCHORS_FULL)
    g = torch.Generator().manual_seed(size + k)
    B = 3
    gs = [size // 32, size // 16, size // 8]
    pred = [torch.randn(B, 15, s, s, generator=g) for s in gs]
    pred[1][1, 5 * 2 + 4, 3, 5] = 9.0
    pred[1][1, 5 * 0 + 4, 7, 1] = 9.0                   # duplicated top confidence in image 1: both ranks -> first cell of the scale
    fv = [torch.randn(B, 512, s, s, generator=g) for s in gs]
    ratio, dw, dh = [0.8, 1.25, 0.5], [10.0, 0.0, 33.0], [20.0, 41.5, 0.0]
    boxes, scores, cells, feats = cache.topk_boxes([p.to(DEV) for p in pred], [f.to(DEV) for f in fv], k, ratio, dw, dh, size)
    for b in range(B):
        p5 = [p[b:b + 1].view(1, 3, 5, p.shape[2], p.shape[3]) for p in pred]
        ob, osc, ocells, of = O.topk_pred_boxes(p5, [f[b:b + 1] for f in fv], k, ratio[b], torch.tensor([dw[b]]), torch.tensor([dh[b]]), size)
        assert cells[b].cpu().tolist() == [list(c) for c in ocells]                       # integer outputs: exact
        assert scores[b].cpu().tolist() == osc                                            # confidences are copied, not computed
        torch.testing.assert_close(boxes[b].cpu(), ob[:, 0], rtol=2e-6, atol=2e-5)
        assert torch.equal(feats[b].cpu(), of[:, 0])
    assert cells[1, 0].tolist() == cells[1, 1].tolist() == [1, 0, 7, 1]


def test_post_rescore_vs_oracle_and_cache_round_trip(tmp_path):
    g = torch.Generator().manual_seed(11)
    k, R, C = 5, 5, 512
    items = []
    for r in range(R):
        b = torch.rand(k, 4, generator=g) * 100
        s = torch.randn(k, generator=g)
        f = torch.nn.functional.normalize(torch.randn(k, C, generator=g), dim=1)
        items.append(cache.cache_item(b, s, f))
        cache.save_cache_item(str(tmp_path), "vid/%06d.JPEG" % r, r, items[-1])
    centre_idx = R // 2
    bt, _, vf, inv = cache.read_data("vid/%06d.JPEG" % centre_idx, centre_idx, centre_idx, cache_dir=str(tmp_path))
    assert inv == -1 and bt.shape == (k, 1, 4) and vf.shape == (k, 1, C)
    refs, invalid = [], []
    for r in range(R + 0):
        name = "vid/%06d.JPEG" % r if r != 1 else "vid/missing.JPEG"                     # frame 1 has no cache: falls back to the centre
        _, sc, f, bad = cache.read_data(name, r, r, center_im="vid/%06d.JPEG" % centre_idx, center_im_idx=centre_idx, cache_dir=str(tmp_path))
        if bad > -1:
            invalid.append(bad)
        refs.append((sc, f))
    assert invalid == [1]
    box, fused, idx = cache.rescore(bt, vf, refs, invalid, device=DEV)
    of, oidx, omatch = O.post_rescore(vf, [f for _, f in refs], [s for s, _ in refs], invalid)
    torch.testing.assert_close(fused.cpu(), of, rtol=1e-5, atol=1e-6)
    assert idx == oidx and torch.equal(box, bt[oidx])
    _, _, match = ops.post_rescore(vf.reshape(k, C).to(DEV), torch.cat([f for _, f in refs], 1).to(DEV),
                                   torch.stack([s for s, _ in refs]).t().contiguous().to(DEV))
    assert torch.equal(match.cpu(), omatch)
