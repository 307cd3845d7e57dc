"""-m "not gpu": the N>1 host logic on CPU with the gloo backend, world size 2 (SURVEY section 8e / task note 5).
The math inside is the oracle's (the CUDA ops cannot run here); what is under test is dcnet_b200.parallel: the
differentiable all-gather, the global partner bookkeeping, and gradient averaging -- i.e. that W ranks with cross-GPU
negatives reproduce the reference loss on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcnet_b200 import parallel, synth
from oracle import dcnet_oracle as O

SIZE, PAIRS_PER_RANK, WORLD = 256, 2, 2


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _global_inputs():
    g = torch.Generator().manual_seed(321)
    B = 2 * PAIRS_PER_RANK * WORLD
    gs = synth.grids(SIZE)
    corr = [torch.nn.functional.normalize(torch.randn(B, 64, x * x, generator=g).abs(), dim=1) for x in gs]
    fa = torch.nn.functional.normalize(torch.randn(B, 64, generator=g).abs(), dim=1)
    bbox = synth.make_boxes(B // 2, SIZE, g)
    return corr, fa, bbox


def _rank_loss_with_partners(corr, fa, fa_neg, gtc, gtc_partner):
    """train_DCNet.py:173-203 with an explicit partner (text vector + target) per sample"""
    B = fa.shape[0]
    pos = torch.cat([(fa[:, :, None] * c).sum(1) for c in corr], 1)
    neg = torch.cat([(fa_neg[:, :, None] * c).sum(1) for c in corr], 1)
    gc = torch.cat([g[:, 4].reshape(B, -1) for g in gtc], 1)
    gp = torch.cat([g[:, 4].reshape(B, -1) for g in gtc_partner], 1)
    p = (pos * gc).sum(-1); n1 = (neg * gc).sum(-1); n2 = (pos * gp).sum(-1)
    return (torch.clamp(0.1 + n1 - p, 0) + torch.clamp(0.1 + n2 - p, 0)).sum() / (B * 2)


def _worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        corr, fa, bbox = _global_inputs()
        Bl = 2 * PAIRS_PER_RANK
        sl = slice(rank * Bl, (rank + 1) * Bl)
        fa_l = fa[sl].clone().requires_grad_(True)
        corr_l = [c[sl].clone().requires_grad_(True) for c in corr]
        gt, gi, gj, bn, gtc = O.build_target(bbox[sl], SIZE)
        fa_neg, partner3 = parallel.global_partners(fa_l, bn, gi, gj)
        # dense centre targets of the partners, rebuilt from the gathered cells
        gs = synth.grids(SIZE)
        gtc_p = [torch.zeros(Bl, 5, x, x) for x in gs]
        for b in range(Bl):
            s = int(partner3[0, b]) // 3
            gtc_p[s][b, 4, int(partner3[2, b]), int(partner3[1, b])] = 1.0
        loss = _rank_loss_with_partners(corr_l, fa_l, fa_neg, gtc, gtc_p)
        loss.backward()
        grads = [fa_l.grad] + [c.grad for c in corr_l]
        parallel.allreduce_mean_([loss.detach()])          # smoke of the flat all-reduce helper (scalar)
        red = loss.detach().clone(); dist.all_reduce(red); red /= WORLD
        out[rank] = dict(loss=float(red), fa_grad=fa_l.grad.clone(), corr_grad=[c.grad.clone() for c in corr_l],
                         partner3=partner3.clone(), fa_neg=fa_neg.detach().clone())
    finally:
        dist.destroy_process_group()


def test_cross_gpu_negatives_match_concatenated_batch():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(port, out), nprocs=WORLD, join=True)
    # single-process reference on the concatenated batch: the reference's own loss (oracle) with the local reversal
    corr, fa, bbox = _global_inputs()
    fa_r = fa.clone().requires_grad_(True)
    corr_r = [c.clone().requires_grad_(True) for c in corr]
    B = fa.shape[0]
    gs = synth.grids(SIZE)
    gt, gi, gj, bn, gtc = O.build_target(bbox, SIZE)
    sim = [O.pix2text(c, fa_r)[0].reshape(B, x, x) for c, x in zip(corr_r, gs)]
    neg = [O.pix2text(c, fa_r)[1].reshape(B, x, x) for c, x in zip(corr_r, gs)]
    ref = O.rank_loss(sim, neg, gtc)
    ref.backward()
    Bl = B // WORLD
    # mean over ranks of the local losses == global loss
    assert abs(out[0]['loss'] - float(ref)) < 1e-6
    for r in range(WORLD):
        sl = slice(r * Bl, (r + 1) * Bl)
        # partner bookkeeping: global reversal
        want = torch.stack([bn, gi, gj], 1).flip(0)[sl].t()
        assert torch.equal(out[r]['partner3'], want)
        torch.testing.assert_close(out[r]['fa_neg'], fa.flip(0)[sl])
        # DDP convention: per-rank gradients are of the LOCAL mean loss; averaged over ranks they are the global gradient,
        # i.e. local grad / WORLD == slice of the single-process gradient
        torch.testing.assert_close(out[r]['fa_grad'] / WORLD, fa_r.grad[sl], rtol=1e-5, atol=1e-7)
        for s in range(3):
            torch.testing.assert_close(out[r]['corr_grad'][s] / WORLD, corr_r[s].grad[sl], rtol=1e-5, atol=1e-7)


def test_all_gather_cat_single_process_is_identity():
    x = torch.randn(3, 4, requires_grad=True)
    assert parallel.all_gather_cat(x) is x
    fa_neg, p3 = parallel.global_partners(x, torch.tensor([0, 4, 8]), torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6]))
    torch.testing.assert_close(fa_neg, x.flip(0))
    assert p3.tolist() == [[8, 4, 0], [3, 2, 1], [6, 5, 4]]
