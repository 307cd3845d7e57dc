"""Generates tests/golden/dcnet_256_b4.pt by running the UNMODIFIED reference (imported via oracle/ref_loader.py) on
seeded synthetic inputs.  Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixture stores reference OUTPUTS only (sub-sampled where large); inputs and weights are regenerated from the same
seeds by the tests (dcnet_b200.synth + the mirror's reference-ordered construction, checked equal to the reference's
state_dict in tests/test_host_cpu.py)."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dcnet_b200 import synth          # noqa: E402
from oracle import ref_loader         # noqa: E402

PAIRS, SIZE, SEED = 2, 256, 13


def disable_dropout(net):
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def inputs():
    g = torch.Generator().manual_seed(1234)
    return synth.make_raw_fvisu(PAIRS, SIZE, g), synth.make_words(PAIRS, gen=g), synth.make_boxes(PAIRS, SIZE, g)


def sub(t):
    """deterministic sub-sample of a big tensor"""
    t = t.detach()
    if t.numel() <= 20000:
        return t.clone()
    return t.flatten()[::max(1, t.numel() // 10007)].clone()


def main():
    T, M, MT = ref_loader.load(SIZE)
    synth.seed_all(SEED)
    net = M.grounding_model(corpus=list(range(1000)), emb_size=512, coordmap=True).train()
    disable_dropout(net)
    maps, wid, bbox = inputs()
    maps = [m.requires_grad_(True) for m in maps]
    net.visumodel.set_maps(maps)
    random.seed(77)
    out = net(torch.zeros(2 * PAIRS, 1, 1, 1), wid, torch.zeros_like(wid))
    pred_anchor, sim_score, loc_score, fvisu, flang_attn, ff, cf, nf, vp, lp, nc = out
    gt_param, gi, gj, best_n_list, gt_center = T.build_target(bbox, pred_anchor)
    pa = [p.view(p.size(0), 3, 5, p.size(2), p.size(3)) for p in pred_anchor]
    neg_sim = [torch.sum(flang_attn[range(flang_attn.size(0) - 1, -1, -1), :, :, :] * fvisu[ii][:, :512], dim=1) for ii in range(3)]
    L = dict(yolo=T.yolo_loss(pa, gt_param, gi, gj, best_n_list),
             rank=T.rank_loss(sim_score, neg_sim, gt_center, gi, gj, best_n_list, w_coord=0.),
             interframe=T.Interframe_contrastive_loss(ff, cf, nf),
             cross=T.Crossmodal_constrastive_loss(vp, lp, nc),
             loc=T.loc_loss(loc_score, sim_score, gt_center))
    loss = L['yolo'] + 100 * L['rank'] + L['loc'] + 100 * L['interframe'] + L['cross']      # train_DCNet.py:642
    loss.backward()
    hot = ['mapping_visu.0.conv.weight', 'mapping_visu.2.bn.weight', 'corr_conv.1.0.conv.weight', 'corr_conv.0.0.bn.bias',
           'fcn_emb.0.0.conv.weight', 'fcn_emb.2.0.bn.weight', 'fcn_out.1.1.weight', 'textmodel.embedding.weight', 'sub_attn.fc.weight']
    params = dict(net.named_parameters())
    fix = dict(
        meta=dict(pairs=PAIRS, size=SIZE, seed=SEED, input_seed=1234, py_seed=77, torch=str(torch.__version__)),
        outbox=[sub(t) for t in pred_anchor], sim_score=[sub(t) for t in sim_score], loc_score=[sub(t) for t in loc_score],
        corr_feat=[sub(t) for t in fvisu], flang_attn=sub(flang_attn), neg_sim=[sub(t) for t in neg_sim],
        frame_feature=sub(torch.stack(ff)), corrspendence_feature=sub(torch.stack(cf)), neg_feature=sub(torch.stack(nf)),
        vit_posit=sub(torch.stack(vp)), lag_posit=sub(torch.stack(lp)), neg_cross=sub(torch.stack(nc)),
        losses={k: float(v) for k, v in L.items()}, loss=float(loss),
        best_n=list(best_n_list), gi=[int(v) for v in gi], gj=[int(v) for v in gj],
        grad_raw=[sub(m.grad) for m in maps], grad_raw_norm=[float(m.grad.norm()) for m in maps],
        grad_param={k: sub(params[k].grad) for k in hot}, grad_param_norm={k: float(params[k].grad.norm()) for k in hot},
    )
    # eval-mode outputs (running statistics after the single training step above, momentum 0.999)
    net.eval()
    with torch.no_grad():
        random.seed(78)
        ev = net(torch.zeros(2 * PAIRS, 1, 1, 1), wid, torch.zeros_like(wid))
    fix['eval'] = dict(outbox=[sub(t) for t in ev[0]], sim_score=[sub(t) for t in ev[1]], only_obj=[sub(t) for t in ev[3]])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dcnet_256_b4.pt")
    torch.save(fix, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; losses", fix['losses'])


if __name__ == "__main__":
    main()
