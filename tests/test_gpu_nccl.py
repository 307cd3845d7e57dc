"""-m gpu, needs >= 2 GPUs (`gpurun --gpus 2`; skipped on the one-GPU box): BASELINE config 5's cross-GPU contrastive negatives on
the real path -- two ranks over NCCL, each running HotPath.step on its slice of a global batch -- against the CPU oracle of the
concatenated batch: the reference rank loss with the partner of global sample g being Bg-1-g (its text vector and target cell
arrive through the all-gather), everything else rank-local like under the reference's DDP (BatchNorm is not synchronised)."""
import copy
import os
import random
import socket
import subprocess
import sys

import pytest
import torch

from dcnet_b200 import synth
from dcnet_b200.hotpath import HotPath
from oracle import dcnet_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_cross_gpu_negatives_two_ranks_vs_concatenated_batch_oracle(tmp_path):
    import nccl_xneg_worker as W
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "nccl_xneg_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = [torch.load(tmp_path / ("rank%d.pt" % k)) for k in range(world)]

    # ---- oracle: the two local computations share ONE global text-vector leaf, so its gradient collects what both ranks send back
    gb = W.global_batch(world)
    fa_g = gb['fa'].clone().requires_grad_(True)
    Bl = 2 * W.PAIRS_PER_RANK
    synth.seed_all(13)
    proto = HotPath(W.SIZE).net.train()
    want = []
    for k in range(world):
        sl = slice(k * Bl, (k + 1) * Bl)
        b = W.rank_slice(gb, k)
        net = copy.deepcopy(proto)
        mk = lambda t: t.clone().requires_grad_(True)
        lv = dict(raw=[mk(t) for t in b['raw']], flang=mk(b['flang']), context=mk(b['context']), head=[mk(t) for t in b['head']],
                  loc=[mk(t) for t in b['loc']])
        random.seed(50 + k)
        o = O.hotpath_restated(net, lv['raw'], lv['flang'], fa_g[sl], lv['context'], lv['head'], lv['loc'], b['dy_head'], b['bbox'], W.SIZE,
                               fa_partner=fa_g.flip(0)[sl], bbox_partner=gb['bbox'].flip(0)[sl])
        want.append((o, lv, dict(net.named_parameters())))
    for k in range(world):
        o, lv, pr = want[k]
        g_ = got[k]
        ref = torch.stack([o['loss'], o['comp']['yolo'], o['comp']['rank'], o['comp']['loc'], o['comp']['interframe'], o['comp']['cross']])
        for i in range(6):
            assert abs(float(g_['out'][i]) - float(ref[i])) < 3e-3 * max(1.0, abs(float(ref[i]))), (k, i, float(g_['out'][i]), float(ref[i]))
        sl = slice(k * Bl, (k + 1) * Bl)
        errs = dict(fa=rel(g_['fa'], fa_g.grad[sl]), flang=rel(g_['flang'], lv['flang'].grad), context=rel(g_['context'], lv['context'].grad))
        for s in range(3):
            errs["raw%d" % s] = rel(g_['raw'][s], lv['raw'][s].grad)
            errs["head%d" % s] = rel(g_['head'][s], lv['head'][s].grad)
            errs["loc%d" % s] = rel(g_['loc'][s], lv['loc'][s].grad)
        for name, p in pr.items():
            if p.grad is not None:
                errs[name] = rel(g_['params'][name], p.grad)
        bad = {n: e for n, e in errs.items() if e > 6e-2}          # the ReLU-pattern noise bar of tests/test_gpu_hotpath.py
        assert not bad, (k, bad)
        # the rank-loss terms that cross the GPUs are not behind a ReLU of the path: tight
        assert errs['fa'] < 2e-2, errs['fa']
