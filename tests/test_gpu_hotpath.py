"""-m gpu: the hot path on its own (what bench.py times) against the CPU oracle: losses, decode, and every gradient that
leaves the path (to the backbone, the text encoder, the head, the location branch) or lands on a hot-path parameter."""
import copy
import random

import pytest
import torch

from dcnet_b200 import synth
from dcnet_b200.hotpath import HotPath
from oracle import dcnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _leaves(batch, dev):
    mk = lambda t: t.clone().to(dev).requires_grad_(True)
    return dict(raw=[mk(t) for t in batch['raw']], flang=mk(batch['flang']), fa=mk(batch['fa']), context=mk(batch['context']),
                head=[mk(t) for t in batch['head']], loc=[mk(t) for t in batch['loc']],
                dy_head=[t.to(dev) for t in batch['dy_head']], bbox=batch['bbox'].to(dev))


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("size,pairs", [(256, 2), (256, 3), (416, 2)])
def test_hotpath_step_vs_oracle(size, pairs, precision):
    synth.seed_all(13)
    hp = HotPath(size)
    hp.net.precision = precision
    TL, TG = (1e-4, 2e-3) if precision == 0 else (3e-3, 6e-2)
    g = torch.Generator().manual_seed(500 + size + pairs)
    batch = synth.make_hotpath_batch(pairs, size, g)
    cpu = copy.deepcopy(hp.net).train()
    hp = hp.to(DEV).train()
    r = _leaves(batch, "cpu")
    random.seed(31)
    o = O.hotpath_restated(cpu, r['raw'], r['flang'], r['fa'], r['context'], r['head'], r['loc'], r['dy_head'], r['bbox'], size)
    c = _leaves(batch, DEV)
    random.seed(31)
    out = hp.step(c['raw'], c['flang'], c['fa'], c['context'], c['head'], c['loc'], c['dy_head'], c['bbox'])
    B = 2 * pairs
    want = torch.stack([o['loss'], o['comp']['yolo'], o['comp']['rank'], o['comp']['loc'], o['comp']['interframe'], o['comp']['cross']])
    for i in range(6):
        assert abs(float(out[i]) - float(want[i])) < TL * max(1.0, abs(float(want[i]))), (i, float(out[i]), float(want[i]))
    torch.testing.assert_close(out[6:].cpu(), o['iou'], rtol=max(1e-4, TL), atol=1e-5)
    # gradients leaving the path
    errs = {}
    for k in ('flang', 'fa', 'context'):
        errs[k] = rel(c[k].grad, r[k].grad)
    for k in ('raw', 'head', 'loc'):
        for s in range(3):
            errs["%s[%d]" % (k, s)] = rel(c[k][s].grad, r[k][s].grad)
    pc, pr = dict(hp.net.named_parameters()), dict(cpu.named_parameters())
    n_checked = 0
    for k, v in pr.items():
        if v.grad is None:
            continue
        n_checked += 1
        errs[k] = rel(pc[k].grad, v.grad)
    assert n_checked == 27
    bad = {k: e for k, e in errs.items() if e > TG}
    assert not bad, bad


def test_hotpath_graph_replay_equals_eager():
    """bench.py replays the step from a CUDA graph; the captured step must reproduce the eager one (same kernels, same
    order).  Not bit for bit: BatchNorm sums and weight gradients are accumulated with fp32 atomics whose order varies from
    run to run, and a last-bit change of a batch statistic can flip a ReLU mask, so gradients agree to ~1e-3, losses to 1e-5."""
    size, pairs = 256, 2
    synth.seed_all(13)
    hp = HotPath(size).to(DEV).train()
    g = torch.Generator().manual_seed(77)
    batch = synth.make_hotpath_batch(pairs, size, g)
    B = 2 * pairs
    random.seed(5)
    negpos, negidx = hp.draw_indices(B)
    static = _leaves(batch, DEV)
    s_negpos = torch.from_numpy(negpos).to(DEV)
    s_negidx = torch.from_numpy(negidx).to(DEV)

    def run():
        return hp.step(static['raw'], static['flang'], static['fa'], static['context'], static['head'], static['loc'], static['dy_head'],
                       static['bbox'], s_negpos, s_negidx)

    def clear():
        for p in hp.parameters():
            p.grad = None
        for k in ('flang', 'fa', 'context'):
            static[k].grad = None
        for k in ('raw', 'head', 'loc'):
            for t in static[k]:
                t.grad = None

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            clear()
            eager = run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager = eager.clone()
    eager_graw = static['raw'][2].grad.clone()
    eager_gw = hp.net.corr_conv[2][0].conv.weight.grad.clone()
    clear()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        res = run()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    torch.testing.assert_close(res, eager, rtol=1e-4, atol=1e-5)
    assert rel(static['raw'][2].grad, eager_graw) < 1e-2
    assert rel(hp.net.corr_conv[2][0].conv.weight.grad, eager_gw) < 1e-2


def test_hotpath_streams_do_not_change_the_result():
    """HotPath.scale_streams: one CUDA stream per pyramid scale plus two auxiliary streams for the sampling blocks, the fusion
    terms and the decode.  Same kernels on the same data as the single-stream step: losses / IoU to 1e-5, gradients to the
    atomics' noise (see test_hotpath_graph_replay_equals_eager); the host RNG stream ends in the same state."""
    size, pairs = 256, 2
    B = 2 * pairs
    results = []
    for streams in (True, False):
        synth.seed_all(13)
        hp = HotPath(size).to(DEV).train()
        hp.scale_streams = streams
        g = torch.Generator().manual_seed(78)
        batch = synth.make_hotpath_batch(pairs, size, g)
        static = _leaves(batch, DEV)
        random.seed(6)
        out = hp.step(static['raw'], static['flang'], static['fa'], static['context'], static['head'], static['loc'], static['dy_head'],
                      static['bbox'])
        torch.cuda.synchronize()
        results.append((out.clone(), static['raw'][2].grad.clone(), static['raw'][0].grad.clone(), static['fa'].grad.clone(),
                        hp.net.fcn_emb[1][0].conv.weight.grad.clone(), random.random()))
    a, b = results
    torch.testing.assert_close(a[0], b[0], rtol=1e-4, atol=1e-5)
    for x, y in zip(a[1:5], b[1:5]):
        assert rel(x, y) < 1e-2, rel(x, y)
    assert a[5] == b[5]
