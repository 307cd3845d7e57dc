"""GPU diagnostic: per-tensor gradient error of HotPath.step against the CPU oracle for each precision mode.
    python scripts/diag_hotpath_grads.py [size] [pairs] [modes: e.g. 0,3,1,4]"""
import copy, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import synth
from dcnet_b200.hotpath import HotPath
from oracle import dcnet_oracle as O

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
modes = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,3,1").split(",")]
torch.set_num_threads(os.cpu_count() or 1)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def leaves(batch, dev):
    mk = lambda t: t.clone().to(dev).requires_grad_(True)
    return dict(raw=[mk(t) for t in batch['raw']], flang=mk(batch['flang']), fa=mk(batch['fa']), context=mk(batch['context']),
                head=[mk(t) for t in batch['head']], loc=[mk(t) for t in batch['loc']],
                dy_head=[t.to(dev) for t in batch['dy_head']], bbox=batch['bbox'].to(dev))


synth.seed_all(13)
hp0 = HotPath(size)
g = torch.Generator().manual_seed(500 + size + pairs)
batch = synth.make_hotpath_batch(pairs, size, g)
cpu = copy.deepcopy(hp0.net).train()
r = leaves(batch, "cpu")
random.seed(31)
o = O.hotpath_restated(cpu, r['raw'], r['flang'], r['fa'], r['context'], r['head'], r['loc'], r['dy_head'], r['bbox'], size)
pr = dict(cpu.named_parameters())
for mode in modes:
    hp = copy.deepcopy(hp0).to("cuda").train()
    hp.net.precision = mode
    c = leaves(batch, "cuda")
    random.seed(31)
    out = hp.step(c['raw'], c['flang'], c['fa'], c['context'], c['head'], c['loc'], c['dy_head'], c['bbox'])
    torch.cuda.synchronize()
    errs = {}
    for k in ('flang', 'fa', 'context'):
        errs[k] = rel(c[k].grad, r[k].grad)
    for k in ('raw', 'head', 'loc'):
        for s in range(3):
            errs["%s[%d]" % (k, s)] = rel(c[k][s].grad, r[k][s].grad)
    pc = dict(hp.net.named_parameters())
    for k, v in pr.items():
        if v.grad is not None:
            errs[k] = rel(pc[k].grad, v.grad)
    print("== size %d pairs %d precision %d: loss %.6f (oracle %.6f)  worst grad err %.2e" % (size, pairs, mode, float(out[0]), float(o['loss']), max(errs.values())))
    for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:12]:
        print("   %-36s %.2e" % (k, e))
    sys.stdout.flush()
