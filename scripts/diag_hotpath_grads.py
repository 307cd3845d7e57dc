"""GPU diagnostic: per-tensor gradient error of HotPath.step against the CPU oracle for each precision mode.
    python scripts/diag_hotpath_grads.py [size] [pairs] [modes: e.g. 0,3,1,4]"""
import copy, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import synth
from dcnet_b200.hotpath import HotPath
from oracle import dcnet_oracle as O
from dcnet_b200 import ops
ops.RN_TF32 = os.environ.get("DCNET_RN", "1") != "0"      # 0: operands truncated by the MMA (round 1 behaviour)
ops.BWD_FP16 = os.environ.get("DCNET_BWD_FP16", "1") != "0"
print("RN_TF32 =", ops.RN_TF32, "BWD_FP16 =", ops.BWD_FP16)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
modes = [x for x in (sys.argv[3] if len(sys.argv) > 3 else "0,3,1").split(",")]      # conv precision[:co-attention precision]
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
    hp.net.precision = int(mode.split(":")[0])
    hp.net.coattn_precision_override = int(mode.split(":")[1]) if ":" in mode else None
    c = leaves(batch, "cuda")
    random.seed(31)
    out, internals = hp.step(c['raw'], c['flang'], c['fa'], c['context'], c['head'], c['loc'], c['dy_head'], c['bbox'], return_internals=True)
    torch.cuda.synchronize()
    # the oracle again, at the product's ReLU patterns (the derivative of the function the product evaluated)
    masks = dict(map=[(t.detach() > 0).cpu() for t in internals['fv']], corr=[(t.detach() > 0).cpu() for t in internals['corr']],
                 fuse=[(t.detach() > 0).cpu() for t in internals['y']])
    cpu2 = copy.deepcopy(hp0.net).train()
    r2 = leaves(batch, "cpu")
    random.seed(31)
    O.hotpath_restated(cpu2, r2['raw'], r2['flang'], r2['fa'], r2['context'], r2['head'], r2['loc'], r2['dy_head'], r2['bbox'], size, relu_masks=masks)
    pr2 = dict(cpu2.named_parameters())
    errs2 = {}
    for k in ('flang', 'fa', 'context'):
        errs2[k] = rel(c[k].grad, r2[k].grad)
    for k in ('raw', 'head', 'loc'):
        for s in range(3):
            errs2["%s[%d]" % (k, s)] = rel(c[k][s].grad, r2[k][s].grad)
    for k, v in pr2.items():
        if v.grad is not None:
            errs2[k] = rel(dict(hp.net.named_parameters())[k].grad, v.grad)
    print("   pinned ReLU masks: worst %.2e ; " % max(errs2.values()) + ", ".join("%s %.1e" % kv for kv in sorted(errs2.items(), key=lambda kv: -kv[1])[:6]))
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
    print("== size %d pairs %d precision %s: loss %.6f (oracle %.6f)  worst grad err %.2e" % (size, pairs, mode, float(out[0]), float(o['loss']), max(errs.values())))
    for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:12]:
        print("   %-36s %.2e" % (k, e))
    sys.stdout.flush()
