"""GPU debugging aid: product vs CPU oracle vs the oracle run with PyTorch ops on the GPU (noise floor)."""
import copy, random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from dcnet_b200 import losses as LS, synth
from dcnet_b200.model.DCNet_model import grounding_model
from oracle import dcnet_oracle as O

class Stub(nn.Module):
    maps = None
    def forward(self, x): return list(self.maps)

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

size, pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 2
synth.seed_all(13)
net = grounding_model(corpus=list(range(1000)), emb_size=512, visumodel=Stub(), size=size)
for m in net.modules():
    if isinstance(m, nn.Dropout): m.p = 0.0
g = torch.Generator().manual_seed(100 + size)
maps = synth.make_raw_fvisu(pairs, size, g); wid = synth.make_words(pairs, gen=g); bbox = synth.make_boxes(pairs, size, g)
cpu_net = copy.deepcopy(net).train(); gpu_ref = copy.deepcopy(net).cuda().train(); net = net.cuda().train()
LS.configure(size=size)

def run_oracle(n, dev):
    mr = [m.clone().to(dev).requires_grad_(True) for m in maps]
    random.seed(21)
    o = O.forward_restated(n, mr, wid.to(dev), return_internals=True)
    for k in ('flang', 'context'): o[k].retain_grad()
    o['flang_attn'].retain_grad()
    ol = O.losses_restated(o, bbox, size)
    ol['loss'].backward()
    return mr, o, ol

mr, o, ol = run_oracle(cpu_net, 'cpu')
mg, og, olg = run_oracle(gpu_ref, 'cuda')
mc = [m.cuda().requires_grad_(True) for m in maps]
net.visumodel.maps = mc; net._capture = {}
random.seed(21)
out = net(torch.zeros(2 * pairs, 1, 1, 1, device='cuda'), wid.cuda(), None)
loss, comp, _ = LS.fused_losses(out[0], out[1], net.last_neg_sim_score, out[2], bbox.cuda(), *out[5:])
loss.backward()
print("loss", float(loss), float(ol['loss']), float(olg['loss']))
for s in range(3):
    print("d raw[%d]: product-vs-cpu %.2e   torch-gpu-vs-cpu %.2e   product-vs-torch-gpu %.2e" % (s, rel(mc[s].grad, mr[s].grad), rel(mg[s].grad, mr[s].grad), rel(mc[s].grad, mg[s].grad)))
cap = net._capture
print("d flang   : product-vs-cpu %.2e  torch-gpu-vs-cpu %.2e  |cpu| %.3e |prod| %.3e" % (rel(cap['flang'].grad, o['flang'].grad), rel(og['flang'].grad, o['flang'].grad), float(o['flang'].grad.norm()), float(cap['flang'].grad.norm())))
print("d fa      : product-vs-cpu %.2e  torch-gpu-vs-cpu %.2e  |cpu| %.3e |prod| %.3e" % (rel(cap['fa'].grad, o['flang_attn'].grad[:, :, 0, 0]), rel(og['flang_attn'].grad, o['flang_attn'].grad), float(o['flang_attn'].grad.norm()), float(cap['fa'].grad.norm())))
print("d context : product-vs-cpu %.2e  torch-gpu-vs-cpu %.2e  |cpu| %.3e |prod| %.3e" % (rel(cap['context'].grad, o['context'].grad), rel(og['context'].grad, o['context'].grad), float(o['context'].grad.norm()), float(cap['context'].grad.norm())))
lens = (wid != 0).sum(1)
mask = (torch.arange(20)[None, :] < lens[:, None]).float()[:, :, None]
print("d context (valid words only): product-vs-cpu %.2e torch-gpu-vs-cpu %.2e" % (rel(cap['context'].grad.cpu() * mask, o['context'].grad * mask), rel(og['context'].grad.cpu() * mask, o['context'].grad * mask)))
pc, pr, pg = dict(net.named_parameters()), dict(cpu_net.named_parameters()), dict(gpu_ref.named_parameters())
for k, v in pr.items():
    if v.grad is None or float(v.grad.norm()) < 1e-12: continue
    e1, e2 = rel(pc[k].grad, v.grad), rel(pg[k].grad, v.grad)
    if e1 > 1e-3 or e2 > 1e-3:
        print("%-40s product-vs-cpu %.2e  torch-gpu-vs-cpu %.2e  |cpu| %.3e" % (k, e1, e2, float(v.grad.norm())))
