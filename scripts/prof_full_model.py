"""Whole-model training step after the backbone -- grounding_model.forward (text encoder, a2-a11, grounding head 8f-1, location branch
8f-2) + the five losses + backward -- on synthetic Darknet maps: this package (kernels through the C ABI, eager) against the restated
reference graph evaluated with PyTorch library ops on the same GPU (oracle port on cuda: cuBLAS / cuDNN / ATen; fp32 and TF32 allowed).
    python scripts/prof_full_model.py [size] [pairs]"""
import copy, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from dcnet_b200 import losses as LS, synth
from dcnet_b200.model.DCNet_model import grounding_model
from oracle import dcnet_oracle as O

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"


class Stub(nn.Module):
    maps = None
    def forward(self, x):
        return list(self.maps)


synth.seed_all(13)
net = grounding_model(corpus=list(range(1000)), emb_size=512, visumodel=Stub(), size=size)
for m in net.modules():
    if isinstance(m, nn.Dropout):
        m.p = 0.0
g = torch.Generator().manual_seed(7)
maps = [m.to(dev) for m in synth.make_raw_fvisu(pairs, size, g)]
wid = synth.make_words(pairs, gen=g).to(dev)
bbox = synth.make_boxes(pairs, size, g).to(dev)
ref = copy.deepcopy(net).to(dev).train()
net = net.to(dev).train()
LS.configure(size=size, anchor_imsize=416, anchors_full=O.ANCHORS_FULL)
img = torch.zeros(2 * pairs, 1, 1, 1, device=dev)


def ours():
    net.zero_grad(set_to_none=True)
    mc = [m.clone().requires_grad_(True) for m in maps]
    net.visumodel.maps = mc
    out = net(img, wid, None)
    outbox, sim, loc, corr, fa, q_if, k_if, neg_if, q_cm, k_cm, neg_cm = out
    loss, comp, _ = LS.fused_losses(outbox, sim, net.last_neg_sim_score, loc, bbox, q_if, k_if, neg_if, q_cm, k_cm, neg_cm)
    loss.backward()
    return float(loss)


def library():
    ref.zero_grad(set_to_none=True)
    mc = [m.clone().requires_grad_(True) for m in maps]
    o = O.forward_restated(ref, mc, wid)
    l = O.losses_restated(o, bbox, size)['loss']
    l.backward()
    return float(l)


def timed(fn, n):
    for _ in range(2):
        random.seed(5); v = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        random.seed(5); v = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, v


t_own, l_own = timed(ours, 10)
res = {}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
    res[tf32] = timed(library, 3)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
print("whole model after the backbone, %d frame-pairs at %dx%d, train step (forward + 5 losses + backward), wall clock with synchronize, eager:" % (pairs, size, size))
print("  dcnet_b200 (kernels through the C ABI; head and location branch included): %8.2f ms/step = %7.1f frame-pairs/s   loss %.5f" % (t_own, pairs / t_own * 1e3, l_own))
for tf32 in (False, True):
    t, l = res[tf32]
    print("  reference graph, PyTorch library ops on the same GPU (%s):            %8.2f ms/step = %7.1f frame-pairs/s   loss %.5f  -> %.1fx" % (
        "TF32 allowed" if tf32 else "fp32        ", t, pairs / t * 1e3, l, t / t_own))
print("  peak memory %.1f GiB" % (torch.cuda.max_memory_allocated() / 2 ** 30))
