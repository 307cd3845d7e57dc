"""Location branch at inference: the materialised PyTorch form (model/DCNet_model.py:556-603) against dcnet_loc_rank8_fwd.
CUDA events, L2 flushed between launches.  python scripts/prof_location.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from dcnet_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / n


for B, SN in ((16, 1344), (16, 3549)):
    C = 512
    g = torch.Generator().manual_seed(1)
    E = F.normalize(torch.rand(SN, 8, generator=g), dim=1).to(dev)
    obj = F.normalize(torch.rand(B, SN, generator=g), dim=1).to(dev)
    W = (torch.randn(C, SN, generator=g) / SN ** 0.5).to(dev)
    bias = torch.randn(C, generator=g).to(dev)
    scale, shift = (torch.rand(C, generator=g) + 0.5).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)
    f = F.normalize(torch.randn(B, C, generator=g), dim=1).to(dev)

    def materialised():
        emb = E[None].expand(B, -1, -1)
        rel = torch.bmm(emb, emb.transpose(1, 2)) * obj[:, None, :]
        y = torch.relu(F.linear(rel.reshape(-1, SN), W, bias) * scale + shift).reshape(B, SN, -1).permute(0, 2, 1)
        m = (F.normalize(y, p=2, dim=1) * f[:, :, None]).sum(1)
        mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
        return (m - mn) / (mx - mn + 1e-6)

    with torch.no_grad():
        ref = materialised()
        got = ops.loc_rank8(E, obj, W, bias, scale, shift, f)
        # both forms replayed from a CUDA graph, so that the numbers are device time and not Python launch overhead
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            materialised(); ops.loc_rank8(E, obj, W, bias, scale, shift, f)
        torch.cuda.current_stream().wait_stream(side)
        g_ref, g_new = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_ref):
            materialised()
        with torch.cuda.graph(g_new):
            ops.loc_rank8(E, obj, W, bias, scale, shift, f)
        t_ref = timed(g_ref.replay)
        t_new = timed(g_new.replay)
    alg = (C * SN + SN * 8 + 2 * B * SN + B * C * 9) * 4
    print("B=%d SN=%d: materialised %.3f ms (rel tensor %.0f MB), rank-8 %.4f ms (%.1fx), max diff %.2e, algorithmic %.1f MB -> %.0f GB/s"
          % (B, SN, t_ref, B * SN * SN * 4 / 1e6, t_new, t_ref / t_new, float((ref - got).abs().max()), alg / 1e6, alg / t_new / 1e6))


# ---- training: forward + backward of the branch (batch statistics), eager launches timed with CUDA events over 10 iterations
print("training (forward + backward, batch statistics):")
for B, SN in ((16, 1344), (32, 3549)):
    C = 512
    g = torch.Generator().manual_seed(2)
    E0 = F.normalize(torch.rand(SN, 8, generator=g), dim=1).to(dev)
    obj0 = F.normalize(torch.rand(B, SN, generator=g), dim=1).to(dev)
    lin = torch.nn.Linear(SN, C).to(dev); bn = torch.nn.BatchNorm1d(C).to(dev).train()
    f0 = F.normalize(torch.randn(B, C, generator=g), dim=1).to(dev)
    wts = torch.linspace(0.5, 1.5, SN, device=dev)

    def run(kernels):
        E, obj, f = E0.clone().requires_grad_(), obj0.clone().requires_grad_(), f0.clone().requires_grad_()
        for p_ in list(lin.parameters()) + list(bn.parameters()):
            p_.grad = None
        if kernels:
            s = ops.loc_rank8_train(E, obj, lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                    bn.momentum, bn.eps, f)
        else:
            emb = E[None].expand(B, -1, -1)
            rel = torch.bmm(emb, emb.transpose(1, 2)) * obj[:, None, :]
            y = torch.relu(bn(lin(rel.reshape(-1, SN)))).reshape(B, SN, -1).permute(0, 2, 1)
            m = (F.normalize(y, p=2, dim=1) * f[:, :, None]).sum(1)
            mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
            s = (m - mn) / (mx - mn + 1e-6)
        (s * wts).sum().backward()
        return s.detach(), obj.grad, lin.weight.grad.clone()

    a, b = run(False), run(True)
    errs = [float((x - y).norm() / y.norm()) for x, y in zip(b, a)]
    res = {}
    for kernels in (False, True):
        for _ in range(3):
            run(kernels)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            run(kernels)
        t1.record(); torch.cuda.synchronize()
        res[kernels] = (t0.elapsed_time(t1) / 10, torch.cuda.max_memory_allocated() / 2 ** 20)
    print("B=%d SN=%d: materialised PyTorch %.3f ms (peak %.0f MiB), kernels %.3f ms (peak %.0f MiB): %.1fx; difference fp32 vs fp32: scores %.1e, "
          "dobj %.1e, dW %.1e" % (B, SN, res[False][0], res[False][1], res[True][0], res[True][1], res[False][0] / res[True][0], *errs))
