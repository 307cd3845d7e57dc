"""one kernel instance of the C3 / C2 step, a few launches in a row (target for `ncu --set full -k regex:... -s N -c 1`):
    python scripts/prof_one.py {gemm_cn|gemm_s|coattn_fwd|coattn_bwd|bn_fwd|bn_bwd} [size] [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops, _lib
what = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 16
B, C, N = 2 * pairs, 512, (size // 8) ** 2
dev = "cuda"
fr = [torch.nn.functional.normalize(torch.randn(B, C, N, device=dev).abs(), dim=1) for _ in range(2)]
reps = 4
ld8 = (N + 7) // 8 * 8
def h16(t):
    o = torch.zeros(*t.shape[:-1], ld8, device=dev, dtype=torch.float16)
    o[..., :N] = ops.cast_f16(t)
    return o
def gemm16(A, a_mn, sA, Bm, b_mn, sB, out, ldc, sC, M, Nn, K, atomic):
    _lib.call("dcnet_gemm_f16", A.data_ptr(), a_mn, ld8, sA, Bm.data_ptr(), b_mn, ld8, sB, out.data_ptr(), ldc, sC, M, Nn, K, B, 1.0, atomic,
              torch.cuda.current_stream().cuda_stream)
if what == "gemm_cn":      # dFa += Fb dS^T on fp16 operands: M = 512, N = K = N2, reduce-add output
    f16, dS = h16(fr[0]), h16(torch.randn(B, N, N, device=dev))
    out = torch.zeros(B, C, N, device=dev)
    for _ in range(reps):
        gemm16(f16, 0, C * ld8, dS, 0, N * ld8, out, N, C * N, C, N, N, 1)
elif what == "gemm_s":     # S = Fa^T Fb on fp16 operands
    fa16, fb16 = h16(fr[0]), h16(fr[1])
    out = torch.empty(B, N, N, device=dev)
    for _ in range(reps):
        gemm16(fa16, 1, C * ld8, fb16, 1, C * ld8, out, N, N * N, N, N, C, 0)
elif what == "coattn_bwd":  # the whole backward of the finest scale (fp16 pipeline): absmax, S/exp, delta16, dP/dS, fix16, three [C,N] contractions
    qa = torch.arange(B, device=dev, dtype=torch.int32)
    x = fr[0].clone().requires_grad_(True)
    o = ops.coattention(x, qa, qa ^ 1, tau=10.0, precision=2)
    g = torch.randn_like(o)
    for _ in range(reps):
        x.grad = None
        o.backward(g, retain_graph=True)
elif what == "coattn_fwd":
    qa = torch.arange(B, device=dev, dtype=torch.int32)
    st = ops.coattn_stage(fr[0])
    for _ in range(reps):
        ops.coattn_fused(st, fr[0].shape, qa, qa ^ 1, tau=10.0)
elif what in ("bn_fwd", "bn_bwd"):
    cv = [torch.rand(C, device=dev) + 0.5 for _ in range(4)]
    fa = torch.nn.functional.normalize(torch.rand(B, C, device=dev), dim=1)
    y, dv = torch.empty_like(fr[0]), torch.empty_like(fr[0])
    sim = [torch.empty(B, N, device=dev) for _ in range(2)]
    sums = torch.zeros(2, C, device=dev); dfa = torch.zeros(B, C, device=dev)
    P = lambda t: t.data_ptr()
    st_ = torch.cuda.current_stream().cuda_stream
    for _ in range(reps):
        if what == "bn_fwd":
            _lib.call("dcnet_bn_act_fwd", P(fr[0]), P(cv[0]), P(cv[1]), P(cv[2]), P(cv[3]), 0.0, 1, P(y), P(fa), None, P(sim[0]), P(sim[1]), B, C, N, st_)
        else:
            _lib.call("dcnet_bn_act_bwd_reduce", P(fr[0]), P(cv[0]), P(cv[1]), P(cv[2]), P(cv[3]), 0.0, 1, P(fr[1]), P(fa), None, P(sim[0]), P(sim[1]),
                      P(dv), P(sums[0]), P(sums[1]), P(dfa), None, B, C, N, st_)
torch.cuda.synchronize()
print("done", what)
