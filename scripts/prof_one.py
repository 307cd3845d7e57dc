"""one kernel instance of the C3 / C2 step, a few launches in a row (target for `ncu --set full -k regex:... -s N -c 1`):
    python scripts/prof_one.py {gemm_cn|gemm_s|coattn_fwd|bn_fwd|bn_bwd} [size] [pairs]"""
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
if what == "gemm_cn":      # dFa += Fb dS^T: M = 512, N = K = N2, reduce-add output
    dS = torch.randn(B, N, N, device=dev)
    out = torch.zeros(B, C, N, device=dev)
    for _ in range(reps):
        ops.gemm_tf32(fr[0], dS, 0, 0, C, N, N, out=out, atomic=1)
elif what == "gemm_s":     # S = Fa^T Fb
    out = torch.empty(B, N, N, device=dev)
    for _ in range(reps):
        ops.gemm_tf32(fr[0], fr[1], 1, 1, N, N, C, out=out)
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
