"""timing of dcnet_bn_act_bwd_reduce variants at the finest scale of C2 (B=16, C=512, N=1024), rotating buffers"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import _lib
B, C, N = 16, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NS = 6
dev = "cuda"
zs = [torch.randn(B, C, N, device=dev) for _ in range(NS)]
dys = [torch.randn(B, C, N, device=dev) for _ in range(NS)]
dvs = [torch.empty(B, C, N, device=dev) for _ in range(NS)]
vec = [torch.rand(C, device=dev) + 0.5 for _ in range(4)]
fa = torch.nn.functional.normalize(torch.rand(B, C, device=dev), dim=1)
sims = [torch.randn(B, N, device=dev) for _ in range(2)]
sums = torch.zeros(2, C, device=dev); dfa = torch.zeros(B, C, device=dev)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()

def run(i, l2, use_fa):
    _lib.call("dcnet_bn_act_bwd_reduce", P(zs[i]), P(vec[0]), P(vec[1]), P(vec[2]), P(vec[3]), 0.0, l2, P(dys[i]), P(fa) if use_fa else None, None,
              P(sims[0]) if use_fa else None, P(sims[1]) if use_fa else None, P(dvs[i]), P(sums[0]), P(sums[1]), P(dfa) if use_fa else None, None, B, C, N, st)

for _ in range(100):
    run(0, 1, True)
for variant in (0, 1):
    _lib.lib().dcnet_bn_bwd_select(variant)
    for l2, use_fa in ((1, True), (1, False), (0, False)):
        for i in range(NS):
            run(i, l2, use_fa)
        tt = []
        for r in range(3):
            for i in range(NS):
                torch.cuda._sleep(200000)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(i, l2, use_fa); b.record(); torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
        ms = sum(tt) / len(tt)
        print("kernel %s l2norm=%d fa=%d: %.1f us  %.0f GB/s" % ("smem-staged" if variant == 0 else "register  ", l2, use_fa, ms * 1e3, 3 * B * C * N * 4 / ms / 1e6))
_lib.lib().dcnet_bn_bwd_select(0)
