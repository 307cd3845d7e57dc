"""persistent tcgen05 tf32 GEMM: single-CTA 128x256 tiles against cta_group::2 pairs (256x256 per cluster) on the shapes of the step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops, _lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = _lib.lib()


def timeit(fn, n=10):
    for _ in range(5):
        fn()
    tt = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(100000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
    return sorted(tt)[len(tt) // 2]


shapes = [  # (a_mn, b_mn, M, N, K, batch, what)
    (1, 1, 1024, 1024, 512, 16, "S = Fa^T Fb, C2 finest scale"),
    (1, 1, 2704, 2704, 512, 8, "S = Fa^T Fb, C3 finest scale (8 of 32 problems)"),
    (0, 1, 512, 2704, 2704, 32, "dFb += dO P ([C,N] output, K = N), C3"),
    (0, 0, 512, 2704, 2704, 32, "dFa += Fb dS^T, C3"),
    (0, 1, 512, 2704, 1024, 32, "corr_conv forward, C3 finest scale"),
    (0, 1, 512, 1024, 1024, 16, "corr_conv forward, C2 finest scale"),
]
for am, bm, M, N, K, B, what in shapes:
    A = torch.randn(B, K, M, device="cuda") if am else torch.randn(B, M, K, device="cuda")
    Bm = torch.randn(B, K, N, device="cuda") if bm else torch.randn(B, N, K, device="cuda")
    o = torch.empty(B, M, N, device="cuda")
    res = {}
    for v, name in ((6, "single"), (0, "pair")):
        L.dcnet_gemm_select(v)
        ms = timeit(lambda: ops.gemm_tf32(A, Bm, am, bm, M, N, K, out=o))
        res[name] = (ms, o.clone() if B * M * N < 5e7 else None)
        print("%-55s %-7s %.4f ms  %7.1f TFLOP/s" % (what, name, ms, 2.0 * M * N * K * B / ms / 1e9))
    if res["single"][1] is not None:
        print("    pair == single:", bool(torch.equal(res["single"][1], res["pair"][1])))
    L.dcnet_gemm_select(0)
    del A, Bm, o, res
    torch.cuda.empty_cache()
