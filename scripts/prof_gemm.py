"""ncu / timing target: the persistent tcgen05 GEMM on the instance bench.py reports (S = Fa^T Fb, M=N=1024, K=512, 16 problems)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
fr = torch.nn.functional.normalize(torch.randn(B, 512, N, device="cuda").abs(), dim=1)
fr2 = fr.flip(0).contiguous()
out = torch.empty(B, N, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(300):      # also lets the clocks ramp up
    ops.gemm_tf32(fr, fr2, 1, 1, N, N, 512, out=out)
tt = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.gemm_tf32(fr, fr2, 1, 1, N, N, 512, out=out); b.record()
    torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
ms = sorted(tt)[len(tt) // 2]
print("gemm2 tf32 M=N=%d K=512 batch=%d: median %.4f ms -> %.1f TFLOP/s" % (N, B, ms, 2.0 * N * N * 512 * B / ms / 1e9))

from dcnet_b200 import _lib
for v, name in ((0, "TMA stores"), (5, "direct stores"), (3, "clusters<=2"), (1, "one tile per CTA")):
    _lib.lib().dcnet_gemm_select(v)
    for _ in range(20):
        ops.gemm_tf32(fr, fr2, 1, 1, N, N, 512, out=out)
    tt = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.gemm_tf32(fr, fr2, 1, 1, N, N, 512, out=out); b.record()
        torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
    print("  variant %-18s median %.4f ms  min %.4f" % (name, sorted(tt)[5], min(tt)))
_lib.lib().dcnet_gemm_select(0)

import numpy as np
tr = torch.zeros(148, 8, 8, dtype=torch.long, device="cuda")
for (am, bm, M, Nn, K, name, dbg) in ((1, 1, N, N, 512, "S = Fa^T Fb (MN,MN) K=512, TMA stores", 0), (1, 1, N, N, 512, "same, direct stores", -5)):
    _lib.lib().dcnet_gemm_select(5 if dbg == -5 else 0)
    A = torch.randn(B, K, M, device="cuda") if am else torch.randn(B, M, K, device="cuda")
    Bm = torch.randn(B, K, Nn, device="cuda") if bm else torch.randn(B, Nn, K, device="cuda")
    o = torch.empty(B, M, Nn, device="cuda")
    for _ in range(3):
        ops.gemm_tf32(A, Bm, am, bm, M, Nn, K, out=o)
    flush.zero_(); torch.cuda.synchronize()
    tr.zero_()
    _lib.lib().dcnet_gemm_trace(tr.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.gemm_tf32(A, Bm, am, bm, M, Nn, K, out=o); b.record(); torch.cuda.synchronize()
    _lib.lib().dcnet_gemm_trace(None)
    t = tr.cpu().numpy()
    _lib.lib().dcnet_gemm_select(0)
    print("%s: %.1f us" % (name, a.elapsed_time(b) * 1e3))
    for cta in (0, 73, 147):
        base = t[cta, 0, 0]
        print("  CTA %3d:" % cta, " | ".join(" ".join("%6d" % (t[cta, j, k] - base) for k in range(6)) for j in range(5) if t[cta, j, 0] > 0))
