"""operand-major dependence of the fp16 / tf32 persistent GEMM on the S-shaped instance (M = N = 2704, K = 512, 32 problems, fp32 output)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops, _lib
B, M, N, K = 32, 2704, 2704, 512
out = torch.empty(B, M, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[3] * 1e3
for dt in ("f16", "tf32"):
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = torch.randn(B, K, M, device="cuda") if a_mn else torch.randn(B, M, K, device="cuda")
            Bm = torch.randn(B, K, N, device="cuda") if b_mn else torch.randn(B, N, K, device="cuda")
            if dt == "f16":
                A, Bm = ops.cast_f16(A), ops.cast_f16(Bm)
                us = t(lambda: ops.gemm_f16(A, Bm, a_mn, b_mn, M, N, K, out=out))
            else:
                us = t(lambda: ops.gemm_tf32(A, Bm, a_mn, b_mn, M, N, K, out=out))
            print("%s  A %s  B %s : %.0f us  %.0f TFLOP/s" % (dt, "MN-major" if a_mn else "K-major ", "MN-major" if b_mn else "K-major ", us, 2.0 * B * M * N * K / us / 1e6))
