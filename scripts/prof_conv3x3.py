"""3x3 head convolution (8f-1): this library's implicit GEMM against cuDNN (fp32 and TF32), forward and backward, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import _lib, ops


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B, h, w in [(16, 8, 8), (16, 16, 16), (16, 32, 32), (32, 52, 52)]:
    C, N = 512, h * w
    x = torch.randn(B, C, N, device="cuda"); W = torch.randn(C, C, 3, 3, device="cuda") / 60; dz = torch.randn(B, C, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    wq = torch.empty(9, C, C, device="cuda"); xm, x0, xp = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    z = torch.empty(B, C, N, device="cuda"); dx = torch.empty_like(x); dWp = torch.empty(C, 9, C, device="cuda"); dW = torch.empty_like(W)
    sums = torch.empty(2 * C, device="cuda")
    t_pack = timeit(lambda: _lib.call("dcnet_conv3x3_pack_weight", W.data_ptr(), wq.data_ptr(), C, C, 0x100, st))
    t_shift = timeit(lambda: _lib.call("dcnet_conv3x3_shift", x.data_ptr(), xm.data_ptr(), xp.data_ptr(), x0.data_ptr(), B * C * h, w, 0x100, st))
    t_f = timeit(lambda: _lib.call("dcnet_conv3x3_fwd", xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), wq.data_ptr(), z.data_ptr(), B, C, C, h, w, sums.data_ptr(), st))
    t_d = timeit(lambda: _lib.call("dcnet_conv3x3_bwd_data", xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), wq.data_ptr(), dx.data_ptr(), B, C, C, h, w, st))
    t_w = timeit(lambda: _lib.call("dcnet_conv3x3_bwd_weight", dz.data_ptr(), xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), dWp.data_ptr(), dW.data_ptr(), B, C, C, h, w, st))
    gf = 2.0 * B * C * C * 9 * N / 1e9
    x4 = x.view(B, C, h, w).clone().requires_grad_(True); Wp = W.clone().requires_grad_(True); g4 = dz.view(B, C, h, w)
    res = {}
    for name, flag in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = flag
        tf = timeit(lambda: torch.nn.functional.conv2d(x4, Wp, padding=1))
        out = torch.nn.functional.conv2d(x4, Wp, padding=1)
        tb = timeit(lambda: torch.autograd.grad(out, [x4, Wp], g4, retain_graph=True))
        res[name] = (tf, tb)
    print("B=%d %dx%d (%.1f GF): own pack %.0f + shift %.0f + fwd %.0f us (%.0f TFLOP/s), bwd data %.0f + weight %.0f (+ shift %.0f) us | cuDNN fp32 fwd %.0f bwd %.0f, "
          "TF32 fwd %.0f bwd %.0f us" % (B, h, w, gf, t_pack, t_shift, t_f, gf / t_f * 1e3, t_d, t_w, t_shift, res["fp32"][0], res["fp32"][1], res["tf32"][0], res["tf32"][1]))
