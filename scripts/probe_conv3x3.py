"""one conv3x3 forward / bwd_data / bwd_weight per process: python scripts/probe_conv3x3.py B C h w what"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import _lib, ops
B, C, h, w = (int(v) for v in sys.argv[1:5]); what = sys.argv[5]
N = h * w
g = torch.Generator().manual_seed(1)
x = torch.randn(B, C, N, generator=g).cuda(); W = (torch.randn(C, C, 3, 3, generator=g) / 60).cuda(); dz = torch.randn(B, C, N, generator=g).cuda()
st = torch.cuda.current_stream().cuda_stream
wq = torch.empty(9, C, C, device="cuda"); _lib.call("dcnet_conv3x3_pack_weight", W.data_ptr(), wq.data_ptr(), C, C, 0x100, st)
xm, x0, xp = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
_lib.call("dcnet_conv3x3_shift", x.data_ptr(), xm.data_ptr(), xp.data_ptr(), x0.data_ptr(), B * C * h, w, 0x100, st)
rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
if what == "fwd":
    z = torch.empty(B, C, N, device="cuda")
    _lib.call("dcnet_conv3x3_fwd", xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), wq.data_ptr(), z.data_ptr(), B, C, C, h, w, None, st)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x0.double().cpu().view(B, C, h, w), ops.round_tf32(W).double().cpu(), padding=1)
    print(what, B, C, h, w, "ok", rel(z.view(B, C, h, w), ref))
elif what == "dw":
    dWp = torch.empty(C, 9, C, device="cuda"); dW = torch.empty(C, C, 3, 3, device="cuda")
    dzr = ops.round_tf32(dz)
    _lib.call("dcnet_conv3x3_bwd_weight", dzr.data_ptr(), xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), dWp.data_ptr(), dW.data_ptr(), B, C, C, h, w, st)
    torch.cuda.synchronize()
    Wr = ops.round_tf32(W).double().cpu().requires_grad_(True)
    torch.nn.functional.conv2d(x0.double().cpu().view(B, C, h, w), Wr, padding=1).backward(dzr.double().cpu().view(B, C, h, w))
    print(what, B, C, h, w, "ok", rel(dW, Wr.grad))
