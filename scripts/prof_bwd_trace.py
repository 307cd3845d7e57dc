"""clock stamps of the first tiles of one contraction of the fp16 co-attention backward (dcnet_gemm_trace):
    python scripts/prof_bwd_trace.py [stage 1..3] [size] [pairs]      stage 1 = S/exp, 2 = dP/dS, 3 = dFb += dOs E"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops, _lib
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 16
B, C, N = 2 * pairs, 512, (size // 8) ** 2
fr = torch.nn.functional.normalize(torch.randn(B, C, N, device="cuda").abs(), dim=1)
qa = torch.arange(B, device="cuda", dtype=torch.int32)
x = fr.clone().requires_grad_(True)
o = ops.coattention(x, qa, qa ^ 1, tau=10.0, precision=2)
g = torch.randn_like(o)
L = _lib.lib()
for _ in range(2):
    x.grad = None; o.backward(g, retain_graph=True)
torch.cuda.synchronize()
tr = torch.zeros(148, 8, 8, dtype=torch.long, device="cuda")
L.dcnet_coattn_bwd_fp16(stage + 1)
L.dcnet_gemm_trace(tr.data_ptr())
x.grad = None; o.backward(g, retain_graph=True)
torch.cuda.synchronize()
L.dcnet_gemm_trace(None)
L.dcnet_coattn_bwd_fp16(1)
t = tr.cpu().numpy()
print("stage %d: per CTA, tiles 0..7: stamps relative to the CTA's first (0 MMA thread at tile, 1 accumulator stage free, 2 first operand stage landed, "
      "3 MMAs issued, 4 epilogue sees the accumulator, 5 epilogue done)" % stage)
for cta in (0, 1, 74, 147):
    base = t[cta, 0, 0]
    for j in range(8):
        if t[cta, j, 0] > 0:
            print("  CTA %3d tile %d: " % (cta, j) + " ".join("%7d" % (t[cta, j, k] - base) for k in range(6)))
