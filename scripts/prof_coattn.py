"""ncu target: co-attention forward + backward at the finest scale of C2 / C3 (the dominant tensor-core kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
N = (size // 8) ** 2
B = 2 * pairs
fr = torch.nn.functional.normalize(torch.randn(B, 512, N, device="cuda").abs(), dim=1).requires_grad_(True)
qa = torch.arange(B, device="cuda", dtype=torch.int32)
for _ in range(3):
    out = ops.coattention(fr, qa, qa ^ 1, tau=10.0)
    out.sum().backward()
torch.cuda.synchronize()
print("done", N, pairs)
