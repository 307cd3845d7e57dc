"""ncu / timing target: the fused co-attention forward kernel alone (staging once, then the kernel `reps` times)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcnet_b200 import ops
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = (size // 8) ** 2
B = 2 * pairs
fr = torch.nn.functional.normalize(torch.randn(B, 512, N, device="cuda").abs(), dim=1)
qa = torch.arange(B, device="cuda", dtype=torch.int32)
kb = qa ^ 1
staged = ops.coattn_stage(fr)
out = torch.empty(B, 512, N, device="cuda"); lse = torch.empty(B, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.coattn_fused(staged, fr.shape, qa, kb, out=out, lse=lse)
tt = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.coattn_fused(staged, fr.shape, qa, kb, out=out, lse=lse); b.record()
    torch.cuda.synchronize(); tt.append(a.elapsed_time(b))
ms = sorted(tt)[len(tt) // 2]
print("fused N=%d pairs=%d: median %.4f ms  ->  %.1f TFLOP/s algorithmic (6cN^2/pair), %.1f executed" % (
    N, pairs, ms, 6.0 * 512 * N * N * pairs / ms / 1e9, 8.0 * 512 * N * N * pairs / ms / 1e9))
variants = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
for variant in (variants if len(sys.argv) > 4 and sys.argv[4] == "trace" else []):
    from dcnet_b200 import _lib
    import numpy as np
    print("=== variant %d" % variant)
    T = (N + 127) // 128
    nct = ((N + 63) // 64) * B
    tr = torch.zeros(nct, T + 1, 8, dtype=torch.long, device="cuda")
    oidx = torch.arange(B, device="cuda", dtype=torch.int32)
    _lib.call("dcnet_coattn_fused_fwd_trace", staged.data_ptr(), B, qa.data_ptr(), kb.data_ptr(), oidx.data_ptr(), B, out.data_ptr(), B,
              lse.data_ptr(), 512, N, 10.0, tr.data_ptr(), variant, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    tfull = tr.cpu().numpy()
    t = tfull[:, :T]
    h = tfull[:, T]
    g0 = h[:, 4].min()
    print('kernel span by globaltimer: %.1f us' % ((h[:, 5].max() - g0) / 1e3))
    for cta in (0, 1, 100, 147, 148, 200, 255):
        if cta < nct:
            print('CTA %3d sm %3d: start +%.1f us, end +%.1f us | cycles: prologue(start->Q landed) %d, loop %d, epilogue(o_full->exit) %d' % (cta, h[cta, 6], (h[cta, 4] - g0) / 1e3, (h[cta, 5] - g0) / 1e3, h[cta, 1] - h[cta, 0], h[cta, 2] - h[cta, 1], h[cta, 3] - h[cta, 2]))
    print('median prologue %d loop %d epilogue %d cycles; CTA duration median %.1f us' % (np.median(h[:, 1] - h[:, 0]), np.median(h[:, 2] - h[:, 1]), np.median(h[:, 3] - h[:, 2]), np.median(h[:, 5] - h[:, 4]) / 1e3))
    for cta in (0, 1, 100, 147, 200):
        if cta >= nct: continue
        base = t[cta, 0, 0]
        print("CTA %d (stamps relative to first block landed; slots: kv0 kv1 kv2 kv3 S-issued P-ready | S-done P-written)" % cta)
        for j in range(T):
            print("  tile %2d: " % j + " ".join("%7d" % (t[cta, j, k] - base) for k in range(8)))
    d = t[:, 1:, 0] - t[:, :-1, 0]
    print("median tile period %d cycles; kv0->kv3 %d; kv3->S-issued %d; S-issued->S-done %d; S-done->P-written %d; P-written->P-ready(mma) %d; P-ready->next kv0 %d" % (
        np.median(d), np.median(t[:, :, 3] - t[:, :, 0]), np.median(t[:, :, 4] - t[:, :, 3]), np.median(t[:, :, 6] - t[:, :, 4]),
        np.median(t[:, :, 7] - t[:, :, 6]), np.median(t[:, :, 5] - t[:, :, 7]), np.median(t[:, 1:, 0] - t[:, :-1, 5])))
