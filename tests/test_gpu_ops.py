"""-m gpu: per-kernel parity of the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.
fp32 paths: <= 1e-5 norm-relative (stated per test); integer / index outputs: exact."""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dcnet_b200 import ops, synth
from oracle import dcnet_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,batch", [(64, 64, 512, 3), (169, 169, 512, 2), (512, 100, 1032, 2), (20, 64, 512, 4), (130, 257, 33, 1)])
def test_sgemm_strided(M, N, K, batch):
    g = gen(1)
    A = torch.randn(batch, K, M, generator=g)        # stored transposed: A(m,k) = A[b,k,m]
    B = torch.randn(batch, K, N, generator=g)
    C = torch.empty(batch, M, N, device=DEV)
    ops.sgemm(A.to(DEV), B.to(DEV), C, M, N, K, batch=batch, sA=(1, M, K * M, 0), sB=(N, 1, K * N, 0), sC=(N, 1, M * N))
    ref = torch.bmm(A.double().transpose(1, 2), B.double())
    assert rel(C, ref) < 2e-6


def test_sgemm_kbatch_atomic_colscale():
    g = gen(2)
    Bt, C_, N, K = 3, 96, 40, 70
    dz = torch.randn(Bt, C_, N, generator=g); x = torch.randn(Bt, K, N, generator=g)
    out = torch.zeros(C_, K, device=DEV)
    # dW = sum_b dz[b] x[b]^T with the batch folded into the reduction
    ops.sgemm(dz.to(DEV), x.to(DEV), out, C_, K, N, batch=1, kbatch=Bt, sA=(N, 1, 0, C_ * N), sB=(1, N, 0, K * N), sC=(K, 1, 0))
    ref = torch.einsum('bcn,bkn->ck', dz.double(), x.double())
    assert rel(out, ref) < 2e-6
    out2 = torch.zeros(C_, K, device=DEV)
    ops.sgemm(dz.to(DEV), x.to(DEV), out2, C_, K, N, batch=Bt, sA=(N, 1, C_ * N, 0), sB=(1, N, K * N, 0), sC=(K, 1, 0), atomic=1)
    assert rel(out2, ref) < 2e-6
    cs = torch.rand(K, generator=g) + 0.5
    out3 = torch.empty(C_, K, device=DEV)
    ops.sgemm(dz.to(DEV), x.to(DEV), out3, C_, K, N, batch=1, kbatch=Bt, sA=(N, 1, 0, C_ * N), sB=(1, N, 0, K * N), sC=(K, 1, 0),
              colscale=cs.to(DEV), alpha=2.0)
    assert rel(out3, 2 * ref * cs.double()[None]) < 2e-6


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 32, 1), (128, 64, 64, 2), (256, 384, 512, 2), (200, 676, 100, 1), (512, 1024, 1032, 1)])
def test_gemm_tf32_all_operand_majors(a_mn, b_mn, M, N, K, batch):
    """tcgen05 TF32 GEMM on random +/- data against fp64: any descriptor / swizzle mistake gives O(1) error; correct results
    sit at the TF32 input-rounding level (~1e-3 of the product of operand norms)."""
    g = gen(7 + M + N + K)
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, N, K, generator=g)
    ref = torch.bmm(A.double(), B.double().transpose(1, 2))
    Ad = (A.transpose(1, 2).contiguous() if a_mn else A).to(DEV)
    Bd = (B.transpose(1, 2).contiguous() if b_mn else B).to(DEV)
    if (Ad.shape[2] % 4) or (Bd.shape[2] % 4):
        pytest.skip("row pitch not TMA-addressable")
    C = ops.gemm_tf32(Ad, Bd, a_mn, b_mn, M, N, K)
    e = rel(C, ref)
    blk = ((C.double().cpu() - ref)[0].abs()[: (M // 32) * 32, : (N // 32) * 32].reshape(M // 32, 32, N // 32, 32).amax(dim=(1, 3)))
    print("gemm_tf32 a_mn=%d b_mn=%d M=%d N=%d K=%d rel err %.2e; worst 32x32 blocks:" % (a_mn, b_mn, M, N, K, e), blk.flatten().topk(3).values.tolist())
    assert e < 1e-3, e


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 64, 1), (128, 64, 128, 2), (256, 384, 512, 2), (200, 680, 104, 1), (512, 1024, 1032, 1)])
def test_gemm_bf16_all_operand_majors(a_mn, b_mn, M, N, K, batch):
    """tcgen05 kind::f16 GEMM (bf16 operands, fp32 accumulation in TMEM) against fp64 on the SAME bf16-rounded operands:
    the only error left is fp32 accumulation order (<= 1e-5); a descriptor / swizzle mistake gives O(1)."""
    g = gen(9 + M + N + K)
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, N, K, generator=g)
    Ad = ops.cast_bf16((A.transpose(1, 2).contiguous() if a_mn else A).to(DEV))
    Bd = ops.cast_bf16((B.transpose(1, 2).contiguous() if b_mn else B).to(DEV))
    assert torch.equal(Ad.cpu(), (A.transpose(1, 2).contiguous() if a_mn else A).to(torch.bfloat16))      # RNE cast
    if (Ad.shape[2] % 8) or (Bd.shape[2] % 8):
        pytest.skip("row pitch not TMA-addressable")
    Ar = (Ad.transpose(1, 2) if a_mn else Ad).double().cpu()
    Br = (Bd.transpose(1, 2) if b_mn else Bd).double().cpu()
    ref = torch.bmm(Ar, Br.transpose(1, 2))
    C = ops.gemm_bf16(Ad, Bd, a_mn, b_mn, M, N, K)
    e = rel(C, ref)
    print("gemm_bf16 a_mn=%d b_mn=%d M=%d N=%d K=%d rel err %.2e" % (a_mn, b_mn, M, N, K, e))
    assert e < 1e-5, e


@pytest.mark.parametrize("a_mn,b_mn,M,N,K,batch,atomic", [(0, 0, 512, 2704, 2704, 2, 1), (0, 1, 512, 1024, 1024, 3, 1), (1, 1, 1024, 1024, 512, 2, 0),
                                                       (1, 1, 256, 256, 512, 4, 0), (0, 0, 512, 680, 672, 2, 1), (1, 0, 384, 512, 128, 1, 0)])
def test_gemm_f16_operands(a_mn, b_mn, M, N, K, batch, atomic):
    """tcgen05 kind::f16 on fp16 operands (what the co-attention backward runs on): single-CTA tiles and cta_group::2 pairs, plain store
    and reduce-add (split reduction), against fp64 on the same fp16-rounded operands -- and, since fp16 and tf32 keep the same 11
    significant bits, the same numbers as the tf32 GEMM on those operands up to the accumulation order."""
    g = gen(19 + M + N + K)
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, N, K, generator=g)
    Ad = ops.cast_f16((A.transpose(1, 2).contiguous() if a_mn else A).to(DEV))
    Bd = ops.cast_f16((B.transpose(1, 2).contiguous() if b_mn else B).to(DEV))
    assert torch.equal(Ad.cpu(), (A.transpose(1, 2).contiguous() if a_mn else A).to(torch.float16))      # RNE cast
    Ar = (Ad.transpose(1, 2) if a_mn else Ad).double().cpu()
    Br = (Bd.transpose(1, 2) if b_mn else Bd).double().cpu()
    ref = torch.bmm(Ar, Br.transpose(1, 2))
    if atomic:
        base = torch.randn(batch, M, N, generator=g)
        C = ops.gemm_f16(Ad, Bd, a_mn, b_mn, M, N, K, alpha=0.5, out=base.to(DEV).clone(), atomic=1)
        ref = base.double() + 0.5 * ref
    else:
        C = ops.gemm_f16(Ad, Bd, a_mn, b_mn, M, N, K)
    e = rel(C, ref)
    Ct = ops.gemm_tf32(Ad.float(), Bd.float(), a_mn, b_mn, M, N, K)
    e2 = rel(ops.gemm_f16(Ad, Bd, a_mn, b_mn, M, N, K), Ct)
    print("gemm_f16 a_mn=%d b_mn=%d M=%d N=%d K=%d atomic=%d: rel err %.2e vs fp64, %.2e vs the tf32 GEMM on the same values" % (a_mn, b_mn, M, N, K, atomic, e, e2))
    assert e < 2e-5 and e2 < 2e-5, (e, e2)


@pytest.mark.parametrize("M,N,K,batch", [(200, 680, 104, 2), (512, 1024, 256, 3), (384, 512, 160, 2), (1024, 256, 64, 1)])
@pytest.mark.parametrize("b_mn", [0, 1])
def test_gemm_kernel_variants_agree_bit_for_bit(M, N, K, batch, b_mn):
    """dcnet_gemm_select: 0 = persistent kernel, TMA-store epilogue; 5 = coalesced st.global epilogue; 4 / 3 = clusters of up to 4 / 2 CTAs multicasting the B tile
    (M=512: 4 M tiles -> 4-CTA clusters; M=384: 3 tiles -> 2-CTA clusters with one idle slot; M=200: 2 tiles);
    1 = one tile per CTA with per-thread stores.  Same MMA order per tile => identical bits."""
    from dcnet_b200 import _lib
    g = gen(5 + M)
    A = torch.randn(batch, M, K, generator=g).to(DEV)
    B = torch.randn(batch, N, K, generator=g)
    Bd = (B.transpose(1, 2).contiguous() if b_mn else B).to(DEV)
    ref = torch.bmm(A.double().cpu(), B.double().transpose(1, 2))
    outs = []
    try:
        for v in (0, 5, 4, 3, 1):
            _lib.lib().dcnet_gemm_select(v)
            outs.append(ops.gemm_tf32(A, Bd, 0, b_mn, M, N, K))
    finally:
        _lib.lib().dcnet_gemm_select(0)
    assert rel(outs[0], ref) < 1e-3
    for o in outs[1:]:
        assert torch.equal(outs[0], o)


def test_conv_linear_backward_forms_tf32():
    """The two backward contractions of the 1x1 conv are linear (no ReLU kink): TF32 must hold 1e-3-class accuracy."""
    from dcnet_b200 import _lib
    g = gen(11)
    B, C, K1, K2, N = 3, 512, 256, 128, 676
    dz = torch.randn(B, C, N, generator=g); W = torch.randn(C, K1 + K2 + 8, generator=g)
    x1 = torch.randn(B, K1, N, generator=g); x2 = torch.randn(B, K2, N, generator=g)
    d = [t.to(DEV) for t in (dz, W, x1, x2)]
    dx1 = torch.empty(B, K1, N, device=DEV); dx2 = torch.empty(B, K2, N, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("dcnet_conv1x1_bwd_data", d[0].data_ptr(), d[1].data_ptr(), K1 + K2 + 8, dx1.data_ptr(), K1, dx2.data_ptr(), K2, B, C, N, 1, st)
    assert rel(dx1, torch.einsum('ck,bcn->bkn', W[:, :K1].double(), dz.double())) < 1e-3
    assert rel(dx2, torch.einsum('ck,bcn->bkn', W[:, K1:K1 + K2].double(), dz.double())) < 1e-3
    dW = torch.full((C, K1 + K2 + 8), 7.0, device=DEV)
    du = torch.empty(B, C, device=DEV); dcc = torch.empty(C, N, device=DEV)
    _lib.call("dcnet_conv1x1_bwd_weight", d[0].data_ptr(), d[2].data_ptr(), K1, d[3].data_ptr(), K2, dW.data_ptr(), K1 + K2 + 8,
              du.data_ptr(), dcc.data_ptr(), B, C, N, 1, st)
    assert rel(dW[:, :K1], torch.einsum('bcn,bkn->ck', dz.double(), x1.double())) < 1e-3
    assert rel(dW[:, K1:K1 + K2], torch.einsum('bcn,bkn->ck', dz.double(), x2.double())) < 1e-3
    assert float((dW[:, K1 + K2:] - 7.0).abs().max()) == 0.0          # columns beyond K1+K2 untouched
    assert rel(du, dz.double().sum(2)) < 1e-5 and rel(dcc, dz.double().sum(0)) < 1e-5


# ------------------------------------------------------------------------------------------------------------------
def _cbr_case(B, K1, K2, N, use_u, use_cc, use_fa, l2, training, seed):
    g = gen(seed)
    C = 512
    d = dict(x1=torch.randn(B, K1, N, generator=g), w=torch.randn(C, K1 + K2 + (8 if use_cc else 0), generator=g) / (K1 + K2) ** 0.5,
             gamma=torch.rand(C, generator=g) + 0.5, beta=torch.randn(C, generator=g) * 0.1,
             rm=torch.randn(C, generator=g) * 0.1, rv=torch.rand(C, generator=g) + 0.5)
    d['x2'] = torch.randn(B, K2, N, generator=g) if K2 else None
    d['u'] = torch.randn(B, C, generator=g) * 0.3 if use_u else None
    d['cc'] = torch.randn(C, N, generator=g) * 0.3 if use_cc else None
    d['fa'] = torch.nn.functional.normalize(torch.randn(B, C, generator=g).abs(), dim=1) if use_fa else None
    return d


def _cbr_oracle(d, l2, training, dtype=torch.float64):
    """oracle of the fused op, built from the oracle's blocks (fp64, or fp32 = the reference's own arithmetic)"""
    t = {k: (v.to(dtype).requires_grad_(k not in ('rm', 'rv')) if v is not None else None) for k, v in d.items()}
    K1 = t['x1'].shape[1]
    K2 = 0 if t['x2'] is None else t['x2'].shape[1]
    x = t['x1'] if t['x2'] is None else torch.cat([t['x1'], t['x2']], 1)
    z = torch.einsum('ck,bkn->bcn', t['w'][:, :K1 + K2], x)
    if t['u'] is not None:
        z = z + t['u'][:, :, None]
    if t['cc'] is not None:
        z = z + t['cc'][None]
    if training:
        mean, var = z.mean(dim=(0, 2)), z.var(dim=(0, 2), unbiased=False)
    else:
        mean, var = t['rm'], t['rv']
    y = torch.relu((z - mean[None, :, None]) / torch.sqrt(var[None, :, None] + O.BN_EPS) * t['gamma'][None, :, None] + t['beta'][None, :, None])
    if l2:
        y = O.l2norm_channels(y)
    outs = [y]
    if t['fa'] is not None:
        outs += list(O.pix2text(y, t['fa']))
    return t, outs, (z, mean, var)


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("B,K1,K2,N,use_u,use_cc,use_fa,l2,training", [
    (4, 1024, 0, 64, False, False, False, True, True),      # mapping_visu scale 0
    (4, 256, 0, 1024, False, False, False, True, True),     # mapping_visu scale 2
    (4, 512, 512, 256, False, False, True, True, True),     # corr_conv: two-source K loop + pix2text
    (4, 512, 0, 169, True, True, False, False, True),       # fusion: split-weight, odd N (416 scale 0)
    (2, 512, 512, 64, False, False, True, True, False),     # eval mode (running statistics)
])
def test_conv_bn_act_forward_backward(B, K1, K2, N, use_u, use_cc, use_fa, l2, training, precision):
    # precision 0: exact fp32 (CUDA cores), <= 1e-5.  precision 1: TF32 tcgen05 GEMMs with fp32 accumulation, <= 1e-3-class
    # (BASELINE north_star: "<= 1e-3 relative for bf16-accumulated-in-fp32 GEMMs"; TF32 keeps 3 more mantissa bits than bf16).
    tol_f, tol_b = (1e-5, 2e-5) if precision == 0 else (1e-3, 4e-3)
    d = _cbr_case(B, K1, K2, N, use_u, use_cc, use_fa, l2, training, seed=10 + N)
    t, outs_ref, (z_ref, mean_ref, var_ref) = _cbr_oracle(d, l2, training)
    c = {k: (v.to(DEV).requires_grad_(k not in ('rm', 'rv')) if v is not None else None) for k, v in d.items()}
    rm0, rv0 = c['rm'].clone(), c['rv'].clone()
    out = ops.conv_bn_act(c['x1'], c['w'], c['gamma'], c['beta'], c['rm'], c['rv'], training, x2=c['x2'], u=c['u'], cc=c['cc'], fa=c['fa'],
                          l2norm=l2, precision=precision)
    outs = list(out) if isinstance(out, tuple) else [out]
    for o, r in zip(outs, outs_ref):
        assert rel(o, r) < tol_f, rel(o, r)
    if training:   # running statistics: momentum 0.999, unbiased variance
        n = B * N
        assert rel(c['rm'], 0.001 * rm0.cpu().double() + 0.999 * mean_ref) < max(tol_f, 1e-5) * 5
        assert rel(c['rv'], 0.001 * rv0.cpu().double() + 0.999 * var_ref * n / (n - 1)) < max(tol_f, 1e-5) * 5
    g = gen(99)
    gouts = [torch.randn(o.shape, generator=g) for o in outs_ref]
    torch.autograd.backward(outs_ref, [x.double() for x in gouts])
    torch.autograd.backward(outs, [x.to(DEV) for x in gouts])
    # fp32 evaluation of the same graph = the reference's own arithmetic.  A pre-activation within fp32 round-off of the
    # ReLU kink flips the mask between fp32 and fp64 (plain fp32 PyTorch shows the same 1e-4 deviation from fp64 on
    # such inputs), so the gradient must match the fp64 OR the fp32 evaluation to 2e-5.
    t32, outs32, _ = _cbr_oracle(d, l2, training, torch.float32)
    torch.autograd.backward(outs32, gouts)
    K = K1 + K2
    floor = {}
    if precision == 1:
        # Reduced-precision forward => ReLU masks differ from the fp64 ones for the fraction f ~ 1e-3 of pre-activations that
        # lie within TF32 rounding of zero, and a gradient's norm-relative change is ~ sqrt(f) ~ 1-3e-2 whatever the
        # implementation.  Bar: no worse than 2x PyTorch's own TF32 evaluation (cuBLAS, allow_tf32) of the same graph.
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            dg = {k: (v.detach().to(DEV) if v is not None else None) for k, v in d.items()}
            tg, outs_g, _ = _cbr_oracle(dg, l2, training, torch.float32)
            torch.autograd.backward(outs_g, [x.to(DEV) for x in gouts])
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        floor = {k: rel(tg[k].grad, t[k].grad) for k in ('x1', 'x2', 'gamma', 'beta', 'u', 'cc', 'fa', 'w') if d[k] is not None}
        print("conv tf32 backward: torch-TF32 floor", {k: "%.1e" % v for k, v in floor.items()})
    for k in ('x1', 'x2', 'gamma', 'beta', 'u', 'cc', 'fa'):
        if d[k] is not None:
            e = min(rel(c[k].grad, t[k].grad), rel(c[k].grad, t32[k].grad))
            assert e < max(tol_b, 2 * floor.get(k, 0.0)), (k, e, floor.get(k))
    e = min(rel(c['w'].grad[:, :K], t['w'].grad[:, :K]), rel(c['w'].grad[:, :K], t32['w'].grad[:, :K]))
    assert e < max(tol_b, 2 * floor.get('w', 0.0)), ('w', e, floor.get('w'))
    assert float(c['w'].grad[:, K:].abs().max()) == 0.0 if c['w'].shape[1] > K else True


@pytest.mark.parametrize("use_terms,use_fa", [(False, False), (True, False), (False, True)])
def test_conv_bn_act_odd_pitch_runs_on_tensor_cores(use_terms, use_fa):
    """N = 169 (13x13, the coarsest scale of 416x416): the row pitch is not a multiple of 16 bytes, so the three contractions run
    on zero-padded copies (pitch 172) on tcgen05.  Same numbers as the exact-fp32 path within tf32 rounding, forward and backward."""
    g = gen(61)
    B, K1, K2, C, N = 4, 512, 512, 512, 169
    x1 = torch.randn(B, K1, N, generator=g).to(DEV); x2 = torch.randn(B, K2, N, generator=g).to(DEV)
    W = (torch.randn(C, K1 + K2 + 8, generator=g) / 32).to(DEV)
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV); beta = (torch.randn(C, generator=g) * 0.1).to(DEV)
    u = (torch.randn(B, C, generator=g) * 0.3).to(DEV); cc = (torch.randn(C, N, generator=g) * 0.3).to(DEV)
    fa = torch.nn.functional.normalize(torch.rand(B, C, generator=g), dim=1).to(DEV)
    go = torch.randn(B, C, N, generator=g).to(DEV)

    def run(prec):
        ins = [t.clone().requires_grad_(True) for t in (x1, x2, W)]
        rm = torch.zeros(C, device=DEV); rv = torch.ones(C, device=DEV)
        out = ops.conv_bn_act(ins[0], ins[2], gamma, beta, rm, rv, True, x2=ins[1], u=u if use_terms else None, cc=cc if use_terms else None,
                              fa=fa if use_fa else None, l2norm=use_fa, precision=prec)
        if use_fa:
            torch.autograd.backward(list(out), [go, torch.ones_like(out[1]), torch.ones_like(out[2])])
            return [out[0].detach(), out[1].detach()] + [t.grad for t in ins] + [rm, rv]
        out.backward(go)
        return [out.detach()] + [t.grad for t in ins] + [rm, rv]

    a, b = run(1), run(0)
    errs = [rel(p, q) for p, q in zip(a, b)]
    print("odd-pitch conv, tf32 padded vs exact fp32:", ["%.1e" % e for e in errs])
    nfwd = 2 if use_fa else 1
    assert max(errs[:nfwd]) < 1e-3 and max(errs) < 3e-2, errs     # gradients: ReLU-mask flips of the tf32 forward (see above)


@pytest.mark.parametrize("B,N,l2norm,use_fa,use_dy", [(3, 64, 1, 1, 1), (2, 676, 1, 1, 1), (4, 1024, 0, 0, 1), (2, 256, 1, 1, 0), (5, 20, 1, 1, 1)])
def test_bn_bwd_reduce_staged_kernel_matches_register_kernel(B, N, l2norm, use_fa, use_dy):
    """dcnet_bn_act_bwd_reduce: the persistent smem-staged kernel (default when C=512, N%4==0) against the register-staged one
    (dcnet_bn_bwd_select(1)) on the same inputs: dv within fp32 rounding (different summation order of the per-position norms),
    channel sums / dfa within 1e-5 of their scale; ragged last tile (N=676, N=20 < one tile)."""
    from dcnet_b200 import _lib
    g = gen(100 + N)
    C = 512
    z = torch.randn(B, C, N, generator=g).to(DEV); dy = torch.randn(B, C, N, generator=g).to(DEV)
    mean = (torch.randn(C, generator=g) * 0.1).to(DEV); invstd = (torch.rand(C, generator=g) + 0.5).to(DEV)
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV); beta = (torch.randn(C, generator=g) * 0.1).to(DEV)
    fa = torch.nn.functional.normalize(torch.rand(B, C, generator=g), dim=1).to(DEV)
    dsim = torch.randn(B, N, generator=g).to(DEV); dneg = torch.randn(B, N, generator=g).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    res = []
    try:
        for v in (0, 1):
            _lib.lib().dcnet_bn_bwd_select(v)
            dv = torch.full((B, C, N), 7.0, device=DEV); sums = torch.zeros(2, C, device=DEV); dfa = torch.zeros(B, C, device=DEV)
            _lib.call("dcnet_bn_act_bwd_reduce", z.data_ptr(), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 0.0, l2norm,
                      dy.data_ptr() if use_dy else None, fa.data_ptr() if use_fa else None, None, dsim.data_ptr() if use_fa else None,
                      dneg.data_ptr() if use_fa else None, dv.data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(),
                      dfa.data_ptr() if use_fa else None, None, B, C, N, st)
            res.append((dv, sums, dfa))
    finally:
        _lib.lib().dcnet_bn_bwd_select(0)
    (dv0, s0, f0), (dv1, s1, f1) = res
    assert rel(dv0, dv1) < 1e-6, rel(dv0, dv1)
    assert rel(s0, s1) < 1e-5 and rel(f0, f1) < 1e-5, (rel(s0, s1), rel(f0, f1))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", [0, 1, 2])
@pytest.mark.parametrize("P,N", [(2, 64), (2, 169), (1, 256), (1, 676), (1, 1024)])
def test_coattention_forward_backward(P, N, precision):
    # precision 1 = tf32 tcgen05 composition, 2 = fused tcgen05 kernel on fp16 operands (backward: the tf32 contractions with fused
    # epilogues): north_star bar 1e-3 relative, forward and backward (no ReLU in this block, so the gradient has no pattern noise)
    tol_f, tol_b = {0: (1e-5, 2e-5), 1: (1e-3, 1e-3), 2: (5e-4, 1e-3)}[precision]
    g = gen(20 + N)
    C = 512
    fr = torch.nn.functional.normalize(torch.randn(2 * P, C, N, generator=g).abs(), dim=1)
    ref_in = fr.double().requires_grad_(True)
    f1, f2 = ref_in.view(P, 2, C, N)[:, 0], ref_in.view(P, 2, C, N)[:, 1]
    o1, o2 = O.coattention(f1, f2, 10.0)
    ref = O.interleave_pairs(o1, o2)
    x = fr.to(DEV).requires_grad_(True)
    qa = torch.arange(2 * P, device=DEV, dtype=torch.int32)
    out = ops.coattention(x, qa, qa ^ 1, tau=10.0, precision=precision)
    print("coattn N=%d precision=%d fwd rel err %.2e" % (N, precision, rel(out, ref)))
    assert rel(out, ref) < tol_f, rel(out, ref)
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go.double())
    out.backward(go.to(DEV))
    print("coattn N=%d precision=%d bwd rel err %.2e" % (N, precision, rel(x.grad, ref_in.grad)))
    assert rel(x.grad, ref_in.grad) < tol_b, rel(x.grad, ref_in.grad)


@pytest.mark.parametrize("C,N", [(512, 200), (256, 1024), (128, 64), (512, 2704)])
def test_coattention_fused_staged_problems(C, N):
    """dcnet_coattn_stage + dcnet_coattn_fused_fwd: one staging of a clip, arbitrary (query frame, key frame, output row) problems,
    ragged N (last key tile / query tile partly out of range), C < 512; lse against the fp64 log-sum-exp."""
    g = gen(77 + N)
    nf = 3
    fr = torch.nn.functional.normalize(torch.randn(nf, C, N, generator=g).abs(), dim=1)
    qa = torch.tensor([0, 2, 1, 1], dtype=torch.int32)
    kb = torch.tensor([1, 0, 1, 2], dtype=torch.int32)
    oidx = torch.tensor([3, 0, 2, 1], dtype=torch.int32)
    staged = ops.coattn_stage(fr.to(DEV))
    out, lse = ops.coattn_fused(staged, fr.shape, qa.to(DEV), kb.to(DEV), oidx.to(DEV), tau=10.0)
    for i in range(4):
        S = 10.0 * fr[qa[i]].double().t() @ fr[kb[i]].double()
        ref = fr[kb[i]].double() @ torch.softmax(S, 1).t()
        # problem 2 is a frame attending to itself: softmax peaked on one key, so the output is essentially one fp16-rounded
        # column of Fb and the error is the operand rounding itself (2^-12 / sqrt(3) = 1.4e-4); distinct frames average it out
        assert rel(out[oidx[i]], ref) < (3e-4 if qa[i] == kb[i] else 2e-4), (i, rel(out[oidx[i]], ref))
        # logits carry tau x the fp16 rounding of a dot product: 10 x 2^-12 worst case on the self-pair diagonal
        assert float((lse[i].double().cpu() - torch.logsumexp(S, 1)).abs().max()) < (4e-3 if qa[i] == kb[i] else 1e-3)
    print("fused staged C=%d N=%d worst fwd rel err %.2e" % (C, N, max(rel(out[oidx[i]], fr[kb[i]].double() @ torch.softmax(
        10.0 * fr[qa[i]].double().t() @ fr[kb[i]].double(), 1).t()) for i in range(4))))


def test_coattention_clip_mode_centre_vs_others():
    """model/test_DCNet_model.py:303-332: centre frame attends to each other frame (one direction), mean of results."""
    g = gen(31)
    C, N, nf = 512, 64, 5
    fr = torch.nn.functional.normalize(torch.randn(nf, C, N, generator=g).abs(), dim=1)
    centre = nf // 2
    others = [i for i in range(nf) if i != centre]
    qa = torch.full((len(others),), centre, device=DEV, dtype=torch.int32)
    kb = torch.tensor(others, device=DEV, dtype=torch.int32)
    out = ops.coattention(fr.to(DEV), qa, kb, tau=10.0, precision=0)
    out_tc = ops.coattention(fr.to(DEV), qa, kb, tau=10.0, precision=1)
    for i, o in enumerate(others):
        o1, _ = O.coattention(fr[centre:centre + 1].double(), fr[o:o + 1].double(), 10.0)
        assert rel(out[i], o1[0]) < 1e-5
        assert rel(out_tc[i], o1[0]) < 1e-3


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,N0", [(3, 64), (2, 169)])
def test_interframe_topk_and_gather(P, N0):
    g = gen(40 + N0)
    C = 512
    fv0 = torch.nn.functional.normalize(torch.relu(synth.make_raw_fvisu(P, 32 * int(N0 ** 0.5), g)[0].flatten(2)[:, :C]), dim=1)
    idx, S0 = ops.interframe_topk(fv0.to(DEV), 30)
    f1, f2 = fv0.view(P, 2, C, N0)[:, 0], fv0.view(P, 2, C, N0)[:, 1]
    S64 = torch.bmm(f1.double().transpose(1, 2), f2.double()).flatten(1)
    assert rel(S0.flatten(1), S64) < 2e-6
    # (1) the kernel's selection is exactly the canonical top-k of ITS OWN fp32 scores (value desc, lower index first)
    for p in range(P):
        _, want = O.canonical_topk(S0[p].flatten().cpu(), 30)
        assert torch.equal(idx[p].cpu(), want)
    # (2) against the fp64 scores the selection agrees except where consecutive values are closer than fp32 round-off
    for p in range(P):
        v, want = O.canonical_topk(S64[p], 31)
        got = idx[p].cpu()
        for r in range(30):
            if got[r] != want[r]:
                assert abs(float(S64[p][got[r]] - S64[p][want[r]])) < 1e-6, (p, r)
    # negative mapping + gather
    random.seed(5)
    negpos = torch.from_numpy(ops.pyrandom_interframe(P, 30, N0, 10)).to(DEV)
    negidx = ops.interframe_negidx(idx, negpos, N0)
    col = (idx % N0).cpu()
    random.seed(5)
    for p in range(P):
        for r in range(30):
            pool = list(range(N0)); pool.remove(int(col[p, r]))
            assert random.sample(pool, 10) == negidx[p, r].cpu().tolist()
    img = torch.full((P * 30,), 1, device=DEV, dtype=torch.int32) + 2 * torch.arange(P, device=DEV, dtype=torch.int32).repeat_interleave(30)
    src = fv0.to(DEV).requires_grad_(True)
    out = ops.gather_cols(src, img, (idx % N0).reshape(-1))
    want = torch.stack([f2[p][:, col[p, r]] for p in range(P) for r in range(30)])
    assert torch.equal(out.cpu(), want)
    go = torch.randn(out.shape, generator=g)
    out.backward(go.to(DEV))
    ref_src = fv0.clone().requires_grad_(True)
    f2r = ref_src.view(P, 2, C, N0)[:, 1]
    torch.stack([f2r[p][:, col[p, r]] for p in range(P) for r in range(30)]).backward(go)
    assert rel(src.grad, ref_src.grad) < 1e-6


def test_topk_tie_rule_lower_index_first():
    # two identical frames with duplicated columns -> exact ties in S0
    g = gen(3)
    C, N0 = 512, 16
    f = torch.nn.functional.normalize(torch.randn(C, N0, generator=g).abs(), dim=0)
    f[:, 5] = f[:, 2]
    fv0 = torch.stack([f, f])
    idx, S0 = ops.interframe_topk(fv0.to(DEV), 30)
    _, want = O.canonical_topk(S0[0].flatten().cpu(), 30)
    assert torch.equal(idx[0].cpu(), want)


@pytest.mark.parametrize("R,Bq,n", [(30, 3, 10), (64, 4, 5)])
def test_infonce_forward_backward(R, Bq, n):
    g = gen(50 + n)
    C = 512
    q = torch.randn(R, Bq, C, generator=g); k = torch.randn(R, Bq, C, generator=g); neg = torch.randn(R, Bq, n, C, generator=g)
    tr = [x.double().requires_grad_(True) for x in (q, k, neg)]
    ref = O.interframe_contrastive_loss(*tr)
    tc = [x.to(DEV).requires_grad_(True) for x in (q, k, neg)]
    rows = ops.infonce_rows(tc[0].reshape(R * Bq, C), tc[1].reshape(R * Bq, C), tc[2].reshape(R * Bq, n, C), 0.07)
    loss = rows.mean()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    ref.backward(); loss.backward()
    for a, b in zip(tc, tr):
        assert rel(a.grad, b.grad) < 2e-5


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N0", [(4, 64), (3, 169)])
def test_crossmodal_block(B, N0):
    g = gen(60 + N0)
    C, T = 512, 20
    fv0 = torch.nn.functional.normalize(torch.randn(B, C, N0, generator=g).abs(), dim=1)
    ctx = torch.randn(B, T, 2 * C, generator=g)
    ctx[1, 12:] = 0                                    # padded words (pad_packed_sequence zeros)
    fw = torch.randn(T, T, 3, generator=g) * 0.2; fb = torch.randn(T, generator=g) * 0.1
    vit_r, lag_r, M_r = O.crossmodal_features(fv0.double(), ctx.double(), fw.double(), fb.double())
    x = fv0.to(DEV).requires_grad_(True); cx = ctx.to(DEV).requires_grad_(True)
    vit = ops.rownorm(x); lag = ops.lagnorm(cx)
    assert rel(vit, vit_r) < 1e-6 and rel(lag, lag_r) < 1e-6
    word, M = ops.crossmodal_words(lag.detach(), vit.detach(), fw.to(DEV), fb.to(DEV))
    word_r = M_r.argmax(1)
    # arg-max agrees except where the two best softmax values are within fp32 round-off of each other
    top2 = M_r.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-6
    assert torch.equal(word.cpu()[safe], word_r[safe])
    assert safe.float().mean() > 0.99
    # backward of the two normalisations
    gv = torch.randn(vit_r.shape, generator=g); gl = torch.randn(lag_r.shape, generator=g)
    xr = fv0.double().requires_grad_(True); cr = ctx.double().requires_grad_(True)
    vr, lr, _ = O.crossmodal_features(xr, cr, fw.double(), fb.double())
    torch.autograd.backward([vr, lr], [gv.double(), gl.double()])
    torch.autograd.backward([vit, lag], [gv.to(DEV), gl.to(DEV)])
    assert rel(x.grad, xr.grad) < 1e-5 and rel(cx.grad, cr.grad) < 1e-5


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [256, 416])
def test_build_target_exact_indices(size):
    g = gen(70 + size)
    bbox = synth.make_boxes(128, size, g)
    bbox[0] = torch.tensor([0., 0., size - 1., size - 1.])           # maximum box
    bbox[1] = torch.tensor([10., 10., 10.5, 10.5])                   # tiny box
    bn, gi, gj, t5, gt, gtc = ops.build_target(bbox.to(DEV), size, 416, O.ANCHORS_FULL, dense=True)
    ogt, ogi, ogj, obn, ogtc = O.build_target(bbox, size)
    assert torch.equal(bn.cpu(), obn) and torch.equal(gi.cpu(), ogi) and torch.equal(gj.cpu(), ogj)
    for s in range(3):
        torch.testing.assert_close(gt[s].cpu(), ogt[s], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(gtc[s].cpu(), ogtc[s], rtol=1e-6, atol=1e-6)
        assert torch.equal(gt[s].cpu() != 0, ogt[s] != 0)


@pytest.mark.parametrize("size,B", [(256, 6), (416, 4)])
def test_ground_losses_forward_backward(size, B):
    g = gen(80 + size)
    gs = [size // 32, size // 16, size // 8]
    bbox = synth.make_boxes(B // 2, size, g)
    pred = [torch.randn(B, 15, x * x, generator=g) for x in gs]
    sim = [torch.rand(B, x * x, generator=g) for x in gs]
    neg = [torch.rand(B, x * x, generator=g) for x in gs]
    loc = [torch.rand(B, x * x, generator=g) for x in gs]
    ogt, ogi, ogj, obn, ogtc = O.build_target(bbox, size)
    tr = [[t.double().requires_grad_(True) for t in grp] for grp in (pred, sim, neg, loc)]
    pr5 = [p.view(B, 3, 5, x, x) for p, x in zip(tr[0], gs)]
    ref = torch.stack([O.yolo_loss(pr5, [t.double() for t in ogt], ogi, ogj, obn),
                       O.rank_loss([s.view(B, x, x) for s, x in zip(tr[1], gs)], [s.view(B, x, x) for s, x in zip(tr[2], gs)],
                                   [t.double() for t in ogtc]),
                       O.loc_loss([s.view(B, x, x) for s, x in zip(tr[3], gs)], [t.double() for t in ogtc])])
    bn, gi, gj, t5, _, _ = ops.build_target(bbox.to(DEV), size, 416, O.ANCHORS_FULL)
    tc = [[t.to(DEV).requires_grad_(True) for t in grp] for grp in (pred, sim, neg, loc)]
    out = ops.ground_losses(tc[0], tc[1], tc[2], tc[3], bn, gi, gj, t5)
    assert rel(out, ref) < 1e-5
    w = torch.tensor([1.0, 100.0, 1.0])
    (ref * w.double()).sum().backward()
    (out * w.to(DEV)).sum().backward()
    for grp_c, grp_r, name in zip(tc, tr, ("pred", "sim", "neg", "loc")):
        for a, b in zip(grp_c, grp_r):
            if b.grad is None:
                assert a.grad is None or float(a.grad.abs().max()) == 0
            else:
                assert rel(a.grad, b.grad) < 2e-5, name


@pytest.mark.parametrize("size", [256, 416])
def test_decode_and_iou(size):
    g = gen(90 + size)
    B = 8
    gs = [size // 32, size // 16, size // 8]
    pred = [torch.randn(B, 15, x * x, generator=g) for x in gs]
    bbox = synth.make_boxes(B // 2, size, g)
    pr5 = [p.view(B, 3, 5, x, x) for p, x in zip(pred, gs)]
    # arg-max decode: integers exact
    boxes, iou, bn, gi, gj = ops.decode([p.to(DEV) for p in pred], size, 416, O.ANCHORS_FULL, target=bbox.to(DEV))
    rb, S, A, GJ, GI = O.decode_argmax(pr5, size)
    assert torch.equal(bn.cpu(), 3 * S + A) and torch.equal(gi.cpu(), GI) and torch.equal(gj.cpu(), GJ)
    torch.testing.assert_close(boxes.cpu(), rb, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(iou.cpu(), O.bbox_iou(rb, bbox), rtol=1e-4, atol=1e-6)
    # decode at the GT cell
    ogt, ogi, ogj, obn, _ = O.build_target(bbox, size)
    boxes2, iou2, _, _, _ = ops.decode([p.to(DEV) for p in pred], size, 416, O.ANCHORS_FULL, obn.to(DEV), ogi.to(DEV), ogj.to(DEV), bbox.to(DEV))
    rb2 = O.decode_at(pr5, ogi, ogj, obn, size)
    torch.testing.assert_close(boxes2.cpu(), rb2, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(ops.bbox_iou(rb2.to(DEV), bbox.to(DEV)).cpu(), O.bbox_iou(rb2, bbox), rtol=1e-6, atol=1e-7)
    # ties: equal maxima -> first in (scale, anchor, gj, gi) order
    tie = [torch.zeros(2, 15, x * x) for x in gs]
    tie[1][0, 9, 7] = 5.0; tie[2][0, 4, 3] = 5.0; tie[0][1, 14, 2] = 1.0; tie[0][1, 4, 60] = 1.0
    _, _, bn, gi, gj = ops.decode([p.to(DEV) for p in tie], size, 416, O.ANCHORS_FULL)
    _, S, A, GJ, GI = O.decode_argmax([p.view(2, 3, 5, x, x) for p, x in zip(tie, gs)], size)
    assert torch.equal(bn.cpu(), 3 * S + A) and torch.equal(gi.cpu(), GI) and torch.equal(gj.cpu(), GJ)


def test_only_obj_and_modulate():
    g = gen(100)
    B, N = 4, 256
    raw = torch.randn(B, 15, N, generator=g); sim = torch.rand(B, N, generator=g); loc = torch.rand(B, N, generator=g)
    tr = [x.double().requires_grad_(True) for x in (raw, sim, loc)]
    tc = [x.to(DEV).requires_grad_(True) for x in (raw, sim, loc)]
    oo_r = O.only_obj(tr[0]); obj_r = oo_r * tr[1]; mod_r = O.modulate_conf(tr[0], tr[1], tr[2])
    oo, obj = ops.only_obj(tc[0], tc[1]); mod = ops.modulate_conf(tc[0], tc[1], tc[2])
    assert rel(oo, oo_r) < 1e-6 and rel(obj, obj_r) < 1e-6 and rel(mod, mod_r) < 1e-6
    gs_ = [torch.randn(x.shape, generator=g) for x in (oo_r, obj_r, mod_r)]
    torch.autograd.backward([oo_r, obj_r, mod_r], [x.double() for x in gs_])
    torch.autograd.backward([oo, obj, mod], [x.to(DEV) for x in gs_])
    for a, b in zip(tc, tr):
        assert rel(a.grad, b.grad) < 1e-5


@pytest.mark.parametrize("g_", [8, 13, 32])
def test_yolo_layer_decode(g_):
    gg = gen(110 + g_)
    anchors = [(116, 90), (156, 198), (373, 326)]
    x = torch.randn(3, 255, g_, g_, generator=gg)
    ref = O.yolo_layer_decode(x, anchors, 80, 256)
    out = ops.yolo_layer_decode(x.to(DEV), anchors, 80, 256)
    torch.testing.assert_close(out.cpu(), ref, rtol=2e-5, atol=1e-5)


def test_iou_loss():
    g = gen(120)
    x = torch.randn(4, 1, 32, 32, generator=g); t = (torch.rand(4, 1, 32, 32, generator=g) > 0.7).float()
    xr = x.double().requires_grad_(True)
    ref = O.iou_loss(xr, t.double())
    xc = x.to(DEV).requires_grad_(True)
    out = ops.iou_loss(xc, t.to(DEV))
    assert abs(float(out) - float(ref)) < 1e-5
    ref.backward(); out.backward()
    assert rel(xc.grad, xr.grad) < 1e-4


def test_coord_map_matches_reference_convention():
    for h, w in [(8, 8), (13, 13), (16, 32)]:
        torch.testing.assert_close(ops.coord_map(h, w, DEV).cpu(), O.coord_map(h, w), rtol=1e-6, atol=1e-7)


def test_empty_inputs_are_noops():
    e = torch.empty(0, device=DEV)
    assert ops.gather_cols(torch.zeros(2, 512, 4, device=DEV), torch.empty(0, device=DEV, dtype=torch.int32),
                           torch.empty(0, device=DEV, dtype=torch.long)).shape == (0, 512)
    assert ops.infonce_rows(torch.empty(0, 512, device=DEV), torch.empty(0, 512, device=DEV), torch.empty(0, 5, 512, device=DEV)).shape == (0,)


def test_explicit_negative_partners_equal_local_reversal():
    """fa_neg / partner3 (cross-GPU negatives) given the local reversal must reproduce the implicit B-1-b path bit for bit,
    forward and backward -- the multi-rank semantics on top are covered by tests/test_parallel_gloo.py."""
    g = gen(777)
    B, C, N, size = 4, 512, 64, 256
    d = _cbr_case(B, 512, 512, N, False, False, True, True, True, seed=5)
    outs = []
    for explicit in (False, True):
        c = {k: (v.to(DEV).requires_grad_(k not in ('rm', 'rv')) if v is not None else None) for k, v in d.items()}
        fa_neg = c['fa'].detach().flip(0).clone().requires_grad_(True) if explicit else None
        y, sim, neg = ops.conv_bn_act(c['x1'], c['w'], c['gamma'], c['beta'], c['rm'], c['rv'], True, x2=c['x2'], fa=c['fa'], l2norm=True,
                                      precision=0, fa_neg=fa_neg)
        (sim.sum() + 2 * neg.sum() + y.mean()).backward()
        dfa = c['fa'].grad.clone()
        if explicit:
            dfa = dfa + fa_neg.grad.flip(0)
        outs.append((sim.detach(), neg.detach(), dfa, c['x1'].grad.clone()))
    for a, b in zip(*outs):        # fp32 atomics: summation order differs between the two paths
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)
    # rank loss with explicit partner cells
    gs = [8, 16, 32]
    bbox = synth.make_boxes(B // 2, size, g)
    pred = [torch.randn(B, 15, x * x, generator=g).to(DEV) for x in gs]
    sim = [torch.rand(B, x * x, generator=g).to(DEV).requires_grad_(True) for x in gs]
    neg = [torch.rand(B, x * x, generator=g).to(DEV) for x in gs]
    loc = [torch.rand(B, x * x, generator=g).to(DEV) for x in gs]
    bn, gi, gj, t5, _, _ = ops.build_target(bbox.to(DEV), size, 416, O.ANCHORS_FULL)
    l0 = ops.ground_losses(pred, sim, neg, loc, bn, gi, gj, t5)
    p3 = torch.stack([bn, gi, gj]).flip(1).contiguous()
    l1 = ops.ground_losses(pred, sim, neg, loc, bn, gi, gj, t5, partner3=p3)
    torch.testing.assert_close(l0, l1, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("B,SN,C", [(2, 1344, 512), (3, 3549, 512), (1, 85, 64), (4, 21, 8)])
def test_loc_rank8_matches_materialised_relation(B, SN, C):
    """the location branch at inference without the [B,SN,SN] tensor, against the materialised formula of
    model/DCNet_model.py:556-603 evaluated in fp64 (fp32 kernel, different summation order: bars 2e-5 on the unit-scale raw scores)."""
    g = torch.Generator().manual_seed(31 + SN)
    E = F.normalize(torch.rand(SN, 8, generator=g), dim=1)
    obj = F.normalize(torch.rand(B, SN, generator=g), dim=1)
    W = torch.randn(C, SN, generator=g) * (3.0 / SN ** 0.5)
    bias = torch.randn(C, generator=g) * 0.1
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    f = F.normalize(torch.randn(B, C, generator=g), dim=1)
    d = lambda t: t.double()
    rel = torch.bmm(d(E)[None].expand(B, -1, -1), d(E).t()[None].expand(B, -1, -1)) * d(obj)[:, None, :]
    y = torch.relu((rel @ d(W).t() + d(bias)) * d(scale) + d(shift)).permute(0, 2, 1)
    m = (F.normalize(y, dim=1) * d(f)[:, :, None]).sum(1)
    mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
    ref = (m - mn) / (mx - mn + 1e-6)
    score, raw, G = ops.loc_rank8(*(t.to(DEV) for t in (E, obj, W, bias, scale, shift, f)), return_raw=True)
    Gref = torch.einsum('cq,bq,qk->bck', d(W), d(obj), d(E))
    assert float((G.double().cpu() - Gref).abs().max()) < 1e-5 * float(Gref.abs().max()) + 1e-7
    assert float((raw.double().cpu() - m).abs().max()) < 2e-5
    spread = float((mx - mn).min())
    assert float((score.double().cpu() - ref).abs().max()) < 4e-5 / max(spread, 1e-3)
    assert float(score.min()) >= 0.0 and float(score.max()) <= 1.0
    # no bias
    score2 = ops.loc_rank8(E.to(DEV), obj.to(DEV), W.to(DEV), None, scale.to(DEV), (shift + scale * bias).to(DEV), f.to(DEV))
    assert float((score2 - score).abs().max()) < 1e-5 / max(spread, 1e-3)


@pytest.mark.parametrize("precision,N", [(0, 256), (1, 256), (0, 169)])
def test_fusion_layer_with_text_and_coordinate_terms(precision, N):
    """a8 (model/DCNet_model.py:489-505): cat([corr_feat, flang tile, coord]) -> 1x1 conv + BN + ReLU in its split-weight form, the text /
    coordinate terms and their gradients on the library's own kernels (dcnet_fuse_terms_fwd/bwd), against the dense form in fp64."""
    B, C, Ct = 4, 512, 512
    g = gen(300 + N)
    h = int(round(N ** 0.5))
    x = torch.randn(B, C, N, generator=g)
    flang = torch.nn.functional.normalize(torch.randn(B, Ct, generator=g).abs(), dim=1)
    w = torch.randn(C, C + Ct + 8, generator=g) / 30
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    coord = O.coord_map(h, h).flatten(1)
    gy = torch.randn(B, C, N, generator=g)

    def dense(dt):
        t = dict(x=x.clone().to(dt).requires_grad_(True), flang=flang.clone().to(dt).requires_grad_(True), w=w.clone().to(dt).requires_grad_(True),
                 gamma=gamma.clone().to(dt).requires_grad_(True), beta=beta.clone().to(dt).requires_grad_(True))
        xin = torch.cat([t['x'], t['flang'][:, :, None].expand(B, Ct, N), coord.to(dt)[None].expand(B, 8, N)], 1)
        y = O.conv1x1_bn_relu(xin, t['w'], t['gamma'], t['beta'], training=True)
        y.backward(gy.to(dt))
        return t, y
    t64, y64 = dense(torch.float64)
    t32, y32 = dense(torch.float32)
    c = dict(x=x.to(DEV).requires_grad_(True), flang=flang.to(DEV).requires_grad_(True), w=w.to(DEV).requires_grad_(True),
             gamma=gamma.to(DEV).requires_grad_(True), beta=beta.to(DEV).requires_grad_(True))
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    y = ops.conv_bn_act(c['x'], c['w'], c['gamma'], c['beta'], rm, rv, True, flang=c['flang'], coords=ops.coord_map(h, h, DEV).flatten(1),
                        precision=precision)
    tol_f, tol_b = (1e-5, 1e-4) if precision == 0 else (1e-3, 3e-2)      # tf32: ReLU-pattern noise, see test_conv_bn_act_forward_backward
    assert rel(y, y64) < tol_f, rel(y, y64)
    y.backward(gy.to(DEV))
    for k in ('x', 'flang', 'w', 'gamma', 'beta'):
        e = min(rel(c[k].grad, t64[k].grad), rel(c[k].grad, t32[k].grad))
        assert e < tol_b, (k, e)
    # the text and coordinate columns of the weight gradient on their own (they are written by dcnet_fuse_terms_bwd)
    for lo, hi in ((C, C + Ct), (C + Ct, C + Ct + 8)):
        e = min(rel(c['w'].grad[:, lo:hi], t64['w'].grad[:, lo:hi]), rel(c['w'].grad[:, lo:hi], t32['w'].grad[:, lo:hi]))
        assert e < tol_b, (lo, hi, e)


# ------------------------------------------------------------------------------------------------------------------
# the grounding head's 3x3 ConvBatchNormReLU and final 1x1 conv with bias on this library's kernels
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,h,w", [(2, 512, 8, 8), (3, 512, 16, 16), (2, 256, 32, 32), (1, 512, 52, 52), (2, 512, 6, 12), (1, 256, 20, 8)])
def test_conv3x3_contractions_vs_conv2d(B, C, h, w):
    """the three contractions of csrc/conv3x3.cu on their own (linear, no activation): forward against F.conv2d(padding=1), data
    gradient against conv_transpose2d, weight gradient against the autograd of conv2d -- all in fp64 on the same tf32-rounded
    operands (what the kernels are handed by their producers), bar 5e-5 (exact tf32 products, the tensor core's fp32 accumulation over 4608 terms: measured 1e-5) and 1e-3
    against the unrounded fp64 result."""
    from dcnet_b200 import _lib
    g = gen(300 + h * w)
    N = h * w
    x = torch.randn(B, C, N, generator=g).to(DEV)
    W = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    dz = torch.randn(B, C, N, generator=g).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    RN = 0x100
    wq = torch.empty(9, C, C, device=DEV)
    _lib.call("dcnet_conv3x3_pack_weight", W.data_ptr(), wq.data_ptr(), C, C, RN, st)
    xm, x0, xp = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    _lib.call("dcnet_conv3x3_shift", x.data_ptr(), xm.data_ptr(), xp.data_ptr(), x0.data_ptr(), B * C * h, w, RN, st)
    # the shifted copies themselves
    x4 = x0.view(B, C, h, w)
    assert torch.equal(xm.view(B, C, h, w)[..., 1:], x4[..., :-1]) and float(xm.view(B, C, h, w)[..., 0].abs().max()) == 0.0
    assert torch.equal(xp.view(B, C, h, w)[..., :-1], x4[..., 1:]) and float(xp.view(B, C, h, w)[..., -1].abs().max()) == 0.0
    assert rel(x0, x) < 3e-4 and torch.equal(wq.permute(1, 2, 0).reshape(C, C, 3, 3), ops.round_tf32(W))
    z = torch.empty(B, C, N, device=DEV)
    sums = torch.empty(2 * C, device=DEV)
    _lib.call("dcnet_conv3x3_fwd", xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), wq.data_ptr(), z.data_ptr(), B, C, C, h, w, sums.data_ptr(), st)
    Wr = ops.round_tf32(W).double().cpu().requires_grad_(True)
    xr = x0.double().cpu().view(B, C, h, w).requires_grad_(True)
    zr = torch.nn.functional.conv2d(xr, Wr, padding=1)
    e_f = rel(z.view(B, C, h, w), zr)
    e_f_exact = rel(z.view(B, C, h, w), torch.nn.functional.conv2d(x.double().cpu().view(B, C, h, w), W.double().cpu(), padding=1))
    assert rel(sums[:C], zr.sum((0, 2, 3))) < 1e-4 and rel(sums[C:], (zr * zr).sum((0, 2, 3))) < 1e-4
    dzr = ops.round_tf32(dz)
    zr.backward(dzr.double().cpu().view(B, C, h, w))
    dzm, dzp = torch.empty_like(dz), torch.empty_like(dz)
    _lib.call("dcnet_conv3x3_shift", dzr.data_ptr(), dzm.data_ptr(), dzp.data_ptr(), None, B * C * h, w, 0, st)
    dx = torch.empty_like(x)
    _lib.call("dcnet_conv3x3_bwd_data", dzm.data_ptr(), dzr.data_ptr(), dzp.data_ptr(), wq.data_ptr(), dx.data_ptr(), B, C, C, h, w, st)
    dWp = torch.empty(C, 9, C, device=DEV)
    dW = torch.empty(C, C, 3, 3, device=DEV)
    _lib.call("dcnet_conv3x3_bwd_weight", dzr.data_ptr(), xm.data_ptr(), x0.data_ptr(), xp.data_ptr(), dWp.data_ptr(), dW.data_ptr(), B, C, C, h, w, st)
    e_dx, e_dw = rel(dx.view(B, C, h, w), xr.grad), rel(dW, Wr.grad)
    print("conv3x3 %dx%d C=%d B=%d: fwd %.1e (vs unrounded fp64 %.1e), dx %.1e, dW %.1e" % (h, w, C, B, e_f, e_f_exact, e_dx, e_dw))
    assert e_f < 5e-5 and e_dx < 5e-5 and e_dw < 5e-5 and e_f_exact < 1e-3, (e_f, e_f_exact, e_dx, e_dw)


@pytest.mark.parametrize("B,h,w,training", [(2, 16, 16, True), (1, 52, 52, True), (3, 8, 8, False), (2, 26, 26, True), (3, 13, 13, True), (2, 13, 13, False)])
def test_conv3x3_bn_act_forward_backward(B, h, w, training):
    """ops.conv3x3_bn_act = ConvBatchNormReLU(512, 512, 3, 1, 1) (model/darknet.py:118-156) against plain PyTorch in fp64: outputs and
    running statistics 1e-3; gradients at the product's own ReLU pattern (the derivative of the function it evaluated) 1e-3."""
    g = gen(310 + h)
    C, N = 512, h * w
    x = torch.randn(B, C, N, generator=g)
    W = torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    rm0, rv0 = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    go = torch.randn(B, C, N, generator=g)
    c = [t.to(DEV).requires_grad_(True) for t in (x, W, gamma, beta)]
    rm, rv = rm0.to(DEV), rv0.to(DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    y = ops.conv3x3_bn_act(c[0], c[1], c[2], c[3], rm, rv, training, h, w, num_batches_tracked=nbt)
    y.backward(go.to(DEV))
    r = [t.double().requires_grad_(True) for t in (x, W, gamma, beta)]
    z = torch.nn.functional.conv2d(r[0].view(B, C, h, w), r[1], padding=1)
    rmr, rvr = rm0.double().clone(), rv0.double().clone()
    zn = torch.nn.functional.batch_norm(z, rmr, rvr, r[2], r[3], training, 0.999, 1e-5)
    yr = torch.relu(zn).view(B, C, N)
    assert rel(y, yr) < 1e-3, rel(y, yr)
    if training:
        assert rel(rm, rmr) < 1e-3 and rel(rv, rvr) < 1e-3 and int(nbt) == 1
    mask = (y.detach() > 0).cpu()
    (zn.view(B, C, N) * mask).backward(go.double())
    errs = [rel(a.grad, b.grad) for a, b in zip(c, r)]
    print("conv3x3_bn_act %dx%d train=%s: fwd %.1e, grads (x, W, gamma, beta) %s" % (h, w, training, rel(y, yr), ["%.1e" % e for e in errs]))
    assert max(errs) < 1e-3, errs


def test_conv3x3_alignment_rule_of_the_c_entry_points():
    """the contractions themselves need w % 4 == 0 (a TMA box shifted by one image row must start 16-byte aligned); ops.conv3x3_bn_act
    runs other widths at the next multiple of 4 (13 -> 16, 26 -> 28: test_conv3x3_bn_act_forward_backward)"""
    from dcnet_b200 import _lib
    L = _lib.lib()
    assert not L.dcnet_conv3x3_supported(512, 512, 26, 26) and not L.dcnet_conv3x3_supported(512, 512, 13, 13)
    assert L.dcnet_conv3x3_supported(512, 512, 52, 52) and L.dcnet_conv3x3_supported(512, 512, 26, 28)
    assert ops.conv3x3_supported(512, 512, 26, 26) and ops.conv3x3_supported(512, 512, 13, 13) and not ops.conv3x3_supported(100, 512, 8, 8)


def test_conv1x1_bias_head_output_layer():
    """fcn_out[s][1] = nn.Conv2d(256, 15, 1) with bias (model/DCNet_model.py:329-337): exact fp32"""
    g = gen(320)
    B, K, C, N = 3, 256, 15, 676
    x, W, b = torch.randn(B, K, N, generator=g), torch.randn(C, K, generator=g) / 16, torch.randn(C, generator=g)
    go = torch.randn(B, C, N, generator=g)
    c = [t.to(DEV).requires_grad_(True) for t in (x, W, b)]
    out = ops.conv1x1_bias(*c)
    out.backward(go.to(DEV))
    r = [t.double().requires_grad_(True) for t in (x, W, b)]
    ref = torch.einsum('ck,bkn->bcn', r[1], r[0]) + r[2][None, :, None]
    ref.backward(go.double())
    assert rel(out, ref) < 1e-5
    for a, b_ in zip(c, r):
        assert rel(a.grad, b_.grad) < 2e-5, rel(a.grad, b_.grad)


# ------------------------------------------------------------------------------------------------------------------
# operand rounding (DCNET_RN_TF32) and the per-problem gradient scale of the fp16 co-attention backward
# ------------------------------------------------------------------------------------------------------------------
def _is_tf32(t):
    return bool(((t.contiguous().view(torch.int32) & 0x1FFF) == 0).all())


def test_producers_round_to_nearest_tf32():
    """DCNET_RN_TF32: what a producer hands to a tf32 contraction has its 13 low mantissa bits clear (the MMA's truncation is then
    exact) and lies within half a tf32 ulp (2^-11 relative) of the unrounded result: dcnet_round_tf32, bn_act_fwd (round_out), the
    co-attention forward (round_out), and the gradient dz that bn_act_bwd_apply hands to the conv backward (seen through dx)."""
    g = gen(400)
    x = torch.randn(3, 777, generator=g).to(DEV)
    r = ops.round_tf32(x)
    assert _is_tf32(r) and float(((r - x).abs() / x.abs()).max()) <= 2.0 ** -11 + 1e-9
    # ties away from zero (cvt.rna): 1 + 2^-11 -> 1 + 2^-10
    t = torch.tensor([1.0 + 2.0 ** -11, -(1.0 + 2.0 ** -11), 1.0 + 2.0 ** -12], device=DEV)
    assert ops.round_tf32(t).tolist() == [1.0 + 2.0 ** -10, -(1.0 + 2.0 ** -10), 1.0]
    B, K, C, N = 2, 256, 512, 256
    x1 = torch.randn(B, K, N, generator=g).to(DEV)
    W = (torch.randn(C, K, generator=g) / 16).to(DEV)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.1).to(DEV)
    outs = {}
    for flag in (False, True):
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        outs[flag] = ops.conv_bn_act(x1, W, gamma, beta, rm, rv, True, l2norm=True, round_out=flag)
    assert _is_tf32(outs[True]) and not _is_tf32(outs[False])
    assert rel(outs[True], outs[False]) < 2.0 ** -11
    fr = torch.nn.functional.normalize(torch.randn(4, 512, 256, generator=g).abs(), dim=1).to(DEV)
    qa = torch.arange(4, device=DEV, dtype=torch.int32)
    from dcnet_b200.ops import _CoAttn
    a = _CoAttn.apply(fr, qa, qa ^ 1, qa, 4, 10.0, 2, True)
    b = _CoAttn.apply(fr, qa, qa ^ 1, qa, 4, 10.0, 2, False)
    assert _is_tf32(a) and not _is_tf32(b) and rel(a, b) < 2.0 ** -11


def test_conv_bwd_data_absmax_from_the_gemm_epilogue():
    """dcnet_conv1x1_bwd_data_absmax: max |dx2[b]| per image out of the data-gradient GEMM's epilogue == the maximum of the tensor it wrote"""
    from dcnet_b200 import _lib
    g = gen(401)
    B, K1, K2, C, N = 5, 512, 512, 512, 676
    dz = (torch.randn(B, C, N, generator=g) * torch.logspace(-6, 2, B)[:, None, None]).to(DEV)       # images on very different scales
    W = (torch.randn(C, K1 + K2, generator=g) / 32).to(DEV)
    dx1, dx2 = torch.empty(B, K1, N, device=DEV), torch.empty(B, K2, N, device=DEV)
    mx = torch.full((B,), -1, device=DEV, dtype=torch.int32)
    _lib.call("dcnet_conv1x1_bwd_data_absmax", dz.data_ptr(), W.data_ptr(), K1 + K2, dx1.data_ptr(), K1, dx2.data_ptr(), K2, B, C, N, 1,
              mx.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(mx.view(torch.float32), dx2.abs().amax(dim=(1, 2)))
    assert rel(dx2, torch.einsum('ck,bcn->bkn', W[:, K1:].double().cpu(), dz.double().cpu())) < 1e-3


@pytest.mark.parametrize("scale", [1e-7, 1e-3, 1.0, 3e4])
def test_coattention_backward_fp16_pipeline_is_scale_free(scale):
    """the fp16 pipeline multiplies everything that follows the incoming gradient by a per-problem power of two: gradients 1e-7 or 3e4
    in magnitude (far outside fp16's range unscaled), and problems of very different magnitude in one batch, give the same relative
    error as gradients of order one; equal to the tf32 pipeline's numbers to rounding."""
    g = gen(402)
    P, C, N = 2, 512, 256
    fr = torch.nn.functional.normalize(torch.randn(2 * P, C, N, generator=g).abs(), dim=1)
    go = torch.randn(2 * P, C, N, generator=g) * scale
    go[1] *= 1e-3                                                # one problem a thousand times smaller than its neighbours
    ref_in = fr.double().requires_grad_(True)
    o1, o2 = O.coattention(ref_in.view(P, 2, C, N)[:, 0], ref_in.view(P, 2, C, N)[:, 1], 10.0)
    O.interleave_pairs(o1, o2).backward(go.double())
    qa = torch.arange(2 * P, device=DEV, dtype=torch.int32)
    errs = {}
    for fp16 in (True, False):
        ops.BWD_FP16 = fp16
        try:
            x = fr.to(DEV).requires_grad_(True)
            ops.coattention(x, qa, qa ^ 1, tau=10.0, precision=2).backward(go.to(DEV))
        finally:
            ops.BWD_FP16 = True
        errs[fp16] = max(rel(x.grad[i], ref_in.grad[i]) for i in range(2 * P))      # per frame: the small problem counts on its own
    print("coattn bwd at gradient scale %.0e: worst per-frame rel err fp16 %.2e, tf32 %.2e" % (scale, errs[True], errs[False]))
    assert errs[True] < 1e-3 and errs[True] < 1.5 * errs[False] + 1e-4, errs


@pytest.mark.parametrize("N", [256, 676, 2704])
def test_fused_forward_keeps_softmax_weights_for_the_backward(N):
    """dcnet_coattn_fused_fwd_keep: E^T[z][key][q] / r[z][q] is the softmax matrix P[q][key] of the problem (to fp16 rounding of the
    weights), for every problem of the batch and at ragged tile edges (N = 676, 2704 are not multiples of the 128-key / 64-query tiles)."""
    from dcnet_b200 import _lib
    g = gen(410 + N)
    nf, C = 3, 512
    fr = torch.nn.functional.normalize(torch.randn(nf, C, N, generator=g).abs(), dim=1)
    x = fr.to(DEV)
    qa = torch.tensor([0, 2, 1], dtype=torch.int32, device=DEV)
    kb = torch.tensor([1, 0, 2], dtype=torch.int32, device=DEV)
    oidx = torch.arange(3, dtype=torch.int32, device=DEV)
    staged = ops.coattn_stage(x)
    L = _lib.lib()
    ld = (N + 7) // 8 * 8
    ek = torch.zeros(L.dcnet_coattn_keep_bytes(3, N), dtype=torch.uint8, device=DEV)
    rk = torch.empty(3, N, device=DEV)
    out = torch.empty(3, C, N, device=DEV)
    lse = torch.empty(3, N, device=DEV)
    _lib.call("dcnet_coattn_fused_fwd_keep", staged.data_ptr(), nf, qa.data_ptr(), kb.data_ptr(), oidx.data_ptr(), 3, out.data_ptr(), 3, lse.data_ptr(),
              C, N, 10.0, 0, ek.data_ptr(), rk.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ET = ek.view(torch.float16).view(3, N, ld)[:, :, :N].float()                # [z][key][q]
    for z in range(3):
        S = 10.0 * fr[int(qa[z])].double().t() @ fr[int(kb[z])].double()
        P = torch.softmax(S, 1)                                                  # [q][key]
        got = (ET[z].t() / rk[z][:, None]).double().cpu()
        assert rel(got, P) < 5e-4, (z, rel(got, P))
        assert float((got.sum(1) - 1).abs().max()) < 1e-5                       # r is the sum of exactly the stored weights
        assert float((lse[z].double().cpu() - torch.logsumexp(S, 1)).abs().max()) < 1e-3


@pytest.mark.parametrize("keep", [True, False])
@pytest.mark.parametrize("N", [256, 1024])
def test_coattention_backward_kept_weights_vs_recomputed(N, keep):
    """the two forms of the fp16 backward -- starting from the weights the forward kept, or recomputing E = exp(tau S - lse) in the
    epilogue of S = Fa^T Fb -- against the fp64 gradient"""
    g = gen(420 + N)
    P, C = 2, 512
    fr = torch.nn.functional.normalize(torch.randn(2 * P, C, N, generator=g).abs(), dim=1)
    go = torch.randn(2 * P, C, N, generator=g)
    ref_in = fr.double().requires_grad_(True)
    o1, o2 = O.coattention(ref_in.view(P, 2, C, N)[:, 0], ref_in.view(P, 2, C, N)[:, 1], 10.0)
    O.interleave_pairs(o1, o2).backward(go.double())
    qa = torch.arange(2 * P, device=DEV, dtype=torch.int32)
    ops.KEEP_E = keep
    try:
        x = fr.to(DEV).requires_grad_(True)
        ops.coattention(x, qa, qa ^ 1, tau=10.0, precision=2).backward(go.to(DEV))
    finally:
        ops.KEEP_E = True
    e = rel(x.grad, ref_in.grad)
    print("coattn bwd N=%d, weights %s: rel err %.2e" % (N, "kept by the forward" if keep else "recomputed", e))
    assert e < 1e-3, e
