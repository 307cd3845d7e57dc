"""CPU-side checks (-m "not gpu"): C-ABI library loads and exports every declared symbol, host RNG emulation is exact,
the model mirror has the reference's parameter names/shapes, and the product path refuses to run without CUDA."""
import json
import os
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

from dcnet_b200 import _lib, ops
from dcnet_b200.model.DCNet_model import grounding_model
from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class StubBackbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.maps = None

    def forward(self, x):
        return list(self.maps)


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 40
    L = _lib.lib()
    for name in protos:
        assert hasattr(L, name), name
    assert L.dcnet_abi_version() == _lib.ABI_VERSION
    assert isinstance(_lib.launch_count(), int)


def test_argument_errors_are_reported_not_thrown_across_abi():
    L = _lib.lib()
    rc = L.dcnet_coord_map(None, 0, 0, None)
    assert rc < 0 and "coord_map" in _lib.last_error()
    with pytest.raises(RuntimeError, match="coord_map"):
        _lib.call("dcnet_coord_map", None, 0, 0, None)


@pytest.mark.parametrize("P,K,N0,n", [(3, 30, 64, 10), (2, 30, 169, 10), (1, 30, 12, 10)])
def test_pyrandom_interframe_matches_cpython(P, K, N0, n):
    random.seed(42)
    got = ops.pyrandom_interframe(P, K, N0, n)
    after = random.random()
    random.seed(42)
    ref = [random.sample(list(range(N0 - 1)), n) for _ in range(P * K)]
    assert (np.array(ref).reshape(P, K, n) == got).all()
    assert after == random.random()          # stream left exactly where the reference would leave it


@pytest.mark.parametrize("B,N0,n", [(4, 64, 5), (6, 169, 5), (2, 16, 5), (16, 64, 5), (3, 65, 5), (1, 64, 0)])
def test_pyrandom_crossmodal_matches_cpython(B, N0, n):
    random.seed(7)
    got = ops.pyrandom_crossmodal(B, N0, n)
    after = random.random()
    random.seed(7)
    ref = []
    for ii in range(B):
        for jj in range(N0):
            for index in range(B):
                pool = list(range(N0))
                if index == ii:
                    pool.remove(jj)
                last = random.sample(pool, n)
            ref.append(last)
    assert (np.array(ref).reshape(B, N0, n) == got).all()
    assert after == random.random()


def test_state_dict_names_and_shapes_match_reference_fixture():
    net = grounding_model(corpus=list(range(1000)), emb_size=512, visumodel=StubBackbone())
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_shapes.json")))
    assert mine == ref


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")
def test_same_seed_gives_reference_weights_and_text_branch():
    """The mirror constructs its modules in the reference's order, so the same seed reproduces the reference's
    random init; with equal weights the stays-PyTorch neighbours (text encoder, head, location branch) restated in
    the mirror must reproduce the reference forward when driven through the oracle's restated blocks."""
    from dcnet_b200 import synth
    from oracle import dcnet_oracle as O
    T, M, MT = ref_loader.load(256)
    synth.seed_all(13)
    ref = M.grounding_model(corpus=list(range(1000)), emb_size=512, coordmap=True)
    synth.seed_all(13)
    mine = grounding_model(corpus=list(range(1000)), emb_size=512, visumodel=StubBackbone())
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    g = torch.Generator().manual_seed(5)
    maps = synth.make_raw_fvisu(2, 256, g)
    wid = synth.make_words(2, gen=g)
    ref.train(); mine.train()
    ref.visumodel.set_maps(maps)
    random.seed(3); torch.manual_seed(4)
    r = ref(torch.zeros(4, 1, 1, 1), wid, torch.zeros_like(wid))
    random.seed(3); torch.manual_seed(4)
    o = O.forward_restated(mine, maps, wid)
    for i, n in enumerate(['outbox', 'sim_score', 'loc_score', 'corr_feat']):
        for s in range(3):
            torch.testing.assert_close(o[n][s], r[i][s], rtol=2e-5, atol=2e-4 if n == 'loc_score' else 2e-5)


def test_product_path_refuses_cpu_tensors():
    net = grounding_model(corpus=list(range(1000)), emb_size=512, visumodel=StubBackbone())
    net.visumodel.maps = [torch.zeros(2, c, g, g) for c, g in ((1024, 8), (512, 16), (256, 32))]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(2, 1, 1, 1), torch.ones(2, 20, dtype=torch.long), None)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ops.rownorm(torch.zeros(4, 8))


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference`: the reference's CPU path (oracle port) on a bounded sample; exactly one JSON line on stdout with
    the keys the driver reads (impl, metric, unit, value, cpu_baseline, e2e with zero transfer bytes)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frame_pairs_per_sec" and d["unit"] == "frame-pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_packed_set_views_alias_one_buffer():
    """bench.py's input sets: tensors of mixed dtype as aligned views of one byte buffer; a copy of the buffer moves them all,
    and a view can be a gradient-requiring leaf whose storage is refreshed through the buffer."""
    from dcnet_b200.synth import PackedSet
    like = [torch.randn(2, 3, 5), torch.arange(7, dtype=torch.int64), torch.arange(9, dtype=torch.int32), torch.randn(4)]
    a, b = PackedSet(like).fill(like), PackedSet(like)
    assert a.nbytes % 256 == 0 and all(o % 256 == 0 for o in a.offsets)
    assert all(v.data_ptr() % 16 == 0 and v.is_contiguous() for v in a.views)
    b.buf.copy_(a.buf)
    for v, t in zip(b.views, like):
        assert v.dtype == t.dtype and v.shape == t.shape and torch.equal(v, t)
    # a subset span: only the two integer tensors
    c = PackedSet(like)
    c.span(1, 3).copy_(a.span(1, 3))
    assert torch.equal(c.views[1], like[1]) and torch.equal(c.views[2], like[2]) and float(c.views[0].abs().sum()) == 0
    assert c.span(3, 4).numel() == 256 and c.span(0, 4).numel() == c.nbytes
    # leaf views: refreshed through the buffer under no_grad, differentiated as usual
    x = b.views[0].requires_grad_(True)
    assert x.is_leaf
    (x * x).sum().backward()
    assert torch.allclose(x.grad, 2 * like[0])
    with torch.no_grad():
        b.buf.zero_()
    x.grad = None
    (x + 1).sum().backward()
    assert float(x.detach().abs().sum()) == 0 and torch.equal(x.grad, torch.ones_like(x))


def test_location_rank8_algebra_matches_materialised_relation():
    """The identity dcnet_loc_rank8_fwd is built on (model/DCNet_model.py:556-603): Linear(bmm(E, E^T) * obj) = E (E^T diag(obj) W^T),
    checked in fp64 against the materialised form, through BN (eval), ReLU, channel normalisation, the phrase dot product and min-max."""
    g = torch.Generator().manual_seed(5)
    B, SN, C = 3, 85, 32
    E = torch.nn.functional.normalize(torch.rand(SN, 8, generator=g, dtype=torch.float64), dim=1)
    obj = torch.nn.functional.normalize(torch.rand(B, SN, generator=g, dtype=torch.float64), dim=1)
    W = torch.randn(C, SN, generator=g, dtype=torch.float64) / SN ** 0.5
    bias = torch.randn(C, generator=g, dtype=torch.float64)
    scale, shift = torch.rand(C, generator=g, dtype=torch.float64) + 0.5, torch.randn(C, generator=g, dtype=torch.float64) * 0.1
    f = torch.nn.functional.normalize(torch.randn(B, C, generator=g, dtype=torch.float64), dim=1)

    def tail(z):                                            # z [B,SN,C]
        y = torch.relu(z * scale + shift).permute(0, 2, 1)
        m = (torch.nn.functional.normalize(y, dim=1) * f[:, :, None]).sum(1)
        mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
        return (m - mn) / (mx - mn + 1e-6)

    rel = torch.bmm(E[None].expand(B, -1, -1), E.t()[None].expand(B, -1, -1)) * obj[:, None, :]
    ref = tail(rel @ W.t() + bias)
    G = torch.einsum('cq,bq,qk->bck', W, obj, E)
    got = tail(torch.einsum('bck,pk->bpc', G, E) + bias)
    assert float((ref - got).abs().max()) < 1e-12


def test_location_rank8_batch_statistics_from_moments():
    """Next step of SURVEY 8f rank 2 (training): BatchNorm1d batch statistics of z[b,p,c] = G[b,c,:].e_p + bias[c] over all (b,p)
    follow from G and the first / second moments of E (8 and 8x8 numbers) -- no [B,SN,C] pass.  fp64 check of that algebra."""
    g = torch.Generator().manual_seed(6)
    B, SN, C = 4, 77, 16
    E = torch.nn.functional.normalize(torch.rand(SN, 8, generator=g, dtype=torch.float64), dim=1)
    G = torch.randn(B, C, 8, generator=g, dtype=torch.float64)
    bias = torch.randn(C, generator=g, dtype=torch.float64)
    z = torch.einsum('bck,pk->bpc', G, E) + bias
    mean_ref, var_ref = z.reshape(-1, C).mean(0), z.reshape(-1, C).var(0, unbiased=False)
    e1 = E.mean(0)                                          # [8]
    e2 = E.t() @ E / SN                                     # [8,8]
    mean = bias + (G @ e1).mean(0)
    second = torch.einsum('bck,kl,bcl->c', G, e2, G) / B + 2 * bias * (G @ e1).mean(0) + bias * bias
    var = second - mean * mean
    assert float((mean - mean_ref).abs().max()) < 1e-12
    assert float((var - var_ref).abs().max()) < 1e-11


def test_location_branch_rank8_training_form():
    """grounding_model.location_branch with rank8_location_train: same scores, same parameter / input gradients and the same
    running-statistics updates as the materialised form of model/DCNet_model.py:556-603 (train mode, fp64, CPU)."""
    import copy
    import torch.nn as nn
    from dcnet_b200.model.DCNet_model import grounding_model
    torch.manual_seed(3)
    size = 96                                               # 3*3 + 6*6 + 12*12 = 189 positions
    net = grounding_model(corpus=list(range(50)), emb_size=512, visumodel=nn.Identity(), size=size).double().train()
    grids = [size // 32, size // 16, size // 8]
    B, T = 3, 7
    g = torch.Generator().manual_seed(4)
    coords = [torch.rand(8, n * n, generator=g, dtype=torch.float64) for n in grids]
    context = torch.randn(B, T, 1024, generator=g, dtype=torch.float64)
    embedded = torch.randn(B, T, net.loc_text_embedding[0].out_features, generator=g, dtype=torch.float64)
    word_id = torch.randint(1, 50, (B, T), generator=g)
    word_id[1, 5:] = 0
    outs = []
    for flag in (False, True):
        m = copy.deepcopy(net)
        m.rank8_location_train = flag
        obj = [torch.rand(B, n * n, generator=torch.Generator().manual_seed(8 + n), dtype=torch.float64).requires_grad_() for n in grids]
        ctx = context.clone().requires_grad_()
        score = m.location_branch(coords, obj, ctx, embedded, word_id)
        (score * torch.linspace(0.5, 1.5, score.shape[1], dtype=torch.float64)).sum().backward()
        grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        stats = {n: b.clone() for n, b in m.named_buffers() if n.startswith(("loc_embedding", "loc_text_embedding"))}
        outs.append((score.detach(), grads, [o.grad for o in obj], ctx.grad, stats))
    (s0, g0, o0, c0, st0), (s1, g1, o1, c1, st1) = outs
    assert s0.shape == (B, sum(n * n for n in grids))
    assert float((s0 - s1).abs().max()) < 1e-10
    assert set(g0) == set(g1) and any(k.startswith("loc_text_embedding.0") for k in g0)
    for k in g0:
        assert float((g0[k] - g1[k]).abs().max()) < 1e-9 * max(1.0, float(g0[k].abs().max())), k
    for a, b in zip(o0, o1):
        assert float((a - b).abs().max()) < 1e-9 * max(1.0, float(a.abs().max()))
    assert float((c0 - c1).abs().max()) < 1e-9 * max(1.0, float(c0.abs().max()))
    for k in st0:
        assert float((st0[k].double() - st1[k].double()).abs().max()) < 1e-10, k


def test_pyrandom_crossmodal_block_sampler_randomised():
    """the block sampler against CPython from arbitrary stream positions (not only a fresh seed): sizes around the set-path /
    pool-path and bit-length boundaries, pre-advanced states, both compaction paths' shared bookkeeping (prefix counts, exact
    raw positions after the own-image sample, block refills inside a sample)."""
    rs = np.random.RandomState(11)
    for trial in range(40):
        B = int(rs.randint(1, 9))
        N0 = int(rs.choice([22, 23, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 169, 255, 256, 257, 676]))
        seed, adv = int(rs.randint(0, 10 ** 6)), int(rs.randint(0, 700))
        random.seed(seed)
        for _ in range(adv):
            random.getrandbits(32)
        state = random.getstate()
        got = ops.pyrandom_crossmodal(B, N0, 5)
        after = random.getstate()
        random.setstate(state)
        ref = np.empty((B, N0, 5), dtype=np.int64)
        for ii in range(B):
            for jj in range(N0):
                for index in range(B):
                    pool = list(range(N0))
                    if index == ii:
                        pool.remove(jj)
                    last = random.sample(pool, 5)
                ref[ii, jj] = last
        assert (ref == got).all(), (trial, B, N0, seed, adv)
        assert random.getstate() == after, (trial, B, N0, seed, adv)
