"""Seeded synthetic inputs for the hot path.  Shared by tests, smoke() and bench.py.
Pure torch-CPU generators: the same seed gives the same tensors here and on the GPU box (same image)."""
import random

import numpy as np
import torch

CH = (1024, 512, 256)


def seed_all(seed=13):
    random.seed(seed)
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 2)


def grids(size):
    return [size // 32, size // 16, size // 8]


def make_raw_fvisu(pairs, size, gen=None, corr=0.9):
    """3 x [2P, C_s, g_s, g_s]; frame 2 = corr*frame 1 + sqrt(1-corr^2)*noise (VID-like temporal correlation)."""
    out = []
    for c, g in zip(CH, grids(size)):
        a = torch.randn(pairs, c, g, g, generator=gen)
        b = corr * a + (1 - corr * corr) ** 0.5 * torch.randn(pairs, c, g, g, generator=gen)
        out.append(torch.stack([a, b], 1).reshape(2 * pairs, c, g, g).contiguous())
    return out


def make_words(pairs, vocab=1000, T=20, gen=None):
    """[2P,T] int64; lengths U{5..20}, zero padded, first pair full length; both frames share the phrase."""
    ids = torch.randint(1, vocab, (pairs, T), generator=gen)
    lens = torch.randint(5, T + 1, (pairs,), generator=gen)
    lens[0] = T
    mask = torch.arange(T)[None, :] < lens[:, None]
    ids = ids * mask
    return ids.repeat_interleave(2, 0).contiguous()


def make_boxes(pairs, size, gen=None):
    """[2P,4] xyxy fp32 clamped to [0,size-1]; frame-2 box = frame-1 box + U(-4,4)."""
    xy = torch.rand(pairs, 2, generator=gen) * size / 2
    wh = size / 8 + torch.rand(pairs, 2, generator=gen) * (size / 2 - size / 8)
    b1 = torch.cat([xy, xy + wh], 1)
    b2 = b1 + (torch.rand(pairs, 4, generator=gen) * 8 - 4)
    return torch.stack([b1, b2], 1).reshape(2 * pairs, 4).clamp(0, size - 1).contiguous()


def make_hotpath_batch(pairs, size, gen=None, T=20, C=512):
    """Synthetic tensors for dcnet_b200.hotpath.HotPath.step (all CPU fp32; see that module for the meaning)."""
    B = 2 * pairs
    gs = grids(size)
    raw = make_raw_fvisu(pairs, size, gen)
    words = make_words(pairs, T=T, gen=gen)
    lens = (words != 0).sum(1)
    unit = lambda t: t / t.norm(dim=1, keepdim=True)
    flang = unit(torch.randn(pairs, C, generator=gen).abs()).repeat_interleave(2, 0).contiguous()
    fa = unit(torch.randn(pairs, C, generator=gen).abs()).repeat_interleave(2, 0).contiguous()
    context = torch.randn(pairs, T, 2 * C, generator=gen).repeat_interleave(2, 0)
    context = (context * (torch.arange(T)[None, :] < lens[:, None])[:, :, None]).contiguous()    # pad_packed_sequence zeros
    head = [torch.randn(B, 15, g * g, generator=gen) * 0.5 for g in gs]
    loc = [torch.rand(B, g * g, generator=gen) for g in gs]
    dy_head = [torch.randn(B, C, g * g, generator=gen) * 1e-3 for g in gs]
    return dict(raw=raw, flang=flang, fa=fa, context=context, head=head, loc=loc, dy_head=dy_head, bbox=make_boxes(pairs, size, gen))


class PackedSet:
    """One step's input tensors laid out back to back in ONE byte buffer (segments aligned to 256 B, so every view keeps the
    16-byte alignment the TMA descriptors need): a whole input set moves host->device or device->device with a single copy
    instead of one launch per tensor.  `views[i]` has the shape / dtype of `like[i]`; `span(i0, i1)` is the byte range that
    holds views i0..i1-1 (for copying a subset, e.g. the maps before the late-drawn indices)."""
    ALIGN = 256

    def __init__(self, like, device=None, pin=False):
        self.offsets, n = [], 0
        for t in like:
            self.offsets.append(n)
            n += (t.numel() * t.element_size() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.nbytes = n
        self.buf = torch.zeros(n, dtype=torch.uint8, device=device, pin_memory=pin)
        self.views = [self.buf[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for o, t in zip(self.offsets, like)]

    def span(self, i0, i1):
        lo = self.offsets[i0]
        hi = self.offsets[i1] if i1 < len(self.offsets) else self.nbytes
        return self.buf[lo:hi]

    def fill(self, tensors):
        for v, t in zip(self.views, tensors):
            v.copy_(t)
        return self
