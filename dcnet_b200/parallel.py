"""Multi-GPU plumbing for the hot path.

The path shards by clip / frame pair: one process per GPU, no collective in the forward data path (the reference uses plain
DDP, BatchNorm is not synchronised, train_DCNet.py:467-483).  The only exchange this module adds is the one BASELINE config 5
asks for and the reference does not have: **cross-GPU contrastive negatives**.  In the reference the rank loss pairs sample b
with sample B-1-b of the *local* batch (train_DCNet.py:195-196, :625): its text vector provides neg_sim and its GT cell
provides the second hinge term.  With global-batch semantics the partner of global sample g is Bg-1-g, which generally lives
on another rank, so every rank needs the partner's text vector flang_attn[512] and target cell (best_n, gi, gj): one NCCL
all-gather of [B_loc,512] floats (256 KB per rank at B_loc=128) and one of [3,B_loc] int64 per step.  The loss computed that
way on W ranks equals the reference loss functions applied to the concatenated batch (tests/test_parallel_gloo.py)."""
import torch
import torch.distributed as dist


class _AllGatherCat(torch.autograd.Function):
    """cat over ranks along dim 0.  backward: sum of every rank's gradient for the local slice (all-reduce + slice, which
    both NCCL and gloo provide).  Gradients follow the DDP convention: each rank back-propagates its *local* loss."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        ctx.rank = dist.get_rank(group)
        ctx.n = x.shape[0]
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, 0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, group=ctx.group)
        return g[ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n], None


def all_gather_cat(x, group=None):
    """[n, ...] on every rank -> [W*n, ...] (rank-major), differentiable."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _AllGatherCat.apply(x, group)


def global_partners(fa, best_n, gi, gj, group=None):
    """fa [B_loc,C] (requires grad), best_n/gi/gj [B_loc] int64 of the local samples.
    Returns (fa_neg [B_loc,C], partner3 [3,B_loc] int64): text vector and target cell of each local sample's partner
    Bg-1-g in the GLOBAL batch (g = rank*B_loc + b).  With one rank this is the reference's local reversal."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    B = fa.shape[0]
    fa_all = all_gather_cat(fa, group)
    cells = torch.stack([best_n, gi, gj], 1)                      # [B_loc,3]
    if world > 1:
        parts = [torch.empty_like(cells) for _ in range(world)]
        dist.all_gather(parts, cells.contiguous(), group=group)
        cells_all = torch.cat(parts, 0)
    else:
        cells_all = cells
    Bg = world * B
    partner = Bg - 1 - (rank * B + torch.arange(B, device=fa.device))
    return fa_all.index_select(0, partner), cells_all.index_select(0, partner).t().contiguous()


def allreduce_mean_(tensors, group=None):
    """data-parallel gradient averaging of the hot-path parameters in one flat collective (what DDP does in buckets)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
