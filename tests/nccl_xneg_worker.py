"""Worker of tests/test_gpu_nccl.py (launched by torch.distributed.run, one rank per GPU, NCCL): one HotPath step with
cross-GPU contrastive negatives on this rank's slice of a seeded global batch; losses and gradients are saved for the parent."""
import os
import random
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcnet_b200 import synth                 # noqa: E402
from dcnet_b200.hotpath import HotPath       # noqa: E402

SIZE, PAIRS_PER_RANK = 256, 2


def global_batch(world):
    g = torch.Generator().manual_seed(4711)
    return synth.make_hotpath_batch(PAIRS_PER_RANK * world, SIZE, g)


def rank_slice(batch, rank):
    Bl = 2 * PAIRS_PER_RANK
    sl = slice(rank * Bl, (rank + 1) * Bl)
    return {k: ([t[sl] for t in v] if isinstance(v, list) else v[sl]) for k, v in batch.items()}


def main():
    outdir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    synth.seed_all(13)
    hp = HotPath(SIZE, cross_gpu_negatives=True).to(dev).train()
    b = rank_slice(global_batch(world), rank)
    mk = lambda t: t.clone().to(dev).requires_grad_(True)
    c = dict(raw=[mk(t) for t in b['raw']], flang=mk(b['flang']), fa=mk(b['fa']), context=mk(b['context']), head=[mk(t) for t in b['head']],
             loc=[mk(t) for t in b['loc']])
    random.seed(50 + rank)
    out = hp.step(c['raw'], c['flang'], c['fa'], c['context'], c['head'], c['loc'], [t.to(dev) for t in b['dy_head']], b['bbox'].to(dev))
    torch.cuda.synchronize()
    save = dict(out=out.cpu(), fa=c['fa'].grad.cpu(), flang=c['flang'].grad.cpu(), context=c['context'].grad.cpu(),
                raw=[t.grad.cpu() for t in c['raw']], head=[t.grad.cpu() for t in c['head']], loc=[t.grad.cpu() for t in c['loc']],
                params={k: p.grad.cpu() for k, p in hp.net.named_parameters() if p.grad is not None})
    torch.save(save, os.path.join(outdir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
