#!/usr/bin/env python
"""bench.py -- frame-pairs/sec through the DCNet dense-correspondence hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c4] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)

A step is one forward+backward pass of the hot path (a2-a18 of SURVEY.md section 8: visual mapping, inter-frame top-k
sampling, co-attention, corr_conv, pixel-to-text, fusion, cross-modal block, targets, five losses, decode + IoU, and the
backward of all of it) over one batch of synthetic VID-shaped frame pairs; Darknet, the text encoder, the 3x3 head and the
location branch are represented by synthetic tensors (dcnet_b200/hotpath.py).

Default workload = C3, the largest single-GPU configuration of BASELINE.json (configs[2]: 16 frame-pairs at 416x416 per GPU);
the default run also carries C2 (configs[1], 8 frame-pairs at 256x256) under "extra", and on N > 1 GPUs the per-GPU share of
configs[4] (64 frame-pairs at 416x416) with NCCL all-gathered cross-GPU negatives.  Prints ONE JSON line (rank 0).

  value : device-timed (CUDA events around each step, inputs resident in HBM, L2 flushed between steps), whole job.
  e2e   : the same metric through HotPath.step with HOST buffers: host RNG draw + pinned H2D of every input + step +
          D2H of the losses/IoU inside the (wall-clock, synchronised) timed region.
  --impl reference : the CPU restatement of the reference's algorithm for this path (oracle/dcnet_oracle.py; the
          reference itself is pure PyTorch and is not present on the GPU box) on the host cores, on a FIXED sample of the
          same workload, same `config`.
"""
import argparse
import gc
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "c2": dict(pairs=8, size=256, name="C2: 8 frame-pairs 256x256 per GPU (32x32 finest map), fwd+bwd of correspondence/fusion/decode path"),
    "c3": dict(pairs=16, size=416, name="C3: 16 frame-pairs 416x416 per GPU (52x52 finest map, 2704x2704 similarity), fwd+bwd"),
    "c5": dict(pairs=64, size=416, name="C5: 64 frame-pairs 416x416 per GPU, fwd+bwd (BASELINE configs[4]; add --xgpu-negatives on N > 1 GPUs for the "
               "NCCL all-gathered cross-GPU contrastive negatives)"),
    "c4": dict(pairs=224, size=256, clips=8, frames=8,
               name="C4: 8 clips x 8 frames 256x256 per GPU, all-pairs inter-frame correspondence (28 unordered = 56 directed pairs per clip), forward"),
}
# frame-pairs per step of the CPU arms (a FIXED sample of the workload, so the ratio is reproducible): the whole C2 batch; 4 of
# the 16 (64) pairs of C3 (C5), whose CPU step otherwise takes ~10 s (the python sampling loops grow with B^2 N0)
CPU_SAMPLE_PAIRS = {"c2": 8, "c3": 4, "c5": 4, "c4": 4}
C_EMB = 512


def config_of(key):
    """the `config` object: identical in both arms (`--impl ours` / `--impl reference`)"""
    wl = WORKLOADS[key]
    return dict(workload=wl["name"], pairs_per_gpu=wl["pairs"], size=wl["size"],
                l2="GPU arm: L2 flushed (256 MiB write) before every timed step, and one step's inputs exceed the 126 MB L2 at 416x416",
                sampling="exact reference random.sample stream", data="synthetic VID-shaped clips, random-init weights, seeds 13/14/15")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms: the oracle port of the reference algorithm on the host cores (bench.py may execute oracle/ only here)
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_step_fn(size, device, seed=4242):
    """-> one(pairs): seconds for one fwd+loss+bwd of the oracle port on `device` (inputs built outside the timed part)"""
    from dcnet_b200 import synth
    from dcnet_b200.hotpath import HotPath
    from oracle import dcnet_oracle as O
    synth.seed_all(13)
    net = HotPath(size).net.to(device).train()      # parameters only; every op below is the oracle's
    g = torch.Generator().manual_seed(seed)
    cuda = torch.device(device).type == "cuda"

    def one(pairs):
        b = synth.make_hotpath_batch(pairs, size, g)
        mk = lambda t: t.clone().to(device).requires_grad_(True)
        args_ = ([mk(t) for t in b['raw']], mk(b['flang']), mk(b['fa']), mk(b['context']), [mk(t) for t in b['head']],
                 [mk(t) for t in b['loc']], [t.to(device) for t in b['dy_head']], b['bbox'].to(device))
        net.zero_grad(set_to_none=True)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        O.hotpath_restated(net, *args_, size)
        if cuda:
            torch.cuda.synchronize()
        return time.perf_counter() - t0
    return one


def reference_arm(args, key, rank, world):
    """`--impl reference`: every step is the fixed CPU sample of the workload (CPU_SAMPLE_PAIRS), fwd+loss+bwd, all host threads."""
    if rank != 0:
        return
    wl = WORKLOADS[key]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size = wl["size"]
    pairs = min(CPU_SAMPLE_PAIRS[key], wl["pairs"])
    one = _oracle_step_fn(size, "cpu")
    random.seed(13)
    # the driver's --steps/--warmup are honoured up to a wall-clock bound (the contract asks for a run of a few minutes)
    t_begin = time.perf_counter()
    for _ in range(args.warmup):
        one(pairs)
        if time.perf_counter() - t_begin > 60.0:
            break
    ts = []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        ts.append(one(pairs))
        if time.perf_counter() - t_begin > 150.0:
            break
    tot = sum(ts)
    val = pairs * len(ts) / tot
    sample = "%d of %d frame-pairs per step at %dx%d (fixed sample), fwd+loss+bwd, torch CPU fp32, %d threads (oracle port of the reference " \
             "algorithm); %d timed steps" % (pairs, wl["pairs"], size, size, cores, len(ts))
    line = dict(metric="frame_pairs_per_sec", value=val, unit="frame-pairs/s", n_gpus=args.gpus, steps=len(ts), warmup=args.warmup,
                ms_per_step=1e3 * tot / len(ts), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference", config=config_of(key),
                cpu_baseline=dict(value=val, unit="frame-pairs/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=val, unit="frame-pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                note="one host, rank 0 only: this arm does not scale with --gpus")
    print(json.dumps(line), flush=True)


def cpu_baseline(key, budget_s=25.0):
    wl = WORKLOADS[key]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size = wl["size"]
    sp = min(CPU_SAMPLE_PAIRS[key], wl["pairs"])
    one = _oracle_step_fn(size, "cpu")
    random.seed(13)
    one(sp)
    ts, t_all = [], time.perf_counter()
    while len(ts) < 8 and (not ts or time.perf_counter() - t_all < budget_s):
        ts.append(one(sp))
    return dict(value=sp * len(ts) / sum(ts), unit="frame-pairs/s", cores=cores, kind="port",
                sample="%d x (%d of %d frame-pairs at %dx%d, fwd+loss+bwd) of the oracle port, torch CPU fp32, %d threads" % (
                    len(ts), sp, wl["pairs"], size, size, cores))


def pytorch_gpu_baseline(key, dev, budget_s=30.0):
    """BASELINE.md section 3 "also report": the same restated reference graph evaluated with PyTorch LIBRARY ops (cuBLAS / cuDNN /
    ATen, fp32, allow_tf32 off like the reference) on the same B200 -- the same-GPU context for the hand-written kernels.  The
    restatement is vectorised where the reference loops in python (its gathers are tensor ops here), so it is a generous
    stand-in for the reference's own GPU run; the host RNG loops are the reference's."""
    wl = WORKLOADS[key]
    size = wl["size"]
    sp = min(CPU_SAMPLE_PAIRS[key], wl["pairs"])
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        one = _oracle_step_fn(size, dev)
        random.seed(13)
        one(sp)
        ts, t_all = [], time.perf_counter()
        while len(ts) < 5 and (not ts or time.perf_counter() - t_all < budget_s):
            ts.append(one(sp))
        return dict(value=sp * len(ts) / sum(ts), unit="frame-pairs/s", ms_per_step=1e3 * sum(ts) / len(ts),
                    kind="oracle port on cuda: PyTorch library ops (cuBLAS/cuDNN/ATen), fp32, allow_tf32=False, eager, wall clock with synchronize",
                    sample="%d x (%d of %d frame-pairs at %dx%d, fwd+loss+bwd)" % (len(ts), sp, wl["pairs"], size, size))
    except Exception as e:   # a baseline leg must not take the bench line down
        return dict(unavailable="%s: %s" % (type(e).__name__, str(e)[:200]))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        gc.collect()
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------------------------
# C4: all ordered frame pairs of 8-frame clips, forward
# ----------------------------------------------------------------------------------------------------------------------
def run_c4(args, wl, rank, world, dev, dist):
    """BASELINE configs[3]: all ordered frame pairs of every 8-frame clip (generalises model/test_DCNet_model.py:303-332 from
    centre-vs-others to all pairs): visual mapping of the 64 frames once, then one co-attention problem per directed pair and
    scale; the maps of a frame are read from HBM/L2 by its 14 problems, never re-mapped.  Forward only, clips shard over GPUs."""
    from dcnet_b200 import _lib, ops, synth
    from dcnet_b200.hotpath import HotPath, all_ordered_pairs
    clips, nf, size = wl["clips"], wl["frames"], wl["size"]
    F_ = clips * nf
    synth.seed_all(13)
    hp = HotPath(size).to(dev).eval()
    g = torch.Generator().manual_seed(7000 + rank)
    host = [[t.pin_memory() for t in synth.make_raw_fvisu(F_ // 2, size, g)] for _ in range(2)]
    static = [t.to(dev) for t in host[0]]
    dev_sets = [static, [t.to(dev) for t in host[1]]]       # two resident input sets: the H2D of step i+1 lands in the other one
    qa, kb = all_ordered_pairs(clips, nf, dev)
    nprob = qa.numel()

    def run_step(maps=static):
        with torch.no_grad():
            fv = hp.net.map_visual(maps)
            outs = [ops.coattention(fv[s], qa, kb, tau=10.0, precision=hp.net.coattn_precision) for s in range(3)]
            return torch.stack([o.sum() for o in outs])

    n0 = _lib.launch_count(); res = run_step(); launches = _lib.launch_count() - n0
    for _ in range(2):
        run_step()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.zero_(); run_step()
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.zero_(); a.record(); res = run_step(); b.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # e2e: every step's 64 frames of maps cross PCIe (pinned host -> the input set the previous step is not reading, on a copy
    # stream, under the previous step's kernels); every step's result is copied back and read by the host one step late.
    main_stream, copy_stream = torch.cuda.current_stream(), torch.cuda.Stream()
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]
    h_outs = [torch.empty(3).pin_memory(), torch.empty(3).pin_memory()]

    def prefetch(i):
        k = i % 2
        copy_stream.wait_event(ev_free[k])
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(dev_sets[k], host[k]):
                dst.copy_(src, non_blocking=True)
            ev_ready[k].record(copy_stream)

    def e2e_step(i):
        k = i % 2
        prefetch(i + 1)                                    # lands in the other set while this step computes
        main_stream.wait_event(ev_ready[k])
        r = run_step(dev_sets[k])
        ev_free[k].record(main_stream)
        h_outs[k].copy_(r, non_blocking=True)
        ev_out[k].record(main_stream)
        if i > 0:
            ev_out[k ^ 1].synchronize()
            return float(h_outs[k ^ 1][0])
        return 0.0

    for k in (0, 1):
        ev_free[k].record(main_stream)
    prefetch(0)
    for i in range(args.warmup):
        e2e_step(i)
    # warm-up ends with the prefetch of step `warmup` issued: keep the parity of the timed loop aligned with it
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        e2e_step(i)
    main_stream.synchronize()                              # the last step's result has reached the host inside the timed region
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    pairs = wl["pairs"]
    if rank == 0:
        peaks = load_peaks()
        flops = sum(4.0 * C_EMB * ((size // st) ** 2) ** 2 for st in (32, 16, 8)) * nprob     # 2 GEMMs of 2 N^2 c per directed pair
        ach = flops / (dev_ms / args.steps * 1e-3) / 1e12
        h2d = sum(t_.numel() * 4 for t_ in host[0])
        cfg = config_of("c4")
        cfg.update(clips_per_gpu=clips, frames=nf, directed_pairs_per_gpu=nprob)
        line = dict(metric="frame_pairs_per_sec", value=world * pairs * args.steps / (dev_ms / 1e3), unit="frame-pairs/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="tf32/fp16", data="synthetic", config=cfg, details=dict(launch="eager"),
                    clocks=clocks,
                    e2e=dict(value=world * pairs * args.steps / (e2e_ms / 1e3), unit="frame-pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=12,
                             ms_per_step=e2e_ms / args.steps,
                             pipeline="H2D of step i+1 (pinned host -> the other resident input set, copy stream) under the kernels of step i; every step's result is copied back, the host reads it one step late"),
                    gpu_launches=int(launches * args.steps), gpu_launches_per_step=int(launches),
                    roofline=dict(bound="tensor", kernel="visual mapping + all-pairs co-attention forward (whole step)", achieved=ach,
                                  peak=peaks["tensor_sustained"], unit="TFLOP/s", frac=ach / peaks["tensor_sustained"], traffic=None,
                                  peak_source=peaks["src"] + " bf16 sustained (kernel timed inside a long step)"),
                    cpu_baseline=None)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# the training step of the hot path: value (device-timed, resident inputs) and e2e (host buffers, wall clock)
# ----------------------------------------------------------------------------------------------------------------------
def measure_hotpath(key, steps, warmup, rank, world, local, dev, dist, xneg=False, allreduce=True, use_graph=True, sample_clocks=False,
                    probe=False):
    """-> dict(value, ms_per_step, e2e, launches_per_step, clocks, ...) for workload `key` on this rank's GPU (max over ranks)."""
    from concurrent.futures import ThreadPoolExecutor
    from dcnet_b200 import _lib, synth
    from dcnet_b200.hotpath import HotPath
    wl = WORKLOADS[key]
    pairs, size = wl["pairs"], wl["size"]
    B = 2 * pairs
    synth.seed_all(13)                       # identical replicas (DDP broadcast equivalent)
    xneg = bool(xneg and world > 1)
    hp = HotPath(size, cross_gpu_negatives=xneg).to(dev).train()
    if xneg and os.environ.get("DCNET_XNEG_EAGER"):
        use_graph = False                    # bring-up switch: NCCL all-gathers outside any CUDA graph
    random.seed(1000 + rank)
    g = torch.Generator().manual_seed(9000 + rank)

    # ---- host batches (pinned) and static device buffers.  Every tensor that crosses PCIe per step (forward inputs of the path,
    # then the two index tensors) is a view of ONE packed byte buffer per set (synth.PackedSet): a set moves with one copy.
    # dy_head is the gradient the head's backward hands to the fusion output: it is produced on the device in a real step,
    # so it stays resident; every forward input of the path (maps, text vectors, head/location outputs, boxes, indices) is copied.
    NB = 2
    H2D_KEYS = ('raw', 'flang', 'fa', 'context', 'head', 'loc', 'bbox')
    grad_keys = ('raw', 'flang', 'fa', 'context', 'head', 'loc')
    batches = []
    for _ in range(NB):
        b = synth.make_hotpath_batch(pairs, size, g)
        batches.append(dict(raw=b['raw'], flang=[b['flang']], fa=[b['fa']], context=[b['context']], head=b['head'], loc=b['loc'],
                            dy_head=b['dy_head'], bbox=[b['bbox']]))
    np_, ni_ = hp.draw_indices(B)
    idx0 = [torch.from_numpy(np_), torch.from_numpy(ni_)]
    slots = [(k, j) for k in H2D_KEYS for j in range(len(batches[0][k]))]
    n_maps = len(slots)                                     # views [0, n_maps) = forward inputs, [n_maps, n_maps+2) = negpos, negidx
    like = [batches[0][k][j] for k, j in slots] + idx0
    host_sets = [synth.PackedSet(like, pin=True).fill([bt[k][j] for k, j in slots] + idx0) for bt in batches]
    static_set = synth.PackedSet(like, device=dev)
    stage_set = synth.PackedSet(like, device=dev)
    static_set.buf.copy_(host_sets[0].buf); stage_set.buf.copy_(host_sets[0].buf)
    static = {k: [] for k in H2D_KEYS}
    for (k, j), v in zip(slots, static_set.views):
        static[k].append(v.requires_grad_(k in grad_keys))
    static['dy_head'] = [t.to(dev) for t in batches[0]['dy_head']]
    s_negpos, s_negidx = static_set.views[n_maps], static_set.views[n_maps + 1]
    h2d_bytes = static_set.nbytes                          # the bytes of the two copies per step (maps span + index span)
    h_out = torch.empty(6 + B, dtype=torch.float32).pin_memory()
    d2h_bytes = h_out.numel() * 4
    del batches

    # Data-parallel gradient all-reduce of the hot-path parameters.  DCNET_AR_BUCKETS=1 (default 0 = one flat all-reduce after the
    # backward): one bucket per pyramid scale, issued on a communication stream from post-accumulate-grad hooks as soon as every
    # parameter of that scale has its gradient -- the coarse scales finish their backward ~40 % before the end of the step.
    ar_buckets = world > 1 and allreduce and os.environ.get("DCNET_AR_BUCKETS", "0") == "1"
    ar_state = dict(pending={}, comm=None)
    if ar_buckets:
        import re
        ar_state["comm"] = torch.cuda.Stream()
        hot = {id(p) for p in hp.hot_parameters}
        groups = {0: [], 1: [], 2: []}
        for n_, p_ in hp.net.named_parameters():
            if id(p_) in hot:
                groups[int(re.search(r"\.(\d)(\.|$)", n_).group(1))].append(p_)

        def make_hook(sc):
            def hook(param):
                comm = ar_state["comm"]
                comm.wait_event(torch.cuda.current_stream().record_event())      # this parameter's gradient is complete on its stream
                ar_state["pending"][sc] -= 1
                if ar_state["pending"][sc] == 0:
                    with torch.cuda.stream(comm):
                        flat = torch.cat([q.grad.reshape(-1) for q in groups[sc] if q.grad is not None])
                        dist.all_reduce(flat)
            return hook
        for sc, ps in groups.items():
            for p_ in ps:
                p_.register_post_accumulate_grad_hook(make_hook(sc))
        ar_state["groups"] = groups

    def run_step():
        if ar_buckets:
            for sc, ps in ar_state["groups"].items():
                ar_state["pending"][sc] = len(ps)
        r_ = hp.step(static['raw'], static['flang'][0], static['fa'][0], static['context'][0], static['head'], static['loc'],
                     static['dy_head'], static['bbox'][0], s_negpos, s_negidx)
        if ar_buckets:
            torch.cuda.current_stream().wait_stream(ar_state["comm"])
        return r_

    def clear_grads():
        for p in hp.parameters():
            p.grad = None
        for k in grad_keys:
            for t in static[k]:
                t.grad = None

    # ---- eager warm-up (also counts this library's launches per step)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            clear_grads()
            n0 = _lib.launch_count()
            res = run_step()
            launches_per_step = _lib.launch_count() - n0
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = None
    ar_in_graph = False
    do_ar = world > 1 and allreduce
    ar_in_graph = bool(ar_buckets and use_graph)
    if do_ar:
        dist.all_reduce(torch.zeros(1, device=dev))          # communicator up before any capture
        torch.cuda.synchronize()
    if use_graph:
        clear_grads()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            res = run_step()
            if do_ar and not ar_buckets:
                # data-parallel gradient all-reduce of the hot-path parameters, one flat NCCL collective, captured with the step
                flat_g = torch.cat([p.grad.reshape(-1) for p in hp.hot_parameters if p.grad is not None])
                dist.all_reduce(flat_g)
                ar_in_graph = True

    def do_step():
        nonlocal res
        if graph is not None:
            graph.replay()
        else:
            clear_grads()
            res = run_step()
        if do_ar and not ar_in_graph and not ar_buckets:
            flat = torch.cat([p.grad.reshape(-1) for p in hp.hot_parameters if p.grad is not None])
            dist.all_reduce(flat)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: device-timed, inputs resident
    for _ in range(warmup):
        flush.zero_()
        do_step()
    barrier()
    sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush.zero_()
        a.record()
        do_step()
        b.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---- e2e: host buffers, wall clock, H2D + D2H inside.  The negative indices follow the reference's sequential random.sample
    # stream, which does not depend on device results: step t+1's draw runs on a worker thread (the C emulation releases the GIL)
    # while the GPU executes step t; the draw of every timed step is inside the timed region (steady-state pipeline).
    pool = ThreadPoolExecutor(1)
    h_idx = [synth.PackedSet(idx0, pin=True) for _ in range(2)]

    def draw(slot):
        npos, nidx = hp.draw_indices(B)
        h_idx[slot].views[0].copy_(torch.from_numpy(npos)); h_idx[slot].views[1].copy_(torch.from_numpy(nidx))
        return slot

    pending = [pool.submit(draw, 0)]

    # Input pipeline: the H2D copy of step i+1 (pinned host -> a staging set on a copy stream) runs under the GPU work of step i;
    # at the start of a step the staged set moves into the graph's static buffers device-to-device (one copy).  Every step's
    # inputs still cross PCIe inside the timed region; the wall clock sees max(copy, compute) instead of their sum.
    main_stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    ev_staged, ev_consumed = torch.cuda.Event(), torch.cuda.Event()
    stage_maps, stage_idx = stage_set.span(0, n_maps), stage_set.span(n_maps, n_maps + 2)
    assert stage_idx.numel() == h_idx[0].nbytes

    def prefetch_maps(i):
        copy_stream.wait_event(ev_consumed)                # the previous contents of the staging set have been consumed
        with torch.cuda.stream(copy_stream), torch.no_grad():
            stage_maps.copy_(host_sets[i % NB].span(0, n_maps), non_blocking=True)

    def prefetch_indices():
        slot = pending.pop().result()                      # drawn on the worker thread while the GPU was busy
        pending.append(pool.submit(draw, slot ^ 1))
        with torch.cuda.stream(copy_stream), torch.no_grad():
            stage_idx.copy_(h_idx[slot].buf, non_blocking=True)
            ev_staged.record(copy_stream)

    ev_consumed.record(main_stream)
    prefetch_maps(0); prefetch_indices()
    h_outs = [torch.empty_like(h_out).pin_memory(), torch.empty_like(h_out).pin_memory()]
    ev_out = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"n": 0, "loss": float("nan")}

    def e2e_step(i):
        main_stream.wait_event(ev_staged)
        with torch.no_grad():
            static_set.buf.copy_(stage_set.buf, non_blocking=True)
        ev_consumed.record(main_stream)
        prefetch_maps(i + 1)                               # next step's maps cross PCIe under this step's kernels
        do_step()
        slot = i & 1
        h_outs[slot].copy_(res, non_blocking=True)         # every step's result goes back to the host ...
        ev_out[slot].record(main_stream)
        prefetch_indices()                                 # next step's indices (host draw finished meanwhile)
        # ... and is read one step late, so the GPU already runs step i while the host waits for the loss of step i-1
        if e2e_state["n"] > 0:
            ev_out[slot ^ 1].synchronize()
            e2e_state["loss"] = float(h_outs[slot ^ 1][0])
        e2e_state["n"] += 1
        return e2e_state["loss"]

    for i in range(warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    main_stream.synchronize()                              # the last step's result has reached the host inside the timed region
    loss_val = float(h_outs[(steps - 1) & 1][0])
    barrier()
    e2e_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    pending.pop().result()
    pool.shutdown()
    clocks = sampler.stop() if sampler else None
    if probe and graph is not None:
        # where the difference between e2e and value sits: wall clock per step of back-to-back replays with pieces of the loop
        def probe_(name, fn, n=200):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            print("probe %-28s %.4f ms/step" % (name, (time.perf_counter() - t1) * 1e3 / n), file=sys.stderr)
        probe_("replay only", graph.replay)
        probe_("replay + D2D", lambda: (static_set.buf.copy_(stage_set.buf, non_blocking=True), graph.replay()))
        probe_("replay + D2H", lambda: (graph.replay(), h_outs[0].copy_(res, non_blocking=True)))

    t = torch.tensor([dev_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    out = dict(value=world * pairs * steps / (dev_ms / 1e3), ms_per_step=dev_ms / steps, steps=steps, warmup=warmup,
               e2e=dict(value=world * pairs * steps / (e2e_ms / 1e3), unit="frame-pairs/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=d2h_bytes,
                        ms_per_step=e2e_ms / steps, last_loss=loss_val,
                        pipeline="H2D of step i+1 (pinned host -> staging set, copy stream) under the kernels of step i; staged -> static inputs "
                                 "device-to-device, each set one packed buffer = one copy; every step's loss is copied back, the host reads it one step late"),
               launches_per_step=int(launches_per_step), clocks=clocks, launch="CUDA graph replay" if graph is not None else "eager",
               grad_allreduce=bool(do_ar), allreduce_buckets=("per scale, from gradient hooks" if ar_buckets else ("one, after the backward" if do_ar else None)),
               cross_gpu_negatives=xneg, nccl_in_graph=bool(graph is not None and (xneg or ar_in_graph)))
    # a CUDA graph that holds captured NCCL kernels must be gone before the communicator is torn down (and before the next capture)
    graph = None
    res = None
    del hp, static, static_set, stage_set, host_sets, flush
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# kernel-level rooflines (rank 0): each kernel timed alone with CUDA events, operands rotating over sets larger than L2
# ----------------------------------------------------------------------------------------------------------------------
# DRAM traffic per launch of the two tensor-core kernels, from `ncu --set full` captures of the same instances
# (dram__bytes_read.sum + dram__bytes_write.sum); (size, pairs) -> (bytes, source file)
NCU_TRAFFIC_GEMM = {(416, 16): (788.396544e6 + 154.899968e6, "profiles/r3b_ncu_full_gemm_cn_f16.txt")}
NCU_TRAFFIC_GEMM_S = {(416, 16): (177.33888e6 + 881.462784e6, "profiles/r3b_ncu_full_gemm_s_f16.txt")}
NCU_TRAFFIC_COATTN = {(416, 16): (89.010432e6 + 127.560192e6, "profiles/r2n_ncu_full_coattn_fwd.txt"), (256, 8): (16.9e6, "profiles/r1w_ncu_full_coattn_fused.txt")}
NCU_TRAFFIC_HBM = {(416, 16): {"bn_act_fwd": (177.355264e6 + 136.50944e6, "profiles/r2n_ncu_full_bn_fwd.txt"),
                              "bn_act_bwd_reduce": (367.188224e6 + 145.653248e6, "profiles/r2n_ncu_full_bn_bwd.txt")}}


def kernel_rooflines(key, dev):
    from dcnet_b200 import _lib, ops
    wl = WORKLOADS[key]
    pairs, size = wl["pairs"], wl["size"]
    B = 2 * pairs
    peaks = load_peaks()
    N2 = (size // 8) ** 2
    # operands and outputs rotate over NSET buffer sets whose total size exceeds the 126 MB L2 several times, so no launch finds its
    # inputs (or the lines it will overwrite) in L2; a spin kernel ahead of each launch lets the host enqueue (event, kernel, event)
    # before the GPU gets there, so the interval holds no launch latency.
    def timed_sets(calls, reps=3):
        for c in calls:
            c()
        tot, n = 0.0, 0
        for _ in range(reps):
            for c in calls:
                torch.cuda._sleep(200000)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); c(); b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b); n += 1
        return tot / n

    NSET = 5 if size <= 256 else 3
    frs = [torch.nn.functional.normalize(torch.randn(B, C_EMB, N2, device=dev).abs(), dim=1) for _ in range(NSET)]
    # The dominant kernel family of the step is the persistent tcgen05 GEMM (cta_group::2 pairs, 256x256 per cluster); its largest
    # instances are the five contractions of the co-attention backward at the finest scale, which run on fp16 operands (kind::f16: the
    # 11 significant bits of tf32 at twice its MMA rate).  Two instance types, each timed alone on all problems of the batch (= how the
    # step launches them):
    #   roofline        : a [C,N]-output contraction (dFa += Fb dS^T: M=512, N=K=N2, reduce-add epilogue), three per scale -- the largest
    #                     share of the step
    #   roofline_gemm_s : S = Fa^T Fb (M=N=N2, K=512), the plain-epilogue sibling of the two N x N-output contractions (S / exp, dP / dS)
    nprob = B
    ld8 = (N2 + 7) // 8 * 8
    f16s = [torch.zeros(B, C_EMB, ld8, device=dev, dtype=torch.float16) for _ in range(NSET)]
    for i in range(NSET):
        f16s[i][:, :, :N2] = ops.cast_f16(frs[i])
    s16 = [torch.zeros(nprob, N2, ld8, device=dev, dtype=torch.float16) for _ in range(NSET)]
    for t in s16:
        t[:, :, :N2].normal_()
    c_bufs = [torch.empty(nprob, N2, N2, device=dev) for _ in range(NSET)]
    set_bytes_g = 2 * nprob * C_EMB * ld8 * 2 + s16[0].numel() * 2 + nprob * C_EMB * N2 * 4
    st0 = torch.cuda.current_stream().cuda_stream

    def gemm16(Aop, a_mn, lda, sA, Bop, b_mn, ldb, sB, out, ldc, sC, M, N, K, atomic):
        _lib.call("dcnet_gemm_f16", Aop.data_ptr(), a_mn, lda, sA, Bop.data_ptr(), b_mn, ldb, sB, out.data_ptr(), ldc, sC, M, N, K, nprob, 1.0, atomic, st0)

    ms_s = timed_sets([(lambda i=i: gemm16(f16s[i], 1, ld8, C_EMB * ld8, f16s[(i + 1) % NSET], 1, ld8, C_EMB * ld8, c_bufs[i], N2, N2 * N2,
                                            N2, N2, C_EMB, 0)) for i in range(NSET)])
    fl_s = 2.0 * N2 * N2 * C_EMB * nprob
    ach_s = fl_s / (ms_s * 1e-3) / 1e12
    d_bufs = [torch.zeros(nprob, C_EMB, N2, device=dev) for _ in range(NSET)]
    ms_g = timed_sets([(lambda i=i: gemm16(f16s[i], 0, ld8, C_EMB * ld8, s16[i], 0, ld8, N2 * ld8, d_bufs[i], N2, C_EMB * N2,
                                            C_EMB, N2, N2, 1)) for i in range(NSET)])
    fl_g = 2.0 * C_EMB * N2 * N2 * nprob
    ach_g = fl_g / (ms_g * 1e-3) / 1e12
    # algorithmic bytes of one launch: A and B read once, the fp32 output read and written once (reduce-add)
    alg_g = nprob * (C_EMB * N2 * 2 + N2 * N2 * 2 + 2 * C_EMB * N2 * 4)
    tr = NCU_TRAFFIC_GEMM.get((size, pairs))
    common = dict(bound="tensor", peak=peaks["tensor_burst"], unit="TFLOP/s", dtype="fp16 operands (kind::f16), fp32 accumulate",
                  l2="operands and outputs rotate over %d sets, %.0f MB in total (> 126 MB L2)" % (NSET, NSET * set_bytes_g / 1e6),
                  note="peak is the measured bf16 figure (fp16 and bf16 share the kind::f16 MMA rate)",
                  peak_source=peaks["src"] + " bf16 burst (kernel timed alone)")
    roof = dict(kernel="umma_gemm2_kernel (tcgen05 kind::f16, cta_group::2 pairs, persistent, TMA reduce-add epilogue; dFa += Fb dS^T of the "
                "co-attention backward on fp16 operands, M=%d N=K=%d, %d problems, one launch)" % (C_EMB, N2, nprob), achieved=ach_g,
                frac=ach_g / peaks["tensor_burst"], ms=ms_g, algorithmic_bytes=alg_g, traffic=tr[0] if tr else None,
                traffic_source=(tr[1] + ": dram__bytes_read.sum + dram__bytes_write.sum") if tr else None, **common)
    tr = NCU_TRAFFIC_GEMM_S.get((size, pairs))
    roof_s = dict(kernel="umma_gemm2_kernel (same kernel; S = Fa^T Fb of the co-attention backward on fp16 operands, fp32 output, M=N=%d K=%d, "
                  "%d problems, one launch)" % (N2, C_EMB, nprob),
                  achieved=ach_s, frac=ach_s / peaks["tensor_burst"], ms=ms_s,
                  traffic=tr[0] if tr else None, traffic_source=(tr[1] + ": dram__bytes_read.sum + dram__bytes_write.sum") if tr else None, **common)
    del f16s, s16
    del d_bufs
    del c_bufs
    qa = torch.arange(B, device=dev, dtype=torch.int32)
    kb = qa ^ 1
    NS2 = 8 if size <= 256 else 3
    stg = [ops.coattn_stage(frs[i % NSET]) for i in range(NS2)]
    o_bufs = [torch.empty(B, C_EMB, N2, device=dev) for _ in range(NS2)]
    l_buf = torch.empty(B, N2, device=dev)
    ms = timed_sets([(lambda i=i: ops.coattn_fused(stg[i], frs[0].shape, qa, kb, tau=10.0, out=o_bufs[i], lse=l_buf)) for i in range(NS2)])
    ms_stage = timed_sets([(lambda i=i: ops.coattn_stage(frs[i])) for i in range(NSET)])
    flops = 6.0 * C_EMB * N2 * N2 * pairs            # 6*c*N^2 per pair forward (both directions share S)
    ach = flops / (ms * 1e-3) / 1e12
    ach_incl = flops / ((ms + ms_stage) * 1e-3) / 1e12
    tr = NCU_TRAFFIC_COATTN.get((size, pairs))
    roof_co = dict(bound="tensor", kernel="coattn_fused_kernel (finest scale, N=%d, %d pairs = %d directed problems, one launch)" % (N2, pairs, B),
                   achieved=ach, peak=peaks["tensor_burst"], unit="TFLOP/s", frac=ach / peaks["tensor_burst"],
                   traffic=tr[0] if tr else None, traffic_source=(tr[1] + ": dram__bytes_read.sum + dram__bytes_write.sum") if tr else None,
                   ms=ms, stage_ms=ms_stage, achieved_incl_staging=ach_incl, frac_incl_staging=ach_incl / peaks["tensor_burst"],
                   frac_of_sustained=ach / peaks["tensor_sustained"], executed_tflops=ach * 4.0 / 3.0,
                   l2="staged operands and outputs rotate over %d sets (%.0f MB)" % (NS2, NS2 * (stg[0].numel() + o_bufs[0].numel() * 4) / 1e6),
                   note="algorithmic 6*c*N^2 per pair; the kernel executes 8*c*N^2 (S recomputed per direction); stage_ms = the stand-alone fp16 staging pass "
                        "that precedes it in the step (frac_incl_staging counts it)",
                   peak_source=peaks["src"] + " bf16 burst (kernel timed alone)")
    del stg, o_bufs
    # ---- HBM-bound kernels of the path against the measured copy bandwidth (north_star: fusion / normalise / decode kernels):
    # algorithmic bytes = every tensor element read or written once, finest scale, rotating buffer sets
    st_ = torch.cuda.current_stream().cuda_stream
    Pp = lambda t: t.data_ptr()
    cvec = [torch.rand(C_EMB, device=dev) + 0.5 for _ in range(4)]           # mean, invstd, gamma, beta stand-ins
    fa_ = torch.nn.functional.normalize(torch.rand(B, C_EMB, device=dev), dim=1)
    ys = [torch.empty_like(frs[0]) for _ in range(NSET)]
    dvs = [torch.empty_like(frs[0]) for _ in range(NSET)]
    sims = [torch.empty(B, N2, device=dev) for _ in range(2)]
    sums = torch.zeros(2, C_EMB, device=dev); dfa = torch.zeros(B, C_EMB, device=dev)
    map_bytes = frs[0].numel() * 4
    hbm = {}
    t_ = timed_sets([(lambda i=i: _lib.call("dcnet_bn_act_fwd", Pp(frs[i]), Pp(cvec[0]), Pp(cvec[1]), Pp(cvec[2]), Pp(cvec[3]), 0.0, 1, Pp(ys[i]),
                                             Pp(fa_), None, Pp(sims[0]), Pp(sims[1]), B, C_EMB, N2, st_)) for i in range(NSET)])
    hbm["bn_act_fwd_kernel (BN + ReLU + channel L2 norm + pixel-to-text dots: z read once, y written once)"] = (2 * map_bytes, t_)
    t_ = timed_sets([(lambda i=i: _lib.call("dcnet_bn_act_bwd_reduce", Pp(frs[i]), Pp(cvec[0]), Pp(cvec[1]), Pp(cvec[2]), Pp(cvec[3]), 0.0, 1,
                                             Pp(ys[i]), Pp(fa_), None, Pp(sims[0]), Pp(sims[1]), Pp(dvs[i]), Pp(sums[0]), Pp(sums[1]), Pp(dfa), None,
                                             B, C_EMB, N2, st_)) for i in range(NSET)])
    hbm["bn_act_bwd_reduce_kernel (reads z, dy; writes dv; channel sums)"] = (3 * map_bytes, t_)
    t_ = timed_sets([(lambda i=i: _lib.call("dcnet_bn_act_bwd_apply", Pp(frs[i]), Pp(cvec[0]), Pp(cvec[1]), Pp(cvec[2]), Pp(dvs[i]), Pp(sums[0]),
                                             Pp(sums[1]), 1, Pp(dvs[i]), B, C_EMB, N2, st_)) for i in range(NSET)])
    hbm["bn_act_bwd_apply_kernel (reads z, dv; writes dz)"] = (3 * map_bytes, t_)
    g2 = size // 8
    yin = [torch.randn(B, 255, g2, g2, device=dev) for _ in range(NSET)]
    anc = [(10, 13), (16, 30), (33, 23)]
    t_ = timed_sets([(lambda i=i: ops.yolo_layer_decode(yin[i], anc, 80, size)) for i in range(NSET)])
    hbm["yolo_decode_kernel (a19, [B,255,g,g] -> [B,3gg,85])"] = (2 * yin[0].numel() * 4, t_)
    # what a plain device copy of the same footprint reaches in this harness (torch copy_: read + write of one map)
    t_copy = timed_sets([(lambda i=i: ys[i].copy_(frs[i])) for i in range(NSET)])
    copy_gbs = 2 * map_bytes / (t_copy * 1e-3) / 1e9
    def ncu_hbm(k):
        for name, v in NCU_TRAFFIC_HBM.get((size, pairs), {}).items():
            if k.startswith(name + "_kernel"):
                return v
        return (None, None)
    roof_hbm = [dict(bound="hbm", kernel=k, achieved=b_ / (t * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s", frac=b_ / (t * 1e-3) / 1e9 / peaks["hbm"],
                     ms=t, bytes=b_, traffic=ncu_hbm(k)[0],
                     traffic_source=(ncu_hbm(k)[1] + ": dram__bytes_read.sum + dram__bytes_write.sum") if ncu_hbm(k)[1] else None,
                     copy_same_size_gbs=copy_gbs) for k, (b_, t) in hbm.items()]
    del frs, ys, dvs, yin
    gc.collect()
    torch.cuda.empty_cache()
    return roof, roof_s, roof_co, roof_hbm


_FULL_AFFINITY = None


def pin_to_gpu_numa_node(local):
    """The end-to-end loop streams every step's inputs from pinned host memory (165 MB per step and GPU at C3: ~25 GB/s per rank, ~200 GB/s
    on 8 GPUs).  Pinned pages are placed on the NUMA node of the thread that first touches them, so each rank binds itself to the CPU
    cores next to its GPU BEFORE it allocates anything: the copies then never cross the socket interconnect.  Best effort (NVML's ideal
    CPU affinity; falls back to sysfs); returns a short description for the JSON line, or None."""
    global _FULL_AFFINITY
    try:
        _FULL_AFFINITY = os.sched_getaffinity(0)
        before = len(_FULL_AFFINITY)
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            how = "nvmlDeviceSetCpuAffinity"
        except Exception:
            bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
            if bus is None:
                return None
            path = "/sys/bus/pci/devices/0000:%02x:00.0/numa_node" % bus
            node = int(open(path).read())
            if node < 0:
                return None
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if not cpus:
                return None
            os.sched_setaffinity(0, cpus)
            how = "sysfs numa_node %d" % node
        return "%s: %d of %d cores" % (how, len(os.sched_getaffinity(0)), before)
    except Exception:
        return None


def step_flops(key):
    """algorithmic GEMM FLOPs of one step per GPU: co-attention 18 c sum N^2 + 1x1 convs 3 x 2 K c sum N per image"""
    wl = WORKLOADS[key]
    size, pairs = wl["size"], wl["pairs"]
    Ns = [(size // s) ** 2 for s in (32, 16, 8)]
    co = 18.0 * C_EMB * sum(n * n for n in Ns)
    conv = 3.0 * 2 * 2.0 * C_EMB * sum(n * (k + 1024 + 512) for n, k in zip(Ns, (1024, 512, 256)))     # x2 images per pair
    return pairs * (co + conv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and pytorch_gpu_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra workloads (C2; the C5 share with cross-GPU negatives on N > 1)")
    ap.add_argument("--no-rooflines", action="store_true", help="skip the kernel-level roofline timings")
    ap.add_argument("--no-allreduce", action="store_true", help="N>1: skip the data-parallel gradient all-reduce of the hot-path parameters")
    ap.add_argument("--xgpu-negatives", action="store_true",
                    help="N>1: BASELINE config 5 -- rank-loss / pixel-to-text negatives from the global batch (NCCL all-gather of text vectors "
                         "and target cells inside the step, captured into the step's CUDA graph with everything else)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    key = args.workload
    wl = WORKLOADS[key]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, key, rank, world)

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own "NCCL version ..." banner (printed to fd 1 when the communicator comes
        # up) is sent to stderr by pointing fd 1 at fd 2 until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if key == "c4":
        return run_c4(args, wl, rank, world, dev, dist)

    m = measure_hotpath(key, args.steps, args.warmup, rank, world, local, dev, dist, xneg=args.xgpu_negatives, allreduce=not args.no_allreduce,
                        use_graph=not args.no_graph, sample_clocks=True, probe=bool(os.environ.get("DCNET_E2E_PROBE")))
    nccl_in_graph = m["nccl_in_graph"]
    extra = {}
    if key == "c3" and not args.no_extras:
        # the other single-GPU config of BASELINE.json, same method, shorter run
        e = measure_hotpath("c2", min(args.steps, 20), 3, rank, world, local, dev, dist, allreduce=not args.no_allreduce, use_graph=not args.no_graph)
        nccl_in_graph |= e["nccl_in_graph"]
        extra["c2"] = dict(config=config_of("c2"), value=e["value"], unit="frame-pairs/s", ms_per_step=e["ms_per_step"], steps=e["steps"],
                           e2e=dict(value=e["e2e"]["value"], ms_per_step=e["e2e"]["ms_per_step"], h2d_bytes_per_step=e["e2e"]["h2d_bytes_per_step"]),
                           gpu_launches_per_step=e["launches_per_step"])
        if world > 1:
            # BASELINE configs[4]: 64 frame-pairs at 416x416 per GPU with NCCL all-gathered cross-GPU contrastive negatives
            e = measure_hotpath("c5", min(args.steps, 5), 3, rank, world, local, dev, dist, xneg=True, allreduce=not args.no_allreduce,
                                use_graph=not args.no_graph)
            nccl_in_graph |= e["nccl_in_graph"]
            cfg5 = config_of("c5"); cfg5.update(cross_gpu_negatives=True)
            extra["c5_xgpu_negatives"] = dict(config=cfg5, value=e["value"], unit="frame-pairs/s", ms_per_step=e["ms_per_step"], steps=e["steps"],
                                              n_gpus=world, e2e=dict(value=e["e2e"]["value"], ms_per_step=e["e2e"]["ms_per_step"],
                                                                     h2d_bytes_per_step=e["e2e"]["h2d_bytes_per_step"]),
                                              grad_allreduce=e["grad_allreduce"], cross_gpu_negatives=e["cross_gpu_negatives"])

    roof = roof_s = roof_co = roof_hbm = cpu_base = gpu_base = None
    if rank == 0:
        if not args.no_rooflines:
            roof, roof_s, roof_co, roof_hbm = kernel_rooflines(key, dev)
        if world == 1 and not args.no_cpu_baseline:
            gpu_base = pytorch_gpu_baseline(key, dev)
            if _FULL_AFFINITY:
                os.sched_setaffinity(0, _FULL_AFFINITY)      # the CPU arm uses every core the process was given, not one NUMA node
            cpu_base = cpu_baseline(key)
        peaks = load_peaks()
        step_tf = step_flops(key) / (m["ms_per_step"] * 1e-3) / 1e12
        line = dict(metric="frame_pairs_per_sec", value=m["value"], unit="frame-pairs/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=m["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="tf32/fp16", data="synthetic",
                    config=config_of(key),
                    details=dict(launch=m["launch"],
                                 arithmetic="fp32 tensors in HBM; 1x1-conv contractions on tcgen05 as tf32 x tf32 -> fp32 on operands rounded to the "
                                            "nearest tf32 by their producers; co-attention forward and backward on fp16 x fp16 -> fp32 (kind::f16: the "
                                            "11 significant bits of tf32, gradient-side operands scaled per problem into fp16's range); "
                                            "index-producing contractions and everything else in fp32",
                                 grad_allreduce=m["grad_allreduce"], allreduce_buckets=m.get("allreduce_buckets"),
                                 cross_gpu_negatives=m["cross_gpu_negatives"],
                                 step_gemm_tflops=step_tf, step_gemm_frac_of_sustained=step_tf / peaks["tensor_sustained"]),
                    clocks=m["clocks"], e2e=m["e2e"], host_affinity=numa,
                    gpu_launches=int(m["launches_per_step"] * args.steps), gpu_launches_per_step=m["launches_per_step"],
                    roofline=roof, roofline_gemm_s=roof_s, roofline_coattn=roof_co, roofline_hbm=roof_hbm, cpu_baseline=cpu_base, pytorch_gpu_baseline=gpu_base,
                    extra=extra or None)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if nccl_in_graph:
            sys.stdout.flush()
            os._exit(0)              # ProcessGroupNCCL teardown after captured collectives can block; nothing is left to flush
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
