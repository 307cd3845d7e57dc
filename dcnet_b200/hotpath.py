"""The hot path on its own: everything between the Darknet feature maps and the scalar loss that
this package replaces (a2-a18), with the stays-PyTorch neighbours -- Darknet, text encoder, the 3x3 head
(fcn_emb[s][1:], fcn_out) and the location branch -- represented by tensors supplied by the caller:

    raw[s]      [B,C_s,h_s,w_s]   Darknet maps                       (input, needs grad -> backbone)
    flang, fa   [B,512]           sentence / attended phrase vectors  (inputs, need grad -> text encoder)
    context     [B,T,1024]        BiLSTM outputs
    head[s]     [B,15,N_s]        what fcn_out(fcn_emb[1:](y)) returns (input, needs grad -> head)
    loc[s]      [B,N_s]           location scores
    dy_head[s]  [B,512,N_s]       the gradient the head's backward sends into the fusion output y
    bbox        [B,4]

One step = forward of a2-a11, targets + the five losses (a12-a17), train-time decode + IoU (a18, a15), and the backward
of all of it.  bench.py captures step() in a CUDA graph; the tests compare it with oracle.hotpath_restated()."""
import torch
import torch.nn as nn

from . import losses as LS
from . import ops, parallel
from .model.DCNet_model import CROSS_NEG_N, NEG_N, TOP_K, grounding_model


def _tensors(o):
    if torch.is_tensor(o):
        yield o
    elif isinstance(o, dict):
        for v in o.values():
            yield from _tensors(v)
    elif isinstance(o, (list, tuple)):
        for v in o:
            yield from _tensors(v)


def all_ordered_pairs(clips, n_frame, device):
    """BASELINE configs[3]: every ordered pair (i, j), i != j, of the n_frame frames of each clip -> (qa, kb) int32 problem lists
    for ops.coattention (queries frame qa attend to frame kb).  Generalises model/test_DCNet_model.py:303-332 (centre frame vs
    the others) to all pairs; frames are numbered clip-major."""
    qa = [c * n_frame + i for c in range(clips) for i in range(n_frame) for j in range(n_frame) if i != j]
    kb = [c * n_frame + j for c in range(clips) for i in range(n_frame) for j in range(n_frame) if i != j]
    return (torch.tensor(qa, device=device, dtype=torch.int32), torch.tensor(kb, device=device, dtype=torch.int32))


class _NoBackbone(nn.Module):
    def forward(self, x):
        raise RuntimeError("HotPath has no backbone: feed raw feature maps to step()")


class HotPath(nn.Module):
    def __init__(self, size=256, vocab=1000, cross_gpu_negatives=False):
        super().__init__()
        self.size = size
        # BASELINE config 5: rank-loss / pixel-to-text negatives taken from the GLOBAL batch (partner Bg-1-g) through an
        # all-gather of the text vectors and target cells (dcnet_b200/parallel.py); off = the reference's local reversal
        self.cross_gpu_negatives = cross_gpu_negatives
        # the mirror model supplies the hot-path parameters (same names/shapes/init as the reference)
        self.net = grounding_model(corpus=list(range(vocab)), emb_size=512, visumodel=_NoBackbone(), size=size)
        self.grids = [size // 32, size // 16, size // 8]
        self.hot_parameters = [p for n, p in self.net.named_parameters()
                               if n.startswith(("mapping_visu", "corr_conv")) or ".0.conv" in n and n.startswith("fcn_emb")
                               or ".0.bn" in n and n.startswith("fcn_emb")]

    # The three pyramid scales are independent from the Darknet maps up to the losses.  The coarse scales launch 16-128 CTAs per
    # kernel on a 148-SM part, so each scale's chain (forward, and its backward: autograd replays every node on the stream its
    # forward ran on) goes to its own CUDA stream; the finest scale stays on the caller's stream.  Fork/join with events, so
    # the whole step still captures into one CUDA graph.
    scale_streams = True
    # Stream priorities (0 = default, -1 = higher).  The auxiliary branches (fusion text terms, the two sampling + InfoNCE
    # branches, train-time decode) are short kernels whose results the chains wait for; at default priority their CTAs queue
    # behind the persistent one-CTA-per-SM GEMMs of the finest scale.  Measured on C2 (profiles/r1z_stream_priorities.txt):
    # aux -1 -> 1.315 ms against 1.371 ms; raising the coarse-scale streams as well, or instead, gives nothing (1.33-1.36 ms).
    side_priority = (0, 0)
    aux_priority = (-1, -1)
    terms_on_aux = True
    # the finest scale's chain (the critical path of the step) on a stream of its own priority; None = the caller's stream (priority 0)
    finest_priority = None
    coarse_on_one_stream = False   # both coarse scales' chains on one side stream (measured: see profiles/r3t_variants_c3.txt)
    finest_first = False     # measured: 7.96 ms against 7.85 ms (C3), 1.475 against 1.439 (C2), profiles/r2l_variants.txt

    def _run_scales(self, chain):
        if not self.scale_streams:
            return [chain(s) for s in range(3)]
        cur = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = [torch.cuda.Stream(priority=p) for p in self.side_priority]
            # parameters are shared by nothing across scales, but their AccumulateGrad nodes live on the stream of the first
            # iteration; the engine synchronises the streams itself, the warning about it is noise here
            if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        outs = [None, None, None]
        fork = cur.record_event()
        for s in (0, 1):
            st = self._side[0 if self.coarse_on_one_stream else s]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                outs[s] = chain(s)
        # (issuing the finest chain first was measured and changes nothing, profiles/r2j_timeline_c3.txt: its first GEMM then shares
        # the SMs with the coarsest scale's exact-fp32 conv instead of waiting behind it)
        if self.finest_priority is None:
            outs[2] = chain(2)
        else:
            if getattr(self, "_fine", None) is None:
                self._fine = torch.cuda.Stream(priority=self.finest_priority)
            self._fine.wait_event(fork)
            with torch.cuda.stream(self._fine):
                outs[2] = chain(2)
            cur.wait_stream(self._fine)
            for t in _tensors(outs[2]):
                t.record_stream(cur)
        if getattr(self, "_aux_pending", False):
            for aux in self._aux:
                cur.wait_stream(aux)
            self._aux_pending = False
        for s in (0, 1):
            if not (self.coarse_on_one_stream and s == 1):
                cur.wait_stream(self._side[s])
            for t in _tensors(outs[s]):
                t.record_stream(cur)          # produced on a side stream, consumed (and later freed) on the caller's stream
        return outs

    def _branch(self, i, fn, *inputs):
        """fn() on auxiliary stream i, forked from the current one; joined by _run_scales (or immediately without streams)"""
        if not self.scale_streams:
            return fn()
        cur = torch.cuda.current_stream()
        if getattr(self, "_aux", None) is None:
            self._aux = [torch.cuda.Stream(priority=p) for p in self.aux_priority]
        aux = self._aux[i]
        aux.wait_stream(cur)
        with torch.cuda.stream(aux):
            out = fn()
        for t in inputs:
            t.record_stream(aux)
        self._aux_pending = True
        return out

    def draw_indices(self, B):
        """host side of the sampling blocks: exact reference RNG stream -> (negpos [P,30,10] int32, negidx [B,N0,5] int64) numpy"""
        N0 = self.grids[0] ** 2
        return ops.pyrandom_interframe(B // 2, TOP_K, N0, NEG_N), ops.pyrandom_crossmodal(B, N0, CROSS_NEG_N)

    def forward_losses(self, raw, flang, fa, context, head, loc, bbox, negpos=None, negidx=None, decode=True):
        net = self.net
        LS.configure(size=self.size)
        hw = [(m.shape[2], m.shape[3]) for m in raw]
        # `fa` feeds every scale's chain.  As a leaf, its AccumulateGrad node lives on the stream of its first consumer -- the
        # coarsest scale's side stream -- and autograd queues every scale's contribution on that stream in the order the CPU walks
        # the graph (finest chain first), so the coarsest chain's own backward sat in the queue behind the finest chain's gradient
        # (profiles/r1z_timeline_graph_replay.txt: it started 400 us late and ended the step).  A view taken on the caller's stream
        # moves that accumulation to the caller's stream, where it queues behind the finest chain's own work.
        if fa.requires_grad:
            fa = fa.view_as(fa)
        if flang.requires_grad:
            flang = flang.view_as(flang)       # same for the sentence vector: every scale's fusion layer consumes it
        best_n, gi, gj, t5, _, _ = ops.build_target(bbox, self.size, LS.args.anchor_imsize, LS.anchors_full)
        fa_neg = partner3 = None
        if self.cross_gpu_negatives:
            fa_neg, partner3 = parallel.global_partners(fa, best_n, gi, gj)

        coords = [ops.coord_map(hw[s][0], hw[s][1], fa.device).flatten(1) for s in range(3)]
        # text / coordinate terms of the three fusion layers (a8): small kernels that depend on nothing of the chains.  Issued up
        # front on an auxiliary stream, their backward (weight / text gradients) also stays off the chains' critical path -- that
        # matters at 256x256 where a chain is a sequence of 10-40 us kernels; terms_on_aux = False computes them inside the fusion
        # layer's own node instead (no extra weight-gradient add).
        terms, ev_terms = [None, None, None], None
        if self.terms_on_aux:
            terms = self._branch(0, lambda: [net.fuse_terms(s, flang, coords[s]) for s in range(3)], flang)
            ev_terms = self._aux[0].record_event() if self.scale_streams else None

        # The finest scale's chain is the critical path of the step; the coarsest scale opens with an exact-fp32 CUDA-core conv whose
        # 256 CTAs hold the SMs for ~350 us at 416x416 (profiles/r2k_timeline_c3.txt: the finest scale's first tcgen05 GEMM waited
        # behind it).  The finest visual mapping is therefore issued BEFORE the fork: the other chains start once it is done and
        # have ~1 ms of slack to the end of the forward anyway.
        fv_first = net.map_visual_scale(raw[2], 2) if self.finest_first else None

        def chain(s):
            """everything of one pyramid scale: a2 -> (a4, a11 on the coarsest scale) -> a5/a6/a9 -> a7/a8 -> a10"""
            o = {}
            o['fv'] = fv_first if (s == 2 and fv_first is not None) else net.map_visual_scale(raw[s], s)
            if s == 0:
                # the two sampling blocks and their InfoNCE losses only need fvisu[0]: a branch of their own next to the
                # co-attention / fusion chain of this scale (forward and, through autograd, backward)
                # (the host RNG stream is consumed in the reference's order: inter-frame draws, then cross-modal draws)
                def inter():
                    r = net.interframe(o['fv'], negpos)
                    return {'if': r, 'l_if': LS.Interframe_contrastive_loss(*r[:3])}

                def cross():
                    r = net.crossmodal(o['fv'], context, negidx)
                    return {'cm': r, 'l_cm': LS.Crossmodal_constrastive_loss(*r[:3])}
                o.update(self._branch(0, inter, o['fv']))
                o.update(self._branch(1, cross, o['fv']))
            o['corr'], o['sim'], o['neg_sim'] = net.correspondence_scale(o['fv'], s, fa, fa_neg)
            if ev_terms is not None:
                torch.cuda.current_stream().wait_event(ev_terms)
                for t in terms[s]:
                    if t is not None:
                        t.record_stream(torch.cuda.current_stream())
            o['y'] = net.fuse_scale(o['corr'], s, flang, coords[s], terms=terms[s])
            o['obj'] = ops.only_obj(head[s], o['sim'])
            o['pred'] = ops.modulate_conf(head[s], o['sim'], loc[s])
            return o

        sc = self._run_scales(chain)
        fv = [o['fv'] for o in sc]
        q_if, k_if, neg_if, idx_if, _ = sc[0]['if']
        q_cm, k_cm, neg_cm, word, _ = sc[0]['cm']
        corr, sim, neg_sim = [o['corr'] for o in sc], [o['sim'] for o in sc], [o['neg_sim'] for o in sc]
        y = [o['y'] for o in sc]
        oo_obj = [o['obj'] for o in sc]
        pred = [o['pred'] for o in sc]
        loss, comp, cell = LS.fused_losses(pred, sim, neg_sim, loc, bbox, q_if, k_if, neg_if, q_cm, k_cm, neg_cm,
                                           target=(best_n, gi, gj, t5), partner3=partner3, l_if=sc[0]['l_if'], l_cm=sc[0]['l_cm'])
        boxes = iou = None
        if decode:
            boxes, iou, _, _, _ = LS.decode_boxes(pred, bbox, cell[:3])
        return dict(loss=loss, comp=comp, y=y, iou=iou, boxes=boxes, cell=cell, corr=corr, sim=sim, pred=pred, bbox=bbox,
                    obj=[o[1] for o in oo_obj], idx_if=idx_if, word=word, fv=fv)

    def step(self, raw, flang, fa, context, head, loc, dy_head, bbox, negpos=None, negidx=None, return_internals=False):
        """forward + backward.  Returns a [6+B] tensor: (loss, yolo, rank, loc, interframe, cross, iou[0..B));
        with return_internals also the dict of forward_losses (tests: activation patterns, indices)."""
        out = self.forward_losses(raw, flang, fa, context, head, loc, bbox, negpos, negidx, decode=False)
        # train-time decode + IoU (a18, a15) do not feed the backward: they go to an auxiliary stream instead of sitting between
        # the loss and the first backward kernel on the critical path
        def dec():
            with torch.no_grad():
                return LS.decode_boxes([p.detach() for p in out['pred']], bbox, out['cell'][:3])
        boxes, iou, _, _, _ = self._branch(1, dec, *out['pred'])
        out['boxes'], out['iou'] = boxes, iou
        torch.autograd.backward([out['loss']] + list(out['y']), [None] + list(dy_head))
        if getattr(self, "_aux_pending", False):
            cur = torch.cuda.current_stream()
            for aux in self._aux:
                cur.wait_stream(aux)
            self._aux_pending = False
            iou.record_stream(cur)
        c = out['comp']
        vec = torch.cat([torch.stack([out['loss'], c['yolo'], c['rank'], c['loc'], c['interframe'], c['cross']]).detach(), out['iou']])
        return (vec, out) if return_internals else vec
