"""Drop-in for the reference's model/test_DCNet_model.py (test-time multi-frame model):
forward(image [b*n_frame,3,S,S], word_id [b,T], word_mask, n_frame=5) -> (outbox, sim_score, loc_score, corr_feat, X) with
X = only_obj in eval / flang_attn in train (model/test_DCNet_model.py:284, :480-483).  The centre frame of every clip attends
to each other frame (one co-attention direction per problem, :303-320 -> cal_corr_feat :247-282), each result goes through
corr_conv + channel L2 norm, and the n_frame-1 normalised maps are averaged (:324-332).  No sampling blocks, no feature_map."""
import torch
import torch.nn.functional as F

from .. import ops
from .DCNet_model import grounding_model as _Base


class grounding_model(_Base):
    # the head's 1x1 layers stay cuDNN fp32 here: with tf32 contractions in the head the clip model's `only_obj` (a mean of
    # near-cancelling confidence logits times sim) moved by 9e-3 norm-relative on a one-clip batch (bar 3e-3); the 2-frame model
    # holds its bars with the head on tcgen05
    head_on_tcgen05 = False

    def __init__(self, corpus=None, emb_size=256, jemb_drop_out=0.1, bert_model='bert-base-uncased',
                 coordmap=True, leaky=False, dataset=None, light=False, visumodel=None, size=256):
        super().__init__(corpus, emb_size, jemb_drop_out, bert_model, coordmap, leaky, dataset, light, visumodel, size,
                         _with_feature_map=False)

    def clip_correspondence(self, fv, n_frame):
        """fv: 3 x [b*n_frame, C, N_s] -> corr_feat 3 x [b, C, N_s] (mean over the n_frame-1 partners of the centre frame)."""
        BF, C, _ = fv[0].shape
        b = BF // n_frame
        centre = n_frame // 2
        dev = fv[0].device
        key = ("clip", b, n_frame, str(dev))
        if key not in self._idx_cache:
            others = [i for i in range(n_frame) if i != centre]
            qa = torch.tensor([c * n_frame + centre for _ in others for c in range(b)], device=dev, dtype=torch.int32)   # partner-major
            kb = torch.tensor([c * n_frame + o for o in others for c in range(b)], device=dev, dtype=torch.int32)
            self._idx_cache[key] = (qa, kb, qa.long())
        qa, kb, qa_l = self._idx_cache[key]
        K = n_frame - 1
        out = []
        for s in range(3):
            attn = ops.coattention(fv[s], qa, kb, tau=self.temperature, precision=self.coattn_precision)          # [K*b, C, N]
            x1 = fv[s].index_select(0, qa_l)                                                                 # centre frames, partner-major
            m = self.corr_conv._modules[str(s)][0]
            if self.training:
                # the reference applies corr_conv once per partner, i.e. BatchNorm statistics per call (:312-320)
                ys = [m.fused(x1[k * b:(k + 1) * b], x2=attn[k * b:(k + 1) * b], l2norm=True, precision=self.precision) for k in range(K)]
                y = torch.stack(ys, 0)
            else:
                y = m.fused(x1, x2=attn, l2norm=True, precision=self.precision).view(K, b, C, -1)
            out.append(y.mean(0))
        return out

    def forward(self, image, word_id, word_mask, n_frame=5):
        raw_fvisu = self.visumodel(image)
        if not raw_fvisu[0].is_cuda:
            raise RuntimeError("dcnet_b200.grounding_model: feature maps are on %s; the hot path has no CPU fallback" % raw_fvisu[0].device)
        b = raw_fvisu[0].shape[0] // n_frame
        hw = [(m.shape[2], m.shape[3]) for m in raw_fvisu]
        fv = self.map_visual(raw_fvisu)
        corr = self.clip_correspondence(fv, n_frame)

        max_len = int((word_id != 0).sum(1).max().item())
        word_id = word_id[:, :max_len]
        raw_flang, context, embedded = self.textmodel(word_id)
        flang = F.normalize(self.mapping_lang(raw_flang), p=2, dim=1)
        _, fa = self.sub_attn(context, embedded, word_id)
        fa = F.normalize(fa, p=2, dim=1)

        coords = [ops.coord_map(h, w, fa.device).flatten(1) for (h, w) in hw]
        inter = self.fuse(corr, flang, coords)
        outbox_raw = []
        for s in range(3):
            y = inter[s].view(b, -1, hw[s][0], hw[s][1])
            for m in list(self.fcn_emb._modules[str(s)])[1:]:
                y = self._head_layer(m, y)
            for m in self.fcn_out._modules[str(s)]:
                y = self._head_layer(m, y)
            outbox_raw.append(y.flatten(2))
        if torch.is_grad_enabled() and any(c.requires_grad for c in corr):
            sim = [(fa[:, :, None] * corr[s]).sum(1) for s in range(3)]          # differentiable form (the test model is used under no_grad)
        else:
            sim = [ops.pix2text(corr[s], fa) for s in range(3)]
        oo, obj = zip(*[ops.only_obj(outbox_raw[s], sim[s]) for s in range(3)])
        locmap = self.location_branch(coords, list(obj), context, embedded, word_id)
        loc, st = [], 0
        for s in range(3):
            n = hw[s][0] * hw[s][1]
            loc.append(locmap[:, st:st + n].contiguous()); st += n
        outbox = [ops.modulate_conf(outbox_raw[s], sim[s], loc[s]) for s in range(3)]
        shp = lambda t, s: t.reshape(t.shape[:-1] + hw[s])
        outbox = [shp(outbox[s], s) for s in range(3)]
        sim_score = [shp(sim[s], s) for s in range(3)]
        loc_score = [shp(loc[s], s) for s in range(3)]
        corr_feat = [shp(corr[s], s) for s in range(3)]
        if self.training:
            return outbox, sim_score, loc_score, corr_feat, fa[:, :, None, None]
        return outbox, sim_score, loc_score, corr_feat, [shp(o, s) for s, o in enumerate(oo)]
