"""Drop-in for the reference's model/DCNet_model.py: same constructor, same forward(image, word_id, word_mask)
signature and return tuples (model/DCNet_model.py:340, :646-650), same parameter names and shapes.
The dense-correspondence hot path (:356-469, :489-505, :525-552, :612-637) runs on hand-written sm_100a kernels through
dcnet_b200.ops; the text encoder, the 3x3 head and the location branch stay PyTorch like in the reference."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .darknet import ConvBatchNormReLU

TOP_K, NEG_N, CROSS_NEG_N = 30, 10, 5      # model/DCNet_model.py:391-392, :53


class PackedList(list):
    """A python list of per-rank / per-pixel tensors (what the reference returns) that remembers the packed tensor
    it is a view of, so the fused losses can consume it without re-stacking."""
    packed = None


def _packed_list(t):
    out = PackedList(t.unbind(0))
    out.packed = t
    return out


def generate_coord(batch, height, width, device="cuda"):
    """[batch,8,h,w] coordinate map (model/DCNet_model.py:23-39), generated on the device."""
    return ops.coord_map(height, width, device).unsqueeze(0).repeat(batch, 1, 1, 1)


class RNNEncoder(nn.Module):
    """model/DCNet_model.py:124-188 (stays PyTorch / cuDNN)."""

    def __init__(self, vocab_size, word_embedding_size, word_vec_size, hidden_size, bidirectional=False,
                 input_dropout_p=0, dropout_p=0, n_layers=1, rnn_type='lstm', variable_lengths=True):
        super().__init__()
        self.variable_lengths = variable_lengths
        self.embedding = nn.Embedding(vocab_size, word_embedding_size)
        self.input_dropout = nn.Dropout(input_dropout_p)
        self.mlp = nn.Sequential(nn.Linear(word_embedding_size, word_vec_size), nn.ReLU())
        self.rnn_type = rnn_type
        self.rnn = getattr(nn, rnn_type.upper())(word_vec_size, hidden_size, n_layers, batch_first=True,
                                                 bidirectional=bidirectional, dropout=dropout_p)
        self.num_dirs = 2 if bidirectional else 1

    def forward(self, input_labels):
        lengths = (input_labels != 0).sum(1).cpu().numpy().tolist()
        assert max(lengths) == input_labels.size(1)
        sort_ixs = np.argsort(lengths)[::-1].tolist()              # descending, same tie order as the reference
        sorted_lengths = [lengths[i] for i in sort_ixs]
        recover = [0] * len(lengths)
        for r, s in enumerate(sort_ixs):
            recover[s] = r
        sort_t = torch.tensor(sort_ixs, device=input_labels.device, dtype=torch.long)
        recover_t = torch.tensor(recover, device=input_labels.device, dtype=torch.long)
        embedded = self.mlp(self.input_dropout(self.embedding(input_labels[sort_t])))
        packed = nn.utils.rnn.pack_padded_sequence(embedded, sorted_lengths, batch_first=True)
        output, _ = self.rnn(packed)
        embedded = embedded[recover_t]           # already padded: pack/unpack of `embedded` is the identity up to zeroed pads
        mask = (torch.arange(embedded.shape[1], device=embedded.device)[None, :] <
                torch.tensor(lengths, device=embedded.device)[:, None])
        embedded = embedded * mask[:, :, None].to(embedded.dtype)
        output, _ = nn.utils.rnn.pad_packed_sequence(output, batch_first=True)
        output = output[recover_t]
        last = torch.tensor([l - 1 for l in lengths], device=output.device, dtype=torch.long)
        sent = output[torch.arange(output.shape[0], device=output.device), last]
        return sent, output, embedded


class PhraseAttention(nn.Module):
    """model/DCNet_model.py:190-219"""

    def __init__(self, input_dim):
        super().__init__()
        self.fc = nn.Linear(input_dim, 1)

    def forward(self, context, embedded, input_labels):
        attn = F.softmax(self.fc(context).squeeze(2), dim=1)
        attn = attn * (input_labels != 0).float()
        attn = attn / attn.sum(1, keepdim=True)
        return attn, torch.bmm(attn.unsqueeze(1), embedded).squeeze(1)


class grounding_model(nn.Module):
    def __init__(self, corpus=None, emb_size=256, jemb_drop_out=0.1, bert_model='bert-base-uncased',
                 coordmap=True, leaky=False, dataset=None, light=False, visumodel=None, size=256, _with_feature_map=True):
        """Extra keyword arguments (both optional, defaults = reference behaviour):
        visumodel -- the Darknet backbone instance (reference: Darknet('./model/yolov3.cfg') + yolov3.weights);
                     when None the surrounding checkout's model.darknet.Darknet is imported (drop-in use).
        size      -- input resolution; sizes the location branch's Linear (1344 positions at 256, :259)."""
        super().__init__()
        self.coordmap = coordmap
        self.light = light
        self.lstm = corpus is not None
        self.emb_size = emb_size
        if not self.lstm:
            raise NotImplementedError("BERT text branch needs pytorch_pretrained_bert (not in scope; use --lstm, README.md:36)")
        if visumodel is None:
            try:
                from model.darknet import Darknet          # the host checkout's backbone (unchanged)
            except Exception as e:                          # pragma: no cover
                raise RuntimeError("grounding_model: pass visumodel=<Darknet instance>; could not import model.darknet (%s)" % e)
            visumodel = Darknet(config_path='./model/yolov3.cfg')
            visumodel.load_weights('./saved_models/yolov3.weights')
        self.visumodel = visumodel
        self.textdim, self.embdim = 1024, 512
        self.textmodel = RNNEncoder(vocab_size=len(corpus), word_embedding_size=self.embdim, word_vec_size=self.textdim // 2,
                                    hidden_size=self.textdim // 2, bidirectional=True, input_dropout_p=0.2, variable_lengths=True)
        self.temperature = 10.
        self.n_pos = sum((size // s) ** 2 for s in (32, 16, 8))
        self.sub_attn = PhraseAttention(self.textdim)
        self.loc_embedding = nn.Sequential(nn.Linear(8, 8), nn.BatchNorm1d(8), nn.ReLU())
        self.loc_text_embedding = nn.Sequential(nn.Linear(self.n_pos, self.embdim), nn.BatchNorm1d(self.embdim), nn.ReLU())
        self.loc_attn = PhraseAttention(self.textdim)
        self.mapping_visu = nn.Sequential(OrderedDict([
            ('0', ConvBatchNormReLU(1024, emb_size, 1, 1, 0, 1, leaky=leaky)),
            ('1', ConvBatchNormReLU(512, emb_size, 1, 1, 0, 1, leaky=leaky)),
            ('2', ConvBatchNormReLU(256, emb_size, 1, 1, 0, 1, leaky=leaky))]))
        self.mapping_lang = nn.Sequential(
            nn.Linear(self.textdim, emb_size), nn.BatchNorm1d(emb_size), nn.ReLU(), nn.Dropout(jemb_drop_out),
            nn.Linear(emb_size, emb_size), nn.BatchNorm1d(emb_size), nn.ReLU())
        self.corr_conv = nn.Sequential(OrderedDict([
            (str(i), nn.Sequential(ConvBatchNormReLU(emb_size * 2, emb_size, 1, 1, 0, 1, leaky=leaky))) for i in range(3)]))
        if _with_feature_map:      # model/test_DCNet_model.py has no feature_map (and draws no init RNG for it)
            self.feature_map = nn.Sequential(nn.Conv1d(20, 20, stride=1, kernel_size=3, padding=1, bias=True), nn.Softmax(dim=1))
        embin_size = emb_size * 2 + (8 if coordmap else 0)
        if light:
            self.fcn_emb = nn.Sequential(OrderedDict([
                (str(i), nn.Sequential(ConvBatchNormReLU(embin_size, emb_size, 1, 1, 0, 1, leaky=leaky))) for i in range(3)]))
            self.fcn_out = nn.Sequential(OrderedDict([
                (str(i), nn.Sequential(nn.Conv2d(emb_size, 3 * 5, kernel_size=1))) for i in range(3)]))
        else:
            self.fcn_emb = nn.Sequential(OrderedDict([
                (str(i), nn.Sequential(ConvBatchNormReLU(embin_size, emb_size, 1, 1, 0, 1, leaky=leaky),
                                       ConvBatchNormReLU(emb_size, emb_size, 3, 1, 1, 1, leaky=leaky),
                                       ConvBatchNormReLU(emb_size, emb_size, 1, 1, 0, 1, leaky=leaky))) for i in range(3)]))
            self.fcn_out = nn.Sequential(OrderedDict([
                (str(i), nn.Sequential(ConvBatchNormReLU(emb_size, emb_size // 2, 1, 1, 0, 1, leaky=leaky),
                                       nn.Conv2d(emb_size // 2, 3 * 5, kernel_size=1))) for i in range(3)]))
        # reproduce the reference's random.sample stream: training draws its negatives from it, eval
        # advances it by what the reference's (discarded) eval-mode sampling consumes.  False: eval leaves `random` untouched.
        self.exact_sampling = True
        self._idx_cache = {}
        self._capture = None
        self.precision = ops.TENSOR_TF32   # GEMM-shaped ops on tcgen05 (TF32 operands, fp32 accumulate); ops.EXACT_FP32 = CUDA cores
        self.fused_coattn = True           # co-attention forward: fused tcgen05 kernel (fp16 operands, S/P never leave the SM)

    # ---------------------------------------------------------------------------------------------------------
    @property
    def coattn_precision(self):
        if getattr(self, "coattn_precision_override", None) is not None:      # experiments / diagnostics
            return self.coattn_precision_override
        if self.precision == ops.EXACT_FP32:
            return ops.EXACT_FP32
        return ops.TENSOR_F16_FUSED if self.fused_coattn else self.precision

    def _pair_index(self, B, device):
        key = ("pair", B, str(device))
        if key not in self._idx_cache:
            qa = torch.arange(B, device=device, dtype=torch.int32)
            self._idx_cache[key] = (qa, qa ^ 1)
        return self._idx_cache[key]

    def map_visual(self, raw_fvisu):
        """a2: fvisu[s] = normalize_c(ConvBNReLU_1x1(raw[s]))   (:356-359) -> 3 x [B,C,N_s].
        Scale 0 runs in exact fp32: the top-30 correspondences (a4) and the arg-max words (a11) are selected from it and
        index parity with the reference needs fp32 scores (SURVEY section 7 "Index parity"); the other scales use tcgen05."""
        return [self.map_visual_scale(raw_fvisu[s], s) for s in range(3)]

    def _head_layer(self, m, y):
        """SURVEY 8(f) row 1, the grounding head (model/DCNet_model.py:316-337, :505-506) on this library's kernels: the 3x3
        ConvBatchNormReLU (fcn_emb[s][1]) as an implicit GEMM on tcgen05 (ops.conv3x3_bn_act), the 1x1 ConvBatchNormReLU layers
        (fcn_emb[s][2]: 512->512, fcn_out[s][0]: 512->256) on the tcgen05 GEMM + fused BN kernels of the rest of the path, the final
        Conv2d(256, 15, 1) with bias as an exact-fp32 pass (ops.conv1x1_bias).  Shapes the 3x3 kernel cannot address (13x13 maps:
        169 positions, row pitch not a multiple of 16 bytes) and the exact-fp32 mode keep the library convolution."""
        B, _, h, w = y.shape
        if not self.head_on_tcgen05 or not y.is_cuda:
            return m(y)
        tf32 = self.precision == ops.TENSOR_TF32
        if isinstance(m, ConvBatchNormReLU) and m.conv.kernel_size == (1, 1) and m.conv.out_channels in (256, 512) and m.conv.in_channels % 128 == 0:
            # inputs rounded by the producing layer of this head, outputs rounded for the next one
            out = m.fused(y.flatten(2), precision=self.precision, round_in=not getattr(y, "_dcnet_rounded", False), round_out=tf32).view(B, -1, h, w)
            out._dcnet_rounded = tf32
            return out
        if tf32 and isinstance(m, ConvBatchNormReLU) and m.conv.kernel_size == (3, 3) and m.conv.stride == (1, 1) and m.conv.padding == (1, 1) \
                and m.conv.dilation == (1, 1) and ops.conv3x3_supported(m.conv.in_channels, m.conv.out_channels, h, w):
            bn = m.bn
            out = ops.conv3x3_bn_act(y.flatten(2), m.conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training, h, w,
                                     momentum=bn.momentum, eps=bn.eps, slope=m.slope, num_batches_tracked=bn.num_batches_tracked,
                                     round_out=True).view(B, -1, h, w)
            out._dcnet_rounded = True
            return out
        if isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1) and m.bias is not None:
            return ops.conv1x1_bias(y.flatten(2), m.weight.view(m.out_channels, -1), m.bias).view(B, -1, h, w)
        return m(y)

    head_on_tcgen05 = True

    def map_visual_scale(self, raw_s, s):
        p0 = ops.EXACT_FP32 if self.precision == ops.EXACT_FP32 else ops.EXACT_FWD_TF32_BWD     # gradients select no index
        # the maps of scales 1, 2 only feed tf32 contractions (corr_conv, co-attention backward): they leave rounded to tf32;
        # scale 0 stays exact fp32 (the index selections read it) and is rounded where the contractions start (correspondence_scale)
        m = self.mapping_visu._modules[str(s)]
        N = raw_s.shape[2] * raw_s.shape[3]
        # ... and the same kernel writes the fused co-attention's staging (fp16 copy + column norms) from its registers
        stage = (s > 0 and self.precision == ops.TENSOR_TF32 and self.coattn_precision == ops.TENSOR_F16_FUSED and N >= ops.FUSED_MIN_N
                 and m.conv.out_channels % 128 == 0 and m.conv.out_channels <= 512)
        # (the Darknet map itself is read truncated: a pass over it would cost more than the layer's BN kernel, and a uniform
        # relative bias of z is removed by the BatchNorm that follows; the weight gradient carries it once, ~2e-4)
        out = m.fused(raw_s.flatten(2), l2norm=True, precision=p0 if s == 0 else self.precision, round_in=False, round_out=s > 0,
                      stage_out=stage)
        if stage:
            out, staged = out
            out._dcnet_staged = staged       # travels with the map to correspondence_scale (same tensor object)
        return out

    def correspondence_scale(self, fv_s, s, fa, fa_neg=None):
        """a5 + a6 + a9 of one scale as one autograd node (ops.correspondence): co-attention both directions, corr_conv on
        [fvisu | attention] without a cat, channel L2 norm, pixel-to-text dots"""
        qa, kb = self._pair_index(fv_s.shape[0], fv_s.device)
        m = self.corr_conv._modules[str(s)][0]
        bn = m.bn
        return ops.correspondence(fv_s, qa, kb, m.conv.weight.view(m.conv.weight.shape[0], -1), bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                  self.training, fa=fa, fa_neg=fa_neg, tau=self.temperature, cprecision=self.coattn_precision,
                                  momentum=bn.momentum, eps=bn.eps, slope=m.slope, precision=self.precision,
                                  num_batches_tracked=bn.num_batches_tracked, round_in=(s == 0), round_out=True,
                                  staged=getattr(fv_s, "_dcnet_staged", None))

    def fuse_terms(self, s, flang, coords_s, kv=512):
        """the text / coordinate terms of scale s on their own (ops.fuse_terms): for callers that issue them ahead of the chain"""
        m = self.fcn_emb._modules[str(s)][0]
        w = m.conv.weight.view(m.conv.weight.shape[0], -1)
        r = ops.fuse_terms(w, flang, coords_s if self.coordmap else None, kv)
        return r if self.coordmap else (r, None)

    def fuse_scale(self, corr_s, s, flang, coords_s, terms=None):
        """a7 + a8 for one scale (:489-505): split-weight form of cat([corr, flang_tile, coord]) -> 1x1 conv + BN + ReLU (SURVEY
        Appendix A.9).  The text / coordinate terms W_l flang, W_c coord and their gradients run on this library's kernels
        (dcnet_fuse_terms_*), the visual term on the tcgen05 GEMM whose epilogue adds them."""
        m = self.fcn_emb._modules[str(s)][0]
        if terms is not None:
            return m.fused(corr_s, u=terms[0], cc=terms[1], l2norm=False, precision=self.precision, round_in=False)
        return m.fused(corr_s, l2norm=False, precision=self.precision, flang=flang, coords=coords_s if self.coordmap else None, round_in=False)

    def interframe(self, fv0, negpos=None):
        """a4 (:381-430) -> packed q [30,P,C], k [30,P,C], neg [30,P,10,C].  negpos: optional pre-drawn device tensor
        [P,30,10] int32 of ops.pyrandom_interframe positions (lets the caller keep the step free of host work)."""
        B, C, N0 = fv0.shape
        P = B // 2
        dev = fv0.device
        idx, _ = ops.interframe_topk(fv0.detach(), TOP_K)
        if negpos is None:
            negpos = torch.from_numpy(ops.pyrandom_interframe(P, TOP_K, N0, NEG_N)).to(dev, non_blocking=True)
        cols = ops.interframe_cols(idx, negpos, N0)                  # rank-major: q [30,P] | k [30,P] | neg [30,P,10]
        key = ("if", P, str(dev))
        if key not in self._idx_cache:
            pair = torch.arange(P, device=dev, dtype=torch.int32)[None, :]
            self._idx_cache[key] = torch.cat([(2 * pair).expand(TOP_K, P).reshape(-1), (2 * pair + 1).expand(TOP_K, P).reshape(-1),
                                              (2 * pair + 1)[:, :, None].expand(TOP_K, P, NEG_N).reshape(-1)]).contiguous()
        img = self._idx_cache[key]
        nq = TOP_K * P
        negidx = cols[2 * nq:].view(TOP_K, P, NEG_N).permute(1, 0, 2)
        g = ops.gather_cols(fv0, img, cols)                          # [30*P*(2+10), C], already in the packed order of the loss
        q = g[:nq].view(TOP_K, P, C)
        k = g[nq:2 * nq].view(TOP_K, P, C)
        neg = g[2 * nq:].view(TOP_K, P, NEG_N, C)
        return q, k, neg, idx, negidx

    def correspondence(self, fv, fa, fa_neg=None):
        """a5 + a6 + a9 (:449-469, :525-535): co-attention both directions, corr_conv on [fvisu | attention] without a
        cat, channel L2 norm, and the pixel-to-text dots fused into the same pass."""
        B = fv[0].shape[0]
        qa, kb = self._pair_index(B, fv[0].device)
        corr, sim, neg_sim = [], [], []
        for s in range(3):
            y, sm, ng = self.correspondence_scale(fv[s], s, fa, fa_neg)
            corr.append(y); sim.append(sm); neg_sim.append(ng)
        return corr, sim, neg_sim

    def fuse(self, corr, flang, coords):
        """a7 + a8 (:489-505) for the three scales.  coords: 3 x [8,N_s] from ops.coord_map."""
        return [self.fuse_scale(corr[s], s, flang, coords[s]) for s in range(3)]

    def crossmodal(self, fv0, context, negidx=None):
        """a11 (:625-637, :41-112) -> packed q [N0,B,C], k [N0,B,1,C], neg [N0,B,5,C].  negidx: optional pre-drawn device
        tensor [B,N0,5] int64 of ops.pyrandom_crossmodal."""
        B, C, N0 = fv0.shape
        dev = fv0.device
        vit = ops.rownorm(fv0)                                   # F.normalize over the spatial axis (:629)
        lag = ops.lagnorm(context)                               # [B,T,C]
        fm = self.feature_map[0]
        word, _ = ops.crossmodal_words(lag.detach(), vit.detach(), fm.weight, fm.bias)
        if negidx is None:
            negidx = torch.from_numpy(ops.pyrandom_crossmodal(B, N0, CROSS_NEG_N)).to(dev, non_blocking=True)
        key = ("cm", B, N0, str(dev))
        if key not in self._idx_cache:
            b = torch.arange(B, device=dev, dtype=torch.int32)
            pix = torch.arange(N0, device=dev, dtype=torch.long)
            self._idx_cache[key] = (b[None, :].expand(N0, B).reshape(-1).contiguous(), pix[:, None].expand(N0, B).reshape(-1).contiguous(),
                                    torch.full((N0 * B * CROSS_NEG_N,), B - 1, device=dev, dtype=torch.int32))
        img_q, col_q, img_n = self._idx_cache[key]
        T = lag.shape[1]
        q = ops.gather_cols(vit, img_q, col_q).view(N0, B, C)
        keyw = ("cmw", B, T, str(dev))
        if keyw not in self._idx_cache:
            self._idx_cache[keyw] = (torch.arange(B, device=dev) * T)[:, None]
        widx = (self._idx_cache[keyw] + word).t().reshape(-1)                                 # rows ordered (pixel, image)
        k = lag.reshape(B * T, C).index_select(0, widx).view(N0, B, 1, C)
        neg = ops.gather_cols(vit, img_n, negidx.permute(1, 0, 2).reshape(-1)).view(N0, B, CROSS_NEG_N, C)
        return q, k, neg, word, negidx

    # inference: the [B,SN,SN] relation tensor of the location branch is never built (ops.loc_rank8)
    rank8_location = True
    # training: the same identity with batch statistics and the backward on this library's kernels (ops.loc_rank8_train: no [B,SN,SN]
    # tensor -- 1.6 GB for 32 images at 416x416 --, no [B*SN,C] activations).  The embedding of a position is the same for every image
    # (same coordinates; BatchNorm1d(8) batch statistics over B copies of the same rows), so the kernels take E = emb[0]: parameter
    # gradients of loc_embedding are sums over the images and do not change (tests/test_gpu_model.py).
    # "torch": the identity through differentiable torch ops (round 1); False: the reference's materialised evaluation.
    rank8_location_train = "kernels"

    def location_branch(self, coords, obj_score, context, embedded, word_id):
        """:556-610; written for any number of positions.  On CUDA the relation matrix is kept in its rank-8 form: three kernels of
        this library at inference, dcnet_loc_rank8_train_fwd / _bwd in training; the tiny coordinate embedding (Linear(8,8) +
        BatchNorm1d(8) + ReLU on the [B*SN, 8] coordinate rows, exactly the reference's evaluation) and the phrase attention stay torch."""
        B = obj_score[0].shape[0]
        _, flang_loc = self.loc_attn(context, embedded, word_id)
        flang_loc = F.normalize(flang_loc, p=2, dim=1)
        obj = F.normalize(torch.cat(obj_score, 1), p=2, dim=1)
        SN = obj.shape[1]
        if self.rank8_location and not self.training and not torch.is_grad_enabled() and obj.is_cuda:
            # eval-mode BatchNorm1d(8): the embedding of a position does not depend on the image
            E = F.normalize(self.loc_embedding(torch.cat([c.t() for c in coords], 0)), p=2, dim=1)
            lin, bn = self.loc_text_embedding[0], self.loc_text_embedding[1]
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            return ops.loc_rank8(E, obj, lin.weight, lin.bias, scale, bn.bias - bn.running_mean * scale, flang_loc)
        coord_map_ = torch.cat([c.t() for c in coords], 0)[None].expand(B, -1, -1)
        emb = self.loc_embedding(coord_map_.reshape(-1, 8)).reshape(B, SN, -1)
        emb = F.normalize(emb, p=2, dim=2)
        lt = self.loc_text_embedding
        if self.rank8_location_train == "kernels" and self.training and obj.is_cuda and isinstance(lt[1], nn.BatchNorm1d) \
                and isinstance(lt[2], nn.ReLU) and lt[1].momentum is not None and lt[0].out_features % 32 == 0:
            lin, bn = lt[0], lt[1]
            return ops.loc_rank8_train(emb[0], obj, lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                       bn.num_batches_tracked, bn.momentum, bn.eps, flang_loc)
        if self.rank8_location_train:
            # same identity with differentiable torch ops: Linear(bmm(E,E^T)*obj) = E (E^T diag(obj) W^T); the modules (and their
            # batch statistics / running-stat updates) are the reference's, only the [B,SN,SN] tensor and its SN-long GEMM are gone
            lin = self.loc_text_embedding[0]
            G = torch.einsum('cq,bq,bqk->bck', lin.weight, obj, emb)
            z = torch.einsum('bck,bpk->bpc', G, emb) + lin.bias
            for m in list(self.loc_text_embedding)[1:]:
                z = m(z.reshape(-1, z.shape[-1]))
            rel = z.reshape(B, SN, -1).permute(0, 2, 1)
        else:
            rel = torch.bmm(emb, emb.transpose(1, 2)) * obj[:, None, :]
            rel = self.loc_text_embedding(rel.reshape(-1, SN)).reshape(B, SN, -1).permute(0, 2, 1)
        rel = F.normalize(rel, p=2, dim=1)
        m = (rel * flang_loc[:, :, None]).sum(1)
        mn, mx = m.min(1)[0][:, None], m.max(1)[0][:, None]
        return (m - mn) / (mx - mn + 1e-6)

    # ---------------------------------------------------------------------------------------------------------
    def forward(self, image, word_id, word_mask):
        raw_fvisu = self.visumodel(image)
        if not raw_fvisu[0].is_cuda:
            raise RuntimeError("dcnet_b200.grounding_model: feature maps are on %s; the hot path has no CPU fallback" % raw_fvisu[0].device)
        B = raw_fvisu[0].shape[0]
        hw = [(m.shape[2], m.shape[3]) for m in raw_fvisu]
        fv = self.map_visual(raw_fvisu)
        if self.training:
            q_if, k_if, neg_if, _, _ = self.interframe(fv[0])

        max_len = int((word_id != 0).sum(1).max().item())
        word_id = word_id[:, :max_len]
        raw_flang, context, embedded = self.textmodel(word_id)
        flang = F.normalize(self.mapping_lang(raw_flang), p=2, dim=1)
        _, fa = self.sub_attn(context, embedded, word_id)
        fa = F.normalize(fa, p=2, dim=1)

        if self._capture is not None:      # debugging / tests: expose the text-side tensors entering the hot path
            for name, t in (("flang", flang), ("fa", fa), ("context", context)):
                t.retain_grad()
                self._capture[name] = t
        corr, sim, neg_sim = self.correspondence(fv, fa)
        coords = [ops.coord_map(h, w, fa.device).flatten(1) for (h, w) in hw]
        inter = self.fuse(corr, flang, coords)
        outbox_raw = []
        for s in range(3):
            y = inter[s].view(B, -1, hw[s][0], hw[s][1])
            for m in list(self.fcn_emb._modules[str(s)])[1:]:
                y = self._head_layer(m, y)
            for m in self.fcn_out._modules[str(s)]:
                y = self._head_layer(m, y)
            outbox_raw.append(y.flatten(2))

        oo, obj = zip(*[ops.only_obj(outbox_raw[s], sim[s]) for s in range(3)])
        locmap = self.location_branch(coords, list(obj), context, embedded, word_id)
        loc, st = [], 0
        for s in range(3):
            n = hw[s][0] * hw[s][1]
            loc.append(locmap[:, st:st + n].contiguous()); st += n
        outbox = [ops.modulate_conf(outbox_raw[s], sim[s], loc[s]) for s in range(3)]

        shp = lambda t, s: t.view(t.shape[:-1] + hw[s])
        outbox = [shp(outbox[s], s) for s in range(3)]
        sim_score = [shp(sim[s], s) for s in range(3)]
        loc_score = [shp(loc[s], s) for s in range(3)]
        if not self.training:
            if self.exact_sampling:
                # The reference runs both sampling blocks in eval mode too and drops their results (:381-430, :625-637 are
                # unconditional).  Their only lasting effect is on Python's global `random` stream: advance it
                # exactly as the reference would (host-only C emulation, no device work), so a run that alternates train and
                # validation epochs keeps drawing the reference's negatives.  exact_sampling = False skips this.
                N0 = hw[0][0] * hw[0][1]
                ops.pyrandom_interframe(B // 2, TOP_K, N0, NEG_N)
                ops.pyrandom_crossmodal(B, N0, CROSS_NEG_N)
            return outbox, sim_score, loc_score, [shp(o, s) for s, o in enumerate(oo)]
        q_cm, k_cm, neg_cm, _, _ = self.crossmodal(fv[0], context)
        self.last_neg_sim_score = [shp(neg_sim[s], s) for s in range(3)]      # train_DCNet.py:623-627, fused
        corr_feat = [shp(corr[s], s) for s in range(3)]
        return (outbox, sim_score, loc_score, corr_feat, fa[:, :, None, None],
                _packed_list(q_if), _packed_list(k_if), _packed_list(neg_if),
                _packed_list(q_cm), _packed_list(k_cm), _packed_list(neg_cm))


def Crossmodal_corrspondence(lag_feature, vit_feature, lag_vit_map, top_k=1):
    """Signature of model/DCNet_model.py:41.  lag_feature [B,T,C], vit_feature [B,C,N0], lag_vit_map [B,T,N0] (after
    feature_map).  Returns the three python lists of the reference; negatives follow the reference's random.sample stream."""
    assert top_k == 1
    B, C, N0 = vit_feature.shape
    dev = vit_feature.device
    T = lag_feature.shape[1]
    word = lag_vit_map.argmax(dim=1)
    negidx = torch.from_numpy(ops.pyrandom_crossmodal(B, N0, CROSS_NEG_N)).to(dev)
    b = torch.arange(B, device=dev, dtype=torch.int32)
    pix = torch.arange(N0, device=dev, dtype=torch.long)
    q = ops.gather_cols(vit_feature, b[None, :].expand(N0, B).reshape(-1).contiguous(),
                        pix[:, None].expand(N0, B).reshape(-1).contiguous()).view(N0, B, C)
    widx = (torch.arange(B, device=dev)[:, None] * T + word).t().reshape(-1)
    k = lag_feature.reshape(B * T, C).index_select(0, widx).view(N0, B, 1, C)
    neg = ops.gather_cols(vit_feature, torch.full((N0 * B * CROSS_NEG_N,), B - 1, device=dev, dtype=torch.int32),
                          negidx.permute(1, 0, 2).reshape(-1).contiguous()).view(N0, B, CROSS_NEG_N, C)
    return _packed_list(q), _packed_list(k), _packed_list(neg)
