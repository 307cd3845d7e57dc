// a10, a14-a19, a21: target assignment, confidence modulation, grounding losses, box decode, YOLO-head decode.
// Warp/block-level kernels; every integer output (best_n, gi, gj, arg-max cell) is bit-exact against the reference:
// the fp32 expressions feeding comparisons use explicit round-to-nearest intrinsics so nvcc cannot contract them
// into FMAs (the reference evaluates them op by op in PyTorch).
#include "common.cuh"

namespace {

struct Anchors9 { float w[9], h[9]; };   // scaled to the grid of their own scale (index i -> scale i/3)
struct Ptr3 { const float* p[3]; };
struct MPtr3 { float* p[3]; };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// utils/utils.py:76-104 with x1y1x2y2=True
__device__ __forceinline__ float iou_xyxy(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2, float by2) {
  const float ix1 = fmaxf(ax1, bx1), iy1 = fmaxf(ay1, by1), ix2 = fminf(ax2, bx2), iy2 = fminf(ay2, by2);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(ix2, ix1), 0.f), fmaxf(__fsub_rn(iy2, iy1), 0.f));
  const float a1 = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
  const float a2 = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), 1e-16f));
}

// train_DCNet.py:265-332.  One thread per sample.
__global__ void build_target_kernel(const float* __restrict__ bbox, int B, int size, Anchors9 an,
                                    long long* __restrict__ best_n, long long* __restrict__ gi_o, long long* __restrict__ gj_o,
                                    float* __restrict__ t5) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float x1 = bbox[b * 4 + 0], y1 = bbox[b * 4 + 1], x2 = bbox[b * 4 + 2], y2 = bbox[b * 4 + 3];
  const float fs = (float)size, f2s = (float)(2 * size);
  const float ncx = __fdiv_rn(__fadd_rn(x1, x2), f2s), ncy = __fdiv_rn(__fadd_rn(y1, y2), f2s);
  const float nw = __fdiv_rn(__fsub_rn(x2, x1), fs), nh = __fdiv_rn(__fsub_rn(y2, y1), fs);
  float best = -INFINITY;
  int bn = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const float grid = (float)(size / (32 >> (i / 3)));
    const float gw = __fmul_rn(nw, grid), gh = __fmul_rn(nh, grid);
    const float v = iou_xyxy(0.f, 0.f, gw, gh, 0.f, 0.f, an.w[i], an.h[i]);
    if (v > best) { best = v; bn = i; }   // np.argmax: first maximum
  }
  const int s = bn / 3;
  const float grid = (float)(size / (32 >> s));
  const float cx = __fmul_rn(ncx, grid), cy = __fmul_rn(ncy, grid);
  const float gw = __fmul_rn(nw, grid), gh = __fmul_rn(nh, grid);
  long long gi = (long long)cx, gj = (long long)cy;   // .long() truncates
  // The reference clamps boxes to [0, size-1] before this point (train_DCNet.py:608), so 0 <= gi, gj < grid there; with an
  // unclamped box its python indexing raises (or wraps for negatives).  Every consumer of (gi, gj) here indexes device memory
  // unchecked (ground loss, scatter, decode), so the cell is saturated into the grid instead of being handed on out of range.
  const long long gmax = (long long)(size / (32 >> s)) - 1;
  gi = gi < 0 ? 0 : (gi > gmax ? gmax : gi);
  gj = gj < 0 ? 0 : (gj > gmax ? gmax : gj);
  best_n[b] = bn;
  gi_o[b] = gi;
  gj_o[b] = gj;
  t5[b * 5 + 0] = __fsub_rn(cx, (float)gi);
  t5[b * 5 + 1] = __fsub_rn(cy, (float)gj);
  t5[b * 5 + 2] = logf(__fadd_rn(__fdiv_rn(gw, an.w[bn]), 1e-16f));
  t5[b * 5 + 3] = logf(__fadd_rn(__fdiv_rn(gh, an.h[bn]), 1e-16f));
  t5[b * 5 + 4] = 1.f;
}

__global__ void scatter_target_kernel(const long long* __restrict__ best_n, const long long* __restrict__ gi, const long long* __restrict__ gj,
                                      const float* __restrict__ t5, int B, int g0, MPtr3 gt, MPtr3 gtc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int s = (int)(best_n[b] / 3), a = (int)(best_n[b] % 3);
  const int g = g0 << s;
  const long long gg = (long long)g * g, cell = gj[b] * g + gi[b];
  for (int k = 0; k < 5; k++) {
    if (gt.p[s]) gt.p[s][(((long long)b * 3 + a) * 5 + k) * gg + cell] = t5[b * 5 + k];
    if (gtc.p[s]) gtc.p[s][((long long)b * 5 + k) * gg + cell] = t5[b * 5 + k];
  }
}

__global__ void only_obj_kernel(const float* __restrict__ raw, const float* __restrict__ sim, float* __restrict__ oo, float* __restrict__ obj,
                                int B, int N) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const long long b = i / N;
  const int n = (int)(i % N);
  const float* r = raw + b * 15 * N + n;
  // torch mean over the anchor dim: sum of 3 then divide
  const float m = __fdiv_rn(__fadd_rn(__fadd_rn(r[4LL * N], r[9LL * N]), r[14LL * N]), 3.f);
  if (oo) oo[i] = m;
  if (obj) obj[i] = __fmul_rn(m, sim[i]);
}

// backward of only_obj / obj_score: d raw[b, 5a+4, n] = (d only_obj + d obj * sim) / 3 for the three anchors (every other channel 0),
// d sim = d obj * only_obj
__global__ void only_obj_bwd_kernel(const float* __restrict__ doo, const float* __restrict__ dobj, const float* __restrict__ sim,
                                    const float* __restrict__ oo, float* __restrict__ draw, float* __restrict__ dsim, int B, int N) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const long long b = i / N;
  const int n = (int)(i % N);
  const float go = doo ? doo[i] : 0.f, gb = dobj ? dobj[i] : 0.f;
  const float d = __fdiv_rn(fmaf(gb, sim[i], go), 3.f);
  float* r = draw + b * 15 * N + n;
#pragma unroll
  for (int k = 0; k < 15; k++) r[(long long)k * N] = (k % 5 == 4) ? d : 0.f;
  dsim[i] = gb * oo[i];
}

__global__ void modulate_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ sim, const float* __restrict__ loc,
                                    float* __restrict__ out, int B, int N) {
  const long long total = (long long)B * 15 * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long bc = i / N;
    const int ch = (int)(bc % 15);
    const long long b = bc / 15;
    float v = raw[i];
    if (ch % 5 == 4) v = __fmul_rn(__fmul_rn(v, sim[b * N + n]), loc[b * N + n]);   // (conf * sim) * loc, :619
    out[i] = v;
  }
}

__global__ void modulate_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ sim, const float* __restrict__ loc,
                                    const float* __restrict__ dout, float* __restrict__ draw, float* __restrict__ dsim,
                                    float* __restrict__ dloc, int B, int N) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const long long b = i / N;
  const int n = (int)(i % N);
  const float s = sim[i], l = loc[i];
  float ds = 0.f, dl = 0.f;
  for (int ch = 0; ch < 15; ch++) {
    const long long o = (b * 15 + ch) * N + n;
    const float d = dout[o];
    if (ch % 5 == 4) {
      const float r = raw[o];
      draw[o] = d * s * l;
      ds = fmaf(d, r * l, ds);
      dl = fmaf(d, r * s, dl);
    } else {
      draw[o] = d;
    }
  }
  dsim[i] = ds;
  dloc[i] = dl;
}

// ------------------------------------------------------------------------------------------------------
// grounding losses, one CTA per sample
// ------------------------------------------------------------------------------------------------------
struct LossIn {
  Ptr3 pred, sim, neg, loc;
  const long long* best_n; const long long* gi; const long long* gj; const float* t5;
  const long long* partner3;   // [3,B] best_n|gi|gj of the rank-loss partner, or NULL = local sample B-1-b
  int B, g0;
  float w_coord, margin;
};

__device__ __forceinline__ float block_lse_conf(const LossIn& a, int b, float* sh) {
  float m = -INFINITY;
  for (int s = 0; s < 3; s++) {
    const int g = a.g0 << s, N = g * g;
    const float* p = a.pred.p[s] + (long long)b * 15 * N;
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) m = fmaxf(m, p[(long long)(5 * (i / N) + 4) * N + i % N]);
  }
  m = block_max(m, sh);
  float sum = 0.f;
  for (int s = 0; s < 3; s++) {
    const int g = a.g0 << s, N = g * g;
    const float* p = a.pred.p[s] + (long long)b * 15 * N;
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) sum += expf(p[(long long)(5 * (i / N) + 4) * N + i % N] - m);
  }
  sum = block_sum(sum, sh);
  return m + logf(sum);
}

__device__ __forceinline__ float block_lse_loc(const LossIn& a, int b, float* sh) {
  float m = -INFINITY;
  for (int s = 0; s < 3; s++) {
    const int g = a.g0 << s, N = g * g;
    const float* p = a.loc.p[s] + (long long)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, p[i]);
  }
  m = block_max(m, sh);
  float sum = 0.f;
  for (int s = 0; s < 3; s++) {
    const int g = a.g0 << s, N = g * g;
    const float* p = a.loc.p[s] + (long long)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sum += expf(p[i] - m);
  }
  sum = block_sum(sum, sh);
  return m + logf(sum);
}

__global__ void __launch_bounds__(256) ground_loss_fwd_kernel(LossIn a, float* __restrict__ losses, float* __restrict__ lse_conf,
                                                              float* __restrict__ lse_loc) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float lc = block_lse_conf(a, b, sh);
  const float ll = block_lse_loc(a, b, sh);
  if (threadIdx.x != 0) return;
  const int s = (int)(a.best_n[b] / 3), an = (int)(a.best_n[b] % 3);
  const int g = a.g0 << s, N = g * g;
  const int cell = (int)(a.gj[b] * g + a.gi[b]);
  const float* p = a.pred.p[s] + ((long long)b * 15 + 5 * an) * N + cell;
  const float* t = a.t5 + b * 5;
  const float ex = sigmoidf_(p[0]) - t[0], ey = sigmoidf_(p[(long long)N]) - t[1];
  const float ew = p[2LL * N] - t[2], eh = p[3LL * N] - t[3];
  const float box = ex * ex + ey * ey + ew * ew + eh * eh;
  const float conf = lc - p[4LL * N];
  // rank loss (train_DCNet.py:173-203): partner sample B-1-b supplies the second negative
  const int rb = a.B - 1 - b;
  const long long pbn = a.partner3 ? a.partner3[b] : a.best_n[rb];
  const long long pgi = a.partner3 ? a.partner3[a.B + b] : a.gi[rb];
  const long long pgj = a.partner3 ? a.partner3[2 * a.B + b] : a.gj[rb];
  const int s2 = (int)(pbn / 3), g2 = a.g0 << s2, N2 = g2 * g2;
  const int cell2 = (int)(pgj * g2 + pgi);
  const float pos = a.sim.p[s][(long long)b * N + cell];
  const float n1 = a.neg.p[s][(long long)b * N + cell];
  const float n2 = a.sim.p[s2][(long long)b * N2 + cell2];
  const float rank = fmaxf(a.margin + n1 - pos, 0.f) + fmaxf(a.margin + n2 - pos, 0.f);
  const float locl = ll - a.loc.p[s][(long long)b * N + cell];
  const float invB = 1.f / (float)a.B;
  atomicAdd(losses + 0, (a.w_coord * box + conf) * invB);
  atomicAdd(losses + 1, rank * 0.5f * invB);
  atomicAdd(losses + 2, locl * invB);
  lse_conf[b] = lc;
  lse_loc[b] = ll;
}

// grid (B, GL_BWD_SPLIT): the elements of a sample are striped over the CTAs of its row; the handful of special cells (GT cell of
// the box / confidence / location terms, the two hinge partners) are resolved inline by whichever thread owns them, so no CTA
// ever touches an element another one writes.
constexpr int GL_BWD_SPLIT = 8;
__global__ void __launch_bounds__(256) ground_loss_bwd_kernel(LossIn a, const float* __restrict__ lse_conf, const float* __restrict__ lse_loc,
                                                              const float* __restrict__ gl, MPtr3 dpred, MPtr3 dsim, MPtr3 dneg, MPtr3 dloc) {
  const int b = blockIdx.x;
  const int t0 = blockIdx.y * blockDim.x + threadIdx.x, tstride = gridDim.y * blockDim.x;
  const float invB = 1.f / (float)a.B;
  const float g_y = gl[0] * invB, g_r = gl[1] * 0.5f * invB, g_l = gl[2] * invB;
  const float lc = lse_conf[b], ll = lse_loc[b];
  // special cells of this sample
  const int sg = (int)(a.best_n[b] / 3), an = (int)(a.best_n[b] % 3);
  const int gg = a.g0 << sg, Ng = gg * gg;
  const int cellg = (int)(a.gj[b] * gg + a.gi[b]);
  const int rb = a.B - 1 - b;
  const long long pbn = a.partner3 ? a.partner3[b] : a.best_n[rb];
  const long long pgi = a.partner3 ? a.partner3[a.B + b] : a.gi[rb];
  const long long pgj = a.partner3 ? a.partner3[2 * a.B + b] : a.gj[rb];
  const int s2 = (int)(pbn / 3), g2 = a.g0 << s2, N2 = g2 * g2;
  const int cell2 = (int)(pgj * g2 + pgi);
  const float pos = a.sim.p[sg][(long long)b * Ng + cellg];
  const float n1 = a.neg.p[sg][(long long)b * Ng + cellg];
  const float n2 = a.sim.p[s2][(long long)b * N2 + cell2];
  const float h1 = (a.margin + n1 - pos >= 0.f) ? g_r : 0.f;   // clamp(min=0) passes the gradient at 0
  const float h2 = (a.margin + n2 - pos >= 0.f) ? g_r : 0.f;
  const float* t5 = a.t5 + b * 5;
  const float kc = 2.f * a.w_coord * g_y;
  for (int s = 0; s < 3; s++) {
    const int g = a.g0 << s, N = g * g;
    const float* p = a.pred.p[s] + (long long)b * 15 * N;
    float* dp = dpred.p[s] + (long long)b * 15 * N;
    for (int i = t0; i < 15 * N; i += tstride) {
      const int ch = i / N, cell = i - ch * N, k5 = ch % 5;
      float v = (k5 == 4) ? g_y * expf(p[i] - lc) : 0.f;
      if (s == sg && cell == cellg && ch / 5 == an) {
        if (k5 == 0 || k5 == 1) {
          const float sx = sigmoidf_(p[i]);
          v = kc * (sx - t5[k5]) * sx * (1.f - sx);
        } else if (k5 == 2 || k5 == 3) {
          v = kc * (p[i] - t5[k5]);
        } else {
          v -= g_y;
        }
      }
      dp[i] = v;
    }
    const float* l = a.loc.p[s] + (long long)b * N;
    for (int i = t0; i < N; i += tstride) {
      const bool atg = (s == sg && i == cellg);
      dloc.p[s][(long long)b * N + i] = g_l * expf(l[i] - ll) - (atg ? g_l : 0.f);
      dsim.p[s][(long long)b * N + i] = (atg ? -(h1 + h2) : 0.f) + ((s == s2 && i == cell2) ? h2 : 0.f);
      dneg.p[s][(long long)b * N + i] = atg ? h1 : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// decode (train_DCNet.py:656-690, :766-816)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_kernel(Ptr3 pred, int B, int g0, Anchors9 an, int mode, long long* __restrict__ best_n,
                                                     long long* __restrict__ gi_o, long long* __restrict__ gj_o, float* __restrict__ boxes,
                                                     const float* __restrict__ target, float* __restrict__ iou) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  const int b = blockIdx.x;
  int s, a, gi, gj;
  if (mode == 1) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    int off = 0;
    for (int sc = 0; sc < 3; sc++) {
      const int g = g0 << sc, N = g * g;
      const float* p = pred.p[sc] + (long long)b * 15 * N;
      for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) {
        const float v = p[(long long)(5 * (i / N) + 4) * N + i % N];
        const int fi = off + i;
        if (v > bv || (v == bv && fi < bi)) { bv = v; bi = fi; }
      }
      off += 3 * N;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[w] = bv; s_i[w] = bi; }
    __syncthreads();
    if (w == 0) {
      bv = lane < (blockDim.x >> 5) ? s_v[lane] : -INFINITY;
      bi = lane < (blockDim.x >> 5) ? s_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) s_i[0] = bi;
    }
    __syncthreads();
    int fi = s_i[0];
    const int n0 = 3 * g0 * g0, n1 = 3 * (2 * g0) * (2 * g0);
    s = fi < n0 ? 0 : (fi < n0 + n1 ? 1 : 2);
    fi -= (s == 0 ? 0 : (s == 1 ? n0 : n0 + n1));
    const int g = g0 << s, N = g * g;
    a = fi / N;
    gj = (fi % N) / g;
    gi = fi % g;
  } else {
    s = (int)(best_n[b] / 3);
    a = (int)(best_n[b] % 3);
    gi = (int)gi_o[b];
    gj = (int)gj_o[b];
  }
  if (threadIdx.x != 0) return;
  const int g = g0 << s, N = g * g;
  const float stride = (float)(32 >> s);
  const float* p = pred.p[s] + ((long long)b * 15 + 5 * a) * N + gj * g + gi;
  const float x = __fmul_rn(__fadd_rn(sigmoidf_(p[0]), (float)gi), stride);
  const float y = __fmul_rn(__fadd_rn(sigmoidf_(p[(long long)N]), (float)gj), stride);
  const float w = __fmul_rn(__fmul_rn(expf(p[2LL * N]), an.w[3 * s + a]), stride);
  const float h = __fmul_rn(__fmul_rn(expf(p[3LL * N]), an.h[3 * s + a]), stride);
  const float bx1 = __fsub_rn(x, __fdiv_rn(w, 2.f)), by1 = __fsub_rn(y, __fdiv_rn(h, 2.f));
  const float bx2 = __fadd_rn(x, __fdiv_rn(w, 2.f)), by2 = __fadd_rn(y, __fdiv_rn(h, 2.f));
  boxes[b * 4 + 0] = bx1; boxes[b * 4 + 1] = by1; boxes[b * 4 + 2] = bx2; boxes[b * 4 + 3] = by2;
  if (mode == 1) { best_n[b] = 3 * s + a; gi_o[b] = gi; gj_o[b] = gj; }
  if (target && iou)
    iou[b] = iou_xyxy(bx1, by1, bx2, by2, target[b * 4 + 0], target[b * 4 + 1], target[b * 4 + 2], target[b * 4 + 3]);
}

// ------------------------------------------------------------------------------------------------------
// 8f-3: test-time cache writer (test_DCNet.py:546-654, get_topk_pred_bbox :657-701) -- per image the top-k confidence
// cells over all scales / anchors, their decoded boxes mapped back to the un-letterboxed image, and the 512-d
// correspondence feature of each cell.  One CTA per image.
//   ranking : torch.topk over the concatenated [3 N0 | 3 N1 | 3 N2] confidences (order anchor, gj, gi inside a scale); ties go to
//             the lower flat index.
//   cell    : the reference takes the SCALE from the flat index and then searches that scale for the FIRST cell whose confidence
//             equals the value (np.where(pred_conf == max_conf)[0], :682) -- with duplicated values both ranks report the first
//             cell; reproduced.
//   box     : (sigmoid(tx)+gi, sigmoid(ty)+gj, exp(tw) aw, exp(th) ah) * stride -> xyxy -> ((x - dw)/ratio, (y - dh)/ratio)
//             -> x1,y1 >= 0, x2 <= img_w, y2 <= img_h (:693-696)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) topk_boxes_kernel(Ptr3 pred, Ptr3 feat, int B, int g0, int C, int k, Anchors9 an,
                                                         const float* __restrict__ meta /* [B,5]: ratio, dw, dh, img_w, img_h */,
                                                         float* __restrict__ boxes, float* __restrict__ scores, long long* __restrict__ cells,
                                                         float* __restrict__ feats) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  __shared__ float s_bv;
  __shared__ int s_bi, s_first;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n0 = 3 * g0 * g0, n1 = 12 * g0 * g0;
  float pv = INFINITY;     // previously selected (value, flat index): the next one is the largest element strictly after it
  int pi = -1;
  for (int r = 0; r < k; r++) {
    float bv = -INFINITY;
    int bi = 0x7fffffff, off = 0;
    for (int sc = 0; sc < 3; sc++) {
      const int g = g0 << sc, N = g * g;
      const float* p = pred.p[sc] + (long long)b * 15 * N;
      for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) {
        const float v = p[(long long)(5 * (i / N) + 4) * N + i % N];
        const int fi = off + i;
        const bool after = v < pv || (v == pv && fi > pi);
        if (after && (v > bv || (v == bv && fi < bi))) { bv = v; bi = fi; }
      }
      off += 3 * N;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[w] = bv; s_i[w] = bi; }
    __syncthreads();
    if (w == 0) {
      bv = lane < nw ? s_v[lane] : -INFINITY;
      bi = lane < nw ? s_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { s_bv = bv; s_bi = bi; s_first = 0x7fffffff; }
    }
    __syncthreads();
    pv = s_bv; pi = s_bi;
    const int s = pi < n0 ? 0 : (pi < n0 + n1 ? 1 : 2);
    const int g = g0 << s, N = g * g;
    // first cell of that scale holding the same confidence
    {
      const float* p = pred.p[s] + (long long)b * 15 * N;
      int first = 0x7fffffff;
      for (int i = threadIdx.x; i < 3 * N; i += blockDim.x)
        if (p[(long long)(5 * (i / N) + 4) * N + i % N] == pv) { first = i; break; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      if (lane == 0 && first != 0x7fffffff) atomicMin(&s_first, first);
    }
    __syncthreads();
    const int fi = s_first;
    const int a = fi / N, gj = (fi % N) / g, gi = fi % g;
    const long long o = (long long)b * k + r;
    // the cell's feature column
    const float* f = feat.p[s] + (long long)b * C * N + gj * g + gi;
    for (int c = threadIdx.x; c < C; c += blockDim.x) feats[o * C + c] = f[(long long)c * N];
    if (threadIdx.x == 0) {
      const float stride = (float)(32 >> s);
      const float* p = pred.p[s] + ((long long)b * 15 + 5 * a) * N + gj * g + gi;
      const float x = __fmul_rn(__fadd_rn(sigmoidf_(p[0]), (float)gi), stride);
      const float y = __fmul_rn(__fadd_rn(sigmoidf_(p[(long long)N]), (float)gj), stride);
      const float bw = __fmul_rn(__fmul_rn(expf(p[2LL * N]), an.w[3 * s + a]), stride);
      const float bh = __fmul_rn(__fmul_rn(expf(p[3LL * N]), an.h[3 * s + a]), stride);
      const float ratio = meta[b * 5 + 0], dw = meta[b * 5 + 1], dh = meta[b * 5 + 2], iw = meta[b * 5 + 3], ih = meta[b * 5 + 4];
      float x1 = __fdiv_rn(__fsub_rn(__fsub_rn(x, __fdiv_rn(bw, 2.f)), dw), ratio), y1 = __fdiv_rn(__fsub_rn(__fsub_rn(y, __fdiv_rn(bh, 2.f)), dh), ratio);
      float x2 = __fdiv_rn(__fsub_rn(__fadd_rn(x, __fdiv_rn(bw, 2.f)), dw), ratio), y2 = __fdiv_rn(__fsub_rn(__fadd_rn(y, __fdiv_rn(bh, 2.f)), dh), ratio);
      boxes[o * 4 + 0] = fmaxf(x1, 0.f); boxes[o * 4 + 1] = fmaxf(y1, 0.f);
      boxes[o * 4 + 2] = fminf(x2, iw); boxes[o * 4 + 3] = fminf(y2, ih);
      scores[o] = pv;
      cells[o * 4 + 0] = s; cells[o * 4 + 1] = a; cells[o * 4 + 2] = gj; cells[o * 4 + 3] = gi;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// 8f-3: post_processing.py:205-270 -- re-scoring of the centre frame's top-k boxes against the cached top-k of R reference
// frames: sim[i,j,r] = <centre_i, ref_{j,r}>; per (i,r) the best matching reference box (first maximum over j) and its
// score; weights = softmax_r(max sim) with invalid frames zeroed AFTER the softmax (:265-268); fused[i] = sum_r w score;
// best = first arg-max of fused.  One CTA; k, R <= 16.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) post_rescore_kernel(const float* __restrict__ centre /* [k,C] */, const float* __restrict__ ref /* [k,R,C] */,
                                                           const float* __restrict__ ref_score /* [k,R] */, const int* __restrict__ invalid /* [R] or null */,
                                                           int k, int R, int C, float* __restrict__ fused, long long* __restrict__ best,
                                                           long long* __restrict__ match /* [k,R] or null */) {
  __shared__ float s_sim[16 * 16 * 16];
  __shared__ float s_w[16 * 16], s_sc[16 * 16], s_f[16];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = w; t < k * k * R; t += nw) {            // t = (i, j, r): one warp per dot product
    const int i = t / (k * R), j = (t / R) % k, r = t % R;
    const float* a = centre + (long long)i * C;
    const float* b = ref + ((long long)j * R + r) * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(a[c], b[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) s_sim[t] = acc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < k * R; t += blockDim.x) {
    const int i = t / R, r = t % R;
    float bv = -INFINITY;
    int bj = 0;
    for (int j = 0; j < k; j++) {
      const float v = s_sim[(i * k + j) * R + r];
      if (v > bv) { bv = v; bj = j; }
    }
    s_w[t] = bv;
    s_sc[t] = ref_score[bj * R + r];
    if (match) match[t] = bj;
  }
  __syncthreads();
  if ((int)threadIdx.x < k) {
    const int i = threadIdx.x;
    float m = -INFINITY;
    for (int r = 0; r < R; r++) m = fmaxf(m, s_w[i * R + r]);
    float z = 0.f;
    for (int r = 0; r < R; r++) z += expf(s_w[i * R + r] - m);
    float f = 0.f;
    for (int r = 0; r < R; r++) {
      const float wgt = (invalid && invalid[r]) ? 0.f : expf(s_w[i * R + r] - m) / z;
      f = fmaf(wgt, s_sc[i * R + r], f);
    }
    s_f[i] = f;
    fused[i] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bi = 0;
    for (int i = 1; i < k; i++)
      if (s_f[i] > s_f[bi]) bi = i;
    *best = bi;
  }
}

__global__ void bbox_iou_kernel(const float* __restrict__ b1, const float* __restrict__ b2, int n, int xyxy, float* __restrict__ iou) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a[4], c[4];
  for (int k = 0; k < 4; k++) { a[k] = b1[i * 4 + k]; c[k] = b2[i * 4 + k]; }
  if (!xyxy) {
    const float ax1 = __fsub_rn(a[0], __fdiv_rn(a[2], 2.f)), ax2 = __fadd_rn(a[0], __fdiv_rn(a[2], 2.f));
    const float ay1 = __fsub_rn(a[1], __fdiv_rn(a[3], 2.f)), ay2 = __fadd_rn(a[1], __fdiv_rn(a[3], 2.f));
    const float cx1 = __fsub_rn(c[0], __fdiv_rn(c[2], 2.f)), cx2 = __fadd_rn(c[0], __fdiv_rn(c[2], 2.f));
    const float cy1 = __fsub_rn(c[1], __fdiv_rn(c[3], 2.f)), cy2 = __fadd_rn(c[1], __fdiv_rn(c[3], 2.f));
    iou[i] = iou_xyxy(ax1, ay1, ax2, ay2, cx1, cy1, cx2, cy2);
  } else {
    iou[i] = iou_xyxy(a[0], a[1], a[2], a[3], c[0], c[1], c[2], c[3]);
  }
}

// ------------------------------------------------------------------------------------------------------
// YOLOLayer decode (model/darknet.py:262-296,365-375): [B, A*(5+nc), g, g] -> [B, A*g*g, 5+nc].
// CTA = (32 cells) x (all attributes) of one (b, anchor): coalesced reads along cells, smem transpose,
// coalesced writes along attributes (the 32 cells of a tile are one contiguous 32*(5+nc)-float span of the output).
// ------------------------------------------------------------------------------------------------------
constexpr int YMAX_ATTR = 96;
struct AnchorsA { float w[8], h[8]; };
// CTA = 128 cells of one (image, anchor): the [nattr][128] slab is read along the cells (coalesced, float4 when the row pitch
// allows), transposed through shared memory, and the contiguous [128][nattr] output slab is written linearly with float4.
// NATTR > 0 fixes the attribute count at compile time (85 for the COCO head) so i / nattr, i % nattr are multiplications.
constexpr int YCELLS = 128;
template <int NATTR>
__global__ void __launch_bounds__(256) yolo_decode_kernel(const float* __restrict__ x, float* __restrict__ out, int A, int nattr_rt, int g,
                                                          float stride, AnchorsA an) {
  extern __shared__ float ytile[];                    // [nattr][YCELLS + 1]
  const int nattr = NATTR > 0 ? NATTR : nattr_rt;
  constexpr int LD = YCELLS + 1;
  const int gg = g * g;
  const int c0 = blockIdx.x * YCELLS;
  const int a = blockIdx.y, b = blockIdx.z;
  const float* xp = x + ((long long)b * A + a) * nattr * gg;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool vec = (gg & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const float aw = an.w[a], ah = an.h[a];
  for (int at = w; at < nattr; at += 8) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int cell = c0 + lane * 4;
    if (vec && cell + 3 < gg) {
      const float4 q = *reinterpret_cast<const float4*>(xp + (long long)at * gg + cell);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++)
        if (cell + e < gg) v[e] = xp[(long long)at * gg + cell + e];
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int ce = cell + e;
      float r = v[e];
      if (at == 0) r = __fmul_rn(__fadd_rn(sigmoidf_(r), (float)(ce % g)), stride);
      else if (at == 1) r = __fmul_rn(__fadd_rn(sigmoidf_(r), (float)(ce / g)), stride);
      else if (at == 2) r = __fmul_rn(__fmul_rn(expf(r), aw), stride);
      else if (at == 3) r = __fmul_rn(__fmul_rn(expf(r), ah), stride);
      else r = sigmoidf_(r);
      // cell 4 lane + e of the tile lives in column 32 e + lane: a warp's 32 stores of one e hit 32 different banks (the natural
      // column 4 lane + e is a 4-way conflict)
      ytile[at * LD + e * 32 + lane] = r;
    }
  }
  __syncthreads();
  const int ncell = min(YCELLS, gg - c0);
  const int total = ncell * nattr;
  float* op = out + (((long long)b * A + a) * gg + c0) * nattr;
  const bool ovec = (reinterpret_cast<uintptr_t>(op) & 15) == 0;
  for (int i = threadIdx.x * 4; i < total; i += blockDim.x * 4) {
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int ii = min(i + e, total - 1);
      const int cl = ii / nattr, at = ii - cl * nattr;
      r[e] = ytile[at * LD + (cl & 3) * 32 + (cl >> 2)];
    }
    if (ovec && i + 3 < total) *reinterpret_cast<float4*>(op + i) = make_float4(r[0], r[1], r[2], r[3]);
    else
      for (int e = 0; e < 4 && i + e < total; e++) op[i + e] = r[e];
  }
}

__global__ void iou_loss_sums_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, float* __restrict__ acc) {
  __shared__ float sh[32];
  float si = 0.f, su = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = sigmoidf_(x[i]), tt = t[i];
    si = fmaf(s, tt, si);
    su += s + tt - s * tt;
  }
  si = block_sum(si, sh);
  su = block_sum(su, sh);
  if (threadIdx.x == 0) { atomicAdd(acc + 0, si); atomicAdd(acc + 1, su); }
}

// d/dx of (-I/U) * gscale
__global__ void iou_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, const float* __restrict__ acc,
                                    const float* __restrict__ g, float gscale, float* __restrict__ dx) {
  const float I = acc[0], U = acc[1];
  if (g) gscale *= g[0];          // upstream gradient read on the device: no host synchronisation, capturable in a CUDA graph
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = sigmoidf_(x[i]), tt = t[i];
    const float dI = tt, dU = 1.f - tt;
    dx[i] = -gscale * (dI * U - I * dU) / (U * U) * s * (1.f - s);
  }
}

Anchors9 scale_anchors9(const float* h, int size, float anchor_imsize) {
  Anchors9 a;
  for (int i = 0; i < 9; i++) {
    const int grid = size / (32 >> (i / 3));
    const double div = (double)anchor_imsize / (double)grid;     // python: x / (anchor_imsize/grid) in float64
    a.w[i] = (float)((double)h[2 * i] / div);
    a.h[i] = (float)((double)h[2 * i + 1] / div);
  }
  return a;
}

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int dcnet_build_target(const float* bbox, int B, int size, float anchor_imsize, const float* h_anchors9x2,
                                  long long* best_n, long long* gi, long long* gj, float* t5,
                                  float* gt0, float* gt1, float* gt2, float* gtc0, float* gtc1, float* gtc2, void* stream) {
  DCNET_CHECK_ARG(bbox && h_anchors9x2 && best_n && gi && gj && t5 && B > 0, "build_target: bad arguments");
  DCNET_CHECK_ARG(size >= 32 && size % 32 == 0, "build_target: size %d must be a positive multiple of 32", size);
  cudaStream_t st = as_stream(stream);
  const Anchors9 an = scale_anchors9(h_anchors9x2, size, anchor_imsize);
  build_target_kernel<<<ceil_div(B, 128), 128, 0, st>>>(bbox, B, size, an, best_n, gi, gj, t5);
  DCNET_LAUNCH_OK("build_target");
  MPtr3 gt{{gt0, gt1, gt2}}, gtc{{gtc0, gtc1, gtc2}};
  bool any = false;
  for (int s = 0; s < 3; s++) {
    const long long g = (size / 32) << s;
    if (gt.p[s]) { DCNET_CUDA(cudaMemsetAsync(gt.p[s], 0, (size_t)B * 15 * g * g * sizeof(float), st), "build_target.memset"); any = true; }
    if (gtc.p[s]) { DCNET_CUDA(cudaMemsetAsync(gtc.p[s], 0, (size_t)B * 5 * g * g * sizeof(float), st), "build_target.memset"); any = true; }
  }
  if (any) {
    scatter_target_kernel<<<ceil_div(B, 128), 128, 0, st>>>(best_n, gi, gj, t5, B, size / 32, gt, gtc);
    DCNET_LAUNCH_OK("build_target.scatter");
  }
  return 0;
}

extern "C" int dcnet_only_obj(const float* raw, const float* sim, float* only_obj, float* obj, int B, int N, void* stream) {
  DCNET_CHECK_ARG(raw && (only_obj || obj) && (!obj || sim) && B > 0 && N > 0, "only_obj: bad arguments");
  only_obj_kernel<<<ceil_div((long long)B * N, 256), 256, 0, as_stream(stream)>>>(raw, sim, only_obj, obj, B, N);
  DCNET_LAUNCH_OK("only_obj");
  return 0;
}

extern "C" int dcnet_only_obj_bwd(const float* d_only_obj, const float* d_obj, const float* sim, const float* only_obj, float* draw, float* dsim,
                                  int B, int N, void* stream) {
  DCNET_CHECK_ARG((d_only_obj || d_obj) && sim && only_obj && draw && dsim && B > 0 && N > 0, "only_obj_bwd: bad arguments");
  only_obj_bwd_kernel<<<ceil_div((long long)B * N, 256), 256, 0, as_stream(stream)>>>(d_only_obj, d_obj, sim, only_obj, draw, dsim, B, N);
  DCNET_LAUNCH_OK("only_obj_bwd");
  return 0;
}

extern "C" int dcnet_modulate_conf_fwd(const float* raw, const float* sim, const float* loc, float* out, int B, int N, void* stream) {
  DCNET_CHECK_ARG(raw && sim && loc && out && B > 0 && N > 0, "modulate_conf_fwd: bad arguments");
  modulate_fwd_kernel<<<ew_grid((long long)B * 15 * N), 256, 0, as_stream(stream)>>>(raw, sim, loc, out, B, N);
  DCNET_LAUNCH_OK("modulate_conf_fwd");
  return 0;
}

extern "C" int dcnet_modulate_conf_bwd(const float* raw, const float* sim, const float* loc, const float* dout,
                                       float* draw, float* dsim, float* dloc, int B, int N, void* stream) {
  DCNET_CHECK_ARG(raw && sim && loc && dout && draw && dsim && dloc && B > 0 && N > 0, "modulate_conf_bwd: bad arguments");
  modulate_bwd_kernel<<<ceil_div((long long)B * N, 256), 256, 0, as_stream(stream)>>>(raw, sim, loc, dout, draw, dsim, dloc, B, N);
  DCNET_LAUNCH_OK("modulate_conf_bwd");
  return 0;
}

extern "C" int dcnet_ground_loss_fwd(const float* pred0, const float* pred1, const float* pred2,
                                     const float* sim0, const float* sim1, const float* sim2,
                                     const float* neg0, const float* neg1, const float* neg2,
                                     const float* loc0, const float* loc1, const float* loc2,
                                     const long long* best_n, const long long* gi, const long long* gj, const float* t5,
                                     const long long* partner3,
                                     int B, int g0, float w_coord, float margin, float* losses, float* lse_conf, float* lse_loc,
                                     void* stream) {
  DCNET_CHECK_ARG(pred0 && pred1 && pred2 && sim0 && sim1 && sim2 && neg0 && neg1 && neg2 && loc0 && loc1 && loc2, "ground_loss_fwd: null input");
  DCNET_CHECK_ARG(best_n && gi && gj && t5 && losses && lse_conf && lse_loc && B > 0 && g0 > 0, "ground_loss_fwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  LossIn a{{{pred0, pred1, pred2}}, {{sim0, sim1, sim2}}, {{neg0, neg1, neg2}}, {{loc0, loc1, loc2}}, best_n, gi, gj, t5, partner3, B, g0, w_coord, margin};
  DCNET_CUDA(cudaMemsetAsync(losses, 0, 3 * sizeof(float), st), "ground_loss_fwd.memset");
  ground_loss_fwd_kernel<<<B, 256, 0, st>>>(a, losses, lse_conf, lse_loc);
  DCNET_LAUNCH_OK("ground_loss_fwd");
  return 0;
}

extern "C" int dcnet_ground_loss_bwd(const float* pred0, const float* pred1, const float* pred2,
                                     const float* sim0, const float* sim1, const float* sim2,
                                     const float* neg0, const float* neg1, const float* neg2,
                                     const float* loc0, const float* loc1, const float* loc2,
                                     const long long* best_n, const long long* gi, const long long* gj, const float* t5,
                                     const long long* partner3,
                                     int B, int g0, float w_coord, float margin, const float* lse_conf, const float* lse_loc,
                                     const float* gl,
                                     float* dpred0, float* dpred1, float* dpred2, float* dsim0, float* dsim1, float* dsim2,
                                     float* dneg0, float* dneg1, float* dneg2, float* dloc0, float* dloc1, float* dloc2,
                                     void* stream) {
  DCNET_CHECK_ARG(pred0 && pred1 && pred2 && sim0 && sim1 && sim2 && neg0 && neg1 && neg2 && loc0 && loc1 && loc2, "ground_loss_bwd: null input");
  DCNET_CHECK_ARG(best_n && gi && gj && t5 && lse_conf && lse_loc && gl && B > 0 && g0 > 0, "ground_loss_bwd: bad arguments");
  DCNET_CHECK_ARG(dpred0 && dpred1 && dpred2 && dsim0 && dsim1 && dsim2 && dneg0 && dneg1 && dneg2 && dloc0 && dloc1 && dloc2, "ground_loss_bwd: null output");
  LossIn a{{{pred0, pred1, pred2}}, {{sim0, sim1, sim2}}, {{neg0, neg1, neg2}}, {{loc0, loc1, loc2}}, best_n, gi, gj, t5, partner3, B, g0, w_coord, margin};
  ground_loss_bwd_kernel<<<dim3(B, GL_BWD_SPLIT), 256, 0, as_stream(stream)>>>(a, lse_conf, lse_loc, gl, MPtr3{{dpred0, dpred1, dpred2}}, MPtr3{{dsim0, dsim1, dsim2}},
                                                             MPtr3{{dneg0, dneg1, dneg2}}, MPtr3{{dloc0, dloc1, dloc2}});
  DCNET_LAUNCH_OK("ground_loss_bwd");
  return 0;
}

extern "C" int dcnet_decode(const float* pred0, const float* pred1, const float* pred2, int B, int g0, int size,
                            float anchor_imsize, const float* h_anchors9x2, int mode,
                            long long* best_n, long long* gi, long long* gj, float* boxes, const float* target, float* iou,
                            void* stream) {
  DCNET_CHECK_ARG(pred0 && pred1 && pred2 && h_anchors9x2 && best_n && gi && gj && boxes && B > 0 && g0 == size / 32 && (mode == 0 || mode == 1),
                  "decode: bad arguments");
  const Anchors9 an = scale_anchors9(h_anchors9x2, size, anchor_imsize);
  decode_kernel<<<B, 256, 0, as_stream(stream)>>>(Ptr3{{pred0, pred1, pred2}}, B, g0, an, mode, best_n, gi, gj, boxes, target, iou);
  DCNET_LAUNCH_OK("decode");
  return 0;
}

extern "C" int dcnet_topk_boxes(const float* pred0, const float* pred1, const float* pred2,
                                const float* feat0, const float* feat1, const float* feat2, int B, int g0, int size, int C, int k,
                                float anchor_imsize, const float* h_anchors9x2, const float* meta,
                                float* boxes, float* scores, long long* cells, float* feats, void* stream) {
  DCNET_CHECK_ARG(pred0 && pred1 && pred2 && feat0 && feat1 && feat2 && h_anchors9x2 && meta && boxes && scores && cells && feats,
                  "topk_boxes: null argument");
  DCNET_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && g0 == size / 32 && k >= 1 && k <= 63 * g0 * g0, "topk_boxes: bad sizes (k <= 3 sum N)");
  const Anchors9 an = scale_anchors9(h_anchors9x2, size, anchor_imsize);
  topk_boxes_kernel<<<B, 256, 0, as_stream(stream)>>>(Ptr3{{pred0, pred1, pred2}}, Ptr3{{feat0, feat1, feat2}}, B, g0, C, k, an, meta, boxes,
                                                      scores, cells, feats);
  DCNET_LAUNCH_OK("topk_boxes");
  return 0;
}

extern "C" int dcnet_post_rescore(const float* centre, const float* ref, const float* ref_score, const int* invalid, int k, int R, int C,
                                  float* fused, long long* best, long long* match, void* stream) {
  DCNET_CHECK_ARG(centre && ref && ref_score && fused && best && k >= 1 && k <= 16 && R >= 1 && R <= 16 && C > 0, "post_rescore: bad arguments (k, R <= 16)");
  post_rescore_kernel<<<1, 256, 0, as_stream(stream)>>>(centre, ref, ref_score, invalid, k, R, C, fused, best, match);
  DCNET_LAUNCH_OK("post_rescore");
  return 0;
}

extern "C" int dcnet_bbox_iou(const float* b1, const float* b2, int n, int x1y1x2y2, float* iou, void* stream) {
  DCNET_CHECK_ARG(b1 && b2 && iou && n >= 0, "bbox_iou: bad arguments");
  if (n == 0) return 0;
  bbox_iou_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(b1, b2, n, x1y1x2y2, iou);
  DCNET_LAUNCH_OK("bbox_iou");
  return 0;
}

extern "C" int dcnet_yolo_layer_decode(const float* x, float* out, int B, int A, int nc, int g, float image_dim,
                                       const float* h_anchorsAx2, void* stream) {
  DCNET_CHECK_ARG(x && out && h_anchorsAx2 && B > 0 && A > 0 && A <= 8 && nc >= 0 && 5 + nc <= YMAX_ATTR && g > 0, "yolo_layer_decode: bad arguments");
  DCNET_CHECK_ARG(B <= 65535, "yolo_layer_decode: B too large");
  AnchorsA an;
  for (int a = 0; a < A; a++) {
    const double div = 416.0 / (double)g;                         // model/darknet.py:287
    an.w[a] = (float)((double)h_anchorsAx2[2 * a] / div);
    an.h[a] = (float)((double)h_anchorsAx2[2 * a + 1] / div);
  }
  const float stride = (float)((double)image_dim / (double)g);   // :266
  dim3 grid(ceil_div(g * g, YCELLS), A, B);
  const int nattr = 5 + nc;
  const size_t smem = (size_t)nattr * (YCELLS + 1) * sizeof(float);
  if (nattr == 85) {
    yolo_decode_kernel<85><<<grid, 256, smem, as_stream(stream)>>>(x, out, A, nattr, g, stride, an);
  } else {
    DCNET_CHECK_ARG(smem <= 48 * 1024, "yolo_layer_decode: %d attributes per anchor exceed the shared-memory tile", nattr);
    yolo_decode_kernel<0><<<grid, 256, smem, as_stream(stream)>>>(x, out, A, nattr, g, stride, an);
  }
  DCNET_LAUNCH_OK("yolo_layer_decode");
  return 0;
}

extern "C" int dcnet_iou_loss_sums(const float* x, const float* t, long long n, float* acc2, void* stream) {
  DCNET_CHECK_ARG(x && t && acc2 && n >= 0, "iou_loss_sums: bad arguments");
  if (n == 0) return 0;
  iou_loss_sums_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, t, n, acc2);
  DCNET_LAUNCH_OK("iou_loss_sums");
  return 0;
}

extern "C" int dcnet_iou_loss_bwd(const float* x, const float* t, long long n, const float* acc2, const float* g, float gscale, float* dx,
                                  void* stream) {
  DCNET_CHECK_ARG(x && t && acc2 && dx && n >= 0, "iou_loss_bwd: bad arguments");
  if (n == 0) return 0;
  iou_loss_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, t, n, acc2, g, gscale, dx);
  DCNET_LAUNCH_OK("iou_loss_bwd");
  return 0;
}
