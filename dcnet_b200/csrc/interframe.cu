// a4 (model/DCNet_model.py:381-430), generic column gather/scatter, a12/a13 InfoNCE (train_DCNet.py:114-166).
#include "common.cuh"

namespace {

// Ordering of the reference's sorted top-k with the documented tie rule: larger value first, then LOWER flat index.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return (v > bv) || (v == bv && i < bi); }

// One CTA per pair.  Round r selects the best entry that is strictly after the previous selection in that ordering;
// no "taken" flags, exact for duplicates.  n = N0*N0 <= 28561 at 416^2 -> <= 28 loads per thread per round.
__global__ void __launch_bounds__(1024) topk_kernel(const float* __restrict__ S0, int n, int top_k, long long* __restrict__ idx) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  __shared__ float last_v;
  __shared__ int last_i;
  const float* s = S0 + (long long)blockIdx.x * n;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { last_v = INFINITY; last_i = -1; }
  __syncthreads();
  for (int r = 0; r < top_k; r++) {
    const float lv = last_v;
    const int li = last_i;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = s[i];
      const bool after = (v < lv) || (v == lv && i > li);
      if (after && better(v, i, bv, bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[w] = bv; s_i[w] = bi; }
    __syncthreads();
    if (w == 0) {
      bv = s_v[lane];
      bi = s_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        last_v = bv;
        last_i = bi;
        idx[(long long)blockIdx.x * top_k + r] = bi;
      }
    }
    __syncthreads();
  }
}

__global__ void negidx_kernel(const long long* __restrict__ idx, const int* __restrict__ negpos, int total, int N0, int neg_n,
                              long long* __restrict__ negidx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = (int)(idx[i / neg_n] % N0);
  const int pos = negpos[i];
  negidx[i] = pos + (pos >= col ? 1 : 0);
}

// cols[p, :] = [ idx//N0 (top_k) | idx%N0 (top_k) | negatives (top_k*neg_n) ]: every gather column of a4 in one launch
// rank-major layout: cols = [q: top_k x P | k: top_k x P | neg: top_k x P x neg_n], so the gathered rows are already the packed
// [rank][pair] tensors the contrastive loss consumes (no transposes / stacks afterwards)
__global__ void interframe_cols_kernel(const long long* __restrict__ idx, const int* __restrict__ negpos, int P, int N0, int top_k, int neg_n,
                                       long long* __restrict__ cols) {
  const int nq = top_k * P;
  const int total = nq * (2 + neg_n);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long v;
  if (i < 2 * nq) {
    const int j = i < nq ? i : i - nq;                 // rank * P + pair
    const int r = j / P, p = j - r * P;
    const long long f = idx[(long long)p * top_k + r];
    v = i < nq ? f / N0 : f % N0;
  } else {
    const int j = i - 2 * nq;                          // (rank * P + pair) * neg_n + t
    const int t = j % neg_n, rp = j / neg_n;
    const int r = rp / P, p = rp - r * P;
    const int col = (int)(idx[(long long)p * top_k + r] % N0);
    const int pos = negpos[((long long)p * top_k + r) * neg_n + t];
    v = pos + (pos >= col ? 1 : 0);
  }
  cols[i] = v;
}

// out[i,:] = src[img[i], :, col[i]]   one warp per i
__global__ void gather_cols_kernel(const float* __restrict__ src, const int* __restrict__ img, const long long* __restrict__ col,
                                   int n, float* __restrict__ out, int C, int N) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* p = src + ((long long)img[i] * C) * N + col[i];
  float* o = out + (long long)i * C;
  for (int c = lane; c < C; c += 32) o[c] = __ldg(p + (long long)c * N);
}

__global__ void scatter_cols_add_kernel(const float* __restrict__ dout, const int* __restrict__ img, const long long* __restrict__ col,
                                        int n, float* __restrict__ dsrc, int C, int N) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  float* p = dsrc + ((long long)img[i] * C) * N + col[i];
  const float* o = dout + (long long)i * C;
  for (int c = lane; c < C; c += 32) atomicAdd(p + (long long)c * N, o[c]);
}

// ------------------------------------------------------------------------------------------------------
// InfoNCE.  One warp per group g: q[g], k[g], neg[g][0..n).  C <= 512 (16 channels per lane), n <= 15.
// ------------------------------------------------------------------------------------------------------
constexpr int NCE_MAXN = 16;   // 1 positive + up to 15 negatives
constexpr int NCE_CPL = 16;    // channels per lane (C = 512)

__device__ __forceinline__ void nce_logits(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ neg,
                                           int n, int C, float T, int lane, float* qv, float& nq, float* logit, float* nrm) {
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NCE_CPL; i++) {
    const int c = lane + 32 * i;
    qv[i] = c < C ? q[c] : 0.f;
    sq = fmaf(qv[i], qv[i], sq);
  }
  nq = fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
  for (int j = 0; j <= n; j++) {
    const float* v = (j == 0) ? k : neg + (long long)(j - 1) * C;
    float d = 0.f, s = 0.f;
#pragma unroll
    for (int i = 0; i < NCE_CPL; i++) {
      const int c = lane + 32 * i;
      const float x = c < C ? v[c] : 0.f;
      d = fmaf(qv[i], x, d);
      s = fmaf(x, x, s);
    }
    d = warp_sum(d);
    s = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    nrm[j] = s;
    logit[j] = d / (nq * s) / T;
  }
}

__global__ void infonce_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ neg,
                                   int G, int n, int C, float T, float* __restrict__ rowloss) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= G) return;
  float qv[NCE_CPL], logit[NCE_MAXN], nrm[NCE_MAXN], nq;
  nce_logits(q + (long long)g * C, k + (long long)g * C, neg + (long long)g * n * C, n, C, T, lane, qv, nq, logit, nrm);
  float m = logit[0];
  for (int j = 1; j <= n; j++) m = fmaxf(m, logit[j]);
  float s = 0.f;
  for (int j = 0; j <= n; j++) s += expf(logit[j] - m);
  if (lane == 0) rowloss[g] = m + logf(s) - logit[0];
}

__global__ void infonce_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ neg,
                                   int G, int n, int C, float T, const float* __restrict__ gscale, int gstride,
                                   float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dneg) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= G) return;
  const float* qg = q + (long long)g * C;
  const float* kg = k + (long long)g * C;
  const float* ng = neg + (long long)g * n * C;
  float qv[NCE_CPL], logit[NCE_MAXN], nrm[NCE_MAXN], nq;
  nce_logits(qg, kg, ng, n, C, T, lane, qv, nq, logit, nrm);
  float m = logit[0];
  for (int j = 1; j <= n; j++) m = fmaxf(m, logit[j]);
  float s = 0.f;
  for (int j = 0; j <= n; j++) s += expf(logit[j] - m);
  const float gs = gscale[(long long)g * gstride];
  float dqh[NCE_CPL];   // d loss / d q_hat
#pragma unroll
  for (int i = 0; i < NCE_CPL; i++) dqh[i] = 0.f;
  const float invq = 1.f / nq;
  for (int j = 0; j <= n; j++) {
    const float pj = expf(logit[j] - m) / s;
    const float dl = gs * (pj - (j == 0 ? 1.f : 0.f)) / T;   // d loss / d cos_j
    const float* v = (j == 0) ? kg : ng + (long long)(j - 1) * C;
    float* dvp = (j == 0) ? dk + (long long)g * C : dneg + ((long long)g * n + (j - 1)) * C;
    const float invn = 1.f / nrm[j];
    const float cosj = logit[j] * T;
#pragma unroll
    for (int i = 0; i < NCE_CPL; i++) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float vh = v[c] * invn;
        const float qh = qv[i] * invq;
        dqh[i] = fmaf(dl, vh, dqh[i]);
        dvp[c] = dl * (qh - vh * cosj) * invn;   // through v_hat = v/|v|
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NCE_CPL; i++) dot = fmaf(dqh[i], qv[i] * invq, dot);
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < NCE_CPL; i++) {
    const int c = lane + 32 * i;
    if (c < C) dq[(long long)g * C + c] = (dqh[i] - qv[i] * invq * dot) * invq;
  }
}

}  // namespace

extern "C" int dcnet_interframe_topk(const float* fv0, int P, int C, int N0, int top_k, float* S0, long long* idx, void* stream) {
  DCNET_CHECK_ARG(fv0 && S0 && idx && P >= 0 && C > 0 && N0 > 0 && top_k > 0, "interframe_topk: bad arguments");
  DCNET_CHECK_ARG((long long)N0 * N0 >= top_k, "interframe_topk: N0^2=%lld < top_k=%d (reference needs >= 128x128 inputs)", (long long)N0 * N0, top_k);
  if (P == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const long long CN = (long long)C * N0;
  DCNET_TRY(sgemm_launch(fv0, fv0 + CN, S0, N0, N0, C, P, 1, 1, N0, 2 * CN, 0, N0, 1, 2 * CN, 0, N0, 1, (long long)N0 * N0,
                         nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 0, st));
  topk_kernel<<<P, 1024, 0, st>>>(S0, N0 * N0, top_k, idx);
  DCNET_LAUNCH_OK("interframe_topk");
  return 0;
}

extern "C" int dcnet_interframe_negidx(const long long* idx, const int* negpos, int P, int N0, int top_k, int neg_n,
                                       long long* negidx, void* stream) {
  DCNET_CHECK_ARG(idx && negpos && negidx && P >= 0 && N0 > 1 && top_k > 0 && neg_n > 0, "interframe_negidx: bad arguments");
  const int total = P * top_k * neg_n;
  if (total == 0) return 0;
  negidx_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(idx, negpos, total, N0, neg_n, negidx);
  DCNET_LAUNCH_OK("interframe_negidx");
  return 0;
}

extern "C" int dcnet_interframe_cols(const long long* idx, const int* negpos, int P, int N0, int top_k, int neg_n, long long* cols, void* stream) {
  DCNET_CHECK_ARG(idx && negpos && cols && P >= 0 && N0 > 1 && top_k > 0 && neg_n > 0, "interframe_cols: bad arguments");
  const int total = P * top_k * (2 + neg_n);
  if (total == 0) return 0;
  interframe_cols_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(idx, negpos, P, N0, top_k, neg_n, cols);
  DCNET_LAUNCH_OK("interframe_cols");
  return 0;
}

extern "C" int dcnet_gather_cols(const float* src, const int* img, const long long* col, int n, float* out, int C, int N, void* stream) {
  if (n == 0) return 0;
  DCNET_CHECK_ARG(src && img && col && out && n > 0 && C > 0 && N > 0, "gather_cols: bad arguments");
  gather_cols_kernel<<<ceil_div((long long)n * 32, 256), 256, 0, as_stream(stream)>>>(src, img, col, n, out, C, N);
  DCNET_LAUNCH_OK("gather_cols");
  return 0;
}

extern "C" int dcnet_scatter_cols_add(const float* dout, const int* img, const long long* col, int n, float* dsrc, int C, int N, void* stream) {
  if (n == 0) return 0;
  DCNET_CHECK_ARG(dout && img && col && dsrc && n > 0 && C > 0 && N > 0, "scatter_cols_add: bad arguments");
  scatter_cols_add_kernel<<<ceil_div((long long)n * 32, 256), 256, 0, as_stream(stream)>>>(dout, img, col, n, dsrc, C, N);
  DCNET_LAUNCH_OK("scatter_cols_add");
  return 0;
}

extern "C" int dcnet_infonce_fwd(const float* q, const float* k, const float* neg, int G, int n, int C, float T, float* rowloss, void* stream) {
  if (G == 0) return 0;
  DCNET_CHECK_ARG(q && k && neg && rowloss && G > 0 && n >= 1 && n < NCE_MAXN && C > 0 && C <= 32 * NCE_CPL, "infonce_fwd: bad arguments (n<=15, C<=512)");
  infonce_fwd_kernel<<<ceil_div((long long)G * 32, 128), 128, 0, as_stream(stream)>>>(q, k, neg, G, n, C, T, rowloss);
  DCNET_LAUNCH_OK("infonce_fwd");
  return 0;
}

extern "C" int dcnet_infonce_bwd(const float* q, const float* k, const float* neg, int G, int n, int C, float T, const float* gscale,
                                 int gstride, float* dq, float* dk, float* dneg, void* stream) {
  if (G == 0) return 0;
  DCNET_CHECK_ARG(q && k && neg && gscale && dq && dk && dneg && G > 0 && n >= 1 && n < NCE_MAXN && C > 0 && C <= 32 * NCE_CPL,
                  "infonce_bwd: bad arguments (n<=15, C<=512)");
  infonce_bwd_kernel<<<ceil_div((long long)G * 32, 128), 128, 0, as_stream(stream)>>>(q, k, neg, G, n, C, T, gscale, gstride, dq, dk, dneg);
  DCNET_LAUNCH_OK("infonce_bwd");
  return 0;
}
