// a11: cross-modal block (model/DCNet_model.py:625-637 and Crossmodal_corrspondence :41-112).
#include "common.cuh"

namespace {

// y = x / max(||x||_2, 1e-12) over the innermost axis; one warp per row.
__global__ void rownorm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ nrm, long long R, int L) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* p = x + r * L;
  float s = 0.f;
  for (int i = lane; i < L; i += 32) s = fmaf(p[i], p[i], s);
  const float n = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  const float inv = 1.f / n;
  for (int i = lane; i < L; i += 32) y[r * L + i] = p[i] * inv;
  if (lane == 0) nrm[r] = n;
}

// dx = (dy - y <dy,y>) / nrm
__global__ void rownorm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ nrm, const float* __restrict__ dy,
                                   float* __restrict__ dx, long long R, int L) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float d = 0.f;
  for (int i = lane; i < L; i += 32) d = fmaf(dy[r * L + i], y[r * L + i], d);
  d = warp_sum(d);
  const float n = nrm[r];
  // when the norm was clamped (||x|| < eps) y = x/eps is linear in x
  const float proj = (n > 1e-12f) ? d : 0.f;
  const float inv = 1.f / n;
  for (int i = lane; i < L; i += 32) dx[r * L + i] = (dy[r * L + i] - y[r * L + i] * proj) * inv;
}

// lag[b,t,c] = context[b,t,2c] / max(sqrt(sum_t context[b,t,2c]^2), 1e-12);  one thread per (b,c)
__global__ void lagnorm_fwd_kernel(const float* __restrict__ ctx, float* __restrict__ lag, float* __restrict__ nrm, int B, int T, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const float* p = ctx + (long long)b * T * 2 * C + 2 * c;
  float s = 0.f;
  for (int t = 0; t < T; t++) { const float v = p[(long long)t * 2 * C]; s = fmaf(v, v, s); }
  const float n = fmaxf(sqrtf(s), 1e-12f);
  const float inv = 1.f / n;
  for (int t = 0; t < T; t++) lag[((long long)b * T + t) * C + c] = p[(long long)t * 2 * C] * inv;
  nrm[i] = n;
}

__global__ void lagnorm_bwd_kernel(const float* __restrict__ lag, const float* __restrict__ nrm, const float* __restrict__ dlag,
                                   float* __restrict__ dctx, int B, int T, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float d = 0.f;
  for (int t = 0; t < T; t++) {
    const long long o = ((long long)b * T + t) * C + c;
    d = fmaf(dlag[o], lag[o], d);
  }
  const float n = nrm[i];
  const float proj = (n > 1e-12f) ? d : 0.f;
  const float inv = 1.f / n;
  float* q = dctx + (long long)b * T * 2 * C + 2 * c;
  for (int t = 0; t < T; t++) {
    const long long o = ((long long)b * T + t) * C + c;
    q[(long long)t * 2 * C] = (dlag[o] - lag[o] * proj) * inv;
    q[(long long)t * 2 * C + 1] = 0.f;   // odd channels are dropped by the nearest 0.5x down-sample
  }
}

// word[b,n] = first arg-max over t of softmax_t( bias[t] + sum_{t',k} w[t,t',k] M[b,t',n+k-1] )
constexpr int MAXT = 32;
__global__ void words_kernel(const float* __restrict__ M, const float* __restrict__ w, const float* __restrict__ bias,
                             long long* __restrict__ word, int B, int T, int N0) {
  extern __shared__ float sw[];   // T*T*3 weights + T bias
  for (int i = threadIdx.x; i < T * T * 3; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < T; i += blockDim.x) sw[T * T * 3 + i] = bias[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N0) return;
  const int b = i / N0, n = i % N0;
  const float* Mb = M + (long long)b * T * N0;
  float o[MAXT];
  for (int t = 0; t < T; t++) o[t] = sw[T * T * 3 + t];
  for (int tp = 0; tp < T; tp++) {
    const float xm = (n > 0) ? Mb[(long long)tp * N0 + n - 1] : 0.f;
    const float x0 = Mb[(long long)tp * N0 + n];
    const float xp = (n + 1 < N0) ? Mb[(long long)tp * N0 + n + 1] : 0.f;
    for (int t = 0; t < T; t++) {
      const float* ww = sw + (t * T + tp) * 3;
      o[t] = fmaf(ww[0], xm, fmaf(ww[1], x0, fmaf(ww[2], xp, o[t])));
    }
  }
  float m = o[0];
  for (int t = 1; t < T; t++) m = fmaxf(m, o[t]);
  float s = 0.f;
  for (int t = 0; t < T; t++) { o[t] = expf(o[t] - m); s += o[t]; }
  int best = 0;
  float bv = o[0] / s;
  for (int t = 1; t < T; t++) {
    const float v = o[t] / s;
    if (v > bv) { bv = v; best = t; }
  }
  word[i] = best;
}

}  // namespace

extern "C" int dcnet_rownorm_fwd(const float* x, float* y, float* nrm, long long R, int L, void* stream) {
  DCNET_CHECK_ARG(x && y && nrm && R >= 0 && L > 0, "rownorm_fwd: bad arguments");
  if (R == 0) return 0;
  rownorm_fwd_kernel<<<ceil_div(R * 32, 256), 256, 0, as_stream(stream)>>>(x, y, nrm, R, L);
  DCNET_LAUNCH_OK("rownorm_fwd");
  return 0;
}

extern "C" int dcnet_rownorm_bwd(const float* y, const float* nrm, const float* dy, float* dx, long long R, int L, void* stream) {
  DCNET_CHECK_ARG(y && nrm && dy && dx && R >= 0 && L > 0, "rownorm_bwd: bad arguments");
  if (R == 0) return 0;
  rownorm_bwd_kernel<<<ceil_div(R * 32, 256), 256, 0, as_stream(stream)>>>(y, nrm, dy, dx, R, L);
  DCNET_LAUNCH_OK("rownorm_bwd");
  return 0;
}

extern "C" int dcnet_lagnorm_fwd(const float* context, float* lag, float* nrm, int B, int T, int C, void* stream) {
  DCNET_CHECK_ARG(context && lag && nrm && B > 0 && T > 0 && C > 0, "lagnorm_fwd: bad arguments");
  lagnorm_fwd_kernel<<<ceil_div((long long)B * C, 256), 256, 0, as_stream(stream)>>>(context, lag, nrm, B, T, C);
  DCNET_LAUNCH_OK("lagnorm_fwd");
  return 0;
}

extern "C" int dcnet_lagnorm_bwd(const float* lag, const float* nrm, const float* dlag, float* dcontext, int B, int T, int C, void* stream) {
  DCNET_CHECK_ARG(lag && nrm && dlag && dcontext && B > 0 && T > 0 && C > 0, "lagnorm_bwd: bad arguments");
  lagnorm_bwd_kernel<<<ceil_div((long long)B * C, 256), 256, 0, as_stream(stream)>>>(lag, nrm, dlag, dcontext, B, T, C);
  DCNET_LAUNCH_OK("lagnorm_bwd");
  return 0;
}

extern "C" int dcnet_crossmodal_words(const float* lag, const float* vit, const float* fm_w, const float* fm_b,
                                      int fm_cout, int fm_cin, float* M, long long* word, int B, int T, int C, int N0, void* stream) {
  DCNET_CHECK_ARG(lag && vit && fm_w && fm_b && M && word && B > 0 && T > 0 && T <= MAXT && C > 0 && N0 > 0, "crossmodal_words: bad arguments (T<=32)");
  // the kernel indexes fm_w as [T][T][3]: like the reference's Conv1d(20, 20, 3) (model/DCNet_model.py:288, :635) any other sentence length is an error
  DCNET_CHECK_ARG(fm_cout == T && fm_cin == T, "crossmodal_words: feature_map is Conv1d(%d -> %d, k=3) but the batch has T=%d words (the reference requires T == in_channels)", fm_cin, fm_cout, T);
  cudaStream_t st = as_stream(stream);
  // M[b] (T x N0) = lag[b] (T x C) . vit[b] (C x N0)
  DCNET_TRY(sgemm_launch(lag, vit, M, T, N0, C, B, 1, C, 1, (long long)T * C, 0, N0, 1, (long long)C * N0, 0, N0, 1, (long long)T * N0,
                         nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 0, st));
  const size_t sh = (size_t)(T * T * 3 + T) * sizeof(float);
  words_kernel<<<ceil_div((long long)B * N0, 128), 128, sh, st>>>(M, fm_w, fm_b, word, B, T, N0);
  DCNET_LAUNCH_OK("crossmodal_words");
  return 0;
}
