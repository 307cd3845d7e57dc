// a1/a2/a6/a8: 1x1 conv + BatchNorm + ReLU (+ L2 norm over channels, + pixel-to-text dots), forward and backward.
// Replaces ConvBatchNormReLU (model/darknet.py:118-156) + F.normalize(dim=1) (model/DCNet_model.py:359,469) and the
// sim_score products (model/DCNet_model.py:530-535, train_DCNet.py:623-627).
#include <cuda_fp16.h>

#include "common.cuh"

#ifndef BNF_CPT
#define BNF_CPT 32      // channels per thread of bn_act_fwd at C = 512 (32 -> 16 groups, 512-thread CTAs; 64 -> 8 groups)
#endif

namespace {

inline int ew_grid(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

// ------------------------------------------------------------------------------------------------------
// z[b,c,n] = u[b,c] + cc[c,n]   (text + coordinate terms of the split-weight fusion)
// ------------------------------------------------------------------------------------------------------
__global__ void init_bias_kernel(float* __restrict__ z, const float* __restrict__ u, const float* __restrict__ cc,
                                 int B, int C, int N) {
  const long long total = (long long)B * C * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long bc = i / N;
    const int c = (int)(bc % C);
    float v = 0.f;
    if (u) v += u[bc];
    if (cc) v += cc[(long long)c * N + n];
    z[i] = v;
  }
}

// du[b,c] = sum_n dz[b,c,n]   (one warp per row)
__global__ void rowsum_kernel(const float* __restrict__ dz, float* __restrict__ du, long long rows, int N) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = dz + row * N;
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += p[n];
  s = warp_sum(s);
  if (lane == 0) du[row] = s;
}

// dcc[c,n] = sum_b dz[b,c,n]
__global__ void batchsum_kernel(const float* __restrict__ dz, float* __restrict__ dcc, int B, long long CN) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < CN; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; b++) s += dz[(long long)b * CN + i];
    dcc[i] = s;
  }
}

// both sums in ONE pass over dz (N % 4 == 0): CTA = (channel c, 1024 positions), thread = 4 consecutive positions, loop over the images.
// dcc leaves as float4; the row sums go through a warp reduction and one atomicAdd per (image, warp) into du (caller zeroes).
__global__ void __launch_bounds__(256) rowbatchsum_kernel(const float* __restrict__ dz, float* __restrict__ du, float* __restrict__ dcc, int B, int C,
                                                          int N) {
  const int c = blockIdx.y;
  const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
  const bool ok = n < N;
  const int lane = threadIdx.x & 31;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; b++) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = *reinterpret_cast<const float4*>(dz + ((long long)b * C + c) * N + n);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    const float r = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) atomicAdd(du + (long long)b * C + c, r);
  }
  if (ok) *reinterpret_cast<float4*>(dcc + (long long)c * N + n) = acc;
}

// ------------------------------------------------------------------------------------------------------
// BatchNorm batch statistics: one CTA per channel, two passes (mean, then centred second moment) so the
// variance has no E[z^2]-E[z]^2 cancellation (fp32 path: <= 1e-5 relative).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, int B, int C, int N, float eps, float momentum,
                                                       float* __restrict__ mean, float* __restrict__ invstd,
                                                       float* __restrict__ rmean, float* __restrict__ rvar, long long* __restrict__ nbt) {
  __shared__ float sh[32];
  const int c = blockIdx.x;
  const long long M = (long long)B * N;
  if (nbt && c == 0 && threadIdx.x == 0) *nbt += 1;
  float s = 0.f;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) {
    const long long b = i / N;
    const int n = (int)(i - b * N);
    s += z[(b * C + c) * N + n];
  }
  const float mu = block_sum(s, sh) / (float)M;
  float q = 0.f;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) {
    const long long b = i / N;
    const int n = (int)(i - b * N);
    const float d = z[(b * C + c) * N + n] - mu;
    q = fmaf(d, d, q);
  }
  const float ssq = block_sum(q, sh);
  if (threadIdx.x == 0) {
    const float var = ssq / (float)M;
    mean[c] = mu;
    invstd[c] = 1.0f / sqrtf(var + eps);
    if (rmean) {
      const float unb = (M > 1) ? ssq / (float)(M - 1) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * unb;
    }
  }
}

// sum / sum of squares per channel (exact-fp32 path feeding dcnet_bn_finalize); one CTA per channel
__global__ void __launch_bounds__(256) chan_sums_kernel(const float* __restrict__ z, int B, int C, int N, float* __restrict__ sums) {
  __shared__ float sh[32];
  const int c = blockIdx.x;
  const long long M = (long long)B * N;
  float s = 0.f, q = 0.f;
  for (long long i = threadIdx.x; i < M; i += blockDim.x) {
    const long long b = i / N;
    const float v = z[(b * C + c) * N + (i - b * N)];
    s += v;
    q = fmaf(v, v, q);
  }
  s = block_sum(s, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) { sums[c] = s; sums[C + c] = q; }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, float count, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ rmean, float* __restrict__ rvar,
                                   long long* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (nbt && c == 0) *nbt += 1;
  const float mu = sums[c] / count;
  const float var = fmaxf(sums[C + c] / count - mu * mu, 0.f);
  mean[c] = mu;
  invstd[c] = 1.0f / sqrtf(var + eps);
  if (rmean) {
    const float unb = count > 1.f ? var * count / (count - 1.f) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * unb;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rmean, const float* __restrict__ rvar, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    mean[c] = rmean[c];
    invstd[c] = 1.0f / sqrtf(rvar[c] + eps);
  }
}

// ------------------------------------------------------------------------------------------------------
// y = act(bn(z)) (+ channel L2 norm, + text dots).  CTA = 32 positions x 8 channel groups (256 threads); a warp
// holds 32 consecutive positions of one channel -> every global access is a full 128-B line.  Each thread keeps
// its CPT = C/8 channel values of one position in registers, so z is read once and y written once.
// ------------------------------------------------------------------------------------------------------
template <int CPT, int G>
__global__ void __launch_bounds__(32 * G) bn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float slope, int l2norm,
                                                         float* __restrict__ y, const float* __restrict__ fa, const float* __restrict__ fa_neg,
                                                         float* __restrict__ sim, float* __restrict__ neg_sim, int B, int N,
                                                         __half* __restrict__ st16, int ld16, float* __restrict__ st_normsq,
                                                         unsigned int* __restrict__ st_maxnorm) {
  constexpr int C = CPT * G;
  __shared__ float s_scale[C], s_shift[C], s_fa[C], s_fr[C];
  __shared__ float red[3][G][33];
  const bool rn = (l2norm & DCNET_RN_TF32) != 0;      // y leaves rounded to the nearest tf32 (it feeds a tf32 contraction)
  l2norm &= 1;
  const int b = blockIdx.y;
  const int pl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + pl;
  for (int c = threadIdx.x; c < C; c += 32 * G) {
    const float sc = gamma[c] * invstd[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - mean[c] * sc;
    if (fa) {
      s_fa[c] = fa[(long long)b * C + c];
      s_fr[c] = fa_neg ? fa_neg[(long long)b * C + c] : fa[(long long)(B - 1 - b) * C + c];
    }
  }
  __syncthreads();
  const bool valid = n < N;
  const float* zp = z + (long long)b * C * N + n;
  float v[CPT];
  float ss = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
  for (int i = 0; i < CPT; i++) {
    const int c = g + G * i;
    float a = valid ? zp[(long long)c * N] : 0.f;
    a = fmaf(a, s_scale[c], s_shift[c]);
    a = a > 0.f ? a : a * slope;
    v[i] = a;
    ss = fmaf(a, a, ss);
    if (fa) {
      d1 = fmaf(a, s_fa[c], d1);
      d2 = fmaf(a, s_fr[c], d2);
    }
  }
  float inv = 1.f;
  if (l2norm || fa || st16) {
    red[0][g][pl] = ss;
    red[1][g][pl] = d1;
    red[2][g][pl] = d2;
    __syncthreads();
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < G; k++) {
      t0 += red[0][k][pl];
      t1 += red[1][k][pl];
      t2 += red[2][k][pl];
    }
    if (l2norm) inv = 1.f / fmaxf(sqrtf(t0), 1e-12f);
    if (fa && g == 0 && valid) {
      sim[(long long)b * N + n] = t1 * inv;
      neg_sim[(long long)b * N + n] = t2 * inv;
    }
    if (st16 && g == 0) {
      // the co-attention staging of this map (umma_coattn.cu): squared column norms and the largest norm of the frame
      const float nsq = valid ? t0 * inv * inv : 0.f;
      if (valid) st_normsq[(long long)b * N + n] = nsq;
      const float mx = warp_max(nsq);
      if (pl == 0) atomicMax(st_maxnorm + b, __float_as_uint(sqrtf(mx)));      // non-negative floats order like their bit patterns
    }
  }
  if (valid) {
    float* yp = y + (long long)b * C * N + n;
#pragma unroll
    for (int i = 0; i < CPT; i++) {
      const float o = rn ? tf32_rn(v[i] * inv) : v[i] * inv;
      yp[(long long)(g + G * i) * N] = o;
      v[i] = o;
    }
    if (st16) {
      // fp16 copy for the fused co-attention forward: a tf32-rounded value converts exactly (both keep 11 significant bits)
      __half* sp = st16 + (long long)b * C * ld16 + n;
#pragma unroll
      for (int i = 0; i < CPT; i++) sp[(long long)(g + G * i) * ld16] = __float2half_rn(v[i]);
    }
  }
}

// Sum over the 32 lanes of a warp of a per-lane vector v[0..CPT): recursive halving -- at every step a lane keeps one half of
// its current values and hands the other half to its partner, so the whole reduction costs CPT-2 shuffles instead of 5*CPT.
// On return lane L holds in v[0..CPT/32) the totals of original indices  vec_reduce_index(L) + {0 .. CPT/32-1}.
template <int CPT>
__device__ __forceinline__ void warp_vec_reduce(float (&v)[CPT], int lane) {
  // CPT >= 32: pure recursive halving.  CPT < 32: halving until one value is left, then plain butterflies for the remaining
  // lane bits (lanes that differ only in those bits end with the same total).
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const int bit = 16 >> s;
    const int h = (CPT >> 1) >> s;
    if (h >= 1) {
      const bool up = (lane & bit) != 0;
#pragma unroll
      for (int k = 0; k < h; k++) {
        const float keep = up ? v[h + k] : v[k];
        const float send = up ? v[k] : v[h + k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], bit);
    }
  }
}
template <int CPT>
__device__ __forceinline__ int vec_reduce_index(int lane) {
  int idx = 0;
#pragma unroll
  for (int s = 0; s < 5; s++) {
    const int bit = 16 >> s;
    const int h = (CPT >> 1) >> s;
    if (h >= 1) idx += ((lane & bit) ? h : 0);
  }
  return idx;
}
// number of totals a lane holds after warp_vec_reduce, and whether this lane is the one that publishes them
template <int CPT> struct VecOut { static constexpr int N = CPT >= 32 ? CPT / 32 : 1; };
template <int CPT>
__device__ __forceinline__ bool vec_reduce_owner(int lane) { return CPT >= 32 || (lane & (32 / (CPT < 32 ? CPT : 32) - 1)) == 0; }

// backward, phase 1 (see header).  Same tiling; recomputes the forward from z.
// G channel groups x 32 positions per CTA (32*G threads); each thread owns CPT = C/G channels of one position.  G = 16 keeps the
// per-thread arrays at 2 x 32 registers so two 512-thread CTAs fit an SM (the 8-group variant needed 255 registers -> 8 warps/SM).
template <int CPT, int G>
__global__ void __launch_bounds__(32 * G) bn_act_bwd_reduce_kernel(
    const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, float slope, int l2norm, const float* __restrict__ dy, const float* __restrict__ fa,
    const float* __restrict__ fa_neg, const float* __restrict__ dsim, const float* __restrict__ dneg, float* __restrict__ dv,
    float* __restrict__ sum_dv, float* __restrict__ sum_dvz, float* __restrict__ dfa, float* __restrict__ dfa_neg, int B, int N) {
  constexpr int C = CPT * G;
  static_assert(CPT == 16 || CPT == 32 || CPT == 64, "CPT");
  __shared__ float s_scale[C], s_shift[C], s_fa[C], s_fr[C], s_mean[C], s_istd[C];
  __shared__ float red[2][G][33];
  const int b = blockIdx.y;
  const int pl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + pl;
  for (int c = threadIdx.x; c < C; c += 32 * G) {
    const float sc = gamma[c] * invstd[c];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - mean[c] * sc;
    s_mean[c] = mean[c];
    s_istd[c] = invstd[c];
    if (fa) {
      s_fa[c] = fa[(long long)b * C + c];
      s_fr[c] = fa_neg ? fa_neg[(long long)b * C + c] : fa[(long long)(B - 1 - b) * C + c];
    }
  }
  __syncthreads();
  const bool valid = n < N;
  const long long base = (long long)b * C * N + n;
  const float ds = (fa && dsim && valid) ? dsim[(long long)b * N + n] : 0.f;
  const float dn = (fa && dneg && valid) ? dneg[(long long)b * N + n] : 0.f;
  float a[CPT], gr[CPT];      // a: pre-activation t, later dpre*zhat;  gr: incoming gradient, later dpre
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int i = 0; i < CPT; i++) {
    const int c = g + G * i;
    float t = valid ? z[base + (long long)c * N] : 0.f;
    t = fmaf(t, s_scale[c], s_shift[c]);
    const float act = t > 0.f ? t : t * slope;
    a[i] = t;
    ss = fmaf(act, act, ss);
    float gg = (valid && dy) ? dy[base + (long long)c * N] : 0.f;
    if (fa) gg = fmaf(s_fa[c], ds, fmaf(s_fr[c], dn, gg));
    gr[i] = gg;
    dot = fmaf(gg, act, dot);
  }
  float inv = 1.f, dotn = 0.f;
  if (l2norm) {
    red[0][g][pl] = ss;
    red[1][g][pl] = dot;
    __syncthreads();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int k = 0; k < G; k++) {
      t0 += red[0][k][pl];
      t1 += red[1][k][pl];
    }
    inv = rsqrtf(fmaxf(t0, 1e-24f));   // 1 / max(|act|, 1e-12) without the sqrt / divide slow paths (gradient only: 2-ulp MUFU.RSQ)
    dotn = t1 * inv * inv;  // <g, yhat> / nrm, with yhat = act * inv
  }
  const int lane = pl;
  const int ridx = vec_reduce_index<CPT>(lane);
  const bool owner = vec_reduce_owner<CPT>(lane);
  if (fa && dfa) {
    // d fa[b] += sum_n dsim * yhat ; d fa_partner += sum_n dneg * yhat   (one vector at a time: register pressure)
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
      const float w = which == 0 ? ds : dn;
      float f[CPT];
#pragma unroll
      for (int i = 0; i < CPT; i++) {
        const float t = a[i];
        f[i] = valid ? w * ((t > 0.f ? t : t * slope) * inv) : 0.f;
      }
      warp_vec_reduce<CPT>(f, lane);
      if (owner) {
        float* dst = which == 0 ? dfa + (long long)b * C
                                : (fa_neg ? (dfa_neg ? dfa_neg + (long long)b * C : nullptr) : dfa + (long long)(B - 1 - b) * C);
        if (dst) {
#pragma unroll
          for (int k = 0; k < VecOut<CPT>::N; k++) atomicAdd(dst + g + G * (ridx + k), f[k]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CPT; i++) {
    const int c = g + G * i;
    const float t = a[i];
    const float act = t > 0.f ? t : t * slope;
    // d(act): l2norm backward  (g - yhat <g,yhat>) / nrm
    const float da = l2norm ? (gr[i] * inv - act * inv * dotn) : gr[i];
    const float dpre = valid ? da * (t > 0.f ? 1.f : slope) : 0.f;
    if (valid) dv[base + (long long)c * N] = dpre;
    // zhat = (z - mean) * invstd, from z itself so it stays exact when gamma == 0
    const float zh = valid ? (z[base + (long long)c * N] - s_mean[c]) * s_istd[c] : 0.f;
    gr[i] = dpre;
    a[i] = dpre * zh;
  }
  warp_vec_reduce<CPT>(gr, lane);
  warp_vec_reduce<CPT>(a, lane);
  if (owner) {
#pragma unroll
    for (int k = 0; k < VecOut<CPT>::N; k++) {
      const int c = g + G * (ridx + k);
      atomicAdd(sum_dv + c, gr[k]);
      atomicAdd(sum_dvz + c, a[k]);
    }
  }
}

// dz = gamma invstd (dv - mean(dv) - zhat mean(dv zhat)).  Rows = (b, c) pairs, VEC consecutive positions per thread; the channel
// constants are resolved once per thread with 32-bit index arithmetic (no 64-bit division per element).
template <int VEC>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ dv, const float* __restrict__ sum_dv,
                                                               const float* __restrict__ sum_dvz, int train, float* __restrict__ dz,
                                                               int rows, int C, int N, int tpr, float invM) {
  const int row = blockIdx.x * (256 / tpr) + threadIdx.x / tpr;
  const int col = (blockIdx.y * tpr + threadIdx.x % tpr) * VEC;
  if (row >= rows || col >= N || (int)threadIdx.x >= (256 / tpr) * tpr) return;
  const bool rn = (train & DCNET_RN_TF32) != 0;       // dz feeds the two tf32 contractions of the conv backward and nothing else
  train &= 1;
  const int c = row % C;
  const float is = invstd[c];
  const float sc = gamma[c] * is;
  float k0 = 0.f, k1 = 0.f, mu = 0.f;
  if (train) {
    mu = mean[c];
    k0 = sum_dv[c] * invM;
    k1 = sum_dvz[c] * invM;
  }
  const long long off = (long long)row * N + col;
  if constexpr (VEC == 4) {
    float4 d = *reinterpret_cast<const float4*>(dv + off);
    if (train) {
      const float4 zz = *reinterpret_cast<const float4*>(z + off);
      d.x = d.x - k0 - (zz.x - mu) * is * k1;
      d.y = d.y - k0 - (zz.y - mu) * is * k1;
      d.z = d.z - k0 - (zz.z - mu) * is * k1;
      d.w = d.w - k0 - (zz.w - mu) * is * k1;
    }
    float4 o = make_float4(sc * d.x, sc * d.y, sc * d.z, sc * d.w);
    if (rn) o = make_float4(tf32_rn(o.x), tf32_rn(o.y), tf32_rn(o.z), tf32_rn(o.w));
    *reinterpret_cast<float4*>(dz + off) = o;
  } else {
    float d = dv[off];
    if (train) d = d - k0 - (z[off] - mu) * is * k1;
    dz[off] = rn ? tf32_rn(sc * d) : sc * d;
  }
}

// sim[b,n] = <fa[b], x[b,:,n]>, neg[b,n] = <fa_neg[b] (or fa[B-1-b]), x[b,:,n]>; one thread per (b,n), coalesced along n
__global__ void pix2text_kernel(const float* __restrict__ x, const float* __restrict__ fa, const float* __restrict__ fa_neg,
                                float* __restrict__ sim, float* __restrict__ neg, int B, int C, int N) {
  extern __shared__ float sf[];   // fa[b] | partner
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sf[c] = fa[(long long)b * C + c];
    sf[C + c] = fa_neg ? fa_neg[(long long)b * C + c] : fa[(long long)(B - 1 - b) * C + c];
  }
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* xp = x + (long long)b * C * N + n;
  float a = 0.f, d = 0.f;
  for (int c = 0; c < C; c++) {
    const float v = xp[(long long)c * N];
    a = fmaf(v, sf[c], a);
    d = fmaf(v, sf[C + c], d);
  }
  sim[(long long)b * N + n] = a;
  if (neg) neg[(long long)b * N + n] = d;
}

// u[b,c] = sum_k W[c, col_l + k] flang[b,k]: one warp per output channel keeps its weight row in registers (KT = Ct/32 values per lane)
// and walks the batch; flang (B x Ct floats) is read through L1/L2 by every warp.  A [B,C] result with B ~ 32: a GEMM tile grid would
// be 8 CTAs deep in a 512-long reduction (measured 50-75 us for the SIMT GEMM against a few us here).
template <int KT>
__global__ void __launch_bounds__(256) text_term_kernel(const float* __restrict__ W, int ldw, const float* __restrict__ flang,
                                                        float* __restrict__ u, int B, int C) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float w[KT];
#pragma unroll
  for (int i = 0; i < KT; i++) w[i] = W[(long long)c * ldw + lane + 32 * i];
  for (int b = 0; b < B; b++) {
    const float* f = flang + (long long)b * (32 * KT) + lane;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KT; i++) acc = fmaf(w[i], f[32 * i], acc);
    acc = warp_sum(acc);
    if (lane == 0) u[(long long)b * C + c] = acc;
  }
}

// dflang[b,k] = sum_c du[b,c] W[c, col_l + k]: one thread per (b,k), coalesced along k, du[b,:] broadcast
__global__ void __launch_bounds__(256) text_grad_kernel(const float* __restrict__ W, int ldw, const float* __restrict__ du,
                                                        float* __restrict__ dflang, int C, int Ct) {
  const int k = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (k >= Ct) return;
  const float* d = du + (long long)b * C;
  float acc = 0.f;
#pragma unroll 8
  for (int c = 0; c < C; c++) acc = fmaf(d[c], W[(long long)c * ldw + k], acc);
  dflang[(long long)b * Ct + k] = acc;
}

// dWc[c,j] = sum_n dcc[c,n] coord[j,n], j < 8: one warp per channel row, coalesced along n (a [512,8] result over a long reduction:
// a GEMM tile would leave all but 8 CTAs idle)
__global__ void __launch_bounds__(256) coord_wgrad_kernel(const float* __restrict__ dcc, const float* __restrict__ coord, float* __restrict__ dW,
                                                          int ldw, int C, int N) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int n = lane; n < N; n += 32) {
    const float d = dcc[(long long)c * N + n];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = fmaf(d, coord[(long long)j * N + n], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const float v = warp_sum(acc[j]);
    if (lane == 0) dW[(long long)c * ldw + j] = v;
  }
}

__global__ void coord_map_kernel(float* __restrict__ coord, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w) return;
  const float r = (float)(i / w), c = (float)(i % w);
  const float fw = (float)w, fh = (float)h;
  const float x0 = (r * 2.f - fw) / fw, y0 = (c * 2.f - fh) / fh;
  const float x1 = ((r + 1.f) * 2.f - fw) / fw, y1 = ((c + 1.f) * 2.f - fh) / fh;
  const int hw = h * w;
  coord[0 * hw + i] = x0;
  coord[1 * hw + i] = y0;
  coord[2 * hw + i] = x1;
  coord[3 * hw + i] = y1;
  coord[4 * hw + i] = (x0 + x1) / 2.f;
  coord[5 * hw + i] = (y0 + y1) / 2.f;
  coord[6 * hw + i] = 1.f / fh;
  coord[7 * hw + i] = 1.f / fw;
}

}  // namespace

static bool tc_shape_ok(int N, const void* a, const void* b) {
  return (N % 4 == 0) && (reinterpret_cast<uintptr_t>(a) % 16 == 0) && (!b || reinterpret_cast<uintptr_t>(b) % 16 == 0);
}

extern "C" int dcnet_conv1x1_fwd(const float* x1, int K1, const float* x2, int K2, const float* W, int ldw,
                                 const float* u, const float* cc, float* z, int B, int C, int N,
                                 float* stat_sums, int precision, void* stream) {
  DCNET_CHECK_ARG(x1 && W && z && K1 > 0 && B > 0 && C > 0 && N > 0, "conv1x1_fwd: bad arguments");
  DCNET_CHECK_ARG((x2 != nullptr) == (K2 > 0) && ldw >= K1 + K2, "conv1x1_fwd: x2/K2/ldw inconsistent");
  cudaStream_t st = as_stream(stream);
  const bool tc = precision == 1 && tc_shape_ok(N, x1, x2) && tc_shape_ok(N, z, W) && ldw % 4 == 0 && C % 128 == 0 && K1 % 32 == 0 && K2 % 32 == 0;
  if (tc) {
    if (stat_sums) DCNET_CUDA(cudaMemsetAsync(stat_sums, 0, 2 * (size_t)C * sizeof(float), st), "conv1x1_fwd.memset");
    UmmaOperand A{W, C, K1 + K2, ldw, 0, 0, false};
    UmmaOperand Bx{x1, K1, N, N, (long long)K1 * N, B, true};
    UmmaOperand Bx2{x2, K2, N, N, (long long)K2 * N, B, true};
    UmmaEpilogue e{};
    e.out = z; e.ldo = N; e.so_b = (long long)C * N; e.alpha = 1.f; e.u = u; e.ldu = C; e.cc = cc; e.ldcc = N;
    e.sum = stat_sums; e.sumsq = stat_sums ? stat_sums + C : nullptr;
    return umma_gemm(A, Bx, x2 ? &Bx2 : nullptr, C, N, K1 + K2, K1, 0, B, e, st);
  }
  float beta = 0.f;
  if (u || cc) {
    init_bias_kernel<<<ew_grid((long long)B * C * N), 256, 0, st>>>(z, u, cc, B, C, N);
    DCNET_LAUNCH_OK("conv1x1_fwd.init");
    beta = 1.f;
  }
  DCNET_TRY(sgemm_launch(W, x1, z, C, N, K1, B, 1, ldw, 1, 0, 0, N, 1, (long long)K1 * N, 0, N, 1, (long long)C * N,
                         nullptr, nullptr, nullptr, 1.f, beta, nullptr, 0, 0, st));
  if (x2)
    DCNET_TRY(sgemm_launch(W + K1, x2, z, C, N, K2, B, 1, ldw, 1, 0, 0, N, 1, (long long)K2 * N, 0, N, 1, (long long)C * N,
                           nullptr, nullptr, nullptr, 1.f, 1.f, nullptr, 0, 0, st));
  if (stat_sums) {
    chan_sums_kernel<<<C, 256, 0, st>>>(z, B, C, N, stat_sums);
    DCNET_LAUNCH_OK("conv1x1_fwd.sums");
  }
  return 0;
}

extern "C" int dcnet_conv1x1_bwd_data(const float* dz, const float* W, int ldw, float* dx1, int K1, float* dx2, int K2,
                                      int B, int C, int N, int precision, void* stream) {
  return dcnet_conv1x1_bwd_data_absmax(dz, W, ldw, dx1, K1, dx2, K2, B, C, N, precision, nullptr, stream);
}

extern "C" int dcnet_conv1x1_bwd_data_absmax(const float* dz, const float* W, int ldw, float* dx1, int K1, float* dx2, int K2,
                                             int B, int C, int N, int precision, unsigned int* dx2_absmax, void* stream) {
  DCNET_CHECK_ARG(dz && W && B > 0 && C > 0 && N > 0 && ldw >= K1 + K2, "conv1x1_bwd_data: bad arguments");
  cudaStream_t st = as_stream(stream);
  const bool tc = precision == 1 && tc_shape_ok(N, dz, W) && tc_shape_ok(N, dx1, dx2) && ldw % 4 == 0 && C % 32 == 0 && K1 % 128 == 0 &&
                  K2 % 128 == 0;
  if (tc && (dx1 || dx2)) {
    // A = W^T as an MN-major operand: rows = C (reduction), cols = input channel (the M index)
    const int moff = dx1 ? 0 : K1;
    const int M = (dx1 ? K1 : 0) + (dx2 ? K2 : 0);
    UmmaOperand A{W + moff, C, M, ldw, 0, 0, true};
    UmmaOperand Bz{dz, C, N, N, (long long)C * N, B, true};
    UmmaEpilogue e{};
    e.alpha = 1.f;
    if (dx1) { e.out = dx1; e.ldo = N; e.so_b = (long long)K1 * N; }
    if (dx1 && dx2) { e.out2 = dx2; e.ldo2 = N; e.so_b2 = (long long)K2 * N; e.m_split = K1; }
    if (!dx1) { e.out = dx2; e.ldo = N; e.so_b = (long long)K2 * N; }
    if (dx2_absmax) {
      DCNET_CHECK_ARG(dx1 && dx2, "conv1x1_bwd_data: dx2_absmax needs both outputs");
      DCNET_CUDA(cudaMemsetAsync(dx2_absmax, 0, (size_t)B * sizeof(unsigned int), st), "conv1x1_bwd_data.memset");
      e.absmax2 = dx2_absmax;
    }
    return umma_gemm(A, Bz, nullptr, M, N, C, 0, 0, B, e, st);
  }
  DCNET_CHECK_ARG(!dx2_absmax, "conv1x1_bwd_data: dx2_absmax is produced by the tensor-core path only");
  if (dx1)
    DCNET_TRY(sgemm_launch(W, dz, dx1, K1, N, C, B, 1, 1, ldw, 0, 0, N, 1, (long long)C * N, 0, N, 1, (long long)K1 * N,
                           nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 0, st));
  if (dx2)
    DCNET_TRY(sgemm_launch(W + K1, dz, dx2, K2, N, C, B, 1, 1, ldw, 0, 0, N, 1, (long long)C * N, 0, N, 1, (long long)K2 * N,
                           nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 0, st));
  return 0;
}

extern "C" int dcnet_conv1x1_bwd_weight(const float* dz, const float* x1, int K1, const float* x2, int K2,
                                        float* dW, int ldw, float* du, float* dcc, int B, int C, int N, int precision, void* stream) {
  DCNET_CHECK_ARG(dz && B > 0 && C > 0 && N > 0 && ldw >= K1 + K2, "conv1x1_bwd_weight: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (dW && x1) {
    // the (b,n) reduction is split over CTAs along b; fp32 atomics into the zeroed columns
    DCNET_CUDA(cudaMemset2DAsync(dW, (size_t)ldw * sizeof(float), 0, (size_t)(K1 + K2) * sizeof(float), C, st), "conv1x1_bwd_weight.memset");
    const bool tc = precision == 1 && tc_shape_ok(N, dz, x1) && tc_shape_ok(N, x2, dW) && C % 128 == 0 && K1 % 128 == 0 && K2 % 128 == 0;
    if (tc) {
      UmmaOperand A{dz, C, N, N, (long long)C * N, B, false};
      UmmaOperand Bx{x1, K1, N, N, (long long)K1 * N, B, false};
      UmmaOperand Bx2{x2, K2, N, N, (long long)K2 * N, B, false};
      UmmaEpilogue e{};
      e.out = dW; e.ldo = ldw; e.so_b = 0; e.alpha = 1.f; e.atomic = 1;
      DCNET_TRY(umma_gemm(A, Bx, x2 ? &Bx2 : nullptr, C, K1 + K2, N, 0, K1, B, e, st));
    } else {
      DCNET_TRY(sgemm_launch(dz, x1, dW, C, K1, N, B, 1, N, 1, (long long)C * N, 0, 1, N, (long long)K1 * N, 0, ldw, 1, 0,
                             nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 1, st));
      if (x2)
        DCNET_TRY(sgemm_launch(dz, x2, dW + K1, C, K2, N, B, 1, N, 1, (long long)C * N, 0, 1, N, (long long)K2 * N, 0, ldw, 1, 0,
                               nullptr, nullptr, nullptr, 1.f, 0.f, nullptr, 0, 1, st));
    }
  }
  if (du && dcc && N % 4 == 0 && reinterpret_cast<uintptr_t>(dz) % 16 == 0 && reinterpret_cast<uintptr_t>(dcc) % 16 == 0 && C <= 65535) {
    // the fusion layer wants both: one pass over dz
    DCNET_CUDA(cudaMemsetAsync(du, 0, (size_t)B * C * sizeof(float), st), "conv1x1_bwd_weight.memset");
    rowbatchsum_kernel<<<dim3(ceil_div(N, 1024), C), 256, 0, st>>>(dz, du, dcc, B, C, N);
    DCNET_LAUNCH_OK("conv1x1_bwd_weight.du_dcc");
    return 0;
  }
  if (du) {
    const long long rows = (long long)B * C;
    rowsum_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(dz, du, rows, N);
    DCNET_LAUNCH_OK("conv1x1_bwd_weight.du");
  }
  if (dcc) {
    batchsum_kernel<<<ew_grid((long long)C * N), 256, 0, st>>>(dz, dcc, B, (long long)C * N);
    DCNET_LAUNCH_OK("conv1x1_bwd_weight.dcc");
  }
  return 0;
}

// a8, the text and coordinate terms of the split-weight fusion (model/DCNet_model.py:489-505 concatenates
// [corr_feat | flang tiled over h x w | coord] before the fusing 1x1 conv): with W = [W_v | W_l | W_c],
//   u[b,c]  = sum_k W_l[c,k] flang[b,k]      (a [B,C] bias per image instead of a [B,512,h,w] tile)
//   cc[c,n] = sum_j W_c[c,j] coord[j,n]      (a [C,N] map shared by the batch instead of a [B,8,h,w] tensor)
// W_l = W[:, col_l : col_l+Ct], W_c = W[:, col_c : col_c+8] inside the [C, ldw] weight.  Exact fp32 (CUDA cores): the products are tiny.
extern "C" int dcnet_fuse_terms_fwd(const float* W, int ldw, int col_l, int Ct, int col_c, const float* flang, const float* coord,
                                    float* u, float* cc, int B, int C, int N, void* stream) {
  DCNET_CHECK_ARG(W && flang && u && B > 0 && C > 0 && Ct > 0 && col_l >= 0 && col_l + Ct <= ldw, "fuse_terms_fwd: bad arguments");
  DCNET_CHECK_ARG((coord != nullptr) == (cc != nullptr) && (!coord || (N > 0 && col_c >= 0 && col_c + 8 <= ldw)), "fuse_terms_fwd: coord/cc inconsistent");
  cudaStream_t st = as_stream(stream);
  if (Ct == 512) {
    text_term_kernel<16><<<ceil_div(C, 8), 256, 0, st>>>(W + col_l, ldw, flang, u, B, C);
    DCNET_LAUNCH_OK("fuse_terms_fwd.text");
  } else {
    DCNET_TRY(sgemm_launch(flang, W + col_l, u, B, C, Ct, 1, 1, Ct, 1, 0, 0, 1, ldw, 0, 0, C, 1, 0, nullptr, nullptr, nullptr, 1.f, 0.f,
                           nullptr, 0, 0, st));
  }
  if (coord)
    DCNET_TRY(sgemm_launch(W + col_c, coord, cc, C, N, 8, 1, 1, ldw, 1, 0, 0, N, 1, 0, 0, N, 1, 0, nullptr, nullptr, nullptr, 1.f, 0.f,
                           nullptr, 0, 0, st));
  return 0;
}

// backward of the above from du [B,C] (= sum_n dz) and dcc [C,N] (= sum_b dz): dflang [B,Ct] (or NULL), and the text / coordinate
// columns of dW [C,ldw] (overwritten; dW NULL = no weight gradient).
extern "C" int dcnet_fuse_terms_bwd(const float* W, int ldw, int col_l, int Ct, int col_c, const float* flang, const float* coord,
                                    const float* du, const float* dcc, float* dflang, float* dW, int B, int C, int N, void* stream) {
  DCNET_CHECK_ARG(W && flang && du && B > 0 && C > 0 && Ct > 0 && col_l >= 0 && col_l + Ct <= ldw, "fuse_terms_bwd: bad arguments");
  DCNET_CHECK_ARG((coord != nullptr) == (dcc != nullptr) && (!coord || (N > 0 && col_c >= 0 && col_c + 8 <= ldw)), "fuse_terms_bwd: coord/dcc inconsistent");
  cudaStream_t st = as_stream(stream);
  if (dflang) {
    DCNET_CHECK_ARG(B <= 65535, "fuse_terms_bwd: B too large");
    text_grad_kernel<<<dim3(ceil_div(Ct, 256), B), 256, 0, st>>>(W + col_l, ldw, du, dflang, C, Ct);
    DCNET_LAUNCH_OK("fuse_terms_bwd.text");
  }
  if (dW) {
    DCNET_TRY(sgemm_launch(du, flang, dW + col_l, C, Ct, B, 1, 1, 1, C, 0, 0, Ct, 1, 0, 0, ldw, 1, 0, nullptr, nullptr, nullptr, 1.f, 0.f,
                           nullptr, 0, 0, st));
    if (coord) {
      coord_wgrad_kernel<<<ceil_div(C, 8), 256, 0, st>>>(dcc, coord, dW + col_c, ldw, C, N);
      DCNET_LAUNCH_OK("fuse_terms_bwd.coord");
    }
  }
  return 0;
}

extern "C" int dcnet_bn_finalize(const float* stat_sums, long long count, int C, float eps, float momentum,
                                 float* mean, float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                                 void* stream) {
  DCNET_CHECK_ARG(stat_sums && mean && invstd && count > 0 && C > 0, "bn_finalize: bad arguments");
  DCNET_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running_mean/var must both be given");
  bn_finalize_kernel<<<ceil_div(C, 256), 256, 0, as_stream(stream)>>>(stat_sums, (float)count, C, eps, momentum, mean, invstd, running_mean, running_var,
                                                                          num_batches_tracked);
  DCNET_LAUNCH_OK("bn_finalize");
  return 0;
}

extern "C" int dcnet_bn_stats(const float* z, int B, int C, int N, float eps, float momentum,
                              float* mean, float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                              void* stream) {
  DCNET_CHECK_ARG(z && mean && invstd && B > 0 && C > 0 && N > 0, "bn_stats: bad arguments");
  DCNET_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_stats: running_mean/var must both be given");
  bn_stats_kernel<<<C, 256, 0, as_stream(stream)>>>(z, B, C, N, eps, momentum, mean, invstd, running_mean, running_var, num_batches_tracked);
  DCNET_LAUNCH_OK("bn_stats");
  return 0;
}

extern "C" int dcnet_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                                   float* mean, float* invstd, void* stream) {
  DCNET_CHECK_ARG(running_mean && running_var && mean && invstd && C > 0, "bn_eval_stats: bad arguments");
  bn_eval_stats_kernel<<<ceil_div(C, 256), 256, 0, as_stream(stream)>>>(running_mean, running_var, C, eps, mean, invstd);
  DCNET_LAUNCH_OK("bn_eval_stats");
  return 0;
}

// (two positions per thread in bn_act_fwd was measured SLOWER than one: 126 us against 87 us at C3, profiles/r2o; removed)
int umma_coattn_stage_layout(void* ws, size_t ws_bytes, int F, int C, int N, void** fp16_maps, int* ld, float** normsq, float** maxnorm);

extern "C" int dcnet_bn_act_fwd(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                float slope, int l2norm, float* y, const float* fa, const float* fa_neg, float* sim, float* neg_sim,
                                int B, int C, int N, void* stream) {
  return dcnet_bn_act_fwd_staged(z, mean, invstd, gamma, beta, slope, l2norm, y, fa, fa_neg, sim, neg_sim, B, C, N, nullptr, 0, stream);
}

extern "C" int dcnet_bn_act_fwd_staged(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                       float slope, int l2norm, float* y, const float* fa, const float* fa_neg, float* sim, float* neg_sim,
                                       int B, int C, int N, void* staged, size_t staged_bytes, void* stream) {
  DCNET_CHECK_ARG(z && mean && invstd && gamma && beta && y && B > 0 && N > 0, "bn_act_fwd: bad arguments");
  __half* st16 = nullptr; int ld16 = 0; float* st_normsq = nullptr; float* st_maxnorm = nullptr;
  if (staged) {
    void* maps;
    DCNET_TRY(umma_coattn_stage_layout(staged, staged_bytes, B, C, N, &maps, &ld16, &st_normsq, &st_maxnorm));
    st16 = reinterpret_cast<__half*>(maps);
    DCNET_CUDA(cudaMemsetAsync(st_maxnorm, 0, (size_t)B * sizeof(float), as_stream(stream)), "bn_act_fwd.memset");
  }
  unsigned int* st_mx = reinterpret_cast<unsigned int*>(st_maxnorm);
  DCNET_CHECK_ARG(C == 512 || C == 256, "bn_act_fwd: C=%d unsupported (512 or 256)", C);
  DCNET_CHECK_ARG(!fa || (sim && neg_sim), "bn_act_fwd: fa given without sim/neg_sim outputs");
  DCNET_CHECK_ARG(B <= 65535, "bn_act_fwd: B too large");
  dim3 grid(ceil_div(N, 32), B);
  if (C == 512)
    bn_act_fwd_kernel<BNF_CPT, 512 / BNF_CPT><<<grid, 32 * (512 / BNF_CPT), 0, as_stream(stream)>>>(z, mean, invstd, gamma, beta, slope, l2norm, y, fa, fa_neg, sim, neg_sim, B, N, st16, ld16, st_normsq, st_mx);
  else
    bn_act_fwd_kernel<32, 8><<<grid, 256, 0, as_stream(stream)>>>(z, mean, invstd, gamma, beta, slope, l2norm, y, fa, fa_neg, sim, neg_sim, B, N, st16, ld16, st_normsq, st_mx);
  DCNET_LAUNCH_OK("bn_act_fwd");
  return 0;
}

int bn_bwd_reduce_staged(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                      int l2norm, const float* dy, const float* fa, const float* fa_neg, const float* dsim, const float* dneg,
                      float* dv, float* sum_dv, float* sum_dvz, float* dfa, float* dfa_neg, int B, int Cc, int N, cudaStream_t st);
static bool g_bn_bwd_no_staged = false;
// test / bring-up knob: 1 = always use the register-staged backward kernel
extern "C" int dcnet_bn_bwd_select(int variant) { g_bn_bwd_no_staged = (variant == 1); return 0; }

extern "C" int dcnet_bn_act_bwd_reduce(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                       float slope, int l2norm, const float* dy, const float* fa, const float* fa_neg, const float* dsim,
                                       const float* dneg_sim, float* dv, float* sum_dv, float* sum_dvz, float* dfa, float* dfa_neg,
                                       int B, int C, int N, void* stream) {
  DCNET_CHECK_ARG(z && mean && invstd && gamma && beta && dv && sum_dv && sum_dvz && B > 0 && N > 0, "bn_act_bwd_reduce: bad arguments");
  DCNET_CHECK_ARG(dy || fa, "bn_act_bwd_reduce: no incoming gradient");
  DCNET_CHECK_ARG(C == 512 || C == 256, "bn_act_bwd_reduce: C=%d unsupported (512 or 256)", C);
  if (!g_bn_bwd_no_staged) {
    // persistent smem-staged kernel (bn_bwd_staged.cu) when the maps are TMA-addressable
    const int r = bn_bwd_reduce_staged(z, mean, invstd, gamma, beta, slope, l2norm, dy, fa, fa_neg, dsim, dneg_sim, dv, sum_dv, sum_dvz, dfa,
                                    dfa_neg, B, C, N, as_stream(stream));
    if (r != -2147483647) return r;
  }
  dim3 grid(ceil_div(N, 32), B);
  if (C == 512)
    bn_act_bwd_reduce_kernel<32, 16><<<grid, 512, 0, as_stream(stream)>>>(z, mean, invstd, gamma, beta, slope, l2norm, dy, fa, fa_neg, dsim,
                                                                             dneg_sim, dv, sum_dv, sum_dvz, dfa, dfa_neg, B, N);
  else
    bn_act_bwd_reduce_kernel<32, 8><<<grid, 256, 0, as_stream(stream)>>>(z, mean, invstd, gamma, beta, slope, l2norm, dy, fa, fa_neg, dsim,
                                                                            dneg_sim, dv, sum_dv, sum_dvz, dfa, dfa_neg, B, N);
  DCNET_LAUNCH_OK("bn_act_bwd_reduce");
  return 0;
}

extern "C" int dcnet_bn_act_bwd_apply(const float* z, const float* mean, const float* invstd, const float* gamma,
                                      const float* dv, const float* sum_dv, const float* sum_dvz, int train,
                                      float* dz, int B, int C, int N, void* stream) {
  DCNET_CHECK_ARG(z && mean && invstd && gamma && dv && dz && B > 0 && C > 0 && N > 0, "bn_act_bwd_apply: bad arguments");
  DCNET_CHECK_ARG(!(train & 1) || (sum_dv && sum_dvz), "bn_act_bwd_apply: train mode needs the channel sums");
  const long long rows = (long long)B * C;
  const float invM = 1.f / (float)((long long)B * N);
  const bool v4 = N % 4 == 0 && reinterpret_cast<uintptr_t>(z) % 16 == 0 && reinterpret_cast<uintptr_t>(dv) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(dz) % 16 == 0;
  const int nv = v4 ? N / 4 : N;
  const int tpr = nv < 256 ? nv : 256;                  // threads along a row
  const long long gx = (rows + (256 / tpr) - 1) / (256 / tpr);     // row blocks on grid.x (2^31 limit), column blocks on grid.y
  DCNET_CHECK_ARG(rows < (1ll << 31) && gx < (1ll << 31), "bn_act_bwd_apply: too many rows");
  dim3 grid((unsigned)gx, (nv + tpr - 1) / tpr);
  if (v4)
    bn_act_bwd_apply_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(z, mean, invstd, gamma, dv, sum_dv, sum_dvz, train, dz, (int)rows, C, N, tpr, invM);
  else
    bn_act_bwd_apply_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(z, mean, invstd, gamma, dv, sum_dv, sum_dvz, train, dz, (int)rows, C, N, tpr, invM);
  DCNET_LAUNCH_OK("bn_act_bwd_apply");
  return 0;
}

__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
  }
  if (i == 0)
    for (long long k = n4 * 4; k < n; k++) y[k] = tf32_rn(x[k]);
}

extern "C" int dcnet_round_tf32(const float* x, float* y, long long n, void* stream) {
  DCNET_CHECK_ARG(x && y && n > 0, "round_tf32: bad arguments");
  const bool v4 = reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0;
  const long long n4 = v4 ? n / 4 : 0;
  DCNET_CHECK_ARG(v4 || n < 4096, "round_tf32: unaligned buffers");
  round_tf32_kernel<<<(unsigned)((n4 > 0 ? n4 + 255 : 256) / 256), 256, 0, as_stream(stream)>>>(x, y, n4, n);
  DCNET_LAUNCH_OK("round_tf32");
  return 0;
}

extern "C" int dcnet_pix2text(const float* x, const float* fa, const float* fa_neg, float* sim, float* neg_sim, int B, int C, int N, void* stream) {
  DCNET_CHECK_ARG(x && fa && sim && B > 0 && C > 0 && N > 0 && B <= 65535, "pix2text: bad arguments");
  dim3 grid(ceil_div(N, 128), B);
  pix2text_kernel<<<grid, 128, 2 * (size_t)C * sizeof(float), as_stream(stream)>>>(x, fa, fa_neg, sim, neg_sim, B, C, N);
  DCNET_LAUNCH_OK("pix2text");
  return 0;
}

extern "C" int dcnet_coord_map(float* coord, int h, int w, void* stream) {
  DCNET_CHECK_ARG(coord && h > 0 && w > 0, "coord_map: bad arguments");
  coord_map_kernel<<<ceil_div(h * w, 256), 256, 0, as_stream(stream)>>>(coord, h, w);
  DCNET_LAUNCH_OK("coord_map");
  return 0;
}
