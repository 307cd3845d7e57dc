// Location branch at inference (model/DCNet_model.py:556-603, same lines in model/test_DCNet_model.py).
//
// The reference materialises rel[b,p,q] = <e_p, e_q> * obj[b,q]  ([B, SN, SN]: 7.2 MB per image at 256x256, 50 MB at 416x416),
// runs Linear(SN -> C) + BatchNorm1d + ReLU over its rows, L2-normalises over the channels, takes the dot product with the
// location phrase vector and min-max normalises over the positions.  rel has rank 8:
//
//     z[b,p,c] = sum_q W[c,q] <e_p, e_q> obj[b,q] + bias[c] = sum_k G[b,c,k] e[p,k] + bias[c],
//     G[b,c,k] = sum_q W[c,q] obj[b,q] e[q,k]                                   ([B, C, 8])
//
// so the [B,SN,SN] tensor and the SN-long rows of the Linear never exist: one pass over W builds G, one fused kernel does the
// 8-term products, the BN affine (eval statistics folded into scale / shift by the caller), ReLU, the channel norm and the dot
// product per position, and a small kernel min-max normalises each image.  HBM traffic: W once (C*SN*4 B, L2-resident across the
// images) + E + obj + the [B,SN] scores, against 2 * B*SN*SN*4 B for the materialised form.  Forward only (no batch statistics).
#include "common.cuh"

namespace {

constexpr int LOC_K = 8;
constexpr int LOC_BG = 4;          // images per warp in loc_g_kernel
constexpr int LOC_PARTS = 4;       // threads per position in loc_score_kernel (each takes every 4th channel)
constexpr int LOC_PP = 4;          // positions per thread in loc_score_kernel

// G[b,c,:] : one warp per (c, group of 4 images), lanes stride over the positions.  E[q,:] (the largest stream: 8 floats per
// position against 1 of W and 1 of obj per image) is read once per 4 images.
__global__ void loc_g_kernel(const float* __restrict__ E, const float* __restrict__ obj, const float* __restrict__ W, int ldw,
                             float* __restrict__ G, int B, int SN, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  const int b0 = blockIdx.y * LOC_BG;
  if (c >= C) return;
  const float* w = W + (long long)c * ldw;
  const float* o[LOC_BG];
#pragma unroll
  for (int i = 0; i < LOC_BG; i++) o[i] = obj + (long long)min(b0 + i, B - 1) * SN;     // tail images recompute the last one
  float acc[LOC_BG][LOC_K];
#pragma unroll
  for (int i = 0; i < LOC_BG; i++)
#pragma unroll
    for (int k = 0; k < LOC_K; k++) acc[i][k] = 0.f;
#pragma unroll 4
  for (int q = lane; q < SN; q += 32) {
    const float wq = w[q];
    const float4 e0 = *reinterpret_cast<const float4*>(E + (long long)q * LOC_K);
    const float4 e1 = *reinterpret_cast<const float4*>(E + (long long)q * LOC_K + 4);
#pragma unroll
    for (int i = 0; i < LOC_BG; i++) {
      const float s = wq * o[i][q];
      acc[i][0] = fmaf(s, e0.x, acc[i][0]); acc[i][1] = fmaf(s, e0.y, acc[i][1]);
      acc[i][2] = fmaf(s, e0.z, acc[i][2]); acc[i][3] = fmaf(s, e0.w, acc[i][3]);
      acc[i][4] = fmaf(s, e1.x, acc[i][4]); acc[i][5] = fmaf(s, e1.y, acc[i][5]);
      acc[i][6] = fmaf(s, e1.z, acc[i][6]); acc[i][7] = fmaf(s, e1.w, acc[i][7]);
    }
  }
#pragma unroll
  for (int i = 0; i < LOC_BG; i++)
#pragma unroll
    for (int k = 0; k < LOC_K; k++) acc[i][k] = warp_sum(acc[i][k]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < LOC_BG; i++) {
      if (b0 + i >= B) break;
      float* g = G + ((long long)(b0 + i) * C + c) * LOC_K;
      *reinterpret_cast<float4*>(g) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(g + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

// raw[b,p] = < normalize_c( relu( a_c * (G[b,c,:] . e_p + bias_c) + s_c ) ), f[b,:] >
// G[b] and the per-channel vectors in smem; LOC_PARTS adjacent threads share one position and interleave the channels.
__global__ void loc_score_kernel(const float* __restrict__ E, const float* __restrict__ G, const float* __restrict__ bias,
                                 const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, const float* __restrict__ flang,
                                 float* __restrict__ raw, float* __restrict__ inrm, int SN, int C) {
  extern __shared__ float4 loc_smem[];
  float4* g4 = loc_smem;                               // [C][2]
  float* a = reinterpret_cast<float*>(g4 + 2 * C);     // [C]  BN scale
  float* s = a + C;                                    // [C]  BN shift with the Linear's bias folded in
  float* f = s + C;                                    // [C]  location phrase vector
  const int b = blockIdx.y;
  const float4* gsrc = reinterpret_cast<const float4*>(G + (long long)b * C * LOC_K);
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) g4[i] = gsrc[i];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float sc = bn_scale[c];
    a[c] = sc;
    s[c] = fmaf(sc, bias ? bias[c] : 0.f, bn_shift[c]);
    f[c] = flang[(long long)b * C + c];
  }
  __syncthreads();
  // thread = (position group pg, channel part); it carries LOC_PP positions (pg, pg + 32, ...) so that one read of G[c,:] and of
  // the per-channel scalars from smem feeds LOC_PP * 12 FMAs (the kernel is bound by the smem pipe otherwise)
  const int part = threadIdx.x % LOC_PARTS;
  const int pg = threadIdx.x / LOC_PARTS;
  const int groups = blockDim.x / LOC_PARTS;
  const int p0 = blockIdx.x * (groups * LOC_PP) + pg;
  float4 e0[LOC_PP], e1[LOC_PP];
  float nrm[LOC_PP], dt[LOC_PP];
#pragma unroll
  for (int j = 0; j < LOC_PP; j++) {
    const int pr = min(p0 + j * groups, SN - 1);       // tail threads stay in the shuffles
    e0[j] = *reinterpret_cast<const float4*>(E + (long long)pr * LOC_K);
    e1[j] = *reinterpret_cast<const float4*>(E + (long long)pr * LOC_K + 4);
    nrm[j] = 0.f;
    dt[j] = 0.f;
  }
#pragma unroll 2
  for (int c = part; c < C; c += LOC_PARTS) {
    const float4 g0 = g4[2 * c], g1 = g4[2 * c + 1];
    const float ac = a[c], sc = s[c], fc = f[c];
#pragma unroll
    for (int j = 0; j < LOC_PP; j++) {
      float z = g0.x * e0[j].x;
      z = fmaf(g0.y, e0[j].y, z); z = fmaf(g0.z, e0[j].z, z); z = fmaf(g0.w, e0[j].w, z);
      z = fmaf(g1.x, e1[j].x, z); z = fmaf(g1.y, e1[j].y, z); z = fmaf(g1.z, e1[j].z, z); z = fmaf(g1.w, e1[j].w, z);
      const float y = fmaxf(fmaf(ac, z, sc), 0.f);
      nrm[j] = fmaf(y, y, nrm[j]);
      dt[j] = fmaf(y, fc, dt[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < LOC_PP; j++) {
#pragma unroll
    for (int o = 1; o < LOC_PARTS; o <<= 1) {
      nrm[j] += __shfl_xor_sync(0xffffffffu, nrm[j], o);
      dt[j] += __shfl_xor_sync(0xffffffffu, dt[j], o);
    }
    const int p = p0 + j * groups;
    if (part == 0 && p < SN) {
      const float in = 1.f / fmaxf(sqrtf(nrm[j]), 1e-12f);     // F.normalize(p=2, dim=1, eps=1e-12)
      raw[(long long)b * SN + p] = dt[j] * in;
      if (inrm) inrm[(long long)b * SN + p] = in;              // kept for the backward (training form)
    }
  }
}

// score[b,:] = (raw[b,:] - min) / (max - min + 1e-6)     one CTA per image
__global__ void loc_minmax_kernel(const float* __restrict__ raw, float* __restrict__ score, int SN) {
  __shared__ float sh[32];
  const float* r = raw + (long long)blockIdx.x * SN;
  float mn = INFINITY, mx = -INFINITY;
  for (int p = threadIdx.x; p < SN; p += blockDim.x) {
    const float v = r[p];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, sh);
  mn = -block_max(-mn, sh);
  const float inv = 1.f / (mx - mn + 1e-6f);
  float* o = score + (long long)blockIdx.x * SN;
  for (int p = threadIdx.x; p < SN; p += blockDim.x) o[p] = (r[p] - mn) * inv;
}


// ------------------------------------------------------------------------------------------------------------------------------
// Training form (batch statistics + backward; model/DCNet_model.py:556-603 under autograd).  Same rank-8 identity; what the
// materialised form gets from a [B*SN, C] tensor comes from G [B,C,8] and the first two moments of E here:
//   z[b,p,c] = G[b,c,:] . E[p,:] + bias[c]       mean_c = bias_c + mean_b G[b,c,:] . ebar
//   var_c = mean_b ( G[b,c,:]^T Mc G[b,c,:]  +  (G[b,c,:] . ebar - mean_b(...))^2 )          (within-image + between-image variance: both >= 0)
//   ebar = mean_p E[p,:],  Mc = mean_p (E[p,:] - ebar)(E[p,:] - ebar)^T
// Backward, with v = scale_c (z - mean_c) + beta_c, y = relu(v), yhat = y / |y|, raw = <yhat, f>:
//   pass 1 (loc_bwd_sums):  dv = [v > 0] draw (f_c - yhat_c raw) / |y| ;  S1_c = sum dv, S2_c = sum dv xhat, df[b,c] = sum_p draw yhat_c
//   pass 2 (loc_bwd_dz):    dz = scale_c (dv - S1_c / M - xhat S2_c / M) ;  dG[b,c,:] += dz E[p,:],  dE[p,:] += dz G[b,c,:]
//   pass 3 (loc_bwd_w):     t = dG[b,c,:] . E[q,:] ;  dW[c,q] = sum_b obj[b,q] t,  dobj[b,q] = sum_c W[c,q] t,
//                           dE[q,:] += sum_c W[c,q] sum_b obj[b,q] dG[b,c,:]
// |y| and raw are kept from the forward, so every (b,p,c) term of pass 1 / 2 is independent: one thread per channel, positions
// broadcast from shared memory.  The [B,SN,C] tensor never exists; pass 3 is the only pass over W besides the forward's.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int LOC_TP = 64;         // positions per CTA in the backward passes 1 and 2

// mom[0..8) = ebar, mom[8..72) = Mc (row-major 8x8).  One CTA.
__global__ void __launch_bounds__(256) loc_moments_kernel(const float* __restrict__ E, int SN, float* __restrict__ mom) {
  __shared__ float sh[32];
  __shared__ float ebar[LOC_K];
  float a[LOC_K];
#pragma unroll
  for (int k = 0; k < LOC_K; k++) a[k] = 0.f;
  for (int p = threadIdx.x; p < SN; p += blockDim.x)
#pragma unroll
    for (int k = 0; k < LOC_K; k++) a[k] += E[(long long)p * LOC_K + k];
#pragma unroll
  for (int k = 0; k < LOC_K; k++) {
    const float t = block_sum(a[k], sh);
    if (threadIdx.x == 0) { ebar[k] = t / (float)SN; mom[k] = t / (float)SN; }
  }
  __syncthreads();
  for (int k = 0; k < LOC_K; k++) {
    float m[LOC_K];
#pragma unroll
    for (int l = 0; l < LOC_K; l++) m[l] = 0.f;
    for (int p = threadIdx.x; p < SN; p += blockDim.x) {
      const float dk = E[(long long)p * LOC_K + k] - ebar[k];
#pragma unroll
      for (int l = 0; l < LOC_K; l++) m[l] = fmaf(dk, E[(long long)p * LOC_K + l] - ebar[l], m[l]);
    }
#pragma unroll
    for (int l = 0; l < LOC_K; l++) {
      const float t = block_sum(m[l], sh);
      if (threadIdx.x == 0) mom[LOC_K + k * LOC_K + l] = t / (float)SN;
    }
  }
}

// batch statistics of z from G and the moments; one thread per channel.  Outputs the BN affine the score kernel applies
// (scale = gamma invstd, shift = beta - mean scale with mean INCLUDING the Linear's bias), mean_nb = mean - bias and invstd for
// the backward; running statistics like nn.BatchNorm1d (momentum, unbiased variance over n = B*SN rows).
__global__ void loc_stats_kernel(const float* __restrict__ G, const float* __restrict__ mom, const float* __restrict__ bias,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                 float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_nb, float* __restrict__ invstd,
                                 float* __restrict__ running_mean, float* __restrict__ running_var, long long* __restrict__ nbt,
                                 int B, int SN, int C) {
  __shared__ float m[LOC_K + LOC_K * LOC_K];
  for (int i = threadIdx.x; i < LOC_K + LOC_K * LOC_K; i += blockDim.x) m[i] = mom[i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float dsum = 0.f;
  for (int b = 0; b < B; b++) {
    const float* g = G + ((long long)b * C + c) * LOC_K;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < LOC_K; k++) d = fmaf(g[k], m[k], d);
    dsum += d;
  }
  const float mu = dsum / (float)B;
  float var = 0.f;
  for (int b = 0; b < B; b++) {
    const float* g = G + ((long long)b * C + c) * LOC_K;
    float d = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < LOC_K; k++) {
      d = fmaf(g[k], m[k], d);
      float r = 0.f;
#pragma unroll
      for (int l = 0; l < LOC_K; l++) r = fmaf(m[LOC_K + k * LOC_K + l], g[l], r);
      q = fmaf(g[k], r, q);
    }
    var += q + (d - mu) * (d - mu);
  }
  var /= (float)B;
  const float is = rsqrtf(var + eps);
  const float bc = bias ? bias[c] : 0.f;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - (mu + bc) * sc;
  mean_nb[c] = mu;
  invstd[c] = is;
  if (running_mean) {
    const float n = (float)B * (float)SN;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mu + bc);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * n / (n - 1.f);
  }
  if (nbt && c == 0) *nbt += 1;
}

// d raw from d score through the per-image min-max normalisation (both extrema take their gradient at the first position that attains
// them, like torch.min / torch.max along a dimension).  One CTA per image.
__global__ void __launch_bounds__(256) loc_minmax_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ dscore,
                                                             float* __restrict__ draw, int SN) {
  __shared__ float sh[32];
  __shared__ int shi[2];
  const float* r = raw + (long long)blockIdx.x * SN;
  const float* ds = dscore + (long long)blockIdx.x * SN;
  float mn = INFINITY, mx = -INFINITY;
  for (int p = threadIdx.x; p < SN; p += blockDim.x) { mn = fminf(mn, r[p]); mx = fmaxf(mx, r[p]); }
  mx = block_max(mx, sh);
  mn = -block_max(-mn, sh);
  if (threadIdx.x == 0) { shi[0] = SN; shi[1] = SN; }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int p = threadIdx.x; p < SN; p += blockDim.x) {
    if (r[p] == mn) atomicMin(&shi[0], p);
    if (r[p] == mx) atomicMin(&shi[1], p);
    s1 += ds[p];
    s2 = fmaf(ds[p], r[p] - mn, s2);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  __syncthreads();
  const float inv = 1.f / (mx - mn + 1e-6f);
  const int imn = shi[0], imx = shi[1];
  float* o = draw + (long long)blockIdx.x * SN;
  for (int p = threadIdx.x; p < SN; p += blockDim.x) {
    float g = ds[p] * inv;
    if (p == imn) g += (s2 * inv - s1) * inv;       // d/d mn: sum ds (-(1/D) + (raw - mn)/D^2)
    if (p == imx) g -= s2 * inv * inv;              // d/d mx: sum ds (-(raw - mn)/D^2)
    o[p] = g;
  }
}

// pass 1 and pass 2 of the backward share their front end: thread = channel c of image b = blockIdx.y, tile of LOC_TP positions
struct LocBwdP {
  const float* E; const float* G; const float* scale; const float* shift; const float* bias; const float* mean_nb; const float* invstd;
  const float* flang; const float* raw; const float* inrm; const float* draw;
  float* S1; float* S2; float* df;       // pass 1 outputs (pass 2 inputs: S1, S2)
  float* dG; float* dE;                  // pass 2 outputs
  int SN, C; float invM;
};

template <int PASS>
__global__ void __launch_bounds__(256) loc_bwd_kernel(const LocBwdP q) {
  __shared__ float sE[LOC_TP][LOC_K], sraw[LOC_TP], sin[LOC_TP], sdr[LOC_TP];
  __shared__ float sdE[LOC_TP][LOC_K];
  const int b = blockIdx.y, p0 = blockIdx.x * LOC_TP;
  const int np = min(LOC_TP, q.SN - p0);
  for (int i = threadIdx.x; i < LOC_TP * LOC_K; i += blockDim.x) {
    (&sE[0][0])[i] = i < np * LOC_K ? q.E[(long long)p0 * LOC_K + i] : 0.f;
    (&sdE[0][0])[i] = 0.f;
  }
  for (int i = threadIdx.x; i < LOC_TP; i += blockDim.x) {
    const bool ok = i < np;
    sraw[i] = ok ? q.raw[(long long)b * q.SN + p0 + i] : 0.f;
    sin[i] = ok ? q.inrm[(long long)b * q.SN + p0 + i] : 0.f;
    sdr[i] = ok ? q.draw[(long long)b * q.SN + p0 + i] : 0.f;        // draw = 0 switches a padded position off
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < q.C; c += blockDim.x) {
    float g[LOC_K];
    const float4 g0 = *reinterpret_cast<const float4*>(q.G + ((long long)b * q.C + c) * LOC_K);
    const float4 g1 = *reinterpret_cast<const float4*>(q.G + ((long long)b * q.C + c) * LOC_K + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    const float sc = q.scale[c], shf = fmaf(sc, q.bias ? q.bias[c] : 0.f, q.shift[c]), fc = q.flang[(long long)b * q.C + c];
    const float mu = q.mean_nb[c], is = q.invstd[c];
    float k1 = 0.f, k2 = 0.f;
    if (PASS == 2) { k1 = q.S1[c] * q.invM; k2 = q.S2[c] * q.invM; }
    float a1 = 0.f, a2 = 0.f, a3 = 0.f, dg[LOC_K];
#pragma unroll
    for (int k = 0; k < LOC_K; k++) dg[k] = 0.f;
    for (int i = 0; i < np; i++) {
      float z = 0.f;
#pragma unroll
      for (int k = 0; k < LOC_K; k++) z = fmaf(g[k], sE[i][k], z);
      const float v = fmaf(sc, z, shf);
      const float yh = fmaxf(v, 0.f) * sin[i];
      const float dv = v > 0.f ? sdr[i] * (fc - yh * sraw[i]) * sin[i] : 0.f;
      const float xh = (z - mu) * is;
      if (PASS == 1) {
        a1 += dv;
        a2 = fmaf(dv, xh, a2);
        a3 = fmaf(sdr[i], yh, a3);
      } else {
        const float dz = sc * (dv - k1 - xh * k2);
#pragma unroll
        for (int k = 0; k < LOC_K; k++) dg[k] = fmaf(dz, sE[i][k], dg[k]);
        // dE[p,:] += dz G[b,c,:]: summed over the channels of the warp, then into the tile's shared accumulator
#pragma unroll
        for (int k = 0; k < LOC_K; k++) {
          const float t = warp_sum(dz * g[k]);
          if (lane == 0) atomicAdd(&sdE[i][k], t);
        }
      }
    }
    if (PASS == 1) {
      atomicAdd(q.S1 + c, a1);
      atomicAdd(q.S2 + c, a2);
      atomicAdd(q.df + (long long)b * q.C + c, a3);
    } else {
#pragma unroll
      for (int k = 0; k < LOC_K; k++) atomicAdd(q.dG + ((long long)b * q.C + c) * LOC_K + k, dg[k]);
    }
  }
  if (PASS == 2) {
    __syncthreads();
    for (int i = threadIdx.x; i < np * LOC_K; i += blockDim.x) atomicAdd(q.dE + (long long)p0 * LOC_K + i, (&sdE[0][0])[i]);
  }
}

// pass 3: thread = position q, CTA = 128 positions x a slice of LOC_CS channels; the images go through in chunks of LOC_BC
constexpr int LOC_CS = 32, LOC_BC = 8;
__global__ void __launch_bounds__(128) loc_bwd_w_kernel(const float* __restrict__ E, const float* __restrict__ obj, const float* __restrict__ W,
                                                        int ldw, const float* __restrict__ dG, float* __restrict__ dW, float* __restrict__ dobj,
                                                        float* __restrict__ dE, int B, int SN, int C) {
  __shared__ float4 sG[LOC_BC][LOC_CS][2];
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const int c0 = blockIdx.y * LOC_CS;
  const bool ok = qi < SN;
  const int qq = ok ? qi : SN - 1;
  float e[LOC_K];
#pragma unroll
  for (int k = 0; k < LOC_K; k++) e[k] = E[(long long)qq * LOC_K + k];
  float wsum[LOC_CS], dEa[LOC_K];
#pragma unroll
  for (int j = 0; j < LOC_CS; j++) wsum[j] = 0.f;
#pragma unroll
  for (int k = 0; k < LOC_K; k++) dEa[k] = 0.f;
  for (int b0 = 0; b0 < B; b0 += LOC_BC) {
    __syncthreads();
    for (int i = threadIdx.x; i < LOC_BC * LOC_CS * 2; i += 128) {
      const int bb = i / (LOC_CS * 2), r = i - bb * (LOC_CS * 2);
      const int cc = c0 + r / 2;
      (&sG[0][0][0])[i] = (b0 + bb < B && cc < C) ? reinterpret_cast<const float4*>(dG + ((long long)(b0 + bb) * C + cc) * LOC_K)[r & 1]
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float ob[LOC_BC], da[LOC_BC];
#pragma unroll
    for (int bb = 0; bb < LOC_BC; bb++) { ob[bb] = (b0 + bb < B) ? obj[(long long)(b0 + bb) * SN + qq] : 0.f; da[bb] = 0.f; }
#pragma unroll
    for (int j = 0; j < LOC_CS; j++) {
      const float wv = (c0 + j < C) ? W[(long long)(c0 + j) * ldw + qq] : 0.f;
      float u[LOC_K];
#pragma unroll
      for (int k = 0; k < LOC_K; k++) u[k] = 0.f;
#pragma unroll
      for (int bb = 0; bb < LOC_BC; bb++) {
        const float4 h0 = sG[bb][j][0], h1 = sG[bb][j][1];
        float t = h0.x * e[0];
        t = fmaf(h0.y, e[1], t); t = fmaf(h0.z, e[2], t); t = fmaf(h0.w, e[3], t);
        t = fmaf(h1.x, e[4], t); t = fmaf(h1.y, e[5], t); t = fmaf(h1.z, e[6], t); t = fmaf(h1.w, e[7], t);
        wsum[j] = fmaf(ob[bb], t, wsum[j]);
        da[bb] = fmaf(wv, t, da[bb]);
        u[0] = fmaf(ob[bb], h0.x, u[0]); u[1] = fmaf(ob[bb], h0.y, u[1]); u[2] = fmaf(ob[bb], h0.z, u[2]); u[3] = fmaf(ob[bb], h0.w, u[3]);
        u[4] = fmaf(ob[bb], h1.x, u[4]); u[5] = fmaf(ob[bb], h1.y, u[5]); u[6] = fmaf(ob[bb], h1.z, u[6]); u[7] = fmaf(ob[bb], h1.w, u[7]);
      }
#pragma unroll
      for (int k = 0; k < LOC_K; k++) dEa[k] = fmaf(wv, u[k], dEa[k]);
    }
    if (ok) {
#pragma unroll
      for (int bb = 0; bb < LOC_BC; bb++)
        if (b0 + bb < B) atomicAdd(dobj + (long long)(b0 + bb) * SN + qi, da[bb]);
    }
  }
  if (ok) {
#pragma unroll
    for (int j = 0; j < LOC_CS; j++)
      if (c0 + j < C) dW[(long long)(c0 + j) * ldw + qi] = wsum[j];
#pragma unroll
    for (int k = 0; k < LOC_K; k++) atomicAdd(dE + (long long)qi * LOC_K + k, dEa[k]);
  }
}

}  // namespace

extern "C" int dcnet_loc_rank8_fwd(const float* E, const float* obj, const float* W, int ldw, const float* bias,
                                   const float* bn_scale, const float* bn_shift, const float* flang,
                                   float* G, float* raw, float* score, int B, int SN, int C, void* stream) {
  DCNET_CHECK_ARG(E && obj && W && bn_scale && bn_shift && flang && G && raw && score, "loc_rank8_fwd: null argument");
  DCNET_CHECK_ARG(B > 0 && SN > 0 && C > 0 && ldw >= SN && B <= 65535, "loc_rank8_fwd: bad sizes B=%d SN=%d C=%d ldw=%d", B, SN, C, ldw);
  DCNET_CHECK_ARG((reinterpret_cast<uintptr_t>(E) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                  "loc_rank8_fwd: E and G must be 16-byte aligned");
  const size_t smem = (size_t)C * (2 * sizeof(float4) + 3 * sizeof(float));
  DCNET_CHECK_ARG(smem <= 48 * 1024, "loc_rank8_fwd: C=%d needs %zu B of shared memory (max 48 KiB)", C, smem);
  cudaStream_t st = as_stream(stream);
  loc_g_kernel<<<dim3((C + 7) / 8, (B + LOC_BG - 1) / LOC_BG), 256, 0, st>>>(E, obj, W, ldw, G, B, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_fwd.g");
  constexpr int PPB = 128 / LOC_PARTS * LOC_PP;     // positions per CTA of 128 threads
  loc_score_kernel<<<dim3((SN + PPB - 1) / PPB, B), 128, smem, st>>>(E, G, bias, bn_scale, bn_shift, flang, raw, nullptr, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_fwd.score");
  loc_minmax_kernel<<<B, 256, 0, st>>>(raw, score, SN);
  DCNET_LAUNCH_OK("loc_rank8_fwd.minmax");
  return 0;
}

// training forward: G, batch statistics (+ running statistics), scores.  Outputs kept for the backward: G [B,C,8], stats [4*C] =
// (scale, shift, mean without bias, invstd), raw / inrm [B,SN]; mom [72] is scratch.
extern "C" int dcnet_loc_rank8_train_fwd(const float* E, const float* obj, const float* W, int ldw, const float* bias,
                                         const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                         float* running_var, long long* num_batches_tracked, const float* flang,
                                         float* G, float* mom, float* stats, float* raw, float* inrm, float* score,
                                         int B, int SN, int C, void* stream) {
  DCNET_CHECK_ARG(E && obj && W && gamma && beta && flang && G && mom && stats && raw && inrm && score, "loc_rank8_train_fwd: null argument");
  DCNET_CHECK_ARG(B > 0 && SN > 1 && C > 0 && ldw >= SN && B <= 65535, "loc_rank8_train_fwd: bad sizes B=%d SN=%d C=%d ldw=%d", B, SN, C, ldw);
  DCNET_CHECK_ARG((reinterpret_cast<uintptr_t>(E) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                  "loc_rank8_train_fwd: E and G must be 16-byte aligned");
  const size_t smem = (size_t)C * (2 * sizeof(float4) + 3 * sizeof(float));
  DCNET_CHECK_ARG(smem <= 48 * 1024, "loc_rank8_train_fwd: C=%d needs %zu B of shared memory (max 48 KiB)", C, smem);
  cudaStream_t st = as_stream(stream);
  loc_moments_kernel<<<1, 256, 0, st>>>(E, SN, mom);
  DCNET_LAUNCH_OK("loc_rank8_train_fwd.moments");
  loc_g_kernel<<<dim3((C + 7) / 8, (B + LOC_BG - 1) / LOC_BG), 256, 0, st>>>(E, obj, W, ldw, G, B, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_train_fwd.g");
  float* scale = stats; float* shift = stats + C; float* mean_nb = stats + 2 * C; float* invstd = stats + 3 * C;
  loc_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(G, mom, bias, gamma, beta, eps, momentum, scale, shift, mean_nb, invstd, running_mean,
                                                    running_var, num_batches_tracked, B, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_train_fwd.stats");
  constexpr int PPB = 128 / LOC_PARTS * LOC_PP;
  loc_score_kernel<<<dim3((SN + PPB - 1) / PPB, B), 128, smem, st>>>(E, G, bias, scale, shift, flang, raw, inrm, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_train_fwd.score");
  loc_minmax_kernel<<<B, 256, 0, st>>>(raw, score, SN);
  DCNET_LAUNCH_OK("loc_rank8_train_fwd.minmax");
  return 0;
}

// backward of the above from dscore [B,SN].  Outputs (all overwritten): dE [SN,8], dobj [B,SN], dW [C,ldw] (columns < SN), dgamma / dbeta
// [C], dflang [B,C]; the Linear's bias has no gradient under batch statistics (BatchNorm removes it).  draw [B,SN], dG [B,C,8] scratch.
extern "C" int dcnet_loc_rank8_train_bwd(const float* E, const float* obj, const float* W, int ldw, const float* bias, const float* flang,
                                         const float* G, const float* stats, const float* raw, const float* inrm, const float* dscore,
                                         float* draw, float* dG, float* dE, float* dobj, float* dW, float* dgamma, float* dbeta, float* dflang,
                                         int B, int SN, int C, void* stream) {
  DCNET_CHECK_ARG(E && obj && W && flang && G && stats && raw && inrm && dscore && draw && dG && dE && dobj && dW && dgamma && dbeta && dflang,
                  "loc_rank8_train_bwd: null argument");
  DCNET_CHECK_ARG(B > 0 && SN > 1 && C > 0 && ldw >= SN && B <= 65535, "loc_rank8_train_bwd: bad sizes");
  DCNET_CHECK_ARG((reinterpret_cast<uintptr_t>(G) & 15) == 0 && (reinterpret_cast<uintptr_t>(dG) & 15) == 0, "loc_rank8_train_bwd: G / dG must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  DCNET_CUDA(cudaMemsetAsync(dG, 0, (size_t)B * C * LOC_K * sizeof(float), st), "loc_rank8_train_bwd.memset");
  DCNET_CUDA(cudaMemsetAsync(dE, 0, (size_t)SN * LOC_K * sizeof(float), st), "loc_rank8_train_bwd.memset");
  DCNET_CUDA(cudaMemsetAsync(dobj, 0, (size_t)B * SN * sizeof(float), st), "loc_rank8_train_bwd.memset");
  DCNET_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * sizeof(float), st), "loc_rank8_train_bwd.memset");
  DCNET_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * sizeof(float), st), "loc_rank8_train_bwd.memset");
  DCNET_CUDA(cudaMemsetAsync(dflang, 0, (size_t)B * C * sizeof(float), st), "loc_rank8_train_bwd.memset");
  loc_minmax_bwd_kernel<<<B, 256, 0, st>>>(raw, dscore, draw, SN);
  DCNET_LAUNCH_OK("loc_rank8_train_bwd.minmax");
  LocBwdP q{};
  q.E = E; q.G = G; q.scale = stats; q.shift = stats + C; q.bias = bias; q.mean_nb = stats + 2 * C; q.invstd = stats + 3 * C;
  q.flang = flang; q.raw = raw; q.inrm = inrm; q.draw = draw;
  q.S1 = dbeta; q.S2 = dgamma; q.df = dflang; q.dG = dG; q.dE = dE;
  q.SN = SN; q.C = C; q.invM = 1.f / ((float)B * (float)SN);
  dim3 grid((SN + LOC_TP - 1) / LOC_TP, B);
  loc_bwd_kernel<1><<<grid, 256, 0, st>>>(q);
  DCNET_LAUNCH_OK("loc_rank8_train_bwd.sums");
  loc_bwd_kernel<2><<<grid, 256, 0, st>>>(q);
  DCNET_LAUNCH_OK("loc_rank8_train_bwd.dz");
  loc_bwd_w_kernel<<<dim3((SN + 127) / 128, (C + LOC_CS - 1) / LOC_CS), 128, 0, st>>>(E, obj, W, ldw, dG, dW, dobj, dE, B, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_train_bwd.w");
  return 0;
}
