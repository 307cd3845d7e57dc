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
                                 float* __restrict__ raw, int SN, int C) {
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
    if (part == 0 && p < SN) raw[(long long)b * SN + p] = dt[j] / fmaxf(sqrtf(nrm[j]), 1e-12f);     // F.normalize(p=2, dim=1, eps=1e-12)
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
  loc_score_kernel<<<dim3((SN + PPB - 1) / PPB, B), 128, smem, st>>>(E, G, bias, bn_scale, bn_shift, flang, raw, SN, C);
  DCNET_LAUNCH_OK("loc_rank8_fwd.score");
  loc_minmax_kernel<<<B, 256, 0, st>>>(raw, score, SN);
  DCNET_LAUNCH_OK("loc_rank8_fwd.minmax");
  return 0;
}
