// a5 / a20: co-attention (model/DCNet_model.py:449-459; model/test_DCNet_model.py:247-274), one direction per problem.
//   S = Fa^T Fb,  P = softmax_j(tau S),  O = Fb P^T.
// This file holds the fp32 composition used for exactness checks and as the backward of round 1: the score matrix
// lives in caller-provided scratch.  The tcgen05 path (umma_coattn.cu) keeps S/P in TMEM/SMEM and is selected by
// dcnet_coattn_fwd when the shape is supported.
#include <cuda_fp16.h>

#include "common.cuh"

int umma_coattn_fwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob, float* out, int n_out, float* lse,
                    int C, int N, float tau, void* ws, size_t ws_bytes, cudaStream_t st, int round_out);  // umma_coattn.cu
bool umma_coattn_supported(int C, int N);
size_t umma_coattn_workspace_bytes(int F, int C, int N);

namespace {

// rows of S' = tau*S (already scaled): P = softmax(row), lse = max + log(sum).  One warp per row.
__global__ void softmax_rows_kernel(float* __restrict__ S, float* __restrict__ lse, long long rows, int N, int ld) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = S + row * ld;
  float m = -INFINITY;
  for (int j = lane; j < N; j += 32) m = fmaxf(m, p[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < N; j += 32) {
    const float e = __expf(p[j] - m);
    p[j] = e;
    s += e;
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int j = lane; j < N; j += 32) p[j] *= inv;
  if (lane == 0 && lse) lse[row] = m + logf(s);
}

// P = exp(S' - lse[row])
__global__ void exp_lse_kernel(float* __restrict__ S, const float* __restrict__ lse, long long rows, int N) {
  const long long total = rows * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    S[i] = __expf(S[i] - lse[i / N]);
}

// dS = tau * P o (dP - delta 1^T), delta_i = sum_j P_ij dP_ij; in place over dP.  One warp per row.
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, long long rows, int N, int ld, float tau) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = P + row * ld;
  float* d = dP + row * ld;
  float dl = 0.f;
  for (int j = lane; j < N; j += 32) dl = fmaf(p[j], d[j], dl);
  dl = warp_sum(dl);
  for (int j = lane; j < N; j += 32) d[j] = tau * p[j] * (d[j] - dl);
}

// dst[r][0..N) = src[(idx ? idx[r / C] : r / C)][r % C][0..N): a copy with the row pitch padded to ld (TMA needs 16-byte pitches)
__global__ void pad_pitch_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, long long rows, int C,
                                 int N, int ld) {
  const long long total = rows * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int n = (int)(i - r * N);
    const long long b = r / C, c = r - b * C;
    const long long sb = idx ? idx[b] : b;
    dst[r * ld + n] = src[(sb * C + c) * N + n];
  }
}

// Backward, between the two N x N contractions: for problem z and query i
//   delta[z,i] = <dO[:,i], O[:,i]>   the softmax-backward row term sum_j P_ij dP_ij, taken from the SAVED output instead of P
//   r[z,i] <- 1 / r[z,i]             (r = row sums of E = exp(tau S - shift))
// block (32, 8): 32 lanes x 4 consecutive positions (float4), 8 channel rows per step; grid (N/128, nprob).
__global__ void __launch_bounds__(256) coattn_delta_kernel(const float* __restrict__ dO, const float* __restrict__ O, const int* __restrict__ oidx,
                                                           float* __restrict__ r, float* __restrict__ delta, int C, int N) {
  const int z = blockIdx.y;
  const int n = (blockIdx.x * 32 + threadIdx.x) * 4;
  // gridDim.z channel slices per (position block, problem): small maps (N = 1024: 8 x 16 blocks) would leave most SMs idle;
  // delta is zero-filled by the caller and the slices add into it
  const int cper = (C + gridDim.z - 1) / gridDim.z, c0 = blockIdx.z * cper, c1 = min(C, c0 + cper);
  __shared__ float part[8][128];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const long long src = (long long)oidx[z] * C * N + n;
#pragma unroll 4
    for (int c = c0 + threadIdx.y; c < c1; c += 8) {
      const float4 g = *reinterpret_cast<const float4*>(dO + src + (long long)c * N);
      const float4 o = *reinterpret_cast<const float4*>(O + src + (long long)c * N);
      acc[0] = fmaf(g.x, o.x, acc[0]); acc[1] = fmaf(g.y, o.y, acc[1]); acc[2] = fmaf(g.z, o.z, acc[2]); acc[3] = fmaf(g.w, o.w, acc[3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) part[threadIdx.y][threadIdx.x * 4 + i] = acc[i];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 128) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += part[k][t];
    const int nn = blockIdx.x * 128 + t;
    if (nn < N) {
      atomicAdd(delta + (long long)z * N + nn, s);
      if (blockIdx.z == 0) r[(long long)z * N + nn] = 1.f / r[(long long)z * N + nn];      // nobody reads r before the next kernel
    }
  }
}

// After dS' = tau (dP - delta) P has been formed with the delta above.  The saved output comes from the forward's bf16 operands, so
// delta is off by a relative ~1e-4 against the tf32 P / dP used here; rows of dS' then sum to rho_i = tau (delta_true - delta)_i
// instead of zero, and because the columns of the maps share a large common component that small row offset would not cancel in
// Fb dS^T (measured: 6e-3 gradient error at N = 1024 against 9e-4).  rho comes for free as the row sums of the dS' epilogue;
// dS = dS' - rho_i P_ij exactly, which moves into the [C,N] maps:
//     dFa[:,i]  = Fb dS'^T[:,i] - rho_i O[:,i]              (O ~ Fb P^T: second-order error now)
//     dFb      += (dO - rho Fa) P  +  Fa dS'
// This kernel writes dOs[z][c,i] = (dO[c,i] - rho_i Fa[c,i]) / r_i (rounded to tf32: the A operand of dOs E) and adds -rho_i O[c,i]
// to dframes[qa[z]].  Same tiling as coattn_delta_kernel.
__global__ void __launch_bounds__(256) coattn_fix_kernel(const float* __restrict__ dO, const float* __restrict__ O, const float* __restrict__ frames,
                                                         const int* __restrict__ oidx, const int* __restrict__ qa, const float* __restrict__ inv_r,
                                                         const float* __restrict__ rho, float* __restrict__ dOs, float* __restrict__ dframes,
                                                         int C, int N) {
  const int z = blockIdx.y;
  const int n = (blockIdx.x * 32 + threadIdx.x) * 4;
  if (n >= N) return;
  const long long so = (long long)oidx[z] * C * N + n;
  const long long sq = (long long)qa[z] * C * N + n;
  float* dst = dOs + (long long)z * C * N + n;
  const float4 iv = *reinterpret_cast<const float4*>(inv_r + (long long)z * N + n);
  const float4 rh = *reinterpret_cast<const float4*>(rho + (long long)z * N + n);
  auto rn = [](float x) { return tf32_rn(x); };
  const int cper = (C + gridDim.z - 1) / gridDim.z, c0 = blockIdx.z * cper, c1 = min(C, c0 + cper);
#pragma unroll 4
  for (int c = c0 + threadIdx.y; c < c1; c += 8) {
    const float4 g = *reinterpret_cast<const float4*>(dO + so + (long long)c * N);
    const float4 o = *reinterpret_cast<const float4*>(O + so + (long long)c * N);
    const float4 f = *reinterpret_cast<const float4*>(frames + sq + (long long)c * N);
    *reinterpret_cast<float4*>(dst + (long long)c * N) =
        make_float4(rn((g.x - rh.x * f.x) * iv.x), rn((g.y - rh.y * f.y) * iv.y), rn((g.z - rh.z * f.z) * iv.z), rn((g.w - rh.w * f.w) * iv.w));
    float* d = dframes + sq + (long long)c * N;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(-rh.x * o.x), "f"(-rh.y * o.y), "f"(-rh.z * o.z), "f"(-rh.w * o.w) : "memory");
  }
}

// ---- fp16 pipeline of the backward.  fp16 keeps 11 significant bits like tf32, so a contraction on fp16 operands is as accurate as the
// tf32 one AS LONG AS the operands stay in fp16's normal range (6e-5 .. 65504) -- at twice the MMA rate and half the operand bytes
// (E and dS' are N x N).  The maps are unit-norm (the forward's own fp16 staging is reused); E carries a factor 2^8 that cancels
// against its row sums; everything that scales with the incoming gradient (dO, dS', dOs) is multiplied by a per-problem power of two
// s_z that brings max |dO| to [4, 8), and the three reduce-add contractions undo it through alpha_z = 1 / s_z.
constexpr float E_SHIFT = 5.545177444479562f;      // 8 ln 2

// mx[z] = max |dO[oidx[z]]| as the bit pattern of a non-negative float (caller zeroes)
__global__ void __launch_bounds__(256) coattn_absmax_kernel(const float* __restrict__ dO, const int* __restrict__ oidx, unsigned int* __restrict__ mx,
                                                            long long CN4) {
  const int z = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(dO) + (long long)oidx[z] * CN4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < CN4; i += (long long)gridDim.x * 256) {
    const float4 v = src[i];
    m = fmaxf(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))), m);
  }
  __shared__ float sh[32];
  m = block_max(m, sh);
  if (threadIdx.x == 0) atomicMax(mx + z, __float_as_uint(m));
}

// the power of two that brings a maximum of m into [4, 8)
__device__ __forceinline__ float coattn_scale(unsigned int mbits) {
  const float m = __uint_as_float(mbits);
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                   // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 3 - e);
}

// coattn_delta_kernel of the fp16 pipeline: delta <- s <dO, O> (the row term in the scaled units of dP), r <- 1 / r, and the scaled fp16
// copy dO16[z][c][n] = fp16(s dO) (pitch ldh); block (0, z, 0) also publishes alpha_z[z] = 1 / s.
__global__ void __launch_bounds__(256) coattn_delta16_kernel(const float* __restrict__ dO, const float* __restrict__ O, const int* __restrict__ oidx,
                                                             const unsigned int* __restrict__ mx, int mx_by_oidx, float* __restrict__ r,
                                                             float* __restrict__ delta, __half* __restrict__ dO16, float* __restrict__ alpha_z,
                                                             int C, int N, int ldh) {
  const int z = blockIdx.y;
  const int n = (blockIdx.x * 32 + threadIdx.x) * 4;
  const float s = coattn_scale(mx[mx_by_oidx ? oidx[z] : z]);
  if (blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) alpha_z[z] = 1.f / s;
  const int cper = (C + gridDim.z - 1) / gridDim.z, c0 = blockIdx.z * cper, c1 = min(C, c0 + cper);
  __shared__ float part[8][128];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const long long src = (long long)oidx[z] * C * N + n;
    __half* dst = dO16 + (long long)z * C * ldh + n;
#pragma unroll 4
    for (int c = c0 + threadIdx.y; c < c1; c += 8) {
      const float4 g = *reinterpret_cast<const float4*>(dO + src + (long long)c * N);
      const float4 o = *reinterpret_cast<const float4*>(O + src + (long long)c * N);
      acc[0] = fmaf(g.x, o.x, acc[0]); acc[1] = fmaf(g.y, o.y, acc[1]); acc[2] = fmaf(g.z, o.z, acc[2]); acc[3] = fmaf(g.w, o.w, acc[3]);
      const __half2 lo = __floats2half2_rn(g.x * s, g.y * s), hi = __floats2half2_rn(g.z * s, g.w * s);
      *reinterpret_cast<uint2*>(dst + (long long)c * ldh) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) part[threadIdx.y][threadIdx.x * 4 + i] = acc[i];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  if (t < 128) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) sum += part[k][t];
    const int nn = blockIdx.x * 128 + t;
    if (nn < N) {
      atomicAdd(delta + (long long)z * N + nn, sum * s);
      if (blockIdx.z == 0) r[(long long)z * N + nn] = 1.f / r[(long long)z * N + nn];
    }
  }
}

// coattn_fix_kernel of the fp16 pipeline: rho arrives scaled by s (row sums of s dS'), the maps Fa come from the fp16 staging (pitch ldh):
//   dOs16[z][c,i] = fp16( (s dO[c,i] - rho_i Fa[c,i]) / r_i ),     dframes[qa[z]][c,i] -= (rho_i / s) O[c,i]
__global__ void __launch_bounds__(256) coattn_fix16_kernel(const float* __restrict__ dO, const float* __restrict__ O, const __half* __restrict__ F16,
                                                           const int* __restrict__ oidx, const int* __restrict__ qa, const float* __restrict__ inv_r,
                                                           const float* __restrict__ rho, const unsigned int* __restrict__ mx, int mx_by_oidx,
                                                           __half* __restrict__ dOs16, float* __restrict__ dframes, int C, int N, int ldh) {
  const int z = blockIdx.y;
  const int n = (blockIdx.x * 32 + threadIdx.x) * 4;
  if (n >= N) return;
  const float s = coattn_scale(mx[mx_by_oidx ? oidx[z] : z]), is = 1.f / s;
  const long long so = (long long)oidx[z] * C * N + n;
  const long long sq = (long long)qa[z] * C * N + n;
  const __half* fq = F16 + (long long)qa[z] * C * ldh + n;
  __half* dst = dOs16 + (long long)z * C * ldh + n;
  const float4 iv = *reinterpret_cast<const float4*>(inv_r + (long long)z * N + n);
  const float4 rh = *reinterpret_cast<const float4*>(rho + (long long)z * N + n);
  const int cper = (C + gridDim.z - 1) / gridDim.z, c0 = blockIdx.z * cper, c1 = min(C, c0 + cper);
#pragma unroll 4
  for (int c = c0 + threadIdx.y; c < c1; c += 8) {
    const float4 g = *reinterpret_cast<const float4*>(dO + so + (long long)c * N);
    const float4 o = *reinterpret_cast<const float4*>(O + so + (long long)c * N);
    const uint2 fu = *reinterpret_cast<const uint2*>(fq + (long long)c * ldh);
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&fu.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&fu.y));
    const __half2 lo = __floats2half2_rn((g.x * s - rh.x * f0.x) * iv.x, (g.y * s - rh.y * f0.y) * iv.y);
    const __half2 hi = __floats2half2_rn((g.z * s - rh.z * f1.x) * iv.z, (g.w * s - rh.w * f1.y) * iv.w);
    *reinterpret_cast<uint2*>(dst + (long long)c * ldh) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    float* d = dframes + sq + (long long)c * N;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(-rh.x * is * o.x), "f"(-rh.y * is * o.y), "f"(-rh.z * is * o.z),
                 "f"(-rh.w * is * o.w) : "memory");
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int pitch8h(int N) { return (N + 7) & ~7; }
inline int pitch4(int N) { return (N + 3) & ~3; }

}  // namespace

// Backward: the N x N tensors (P, then dP -> dS in place) are produced and consumed chunk by chunk of problems, in ONE reused
// scratch that is sized to stay resident in the 126 MB L2 (2 x chunk x N^2 fp32 <= g_bwd_l2_budget): the five contractions and
// the two row kernels of a chunk run back to back, every N^2 line is rewritten while still dirty in L2, so S / P / dP / dS do not
// travel to HBM and the workspace no longer grows with the number of problems.  A chunk keeps >= one wave of 128x256 tiles.
static long long g_bwd_l2_budget = 1ll << 62;      // default: one chunk (see profiles/r2b: per-problem chunks starve the [C,N]-output contractions)
extern "C" int dcnet_coattn_bwd_l2_budget(long long bytes) { g_bwd_l2_budget = bytes > 0 ? bytes : (1ll << 62); return 0; }
// 1 (default): dcnet_coattn_bwd at precision 2 runs its contractions on fp16 operands when the caller hands over the forward's staging;
// 0: always the tf32 contractions (bring-up / comparison knob, process-wide)
static int g_bwd_fp16 = 1;
static int g_bwd_use_keep = 1;    // 0 (dcnet_coattn_bwd_fp16(-1)): recompute E even when the forward kept it (comparison)
static int g_bwd_stop = 0;         // profiling: on > 1 stops the fp16 pipeline after its (on - 1)-th contraction (dcnet_gemm_trace then holds that launch)
extern "C" int dcnet_coattn_bwd_fp16(int on) {
  g_bwd_fp16 = on ? 1 : 0;
  g_bwd_stop = on > 1 ? on - 1 : 0;
  g_bwd_use_keep = on == -1 ? 0 : 1;
  return 0;
}
static int coattn_bwd_chunk(int nprob, int N) {
  const long long per = 2ll * N * N * (long long)sizeof(float);
  long long c = g_bwd_l2_budget / per;
  if (c < 1) c = 1;
  if (c > nprob) c = nprob;
  return (int)c;
}

extern "C" size_t dcnet_coattn_workspace_bytes(int F, int nprob, int C, int N, int precision) {
  if (nprob <= 0 || N <= 0 || F <= 0) return 256;
  // S / P and dP scratch: all problems for the unfused forward (precision <= 1), one L2-resident chunk for the backward
  const int chunk = (precision == 2) ? coattn_bwd_chunk(nprob, N) : nprob;
  size_t unfused = 2 * align256((size_t)chunk * N * N * sizeof(float)) + 256;
  if (precision == 2 && N % 4 == 0)    // fused-epilogue backward: dO / r [chunk,C,N] + row sums and delta [chunk,N] each
    unfused += align256((size_t)chunk * C * N * sizeof(float)) + 3 * align256((size_t)chunk * N * sizeof(float));
  if (N % 4 != 0 && precision >= 1)    // odd pitch: P / dP with the pitch padded to 4, plus padded copies of the maps and of dout
    unfused = 2 * align256((size_t)nprob * N * pitch4(N) * sizeof(float)) + align256((size_t)F * C * pitch4(N) * sizeof(float)) +
              align256((size_t)nprob * C * pitch4(N) * sizeof(float)) + 256;
  if (precision == 2 && N % 4 == 0) {  // fp16 pipeline (dcnet_coattn_bwd with the forward's staging): E16, dS16 [chunk,N,ldh], dO16, dOs16 [chunk,C,ldh]
    const size_t ldh = (size_t)((N + 7) & ~7);
    const size_t h16 = 2 * align256((size_t)chunk * N * ldh * 2) + 2 * align256((size_t)chunk * C * ldh * 2) +
                       3 * align256((size_t)chunk * N * sizeof(float)) + 2 * align256((size_t)chunk * sizeof(float)) + 256;
    if (h16 > unfused) unfused = h16;
  }
  const size_t fused = umma_coattn_workspace_bytes(F, C, N);                           // fp16 staging of the maps + column norms
  return precision == 2 ? (unfused > fused ? unfused : fused) : unfused;
}

static bool coattn_tc_ok(int precision, int C, int N, const void* a, const void* b, const void* c) {
  auto al = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  return precision >= 1 && N % 4 == 0 && C % 128 == 0 && al(a) && al(b) && al(c);
}

extern "C" int dcnet_coattn_fwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                float* out, int n_out, float* lse, int C, int N, float tau, int precision,
                                void* workspace, size_t workspace_bytes, void* stream) {
  DCNET_CHECK_ARG(frames && qa && kb && oidx && out && lse && nprob >= 0 && C > 0 && N > 0 && F > 0 && n_out > 0, "coattn_fwd: bad arguments");
  if (nprob == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (precision & DCNET_RN_TF32) {
    // the attention maps feed a tf32 contraction (corr_conv): round them once here.  The fused kernel does it in its epilogue;
    // the other forms (small maps) get a pass over `out` -- rows no problem writes (n_out > nprob) are zero either way
    precision &= ~DCNET_RN_TF32;
    if (precision == 2 && umma_coattn_supported(C, N))
      return umma_coattn_fwd(frames, F, qa, kb, oidx, nprob, out, n_out, lse, C, N, tau, workspace, workspace_bytes, st, 1);
    DCNET_TRY(dcnet_coattn_fwd(frames, F, qa, kb, oidx, nprob, out, n_out, lse, C, N, tau, precision, workspace, workspace_bytes, stream));
    return dcnet_round_tf32(out, out, (long long)n_out * C * N, stream);
  }
  if (precision == 2 && umma_coattn_supported(C, N))
    return umma_coattn_fwd(frames, F, qa, kb, oidx, nprob, out, n_out, lse, C, N, tau, workspace, workspace_bytes, st, 0);
  DCNET_CHECK_ARG(workspace && workspace_bytes >= dcnet_coattn_workspace_bytes(F, nprob, C, N, precision), "coattn_fwd: workspace too small");
  float* S = (float*)workspace;
  const long long CN = (long long)C * N, NN = (long long)N * N;
  const long long rows = (long long)nprob * N;
  if (coattn_tc_ok(precision, C, N, frames, out, S)) {
    // S' = tau Fa^T Fb : both operands MN-major views of the [C,N] maps
    UmmaOperand A{frames, C, N, N, CN, F, true}, B{frames, C, N, N, CN, F, true};
    UmmaEpilogue e{};
    e.out = S; e.ldo = N; e.so_b = NN; e.alpha = tau; e.idxA = qa; e.idxB = kb;
    DCNET_TRY(umma_gemm(A, B, nullptr, N, N, C, 0, 0, nprob, e, st));
    softmax_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(S, lse, rows, N, N);
    DCNET_LAUNCH_OK("coattn_fwd.softmax");
    // O[c,i] = sum_j Fb[c,j] P[i,j] : both K-major
    UmmaOperand A2{frames, C, N, N, CN, F, false}, B2{S, N, N, N, NN, nprob, false};
    UmmaEpilogue e2{};
    e2.out = out; e2.ldo = N; e2.so_b = CN; e2.alpha = 1.f; e2.idxA = kb; e2.idxC = oidx;
    return umma_gemm(A2, B2, nullptr, C, N, N, 0, 0, nprob, e2, st);
  }
  // S' = tau Fa^T Fb
  DCNET_TRY(sgemm_launch(frames, frames, S, N, N, C, nprob, 1, 1, N, CN, 0, N, 1, CN, 0, N, 1, NN, qa, kb, nullptr, tau, 0.f,
                         nullptr, 0, 0, st));
  softmax_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(S, lse, rows, N, N);
  DCNET_LAUNCH_OK("coattn_fwd.softmax");
  // O[c,i] = sum_j Fb[c,j] P[i,j]
  DCNET_TRY(sgemm_launch(frames, S, out, C, N, N, nprob, 1, N, 1, CN, 0, 1, N, NN, 0, N, 1, CN, kb, nullptr, oidx, 1.f, 0.f,
                         nullptr, 0, 0, st));
  return 0;
}

extern "C" int dcnet_coattn_bwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                const float* out, int n_out, const float* lse, const float* dout, float* dframes,
                                int C, int N, float tau, int precision, const void* staged, void* workspace, size_t workspace_bytes,
                                void* stream) {
  return dcnet_coattn_bwd_ex(frames, F, qa, kb, oidx, nprob, out, n_out, lse, dout, nullptr, dframes, C, N, tau, precision, staged, nullptr,
                             nullptr, workspace, workspace_bytes, stream);
}

extern "C" int dcnet_coattn_bwd_ex(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                   const float* out, int n_out, const float* lse, const float* dout, const unsigned int* dout_absmax,
                                   float* dframes, int C, int N, float tau, int precision, const void* staged, const void* e_keep,
                                   const float* r_keep, void* workspace, size_t workspace_bytes, void* stream) {
  DCNET_CHECK_ARG(frames && qa && kb && oidx && lse && dout && dframes && nprob >= 0 && C > 0 && N > 0 && F > 0 && n_out > 0, "coattn_bwd: bad arguments");
  DCNET_CHECK_ARG(out || precision != 2, "coattn_bwd: the saved forward output is needed (delta = <dO, O>)");
  if (nprob == 0) return 0;
  DCNET_CHECK_ARG(workspace && workspace_bytes >= dcnet_coattn_workspace_bytes(F, nprob, C, N, precision), "coattn_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const long long CN = (long long)C * N, NN = (long long)N * N;
  if (precision >= 1 && N % 4 != 0 && C % 128 == 0 && reinterpret_cast<uintptr_t>(workspace) % 16 == 0) {
    // odd row pitch (N = 169 at 416x416): the five contractions still run on tcgen05, on copies of the maps / dout whose pitch is
    // padded to 4 positions and with P / dP at that pitch; the gradients land in dframes (pitch N) through the per-thread
    // atomicAdd epilogue of the one-tile-per-CTA kernel
    const int ld = pitch4(N);
    const long long NL = (long long)N * ld, CL = (long long)C * ld;
    char* w = (char*)workspace;
    float* Pp = (float*)w; w += align256((size_t)nprob * NL * sizeof(float));
    float* dPp = (float*)w; w += align256((size_t)nprob * NL * sizeof(float));
    float* Fp = (float*)w; w += align256((size_t)F * CL * sizeof(float));
    float* Gp = (float*)w;
    const long long rows_ = (long long)nprob * N;
    pad_pitch_kernel<<<148 * 8, 256, 0, st>>>(frames, nullptr, Fp, (long long)F * C, C, N, ld);
    DCNET_LAUNCH_OK("coattn_bwd.pad");
    pad_pitch_kernel<<<148 * 8, 256, 0, st>>>(dout, oidx, Gp, (long long)nprob * C, C, N, ld);
    DCNET_LAUNCH_OK("coattn_bwd.pad");
    UmmaOperand Fmn{Fp, C, N, ld, CL, F, true}, Fk{Fp, C, N, ld, CL, F, false};
    UmmaOperand Gmn{Gp, C, N, ld, CL, nprob, true}, Gk{Gp, C, N, ld, CL, nprob, false};
    UmmaOperand Pmn{Pp, N, N, ld, NL, nprob, true};
    UmmaOperand dSk{dPp, N, N, ld, NL, nprob, false}, dSmn{dPp, N, N, ld, NL, nprob, true};
    UmmaEpilogue e{};
    e.out = Pp; e.ldo = ld; e.so_b = NL; e.alpha = tau; e.idxA = qa; e.idxB = kb;
    DCNET_TRY(umma_gemm(Fmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    softmax_rows_kernel<<<ceil_div(rows_ * 32, 256), 256, 0, st>>>(Pp, nullptr, rows_, N, ld);   // re-normalised from these logits
    DCNET_LAUNCH_OK("coattn_bwd.softmax");
    e = UmmaEpilogue{}; e.out = dPp; e.ldo = ld; e.so_b = NL; e.alpha = 1.f; e.idxB = kb;
    DCNET_TRY(umma_gemm(Gmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxC = kb;
    DCNET_TRY(umma_gemm(Gk, Pmn, nullptr, C, N, N, 0, 0, nprob, e, st));
    softmax_bwd_rows_kernel<<<ceil_div(rows_ * 32, 256), 256, 0, st>>>(Pp, dPp, rows_, N, ld, tau);
    DCNET_LAUNCH_OK("coattn_bwd.softmax");
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = kb; e.idxC = qa;
    DCNET_TRY(umma_gemm(Fk, dSk, nullptr, C, N, N, 0, 0, nprob, e, st));
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = qa; e.idxC = kb;
    return umma_gemm(Fk, dSmn, nullptr, C, N, N, 0, 0, nprob, e, st);
  }
  if (precision == 2 && staged && g_bwd_fp16 && N % 4 == 0 && C % 128 == 0 && N >= 64 && coattn_tc_ok(precision, C, N, frames, dout, dframes) &&
      reinterpret_cast<uintptr_t>(workspace) % 256 == 0 && reinterpret_cast<uintptr_t>(staged) % 256 == 0) {
    // ---- fp16 pipeline (see above): the five contractions on kind::f16 with operands that carry tf32's precision
    const int ldh = pitch8h(N);
    const long long NLh = (long long)N * ldh, CLh = (long long)C * ldh;
    const __half* F16 = reinterpret_cast<const __half*>(staged);            // [F][C][ldh]: the forward's staging (same pitch rule)
    char* w = (char*)workspace;
    __half* E16 = (__half*)w; w += align256((size_t)nprob * NLh * 2);
    __half* dS16 = (__half*)w; w += align256((size_t)nprob * NLh * 2);
    __half* dO16 = (__half*)w; w += align256((size_t)nprob * CLh * 2);
    __half* dOs16 = (__half*)w; w += align256((size_t)nprob * CLh * 2);
    float* rsum = (float*)w; w += align256((size_t)nprob * N * sizeof(float));
    float* delta = (float*)w; w += align256((size_t)nprob * N * sizeof(float));
    float* rho = (float*)w; w += align256((size_t)nprob * N * sizeof(float));
    unsigned int* mx = (unsigned int*)w; w += align256((size_t)nprob * sizeof(float));
    float* alpha_z = (float*)w;
    DCNET_CUDA(cudaMemsetAsync(rsum, 0, (size_t)((char*)alpha_z - (char*)rsum), st), "coattn_bwd.memset");
    auto H = [](const __half* p_, long long rows, long long cols, long long ld, long long bs, int nb, bool mn) {
      UmmaOperand o{reinterpret_cast<const float*>(p_), rows, cols, ld, bs, nb, mn};
      o.bf16 = true; o.f16 = true;
      return o;
    };
    const UmmaOperand Fmn = H(F16, C, N, ldh, CLh, F, true), Fk = H(F16, C, N, ldh, CLh, F, false);
    const UmmaOperand Gmn = H(dO16, C, N, ldh, CLh, nprob, true), Gsk = H(dOs16, C, N, ldh, CLh, nprob, false);

    const UmmaOperand dSk = H(dS16, N, N, ldh, NLh, nprob, false), dSmn = H(dS16, N, N, ldh, NLh, nprob, true);
    const int mblk = ceil_div(N, 128) * nprob;
    const int zsl = mblk >= 592 ? 1 : (592 / mblk > 8 ? 8 : 592 / mblk);
    const unsigned int* mxp = mx;
    const int mx_by_oidx = dout_absmax ? 1 : 0;
    if (dout_absmax) {
      mxp = dout_absmax;            // max |dout[row]| per output row, left by the GEMM that produced dout (dcnet_conv1x1_bwd_data_absmax)
    } else {
      coattn_absmax_kernel<<<dim3(8, nprob), 256, 0, st>>>(dout, oidx, mx, CN / 4);
      DCNET_LAUNCH_OK("coattn_bwd.absmax");
    }
    UmmaEpilogue e{};
    int e_t = 0;
    if (e_keep && r_keep && g_bwd_use_keep) {
      // the forward kept its unnormalised weights -- transposed, E^T[z][key][q] (fp16, same pitch) -- and their row sums: no recomputation
      // of S.  P = E / r holds exactly for the values the MMAs read, whatever common factor E carries.  The dS epilogue reads its E tile
      // transposed, and dFb += dOs E takes E^T as a K-major operand
      E16 = const_cast<__half*>(reinterpret_cast<const __half*>(e_keep));
      e_t = 1;
      DCNET_CUDA(cudaMemcpyAsync(rsum, r_keep, (size_t)nprob * N * sizeof(float), cudaMemcpyDeviceToDevice, st), "coattn_bwd.rsum");
    } else {
      // E' = 2^8 exp(tau S - lse_fwd) (fp16) with row sums r'
      e.out = reinterpret_cast<float*>(E16); e.out_f16 = 1; e.ldo = ldh; e.so_b = NLh; e.alpha = tau; e.idxA = qa; e.idxB = kb;
      e.epi_exp = 1; e.u = lse; e.ldu = N; e.exp_shift = E_SHIFT; e.sum = rsum; e.sum_ldz = N;
      DCNET_TRY(umma_gemm(Fmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    }
    if (g_bwd_stop == 1) return 0;
    // E as the B operand of dFb += dOs E (reduction over the queries): [q][k] memory = MN-major, [k][q] memory = K-major
    const UmmaOperand Emn = H(E16, N, N, ldh, NLh, nprob, e_t == 0);
    coattn_delta16_kernel<<<dim3(ceil_div(N, 128), nprob, zsl), dim3(32, 8), 0, st>>>(dout, out, oidx, mxp, mx_by_oidx, rsum, delta, dO16, alpha_z, C, N, ldh);
    DCNET_LAUNCH_OK("coattn_bwd.delta16");
    // s dS' = tau (s dP - s delta) E' / r' (fp16) with row sums s rho
    e = UmmaEpilogue{};
    e.out = reinterpret_cast<float*>(dS16); e.out_f16 = 1; e.ldo = ldh; e.so_b = NLh; e.alpha = tau; e.idxB = kb;
    e.epi_exp = 2; e.u = delta; e.u2 = rsum; e.ldu = N; e.cc = reinterpret_cast<const float*>(E16); e.ldcc = ldh; e.cc_sb = NLh; e.cc_t = e_t; e.sum = rho; e.sum_ldz = N;
    DCNET_TRY(umma_gemm(Gmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    if (g_bwd_stop == 2) return 0;
    coattn_fix16_kernel<<<dim3(ceil_div(N, 128), nprob, zsl), dim3(32, 8), 0, st>>>(dout, out, F16, oidx, qa, rsum, rho, mxp, mx_by_oidx, dOs16, dframes, C, N, ldh);
    DCNET_LAUNCH_OK("coattn_bwd.fix16");
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.alpha_z = alpha_z; e.atomic = 1; e.idxC = kb; e.k_chunks = -1;
    DCNET_TRY(umma_gemm(Gsk, Emn, nullptr, C, N, N, 0, 0, nprob, e, st));
    if (g_bwd_stop == 3) return 0;
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.alpha_z = alpha_z; e.atomic = 1; e.idxA = kb; e.idxC = qa; e.k_chunks = -1;
    DCNET_TRY(umma_gemm(Fk, dSk, nullptr, C, N, N, 0, 0, nprob, e, st));
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.alpha_z = alpha_z; e.atomic = 1; e.idxA = qa; e.idxC = kb; e.k_chunks = -1;
    return umma_gemm(Fk, dSmn, nullptr, C, N, N, 0, 0, nprob, e, st);
  }
  const int chunk = (precision == 2) ? coattn_bwd_chunk(nprob, N) : nprob;
  float* P = (float*)workspace;
  float* dP = (float*)((char*)workspace + align256((size_t)chunk * NN * sizeof(float)));
  const bool tc = coattn_tc_ok(precision, C, N, frames, dout, dframes) && coattn_tc_ok(precision, C, N, P, dP, P);
  const int nprob_all = nprob;
  const int* qa_all = qa; const int* kb_all = kb; const int* oidx_all = oidx; const float* lse_all = lse;
  for (int p0 = 0; p0 < nprob_all; p0 += chunk) {
  nprob = (nprob_all - p0 < chunk) ? nprob_all - p0 : chunk;
  qa = qa_all + p0; kb = kb_all + p0; oidx = oidx_all + p0; lse = lse_all + (long long)p0 * N;
  const long long rows = (long long)nprob * N;
  // precision 2: the forward's lse comes from bf16 logits; P is re-normalised from the logits recomputed here so that its rows
  // sum to one in this precision (a 5e-4 logit mismatch times tau would otherwise show up as a 5e-3 error in P)
  auto exp_launch = [&]() {
    if (precision == 2) {
      softmax_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(P, nullptr, rows, N, N);
    } else {
      long long g = (rows * N + 255) / 256;
      exp_lse_kernel<<<(int)(g > 148 * 16 ? 148 * 16 : g), 256, 0, st>>>(P, lse, rows, N);
    }
  };
  if (tc) {
    UmmaOperand Fmn{frames, C, N, N, CN, F, true}, Fk{frames, C, N, N, CN, F, false};
    UmmaOperand Gmn{dout, C, N, N, CN, n_out, true}, Gk{dout, C, N, N, CN, n_out, false};
    UmmaOperand Pmn{P, N, N, N, NN, nprob, true};
    UmmaOperand dSk{dP, N, N, N, NN, nprob, false}, dSmn{dP, N, N, N, NN, nprob, true};
    UmmaEpilogue e{};
    e.out = P; e.ldo = N; e.so_b = NN; e.alpha = tau; e.idxA = qa; e.idxB = kb;
    if (precision == 2) {
      // Fused epilogues: no pass over an N x N tensor outside the five contractions.
      //   E = exp(tau S - lse_fwd) with row sums r        (epilogue of S = Fa^T Fb; lse_fwd is only a shift: P = E / r exactly sums to 1
      //                                                     in THIS precision, whatever the forward's bf16 logits gave)
      //   delta = <dO, O>, r <- 1/r                         (one pass over two [C,N] maps)
      //   dS' = tau (dP - delta) E / r with row sums rho    (epilogue of dP = dO^T Fb, E tile read back from L2 / HBM)
      //   dOs = (dO - rho Fa) / r, dFa -= rho O             (first-order repair of delta, see coattn_fix_kernel)
      //   dFb += dOs E ; dFa += Fb dS'^T ; dFb += Fa dS'   (TMA reduce-add; the reduction split over CTAs when a problem has few tiles)
      char* w = (char*)dP + align256((size_t)chunk * NN * sizeof(float));
      float* dOs = (float*)w; w += align256((size_t)chunk * CN * sizeof(float));
      float* rsum = (float*)w; w += align256((size_t)chunk * N * sizeof(float));
      float* delta = (float*)w; w += align256((size_t)chunk * N * sizeof(float));
      float* rho = (float*)w;
      // rsum | delta | rho are contiguous: one fill for the three accumulators
      DCNET_CUDA(cudaMemsetAsync(rsum, 0, (size_t)((char*)rho - (char*)rsum) + (size_t)nprob * N * sizeof(float), st), "coattn_bwd.memset");
      const int mblk = ceil_div(N, 128) * nprob;
      const int zsl = mblk >= 592 ? 1 : (592 / mblk > 8 ? 8 : 592 / mblk);     // >= 4 blocks per SM
      e.epi_exp = 1; e.u = lse; e.ldu = N; e.sum = rsum; e.sum_ldz = N;
      DCNET_TRY(umma_gemm(Fmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
      coattn_delta_kernel<<<dim3(ceil_div(N, 128), nprob, zsl), dim3(32, 8), 0, st>>>(dout, out, oidx, rsum, delta, C, N);
      DCNET_LAUNCH_OK("coattn_bwd.delta");
      e = UmmaEpilogue{}; e.out = dP; e.ldo = N; e.so_b = NN; e.alpha = tau; e.idxA = oidx; e.idxB = kb;
      e.epi_exp = 2; e.u = delta; e.u2 = rsum; e.ldu = N; e.cc = P; e.ldcc = N; e.cc_sb = NN; e.sum = rho; e.sum_ldz = N;
      DCNET_TRY(umma_gemm(Gmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
      coattn_fix_kernel<<<dim3(ceil_div(N, 128), nprob, zsl), dim3(32, 8), 0, st>>>(dout, out, frames, oidx, qa, rsum, rho, dOs, dframes, C, N);
      DCNET_LAUNCH_OK("coattn_bwd.fix");
      UmmaOperand Gsk{dOs, C, N, N, CN, nprob, false};
      e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxC = kb; e.k_chunks = -1;
      DCNET_TRY(umma_gemm(Gsk, Pmn, nullptr, C, N, N, 0, 0, nprob, e, st));
      e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = kb; e.idxC = qa; e.k_chunks = -1;
      DCNET_TRY(umma_gemm(Fk, dSk, nullptr, C, N, N, 0, 0, nprob, e, st));
      e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = qa; e.idxC = kb; e.k_chunks = -1;
      DCNET_TRY(umma_gemm(Fk, dSmn, nullptr, C, N, N, 0, 0, nprob, e, st));
      continue;
    }
    // precision 1: P = exp(tau S - lse) with the (tf32) forward's own lse
    DCNET_TRY(umma_gemm(Fmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    exp_launch();
    DCNET_LAUNCH_OK("coattn_bwd.exp");
    // dP[i,j] = sum_c dO[c,i] Fb[c,j]
    e = UmmaEpilogue{}; e.out = dP; e.ldo = N; e.so_b = NN; e.alpha = 1.f; e.idxA = oidx; e.idxB = kb;
    DCNET_TRY(umma_gemm(Gmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    // dFb[c,j] += sum_i dO[c,i] P[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = oidx; e.idxC = kb; e.k_chunks = -1;
    DCNET_TRY(umma_gemm(Gk, Pmn, nullptr, C, N, N, 0, 0, nprob, e, st));
    softmax_bwd_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(P, dP, rows, N, N, tau);
    DCNET_LAUNCH_OK("coattn_bwd.softmax");
    // dFa[c,i] += sum_j Fb[c,j] dS[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = kb; e.idxC = qa; e.k_chunks = -1;
    DCNET_TRY(umma_gemm(Fk, dSk, nullptr, C, N, N, 0, 0, nprob, e, st));
    // dFb[c,j] += sum_i Fa[c,i] dS[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = qa; e.idxC = kb; e.k_chunks = -1;
    DCNET_TRY(umma_gemm(Fk, dSmn, nullptr, C, N, N, 0, 0, nprob, e, st));
    continue;
  }
  // recompute P = exp(tau S - lse)
  DCNET_TRY(sgemm_launch(frames, frames, P, N, N, C, nprob, 1, 1, N, CN, 0, N, 1, CN, 0, N, 1, NN, qa, kb, nullptr, tau, 0.f,
                         nullptr, 0, 0, st));
  exp_launch();
  DCNET_LAUNCH_OK("coattn_bwd.exp");
  // dP[i,j] = sum_c dO[c,i] Fb[c,j]
  DCNET_TRY(sgemm_launch(dout, frames, dP, N, N, C, nprob, 1, 1, N, CN, 0, N, 1, CN, 0, N, 1, NN, oidx, kb, nullptr, 1.f, 0.f,
                         nullptr, 0, 0, st));
  // dFb += dO P
  DCNET_TRY(sgemm_launch(dout, P, dframes, C, N, N, nprob, 1, N, 1, CN, 0, N, 1, NN, 0, N, 1, CN, oidx, nullptr, kb, 1.f, 0.f,
                         nullptr, 0, 1, st));
  softmax_bwd_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(P, dP, rows, N, N, tau);
  DCNET_LAUNCH_OK("coattn_bwd.softmax");
  // dFa[c,i] += sum_j Fb[c,j] dS[i,j]
  DCNET_TRY(sgemm_launch(frames, dP, dframes, C, N, N, nprob, 1, N, 1, CN, 0, 1, N, NN, 0, N, 1, CN, kb, nullptr, qa, 1.f, 0.f,
                         nullptr, 0, 1, st));
  // dFb[c,j] += sum_i Fa[c,i] dS[i,j]
  DCNET_TRY(sgemm_launch(frames, dP, dframes, C, N, N, nprob, 1, N, 1, CN, 0, N, 1, NN, 0, N, 1, CN, qa, nullptr, kb, 1.f, 0.f,
                         nullptr, 0, 1, st));
  }   // chunks
  return 0;
}
