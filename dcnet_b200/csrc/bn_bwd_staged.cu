// Backward of bn_act_fwd, phase 1 (dcnet_bn_act_bwd_reduce) for C = 512 and 16-byte-addressable maps: a persistent kernel whose
// position tiles are staged asynchronously in shared memory.
//
// Same math as bn_act_bwd_reduce_kernel (conv_bn.cu): recompute the forward from z, form d(pre-activation) from dy (+ the
// pixel-to-text gradients dsim / dneg_sim), write it to dv, and accumulate sum_dv[c], sum_dvz[c] (and dfa[b,c]).
// What changes is the data movement.  The register-staged kernel keeps 64 loads per thread in flight, then computes, then
// stores -- with 128 registers one 512-thread CTA fits an SM, so nothing overlaps the three phases (23 % of the HBM roofline).
// Here one CTA per SM walks (image, 16-position tile) items; two producer warps keep three 64 KiB stages (z tile + dy tile,
// [512 channels][16 positions]) in flight with 16-byte cp.async (LDGSTS) completing on an mbarrier, sixteen compute warps work
// out of shared memory, dv is written over the z tile in place and leaves as float4 rows, and the per-channel sums stay in
// registers across all items of the CTA (2 atomics per thread at the end instead of 2 per thread per tile).
// (A first version staged the tiles with cp.async.bulk.tensor boxes of 256 rows x 64 B: 10 k cycles per item even without the
// normalisation terms -- the TMA unit is bound by the number of 64-byte rows, not by bytes.  Measured: scripts/prof_bn.py.)
#include "common.cuh"
#include "umma.cuh"

using namespace umma;

namespace {

constexpr int C = 512;
constexpr int PT = 16;                      // positions per tile
#ifndef BNB_NG
#define BNB_NG 32
#endif
constexpr int NG = BNB_NG;                  // channel groups = half-warps (2 per compute warp); a thread owns channels g + NG i
constexpr int CPT = C / NG;                 // 16 (32 groups, 16 compute warps) or 32 (16 groups, 8 compute warps)
constexpr int OUTN = CPT / 16;              // totals a lane holds after the half-warp vector reduction
constexpr int TILE_BYTES = C * PT * 4;      // 32 KiB
constexpr int STAGE_BYTES = 2 * TILE_BYTES; // z + dy
constexpr int NST = 3;
constexpr int NCOMPUTE = NG * 16;
constexpr int NPROD = 64;                    // two producer warps
constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;

struct BwdP {
  const float* z; const float* dy; float* dv;
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  const float* fa; const float* fa_neg; const float* dsim; const float* dneg;
  float* sum_dv; float* sum_dvz; float* dfa; float* dfa_neg;
  float slope; int l2norm, has_dy, B, N, tiles, items;
};

// sum over the 16 lanes of a half-warp of a per-lane vector v[0..CPT): recursive halving, CPT-OUTN shuffles; on return the lane
// holds in v[0..OUTN) the totals of the original indices half_reduce_index(lane) + {0 .. OUTN-1}
__device__ __forceinline__ void half_vec_reduce(float (&v)[CPT], int lane) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int bit = 8 >> s;
    const int h = (CPT / 2) >> s;
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int k = 0; k < h; k++) {
      const float keep = up ? v[h + k] : v[k];
      const float send = up ? v[k] : v[h + k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
}
__device__ __forceinline__ int half_reduce_index(int lane) {
  int idx = 0;
#pragma unroll
  for (int s = 0; s < 4; s++) idx += (lane & (8 >> s)) ? ((CPT / 2) >> s) : 0;
  return idx;
}

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory"); }

// 16-byte asynchronous copy global -> shared; src_bytes = 0 zero-fills (positions beyond N)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all its cp.async issued so far have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(NCOMPUTE + NPROD, 1)
bn_bwd_reduce_staged_kernel(const BwdP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ float s_scale[C], s_shift[C], s_mean[C], s_istd[C], s_fa[C], s_fr[C];
  __shared__ float red[2][NG][PT];
  __shared__ uint64_t full[NST], empty[NST];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], NCOMPUTE); }
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float is = p.invstd[c], sc = p.gamma[c] * is;
    s_scale[c] = sc;
    s_shift[c] = p.beta[c] - p.mean[c] * sc;
    s_mean[c] = p.mean[c];
    s_istd[c] = is;
  }
  __syncthreads();

  // every CTA walks a CONTIGUOUS range of (image, tile) items: consecutive items mostly belong to the same image, so the text-vector
  // gradients of an image accumulate in registers and leave with one round of atomics per image instead of one per tile
  const int per_cta = (p.items + gridDim.x - 1) / gridDim.x;
  const int it_begin = blockIdx.x * per_cta, it_end = min(p.items, it_begin + per_cta);
  if (warp >= NCOMPUTE / 32) {
    // ------------------------------------------------------------ producers: 64 threads x 16-byte chunks
    const int pt = threadIdx.x - NCOMPUTE;
    uint32_t k = 0;
    for (int it = it_begin; it < it_end; it++, k++) {
      const int s = k % NST;
      const uint32_t ph = (k / NST) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      const int b = it / p.tiles, n0 = (it - b * p.tiles) * PT;
      uint8_t* st = smem + s * STAGE_BYTES;
      const float* zb = p.z + (long long)b * C * p.N + n0;
      const float* db = p.dy + (long long)b * C * p.N + n0;
      // chunk id = 4 c + q  (channel c, positions 4q .. 4q+3); consecutive threads take consecutive chunks: 64-byte rows
#pragma unroll 8
      for (int id = pt; id < C * (PT / 4); id += NPROD) {
        const int c = id >> 2, q = id & 3;
        const uint32_t nb = (n0 + 4 * q < p.N) ? 16u : 0u;
        cp_async16(st + id * 16, zb + (long long)c * p.N + 4 * q, nb);
      }
      if (p.has_dy) {
#pragma unroll 8
        for (int id = pt; id < C * (PT / 4); id += NPROD) {
          const int c = id >> 2, q = id & 3;
          const uint32_t nb = (n0 + 4 * q < p.N) ? 16u : 0u;
          cp_async16(st + TILE_BYTES + id * 16, db + (long long)c * p.N + 4 * q, nb);
        }
      }
      cp_async_arrive(&full[s]);
    }
    return;
  }

  // -------------------------------------------------------------- compute warps: thread = (position pp, channel group g)
  const int pp = lane & 15;
  const int g = 2 * warp + (lane >> 4);
  const int ridx = half_reduce_index(lane);
  float acc_dv[OUTN], acc_dvz[OUTN];
#pragma unroll
  for (int o = 0; o < OUTN; o++) acc_dv[o] = acc_dvz[o] = 0.f;
  float acc_fa[2][OUTN];
#pragma unroll
  for (int o = 0; o < OUTN; o++) acc_fa[0][o] = acc_fa[1][o] = 0.f;
  auto flush_fa = [&](int b) {
    if (b < 0 || !(p.fa && p.dfa)) return;
#pragma unroll
    for (int which = 0; which < 2; which++) {
      float* dst = which == 0 ? p.dfa + (long long)b * C
                              : (p.fa_neg ? (p.dfa_neg ? p.dfa_neg + (long long)b * C : nullptr) : p.dfa + (long long)(p.B - 1 - b) * C);
#pragma unroll
      for (int o = 0; o < OUTN; o++) {
        if (dst) atomicAdd(dst + g + NG * (ridx + o), acc_fa[which][o]);
        acc_fa[which][o] = 0.f;
      }
    }
  };
  uint32_t k = 0;
  int cur_b = -1;
  for (int it = it_begin; it < it_end; it++, k++) {
    const int s = k % NST;
    const uint32_t ph = (k / NST) & 1u;
    const int b = it / p.tiles, n0 = (it - b * p.tiles) * PT;
    const int n = n0 + pp;
    const bool valid = n < p.N;
    if (p.fa && b != cur_b) {
      flush_fa(cur_b);
      // text vectors of this image (and of its negative partner)
      compute_sync();      // every thread is done with the previous image's vectors
      for (int c = threadIdx.x; c < C; c += NCOMPUTE) {
        s_fa[c] = p.fa[(long long)b * C + c];
        s_fr[c] = p.fa_neg ? p.fa_neg[(long long)b * C + c] : p.fa[(long long)(p.B - 1 - b) * C + c];
      }
      compute_sync();
      cur_b = b;
    }
    const float ds = (p.fa && p.dsim && valid) ? p.dsim[(long long)b * p.N + n] : 0.f;
    const float dn = (p.fa && p.dneg && valid) ? p.dneg[(long long)b * p.N + n] : 0.f;
    float* zt = reinterpret_cast<float*>(smem + s * STAGE_BYTES);
    const float* dyt = zt + C * PT;
    mbar_wait(&full[s], ph);

    float a[CPT], gr[CPT];
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int i = 0; i < CPT; i++) {
      const int c = g + NG * i;
      const float t = fmaf(zt[c * PT + pp], s_scale[c], s_shift[c]);
      const float act = t > 0.f ? t : t * p.slope;
      a[i] = t;
      ss = fmaf(act, act, ss);
      float gg = p.has_dy ? dyt[c * PT + pp] : 0.f;
      if (p.fa) gg = fmaf(s_fa[c], ds, fmaf(s_fr[c], dn, gg));
      gr[i] = gg;
      dot = fmaf(gg, act, dot);
    }
    float inv = 1.f, dotn = 0.f;
    if (p.l2norm) {
      red[0][g][pp] = ss;
      red[1][g][pp] = dot;
      compute_sync();
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int q = 0; q < NG; q++) {
        t0 += red[0][q][pp];
        t1 += red[1][q][pp];
      }
      inv = rsqrtf(fmaxf(t0, 1e-24f));
      dotn = t1 * inv * inv;
    }
    if (p.fa && p.dfa) {
#pragma unroll
      for (int which = 0; which < 2; which++) {
        const float w = which == 0 ? ds : dn;
        float f[CPT];
#pragma unroll
        for (int i = 0; i < CPT; i++) {
          const float t = a[i];
          f[i] = w * ((t > 0.f ? t : t * p.slope) * inv);      // w = 0 at positions beyond N
        }
        half_vec_reduce(f, lane);
#pragma unroll
        for (int o = 0; o < OUTN; o++) acc_fa[which][o] += f[o];
      }
    }
#pragma unroll
    for (int i = 0; i < CPT; i++) {
      const int c = g + NG * i;
      const float t = a[i];
      const float act = t > 0.f ? t : t * p.slope;
      const float da = p.l2norm ? (gr[i] * inv - act * inv * dotn) : gr[i];
      // positions beyond N hold zero-filled dy and ds = dn = 0, so gr = dot = 0 there and dpre is exactly 0 without a select
      const float dpre = da * (t > 0.f ? 1.f : p.slope);
      const float zh = (zt[c * PT + pp] - s_mean[c]) * s_istd[c];
      zt[c * PT + pp] = dpre;                 // dv over the z tile, in place (this thread is the only reader of the element)
      gr[i] = dpre;
      a[i] = dpre * zh;
    }
    half_vec_reduce(gr, lane);
    half_vec_reduce(a, lane);
#pragma unroll
    for (int o = 0; o < OUTN; o++) { acc_dv[o] += gr[o]; acc_dvz[o] += a[o]; }
    // dv tile -> global as float4 rows (64 B per channel row), then the stage goes back to the producers
    compute_sync();
    {
      float* ob = p.dv + (long long)b * C * p.N + n0;
#pragma unroll
      for (int id = threadIdx.x; id < C * (PT / 4); id += NCOMPUTE) {
        const int c = id >> 2, q = id & 3;
        if (n0 + 4 * q < p.N) *reinterpret_cast<float4*>(ob + (long long)c * p.N + 4 * q) = *reinterpret_cast<const float4*>(zt + id * 4);
      }
    }
    mbar_arrive(&empty[s]);
  }
  flush_fa(cur_b);
  // channel sums of this CTA
#pragma unroll
  for (int o = 0; o < OUTN; o++) {
    atomicAdd(p.sum_dv + g + NG * (ridx + o), acc_dv[o]);
    atomicAdd(p.sum_dvz + g + NG * (ridx + o), acc_dvz[o]);
  }
}

}  // namespace

// returns BN_BWD_TMA_UNSUPPORTED if the shape / pointers are outside what this kernel handles (caller falls back), 0 on success
int bn_bwd_reduce_staged(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                      int l2norm, const float* dy, const float* fa, const float* fa_neg, const float* dsim, const float* dneg,
                      float* dv, float* sum_dv, float* sum_dvz, float* dfa, float* dfa_neg, int B, int Cc, int N, cudaStream_t st) {
  auto al = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  if (Cc != C || N % 4 != 0 || !al(z) || !al(dv) || (dy && !al(dy)) || B > 65535) return -2147483647;
  BwdP p{};
  p.z = z; p.dy = dy ? dy : z; p.dv = dv;
  p.mean = mean; p.invstd = invstd; p.gamma = gamma; p.beta = beta; p.fa = fa; p.fa_neg = fa_neg; p.dsim = dsim; p.dneg = dneg;
  p.sum_dv = sum_dv; p.sum_dvz = sum_dvz; p.dfa = dfa; p.dfa_neg = dfa_neg; p.slope = slope; p.l2norm = l2norm; p.has_dy = dy ? 1 : 0;
  p.B = B; p.N = N; p.tiles = ceil_div(N, PT); p.items = B * p.tiles;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  DCNET_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES), "bn_bwd_reduce_staged.attr");
  int grid = p.items < sms ? p.items : sms;
  grid = ceil_div(p.items, ceil_div(p.items, grid));      // no CTA with an empty item range
  bn_bwd_reduce_staged_kernel<<<grid, NCOMPUTE + NPROD, SMEM_BYTES, st>>>(p);
  DCNET_LAUNCH_OK("bn_act_bwd_reduce");
  return 0;
}
