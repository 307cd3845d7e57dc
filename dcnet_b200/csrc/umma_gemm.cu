// Batched TF32 GEMM on the 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM) fed by TMA.
//
// Operands are the fp32 tensors exactly as the reference lays them out ([.., C, N] with N innermost): TMA brings 128-byte-swizzled
// tiles into shared memory and the MMA reads them as TF32 (the low 13 mantissa bits are ignored by the hardware), so there is
// no conversion pass and no extra copy in HBM.  Either operand can be K-major (reduction axis contiguous) or MN-major
// (output axis contiguous); that covers every contraction of the hot path without a transpose:
//     conv fwd      z[b]  = W x[b]                A: W      (K-major)   B: x[b]   (MN-major)
//     conv bwd data dx[b] = W^T dz[b]             A: W^T    (MN-major)  B: dz[b]  (MN-major)
//     conv bwd wgt  dW   += dz[b] x[b]^T          A: dz[b]  (K-major)   B: x[b]   (K-major)
//     co-attention  S = Fa^T Fb  (MN,MN);  O = Fb P^T (K,K);  dFb += dO P (K,MN);  dFa += Fb dS^T (K,K);  dFb += Fa dS (K,MN)
//
// CTA = one 128 x BN output tile.  Warp roles (192 threads): warps 0-3 epilogue (TMEM -> registers -> global, warp w owns TMEM
// lanes 32w..32w+31), warp 4 TMA producer (one elected lane), warp 5 TMEM allocator + MMA issuer (one elected lane).
// A STAGES-deep ring of {A tile, B tile} with full/empty mbarriers decouples TMA from MMA; tcgen05.commit releases a slot.
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

using namespace umma;

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
static long long* g_gemm_trace = nullptr;   // profiling: clock stamps of the next persistent GEMM launches
extern "C" int dcnet_gemm_trace(long long* buf) { g_gemm_trace = buf; return 0; }
static int g_gemm_tma_store = 1;   // dcnet_gemm_select(5): epilogue through coalesced st.global / red.global.add.v4 instead of TMA store / reduce-add
static bool g_force_v1 = false;   // tests: run the one-tile-per-CTA kernel with direct stores
// largest cluster the persistent kernel may use.  Measured on B200 (scripts/prof_gemm.py, M=N=1024 K=512 x16, L2 flushed): no
// clusters 44.2 us, clusters of 2 44.2 us, clusters of 4 46 us, and the C2 step 1.73 ms vs 1.79 ms -- the kernel is not bound by
// the L2 -> SM operand traffic that multicast removes, and the cluster lock-step costs a little.  Default: off.
static int g_cluster_max = 1;
// 128x256 tf32 tiles as cta_group::2 pairs (256x256 per cluster); dcnet_gemm_select(6) or DCNET_GEMM_PAIR=0 turns it off
static int g_pair_mode = []() { const char* v = getenv("DCNET_GEMM_PAIR"); return (v && v[0] == '0') ? 0 : 1; }();
// fp16 exp / dS epilogues with 8 epilogue warps (two per SM sub-partition); dcnet_gemm_select(7) keeps 4
static int g_epi_warps8 = 1;
extern "C" int dcnet_gemm_select(int variant) {
  g_epi_warps8 = (variant == 7) ? 0 : 1;
  g_force_v1 = (variant == 1);
  g_cluster_max = (variant == 4) ? 4 : ((variant == 3) ? 2 : 1);
  g_gemm_tma_store = (variant == 5) ? 0 : 1;
  g_pair_mode = (variant == 6) ? 0 : 1;
  return 0;
}


namespace {

#ifndef GEMM2_NSLOT
#define GEMM2_NSLOT 2
#endif
#ifndef GEMM2_STAGES256
#define GEMM2_STAGES256 4
#endif
#ifndef GEMM2_STAGES_PAIR
#define GEMM2_STAGES_PAIR 6        // 32 KiB per stage and CTA in pair mode (+ 32 KiB output slots; one stage less with the 32 KiB E slots)
#endif
constexpr int BM = 128;
constexpr int A_BYTES = BM * 128;    // 16 KiB: 128 rows x one 128-byte swizzle row (32 tf32 or 64 bf16 along the reduction)
// Element-size dependent geometry (EB = 4: fp32 read as TF32, EB = 2: bf16):
//   BK       reduction elements per stage = 128 B / EB (32 | 64); one MMA consumes 32 B of K (8 tf32 | 16 bf16) -> 4 MMAs per stage
//   MN-major operands are stored as blocks of (128 B along MN) x (BK reduction rows): 4 KiB | 8 KiB, LBO = block size.
//   tf32 MN-major must use SWIZZLE_128B_BASE32B (4-row groups, SBO 512 B); bf16 uses plain SWIZZLE_128B (8-row groups, SBO 1024 B).
template <int EB> struct Geo {
  static constexpr int BK = 128 / EB;
  static constexpr int MN_ELEMS = 128 / EB;            // MN elements per block
  static constexpr int BLK_BYTES = BK * 128;
  static constexpr int KSTEP_ROWS = 32 / EB;           // reduction rows per MMA
  static constexpr uint32_t MN_LAYOUT = (EB == 4) ? 1u : 2u;
  static constexpr uint32_t MN_SBO = (EB == 4) ? 512u : 1024u;
  static constexpr uint32_t FMT = (EB == 4) ? FMT_TF32 : FMT_BF16;
};

struct GemmP {
  int M_valid, N_valid;     // store predicates
  int k_iters;              // BK-steps
  int k_split;              // B MN-major: iterations >= k_split read mapB2 (second K source), 0 = off
  int n_split;              // B K-major : output columns >= n_split read mapB2 (rows n - n_split), 0 = off
  int m_split;              // store: rows >= m_split go to out2 (row m - m_split), 0 = off
  int a_batched, b_batched; // coordinate 2 of the operand = batch index (else 0)
  const int* idxA; const int* idxB; const int* idxC;   // optional batch indirection (frame / problem indices)
  float* out; long long ldo, so_b;
  float* out2; long long ldo2, so_b2;
  float alpha;
  int atomic;               // 1: atomicAdd into out (split reductions / shared gradients)
  const float* u; int ldu;  // + u[b*ldu + m]
  const float* cc; long long ldcc;   // + cc[m*ldcc + n]
  float* sum; float* sumsq; // per-row (channel) sums of the stored values and their squares (BatchNorm statistics)
};

template <bool A_MN, bool B_MN, int BN, int STAGES, int EB>
__global__ void __launch_bounds__(192, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapB2, const GemmP p) {
  using G = Geo<EB>;
  constexpr int BK = G::BK;
  constexpr int BLK_BYTES = G::BLK_BYTES;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, z = blockIdx.z;

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapB2);
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      const int a_b = p.a_batched ? (p.idxA ? p.idxA[z] : z) : 0;
      const int b_b = p.b_batched ? (p.idxB ? p.idxB[z] : z) : 0;
      for (int it = 0; it < p.k_iters; it++) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sA = smem + s * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        if constexpr (A_MN) {
#pragma unroll
          for (int j = 0; j < BM / G::MN_ELEMS; j++) tma_load_3d(sA + j * BLK_BYTES, &mapA, &full[s], m0 + G::MN_ELEMS * j, it * BK, a_b);
        } else {
          tma_load_3d(sA, &mapA, &full[s], it * BK, m0, a_b);
        }
        if constexpr (B_MN) {
          const bool src2 = p.k_split > 0 && it >= p.k_split;
          const CUtensorMap* mb = src2 ? &mapB2 : &mapB;
          const int kc = (src2 ? it - p.k_split : it) * BK;
#pragma unroll
          for (int j = 0; j < BN / G::MN_ELEMS; j++) tma_load_3d(sB + j * BLK_BYTES, mb, &full[s], n0 + G::MN_ELEMS * j, kc, b_b);
        } else {
          const bool src2 = p.n_split > 0 && n0 >= p.n_split;
          tma_load_3d(sB, src2 ? &mapB2 : &mapB, &full[s], it * BK, src2 ? n0 - p.n_split : n0, b_b);
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      constexpr uint32_t idesc = instr_desc(G::FMT, BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int it = 0; it < p.k_iters; it++) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sB = sA + A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {          // one MMA = 32 bytes of K (8 tf32 / 16 bf16)
          // K-major : SWIZZLE_128B; +32 B inside the 128-B swizzle row; rows 128 B apart, 8-row atoms 1024 B apart
          // MN-major: +KSTEP_ROWS reduction rows = +KSTEP_ROWS*128 B; k-row groups MN_SBO apart; MN blocks BLK_BYTES apart (LBO)
          constexpr uint32_t kadv = G::KSTEP_ROWS * 128;
          const uint64_t ad = A_MN ? smem_desc(sA + ks * kadv, BLK_BYTES, G::MN_SBO, G::MN_LAYOUT) : smem_desc(sA + ks * 32, 16, 1024, 2);
          const uint64_t bd = B_MN ? smem_desc(sB + ks * kadv, BLK_BYTES, G::MN_SBO, G::MN_LAYOUT) : smem_desc(sB + ks * 32, 16, 1024, 2);
          if constexpr (EB == 4) mma_tf32(tmem_base, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
          else mma_bf16(tmem_base, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
        }
        mma_commit(&empty[s]);      // slot reusable once these MMAs have read it
      }
      mma_commit(accum);            // accumulator complete
    }
  } else {
    // ---------------- epilogue: warp w <-> TMEM lanes 32w..32w+31 <-> output rows m0+32w..
    mbar_wait(accum, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    const bool row_ok = row < p.M_valid;
    const int zc = p.idxC ? p.idxC[z] : z;
    float* orow;
    if (p.m_split > 0 && m0 >= p.m_split) orow = p.out2 + (long long)zc * p.so_b2 + (long long)(row - p.m_split) * p.ldo2;
    else orow = p.out + (long long)zc * p.so_b + (long long)row * p.ldo;
    const float bias = (p.u && row_ok) ? p.u[(long long)z * p.ldu + row] : 0.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; c++) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (row_ok && nb < p.N_valid) {
      if (p.cc) {
        const float* cr = p.cc + (long long)row * p.ldcc + nb;
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = fmaf(p.alpha, v[e], bias + ((nb + e < p.N_valid) ? cr[e] : 0.f));
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = fmaf(p.alpha, v[e], bias);
      }
      if (p.atomic) {
#pragma unroll
        for (int e = 0; e < 32; e++)
          if (nb + e < p.N_valid) atomicAdd(orow + nb + e, v[e]);
      } else if (nb + 32 <= p.N_valid) {
        float4* o4 = reinterpret_cast<float4*>(orow + nb);
#pragma unroll
        for (int e = 0; e < 8; e++) o4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
#pragma unroll
        for (int e = 0; e < 32; e++) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
      } else {
        for (int e = 0; e < 32 && nb + e < p.N_valid; e++) { orow[nb + e] = v[e]; s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
      }
      }
      __syncwarp();
    }
    if (p.sum && row_ok) {
      atomicAdd(p.sum + row, s1);
      atomicAdd(p.sumsq + row, s2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, BN);
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant (the default): each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... ; two TMEM accumulator
// stages let the epilogue of tile i run under the main loop of tile i+1; the epilogue leaves through 128-byte-swizzled smem
// slots and TMA (plain store, or fp32 reduce-add for split reductions / shared gradients) instead of per-thread stores whose
// lanes hit 32 different rows.  Each epilogue warp owns its 32 rows end to end (2 x 4 KiB slots), so no cross-warp barrier.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Gemm2P {
  int M_valid, N_valid, k_iters, k_split, n_split, m_split;
  int a_batched, b_batched, out_batched;
  const int* idxA; const int* idxB; const int* idxC;
  int tiles_m, tiles_n, ntiles;
  float alpha; int atomic;
  const float* u; int ldu; const float* cc; long long ldcc; float* sum; float* sumsq;
  int epi_exp;
  const float* u2; long long cc_sb; long long sum_ldz;
  int k_chunks, k_per;   // split-K (reduce-add outputs): work item = (tile, chunk); chunk c covers k iterations [c*k_per, min(k_iters, (c+1)*k_per))
  // 3x3 convolution as an implicit GEMM over three column-shifted copies of the map (B = x(dx=-1), B2 = x, B3 = x(dx=+1); conv3x3.cu):
  //   tap_kper > 0 (B MN-major; forward / data gradient): k iteration `it` belongs to tap t = it / tap_kper = (dy+1)*3 + (dx+1); A tile
  //     from "batch" t of A (the weight is stored [tap][Cout][Cin]); B tile from copy dx at positions shifted by dy * tap_w (TMA
  //     zero-fills rows above / below the image); tap_flip: the B side uses tap 8 - t (transposed convolution of the data gradient)
  //   tap_n > 0 (B K-major; weight gradient): output column block n0 belongs to tap t = n0 / tap_n; B rows from copy dx, reduction
  //     coordinate (positions) shifted by dy * tap_w
  int tap_kper, tap_w, tap_flip, tap_n;
  int f16;             // 2-byte operands are __half (FMT_F16) instead of __nv_bfloat16
  int cc_t;            // fp16 dS epilogue: E is stored transposed ([col][row]); its tile arrives as [64 k rows][32 q] without swizzle
  int out_f16;         // out / the E tile are __half tensors: two 32-column chunks share one 128-byte slot row, values rounded to fp16
  float exp_shift; const float* alpha_z;
  unsigned int* absmax2;
  int direct_store;    // 1: epilogue leaves through coalesced st.global / red.global.add.v4 instead of TMA store / reduce
  float* out; long long ldo, so_b; float* out2; long long ldo2, so_b2;
  long long* trace;    // optional [CTA][tile slot < 8][8] clock stamps (profiling entry point dcnet_gemm_tf32_trace); nullptr = off
};
#define GTRACE(tl, slot) do { if (p.trace && (tl) < 8) p.trace[((long long)blockIdx.x * 8 + (tl)) * 8 + (slot)] = clock64(); } while (0)

// CS = cluster size: CS CTAs with consecutive M tiles of the same (batch, N tile) share the B tile -- each loads 1/CS of it and
// multicasts (the kernel is bound by the L2 -> SM operand traffic, profiles/r1i_ncu_full_umma_gemm2.txt).
// TWO: the CTAs of a cluster of 2 form one cta_group::2 pair: one MMA covers 256 (M) x BN, each CTA's TMEM holds its 128 rows, each
// CTA loads its own A tile and HALF of the B tile (no multicast: the MMA reads both halves), so a 256 x 256 output tile costs
// (256 + 256) x K operand elements from L2 instead of 2 x (128 + 256) x K -- the tf32 GEMMs here are bound by that traffic
// (fp32 operands: 4 bytes per element at half the bf16 MMA rate).  The leader CTA (rank 0) issues every MMA.
// ESLOT (pair mode, dS epilogue only): per-warp landing slots for the E tile (one pipeline stage less)
// EW: epilogue warps, 4 or 8.  tcgen05.ld ties a warp to the TMEM lane quarter (warp % 4), so with 8 the warps w and w + 4 share 32 rows
// and split the columns (chunk pairs alternate between them): two epilogue warps per SM sub-partition hide each other's TMEM / barrier /
// store latencies.  Used by the fp16 exp / dS epilogues, which are bound by the epilogue (K = 512: 4 k MMA cycles per tile).
template <bool A_MN, bool B_MN, int BN, int STAGES, int EB, int CS, bool TWO, bool ESLOT = false, int EW = 4>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
umma_gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapB2, const __grid_constant__ CUtensorMap mapB3,
                  const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapO2, const Gemm2P p) {
  using G = Geo<EB>;
  constexpr int BK = G::BK;
  constexpr int BLK_BYTES = G::BLK_BYTES;
  static_assert(!TWO || CS == 2, "a CTA pair is a cluster of 2");
  constexpr int B_BYTES = (TWO ? BN / 2 : BN) * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int SLOT_BYTES = 32 * 128;           // 32 rows x 32 fp32
  constexpr int NSLOT = EW == 8 ? 1 : GEMM2_NSLOT;   // staging slots per epilogue warp (TMA stores in flight per warp)
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + STAGES * STAGE_BYTES;                 // [4 warps][NSLOT][SLOT_BYTES]
  // pair mode only: [4 warps][2][SLOT_BYTES] landing slots for the E tile of the dS epilogue (TMA, same 128-byte-swizzled 32x32 boxes)
  static_assert(!ESLOT || TWO, "E landing slots exist in pair mode only");
  constexpr int E_BYTES = ESLOT ? EW * 2 * SLOT_BYTES : 0;
  uint8_t* stage_e = stage_out + EW * NSLOT * SLOT_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_e + E_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* e_full = acc_empty + 2;        // [4 warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_full + 2 * EW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == EW && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapB2);
    prefetch_tmap(&mapB3);
    prefetch_tmap(&mapO);
    // pair: a slot is released by ONE commit (multicast to both CTAs); the leader's accumulator stage is free once the epilogue
    // warps of BOTH CTAs have drained their halves
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], TWO ? 1 : CS); }
    for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TWO ? 2 * EW : EW); }
    for (int s = 0; s < 2 * EW; s++) mbar_init(&e_full[s], 1);
    fence_barrier_init();
  }
  if constexpr (TWO) cluster_sync_all();             // both CTAs are resident before the pair-wide TMEM allocation
  if (warp == EW + 1) {
    if constexpr (TWO) { tmem_alloc2(tmem_slot, 2 * BN); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();          // every CTA's barriers exist before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // work item = (batch z, N tile, group of CS consecutive M tiles); the CTAs of a cluster take the M tiles of one group
  const int mgroups = (p.tiles_m + CS - 1) / CS;
  const int per_z = mgroups * p.tiles_n;
  const int ngroups = per_z * (p.ntiles / (p.tiles_m * p.tiles_n)) * p.k_chunks;
  const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / CS, ncl = gridDim.x / CS;
  constexpr uint16_t cmask = (uint16_t)((1u << CS) - 1);

  if (warp == EW) {
    if (elect_one()) {
      uint32_t itg = 0;
      for (int t = cid; t < ngroups; t += ncl) {
        const int kch = t % p.k_chunks, tt = t / p.k_chunks;
        const int z = tt / per_z, r = tt - z * per_z;
        const int m0 = ((r / p.tiles_n) * CS + crank) * BM, n0 = (r % p.tiles_n) * BN;
        const int a_b = p.a_batched ? (p.idxA ? p.idxA[z] : z) : 0;
        const int b_b = p.b_batched ? (p.idxB ? p.idxB[z] : z) : 0;
        const int it0 = kch * p.k_per, it1 = min(p.k_iters, it0 + p.k_per);
        for (int it = it0; it < it1; it++, itg++) {
          const int s = itg % STAGES;
          const uint32_t ph = (itg / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sA = smem + s * STAGE_BYTES;
          uint8_t* sB = sA + A_BYTES;
          // where this k step's tiles come from
          int a_k = it * BK, a_z = a_b, b_k = it * BK, b_n = n0;
          const CUtensorMap* mb = &mapB;
          if constexpr (B_MN) {
            if (p.tap_kper > 0) {
              const int tap = it / p.tap_kper, tb = p.tap_flip ? 8 - tap : tap;
              a_k = b_k = (it - tap * p.tap_kper) * BK;
              a_z = tap;
              const int dxi = tb % 3;
              mb = dxi == 0 ? &mapB : (dxi == 1 ? &mapB2 : &mapB3);
              b_n = n0 + (tb / 3 - 1) * p.tap_w;
            } else if (p.k_split > 0 && it >= p.k_split) {
              mb = &mapB2;
              b_k = (it - p.k_split) * BK;
            }
          } else {
            if (p.tap_n > 0) {
              const int tap = n0 / p.tap_n, dxi = tap % 3;
              mb = dxi == 0 ? &mapB : (dxi == 1 ? &mapB2 : &mapB3);
              b_n = n0 - tap * p.tap_n;
              b_k = it * BK + (tap / 3 - 1) * p.tap_w;
            } else if (p.n_split > 0 && n0 >= p.n_split) {
              mb = &mapB2;
              b_n = n0 - p.n_split;
            }
          }
          if constexpr (TWO) {
            // both CTAs' bytes complete on the leader's barrier; only the leader arms it
            if (crank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
            if constexpr (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / G::MN_ELEMS; j++) tma_load_3d_2cta(sA + j * BLK_BYTES, &mapA, &full[s], m0 + G::MN_ELEMS * j, a_k, a_z);
            } else {
              tma_load_3d_2cta(sA, &mapA, &full[s], a_k, m0, a_z);
            }
            if constexpr (B_MN) {
              constexpr int HB = BN / 2 / G::MN_ELEMS;        // MN blocks of this CTA's half of the B tile
#pragma unroll
              for (int j = 0; j < HB; j++) tma_load_3d_2cta(sB + j * BLK_BYTES, mb, &full[s], b_n + G::MN_ELEMS * (crank * HB + j), b_k, b_b);
            } else {
              tma_load_3d_2cta(sB, mb, &full[s], b_k, b_n + crank * (BN / 2), b_b);
            }
            continue;
          }
          mbar_expect_tx(&full[s], STAGE_BYTES);
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / G::MN_ELEMS; j++) tma_load_3d(sA + j * BLK_BYTES, &mapA, &full[s], m0 + G::MN_ELEMS * j, a_k, a_z);
          } else {
            tma_load_3d(sA, &mapA, &full[s], a_k, m0, a_z);
          }
          if constexpr (B_MN) {
            constexpr int NBLK = BN / G::MN_ELEMS;
            if constexpr (CS > 1) {
#pragma unroll
              for (int j = 0; j < NBLK / CS; j++) {
                const int jj = crank * (NBLK / CS) + j;
                tma_load_3d_mc(sB + jj * BLK_BYTES, mb, &full[s], b_n + G::MN_ELEMS * jj, b_k, b_b, cmask);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NBLK; j++) tma_load_3d(sB + j * BLK_BYTES, mb, &full[s], b_n + G::MN_ELEMS * j, b_k, b_b);
            }
          } else {
            if constexpr (CS > 1)
              tma_load_3d_mc(sB + crank * (BN / CS) * 128, mb, &full[s], b_k, b_n + crank * (BN / CS), b_b, cmask);
            else
              tma_load_3d(sB, mb, &full[s], b_k, b_n, b_b);
          }
        }
      }
    }
  } else if (warp == EW + 1) {
    if ((!TWO || crank == 0) && elect_one()) {
      const uint32_t idesc = instr_desc((EB == 2 && p.f16) ? (uint32_t)FMT_F16 : G::FMT, TWO ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t itg = 0, tl = 0;
      for (int t = cid; t < ngroups; t += ncl, tl++) {
        const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
        GTRACE(tl, 0);
        mbar_wait(&acc_empty[as], aph ^ 1u);          // the epilogue has drained this accumulator stage
        tc_fence_after();
        GTRACE(tl, 1);
        const uint32_t dcol = tmem_base + as * BN;
        const int kc = t % p.k_chunks;
        const int nit = min(p.k_iters, (kc + 1) * p.k_per) - kc * p.k_per;
        for (int it = 0; it < nit; it++, itg++) {
          const int s = itg % STAGES;
          const uint32_t ph = (itg / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (it == 0) GTRACE(tl, 2);
          const uint32_t sA = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            constexpr uint32_t kadv = G::KSTEP_ROWS * 128;
            const uint64_t ad = A_MN ? smem_desc(sA + ks * kadv, BLK_BYTES, G::MN_SBO, G::MN_LAYOUT) : smem_desc(sA + ks * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? smem_desc(sB + ks * kadv, BLK_BYTES, G::MN_SBO, G::MN_LAYOUT) : smem_desc(sB + ks * 32, 16, 1024, 2);
            if constexpr (TWO) {
              if constexpr (EB == 4) mma_tf32_2cta(dcol, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
              else mma_bf16_2cta(dcol, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
            } else {
              if constexpr (EB == 4) mma_tf32(dcol, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
              else mma_bf16(dcol, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
            }
          }
          if constexpr (TWO) mma_commit_2cta(&empty[s], cmask);    // both CTAs' producers refill their part of the slot
          else if constexpr (CS > 1) mma_commit_mc(&empty[s], cmask);   // the slot is refilled by every CTA of the cluster
          else mma_commit(&empty[s]);
        }
        if constexpr (TWO) mma_commit_2cta(&acc_full[as], cmask);  // each CTA's epilogue drains its own 128 rows
        else mma_commit(&acc_full[as]);
        GTRACE(tl, 3);
      }
    }
  } else {
    // ---------------- epilogue warps 0..3: TMEM lanes 32w.. <-> output rows m0 + 32w..
    uint8_t* slots = stage_out + warp * NSLOT * SLOT_BYTES;
    const int q4 = warp & 3;            // TMEM lane quarter = 32-row block of the tile
    const int chalf = warp >> 2;        // EW == 8: which chunk pairs of a tile are this warp's (pair index parity)
    uint32_t tl = 0, chunk = 0, echunk = 0;
    for (int t = cid; t < ngroups; t += ncl, tl++) {
      const int kc = t % p.k_chunks, tt = t / p.k_chunks;
      const int z = tt / per_z, r = tt - z * per_z;
      const int m0 = ((r / p.tiles_n) * CS + crank) * BM, n0 = (r % p.tiles_n) * BN;
      const uint32_t as = tl & 1u, aph = (tl >> 1) & 1u;
      const int row0 = m0 + q4 * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M_valid;
      const int zc = p.out_batched ? (p.idxC ? p.idxC[z] : z) : 0;
      const bool second = p.m_split > 0 && m0 >= p.m_split;
      const CUtensorMap* mo = second ? &mapO2 : &mapO;
      const int orow0 = second ? row0 - p.m_split : row0;
      // the bias / coordinate terms belong to the whole reduction: with split-K only chunk 0 adds them
      const float bias = (p.u && row_ok && kc == 0) ? p.u[(long long)z * p.ldu + row] : 0.f;
      const float rowmul = (p.epi_exp == 2 && row_ok) ? p.alpha * p.u2[(long long)z * p.ldu + row] : 0.f;
      const float alpha = p.alpha_z ? p.alpha * p.alpha_z[z] : p.alpha;
      const bool h16 = p.out_f16 != 0;       // fp16 pipeline of the co-attention backward: its own chunk code below
      float s1 = 0.f, s2 = 0.f, amax = 0.f;
      // dS epilogue: the E chunk of a thread's row (128 contiguous bytes) is fetched ONE CHUNK AHEAD -- the first one before the
      // accumulator is even complete -- so its L2 / HBM latency hides under the TMEM load, the arithmetic and the store of the
      // previous chunk (profiles/r2i: with the load inside the chunk this GEMM took 824 us against 525 us for its exp sibling)
      // dS epilogue: the 32x32 E block of a chunk.  A thread needs ITS row (128 contiguous bytes): read directly that is 32 cache lines
      // per warp instruction, and the epilogue became L1-wavefront bound (profiles/r2i, r2j: 825 us against 476 us for the exp
      // sibling, with or without prefetching).  Pair mode: TMA brings the block into a per-warp swizzled slot one chunk ahead (the
      // first one before the accumulator is complete) and each thread reads its row from shared memory; columns / rows beyond the
      // tensor arrive as zeros.  mapO2 is the E tensor in that mode (m_split is never used together with this epilogue).
      const bool e_tma = ESLOT && p.epi_exp == 2;
      uint8_t* eslots = stage_e + warp * 2 * SLOT_BYTES;
      uint64_t* ebar = e_full + warp * 2;
      float4 ecur[8];
      const float* erow = (p.epi_exp == 2 && row_ok && !e_tma && !h16) ? p.cc + (long long)z * p.cc_sb + (long long)row * p.ldcc : nullptr;
      const __half* erow_h = (p.epi_exp == 2 && row_ok && !e_tma && h16)
                                 ? reinterpret_cast<const __half*>(p.cc) + (long long)z * p.cc_sb + (p.cc_t ? (long long)row : (long long)row * p.ldcc)
                                 : nullptr;
      auto load_e = [&](int c, float4 (&dst)[8]) {
        const int nb_ = n0 + c * 32;
        if (erow && nb_ + 32 <= p.N_valid) {
#pragma unroll
          for (int e = 0; e < 8; e++) dst[e] = __ldg(reinterpret_cast<const float4*>(erow + nb_ + 4 * e));
        }
      };
      // fp32 E: one 32 x 32 block per chunk; fp16 E: one 32 x 64 block per PAIR of chunks (same 4 KiB slot, same swizzle)
      auto tma_e = [&](int c) {          // lane 0: E block of chunk c -> slot (echunk + c) & 1
        const uint32_t k = echunk + (uint32_t)(h16 ? (EW == 8 ? c >> 2 : c >> 1) : c);
        mbar_expect_tx(&ebar[k & 1u], SLOT_BYTES);
        if (p.cc_t) tma_load_3d(eslots + (k & 1u) * SLOT_BYTES, &mapO2, &ebar[k & 1u], row0, n0 + c * 32, z);   // box [64 k][32 q]
        else tma_load_3d(eslots + (k & 1u) * SLOT_BYTES, &mapO2, &ebar[k & 1u], n0 + c * 32, row0, z);
      };
      if (e_tma) { if (lane == 0) tma_e(EW == 8 ? 2 * chalf : 0); }
      else if (p.epi_exp == 2 && !h16) load_e(0, ecur);
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      if (threadIdx.x == 0) GTRACE(tl, 4);
#pragma unroll 1
      for (int c = 0; c < BN / 32; c++, chunk++) {
        float4 enext[8];
        const bool mine = EW == 4 || ((c >> 1) & 1) == chalf;
        if (e_tma) {
          constexpr int estep = EW == 8 ? 4 : 2;      // chunks to this warp's next pair
          if (h16) { if (lane == 0 && (c & 1) == 0 && mine && c + estep < BN / 32) tma_e(c + estep); }
          else if (lane == 0 && c + 1 < BN / 32) tma_e(c + 1);
        } else if (p.epi_exp == 2 && !h16 && c + 1 < BN / 32) load_e(c + 1, enext);
        if (!mine) continue;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + as * BN + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        const int nb = n0 + c * 32;
        if (h16) {
          // ---- fp16 pipeline (co-attention backward, epi_exp 1 / 2): the chunk's 32 values are produced, rounded to fp16, summed (the
          // row sum is the sum of exactly the values the next MMA reads) and packed in ONE pass -- ~5 instructions per element; the
          // epilogue warp is the bottleneck of these N x N-output contractions (one warp per SM sub-partition).  Chunks 2j and 2j+1
          // fill the two halves of one 128-byte slot row (64 fp16), one TMA store per pair; BN / 32 is even, so the parity of the
          // running chunk counter is the parity of c.
          uint32_t pk[16];
          float rs = 0.f, rs2 = 0.f;
          const bool full = nb + 32 <= p.N_valid;
          if (p.epi_exp == 1) {
            const float a2 = p.alpha * 1.4426950408889634f, c2 = (p.exp_shift - bias) * 1.4426950408889634f;
            if (full) {          // (a separate copy: the column masks of the ragged chunk cost 3 predicated instructions per element)
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const __half2 hh = __floats2half2_rn(ex2_approx(fmaf(v[2 * e], a2, c2)), ex2_approx(fmaf(v[2 * e + 1], a2, c2)));
                pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
                const float2 f = __half22float2(hh);
                rs += f.x; rs2 += f.y;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) {
                float x0 = ex2_approx(fmaf(v[2 * e], a2, c2)), x1 = ex2_approx(fmaf(v[2 * e + 1], a2, c2));
                if (nb + 2 * e >= p.N_valid) x0 = 0.f;
                if (nb + 2 * e + 1 >= p.N_valid) x1 = 0.f;
                const __half2 hh = __floats2half2_rn(x0, x1);
                pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
                const float2 f = __half22float2(hh);
                rs += f.x; rs += f.y;
              }
            }
          } else {
            // dS = tau (dP - delta) E / r : per-row delta (bias) and tau / r (rowmul); E tile: this chunk's half of the 128-byte row
            uint32_t eh[16];
            if (e_tma && p.cc_t) {
              // transposed E: the slot holds [64 k rows][32 q] (64-byte rows, no swizzle); this thread's query is column `lane`, the
              // 32 lanes of a load read 64 consecutive bytes of one row
              const uint32_t k = echunk + (uint32_t)(EW == 8 ? c >> 2 : c >> 1);
              mbar_wait(&ebar[k & 1u], (k >> 1) & 1u);
              const __half* er = reinterpret_cast<const __half*>(eslots + (k & 1u) * SLOT_BYTES) + (c & 1) * 32 * 32 + lane;
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const __half2 hh = __halves2half2(er[(2 * e) * 32], er[(2 * e + 1) * 32]);
                eh[e] = *reinterpret_cast<const uint32_t*>(&hh);
              }
            } else if (e_tma) {
              const uint32_t k = echunk + (uint32_t)(EW == 8 ? c >> 2 : c >> 1);
              mbar_wait(&ebar[k & 1u], (k >> 1) & 1u);
              const uint8_t* er = eslots + (k & 1u) * SLOT_BYTES + lane * 128;
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const uint4 q = *reinterpret_cast<const uint4*>(er + ((((c & 1) * 4 + e) ^ (lane & 7)) * 16));
                eh[4 * e] = q.x; eh[4 * e + 1] = q.y; eh[4 * e + 2] = q.z; eh[4 * e + 3] = q.w;
              }
            } else if (erow_h && p.cc_t) {
              // transposed E from global memory (single-CTA tiles): element (row, nb + e) lives at cc[(nb + e) * ldcc + row]
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const __half z0 = __float2half(0.f);
                const __half2 hh = __halves2half2(nb + 2 * e < p.N_valid ? erow_h[(long long)(nb + 2 * e) * p.ldcc] : z0,
                                                  nb + 2 * e + 1 < p.N_valid ? erow_h[(long long)(nb + 2 * e + 1) * p.ldcc] : z0);
                eh[e] = *reinterpret_cast<const uint32_t*>(&hh);
              }
            } else if (erow_h && full) {
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(erow_h + nb + 8 * e));
                eh[4 * e] = q.x; eh[4 * e + 1] = q.y; eh[4 * e + 2] = q.z; eh[4 * e + 3] = q.w;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) {
                const __half2 hh = __halves2half2((erow_h && nb + 2 * e < p.N_valid) ? erow_h[nb + 2 * e] : __float2half(0.f),
                                                  (erow_h && nb + 2 * e + 1 < p.N_valid) ? erow_h[nb + 2 * e + 1] : __float2half(0.f));
                eh[e] = *reinterpret_cast<const uint32_t*>(&hh);
              }
            }
#pragma unroll
            for (int e = 0; e < 16; e++) {
              const float2 ev = __half22float2(*reinterpret_cast<const __half2*>(&eh[e]));
              const __half2 hh = __floats2half2_rn((v[2 * e] - bias) * (rowmul * ev.x), (v[2 * e + 1] - bias) * (rowmul * ev.y));
              pk[e] = *reinterpret_cast<const uint32_t*>(&hh);
              const float2 f = __half22float2(hh);
              rs += f.x; rs2 += f.y;
            }
          }
          if (p.sum && row_ok) s1 += rs + rs2;
          uint8_t* slot16 = slots + ((chunk >> 1) % NSLOT) * SLOT_BYTES;
          if ((c & 1) == 0) {
            if (lane == 0) tma_store_wait_read_n<NSLOT - 1>();
            __syncwarp();
          }
          uint8_t* srow16 = slot16 + lane * 128;
#pragma unroll
          for (int e = 0; e < 4; e++)
            *reinterpret_cast<uint4*>(srow16 + ((((c & 1) * 4 + e) ^ (lane & 7)) * 16)) = make_uint4(pk[4 * e], pk[4 * e + 1], pk[4 * e + 2], pk[4 * e + 3]);
          if (c & 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (row0 < p.M_valid && nb - 32 < p.N_valid) tma_store_3d(mo, slot16, nb - 32, orow0, zc);
              tma_store_commit();
            }
          }
          continue;
        }
        if (p.epi_exp == 1) {
#pragma unroll
          for (int e = 0; e < 32; e++) v[e] = tf32_rn(__expf(fmaf(p.alpha, v[e], p.exp_shift - bias)));
        } else if (e_tma) {
          // dS = tau (dP - delta) E / r : per-row delta (bias) and tau / r (rowmul)
          const uint32_t k = echunk + (uint32_t)c;
          mbar_wait(&ebar[k & 1u], (k >> 1) & 1u);
          const uint8_t* er = eslots + (k & 1u) * SLOT_BYTES + lane * 128;
#pragma unroll
          for (int e = 0; e < 8; e++) {
            const float4 q = *reinterpret_cast<const float4*>(er + ((e ^ (lane & 7)) * 16));
            v[4 * e] = tf32_rn((v[4 * e] - bias) * rowmul * q.x); v[4 * e + 1] = tf32_rn((v[4 * e + 1] - bias) * rowmul * q.y);
            v[4 * e + 2] = tf32_rn((v[4 * e + 2] - bias) * rowmul * q.z); v[4 * e + 3] = tf32_rn((v[4 * e + 3] - bias) * rowmul * q.w);
          }
        } else if (p.epi_exp == 2) {
          // same from global memory (single-CTA tiles: no room for the landing slots)
          if (row_ok && nb < p.N_valid) {
            if (nb + 32 <= p.N_valid) {
#pragma unroll
              for (int e = 0; e < 8; e++) {
                const float4 q = ecur[e];
                v[4 * e] = tf32_rn((v[4 * e] - bias) * rowmul * q.x); v[4 * e + 1] = tf32_rn((v[4 * e + 1] - bias) * rowmul * q.y);
                v[4 * e + 2] = tf32_rn((v[4 * e + 2] - bias) * rowmul * q.z); v[4 * e + 3] = tf32_rn((v[4 * e + 3] - bias) * rowmul * q.w);
              }
            } else {
              const float* er = erow + nb;      // ragged last chunk of the row: element by element
#pragma unroll
              for (int e = 0; e < 32; e++) v[e] = (nb + e < p.N_valid) ? tf32_rn((v[e] - bias) * rowmul * er[e]) : 0.f;
            }
          }
#pragma unroll
          for (int e = 0; e < 8; e++) ecur[e] = enext[e];
        } else if (p.cc && row_ok && nb < p.N_valid && kc == 0) {
          const float* cr = p.cc + (long long)row * p.ldcc + nb;
          if (nb + 32 <= p.N_valid) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
              const float4 q = *reinterpret_cast<const float4*>(cr + 4 * e);
              v[4 * e] = fmaf(alpha, v[4 * e], bias + q.x); v[4 * e + 1] = fmaf(alpha, v[4 * e + 1], bias + q.y);
              v[4 * e + 2] = fmaf(alpha, v[4 * e + 2], bias + q.z); v[4 * e + 3] = fmaf(alpha, v[4 * e + 3], bias + q.w);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e++) v[e] = fmaf(alpha, v[e], bias + ((nb + e < p.N_valid) ? cr[e] : 0.f));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e++) v[e] = fmaf(alpha, v[e], bias);
          if (p.absmax2 && second) {      // rows / columns beyond the tensor come from zero-filled operands: they contribute 0
#pragma unroll
            for (int e = 0; e < 32; e++) amax = fmaxf(amax, fabsf(v[e]));
          }
        }
        if (p.sum && row_ok) {
          if (nb + 32 <= p.N_valid) {
#pragma unroll
            for (int e = 0; e < 32; e++) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e++)
              if (nb + e < p.N_valid) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
          }
        }
        uint8_t* slot = slots + (chunk % NSLOT) * SLOT_BYTES;
        if (p.direct_store) {
          // variant 5: transpose through the warp's swizzled slot, then row-major 16-byte stores (a warp instruction writes
          // 4 rows x 128 B).  Measured against the TMA-store epilogue on the 128x256x512 tiles of the co-attention backward
          // (scripts/prof_gemm.py): epilogue 14 k instead of 8.5 k cycles per tile, kernel 65 us instead of 60 us.  Either way the
          // main loop runs at ~10.3 k cycles per tile without the output stores and ~15 k with them: these GEMMs write 64 MB for
          // 17 GFLOP and are bound by that traffic, not by the tensor pipe.
          __syncwarp();                                   // previous chunk's reads of this slot are done
          uint8_t* srow = slot + lane * 128;
#pragma unroll
          for (int e = 0; e < 8; e++)
            *reinterpret_cast<float4*>(srow + ((e ^ (lane & 7)) * 16)) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          __syncwarp();
          const int cq = lane & 7, rsub = lane >> 3;
          const int col = nb + 4 * cq;
          float* obase = (second ? p.out2 + (long long)zc * p.so_b2 : p.out + (long long)zc * p.so_b);
          const long long ldo_ = second ? p.ldo2 : p.ldo;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int r = 4 * i + rsub;
            const float4 q = *reinterpret_cast<const float4*>(slot + r * 128 + ((cq ^ (r & 7)) * 16));
            if (row0 + r < p.M_valid && col < p.N_valid) {
              float* dst = obase + (long long)(orow0 + r) * ldo_ + col;
              if (col + 4 <= p.N_valid) {
                if (p.atomic) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w) : "memory");
                else *reinterpret_cast<float4*>(dst) = q;
              } else {
                const float qq[4] = {q.x, q.y, q.z, q.w};
                for (int t = 0; t < 4 && col + t < p.N_valid; t++) {
                  if (p.atomic) atomicAdd(dst + t, qq[t]);
                  else dst[t] = qq[t];
                }
              }
            }
          }
          continue;
        }
        if (lane == 0) tma_store_wait_read_n<NSLOT - 1>();   // the store that last read this slot (NSLOT chunks ago) is done with it
        __syncwarp();
        uint8_t* srow = slot + lane * 128;
#pragma unroll
        for (int e = 0; e < 8; e++)
          *reinterpret_cast<float4*>(srow + ((e ^ (lane & 7)) * 16)) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (row0 < p.M_valid && nb < p.N_valid) {       // rows / columns beyond the tensor are clipped by the map
            if (p.atomic) tma_reduce_add_3d(mo, slot, nb, orow0, zc);
            else tma_store_3d(mo, slot, nb, orow0, zc);
          }
          tma_store_commit();
        }
      }
      if (e_tma) echunk += h16 ? BN / 64 / (EW / 4) : BN / 32;
      if (p.absmax2 && second) {
        amax = warp_max(amax);
        if (lane == 0) atomicMax(p.absmax2 + z, __float_as_uint(amax));
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 0) GTRACE(tl, 5);
      if (lane == 0) {
        if constexpr (TWO) mbar_arrive_remote(&acc_empty[as], 0);   // the leader's MMA thread waits for both halves
        else mbar_arrive(&acc_empty[as]);
      }
      if (p.sum && row_ok) {
        atomicAdd(p.sum + (long long)z * p.sum_ldz + row, s1);
        if (p.sumsq) atomicAdd(p.sumsq + (long long)z * p.sum_ldz + row, s2);
      }
    }
    if (lane == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CS > 1) cluster_sync_all();          // no CTA leaves while a peer may still multicast into it / arrive on its barriers
  if (warp == EW + 1) {
    if constexpr (TWO) tmem_dealloc2(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <bool A_MN, bool B_MN, int BN, int STAGES, int EB, int CS, bool TWO = false, bool ESLOT = false, int EW = 4>
int launch_cfg2c(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const CUtensorMap& mb3, const CUtensorMap& mo,
                 const CUtensorMap& mo2, const Gemm2P& p, cudaStream_t st) {
  constexpr int smem = STAGES * (A_BYTES + (TWO ? BN / 2 : BN) * 128) + EW * (EW == 8 ? 1 : GEMM2_NSLOT) * 32 * 128 +
                       (ESLOT ? EW * 2 * 32 * 128 : 0) + 1024 + 256;
  static_assert(smem <= 232448, "umma_gemm2: shared memory budget");
  auto kern = umma_gemm2_kernel<A_MN, B_MN, BN, STAGES, EB, CS, TWO, ESLOT, EW>;
  DCNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "umma_gemm2.attr");
  const int mgroups = (p.tiles_m + CS - 1) / CS;
  const int ngroups = mgroups * (p.ntiles / p.tiles_m) * p.k_chunks;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((EW + 2) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  int ncl = sm_count() / CS;
  if (CS > 1) {
    // clusters are placed inside one GPC: ask how many fit (per template instance, cached)
    static int max_cl = 0;
    if (!max_cl) {
      cfg.gridDim = dim3(sm_count() / CS * CS);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) max_cl = n; else { cudaGetLastError(); max_cl = sm_count() / CS; }
    }
    ncl = max_cl;
  }
  if (ncl > ngroups) ncl = ngroups;
  cfg.gridDim = dim3(ncl * CS);
  DCNET_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mb2, mb3, mo, mo2, p), "umma_gemm2.launch");
  DCNET_LAUNCH_OK("umma_gemm2");
  return 0;
}

// cluster size: the M tiles of a group share B.  4 when M has >= 4 tiles (C = 512 channels, N >= 512 positions), else 2, else 1.
template <bool A_MN, bool B_MN, int BN, int STAGES, int EB>
int launch_cfg2(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const CUtensorMap& mb3, const CUtensorMap& mo,
                const CUtensorMap& mo2, const Gemm2P& p, int cs, cudaStream_t st) {
  if constexpr (BN == 256 && EB == 2 && A_MN && B_MN) {
    // fp16 exp / dS epilogues of the co-attention backward (S = Fa^T Fb, dP = dO^T Fb: both operands MN-major): 8 epilogue warps
    if (cs == -2 && p.out_f16 && g_epi_warps8) {
      if (p.epi_exp == 2) return launch_cfg2c<A_MN, B_MN, BN, GEMM2_STAGES_PAIR - 2, EB, 2, true, true, 8>(ma, mb, mb2, mb3, mo, mo2, p, st);
      return launch_cfg2c<A_MN, B_MN, BN, GEMM2_STAGES_PAIR, EB, 2, true, false, 8>(ma, mb, mb2, mb3, mo, mo2, p, st);
    }
  }
  if constexpr (BN == 256) {
    if constexpr (A_MN && B_MN) {      // the dS epilogue belongs to dP = dO^T Fb: both operands MN-major
      if (cs == -2 && p.epi_exp == 2) return launch_cfg2c<A_MN, B_MN, BN, GEMM2_STAGES_PAIR - 1, EB, 2, true, true>(ma, mb, mb2, mb3, mo, mo2, p, st);
    }
    if (cs == -2) return launch_cfg2c<A_MN, B_MN, BN, GEMM2_STAGES_PAIR, EB, 2, true>(ma, mb, mb2, mb3, mo, mo2, p, st);
  }
  if (cs == 4) return launch_cfg2c<A_MN, B_MN, BN, STAGES, EB, 4>(ma, mb, mb2, mb3, mo, mo2, p, st);
  if (cs == 2) return launch_cfg2c<A_MN, B_MN, BN, STAGES, EB, 2>(ma, mb, mb2, mb3, mo, mo2, p, st);
  return launch_cfg2c<A_MN, B_MN, BN, STAGES, EB, 1>(ma, mb, mb2, mb3, mo, mo2, p, st);
}

template <bool A_MN, bool B_MN, int BN, int STAGES, int EB>
int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, const GemmP& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * (A_BYTES + BN * 128) + 1024 + 256;
  auto kern = umma_gemm_kernel<A_MN, B_MN, BN, STAGES, EB>;
  DCNET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "umma_gemm.attr");   // per device
  kern<<<grid, 192, smem, st>>>(ma, mb, mb2, p);
  DCNET_LAUNCH_OK("umma_gemm");
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Host-side description of one operand: fp32 tensor [batch][rows][cols] (cols contiguous).
//   K-major  use: rows = the M (or N) index, cols = the reduction index
//   MN-major use: rows = the reduction index, cols = the M (or N) index
// ---------------------------------------------------------------------------------------------------------------------
static bool operand_ok(const UmmaOperand& o) {
  const int per16 = o.bf16 ? 8 : 4;     // elements per 16 bytes
  return o.ptr && (reinterpret_cast<uintptr_t>(o.ptr) % 16 == 0) && (o.ld % per16 == 0) && (o.batch_stride % per16 == 0) && o.rows > 0 && o.cols > 0;
}

static int make_operand_map(CUtensorMap* m, const UmmaOperand& o, int tile_rows_kmajor) {
  const uint64_t nb = o.batches > 0 ? (uint64_t)o.batches : 1;
  const uint64_t bs = o.batches > 0 ? (uint64_t)o.batch_stride : (uint64_t)(o.rows * o.ld);
  const uint32_t inner = o.bf16 ? 64u : 32u;        // 128 bytes
  // box: K-major {128 B of k, tile rows}; MN-major {128 B of mn, BK reduction rows}
  const int r = make_tmap(m, o.ptr, o.bf16 ? 2 : 4, (uint64_t)o.cols, (uint64_t)o.rows, nb, (uint64_t)o.ld, bs, inner,
                          o.mn_major ? inner : (uint32_t)tile_rows_kmajor, /*atom32b=*/o.mn_major && !o.bf16);
  if (r != 0) return dcnet_set_error(-3, "umma_gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", r, o.rows, o.cols, o.ld);
  return 0;
}

bool umma_gemm_usable(const UmmaOperand& A, const UmmaOperand& B, const UmmaOperand* B2, int K) {
  if (!operand_ok(A) || !operand_ok(B) || (B2 && !operand_ok(*B2))) return false;
  return K > 0;
}

// D[z] (M x N) = alpha * A[z] (M x K) B[z] (K x N) (+ second source) ; grid.z = batch
int umma_gemm(const UmmaOperand& A, const UmmaOperand& B, const UmmaOperand* B2, int M, int N, int K, int k_split_elems, int n_split,
              int batch, const UmmaEpilogue& e, cudaStream_t st) {
  DCNET_CHECK_ARG(umma_gemm_usable(A, B, B2, K), "umma_gemm: operand not TMA-compatible (16-B aligned base, row pitch multiple of 4 floats)");
  DCNET_CHECK_ARG(A.bf16 == B.bf16 && (!B2 || B2->bf16 == B.bf16), "umma_gemm: mixed operand element types");
  DCNET_CHECK_ARG(A.f16 == B.f16 && (!A.f16 || A.bf16), "umma_gemm: fp16 operands: both, and flagged as 2-byte");
  const int BK = A.bf16 ? 64 : 32;
  DCNET_CHECK_ARG(k_split_elems % BK == 0, "umma_gemm: k_split must be a multiple of %d", BK);
  DCNET_CHECK_ARG(batch >= 1 && batch <= 65535, "umma_gemm: batch %d", batch);
  // wide tiles cut the L2->SM operand traffic per FLOP (the kernel is L2-bound at 128x128, profiles/r1b_ncu_full_umma_gemm.txt);
  // use them once there are enough 128x256 tiles to fill the 148 SMs
  const long long tiles256 = (long long)ceil_div(N, 256) * ceil_div(M, BM) * batch;
  // k_chunks < 0 (auto split-K, reduce-add outputs): the wide tile is kept even when there are fewer than 148 of them -- the
  // reduction is then split so that every SM still gets work (long-K contractions onto a [C,N] map: 44 tiles per problem at N=2704)
  const bool auto_split = e.k_chunks < 0 && e.atomic;
  const int BN = (N <= 64) ? 64 : ((N >= 256 && (tiles256 >= 148 || auto_split)) ? 256 : 128);
  // persistent kernel with TMA-store epilogue whenever the output rows are TMA-addressable
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  const int oal = e.out_f16 ? 8 : 4;        // elements per 16 bytes of the output
  DCNET_CHECK_ARG(!e.out_f16 || (!e.atomic && e.m_split == 0 && (e.epi_exp == 1 || e.epi_exp == 2)), "umma_gemm: fp16 output belongs to the exp / dS epilogues (plain store, no m_split)");
  const bool out_ok = al16(e.out) && e.ldo % oal == 0 && e.so_b % oal == 0 &&
                      (e.m_split == 0 || (e.m_split % BM == 0 && e.out2 && al16(e.out2) && e.ldo2 % 4 == 0 && e.so_b2 % 4 == 0)) &&
                      (!e.cc || (al16(e.cc) && e.ldcc % oal == 0)) && !(e.so_b == 0 && batch > 1 && !e.atomic) && !g_force_v1;
  DCNET_CHECK_ARG(out_ok || (!e.out_f16 && !e.alpha_z && e.exp_shift == 0.f), "umma_gemm: fp16 output / alpha_z / exp_shift need the persistent kernel");
  DCNET_CHECK_ARG(!e.absmax2 || (out_ok && e.m_split > 0 && !e.cc && !e.epi_exp && !e.atomic), "umma_gemm: absmax2 belongs to the plain two-output epilogue of the persistent kernel");
  // cluster size of the persistent kernel: CTAs with consecutive M tiles share (multicast) the B tile
  const int tiles_m = ceil_div(M, BM);
  int cs = (!out_ok || g_cluster_max < 2 || tiles_m < 2) ? 1 : ((g_cluster_max >= 4 && tiles_m % 4 == 0) ? 4 : 2);
  if (B.mn_major) {            // an MN-major B tile is loaded in blocks of 128 B along N: every CTA of the cluster needs >= 1 block
    const int nblk = BN / (B.bf16 ? 64 : 32);
    while (cs > 1 && nblk % cs != 0) cs >>= 1;
  }
  // CTA pairs (cta_group::2) for the wide tf32 tiles: halves the B traffic per output tile
  const bool pair = out_ok && g_pair_mode && BN == 256 && tiles_m >= 2;
  if (pair) cs = 2;
  CUtensorMap ma, mb, mb2, mb3;
  DCNET_TRY(make_operand_map(&ma, A, BM));
  DCNET_TRY(make_operand_map(&mb, B, BN / cs));
  if (B2) DCNET_TRY(make_operand_map(&mb2, *B2, BN / cs)); else mb2 = mb;
  if (e.B3) DCNET_TRY(make_operand_map(&mb3, *e.B3, BN / cs)); else mb3 = mb;
  const bool taps = e.tap_kper > 0 || e.tap_n > 0;
  DCNET_CHECK_ARG(!taps || (out_ok && B2 && e.B3 && !A.bf16 && (e.tap_kper == 0) != (e.tap_n == 0)), "umma_gemm: tap mode needs three B sources and the persistent kernel");
  DCNET_CHECK_ARG(e.tap_kper == 0 || (B.mn_major && K == 9 * e.tap_kper * BK), "umma_gemm: tap_kper: B must be MN-major and K = 9 taps");
  DCNET_CHECK_ARG(e.tap_n == 0 || (!B.mn_major && e.tap_n % BN == 0 && N == 9 * e.tap_n), "umma_gemm: tap_n: B must be K-major, N = 9 taps of whole tiles");
  if (pair) cs = -2;
  GemmP p{};
  p.M_valid = M; p.N_valid = N;
  p.k_iters = (K + BK - 1) / BK;
  p.k_split = (B.mn_major && B2 && !taps) ? k_split_elems / BK : 0;
  p.n_split = (!B.mn_major && B2 && !taps) ? n_split : 0;
  p.m_split = e.m_split;
  p.a_batched = A.batches > 0; p.b_batched = B.batches > 0;
  p.idxA = e.idxA; p.idxB = e.idxB; p.idxC = e.idxC;
  p.out = e.out; p.ldo = e.ldo; p.so_b = e.so_b; p.out2 = e.out2; p.ldo2 = e.ldo2; p.so_b2 = e.so_b2;
  p.alpha = e.alpha; p.atomic = e.atomic; p.u = e.u; p.ldu = e.ldu; p.cc = e.cc; p.ldcc = e.ldcc; p.sum = e.sum; p.sumsq = e.sumsq;
  const int am = A.mn_major ? 1 : 0, bm = B.mn_major ? 1 : 0;
  DCNET_CHECK_ARG(e.epi_exp != 1 || (out_ok && e.u && !e.cc && !e.atomic), "umma_gemm: the exp epilogue needs the persistent kernel and a row term");
  DCNET_CHECK_ARG(e.epi_exp != 2 || (out_ok && e.u && e.u2 && e.cc && !e.atomic), "umma_gemm: the dS epilogue needs the persistent kernel, two row terms and the E tile");
  DCNET_CHECK_ARG(e.k_chunks <= 1 || (out_ok && e.atomic && !e.sum && !e.epi_exp), "umma_gemm: split-K needs a reduce-add output on the persistent kernel");
  DCNET_CHECK_ARG(!(B.mn_major && B2) || e.k_chunks == 1 || e.k_chunks == 0, "umma_gemm: split-K with a second K source is not supported");
  DCNET_CHECK_ARG(out_ok || (e.sum_ldz == 0 && e.cc_sb == 0), "umma_gemm: batched sums / cc need the persistent kernel");
  if (out_ok) {
    Gemm2P q{};
    q.M_valid = M; q.N_valid = N; q.k_iters = p.k_iters; q.k_split = p.k_split; q.n_split = p.n_split; q.m_split = p.m_split;
    q.a_batched = p.a_batched; q.b_batched = p.b_batched; q.out_batched = (e.so_b != 0) ? 1 : 0;
    q.idxA = e.idxA; q.idxB = e.idxB; q.idxC = e.idxC;
    q.tiles_m = ceil_div(M, BM); q.tiles_n = ceil_div(N, BN); q.ntiles = q.tiles_m * q.tiles_n * batch;
    q.alpha = e.alpha; q.atomic = e.atomic; q.u = e.u; q.ldu = e.ldu; q.cc = e.cc; q.ldcc = e.ldcc; q.sum = e.sum; q.sumsq = e.sumsq;
    q.epi_exp = e.epi_exp; q.u2 = e.u2; q.cc_sb = e.cc_sb; q.sum_ldz = e.sum_ldz;
    q.tap_kper = e.tap_kper; q.tap_w = e.tap_w; q.tap_flip = e.tap_flip; q.tap_n = e.tap_n;
    q.f16 = A.f16 ? 1 : 0; q.out_f16 = e.out_f16; q.exp_shift = e.exp_shift; q.alpha_z = e.alpha_z;
    q.absmax2 = e.absmax2; q.cc_t = e.cc_t;
    int want_chunks = e.k_chunks;
    if (auto_split) {
      const long long tiles = (long long)q.ntiles;
      want_chunks = tiles >= 2 * sm_count() ? 1 : (int)((2 * sm_count() + tiles - 1) / tiles);
      if (want_chunks > q.k_iters / 8) want_chunks = q.k_iters / 8;       // keep >= 8 k-steps (256 tf32 / 512 bf16 reduction elements) per item
    }
    q.k_chunks = want_chunks > 1 ? (want_chunks < q.k_iters ? want_chunks : q.k_iters) : 1;
    q.k_per = (q.k_iters + q.k_chunks - 1) / q.k_chunks;
    q.k_chunks = (q.k_iters + q.k_per - 1) / q.k_per;        // no empty chunk
    q.trace = g_gemm_trace;
    q.direct_store = g_gemm_tma_store ? 0 : 1;
    q.out = e.out; q.ldo = e.ldo; q.so_b = e.so_b; q.out2 = e.out2; q.ldo2 = e.ldo2; q.so_b2 = e.so_b2;
    const uint64_t nbo = q.out_batched ? 65535u : 1u;
    const uint64_t rows1 = e.m_split > 0 ? (uint64_t)e.m_split : (uint64_t)M;
    CUtensorMap mo, mo2;
    // fp16 output: 32 rows x 64 fp16 per store (the same 128-byte slot rows)
    const int oeb = e.out_f16 ? 2 : 4;
    const uint32_t obox = e.out_f16 ? 64u : 32u;
    int r = make_tmap(&mo, e.out, oeb, (uint64_t)N, rows1, nbo, (uint64_t)e.ldo, q.out_batched ? (uint64_t)e.so_b : rows1 * (uint64_t)e.ldo, obox, 32);
    if (r != 0) return dcnet_set_error(-3, "umma_gemm: cuTensorMapEncodeTiled(out) failed (%d)", r);
    mo2 = mo;
    if (e.epi_exp == 2) {
      DCNET_CHECK_ARG(e.m_split == 0 && al16(e.cc) && e.ldcc % oal == 0 && e.cc_sb % oal == 0, "umma_gemm: dS epilogue: E must be TMA-addressable, no m_split");
      if (e.cc_t) {
        DCNET_CHECK_ARG(e.out_f16, "umma_gemm: the transposed E tensor belongs to the fp16 dS epilogue");
        // E^T [z][N cols of the output (k)][M rows of the output (q)]: tile [64 k][32 q], 64-byte rows, no swizzle
        r = make_tmap(&mo2, e.cc, 2, (uint64_t)M, (uint64_t)N, e.cc_sb ? 65535u : 1u, (uint64_t)e.ldcc, e.cc_sb ? (uint64_t)e.cc_sb : (uint64_t)N * e.ldcc,
                      32, 64, false, /*no_swizzle=*/true);
      } else
      r = make_tmap(&mo2, e.cc, oeb, (uint64_t)N, (uint64_t)M, e.cc_sb ? 65535u : 1u, (uint64_t)e.ldcc, e.cc_sb ? (uint64_t)e.cc_sb : (uint64_t)M * e.ldcc, obox, 32);
      if (r != 0) return dcnet_set_error(-3, "umma_gemm: cuTensorMapEncodeTiled(E) failed (%d)", r);
    }
    if (e.m_split > 0) {
      const uint64_t rows2 = (uint64_t)(M - e.m_split);
      r = make_tmap(&mo2, e.out2, 4, (uint64_t)N, rows2, nbo, (uint64_t)e.ldo2, q.out_batched ? (uint64_t)e.so_b2 : rows2 * (uint64_t)e.ldo2, 32, 32);
      if (r != 0) return dcnet_set_error(-3, "umma_gemm: cuTensorMapEncodeTiled(out2) failed (%d)", r);
    }
    #define DISPATCH2(AM, BMJ)                                                                           \
    if (am == AM && bm == BMJ) {                                                                     \
      if (A.bf16) {                                                                                  \
        if (BN == 64) return launch_cfg2<AM, BMJ, 64, 4, 2>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);     \
        if (BN == 256) return launch_cfg2<AM, BMJ, 256, GEMM2_STAGES256, 2>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);   \
        return launch_cfg2<AM, BMJ, 128, 4, 2>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);                  \
      }                                                                                              \
      if (BN == 64) return launch_cfg2<AM, BMJ, 64, 4, 4>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);       \
      if (BN == 256) return launch_cfg2<AM, BMJ, 256, GEMM2_STAGES256, 4>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);     \
      return launch_cfg2<AM, BMJ, 128, 4, 4>(ma, mb, mb2, mb3, mo, mo2, q, cs, st);                    \
    }
    DISPATCH2(0, 0) DISPATCH2(0, 1) DISPATCH2(1, 0) DISPATCH2(1, 1)
#undef DISPATCH2
  }
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), batch);
#define DISPATCH(AM, BMJ)                                                                   \
  if (am == AM && bm == BMJ) {                                                              \
    if (A.bf16) {                                                                           \
      if (BN == 64) return launch_cfg<AM, BMJ, 64, 4, 2>(ma, mb, mb2, p, grid, st);         \
      if (BN == 256) return launch_cfg<AM, BMJ, 256, 4, 2>(ma, mb, mb2, p, grid, st);       \
      return launch_cfg<AM, BMJ, 128, 3, 2>(ma, mb, mb2, p, grid, st);                      \
    }                                                                                       \
    if (BN == 64) return launch_cfg<AM, BMJ, 64, 4, 4>(ma, mb, mb2, p, grid, st);           \
    if (BN == 256) return launch_cfg<AM, BMJ, 256, 4, 4>(ma, mb, mb2, p, grid, st);         \
    return launch_cfg<AM, BMJ, 128, 3, 4>(ma, mb, mb2, p, grid, st);                        \
  }
  DISPATCH(0, 0) DISPATCH(0, 1) DISPATCH(1, 0) DISPATCH(1, 1)
#undef DISPATCH
  return dcnet_set_error(-1, "umma_gemm: unreachable");
}


// C ABI: direct access to the tensor-core GEMM (tests, and callers with their own contractions)
extern "C" int dcnet_gemm_tf32(const float* A, int a_mn_major, long long lda, long long strideA,
                               const float* B, int b_mn_major, long long ldb, long long strideB,
                               float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                               void* stream) {
  DCNET_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm_tf32: bad arguments");
  // K-major operand: [rows = M|N][cols = K]; MN-major operand: [rows = K][cols = M|N]
  UmmaOperand a{A, a_mn_major ? K : M, a_mn_major ? M : K, lda, strideA, batch, a_mn_major != 0, false};
  UmmaOperand b{B, b_mn_major ? K : N, b_mn_major ? N : K, ldb, strideB, batch, b_mn_major != 0, false};
  UmmaEpilogue e{};
  e.out = C; e.ldo = ldc; e.so_b = strideC; e.alpha = alpha; e.atomic = atomic;
  if (atomic) e.k_chunks = -1;     // reduce-add output: split the reduction when the tiles do not fill the SMs
  return umma_gemm(a, b, nullptr, M, N, K, 0, 0, batch, e, as_stream(stream));
}


// same contraction with bf16 operands (kind::f16, fp32 accumulation); A/B point to __nv_bfloat16 data, pitches in elements
extern "C" int dcnet_gemm_bf16(const void* A, int a_mn_major, long long lda, long long strideA,
                               const void* B, int b_mn_major, long long ldb, long long strideB,
                               float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                               void* stream) {
  DCNET_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm_bf16: bad arguments");
  UmmaOperand a{(const float*)A, a_mn_major ? K : M, a_mn_major ? M : K, lda, strideA, batch, a_mn_major != 0, true};
  UmmaOperand b{(const float*)B, b_mn_major ? K : N, b_mn_major ? N : K, ldb, strideB, batch, b_mn_major != 0, true};
  UmmaEpilogue e{};
  e.out = C; e.ldo = ldc; e.so_b = strideC; e.alpha = alpha; e.atomic = atomic;
  return umma_gemm(a, b, nullptr, M, N, K, 0, 0, batch, e, as_stream(stream));
}


// same contraction with fp16 operands (kind::f16, fp32 accumulation): __half data, pitches in elements.  fp16 keeps the 11 significant
// bits of tf32, so for operands inside its normal range this is the tf32 contraction at twice the MMA rate (co-attention backward)
extern "C" int dcnet_gemm_f16(const void* A, int a_mn_major, long long lda, long long strideA,
                              const void* B, int b_mn_major, long long ldb, long long strideB,
                              float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                              void* stream) {
  DCNET_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm_f16: bad arguments");
  UmmaOperand a{(const float*)A, a_mn_major ? K : M, a_mn_major ? M : K, lda, strideA, batch, a_mn_major != 0, true, true};
  UmmaOperand b{(const float*)B, b_mn_major ? K : N, b_mn_major ? N : K, ldb, strideB, batch, b_mn_major != 0, true, true};
  UmmaEpilogue e{};
  e.out = C; e.ldo = ldc; e.so_b = strideC; e.alpha = alpha; e.atomic = atomic;
  if (atomic) e.k_chunks = -1;
  return umma_gemm(a, b, nullptr, M, N, K, 0, 0, batch, e, as_stream(stream));
}
