// Generic strided, batched fp32 GEMM on the CUDA cores.  It is the exact-fp32 path of the library: the
// scale-0 similarity that feeds the top-30 selection (indices must match the reference, SURVEY.md section 7
// "Index parity"), the tiny K=8 / GEMV-shaped products of the fusion, and the weight-gradient reductions.
// The large dense contractions run on tcgen05 (umma_*.cu).
#include "common.cuh"

namespace {

struct GemmArgs {
  const float* A; const float* B; float* C;
  int M, N, K, kbatch;
  long long sAm, sAk, sAb, sAkb, sBk, sBn, sBb, sBkb, sCm, sCn, sCb;
  const int* idxA; const int* idxB; const int* idxC;
  float alpha, beta;
  const float* colscale; long long sColB;
  int atomic;
};

// BM x BN tile, 256 threads, TM x TN register micro-tile per thread.  The exact-fp32 problems of the hot path are small and
// deep (K up to 1024 on a handful of CTAs), i.e. bound by the latency of one global->smem tile per iteration: deep BK tiles
// (32 for the 64x64 tile, 16 for 128x128) keep 4x / 2x the bytes in flight per iteration and cut the barrier count alike.
template <int BM, int BN, int TM, int TN, int BK>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "tile/thread mismatch");
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];

  const int b = blockIdx.z;
  const long long offA = (g.idxA ? (long long)g.idxA[b] : (long long)b) * g.sAb;
  const long long offB = (g.idxB ? (long long)g.idxB[b] : (long long)b) * g.sBb;
  const long long offC = (g.idxC ? (long long)g.idxC[b] : (long long)b) * g.sCb;
  const float* __restrict__ A = g.A + offA;
  const float* __restrict__ B = g.B + offB;
  float* __restrict__ C = g.C + offC;

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  // global->smem mapping: consecutive threads run along the contiguous axis of each operand
  const bool a_kfast = (g.sAk == 1);   // K contiguous
  const bool b_nfast = (g.sBn == 1);   // N contiguous
  constexpr int A_PER = BM * BK / NT;  // elements per thread
  constexpr int B_PER = BN * BK / NT;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  float ra[A_PER], rb[B_PER];
  const int ktiles = (g.K + BK - 1) / BK;
  const int total = ktiles * g.kbatch;

  auto load_tile = [&](int t) {
    const int kb = t / ktiles, k0 = (t % ktiles) * BK;
    const float* Ab = A + (long long)kb * g.sAkb;
    const float* Bb = B + (long long)kb * g.sBkb;
#pragma unroll
    for (int e = 0; e < A_PER; e++) {
      const int l = tid + e * NT;
      int m, k;
      if (a_kfast) { k = l % BK; m = l / BK; } else { m = l % BM; k = l / BM; }
      const int gm = m0 + m, gk = k0 + k;
      ra[e] = (gm < g.M && gk < g.K) ? __ldg(Ab + (long long)gm * g.sAm + (long long)gk * g.sAk) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < B_PER; e++) {
      const int l = tid + e * NT;
      int n, k;
      if (b_nfast) { n = l % BN; k = l / BN; } else { k = l % BK; n = l / BK; }
      const int gn = n0 + n, gk = k0 + k;
      rb[e] = (gn < g.N && gk < g.K) ? __ldg(Bb + (long long)gk * g.sBk + (long long)gn * g.sBn) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int e = 0; e < A_PER; e++) {
      const int l = tid + e * NT;
      int m, k;
      if (a_kfast) { k = l % BK; m = l / BK; } else { m = l % BM; k = l / BM; }
      As[buf][k][m] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < B_PER; e++) {
      const int l = tid + e * NT;
      int n, k;
      if (b_nfast) { n = l % BN; k = l / BN; } else { k = l % BK; n = l / BK; }
      Bs[buf][k][n] = rb[e];
    }
  };

  if (total > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int t = 0; t < total; t++) {
    const int buf = t & 1;
    if (t + 1 < total) load_tile(t + 1);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; j++) bb[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (t + 1 < total) store_tile(buf ^ 1);
    __syncthreads();
  }

  const float* cs = g.colscale ? g.colscale + (long long)b * g.sColB : nullptr;
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (cs) v *= cs[gn];
      float* p = C + (long long)gm * g.sCm + (long long)gn * g.sCn;
      if (g.atomic) atomicAdd(p, v);
      else *p = (g.beta != 0.f) ? fmaf(g.beta, *p, v) : v;
    }
  }
}

}  // namespace

int sgemm_launch(const float* A, const float* B, float* C, int M, int N, int K, int batch, int kbatch,
                 long long sAm, long long sAk, long long sAb, long long sAkb,
                 long long sBk, long long sBn, long long sBb, long long sBkb,
                 long long sCm, long long sCn, long long sCb,
                 const int* idxA, const int* idxB, const int* idxC,
                 float alpha, float beta, const float* colscale, long long sColB, int atomic, cudaStream_t st) {
  DCNET_CHECK_ARG(A && B && C, "sgemm: null operand");
  DCNET_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && batch >= 0 && kbatch >= 1,
                  "sgemm: bad sizes M=%d N=%d K=%d batch=%d kbatch=%d", M, N, K, batch, kbatch);
  if (M == 0 || N == 0 || batch == 0) return 0;
  DCNET_CHECK_ARG(batch <= 65535, "sgemm: batch %d > 65535", batch);
  GemmArgs g{A, B, C, M, N, K, kbatch, sAm, sAk, sAb, sAkb, sBk, sBn, sBb, sBkb, sCm, sCn, sCb,
             idxA, idxB, idxC, alpha, beta, colscale, sColB, atomic};
  const long long tiles128 = (long long)ceil_div(M, 128) * ceil_div(N, 128) * batch;
  if (M >= 96 && N >= 96 && tiles128 >= 96) {
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128), batch);
    sgemm_kernel<128, 128, 8, 8, 16><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64), batch);
    sgemm_kernel<64, 64, 4, 4, 32><<<grid, 256, 0, st>>>(g);
  }
  DCNET_LAUNCH_OK("sgemm");
  return 0;
}

extern "C" int dcnet_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int batch, int kbatch,
                           long long sAm, long long sAk, long long sAb, long long sAkb,
                           long long sBk, long long sBn, long long sBb, long long sBkb,
                           long long sCm, long long sCn, long long sCb,
                           const int* idxA, const int* idxB, const int* idxC,
                           float alpha, float beta, const float* colscale, long long sColscaleB,
                           int atomic, void* stream) {
  return sgemm_launch(A, B, C, M, N, K, batch, kbatch, sAm, sAk, sAb, sAkb, sBk, sBn, sBb, sBkb, sCm, sCn, sCb,
                      idxA, idxB, idxC, alpha, beta, colscale, sColscaleB, atomic, as_stream(stream));
}
