// a5 / a20: co-attention (model/DCNet_model.py:449-459; model/test_DCNet_model.py:247-274), one direction per problem.
//   S = Fa^T Fb,  P = softmax_j(tau S),  O = Fb P^T.
// This file holds the fp32 composition used for exactness checks and as the backward of round 1: the score matrix
// lives in caller-provided scratch.  The tcgen05 path (umma_coattn.cu) keeps S/P in TMEM/SMEM and is selected by
// dcnet_coattn_fwd when the shape is supported.
#include "common.cuh"

int umma_coattn_fwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob, float* out, int n_out, float* lse,
                    int C, int N, float tau, void* ws, size_t ws_bytes, cudaStream_t st);  // umma_coattn.cu
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

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int pitch4(int N) { return (N + 3) & ~3; }

}  // namespace

extern "C" size_t dcnet_coattn_workspace_bytes(int F, int nprob, int C, int N, int precision) {
  if (nprob <= 0 || N <= 0 || F <= 0) return 256;
  size_t unfused = 2 * align256((size_t)nprob * N * N * sizeof(float)) + 256;   // S / P and dP scratch (backward; unfused forward)
  if (N % 4 != 0 && precision >= 1)    // odd pitch: P / dP with the pitch padded to 4, plus padded copies of the maps and of dout
    unfused = 2 * align256((size_t)nprob * N * pitch4(N) * sizeof(float)) + align256((size_t)F * C * pitch4(N) * sizeof(float)) +
              align256((size_t)nprob * C * pitch4(N) * sizeof(float)) + 256;
  const size_t fused = umma_coattn_workspace_bytes(F, C, N);                           // bf16 staging of the maps + column norms
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
  if (precision == 2 && umma_coattn_supported(C, N))
    return umma_coattn_fwd(frames, F, qa, kb, oidx, nprob, out, n_out, lse, C, N, tau, workspace, workspace_bytes, st);
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
  (void)out;
  DCNET_CHECK_ARG(frames && qa && kb && oidx && lse && dout && dframes && nprob >= 0 && C > 0 && N > 0 && F > 0 && n_out > 0, "coattn_bwd: bad arguments");
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
  float* P = (float*)workspace;
  float* dP = (float*)((char*)workspace + align256((size_t)nprob * NN * sizeof(float)));
  const long long rows = (long long)nprob * N;
  const bool tc = coattn_tc_ok(precision, C, N, frames, dout, dframes) && coattn_tc_ok(precision, C, N, P, dP, P);
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
    if (precision == 2 && staged && reinterpret_cast<uintptr_t>(staged) % 256 == 0 && umma_coattn_supported(C, N)) {
      // P = exp(tau S - lse) in one kernel: S from the forward's bf16 staging of the maps -- the same operand bits the fused
      // forward took its lse from, so the rows of P sum to one without a renormalisation pass -- and exp in the GEMM epilogue
      const int ld = (N + 7) & ~7;
      const float* s16 = reinterpret_cast<const float*>(staged);
      UmmaOperand Smn{s16, C, N, ld, (long long)C * ld, F, true, true};
      e.epi_exp = 1; e.u = lse; e.ldu = N;
      DCNET_TRY(umma_gemm(Smn, Smn, nullptr, N, N, C, 0, 0, nprob, e, st));
    } else {
      // P = exp(tau S - lse) (precision 1) / softmax(tau S) (precision 2 without the forward's staging)
      DCNET_TRY(umma_gemm(Fmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
      exp_launch();
      DCNET_LAUNCH_OK("coattn_bwd.exp");
    }
    // dP[i,j] = sum_c dO[c,i] Fb[c,j]
    e = UmmaEpilogue{}; e.out = dP; e.ldo = N; e.so_b = NN; e.alpha = 1.f; e.idxA = oidx; e.idxB = kb;
    DCNET_TRY(umma_gemm(Gmn, Fmn, nullptr, N, N, C, 0, 0, nprob, e, st));
    // dFb[c,j] += sum_i dO[c,i] P[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = oidx; e.idxC = kb;
    DCNET_TRY(umma_gemm(Gk, Pmn, nullptr, C, N, N, 0, 0, nprob, e, st));
    softmax_bwd_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, st>>>(P, dP, rows, N, N, tau);
    DCNET_LAUNCH_OK("coattn_bwd.softmax");
    // dFa[c,i] += sum_j Fb[c,j] dS[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = kb; e.idxC = qa;
    DCNET_TRY(umma_gemm(Fk, dSk, nullptr, C, N, N, 0, 0, nprob, e, st));
    // dFb[c,j] += sum_i Fa[c,i] dS[i,j]
    e = UmmaEpilogue{}; e.out = dframes; e.ldo = N; e.so_b = CN; e.alpha = 1.f; e.atomic = 1; e.idxA = qa; e.idxC = kb;
    return umma_gemm(Fk, dSmn, nullptr, C, N, N, 0, 0, nprob, e, st);
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
  return 0;
}
