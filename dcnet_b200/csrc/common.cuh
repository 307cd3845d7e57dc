// Shared device/host helpers for libdcnet_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcnet_b200.h"
#include "host_error.h"

#define DCNET_CHECK_ARG(cond, ...)                         \
  do {                                                     \
    if (!(cond)) return dcnet_set_error(-1, __VA_ARGS__);  \
  } while (0)

// after a kernel launch: count it and surface launch errors (no synchronisation)
#define DCNET_LAUNCH_OK(name)                                                              \
  do {                                                                                     \
    dcnet_count_launch(1);                                                                 \
    cudaError_t e__ = cudaPeekAtLastError();                                               \
    if (e__ != cudaSuccess) {                                                              \
      cudaGetLastError();                                                                  \
      return dcnet_set_error((int)e__, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    }                                                                                      \
  } while (0)

#define DCNET_CUDA(call, name)                                                             \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) return dcnet_set_error((int)e__, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define DCNET_TRY(call)          \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `sh` must hold >= 32 floats; result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// round to nearest tf32 (10 mantissa bits).  tcgen05.mma.kind::tf32 reads fp32 operands by TRUNCATION (a relative bias of about
// -2^-12 per operand that adds up along a chain of contractions); a value rounded here reaches the MMA unchanged, so the rounding error
// is unbiased and half as large.  Producers that hand a tensor to a tf32 contraction round it when DCNET_RN_TF32 is set.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// internal launcher of the generic fp32 GEMM (sgemm.cu)
int sgemm_launch(const float* A, const float* B, float* C, int M, int N, int K, int batch, int kbatch,
                 long long sAm, long long sAk, long long sAb, long long sAkb,
                 long long sBk, long long sBn, long long sBb, long long sBkb,
                 long long sCm, long long sCn, long long sCb,
                 const int* idxA, const int* idxB, const int* idxC,
                 float alpha, float beta, const float* colscale, long long sColB, int atomic, cudaStream_t st);

// tcgen05 TF32 GEMM (umma_gemm.cu)
struct UmmaOperand {
  const float* ptr; long long rows, cols, ld, batch_stride; int batches;   // batches = 0: not batched; ld/stride in elements
  bool mn_major;
  bool bf16 = false;   // ptr is 2-byte data (kind::f16 MMA) instead of fp32 read as TF32: __nv_bfloat16 ...
  bool f16 = false;    // ... or, with bf16 = true as well, __half (11 significant bits like tf32, at twice the tf32 MMA rate)
};
struct UmmaEpilogue {
  float* out; long long ldo, so_b; float* out2; long long ldo2, so_b2; int m_split;
  float alpha; int atomic; const float* u; int ldu; const float* cc; long long ldcc; float* sum; float* sumsq;
  const int* idxA; const int* idxB; const int* idxC;
  // persistent kernel only (co-attention backward, coattn.cu):
  //   epi_exp = 1: v = exp(alpha*acc - u[z*ldu + row])                                  E = exp(tau S - shift)
  //   epi_exp = 2: v = alpha * (acc - u[z*ldu + row]) * u2[z*ldu + row] * cc[z*cc_sb + row*ldcc + col]      dS = tau (dP - delta) E / r
  //   both round v to the nearest tf32 (what the next MMA reads), and `sum` then receives row sums of exactly those values
  int epi_exp = 0;
  const float* u2 = nullptr;     // second per-row term (epi_exp = 2)
  long long cc_sb = 0;           // batch stride of cc (0: cc is shared by the batch, the conv use)
  long long sum_ldz = 0;         // sum / sumsq are indexed [z*sum_ldz + row] (0: one vector for the whole batch = BatchNorm statistics)
  int k_chunks = 1;              // reduce-add outputs (atomic = 1) only: split the reduction over this many work items per output tile
  // fp16 pipeline of the co-attention backward (coattn.cu): out (and the E tile `cc` of the dS epilogue) are __half tensors -- ldo /
  // ldcc / so_b / cc_sb in elements, multiples of 8 --, epi_exp values are rounded to fp16 instead of tf32; exp_shift is added to the
  // exponent of epi_exp = 1; alpha_z[z] (optional) multiplies alpha per batch item (undoes a per-problem operand scale)
  int out_f16 = 0;
  int cc_t = 0;                  // dS epilogue, fp16: the E tensor is stored transposed, cc[z][col][row] (what the fused forward keeps)
  float exp_shift = 0.f;
  const float* alpha_z = nullptr;
  // plain epilogue with m_split: absmax2[z] = max |out2[z]| as the bit pattern of a non-negative float (caller zeroes) -- the
  // per-problem scale of the co-attention backward's fp16 pipeline comes out of the GEMM that produces its dO
  unsigned int* absmax2 = nullptr;
  // 3x3 convolution as an implicit GEMM (conv3x3.cu; see Gemm2P in umma_gemm.cu): B, B2, B3 = the map shifted by dx = -1, 0, +1
  const UmmaOperand* B3 = nullptr;
  int tap_kper = 0, tap_w = 0, tap_flip = 0, tap_n = 0;
};
bool umma_gemm_usable(const UmmaOperand& A, const UmmaOperand& B, const UmmaOperand* B2, int K);
int umma_gemm(const UmmaOperand& A, const UmmaOperand& B, const UmmaOperand* B2, int M, int N, int K, int k_split_elems, int n_split,
              int batch, const UmmaEpilogue& e, cudaStream_t st);
