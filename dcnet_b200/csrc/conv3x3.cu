// the 3x3 convolution of the grounding head (fcn_emb[s][1] = ConvBatchNormReLU(512, 512, 3, 1, 1),
// model/DCNet_model.py:316-337, applied at :505-506) as an implicit GEMM on the tcgen05 kernel of umma_gemm.cu.
//
//   z[b, co, (y,x)] = sum_{dy,dx in -1..1} sum_ci W[co, ci, dy+1, dx+1] * x[b, ci, y+dy, x+dx]          (zero padding)
//
// The maps keep the [B, C, N = h*w] layout of the rest of the path.  A shift by dy is a shift of the flat position by dy*w, and the
// rows that fall off the top / bottom of the image are positions outside [0, N): TMA zero-fills them (the position axis is a
// dimension of its own in the tensor map, so nothing leaks in from the neighbouring channel).  A shift by dx would wrap around the
// row ends, so the three column shifts are materialised once (conv3x3_shift_kernel: x read once, x(dx=-1) and x(dx=+1) written, the
// border column zeroed) and every (dy, dx) tap of the K loop reads copy dx at position offset dy*w:
//   forward        z  = sum_t  Wq[t]   . shift_dy(x_dx)              K = 9*Cin    A = Wq [9][Cout][Cin] (tap = "batch" of the A map)
//   data gradient  dx = sum_t  Wq[t]^T . shift_-dy(dz_-dx)           K = 9*Cout   same Wq read MN-major, B side uses tap 8 - t
//   weight grad.   dWp[co][t][ci] = sum_{b,p} dz[b,co,p] x_dx[b,ci,p+dy*w]        N = 9*Cin, tap = output column block, K = positions
// The start of a TMA box must be 16-byte aligned, so the row shift needs w % 4 == 0 (8, 16, 32 at 256x256 -- the only size the reference
// itself runs, model/DCNet_model.py:584 --, 52 at 416x416).  Other widths (13, 26 at 416x416) run at a padded width wp = 16, 28: the
// shifted copies are written at pitch wp with zero pad columns (dcnet_conv3x3_shift_padded), the contractions see an h x wp image whose
// pad columns act as the border, and the pad columns of the result are dropped (dcnet_conv3x3_unpad).
// Wq is the weight permuted to [tap][Cout][Cin] (and rounded to tf32) once per step; dWp is permuted back to [Cout][Cin][3][3].
// BatchNorm statistics come out of the forward's epilogue like in the 1x1 layers.
#include "common.cuh"

namespace {

// rows = (b, c, y); one thread per element of an OUTPUT row of pitch wp >= w (the pad columns w .. wp-1 are written as zeros)
__global__ void __launch_bounds__(256) conv3x3_shift_kernel(const float* __restrict__ x, float* __restrict__ xm, float* __restrict__ xp,
                                                            float* __restrict__ x0, long long total, int w, int wp, int rn) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // index into the padded outputs
  if (i >= total) return;
  const long long row = i / wp;
  const int col = (int)(i - row * wp);
  float v = 0.f, l = 0.f, r = 0.f;
  if (col < w) {
    const float* xr = x + row * w;
    v = xr[col];
    l = col > 0 ? xr[col - 1] : 0.f;          // x(dx=-1)[p] = x[p-1]
    r = col < w - 1 ? xr[col + 1] : 0.f;      // x(dx=+1)[p] = x[p+1]
    if (rn) { v = tf32_rn(v); l = tf32_rn(l); r = tf32_rn(r); }
  }
  xm[i] = l;
  xp[i] = r;
  if (x0) x0[i] = v;
}

// [rows][wp] -> [rows][w]
__global__ void __launch_bounds__(256) conv3x3_unpad_kernel(const float* __restrict__ zp, float* __restrict__ z, long long total, int w, int wp) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;       // index into the unpadded output
  if (i >= total) return;
  const long long row = i / w;
  z[i] = zp[row * wp + (i - row * w)];
}

// W [Cout][Cin][9] -> Wq [9][Cout][Cin]
__global__ void __launch_bounds__(256) conv3x3_pack_weight_kernel(const float* __restrict__ W, float* __restrict__ Wq, int Cout, int Cin, int rn) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;      // index into Wq
  const long long per = (long long)Cout * Cin;
  if (i >= 9 * per) return;
  const int t = (int)(i / per);
  const long long r = i - t * per;        // co * Cin + ci
  const float v = W[r * 9 + t];
  Wq[i] = rn ? tf32_rn(v) : v;
}

// dWp [Cout][9][Cin] -> dW [Cout][Cin][9]
__global__ void __launch_bounds__(256) conv3x3_unpack_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int Cout, int Cin) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;      // index into dW
  if (i >= (long long)Cout * Cin * 9) return;
  const int t = (int)(i % 9);
  const long long r = i / 9;
  const int ci = (int)(r % Cin);
  const long long co = r / Cin;
  dW[i] = dWp[(co * 9 + t) * Cin + ci];
}

bool shape_ok(int Cin, int Cout, int h, int w, const void* a, const void* b, const void* c, const void* d) {
  auto al = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  // w % 4: the dy shift moves a TMA box by w positions, and the start of a box must stay 16-byte aligned (measured: w = 26 or 10 never
  // completes the load's transaction bytes)
  return Cin % 256 == 0 && Cout % 128 == 0 && w % 4 == 0 && h >= 2 && al(a) && al(b) && al(c) && al(d);
}

}  // namespace

extern "C" int dcnet_conv3x3_supported(int Cin, int Cout, int h, int w) {
  return (Cin % 256 == 0 && Cout % 256 == 0 && w % 4 == 0 && h >= 2) ? 1 : 0;
}

extern "C" int dcnet_conv3x3_shift(const float* x, float* x_m, float* x_p, float* x_0, long long rows, int w, int flags, void* stream) {
  return dcnet_conv3x3_shift_padded(x, x_m, x_p, x_0, rows, w, w, flags, stream);
}

// the same with the outputs at row pitch wp >= w (pad columns zero): maps whose width is not a multiple of 4 run the implicit GEMM at a
// padded width -- every row shift is then 16-byte aligned for TMA, the zero pad columns play the part of the image border, and the
// results at the pad columns are dropped (dcnet_conv3x3_unpad).  x_0 is required when wp != w.
extern "C" int dcnet_conv3x3_shift_padded(const float* x, float* x_m, float* x_p, float* x_0, long long rows, int w, int wp, int flags,
                                          void* stream) {
  DCNET_CHECK_ARG(x && x_m && x_p && rows > 0 && w >= 2 && wp >= w && (wp == w || x_0), "conv3x3_shift: bad arguments");
  const long long total = rows * wp;
  conv3x3_shift_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x, x_m, x_p, x_0, total, w, wp, (flags & DCNET_RN_TF32) ? 1 : 0);
  DCNET_LAUNCH_OK("conv3x3_shift");
  return 0;
}

extern "C" int dcnet_conv3x3_unpad(const float* zp, float* z, long long rows, int w, int wp, void* stream) {
  DCNET_CHECK_ARG(zp && z && rows > 0 && w >= 1 && wp >= w, "conv3x3_unpad: bad arguments");
  const long long total = rows * w;
  conv3x3_unpad_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(zp, z, total, w, wp);
  DCNET_LAUNCH_OK("conv3x3_unpad");
  return 0;
}

extern "C" int dcnet_conv3x3_pack_weight(const float* W, float* Wq, int Cout, int Cin, int flags, void* stream) {
  DCNET_CHECK_ARG(W && Wq && Cout > 0 && Cin > 0, "conv3x3_pack_weight: bad arguments");
  conv3x3_pack_weight_kernel<<<ceil_div(9ll * Cout * Cin, 256), 256, 0, as_stream(stream)>>>(W, Wq, Cout, Cin, (flags & DCNET_RN_TF32) ? 1 : 0);
  DCNET_LAUNCH_OK("conv3x3_pack_weight");
  return 0;
}

extern "C" int dcnet_conv3x3_fwd(const float* x_m, const float* x_0, const float* x_p, const float* Wq, float* z,
                                 int B, int Cin, int Cout, int h, int w, float* stat_sums, void* stream) {
  DCNET_CHECK_ARG(x_m && x_0 && x_p && Wq && z && B > 0, "conv3x3_fwd: bad arguments");
  DCNET_CHECK_ARG(shape_ok(Cin, Cout, h, w, x_m, x_0, x_p, z) && reinterpret_cast<uintptr_t>(Wq) % 16 == 0,
                  "conv3x3_fwd: needs Cin %% 256 == 0, Cout %% 128 == 0, w %% 4 == 0 and 16-byte aligned buffers");
  cudaStream_t st = as_stream(stream);
  const int N = h * w;
  if (stat_sums) DCNET_CUDA(cudaMemsetAsync(stat_sums, 0, 2 * (size_t)Cout * sizeof(float), st), "conv3x3_fwd.memset");
  UmmaOperand A{Wq, Cout, Cin, Cin, (long long)Cout * Cin, 9, false};
  UmmaOperand Bm{x_m, Cin, N, N, (long long)Cin * N, B, true}, B0{x_0, Cin, N, N, (long long)Cin * N, B, true},
      Bp{x_p, Cin, N, N, (long long)Cin * N, B, true};
  UmmaEpilogue e{};
  e.out = z; e.ldo = N; e.so_b = (long long)Cout * N; e.alpha = 1.f;
  e.sum = stat_sums; e.sumsq = stat_sums ? stat_sums + Cout : nullptr;
  e.B3 = &Bp; e.tap_kper = Cin / 32; e.tap_w = w;
  return umma_gemm(A, Bm, &B0, Cout, N, 9 * Cin, 0, 0, B, e, st);
}

extern "C" int dcnet_conv3x3_bwd_data(const float* dz_m, const float* dz_0, const float* dz_p, const float* Wq, float* dx,
                                      int B, int Cin, int Cout, int h, int w, void* stream) {
  DCNET_CHECK_ARG(dz_m && dz_0 && dz_p && Wq && dx && B > 0, "conv3x3_bwd_data: bad arguments");
  DCNET_CHECK_ARG(shape_ok(Cout, Cin, h, w, dz_m, dz_0, dz_p, dx) && reinterpret_cast<uintptr_t>(Wq) % 16 == 0,
                  "conv3x3_bwd_data: needs Cout %% 256 == 0, Cin %% 128 == 0, w %% 4 == 0 and 16-byte aligned buffers");
  const int N = h * w;
  // A = Wq[t]^T: an MN-major operand whose reduction rows are the output channels and whose M columns are the input channels
  UmmaOperand A{Wq, Cout, Cin, Cin, (long long)Cout * Cin, 9, true};
  UmmaOperand Bm{dz_m, Cout, N, N, (long long)Cout * N, B, true}, B0{dz_0, Cout, N, N, (long long)Cout * N, B, true},
      Bp{dz_p, Cout, N, N, (long long)Cout * N, B, true};
  UmmaEpilogue e{};
  e.out = dx; e.ldo = N; e.so_b = (long long)Cin * N; e.alpha = 1.f;
  e.B3 = &Bp; e.tap_kper = Cout / 32; e.tap_w = w; e.tap_flip = 1;
  return umma_gemm(A, Bm, &B0, Cin, N, 9 * Cout, 0, 0, B, e, as_stream(stream));
}

extern "C" int dcnet_conv3x3_bwd_weight(const float* dz, const float* x_m, const float* x_0, const float* x_p, float* dWp, float* dW,
                                        int B, int Cin, int Cout, int h, int w, void* stream) {
  DCNET_CHECK_ARG(dz && x_m && x_0 && x_p && dWp && dW && B > 0, "conv3x3_bwd_weight: bad arguments");
  DCNET_CHECK_ARG(shape_ok(Cin, Cout, h, w, x_m, x_0, x_p, dz) && reinterpret_cast<uintptr_t>(dWp) % 16 == 0,
                  "conv3x3_bwd_weight: needs Cin %% 256 == 0, Cout %% 128 == 0, w %% 4 == 0 and 16-byte aligned buffers");
  cudaStream_t st = as_stream(stream);
  const int N = h * w;
  // the (b, position) reduction is split over the images: fp32 reduce-add into the zeroed packed gradient
  DCNET_CUDA(cudaMemsetAsync(dWp, 0, (size_t)Cout * 9 * Cin * sizeof(float), st), "conv3x3_bwd_weight.memset");
  UmmaOperand A{dz, Cout, N, N, (long long)Cout * N, B, false};
  UmmaOperand Bm{x_m, Cin, N, N, (long long)Cin * N, B, false}, B0{x_0, Cin, N, N, (long long)Cin * N, B, false},
      Bp{x_p, Cin, N, N, (long long)Cin * N, B, false};
  UmmaEpilogue e{};
  e.out = dWp; e.ldo = 9ll * Cin; e.so_b = 0; e.alpha = 1.f; e.atomic = 1;
  e.B3 = &Bp; e.tap_n = Cin; e.tap_w = w;
  DCNET_TRY(umma_gemm(A, Bm, &B0, Cout, 9 * Cin, N, 0, 0, B, e, st));
  conv3x3_unpack_wgrad_kernel<<<ceil_div(9ll * Cout * Cin, 256), 256, 0, st>>>(dWp, dW, Cout, Cin);
  DCNET_LAUNCH_OK("conv3x3_unpack_wgrad");
  return 0;
}
