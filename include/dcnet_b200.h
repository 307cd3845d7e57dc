/*
 * dcnet_b200.h -- C ABI of libdcnet_sm100.so, the B200-native (sm_100a) implementation of the
 * DCNet dense-correspondence hot path (mengcaopku/DCNet, model/DCNet_model.py:356-650 and the loss /
 * target / decode functions of train_DCNet.py:45-332).
 *
 * The reference has no FFI of its own (it is 100 % PyTorch); each entry point below names the reference
 * lines it replaces.  Conventions:
 *   - plain pointers and sizes only; every buffer (inputs, outputs, workspace) is DEVICE memory owned by
 *     the caller unless the name starts with h_ (host memory); the library never allocates, frees or
 *     retains pointers;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit synchronisation;
 *   - return 0 on success, <0 for an argument error, >0 = cudaError_t of the failed launch;
 *     dcnet_last_error() returns a thread-local message;
 *   - feature maps are fp32 [B, C, N] = flattened NCHW exactly as the reference holds them
 *     (N = h*w innermost); images 2p and 2p+1 are the two frames of pair p (model/DCNet_model.py:365-374);
 *   - int64 index outputs match torch.long in the reference API.
 */
#ifndef DCNET_B200_H_
#define DCNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCNET_ABI_VERSION 4
#define DCNET_API __attribute__((visibility("default")))

/* ---- library ---------------------------------------------------------------------------------------- */
DCNET_API int dcnet_abi_version(void);
DCNET_API const char* dcnet_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
DCNET_API long long dcnet_launch_count(void);

/* ---- generic strided batched fp32 GEMM (building block; exact-fp32 path) ------------------------------
 * C[b](m,n) = alpha * sum_{kb<kbatch} sum_k A[b](m,k;kb) * B[b](k,n;kb) * colscale[b](n) + beta * C[b](m,n)
 * element (m,k) of A for batch b, k-batch kb: A + ia(b)*sAb + kb*sAkb + m*sAm + k*sAk   (ia(b) = idxA ? idxA[b] : b)
 * atomic != 0: C is updated with atomicAdd (beta ignored, caller pre-zeroes).                           */
DCNET_API int dcnet_sgemm(const float* A, const float* B, float* C, int M, int N, int K, int batch, int kbatch,
                          long long sAm, long long sAk, long long sAb, long long sAkb,
                          long long sBk, long long sBn, long long sBb, long long sBkb,
                          long long sCm, long long sCn, long long sCb,
                          const int* idxA, const int* idxB, const int* idxC,
                          float alpha, float beta, const float* colscale, long long sColscaleB,
                          int atomic, void* stream);

/* ---- batched TF32 tensor-core GEMM (tcgen05.mma, TMA-fed, fp32 accumulation in TMEM) ----------------------------
 * C[b] (M x N, row pitch ldc) (+)= alpha * A[b] (M x K) * B[b] (K x N).  Operand storage (fp32, innermost axis contiguous):
 *   a_mn_major = 0: A[b] stored [M][K] (pitch lda)     a_mn_major = 1: stored [K][M]
 *   b_mn_major = 0: B[b] stored [N][K] (pitch ldb)     b_mn_major = 1: stored [K][N]
 * Requirements: 16-byte aligned bases, pitches and batch strides multiples of 4 floats.  atomic != 0: atomicAdd into C. */
DCNET_API int dcnet_gemm_tf32(const float* A, int a_mn_major, long long lda, long long strideA,
                              const float* B, int b_mn_major, long long ldb, long long strideB,
                              float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                              void* stream);

/* same contraction with bf16 operands (tcgen05 kind::f16, fp32 accumulation).  A, B: __nv_bfloat16 data; pitches / strides in
 * elements, multiples of 8.  dcnet_cast_bf16 converts fp32 -> bf16 with round-to-nearest-even.                            */
DCNET_API int dcnet_gemm_bf16(const void* A, int a_mn_major, long long lda, long long strideA,
                              const void* B, int b_mn_major, long long ldb, long long strideB,
                              float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                              void* stream);
DCNET_API int dcnet_cast_bf16(const float* x, void* y, long long n, void* stream);
/* and with fp16 operands (__half; dcnet_cast_f16 rounds to nearest even): the 11 significant bits of tf32 at twice its MMA rate, for
 * operands inside fp16's normal range.  This is what the co-attention backward runs on (dcnet_coattn_bwd with `staged`).  atomic = 1:
 * reduce-add into C, the reduction split over CTAs when the tiles do not fill the SMs.                                          */
DCNET_API int dcnet_gemm_f16(const void* A, int a_mn_major, long long lda, long long strideA,
                             const void* B, int b_mn_major, long long ldb, long long strideB,
                             float* C, long long ldc, long long strideC, int M, int N, int K, int batch, float alpha, int atomic,
                             void* stream);
DCNET_API int dcnet_cast_f16(const float* x, void* y, long long n, void* stream);
/* which tensor-core GEMM kernel every entry point of this library uses: 0 (default) = persistent kernel, two TMEM accumulator
 * stages, epilogue through swizzled smem slots and TMA store / reduce-add; 5 = the same with coalesced 16-byte st.global /
 * red.global.add.v4.f32 from the slots; 3 / 4 = with thread-block clusters of 2 / 4 CTAs (consecutive M tiles) that multicast
 * the B tile; 1 = one tile per CTA with per-thread row stores (outputs whose rows are not 16-byte addressable).  All variants give
 * identical bits for plain stores.  6 = no cta_group::2 pairs; 7 = 4 instead of 8 epilogue warps in the fp16 exp / dS epilogues of the
 * co-attention backward.  Process-wide switch, not thread-safe: a test / bring-up knob.                                          */
DCNET_API int dcnet_gemm_select(int variant);
/* profiling: while buf != NULL every persistent GEMM launch writes clock64 stamps to buf [148 CTAs][8 tiles][8]: 0 tile start (MMA
 * thread), 1 accumulator stage free, 2 first operands landed, 3 last MMA issued, 4 accumulator complete (epilogue), 5 epilogue done */
DCNET_API int dcnet_gemm_trace(long long* buf);

/* ---- operand rounding for the tf32 contractions.  tcgen05.mma.kind::tf32 reads fp32 operands by truncating them to 10 mantissa
 * bits: a relative bias of about -2^-12 per operand, which adds up along a chain of contractions (the backward of the path is ~10
 * of them: 5.7e-3 after 8 layers against 8e-4 with rounded operands).  A value already rounded to the nearest tf32 passes the MMA
 * unchanged.  DCNET_RN_TF32, or-ed into the flag argument named at each entry point, makes a producer round what it hands to a tf32
 * contraction (dcnet_bn_act_fwd: `l2norm`; dcnet_bn_act_bwd_apply: `train`; dcnet_coattn_fwd: `precision`); dcnet_round_tf32 rounds
 * a tensor the library did not produce (weights, the Darknet maps).  y may alias x.                                              */
#define DCNET_RN_TF32 0x100
DCNET_API int dcnet_round_tf32(const float* x, float* y, long long n, void* stream);

/* ---- a1/a2/a6/a8: 1x1 conv (no bias) + BatchNorm + ReLU (+ L2 norm over channels) ---------------------
 * replaces ConvBatchNormReLU (model/darknet.py:118-156) as used by mapping_visu (:356-359), corr_conv
 * (:467-469) and fcn_emb[s][0] (:505), and F.normalize(dim=1).
 *
 * conv1x1_fwd:  z[b] = W[:, 0:K1] x1[b] + W[:, K1:K1+K2] x2[b] + u[b] 1^T + cc
 *   x1 [B,K1,N]; x2 [B,K2,N] or NULL (two-source K loop: corr_conv reads [fvisu | attention] without a cat);
 *   W [C, ldw] row-major (ldw >= K1+K2); u [B,C] or NULL (text term W_l flang of the fusion, Appendix A.9);
 *   cc [C,N] or NULL (coordinate term W_c coord); z [B,C,N].                                              */
DCNET_API int dcnet_conv1x1_fwd(const float* x1, int K1, const float* x2, int K2, const float* W, int ldw,
                                const float* u, const float* cc, float* z, int B, int C, int N,
                                float* stat_sums, int precision, void* stream);
/* precision (all three conv entry points): 0 = exact fp32 on the CUDA cores (used where indices depend on the result:
 * the scale-0 visual mapping that feeds the top-30 / arg-max selections); 1 = TF32 tcgen05 tensor-core path (fp32 operands
 * read as TF32 by the MMA, fp32 accumulation in TMEM).  Shapes TMA cannot address (N % 4 != 0) run at precision 0.
 * stat_sums (optional, [2*C]): receives sum_z[C] and sum_z2[C] over (b,n) for dcnet_bn_finalize -- accumulated in the GEMM
 * epilogue on the tensor-core path, so z is not re-read for the BatchNorm statistics.                                   */
/* dx1 = W[:,0:K1]^T dz, dx2 = W[:,K1:]^T dz (either may be NULL) */
DCNET_API int dcnet_conv1x1_bwd_data(const float* dz, const float* W, int ldw, float* dx1, int K1, float* dx2, int K2,
                                     int B, int C, int N, int precision, void* stream);
/* same; dx2_absmax [B] (optional, tensor-core path with both outputs) receives max |dx2[b]| as the bit pattern of a non-negative float,
 * out of the GEMM's epilogue: corr_conv's dx2 is the co-attention's dout, whose per-problem scale dcnet_coattn_bwd_ex then needs no pass for */
DCNET_API int dcnet_conv1x1_bwd_data_absmax(const float* dz, const float* W, int ldw, float* dx1, int K1, float* dx2, int K2,
                                            int B, int C, int N, int precision, unsigned int* dx2_absmax, void* stream);
/* dW[:,0:K1] = sum_b dz[b] x1[b]^T, dW[:,K1:] = sum_b dz[b] x2[b]^T  (overwrites those columns of dW [C,ldw]);
 * du [B,C] = sum_n dz (or NULL); dcc [C,N] = sum_b dz (or NULL)                                           */
DCNET_API int dcnet_conv1x1_bwd_weight(const float* dz, const float* x1, int K1, const float* x2, int K2,
                                       float* dW, int ldw, float* du, float* dcc, int B, int C, int N, int precision, void* stream);

/* ---- 8(f) row 1: the grounding head's 3x3 convolution (fcn_emb[s][1] = ConvBatchNormReLU(C, C, 3, 1, 1), model/DCNet_model.py:316-337,
 * :505-506; ConvBatchNormReLU model/darknet.py:118-156) as an implicit GEMM on tcgen05 (csrc/conv3x3.cu).  Maps are [B, C, N = h*w].
 *   dcnet_conv3x3_shift:       x_m[p] = x[p-1], x_p[p] = x[p+1] inside each image row of w positions (zero at the border column); x_0
 *                              (optional) = copy of x; flags = DCNET_RN_TF32 rounds all outputs.  rows = B*C*h.
 *   dcnet_conv3x3_pack_weight: W [Cout,Cin,3,3] -> Wq [9,Cout,Cin] (flags = DCNET_RN_TF32: rounded)
 *   dcnet_conv3x3_fwd:         z [B,Cout,N] = conv3x3(x) (stride 1, zero padding 1, no bias); stat_sums like dcnet_conv1x1_fwd
 *   dcnet_conv3x3_bwd_data:    dx [B,Cin,N] from the three shifted copies of dz [B,Cout,N]
 *   dcnet_conv3x3_bwd_weight:  dW [Cout,Cin,3,3] (overwritten); dWp [Cout,9,Cin] is scratch
 * Needs Cin, Cout multiples of 256 and w % 4 == 0 (a TMA box shifted by one image row must start 16-byte aligned:
 * dcnet_conv3x3_supported).  Other widths run at a padded width wp (multiple of 4): dcnet_conv3x3_shift_padded writes the copies at
 * row pitch wp with zero pad columns, the three contractions are called with (h, wp), dcnet_conv3x3_unpad drops the pad columns of
 * z / dx (stat_sums = NULL there: the BatchNorm statistics come from the unpadded z).                                             */
DCNET_API int dcnet_conv3x3_supported(int Cin, int Cout, int h, int w);
DCNET_API int dcnet_conv3x3_shift(const float* x, float* x_m, float* x_p, float* x_0, long long rows, int w, int flags, void* stream);
DCNET_API int dcnet_conv3x3_shift_padded(const float* x, float* x_m, float* x_p, float* x_0, long long rows, int w, int wp, int flags,
                                         void* stream);
DCNET_API int dcnet_conv3x3_unpad(const float* zp, float* z, long long rows, int w, int wp, void* stream);
DCNET_API int dcnet_conv3x3_pack_weight(const float* W, float* Wq, int Cout, int Cin, int flags, void* stream);
DCNET_API int dcnet_conv3x3_fwd(const float* x_m, const float* x_0, const float* x_p, const float* Wq, float* z,
                                int B, int Cin, int Cout, int h, int w, float* stat_sums, void* stream);
DCNET_API int dcnet_conv3x3_bwd_data(const float* dz_m, const float* dz_0, const float* dz_p, const float* Wq, float* dx,
                                     int B, int Cin, int Cout, int h, int w, void* stream);
DCNET_API int dcnet_conv3x3_bwd_weight(const float* dz, const float* x_m, const float* x_0, const float* x_p, float* dWp, float* dW,
                                       int B, int Cin, int Cout, int h, int w, void* stream);

/* ---- a8: text / coordinate terms of the split-weight fusion (model/DCNet_model.py:489-505; ancestor
 * model/grounding_model_semantic_attn.py:266-281).  W [C,ldw] = [W_v | W_l | W_c]: u [B,C] = flang [B,Ct] W_l^T with
 * W_l = W[:, col_l:col_l+Ct]; cc [C,N] = W_c coord with W_c = W[:, col_c:col_c+8], coord [8,N] (dcnet_coord_map).  coord / cc may be
 * NULL (coordmap=False).  u and cc are what dcnet_conv1x1_fwd adds in its epilogue, so the [B,1032,h,w] concatenation of the
 * reference is never built.                                                                                               */
DCNET_API int dcnet_fuse_terms_fwd(const float* W, int ldw, int col_l, int Ct, int col_c, const float* flang, const float* coord,
                                   float* u, float* cc, int B, int C, int N, void* stream);
/* backward from du [B,C] / dcc [C,N] (dcnet_conv1x1_bwd_weight): dflang [B,Ct] (or NULL) and the text / coordinate columns of
 * dW [C,ldw] (overwritten; NULL = no weight gradient)                                                                     */
DCNET_API int dcnet_fuse_terms_bwd(const float* W, int ldw, int col_l, int Ct, int col_c, const float* flang, const float* coord,
                                   const float* du, const float* dcc, float* dflang, float* dW, int B, int C, int N, void* stream);

/* mean/invstd (+ running statistics, momentum, unbiased variance) from the epilogue sums: count = B*N.
 * num_batches_tracked (optional, one int64 on the device) is incremented like nn.BatchNorm2d does per training step. */
DCNET_API int dcnet_bn_finalize(const float* stat_sums, long long count, int C, float eps, float momentum,
                                float* mean, float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                                void* stream);

/* BatchNorm statistics of z [B,C,N] over (B,N): mean[C], invstd[C] = 1/sqrt(biased var + eps); when
 * running_mean/var are non-NULL they are updated in place with `momentum` and the UNBIASED variance
 * (PyTorch convention; the reference uses momentum 0.999, eps 1e-5, model/darknet.py:145).               */
DCNET_API int dcnet_bn_stats(const float* z, int B, int C, int N, float eps, float momentum,
                             float* mean, float* invstd, float* running_mean, float* running_var, long long* num_batches_tracked,
                             void* stream);
/* eval mode: mean = running_mean, invstd = 1/sqrt(running_var + eps) */
DCNET_API int dcnet_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                                  float* mean, float* invstd, void* stream);
/* y = act(gamma (z-mean) invstd + beta), act = ReLU (slope 0) or LeakyReLU(slope); if l2norm: y /= max(||y||_c, 1e-12).
 * If fa != NULL (text vectors [B,C], a9): sim[b,n] = <fa[b], y[b,:,n]>, neg_sim[b,n] = <fa[B-1-b], y[b,:,n]>
 * (model/DCNet_model.py:530-535, train_DCNet.py:623-627), fused so corr_feat is read once.               */
DCNET_API int dcnet_bn_act_fwd(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                               float slope, int l2norm, float* y, const float* fa, const float* fa_neg, float* sim, float* neg_sim,
                               int B, int C, int N, void* stream);
/* same, and fills `staged` (dcnet_coattn_stage_bytes(B, C, N) bytes, 256-byte aligned) with what dcnet_coattn_stage would make of y
 * -- fp16 copy, column norms, largest norm per frame -- so the fused co-attention forward starts from the producer's registers
 * instead of re-reading the map (staged = NULL: plain dcnet_bn_act_fwd).  Follow with dcnet_coattn_fused_fwd.                     */
DCNET_API int dcnet_bn_act_fwd_staged(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                      float slope, int l2norm, float* y, const float* fa, const float* fa_neg, float* sim, float* neg_sim,
                                      int B, int C, int N, void* staged, size_t staged_bytes, void* stream);
/* fa_neg (optional, [B,C]): explicit text vector of each image's negative partner (cross-GPU negatives: the partner of global
 * sample g is Bg-1-g and may live on another rank); NULL = the reference's local batch reversal fa[B-1-b].                 */
/* Backward of bn_act_fwd in train mode (batch statistics).  Two launches:
 *   reduce: dv = d(pre-activation) from dy (+ dsim/dneg_sim), written to `dv` [B,C,N]; accumulates
 *           sum_dv[C], sum_dvz[C] (caller zeroes), dfa [B,C] (caller zeroes, may be NULL);
 *   apply : dz = gamma invstd (dv - sum_dv/M - zhat sum_dvz/M)  (in place allowed: dz == dv);
 *           dgamma = sum_dvz, dbeta = sum_dv.
 * train == 0 (running statistics): dz = gamma invstd dv.                                                  */
DCNET_API int dcnet_bn_act_bwd_reduce(const float* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                      float slope, int l2norm, const float* dy, const float* fa, const float* fa_neg, const float* dsim,
                                      const float* dneg_sim, float* dv, float* sum_dv, float* sum_dvz, float* dfa, float* dfa_neg,
                                      int B, int C, int N, void* stream);
/* 0 (default): dcnet_bn_act_bwd_reduce uses the persistent smem-staged kernel when C == 512 and the maps are 16-byte addressable
 * (N % 4 == 0, 16-byte aligned); 1: always the register-staged kernel.  Test / bring-up knob, process-wide.                */
DCNET_API int dcnet_bn_bwd_select(int variant);
DCNET_API int dcnet_bn_act_bwd_apply(const float* z, const float* mean, const float* invstd, const float* gamma,
                                     const float* dv, const float* sum_dv, const float* sum_dvz, int train,
                                     float* dz, int B, int C, int N, void* stream);

/* ---- a9 stand-alone (model/DCNet_model.py:530-535): sim[b,n] = <fa[b], x[b,:,n]>; neg_sim (optional) with fa_neg[b] or fa[B-1-b].
 * The training path gets these from dcnet_bn_act_fwd's epilogue; the clip path (mean of several corr maps) calls this.  Forward only. */
DCNET_API int dcnet_pix2text(const float* x, const float* fa, const float* fa_neg, float* sim, float* neg_sim, int B, int C, int N, void* stream);

/* ---- location branch at inference, rank-8 form (model/DCNet_model.py:556-603, model/test_DCNet_model.py same
 * lines).  Replaces   rel = bmm(E, E^T) * obj[:,None,:]  ->  Linear(SN -> C)  ->  BatchNorm1d (eval)  ->  ReLU  ->  normalize over C
 * -> dot with the location phrase vector -> min-max over the positions   without the [B,SN,SN] tensor:
 *   E [SN,8] normalised coordinate embeddings (batch independent), obj [B,SN] L2-normalised objectness, W [C,ldw>=SN] + bias [C]
 *   (NULL = none) the Linear, bn_scale = gamma / sqrt(running_var + eps), bn_shift = beta - running_mean * bn_scale, flang [B,C].
 *   Workspaces: G [B,C,8], raw [B,SN] (the scores before min-max).  Output score [B,SN] in [0,1].  C <= 1024.  Forward only. */
DCNET_API int dcnet_loc_rank8_fwd(const float* E, const float* obj, const float* W, int ldw, const float* bias,
                                  const float* bn_scale, const float* bn_shift, const float* flang,
                                  float* G, float* raw, float* score, int B, int SN, int C, void* stream);

/* ---- a7: coordinate map (model/DCNet_model.py:23-39), [8,h,w], batch independent ---------------------- */
/* the same branch in training mode (batch statistics of the BatchNorm1d, running statistics updated like nn.BatchNorm1d, and the
 * backward): forward keeps G [B,C,8], stats [4*C] = (BN scale, BN shift, mean of z without the Linear's bias, invstd), raw and
 * inrm [B,SN] (score before min-max, 1 / channel norm) for the backward; mom [72] is scratch.  Backward from dscore [B,SN]: dE [SN,8],
 * dobj [B,SN], dW [C,ldw] (columns < SN), dgamma / dbeta [C], dflang [B,C], all overwritten (the Linear's bias has no gradient under
 * batch statistics); draw [B,SN] and dG [B,C,8] are scratch.  The [B,SN,SN] relation tensor and the [B*SN, C] activations never exist. */
DCNET_API int dcnet_loc_rank8_train_fwd(const float* E, const float* obj, const float* W, int ldw, const float* bias,
                                        const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                                        float* running_var, long long* num_batches_tracked, const float* flang,
                                        float* G, float* mom, float* stats, float* raw, float* inrm, float* score,
                                        int B, int SN, int C, void* stream);
DCNET_API int dcnet_loc_rank8_train_bwd(const float* E, const float* obj, const float* W, int ldw, const float* bias, const float* flang,
                                        const float* G, const float* stats, const float* raw, const float* inrm, const float* dscore,
                                        float* draw, float* dG, float* dE, float* dobj, float* dW, float* dgamma, float* dbeta, float* dflang,
                                        int B, int SN, int C, void* stream);
DCNET_API int dcnet_coord_map(float* coord, int h, int w, void* stream);

/* ---- a5/a20: co-attention (model/DCNet_model.py:449-459, model/test_DCNet_model.py:247-274) -----------
 * One "problem" i is one direction: queries = frame qa[i], keys/values = frame kb[i] of frames [F,C,N]:
 *   S = Fa^T Fb,  P = softmax_j(tau S),  out[oidx[i]] = Fb P^T  ([C,N]),  lse[i][n] = log sum_j exp(tau S[n,j]).
 * A training pair p is the two problems (2p,2p+1) and (2p+1,2p); the test-time clip path uses the centre
 * direction only.  precision: 0 = exact fp32 (CUDA cores), 1 = tcgen05 TF32 GEMMs with S/P in workspace,
 * 2 = fused tcgen05 kernel (fp16 operands -- 11 significant bits like tf32 --, fp32 accumulation in TMEM): S / P never leave
 * the SM; workspace holds the fp16 staging of the maps (|values| < 65504; the model's maps are unit-norm).  The fused forward assumes every logit of a row lies within ~80 of
 * tau |Fa_q| max_k |Fb_k| (always true for the unit-norm maps of the model, model/DCNet_model.py:359).
 * The backward of precision 2 is the precision-1 composition.                                            */
DCNET_API size_t dcnet_coattn_workspace_bytes(int F, int nprob, int C, int N, int precision);
DCNET_API int dcnet_coattn_fwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                               float* out, int n_out, float* lse, int C, int N, float tau, int precision,
                               void* workspace, size_t workspace_bytes, void* stream);
/* the two halves of the precision-2 forward, callable separately (the clip path stages a clip once and runs many problems):
 * dcnet_coattn_stage: staged <- fp16 copy of frames (row pitch padded to 8 elements) + column norms; dcnet_coattn_fused_fwd: the
 * fused TMA/tcgen05 kernel over a staged buffer (grid = 64-query tiles x problems).  C % 128 == 0, C <= 512.                 */
DCNET_API size_t dcnet_coattn_stage_bytes(int F, int C, int N);
DCNET_API int dcnet_coattn_stage(const float* frames, int F, int C, int N, void* staged, size_t staged_bytes, void* stream);
DCNET_API int dcnet_coattn_fused_fwd(const void* staged, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                     float* out, int n_out, float* lse, int C, int N, float tau, int flags, void* stream);
/* training form: the kernel also keeps its unnormalised softmax weights, transposed as it holds them -- E^T[z][key][q] (fp16,
 * dcnet_coattn_keep_bytes(nprob, N) bytes, row pitch N rounded up to 8; one TMA store per key tile from the shared-memory tile the
 * second MMA reads) -- and their row sums r [nprob,N] for dcnet_coattn_bwd_ex, whose contractions then start without recomputing
 * S = Fa^T Fb (HBM is not scarce: 468 MB for 32 problems at N = 2704 against a 240 GFLOP contraction per step)                        */
DCNET_API size_t dcnet_coattn_keep_bytes(int nprob, int N);
DCNET_API int dcnet_coattn_fused_fwd_keep(const void* staged, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                          float* out, int n_out, float* lse, int C, int N, float tau, int flags, void* e_keep, float* r_keep,
                                          void* stream);
/* flags: 0 or DCNET_RN_TF32 (out leaves rounded to the nearest tf32) */
/* profiling variant: trace [ceil(N/64) * nprob CTAs][ceil(N/128) key tiles + 1][8] int64 receives clock64 stamps (see umma_coattn.cu) */
DCNET_API int dcnet_coattn_fused_fwd_trace(const void* staged, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                           float* out, int n_out, float* lse, int C, int N, float tau, long long* trace, int variant, void* stream);
/* dframes [F,C,N] += gradient (caller zeroes, or hands over a buffer that already holds another consumer's gradient of the
 * frames); dout/out indexed by oidx like the forward.  P is re-normalised from tf32 logits recomputed here (the forward's lse is
 * only a shift).  staged (optional, precision 2): the fp16 staging of `frames` the forward used (dcnet_coattn_stage /
 * dcnet_bn_act_fwd_staged).  With it the five contractions run on fp16 operands -- kind::f16: the 11 significant bits of tf32 at twice
 * the MMA rate and half the bytes for the two N x N tensors; everything that scales with dout carries a per-problem power of two
 * that keeps it in fp16's normal range --; NULL: tf32 contractions on the fp32 maps.                                           */
DCNET_API int dcnet_coattn_bwd(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                               const float* out, int n_out, const float* lse, const float* dout, float* dframes,
                               int C, int N, float tau, int precision, const void* staged, void* workspace, size_t workspace_bytes,
                               void* stream);
/* same with dout_absmax [n_out] (optional): max |dout[row]| per output row as float bit patterns (dcnet_conv1x1_bwd_data_absmax); the
 * fp16 pipeline then skips its own pass over dout */
DCNET_API int dcnet_coattn_bwd_ex(const float* frames, int F, const int* qa, const int* kb, const int* oidx, int nprob,
                                  const float* out, int n_out, const float* lse, const float* dout, const unsigned int* dout_absmax,
                                  float* dframes, int C, int N, float tau, int precision, const void* staged, const void* e_keep,
                                  const float* r_keep, void* workspace, size_t workspace_bytes, void* stream);
/* e_keep / r_keep (optional, with `staged`): what dcnet_coattn_fused_fwd_keep left for the same problems */
/* The backward keeps its N x N scratch (P, dP -> dS) resident in L2 by working through the problems in chunks whose scratch
 * fits `bytes` (default 64 MiB of the 126 MB L2; <= 0 = unlimited = one chunk); dcnet_coattn_workspace_bytes follows it. */
DCNET_API int dcnet_coattn_bwd_l2_budget(long long bytes);
/* 1 (default): the fp16 pipeline is used whenever `staged` is given; 0: always tf32 (comparison knob, process-wide); n > 1
 * (profiling): the fp16 pipeline returns after its (n-1)-th contraction, so that dcnet_gemm_trace holds that launch; -1: fp16 pipeline
 * that recomputes E even when the forward kept it (comparison) */
DCNET_API int dcnet_coattn_bwd_fp16(int on);

/* ---- a4: inter-frame patch correspondence (model/DCNet_model.py:381-430) ------------------------------
 * fv0 [2P,C,N0].  S0[p] = F1^T F2 in exact fp32; idx[p, r] = flat index (row*N0+col) of the r-th largest
 * entry (descending; ties -> lower flat index first).  S0 scratch [P,N0,N0] is caller-provided.           */
DCNET_API int dcnet_interframe_topk(const float* fv0, int P, int C, int N0, int top_k, float* S0, long long* idx, void* stream);
/* negative index mapping: the host draws POSITIONS pos in [0,N0-2] with the reference's RNG stream
 * (random.sample over a population of N0-1, :411-413); the kernel maps them to pixels skipping col:
 * pix = pos + (pos >= col).  negidx [P,top_k,neg_n] int64 out.                                            */
DCNET_API int dcnet_interframe_negidx(const long long* idx, const int* negpos, int P, int N0, int top_k, int neg_n,
                                      long long* negidx, void* stream);

/* cols [top_k*P*(2+neg_n)] int64, rank-major: [frame-1 column idx//N0 : top_k x P | frame-2 column idx%N0 : top_k x P |
 * negative columns : top_k x P x neg_n] -- the gather columns of q, k and neg (model/DCNet_model.py:407-420) in one launch,
 * ordered so that the gathered rows are the packed [rank][pair] tensors of the contrastive loss.                        */
DCNET_API int dcnet_interframe_cols(const long long* idx, const int* negpos, int P, int N0, int top_k, int neg_n, long long* cols, void* stream);

/* ---- generic column gather / scatter-add (a4, a11 gathers and their backward) -------------------------
 * out[i, :] = src[img[i], :, col[i]]   (src [F,C,N]; out [n,C]);  backward: dsrc[img[i], :, col[i]] += dout[i,:] */
DCNET_API int dcnet_gather_cols(const float* src, const int* img, const long long* col, int n, float* out, int C, int N, void* stream);
DCNET_API int dcnet_scatter_cols_add(const float* dout, const int* img, const long long* col, int n, float* dsrc, int C, int N, void* stream);

/* ---- a12/a13: InfoNCE-style contrastive loss (train_DCNet.py:114-166) ----------------------------------
 * q [G,C], k [G,C], neg [G,n,C] (G = ranks*pairs or pixels*images).  loss_g = CE([q^.k^, q^.n^_j]/T, 0) with every
 * vector L2-normalised (eps 1e-12).  rowloss [G]; the mean is taken by the caller (all groups have equal size).
 * bwd: dq,dk,dneg given gscale[g*gstride] = dL/d(rowloss_g) (gstride 0 = one shared scalar).                                            */
DCNET_API int dcnet_infonce_fwd(const float* q, const float* k, const float* neg, int G, int n, int C, float T, float* rowloss, void* stream);
DCNET_API int dcnet_infonce_bwd(const float* q, const float* k, const float* neg, int G, int n, int C, float T, const float* gscale,
                                int gstride, float* dq, float* dk, float* dneg, void* stream);

/* ---- a11: cross-modal block (model/DCNet_model.py:625-637, :41-112) ------------------------------------
 * axis-normalisations used only there: rows = F.normalize over the innermost axis of [R, L] rows
 * (vit: over the spatial axis, :629); lag: context [B,T,2C] -> even channels (nearest 0.5x, :631) normalised over
 * the WORD axis (:632) -> lag [B,T,C].                                                                    */
DCNET_API int dcnet_rownorm_fwd(const float* x, float* y, float* nrm, long long R, int L, void* stream);
DCNET_API int dcnet_rownorm_bwd(const float* y, const float* nrm, const float* dy, float* dx, long long R, int L, void* stream);
DCNET_API int dcnet_lagnorm_fwd(const float* context, float* lag, float* nrm, int B, int T, int C, void* stream);
DCNET_API int dcnet_lagnorm_bwd(const float* lag, const float* nrm, const float* dlag, float* dcontext, int B, int T, int C, void* stream);
/* word[b,n] = argmax_t softmax_t(Conv1d_{T->T,k=3,pad=1}(M)[b,t,n]) (first maximum), M = lag . vit [B,T,N0]
 * computed inside (scratch M [B,T,N0] caller-provided).  fm_w [fm_cout,fm_cin,3], fm_b [fm_cout]; both must equal T
 * (the reference's Conv1d(20,20,3) raises on any other sentence length, model/DCNet_model.py:288,:635).          */
DCNET_API int dcnet_crossmodal_words(const float* lag, const float* vit, const float* fm_w, const float* fm_b,
                                     int fm_cout, int fm_cin, float* M, long long* word, int B, int T, int C, int N0, void* stream);

/* ---- a14: build_target (train_DCNet.py:265-332) -------------------------------------------------------
 * bbox [B,4] xyxy (pixels, already clamped).  Per sample: best_n (0..8, first maximum of the 9 anchor IoUs),
 * gi, gj (cell of the box centre at the best scale), t = (tx,ty,tw,th,1).  anchors: 9 (w,h) pairs, largest first.
 * If gt0..2 / gtc0..2 are non-NULL they receive the dense targets [B,3,5,g,g] / [B,5,g,g] (zero filled + scatter). */
DCNET_API int dcnet_build_target(const float* bbox, int B, int size, float anchor_imsize, const float* h_anchors9x2,
                                 long long* best_n, long long* gi, long long* gj, float* t5,
                                 float* gt0, float* gt1, float* gt2, float* gtc0, float* gtc1, float* gtc2, void* stream);

/* ---- a10: objectness / confidence modulation (model/DCNet_model.py:545-552, :612-621) -------------------
 * raw [B,15,N] -> only_obj [B,N] = mean_a raw[b,5a+4,n]; obj [B,N] = only_obj*sim;  (either out may be NULL)  */
DCNET_API int dcnet_only_obj(const float* raw, const float* sim, float* only_obj, float* obj, int B, int N, void* stream);
/* backward of dcnet_only_obj: draw [B,15,N] (confidence channels 5a+4 = (d_only_obj + d_obj sim)/3, the rest 0), dsim = d_obj only_obj;
 * either incoming gradient may be NULL                                                                     */
DCNET_API int dcnet_only_obj_bwd(const float* d_only_obj, const float* d_obj, const float* sim, const float* only_obj, float* draw, float* dsim,
                                 int B, int N, void* stream);
/* out = raw with channels 5a+4 multiplied by sim*loc */
DCNET_API int dcnet_modulate_conf_fwd(const float* raw, const float* sim, const float* loc, float* out, int B, int N, void* stream);
DCNET_API int dcnet_modulate_conf_bwd(const float* raw, const float* sim, const float* loc, const float* dout,
                                      float* draw, float* dsim, float* dloc, int B, int N, void* stream);

/* ---- a16/a17: grounding losses (train_DCNet.py:45-72, :173-220) ----------------------------------------
 * pred0..2 [B,15,N_s] (modulated), sim/neg_sim/loc 0..2 [B,N_s]; targets from dcnet_build_target.
 * losses[0..2] = yolo_loss, rank_loss, loc_loss (batch means as in the reference).
 * bwd: gl[3] = dL/d(yolo,rank,loc) (device); writes dpred (dense, zero elsewhere), dsim, dneg_sim, dloc.      */
DCNET_API int dcnet_ground_loss_fwd(const float* pred0, const float* pred1, const float* pred2,
                                    const float* sim0, const float* sim1, const float* sim2,
                                    const float* neg0, const float* neg1, const float* neg2,
                                    const float* loc0, const float* loc1, const float* loc2,
                                    const long long* best_n, const long long* gi, const long long* gj, const float* t5,
                                    const long long* partner3,
                                    int B, int g0, float w_coord, float margin, float* losses, float* lse_conf, float* lse_loc,
                                    void* stream);
/* partner3 (optional, [3,B] = best_n | gi | gj of each sample's rank-loss partner); NULL = local partner B-1-b (:195-196) */
DCNET_API int dcnet_ground_loss_bwd(const float* pred0, const float* pred1, const float* pred2,
                                    const float* sim0, const float* sim1, const float* sim2,
                                    const float* neg0, const float* neg1, const float* neg2,
                                    const float* loc0, const float* loc1, const float* loc2,
                                    const long long* best_n, const long long* gi, const long long* gj, const float* t5,
                                    const long long* partner3,
                                    int B, int g0, float w_coord, float margin, const float* lse_conf, const float* lse_loc,
                                    const float* gl,
                                    float* dpred0, float* dpred1, float* dpred2, float* dsim0, float* dsim1, float* dsim2,
                                    float* dneg0, float* dneg1, float* dneg2, float* dloc0, float* dloc1, float* dloc2,
                                    void* stream);

/* ---- 8f-3: test-time cache writer (test_DCNet.py:546-654, get_topk_pred_bbox :657-701) -----------------------------
 * pred0..2 [B,15,N_s] (modulated confidences at channels 5a+4), feat0..2 [B,C,N_s] (corr_feat), meta [B,5] = (ratio, dw, dh,
 * img_w, img_h) of the letterbox (test_DCNet.py:615-627).  Per image the k best (scale, anchor, cell) by confidence (ties ->
 * lower flat index; the cell is the FIRST of its scale with that confidence, :682): boxes [B,k,4] xyxy in the original image
 * (clamped like :693-696), scores [B,k], cells [B,k,4] = (scale, anchor, gj, gi), feats [B,k,C].                          */
DCNET_API int dcnet_topk_boxes(const float* pred0, const float* pred1, const float* pred2,
                               const float* feat0, const float* feat1, const float* feat2, int B, int g0, int size, int C, int k,
                               float anchor_imsize, const float* h_anchors9x2, const float* meta,
                               float* boxes, float* scores, long long* cells, float* feats, void* stream);
/* ---- 8f-3: post_processing.py:205-270.  centre [k,C]; ref [k,R,C] / ref_score [k,R] = cached features / scores of the top-k boxes
 * of R reference frames; invalid [R] (or NULL) marks frames whose cache was missing (:234-236, zeroed after the softmax :267).
 * fused [k] = sum_r softmax_r(max_j <centre_i, ref_jr>) * score of the arg-max box; best = first arg-max of fused;
 * match [k,R] (or NULL) the matched reference box per (i,r).  k, R <= 16.                                                  */
DCNET_API int dcnet_post_rescore(const float* centre, const float* ref, const float* ref_score, const int* invalid, int k, int R, int C,
                                 float* fused, long long* best, long long* match, void* stream);

/* ---- a18: decode (train_DCNet.py:656-690 at the GT cell; :766-816 arg-max) + a15 bbox_iou ---------------
 * mode 0: decode at the given (best_n, gi, gj); mode 1: arg-max over the 3*SN conf logits (first maximum in
 * (scale, anchor, gj, gi) order) and the cell is written to best_n/gi/gj.  boxes [B,4] xyxy pixels;
 * iou [B] vs target boxes (NULL to skip).                                                                  */
DCNET_API int dcnet_decode(const float* pred0, const float* pred1, const float* pred2, int B, int g0, int size,
                           float anchor_imsize, const float* h_anchors9x2, int mode,
                           long long* best_n, long long* gi, long long* gj, float* boxes, const float* target, float* iou,
                           void* stream);
DCNET_API int dcnet_bbox_iou(const float* b1, const float* b2, int n, int x1y1x2y2, float* iou, void* stream);

/* ---- a19: YOLOLayer COCO-head decode (model/darknet.py:262-296, :365-375) -------------------------------
 * x [B, A*(5+nc), g, g] -> out [B, A*g*g, 5+nc]; anchors A (w,h) pairs (un-scaled, pixels at 416).         */
DCNET_API int dcnet_yolo_layer_decode(const float* x, float* out, int B, int A, int nc, int g, float image_dim,
                                      const float* h_anchorsAx2, void* stream);

/* ---- a21: IoULoss (utils/losses.py:26-34): acc[0] += sum(sig(x) t), acc[1] += sum(sig(x)+t-sig(x) t) -- */
DCNET_API int dcnet_iou_loss_sums(const float* x, const float* t, long long n, float* acc2, void* stream);
/* dx = g[0] * gscale * d(-I/U)/dx; g (device scalar, may be NULL = 1) is the upstream gradient: read on the device, no host sync */
DCNET_API int dcnet_iou_loss_bwd(const float* x, const float* t, long long n, const float* acc2, const float* g, float gscale, float* dx,
                                 void* stream);

/* ---- host: exact emulation of CPython's random.sample() stream (model/DCNet_model.py:87, :413) --------
 * state625: the 625 uint32 words of random.getstate()[1] (624 MT19937 words + position), updated in place so
 * the caller can random.setstate() afterwards.  No device work.                                           */
/* negpos [P*top_k*neg_n]: random.sample(population of N0-1, neg_n) positions, pairs outer, ranks inner */
DCNET_API int dcnet_pyrandom_interframe(uint32_t* h_state625, int P, int top_k, int N0, int neg_n, int* h_negpos);
/* negidx [B*N0*neg_n]: for every (image ii, pixel jj) the reference draws B samples (population N0-1 when
 * index==ii else N0) and keeps the last (index=B-1); returned as pixel indices of image B-1.               */
DCNET_API int dcnet_pyrandom_crossmodal(uint32_t* h_state625, int B, int N0, int neg_n, long long* h_negidx);

#ifdef __cplusplus
}
#endif
#endif  /* DCNET_B200_H_ */
