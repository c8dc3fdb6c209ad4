/* cmx_b200 — C ABI of the B200-native CMX RGB-X segmentation hot path.
 *
 * The reference (ynalcakan/RGBX_Semantic_Segmentation) is pure Python/PyTorch and has NO FFI of its
 * own; the boundary a user sees is `models/builder.py::EncoderDecoder` (builder.py:14-253) and
 * `utils/metric.py::hist_info` (metric.py:8-15).  This header is the thin C ABI *below* that
 * boundary (last row): every entry point replaces the ATen/cuDNN/cuBLAS call(s) that
 * the cited reference line issues.  Conventions:
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless stated; caller owns every
 *     buffer (incl. workspaces); kernels never allocate, never synchronise, and enqueue on `stream`
 *     (a cudaStream_t passed as void*).
 *   - return 0 on success, <0 for argument errors, >0 = cudaError_t; message via cmx_last_error().
 *   - activations are token-major ("NLC" == NHWC): row = b*H*W + h*W + w, channels contiguous,
 *     `ld*` = row stride in elements.  dtype tags: CMX_BF16 = 0, CMX_F32 = 1.
 */
#ifndef CMX_B200_H
#define CMX_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* cmx_last_error(void);
int cmx_version(void);
/* number of kernel launches issued through this library by the calling process (bench `gpu_launches`) */
long long cmx_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = residual + row_scale[row/rows_per_sample] * act(alpha * A.B + bias),  act in {none, ReLU}
 * Replaces every nn.Linear / 1x1 conv / patchified conv / bmm of the path:
 *   dual_segformer.py:68,72 (fc1, fc2) :119-135 (q, sr, kv, q@k^T, attn@v, proj) :219 (patch embed proj)
 *   net_utils.py:79-83, 206-212, 274-279, 325-328 ; MLPDecoder.py:18, 76, 79 ; and their autograd
 *   backward (dgrad / wgrad).
 * A(m,k): trans_a==0 -> A[m*lda + k]  ("K-major"),  trans_a==1 -> A[k*lda + m]  ("MN-major")
 * B(k,n): trans_b==0 -> B[n*ldb + k]  (nn.Linear weight layout [N,K]),  trans_b==1 -> B[k*ldb + n]
 * A, B are bf16.  C dtype = c_dtype; residual dtype = r_dtype.  bias fp32[N] or NULL.
 * batch: two-level (b1,b2) with element strides; batch1=batch2=1 for a plain GEMM.
 * split_k>1 or accumulate!=0: C must be fp32 and is accumulated with atomic adds (caller zeroes it).
 * impl: 0 = auto (tcgen05/TMA kernel when eligible, else generic tensor-core fallback),
 *       1 = force generic fallback, 2 = require tcgen05 (error if not eligible).
 */
typedef struct CmxGemm {
  const void* A;
  const void* B;
  void* C;
  const float* bias;
  const void* residual;
  const float* row_scale;
  int64_t M, N, K;
  int64_t lda, ldb, ldc, ldr;
  int32_t trans_a, trans_b;
  int32_t batch1, batch2;
  int64_t sA1, sA2, sB1, sB2, sC1, sC2;
  int32_t c_dtype, r_dtype;
  int32_t act;
  float alpha;
  int32_t accumulate;
  int32_t split_k;
  int32_t rows_per_sample;
  int32_t impl;
  /* grouped launches (the RGB and X branches of a stage = batch1 index 0 / 1, same shapes, different weights): element
   * strides per batch1 index of bias, residual and row_scale (0 = shared); only with batch2 == 1 */
  int64_t sBias1, sR1, sS1;
} CmxGemm;
int cmx_gemm(const CmxGemm* g, void* stream);
/* debug aid: device buffer (5*64*4 int64) receiving clock64 stamps of CTA 0's producer / MMA / epilogue roles, or NULL */
int cmx_debug_set_gemm_trace(void* buf);
/* which implementation cmx_gemm would pick: 2 = tcgen05, 1 = fallback */
int cmx_gemm_which(const CmxGemm* g);

/* ---- LayerNorm (dual_segformer.py:177-178, 123, 221-223, 382-383; net_utils.py:279-280) ------ */
/* groups > 1 (grouped launch, the RGB and X branch of a stage at once): x, y, mean, rstd hold `groups` stacked blocks of M
 * rows; gamma / beta of group g lie g * param_gs elements behind the given pointers (flat parameter buffer). */
int cmx_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, float eps,
                      void* y, int y_dtype, int64_t ldy, float* mean, float* rstd, int64_t M, int C, int groups,
                      int64_t param_gs, void* stream);
/* dx = dres + LN'(dy + dy2);  dx_bf = bf16(dx * scale[row / rows_per_sample]) (optional);
 * dgamma/dbeta (fp32[C]) are ACCUMULATED (atomic). dy_dtype applies to dy; dy2 is bf16; dres fp32.
 * dbias (optional, fp32[C], accumulated): column sums of the bf16 output (dx_bf if given, else dx) = bias gradient of
 * the Linear/conv layer whose output this LayerNorm normalised.
 * groups > 1: every row tensor holds `groups` stacked blocks of M rows; gamma / dgamma / dbeta / dbias of group g lie
 * g * param_gs elements behind the given pointers, scale g * scale_gs elements. */
int cmx_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* dy2, int64_t lddy2,
                      const void* x, int x_dtype, int64_t ldx, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, int64_t lddres,
                      void* dx, int dx_dtype, int64_t lddx, void* dx_bf, int64_t lddxbf,
                      const float* scale, int rows_per_sample,
                      float* dgamma, float* dbeta, float* dbias, int64_t M, int C, int groups, int64_t param_gs,
                      int64_t scale_gs, void* stream);

/* ---- BatchNorm2d over token-major [M,C] (net_utils.py:318,321; MLPDecoder.py:52-53) ----------- */
int cmx_colstats(const void* x, int x_dtype, int64_t ldx, double* sum, double* sumsq, int64_t M, int C, void* stream);
int cmx_bn_finalize(const double* sum, const double* sumsq, int64_t count, float eps, float momentum,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked,
                    float* mean, float* invstd, int C, void* stream);
int cmx_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean, float* invstd,
                      int C, void* stream);
/* y = mask[b,c] * relu?( (x-mean)*invstd*gamma+beta + residual ) */
int cmx_bn_apply(const void* x, int x_dtype, int64_t ldx, const float* mean, const float* invstd,
                 const float* gamma, const float* beta, const void* residual, int r_dtype, int64_t ldr,
                 int relu, const float* mask, int rows_per_sample,
                 void* y, int y_dtype, int64_t ldy, int64_t M, int C, void* stream);
/* backward of cmx_bn_apply in batch-statistics mode. Pass 1 accumulates sum_dy / sum_dy_xhat (double[C]). */
int cmx_bn_bwd_reduce(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                      const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const void* residual, int r_dtype, int64_t ldr,
                      int relu, const float* mask, int rows_per_sample,
                      double* sum_dy, double* sum_dy_xhat, int64_t M, int C, void* stream);
/* Pass 2: dx (and dres = effective dy, optional); dgamma/dbeta accumulated into fp32[C]. */
int cmx_bn_bwd_apply(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                     const float* mean, const float* invstd, const float* gamma, const float* beta,
                     const void* residual, int r_dtype, int64_t ldr,
                     int relu, const float* mask, int rows_per_sample,
                     const double* sum_dy, const double* sum_dy_xhat,
                     void* dx, int dx_dtype, int64_t lddx, void* dres, int dres_dtype, int64_t lddres,
                     float* dgamma, float* dbeta, int64_t M, int C, void* stream);

/* ---- depthwise 3x3 (pad 1) + bias + activation on NHWC bf16 (dual_segformer.py:25-33,69-70;
 *      net_utils.py:314-315).  w: fp32 [C,9] (== Conv2d weight [C,1,3,3]).  flip: correlate with the
 *      180-degree rotated kernel (data-gradient).
 *      groups > 1 (grouped launch): x / y / dy / du hold `groups` stacked blocks of B samples; the parameters (w, bias,
 *      ysum, dw, db) of group g lie g * param_gs ELEMENTS behind the given pointers (flat parameter / gradient buffer). */
int cmx_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w, const float* bias, int act, int flip,
                      void* y, int64_t ldy, float* ysum, int B, int H, int W, int C, int groups, int64_t param_gs, void* stream);
/* ysum (optional, fp32[C], accumulated): per-channel sums of the output (bias gradient of the producer layer when
 * this call computes a data gradient) */
/* du = dy * act'(conv(x)+b) -> bf16; dW[C,9], db[C] accumulated (fp32 atomics). */
int cmx_dwconv3x3_bwd_pre(const void* x, int64_t ldx, const float* w, const float* bias, int act,
                          const void* dy, int64_t lddy, void* du, int64_t lddu, float* dw, float* db,
                          int B, int H, int W, int C, int groups, int64_t param_gs, void* stream);

/* ---- layout movers ---------------------------------------------------------------------------- */
/* stage-1 OverlapPatchEmbed input: NCHW fp32 image -> bf16 im2col rows [B*Ho*Wo, kpad],
 * column = (kh*k+kw)*Cin+ci (dual_segformer.py:219, conv k=7 s=4 p=3) */
int cmx_im2col_nchw(const float* x, void* col, int B, int Cin, int H, int W, int k, int s, int p,
                    int Ho, int Wo, int kpad, void* stream);
/* the same gather straight from the RAW uint8 image [B,H,W,ch] (ch = 3 HWC, or 1 = grey X): the reference's host-side input
 * pipeline - float64 normalisation ((v / 255) - mean[c]) / std[c] rounded to fp32 (utils/transforms.py:182-187), thermal 1 -> 3
 * replication (RGBXDataset.py:57-59), HWC -> CHW (dataloader.py:85-112) - fused into the stage-1 patch-embed load.  ch = 1 (grey
 * X): the replicated channels are affine images of one value, so TWO columns per tap are written, (v / 255, 1-inside-the-image);
 * the caller multiplies by the 7x7x3 weights folded to 7x7x2 (sum_c W_c / std_c, -sum_c W_c mean_c / std_c). */
int cmx_im2col_u8(const uint8_t* x, void* col, int B, int ch, int H, int W, int k, int s, int p, int Ho, int Wo, int kpad,
                  double mean0, double mean1, double mean2, double std0, double std1, double std2, void* stream);
/* NHWC bf16 -> im2col rows [B*Ho*Wo, k*k*C] (stage 2-4 patch embeds k=3 s=2 p=1; SR conv k=s=R p=0) */
int cmx_im2col_nhwc(const void* x, int64_t ldx, void* col, int B, int H, int W, int C, int k, int s, int p,
                    int Ho, int Wo, void* stream);
/* adjoint of cmx_im2col_nhwc (gather form): dx[b,y,x,:] = add + sum of dcol entries */
int cmx_col2im_nhwc(const void* dcol, const void* add, int add_dtype, int64_t ldadd, void* dx, int dx_dtype,
                    int64_t lddx, int B, int H, int W, int C, int k, int s, int p, int Ho, int Wo, void* stream);
/* Conv2d weight [Co,Ci,kh,kw] fp32 -> bf16 [Co,kpad] with column (kh*kw_+kw)*Ci+ci, and the adjoint
 * (fp32 [Co,kpad] grad -> ACCUMULATED into [Co,Ci,kh,kw]) */
int cmx_convw_pack(const float* w, void* wp, int Co, int Ci, int kh, int kw, int kpad, void* stream);
int cmx_convw_unpack_grad(const float* gp, float* gw, int Co, int Ci, int kh, int kw, int kpad, void* stream);
/* the same two operations for ALL real convolutions of the model in one launch each (start of the step / end of the
 * backward pass).  descs: DEVICE array of n descriptors; gp/gw are only read by the unpack, w/wp only by the pack. */
typedef struct CmxConvDesc {
  const float* w;   /* [Co, Ci, kh, kw] fp32 master weight */
  void* wp;         /* [Co, kpad] bf16 packed (kh, kw, ci)-major operand */
  const float* gp;  /* [Co, kpad] fp32 packed gradient (accumulated by the wgrad GEMM) */
  float* gw;        /* [Co, Ci, kh, kw] fp32 gradient, += */
  int32_t Co, Ci, kh, kw, kpad, reserved;
} CmxConvDesc;
int cmx_convw_pack_multi(const CmxConvDesc* descs, int n, void* stream);
int cmx_convw_unpack_grad_multi(const CmxConvDesc* descs, int n, void* stream);
/* fp32 -> bf16 cast (weights), optionally many at once is done by the caller on a flat buffer */
int cmx_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
int cmx_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream);
/* column sums of a [M,N] matrix (bias gradients): out[n] += sum_m x[m,n] */
/* groups > 1: x = `groups` stacked blocks of M rows; out of group g lies g * out_gs elements behind the given pointer */
int cmx_colsum(const void* x, int x_dtype, int64_t ldx, float* out, int64_t M, int N, int groups, int64_t out_gs, void* stream);
/* dy *= (y > 0)   (ReLU backward, in place on bf16) */
int cmx_relu_bwd(void* dy, int64_t lddy, const void* y, int64_t ldy, int64_t M, int N, void* stream);
/* out = a*x + b*y elementwise on fp32 flat buffers (grad scaling) */
int cmx_axpby_f32(float a, const float* x, float b, const float* y, float* out, int64_t n, void* stream);
/* AdamW (torch.optim.AdamW semantics: decoupled decay, bias correction, no amsgrad) over the engine's flat fp32
 * parameter storage in ONE launch - the optimizer step of train.py:95-100,176-178.  p/g/m/v: n fp32 elements (n a
 * multiple of 64); block_group[n/64] (DEVICE, uint8): parameter group of each 64-element block, 255 = leave untouched;
 * lr/wd: HOST arrays of ngroups (<= 8) values; step counts from 1; grads are multiplied by grad_scale first;
 * w_bf16 (optional, device): receives the bf16 copy of the updated parameters. */
int cmx_adamw_flat(float* p, const float* g, float* m, float* v, void* w_bf16, const uint8_t* block_group, int64_t n,
                   const float* lr, const float* wd, int ngroups, float beta1, float beta2, float eps, float grad_scale,
                   int64_t step, void* stream);

/* ---- fused spatial-reduction self-attention forward (dual_segformer.py:127-134), head_dim 64, Nkv <= 320 ----------
 * O[b,n,h,:] = softmax_k(scale * q[b,n,h,:].k[b,k,h,:]) v[b,k,h,:]   — flash style (scores only in tensor memory).
 * q [B*N, ldq] (head h at columns h*64..), kv [B*Nk, ldkv] (K at h*64, V at heads*64 + h*64), o [B*N, ldo], all bf16.
 * p_out (optional, bf16 [B*heads*N, ldp]): the normalised probabilities, written by TMA for the backward pass.
 * lse (optional, fp32 [B*heads*N]): natural-log row normaliser. */
int cmx_attn_fwd(const void* q, int64_t ldq, const void* kv, int64_t ldkv, void* o, int64_t ldo, void* p_out, int64_t ldp,
                 float* lse, int B, int N, int Nk, int heads, float scale, int64_t kv_sample_rows, void* stream);
/* kv_sample_rows (0 = Nk): key rows per sample in kv.  Larger than Nk when the call covers ONE CHUNK of a longer key axis (kv
 * then points at the chunk's first key row of sample 0): key-chunked attention for Nkv above the resident-K/V limit, combined by
 * cmx_attn_combine:  o = sum_c exp(lse_c - lse) o_c,  lse = log sum_c exp(lse_c);  o_parts / lse_parts hold nc stacked parts
 * (part_stride / lse_stride elements apart). */
int cmx_attn_combine(const void* o_parts, int64_t part_stride, int64_t ldp, const float* lse_parts, int64_t lse_stride, int nc,
                     void* o, int64_t ldo, float* lse, int B, int N, int heads, void* stream);
/* out[n] (bf16) = sum over nc stacked bf16 tensors, fp32 accumulation (the per-chunk dQ partials of the chunked backward) */
int cmx_sum_parts_bf16(const void* parts, int64_t part_stride, int nc, void* out, int64_t n, void* stream);

/* fused attention backward core: dP = dO V^T (tensor memory only), dS = scale * P .* (dP - rowsum(P .* dP)) -> ds_out
 * (bf16, same layout as p; consumed by the split-K dK = dS^T Q GEMM), dQ = dS K -> dq [B*N, lddq].  dV = P^T dO and
 * dK stay ordinary batched split-K cmx_gemm calls (they contract over all query tokens). */
int cmx_attn_bwd(const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const void* p, int64_t ldp, void* ds_out,
                 int64_t ldds, void* dq, int64_t lddq, int B, int N, int Nk, int heads, float scale, void* stream);

/* EXPERIMENTAL (kernel-level parity green on B200, 9 shapes; the engine calls it only under CMX_ATTN_DKV_RECOMPUTE=1 until the
 * model-level tests and the bench A/B under that flag have run):
 * key-major dK / dV of the same attention with the probabilities recomputed from q, k and the forward's lse instead of
 * read back from HBM.  delta [B*heads*N] = rowsum(dO .* O) from cmx_attn_delta; dkv_acc: zero-initialised fp32
 * [B*Nk, lddkv] (dK at columns h*64.., dV at heads*64 + h*64.., the layout of the kv projection output), added to
 * with fp32 RED.  Replaces the two batched split-K cmx_gemm calls dV = P^T dO and dK = dS^T Q. */
int cmx_attn_delta(const void* d_o, int64_t lddo, const void* o, int64_t ldo, float* delta, int B, int N, int heads,
                   void* stream);
int cmx_attn_dkv(const void* q, int64_t ldq, const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const float* lse,
                 const float* delta, float* dkv_acc, int64_t lddkv, int B, int N, int Nk, int heads, float scale,
                 void* stream);
/* EXPERIMENTAL companion (same status): query-major dQ [B*N, lddq] (bf16) with recomputed probabilities, Nkv <= 384; with
 * cmx_attn_dkv it makes the stored probabilities (p_out of cmx_attn_fwd) and the dS round trip of cmx_attn_bwd unnecessary. */
int cmx_attn_dq(const void* q, int64_t ldq, const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const float* lse,
                const float* delta, void* dq, int64_t lddq, int B, int N, int Nk, int heads, float scale,
                int64_t kv_sample_rows, void* stream);

/* ---- softmax ---------------------------------------------------------------------------------- */
/* row softmax of fp32 S [rows, n] (ld) -> bf16 P  (dual_segformer.py:131) and its backward
 * dS = scale * P .* (dP - rowsum(P.*dP)) -> bf16 */
int cmx_softmax_rows_fwd(const float* s, int64_t lds, void* p, int64_t ldp, int64_t rows, int n, void* stream);
int cmx_softmax_rows_bwd(const void* p, int64_t ldp, const float* dp, int64_t lddp, float scale,
                         void* ds, int64_t ldds, int64_t rows, int n, void* stream);
/* FFM context softmax over dim -2 of [nb, d, d] fp32 (net_utils.py:207,209): P = softmax_rows-index(scale*C)
 * writes fp32 P and bf16 P^T-free copy; backward dC = scale * P .* (dP - colsum(P.*dP)) */
int cmx_softmax_dim2_fwd(const float* c, float scale, float* p32, void* p16, int nb, int d, void* stream);
int cmx_softmax_dim2_bwd(const float* p32, const float* dp, float scale, void* dc16, int nb, int d, void* stream);

/* ---- FRM (net_utils.py:22-30, 79-83, 147-152) ------------------------------------------------- */
/* global avg + max pool per (b,c) of x1|x2 packed as one [B*HW, C2] matrix (C2 = 2C). y[b] = [avg(C2) | max(C2)],
 * argmax (row index within the sample) saved for backward. */
int64_t cmx_pool_avgmax_ws_bytes(int B, int C2);   /* size of the partial-result workspace `ws` */
int cmx_pool_avgmax_fwd(const void* x, int64_t ldx, float* y, int32_t* argmax, void* ws, int B, int HW, int C2, void* stream);
/* dx[row,c] += dy_avg[b,c]/HW + (row==argmax[b,c]) * dy_max[b,c]   (dx fp32, in place) */
int cmx_pool_avgmax_bwd(const float* dy, const int32_t* argmax, float* dx, int64_t lddx, int B, int HW, int C2, void* stream);
/* small-M fp32 linear: y = act(x W^T + b), x [Mb,K], W [N,K]; act: 0 none, 1 relu, 3 sigmoid */
int cmx_smallm_linear_fwd(const float* x, const float* w, const float* b, int act, float* y, int Mb, int N, int K, void* stream);
/* backward: given dy (grad wrt y) and y: dpre = dy*act'(y); dx = dpre W ; dW += dpre^T x ; db += colsum(dpre) */
int cmx_smallm_linear_bwd(const float* dy, const float* y, int act, const float* x, const float* w,
                          float* dx, float* dw, float* db, float* dpre_ws, int Mb, int N, int K, void* stream);
/* sw = sigmoid(t W2^T + b2) ([M,2], saved);  r1 = a1 + .5*(cw1[b,:]+sw1)*a2 ; r2 = a2 + .5*(cw0[b,:]+sw0)*a1
 * a = [a1|a2] bf16 [M,2C] ; t bf16 [M,C] ; cw fp32 [B,2C] ; w2 fp32 [2,C] */
int cmx_frm_rectify_fwd(const void* a, int64_t lda, const void* t, int64_t ldt, const float* w2, const float* b2,
                        const float* cw, float* sw, void* r1, int64_t ldr1, void* r2, int64_t ldr2,
                        int B, int HW, int C, void* stream);
/* backward: dr1, dr2 fp32 [M,C] -> da fp32 [M,2C] (written), dt bf16 [M,C] (relu-masked by t>0),
 * dcw fp32[B,2C], dw2 fp32[2,C], db2 fp32[2] ACCUMULATED */
int cmx_frm_rectify_bwd(const float* dr1, int64_t lddr1, const float* dr2, int64_t lddr2,
                        const void* a, int64_t lda, const void* t, int64_t ldt, const float* w2,
                        const float* cw, const float* sw, float* da, int64_t ldda, void* dt, int64_t lddt,
                        float* dcw, float* dw2, float* db2, int B, int HW, int C, void* stream);

/* ---- decoder (MLPDecoder.py:66-77) ------------------------------------------------------------- */
/* out[b,y,x,:] (fp32) = bias + z0[b,y,x,:] + sum_i bilinear(z_i)[b,y,x,:]  (align_corners=False) */
int cmx_upsample_sum_fwd(const void* z0, const void* z1, const void* z2, const void* z3,
                         int H0, int W0, int H1, int W1, int H2, int W2, int H3, int W3,
                         const float* bias, void* out, int out_dtype, int B, int C, void* stream);
/* adjoint for ONE source: dz[b,yi,xi,:] = sum_{y,x} w(y,x;yi,xi) * dout[b,y,x,:]  (bf16 in, bf16 out) */
int cmx_upsample_bwd(const void* dout, int Ho, int Wo, void* dz, int Hi, int Wi, int B, int C, void* stream);
/* the same adjoint for up to three sources (dz2 / dz3 may be NULL) in one pass over dout; C % 64 == 0
 * (MLPDecoder.py:66-77 backward: the upsampled c2..c4 branches share the gradient of the fused map) */
int cmx_upsample_bwd_multi(const void* dout, int Ho, int Wo, void* dz1, int H1, int W1, void* dz2, int H2, int W2,
                           void* dz3, int H3, int W3, int B, int C, void* stream);

/* ---- loss / logits / metric (builder.py:233,249; evaluator.py:393; utils/metric.py:8-15) ------- */
/* low-res logits [B,h,w,ncls] fp32 (channels-last, pixel stride ld >= ncls elements: the engine pads the class axis to a
 * multiple of 8 so the prediction layer runs on the tcgen05 path) -> bilinear x to [H,W] -> CE(ignore).
 * acc[0] += sum of -log p (double), acc[1] += valid count (double). If dlogits != NULL also accumulates the
 * UNNORMALISED gradient sum_{pixels} w * (softmax - onehot) into dlogits (fp32, same layout and stride as logits). */
int cmx_ce_upsampled_fwd_bwd(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc,
                             float* dlogits, int B, int h, int w, int H, int W, int ncls, void* stream);
/* same pass for the reference's alternate criteria (train.py:70-93): w_ce * CE + w_focal * FocalLoss(gamma, alpha) with the
 * all-classes focal form of utils/loss_opr.py:157-196 ('FocalLoss': w_ce 0, w_focal 1; 'CE_Focal': 1 and 0.2, builder.py:246-247);
 * acc[0] += weighted loss sum, acc[1] += valid count, dlogits accumulates the weighted unnormalised gradient. */
int cmx_ce_focal_upsampled_fwd_bwd(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc,
                                   float* dlogits, int B, int h, int w, int H, int W, int ncls, float w_ce, float w_focal,
                                   float gamma, float alpha, void* stream);
/* DiceCELoss (utils/loss_opr.py:103-156; train.py:79-80): alpha * (1 - mean_{b,k} dice_bk) + (1 - alpha) * CE on the bilinearly
 * upsampled logits, two passes over the pixels.  cmx_dice_ce_stats: acc (double[2], zeroed) += (CE sum, valid count), dstats
 * (double[B*3*ncls], zeroed) += per (sample, class) (sum p, sum p*onehot, sum onehot) over the valid pixels.
 * cmx_dice_ce_finalize: loss (optional) and coef (optional, float[B*2*ncls + 1]) from the sums.  cmx_dice_ce_grad: dlogits
 * (fp32, zeroed, same layout as logits) += d loss / d logits (final scale; no further normalisation). */
int cmx_dice_ce_stats(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc, double* dstats,
                      int B, int h, int w, int H, int W, int ncls, void* stream);
int cmx_dice_ce_finalize(const double* acc, const double* dstats, int B, int ncls, float alpha, float smooth, float* loss,
                         float* coef, void* stream);
int cmx_dice_ce_grad(const float* logits, int64_t ld, const int64_t* label, int ignore_index, const float* coef,
                     float* dlogits, int B, int h, int w, int H, int W, int ncls, void* stream);
/* loss = acc[0]/acc[1];  dlogits_out(bf16/f32) = dlogits * gscale/acc[1] */
int cmx_ce_finalize(const double* acc, float* loss, const float* dlogits, const float* gscale,
                    void* dlogits_out, int out_dtype, int64_t n, void* stream);
/* low-res channels-last logits -> full-res NCHW fp32 logits (eval output of EncoderDecoder.forward) */
int cmx_logits_upsample_nchw(const float* logits, int64_t ld, float* out, int B, int h, int w, int H, int W, int ncls,
                             void* stream);
/* confusion matrix: hist[n_cl*gt+pred] += 1 for 0<=gt<n_cl (int64), stats[0]=labeled, stats[1]=correct.
 * pred_dtype/gt_dtype: 0 = uint8, 1 = int32, 2 = int64.  */
int cmx_confusion(const void* pred, int pred_dtype, const void* gt, int gt_dtype, int64_t n, int n_cl,
                  int64_t* hist, int64_t* stats, void* stream);
/* fused argmax over channel of NCHW scores [ncls,H*W] (one image; fp32, or fp64 when score_f64 != 0) + confusion;
 * pred_out optional (uint8); gt optional (argmax only) */
int cmx_argmax_confusion(const void* scores, int score_f64, const void* gt, int gt_dtype, int64_t npix, int n_cl,
                         uint8_t* pred_out, int64_t* hist, int64_t* stats, void* stream);

/* ---- sliding-window / multi-scale evaluation driver on the device (engine/evaluator.py:306-431) -------------------------
 * cmx_eval_pack_crop: one network input crop [out_ch, crop_h, crop_w] fp32 (CHW) cut from a uint8 HWC (ch = 3) or HW (ch = 1)
 * image resident on the device.  Restates evaluator.py:337-360 + 398-431 + utils/transforms.py:61-75,182-187:
 *   crop pixel (y, x) -> window pixel (y - out_top, x - out_left); outside the win_h x win_w window the output is 0.0 (the zero
 *   padding process_image_rgbX applies AFTER normalisation); inside, the canvas pixel (s_y + ., s_x + .) of the zero-padded RAW
 *   image (pad_top / pad_left black pixels before the image) is normalised in float64, ((v / 255) - mean[c]) / std[c], and
 *   rounded to fp32 like the reference's astype(float32).  flip != 0 mirrors the crop horizontally (flip-TTA input). */
int cmx_eval_pack_crop(const uint8_t* img, int rows, int cols, int ch, int pad_top, int pad_left, int s_y, int s_x,
                       int win_h, int win_w, int out_top, int out_left, double mean0, double mean1, double mean2,
                       double std0, double std1, double std2, int flip, float* out, int crop_h, int crop_w, void* stream);
/* one tile of a scale's score canvas: crop index in the logits batch, canvas window [s_y, e_y) x [s_x, e_x) and the
 * tile-internal margins (tm_top, tm_left) that process_image_rgbX added */
typedef struct CmxEvalTile {
  int32_t crop, s_y, s_x, e_y, e_x, tm_top, tm_left, pad_;
} CmxEvalTile;
/* cmx_eval_accumulate_scale: processed[ncls, ori_rows, ori_cols] (fp64) += resize_bilinear(score)[...] for ONE scale, where
 *   score(c, y, x) = sum over the tiles t covering canvas pixel (y + m_top, x + m_left), in table order and in fp32, of
 *   exp(logits[t.crop, c, ...] (+ logits_flip[t.crop, c, ..., mirrored] when logits_flip != NULL))   (evaluator.py:361-393)
 *   is the rows x cols score map of the scale (canvas minus the margins) and the resize is cv2.INTER_LINEAR's sampling
 *   (half-pixel centres, edge clamp) evaluated in fp32, horizontal then vertical (evaluator.py:318-320).
 * logits / logits_flip: fp32 [n_crops, ncls, crop_h, crop_w] (the model's NCHW output); tiles: device table of n_tiles. */
int cmx_eval_accumulate_scale(const float* logits, const float* logits_flip, int crop_h, int crop_w, int ncls,
                              const CmxEvalTile* tiles, int n_tiles, int m_top, int m_left, int rows, int cols,
                              double* processed, int ori_rows, int ori_cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif
