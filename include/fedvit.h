/*
 * fedvit.h — C ABI of libfedvit.so (B200 / sm_100a).
 *
 * The reference (apurbaaaa/Federated-Vit-Skin-Lesion-Classification) is pure Python and has no
 * FFI of its own: its seam for this path is the set of PyTorch calls made by
 *   model.py:112-117,178-207   (timm ViT backbone + head forward)
 *   losses.py:41-67            (AsymmetricFocalLoss.forward)
 *   train.py:144-162           (autocast fwd, backward, clip, AdamW step, EMA)
 *   utils.py:76-83,192-193     (EMA.update, clip_grad_norm)
 * Every entry point below replaces one of those library calls (the reference line it stands in
 * for is cited per function). INTEGRATION.md shows the ctypes / torch.library binding.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, enums; no torch / C++ types in any signature.
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is allocated here.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises, and is CUDA-graph capturable.
 *   - return value: 0 on success, a negative FV_ERR_* code otherwise; fv_last_error() returns a
 *     thread-local message. Nothing is thrown across the boundary.
 *   - all matrices are row-major and densely packed unless a leading dimension is given.
 */
#ifndef FEDVIT_H_
#define FEDVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FV_OK 0
#define FV_ERR_INVALID_ARG (-1)
#define FV_ERR_CUDA (-2)
#define FV_ERR_UNSUPPORTED (-3)

/* element types */
#define FV_F32 0
#define FV_BF16 1

/* GEMM epilogues (see fv_gemm_bf16 / fv_gemm_f32) */
#define FV_EPI_NONE 0          /* C = acc (+bias)                                         */
#define FV_EPI_RESIDUAL 1      /* C = acc (+bias) + R            R fp32, same shape as C   */
#define FV_EPI_GELU 2          /* u = acc+bias; C = gelu_erf(u); AUX = gelu_erf'(u) (for bwd)*/
#define FV_EPI_DGELU 3         /* C = acc * AUX                  AUX = gelu_erf'(u) from fwd */
#define FV_EPI_ACCUM 4         /* C (fp32) += acc                 weight-gradient accumulate*/
#define FV_EPI_PATCH 5         /* C[row + row/tokens_per_img + 1] = acc + bias + pos[...]   */

/* operand storage for fv_gemm_*:  op(A) is M x K, op(B) is N x K (both "K-major" when 0).   */
#define FV_MAJOR_K 0           /* stored [rows, K]   (K contiguous)                        */
#define FV_MAJOR_MN 1          /* stored [K, rows]   (rows contiguous)                     */

int fv_version(void);
const char* fv_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t fv_launch_count(void);
/* launches per kernel family since load — lets a test assert WHICH kernel served a call (e.g. that a
 * ViT-B step really ran on the CTA-pair GEMM) without a profiler attached */
#define FV_KERNEL_GEMM_TC 0        /* single-CTA tcgen05 GEMM                     */
#define FV_KERNEL_GEMM_TC_PAIR 1   /* CTA-pair (cta_group::2) tcgen05 GEMM        */
#define FV_KERNEL_ATTN_FWD 2       /* tcgen05 attention forward, N <= 256         */
#define FV_KERNEL_ATTN_BWD 3       /* tcgen05 attention backward, N <= 256        */
#define FV_KERNEL_ATTN_FWD_LONG 4  /* tcgen05 attention forward, 256 < N <= 768   */
#define FV_KERNEL_ATTN_BWD_LONG 5  /* tcgen05 attention backward, 256 < N <= 768  */
#define FV_KERNEL_ATTN_LEGACY 6    /* mma.sync flash kernels (N > 768)            */
#define FV_KERNEL_FAMILIES 8
int64_t fv_kernel_launches(int family);

/* ------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = op(A)[M,K] * op(B)[N,K]^T  with a fused epilogue.
 * Replaces the cuBLAS calls behind timm's nn.Linear layers — Attention.qkv/.proj, Mlp.fc1/.fc2
 * (reached from model.py:193) — and their autograd backward (train.py:153).
 *
 * fv_gemm_bf16: bf16 operands, fp32 accumulate in TMEM (tcgen05.mma, TMA-fed).
 * fv_gemm_f32 : fp32 operands, fp32 FFMA (the 1e-4 parity path).
 *   a_major/b_major : FV_MAJOR_K or FV_MAJOR_MN (how the operand is laid out in memory)
 *   lda/ldb         : leading dimension in elements of the stored matrix
 *   bias            : fp32 [N] or NULL
 *   c, c_dtype, ldc : output (FV_F32 or FV_BF16)
 *   aux             : FV_EPI_RESIDUAL -> fp32 residual [M,ldaux]; FV_EPI_GELU -> second output, the
 *                     GELU derivative at the pre-activation (c_dtype), or NULL when only the activation
 *                     is wanted (forward-only / eval); FV_EPI_DGELU -> that tensor;
 *                     FV_EPI_PATCH -> fp32 pos_embed [(tokens_per_img+1), N]
 *   split_k         : >1 only with FV_EPI_ACCUM: K is cut in split_k slices, each atomically
 *                     accumulated into fp32 C.
 *   batch / strides : fv_gemm_f32 only (strided-batched, used by the fp32 attention path).
 * ---------------------------------------------------------------------------------------- */
int fv_gemm_bf16(const void* a, int a_major, int64_t lda,
                 const void* b, int b_major, int64_t ldb,
                 const float* bias,
                 void* c, int c_dtype, int64_t ldc,
                 void* aux, int64_t ldaux,
                 int64_t m, int64_t n, int64_t k,
                 int epilogue, int split_k, int tokens_per_img, void* stream);

/* Weight gradient of y = x W^T + b in one kernel: dW[out,in] += dY^T X (split-K, fp32 atomics) and
 * dbias[out] += column sums of dY — the epilogue warps add up the dY tiles while the tensor core
 * consumes them, so the bias gradient costs no extra pass over HBM. dbias may be NULL.
 *   dy [tokens, out] bf16 (lddy), x [tokens, in] bf16 (ldx), dw fp32 [out, in] (lddw).            */
int fv_wgrad_bf16(const void* dy, int64_t lddy, const void* x, int64_t ldx, float* dw, int64_t lddw,
                  float* dbias, int64_t tokens, int64_t out_features, int64_t in_features,
                  int split_k, void* stream);

/* Branch + residual of a transformer block with per-sample stochastic depth (timm DropPath):
 *   out[r,:] = residual[r,:] + row_scale[r / rows_per_scale] * (a[r,:] w^T + bias)
 * row_scale (fp32, one value per image: keep_mask / keep_prob) may be NULL (plain residual).
 * bf16: a [m,k], w [n,k] bf16; f32: fp32 operands. residual / out fp32.                          */
int fv_linear_residual_bf16(const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias,
                            const float* residual, int64_t ldr, const float* row_scale,
                            int64_t rows_per_scale, float* out, int64_t ldo, int64_t m, int64_t n,
                            int64_t k, void* stream);
int fv_linear_residual_f32(const float* a, int64_t lda, const float* w, int64_t ldw, const float* bias,
                           const float* residual, int64_t ldr, const float* row_scale,
                           int64_t rows_per_scale, float* out, int64_t ldo, int64_t m, int64_t n,
                           int64_t k, void* stream);

int fv_gemm_f32(const float* a, int64_t a_row_stride, int64_t a_col_stride, int64_t a_batch_stride,
                const float* b, int64_t b_row_stride, int64_t b_col_stride, int64_t b_batch_stride,
                const float* bias,
                float* c, int64_t ldc, int64_t c_batch_stride,
                float* aux, int64_t ldaux,
                int64_t m, int64_t n, int64_t k, int64_t batch,
                float alpha, int epilogue, int tokens_per_img, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (eps inside the sqrt, affine).
 * Replaces timm Block.norm1/norm2 and VisionTransformer.norm (F.layer_norm, model.py:193).
 *   x fp32 [rows, cols]; y is y_dtype; mean/rstd fp32 [rows] are saved for the backward.
 * Backward: dx = dres + LN'(dy) in fp32, optional bf16 copy of dx (dx_lp) for the next GEMM;
 *   dgamma/dbeta are ACCUMULATED (+=) into fp32 [cols] (zero them first for a fresh gradient).
 *   lp_row_scale (optional, fp32, one value per rows_per_scale rows): the bf16 copy is written as
 *   dx * scale — the per-sample stochastic-depth factor of the sub-layer that consumes it.
 * ---------------------------------------------------------------------------------------- */
int fv_layernorm_fwd(const float* x, const float* gamma, const float* beta,
                     void* y, int y_dtype, float* mean, float* rstd,
                     int64_t rows, int64_t cols, float eps, void* stream);
int fv_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* gamma,
                     const float* mean, const float* rstd, const float* dres,
                     float* dx, void* dx_lp, float* dgamma, float* dbeta,
                     int64_t rows, int64_t cols,
                     const float* lp_row_scale, int64_t rows_per_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-head self-attention core, softmax(Q K^T * scale) V, no mask, no dropout.
 * Replaces F.scaled_dot_product_attention inside timm Attention.forward (model.py:193).
 *   qkv : [B, N, 3, H, 64] (the un-permuted output of Attention.qkv), dtype bf16 or fp32
 *   out : [B, N, H, 64] token-major (what Attention.proj consumes)
 *   lse : fp32 [B, H, N] log-sum-exp of the scaled scores, saved for the backward
 * Backward writes dqkv in the same [B, N, 3, H, 64] layout.
 *   delta: fp32 [B, H, N] scratch (rowsum(dO * O)), caller-owned.
 *   workspace: caller-owned scratch of fv_attention_bwd_workspace(batch, tokens, heads) bytes, 128-byte
 *     aligned (0 bytes / NULL for tokens <= 256 and > 768; the long-sequence tcgen05 path accumulates
 *     the query gradient there in fp32 with TMA reduce-adds before converting it into dqkv).
 * head_dim is 64 for every ViT the path covers (Tiny/Base/Large).
 * ---------------------------------------------------------------------------------------- */
int fv_attention_fwd(const void* qkv, void* out, float* lse, int dtype,
                     int64_t batch, int64_t tokens, int64_t heads, float scale, void* stream);
int64_t fv_attention_bwd_workspace(int64_t batch, int64_t tokens, int64_t heads);
int fv_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                     float* delta, void* dqkv, int dtype, int64_t batch, int64_t tokens,
                     int64_t heads, float scale, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Patch embedding helpers (timm PatchEmbed.proj = Conv2d(C,D,16,16) + cls/pos, model.py:193).
 *   fv_patchify: NCHW fp32 image -> [B*(H/16)*(W/16), C*256] rows in (c,py,px) order, bf16/fp32
 *   fv_cls_pos_rows: writes row 0 of every image: x[b,0,:] = cls + pos[0]
 * ---------------------------------------------------------------------------------------- */
/* im2col-free patch embedding (timm PatchEmbed.proj + cls/pos prologue, rows 1..N-1):
 *   x[b, 1 + t, :] = patch(b, t) . weight^T + bias + pos[1 + t, :]
 * The A operand is never materialised: a 5-D TMA tensor map (px, py, pw, ph, b*c) over the NCHW
 * fp32 image delivers [patches x 32] K-slices straight into 128B-swizzled smem in token order;
 * tf32 tcgen05.mma, fp32 accumulate. weight is the conv weight [dim, chans, 16, 16] (fp32), pos the
 * fp32 [N, dim] position table, x the fp32 [B*N, dim] residual stream (row 0 of each image is
 * written by fv_cls_pos_rows).                                                                  */
int fv_patch_embed_tf32(const float* img, const float* weight, const float* bias, const float* pos,
                        float* x, int64_t batch, int64_t chans, int64_t height, int64_t width,
                        int64_t dim, void* stream);
int fv_patchify(const float* img, void* out, int out_dtype,
                int64_t batch, int64_t chans, int64_t height, int64_t width, void* stream);
/* the same with `lead_rows` zero rows in front of every image's patches: [B*(lead_rows + (H/16)*(W/16)), C*256].
 * With lead_rows = 1 the rows line up with the token rows [B*N, D] of the residual stream (row 0 = the cls slot),
 * so the patch-embedding weight gradient contracts the stream's gradient as it lies in memory — the cls rows meet
 * zeros — instead of a copy of its patch rows (the backward of timm PatchEmbed.proj, model.py:193). */
int fv_patchify_rows(const float* img, void* out, int out_dtype,
                     int64_t batch, int64_t chans, int64_t height, int64_t width, int64_t lead_rows, void* stream);
int fv_cls_pos_rows(const float* cls, const float* pos, float* x,
                    int64_t batch, int64_t tokens, int64_t dim, void* stream);

/* Head of the backbone backward (the head reads timm's global_pool='token' only, model.py:193):
 *   dx[b*tokens + t, :] = t == 0 ? dcls[b, :] : 0                    (fp32, may be NULL)
 *   dy[b*tokens + t, :] = t == 0 ? dcls[b, :] * row_scale[b] : 0     (dy_dtype, may be NULL)
 * row_scale = the last block's per-sample stochastic-depth factor (NULL: 1).                   */
int fv_cls_grad_rows(const float* dcls, const float* row_scale, float* dx, void* dy, int dy_dtype,
                     int64_t batch, int64_t tokens, int64_t dim, void* stream);

/* column sums: out[n] (+)= sum_m a[m,n]   (bias gradients, pos_embed / cls_token gradients)   */
int fv_colsum(const void* a, int a_dtype, int64_t lda, float* out, int accumulate,
              int64_t rows, int64_t cols, void* stream);

/* ------------------------------------------------------------------------------------------
 * Device-side batch assembly (scope row f3).
 * fv_assemble_batch replaces the per-sample tensor work of the reference Dataset
 * (data.py:148-155 TF.to_tensor + TF.normalize + (mask-0.5)/0.5, data.py:222-224 torch.cat to 4
 * channels) for a whole batch that crossed PCIe as bytes, optionally mixing in the same pass.
 * fv_mix_batch replaces MixUp.__call__ / CutMix.__call__ (utils.py:112-150) on an fp32 batch.
 *   img   uint8, nhwc ? [B,H,W,3] : [B,3,H,W];  mask uint8 [B,H,W] or NULL (-> 3 channels)
 *   mean / stdv  HOST pointers to 3 floats (IMAGENET_MEAN / IMAGENET_STD in the reference)
 *   perm  int64 [B] device: partner sample of each sample (torch.randperm), NULL when mode == 0
 *   mode  0 none, 1 mixup: out = lam*x + one_minus_lam*x[perm] (two rounded products + rounded sum,
 *         bit-identical to torch), 2 cutmix: rows [x1,x2) x columns [y1,y2) taken from x[perm]
 *   out   fp32 [B, 3|4, H, W];  W % 4 == 0
 * ---------------------------------------------------------------------------------------- */
int fv_assemble_batch(const uint8_t* img, int nhwc, const uint8_t* mask, const float* mean, const float* stdv,
                      const int64_t* perm, float lam, float one_minus_lam, int mode, int x1, int y1, int x2, int y2,
                      float* out, int64_t batch, int64_t height, int64_t width, void* stream);
int fv_mix_batch(const float* x, const int64_t* perm, float lam, float one_minus_lam, int mode, int x1, int y1,
                 int x2, int y2, float* out, int64_t batch, int64_t chans, int64_t height, int64_t width,
                 void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses. fv_asl_loss replaces AsymmetricFocalLoss.forward (losses.py:41-67) AND its autograd
 * backward in one pass; fv_ce_loss replaces F.cross_entropy (utils.py:262).
 *   logits fp32 [B,C] (C <= 32), targets int64 [B]
 *   loss   fp32 [1]  (mean over B; written, not accumulated)
 *   dlogits fp32 [B,C] = d loss / d logits (may be NULL)
 * ---------------------------------------------------------------------------------------- */
int fv_asl_loss(const float* logits, const int64_t* targets, float* loss, float* dlogits,
                int64_t batch, int64_t classes, float gamma_neg, float gamma_pos, float clip,
                float eps, void* stream);
int fv_ce_loss(const float* logits, const int64_t* targets, float* loss, float* dlogits,
               int64_t batch, int64_t classes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser sweep over one flat fp32 parameter arena.
 * Replaces clip_grad_norm_ (utils.py:192-193), torch.optim.AdamW.step over the LLRD groups
 * (train.py:158,253-261; model.py:228-270), EMA.update (utils.py:76-83) and the bf16 re-cast
 * of the weights, in one pass.
 *   seg_end  : int64 [nseg] exclusive end offset of each segment (ascending, last == n)
 *   seg_lr/seg_wd : fp32 [nseg] per-segment learning rate / weight decay; lr < 0 marks a
 *              segment that is not optimised (cls_token / pos_embed quirk, model.py:228-270)
 *   sumsq    : fp32 [1] device scalar, sum of squares of ALL gradients (fv_sumsq)
 *   max_norm : clip threshold (<=0 disables); coefficient = min(1, max_norm/(sqrt(sumsq)+1e-6))
 *   step     : 1-based step count for bias correction
 *   ema/ema_decay : optional shadow arena (NULL to skip)
 *   p_lp     : optional bf16 copy of the updated parameters (NULL to skip)
 *   zero_grad: non-zero -> the sweep, the gradient's last reader, writes zeros back into g: the
 *              optimizer.zero_grad of the next step (train.py:160) costs no pass of its own
 * ---------------------------------------------------------------------------------------- */
int fv_sumsq(const float* g, int64_t n, float* sumsq, int accumulate, void* stream);
int fv_adamw_flat(float* p, float* g, float* m, float* v,
                  const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int nseg,
                  const float* sumsq, float max_norm, float beta1, float beta2, float eps,
                  int64_t step, float* ema, float ema_decay, void* p_lp,
                  int64_t n, int zero_grad, void* stream);
/* The same sweep with the two step-dependent scalars in DEVICE memory — bias_corr[0] = 1 - beta1^t,
 * bias_corr[1] = sqrt(1 - beta2^t) — so a launch captured in a CUDA graph
 * (fedvit_b200.graphs.GraphedTrainStep) replays correctly for every step t. fv_adamw_tick is the
 * captured launch that advances them: *step += 1, bias_corr <- the corrections of that step (device
 * arithmetic in fp64, one thread), so a replay never depends on host writes racing the device. */
int fv_adamw_flat_dev(float* p, float* g, float* m, float* v,
                      const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int nseg,
                      const float* sumsq, float max_norm, float beta1, float beta2, float eps,
                      const float* bias_corr, float* ema, float ema_decay, void* p_lp,
                      int64_t n, int zero_grad, void* stream);
int fv_adamw_tick(int64_t* step, float* bias_corr, float beta1, float beta2, void* stream);
int fv_scale_inplace(float* x, const float* sumsq, float max_norm, int64_t n, void* stream);
int fv_ema_update(float* shadow, const float* p, float decay, int64_t n, void* stream);
int fv_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * FedAvg local fold: acc = (init ? 0 : acc) + weight * w over a flat fp32 arena.
 * There is no FedAvg in the reference; spec in SURVEY.md §8.2:
 *   w_global = sum_k (n_k / sum_j n_j) * w_k, summed in client order in fp32.
 * The cross-GPU sum is an NCCL allreduce issued by the host (torch.distributed).
 * ---------------------------------------------------------------------------------------- */
int fv_fedavg_accum(float* acc, const float* w, float weight, int init, int64_t n, void* stream);
/* The last fold of a round: out = (acc ? acc : 0) + weight * w with the same rounding, where out may
 * alias w — the result lands in the parameter arena the allreduce then runs on in place (no copy of
 * the accumulator back) — and out_lp (optional) receives its bf16 copy (single-rank rounds).      */
int fv_fedavg_fold_into(const float* acc, const float* w, float weight, float* out, void* out_lp,
                        int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * One stage of the reference's MetadataBranch (model.py:27-60), fused:
 *     y = dropout_mask * GELU_erf(BatchNorm1d(x W^T + b))
 * Replaces nn.Linear + nn.BatchNorm1d (training: batch statistics over the batch dimension, biased
 * variance for the normalisation, running_mean / running_var updated with momentum and the UNBIASED
 * variance; eval: running statistics) + nn.GELU + nn.Dropout's multiply (scope row f2; the metadata
 * branch is enabled in the reference's default config.yaml:34-40). fp32.
 *   x [batch, in] (ldx), w [out, in], bias / gamma / beta [out], running_* [out] (updated in place when
 *   training), drop_mask [batch, out] (keep / keep_prob, or NULL)
 *   y [batch, out] (ldy: may be a column slice of a wider buffer)
 *   saved for the backward: xhat [batch, out] (normalised pre-affine values), dact [batch, out]
 *   (= drop_mask * GELU'(z); may be NULL in inference), rstd [out] (may be NULL)
 * Backward: dh [batch, out] (gradient w.r.t. the Linear output; the caller multiplies it by W for dx),
 *   and ACCUMULATES (+=) dw [out, in], dbias, dgamma, dbeta (each may be NULL).
 * One CTA owns 8 features for all rows, so every batch reduction stays inside a CTA: no atomics,
 * bit-reproducible.
 * ---------------------------------------------------------------------------------------- */
int fv_linear_bn_gelu_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                          int training, const float* drop_mask, float* y, int64_t ldy, float* xhat, float* dact,
                          float* rstd, int64_t batch, int64_t in_features, int64_t out_features, void* stream);
int fv_linear_bn_gelu_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* xhat,
                          const float* dact, const float* rstd, const float* gamma, int training, float* dh,
                          float* dw, float* dbias, float* dgamma, float* dbeta, int64_t batch,
                          int64_t in_features, int64_t out_features, void* stream);

/* fp32 attention helpers for the parity path: row softmax fwd/bwd on [rows, cols] scores      */
int fv_softmax_rows(const float* s, float* p, int64_t rows, int64_t cols, float scale, void* stream);
int fv_softmax_rows_bwd(const float* p, const float* dp, float* ds, int64_t rows, int64_t cols,
                        float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEDVIT_H_ */
