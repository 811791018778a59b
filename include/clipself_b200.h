/*
 * clipself_b200 — C ABI of the B200-native CLIPSelf distillation hot path.
 *
 * Drop-in boundary for the path  src/training/clipself.py:7-49  of wusize/CLIPSelf and the model
 * entry points it calls (eva_clip/model.py:313-346, eva_clip/eva_vit_model.py:533-664).
 * The reference has no native layer (everything is Python -> ATen / torchvision / xformers), so
 * the "FFI" a maintainer binds is ctypes: see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch's
 *     caching allocator in our host code) owns all buffers, including outputs and workspaces;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device or
 *     touches a hidden stream; all entry points are re-entrant;
 *   - return value: 0 (CS_OK) or a CS_ERR_* code; cs_last_error() gives the thread-local text;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - dtype arguments use cs_dtype_t.
 * All kernels are compiled for sm_100a only; there is no CPU path and no fallback.
 */
#ifndef CLIPSELF_B200_H_
#define CLIPSELF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_OK 0
#define CS_ERR_INVALID_ARGUMENT 1
#define CS_ERR_CUDA 2
#define CS_ERR_UNSUPPORTED 3

typedef enum { CS_F32 = 0, CS_BF16 = 1 } cs_dtype_t;

const char* cs_last_error(void);
/* Library/ABI version (bumped on any signature change). */
int cs_abi_version(void);
/* Fills sm (e.g. 100), SM count, and total HBM bytes of the current device; CS_ERR_CUDA if no
 * usable sm_100 device is present (the product path must fail loudly, never fall back). */
int cs_device_info(int* sm_out, int* num_sms_out, int64_t* hbm_bytes_out);
/* Number of cuTensorMapEncodeTiled driver calls made so far by this process (TMA descriptors are cached per
 * (address, geometry): a steady-state step makes none). */
int64_t cs_tensor_map_encodes(void);

/* ------------------------------------------------------------------------------------------
 * Region path (HBM-bound, no tensor cores)
 * ---------------------------------------------------------------------------------------- */

/* Valid-box / crop index extraction — replaces clipself.py:29-36.
 *   normed_boxes [B,K,5] f32 (x0,y0,x1,y1,valid).  valid := box[4] > 0.5.
 * Outputs, image-major and order-preserving (bit-exact copies of the inputs):
 *   rois      [B*K,4] f32  first R rows filled, the rest zeroed   (== cat(rois_list))
 *   crop_index[B*K]   i32  flat index b*K+k of each kept row (== index into image_crops.flatten(0,1))
 *   roi_batch [B*K]   i32  image index of each kept row
 *   img_offsets[B+1]  i32  exclusive prefix sum of per-image counts; img_offsets[B] == R      */
int cs_extract_rois(const float* normed_boxes, int B, int K, float* rois, int32_t* crop_index,
                    int32_t* roi_batch, int32_t* img_offsets, void* stream);

/* Gather rows: dst[r,:] = src[index[r],:], row_bytes % 16 == 0 — the crop gather of
 * clipself.py:34-36 (torch.cat(crops_list)) when the crops are device resident. */
int cs_gather_rows(const void* src, const int32_t* index, int R, int64_t row_bytes, void* dst,
                   void* stream);

/* RoIAlign, output 1x1, spatial_scale 1, sampling_ratio -1 (adaptive), aligned=True, on an NHWC
 * f32 map — replaces _denormalize_boxes + torchvision.ops.roi_align at eva_vit_model.py:625-629,
 * 655-664.  rois are the NORMALISED boxes of cs_extract_rois; the x*=W, y*=H denormalisation
 * happens inside in f32 exactly like the reference.
 *   fmap [B,H,W,C] f32, out [R,C] f32.  The separable sampling weights are written to
 *   wy [R,H] / wx [R,W] (f32, caller provided) and reused by the backward. */
int cs_roi_align_fwd(const float* fmap, int B, int H, int W, int C, const float* rois,
                     const int32_t* img_offsets, int R, float* wy, float* wx, float* out,
                     void* stream);
/* d_fmap [B,H,W,C] f32 is fully overwritten (deterministic gather form, no atomics) —
 * replaces torchvision's roi_align_backward_kernel. */
int cs_roi_align_bwd(const float* d_out, int B, int H, int W, int C, const int32_t* img_offsets,
                     int R, const float* wy, const float* wx, float* d_fmap, void* stream);

/* Mask pooling — replaces EVAVisionTransformer.mask_pool, eva_vit_model.py:645-653, without the
 * repeat_interleave materialisation.  fmap [B,HW,C] f32, masks [R,HW] f32 (image-major),
 * out[r] = sum_p f[b(r),p]*m[r,p] / (sum_p m[r,p] + 1e-12). */
int cs_mask_pool_fwd(const float* fmap, int B, int HW, int C, const float* masks,
                     const int32_t* img_offsets, int R, float* out, void* stream);

/* L2-normalise + cosine loss — replaces clipself.py:42-47.
 *   loss = (1 - mean_r <s_r/max(|s_r|,1e-12), t_r/max(|t_r|,1e-12)>) * weight
 * s,t [R,C] f32; loss: 1 f32; row_stats [R,3] f32 = (1/max|s|, 1/max|t|, cos_r) kept for the
 * backward; `scratch` >= 4 bytes zero-initialised by the callee (ticket counter). Deterministic
 * (fixed-order final reduction). */
int cs_cosine_loss_fwd(const float* s, const float* t, int R, int C, float weight, float* loss,
                       float* row_stats, void* stream);
/* d_s [R,C] = d_loss * dloss/ds ; d_loss is a DEVICE scalar (f32). */
int cs_cosine_loss_bwd(const float* s, const float* t, const float* row_stats, int R, int C,
                       float weight, const float* d_loss, float* d_s, void* stream);

/* Row-wise L2 normalisation y = x / max(|x|,1e-12) (F.normalize, eva_vit_model.py:620) over
 * [M,C] f32; inv_norm [M] is saved for the backward. */
int cs_l2norm_fwd(const float* x, int64_t M, int C, float* y, float* inv_norm, void* stream);
int cs_l2norm_bwd(const float* y, const float* inv_norm, const float* d_y, int64_t M, int C,
                  float* d_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Crop generation — replaces the CPU PIL path of
 * GridDistillDataset._obtain_image_crops (training/data.py:226-245: image.crop(box) -> transforms[1]) and
 * the transforms of open_clip/transform.py:26-49,119-133 (ResizeMaxSize, centre pad) / :136-191
 * (ResizeLongest, pad right/bottom), bit-exact with Pillow's bicubic ImagingResample.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t x0, y0, x1, y1;      /* source rectangle, pixels (Image.crop's rounded box; outside the image = 0) */
    int32_t out_w, out_h;        /* size after the bicubic resize (round(side * max_size / longest side))        */
    int32_t pad_left, pad_top;   /* where the resized crop sits inside the size x size zero canvas               */
} cs_crop_desc_t;

/* bytes of scratch cs_crop_resize_normalize needs: coefficient tables + the uint8 intermediate of the
 * horizontal pass; ksize_max >= 2*ceil(2*max(in/out,1))+1 over all crops and axes, tmp_rows_max >= max crop height. */
int cs_crop_workspace_bytes(int K, int size, int ksize_max, int tmp_rows_max, int64_t* bytes);

/* image_hwc: uint8 [H,W,3] on the device; descs: K cs_crop_desc_t on the device; mean3 / std3: host floats;
 * out: f32 [K,3,size,size] = Normalize(ToTensor(pad(resize(crop)))). */
int cs_crop_resize_normalize(const uint8_t* image_hwc, int H, int W, const void* descs, int K, int size,
                             int ksize_max, int tmp_rows_max, const float* mean3, const float* std3, float* out,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* The same for a BATCH of images in one call (one training batch of GridDistillDataset samples, training/data.py:226-281):
 * images_blob holds the uint8 [H_i,W_i,3] images back to back, image_offsets[i] (bytes) and image_hw[2i], image_hw[2i+1]
 * (H_i, W_i) describe image i, desc_image[k] is the image crop k is cut from.  All tables on the device. */
int cs_crop_resize_normalize_batched(const uint8_t* images_blob, const int64_t* image_offsets, const int32_t* image_hw,
                                     const int32_t* desc_image, const void* descs, int K, int size, int ksize_max,
                                     int tmp_rows_max, const float* mean3, const float* std3, float* out, void* workspace,
                                     int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tower kernels (EVA02 ViT, eva_vit_model.py)
 * ---------------------------------------------------------------------------------------- */

/* Non-overlapping patch gather (im2col of the k=s=P conv, eva_vit_model.py:348-356):
 *   images [B,3,S,S] (f32 or bf16) -> patches [B*g*g, ldp] bf16, column order (c,py,px) matching
 *   the flattened conv weight [D, 3*P*P]; columns [3*P*P, ldp) are zero padding. */
int cs_im2col_patches(const void* images, cs_dtype_t dtype, int B, int S, int P,
                      void* patches_bf16, int64_t ldp, void* stream);

/* Bilinear resize of `planes` image planes [Hin,Win] -> [Hout,Wout] (f32 or bf16, same dtype out),
 * align_corners = false, no antialiasing: the --multiscale student input
 * (training/clipself.py:17-27, F.interpolate(images, size=(t,t), mode='bilinear')). */
int cs_resize_bilinear(const void* src, cs_dtype_t dtype, int64_t planes, int Hin, int Win, int Hout, int Wout,
                       void* dst, void* stream);

/* x[b,0,:] = cls_token + pos_embed[0]  (eva_vit_model.py:540-543); x [B,N,D] f32. */
int cs_fill_cls_rows(const float* cls_token, const float* pos_embed, int B, int N, int D, float* x,
                     void* stream);

/* LayerNorm forward (eva_clip/transformer.py:52-58, eps from model.py:123), one row per warp.
 *   x: [*, ldx] f32 or bf16;  logical row m reads physical row
 *        m*row_mul + (row_div > 0 ? m/row_div : 0) + row_off
 *   (identity: mul 1, div 0, off 0;  CLS rows only: mul = tokens;  patch rows only:
 *    div = tokens-1, off = 1);
 *   y [M, ldy] bf16;  mean/rstd [M] f32 optional (NULL to skip; needed for backward). */
int cs_layernorm_fwd(const void* x, cs_dtype_t x_dtype, int64_t ldx, int64_t M, int D,
                     int row_div, int row_mul, int row_off,
                     const float* gamma, const float* beta, float eps, void* y_bf16, int64_t ldy,
                     float* mean, float* rstd, void* stream);

/* x [M, ldx] f32 -> xb [M, ldb] bf16 (round to nearest even) and stats [M, parts, 2] f32 = the (sum, sum of squares)
 * of every row in part 0, zeros in the other parts: the operands of a LayerNorm folded into the next GEMM
 * (cs_gemm_epilogue_t.ln_stats) for a residual stream that was not produced by a GEMM epilogue (the patch embedding
 * output, eva_vit_model.py:540-544).  D, ldx, ldb multiples of 4. */
int cs_row_stats_cast(const float* x, int64_t ldx, int64_t M, int D, void* xb_bf16, int64_t ldb, float* stats,
                      int parts, void* stream);

/* Epilogue description of cs_gemm_bf16 (all pointers device, may be NULL when unused). */
typedef enum {
    CS_EPI_STORE = 0,      /* out = acc + bias (+ residual)                                     */
    CS_EPI_QKV_ROPE = 1,   /* out(bf16) = rope(acc + bias) on columns < rope_cols, patch rows   */
    CS_EPI_SWIGLU = 2,     /* out(bf16)[m, n/2-ish] = silu(acc1+b1) * (acc2+b2), gate/up packed */
    CS_EPI_TOKENS = 3      /* patch-embed: row remap past CLS, + bias + pos_embed               */
} cs_epilogue_mode_t;

typedef struct {
    int32_t mode;            /* cs_epilogue_mode_t */
    int32_t out_dtype;       /* cs_dtype_t */
    void* out;               /* [M(+), ldo] */
    int64_t ldo;
    const float* bias;       /* [N] or NULL */
    const float* residual;   /* [M, ldr] f32 or NULL; may alias out (in-place x += ...) */
    int64_t ldr;
    const float* rope_pos;   /* [rope_grid] f32: position of grid index i, i/ft*pt (rope.py:127) */
    const float* rope_freq;  /* [16] f32: theta^(-2n/32) (rope.py:118); angle(token,d) = pos * freq   */
    int32_t rope_grid;       /* grid side; tokens == rope_grid^2 + 1 */
    int32_t tokens;          /* tokens per image incl. CLS (QKV_ROPE, TOKENS) */
    int32_t rope_cols;       /* columns [0, rope_cols) are rotated (= 2*D for q|k|v) */
    const float* pos_embed;  /* [tokens, N] f32 (TOKENS) */
    float alpha;             /* scale applied to acc before everything else (1.0 default) */
    int32_t reserved;        /* must be 0 (debug ablation switches) */
    /* LayerNorm folding (frozen teacher): the GEMM runs on the UN-normalised rows a with weights
     * W' = W*diag(gamma); the epilogue applies  y = rstd*(acc - mean*ln_c1[n]) + bias[n]  with
     * ln_c1 = rowsum(W'), bias = W*beta + b, and (mean, rstd) of each input row rebuilt from
     * ln_parts partial (sum, sumsq) pairs:  ln_stats [M, ln_parts, 2] (ln_parts even).  Every mode
     * but TOKENS (the folded value then goes through RoPE / SwiGLU / the residual add like a plain
     * accumulator); alpha must be 1.  NULL ln_stats disables it. */
    const float* ln_stats;
    const float* ln_c1;      /* [N] */
    int32_t ln_parts;
    int32_t ln_dim;          /* number of elements the statistics run over (K of the GEMM) */
    float ln_eps;
    int32_t reserved2;       /* split-K: 0 = off, -1 = choose, n = n splits; partial products are red.add'ed into
                              * `out`, which must be zero on entry (f32 STORE, no residual) */
    /* Row statistics for a LayerNorm folded into the NEXT GEMM: partial (sum, sumsq) of the f32 output values
     * per (row, n-tile, column half of the tile)  ->  stats_out [M, 2*ceil(N/T), 2], T = 256 if N % 256 == 0 else 128
     * (SWIGLU: T = 256 packed columns = 128 outputs; the statistics are over the outputs).  Produced by the
     * SWIGLU epilogue (ffn_ln folded into w3) and by the out2_bf16 epilogue (norm1 / norm2 folded into the
     * qkv / w1|w2 GEMMs).  Optional. */
    float* stats_out;
    /* STORE / f32 / residual only, optional: additionally write the output rounded to bf16
     * (out2_bf16 [M, ldo2]) — the new residual stream as the next GEMM's A operand, so no separate
     * LayerNorm / cast pass reads the f32 stream again. */
    void* out2_bf16;
    int64_t ldo2;
} cs_gemm_epilogue_t;

/* C[M,N] = A[M,K] · W[N,K]^T on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM ->
 * epilogue).  A, W bf16 row-major, lda/ldw multiples of 8; K % 8 == 0; N % 32 == 0.
 * This is the one GEMM behind every F.linear / conv of the tower (eva_vit_model.py:177-179,
 * 220, 98-105, 355, 569) and their backward products. */
int cs_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int N, int K,
                 const cs_gemm_epilogue_t* epi, void* stream);

/* C[M,N] (f32) = At^T · Bt with BOTH operands stored reduction-major: At [K, M] (row stride lda), Bt [K, N] (row stride
 * ldb), bf16.  This is the weight gradient of every nn.Linear of the student (the autograd backward of
 * eva_vit_model.py:177-179, 220, 98-105 under train.py:104): dW[out,in] = dY[tokens,out]^T · X[tokens,in] reads dY and X
 * as they lie in memory — the tensor cores take both tiles MN-major, no transposed copies are made.
 * Epilogue: plain f32 STORE only (mode STORE, no bias / residual / folding); epi->reserved2 selects split-K exactly as
 * in cs_gemm_bf16 (-1 = choose; `out` must then be zero on entry, partial products are accumulated with red.add). */
int cs_gemm_bf16_tn(const void* At, int64_t lda, const void* Bt, int64_t ldb, int64_t M, int N, int K,
                    const cs_gemm_epilogue_t* epi, void* stream);

/* Pack the SwiGLU gate/up weights so that one GEMM tile holds matching x1/x2 columns:
 *   packed row (t*2*half + j)        = w1 row (t*half + j)
 *   packed row (t*2*half + half + j) = w2 row (t*half + j),   half = 128, j < half
 * w1,w2 [Hd, K] f32 or bf16 -> packed [2*Hd_pad, ldk] bf16; biases likewise into bias12 [2*Hd_pad]. */
int cs_pack_swiglu_weights(const void* w1, const void* w2, cs_dtype_t dtype, int Hd, int K,
                           const float* b1, const float* b2, void* packed_bf16, int64_t ldk,
                           float* bias12, void* stream);

/* Non-causal softmax attention over the packed projections (eva_vit_model.py:206-217 / 221-246):
 *   qkv [B*N, 3*D] bf16 (q | k | v, each H heads of 64, RoPE already applied to q,k),
 *   out [B*N, D] bf16;  lse [B,H,N] f32 optional (needed for the backward).  head_dim must be 64.
 *   row_stats (optional): [B*N, 4H, 2] f32 partial (sum, sum of squares) of every
 *   output row per (head, 16-dim quarter) — consumed by the LayerNorm-folded proj GEMM (ln_* fields of
 *   cs_gemm_epilogue_t) so inner_attn_ln (eva_vit_model.py:218) needs no pass of its own.
 *   N <= 224 runs the tcgen05 kernel (S/O in TMEM), longer sequences the streaming kernel. */
int cs_attention_fwd(const void* qkv_bf16, int B, int N, int H, float scale, void* out_bf16,
                     float* lse, float* row_stats, void* stream);

/* f32 -> bf16 cast of a [rows, cols] matrix into a (possibly wider, zero padded) bf16 matrix. */
int cs_cast_pad_bf16(const float* src, int64_t rows, int64_t cols, int64_t lds, void* dst_bf16,
                     int64_t ldd, void* stream);

/* ------------------------------------------------------------------------------------------
 * Student backward (autograd of the tower as driven by train.py:96 `backward(total_loss)`)
 * ---------------------------------------------------------------------------------------- */

/* Softmax-attention backward for head_dim 64 (replaces the xformers / ATen backward of
 * eva_vit_model.py:206-217).  Inputs are the tensors of cs_attention_fwd plus d_out [B*N,D] bf16;
 * output dqkv [B*N,3D] bf16 holds dq | dk | dv.  When rope tables are given the rotation applied
 * by the CS_EPI_QKV_ROPE epilogue is undone on dq/dk (patch tokens only), i.e. the result is the
 * gradient w.r.t. the raw projections.  delta_ws: B*H*N floats of scratch. Deterministic. */
int cs_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                     int B, int N, int H, float scale, const float* rope_cos, const float* rope_sin,
                     float* delta_ws, void* dqkv_bf16, void* stream);

/* src [M,N] (f32|bf16, ld lds) -> dst [M,ldd] bf16 (optional) and its transpose dst_t [N,ldt] bf16
 * (optional): operand staging for the dgrad (dY·W) and wgrad (dY^T·X) GEMMs. */
int cs_cast_transpose_bf16(const void* src, cs_dtype_t dtype, int64_t M, int N, int64_t lds,
                           void* dst_bf16, int64_t ldd, void* dst_t_bf16, int64_t ldt, void* stream);

/* LayerNorm backward w.r.t. the input:  dx = rstd*(g*dy - mean(g*dy) - xhat*mean(g*dy*xhat)) (+ add).
 * x uses the forward's row map; when dx_row_mapped != 0, dx/add rows use it too (scatter back
 * into the [B,N,D] residual gradient).  */
int cs_layernorm_bwd_dx(const void* dy, cs_dtype_t dy_dtype, int64_t lddy, const void* x,
                        cs_dtype_t x_dtype, int64_t ldx, int64_t M, int D, int row_div, int row_mul,
                        int row_off, const float* mean, const float* rstd, const float* gamma,
                        const float* add, int64_t ldadd, void* dx, cs_dtype_t dx_dtype, int64_t lddx,
                        int dx_row_mapped, void* stream);

/* Column reductions over the M rows, deterministic two-stage:
 *   dbeta[c]  = sum_m dy[m,c]                       (bias gradients when x == NULL)
 *   dgamma[c] = sum_m dy[m,c] * (x[m,c]-mean[m])*rstd[m]   (LayerNorm weight gradient, x != NULL)
 * workspace: at least 128*2*D floats. */
int cs_col_reduce(const void* dy, cs_dtype_t dy_dtype, int64_t lddy, const void* x, cs_dtype_t x_dtype,
                  int64_t ldx, int64_t M, int D, int row_div, int row_mul, int row_off,
                  const float* mean, const float* rstd, float* dgamma, float* dbeta,
                  float* workspace, int64_t workspace_floats, void* stream);

/* SwiGLU (eva_vit_model.py:98-101) on pre-activations x12 [M, 2*Hd] bf16:
 *   split_layout == 1: gate = x12[:, j], up = x12[:, Hd + j]           (weights [w1; w2] stacked)
 *   split_layout >= 2: gate = x12[:, j], up = x12[:, split_layout + j] (padded halves, ViT-L: Hd 2730 -> 2816)
 *   split_layout == 0: the packed layout of cs_pack_swiglu_weights (gate|up interleaved by 128)
 *   h = silu(gate) * up  and its backward (dx12 in the same layout as x12). */
int cs_swiglu_fwd(const void* x12_bf16, int64_t M, int Hd, int64_t ld12, void* h_bf16, int64_t ldh,
                  int split_layout, void* stream);
int cs_swiglu_bwd(const void* x12_bf16, const void* dh_bf16, int64_t M, int Hd, int64_t ld12,
                  int64_t lddh, void* dx12_bf16, int split_layout, void* stream);

/* Attention for the CLS query only: qkv [B*N, 3*H*64] bf16 (CLS = row 0 of each image) -> out_cls [B, H*64] bf16 and,
 * optionally, row_stats_cls [B, 4H, 2] (sum / sum of squares per 16-dim quarter of the f32 output, for the folded
 * inner_attn_ln).  eva_vit_model.py:206-217 for the LAST block of a tower that is read at the CLS token only
 * (encode_image, eva_vit_model.py:565-569): cs_vit_forward_cls runs that block on the CLS rows alone.  N <= 1024. */
int cs_attention_cls_fwd(const void* qkv_bf16, int B, int N, int H, float scale, void* out_cls_bf16, float* row_stats_cls,
                         void* stream);

/* Fused AdamW over a flat f32 buffer — torch.optim.AdamW semantics as configured at
 * main.py:205-213 (decoupled weight decay, bias correction); grad is multiplied by grad_scale
 * first (1/world_size after a sum all-reduce). `step` is 1-based. */
int cs_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int step,
                  double grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tower level: a frozen EVA02 vision tower as one handle and one call.
 * What a reference maintainer binds instead of EVAVisionTransformer.forward / encode_dense
 * (eva_vit_model.py:533-623) when the tower is frozen (the CLIPSelf teacher, clipself.py:37-38, and
 * every inference / eval call).  Kernel sequence and LayerNorm folding: csrc/tower.cu, DESIGN.md §5.2.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t image_size, patch, width, heads, layers, hidden, embed_dim, pt_seq_len;   /* CLIPVisionCfg, eva_clip/model.py:36-62 */
    float ln_eps;                                                                     /* 1e-6, eva_clip/model.py:123 */
} cs_tower_cfg_t;
typedef struct cs_tower cs_tower_t;

/* Bytes of device memory the packed (bf16, LayerNorm-folded) weights of one tower need. */
int cs_pack_weights_bytes(const cs_tower_cfg_t* cfg, int64_t* bytes);
/* Pack the tower's weights.  names[i] / tensors_f32[i]: the `visual.*` state_dict entries WITHOUT the prefix
 * ("blocks.0.attn.q_proj.weight", "pos_embed", ...; D.1 of SURVEY.md), contiguous f32 device tensors; extra entries
 * (rope buffers) are ignored, a missing one is an error.  pack_buffer: caller-owned device memory of
 * cs_pack_weights_bytes() bytes, 256 B aligned, which must outlive the handle.  Synchronises `stream` once. */
int cs_pack_weights_create(const cs_tower_cfg_t* cfg, const char* const* names, const void* const* tensors_f32, int count,
                           void* pack_buffer, int64_t pack_bytes, void* stream, cs_tower_t** out);
/* Re-pack after the weights changed (same buffers; CUDA graphs captured by the forward calls stay valid). */
int cs_pack_weights_update(cs_tower_t* tower, const char* const* names, const void* const* tensors_f32, int count, void* stream);
int cs_pack_weights_destroy(cs_tower_t* tower);

/* Bytes of activation scratch for chunks of `chunk_images` images at `image_size` (0 = the tower's own size). */
int cs_query_workspace(const cs_tower_cfg_t* cfg, int chunk_images, int image_size, int64_t* bytes);

/* encode_image(normalize=False) of a frozen tower: images [n,3,S,S] (f32 or bf16, S = cfg.image_size) -> out [n, embed_dim] f32.
 * Processed in chunks of chunk_images through the caller's workspace (256 B aligned, cs_query_workspace bytes).  The third and
 * later calls with the same (images, workspace, out) addresses and shapes replay a CUDA graph (CLIPSELF_NO_GRAPH=1 disables). */
int cs_vit_forward_cls(cs_tower_t* tower, const void* images, cs_dtype_t dtype, int n_images, void* workspace,
                       int64_t workspace_bytes, int chunk_images, float* out, void* stream);
/* encode_dense (eva_vit_model.py:588-623) without a tape: -> out NHWC [n, S/p, S/p, embed_dim] f32, unit norm per token.
 * image_size: any multiple of the patch size up to a 64 x 64 grid (0 = native); for a non-native size the caller passes the
 * bicubically rescaled pos_embed [1 + grid^2, width] (eva_vit_model.py:631-643), else NULL. */
int cs_vit_forward_dense(cs_tower_t* tower, const void* images, cs_dtype_t dtype, int n_images, int image_size,
                         const float* pos_embed_rescaled, void* workspace, int64_t workspace_bytes, int chunk_images,
                         float* out_nhwc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPSELF_B200_H_ */
