/* libuc_b200 -- C ABI of the B200-native DUSt3R two-view hot path.
 *
 * Drop-in boundary.  The reference (castacks/UniCeption @ 802ebc17) is pure
 * PyTorch with ONE native entry point, the pybind function
 *     curope.rope_2d(tokens[B,N,H,D], positions[B,N,2] int64, base, +-F0)        (curope.cpp:49-69)
 * which `uc_rope2d` replaces 1:1.  Every other entry point below replaces a torch library call made
 * by a reference module on the path; the file:line of the call it replaces is cited per function.
 *
 * Conventions (all functions):
 *   - plain C, no torch types; raw DEVICE pointers + explicit sizes / leading dimensions (in elements);
 *   - the caller allocates every output and workspace and passes its CUDA stream
 *     (`torch.cuda.current_stream().cuda_stream`); functions are asynchronous, never synchronise,
 *     never allocate device memory, and are CUDA-graph capturable;
 *   - return 0 on success or a negative UC_ERR_* code; `uc_last_error` returns the message of the
 *     last failure on the calling thread;
 *   - there is NO CPU fallback: on a machine without an sm_100 GPU every compute call fails with
 *     UC_ERR_CUDA.
 */
#ifndef UC_B200_H
#define UC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* uc_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define UC_API __attribute__((visibility("default")))
#else
#define UC_API
#endif

#define UC_OK 0
#define UC_ERR_BAD_SHAPE (-1)
#define UC_ERR_BAD_DTYPE (-2)
#define UC_ERR_CUDA (-3)
#define UC_ERR_UNSUPPORTED (-4)

#define UC_DTYPE_BF16 0
#define UC_DTYPE_F32 1
#define UC_DTYPE_F16 2

UC_API int uc_version(void);
/* Copies the last error message of this thread into buf (NUL-terminated); returns its length. */
UC_API size_t uc_last_error(char* buf, size_t cap);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
UC_API uint64_t uc_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM with fused epilogue: the tcgen05 tensor-core kernel behind every nn.Linear / 1x1 conv /
 * patch-embed conv of the path (libs/croco/blocks.py:74-77,97-99; utils/transformer_blocks.py:76-79,
 * 191-199,305-311; info_sharing/cross_attention_transformer.py:116; prediction_heads/linear.py:47-54;
 * libs/croco/patch_embed.py:47) and their autograd dgrad / wgrad.
 *
 *   C[m,n] = epilogue( sum_k A(m,k) * B(n,k) )          bf16 operands, fp32 accumulation in TMEM
 *
 *   a_layout 0: A stored [m][k] (k contiguous, lda = row pitch)   1: A stored [k][m] (m contiguous)
 *   b_layout 0: B stored [n][k] (k contiguous, ldb = row pitch)   1: B stored [k][n] (n contiguous)
 *     forward  y = x W^T : A=x (0),  B=W (0)
 *     dgrad   dx = dy W  : A=dy (0), B=W (1)   [k = out features]
 *     wgrad   dW = dy^T x: A=dy (1), B=x (1)   [k = tokens], fp32 output, split-k + atomics
 * Epilogue, applied in this order on the fp32 accumulator v (row m, column n):
 *   UC_EPI_BIAS       v += bias[n]                          (fp32 bias)
 *   UC_EPI_ROPE       2-D RoPE on columns n < rope_cols (head_dim 64; q|k thirds of a packed qkv):
 *                     pair (i, i+16) inside each 32-column half-head rotated by
 *                     rope_table[pos[m][half]][i] = (cos, sin)   (curope/kernels.cu:39-80)
 *   UC_EPI_GELU       aux_out[m,n] = bf16(v);  v = gelu_erf(bf16(v))     (blocks.py:80-86)
 *   UC_EPI_GELU_BWD   v *= gelu_erf'(aux_in[m,n])
 *   UC_EPI_RELU / UC_EPI_RELU_BWD   v = max(v,0) / v *= (aux_in[m,n] > 0)
 *   UC_EPI_RESIDUAL   v += residual[m,n]                    (bf16, pitch ldc)
 *   UC_EPI_ATOMIC     C += v with red.global.add (fp32 C only; implied when split_k > 1)
 * ------------------------------------------------------------------------------------------ */
#define UC_EPI_BIAS 1
#define UC_EPI_ROPE 2
#define UC_EPI_GELU 4
#define UC_EPI_GELU_BWD 8
#define UC_EPI_RESIDUAL 16
#define UC_EPI_ATOMIC 32
#define UC_EPI_RELU 64      /* v = max(v, 0)                      (DPT convs, dpt_block.py:114-177) */
#define UC_EPI_RELU_BWD 128 /* v *= (aux_in[m,n] > 0)             (aux_in = the ReLU output) */
#define UC_EPI_RESIDUAL_F32 256 /* with UC_EPI_RESIDUAL and fp32 C: `residual` is fp32 [m][ldc] -- an fp32 residual stream, which
                                 * is what torch.autocast gives the reference's global / alternating transformers once the fp32
                                 * view encoding has been added to the bf16 projection (global_attention_transformer.py:338-351) */

typedef struct {
  const void* a;
  const void* b;
  void* c;
  int32_t m, n, k;
  int32_t a_layout, b_layout;
  int64_t lda, ldb, ldc;
  int32_t c_dtype;  /* UC_DTYPE_BF16 or UC_DTYPE_F32 */
  int32_t epilogue; /* UC_EPI_* flags */
  int32_t split_k;  /* >= 1; 0 = choose */
  int32_t rope_cols;
  const float* bias;          /* [n] fp32 */
  const void* residual;       /* [m][ldc] bf16 (fp32 with UC_EPI_RESIDUAL_F32) */
  void* aux_out;              /* [m][ldc] bf16 (UC_EPI_GELU pre-activation) */
  const void* aux_in;         /* [m][ldc] bf16 (UC_EPI_GELU_BWD pre-activation) */
  const int32_t* positions;   /* [m][2] int32 (y,x) per row (UC_EPI_ROPE) */
  const float* rope_table;    /* [P][16][2] fp32 (cos,sin) from uc_rope2d_table */
  float* c_colsum;            /* optional [n] fp32, accumulated: column sums of the bf16 C this call writes (the bias gradient
                                 of the Linear whose output gradient C is, e.g. fc1 when C = dgrad(fc2) * GELU'); bf16 C only */
} uc_gemm_params;

UC_API int uc_gemm(const uc_gemm_params* p, uc_stream_t stream);
/* Tile distribution of the CTA-pair GEMM: 0 (default) static round robin over the persistent clusters; 1 an atomic tile queue, so
 * that clusters whose SMs are temporarily held by other kernels (the NCCL all-reduce that data parallelism overlaps with the
 * backward pass -- the reference's users get this from DistributedDataParallel, prediction_heads/dpt.py:82 -- or another
 * stream's kernel) take fewer tiles instead of adding a wave.  Returns the previous setting.  Process-wide. */
UC_API int uc_set_gemm_dynamic(int on);

/* ------------------------------------------------------------------------------------------
 * 2-D RoPE.  uc_rope2d == curope.rope_2d (curope.cpp:49-69, kernels.cu:17-108): in place on
 * tokens [B,N,H,D] (element strides given, so the [B,H,N,D] view of curope2d.py:38 works),
 * positions [B,N,2] int64, fwd = +F0 (forward) / -F0 (backward).  fp32 angle math.
 * uc_rope2d_table fills table[p][i] = (cos, sin)(p * fwd / base^(i/Q)) for p < P, i < Q.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  void* tokens;
  const int64_t* positions;
  int32_t B, N, H, D;
  int64_t stride_b, stride_n, stride_h; /* element strides of tokens; D is contiguous */
  int32_t dtype;
  float base, fwd;
} uc_rope2d_params;
UC_API int uc_rope2d(const uc_rope2d_params* p, uc_stream_t stream);
UC_API int uc_rope2d_table(float* table, int32_t P, int32_t Q, float base, float fwd, uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (nn.LayerNorm(C, eps=1e-6): encoders/croco.py:32,127;
 * libs/croco/blocks.py:148,154; utils/transformer_blocks.py:569,587,589,607).
 * fwd: y = (x-mean)*rstd*gamma + beta; x bf16/fp32 [rows][C], y bf16/fp32, stats fp32 [rows].
 * bwd: dx = LN'(dy) (+ dres);  dgamma/dbeta are ACCUMULATED (atomicAdd) into fp32 [C] buffers.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  float* mean;
  float* rstd;
  int32_t rows, C;
  int32_t x_dtype, y_dtype;
  float eps;
} uc_layernorm_fwd_params;
UC_API int uc_layernorm_fwd(const uc_layernorm_fwd_params* p, uc_stream_t stream);

typedef struct {
  const void* dy;   /* [rows][C] */
  const void* x;    /* [rows][C] LN input */
  const void* dres; /* optional [rows][C] bf16: extra gradient added to dx (residual path) */
  void* dx;         /* [rows][C] bf16 */
  const float* gamma;
  const float* mean;
  const float* rstd;
  float* dgamma; /* [C] fp32, accumulated */
  float* dbeta;  /* [C] fp32, accumulated */
  int32_t rows, C;
  int32_t dy_dtype, x_dtype;
  float* dx_colsum; /* optional [C] fp32, accumulated: column sums of the bf16 dx this call writes -- the bias gradient of
                       the Linear whose output gradient dx is (fc2 / proj of the next sub-block in backward order), so that
                       tensor is not re-read by uc_colsum */
} uc_layernorm_bwd_params;
UC_API int uc_layernorm_bwd(const uc_layernorm_bwd_params* p, uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attention core softmax(q k^T * scale) v, head_dim 64, no mask, replacing
 * F.scaled_dot_product_attention (libs/croco/blocks.py:123; utils/transformer_blocks.py:244,373).
 * q/k/v/o are token-major bf16 matrices: element (b, token, head, d) at
 *   base + (b*N + token) * ld + head*64 + d     (so a packed qkv [B*N, 3C] works with 3 base pointers).
 * RoPE has already been applied to q and k (fused into the producing GEMM's epilogue).
 * lse [B][H][Nq] fp32 = log-sum-exp of the scaled scores (natural log), saved for backward.
 * uc_attn_bwd: dq, dk, dv with recomputation of the probabilities (two kernels: key-outer for dk / dv, query-outer for dq;
 * both bit-reproducible); scale and the optional inverse RoPE are applied in the epilogues.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;
  int32_t B, H, Nq, Nk;
  int64_t ldq, ldk, ldv, ldo;
  float scale;
} uc_attn_fwd_params;
UC_API int uc_attn_fwd(const uc_attn_fwd_params* p, uc_stream_t stream);

typedef struct {
  const void* q;
  const void* k;
  const void* v;
  const void* o;
  const void* d_o;
  const float* lse;
  float* delta;   /* workspace, fp32, B*H*ceil(Nq/64)*128 elements: per (b,h) and 64-query tile, 64 x lse*log2(e) then
                     64 x delta = rowsum(dO o O) (rows >= Nq: +inf, 0), written by the call's statistics kernel */
  float* dq_acc;  /* unused (may be NULL): dQ is produced by a query-outer kernel, no fp32 accumulator / atomics.  Only the
                     first-generation kernel (UC_ATTN_BWD=1, A/B baseline) needs [B*Nq][H*64] fp32 here, zeroed by the call */
  void* dq;
  void* dk;
  void* dv;       /* bf16 outputs, same addressing as q/k/v with lddq/lddk/lddv */
  int32_t B, H, Nq, Nk;
  int64_t ldq, ldk, ldv, ldo, lddq, lddk, lddv;
  float scale;
  /* optional inverse RoPE on dq / dk (gradient w.r.t. the un-rotated projections) */
  const int32_t* q_positions; /* [B*Nq][2] or NULL */
  const int32_t* k_positions; /* [B*Nk][2] or NULL */
  const float* rope_table;    /* forward table; the kernel uses (cos, -sin) */
} uc_attn_bwd_params;
UC_API int uc_attn_bwd(const uc_attn_bwd_params* p, uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Patch-embed front end (libs/croco/patch_embed.py:68-82): gathers non-overlapping p x p patches of
 * an fp32 NCHW image into bf16 rows [B*h*w][3*p*p] (column order c, i, j == conv weight layout),
 * so that the conv is `uc_gemm`.  Bit-exact gather + one bf16 rounding.  Row pitch of `cols_bf16`: 3*p*p when p % 8 == 0,
 * otherwise 3*p*p rounded up to a multiple of 64 elements with zero pad columns (p = 14: 588 -> 640).
 * ------------------------------------------------------------------------------------------ */
UC_API int uc_patchify(const float* img, void* cols_bf16, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Block option flags.
 * qk_norm: `q, k = self.q_norm(q), self.k_norm(k)` with norm_layer(head_dim) on [B,H,N,64] (utils/transformer_blocks.py:199-200,
 * :222, :306-307, :347), followed by the positional encoding (:224-229).  On the packed projection buffer that is a
 * LayerNorm over every 64-column head segment of a row:
 *   fwd: y[r, h*64 + :] = rope2d(LN_64(x[r, h*64 + :]) * gamma + beta)      (RoPE only with positions + rope_table)
 *   bwd: y holds dL/d(normalised, UN-rotated values) on entry (uc_attn_bwd applies the inverse RoPE) and dL/dx on exit;
 *        statistics are recomputed from x; dgamma / dbeta [64] are ACCUMULATED.
 * LayerScale (utils/transformer_blocks.py:389-412, used as x + ls(f(norm(x))), :484-485, :643-646):
 *   fwd: out = res + gamma[c] * z (res optional);  bwd: dz = gamma[c] * dy, dgamma[c] += sum_rows dy * z.  bf16 [rows][cols].
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x;      /* bf16 [rows][ldx]: the raw projection (q or k columns of a packed buffer) */
  void* y;            /* bf16 [rows][ldy]: fwd output / bwd gradient (in place) */
  int64_t ldx, ldy;
  const float* gamma; /* [64] */
  const float* beta;  /* [64] (fwd) */
  float* dgamma;      /* [64] fp32, accumulated (bwd) */
  float* dbeta;       /* [64] fp32, accumulated (bwd) */
  const int32_t* positions; /* optional [rows][2] int32 (y,x) (fwd) */
  const float* rope_table;  /* optional [P][16][2] from uc_rope2d_table (fwd) */
  int32_t rows, heads;
  float eps;
} uc_headnorm_params;
UC_API int uc_headnorm_fwd(const uc_headnorm_params* p, uc_stream_t stream);
UC_API int uc_headnorm_bwd(const uc_headnorm_params* p, uc_stream_t stream);
UC_API int uc_layerscale_fwd(const void* z, const void* res, const float* gamma, void* out, int32_t rows, int32_t cols, uc_stream_t stream);
UC_API int uc_layerscale_bwd(const void* dy, const void* z, const float* gamma, void* dz, float* dgamma, int32_t rows, int32_t cols,
                             uc_stream_t stream);

/* Row softmax of a score matrix produced by uc_gemm and its backward: the un-fused attention of head dims other than 64
 * (DiffAttention family: 128-wide self-attention heads, 64-wide q/k against 128-wide v; utils/transformer_blocks.py:686-945,
 * info_sharing/diff_cross_attention_transformer.py:110-113).  Columns >= valid (key padding up to the GEMM's multiple of 64)
 * are written as zeros.   fwd: P = softmax(scale * S[:, :valid]);   bwd: dS = scale * P o (dP - rowsum(P o dP)). */
UC_API int uc_softmax_rows_fwd(const float* s, void* p_bf16, int32_t rows, int32_t valid, int32_t ld, float scale, uc_stream_t stream);
UC_API int uc_softmax_rows_bwd(const void* p_bf16, const float* dp, void* ds_bf16, int32_t rows, int32_t valid, int32_t ld, float scale,
                               uc_stream_t stream);

/* column sums of a [rows][cols] matrix (bias gradients), ACCUMULATED into fp32 out[cols] */
UC_API int uc_colsum(const void* x, int32_t x_dtype, int64_t ld, int32_t rows, int32_t cols, float* out, uc_stream_t stream);

/* fp32 -> bf16 cast (weights), optionally transposing a [rows][cols] matrix */
UC_API int uc_cast_bf16(const float* src, void* dst, int64_t n, uc_stream_t stream);
/* bf16 [rows][C] <-> fp32 NCHW [B][C][hw] layout converters used at module boundaries
 * (encoders/croco.py:177-180; info_sharing/cross_attention_transformer.py:222-225, :270-273) */
UC_API int uc_nlc_to_nchw(const void* src, int32_t src_dtype, float* dst, int32_t B, int32_t L, int32_t C, uc_stream_t stream);
UC_API int uc_nchw_to_nlc(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t L, int32_t C, uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Linear-head post-processing: pixel_shuffle(p) (prediction_heads/linear.py:81-82, bit-exact gather)
 * fused with PointMapWithConfidenceAdaptor in exp/exp mode (prediction_heads/adaptors.py:337-342,
 * :1080-1083) and the BCHW->BHWC permutes of factory/dust3r.py:323-330.
 *   y  [B*h*w][4*p*p] fp32 (head GEMM output, column = c*p*p + i*p + j)
 *   pts [B][h*p][w*p][3], conf [B][h*p][w*p][1] fp32
 * bwd: dy from (dpts, dconf) and y.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* y;
  float* pts;
  float* conf;
  int32_t B, h, w, patch;
  float conf_min, conf_max; /* confidence = conf_min + min(exp(x), conf_max - conf_min) */
  int64_t ldy;              /* row pitch of y in elements; 0 = 4*patch*patch (dense).  patch = 1 + ldy = 64 serves the DPT head */
} uc_head_post_fwd_params;
UC_API int uc_head_post_fwd(const uc_head_post_fwd_params* p, uc_stream_t stream);

typedef struct {
  const float* y;
  const float* dpts;
  const float* dconf;
  void* dy; /* [B*h*w][4*p*p] */
  int32_t dy_dtype;
  int32_t B, h, w, patch;
  float conf_min, conf_max;
  int64_t ldy; /* row pitch of y AND dy; 0 = dense */
} uc_head_post_bwd_params;
UC_API int uc_head_post_bwd(const uc_head_post_bwd_params* p, uc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DPT head support (prediction_heads/dpt.py:94-311, libs/croco/dpt_block.py:114-255).  Feature maps are NHWC
 * bf16 (== token-major [B*H*W, C], C % 8 == 0), so every convolution is uc_gemm:
 *   3x3 conv, pad 1, stride 1|2 : uc_im2col3x3 -> uc_gemm;  dgrad: uc_gemm -> uc_col2im3x3;  wgrad: uc_gemm on the columns
 *   ConvTranspose2d k == s      : uc_gemm to [(i,j,co)] columns -> uc_depth_space(to_space=1); bwd: to_space=0 -> uc_gemm
 *   F.interpolate(bilinear, align_corners=True) : uc_bilinear_fwd / uc_bilinear_bwd
 *   uc_elementwise op 0: a+b  1: relu(a)  2: a*(b>0)  3: a+b+c
 * ------------------------------------------------------------------------------------------ */
/* Patch embedding without a column buffer == PatchEmbedDust3R.proj + flatten(2).transpose(1, 2)
 * (libs/croco/patch_embed.py:47, :77-80): Conv2d(3, n, kernel = stride = patch) + bias on the fp32 NCHW image, tokens out
 * (bf16 [B * (H/patch) * (W/patch), n], row = (b, patch y, patch x)).  The A operand is fetched by 5-D TMA boxes
 * (dx, dy, patch x, patch y, 3 b + channel) straight from the image and multiplied in TF32 with the fp32 weight
 * [n, 3 * patch * patch] (= proj.weight.flatten(1)); patch 16 or 32, W % 4 == 0, n % 128 == 0.  Other patch sizes (14:
 * 56-byte patch rows cannot be TMA boxes) and the ManyAR transposed samples use uc_patchify + uc_gemm. */
typedef struct {
  const float* img;  /* fp32 [B, 3, H, W] */
  const float* w;    /* fp32 [n, 3 * patch * patch] */
  const float* bias; /* fp32 [n] or NULL */
  void* out;         /* bf16 [B * (H / patch) * (W / patch), n] */
  int32_t B, H, W, patch, n;
} uc_patch_embed_params;
UC_API int uc_patch_embed(const uc_patch_embed_params* p, uc_stream_t stream);

/* 3x3 / stride 1 / pad 1 convolution on NHWC bf16 maps as an implicit GEMM (replaces nn.Conv2d(k=3, padding=1) of
 * dpt_block.py:133-152 (ResidualConvUnit), prediction_heads/dpt.py:131-140 (layer_rn) and :262-283 (regression head) and their
 * autograd backward).  No column buffer: per output tile the 9 taps are 9 shifted 4-D TMA boxes of the map (out-of-image
 * pixels zero-fill = the padding) accumulated in TMEM by the CTA-pair tcgen05 kernel of uc_gemm; outputs leave through 4-D TMA
 * stores clipped at the image border.  Weight layout [cout, 9 * cin], k = (r * 3 + t) * cin + ci  (= conv.weight.permute(0,2,3,1)).
 *   mode 0  y  = conv(x, w) [+ bias] [ReLU | + residual]                 x [B*H*W, cin]  -> y  [B*H*W, cout]
 *   mode 1  dx = conv(dy, flipped w^T) [* (relu_out > 0)]                dy [B*H*W, cout] -> dx [B*H*W, cin]
 *   mode 2  dw += dy^T * patches(x)   (fp32, accumulated, split over the pixels with TMA reduce-add)   dw [cout, 9 * cin]
 * Channel counts are multiples of 64 (the engine pads); stride-2 convolutions keep the uc_im2col3x3 path. */
typedef struct {
  int32_t mode;
  int32_t B, H, W;
  int32_t cin, cout;
  const void* x;        /* mode 0, 2 */
  const void* w;        /* mode 0, 1: bf16 [cout, 9 * cin] */
  void* y;              /* mode 0 */
  const void* dy;       /* mode 1, 2 */
  void* dx;             /* mode 1 */
  float* dw;            /* mode 2: fp32 [cout, 9 * cin], accumulated */
  const float* bias;    /* mode 0, optional [cout] */
  const void* residual; /* mode 0, optional bf16 [B*H*W, cout] */
  const void* relu_out; /* mode 1, optional bf16 [B*H*W, cin]: the ReLU output whose input gradient dx is */
  int32_t relu;         /* mode 0: fused ReLU */
} uc_conv3x3_params;
UC_API int uc_conv3x3(const uc_conv3x3_params* p, uc_stream_t stream);
UC_API int uc_im2col3x3(const void* x, void* cols, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride, uc_stream_t stream);
UC_API int uc_col2im3x3(const void* dcols, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride, uc_stream_t stream);
UC_API int uc_depth_space(const void* src, void* dst, int32_t B, int32_t h, int32_t w, int32_t C, int32_t s, int32_t to_space,
                          uc_stream_t stream);
UC_API int uc_bilinear_fwd(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                           uc_stream_t stream);
/* the same resampling from an fp32 map (bf16 out): the 1x1 `out_conv` of a fusion block commutes with the interpolation that
 * precedes it (dpt_block.py:251-255; both are linear and the interpolation weights sum to 1), so the engine runs the conv at
 * the LOW resolution with fp32 output and resamples that -- a quarter of the conv's FLOPs and one bf16 rounding less */
UC_API int uc_bilinear_fwd_f32in(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                                 uc_stream_t stream);
UC_API int uc_bilinear_bwd(const void* dout, void* din, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                           uc_stream_t stream);
UC_API int uc_elementwise(int32_t op, const void* a, const void* b, const void* c, void* out, int64_t n, uc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UC_B200_H */
