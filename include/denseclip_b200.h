/*
 * denseclip_b200.h -- C ABI of libdenseclip_b200.so: hand-written sm_100a kernels for DenseCLIP's language-guided
 * dense-prediction forward path.  Plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer
 * unless stated otherwise; `stream` is a cudaStream_t passed as void*.  Every entry point returns 0 on success and a
 * non-zero code on failure; dclip_last_error() returns the message.  No entry point synchronises the device, and all
 * are CUDA-graph capturable once their plans exist (plans = pre-encoded TMA tensor maps, cached per argument set).
 *
 * The reference has no native layer: each entry point replaces stock torch ops on the reference's hot path
 * (reference files are relative to segmentation/denseclip/):
 *
 *   dclip_gemm                nn.Linear / in_proj / out_proj / 1x1 Conv2d          models.py:275-281,287-289; denseclip.py:198,616
 *   dclip_layernorm           LayerNorm (fp32 stats, eps 1e-5)                     models.py:243-249
 *   dclip_attention           softmax(QK^T/sqrt(d))V inside nn.MultiheadAttention  models.py:287-289
 *   dclip_attention_split     the same op in fp32-class precision (3-pass hi|lo split)  models.py:287-289
 *   dclip_attention_small     Attention.forward einsum/softmax/einsum (19 queries), causal text attention  models.py:328-344, 836-842
 *   dclip_im2col_patches      Conv2d(3,D,ps,stride=ps) operand gather               models.py:407,546-548
 *   dclip_posemb_interp       interpolate_pos_encoding                              models.py:514-540
 *   dclip_tap_nchw            LND->NLC, drop CLS, reshape [B,D,H,W]                 models.py:568-582
 *   dclip_token_mean          F.adaptive_avg_pool2d(x,(1,1)).flatten(1)             denseclip.py:596
 *   dclip_score_map           F.normalize x2 + einsum('bchw,bkc->bkhw')             denseclip.py:672-675
 *   dclip_upsample_bilinear   F.interpolate(bilinear, align_corners=False)          denseclip.py:894-916
 *   dclip_conv3x3_gather      3x3/pad-1 conv operand gather (neck, FCN heads)       models.py:13-20,761-782; torchvision FCNHead
 *   dclip_vit_*               CLIPVisionTransformer.forward                         models.py:543-597
 */
#ifndef DENSECLIP_B200_H_
#define DENSECLIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCLIP_ABI_VERSION 1

typedef struct dclip_handle_s* dclip_handle_t;
typedef struct dclip_vit_s* dclip_vit_t;

/* ---- lifetime / errors ------------------------------------------------------------------------------------- */
int dclip_abi_version(void);
int dclip_create(int device, dclip_handle_t* out);
int dclip_destroy(dclip_handle_t h);
/* message of the last failing call on this handle (or of the last failing dclip_create when h == NULL) */
const char* dclip_last_error(dclip_handle_t h);
/* number of kernels launched through this handle since creation / since the last reset (bench.py `gpu_launches`) */
long long dclip_launch_count(dclip_handle_t h);
int dclip_reset_launch_count(dclip_handle_t h);

/* ---- GEMM: C[M,N] = A[M,K] * W[N,K]^T with fused epilogue ---------------------------------------------------- */
enum { DCLIP_ACT_NONE = 0, DCLIP_ACT_QUICKGELU = 1, DCLIP_ACT_QUICKGELU_PRECISE = 2, DCLIP_ACT_GELU_ERF = 3, DCLIP_ACT_RELU = 4 };

typedef struct {
  const void* A;  long long lda;   /* bf16 [M, K] (split_in: [M, 2K] = hi|lo), lda in elements, multiple of 8 */
  const void* W;  long long ldw;   /* bf16 [N, K] (split_in: [N, 2K] = hi|lo) -- torch Linear.weight layout   */
  int M, N, K;                     /* N multiple of 4 */
  int split_in;                    /* 1: three-pass split-bf16 product (fp32-class accuracy).  K % 64 == 0 keeps the halves apart exactly;
                                    * otherwise the last hi K block also multiplies the first 64 - K % 64 lo columns (an extra lo*lo term,
                                    * <= 2^-17 relative: below the split's own truncation error) */
  const float* bias;               /* [N] or NULL */
  int act;                         /* DCLIP_ACT_* applied to (acc + bias) */
  float out_scale;                 /* multiplies the activated value (use 1.0f) */
  const float* residual; long long ldr; /* fp32, added last; may alias out_f32 (in-place residual stream) */
  int res_mod;                     /* with remap_P: residual row = 1 + m % remap_P (positional embedding) */
  int remap_P, remap_Nt;           /* remap_P > 0: output row = (m / P) * Nt + 1 + m % P (patch rows -> token rows) */
  float* out_f32; long long ldc;   /* optional fp32 output */
  void* out_bf16; long long ldcb;  /* optional bf16 output */
  int split_out; long long split_out_off; /* 1: also write lo = bf16(v - hi) at column offset split_out_off */
  int block_n;                     /* 0 = auto, else 64 / 128 / 256 */
  /* implicit 3x3 / pad-1 / stride-1 convolution (conv_C > 0): A points at pixel (0,0) of image 0 of a token-major
   * activation [B][gh*gw][lda] (batch stride a_bs elements, conv_C channels, x2 = hi|lo when split_in); then
   * M = conv_B*conv_gh*conv_gw, K = 9*conv_C and W is [N, 9*C] with K order (ky, kx, c).  Needs gh*gw % 128 == 0 and
   * gw | 128 or 128 | gw; otherwise use dclip_conv3x3_gather + a plain GEMM. */
  int conv_C, conv_gw, conv_gh, conv_B;
  long long a_bs;
  /* grouped conv (conv_G > 1): G activation tensors (element stride a_gs) convolved with G filter banks of block_n
   * outputs each; W is [G*block_n, 9*C], output column g*block_n + n.  One launch for the neck's 12 taps. */
  int conv_G;
  long long a_gs;
  /* weight gradient of a 3x3 / pad-1 conv (wg_C > 0; backward of models.py:717-760 / torchvision FCNHead, the trainable tail
   * of train_denseclip.py:1040-1044): A = dY^T [M filters, K] and W = three copies of X^T, [3][wg_rows channels][K + 2*wg_pitch],
   * channel-major over ZERO-PADDED images (dclip_transpose_pad: row pitch wg_pitch % 8 == 0, one zero row per image; copy kx holds
   * X^T shifted by kx - 1 pixels, with wg_pitch leading zeros).  Then C[m, t*wg_C + c] = sum_k A[m,k] * Wcopy[t%3][c, k + (t/3)*wg_pitch],
   * t = 0..8 = (ky,kx): dW in the (ky, kx, c) order of the forward operand, every TMA box start 16-byte aligned.
   * N = 9*wg_C, wg_C % block_n == 0.  wg_grouped = 1: rows 128g..128g+127 of A pair with rows g*wg_C.. of each copy
   * (G independent convs in one launch, wg_rows = G*wg_C). */
  int wg_C, wg_pitch, wg_grouped, wg_rows;
} dclip_gemm_args;

int dclip_gemm(dclip_handle_t h, const dclip_gemm_args* a, void* stream);
/* sizeof(dclip_gemm_args) as compiled into the library: bindings check their struct mirror against it */
size_t dclip_sizeof_gemm_args(void);

/* ---- LayerNorm over the last dim; D % 128 == 0, D <= 1024 ----------------------------------------------------- */
int dclip_layernorm(dclip_handle_t h, const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                    int M, int D, float* out_f32, long long ldo, void* out_bf16, long long ldb, int split,
                    long long split_off, void* stream);

/* ---- fp32 -> bf16 cast with optional hi|lo split and scale ---------------------------------------------------- */
int dclip_cast_bf16(dclip_handle_t h, const float* x, long long ldx, void* out, long long ldo, int rows, int cols,
                    int split, long long split_off, float scale, void* stream);

/* ---- attention (head_dim 64) ---------------------------------------------------------------------------------- */
/* tensor-core flash attention: q/k/v bf16 token-major [B][N][ld]; head h uses columns col0 + 64h .. +64;
 * computes query rows [q_start, Nq) against all Nk keys; out bf16 [B][Nq][ldo] (column 64h of head h) */
int dclip_attention(dclip_handle_t h, const void* q, const void* k, const void* v, long long ldq, long long ldk,
                    long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0, int k_col0, int v_col0,
                    int B, int H, int Nq, int q_start, int Nk, float scale, void* out, long long ldo, long long out_bs,
                    void* stream);
/* fp32-class tensor-core flash attention for precision="fp32" (same reference op, models.py:287-289): every operand row
 * holds a bf16 hi half and, lo_off columns further, its lo half (x = hi + lo); S = QhKh + QlKh + QhKl and
 * O = PhVh + PlVh + PhVl are three tcgen05 passes each, softmax in fp32.  out bf16 [B][Nq][ldo]: hi at column 64h,
 * lo at out_lo_off + 64h (the hi|lo A operand of the split out-proj GEMM). */
int dclip_attention_split(dclip_handle_t h, const void* q, const void* k, const void* v, long long ldq, long long ldk,
                          long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0, int k_col0, int v_col0,
                          long long lo_off, int B, int H, int Nq, int Nk, float scale, void* out, long long ldo, long long out_bs,
                          long long out_lo_off, void* stream);
/* few-query fp32 attention on CUDA cores; inputs bf16 (is_f32 = 0) or fp32; rows [q_first, q_first + q_count) */
int dclip_attention_small(dclip_handle_t h, const void* q, const void* k, const void* v, int is_f32, long long ldq,
                          long long ldk, long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0,
                          int k_col0, int v_col0, int B, int H, int q_first, int q_count, int Nk, float scale, int causal,
                          void* out, int out_f32, long long ldo, long long out_bs, long long out_split_off, void* stream);

/* ---- ViT front end -------------------------------------------------------------------------------------------- */
int dclip_im2col_patches(dclip_handle_t h, const float* img, int B, int H, int W, int ps, void* out, long long lda,
                         int split, long long split_off, void* stream);
int dclip_posemb_interp(dclip_handle_t h, const float* pos, int g0, int gh, int gw, int D, float* out, void* stream);

/* ---- layout / reductions / tail ------------------------------------------------------------------------------- */
int dclip_tap_nchw(dclip_handle_t h, const float* tokens, int B, int Ntok, int D, float* out_nchw, void* stream);
int dclip_nchw_to_tokens(dclip_handle_t h, const float* in_nchw, int B, int C, int P, float* out_f32, void* out_bf16,
                         long long ld, long long out_bs, int row_off, void* stream);
int dclip_token_mean(dclip_handle_t h, const float* x, int B, int row0, int P, long long ld, long long bs, int D,
                     float* out, void* stream);
int dclip_score_map(dclip_handle_t h, const float* vis, long long ld, long long bs, int row0, const float* text, int B,
                    int K, int C, int P, float eps, float* score, void* stream);
int dclip_upsample_bilinear(dclip_handle_t h, const float* in, int in_nchw, long long ldi, long long in_bs, int B, int C,
                            int hh, int ww, int H, int W, float* out, void* stream);
/* fused bilinear upsample + argmax over K channels: in token-major fp32 [B, hh*ww, ldi] -> uint8 class map [B, H, W]
 * (simple_test's seg_logit.argmax(dim=1), denseclip.py:987-1000, without materialising the logits) */
int dclip_upsample_argmax(dclip_handle_t h, const float* in, long long ldi, long long in_bs, int B, int K, int hh, int ww,
                          int H, int W, uint8_t* out, void* stream);
/* evaluation statistics of one shard, accumulated into caller-zeroed buffers (train_denseclip.py:351-355, 582-593:
 * torchmetrics JaccardIndex / Accuracy with ignore_index, MeanSquaredError(squared=False) over depth_mask):
 * conf[t*K + p] += 1 for every pixel with target t != ignore_index (int64 [K*K], K <= 64; target uint8 or int64);
 * depth_stats[0] += sum (depth_pred - depth_gt)^2 and depth_stats[1] += count over pixels with depth_mask != 0
 * (mask may be NULL = all pixels).  Either half may be skipped by passing pred == NULL / depth_pred == NULL. */
int dclip_eval_stats(dclip_handle_t h, const uint8_t* pred, const void* target, int target_is_i64, long long n, int K,
                     int ignore_index, const float* depth_pred, const float* depth_gt, const uint8_t* depth_mask,
                     long long n_depth, long long* conf, double* depth_stats, void* stream);
int dclip_gamma_residual(dclip_handle_t h, const float* a, const float* gamma, const float* d, float* out, long long n,
                         int C, void* stream);
/* 3x3 / pad 1 / stride 1 conv operand gather: in token-major [B][row0 + y*w + x][ld] (fp32 or bf16, C channels) ->
 * out bf16 [B*h*w, 9*C] with K index = (ky*3 + kx)*C + c; the conv weight must be packed to the same K order */
int dclip_conv3x3_gather(dclip_handle_t h, const void* in, int in_f32, long long ld, long long bs, int row0, int B, int hh,
                         int ww, int C, void* out, long long ldo, void* stream);

/* ---- training mode of the trainable tail -------------------------------------------------------
 * What loss.backward() reaches in the reference's training step (train_denseclip.py:1226-1330; backbone and text tower
 * frozen, :1040-1044; the heads read the neck output of the original backbone features, denseclip.py:755-812):
 * ViTFeatureFusionNeck (models.py:717-782) and the two FCNHeads (denseclip.py:305-349) with BatchNorm on BATCH statistics,
 * the bilinear resize to the ground-truth size (denseclip.py:838, 849) and CE(ignore_index) + SILog (losses.py:21-79).
 * Matrix products go through dclip_gemm (implicit conv forward / input gradient, wg_* weight gradient); these entry points
 * are everything around them.  Activations are token-major fp32 [M = B*gh*gw, N]; reductions are deterministic. */
typedef struct {
  const float* a; long long lda;      /* mode 0: the matrix to reduce; mode 1: upstream gradient */
  const float* x; long long ldx;      /* mode 1: pre-BatchNorm activations (NULL: plain column sums of a = a conv bias gradient) */
  const float *mean, *rstd, *gamma, *beta;             /* mode 1 with x */
  const uint8_t* mask; long long ldm; float mask_scale; /* mode 1: dropout keep-mask and 1/(1-p), or NULL */
  int relu, M, N;
  int mode;      /* 0: nn.BatchNorm2d training statistics: out0 = mean, out1 = biased var, out2 = rstd (optional), running stats updated
                  * 1: out0 = sum_rows g (dbeta), out1 = sum_rows g * xhat (dgamma, optional), g = a [* mask] [* (BN(x) > 0)] */
  void* workspace; size_t workspace_bytes;              /* >= dclip_col_reduce_workspace(M, N) */
  float *out0, *out1, *out2;
  float eps;
  float *run_mean, *run_var; float momentum;            /* mode 0, optional (BatchNorm2d.running_mean / running_var, momentum 0.1) */
} dclip_col_reduce_args;
size_t dclip_col_reduce_workspace(int M, int N);
int dclip_col_reduce(dclip_handle_t h, const dclip_col_reduce_args* a, void* stream);

typedef struct {
  const float* a; long long lda;      /* modes 1, 2: upstream gradient */
  const float* x; long long ldx;      /* pre-BatchNorm activations */
  const float *mean, *rstd, *gamma, *beta;              /* NULL mean: no BatchNorm (y = x) */
  const float *sum_g, *sum_gx;                          /* mode 1: the two dclip_col_reduce(mode 1) outputs */
  const uint8_t* mask; long long ldm; float mask_scale;
  int relu, M, N;
  int mode;      /* 0: y = dropout(relu(gamma*(x-mean)*rstd + beta));  1: BatchNorm(+ReLU, +dropout) input gradient;  2: g only */
  float* out_f32; long long ldo;
  void* out_bf16; long long ldb;
} dclip_bn_apply_args;
int dclip_bn_apply(dclip_handle_t h, const dclip_bn_apply_args* a, void* stream);

/* token-major [B][gh*gw][C] (fp32 or bf16; row pitch ld, image pitch bs, in elements) -> channel-major bf16 [C][ldk] over
 * zero-padded images: out[c][k] = padded[c][k - lead + shift], padded index = (b*(gh+pad) + y)*pitch + x, pitch >= gw + pad,
 * everything that is not a pixel zeroed up to ldk.  The operand layout of the weight-gradient GEMMs (dclip_gemm_args.wg_*):
 * pad = 1, pitch % 8 == 0, lead = pitch and shift = -1 / 0 / +1 for the three copies of X^T of a 3x3 conv (dY^T: lead = shift = 0);
 * pad = 0, pitch = gw for a 1x1 conv (a plain transpose). */
int dclip_transpose_pad(dclip_handle_t h, const void* in, int in_f32, long long ld, long long bs, int B, int gh, int gw, int C, int pad,
                        int pitch, int lead, int shift, int nshift, long long plane, void* out_bf16, long long ldk, void* stream);
/* nshift = 3: ONE read of the input writes the three copies shift = -1, 0, +1 to out_bf16 + j * plane elements (j = 0, 1, 2); `shift` is ignored */
/* backward of F.interpolate(bilinear, align_corners=False): dout NCHW fp32 [B,K,H,W] -> dtok token-major fp32 [B*gh*gw, ldc] */
int dclip_upsample_bilinear_bwd(dclip_handle_t h, const float* dout, int B, int K, int H, int W, int gh, int gw, float* dtok,
                                long long ldc, void* stream);
/* losses: stats = float[4] on the device.  CE: {mean loss over counted pixels, count}; SILog: {loss, T, sum d}.
 * workspace >= dclip_loss_workspace() bytes.  The *_bwd calls read stats and the upstream scalar gradient gout[0] on the device. */
size_t dclip_loss_workspace(void);
int dclip_ce_loss(dclip_handle_t h, const float* logits, const long long* target, int B, int K, long long HW, int ignore_index,
                  void* workspace, float* stats, void* stream);
int dclip_ce_loss_bwd(dclip_handle_t h, const float* logits, const long long* target, int B, int K, long long HW, int ignore_index,
                      const float* stats, const float* gout, float* grad, void* stream);
int dclip_silog_loss(dclip_handle_t h, const float* pred, const float* target, const uint8_t* mask, long long n, float lambd,
                     float eps, void* workspace, float* stats, void* stream);
int dclip_silog_loss_bwd(dclip_handle_t h, const float* pred, const float* target, const uint8_t* mask, long long n, float lambd,
                         float eps, const float* stats, const float* gout, float* grad, void* stream);

/* ---- CLIPVisionTransformer.forward ---------------------------------------------------------------------------- */
typedef struct {
  int width, layers, heads, patch_size, grid0; /* grid0 = input_resolution / patch_size (stored pos-emb grid) */
  int precise;                                 /* 0: bf16 tensor-core path; 1: fp32-class path (3-pass hi|lo split GEMMs and attention) */
  int ln_fold;                                 /* bf16 path: ln_1 / ln_2 folded into the QKV / c_fc GEMMs (see dclip_vit_weights) */
} dclip_vit_config;

/* All weights stay owned by the caller and must outlive the object.  bf16 matrices are [out, in] row-major; in
 * precise mode they are [out, 2*in] = hi|lo halves.  conv1_w is [width, kp] with kp = 3*ps*ps rounded up to 8
 * (to 64 in precise mode: each half of a hi|lo operand is a whole number of 64-column K blocks). */
typedef struct {
  const void* conv1_w;
  const float* class_embedding;       /* [width] */
  const float* positional_embedding;  /* [grid0*grid0 + 1, width] */
  const float* ln_pre_g;  const float* ln_pre_b;
  const float* ln_post_g; const float* ln_post_b;
  const float* const* ln1_g; const float* const* ln1_b;   /* host arrays of `layers` device pointers */
  const float* const* ln2_g; const float* const* ln2_b;
  const void* const* in_proj_w;  const float* const* in_proj_b;   /* [3D, D], [3D] */
  const void* const* out_proj_w; const float* const* out_proj_b;  /* [D, D], [D]   */
  const void* const* fc_w;       const float* const* fc_b;        /* [4D, D], [4D] */
  const void* const* proj_w;     const float* const* proj_b;      /* [D, 4D], [D]  */
  /* ln_fold = 1 (models.py:291-293 ln_1 / ln_2 applied algebraically): in_proj_w / fc_w hold bf16(W * gamma) (gamma of ln_1 /
   * ln_2 along the input dim), in_proj_b / fc_b hold b + W beta, and ln1_c / ln2_c the fp32 row sums of those bf16 folded
   * weights ([3D] / [4D]); then LN(x) W^T + b = rstd (x Wf^T) - rstd mean c + bf.  ln1_g .. ln2_b are unused in this mode. */
  const float* const* ln1_c;     const float* const* ln2_c;
} dclip_vit_weights;

typedef struct {
  int n_taps;
  const int* tap_layers;           /* host array, sorted ascending, unique; ln_post is applied iff layer == layers-1 */
  float* const* taps_nchw;         /* host array of n_taps device pointers: fp32 [B, width, gh, gw] or NULL */
  void* const* taps_tokens_bf16;   /* host array of n_taps device pointers: bf16 [B, 1+gh*gw, width] (CLS row included) or NULL */
  float* last_tokens_f32;          /* optional fp32 [B, 1+gh*gw, width]: ln_post(final layer) token-major */
} dclip_vit_outputs;

int dclip_vit_create(dclip_handle_t h, const dclip_vit_config* cfg, dclip_vit_t* out);
int dclip_vit_destroy(dclip_vit_t v);
int dclip_vit_set_weights(dclip_vit_t v, const dclip_vit_weights* w);
int dclip_vit_workspace_bytes(dclip_vit_t v, int B, int H, int W, size_t* bytes);
int dclip_vit_forward(dclip_vit_t v, const float* img, int B, int H, int W, void* workspace, size_t workspace_bytes,
                      const dclip_vit_outputs* outs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENSECLIP_B200_H_ */
