// C ABI of libdenseclip_b200.so (see include/denseclip_b200.h).  Thin, exception-free boundary over the kernels in
// gemm_tcgen05.cuh / attn_tcgen05.cuh / rowwise.cuh, plus the CLIPVisionTransformer forward as one native call.
#include "../../include/denseclip_b200.h"

#include <map>
#include <memory>
#include <vector>

#include "host_utils.cuh"
#include "rowwise.cuh"
#include "train_tail.cuh"
#include "vit_encoder.cuh"

using namespace dclip;

struct dclip_handle_s {
  int device = 0;
  std::string err;
  long long launches = 0;
  // plan caches: tensor maps are encoded once per distinct argument set
  std::map<std::string, GemmPlan> gemm_plans;
  std::map<std::string, AttnPlan> attn_plans;
  std::map<std::string, AttnSplitPlan> attn_split_plans;
  std::map<void*, AttnSmallScratch> attn_small_scratch;  // key-split partials, one per stream (never freed before destroy)
};

struct dclip_vit_s {
  dclip_handle_t h;
  VitEncoder enc;
};

static thread_local std::string g_create_err;

// Runs f with the handle's device current and RESTORES the caller's device afterwards (the host framework's notion of
// the current device must not change behind its back).
template <class F>
static int guarded(dclip_handle_t h, F&& f) {
  int cur = -1;
  bool switched = false;
  int rc = 0;
  try {
    if (!h) throw Error{"null handle"};
    DCLIP_CHECK_CUDA(cudaGetDevice(&cur));
    if (cur != h->device) {
      DCLIP_CHECK_CUDA(cudaSetDevice(h->device));
      switched = true;
    }
    f();
  } catch (const Error& e) {
    if (h) h->err = e.msg; else g_create_err = e.msg;
    rc = 1;
  } catch (const std::exception& e) {
    if (h) h->err = e.what(); else g_create_err = e.what();
    rc = 2;
  }
  if (switched) cudaSetDevice(cur);
  return rc;
}

template <class T>
static std::string key_of(const T& v) {
  return std::string(reinterpret_cast<const char*>(&v), sizeof(T));
}

static void check_launch(dclip_handle_t h, int n = 1) {
  DCLIP_CHECK_CUDA(cudaGetLastError());
  h->launches += n;
}

extern "C" {

int dclip_abi_version(void) { return DCLIP_ABI_VERSION; }
size_t dclip_sizeof_gemm_args(void) { return sizeof(dclip_gemm_args); }

int dclip_create(int device, dclip_handle_t* out) {
  try {
    if (!out) throw Error{"dclip_create: out is null"};
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) throw Error{std::string("dclip_create: no CUDA device (") + cudaGetErrorString(e) + ")"};
    if (device < 0 || device >= n) throw Error{"dclip_create: bad device index"};
    cudaDeviceProp prop;
    DCLIP_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
      char buf[160];
      snprintf(buf, sizeof(buf), "dclip_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
               prop.major, prop.minor);
      throw Error{buf};
    }
    int cur = 0;
    DCLIP_CHECK_CUDA(cudaGetDevice(&cur));
    DCLIP_CHECK_CUDA(cudaSetDevice(device));
    get_encode_fn();
    DCLIP_CHECK_CUDA(cudaSetDevice(cur));
    auto* h = new dclip_handle_s;
    h->device = device;
    *out = h;
    return 0;
  } catch (const Error& e) {
    g_create_err = e.msg;
    return 1;
  }
}

int dclip_destroy(dclip_handle_t h) {
  if (h) {
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(h->device);
    for (auto& kv : h->attn_small_scratch) kv.second.release();
    cudaSetDevice(cur);
  }
  delete h;
  return 0;
}

const char* dclip_last_error(dclip_handle_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }
long long dclip_launch_count(dclip_handle_t h) { return h ? h->launches : 0; }
int dclip_reset_launch_count(dclip_handle_t h) {
  if (h) h->launches = 0;
  return 0;
}

int dclip_gemm(dclip_handle_t h, const dclip_gemm_args* a, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(a != nullptr, "null args");
    const std::string key = key_of(*a);
    auto it = h->gemm_plans.find(key);
    if (it == h->gemm_plans.end()) {
      GemmOperands op{static_cast<const __nv_bfloat16*>(a->A), int(a->lda), static_cast<const __nv_bfloat16*>(a->W), int(a->ldw)};
      op.a_bs = a->a_bs; op.conv_B = a->conv_B; op.conv_gh = a->conv_gh; op.a_gs = a->a_gs;
      GemmParams p{};
      p.M = a->M; p.N = a->N; p.K = a->K; p.split_in = a->split_in;
      p.bias = a->bias; p.act = a->act; p.out_scale = a->out_scale;
      p.residual = a->residual; p.ldr = int(a->ldr); p.res_mod = a->res_mod;
      p.remap_P = a->remap_P; p.remap_Nt = a->remap_Nt;
      p.out_f32 = a->out_f32; p.ldc = int(a->ldc);
      p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); p.ldcb = int(a->ldcb);
      p.split_out = a->split_out; p.split_out_off = int(a->split_out_off);
      if (a->conv_C > 0) {
        p.conv_C = a->conv_C; p.conv_gw = a->conv_gw;
        p.conv_tiles_per_img = a->conv_gh * a->conv_gw / 128;
        p.conv_G = a->conv_G;
      }
      p.wg_C = a->wg_C; p.wg_pitch = a->wg_pitch; p.wg_grouped = a->wg_grouped; p.wg_rows = a->wg_rows;
      DCLIP_REQUIRE(p.out_f32 || p.out_bf16, "GEMM needs at least one output");
      if (h->gemm_plans.size() > 4096) h->gemm_plans.clear();
      it = h->gemm_plans.emplace(key, make_gemm_plan(op, p, a->block_n)).first;
    }
    run_gemm(it->second, static_cast<cudaStream_t>(stream));
    h->launches += 1;
  });
}

int dclip_layernorm(dclip_handle_t h, const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                    int M, int D, float* out_f32, long long ldo, void* out_bf16, long long ldb, int split,
                    long long split_off, void* stream) {
  return guarded(h, [&] {
    LayerNormParams p{x, ldx, gamma, beta, eps, M, D, out_f32, ldo, static_cast<__nv_bfloat16*>(out_bf16), ldb, split, int(split_off)};
    launch_layernorm(p, static_cast<cudaStream_t>(stream));
    check_launch(h);
  });
}

int dclip_cast_bf16(dclip_handle_t h, const float* x, long long ldx, void* out, long long ldo, int rows, int cols,
                    int split, long long split_off, float scale, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(cols % 2 == 0 && ldx % 2 == 0 && ldo % 2 == 0 && split_off % 2 == 0, "cast: even cols/ld required");
    CastParams p{x, ldx, static_cast<__nv_bfloat16*>(out), ldo, rows, cols, split, int(split_off), scale};
    const long long total = (long long)rows * (cols / 2);
    const int grid = int(std::min<long long>((total + 255) / 256, 148 * 16));
    cast_bf16_kernel<<<std::max(grid, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_attention(dclip_handle_t h, const void* q, const void* k, const void* v, long long ldq, long long ldk,
                    long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0, int k_col0, int v_col0,
                    int B, int H, int Nq, int q_start, int Nk, float scale, void* out, long long ldo, long long out_bs,
                    void* stream) {
  return guarded(h, [&] {
    // the plan-cache key is the raw bytes of (op, p): zero them first so the padding bytes are deterministic
    AttnOperands op;
    memset(&op, 0, sizeof(op));
    op.q = static_cast<const __nv_bfloat16*>(q); op.k = static_cast<const __nv_bfloat16*>(k); op.v = static_cast<const __nv_bfloat16*>(v);
    op.ldq = int(ldq); op.ldk = int(ldk); op.ldv = int(ldv); op.q_bs = q_bs; op.k_bs = k_bs; op.v_bs = v_bs; op.Nq_total = Nq;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.Nq_total = Nq; p.q_start = q_start; p.Nk = Nk;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out = static_cast<__nv_bfloat16*>(out); p.out_batch_stride = out_bs; p.ldo = int(ldo);
    const std::string key = key_of(op) + key_of(p);
    auto it = h->attn_plans.find(key);
    if (it == h->attn_plans.end()) {
      if (h->attn_plans.size() > 1024) h->attn_plans.clear();
      it = h->attn_plans.emplace(key, make_attn_plan(op, p)).first;
    }
    run_attn(it->second, static_cast<cudaStream_t>(stream));
    h->launches += 1;
  });
}

int dclip_attention_small(dclip_handle_t h, const void* q, const void* k, const void* v, int is_f32, long long ldq,
                          long long ldk, long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0,
                          int k_col0, int v_col0, int B, int H, int q_first, int q_count, int Nk, float scale, int causal,
                          void* out, int out_f32, long long ldo, long long out_bs, long long out_split_off, void* stream) {
  return guarded(h, [&] {
    SmallAttnParams p{};
    p.q = q; p.k = k; p.v = v; p.is_f32 = is_f32; p.B = B; p.H = H; p.Nk = Nk; p.q_first = q_first; p.q_count = q_count;
    p.ldq = int(ldq); p.ldk = int(ldk); p.ldv = int(ldv); p.q_bs = q_bs; p.k_bs = k_bs; p.v_bs = v_bs;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.scale = scale; p.causal = causal;
    p.out = out; p.out_f32 = out_f32; p.ldo = int(ldo); p.out_bs = out_bs; p.out_split_off = int(out_split_off);
    const int align = is_f32 ? 4 : 8;
    DCLIP_REQUIRE(ldk % align == 0 && ldv % align == 0 && k_col0 % align == 0 && v_col0 % align == 0 && k_bs % align == 0 &&
                      v_bs % align == 0, "small attention: K/V rows must be 16B aligned");
    h->launches += run_attn_small(p, static_cast<cudaStream_t>(stream), &h->attn_small_scratch[stream]);
  });
}

int dclip_attention_split(dclip_handle_t h, const void* q, const void* k, const void* v, long long ldq, long long ldk,
                          long long ldv, long long q_bs, long long k_bs, long long v_bs, int q_col0, int k_col0, int v_col0,
                          long long lo_off, int B, int H, int Nq, int Nk, float scale, void* out, long long ldo, long long out_bs,
                          long long out_lo_off, void* stream) {
  return guarded(h, [&] {
    AttnSplitOperands op;
    memset(&op, 0, sizeof(op));
    op.q = static_cast<const __nv_bfloat16*>(q); op.k = static_cast<const __nv_bfloat16*>(k); op.v = static_cast<const __nv_bfloat16*>(v);
    op.ldq = int(ldq); op.ldk = int(ldk); op.ldv = int(ldv); op.q_bs = q_bs; op.k_bs = k_bs; op.v_bs = v_bs;
    AttnSplitParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.lo_off = int(lo_off);
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out = static_cast<__nv_bfloat16*>(out); p.out_bs = out_bs; p.ldo = int(ldo); p.out_lo_off = int(out_lo_off);
    const std::string key = key_of(op) + key_of(p);
    auto it = h->attn_split_plans.find(key);
    if (it == h->attn_split_plans.end()) {
      if (h->attn_split_plans.size() > 1024) h->attn_split_plans.clear();
      it = h->attn_split_plans.emplace(key, make_attn_split_plan(op, p)).first;
    }
    run_attn_split(it->second, static_cast<cudaStream_t>(stream));
    h->launches += 1;
  });
}

int dclip_im2col_patches(dclip_handle_t h, const float* img, int B, int H, int W, int ps, void* out, long long lda,
                         int split, long long split_off, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(ps > 0 && ps % 2 == 0 && H >= ps && W >= ps, "im2col: even patch size <= image required");
    Im2colParams p{img, static_cast<__nv_bfloat16*>(out), lda, B, H, W, ps, H / ps, W / ps, split, int(split_off)};
    launch_im2col(p, static_cast<cudaStream_t>(stream));
    check_launch(h);
  });
}

int dclip_posemb_interp(dclip_handle_t h, const float* pos, int g0, int gh, int gw, int D, float* out, void* stream) {
  return guarded(h, [&] {
    posemb_interp_kernel<<<1 + gh * gw, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, out, g0, gh, gw, D);
    check_launch(h);
  });
}

int dclip_tap_nchw(dclip_handle_t h, const float* tokens, int B, int Ntok, int D, float* out_nchw, void* stream) {
  return guarded(h, [&] {
    launch_tap_nchw(tokens, out_nchw, B, Ntok, D, static_cast<cudaStream_t>(stream));
    check_launch(h);
  });
}

int dclip_nchw_to_tokens(dclip_handle_t h, const float* in_nchw, int B, int C, int P, float* out_f32, void* out_bf16,
                         long long ld, long long out_bs, int row_off, void* stream) {
  return guarded(h, [&] {
    dim3 grid((P + 31) / 32, (C + 31) / 32, B);
    nchw_to_tokens_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_nchw, out_f32, static_cast<__nv_bfloat16*>(out_bf16),
                                                                               C, P, ld, out_bs, row_off);
    check_launch(h);
  });
}

int dclip_token_mean(dclip_handle_t h, const float* x, int B, int row0, int P, long long ld, long long bs, int D,
                     float* out, void* stream) {
  return guarded(h, [&] {
    dim3 grid((D + 31) / 32, B);
    token_mean_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, row0, P, ld, bs, D);
    check_launch(h);
  });
}

int dclip_score_map(dclip_handle_t h, const float* vis, long long ld, long long bs, int row0, const float* text, int B,
                    int K, int C, int P, float eps, float* score, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(C % 128 == 0 && C <= 1024, "score map: C=%d must be a multiple of 128 and <= 1024", C);
    DCLIP_REQUIRE(ld % 4 == 0 && bs % 4 == 0 && (reinterpret_cast<uintptr_t>(vis) & 15) == 0, "score map: vis rows must be 16B aligned");
    const size_t smem = size_t(K) * C * 4;
    DCLIP_REQUIRE(smem <= 200 * 1024, "score map: K*C too large for shared memory");
    ScoreParams p{vis, ld, bs, row0, text, score, B, K, C, P, eps};
    // 32 pixels per block and trip (8 warps x 4 pixels); ~4 blocks per SM over all images, so the per-block
    // staging of the image's text rows (K*C*4 bytes) is amortised over as many pixels as possible
    DCLIP_REQUIRE((reinterpret_cast<uintptr_t>(text) & 15) == 0, "score map: text must be 16B aligned");
    const int per_img = std::max(1, std::min((P + 31) / 32, (4 * sm_count() + B - 1) / B));
    dim3 grid(per_img, B);
    auto launch = [&](auto kern) {
      ensure_dyn_smem(kern, 200 * 1024);
      kern<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
    };
    switch (C / 128) {
      case 1: launch(score_map_kernel<1>); break;
      case 2: launch(score_map_kernel<2>); break;
      case 4: launch(score_map_kernel<4>); break;
      case 6: launch(score_map_kernel<6>); break;
      case 8: launch(score_map_kernel<8>); break;
      default: throw Error{"score map: C must be one of 128, 256, 512, 768, 1024"};
    }
    check_launch(h);
  });
}

int dclip_upsample_bilinear(dclip_handle_t h, const float* in, int in_nchw, long long ldi, long long in_bs, int B, int C,
                            int hh, int ww, int H, int W, float* out, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(W % 4 == 0, "upsample: output width must be a multiple of 4");
    UpsampleParams p{in, in_nchw, ldi, in_bs, out, B, C, hh, ww, H, W};
    // token-major input whose rows hold (C rounded up to 4) readable floats: all-channel fast path
    if (!in_nchw && ldi % 4 == 0 && ldi >= ((C + 3) & ~3) && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && in_bs % 4 == 0) {
      const long long total = (long long)B * ((H + UP_ROWS - 1) / UP_ROWS) * (W / 4);
      const int grid = int(std::min<long long>((total + 255) / 256, 148 * 64));
      upsample_bilinear_tok_kernel<<<std::max(grid, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    } else {
      const long long total = (long long)B * C * H * (W / 4);
      const int grid = int(std::min<long long>((total + 255) / 256, 148 * 32));
      upsample_bilinear_kernel<<<std::max(grid, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    }
    check_launch(h);
  });
}

int dclip_upsample_argmax(dclip_handle_t h, const float* in, long long ldi, long long in_bs, int B, int K, int hh, int ww,
                          int H, int W, uint8_t* out, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(W % 4 == 0 && K > 0 && K <= 256, "upsample_argmax: W %% 4 == 0 and K <= 256 required");
    UpsampleArgmaxParams p{in, ldi, in_bs, out, B, K, hh, ww, H, W};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int k4 = (K + 3) / 4;
    if (k4 <= 5 && ldi % 4 == 0 && ldi >= 4 * k4 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && in_bs % 4 == 0) {
      // class scores cached in registers per thread strip (the 19-class Cityscapes head: K4 = 5)
      const long long total = (long long)B * ((H + UP_ROWS - 1) / UP_ROWS) * (W / 4);
      const int grid = std::max(int(std::min<long long>((total + 255) / 256, 148 * 64)), 1);
      switch (k4) {
        case 1: upsample_argmax_strip_kernel<1><<<grid, 256, 0, st>>>(p); break;
        case 2: upsample_argmax_strip_kernel<2><<<grid, 256, 0, st>>>(p); break;
        case 3: upsample_argmax_strip_kernel<3><<<grid, 256, 0, st>>>(p); break;
        case 4: upsample_argmax_strip_kernel<4><<<grid, 256, 0, st>>>(p); break;
        default: upsample_argmax_strip_kernel<5><<<grid, 256, 0, st>>>(p); break;
      }
    } else {
      const long long total = (long long)B * H * (W / 4);
      const int grid = int(std::min<long long>((total + 255) / 256, 148 * 32));
      upsample_argmax_kernel<<<std::max(grid, 1), 256, 0, st>>>(p);
    }
    check_launch(h);
  });
}

int dclip_eval_stats(dclip_handle_t h, const uint8_t* pred, const void* target, int target_is_i64, long long n, int K,
                     int ignore_index, const float* depth_pred, const float* depth_gt, const uint8_t* depth_mask,
                     long long n_depth, long long* conf, double* depth_stats, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(pred || depth_pred, "eval_stats: nothing to do");
    if (pred) DCLIP_REQUIRE(target && conf && K > 0 && K <= 64 && n >= 0, "eval_stats: segmentation needs target, conf and 0 < K <= 64");
    if (depth_pred) DCLIP_REQUIRE(depth_gt && depth_stats && n_depth >= 0, "eval_stats: depth needs depth_gt and depth_stats");
    EvalStatsParams p{pred, target, target_is_i64, n, pred ? K : 1, ignore_index, depth_pred, depth_gt, depth_mask, n_depth,
                      reinterpret_cast<unsigned long long*>(conf), depth_stats};
    const long long work = std::max(pred ? n : 0, depth_pred ? n_depth : 0);
    const int grid = int(std::max<long long>(1, std::min<long long>((work + 256 * 16 - 1) / (256 * 16), 148 * 8)));
    eval_stats_kernel<<<grid, 256, size_t(p.K) * p.K * sizeof(unsigned int), static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_gamma_residual(dclip_handle_t h, const float* a, const float* gamma, const float* d, float* out, long long n,
                         int C, void* stream) {
  return guarded(h, [&] {
    gamma_residual_kernel<<<int((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, gamma, d, out, n, C);
    check_launch(h);
  });
}

int dclip_conv3x3_gather(dclip_handle_t h, const void* in, int in_f32, long long ld, long long bs, int row0, int B, int hh,
                         int ww, int C, void* out, long long ldo, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(C % 2 == 0 && ldo % 2 == 0, "conv gather: even C required");
    Conv3x3GatherParams p{in, in_f32, ld, bs, row0, B, hh, ww, C, static_cast<__nv_bfloat16*>(out), ldo};
    const long long total = (long long)B * hh * ww * 9 * (C / 2);
    const int grid = int(std::min<long long>((total + 255) / 256, 148 * 32));
    conv3x3_gather_kernel<<<std::max(grid, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

// ------------------------------------------------------------------------------------------------------------
// ViT encoder
// ------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// training mode of the trainable tail (train_tail.cuh)
// ---------------------------------------------------------------------------------------------------------
static int col_reduce_blocks(int M) {   // row blocks of the partial pass: few enough that the finishing pass (one thread per column
  int nblk = (M + 255) / 256;            // walking the partials) stays in the microseconds, enough to fill the GPU with N / 32 column groups
  return nblk < 1 ? 1 : (nblk > 64 ? 64 : nblk);
}

size_t dclip_col_reduce_workspace(int M, int N) { return size_t(col_reduce_blocks(M)) * 2 * size_t(N > 0 ? N : 1) * sizeof(double); }

int dclip_col_reduce(dclip_handle_t h, const dclip_col_reduce_args* a, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(a && a->a && a->M > 0 && a->N > 0 && a->out0 && (a->mode == 0 || a->mode == 1), "col_reduce: bad arguments");
    DCLIP_REQUIRE(a->mode == 1 || a->out1, "col_reduce: mode 0 needs out1 (variance)");
    DCLIP_REQUIRE(!(a->mode == 1 && a->x) || (a->mean && a->rstd && a->gamma && a->beta), "col_reduce: mode 1 with x needs mean / rstd / gamma / beta");
    ColReduceParams p{};
    p.a = a->a; p.lda = a->lda; p.x = a->x; p.ldx = a->ldx;
    p.mean = a->mean; p.rstd = a->rstd; p.gamma = a->gamma; p.beta = a->beta;
    p.mask = a->mask; p.ldm = a->ldm; p.mask_scale = a->mask_scale;
    p.relu = a->relu; p.M = a->M; p.N = a->N; p.mode = a->mode;
    p.nblk = col_reduce_blocks(a->M);
    p.rows_per_blk = (a->M + p.nblk - 1) / p.nblk;
    DCLIP_REQUIRE(a->workspace && a->workspace_bytes >= dclip_col_reduce_workspace(a->M, a->N) &&
                      (reinterpret_cast<uintptr_t>(a->workspace) & 7) == 0, "col_reduce: workspace too small or misaligned");
    p.part = static_cast<double*>(a->workspace);
    p.out0 = a->out0; p.out1 = a->out1; p.out2 = a->out2; p.eps = a->eps;
    p.run_mean = a->run_mean; p.run_var = a->run_var; p.momentum = a->momentum;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    col_reduce_partial_kernel<<<dim3((a->N + 31) / 32, p.nblk), 256, 0, st>>>(p);
    col_reduce_final_kernel<<<(a->N + 255) / 256, 256, 0, st>>>(p);
    check_launch(h, 2);
  });
}

int dclip_bn_apply(dclip_handle_t h, const dclip_bn_apply_args* a, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(a && a->M > 0 && a->N > 0 && (a->out_f32 || a->out_bf16) && a->mode >= 0 && a->mode <= 2, "bn_apply: bad arguments");
    DCLIP_REQUIRE(a->mode != 0 || a->x, "bn_apply: mode 0 needs x");
    DCLIP_REQUIRE(a->mode == 0 || a->a, "bn_apply: backward modes need the upstream gradient");
    DCLIP_REQUIRE(a->mode != 1 || (a->x && a->mean && a->rstd && a->gamma && a->beta && a->sum_g && a->sum_gx), "bn_apply: mode 1 needs x, the statistics and both sums");
    DCLIP_REQUIRE(!a->mean || (a->rstd && a->gamma && a->beta), "bn_apply: BatchNorm needs mean, rstd, gamma and beta");
    BnApplyParams p{};
    p.a = a->a; p.lda = a->lda; p.x = a->x; p.ldx = a->ldx;
    p.mean = a->mean; p.rstd = a->rstd; p.gamma = a->gamma; p.beta = a->beta; p.sum_g = a->sum_g; p.sum_gx = a->sum_gx;
    p.mask = a->mask; p.ldm = a->ldm; p.mask_scale = a->mask_scale;
    p.relu = a->relu; p.M = a->M; p.N = a->N; p.mode = a->mode;
    p.out_f32 = a->out_f32; p.ldo = a->ldo; p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16); p.ldb = a->ldb;
    const long long total = (long long)a->M * a->N;
    const int grid = int(std::min<long long>((total + 255) / 256, 16LL * sm_count()));
    bn_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_transpose_pad(dclip_handle_t h, const void* in, int in_f32, long long ld, long long bs, int B, int gh, int gw, int C, int pad,
                        int pitch, int lead, int shift, int nshift, long long plane, void* out_bf16, long long ldk, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(in && out_bf16 && B > 0 && gh > 0 && gw > 0 && C > 0 && (pad == 0 || pad == 1), "transpose_pad: bad arguments");
    DCLIP_REQUIRE(pitch >= gw + pad && lead >= 0 && shift >= -1 && shift <= 1, "transpose_pad: pitch %d < gw + pad, or bad lead / shift", pitch);
    DCLIP_REQUIRE(nshift == 1 || (nshift == 3 && plane >= (long long)C * ldk), "transpose_pad: nshift must be 1, or 3 with plane >= C * ldk");
    const long long K = (long long)B * (gh + pad) * pitch;
    DCLIP_REQUIRE(K < (1ll << 31), "transpose_pad: %lld padded pixels exceed 2^31", K);
    DCLIP_REQUIRE(ldk >= K + lead && ldk % 8 == 0, "transpose_pad: ldk (%lld) must cover lead + %lld padded pixels and be a multiple of 8", ldk, K);
    TransposePadParams p{in, in_f32, ld, bs, B, gh, gw, C, pad, pitch, lead, shift, static_cast<__nv_bfloat16*>(out_bf16), ldk, nshift, plane};
    transpose_pad_kernel<<<dim3(unsigned((ldk + 31) / 32), unsigned((C + 31) / 32)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_upsample_bilinear_bwd(dclip_handle_t h, const float* dout, int B, int K, int H, int W, int gh, int gw, float* dtok,
                                long long ldc, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(dout && dtok && B > 0 && K > 0 && H > 0 && W > 0 && gh > 0 && gw > 0 && ldc >= K, "upsample_bwd: bad arguments");
    DCLIP_REQUIRE(size_t(W) * 4 <= 48 * 1024, "upsample_bwd: W = %d too wide for the shared-memory line", W);
    UpsampleBwdParams p{dout, B, K, H, W, gh, gw, dtok, ldc};
    upsample_bilinear_bwd_kernel<<<unsigned(B) * K * gh, 256, size_t(W) * 4, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

static constexpr int kLossBlocks = 1184;   // 8 x 148
size_t dclip_loss_workspace(void) { return size_t(kLossBlocks) * 3 * sizeof(double); }

int dclip_ce_loss(dclip_handle_t h, const float* logits, const long long* target, int B, int K, long long HW, int ignore_index,
                  void* workspace, float* stats, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(logits && target && workspace && stats && B > 0 && K > 0 && HW > 0, "ce_loss: bad arguments");
    LossParams p{};
    p.pred = logits; p.target = target; p.B = B; p.K = K; p.HW = HW; p.ignore_index = ignore_index;
    p.part = static_cast<double*>(workspace); p.nblk = kLossBlocks; p.stats = stats;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ce_loss_partial_kernel<<<kLossBlocks, 256, 0, st>>>(p);
    loss_final_kernel<<<1, 256, 0, st>>>(p, 0);
    check_launch(h, 2);
  });
}

int dclip_ce_loss_bwd(dclip_handle_t h, const float* logits, const long long* target, int B, int K, long long HW, int ignore_index,
                      const float* stats, const float* gout, float* grad, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(logits && target && stats && gout && grad && B > 0 && K > 0 && HW > 0, "ce_loss_bwd: bad arguments");
    LossParams p{};
    p.pred = logits; p.target = target; p.B = B; p.K = K; p.HW = HW; p.ignore_index = ignore_index;
    p.stats = const_cast<float*>(stats); p.gout = gout; p.grad = grad;
    ce_loss_bwd_kernel<<<kLossBlocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_silog_loss(dclip_handle_t h, const float* pred, const float* target, const uint8_t* mask, long long n, float lambd,
                     float eps, void* workspace, float* stats, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(pred && target && workspace && stats && n > 0, "silog_loss: bad arguments");
    LossParams p{};
    p.pred = pred; p.ftarget = target; p.mask = mask; p.B = 1; p.HW = n; p.lambd = lambd; p.eps = eps;
    p.part = static_cast<double*>(workspace); p.nblk = kLossBlocks; p.stats = stats;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    silog_loss_partial_kernel<<<kLossBlocks, 256, 0, st>>>(p);
    loss_final_kernel<<<1, 256, 0, st>>>(p, 1);
    check_launch(h, 2);
  });
}

int dclip_silog_loss_bwd(dclip_handle_t h, const float* pred, const float* target, const uint8_t* mask, long long n, float lambd,
                         float eps, const float* stats, const float* gout, float* grad, void* stream) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(pred && target && stats && gout && grad && n > 0, "silog_loss_bwd: bad arguments");
    LossParams p{};
    p.pred = pred; p.ftarget = target; p.mask = mask; p.B = 1; p.HW = n; p.lambd = lambd; p.eps = eps;
    p.stats = const_cast<float*>(stats); p.gout = gout; p.grad = grad;
    silog_loss_bwd_kernel<<<kLossBlocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    check_launch(h);
  });
}

int dclip_vit_create(dclip_handle_t h, const dclip_vit_config* cfg, dclip_vit_t* out) {
  return guarded(h, [&] {
    DCLIP_REQUIRE(cfg && out, "null argument");
    auto v = std::make_unique<dclip_vit_s>();
    v->h = h;
    v->enc.configure(VitConfig{cfg->width, cfg->layers, cfg->heads, cfg->patch_size, cfg->grid0, cfg->precise, cfg->ln_fold});
    *out = v.release();
  });
}

int dclip_vit_destroy(dclip_vit_t v) {
  delete v;
  return 0;
}

int dclip_vit_set_weights(dclip_vit_t v, const dclip_vit_weights* w) {
  if (!v) return 1;
  return guarded(v->h, [&] {
    DCLIP_REQUIRE(w != nullptr, "null weights");
    VitWeights& dst = v->enc.weights;
    const int L = v->enc.cfg.layers;
    dst.conv1_w = static_cast<const __nv_bfloat16*>(w->conv1_w);
    dst.class_embedding = w->class_embedding;
    dst.positional_embedding = w->positional_embedding;
    dst.ln_pre_g = w->ln_pre_g; dst.ln_pre_b = w->ln_pre_b; dst.ln_post_g = w->ln_post_g; dst.ln_post_b = w->ln_post_b;
    dst.layers.resize(L);
    for (int i = 0; i < L; ++i) {
      VitLayerWeights& l = dst.layers[i];
      l.ln1_g = w->ln1_g[i]; l.ln1_b = w->ln1_b[i]; l.ln2_g = w->ln2_g[i]; l.ln2_b = w->ln2_b[i];
      l.in_proj_w = static_cast<const __nv_bfloat16*>(w->in_proj_w[i]); l.in_proj_b = w->in_proj_b[i];
      l.out_proj_w = static_cast<const __nv_bfloat16*>(w->out_proj_w[i]); l.out_proj_b = w->out_proj_b[i];
      l.fc_w = static_cast<const __nv_bfloat16*>(w->fc_w[i]); l.fc_b = w->fc_b[i];
      l.proj_w = static_cast<const __nv_bfloat16*>(w->proj_w[i]); l.proj_b = w->proj_b[i];
      l.ln1_c = w->ln1_c ? w->ln1_c[i] : nullptr;
      l.ln2_c = w->ln2_c ? w->ln2_c[i] : nullptr;
    }
    v->enc.invalidate_plans();
  });
}

int dclip_vit_workspace_bytes(dclip_vit_t v, int B, int H, int W, size_t* bytes) {
  if (!v) return 1;
  return guarded(v->h, [&] {
    DCLIP_REQUIRE(bytes != nullptr, "null argument");
    *bytes = v->enc.workspace_bytes(B, H, W);
  });
}

int dclip_vit_forward(dclip_vit_t v, const float* img, int B, int H, int W, void* workspace, size_t workspace_bytes,
                      const dclip_vit_outputs* outs, void* stream) {
  if (!v) return 1;
  return guarded(v->h, [&] {
    DCLIP_REQUIRE(outs != nullptr && img != nullptr && workspace != nullptr, "null argument");
    VitOutputs o;
    o.last_tokens_f32 = outs->last_tokens_f32;
    for (int i = 0; i < outs->n_taps; ++i) {
      VitTap t;
      t.layer = outs->tap_layers[i];
      t.nchw = outs->taps_nchw ? outs->taps_nchw[i] : nullptr;
      t.tokens_bf16 = outs->taps_tokens_bf16 ? static_cast<__nv_bfloat16*>(outs->taps_tokens_bf16[i]) : nullptr;
      o.taps.push_back(t);
    }
    v->h->launches += v->enc.forward(img, B, H, W, workspace, workspace_bytes, o, static_cast<cudaStream_t>(stream));
  });
}

}  // extern "C"
