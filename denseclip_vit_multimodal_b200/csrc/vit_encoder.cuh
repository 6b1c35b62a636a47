// CLIPVisionTransformer.forward (segmentation/denseclip/models.py:543-597) as one native call.
//
// Data layout in HBM (token-major, B images flattened into M = B * Ntok rows; Ntok = 1 + gh*gw, CLS at row 0):
//   x     fp32 [M, D]        residual stream, never rounded to bf16
//   h     bf16 [M, D*s]      LayerNorm output / attention output (A operand of the next GEMM), s = 2 in precise mode (hi|lo)
//   qkv   bf16 [M, 3D]       fused in_proj output, head-interleaved columns [q | k | v]   (precise: bf16 [M, 6D] = hi | lo)
//   g     bf16 [M, 4D*s]     QuickGELU(c_fc) output; the patch im2col matrix aliases it before layer 0
//   pos   fp32 [Ntok, D]     positional embedding interpolated to (gh, gw)
//   lnp   fp32 [M, D]        ln_post output (only when the last layer is tapped)
// Per layer: LN -> QKV GEMM -> flash attention (+ CLS-query side kernel) -> out-proj GEMM (+bias +residual, in place on x)
//            -> LN -> c_fc GEMM (+bias +QuickGELU) -> c_proj GEMM (+bias +residual, in place) -> optional taps.
#pragma once
#include <vector>

#include "host_utils.cuh"
#include "rowwise.cuh"

namespace dclip {

struct VitConfig {
  int width = 768, layers = 12, heads = 12, patch = 16, grid0 = 14, precise = 0;
  // bf16 path only: ln_1 / ln_2 are folded into the QKV / c_fc GEMMs (weights arrive pre-multiplied by gamma, biases with
  // W beta added, ln1_c / ln2_c = row sums of the folded weights); the residual GEMMs publish bf16(x) and per-row statistics.
  // Removes the 2 x layers stand-alone LayerNorm passes (each re-read the 100 MB fp32 stream the epilogue had just written).
  int ln_fold = 0;
};

struct VitLayerWeights {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const __nv_bfloat16 *in_proj_w, *out_proj_w, *fc_w, *proj_w;
  const float *in_proj_b, *out_proj_b, *fc_b, *proj_b;
  const float *ln1_c = nullptr, *ln2_c = nullptr;   // ln_fold: [3D] / [4D] row sums of the folded in_proj / c_fc weights
};

struct VitWeights {
  const __nv_bfloat16* conv1_w = nullptr;
  const float *class_embedding = nullptr, *positional_embedding = nullptr;
  const float *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr, *ln_post_b = nullptr;
  std::vector<VitLayerWeights> layers;
};

struct VitTap {
  int layer;
  float* nchw;
  __nv_bfloat16* tokens_bf16;
};

struct VitOutputs {
  std::vector<VitTap> taps;
  float* last_tokens_f32 = nullptr;
};

class VitEncoder {
 public:
  VitConfig cfg;
  VitWeights weights;

  void configure(const VitConfig& c) {
    DCLIP_REQUIRE(c.width % 128 == 0 && c.width <= 1024, "ViT width %d must be a multiple of 128 and <= 1024", c.width);
    DCLIP_REQUIRE(c.heads * 64 == c.width, "ViT head_dim must be 64 (width %d, heads %d)", c.width, c.heads);
    DCLIP_REQUIRE(c.patch % 2 == 0 && c.layers > 0 && c.grid0 > 0, "bad ViT config");
    cfg = c;
    invalidate_plans();
  }
  void invalidate_plans() { plan_key_ = PlanKey{}; }

  // im2col row pitch (elements, per hi/lo half).  precise: a multiple of the 64-column K block, so the last hi block of the hi|lo
  // operand never reaches into the lo half (patch 14: 588 -> 640; the zero columns cost nothing measurable)
  int kp() const { const int q = cfg.precise ? 64 : 8; return (3 * cfg.patch * cfg.patch + q - 1) / q * q; }

  struct Layout {
    size_t x, h, qkv, g, pos, lnp, xb, xb2, st1, st2, total;
    int gh, gw, P, Ntok, M;
  };
  bool fold() const { return cfg.ln_fold && !cfg.precise; }
  static constexpr int kStatSlots = 12;   // per-row partial-statistics slots: 2 per producer n-block (N = width: <= 6 n-blocks)

  Layout layout(int B, int H, int W) const {
    Layout L{};
    L.gh = H / cfg.patch; L.gw = W / cfg.patch; L.P = L.gh * L.gw; L.Ntok = L.P + 1; L.M = B * L.Ntok;
    const size_t D = cfg.width, s = cfg.precise ? 2 : 1, M = L.M;
    auto up = [](size_t v) { return (v + 1023) & ~size_t(1023); };
    size_t off = 0;
    L.x = off; off += up(M * D * 4);
    L.h = off; off += up(M * D * s * 2);
    L.qkv = off; off += up(M * 3 * D * (cfg.precise ? 4 : 2));
    const size_t g_bytes = M * 4 * D * s * 2, patch_bytes = size_t(B) * L.P * kp() * s * 2;
    L.g = off; off += up(g_bytes > patch_bytes ? g_bytes : patch_bytes);
    L.pos = off; off += up(size_t(L.Ntok) * D * 4);
    L.lnp = off; off += up(M * D * 4);
    if (fold()) {   // bf16 copies of the residual stream (A operands of QKV / c_fc) and the per-row partial statistics
      L.xb = off; off += up(M * D * 2);
      L.xb2 = off; off += up(M * D * 2);
      L.st1 = off; off += up(M * size_t(kStatSlots) * 8);
      L.st2 = off; off += up(M * size_t(kStatSlots) * 8);
    }
    L.total = off;
    return L;
  }

  size_t workspace_bytes(int B, int H, int W) const {
    DCLIP_REQUIRE(B > 0 && H >= cfg.patch && W >= cfg.patch, "bad image shape %dx%dx%d", B, H, W);
    return layout(B, H, W).total;
  }

  // returns the number of kernel launches
  int forward(const float* img, int B, int H, int W, void* ws, size_t ws_bytes, const VitOutputs& outs, cudaStream_t st) {
    DCLIP_REQUIRE(int(weights.layers.size()) == cfg.layers && weights.conv1_w, "ViT weights not set");
    const Layout L = layout(B, H, W);
    DCLIP_REQUIRE(ws_bytes >= L.total, "workspace too small: %zu < %zu", ws_bytes, L.total);
    DCLIP_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024B aligned");
    for (const VitTap& t : outs.taps) DCLIP_REQUIRE(t.layer >= 0 && t.layer < cfg.layers, "tap layer %d out of range", t.layer);
    build_plans(B, H, W, ws, L, outs);
    int n = 0;
    const int D = cfg.width, s = cfg.precise ? 2 : 1;
    uint8_t* w8 = static_cast<uint8_t*>(ws);
    float* x = reinterpret_cast<float*>(w8 + L.x);
    __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(w8 + L.h);
    float* pos = reinterpret_cast<float*>(w8 + L.pos);
    float* lnp = reinterpret_cast<float*>(w8 + L.lnp);
    __nv_bfloat16* patches = reinterpret_cast<__nv_bfloat16*>(w8 + L.g);

    // ---- patch embed + CLS + positional embedding (models.py:546-556) ----
    if (L.P == cfg.grid0 * cfg.grid0) {
      DCLIP_CHECK_CUDA(cudaMemcpyAsync(pos, weights.positional_embedding, size_t(L.Ntok) * D * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      posemb_interp_kernel<<<L.Ntok, 256, 0, st>>>(weights.positional_embedding, pos, cfg.grid0, L.gh, L.gw, D);
      ++n;
    }
    Im2colParams ic{img, patches, (long long)kp() * s, B, H, W, cfg.patch, L.gh, L.gw, cfg.precise, kp()};
    if (kp() != 3 * cfg.patch * cfg.patch) DCLIP_CHECK_CUDA(cudaMemsetAsync(patches, 0, size_t(B) * L.P * kp() * s * 2, st));
    launch_im2col(ic, st); ++n;
    run_gemm(patch_plan_, st); ++n;
    cls_row_kernel<<<(B * D + 255) / 256, 256, 0, st>>>(x, weights.class_embedding, pos, B, L.Ntok, D); ++n;
    // ---- ln_pre (in place on the residual stream) ----
    LayerNormParams lp{x, D, weights.ln_pre_g, weights.ln_pre_b, 1e-5f, L.M, D, x, D, nullptr, 0, 0, 0};
    if (fold()) {   // also the bf16 copy and the row statistics of its output: layer 0's QKV normalises in its epilogue
      lp.out_bf16 = reinterpret_cast<__nv_bfloat16*>(w8 + L.xb); lp.ldb = D;
      lp.stats_out = reinterpret_cast<float2*>(w8 + L.st1); lp.stats_ld = L.M; lp.stats_n = 1;
    }
    launch_layernorm(lp, st); ++n;

    size_t tap_i = 0;
    for (int li = 0; li < cfg.layers; ++li) {
      const VitLayerWeights& lw = weights.layers[li];
      if (!fold()) {
        LayerNormParams l1{x, D, lw.ln1_g, lw.ln1_b, 1e-5f, L.M, D, nullptr, 0, hbuf, (long long)D * s, cfg.precise, D};
        launch_layernorm(l1, st); ++n;
      }
      run_gemm(layer_plans_[li].qkv, st); ++n;
      if (cfg.precise) {
        run_attn_split(attn_split_plan_, st); ++n;   // 3-pass hi|lo tensor-core attention, writes hbuf as hi | lo
      } else {
        run_attn(attn_plan_, st); ++n;
        if (attn_plan_.p.q_start == 1) { run_attn_small(cls_attn_, st); ++n; }
      }
      run_gemm(layer_plans_[li].out_proj, st); ++n;
      if (!fold()) {
        LayerNormParams l2{x, D, lw.ln2_g, lw.ln2_b, 1e-5f, L.M, D, nullptr, 0, hbuf, (long long)D * s, cfg.precise, D};
        launch_layernorm(l2, st); ++n;
      }
      run_gemm(layer_plans_[li].fc, st); ++n;
      // ---- feature taps (models.py:568-582) ----
      while (tap_i < outs.taps.size() && outs.taps[tap_i].layer < li) ++tap_i;
      const bool last = li == cfg.layers - 1;
      // the bf16 token-major tap of a non-final layer is written by the c_proj epilogue itself (second output)
      __nv_bfloat16* fused_tap = nullptr;
      if (!last && !cfg.precise)
        for (size_t t = tap_i; t < outs.taps.size() && outs.taps[t].layer == li; ++t)
          if (outs.taps[t].tokens_bf16 && !fused_tap) fused_tap = outs.taps[t].tokens_bf16;
      if (fold()) {   // (the plan already writes bf16(x) to the tap, or to the shared copy, and the row statistics)
        run_gemm(layer_plans_[li].proj, st); ++n;
      } else if (fused_tap) {
        GemmPlan pl = layer_plans_[li].proj;
        pl.p.out_bf16 = fused_tap;
        pl.p.ldcb = D;
        run_gemm(pl, st); ++n;
      } else {
        run_gemm(layer_plans_[li].proj, st); ++n;
      }
      bool lnp_done = false;
      for (size_t t = tap_i; t < outs.taps.size() && outs.taps[t].layer == li; ++t) {
        const VitTap& tp = outs.taps[t];
        const float* src = x;
        if (last) {
          if (!lnp_done) {  // ln_post once; its bf16 copy (if wanted) comes out of the same kernel
            LayerNormParams lq{x, D, weights.ln_post_g, weights.ln_post_b, 1e-5f, L.M, D, lnp, D, tp.tokens_bf16, D, 0, 0};
            launch_layernorm(lq, st); ++n;
            lnp_done = true;
            src = lnp;
            if (tp.nchw) { launch_tap_nchw(src, tp.nchw, B, L.Ntok, D, st); ++n; }
            continue;
          }
          src = lnp;
        }
        if (tp.nchw) { launch_tap_nchw(src, tp.nchw, B, L.Ntok, D, st); ++n; }
        if (tp.tokens_bf16 && tp.tokens_bf16 != fused_tap) {
          CastParams cp{src, D, tp.tokens_bf16, D, L.M, D, 0, 0, 1.0f};
          cast_bf16_kernel<<<148 * 8, 256, 0, st>>>(cp); ++n;
        }
      }
      auto ensure_lnp = [&] {
        if (lnp_done) return;
        LayerNormParams lq{x, D, weights.ln_post_g, weights.ln_post_b, 1e-5f, L.M, D, lnp, D, nullptr, 0, 0, 0};
        launch_layernorm(lq, st); ++n;
        lnp_done = true;
      };
      if (last && outs.last_tokens_f32) {
        ensure_lnp();
        DCLIP_CHECK_CUDA(cudaMemcpyAsync(outs.last_tokens_f32, lnp, size_t(L.M) * D * 4, cudaMemcpyDeviceToDevice, st));
      }
    }
    DCLIP_CHECK_CUDA(cudaGetLastError());
    return n;
  }

 private:
  struct PlanKey {
    int B = 0, H = 0, W = 0;
    void* ws = nullptr;
    std::vector<void*> taps;   // ln_fold: the bf16 token taps are GEMM operands of the next layer (baked into its tensor map)
    bool operator==(const PlanKey& o) const { return B == o.B && H == o.H && W == o.W && ws == o.ws && taps == o.taps; }
  };
  struct LayerPlans {
    GemmPlan qkv, out_proj, fc, proj;
  };
  PlanKey plan_key_;
  GemmPlan patch_plan_;
  std::vector<LayerPlans> layer_plans_;
  AttnPlan attn_plan_;
  AttnSplitPlan attn_split_plan_;
  SmallAttnParams cls_attn_;

  // slots a residual GEMM publishes per row: one per (n-block, epilogue warp half)
  static int stat_slots(const GemmPlan& producer) {
    const int n = 2 * ((producer.p.N + producer.bn - 1) / producer.bn);
    DCLIP_REQUIRE(n <= kStatSlots, "ln_fold: %d statistics slots per row (max %d)", n, kStatSlots);
    return n;
  }

  void build_plans(int B, int H, int W, void* ws, const Layout& L, const VitOutputs& outs) {
    PlanKey key{B, H, W, ws, {}};
    // bf16 destination of layer li's updated residual stream: its token tap if one is requested, else the shared copy
    std::vector<__nv_bfloat16*> xb_dst(cfg.layers, nullptr);
    if (fold()) {
      for (int li = 0; li + 1 < cfg.layers; ++li) {
        xb_dst[li] = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(ws) + L.xb);
        for (const VitTap& t : outs.taps)
          if (t.layer == li && t.tokens_bf16) { xb_dst[li] = t.tokens_bf16; break; }
        key.taps.push_back(xb_dst[li]);
      }
    }
    if (key == plan_key_ && int(layer_plans_.size()) == cfg.layers) return;
    const int D = cfg.width, s = cfg.precise ? 2 : 1, M = L.M;
    uint8_t* w8 = static_cast<uint8_t*>(ws);
    float* x = reinterpret_cast<float*>(w8 + L.x);
    __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(w8 + L.h);
    __nv_bfloat16* gbuf = reinterpret_cast<__nv_bfloat16*>(w8 + L.g);
    float* pos = reinterpret_cast<float*>(w8 + L.pos);
    const int K0 = kp();
    {
      // patch-embed GEMM: rows m = b*P + p -> token row b*Ntok + 1 + p, + pos[1 + p]
      GemmOperands op{gbuf, K0 * s, weights.conv1_w, K0 * s};
      GemmParams p{};
      p.M = B * L.P; p.N = D; p.K = K0; p.split_in = cfg.precise; p.out_scale = 1.f;
      p.residual = pos; p.ldr = D; p.res_mod = 1; p.remap_P = L.P; p.remap_Nt = L.Ntok;
      p.out_f32 = x; p.ldc = D;
      patch_plan_ = make_gemm_plan(op, p);
    }
    layer_plans_.assign(cfg.layers, LayerPlans{});
    for (int li = 0; li < cfg.layers; ++li) {
      const VitLayerWeights& lw = weights.layers[li];
      LayerPlans& lp = layer_plans_[li];
      float2* st1 = fold() ? reinterpret_cast<float2*>(w8 + L.st1) : nullptr;
      float2* st2 = fold() ? reinterpret_cast<float2*>(w8 + L.st2) : nullptr;
      const __nv_bfloat16* a_qkv = hbuf;
      const __nv_bfloat16* a_fc = hbuf;
      if (fold()) {
        DCLIP_REQUIRE(lw.ln1_c && lw.ln2_c, "ln_fold: folded-weight row sums (ln1_c / ln2_c) not set");
        a_qkv = li == 0 ? reinterpret_cast<const __nv_bfloat16*>(w8 + L.xb) : xb_dst[li - 1];
        a_fc = reinterpret_cast<const __nv_bfloat16*>(w8 + L.xb2);
      }
      {
        GemmOperands op{a_qkv, D * s, lw.in_proj_w, D * s};
        GemmParams p{};
        p.M = M; p.N = 3 * D; p.K = D; p.split_in = cfg.precise; p.bias = lw.in_proj_b; p.out_scale = 1.f;
        if (fold()) { p.row_stats_in = st1; p.stats_ld = M; p.stats_n = li == 0 ? 1 : stat_slots(layer_plans_[li - 1].proj); p.ln_c = lw.ln1_c; p.ln_inv_d = 1.f / D; p.ln_eps = 1e-5f; }
        p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(w8 + L.qkv);
        if (cfg.precise) { p.ldcb = 6 * D; p.split_out = 1; p.split_out_off = 3 * D; }   // [q k v]_hi | [q k v]_lo
        else p.ldcb = 3 * D;
        lp.qkv = make_gemm_plan(op, p);
      }
      {
        GemmOperands op{hbuf, D * s, lw.out_proj_w, D * s};
        GemmParams p{};
        p.M = M; p.N = D; p.K = D; p.split_in = cfg.precise; p.bias = lw.out_proj_b; p.out_scale = 1.f;
        p.residual = x; p.ldr = D; p.out_f32 = x; p.ldc = D;
        if (fold()) {
          p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(w8 + L.xb2); p.ldcb = D;
          p.row_stats_out = st2; p.stats_ld = M;
        }
        lp.out_proj = make_gemm_plan(op, p);
      }
      {
        GemmOperands op{a_fc, D * s, lw.fc_w, D * s};
        GemmParams p{};
        p.M = M; p.N = 4 * D; p.K = D; p.split_in = cfg.precise; p.bias = lw.fc_b; p.out_scale = 1.f;
        if (fold()) { p.row_stats_in = st2; p.stats_ld = M; p.stats_n = stat_slots(lp.out_proj); p.ln_c = lw.ln2_c; p.ln_inv_d = 1.f / D; p.ln_eps = 1e-5f; }
        p.act = cfg.precise ? ACT_QUICKGELU_PRECISE : ACT_QUICKGELU;
        p.out_bf16 = gbuf; p.ldcb = 4 * D * s; p.split_out = cfg.precise; p.split_out_off = 4 * D;
        lp.fc = make_gemm_plan(op, p);
      }
      {
        GemmOperands op{gbuf, 4 * D * s, lw.proj_w, 4 * D * s};
        GemmParams p{};
        p.M = M; p.N = D; p.K = 4 * D; p.split_in = cfg.precise; p.bias = lw.proj_b; p.out_scale = 1.f;
        p.residual = x; p.ldr = D; p.out_f32 = x; p.ldc = D;
        if (fold() && li + 1 < cfg.layers) {
          p.out_bf16 = xb_dst[li]; p.ldcb = D;
          p.row_stats_out = st1; p.stats_ld = M;
        }
        lp.proj = make_gemm_plan(op, p);
      }
    }
    if (cfg.precise) {
      const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(w8 + L.qkv);
      const long long bs = (long long)L.Ntok * 6 * D;
      AttnSplitOperands op{qkv, qkv, qkv, 6 * D, 6 * D, 6 * D, bs, bs, bs};
      AttnSplitParams p{};
      p.B = B; p.H = cfg.heads; p.Nq = L.Ntok; p.Nk = L.Ntok;
      p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D; p.lo_off = 3 * D;
      p.scale_log2 = 0.125f * 1.4426950408889634f;
      p.out = hbuf; p.out_bs = (long long)L.Ntok * 2 * D; p.ldo = 2 * D; p.out_lo_off = D;
      attn_split_plan_ = make_attn_split_plan(op, p);
    } else {
      const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(w8 + L.qkv);
      const long long bs = (long long)L.Ntok * 3 * D;
      AttnOperands op{qkv, qkv, qkv, 3 * D, 3 * D, 3 * D, bs, bs, bs, L.Ntok};
      AttnParams p{};
      p.B = B; p.H = cfg.heads; p.Nq_total = L.Ntok; p.Nk = L.Ntok;
      // All query rows (CLS included) go through the tensor-core kernel.  (Measured on B200: peeling the CLS query
      // off into the few-query kernel so that N - 1 = 8 x 256 rows tile exactly was slower, 0.370 vs 0.343 ms/layer at
      // B = 16, because the side kernel is latency-bound; see profiles/r01_attention_notes.md.)
      p.q_start = 0;
      p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D;
      p.scale_log2 = 0.125f * 1.4426950408889634f;
      p.out = hbuf; p.out_batch_stride = (long long)L.Ntok * D; p.ldo = D;
      attn_plan_ = make_attn_plan(op, p);
      SmallAttnParams sp{};
      sp.q = sp.k = sp.v = qkv; sp.is_f32 = 0; sp.B = B; sp.H = cfg.heads; sp.Nk = L.Ntok; sp.q_first = 0; sp.q_count = 1;
      sp.ldq = sp.ldk = sp.ldv = 3 * D; sp.q_bs = sp.k_bs = sp.v_bs = bs;
      sp.q_col0 = 0; sp.k_col0 = D; sp.v_col0 = 2 * D; sp.scale = 0.125f; sp.causal = 0;
      sp.out = hbuf; sp.out_f32 = 0; sp.ldo = D; sp.out_bs = (long long)L.Ntok * D; sp.out_split_off = 0;
      cls_attn_ = sp;
    }
    plan_key_ = key;
  }
};

}  // namespace dclip
