// Column-split persistent flash attention (round 2): the same pipeline as attn_fwd_persistent_kernel (attn_tcgen05.cuh) --
// one CTA per SM, static (image, head, 256-query block) item list, TMEM / barriers / K/V ring alive across items, S = Q K^T
// and O += P V on tcgen05.mma with P kept in TMEM, key 0 peeled, tail rows in the background, TMA-store output pass -- but
// with FOUR softmax warpgroups instead of two: warpgroup (i, hf) owns query tile i and the key half hf of every 128-key tile,
// i.e. a thread handles one query row x 64 keys.
//
// Why (profiles/r01_attention_notes.md, r02_attention_notes.md): with one warp per scheduler and query tile, a KV step took
// ~2850 cycles for 2 x 1024 SFU cycles because the load -> max -> exp -> publish chain of a lone in-order warp is as long as
// its SFU work and the MUFU stream has no slack for the sibling warp's instructions.  Splitting the COLUMNS (not the keys of
// the MMA: S and P tiles, and therefore every tcgen05.mma, keep their round-1 shapes -- the 4-warpgroup split-K experiment
// lost to the fixed latency of short MMA groups) halves every warp's chain and puts four warps on each scheduler, so the
// SFU pipe always finds an exponential to issue; no MUFU token is needed.
//
// The two threads that share a query row exchange their 64-key maxima through shared memory and a 256-thread named barrier
// per tile, so the reference max, the lazy-rescale decisions and every P value are exactly those of the two-warpgroup
// kernel (row sums differ only in fp32 summation order).
#pragma once
#include "../../denseclip_vit_multimodal_b200/csrc/attn_tcgen05.cuh"

namespace dclip {

struct AttnCsCfg {
  static constexpr int TQ = 128, TKV = 128, HD = 64, KV_STAGES = 4;
  static constexpr int Q_OFF = 0;                               // 2 buffers x 2 tiles x 16 KB
  static constexpr int K_OFF = 4 * 16384;
  static constexpr int V_OFF = K_OFF + KV_STAGES * 16384;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * 16384;
  static constexpr int NUM_BARS = 4 + 3 * KV_STAGES + 8;
  static constexpr int MERGE_OFF = BAR_OFF + 256;               // background tail-row merge buffer (8 x 66 floats)
  static constexpr int XMAX_OFF = MERGE_OFF + 8 * 66 * 4;       // [parity 2][tile 2][half 2][128] row maxima
  static constexpr int XSUM_OFF = XMAX_OFF + 2 * 2 * 2 * 128 * 4;   // [tile 2][half 2][128] row sums (output pass)
  static constexpr int SMEM_BYTES = XSUM_OFF + 2 * 2 * 128 * 4;
  static constexpr int SOFTMAX_WARPS = 16;
  static constexpr int THREADS = 640;                           // 16 softmax warps + MMA + TMA + 2 tail / allocator warps
  static constexpr int TMEM_COLS = 512;  // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
};
static_assert(AttnCsCfg::NUM_BARS * 8 + 16 <= 256 && AttnCsCfg::SMEM_BYTES <= 232448, "column-split attention smem budget");

// One KV tile for one (query row, key half): 64 S columns (fp32, TMEM) -> 64 P values (bf16 pairs, 32 TMEM columns).
// `valid` (0..64, warp-uniform) masks the trailing columns of a ragged last tile.
template <int POLY>
__device__ __forceinline__ void attn_cs_softmax_tile(uint32_t tS, uint32_t tO, uint32_t tP, int lane, int valid, bool first, float sc,
                                                     float& m_used, float& l, uint64_t* s_free_bar, uint64_t* o_done_bar,
                                                     uint32_t o_done_parity, float* x_mine, const float* x_partner, int pair_bar) {
  constexpr int NC = 64;
  uint32_t su[NC];
  if (valid > 0) {
    tmem_ld_32x32b_x32(tS, reinterpret_cast<uint32_t(&)[32]>(su[0]));
    tmem_ld_32x32b_x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(su[32]));
    tmem_wait_ld();
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_free_bar);
  float mx = -INFINITY;
  if (valid > 0) {
    if (valid < NC) {
#pragma unroll
      for (int e = 0; e < NC; ++e)
        if (e >= valid) su[e] = 0xff800000u;  // -inf
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int e = 0; e < NC; e += 8) {
      mx0 = fmaxf(mx0, fmaxf(__uint_as_float(su[e]), __uint_as_float(su[e + 1])));
      mx1 = fmaxf(mx1, fmaxf(__uint_as_float(su[e + 2]), __uint_as_float(su[e + 3])));
      mx2 = fmaxf(mx2, fmaxf(__uint_as_float(su[e + 4]), __uint_as_float(su[e + 5])));
      mx3 = fmaxf(mx3, fmaxf(__uint_as_float(su[e + 6]), __uint_as_float(su[e + 7])));
    }
    mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  }
  // the row's other half: exchange the partial maxima (double-buffered by tile parity: the partner reads this slot before
  // it arrives on the NEXT tile's barrier, and this thread rewrites it only after that barrier)
  *x_mine = mx;
  named_bar_sync(pair_bar, 256);
  const float m_new = fmaxf(m_used, fmaxf(mx, *x_partner));
  // lazy rescale: keep the old reference max unless it grew by more than 2^8 (first tile: m_used = -inf -> always).
  // Both halves see the same m_new and the same 32 rows per warp, so they take the same decision.
  const bool need = (m_new - m_used) * sc > 8.0f;
  const bool rescale = __any_sync(0xffffffffu, need);
  float alpha = 1.0f;
  if (rescale) {
    alpha = ex2_approx((m_used - m_new) * sc);
    m_used = m_new;
    l *= alpha;
  }
  if (!first) {   // PV of the previous tile must be done before P is overwritten or O is touched
    mbar_wait(o_done_bar, o_done_parity);
    tc_fence_after();
    if (rescale) {   // this half's 32 of the 64 output columns
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO, o);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
      tmem_st_32x32b_x32(tO, o);
      tmem_wait_st();
    }
  }
  if (valid > 0) {
    const uint64_t sc2 = pack_f32x2(sc, sc);
    const float nmc = -m_used * sc;
    const uint64_t nmc2 = pack_f32x2(nmc, nmc);
    uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
    uint32_t pk[16];
#pragma unroll
    for (int c16 = 0; c16 < NC / 8; ++c16) {
      float pv[8];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const uint64_t t = fma_f32x2(pack_f32x2(__uint_as_float(su[c16 * 8 + e]), __uint_as_float(su[c16 * 8 + e + 1])), sc2, nmc2);
        if (e >= 8 - 2 * POLY) {
          exp2_poly_x2(t, pv[e], pv[e + 1]);
        } else {
          float t0, t1;
          unpack_f32x2(t, t0, t1);
          pv[e] = ex2_approx(t0);
          pv[e + 1] = ex2_approx(t1);
        }
      }
      acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(pv[0], pv[1]), pack_f32x2(pv[2], pv[3])));
      acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(pv[4], pv[5]), pack_f32x2(pv[6], pv[7])));
      // P stays on-chip in TMEM: lane = query row, 32-bit column c holds (P[2c], P[2c+1]) -- the A operand of the TS MMA
      pk[(c16 & 3) * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
      pk[(c16 & 3) * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
      pk[(c16 & 3) * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
      pk[(c16 & 3) * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
      if ((c16 & 3) == 3) tmem_st_32x32b_x16(tP + (c16 >> 2) * 16, pk);
    }
    float a0, a1, a2, a3;
    unpack_f32x2(acc0, a0, a1);
    unpack_f32x2(acc1, a2, a3);
    l += (a0 + a1) + (a2 + a3);
  }
}

// Output pass of one (query row, column half): 32 columns of O (TMEM) * alpha + p0 * v_0 -> bf16 -> this thread's half of its row
// of the finished Q tile in shared memory (SWIZZLE_128B layout); one TMA store per query tile follows.
__device__ __forceinline__ void attn_cs_store_half(uint32_t tO, uint8_t* q_tile, int r, int hf, float alpha, float p0, bool fold,
                                                   const uint4 (&v0r)[4]) {
  uint8_t* srow = q_tile + (r >> 3) * 1024 + (r & 7) * 128;
  uint32_t o[32];
  tmem_ld_32x32b_x32(tO, o);
  tmem_wait_ld();
#pragma unroll
  for (int e = 0; e < 32; e += 8) {
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(o[e + u]) * alpha;
    if (fold) {
      const uint4 vv = v0r[e >> 3];
      const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        f[2 * u] = fmaf(p0, __uint_as_float(vw[u] << 16), f[2 * u]);
        f[2 * u + 1] = fmaf(p0, __uint_as_float(vw[u] & 0xffff0000u), f[2 * u + 1]);
      }
    }
    *reinterpret_cast<uint4*>(srow + (((hf * 4 + (e >> 3)) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
  fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
}

template <int POLY = 0>
__global__ void __launch_bounds__(AttnCsCfg::THREADS, 1)
attn_fwd_cs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  using Cfg = AttnCsCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;                       // [2]  producer -> MMA / softmax : Q buffer landed
  uint64_t* q_empty = bars + 2;                  // [2]  MMA commit + the 2 output-store threads -> producer : Q buffer reusable
  uint64_t* k_full = bars + 4;
  uint64_t* v_full = k_full + Cfg::KV_STAGES;
  uint64_t* kv_empty = v_full + Cfg::KV_STAGES;
  uint64_t* s_full = kv_empty + Cfg::KV_STAGES;
  uint64_t* s_free = s_full + 2;
  uint64_t* p_ready = s_free + 2;
  uint64_t* o_done = p_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);
  float* x_max = reinterpret_cast<float*>(smem + Cfg::XMAX_OFF);
  float* x_sum = reinterpret_cast<float*>(smem + Cfg::XSUM_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G_CTAS = gridDim.x, cta = blockIdx.x;
  const int nqb = (p.Nq_total - p.q_start + 2 * Cfg::TQ - 1) / (2 * Cfg::TQ);
  const int rows_last = p.Nq_total - p.q_start - (nqb - 1) * 2 * Cfg::TQ;
  const int has_tail = rows_last <= p.tail_rows_max ? 1 : 0;
  const int nqb_reg = nqb - has_tail;
  const int n_reg = p.B * p.H * nqb_reg;          // host guarantees n_reg >= gridDim.x
  const int n_tail = p.B * p.H * has_tail;        // host guarantees tail_overlap whenever n_tail > 0
  const int koff = p.peel_key0 ? 1 : 0;
  const int nk_eff = p.Nk - koff;
  const int T = (nk_eff + Cfg::TKV - 1) / Cfg::TKV;
  const int last_valid = nk_eff - (T - 1) * Cfg::TKV;
  const int last_cols16 = (last_valid + 15) & ~15;
  const int first_r = G_CTAS - 1 - cta;           // (see attn_fwd_persistent_kernel: CTAs with a tail row get fewer regular items)
  const int n_items = (n_reg - first_r + G_CTAS - 1) / G_CTAS;
  const int G = n_items * T;                      // KV tiles this CTA processes per query tile

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("dclip attn: dynamic smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 3);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_ready[i], 8);
      mbar_init(&o_done[i], 1);
    }
    for (int s = 0; s < Cfg::KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 18) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= Cfg::SOFTMAX_WARPS) {
    // register pool of the CTA = 640 x 96 (launch bounds): 128 x 64 + 512 x 104 = 61440 uses it exactly (setmaxnreg.inc blocks
    // forever if the CTA's pool cannot supply the increase)
    setmaxnreg_dec<64>();
    if (warp == 17) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        int g = 0;
        for (int n = 0; n < n_items; ++n) {
          const int r = first_r + n * G_CTAS;
          const int qb = r % nqb_reg, h = (r / nqb_reg) % p.H, b = r / (nqb_reg * p.H);
          const int q_row0 = p.q_start + qb * 2 * Cfg::TQ;
          const int qbuf = n & 1;
          if (n >= 2) mbar_wait_relaxed(&q_empty[qbuf], ((n >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&q_full[qbuf], 2 * 16384);
          tma_load_3d(smem + Cfg::Q_OFF + qbuf * 32768, &tmQ, &q_full[qbuf], p.q_col0 + h * Cfg::HD, q_row0, b);
          tma_load_3d(smem + Cfg::Q_OFF + qbuf * 32768 + 16384, &tmQ, &q_full[qbuf], p.q_col0 + h * Cfg::HD, q_row0 + Cfg::TQ, b);
          for (int j = 0; j < T; ++j, ++g) {
            const int s = g % Cfg::KV_STAGES;
            const uint32_t ph = (g / Cfg::KV_STAGES) & 1;
            mbar_wait_relaxed(&kv_empty[s], ph ^ 1);
            mbar_arrive_expect_tx(&k_full[s], 16384);
            tma_load_3d(smem + Cfg::K_OFF + s * 16384, &tmK, &k_full[s], p.k_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
            mbar_arrive_expect_tx(&v_full[s], 16384);
            tma_load_3d(smem + Cfg::V_OFF + s * 16384, &tmV, &v_full[s], p.v_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
          }
        }
      }
    } else if (warp >= 18) {
      // ------------------------------- tail rows in the background ----------------
      if (p.tail_overlap) {
        float* s_merge = reinterpret_cast<float*>(smem + Cfg::MERGE_OFF);
        for (int t = cta; t < n_tail; t += G_CTAS) {
          const int h = t % p.H, b = t / p.H;
          attn_tail_rows_bg(p.q + (long long)b * p.q_bs + p.q_col0 + h * 64, p.k + (long long)b * p.k_bs + p.k_col0 + h * 64,
                            p.v + (long long)b * p.v_bs + p.v_col0 + h * 64, p.out + (long long)b * p.out_batch_stride + h * 64,
                            p.ldq, p.ldk, p.ldv, p.ldo, p.Nk, p.scale_log2, p.q_start + nqb_reg * 2 * Cfg::TQ, rows_last,
                            s_merge, threadIdx.x - 18 * 32);
        }
      }
    } else {
      // ------------------------------- MMA issuer (warp 16) ------------------------
      const uint64_t dQ = make_smem_desc_sw128(smem_u32(smem + Cfg::Q_OFF), 16, 1024);
      const uint64_t dK = make_smem_desc_sw128(smem_u32(smem + Cfg::K_OFF), 16, 1024);
      const uint64_t dV = make_smem_desc_sw128(smem_u32(smem + Cfg::V_OFF), 16, 1024);
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t idesc_qk_last = make_idesc_bf16(128, last_cols16);
      auto issue_qk = [&](int i, int g, int n, int j) {
        const int stage = g % Cfg::KV_STAGES;
        const uint64_t a = dQ + uint64_t(n & 1) * 2048 + uint64_t(i) * 1024, bb = dK + uint64_t(stage) * 1024;
        const uint32_t idesc = (j == T - 1) ? idesc_qk_last : idesc_qk;
        const uint32_t d = tmem_base + i * 128;
        const bool release_q = (i == 1 && j == T - 1);  // last QK^T that reads this item's Q buffer
        if (elect_one_sync()) {
          umma_ss_f16(d, a, bb, idesc, 0u);
          umma_ss_f16(d, a + 2, bb + 2, idesc, 1u);
          umma_ss_f16(d, a + 4, bb + 4, idesc, 1u);
          umma_ss_f16(d, a + 6, bb + 6, idesc, 1u);
          umma_commit(&s_full[i]);
          if (release_q) umma_commit(&q_empty[n & 1]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int i, int stage, bool is_last, uint32_t acc, bool release_kv) {
        const uint64_t bb = dV + uint64_t(stage) * 1024;
        const uint32_t d = tmem_base + 256 + i * 64;
        const uint32_t ta = tmem_base + 384 + i * 64;
        if (elect_one_sync()) {
          if (!is_last || last_cols16 == 128) {
            umma_ts_f16(d, ta, bb, idesc_pv, acc);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) umma_ts_f16(d, ta + ks * 8, bb + ks * 128, idesc_pv, 1u);
          } else {
            for (int ks = 0; ks < last_cols16 / 16; ++ks)
              umma_ts_f16(d, ta + ks * 8, bb + ks * 128, idesc_pv, (acc | ks) ? 1u : 0u);
          }
          umma_commit(&o_done[i]);
          if (release_kv) umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      };
      // prologue: tiles g = 0 and g = 1
      mbar_wait(&q_full[0], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, 0, 0);
      issue_qk(1, 0, 0, 0);
      if (G > 1) {
        const int n1 = (T == 1) ? 1 : 0, j1 = (T == 1) ? 0 : 1;
        if (n1) mbar_wait(&q_full[1], 0);
        mbar_wait(&k_full[1 % Cfg::KV_STAGES], (1 / Cfg::KV_STAGES) & 1);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&s_free[i], 0);
          tc_fence_after();
          issue_qk(i, 1, n1, j1);
        }
      }
      int j = 0;                                  // tile index of g inside its item
      int n2 = (T <= 2) ? 2 / T : 0, j2 = 2 % T;  // item / tile index of g + 2
      for (int g = 0; g < G; ++g) {
        const int s = g % Cfg::KV_STAGES;
        mbar_wait(&v_full[s], (g / Cfg::KV_STAGES) & 1);
        const int g2 = g + 2;
        const bool more = g2 < G;
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_ready[i], g & 1);
          tc_fence_after();
          issue_pv(i, s, j == T - 1, j > 0 ? 1u : 0u, i == 1);
        }
        if (more) {
          mbar_wait(&k_full[g2 % Cfg::KV_STAGES], (g2 / Cfg::KV_STAGES) & 1);
          if (j2 == 0) mbar_wait(&q_full[n2 & 1], (n2 >> 1) & 1);
          for (int i = 0; i < 2; ++i) {
            mbar_wait(&s_free[i], (g + 1) & 1);
            tc_fence_after();
            issue_qk(i, g2, n2, j2);
          }
        }
        if (++j == T) j = 0;
        if (++j2 == T) { j2 = 0; ++n2; }
      }
    }
  } else {
    // ------------------------------- softmax warpgroups --------------------------
    setmaxnreg_inc<104>();
    const int i = warp >> 3;          // query tile
    const int hf = (warp >> 2) & 1;   // key half of every KV tile (and output-column half in the output pass)
    const int q = warp & 3;           // TMEM lane quarter (hardware: warp w reaches lanes 32 (w % 4) ..)
    const int r = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + i * 128 + hf * 64 + lane_off;
    const uint32_t tO = tmem_base + 256 + i * 64 + hf * 32 + lane_off;
    const uint32_t tP = tmem_base + 384 + i * 64 + hf * 32 + lane_off;
    const float sc = p.scale_log2;
    const int pair_bar = 1 + i, store_bar = 3 + i;
    int g = 0;
    for (int n = 0; n < n_items; ++n) {
      const int ritem = first_r + n * G_CTAS;
      const int qb = ritem % nqb_reg, h = (ritem / nqb_reg) % p.H, b = ritem / (nqb_reg * p.H);
      const int q_row0 = p.q_start + qb * 2 * Cfg::TQ;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < T; ++j, ++g) {
        mbar_wait(&s_full[i], g & 1);
        tc_fence_after();
        const int valid_tile = (j + 1 == T) ? last_valid : 128;
        const int valid = min(64, max(0, valid_tile - hf * 64));
        if (koff && lane == 0 && hf == 0 && j + 1 == T) {   // k_0 / v_0 into L1 for the output pass
          asm volatile("prefetch.global.L1 [%0];" ::"l"(p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD));
        }
        float* xm = x_max + (((g & 1) * 2 + i) * 2) * 128;
        attn_cs_softmax_tile<POLY>(tS, tO, tP, lane, valid, j == 0, sc, m_used, l, &s_free[i], &o_done[i], (g - 1) & 1,
                                   xm + hf * 128 + r, xm + (hf ^ 1) * 128 + r, pair_bar);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[i]);
        if (j == 0 && n > 0 && hf == 0 && q == 0 && lane == 0) {
          // the previous item's output tile sits in its Q buffer until the TMA store has read it (long done by now)
          tma_store_wait_read();
          mbar_arrive(&q_empty[(n - 1) & 1]);
        }
      }
      // ---- output pass of this item (the MMA warp is already on the next item's first tiles) ----
      float z0 = 0.f;
      uint4 v0r[4];
      if (koff) {
        const uint4* k0 = reinterpret_cast<const uint4*>(p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD);
        const uint4* v0 = reinterpret_cast<const uint4*>(p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD);
        const uint8_t* qrow = smem + Cfg::Q_OFF + (n & 1) * 32768 + i * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {   // (both halves compute the full 64-dim score of key 0: same operands, same order)
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
          const uint4 kv = __ldg(k0 + c);
          const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            s0 = fmaf(__uint_as_float(qq[e] << 16), __uint_as_float(kk[e] << 16), s0);
            s1 = fmaf(__uint_as_float(qq[e] & 0xffff0000u), __uint_as_float(kk[e] & 0xffff0000u), s1);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) v0r[c] = __ldg(v0 + hf * 4 + c);
        z0 = (s0 + s1) * sc;
      }
      // total row sum = half 0 + half 1 (fixed order: both threads of a row get the same bits)
      x_sum[(i * 2 + hf) * 128 + r] = l;
      named_bar_sync(pair_bar, 256);
      l = x_sum[(i * 2) * 128 + r] + x_sum[(i * 2 + 1) * 128 + r];
      mbar_wait(&o_done[i], (g - 1) & 1);
      tc_fence_after();
      float alpha = 1.0f, p0 = 0.f;
      if (koff) {
        const float e0 = z0 - m_used * sc;
        alpha = e0 > 0.f ? ex2_approx(-e0) : 1.0f;
        p0 = e0 > 0.f ? 1.0f : ex2_approx(e0);
        l = l * alpha + p0;
      }
      const float inv = 1.0f / l;
      alpha *= inv;
      p0 *= inv;
      uint8_t* q_tile = smem + Cfg::Q_OFF + (n & 1) * 32768 + i * 16384;
      attn_cs_store_half(tO, q_tile, r, hf, alpha, p0, koff != 0, v0r);
      named_bar_sync(store_bar, 256);
      if (hf == 0 && q == 0 && lane == 0) {
        tma_store_3d(&tmO, q_tile, h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
        tma_store_commit();
      }
      tc_fence_before();  // the O reads are ordered before this warp's next p_ready arrive (next item's first PV overwrites O)
    }
    if (hf == 0 && q == 0 && lane == 0) tma_store_wait_read();  // last item's tile: shared memory must outlive the store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 18) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dclip
