// fp32-class flash attention on the tensor cores (head_dim 64, non-causal) for precision="fp32": the same
// softmax(QK^T)V of nn.MultiheadAttention (segmentation/denseclip/models.py:287-289), with every operand carried as a
// bf16 hi|lo pair (x = hi + lo, 16 mantissa bits) and every product evaluated as three tcgen05 passes
//     S = Qh Kh^T + Ql Kh^T + Qh Kl^T          O += Ph Vh + Pl Vh + Ph Vl
// accumulated in fp32 in TMEM (the dropped lo*lo term is 2^-18 relative).  Softmax statistics, the exponentials and the
// row sums are fp32.  This replaces the CUDA-core fp32 kernel of round 1 (8 ms per layer at B = 16) in the one mode that
// meets the north star's >= 99.9% argmax gate (no bf16 block may remain on the path).
//
//   * one CTA = one (image, head, 256-query block); warps 0-3 / 4-7 = softmax warpgroups of query tile 0 / 1 (one thread
//     per query row), warp 8 = MMA issuer, warp 9 = TMA producer, warp 10 = TMEM allocator
//   * TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384).  P ALIASES S: once a thread has pulled its S row into
//     registers it writes P_hi into columns [0,64) and P_lo into [64,128) of the same buffer (thread = TMEM lane, so no
//     cross-thread hazard); the MMA warp issues PV(j) and then QK^T(j+1) into that buffer -- tcgen05.mma ops of one thread
//     execute in order, so "S(j+1) ready" implies "PV(j) done" and no separate PV-done barrier is needed inside the loop
//   * the kernel is tensor-bound (3 x the MMA work of the bf16 kernel against the same exponentials), so there is no MUFU
//     token, no key peeling and no CUDA-core tail path: ragged edges are a narrow last MMA and masked columns / rows
#pragma once
#include "attn_tcgen05.cuh"

namespace dclip {

struct AttnSplitParams {
  int B, H;
  int Nq, Nk;
  int q_col0, k_col0, v_col0;  // column of head 0's hi half inside a token row of the Q / K / V tensor
  int lo_off;                  // column distance from a hi half to its lo half
  float scale_log2;            // head_dim^-0.5 * log2(e)
  __nv_bfloat16* out;          // out[b][row][h*64 + d] = hi, [.. + out_lo_off] = lo
  long long out_bs;            // elements
  int ldo;
  int out_lo_off;
};

struct AttnSplitCfg {
  static constexpr int TQ = 128, TKV = 128, HD = 64, KV_STAGES = 2;
  static constexpr int Q_OFF = 0;                              // [tile 0 hi | tile 0 lo | tile 1 hi | tile 1 lo] x 16 KB
  static constexpr int K_OFF = 4 * 16384;                      // KV_STAGES x [hi | lo] x 16 KB
  static constexpr int V_OFF = K_OFF + KV_STAGES * 32768;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * 32768;
  static constexpr int NUM_BARS = 1 + 4 * KV_STAGES + 6;
  static constexpr int SMEM_BYTES = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr int THREADS = 384;
  static constexpr int TMEM_COLS = 512;
};
static_assert(AttnSplitCfg::SMEM_BYTES <= 232448, "split attention smem budget");

// One KV tile of the online softmax for one query row: S (NC fp32 columns, TMEM) -> P_hi | P_lo (bf16, same TMEM buffer).
template <int NC>
__device__ __forceinline__ void attn_split_softmax_tile(uint32_t tS, uint32_t tO, int valid, bool first, float sc, float& m_used,
                                                        float& l) {
  uint32_t su[NC];
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, reinterpret_cast<uint32_t(&)[32]>(su[c * 32]));
  tmem_wait_ld();
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
  const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
  // lazy rescale (same rule as the bf16 kernel): keep the reference max unless it grew by more than 2^8.  PV(j-1) is
  // complete here (the S tile this thread just read was produced by an MMA issued after it), so O may be touched.
  const bool need = (m_new - m_used) * sc > 8.0f;
  if (__any_sync(0xffffffffu, need)) {
    const float alpha = ex2_approx((m_used - m_new) * sc);
    m_used = m_new;
    l *= alpha;
    if (!first) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st_32x32b_x32(tO + c * 32, o);
      }
    }
  }
  const uint64_t sc2 = pack_f32x2(sc, sc);
  const float nmc = -m_used * sc;
  const uint64_t nmc2 = pack_f32x2(nmc, nmc);
  uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
  uint32_t ph[16], pl[16];
#pragma unroll
  for (int c8 = 0; c8 < NC / 8; ++c8) {
    float pv[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const uint64_t t = fma_f32x2(pack_f32x2(__uint_as_float(su[c8 * 8 + e]), __uint_as_float(su[c8 * 8 + e + 1])), sc2, nmc2);
      float t0, t1;
      unpack_f32x2(t, t0, t1);
      pv[e] = ex2_approx(t0);
      pv[e + 1] = ex2_approx(t1);
    }
    acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(pv[0], pv[1]), pack_f32x2(pv[2], pv[3])));
    acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(pv[4], pv[5]), pack_f32x2(pv[6], pv[7])));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t h = pack_bf16x2(pv[2 * e], pv[2 * e + 1]);
      ph[(c8 & 3) * 4 + e] = h;
      pl[(c8 & 3) * 4 + e] = pack_bf16x2(pv[2 * e] - __uint_as_float(h << 16), pv[2 * e + 1] - __uint_as_float(h & 0xffff0000u));
    }
    if ((c8 & 3) == 3) {  // 32 keys = 16 packed columns of each half
      tmem_st_32x32b_x16(tS + (c8 >> 2) * 16, ph);
      tmem_st_32x32b_x16(tS + 64 + (c8 >> 2) * 16, pl);
    }
  }
  float a0, a1, a2, a3;
  unpack_f32x2(acc0, a0, a1);
  unpack_f32x2(acc1, a2, a3);
  l += (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(AttnSplitCfg::THREADS, 1)
attn_fwd_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnSplitParams p) {
  using Cfg = AttnSplitCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + Cfg::KV_STAGES;
  uint64_t* k_empty = v_full + Cfg::KV_STAGES;
  uint64_t* v_empty = k_empty + Cfg::KV_STAGES;
  uint64_t* s_full = v_empty + Cfg::KV_STAGES;   // [2]  MMA -> softmax : S tile ready (and every earlier MMA complete)
  uint64_t* p_ready = s_full + 2;                // [2]  softmax -> MMA : P_hi | P_lo written over the S tile (and O rescaled)
  uint64_t* o_done = p_ready + 2;                // [2]  MMA -> softmax : the last PV of the item finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = (p.Nq + 2 * Cfg::TQ - 1) / (2 * Cfg::TQ);
  const int qb = blockIdx.x % nqb;
  const int h = (blockIdx.x / nqb) % p.H;
  const int b = blockIdx.x / (nqb * p.H);
  const int q_row0 = qb * 2 * Cfg::TQ;
  const int T = (p.Nk + Cfg::TKV - 1) / Cfg::TKV;
  const int last_valid = p.Nk - (T - 1) * Cfg::TKV;
  const int last_cols16 = (last_valid + 15) & ~15;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("dclip attn split: dynamic smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<80>();
    if (warp == 9) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 4 * 16384);
        for (int i = 0; i < 2; ++i) {
          tma_load_3d(smem + Cfg::Q_OFF + i * 32768, &tmQ, q_full, p.q_col0 + h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
          tma_load_3d(smem + Cfg::Q_OFF + i * 32768 + 16384, &tmQ, q_full, p.q_col0 + p.lo_off + h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
        }
        for (int j = 0; j < T; ++j) {
          const int s = j % Cfg::KV_STAGES;
          const uint32_t ph = (j / Cfg::KV_STAGES) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], 2 * 16384);
          tma_load_3d(smem + Cfg::K_OFF + s * 32768, &tmK, &k_full[s], p.k_col0 + h * Cfg::HD, j * Cfg::TKV, b);
          tma_load_3d(smem + Cfg::K_OFF + s * 32768 + 16384, &tmK, &k_full[s], p.k_col0 + p.lo_off + h * Cfg::HD, j * Cfg::TKV, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], 2 * 16384);
          tma_load_3d(smem + Cfg::V_OFF + s * 32768, &tmV, &v_full[s], p.v_col0 + h * Cfg::HD, j * Cfg::TKV, b);
          tma_load_3d(smem + Cfg::V_OFF + s * 32768 + 16384, &tmV, &v_full[s], p.v_col0 + p.lo_off + h * Cfg::HD, j * Cfg::TKV, b);
        }
      }
    } else if (warp == 8) {
      // ------------------------------- MMA issuer ----------------------------------
      const uint64_t dQ = make_smem_desc_sw128(smem_u32(smem + Cfg::Q_OFF), 16, 1024);
      const uint64_t dK = make_smem_desc_sw128(smem_u32(smem + Cfg::K_OFF), 16, 1024);
      const uint64_t dV = make_smem_desc_sw128(smem_u32(smem + Cfg::V_OFF), 16, 1024);
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major: head_dim contiguous
      const uint32_t idesc_qk_last = make_idesc_bf16(128, last_cols16);
      // (descriptor address fields are in units of 16 B: a 16 KB tile = 1024, a hi|lo pair = 2048)
      auto issue_qk = [&](int i, int stage, bool is_last, bool release_k) {
        const uint64_t ah = dQ + uint64_t(i) * 2048, al = ah + 1024;
        const uint64_t bh = dK + uint64_t(stage) * 2048, bl = bh + 1024;
        const uint32_t idesc = is_last ? idesc_qk_last : idesc_qk;
        const uint32_t d = tmem_base + i * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(d, ah + 2 * k, bh + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(d, al + 2 * k, bh + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(d, ah + 2 * k, bl + 2 * k, idesc, 1u);
          umma_commit(&s_full[i]);
          if (release_k) umma_commit(&k_empty[stage]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int i, int stage, bool is_last, uint32_t acc, bool release_v) {
        const uint64_t bh = dV + uint64_t(stage) * 2048, bl = bh + 1024;
        const uint32_t d = tmem_base + 256 + i * 64;
        const uint32_t th = tmem_base + i * 128, tl = th + 64;   // 16 bf16 of K per MMA = 8 TMEM columns
        const int nks = is_last ? last_cols16 / 16 : 8;
        if (elect_one_sync()) {
          for (int ks = 0; ks < nks; ++ks) umma_ts_f16(d, th + ks * 8, bh + ks * 128, idesc_pv, (acc | ks) ? 1u : 0u);
          for (int ks = 0; ks < nks; ++ks) umma_ts_f16(d, tl + ks * 8, bh + ks * 128, idesc_pv, 1u);
          for (int ks = 0; ks < nks; ++ks) umma_ts_f16(d, th + ks * 8, bl + ks * 128, idesc_pv, 1u);
          if (is_last) umma_commit(&o_done[i]);
          if (release_v) umma_commit(&v_empty[stage]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, T == 1, false);
      issue_qk(1, 0, T == 1, true);
      for (int j = 0; j < T; ++j) {
        const int s = j % Cfg::KV_STAGES;
        mbar_wait(&v_full[s], (j / Cfg::KV_STAGES) & 1);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_ready[i], j & 1);
          tc_fence_after();
          issue_pv(i, s, j + 1 == T, j > 0 ? 1u : 0u, i == 1);
          if (j + 1 < T) {
            const int s1 = (j + 1) % Cfg::KV_STAGES;
            if (i == 0) {
              mbar_wait(&k_full[s1], ((j + 1) / Cfg::KV_STAGES) & 1);
              tc_fence_after();
            }
            issue_qk(i, s1, j + 2 == T, i == 1);
          }
        }
      }
    }
  } else {
    // ------------------------------- softmax warpgroups --------------------------
    setmaxnreg_inc<208>();
    const int i = warp >> 2;        // query tile
    const int q = warp & 3;         // TMEM lane quarter
    const int r = q * 32 + lane;    // row inside the tile
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + i * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + i * 64 + lane_off;
    const float sc = p.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      const int valid = (j + 1 == T) ? last_valid : 128;
      if (valid > 32) attn_split_softmax_tile<128>(tS, tO, valid, j == 0, sc, m_used, l);
      else attn_split_softmax_tile<32>(tS, tO, valid, j == 0, sc, m_used, l);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[i]);
    }
    // ------------------------------- output ---------------------------------------
    mbar_wait(&o_done[i], 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q_row0 + i * Cfg::TQ + r;
    __nv_bfloat16* orow = p.out + (long long)b * p.out_bs + (long long)row * p.ldo + h * Cfg::HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + c * 32, o);
      tmem_wait_ld();
      if (row < p.Nq) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float a = __uint_as_float(o[e + 2 * u]) * inv, bb = __uint_as_float(o[e + 2 * u + 1]) * inv;
            hi[u] = pack_bf16x2(a, bb);
            lo[u] = pack_bf16x2(a - __uint_as_float(hi[u] << 16), bb - __uint_as_float(hi[u] & 0xffff0000u));
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + e) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(orow + p.out_lo_off + c * 32 + e) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dclip
