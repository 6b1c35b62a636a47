// Flash-attention forward, round-2 structure ("P4"): persistent CTA with FOUR softmax warpgroups.
//
// Why (profiles/r01_attention_notes.md, r02_attention_notes.md): at head_dim 64 a 128x128 tile costs 1024 SFU cycles but
// only 512 tensor cycles, and in the round-1 kernel ONE in-order warp per scheduler ran a tile's whole exponential stream
// (11.3 cycles per MUFU instead of 8) while its sibling warpgroup waited for the "MUFU token": the KV step was the serial
// chain of one warpgroup (2850 cycles for 2 x 1024 SFU cycles).  Here every scheduler holds four softmax warps that never
// synchronise with each other inside an item:
//
//   * work item = (image, head, 256 queries) as before; warpgroup w = 2*i + par owns query tile i (128 rows, one thread per
//     row) and the key half `par` of every 128-key K/V stage: split-K inside the CTA, flash-decoding style.  Each
//     warpgroup keeps its own running max / sum and its own O accumulator; the two halves of a query tile are merged
//     once per item in the output pass (O_b is read straight from TMEM, (m_b, l_b) travel through 2 KB of smem)
//   * 64-key sub-tiles: S_w (128 x 64 fp32) and O_w (128 x 64 fp32) = 128 TMEM columns per warpgroup, 512 in total.  P
//     ALIASES S: a thread pulls its S row into registers and writes the packed bf16 probabilities over the same columns;
//     the MMA warp issues PV(w, j) and then QK^T(w, j+1) into that buffer.  tcgen05.mma ops of one thread execute in
//     order, so "S(j+1) ready" implies "PV(j) done": no PV-done wait and no separate P buffer
//   * no MUFU token, no named barriers between warpgroups: four independent instruction streams per scheduler hide the
//     SFU / TMEM latencies of each other; the warpgroups drift freely (bounded by the 5-stage K/V ring)
//   * there is NO MMA warp: after a warpgroup-local named barrier ("all four warps have written their P rows") warp 0 of
//     the warpgroup issues PV(j) and QK^T(j+1) itself through one elected lane.  tcgen05 ordering is per issuing thread
//     and the warpgroups touch disjoint TMEM columns, so nothing else is needed; a dedicated MMA warp polling four
//     warpgroups was measured at ~1800 cycles per service (dependent probes + R2UR chains) and made the kernel 1.5x
//     SLOWER than round 1 (profiles/r02_attention_notes.md)
//   * everything else is inherited from the round-1 persistent kernel: static item list, Q double buffer, K/V TMA ring
//     alive across items, peeled key 0, output tile staged in the finished Q buffer + one TMA store per query tile, the
//     2049th query row served on CUDA cores by two otherwise idle warps in the background of the pipeline
#pragma once
#include "../../denseclip_vit_multimodal_b200/csrc/attn_tcgen05.cuh"

namespace dclip {

struct AttnP4Cfg {
  static constexpr int TQ = 128, TKV = 128, SUB = 64, HD = 64, KV_STAGES = 5;
  static constexpr int Q_OFF = 0;                               // 2 buffers x 2 tiles x 16 KB
  static constexpr int K_OFF = 4 * 16384;
  static constexpr int V_OFF = K_OFF + KV_STAGES * 16384;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * 16384;
  static constexpr int NUM_BARS = 4 + 2 * KV_STAGES + 8 + 4;
  static constexpr int ML_OFF = BAR_OFF + 320;                  // [2 query tiles][128 rows] float2 (m_b * sc, l_b)
  static constexpr int TAIL_OFF = ML_OFF + 2 * 128 * 8;         // 66 floats: the second tail warp's partial result
  static constexpr int SMEM_BYTES = TAIL_OFF + 272;
  static constexpr int THREADS = 640;                           // 16 softmax warps + TMA + TMEM allocator + 2 tail-row warps
  static constexpr int TMEM_COLS = 512;                         // S_w at 64 w, O_w at 256 + 64 w
};
static_assert(AttnP4Cfg::NUM_BARS * 8 + 8 <= 320 && AttnP4Cfg::SMEM_BYTES <= 232448, "P4 attention smem budget");

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Eight non-blocking mbarrier probes issued back to back (ONE shared-memory round trip for all of them; eight dependent
// test_wait + branch pairs cost ~8 round trips, and the MMA warp's service latency is on every warpgroup's critical path).
// Bit i of the result = phase `par[i]` of barrier `addr[i]` has completed.
__device__ __forceinline__ uint32_t mbar_test_wait_x8(const uint32_t (&addr)[8], const uint32_t (&par)[8]) {
  uint32_t mask;
  asm volatile(
      "{\n\t.reg .pred p0, p1, p2, p3, p4, p5, p6, p7;\n\t.reg .b32 t0, t1, t2, t3, t4, t5, t6, t7;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p0, [%1], %9;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p1, [%2], %10;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p2, [%3], %11;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p3, [%4], %12;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p4, [%5], %13;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p5, [%6], %14;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p6, [%7], %15;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p7, [%8], %16;\n\t"
      "selp.b32 t0, 1, 0, p0;\n\tselp.b32 t1, 2, 0, p1;\n\tselp.b32 t2, 4, 0, p2;\n\tselp.b32 t3, 8, 0, p3;\n\t"
      "selp.b32 t4, 16, 0, p4;\n\tselp.b32 t5, 32, 0, p5;\n\tselp.b32 t6, 64, 0, p6;\n\tselp.b32 t7, 128, 0, p7;\n\t"
      "or.b32 t0, t0, t1;\n\tor.b32 t2, t2, t3;\n\tor.b32 t4, t4, t5;\n\tor.b32 t6, t6, t7;\n\t"
      "or.b32 t0, t0, t2;\n\tor.b32 t4, t4, t6;\n\tor.b32 %0, t0, t4;\n\t}\n"
      : "=r"(mask)
      : "r"(addr[0]), "r"(addr[1]), "r"(addr[2]), "r"(addr[3]), "r"(addr[4]), "r"(addr[5]), "r"(addr[6]), "r"(addr[7]),
        "r"(par[0]), "r"(par[1]), "r"(par[2]), "r"(par[3]), "r"(par[4]), "r"(par[5]), "r"(par[6]), "r"(par[7])
      : "memory");
  return mask;
}

// One 64-key sub-tile of the online softmax for one query row: S (64 fp32 columns, TMEM) -> P (32 packed bf16x2 columns,
// written over the same TMEM buffer).  POLY of every 4 column pairs take their exp2 on the FMA pipe (exp2_poly_x2).
template <int POLY>
__device__ __forceinline__ void attn_p4_softmax_sub(uint32_t tS, uint32_t tO, int valid, bool first, float sc, float& m_used,
                                                    float& l) {
  constexpr int NC = AttnP4Cfg::SUB;
  uint32_t su[NC];
  tmem_ld_32x32b_x32(tS, reinterpret_cast<uint32_t(&)[32]>(su[0]));
  tmem_ld_32x32b_x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(su[32]));
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
  // lazy rescale: keep the old reference max unless it grew by more than 2^8 (first sub-tile: m_used = -inf -> always).
  // PV(j-1) of this warpgroup is complete (the S tile just read was produced by an MMA issued after it).
  const bool need = (m_new - m_used) * sc > 8.0f;
  if (__any_sync(0xffffffffu, need)) {
    const float alpha = ex2_approx((m_used - m_new) * sc);
    m_used = m_new;
    l *= alpha;
    if (!first) {   // (rare; 16 columns at a time: the whole S row is live in registers here)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c * 16, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st_32x32b_x16(tO + c * 16, o);
      }
    }
  }
  const uint64_t sc2 = pack_f32x2(sc, sc);
  const float nmc = -m_used * sc;
  const uint64_t nmc2 = pack_f32x2(nmc, nmc);
  uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
  uint32_t pk[8];
#pragma unroll
  for (int c8 = 0; c8 < NC / 8; ++c8) {
    float pv[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const uint64_t t = fma_f32x2(pack_f32x2(__uint_as_float(su[c8 * 8 + e]), __uint_as_float(su[c8 * 8 + e + 1])), sc2, nmc2);
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
    // lane = query row, 32-bit column c holds (P[2c], P[2c+1]): the A operand of the TS MMA
    pk[(c8 & 1) * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
    pk[(c8 & 1) * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
    pk[(c8 & 1) * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
    pk[(c8 & 1) * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
    if (c8 & 1) tmem_st_32x32b_x8(tS + (c8 >> 1) * 8, pk);   // 16 keys = 8 packed columns (= one K step of the PV MMA)
  }
  float a0, a1, a2, a3;
  unpack_f32x2(acc0, a0, a1);
  unpack_f32x2(acc1, a2, a3);
  l += (a0 + a1) + (a2 + a3);
}

// Tail rows in the background (two warps = 8 key-row groups of 8 threads), as attn_tail_rows_bg, but the four groups of a
// warp are merged with shuffles and only the second warp's partial goes through shared memory (272 B instead of 2.1 KB).
__device__ __forceinline__ void attn_tail_rows_bg2(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                                   __nv_bfloat16* out, int ldq, int ldk, int ldv, int ldo, int Nk, float scale_log2,
                                                   int row0, int nrows, float* s_part, int tid64) {
  const int g = tid64 >> 3, dq = tid64 & 7, lane = tid64 & 31, wsel = tid64 >> 5;
  for (int rr = 0; rr < nrows; ++rr) {
    const int row = row0 + rr;
    float qf[8];
    {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(q + (long long)row * ldq) + dq);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qf[2 * e] = __uint_as_float(ww[e] << 16) * scale_log2;
        qf[2 * e + 1] = __uint_as_float(ww[e] & 0xffff0000u) * scale_log2;
      }
    }
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const int trips = (Nk + 31) / 32;  // 8 groups x 4 keys per trip, warp-uniform
    for (int t = 0; t < trips; ++t) {
      uint4 kw[4], vw[4];
      float sj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = t * 32 + u * 8 + g;
        kw[u] = make_uint4(0u, 0u, 0u, 0u);
        vw[u] = kw[u];
        if (j < Nk) {
          kw[u] = __ldg(reinterpret_cast<const uint4*>(k + (long long)j * ldk) + dq);
          vw[u] = __ldg(reinterpret_cast<const uint4*>(v + (long long)j * ldv) + dq);
        }
      }
      float mt = m;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = t * 32 + u * 8 + g;
        const uint32_t ww[4] = {kw[u].x, kw[u].y, kw[u].z, kw[u].w};
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s0 = fmaf(qf[2 * e], __uint_as_float(ww[e] << 16), s0);
          s1 = fmaf(qf[2 * e + 1], __uint_as_float(ww[e] & 0xffff0000u), s1);
        }
        float x = s0 + s1;
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        sj[u] = j < Nk ? x : -INFINITY;
        mt = fmaxf(mt, sj[u]);
      }
      const float corr = ex2_approx(m - mt);  // m = -inf on the first trip -> 0 (mt finite: the host requires Nk >= 32)
      m = mt;
      l *= corr;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float pu = ex2_approx(sj[u] - m);
        l += pu;
        const uint32_t ww[4] = {vw[u].x, vw[u].y, vw[u].z, vw[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] = fmaf(pu, __uint_as_float(ww[e] << 16), acc[2 * e]);
          acc[2 * e + 1] = fmaf(pu, __uint_as_float(ww[e] & 0xffff0000u), acc[2 * e + 1]);
        }
      }
    }
    // merge the 4 groups of this warp (lanes differing in bits 3 and 4) with shuffles; l is held per group (all dq agree)
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mm = fmaxf(m, mo);
      const float wa = ex2_approx(m - mm), wb = ex2_approx(mo - mm);
      l = l * wa + lo * wb;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = acc[e] * wa + __shfl_xor_sync(0xffffffffu, acc[e], o) * wb;
      m = mm;
    }
    named_bar_sync(5, 64);  // previous row's exchange buffer fully consumed
    if (wsel == 1 && lane < 8) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s_part[dq * 8 + e] = acc[e];
      if (dq == 0) { s_part[64] = m; s_part[65] = l; }
    }
    named_bar_sync(5, 64);
    if (wsel == 0 && lane < 8) {
      const float mo = s_part[64], lo = s_part[65];
      const float mm = fmaxf(m, mo);
      const float wa = ex2_approx(m - mm), wb = ex2_approx(mo - mm);
      const float inv = 1.0f / (l * wa + lo * wb);
      uint32_t pkd[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        pkd[e] = pack_bf16x2((acc[2 * e] * wa + s_part[dq * 8 + 2 * e] * wb) * inv, (acc[2 * e + 1] * wa + s_part[dq * 8 + 2 * e + 1] * wb) * inv);
      *reinterpret_cast<uint4*>(out + (long long)row * ldo + dq * 8) = make_uint4(pkd[0], pkd[1], pkd[2], pkd[3]);
    }
  }
}

template <int POLY = 0>
__global__ void __launch_bounds__(AttnP4Cfg::THREADS, 1)
attn_fwd_p4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  using Cfg = AttnP4Cfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;                       // [2]  producer -> issuers : Q buffer landed
  uint64_t* q_empty = bars + 2;                  // [2]  4 issuer commits (last QK^T of the item) + the 2 output-store threads -> producer
  uint64_t* kv_full = bars + 4;                  // [S]  producer -> issuers : K and V tile of the stage landed (one barrier, 32 KB)
  uint64_t* kv_empty = kv_full + Cfg::KV_STAGES; // [S]  4 issuer commits (PV on the stage) -> producer
  uint64_t* s_full = kv_empty + Cfg::KV_STAGES;  // [4]  MMA -> softmax w : S sub-tile ready (and every earlier MMA of w complete)
  uint64_t* o_done = s_full + 4;                 // [4]  MMA -> softmax w : last PV of the item finished
  uint64_t* ob_ready = o_done + 4;               // [2]  softmax (i, 1) -> softmax (i, 0) : O_b final, (m_b, l_b) in smem
  uint64_t* ob_free = ob_ready + 2;              // [2]  softmax (i, 0) -> issuer (i, 1) : O_b and (m_b, l_b) consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ob_free + 2);
  float2* s_ml = reinterpret_cast<float2*>(smem + Cfg::ML_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G_CTAS = gridDim.x, cta = blockIdx.x;
  const int nqb = (p.Nq_total - p.q_start + 2 * Cfg::TQ - 1) / (2 * Cfg::TQ);
  const int rows_last = p.Nq_total - p.q_start - (nqb - 1) * 2 * Cfg::TQ;
  const int has_tail = rows_last <= p.tail_rows_max ? 1 : 0;
  const int nqb_reg = nqb - has_tail;
  const int n_reg = p.B * p.H * nqb_reg;          // host guarantees n_reg >= gridDim.x
  const int n_tail = p.B * p.H * has_tail;
  const int koff = p.peel_key0 ? 1 : 0;
  const int nk_eff = p.Nk - koff;
  const int T = (nk_eff + Cfg::TKV - 1) / Cfg::TKV;               // K/V stages per item
  const int last_valid = nk_eff - (T - 1) * Cfg::TKV;              // valid keys of the last stage (1..128)
  const int Tb = T - (last_valid <= Cfg::SUB ? 1 : 0);             // sub-tiles per item of the odd-half warpgroups
  // regular items r = first_r, first_r + G_CTAS, ...: the CTAs are walked backwards so that the ones that serve an extra
  // tail row (low ids) are not the ones that get the remainder of the regular items
  const int first_r = G_CTAS - 1 - cta;
  const int n_items = (n_reg - first_r + G_CTAS - 1) / G_CTAS;

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
      mbar_init(&q_empty[i], 6);
      mbar_init(&ob_ready[i], 4);
      mbar_init(&ob_free[i], 4);
    }
    for (int w = 0; w < 4; ++w) {
      mbar_init(&s_full[w], 1);
      mbar_init(&o_done[w], 1);
    }
    for (int s = 0; s < Cfg::KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    // (setmaxnreg must be the same instruction in all four warps of a warpgroup: 512 x 104 + 128 x 64 = the 640 x 96 pool)
    setmaxnreg_dec<64>();
    if (warp == 16) {
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
            mbar_arrive_expect_tx(&kv_full[s], 2 * 16384);
            tma_load_3d(smem + Cfg::K_OFF + s * 16384, &tmK, &kv_full[s], p.k_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
            tma_load_3d(smem + Cfg::V_OFF + s * 16384, &tmV, &kv_full[s], p.v_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
          }
        }
      }
    } else if (warp >= 18) {
      // ------------------------------- tail rows in the background ----------------
      float* s_part = reinterpret_cast<float*>(smem + Cfg::TAIL_OFF);
      for (int t = cta; t < n_tail; t += G_CTAS) {
        const int h = t % p.H, b = t / p.H;
        attn_tail_rows_bg2(p.q + (long long)b * p.q_bs + p.q_col0 + h * 64, p.k + (long long)b * p.k_bs + p.k_col0 + h * 64,
                           p.v + (long long)b * p.v_bs + p.v_col0 + h * 64, p.out + (long long)b * p.out_batch_stride + h * 64,
                           p.ldq, p.ldk, p.ldv, p.ldo, p.Nk, p.scale_log2, p.q_start + nqb_reg * 2 * Cfg::TQ, rows_last, s_part,
                           threadIdx.x - 576);
      }
    }
  } else {
    // ------------------------------- softmax warpgroups (each issues its own MMAs) --------------------------
    setmaxnreg_inc<104>();
    const int w = warp >> 2;        // warpgroup: query tile i = w >> 1, key half par = w & 1
    const int i = w >> 1, par = w & 1;
    const int q = warp & 3;         // TMEM lane quarter
    const int r = q * 32 + lane;    // row inside the query tile
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + w * 64 + lane_off;
    const uint32_t tO = tmem_base + 256 + w * 64 + lane_off;
    const float sc = p.scale_log2;
    const int Tw = par ? Tb : T;
    // MMA operands of this warpgroup (descriptor address fields are in units of 16 B: a 16 KB tile = 1024; the key half
    // `par` of a stage starts 64 rows x 128 B = 512 units into the K / V tile)
    const uint64_t dQw = make_smem_desc_sw128(smem_u32(smem + Cfg::Q_OFF), 16, 1024) + uint64_t(i) * 1024;
    const uint64_t dKw = make_smem_desc_sw128(smem_u32(smem + Cfg::K_OFF), 16, 1024) + uint64_t(par) * 512;
    const uint64_t dVw = make_smem_desc_sw128(smem_u32(smem + Cfg::V_OFF), 16, 1024) + uint64_t(par) * 512;
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major: head_dim contiguous
    const uint32_t dS = tmem_base + w * 64, dO = tmem_base + 256 + w * 64;
    // QK^T of sub-tile (item n, stage slot s): S_w = Q_i K_half^T
    auto issue_qk = [&](int n, int s, bool last_of_item) {
      const uint64_t a = dQw + uint64_t(n & 1) * 2048, bb = dKw + uint64_t(s) * 1024;
      if (elect_one_sync()) {
        umma_ss_f16(dS, a, bb, idesc_qk, 0u);
        umma_ss_f16(dS, a + 2, bb + 2, idesc_qk, 1u);
        umma_ss_f16(dS, a + 4, bb + 4, idesc_qk, 1u);
        umma_ss_f16(dS, a + 6, bb + 6, idesc_qk, 1u);
        umma_commit(&s_full[w]);
        if (last_of_item) umma_commit(&q_empty[n & 1]);   // this warpgroup no longer reads the item's Q buffer
      }
      __syncwarp();
    };
    int c = 0;          // sub-tiles processed by this warpgroup (s_full parity)
    int slot = 0;       // K/V ring slot of the pending sub-tile, and the parity of its kv_full phase
    uint32_t kvph = 0;
    if (q == 0 && n_items > 0 && Tw > 0) {   // the first QK^T of the first item
      mbar_wait(&q_full[0], 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, Tw == 1);
    }
    for (int n = 0; n < n_items; ++n) {
      const int ritem = first_r + n * G_CTAS;
      const int qb = ritem % nqb_reg, h = (ritem / nqb_reg) % p.H, b = ritem / (nqb_reg * p.H);
      const int q_row0 = p.q_start + qb * 2 * Cfg::TQ;
      float m_used = -INFINITY, l = 0.f;
      if (Tw == 0) {
        // odd half and no stage holds more than 64 keys (then T == 1): nothing to compute, but the producer counts this
        // warpgroup's arrival on every stage and Q buffer
        if (q == 0) {
          mbar_wait(&kv_full[slot], kvph);
          mbar_wait(&q_full[n & 1], (n >> 1) & 1);
          if (lane == 0) {
            mbar_arrive(&kv_empty[slot]);
            mbar_arrive(&q_empty[n & 1]);
          }
        }
        if (++slot == Cfg::KV_STAGES) { slot = 0; kvph ^= 1; }
        continue;
      }
      for (int j = 0; j < Tw; ++j, ++c) {
        long long* dbg = nullptr;
        DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && q == 0 && lane == 0 && c < 40) dbg = p.dbg + (w * 40 + c) * 8;)
        DCLIP_TL(if (dbg) dbg[0] = clock64();)
        mbar_wait(&s_full[w], c & 1);
        tc_fence_after();
        DCLIP_TL(if (dbg) dbg[1] = clock64();)
        int valid = Cfg::SUB;
        if (j + 1 == T) valid = min(Cfg::SUB, last_valid - par * Cfg::SUB);
        if (koff && par == 0 && lane == 0 && (j + 1 == Tw || j == 0)) {
          // k_0 / v_0 of this head are touched by nobody else: pull them into L2 at the start of the item and into L1 on
          // the last sub-tile, so the output pass finds them there
          const __nv_bfloat16* k0p = p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD;
          const __nv_bfloat16* v0p = p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD;
          if (j + 1 == Tw) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(k0p));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(v0p));
          } else {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(k0p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(v0p));
          }
        }
        attn_p4_softmax_sub<POLY>(tS, tO, valid, j == 0, sc, m_used, l);
        DCLIP_TL(if (dbg) dbg[2] = clock64();)
        tmem_wait_st();
        tc_fence_before();
        named_bar_sync(1 + w, 128);   // all four warps of the warpgroup have written their P rows (and rescaled O)
        DCLIP_TL(if (dbg) dbg[3] = clock64();)
        // position of the NEXT sub-tile of this warpgroup in the K/V ring (an odd-half warpgroup that skips the ragged
        // last stage of every item advances by two slots at the item boundary)
        const bool item_end = (j + 1 == Tw);
        int nslot = slot + 1;
        uint32_t nph = kvph;
        if (nslot == Cfg::KV_STAGES) { nslot = 0; nph ^= 1; }
        if (q == 0) {
          // ---- this warp issues the warpgroup's MMAs: PV(n, j), then QK^T of the next sub-tile into the freed S buffer ----
          tc_fence_after();
          if (par == 1 && j == 0 && n > 0) {   // the first PV of an item overwrites O_b: the partner must have merged it
            mbar_wait(&ob_free[i], (n - 1) & 1);
            tc_fence_after();
          }
          {
            const uint64_t bb = dVw + uint64_t(slot) * 1024;
            if (elect_one_sync()) {
              umma_ts_f16(dO, dS, bb, idesc_pv, j > 0 ? 1u : 0u);
              umma_ts_f16(dO, dS + 8, bb + 128, idesc_pv, 1u);
              umma_ts_f16(dO, dS + 16, bb + 256, idesc_pv, 1u);
              umma_ts_f16(dO, dS + 24, bb + 384, idesc_pv, 1u);
              if (item_end) umma_commit(&o_done[w]);
              umma_commit(&kv_empty[slot]);
            }
            __syncwarp();
          }
          DCLIP_TL(if (dbg) dbg[4] = clock64();)
          if (item_end && Tw < T) {
            // ragged last stage without keys for this half: wait until it has landed (so that the arrival below counts
            // for ITS phase of kv_empty) and release it unused
            mbar_wait(&kv_full[nslot], nph);
            if (lane == 0) mbar_arrive(&kv_empty[nslot]);
          }
        }
        if (item_end && Tw < T) {
          if (++nslot == Cfg::KV_STAGES) { nslot = 0; nph ^= 1; }
        }
        if (q == 0 && !(item_end && n + 1 == n_items)) {
          const int nn = item_end ? n + 1 : n;
          const int njj = item_end ? 0 : j + 1;
          mbar_wait(&kv_full[nslot], nph);
          if (njj == 0) mbar_wait(&q_full[nn & 1], (nn >> 1) & 1);
          tc_fence_after();
          issue_qk(nn, nslot, njj == Tw - 1);
          DCLIP_TL(if (dbg) dbg[5] = clock64();)
        }
        slot = nslot;
        kvph = nph;
        if (par == 0 && j == 0 && n > 0 && q == 0 && lane == 0) {
          // the previous item's output tile sits in its Q buffer until the TMA store has read it (long done by now);
          // only then may the producer refill that buffer (it needs it a whole item from now)
          tma_store_wait_read();
          mbar_arrive(&q_empty[(n - 1) & 1]);
        }
      }
      if (par == 1) {
        // ---- odd key half: publish (m_b, l_b); O_b stays in TMEM for the partner warpgroup ----
        if (Tw > 0) {
          mbar_wait(&o_done[w], n & 1);
          tc_fence_after();
          s_ml[i * 128 + r] = make_float2(m_used * sc, l);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ob_ready[i]);
        }
        continue;
      }
      // ---- even key half: merge the two halves (+ the peeled key 0) and write the output tile ----
      float z0 = 0.f;
      const uint4* v0 = reinterpret_cast<const uint4*>(p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD);
      if (koff) {
        const uint4* k0 = reinterpret_cast<const uint4*>(p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD);
        const uint8_t* qrow = smem + Cfg::Q_OFF + (n & 1) * 32768 + i * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((cc ^ (r & 7)) << 4));
          const uint4 kv = __ldg(k0 + cc);
          const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            s0 = fmaf(__uint_as_float(qq[e] << 16), __uint_as_float(kk[e] << 16), s0);
            s1 = fmaf(__uint_as_float(qq[e] & 0xffff0000u), __uint_as_float(kk[e] & 0xffff0000u), s1);
          }
        }
        z0 = (s0 + s1) * sc;
      }
      mbar_wait(&o_done[w], n & 1);
      float Mb = -INFINITY, lb = 0.f;
      if (Tb > 0) {
        mbar_wait(&ob_ready[i], n & 1);
        const float2 ml = s_ml[i * 128 + r];
        Mb = ml.x;
        lb = ml.y;
      }
      tc_fence_after();
      // common reference: M = max(M_a, M_b, z0); everything is renormalised to it, so nothing can overflow
      const float Ma = m_used * sc;
      float M = fmaxf(Ma, Mb);
      if (koff) M = fmaxf(M, z0);
      float wa = ex2_approx(Ma - M), wb = ex2_approx(Mb - M);      // (Mb = -inf -> 0)
      float p0 = koff ? ex2_approx(z0 - M) : 0.f;
      const float inv = 1.0f / (l * wa + lb * wb + p0);
      wa *= inv; wb *= inv; p0 *= inv;
      uint8_t* q_tile = smem + Cfg::Q_OFF + (n & 1) * 32768 + i * 16384;
      uint8_t* srow = q_tile + (r >> 3) * 1024 + (r & 7) * 128;
      const uint32_t tOb = tO + 64;   // the partner's accumulator (same lane quarter)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {   // 16 output columns at a time (register budget: 104 per thread)
        uint32_t oa[16], ob[16];
        tmem_ld_32x32b_x16(tO + cc * 16, oa);
        if (Tb > 0) tmem_ld_32x32b_x16(tOb + cc * 16, ob);
        tmem_wait_ld();
        if (cc == 3) {   // all of O_b (and (m_b, l_b)) is in registers: the partner's next item may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ob_free[i]);
        }
#pragma unroll
        for (int e = 0; e < 16; e += 8) {
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            f[u] = __uint_as_float(oa[e + u]) * wa;
            if (Tb > 0) f[u] = fmaf(__uint_as_float(ob[e + u]), wb, f[u]);
          }
          if (koff) {
            const uint4 vv = __ldg(v0 + cc * 2 + (e >> 3));   // (L1 hit: prefetched on the last sub-tile)
            const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              f[2 * u] = fmaf(p0, __uint_as_float(vw[u] << 16), f[2 * u]);
              f[2 * u + 1] = fmaf(p0, __uint_as_float(vw[u] & 0xffff0000u), f[2 * u + 1]);
            }
          }
          *reinterpret_cast<uint4*>(srow + (((cc * 2 + (e >> 3)) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
      named_bar_sync(6 + i, 128);
      if (q == 0 && lane == 0) {
        tma_store_3d(&tmO, q_tile, h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
        tma_store_commit();
      }
      tc_fence_before();  // the O reads are ordered before this warpgroup's next PV (which overwrites O)
    }
    if (par == 0 && q == 0 && lane == 0) tma_store_wait_read();  // last item's tile: shared memory must outlive the store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dclip
