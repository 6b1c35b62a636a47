// Flash-attention forward for sm_100a (head_dim 64, non-causal), replacing the softmax(QK^T)V inside
// nn.MultiheadAttention of the reference ViT block (segmentation/denseclip/models.py:287-289).
//
//   * one CTA = one (image, head, 256-query block): two 128-row query tiles share every K/V tile
//   * S = Q K^T and O += P V run on tcgen05.mma (bf16 in, fp32 accumulate in TMEM); S and O never leave the SM
//   * warp roles: warps 0-3 / 4-7 = softmax warpgroups for query tile 0 / 1 (one thread per query row), warp 8 MMA
//     issuer (one thread), warp 9 TMA producer (Q once, K/V 4-stage ring), warp 10 TMEM allocator
//   * softmax: fp32 online softmax in the exp2 domain with the 1/sqrt(d) scale folded in; O stays in TMEM and is
//     rescaled lazily (only when the running max grows by more than 2^8), P goes through swizzled smem as the
//     A operand of the PV MMA; the two warpgroups ping-pong so MMA time hides behind the MUFU-bound softmax
//   * the ragged tail (N = 2049 = 16*128 + 1) costs a 16-wide MMA, not a 17th full tile
#pragma once
#include "ptx.cuh"

// -DDCLIP_ATTN_TIMELINE compiles in clock64() stamps (selftest_attn timeline); off in the product build
#ifdef DCLIP_ATTN_TIMELINE
#define DCLIP_TL(...) __VA_ARGS__
#else
#define DCLIP_TL(...)
#endif

namespace dclip {

struct AttnParams {
  int B, H;
  int Nq_total;  // rows in the Q tensor (TMA bound along tokens)
  int q_start;   // first query row processed by this launch
  int Nk;        // number of keys/values
  int q_col0, k_col0, v_col0;  // column of head 0 inside a token row of the Q / K / V tensor
  float scale_log2;            // head_dim^-0.5 * log2(e)
  __nv_bfloat16* out;          // out[b][row][h*64 + d]
  long long out_batch_stride;  // elements
  int ldo;                     // elements
  int token_mode;              // 0: no MUFU token, 1: hand over after the exp phase, 2: hand over at 3/4 of it
  // raw operand pointers (same tensors as the TMA maps): used by the CUDA-core path that serves a query block with only a
  // few valid rows (the 2049th token of a 256-row blocking), see attn_tail_rows
  const __nv_bfloat16 *q, *k, *v;
  int ldq, ldk, ldv;
  long long q_bs, k_bs, v_bs;
  int tail_rows_max;           // query blocks with <= this many valid rows take the CUDA-core path (0 = never)
  int tail_overlap;            // persistent kernel: 1 = the tail rows are served by the two idle control warps WHILE the tensor-core
                               //    pipeline runs (attn_tail_rows_bg), 0 = by the whole CTA before the pipeline starts
  int peel_key0;               // 1: key 0 is handled on CUDA cores (score in the prologue, P*V in the output pass) and the KV tiles
                               //    cover keys 1..Nk-1 -- for Nk = 128 t + 1 (2049 tokens) this saves the whole ragged last tile
  long long* dbg;              // optional timeline buffer (selftest only): clock64 stamps of CTA `dbg_cta`
  int dbg_cta;
};

template <bool PT>  // PT: P lives in TMEM (TS MMA); otherwise P goes through swizzled shared memory (SS MMA)
struct AttnCfgT {
  static constexpr int TQ = 128, TKV = 128, HD = 64, KV_STAGES = PT ? 6 : 4;
  static constexpr int Q_OFF = 0;                               // 2 x 16 KB
  static constexpr int K_OFF = 2 * 16384;                       // KV_STAGES x 16 KB
  static constexpr int V_OFF = K_OFF + KV_STAGES * 16384;       // KV_STAGES x 16 KB
  static constexpr int P_OFF = V_OFF + KV_STAGES * 16384;       // 2 x 32 KB
  static constexpr int BAR_OFF = P_OFF + (PT ? 0 : 2 * 32768);
  static constexpr int NUM_BARS = 1 + 3 * KV_STAGES + 8;
  static constexpr int SMEM_BYTES = BAR_OFF + NUM_BARS * 8 + 16;
  static constexpr int THREADS = 384;
  static constexpr int TMEM_COLS = 512;  // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
};
using AttnCfg = AttnCfgT<true>;

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// One KV tile of the online softmax for one query row (= one thread): S (NC columns, fp32, TMEM) -> P (bf16, smem).
// NC = 128 for regular tiles, 32 for a short ragged tail.  `valid` < NC masks the trailing columns.
// exp2 of two values on the FMA / ALU pipes (no MUFU): Cody-Waite split t = n + f with the 1.5*2^23 rounding constant,
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 1.0e-4, 38x below bf16 rounding), exponent
// re-inserted with one integer multiply-add.  Offloads a fraction of the softmax exponentials from the SFU, which
// is the bottleneck of head_dim-64 attention (profiles/r01_attention_notes.md).
__device__ __forceinline__ void exp2_poly_x2(uint64_t t2, float& p0, float& p1) {
  float ta, tb;
  unpack_f32x2(t2, ta, tb);
  t2 = pack_f32x2(fmaxf(ta, -126.f), fmaxf(tb, -126.f));
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f);
  const uint64_t r2 = add_f32x2(t2, magic);                                        // low mantissa bits = round(t)
  const uint64_t n2 = add_f32x2(r2, pack_f32x2(-12582912.f, -12582912.f));         // round(t) as float
  const uint64_t f2 = fma_f32x2(n2, pack_f32x2(-1.f, -1.f), t2);                   // t - round(t)
  uint64_t q2 = fma_f32x2(pack_f32x2(0.05592203512787819f, 0.05592203512787819f), f2, pack_f32x2(0.24264007806777954f, 0.24264007806777954f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.6931210160255432f, 0.6931210160255432f));
  q2 = fma_f32x2(q2, f2, pack_f32x2(0.9999244809150696f, 0.9999244809150696f));
  float qa, qb, ra, rb;
  unpack_f32x2(q2, qa, qb);
  unpack_f32x2(r2, ra, rb);
  p0 = __uint_as_float(__float_as_uint(qa) + (__float_as_uint(ra) << 23));
  p1 = __uint_as_float(__float_as_uint(qb) + (__float_as_uint(rb) << 23));
}

// POLY = number of column pairs (of the 4 pairs in every group of 8 columns) whose exp2 runs on the FMA pipe
// DEFER (P in TMEM only): P(j) may only be written once PV(j-1) has consumed P(j-1) (single P buffer), and that MMA is
// issued ~300 cycles after this warpgroup published P(j-1) and completes ~600 cycles later -- on the critical chain of the
// warpgroup if waited for before the exponentials.  With DEFER the packed bf16 probabilities of the first 3/4 of the tile
// stay in registers (48), the PV-done barrier is probed at 1/2 and waited for at 3/4 of the exponential phase (by then
// it has long completed), and the TMEM stores of the first three chunks overlap the last quarter's exponentials.
template <int NC, bool PT, int POLY, bool DEFER = false>
__device__ __forceinline__ void attn_softmax_tile(uint32_t tS, uint32_t tO, uint32_t tP, uint8_t* sProw, int r, int lane, int valid,
                                                  bool first, float sc, float& m_used, float& l, uint64_t* s_free_bar,
                                                  uint64_t* o_done_bar, uint32_t o_done_parity, int wg, bool last_tile,
                                                  int token_mode, long long* dbg) {
  uint32_t su[NC];
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, reinterpret_cast<uint32_t(&)[32]>(su[c * 32]));
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_free_bar);
  DCLIP_TL(if (dbg) dbg[1] = clock64();)
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
  DCLIP_TL(if (dbg) dbg[2] = clock64();)
  // lazy rescale: keep the old reference max unless it grew by more than 2^8 (first tile: m_used = -inf -> always)
  const bool need = (m_new - m_used) * sc > 8.0f;
  const bool rescale = __any_sync(0xffffffffu, need);
  float alpha = 1.0f;
  if (rescale) {
    alpha = ex2_approx((m_used - m_new) * sc);
    m_used = m_new;
    l *= alpha;
  }
  // PV of the previous tile must be done before P is overwritten or O is touched.  It was issued when this warpgroup
  // arrived on p_ready a whole load+max phase ago, so this wait normally returns at once.
  constexpr bool kDefer = DEFER && PT;
  if (!first && (!kDefer || rescale)) {
    mbar_wait(o_done_bar, o_done_parity);
    tc_fence_after();
    if (rescale) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
        tmem_st_32x32b_x32(tO + c * 32, o);
      }
      tmem_wait_st();
    }
  }
  if (!kDefer) { DCLIP_TL(if (dbg) dbg[4] = clock64();) }
  // MUFU token: only one warpgroup at a time runs its exponential phase (the SFU pipe is the bottleneck at head_dim 64:
  // 128 ex2 per row per tile); the other one overlaps its TMEM loads / max / barrier work with it.
  if (token_mode) named_bar_sync(1 + wg, 256);
  // P = exp2(S * sc - m * sc) -> bf16 -> swizzled smem (A operand of the PV MMA), 8 columns (16 B) at a time
  const uint64_t sc2 = pack_f32x2(sc, sc);
  const float nmc = -m_used * sc;
  const uint64_t nmc2 = pack_f32x2(nmc, nmc);
  uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
  uint32_t pk[kDefer ? NC / 2 : 16];
  constexpr int WAIT_AT = NC == 128 ? 11 : NC / 8 - 1;  // 8-column group (of NC / 8) after which the deferred chunks are stored
  constexpr int PROBE_AT = NC == 128 ? 6 : -1;
  bool pv_done = false;
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
    if constexpr (kDefer) {
      pk[c16 * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
      pk[c16 * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
      pk[c16 * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
      pk[c16 * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
      if (c16 == PROBE_AT) pv_done = mbar_test_wait(o_done_bar, o_done_parity);  // parity of a fresh barrier's "previous" phase passes at j = 0
      if (c16 == WAIT_AT) {
        if (!pv_done) mbar_wait(o_done_bar, o_done_parity);
        tc_fence_after();
        DCLIP_TL(if (dbg) dbg[4] = clock64();)
#pragma unroll
        for (int ch = 0; ch <= WAIT_AT / 4; ++ch) tmem_st_32x32b_x16(tP + ch * 16, reinterpret_cast<uint32_t(&)[16]>(pk[ch * 16]));
      } else if (c16 > WAIT_AT && (c16 & 3) == 3) {
        tmem_st_32x32b_x16(tP + (c16 >> 2) * 16, reinterpret_cast<uint32_t(&)[16]>(pk[(c16 >> 2) * 16]));
      }
    } else if constexpr (PT) {
      // P stays on-chip in TMEM: lane = query row, 32-bit column c holds (P[2c], P[2c+1]) -- the A operand of the TS MMA
      pk[(c16 & 3) * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
      pk[(c16 & 3) * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
      pk[(c16 & 3) * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
      pk[(c16 & 3) * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
      if ((c16 & 3) == 3) tmem_st_32x32b_x16(tP + (c16 >> 2) * 16, pk);
    } else {
      *reinterpret_cast<uint4*>(sProw + (c16 >> 3) * 16384 + (((c16 & 7) ^ (r & 7)) << 4)) =
          make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
    }
    // token_mode 2: early hand-over once 3/4 of this tile's exponentials are issued
    if (token_mode == 2 && c16 == (3 * NC / 32) && (!last_tile || wg == 0)) named_bar_arrive(2 - wg, 256);
  }
  if (token_mode == 1 && (!last_tile || wg == 0)) named_bar_arrive(2 - wg, 256);  // hand the token over (WG1 skips its final, unmatched arrive)
  float a0, a1, a2, a3;
  unpack_f32x2(acc0, a0, a1);
  unpack_f32x2(acc1, a2, a3);
  l += (a0 + a1) + (a2 + a3);
  DCLIP_TL(if (dbg) dbg[3] = clock64();)
}

#ifdef DCLIP_EXPERIMENTS
// Speculative-max variant of attn_softmax_tile for KV tiles j >= 1.  The serial chain of one warpgroup per tile
// (barrier wait -> TMEM load -> row max -> PV-done wait -> exponentials -> publish) is longer than the SFU work of the two
// warpgroups together, so the chain, not the SFU, sets the step time (profiles/r01_attention_notes.md).  Here the
// exponentials run against the STALE reference maximum m_used straight after the load; the tile's row maximum is reduced
// inside the same instruction stream (FMNMX3 in the shadow of the MUFU issue slots) and only checked afterwards.  The
// lazy-rescale rule is unchanged (rescale when the maximum grew by more than 2^8), so the result is identical to the
// classic order: in the rare case a rescale is due, O and l are rescaled and the tile's exponentials are redone (the
// scores are still in registers).  The PV(j-1)-done round trip is issued right behind the TMEM loads and overlaps their
// latency.
// DEFER: as in attn_softmax_tile, but the scores stay live for the redo, so only the first half of the tile is held back.
template <int NC, bool PT, int POLY, bool DEFER = false>
__device__ __forceinline__ void attn_softmax_tile_spec(uint32_t tS, uint32_t tO, uint32_t tP, uint8_t* sProw, int r, int lane,
                                                       int valid, float sc, float& m_used, float& l, uint64_t* s_free_bar,
                                                       uint64_t* o_done_bar, uint32_t o_done_parity, int wg, bool last_tile,
                                                       int token_mode, long long* dbg) {
  uint32_t su[NC];
#pragma unroll
  for (int c = 0; c < NC / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, reinterpret_cast<uint32_t(&)[32]>(su[c * 32]));
  constexpr bool kDefer = DEFER && PT;
  if constexpr (!kDefer) {
    mbar_wait(o_done_bar, o_done_parity);  // P (and O) are free again; ~120-cycle round trip hidden behind the loads
    tc_fence_after();
    DCLIP_TL(if (dbg) dbg[4] = clock64();)
  }
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(s_free_bar);
  DCLIP_TL(if (dbg) dbg[1] = clock64();)
  if (valid < NC) {
#pragma unroll
    for (int e = 0; e < NC; ++e)
      if (e >= valid) su[e] = 0xff800000u;  // -inf
  }
  DCLIP_TL(if (dbg) dbg[2] = clock64();)
  if (token_mode) named_bar_sync(1 + wg, 256);
  const uint64_t sc2 = pack_f32x2(sc, sc);
  uint64_t acc0, acc1;
#pragma unroll 1
  for (int pass = 0;; ++pass) {
    const float nmc = -m_used * sc;
    const uint64_t nmc2 = pack_f32x2(nmc, nmc);
    acc0 = pack_f32x2(0.f, 0.f);
    acc1 = acc0;
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
    uint32_t pk[kDefer ? NC / 2 : 16];
    constexpr int WAIT_AT = NC == 128 ? 7 : NC / 8 - 1;
    constexpr int PROBE_AT = NC == 128 ? 3 : -1;
    bool pv_done = pass != 0;
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
      mx0 = fmaxf(mx0, fmaxf(__uint_as_float(su[c16 * 8 + 0]), __uint_as_float(su[c16 * 8 + 1])));
      mx1 = fmaxf(mx1, fmaxf(__uint_as_float(su[c16 * 8 + 2]), __uint_as_float(su[c16 * 8 + 3])));
      mx2 = fmaxf(mx2, fmaxf(__uint_as_float(su[c16 * 8 + 4]), __uint_as_float(su[c16 * 8 + 5])));
      mx3 = fmaxf(mx3, fmaxf(__uint_as_float(su[c16 * 8 + 6]), __uint_as_float(su[c16 * 8 + 7])));
      acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(pv[0], pv[1]), pack_f32x2(pv[2], pv[3])));
      acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(pv[4], pv[5]), pack_f32x2(pv[6], pv[7])));
      if constexpr (kDefer) {
        pk[c16 * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
        pk[c16 * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
        pk[c16 * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
        pk[c16 * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
        if (c16 == PROBE_AT && !pv_done) pv_done = mbar_test_wait(o_done_bar, o_done_parity);
        if (c16 == WAIT_AT) {
          if (!pv_done) mbar_wait(o_done_bar, o_done_parity);
          tc_fence_after();
          DCLIP_TL(if (dbg && !pass) dbg[4] = clock64();)
#pragma unroll
          for (int ch = 0; ch <= WAIT_AT / 4; ++ch) tmem_st_32x32b_x16(tP + ch * 16, reinterpret_cast<uint32_t(&)[16]>(pk[ch * 16]));
        } else if (c16 > WAIT_AT && (c16 & 3) == 3) {
          tmem_st_32x32b_x16(tP + (c16 >> 2) * 16, reinterpret_cast<uint32_t(&)[16]>(pk[(c16 >> 2) * 16]));
        }
      } else if constexpr (PT) {
        pk[(c16 & 3) * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
        pk[(c16 & 3) * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
        pk[(c16 & 3) * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
        pk[(c16 & 3) * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
        if ((c16 & 3) == 3) tmem_st_32x32b_x16(tP + (c16 >> 2) * 16, pk);
      } else {
        *reinterpret_cast<uint4*>(sProw + (c16 >> 3) * 16384 + (((c16 & 7) ^ (r & 7)) << 4)) =
            make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
      }
    }
    if (pass) break;
    const float m_new = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
    const bool need = (m_new - m_used) * sc > 8.0f;
    if (!__any_sync(0xffffffffu, need)) break;
    // rare: the reference maximum moved by more than 2^8 -- rescale O and l, redo this tile's exponentials
    const float alpha = ex2_approx((m_used - m_new) * sc);
    m_used = m_new;
    l *= alpha;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + c * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
      tmem_st_32x32b_x32(tO + c * 32, o);
    }
    tmem_wait_st();
  }
  if (token_mode && (!last_tile || wg == 0)) named_bar_arrive(2 - wg, 256);
  float a0, a1, a2, a3;
  unpack_f32x2(acc0, a0, a1);
  unpack_f32x2(acc1, a2, a3);
  l += (a0 + a1) + (a2 + a3);
  DCLIP_TL(if (dbg) dbg[3] = clock64();)
}

#endif  // DCLIP_EXPERIMENTS

// ---------------------------------------------------------------------------------------------------------
// CUDA-core path for a query block that holds only a few valid rows (N = 2049 tokens = 8 x 256 + 1: without it the ninth
// CTA of every (image, head) would sweep all keys on the tensor cores for a single row, 11% of the grid).  The whole CTA
// (384 threads) serves one query row at a time: thread-per-key dot products -> fp32 scores in shared memory -> block
// softmax -> key-group-parallel PV with a shared-memory reduction.  Runs before any barrier / TMEM set-up.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float attn_block_reduce(float v, bool is_max, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();  // s_red reuse
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float r = s_red[0];
  for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, s_red[w]) : r + s_red[w];
  return r;
}

struct AttnTailArgs {  // by-value subset of AttnParams, pre-offset to (image, head) -- keeps the kernel parameters out of local memory
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  int ldq, ldk, ldv, ldo, Nk;
  float scale_log2;
};

__device__ __noinline__ void attn_tail_rows(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, __nv_bfloat16* out,
                                            int ldq, int ldk, int ldv, int ldo, int Nk, float scale_log2, int row0, int nrows,
                                            uint8_t* smem) {
  const AttnTailArgs p{q, k, v, out, ldq, ldk, ldv, ldo, Nk, scale_log2};
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int KG = nthr / 8;                                   // key groups of the PV phase (48)
  float* s_part = reinterpret_cast<float*>(smem);            // [KG][64] partial outputs
  float* s_red = s_part + KG * 64;                           // [32]
  float* s_sc = s_red + 32;                                  // [Nk] scores / probabilities
  const __nv_bfloat16* kb = p.k;
  const __nv_bfloat16* vb = p.v;
  for (int rr = 0; rr < nrows; ++rr) {
    const int row = row0 + rr;
    // 8 threads per key row (16 B each): a warp instruction touches 4 whole 128 B lines
    const int g = tid >> 3, dq = tid & 7;
    float qf[8];
    {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.q + (long long)row * p.ldq) + dq);
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qf[2 * e] = __uint_as_float(ww[e] << 16) * p.scale_log2;
        qf[2 * e + 1] = __uint_as_float(ww[e] & 0xffff0000u) * p.scale_log2;
      }
    }
    float lmax = -INFINITY;
    const int trips = (p.Nk + KG - 1) / KG;  // warp-uniform trip count: the shuffles below need the whole warp
    constexpr int NB = 16;  // independent 16 B loads in flight per thread before the first use
    for (int t0 = 0; t0 < trips; t0 += NB) {
      uint4 w[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = g + (t0 + u) * KG;
        w[u] = make_uint4(0u, 0u, 0u, 0u);
        if (j < p.Nk) w[u] = __ldg(reinterpret_cast<const uint4*>(kb + (long long)j * p.ldk) + dq);
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = g + (t0 + u) * KG;
        const uint32_t ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s0 = fmaf(qf[2 * e], __uint_as_float(ww[e] << 16), s0);
          s1 = fmaf(qf[2 * e + 1], __uint_as_float(ww[e] & 0xffff0000u), s1);
        }
        float sj = s0 + s1;
        sj += __shfl_xor_sync(0xffffffffu, sj, 1);
        sj += __shfl_xor_sync(0xffffffffu, sj, 2);
        sj += __shfl_xor_sync(0xffffffffu, sj, 4);
        if (j < p.Nk) {
          if (dq == 0) s_sc[j] = sj;
          lmax = fmaxf(lmax, sj);
        }
      }
    }
    const float m = attn_block_reduce(lmax, true, s_red);
    // (the reduction's barriers also publish s_sc)  P = exp2(s - m) is evaluated by all 8 threads of a key row; the row
    // sum counts it once
    float lsum = 0.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int t0 = 0; t0 < trips; t0 += NB) {
      uint4 w[NB];
      float pj[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int j = g + (t0 + u) * KG;
        w[u] = make_uint4(0u, 0u, 0u, 0u);
        pj[u] = -INFINITY;
        if (j < p.Nk) {
          w[u] = __ldg(reinterpret_cast<const uint4*>(vb + (long long)j * p.ldv) + dq);
          pj[u] = s_sc[j];
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const float pu = ex2_approx(pj[u] - m);
        lsum += pu;
        const uint32_t ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] = fmaf(pu, __uint_as_float(ww[e] << 16), acc[2 * e]);
          acc[2 * e + 1] = fmaf(pu, __uint_as_float(ww[e] & 0xffff0000u), acc[2 * e + 1]);
        }
      }
    }
    const float l = attn_block_reduce(dq == 0 ? lsum : 0.f, false, s_red);
#pragma unroll
    for (int e = 0; e < 8; ++e) s_part[g * 64 + dq * 8 + e] = acc[e];
    __syncthreads();
    {  // two-level sum over the key groups: nthr / 64 slices first
      const int d = tid & 63, slice = tid >> 6, nslice = nthr >> 6;
      float o = 0.f;
      for (int gg = slice; gg < KG; gg += nslice) o += s_part[gg * 64 + d];
      __syncthreads();
      s_part[slice * 64 + d] = o;
      __syncthreads();
      if (tid < 64) {
        o = 0.f;
        for (int sl = 0; sl < nslice; ++sl) o += s_part[sl * 64 + tid];
        p.out[(long long)row * p.ldo + tid] = __float2bfloat16(o / l);
      }
    }
    __syncthreads();
  }
}


// Output pass of one query tile: O (TMEM) * alpha + p0 * v_0 -> bf16 -> the thread's own row of the (finished) Q tile in
// shared memory (same SWIZZLE_128B layout the TMA load produced) -> ONE TMA store per warpgroup.  The direct version
// (each thread storing its 128-byte row with 16-byte st.global) cost ~4500 cycles per CTA: every warp instruction
// touched 32 different lines, 4096 LSU wavefronts per CTA.  Rows beyond the tensor are clipped by the TMA unit.
__device__ __forceinline__ void attn_store_tile(uint32_t tO, uint8_t* q_tile, int r, float alpha, float p0, bool fold,
                                                const uint4 (&v0r)[8]) {
  uint8_t* srow = q_tile + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // (both 32-column loads in flight at once measured slower: register pressure)
    uint32_t o[32];
    tmem_ld_32x32b_x32(tO + c * 32, o);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 32; e += 8) {
      float f[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(o[e + u]) * alpha;
      if (fold) {
        const uint4 vv = v0r[c * 4 + (e >> 3)];
        const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          f[2 * u] = fmaf(p0, __uint_as_float(vw[u] << 16), f[2 * u]);
          f[2 * u + 1] = fmaf(p0, __uint_as_float(vw[u] & 0xffff0000u), f[2 * u + 1]);
        }
      }
      *reinterpret_cast<uint4*>(srow + (((c * 4 + (e >> 3)) ^ (r & 7)) << 4)) =
          make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
  }
  fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
}

// MODE bit 0: speculative-max tiles for j >= 1 (attn_softmax_tile_spec); bit 1: deferred P stores (DEFER)
template <bool PT, int POLY = 0, int MODE = 0>
__global__ void __launch_bounds__(AttnCfgT<PT>::THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  using Cfg = AttnCfgT<PT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + Cfg::KV_STAGES;
  uint64_t* kv_empty = v_full + Cfg::KV_STAGES;
  uint64_t* s_full = kv_empty + Cfg::KV_STAGES;  // [2]  MMA  -> softmax : S tile ready in TMEM
  uint64_t* s_free = s_full + 2;                 // [2]  softmax -> MMA : S tile copied to registers
  uint64_t* p_ready = s_free + 2;                // [2]  softmax -> MMA : P tile in smem (and O rescaled)
  uint64_t* o_done = p_ready + 2;                // [2]  MMA  -> softmax : PV accumulate finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqb = (p.Nq_total - p.q_start + 2 * Cfg::TQ - 1) / (2 * Cfg::TQ);
  const int qb = blockIdx.x % nqb;
  const int h = (blockIdx.x / nqb) % p.H;
  const int b = blockIdx.x / (nqb * p.H);
  const int q_row0 = p.q_start + qb * 2 * Cfg::TQ;
  const int koff = p.peel_key0 ? 1 : 0;                // first key of the tensor-core tiles
  const int nk_eff = p.Nk - koff;
  const int T = (nk_eff + Cfg::TKV - 1) / Cfg::TKV;
  const int last_valid = nk_eff - (T - 1) * Cfg::TKV;   // valid columns of the last KV tile (1..128)
  const int last_cols16 = (last_valid + 15) & ~15;      // MMA extent of the last KV tile

  if (p.Nq_total - q_row0 <= p.tail_rows_max) {  // CTA-uniform: a block with a handful of valid rows skips the tensor-core machinery
    attn_tail_rows(p.q + (long long)b * p.q_bs + p.q_col0 + h * 64, p.k + (long long)b * p.k_bs + p.k_col0 + h * 64,
                   p.v + (long long)b * p.v_bs + p.v_col0 + h * 64, p.out + (long long)b * p.out_batch_stride + h * 64, p.ldq, p.ldk,
                   p.ldv, p.ldo, p.Nk, p.scale_log2, q_row0, p.Nq_total - q_row0, smem);
    return;
  }
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("dclip attn: dynamic smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  // warp roles: 0-3 softmax WG0, 4-7 softmax WG1, 8 MMA issuer, 9 TMA producer, 10 TMEM allocator, 11 idle.
  // The control warps sit at the HIGH warp ids: the issue arbiter favours higher warp ids, and a late MMA issue
  // stalls both softmax warpgroups.
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
        mbar_arrive_expect_tx(q_full, 2 * 16384);
        tma_load_3d(smem + Cfg::Q_OFF, &tmQ, q_full, p.q_col0 + h * Cfg::HD, q_row0, b);
        tma_load_3d(smem + Cfg::Q_OFF + 16384, &tmQ, q_full, p.q_col0 + h * Cfg::HD, q_row0 + Cfg::TQ, b);
        for (int j = 0; j < T; ++j) {
          const int s = j % Cfg::KV_STAGES;
          const uint32_t ph = (j / Cfg::KV_STAGES) & 1;
          mbar_wait(&kv_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], 16384);
          tma_load_3d(smem + Cfg::K_OFF + s * 16384, &tmK, &k_full[s], p.k_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
          mbar_arrive_expect_tx(&v_full[s], 16384);
          tma_load_3d(smem + Cfg::V_OFF + s * 16384, &tmV, &v_full[s], p.v_col0 + h * Cfg::HD, koff + j * Cfg::TKV, b);
        }
      }
    } else if (warp == 8) {
      // ------------------------------- MMA issuer ----------------------------------
      // The whole warp walks the schedule (so every value stays warp-uniform and lives in uniform registers); the
      // tcgen05 instructions themselves are issued by one elected lane.
      // descriptors are built once; per-MMA operands are base + small constants (units of 16 B in the address field)
      const uint64_t dQ = make_smem_desc_sw128(smem_u32(smem + Cfg::Q_OFF), 16, 1024);
      const uint64_t dK = make_smem_desc_sw128(smem_u32(smem + Cfg::K_OFF), 16, 1024);
      const uint64_t dV = make_smem_desc_sw128(smem_u32(smem + Cfg::V_OFF), 16, 1024);
      const uint64_t dP = make_smem_desc_sw128(smem_u32(smem + Cfg::P_OFF), 16, 1024);
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major: head_dim contiguous
      const uint32_t idesc_qk_last = make_idesc_bf16(128, last_cols16);
      auto issue_qk = [&](int i, int stage, bool is_last) {
        const uint64_t a = dQ + uint64_t(i) * 1024, bb = dK + uint64_t(stage) * 1024;
        const uint32_t idesc = is_last ? idesc_qk_last : idesc_qk;
        const uint32_t d = tmem_base + i * 128;
        if (elect_one_sync()) {
          umma_ss_f16(d, a, bb, idesc, 0u);
          umma_ss_f16(d, a + 2, bb + 2, idesc, 1u);
          umma_ss_f16(d, a + 4, bb + 4, idesc, 1u);
          umma_ss_f16(d, a + 6, bb + 6, idesc, 1u);
          umma_commit(&s_full[i]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int i, int stage, bool is_last, uint32_t acc, bool release_kv) {
        const uint64_t a = dP + uint64_t(i) * 2048, bb = dV + uint64_t(stage) * 1024;
        const uint32_t d = tmem_base + 256 + i * 64;
        if (elect_one_sync()) {
        if constexpr (PT) {
          const uint32_t ta = tmem_base + 384 + i * 64;  // 16 bf16 of K per MMA = 8 TMEM columns
          if (!is_last || last_cols16 == 128) {
            umma_ts_f16(d, ta, bb, idesc_pv, acc);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) umma_ts_f16(d, ta + ks * 8, bb + ks * 128, idesc_pv, 1u);
          } else {
            for (int ks = 0; ks < last_cols16 / 16; ++ks)
              umma_ts_f16(d, ta + ks * 8, bb + ks * 128, idesc_pv, (acc | ks) ? 1u : 0u);
          }
        } else if (!is_last || last_cols16 == 128) {
          umma_ss_f16(d, a, bb, idesc_pv, acc);
#pragma unroll
          for (int ks = 1; ks < 8; ++ks)
            umma_ss_f16(d, a + (ks >> 2) * 1024 + (ks & 3) * 2, bb + ks * 128, idesc_pv, 1u);
        } else {
          for (int ks = 0; ks < last_cols16 / 16; ++ks)
            umma_ss_f16(d, a + (ks >> 2) * 1024 + (ks & 3) * 2, bb + ks * 128, idesc_pv, (acc | ks) ? 1u : 0u);
        }
        umma_commit(&o_done[i]);
        if (release_kv) umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
      };
      // Issue order: S tiles are produced two KV tiles ahead of their use (S_i(j+2) is issued as soon as warpgroup i
      // has pulled S_i(j+1) into registers), so the softmax warpgroups never wait on QK^T; PV(i,j) goes out as soon
      // as P_i(j) is in smem.  The two warpgroups may drift by up to a tile without blocking each other.
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0, T == 1);
      issue_qk(1, 0, T == 1);
      if (T > 1) {
        mbar_wait(&k_full[1 % Cfg::KV_STAGES], (1 / Cfg::KV_STAGES) & 1);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&s_free[i], 0);
          tc_fence_after();
          issue_qk(i, 1 % Cfg::KV_STAGES, T == 2);
        }
      }
      for (int j = 0; j < T; ++j) {
        const int s = j % Cfg::KV_STAGES;
        mbar_wait(&v_full[s], (j / Cfg::KV_STAGES) & 1);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_ready[i], j & 1);
          tc_fence_after();
          DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta) p.dbg[512 + (i * 32 + j) * 2] = clock64();)
          issue_pv(i, s, j + 1 == T, j > 0 ? 1u : 0u, i == 1);
          if (j + 2 < T) {
            const int s2 = (j + 2) % Cfg::KV_STAGES;
            if (i == 0) mbar_wait(&k_full[s2], ((j + 2) / Cfg::KV_STAGES) & 1);
            mbar_wait(&s_free[i], (j + 1) & 1);
            tc_fence_after();
            DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta) p.dbg[512 + (i * 32 + j + 2) * 2 + 1] = clock64();)
            issue_qk(i, s2, j + 3 == T);
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
    uint8_t* sProw = smem + Cfg::P_OFF + i * 32768 + (r >> 3) * 1024 + (r & 7) * 128;  // (unused when PT)
    const uint32_t tP = tmem_base + 384 + i * 64 + lane_off;
    const float sc = p.scale_log2;
    float m_used = -INFINITY, l = 0.f;
    if (i == 1 && p.token_mode) named_bar_arrive(1, 256);  // WG0 owns the MUFU token first
    for (int j = 0; j < T; ++j) {
      DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && (warp & 3) == 0 && lane == 0) p.dbg[(i * 32 + j) * 8 + 6] = clock64();)
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      const int valid = (j + 1 == T) ? last_valid : 128;
      // peeled key 0 (koff): its score and its P*V term are computed on CUDA cores in the output pass; one tile ahead, park
      // the two 128 B lines (k_0, v_0 of this head) in L1 -- the K/V stream goes through TMA and never touches L1
      if (koff && lane == 0 && (j + 1 == T || (p.peel_key0 == 2 && (j == 0 || j + 2 == T)))) {
        const __nv_bfloat16* k0p = p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD;
        const __nv_bfloat16* v0p = p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD;
        if (j == 0) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(k0p));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(v0p));
        } else {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(k0p));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(v0p));
        }
      }
      long long* dbg = nullptr;
      DCLIP_TL(dbg = (p.dbg && blockIdx.x == p.dbg_cta && (warp & 3) == 0 && lane == 0) ? p.dbg + (i * 32 + j) * 8 : nullptr;)
      DCLIP_TL(if (dbg) dbg[0] = clock64();)
      constexpr bool DEFER = (MODE & 2) != 0;
#ifdef DCLIP_EXPERIMENTS
      constexpr bool SPEC = (MODE & 1) != 0;
      if (SPEC && j > 0) {
        if (valid > 32)
          attn_softmax_tile_spec<128, PT, POLY, DEFER>(tS, tO, tP, sProw, r, lane, valid, sc, m_used, l, &s_free[i], &o_done[i], (j - 1) & 1, i, j + 1 == T, p.token_mode, dbg);
        else
          attn_softmax_tile_spec<32, PT, POLY, DEFER>(tS, tO, tP, sProw, r, lane, valid, sc, m_used, l, &s_free[i], &o_done[i], (j - 1) & 1, i, j + 1 == T, p.token_mode, dbg);
      } else
#endif
      if (valid > 32)
        attn_softmax_tile<128, PT, POLY, DEFER>(tS, tO, tP, sProw, r, lane, valid, j == 0, sc, m_used, l, &s_free[i], &o_done[i], (j - 1) & 1, i, j + 1 == T, p.token_mode, dbg);
      else
        attn_softmax_tile<32, PT, POLY, DEFER>(tS, tO, tP, sProw, r, lane, valid, j == 0, sc, m_used, l, &s_free[i], &o_done[i], (j - 1) & 1, i, j + 1 == T, p.token_mode, dbg);
      if constexpr (PT) tmem_wait_st(); else fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[i]);
      DCLIP_TL(if (dbg) dbg[5] = clock64();)
    }

    // ------------------------------- output ---------------------------------------
    // peeled key: z0 = scale_log2 * (q_r . k_0) from the Q row still sitting in the swizzled smem tile and k_0 from L1;
    // everything is issued before the PV-done wait so it overlaps the last MMA
    float z0 = 0.f;
    uint4 v0r[8];
    if (koff) {
      const uint4* k0 = reinterpret_cast<const uint4*>(p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD);
      const uint4* v0 = reinterpret_cast<const uint4*>(p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD);
      const uint8_t* qrow = smem + Cfg::Q_OFF + i * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
        const uint4 kv = __ldg(k0 + c);
        v0r[c] = __ldg(v0 + c);
        const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s0 = fmaf(__uint_as_float(qq[e] << 16), __uint_as_float(kk[e] << 16), s0);
          s1 = fmaf(__uint_as_float(qq[e] & 0xffff0000u), __uint_as_float(kk[e] & 0xffff0000u), s1);
        }
      }
      z0 = (s0 + s1) * sc;
    }
    mbar_wait(&o_done[i], (T - 1) & 1);
    tc_fence_after();
    // fold the peeled key in: renormalise to max(m_used, z0) so nothing can overflow
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
    // this query tile's smem (all QK^T that read it are complete) stages the bf16 output for one TMA store per warpgroup
    uint8_t* q_tile = smem + Cfg::Q_OFF + i * 16384;
    attn_store_tile(tO, q_tile, r, alpha, p0, koff != 0, v0r);
    named_bar_sync(3 + i, 128);
    if (q == 0 && lane == 0) {
      tma_store_3d(&tmO, q_tile, h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
      tma_store_commit();
      tma_store_wait_read();  // the CTA (and its shared memory) may go away once every thread is through
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// Tail rows in the background of the persistent pipeline: warps 10 and 11 (64 threads, 80 registers) = 8 key-row groups of
// 8 threads.  Group g walks keys g, g + 8, ... four at a time with an online softmax (running max / sum / 8 output dims per
// thread), so no per-key score buffer is needed; the 8 partial results are merged through 2.1 KB of shared memory.
// ~60 us per row, hidden behind the ~300 us the tensor-core pipeline of the same CTA runs.
__device__ __forceinline__ void attn_tail_rows_bg(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                                  __nv_bfloat16* out, int ldq, int ldk, int ldv, int ldo, int Nk, float scale_log2,
                                                  int row0, int nrows, float* s_merge, int tid64) {
  const int g = tid64 >> 3, dq = tid64 & 7;
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
      const float corr = ex2_approx(m - mt);  // m = -inf on the first trip -> 0 (mt is finite: key g < Nk exists for Nk >= 8)
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
    // merge the 8 groups: s_merge[g][0..63] = acc, [64] = m, [65] = l
    named_bar_sync(5, 64);  // previous row's merge buffer fully consumed
#pragma unroll
    for (int e = 0; e < 8; ++e) s_merge[g * 66 + dq * 8 + e] = acc[e];
    if (dq == 0) { s_merge[g * 66 + 64] = m; s_merge[g * 66 + 65] = l; }
    named_bar_sync(5, 64);
    float mm = -INFINITY;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) mm = fmaxf(mm, s_merge[gg * 66 + 64]);
    float o = 0.f, lt = 0.f;
#pragma unroll
    for (int gg = 0; gg < 8; ++gg) {
      const float w = ex2_approx(s_merge[gg * 66 + 64] - mm);
      o = fmaf(w, s_merge[gg * 66 + tid64], o);
      lt = fmaf(w, s_merge[gg * 66 + 65], lt);
    }
    out[(long long)row * ldo + tid64] = __float2bfloat16(o / lt);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks a static list of (image, head, 256-query block) items.  Per item the
// non-persistent kernel spends ~18% of its ~30 us outside the steady-state KV loop (CTA launch, barrier / TMEM set-up,
// cold Q/K loads, first tiles without overlap, output pass with idle tensor cores).  Here TMEM, barriers and the K/V
// ring live across items: the producer streams the next item's Q (double-buffered) and K/V tiles while the current item
// finishes, the MMA warp issues the next item's first two QK^T as soon as the S buffers are free, and the softmax
// warpgroups' output pass overlaps them.  All barrier parities run on a per-CTA global tile counter g = item * T + j.
// Items whose block holds <= tail_rows_max valid rows are served first (CUDA-core path), before the pipeline exists.
// ---------------------------------------------------------------------------------------------------------
struct AttnPersistCfg {
  static constexpr int TQ = 128, TKV = 128, HD = 64, KV_STAGES = 5;
  static constexpr int Q_OFF = 0;                               // 2 buffers x 2 tiles x 16 KB
  static constexpr int K_OFF = 4 * 16384;
  static constexpr int V_OFF = K_OFF + KV_STAGES * 16384;
  static constexpr int BAR_OFF = V_OFF + KV_STAGES * 16384;
  static constexpr int NUM_BARS = 4 + 3 * KV_STAGES + 8;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 8 * 66 * 4;  // barriers (< 256 B) + background tail-row merge buffer
  static constexpr int THREADS = 384;
  static constexpr int TMEM_COLS = 512;
};
static_assert(AttnPersistCfg::NUM_BARS * 8 + 16 <= 256 && AttnPersistCfg::SMEM_BYTES <= 232448, "persistent attention smem budget");

template <int POLY = 0, int MODE = 0>
__global__ void __launch_bounds__(AttnPersistCfg::THREADS, 1)
attn_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  using Cfg = AttnPersistCfg;
  constexpr bool PT = true;
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
  const int T = (nk_eff + Cfg::TKV - 1) / Cfg::TKV;
  const int last_valid = nk_eff - (T - 1) * Cfg::TKV;
  const int last_cols16 = (last_valid + 15) & ~15;
  // regular items r = first_r, first_r + G_CTAS, ...: the CTAs are walked backwards so that the ones that serve an extra
  // tail row (low ids) are not the ones that get the remainder of the regular items
  const int first_r = G_CTAS - 1 - cta;
  const int n_items = (n_reg - first_r + G_CTAS - 1) / G_CTAS;
  const int G = n_items * T;                      // KV tiles this CTA processes per query tile

  // ---- tail rows first (whole CTA, plain loads, shared memory not yet in use) ----
  for (int t = cta; t < n_tail && !p.tail_overlap; t += G_CTAS) {
    const int h = t % p.H, b = t / p.H;
    attn_tail_rows(p.q + (long long)b * p.q_bs + p.q_col0 + h * 64, p.k + (long long)b * p.k_bs + p.k_col0 + h * 64,
                   p.v + (long long)b * p.v_bs + p.v_col0 + h * 64, p.out + (long long)b * p.out_batch_stride + h * 64, p.ldq, p.ldk,
                   p.ldv, p.ldo, p.Nk, p.scale_log2, p.q_start + nqb_reg * 2 * Cfg::TQ, rows_last, smem);
  }
  __syncthreads();

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
      mbar_init(&s_free[i], 4);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_done[i], 1);
    }
    for (int s = 0; s < Cfg::KV_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
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
    } else if (warp >= 10) {
      // ------------------------------- tail rows in the background ----------------
      if (p.tail_overlap) {  // (the host enables it only for Nk >= 32: every key group sees a valid key in its first trip)
        float* s_merge = reinterpret_cast<float*>(smem + Cfg::BAR_OFF + 256);
        for (int t = cta; t < n_tail; t += G_CTAS) {
          const int h = t % p.H, b = t / p.H;
          attn_tail_rows_bg(p.q + (long long)b * p.q_bs + p.q_col0 + h * 64, p.k + (long long)b * p.k_bs + p.k_col0 + h * 64,
                            p.v + (long long)b * p.v_bs + p.v_col0 + h * 64, p.out + (long long)b * p.out_batch_stride + h * 64,
                            p.ldq, p.ldk, p.ldv, p.ldo, p.Nk, p.scale_log2, p.q_start + nqb_reg * 2 * Cfg::TQ, rows_last,
                            s_merge, threadIdx.x - 320);
        }
      }
    } else if (warp == 8) {
      // ------------------------------- MMA issuer ----------------------------------
      const uint64_t dQ = make_smem_desc_sw128(smem_u32(smem + Cfg::Q_OFF), 16, 1024);
      const uint64_t dK = make_smem_desc_sw128(smem_u32(smem + Cfg::K_OFF), 16, 1024);
      const uint64_t dV = make_smem_desc_sw128(smem_u32(smem + Cfg::V_OFF), 16, 1024);
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t idesc_qk_last = make_idesc_bf16(128, last_cols16);
      // QK^T of query tile i against global KV tile g (item n = g / T, tile j = g % T of that item)
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
        // Both PVs first, then both QK^T of tile g + 2 (not needed for another ~2 steps): a warpgroup that is late with
        // s_free -- its output pass at an item boundary -- must not hold up the other warpgroup's PV.  (Keeping the
        // non-persistent order PV0 QK0 PV1 QK1 in steady state and reordering only the boundary tile measured slower:
        // 0.318 vs 0.312 ms.)
        const int g2 = g + 2;
        const bool more = g2 < G;
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_ready[i], g & 1);
          tc_fence_after();
          DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && g >= T - 2 && g < T - 2 + 31) p.dbg[512 + (i * 32 + g - (T - 2)) * 2] = clock64();)
          issue_pv(i, s, j == T - 1, j > 0 ? 1u : 0u, i == 1);
        }
        if (more) {
          mbar_wait(&k_full[g2 % Cfg::KV_STAGES], (g2 / Cfg::KV_STAGES) & 1);
          if (j2 == 0) mbar_wait(&q_full[n2 & 1], (n2 >> 1) & 1);
          for (int i = 0; i < 2; ++i) {
            mbar_wait(&s_free[i], (g + 1) & 1);
            tc_fence_after();
            DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && g2 >= T - 2 && g2 < T - 2 + 31) p.dbg[512 + (i * 32 + g2 - (T - 2)) * 2 + 1] = clock64();)
            issue_qk(i, g2, n2, j2);
          }
        }
        if (++j == T) j = 0;
        if (++j2 == T) { j2 = 0; ++n2; }
      }
    }
  } else {
    // ------------------------------- softmax warpgroups --------------------------
    setmaxnreg_inc<208>();
    const int i = warp >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + i * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + i * 64 + lane_off;
    const uint32_t tP = tmem_base + 384 + i * 64 + lane_off;
    const float sc = p.scale_log2;
    if (i == 1 && p.token_mode) named_bar_arrive(1, 256);  // WG0 owns the MUFU token first
    int g = 0;
    for (int n = 0; n < n_items; ++n) {
      const int ritem = first_r + n * G_CTAS;
      const int qb = ritem % nqb_reg, h = (ritem / nqb_reg) % p.H, b = ritem / (nqb_reg * p.H);
      const int q_row0 = p.q_start + qb * 2 * Cfg::TQ;
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < T; ++j, ++g) {
        long long* dbg = nullptr;
        DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && (warp & 3) == 0 && lane == 0 && g >= T - 2 && g < T - 2 + 31) dbg = p.dbg + (i * 32 + g - (T - 2)) * 8;)
        DCLIP_TL(if (dbg) dbg[6] = clock64();)
        mbar_wait(&s_full[i], g & 1);
        tc_fence_after();
        DCLIP_TL(if (dbg) dbg[0] = clock64();)
        const int valid = (j + 1 == T) ? last_valid : 128;
        const bool last_tile = g + 1 == G;
        if (koff && lane == 0 && (j + 1 == T || (p.peel_key0 == 2 && (j == 0 || j + 2 == T)))) {
          // k_0 / v_0 of this head are touched by nobody else: pull them into L2 at the start of the item (HBM miss,
          // ~1.5 us) and into L1 over the last two tiles, so the output pass finds them there
          const __nv_bfloat16* k0p = p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD;
          const __nv_bfloat16* v0p = p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD;
          if (j == 0) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(k0p));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(v0p));
          } else {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(k0p));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(v0p));
          }
        }
        const uint32_t odp = (g - 1) & 1;
#ifdef DCLIP_EXPERIMENTS
        if ((MODE & 1) != 0 && j > 0) {
          if (valid > 32)
            attn_softmax_tile_spec<128, PT, POLY, (MODE & 2) != 0>(tS, tO, tP, nullptr, r, lane, valid, sc, m_used, l, &s_free[i], &o_done[i], odp, i, last_tile, p.token_mode, nullptr);
          else
            attn_softmax_tile_spec<32, PT, POLY, (MODE & 2) != 0>(tS, tO, tP, nullptr, r, lane, valid, sc, m_used, l, &s_free[i], &o_done[i], odp, i, last_tile, p.token_mode, nullptr);
        } else
#endif
        if (valid > 32)
          attn_softmax_tile<128, PT, POLY, false>(tS, tO, tP, nullptr, r, lane, valid, j == 0, sc, m_used, l, &s_free[i], &o_done[i], odp, i, last_tile, p.token_mode, dbg);
        else
          attn_softmax_tile<32, PT, POLY, false>(tS, tO, tP, nullptr, r, lane, valid, j == 0, sc, m_used, l, &s_free[i], &o_done[i], odp, i, last_tile, p.token_mode, dbg);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[i]);
        DCLIP_TL(if (dbg) dbg[5] = clock64();)
        if (j == 0 && n > 0 && q == 0 && lane == 0) {
          // the previous item's output tile sits in its Q buffer until the TMA store has read it (long done by now);
          // only then may the producer refill that buffer (it needs it a whole item from now)
          tma_store_wait_read();
          mbar_arrive(&q_empty[(n - 1) & 1]);
        }
      }
      // ---- output pass of this item (the MMA warp is already on the next item's first tiles) ----
      long long* edbg = nullptr;
      DCLIP_TL(if (p.dbg && blockIdx.x == p.dbg_cta && (warp & 3) == 0 && lane == 0 && n == 1) edbg = p.dbg + (i * 32 + 31) * 8;)
      DCLIP_TL(if (edbg) edbg[6] = clock64();)
      float z0 = 0.f;
      uint4 v0r[8];
      if (koff) {
        const uint4* k0 = reinterpret_cast<const uint4*>(p.k + (long long)b * p.k_bs + p.k_col0 + h * Cfg::HD);
        const uint4* v0 = reinterpret_cast<const uint4*>(p.v + (long long)b * p.v_bs + p.v_col0 + h * Cfg::HD);
        const uint8_t* qrow = smem + Cfg::Q_OFF + (n & 1) * 32768 + i * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
          const uint4 kv = __ldg(k0 + c);
          v0r[c] = __ldg(v0 + c);
          const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w}, kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            s0 = fmaf(__uint_as_float(qq[e] << 16), __uint_as_float(kk[e] << 16), s0);
            s1 = fmaf(__uint_as_float(qq[e] & 0xffff0000u), __uint_as_float(kk[e] & 0xffff0000u), s1);
          }
        }
        z0 = (s0 + s1) * sc;
      }
      DCLIP_TL(if (edbg) edbg[0] = clock64();)
      mbar_wait(&o_done[i], (g - 1) & 1);
      tc_fence_after();
      DCLIP_TL(if (edbg) edbg[1] = clock64();)
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
      attn_store_tile(tO, q_tile, r, alpha, p0, koff != 0, v0r);
      DCLIP_TL(if (edbg) edbg[2] = clock64();)
      named_bar_sync(3 + i, 128);
      DCLIP_TL(if (edbg) edbg[3] = clock64();)
      if (q == 0 && lane == 0) {
        tma_store_3d(&tmO, q_tile, h * Cfg::HD, q_row0 + i * Cfg::TQ, b);
        tma_store_commit();
      }
      DCLIP_TL(if (edbg) edbg[5] = clock64();)
      tc_fence_before();  // the O reads are ordered before this warp's next p_ready arrive (next item's first PV overwrites O)
    }
    if (q == 0 && lane == 0) tma_store_wait_read();  // last item's tile: shared memory must outlive the store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Few-query attention on CUDA cores (fp32 math): one CTA per (batch, head, query).  Used for the CLS query of the
// ViT (so the tensor-core kernel sees 8 aligned 256-row blocks instead of 8 + a 1-row block), for the
// ContextDecoder's 19-query self/cross attention (models.py:328-344) and for the causal 22-token text tower
// (models.py:836-842).  Phase 1: thread-per-key dot products -> smem scores; block softmax; phase 2: column-parallel PV.
// ---------------------------------------------------------------------------------------------------------
struct SmallAttnParams {
  const void* q; const void* k; const void* v;  // bf16 (is_f32 = 0) or fp32 (is_f32 = 1); [B][N][ld] token-major
  int is_f32;
  int B, H, Nk;
  int q_first, q_count;        // query rows [q_first, q_first + q_count) are computed
  int ldq, ldk, ldv;           // elements per token row
  long long q_bs, k_bs, v_bs;  // batch strides (elements)
  int q_col0, k_col0, v_col0;
  float scale;                 // head_dim^-0.5 (applied to the logits)
  int causal;                  // 1: key j allowed iff j <= query index
  void* out; int out_f32; int ldo; long long out_bs;  // out[b][row][h*64 + d]
  int out_split_off;           // > 0 (bf16 out only): also write lo = bf16(v - hi) at this column offset
  int key_splits;              // > 1: the keys are split over this many CTAs per (batch, head, query block); partial results
  float* ws;                   //      (unnormalised O[64], running max, row sum per query) go to ws, attn_small_combine_kernel merges them
};

// one output element: fp32, or bf16 (+ optional lo half)
__device__ __forceinline__ void attn_small_store(const SmallAttnParams& p, long long ooff, float o) {
  if (p.out_f32) {
    reinterpret_cast<float*>(p.out)[ooff] = o;
  } else {
    const __nv_bfloat16 hi = __float2bfloat16(o);
    reinterpret_cast<__nv_bfloat16*>(p.out)[ooff] = hi;
    if (p.out_split_off > 0)
      reinterpret_cast<__nv_bfloat16*>(p.out)[ooff + p.out_split_off] = __float2bfloat16(o - __bfloat162float(hi));
  }
}

__device__ __forceinline__ float ld_elem(const void* base, int is_f32, long long idx) {
  return is_f32 ? reinterpret_cast<const float*>(base)[idx] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
}

// dynamic smem: QB*Nk floats (scores) + QB*64 (q) + 8*QB*64 (partial O) + 16*QB (reduction scratch).
// QB = queries handled per CTA (they share every K/V load); a CTA covers queries [q0, q0 + QB) of one (batch, head).
// DPT = output dims per thread in phase 2 (8, or 4 for the wide-QB variant whose accumulators would not fit otherwise).
template <int QB, int DPT = (QB > 8 ? 4 : 8)>
__global__ void __launch_bounds__(256) attn_small_kernel(const SmallAttnParams p) {
  extern __shared__ float sm[];
  const int S = p.key_splits > 1 ? p.key_splits : 1;
  const int span = (p.Nk + S - 1) / S;  // keys handled by one CTA
  float* sc = sm;                       // [QB][span]
  float* sq = sm + ((QB * span + 3) & ~3);  // [QB][64], 16-byte aligned (read as float4)
  float* so = sq + QB * 64;             // [8 warps][QB][64]
  float* red = so + 8 * QB * 64;        // [2][8][QB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nqb = (p.q_count + QB - 1) / QB;
  const int ks = blockIdx.x % S, bid = blockIdx.x / S;
  const int q0 = p.q_first + (bid % nqb) * QB;
  const int h = (bid / nqb) % p.H;
  const int b = bid / (nqb * p.H);
  const int nq = min(QB, p.q_first + p.q_count - q0);
  // causal: query qi may see keys j <= qi; the CTA scans up to its last query's limit and masks per query
  const int nk_all = p.causal ? min(p.Nk, q0 + nq) : p.Nk;
  // key range of this CTA (the whole range unless the launch splits the keys); scores are stored relative to k0
  const int k0 = ks * span, nk = min(nk_all, k0 + span);
  for (int i = tid; i < QB * 64; i += 256) {
    const int qq = i >> 6, d = i & 63;
    sq[i] = qq < nq ? ld_elem(p.q, p.is_f32, b * p.q_bs + (long long)(q0 + qq) * p.ldq + p.q_col0 + h * 64 + d) * p.scale : 0.f;
  }
  __syncthreads();
  // phase 1: scores, one key per thread per iteration, all QB queries against the same K row
  float mx[QB];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) mx[qq] = -INFINITY;
  for (int j = k0 + tid; j < nk; j += 256) {
    const long long koff = b * p.k_bs + (long long)j * p.ldk + p.k_col0 + h * 64;
    float kf[64];
    if (p.is_f32) {
      const float4* kp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.k) + koff);
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 kv = kp[d];
        kf[4 * d] = kv.x; kf[4 * d + 1] = kv.y; kf[4 * d + 2] = kv.z; kf[4 * d + 3] = kv.w;
      }
    } else {
      const uint4* kp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.k) + koff);
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const uint4 kv = kp[d];
        const uint32_t w[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          kf[8 * d + 2 * e] = __uint_as_float(w[e] << 16);
          kf[8 * d + 2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int qq = 0; qq < QB; ++qq) {
      float s0 = 0.f, s1 = 0.f;
      const float4* q4 = reinterpret_cast<const float4*>(sq + qq * 64);  // broadcast 128-bit smem reads: 16 per query instead of 64
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 qv = q4[d];
        s0 += qv.x * kf[4 * d];
        s1 += qv.y * kf[4 * d + 1];
        s0 += qv.z * kf[4 * d + 2];
        s1 += qv.w * kf[4 * d + 3];
      }
      float s = s0 + s1;
      if (p.causal && j > q0 + qq) s = -INFINITY;
      sc[qq * span + (j - k0)] = s;
      mx[qq] = fmaxf(mx[qq], s);
    }
  }
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[qq] = fmaxf(mx[qq], __shfl_xor_sync(0xffffffffu, mx[qq], o));
    if (lane == 0) red[warp * QB + qq] = mx[qq];
  }
  __syncthreads();
  float sum[QB];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
    float m = red[qq];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w * QB + qq]);
    mx[qq] = m;
    sum[qq] = 0.f;
  }
  for (int j = k0 + tid; j < nk; j += 256) {
#pragma unroll
    for (int qq = 0; qq < QB; ++qq) {
      const float e = __expf(sc[qq * span + (j - k0)] - mx[qq]);
      sc[qq * span + (j - k0)] = e;
      sum[qq] += e;
    }
  }
#pragma unroll
  for (int qq = 0; qq < QB; ++qq) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum[qq] += __shfl_xor_sync(0xffffffffu, sum[qq], o);
    if (lane == 0) red[8 * QB + warp * QB + qq] = sum[qq];
  }
  __syncthreads();
  // phase 2: O[q][d] = sum_j p[q][j] V[j][d]; thread (g, dq) covers keys j = g, g + KG, ... and DPT consecutive dims
  constexpr int TPR = 64 / DPT, KG = 256 / TPR;  // threads per V row, key groups
  const int dq = tid % TPR, g = tid / TPR;
  float acc[QB][DPT];
#pragma unroll
  for (int qq = 0; qq < QB; ++qq)
#pragma unroll
    for (int e = 0; e < DPT; ++e) acc[qq][e] = 0.f;
  const long long vbase = b * p.v_bs + p.v_col0 + h * 64 + dq * DPT;
  for (int j0 = k0 + g; j0 < nk; j0 += 2 * KG) {
    float vf[2][DPT];
    int jj[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = j0 + KG * u;
      jj[u] = j;
      const long long voff = vbase + (long long)(j < nk ? j : j0) * p.ldv;
      if (p.is_f32) {
        const float4* vp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.v) + voff);
#pragma unroll
        for (int e4 = 0; e4 < DPT / 4; ++e4) {
          const float4 a = vp[e4];
          vf[u][4 * e4] = a.x; vf[u][4 * e4 + 1] = a.y; vf[u][4 * e4 + 2] = a.z; vf[u][4 * e4 + 3] = a.w;
        }
      } else {
        uint32_t w[DPT / 2];
        if constexpr (DPT == 8) {
          const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.v) + voff);
          w[0] = raw.x; w[1] = raw.y; w[2] = raw.z; w[3] = raw.w;
        } else {
          const uint2 raw = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.v) + voff);
          w[0] = raw.x; w[1] = raw.y;
        }
#pragma unroll
        for (int e = 0; e < DPT / 2; ++e) {
          vf[u][2 * e] = __uint_as_float(w[e] << 16);
          vf[u][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (jj[u] < nk) {
#pragma unroll
        for (int qq = 0; qq < QB; ++qq) {
          const float pj = sc[qq * span + (jj[u] - k0)];
#pragma unroll
          for (int e = 0; e < DPT; ++e) acc[qq][e] += pj * vf[u][e];
        }
      }
    }
  }
#pragma unroll
  for (int qq = 0; qq < QB; ++qq)
#pragma unroll
    for (int e = 0; e < DPT; ++e) {
      if constexpr (TPR == 8) acc[qq][e] += __shfl_xor_sync(0xffffffffu, acc[qq][e], 8);
      acc[qq][e] += __shfl_xor_sync(0xffffffffu, acc[qq][e], 16);
    }
  if (lane < TPR) {
#pragma unroll
    for (int qq = 0; qq < QB; ++qq)
#pragma unroll
      for (int e = 0; e < DPT; ++e) so[(warp * QB + qq) * 64 + lane * DPT + e] = acc[qq][e];
  }
  __syncthreads();
  for (int i = tid; i < nq * 64; i += 256) {
    const int qq = i >> 6, d = i & 63;
    float o = 0.f, ssum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      o += so[(w * QB + qq) * 64 + d];
      ssum += red[8 * QB + w * QB + qq];
    }
    if (S > 1) {  // partial result of this key range: [O (unnormalised) x 64 | max | sum] per query
      float* w = p.ws + ((size_t(bid) * S + ks) * QB + qq) * 66;
      w[d] = o;
      if (d == 0) {
        float m = red[qq];
#pragma unroll
        for (int ww = 1; ww < 8; ++ww) m = fmaxf(m, red[ww * QB + qq]);
        w[64] = m;
        w[65] = ssum;
      }
    } else {
      attn_small_store(p, b * p.out_bs + (long long)(q0 + qq) * p.ldo + h * 64 + d, o / ssum);
    }
  }
}

// merges the key-range partials of attn_small_kernel: one 64-thread CTA per (batch, head, query block, query), thread = dim
template <int QB>
__global__ void __launch_bounds__(64) attn_small_combine_kernel(const SmallAttnParams p) {
  const int nqb = (p.q_count + QB - 1) / QB, S = p.key_splits;
  const int bid = blockIdx.x / QB, qq = blockIdx.x % QB, d = threadIdx.x;
  const int q0 = p.q_first + (bid % nqb) * QB;
  const int h = (bid / nqb) % p.H;
  const int b = bid / (nqb * p.H);
  if (q0 + qq >= p.q_first + p.q_count) return;
  const float* w = p.ws + (size_t(bid) * S * QB + qq) * 66;
  float ms[8], ls[8], os[8];  // S <= 8: all partials are fetched before any of them is used
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const bool on = s < S;
    const float* wsp = w + size_t(on ? s : 0) * QB * 66;
    ms[s] = on ? wsp[64] : -INFINITY;
    ls[s] = on ? wsp[65] : 0.f;
    os[s] = on ? wsp[d] : 0.f;
  }
  float m = -INFINITY;
#pragma unroll
  for (int s = 0; s < 8; ++s) m = fmaxf(m, ls[s] > 0.f ? ms[s] : -INFINITY);
  float num = 0.f, den = 0.f;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const float f = ls[s] > 0.f ? __expf(ms[s] - m) : 0.f;  // an empty key range contributes nothing
    num += f * os[s];
    den += f * ls[s];
  }
  attn_small_store(p, b * p.out_bs + (long long)(q0 + qq) * p.ldo + h * 64 + d, num / den);
}

}  // namespace dclip
