// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * operands: TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring, K-major for both A and W
//     (W is a torch `nn.Linear.weight` [out,in], i.e. already K-major: no transposes anywhere)
//   * math: tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), 128 x BN x 16 per instruction, accumulators in TMEM,
//     double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   * roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = epilogue
//   * epilogue: tcgen05.ld -> per-warp smem transpose -> coalesced global access with
//     bias / activation / fp32 residual / row remap fused; writes fp32 and/or bf16 (optionally bf16 hi+lo split)
//
// Replaces the cuBLAS/MKL calls behind nn.Linear / nn.MultiheadAttention in_proj,out_proj / 1x1 Conv2d of the
// reference (segmentation/denseclip/models.py:275-281,287-289; denseclip.py:198,616).
#pragma once
#include "ptx.cuh"

#ifdef DCLIP_GEMM_TIMELINE
#define DCLIP_GTL(...) __VA_ARGS__
#else
#define DCLIP_GTL(...)
#endif

namespace dclip {

enum GemmAct : int {
  ACT_NONE = 0,
  ACT_QUICKGELU = 1,          // x * sigmoid(1.702 x)  (models.py:252-254), tanh.approx form (bf16-accurate)
  ACT_QUICKGELU_PRECISE = 2,  // same, exp/rcp form (fp32-accurate; used by the split-bf16 "fp32" path)
  ACT_GELU_ERF = 3,           // nn.GELU() exact (models.py:363)
  ACT_RELU = 4,
};

struct GemmParams {
  int M, N, K;            // logical problem (K per segment when split_in)
  int split_in;           // 1: A=[Ah|Al] (2K cols), W=[Wh|Wl] (2K cols); computes Ah*Wh + Al*Wh + Ah*Wl
  const float* bias;      // [N] or null
  int act;                // GemmAct
  const float* residual;  // fp32 or null; row = res_mod ? 1 + m % remap_P : out_row
  int ldr;
  int res_mod;
  int remap_P, remap_Nt;  // remap_P > 0: out_row = (m / P) * Nt + 1 + m % P   (patch-embed -> token rows)
  float* out_f32;         // [*, ldc] or null
  int ldc;
  __nv_bfloat16* out_bf16;  // [*, ldcb] or null
  int ldcb;
  int split_out;          // 1: also write lo = bf16(v - hi) at column offset split_out_off
  int split_out_off;
  float out_scale;        // applied after activation, before residual (1.0f default)
  // implicit 3x3 / pad-1 convolution (conv_C > 0): A is a token-major image [B][gh][gw][C(x2 if split_in)] read through a
  // 4D tensor map; K = 9*C ordered (ky, kx, c); M = B*gh*gw with 128-pixel tiles that never straddle an image.
  int conv_C, conv_gw, conv_tiles_per_img;
  int conv_G;    // > 1: grouped conv -- n-block g reads activation group g (5-D map), N = G * BLOCK_N (the neck's 12 taps in one launch)
  int cluster;   // informational: 2 when launched as CTA pairs (PAIR template), else 1
  // 3x3-conv weight-gradient mode (wg_C > 0; train_tail.py): C[m, t*wg_C + c] = sum_k A[m, k] * Wcopy[t%3][c, k + (t/3)*wg_pitch],
  // t = 0..8 = (ky, kx).  A = dY^T [filters][padded pixels]; W = three copies of X^T, [3][wg_rows][wg_pitch + padded pixels + ...]:
  // copy kx is X^T shifted by kx - 1 pixels behind wg_pitch leading zeros (transpose_pad_kernel), the pixel index runs over
  // zero-padded images with row pitch wg_pitch (% 8 == 0).  The horizontal tap offset is baked into the copies because a TMA box
  // must start 16-byte aligned in the innermost dimension; the vertical one is a K-coordinate shift by a multiple of wg_pitch.
  // N = 9 * wg_C, wg_C % BLOCK_N == 0.  wg_grouped: m-block g (the 128 filters of group g) pairs with rows g*wg_C .. of each copy.
  int wg_C, wg_pitch, wg_grouped, wg_rows;
  // LayerNorm folding (bf16 ViT blocks, vit_encoder.cuh): the residual GEMMs (out-proj / c_proj) publish per-row partial
  // (sum, sum of squares) of the UPDATED residual stream, one float2 slot per (n-block, epilogue warp half), SLOT-MAJOR
  // (row_stats_out[slot * stats_ld + row], stats_ld >= M: the consumer's lanes = consecutive rows read 256 contiguous bytes
  // per slot).  Every slot has exactly one writer and the consumer adds the slots in order, so the statistics are
  // deterministic.  The GEMM that follows (QKV / c_fc) takes the raw bf16 stream as its A operand and normalises in its epilogue:
  //   LN(x) W^T + b = rstd * (x (gamma*W)^T) - rstd * mean * rowsum(gamma*W) + (b + W beta)
  // with W := gamma*W and bias := b + W beta folded on the host, ln_c[n] = rowsum of the (bf16-rounded) folded weight.
  float2* row_stats_out;        // producer side (ld/st epilogues), or null
  const float2* row_stats_in;   // consumer side (TMA-store epilogue), or null
  int stats_ld, stats_n;        // slot stride in rows (>= M) / slots to sum on the consumer side (<= 12)
  const float* ln_c;            // [N]
  float ln_inv_d, ln_eps;       // 1 / (LayerNorm width), eps
  long long* dbg;  // selftest only (-DDCLIP_GEMM_TIMELINE): clock64 stamps of CTA 0's epilogue warp 4 and MMA warp
  int dbg_mode;  // selftest only: 1 = epilogue drains TMEM but skips staging and stores; 2 = stage but skip global stores
};

// Compile-time epilogue specialisation. ACT < 0 / FLAGS < 0 select the generic (runtime-checked) epilogue.
// EPI_CONV_SWAP: grouped implicit 3x3 conv with the operand roles swapped -- the 128 filters of a group are the MMA's M rows
// (A operand = weights), 256 PIXELS are its N columns (B operand = the TMA-gathered activations), so one tcgen05.mma is
// 128 x 256 x 16 instead of 128 x 128 x 16: the per-instruction cost of the tensor pipe is the same for N = 128 and 256
// (profiles/r02_mma_probe.txt), and the 128-filter tile ran at 47% of peak.  The accumulator is then [filter][pixel]; the epilogue
// writes it transposed (2-byte stores, 64 contiguous bytes per warp instruction) -- negligible next to the K = 6912 mainloop.
enum GemmEpiFlags : int { EPI_RESID = 1, EPI_OUT_F32 = 2, EPI_OUT_BF16 = 4, EPI_SPLIT = 8, EPI_REMAP = 16, EPI_TMA_STORE = 32, EPI_CONV_SWAP = 64 };

// PAIR: two CTAs of a 2x1x1 cluster form one tcgen05 cta_group::2 unit computing a 256 x BN tile; each CTA stages its
// own 128 A rows and only HALF of the W tile (BN/2 rows), which cuts the per-SM operand ingest from 48 KB to 32 KB per
// k-block -- the 1-CTA mainloop is bound by exactly that ingest (profiles/r01_gemm_notes.md).
template <int BN, bool PAIR = false>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 196608 / STAGE_BYTES;
  static constexpr int EPI_WARPS = 8;
  static constexpr int STG_WARP_BYTES = 32 * 32 * 4;  // one 32x32 fp32 chunk per epilogue warp, XOR-swizzled (no padding)
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES + STG_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256;
  static constexpr int TMEM_COLS = BN == 192 ? 512 : 2 * BN;  // tcgen05.alloc wants a power of two; the two accumulators sit at 0 and BN
  static constexpr int THREADS = 128 + EPI_WARPS * 32;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_QUICKGELU: return x * (0.5f + 0.5f * tanh_approx(0.851f * x));
    case ACT_QUICKGELU_PRECISE: return x / (1.0f + __expf(-1.702f * x));
    case ACT_GELU_ERF: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}

template <int ACT>
__device__ __forceinline__ float apply_act_t(float x, int act_rt) {
  if constexpr (ACT < 0) return apply_act(x, act_rt);
  else if constexpr (ACT == ACT_QUICKGELU) return x * (0.5f + 0.5f * tanh_approx(0.851f * x));
  else if constexpr (ACT == ACT_NONE) return x;
  else return apply_act(x, ACT);
}

template <int BN, int ACT, int FLAGS, bool PAIR = false>
__global__ void __launch_bounds__(GemmCfg<BN, PAIR>::THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool SWAP = FLAGS >= 0 && (FLAGS & EPI_CONV_SWAP) != 0;   // (BN = 256 pixels per tile, BM = 128 filters per group)
  const int num_m = SWAP ? p.M / BN : (p.M + BM - 1) / BM, num_n = SWAP ? p.conv_G : (p.N + BN - 1) / BN;
  // pair mode: the two CTAs of a pair own m-blocks (2*mp, 2*mp + 1) of the same n-block
  constexpr int cl = PAIR ? 2 : 1;
  const int crank = PAIR ? int(cluster_ctarank()) : 0;
  const int num_units = cl == 2 ? ((num_m + 1) / 2) * num_n : num_m * num_n;  // work items per CTA stream
  const int unit0 = cl == 2 ? blockIdx.x / 2 : blockIdx.x;
  const int unit_step = cl == 2 ? gridDim.x / 2 : gridDim.x;
  const int num_tiles = num_units;
  const int kseg = (p.K + BK - 1) / BK;
  const int num_k = p.split_in ? 3 * kseg : kseg;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("dclip gemm: dynamic smem base not 1024B aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], cl);    // pair mode: leader's producer (with the tx bytes of both CTAs) + the peer's producer
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], Cfg::EPI_WARPS * cl);  // pair mode: the leader's MMA waits for both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) { tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // peer barriers / TMEM must exist before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int m_blk = (tile / num_n) * cl + crank, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (PAIR) {
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            else mbar_arrive_remote(&full_bar[stage], 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          }
          int a_col = kb * BK, b_col = kb * BK;
          if (p.split_in) {
            const int seg = kb / kseg, off = (kb - seg * kseg) * BK;
            a_col = (seg == 1 ? p.K : 0) + off;
            b_col = (seg == 2 ? p.K : 0) + off;
          }
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if constexpr (SWAP) {
            // tile = (256-pixel block m_blk, group n_blk): weights of the group -> A slot (16 KB), pixels -> B slot (32 KB)
            const int cpb = p.conv_C / BK;
            const int tap = kb / cpb, c0 = (kb - tap * cpb) * BK;
            const int img = m_blk / p.conv_tiles_per_img, p0 = (m_blk - img * p.conv_tiles_per_img) * BN;
            const int y0 = p0 / p.conv_gw, x0 = p0 - y0 * p.conv_gw;
            tma_load_2d(sa, &tmB, &full_bar[stage], kb * BK, n_blk * BM);
            tma_load_5d(sa + Cfg::A_BYTES, &tmA, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img, n_blk);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (p.conv_C > 0) {
            const int kc = kb % kseg;                     // k-block inside the (ky, kx, c) K range
            const int cpb = p.conv_C / BK;                // k-blocks per filter tap
            const int tap = kc / cpb, c0 = (kc - tap * cpb) * BK + (a_col >= p.K ? p.conv_C : 0);
            const int img = m_blk / p.conv_tiles_per_img, p0 = (m_blk - img * p.conv_tiles_per_img) * BM;
            const int y0 = p0 / p.conv_gw, x0 = p0 - y0 * p.conv_gw;
            if (p.conv_G > 1) tma_load_5d(sa, &tmA, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img, n_blk);
            else tma_load_4d(sa, &tmA, &full_bar[stage], c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, img);
          } else if constexpr (PAIR) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], a_col, m_blk * BM);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], a_col, m_blk * BM);
          }
          if constexpr (PAIR) {
            tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], b_col, n_blk * BN + crank * (BN / 2));
          } else {
            int b_row = n_blk * BN;
            if (p.wg_C > 0) {
              const int cpb = p.wg_C / BN, t = n_blk / cpb;
              b_row = (t % 3) * p.wg_rows + (p.wg_grouped ? m_blk * p.wg_C : 0) + (n_blk - t * cpb) * BN;
              b_col += (t / 3) * p.wg_pitch;
            }
            tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], b_col, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ------------------------------- MMA issuer (leader CTA only in pair mode) ----
    constexpr uint32_t idesc = make_idesc_bf16(BM * cl, BN);
    int stage = 0;
    uint32_t phase = 0, it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            if constexpr (PAIR) umma_ss_f16_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_ss_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          DCLIP_GTL(if (p.dbg && blockIdx.x == 0 && it < 16 && (kb == 0 || kb == num_k - 1)) p.dbg[256 + it * 2 + (kb ? 1 : 0)] = clock64();)
          if constexpr (PAIR) {  // both CTAs' producers / epilogues are released by the same commit
            umma_commit_2sm_mcast(&empty_bar[stage], uint16_t(3));
            if (kb == num_k - 1) umma_commit_2sm_mcast(&tfull_bar[as], uint16_t(3));
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == num_k - 1) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------
    // 8 warps: warp pair (w, w+4) shares TMEM lane quarter q = w & 3 and splits the tile's 32-column chunks.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    uint8_t* stg = smem + STAGES * Cfg::STAGE_BYTES + (warp - 4) * Cfg::STG_WARP_BYTES;
    const bool has_res = FLAGS >= 0 ? bool(FLAGS & EPI_RESID) : p.residual != nullptr;
    const bool has_f32 = FLAGS >= 0 ? bool(FLAGS & EPI_OUT_F32) : p.out_f32 != nullptr;
    const bool has_b16 = FLAGS >= 0 ? bool(FLAGS & EPI_OUT_BF16) : p.out_bf16 != nullptr;
    const bool has_split = FLAGS >= 0 ? bool(FLAGS & EPI_SPLIT) : p.split_out != 0;
    const bool has_remap = FLAGS >= 0 ? bool(FLAGS & EPI_REMAP) : p.remap_P > 0;
    const int sub = lane >> 3;      // row within a group of 4 rows handled per warp instruction
    const int cq = lane & 7;        // 4-column group within the 32-column chunk
    uint32_t it = 0;
    if constexpr (SWAP) {
      // accumulator [filter = TMEM lane][pixel = column]: this thread owns filter q*32 + lane of the group and writes its 256
      // pixels transposed into out_bf16[pixel][group * 128 + filter] (a warp instruction = 32 consecutive filters = 64 B)
      for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        const int col = n_blk * BM + q * 32 + lane;
        const float bias = p.bias ? __ldg(p.bias + col) : 0.f;
        __nv_bfloat16* obase = p.out_bf16 + size_t(m_blk) * BN * p.ldcb + col;
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = half; chunk < BN / 32; chunk += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + as * BN + chunk * 32 + (uint32_t(q * 32) << 16), r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            obase[size_t(chunk * 32 + e) * p.ldcb] = __float2bfloat16(apply_act_t<ACT>(__uint_as_float(r[e]) + bias, p.act) * p.out_scale);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
    } else
    if constexpr (FLAGS >= 0 && (FLAGS & EPI_TMA_STORE)) {
      // bf16 output through TMA stores: each warp converts its 32 rows x 64 columns (bias + activation applied in the
      // row-per-lane TMEM layout), writes them as one SWIZZLE_128B smem tile and hands it to the TMA engine, which
      // emits full 128 B lines asynchronously (no LSU store traffic, M/N tails clipped by the tensor map).
      for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
        const int m_blk = (tile / num_n) * cl + crank, n_blk = tile % num_n;
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        const int row0 = m_blk * BM + q * 32;
        // folded LayerNorm: this thread's row statistics (summed in slot order: deterministic), before the accumulator wait
        float ln_rstd = 1.f, ln_nm = 0.f;
        if (p.row_stats_in) {
          float s1 = 0.f, s2 = 0.f;
          if (row0 + lane < p.M) {
            const float2* sp = p.row_stats_in + row0 + lane;
            // <= 12 slots (2 per producer n-block).  Unconditional loads (the slot index is clamped, surplus slots get weight 0):
            // a predicated load per slot made the compiler serialise them behind branches (12 exposed DRAM/L2 latencies per tile)
            float2 t[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) t[j] = __ldg(sp + size_t(min(j, p.stats_n - 1)) * p.stats_ld);
#pragma unroll
            for (int j = 0; j < 12; ++j) { const float wj = j < p.stats_n ? 1.f : 0.f; s1 = fmaf(wj, t[j].x, s1); s2 = fmaf(wj, t[j].y, s2); }
          }
          const float mean = s1 * p.ln_inv_d;
          ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_d - mean * mean, 0.f) + p.ln_eps);
          ln_nm = -ln_rstd * mean;
        }
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after();
#pragma unroll 1
        for (int cg = half; cg < BN / 64; cg += 2) {
          const int c0 = n_blk * BN + cg * 64;
          if (lane == 0) tma_store_wait_read();  // the previous store of this warp has finished reading the staging tile
          __syncwarp();
#pragma unroll
          for (int sc = 0; sc < 2; ++sc) {
            // per-column epilogue constants of this 32-column chunk (same address in every lane: one L1 wavefront each), all
            // issued BEFORE the TMEM load so their latencies overlap it and each other (loading them per 8 columns inside the
            // loop below exposed one L1/L2 latency per load: the epilogue, not the MMA, bounded the tile).  Columns >= N are
            // clamped to the last valid group: the TMA store clips them.
            float4 bv[8], cv[8];
            {
              const int cN = p.N >= 8 ? p.N - 8 : 0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = min(c0 + sc * 32 + 8 * j, cN);
                bv[2 * j] = bv[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) {
                  bv[2 * j] = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                  bv[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
                }
                cv[2 * j] = cv[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.row_stats_in) {
                  cv[2 * j] = __ldg(reinterpret_cast<const float4*>(p.ln_c + c));
                  cv[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p.ln_c + c + 4));
                }
              }
            }
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + as * BN + cg * 64 + sc * 32 + (uint32_t(q * 32) << 16), r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
              // v = rstd * acc - rstd * mean * ln_c[c] + bias  (plain GEMM: rstd = 1, the ln_c term is 0)
              const float4 b0 = bv[2 * j], b1 = bv[2 * j + 1], k0 = cv[2 * j], k1 = cv[2 * j + 1];
              v[0] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 0]), fmaf(ln_nm, k0.x, b0.x));
              v[1] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 1]), fmaf(ln_nm, k0.y, b0.y));
              v[2] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 2]), fmaf(ln_nm, k0.z, b0.z));
              v[3] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 3]), fmaf(ln_nm, k0.w, b0.w));
              v[4] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 4]), fmaf(ln_nm, k1.x, b1.x));
              v[5] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 5]), fmaf(ln_nm, k1.y, b1.y));
              v[6] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 6]), fmaf(ln_nm, k1.z, b1.z));
              v[7] = fmaf(ln_rstd, __uint_as_float(r[8 * j + 7]), fmaf(ln_nm, k1.w, b1.w));
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = apply_act_t<ACT>(v[e], p.act) * p.out_scale;
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((sc * 4 + j) ^ (lane & 7)) << 4)) =
                  make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M && c0 < p.N) {
            tma_store_2d(&tmC, stg, c0, p.dbg_mode == 5 ? (row0 & 4095) : row0);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[as], 0);
          else mbar_arrive(&tempty_bar[as]);
        }
      }
      if (lane == 0) tma_store_wait_all();
    } else
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
      const int m_blk = (tile / num_n) * cl + crank, n_blk = tile % num_n;
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int row0 = m_blk * BM + q * 32;
      const int rows_valid = min(32, p.M - row0);  // may be <= 0
      // per-lane output / residual row offsets for its 8 rows (row = row0 + 4*k + sub)
      size_t orow[8], rrow[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int m = row0 + 4 * k + sub;
        orow[k] = m;
        rrow[k] = m;
        if (has_remap) {
          const int bi = m / p.remap_P, pm = m - bi * p.remap_P;
          orow[k] = size_t(bi) * p.remap_Nt + 1 + pm;
          rrow[k] = p.res_mod ? size_t(1 + pm) : orow[k];
        }
      }
      // The residual does not depend on the accumulator: its loads for chunk 0 are issued BEFORE waiting for the MMA,
      // and those of chunk i+1 before chunk i is processed, so the DRAM latency of the fp32 residual stream is hidden.
      auto load_resid = [&](int chunk, float4 (&dst)[8]) {
        const int cc = n_blk * BN + chunk * 32 + 4 * cq;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_res && chunk < BN / 32 && cc < p.N && 4 * k + sub < rows_valid)
            dst[k] = *reinterpret_cast<const float4*>(p.residual + rrow[k] * p.ldr + cc);
        }
      };
      // ... and the NEXT tile's residual rows are pulled into L2 now (one tile of lead time): lane = row, 4 lines per lane
      if (has_res && tile + unit_step < num_tiles) {
        const int nt = tile + unit_step;
        const int m2 = ((nt / num_n) * cl + crank) * BM + q * 32 + lane;
        if (m2 < p.M) {
          size_t r2 = m2;
          if (has_remap) {
            const int bi = m2 / p.remap_P, pm = m2 - bi * p.remap_P;
            r2 = p.res_mod ? size_t(1 + pm) : size_t(bi) * p.remap_Nt + 1 + pm;
          }
          const float* base = p.residual + r2 * p.ldr + (nt % num_n) * BN;
          for (int ch = half; ch < BN / 32; ch += 2)
            if ((nt % num_n) * BN + ch * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ch * 32));
        }
      }
      float4 rr[8];
      load_resid(half, rr);
      float4 bias_t[(BN / 64 + 1) / 2 * 2];
#pragma unroll
      for (int ci = 0; ci < (BN / 64 + 1) / 2 * 2; ++ci) {
        const int cc = n_blk * BN + (half + 2 * ci) * 32 + 4 * cq;
        bias_t[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && half + 2 * ci < BN / 32 && cc < p.N) bias_t[ci] = *reinterpret_cast<const float4*>(p.bias + cc);
      }
      DCLIP_GTL(if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 16) p.dbg[it * 8 + 0] = clock64();)
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      DCLIP_GTL(if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 16) p.dbg[it * 8 + 1] = clock64();)
      // (row_stats_out) this lane's share of the row sums of its 8 rows, accumulated over all chunks of this warp
      float st1[8], st2[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { st1[k] = 0.f; st2[k] = 0.f; }
      auto process = [&](int chunk, const float4 (&rres)[8], const float4 b4) {
        const int c = n_blk * BN + chunk * 32 + 4 * cq;
        const bool col_ok = c < p.N;  // N % 4 == 0 is required by the launcher
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + as * BN + chunk * 32 + (uint32_t(q * 32) << 16), r);
        tmem_wait_ld();
        if (p.dbg_mode == 1) return;
        // row-per-lane -> staging (16B chunk j of row `lane` lands at physical chunk j ^ (lane & 7))
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        __syncwarp();
        if (col_ok && p.dbg_mode != 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = 4 * k + sub;
            if (i < rows_valid) {
              float4 v = *reinterpret_cast<const float4*>(stg + i * 128 + ((cq ^ (i & 7)) << 4));
              v.x = apply_act_t<ACT>(v.x + b4.x, p.act) * p.out_scale + rres[k].x;
              v.y = apply_act_t<ACT>(v.y + b4.y, p.act) * p.out_scale + rres[k].y;
              v.z = apply_act_t<ACT>(v.z + b4.z, p.act) * p.out_scale + rres[k].z;
              v.w = apply_act_t<ACT>(v.w + b4.w, p.act) * p.out_scale + rres[k].w;
              if constexpr (FLAGS >= 0 && (FLAGS & EPI_RESID)) {   // (row_stats_out: only the residual GEMMs publish statistics)
                st1[k] += (v.x + v.y) + (v.z + v.w);
                st2[k] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
              }
              if (has_f32) *reinterpret_cast<float4*>(p.out_f32 + orow[k] * p.ldc + c) = v;
              if (has_b16) {
                const uint32_t h0 = pack_bf16x2(v.x, v.y), h1 = pack_bf16x2(v.z, v.w);
                *reinterpret_cast<uint2*>(p.out_bf16 + orow[k] * p.ldcb + c) = make_uint2(h0, h1);
                if (has_split) {
                  const uint32_t l0 = pack_bf16x2(v.x - __uint_as_float(h0 << 16), v.y - __uint_as_float(h0 & 0xffff0000u));
                  const uint32_t l1 = pack_bf16x2(v.z - __uint_as_float(h1 << 16), v.w - __uint_as_float(h1 & 0xffff0000u));
                  *reinterpret_cast<uint2*>(p.out_bf16 + orow[k] * p.ldcb + p.split_out_off + c) = make_uint2(l0, l1);
                }
              }
            }
          }
        }
        __syncwarp();
      };
      // ping-pong residual buffers (no register copies: a copy would wait for the loads it is meant to hide); the
      // per-column bias of all this warp's chunks was loaded before the accumulator wait as well
#pragma unroll
      for (int ci = 0; ci < (BN / 64 + 1) / 2; ++ci) {
        const int chunk = half + 4 * ci;
        if (chunk < BN / 32) {
          float4 rr2[8];
          load_resid(chunk + 2, rr2);
          process(chunk, rr, bias_t[2 * ci]);
          if (chunk + 2 < BN / 32) {
            load_resid(chunk + 4, rr);
            process(chunk + 2, rr2, bias_t[2 * ci + 1]);
          }
        }
        DCLIP_GTL(if (p.dbg && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 16 && chunk < 8) p.dbg[it * 8 + 2 + chunk / 2] = clock64();)
      }
      tc_fence_before();
      if (lane == 0) {
        if (PAIR && crank != 0) mbar_arrive_remote(&tempty_bar[as], 0);
        else mbar_arrive(&tempty_bar[as]);
      }
      if constexpr (FLAGS >= 0 && (FLAGS & EPI_RESID)) {
        if (p.row_stats_out) {   // warp-uniform: sum over the 8 column groups (lanes that share `sub`); slot = (n-block, warp half)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              st1[k] += __shfl_xor_sync(0xffffffffu, st1[k], o);
              st2[k] += __shfl_xor_sync(0xffffffffu, st2[k], o);
            }
          }
          if (cq == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (4 * k + sub < rows_valid)
                p.row_stats_out[size_t(n_blk * 2 + half) * p.stats_ld + orow[k]] = make_float2(st1[k], st2[k]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dclip
