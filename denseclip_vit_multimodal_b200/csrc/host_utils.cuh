// Host-side helpers: error plumbing, driver entry point for cuTensorMapEncodeTiled (no link-time libcuda
// dependency, so the library loads on a machine without a driver), tensor-map builders, GEMM launcher.
#pragma once
#include <set>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include <map>
#include <utility>
#include <vector>

#ifdef DCLIP_WITH_ATTN_P4   // round-2 experiment (scripts/experiments/, selftest builds only; measured slower: profiles/r02_attention_notes.md)
#include "../../scripts/experiments/attn_p4_tcgen05.cuh"
#endif
#include "attn_split_tcgen05.cuh"
#include "attn_tcgen05.cuh"
#include "gemm_tcgen05.cuh"
#ifdef DCLIP_EXPERIMENTS   // round-2 experiment: column-split softmax (four warpgroups); measured slower (0.424 vs 0.311 ms): profiles/r02_attention_notes.md
#include "../../scripts/experiments/attn_cs_tcgen05.cuh"
#endif

namespace dclip {

struct Error {
  std::string msg;
};

#define DCLIP_CHECK_CUDA(expr)                                                                       \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      char _buf[512];                                                                                \
      snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
               __LINE__);                                                                            \
      throw ::dclip::Error{_buf};                                                                    \
    }                                                                                                \
  } while (0)

#define DCLIP_REQUIRE(cond, ...)                                                  \
  do {                                                                            \
    if (!(cond)) {                                                                \
      char _buf[512];                                                             \
      int _n = snprintf(_buf, sizeof(_buf), "requirement failed: %s: ", #cond);   \
      snprintf(_buf + _n, sizeof(_buf) - _n, __VA_ARGS__);                        \
      throw ::dclip::Error{_buf};                                                 \
    }                                                                             \
  } while (0)

inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  if (!fn) throw Error{"cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)"};
  return fn;
}

// bf16 tensor map, up to 4 dims (innermost first). strides_bytes has rank-1 entries (dims 1..rank-1).
inline CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                  const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  DCLIP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16B aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    DCLIP_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map stride[%d]=%llu not a multiple of 16B", i,
                  (unsigned long long)strides_bytes[i]);
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  cuuint64_t gdim[5] = {1, 1, 1, 1, 1};
  cuuint64_t gstr[4] = {0, 0, 0, 0};
  cuuint32_t bdim[5] = {1, 1, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = get_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", int(r),
             rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1], bdim[0], bdim[1]);
    throw Error{buf};
  }
  return tm;
}

// row-major [rows, cols] bf16 matrix with leading dimension ld (elements); box = {64 cols, box_rows}
inline CUtensorMap make_tmap_2d_bf16(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_bf16(base, 2, dims, str, box);
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs through several handles)
inline int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of the function: opt in once per (kernel, device)
template <class Kern>
inline void ensure_dyn_smem(Kern kern, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  DCLIP_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(reinterpret_cast<const void*>(kern), dev);
  if (done.count(key)) return;
  DCLIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert(key);
}

// A/B knobs of the round-1 / round-2 experiments are read from the environment ONLY in selftest builds
// (-DDCLIP_EXPERIMENTS); the shipped library compiles the measured-best setting in and contains no getenv.
#ifdef DCLIP_EXPERIMENTS
inline int env_knob(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#define DCLIP_KNOB(name, dflt) ([] { static const int v = ::dclip::env_knob(name, dflt); return v; }())
#else
#define DCLIP_KNOB(name, dflt) (dflt)
#endif

struct GemmOperands {
  const __nv_bfloat16* A;
  int lda;  // elements
  const __nv_bfloat16* W;
  int ldw;
  // implicit-conv mode (GemmParams::conv_C > 0): A is [B][row0 + gh*gw][lda] token-major with batch stride a_bs
  long long a_bs = 0;
  int conv_B = 0, conv_gh = 0;
  long long a_gs = 0;  // grouped conv: element stride between activation groups
};

// A GEMM launch with its tensor maps pre-encoded (cuTensorMapEncodeTiled is a driver call; plans are built once per
// (buffer, shape) and replayed, which also makes the forward CUDA-graph capturable without host work).
struct GemmPlan {
  CUtensorMap tmA, tmB, tmC;   // tmC: bf16 output map for the TMA-store epilogue (valid iff tma_store)
  bool tma_store = false;
  const void* tma_store_ptr = nullptr;
  GemmParams p;
  int bn = 0;
  int grid = 0;
  bool conv_swap = false;   // grouped 3x3 conv with swapped operand roles (EPI_CONV_SWAP)
};

template <int BN, int ACT, int FLAGS, bool PAIR = false>
inline void launch_gemm_inst(const GemmPlan& plan, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, ACT, FLAGS, PAIR>;
  ensure_dyn_smem(kern, Cfg::SMEM_BYTES);
  if (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DCLIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, plan.tmA, plan.tmB, plan.tmC, plan.p));
  } else {
    kern<<<plan.grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(plan.tmA, plan.tmB, plan.tmC, plan.p);
  }
  DCLIP_CHECK_CUDA(cudaGetLastError());
}

// DCLIP_GEMM_TRACE=1: print every (BLOCK_N, act, epilogue flags, pair) combination that falls through to the generic
// runtime-checked epilogue, once -- those are ~1.5x slower per tile than a specialised instantiation
inline void trace_generic_gemm(int bn, int act, int flags, int pair, const GemmParams& p) {
  if (!DCLIP_KNOB("DCLIP_GEMM_TRACE", 0)) return;
  static std::set<long long> seen;
  const long long key = ((long long)bn << 32) | (act << 16) | (flags << 4) | pair;
  if (seen.insert(key).second)
    fprintf(stderr, "dclip gemm: generic epilogue bn=%d act=%d flags=%d pair=%d (M=%d N=%d K=%d split_in=%d)\n", bn, act, flags, pair, p.M,
            p.N, p.K, p.split_in);
}

// Fallback for epilogue combinations without a full specialisation: the output flags are checked at run time, but the
// ACTIVATION is still a template parameter -- a run-time activation switch inside the element loops costs 5.7x per tile
// (measured: 125 us vs 22 us for a 32784 x 512 x 256 fp32-output GEMM), run-time flags only 1.5x.
template <int BN, bool PAIR>
inline void launch_gemm_generic(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.p.act) {
    case ACT_NONE: return launch_gemm_inst<BN, ACT_NONE, -1, PAIR>(plan, stream);
    case ACT_QUICKGELU: return launch_gemm_inst<BN, ACT_QUICKGELU, -1, PAIR>(plan, stream);
    case ACT_QUICKGELU_PRECISE: return launch_gemm_inst<BN, ACT_QUICKGELU_PRECISE, -1, PAIR>(plan, stream);
    case ACT_GELU_ERF: return launch_gemm_inst<BN, ACT_GELU_ERF, -1, PAIR>(plan, stream);
    case ACT_RELU: return launch_gemm_inst<BN, ACT_RELU, -1, PAIR>(plan, stream);
    default: throw Error{"unknown activation"};
  }
}

template <int BN>
inline void launch_gemm_bn(const GemmPlan& plan, cudaStream_t stream) {
  const GemmParams& p = plan.p;
  const int flags = (p.residual ? EPI_RESID : 0) | (p.out_f32 ? EPI_OUT_F32 : 0) | (p.out_bf16 ? EPI_OUT_BF16 : 0) |
                    (p.split_out ? EPI_SPLIT : 0) | (p.remap_P > 0 ? EPI_REMAP : 0);
  // BLOCK_N 192 exists only as CTA-pair instantiations of the fp32-residual epilogues (N = 768: 7 full waves of
  // 256x192 pair tiles instead of 5.2 -> 6 waves of 256x256)
  if constexpr (BN == 192) {
    DCLIP_REQUIRE(p.cluster == 2, "BLOCK_N 192 is a CTA-pair configuration");
    if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32))
      return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32, true>(plan, stream);
    if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16))
      return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16, true>(plan, stream);
    trace_generic_gemm(BN, p.act, flags, 1, p);
    return launch_gemm_generic<BN, true>(plan, stream);
  } else {
  if constexpr (BN == 256) {
    if (plan.conv_swap) {
      DCLIP_REQUIRE(p.act == ACT_RELU && flags == EPI_OUT_BF16, "swapped grouped conv: bf16 ReLU epilogue only");
      return launch_gemm_inst<BN, ACT_RELU, EPI_OUT_BF16 | EPI_CONV_SWAP>(plan, stream);
    }
  }
  // CTA-pair (cta_group::2) instantiations exist for the 256-wide tile and the hot ViT-block epilogues
  if constexpr (BN == 256) {
    if (p.cluster == 2) {
      if (plan.tma_store && flags == EPI_OUT_BF16 && p.out_bf16 == plan.tma_store_ptr) {
        if (p.act == ACT_NONE) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_BF16 | EPI_TMA_STORE, true>(plan, stream);
        if (p.act == ACT_QUICKGELU) return launch_gemm_inst<BN, ACT_QUICKGELU, EPI_OUT_BF16 | EPI_TMA_STORE, true>(plan, stream);
      }
      if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32))
        return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32, true>(plan, stream);
      if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16))
        return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16, true>(plan, stream);
      if (p.act == ACT_NONE && flags == EPI_OUT_F32) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_F32, true>(plan, stream);
      if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32 | EPI_REMAP))  // patch embedding (+ positional embedding, row remap)
        return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32 | EPI_REMAP, true>(plan, stream);
      // fp32-class path (3-pass split products): fused QKV and c_fc write bf16 hi | lo
      if (p.act == ACT_NONE && flags == (EPI_OUT_BF16 | EPI_SPLIT)) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_BF16 | EPI_SPLIT, true>(plan, stream);
      if (p.act == ACT_QUICKGELU_PRECISE && flags == (EPI_OUT_BF16 | EPI_SPLIT))
        return launch_gemm_inst<BN, ACT_QUICKGELU_PRECISE, EPI_OUT_BF16 | EPI_SPLIT, true>(plan, stream);
      trace_generic_gemm(BN, p.act, flags, 1, p);
      return launch_gemm_generic<BN, true>(plan, stream);
    }
  }
  // hot ViT-block epilogues get compile-time specialisations; everything else takes the generic instantiation
  if (plan.tma_store && flags == EPI_OUT_BF16 && p.out_bf16 == plan.tma_store_ptr) {
    if (p.act == ACT_NONE) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_BF16 | EPI_TMA_STORE>(plan, stream);
    if (p.act == ACT_QUICKGELU) return launch_gemm_inst<BN, ACT_QUICKGELU, EPI_OUT_BF16 | EPI_TMA_STORE>(plan, stream);
    if (p.act == ACT_RELU) return launch_gemm_inst<BN, ACT_RELU, EPI_OUT_BF16 | EPI_TMA_STORE>(plan, stream);
  }
  if (p.act == ACT_NONE && flags == EPI_OUT_BF16) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_BF16>(plan, stream);
  if (p.act == ACT_QUICKGELU && flags == EPI_OUT_BF16) return launch_gemm_inst<BN, ACT_QUICKGELU, EPI_OUT_BF16>(plan, stream);
  if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32))
    return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32>(plan, stream);
  if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16))
    return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32 | EPI_OUT_BF16>(plan, stream);
  if (p.act == ACT_RELU && flags == EPI_OUT_BF16) return launch_gemm_inst<BN, ACT_RELU, EPI_OUT_BF16>(plan, stream);
  if (p.act == ACT_NONE && flags == EPI_OUT_F32) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_F32>(plan, stream);
  if (p.act == ACT_RELU && flags == (EPI_OUT_F32 | EPI_OUT_BF16)) return launch_gemm_inst<BN, ACT_RELU, EPI_OUT_F32 | EPI_OUT_BF16>(plan, stream);
  if (p.act == ACT_RELU && flags == EPI_OUT_F32) return launch_gemm_inst<BN, ACT_RELU, EPI_OUT_F32>(plan, stream);
  // split-precision MLP hidden layers: ContextDecoder (erf GELU) and text tower (exact QuickGELU), bf16 hi|lo output
  if (p.act == ACT_NONE && flags == (EPI_OUT_BF16 | EPI_SPLIT)) return launch_gemm_inst<BN, ACT_NONE, EPI_OUT_BF16 | EPI_SPLIT>(plan, stream);
  if (p.act == ACT_GELU_ERF && flags == (EPI_OUT_BF16 | EPI_SPLIT)) return launch_gemm_inst<BN, ACT_GELU_ERF, EPI_OUT_BF16 | EPI_SPLIT>(plan, stream);
  if (p.act == ACT_QUICKGELU_PRECISE && flags == (EPI_OUT_BF16 | EPI_SPLIT))
    return launch_gemm_inst<BN, ACT_QUICKGELU_PRECISE, EPI_OUT_BF16 | EPI_SPLIT>(plan, stream);
  if (p.act == ACT_NONE && flags == (EPI_RESID | EPI_OUT_F32 | EPI_REMAP))
    return launch_gemm_inst<BN, ACT_NONE, EPI_RESID | EPI_OUT_F32 | EPI_REMAP>(plan, stream);
  trace_generic_gemm(BN, p.act, flags, 0, p);
  return launch_gemm_generic<BN, false>(plan, stream);
  }
}

// bn = 0: choose automatically.
inline GemmPlan make_gemm_plan(const GemmOperands& op, const GemmParams& p, int bn = 0, int max_ctas = 0) {
  DCLIP_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "bad GEMM shape %d %d %d", p.M, p.N, p.K);
  DCLIP_REQUIRE(p.N % 4 == 0, "GEMM N=%d must be a multiple of 4 (vectorised epilogue)", p.N);
  DCLIP_REQUIRE(p.K % 8 == 0 && op.lda % 8 == 0 && op.ldw % 8 == 0, "GEMM K/lda/ldw must be multiples of 8 (K=%d)", p.K);
  if (p.out_f32) DCLIP_REQUIRE(p.ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0, "out_f32 alignment");
  if (p.out_bf16) DCLIP_REQUIRE(p.ldcb % 4 == 0 && p.split_out_off % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 7) == 0, "out_bf16 alignment");
  if (p.residual) DCLIP_REQUIRE(p.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0, "residual alignment");
  if (p.bias) DCLIP_REQUIRE((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "bias alignment");
  const int cluster_env = DCLIP_KNOB("DCLIP_GEMM_CLUSTER", -1);
  const int num_m_pairs = ((p.M + 127) / 128 + 1) / 2;
  if (bn == 0) {
    bn = p.N > 128 ? 256 : (p.N > 64 ? 128 : 64);
    // latency-bound tiny problems (the ContextDecoder's M = 304 rows): narrow tiles spread the W traffic and the epilogue
    // over 4x more CTAs (measured in a CUDA graph: 10.9 -> 7.2 us, 14.8 -> 8.7 us per launch)
    const bool no_small = DCLIP_KNOB("DCLIP_GEMM_NO_SMALL_BN", 0) != 0;
    if (!no_small && p.conv_C == 0 && (long long)((p.M + 127) / 128) * ((p.N + 255) / 256) * 8 <= sm_count()) bn = 64;
    // wave quantisation of the CTA-pair grid: N = 768 with M = 32784 is 387 pair tiles of 256x256 on 74 pairs (6 waves, 87%
    // full) but 516 tiles of 256x192 (7 waves, 99.6% full).  Only the fp32-residual epilogues are instantiated at 192, and
    // only short-K GEMMs win (measured, B200: K = 768 0.062 -> 0.054 ms; K = 3072 0.136 -> 0.140 ms, the narrower tile
    // costs more operand traffic per MMA than the fuller last wave saves).
    const bool no192 = DCLIP_KNOB("DCLIP_GEMM_NO_BN192", 0) != 0;
    if (bn == 256 && !no192 && cluster_env != 0 && p.K <= 1536 && p.N % 192 == 0 && p.conv_C == 0 && p.residual && p.out_f32 && !p.split_out && p.remap_P == 0) {
      const int pairs = sm_count() / 2;
      const long long u256 = (long long)num_m_pairs * ((p.N + 255) / 256), u192 = (long long)num_m_pairs * (p.N / 192);
      const long long c256 = (u256 + pairs - 1) / pairs * 256, c192 = (u192 + pairs - 1) / pairs * 192;
      if (u192 >= sm_count() && c192 * 100 < c256 * 95) bn = 192;
    }
  }
  DCLIP_REQUIRE(bn == 256 || bn == 192 || bn == 128 || bn == 64, "unsupported BLOCK_N %d", bn);
  GemmPlan plan;
  plan.p = p;
  plan.bn = bn;
  const uint64_t kcols = p.split_in ? 2ull * p.K : uint64_t(p.K);
  // grouped conv, 128 filters per group, bf16 ReLU output, grid tileable by 256 pixels: swapped operand roles (see EPI_CONV_SWAP)
  if (p.conv_C > 0 && p.conv_G > 1 && bn == 128 && !p.split_in && !p.split_out && p.act == ACT_RELU && p.out_bf16 && !p.out_f32 &&
      !p.residual && p.remap_P == 0 && p.conv_C % 64 == 0 && (op.conv_gh * p.conv_gw) % 256 == 0 && 256 % p.conv_gw == 0 &&
      DCLIP_KNOB("DCLIP_GEMM_CONV_SWAP", 1)) {
    const int gw = p.conv_gw, gh = op.conv_gh;
    DCLIP_REQUIRE(p.K == 9 * p.conv_C && p.M == op.conv_B * gh * gw && p.N == p.conv_G * 128, "grouped conv: inconsistent shape");
    plan.conv_swap = true;
    plan.bn = 256;
    plan.p.conv_tiles_per_img = gh * gw / 256;
    plan.p.cluster = 1;
    uint64_t dims[5] = {uint64_t(p.conv_C), uint64_t(gw), uint64_t(gh), uint64_t(op.conv_B), uint64_t(p.conv_G)};
    uint64_t str[4] = {uint64_t(op.lda) * 2, uint64_t(op.lda) * 2 * gw, uint64_t(op.a_bs) * 2, uint64_t(op.a_gs) * 2};
    uint32_t box[5] = {64, uint32_t(gw), uint32_t(256 / gw), 1, 1};
    plan.tmA = make_tmap_bf16(op.A, 5, dims, str, box);
    plan.tmB = make_tmap_2d_bf16(op.W, p.N, kcols, op.ldw, 128);
    memset(&plan.tmC, 0, sizeof(plan.tmC));
    const int tiles = (p.M / 256) * p.conv_G;
    plan.grid = tiles < sm_count() ? tiles : sm_count();
    return plan;
  }
  if (p.conv_C > 0) {
    const int gw = p.conv_gw, gh = op.conv_gh;
    DCLIP_REQUIRE(p.conv_C % 64 == 0 && p.K == 9 * p.conv_C, "implicit conv: C %% 64 == 0 and K == 9*C required");
    DCLIP_REQUIRE((gh * gw) % 128 == 0 && (128 % gw == 0 || gw % 128 == 0), "implicit conv: grid %dx%d not tileable by 128 pixels", gh, gw);
    DCLIP_REQUIRE(p.M == op.conv_B * gh * gw && p.conv_tiles_per_img == gh * gw / 128, "implicit conv: inconsistent M");
    const uint32_t bw = gw < 128 ? gw : 128, bh = 128 / bw;
    uint64_t dims[5] = {uint64_t(p.conv_C) * (p.split_in ? 2 : 1), uint64_t(gw), uint64_t(gh), uint64_t(op.conv_B), uint64_t(p.conv_G > 1 ? p.conv_G : 1)};
    uint64_t str[4] = {uint64_t(op.lda) * 2, uint64_t(op.lda) * 2 * gw, uint64_t(op.a_bs) * 2, uint64_t(op.a_gs) * 2};
    uint32_t box[5] = {64, bw, bh, 1, 1};
    if (p.conv_G > 1) DCLIP_REQUIRE(p.N == p.conv_G * bn, "grouped conv: N (%d) must equal groups (%d) x BLOCK_N (%d)", p.N, p.conv_G, bn);
    plan.tmA = make_tmap_bf16(op.A, p.conv_G > 1 ? 5 : 4, dims, str, box);
  } else {
    plan.tmA = make_tmap_2d_bf16(op.A, p.M, kcols, op.lda, 128);
  }
  // CTA-pair mode (cta_group::2, 256 x 256 tile per pair, each CTA stages half of W): when there are enough pair-units
  const int num_m_blocks = (p.M + 127) / 128;
  bool use_cluster = bn == 256 && p.conv_C == 0 && ((num_m_blocks + 1) / 2) * ((p.N + bn - 1) / bn) >= sm_count();
  if (cluster_env == 0) use_cluster = false;
  if (cluster_env == 2 && bn == 256 && p.conv_C == 0) use_cluster = true;
  if (p.wg_C > 0) {
    DCLIP_REQUIRE(p.conv_C == 0 && !p.split_in && bn != 192 && p.wg_C % bn == 0 && p.N == 9 * p.wg_C && p.wg_rows >= p.wg_C &&
                      p.wg_pitch > 0 && p.wg_pitch % 8 == 0 && op.ldw >= (long long)p.K + 2 * p.wg_pitch,
                  "conv weight-gradient GEMM: N == 9 * wg_C, wg_C %% BLOCK_N == 0, wg_pitch %% 8 == 0 and ldw >= K + 2 * wg_pitch required "
                  "(wg_C=%d, N=%d, BLOCK_N=%d, pitch=%d)", p.wg_C, p.N, bn, p.wg_pitch);
    if (p.wg_grouped) DCLIP_REQUIRE(p.M % 128 == 0 && p.wg_rows == (p.M / 128) * p.wg_C, "grouped weight-gradient GEMM: 128 filters per group");
    use_cluster = false;
  }
  if (bn == 192) {
    DCLIP_REQUIRE(p.conv_C == 0, "BLOCK_N 192: plain GEMM only");
    use_cluster = true;
  }
  plan.p.cluster = use_cluster ? 2 : 1;
  plan.tmB = p.wg_C > 0 ? make_tmap_2d_bf16(op.W, 3ull * p.wg_rows, uint64_t(op.ldw), op.ldw, bn)
                        : make_tmap_2d_bf16(op.W, p.N, kcols, op.ldw, use_cluster ? bn / 2 : bn);
  memset(&plan.tmC, 0, sizeof(plan.tmC));
  const bool no_tma_store = DCLIP_KNOB("DCLIP_GEMM_NO_TMA_STORE", 0) != 0;
  if (!no_tma_store && bn != 192 && p.out_bf16 && !p.out_f32 && !p.residual && !p.split_out && p.remap_P == 0 && p.ldcb % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0 && (p.dbg_mode == 0 || p.dbg_mode == 5)) {
    uint64_t dims[2] = {uint64_t(p.N), uint64_t(p.M)};
    uint64_t str[1] = {uint64_t(p.ldcb) * 2};
    uint32_t box[2] = {64, 32};
    plan.tmC = make_tmap_bf16(p.out_bf16, 2, dims, str, box);
    plan.tma_store = true;
    plan.tma_store_ptr = p.out_bf16;
  }
  // LayerNorm folding (GemmParams::row_stats_*): only the epilogues that implement it may be selected
  if (p.row_stats_in)
    DCLIP_REQUIRE(plan.tma_store && p.ln_c && p.stats_n >= 1 && p.stats_n <= 12 && p.stats_ld >= p.M && p.N >= 8,
                  "folded-LayerNorm consumer GEMM needs the TMA-store epilogue (plain bf16 output), ln_c and 1..12 statistics slots");
  if (p.row_stats_out)
    DCLIP_REQUIRE(p.residual && p.out_f32 && p.act == ACT_NONE && !p.split_in && !p.split_out && p.remap_P == 0 && p.conv_C == 0 &&
                      p.stats_ld >= p.M && 2 * ((p.N + bn - 1) / bn) <= 12,
                  "row statistics are published by the fp32-residual epilogues only (N / BLOCK_N <= 6)");
  const int num_tiles = ((p.M + 127) / 128) * ((p.N + bn - 1) / bn);
  plan.grid = num_tiles < sm_count() ? num_tiles : sm_count();
  if (max_ctas > 0 && plan.grid > max_ctas) plan.grid = max_ctas;
  if (use_cluster) {
    const int units = ((num_m_blocks + 1) / 2) * ((p.N + bn - 1) / bn);
    int pairs = sm_count() / 2;
    if (units < pairs) pairs = units;
    plan.grid = 2 * pairs;
  }
  return plan;
}

inline void run_gemm(const GemmPlan& plan, cudaStream_t stream) {
  switch (plan.bn) {
    case 256: launch_gemm_bn<256>(plan, stream); break;
    case 192: launch_gemm_bn<192>(plan, stream); break;
    case 128: launch_gemm_bn<128>(plan, stream); break;
    case 64: launch_gemm_bn<64>(plan, stream); break;
    default: throw Error{"unsupported BLOCK_N"};
  }
}

inline void launch_gemm(const GemmOperands& op, const GemmParams& p, cudaStream_t stream, int bn = 0, int max_ctas = 0) {
  run_gemm(make_gemm_plan(op, p, bn, max_ctas), stream);
}

// ---------------------------------------------------------------------------------------------------------
// attention
// ---------------------------------------------------------------------------------------------------------
// [B][N][ld] bf16 token-major tensor -> 3D map {ld cols, N, B} with box {64, 128, 1}
inline CUtensorMap make_tmap_tokens_bf16(const void* base, uint64_t B, uint64_t N, uint64_t ld, uint64_t batch_stride) {
  uint64_t dims[3] = {ld, N, B};
  uint64_t str[2] = {ld * 2, batch_stride * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(base, 3, dims, str, box);
}

using Cfg128 = AttnCfg;
struct AttnPlan {
  CUtensorMap tmQ, tmK, tmV, tmO;
  AttnParams p;
  int grid = 0;
};

struct AttnOperands {
  const __nv_bfloat16 *q, *k, *v;  // token-major [B][N][ld]
  int ldq, ldk, ldv;
  long long q_bs, k_bs, v_bs;      // batch strides in elements
  int Nq_total;
};

inline AttnPlan make_attn_plan(const AttnOperands& op, const AttnParams& p) {
  DCLIP_REQUIRE(p.B > 0 && p.H > 0 && p.Nk > 0 && p.Nq_total > p.q_start, "bad attention shape");
  DCLIP_REQUIRE(p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && p.out_batch_stride % 8 == 0, "attention out alignment");
  AttnPlan plan;
  plan.p = p;
  plan.p.token_mode = DCLIP_KNOB("DCLIP_ATTN_TOKEN", 1);
  const int tail_env = DCLIP_KNOB("DCLIP_ATTN_TAIL_ROWS", 4);
  plan.p.q = op.q; plan.p.k = op.k; plan.p.v = op.v;
  plan.p.ldq = op.ldq; plan.p.ldk = op.ldk; plan.p.ldv = op.ldv;
  plan.p.q_bs = op.q_bs; plan.p.k_bs = op.k_bs; plan.p.v_bs = op.v_bs;
  const int tail_ov_env = DCLIP_KNOB("DCLIP_ATTN_TAIL_OVERLAP", 1);
  plan.p.tail_overlap = (tail_ov_env && p.Nk >= 32) ? 1 : 0;  // (the online softmax of the background path assumes every key group sees a valid key in its first trip)
  const int peel_env = DCLIP_KNOB("DCLIP_ATTN_PEEL", 1);
  plan.p.peel_key0 = (peel_env && p.Nk > Cfg128::TKV && (p.Nk - 1) % Cfg128::TKV == 0) ? peel_env : 0;  // 2: also L2 / early L1 prefetch of k_0, v_0
  // the CUDA-core tail path keeps one fp32 score per key in shared memory
  plan.p.tail_rows_max = (size_t(p.Nk) * 4 + 16384 <= size_t(AttnCfg::SMEM_BYTES)) ? tail_env : 0;
  DCLIP_REQUIRE(op.ldq % 8 == 0 && op.ldk % 8 == 0 && op.ldv % 8 == 0 && p.q_col0 % 8 == 0 && p.k_col0 % 8 == 0 && p.v_col0 % 8 == 0 &&
                op.q_bs % 8 == 0 && op.k_bs % 8 == 0 && op.v_bs % 8 == 0, "attention operand alignment");
  plan.tmQ = make_tmap_tokens_bf16(op.q, p.B, op.Nq_total, op.ldq, op.q_bs);
  plan.tmK = make_tmap_tokens_bf16(op.k, p.B, p.Nk, op.ldk, op.k_bs);
  plan.tmV = make_tmap_tokens_bf16(op.v, p.B, p.Nk, op.ldv, op.v_bs);
  plan.tmO = make_tmap_tokens_bf16(p.out, p.B, op.Nq_total, p.ldo, p.out_batch_stride);
  const int nqb = (p.Nq_total - p.q_start + 255) / 256;
  plan.grid = nqb * p.H * p.B;
  return plan;
}

template <bool PT, int POLY, int MODE>
inline void run_attn_variant(const AttnPlan& plan, cudaStream_t stream) {
  using Cfg = AttnCfgT<PT>;
  ensure_dyn_smem(attn_fwd_tcgen05_kernel<PT, POLY, MODE>, Cfg::SMEM_BYTES);
  attn_fwd_tcgen05_kernel<PT, POLY, MODE><<<plan.grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(plan.tmQ, plan.tmK, plan.tmV, plan.tmO, plan.p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
}

#ifdef DCLIP_WITH_ATTN_P4
template <int POLY>
inline void run_attn_p4(const AttnPlan& plan, int grid, cudaStream_t stream) {
  using Cfg = AttnP4Cfg;
  ensure_dyn_smem(attn_fwd_p4_kernel<POLY>, Cfg::SMEM_BYTES);
  attn_fwd_p4_kernel<POLY><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(plan.tmQ, plan.tmK, plan.tmV, plan.tmO, plan.p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
}
#endif

template <int POLY>
inline void run_attn_persistent(const AttnPlan& plan, int grid, cudaStream_t stream) {
  using Cfg = AttnPersistCfg;
  ensure_dyn_smem(attn_fwd_persistent_kernel<POLY, 0>, Cfg::SMEM_BYTES);
  attn_fwd_persistent_kernel<POLY, 0><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(plan.tmQ, plan.tmK, plan.tmV, plan.tmO, plan.p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
}

// Production dispatch: the persistent kernel (P in TMEM, MUFU token, background tail rows, peeled key 0) whenever there is
// at least one regular 256-query block; the one-CTA-per-item kernel only serves launches without any (a handful of rows).
// The variants measured slower in round 1 / 2 (speculative max, deferred P stores, P through shared memory, polynomial exp2,
// the 4-warpgroup P4 structure, the column-split CS structure) exist only in selftest builds (-DDCLIP_EXPERIMENTS): profiles/r01_attention_notes.md,
// profiles/r02_attention_notes.md.
#ifndef DCLIP_EXPERIMENTS
inline void run_attn(const AttnPlan& plan, cudaStream_t stream) {
  const AttnParams& q = plan.p;
  const int nqb = (q.Nq_total - q.q_start + 255) / 256;
  const int rows_last = q.Nq_total - q.q_start - (nqb - 1) * 256;
  const int n_reg = q.B * q.H * (nqb - (rows_last <= q.tail_rows_max ? 1 : 0));
  if (n_reg > 0) return run_attn_persistent<0>(plan, n_reg < sm_count() ? n_reg : sm_count(), stream);
  return run_attn_variant<true, 0, 0>(plan, stream);
}
#else
template <int POLY>
inline void run_attn_cs(const AttnPlan& plan, int grid, cudaStream_t stream) {
  using Cfg = AttnCsCfg;
  ensure_dyn_smem(attn_fwd_cs_kernel<POLY>, Cfg::SMEM_BYTES);
  attn_fwd_cs_kernel<POLY><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(plan.tmQ, plan.tmK, plan.tmV, plan.tmO, plan.p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
}


// Production variant: P in TMEM (TS MMA) + DCLIP_ATTN_POLY_DEFAULT of every 4 exp2 pairs on the FMA pipe.
// A/B knobs (selftests only): DCLIP_ATTN_P_SMEM=1 (P through shared memory), DCLIP_ATTN_POLY=0|1|2.
#ifndef DCLIP_ATTN_POLY_DEFAULT
#define DCLIP_ATTN_POLY_DEFAULT 0
#endif
#ifndef DCLIP_ATTN_SPEC_DEFAULT
#define DCLIP_ATTN_SPEC_DEFAULT 0  // speculative-max softmax tiles (measured slower with the MUFU token: 0.366 vs 0.349 ms)
#endif
#ifndef DCLIP_ATTN_PERSIST_DEFAULT
#define DCLIP_ATTN_PERSIST_DEFAULT 1  // persistent one-CTA-per-SM kernel (attn_fwd_persistent_kernel)
#endif
#ifndef DCLIP_ATTN_DEFER_DEFAULT
#define DCLIP_ATTN_DEFER_DEFAULT 0  // deferred P stores (measured slower: 0.373 vs 0.349 ms, profiles/r01_attention_notes.md)
#endif
inline void run_attn(const AttnPlan& plan, cudaStream_t stream) {
  static const bool p_smem = [] { const char* e = getenv("DCLIP_ATTN_P_SMEM"); return e && e[0] == '1'; }();
  static const int poly = [] { const char* e = getenv("DCLIP_ATTN_POLY"); return e ? atoi(e) : DCLIP_ATTN_POLY_DEFAULT; }();
  static const int spec = [] { const char* e = getenv("DCLIP_ATTN_SPEC"); return e ? atoi(e) : DCLIP_ATTN_SPEC_DEFAULT; }();
  static const int defer = [] { const char* e = getenv("DCLIP_ATTN_DEFER"); return e ? atoi(e) : DCLIP_ATTN_DEFER_DEFAULT; }();
  static const int persist = [] { const char* e = getenv("DCLIP_ATTN_PERSIST"); return e ? atoi(e) : DCLIP_ATTN_PERSIST_DEFAULT; }();
  if (persist && !p_smem && !spec && !defer) {
    // persistent kernel: one CTA per SM over a static item list (needs at least one regular 256-query block per CTA)
    const AttnParams& q = plan.p;
    const int nqb = (q.Nq_total - q.q_start + 255) / 256;
    const int rows_last = q.Nq_total - q.q_start - (nqb - 1) * 256;
    const int n_reg = q.B * q.H * (nqb - (rows_last <= q.tail_rows_max ? 1 : 0));
    if (n_reg > 0) {
      const int grid = n_reg < sm_count() ? n_reg : sm_count();
#ifdef DCLIP_WITH_ATTN_P4
      static const int impl = [] { const char* e = getenv("DCLIP_ATTN_IMPL"); return e ? atoi(e) : 2; }();   // 4: P4 experiment, 2: persistent kernel
      if (impl == 4 && q.tail_overlap) {
        if (poly == 2) return run_attn_p4<2>(plan, grid, stream);
        if (poly) return run_attn_p4<1>(plan, grid, stream);
        return run_attn_p4<0>(plan, grid, stream);
      }
#endif
      static const int cs = [] { const char* e = getenv("DCLIP_ATTN_CS"); return e ? atoi(e) : 0; }();
      if (cs && (q.tail_overlap || rows_last > q.tail_rows_max)) {
        if (poly == 2) return run_attn_cs<2>(plan, grid, stream);
        if (poly) return run_attn_cs<1>(plan, grid, stream);
        return run_attn_cs<0>(plan, grid, stream);
      }
      if (poly) return run_attn_persistent<1>(plan, grid, stream);
      return run_attn_persistent<0>(plan, grid, stream);
    }
  }
  if (p_smem) return run_attn_variant<false, 0, 0>(plan, stream);
  switch ((poly ? 4 : 0) | (defer ? 2 : 0) | (spec ? 1 : 0)) {
    case 0: return run_attn_variant<true, 0, 0>(plan, stream);
    case 1: return run_attn_variant<true, 0, 1>(plan, stream);
    case 2: return run_attn_variant<true, 0, 2>(plan, stream);
    case 3: return run_attn_variant<true, 0, 3>(plan, stream);
    case 4: return run_attn_variant<true, 1, 0>(plan, stream);
    case 5: return run_attn_variant<true, 1, 1>(plan, stream);
    case 6: return run_attn_variant<true, 1, 2>(plan, stream);
    default: return run_attn_variant<true, 1, 3>(plan, stream);
  }
}

#endif  // DCLIP_EXPERIMENTS

// ---- fp32-class (hi|lo split) tensor-core attention ------------------------------------------------------------
struct AttnSplitPlan {
  CUtensorMap tmQ, tmK, tmV;
  AttnSplitParams p;
  int grid = 0;
};

struct AttnSplitOperands {
  const __nv_bfloat16 *q, *k, *v;  // token-major [B][N][ld], each row holding hi and (lo_off columns further) lo halves
  int ldq, ldk, ldv;
  long long q_bs, k_bs, v_bs;
};

inline AttnSplitPlan make_attn_split_plan(const AttnSplitOperands& op, const AttnSplitParams& p) {
  DCLIP_REQUIRE(p.B > 0 && p.H > 0 && p.Nk > 0 && p.Nq > 0, "bad attention shape");
  DCLIP_REQUIRE(p.ldo % 8 == 0 && p.out_lo_off % 8 == 0 && p.out_bs % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0,
                "split attention out alignment");
  DCLIP_REQUIRE(op.ldq % 8 == 0 && op.ldk % 8 == 0 && op.ldv % 8 == 0 && p.q_col0 % 8 == 0 && p.k_col0 % 8 == 0 && p.v_col0 % 8 == 0 &&
                p.lo_off % 8 == 0 && op.q_bs % 8 == 0 && op.k_bs % 8 == 0 && op.v_bs % 8 == 0, "split attention operand alignment");
  AttnSplitPlan plan;
  plan.p = p;
  plan.tmQ = make_tmap_tokens_bf16(op.q, p.B, p.Nq, op.ldq, op.q_bs);
  plan.tmK = make_tmap_tokens_bf16(op.k, p.B, p.Nk, op.ldk, op.k_bs);
  plan.tmV = make_tmap_tokens_bf16(op.v, p.B, p.Nk, op.ldv, op.v_bs);
  plan.grid = ((p.Nq + 255) / 256) * p.H * p.B;
  return plan;
}

inline void run_attn_split(const AttnSplitPlan& plan, cudaStream_t stream) {
  ensure_dyn_smem(attn_fwd_split_kernel, AttnSplitCfg::SMEM_BYTES);
  attn_fwd_split_kernel<<<plan.grid, AttnSplitCfg::THREADS, AttnSplitCfg::SMEM_BYTES, stream>>>(plan.tmQ, plan.tmK, plan.tmV, plan.p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
}

// scratch for the key-split partials: grown on demand.  dclip_api.cu keeps one per (handle, stream), so calls on
// different streams never share a buffer.  A block that has been handed to a launch is NEVER freed before the handle is
// destroyed: a captured CUDA graph may have its address baked in (growing allocates a new block and retires the old one).
struct AttnSmallScratch {
  float* ptr = nullptr;
  size_t bytes = 0;
  std::vector<float*> retired;
  void release() {
    if (ptr) cudaFree(ptr);
    for (float* r : retired) cudaFree(r);
    ptr = nullptr; bytes = 0; retired.clear();
  }
};

template <int QB>
inline int launch_attn_small(SmallAttnParams p, cudaStream_t stream, AttnSmallScratch* scratch) {
  const int nqb = (p.q_count + QB - 1) / QB;
  const int ctas = p.B * p.H * nqb;
  // long key ranges with few (batch, head, query-block) CTAs -- the ContextDecoder cross attention: 19 queries over 2048
  // keys -- are split over the keys (flash-decoding style) so the K/V stream is spread over ~4 waves of CTAs
  int S = 1;
  const bool no_split = DCLIP_KNOB("DCLIP_ATTN_SMALL_NO_SPLIT", 0) != 0;
  if (scratch && !no_split && !p.causal && p.Nk >= 1024) {
    S = (4 * sm_count() + ctas / 2) / ctas;
    S = S < 1 ? 1 : (S > 8 ? 8 : S);
    while (S > 1 && (p.Nk + S - 1) / S < 256) --S;
  }
  if (S > 1) {
    const size_t need = size_t(ctas) * S * QB * 66 * sizeof(float);
    if (scratch->bytes < need) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      DCLIP_CHECK_CUDA(cudaStreamIsCapturing(stream, &cs));
      DCLIP_REQUIRE(cs == cudaStreamCaptureStatusNone, "small attention: scratch must be sized by an eager warm-up call before graph capture");
      if (scratch->ptr) scratch->retired.push_back(scratch->ptr);  // (kept alive: graphs captured earlier still point at it)
      scratch->ptr = nullptr; scratch->bytes = 0;
      DCLIP_CHECK_CUDA(cudaMalloc(&scratch->ptr, need));
      scratch->bytes = need;
    }
    p.key_splits = S;
    p.ws = scratch->ptr;
  } else {
    p.key_splits = 1;
    p.ws = nullptr;
  }
  const int span = (p.Nk + S - 1) / S;
  const size_t smem = (((size_t(QB) * span + 3) & ~size_t(3)) + QB * 64 + 8 * QB * 64 + 16 * QB) * 4;
  if constexpr (QB > 8) {
    if (smem > 200 * 1024) return launch_attn_small<8>(p, stream, scratch);  // (key splitting disabled: scores of 20 queries do not fit)
  }
  DCLIP_REQUIRE(smem <= 200 * 1024, "small attention: Nk=%d too large for the smem score buffer", p.Nk);
  ensure_dyn_smem(attn_small_kernel<QB>, 200 * 1024);
  attn_small_kernel<QB><<<ctas * S, 256, smem, stream>>>(p);
  DCLIP_CHECK_CUDA(cudaGetLastError());
  if (S > 1) {
    attn_small_combine_kernel<QB><<<ctas * QB, 64, 0, stream>>>(p);
    DCLIP_CHECK_CUDA(cudaGetLastError());
  }
  return S > 1 ? 2 : 1;
}

// returns the number of kernels launched
inline int run_attn_small(const SmallAttnParams& p, cudaStream_t stream, AttnSmallScratch* scratch = nullptr) {
  DCLIP_REQUIRE(p.q_count > 0 && p.Nk > 0, "bad small-attention shape");
  if (p.q_count == 1) return launch_attn_small<1>(p, stream, scratch);
  // 9..20 queries over a long key range (ContextDecoder cross attention: 19 class queries x 2048 visual tokens): one query
  // block per (batch, head), so K and V are streamed once instead of once per 8 queries; the keys are split over CTAs
  if (scratch && !p.causal && p.q_count > 8 && p.q_count <= 20 && p.Nk >= 1024) return launch_attn_small<20>(p, stream, scratch);
  if (p.q_count <= 4 || p.Nk > 4096) return launch_attn_small<4>(p, stream, scratch);
  return launch_attn_small<8>(p, stream, scratch);
}

}  // namespace dclip
