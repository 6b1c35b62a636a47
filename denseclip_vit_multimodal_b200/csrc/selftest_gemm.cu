// Standalone GPU sanity/timing binary for the tcgen05 GEMM (no torch): compares against a CUDA-core fp32 kernel.
// Build: make -C denseclip_vit_multimodal_b200/csrc selftest_gemm ; run on a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_utils.cuh"

using namespace dclip;

__global__ void ref_gemm_kernel(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, float* C, int M, int N,
                                int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(A[size_t(m) * lda + k]), __bfloat162float(W[size_t(n) * ldw + k]), acc);
  C[size_t(m) * N + n] = acc;
}

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / float(0x1000000) - 0.5f;
  p[i] = __float2bfloat16(u * scale);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  p[i] = ((x & 0xffffff) / float(0x1000000) - 0.5f) * scale;
}

static int g_fail = 0;

static float host_act(float x, int act) {
  switch (act) {
    case ACT_QUICKGELU:
    case ACT_QUICKGELU_PRECISE: return x / (1.f + expf(-1.702f * x));
    case ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.70710678f));
    case ACT_RELU: return x > 0 ? x : 0;
    default: return x;
  }
}

// mode: 0 plain bf16 out, 1 bias+quickgelu bf16, 2 bias+residual f32 (in place), 3 patch remap + pos, 4 f32+bf16 split out
static void run_case(int M, int N, int K, int bn, int mode, bool timing) {
  __nv_bfloat16 *A, *W, *Cb;
  float *Cref, *bias, *resid, *Cf;
  int P = 0, Nt = 0, rows_out = M;
  if (mode == 3) { P = 50; Nt = P + 1; rows_out = (M / P) * Nt; M = (M / P) * P; }
  cudaMalloc(&A, size_t(M) * K * 2);
  cudaMalloc(&W, size_t(N) * K * 2);
  cudaMalloc(&Cb, size_t(rows_out) * 2 * N * 2);
  cudaMalloc(&Cf, size_t(rows_out) * N * 4);
  cudaMalloc(&Cref, size_t(M) * N * 4);
  cudaMalloc(&bias, N * 4);
  cudaMalloc(&resid, size_t(rows_out) * N * 4);
  fill_bf16<<<(size_t(M) * K + 255) / 256, 256>>>(A, size_t(M) * K, 1u, 2.0f);
  fill_bf16<<<(size_t(N) * K + 255) / 256, 256>>>(W, size_t(N) * K, 2u, 0.25f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 3u, 1.0f);
  fill_f32<<<(size_t(rows_out) * N + 255) / 256, 256>>>(resid, size_t(rows_out) * N, 4u, 2.0f);
  cudaMemset(Cb, 0, size_t(rows_out) * 2 * N * 2);
  cudaMemset(Cf, 0, size_t(rows_out) * N * 4);
  std::vector<float> h_resid(size_t(rows_out) * N), h_bias(N);
  cudaMemcpy(h_resid.data(), resid, h_resid.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_bias.data(), bias, N * 4, cudaMemcpyDeviceToHost);

  GemmOperands op{A, K, W, K};
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.out_scale = 1.0f;
  switch (mode) {
    case 0: p.out_bf16 = Cb; p.ldcb = N; break;
    case 1: p.bias = bias; p.act = ACT_QUICKGELU; p.out_bf16 = Cb; p.ldcb = N; break;
    case 2: p.bias = bias; p.residual = Cf; p.ldr = N; p.out_f32 = Cf; p.ldc = N;
            cudaMemcpy(Cf, resid, size_t(rows_out) * N * 4, cudaMemcpyDeviceToDevice); break;
    case 3: p.residual = resid; p.ldr = N; p.res_mod = 1; p.remap_P = P; p.remap_Nt = Nt; p.out_f32 = Cf; p.ldc = N; break;
    case 4: p.bias = bias; p.act = ACT_QUICKGELU_PRECISE; p.out_f32 = Cf; p.ldc = N; p.out_bf16 = Cb; p.ldcb = 2 * N;
            p.split_out = 1; p.split_out_off = N; break;
  }
  try {
    launch_gemm(op, p, 0, bn);
  } catch (Error& e) {
    printf("FAIL launch: %s\n", e.msg.c_str());
    g_fail++;
    return;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL M=%d N=%d K=%d bn=%d mode=%d: CUDA error %s\n", M, N, K, bn, mode, cudaGetErrorString(e));
    g_fail++;
    exit(2);
  }
  ref_gemm_kernel<<<dim3((N + 127) / 128, M), 128>>>(A, K, W, K, Cref, M, N, K);
  cudaDeviceSynchronize();
  std::vector<float> h_ref(size_t(M) * N), h_cf(size_t(rows_out) * N);
  std::vector<__nv_bfloat16> h_cb(size_t(rows_out) * 2 * N);
  cudaMemcpy(h_ref.data(), Cref, h_ref.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_cf.data(), Cf, h_cf.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_cb.data(), Cb, h_cb.size() * 2, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  size_t bad = 0, first_bad = size_t(-1);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = h_ref[size_t(m) * N + n], got = 0, tol = 0;
      size_t orow = m;
      switch (mode) {
        case 0: got = __bfloat162float(h_cb[size_t(m) * N + n]); tol = fabsf(ref) * 0.008f + 1e-3f; break;
        case 1: ref = host_act(ref + h_bias[n], ACT_QUICKGELU); got = __bfloat162float(h_cb[size_t(m) * N + n]);
                tol = fabsf(ref) * 0.01f + 2e-2f; break;
        case 2: ref = ref + h_bias[n] + h_resid[size_t(m) * N + n]; got = h_cf[size_t(m) * N + n];
                tol = fabsf(ref) * 1e-4f + 2e-3f; break;
        case 3: { int bi = m / P, pm = m % P; orow = size_t(bi) * Nt + 1 + pm;
                ref = ref + h_resid[size_t(1 + pm) * N + n]; got = h_cf[orow * N + n];
                tol = fabsf(ref) * 1e-4f + 2e-3f; break; }
        case 4: { ref = host_act(ref + h_bias[n], ACT_QUICKGELU_PRECISE);
                float hi = __bfloat162float(h_cb[size_t(m) * 2 * N + n]), lo = __bfloat162float(h_cb[size_t(m) * 2 * N + N + n]);
                got = hi + lo; float gf = h_cf[size_t(m) * N + n];
                tol = fabsf(ref) * 1e-4f + 2e-3f;
                if (fabsf(gf - got) > fabsf(gf) * 2e-5f + 1e-6f) { bad++; }
                break; }
      }
      double err = fabs(double(got) - ref);
      if (err > max_err) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (!(err <= tol)) {
        if (first_bad == size_t(-1)) first_bad = size_t(m) * N + n;
        bad++;
      }
    }
  printf("%s M=%d N=%d K=%d bn=%d mode=%d max_abs_err=%.3e max_ref=%.3e bad=%zu", bad ? "FAIL" : "ok  ", M, N, K, bn, mode,
         max_err, max_ref, bad);
  if (bad) {
    g_fail++;
    printf(" first_bad=(m=%zu,n=%zu)", first_bad / N, first_bad % N);
  }
  if (timing && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    GemmPlan plan = make_gemm_plan(op, p, bn);
    for (int i = 0; i < 3; ++i) run_gemm(plan, 0);
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; ++i) run_gemm(plan, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("  time=%.3f ms  %.1f TFLOP/s", ms, 2.0 * M * N * K / ms * 1e-9);
  }
  printf("\n");
  fflush(stdout);
  cudaFree(A); cudaFree(W); cudaFree(Cb); cudaFree(Cf); cudaFree(Cref); cudaFree(bias); cudaFree(resid);
}

// split-bf16 ("fp32 path") check: A, W fp32 -> [hi|lo] bf16, GEMM with split_in vs fp64 host reference
static void run_split_case(int M, int N, int K) {
  std::vector<float> hA(size_t(M) * K), hW(size_t(N) * K);
  srand(7);
  for (auto& v : hA) v = (rand() / float(RAND_MAX) - 0.5f) * 2.f;
  for (auto& v : hW) v = (rand() / float(RAND_MAX) - 0.5f) * 0.25f;
  std::vector<__nv_bfloat16> hA2(size_t(M) * 2 * K), hW2(size_t(N) * 2 * K);
  auto split = [](const std::vector<float>& src, std::vector<__nv_bfloat16>& dst, int R, int K) {
    for (int r = 0; r < R; ++r)
      for (int k = 0; k < K; ++k) {
        float v = src[size_t(r) * K + k];
        __nv_bfloat16 hi = __float2bfloat16(v);
        dst[size_t(r) * 2 * K + k] = hi;
        dst[size_t(r) * 2 * K + K + k] = __float2bfloat16(v - __bfloat162float(hi));
      }
  };
  split(hA, hA2, M, K);
  split(hW, hW2, N, K);
  __nv_bfloat16 *A, *W;
  float* C;
  cudaMalloc(&A, hA2.size() * 2); cudaMalloc(&W, hW2.size() * 2); cudaMalloc(&C, size_t(M) * N * 4);
  cudaMemcpy(A, hA2.data(), hA2.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(W, hW2.data(), hW2.size() * 2, cudaMemcpyHostToDevice);
  GemmOperands op{A, 2 * K, W, 2 * K};
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.split_in = 1; p.out_f32 = C; p.ldc = N; p.out_scale = 1.f;
  try { launch_gemm(op, p, 0, 0); } catch (Error& e) { printf("FAIL launch: %s\n", e.msg.c_str()); g_fail++; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL split: %s\n", cudaGetErrorString(e)); g_fail++; exit(2); }
  std::vector<float> hC(size_t(M) * N);
  cudaMemcpy(hC.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += double(hA[size_t(m) * K + k]) * hW[size_t(n) * K + k];
      max_err = fmax(max_err, fabs(acc - hC[size_t(m) * N + n]));
      max_ref = fmax(max_ref, fabs(acc));
    }
  bool ok = max_err <= 2e-5 * max_ref + 1e-6;
  printf("%s split-bf16x3 M=%d N=%d K=%d max_abs_err=%.3e max_ref=%.3e rel=%.2e\n", ok ? "ok  " : "FAIL", M, N, K, max_err,
         max_ref, max_err / max_ref);
  if (!ok) g_fail++;
  cudaFree(A); cudaFree(W); cudaFree(C);
}

int main(int argc, char** argv) {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
    printf("no CUDA device\n");
    return 3;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  bool quick = argc > 1 && !strcmp(argv[1], "quick");
  if (argc > 1 && !strcmp(argv[1], "dbg")) {  // where does the GEMM time go? (epilogue variants)
    const int M = 16 * 2049, N = 2304, K = 768;
    __nv_bfloat16 *A, *W, *Cb;
    cudaMalloc(&A, size_t(M) * K * 2); cudaMalloc(&W, size_t(N) * K * 2); cudaMalloc(&Cb, size_t(M) * N * 2);
    fill_bf16<<<(size_t(M) * K + 255) / 256, 256>>>(A, size_t(M) * K, 1u, 2.0f);
    fill_bf16<<<(size_t(N) * K + 255) / 256, 256>>>(W, size_t(N) * K, 2u, 0.25f);
    for (int mode = 0; mode < 6; ++mode) {
      GemmOperands op{A, K, W, K};
      GemmParams p{};
      p.M = M; p.N = N; p.K = K; p.out_scale = 1.f; p.out_bf16 = Cb; p.ldcb = N; p.dbg_mode = mode;
      if (mode == 3) { p.dbg_mode = 0; p.ldcb = 0; }            // all rows land on one 4.6 KB row: stores hit L2 only
      if (mode == 4) { p.dbg_mode = 0; p.out_bf16 = nullptr; p.out_f32 = reinterpret_cast<float*>(Cb); p.ldc = 0; }  // fp32 stores, L2 only
      GemmPlan plan = make_gemm_plan(op, p, 256);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; ++i) run_gemm(plan, 0);
      cudaEventRecord(e0);
      for (int i = 0; i < 20; ++i) run_gemm(plan, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
      printf("dbg_mode=%d (0 full, 1 no staging/stores, 2 staging only, 3 bf16 stores to one row, 4 fp32 stores to one row)  %.3f ms  %.1f TFLOP/s\n", mode, ms, 2.0 * M * N * K / ms * 1e-9);
    }
#ifdef DCLIP_GEMM_TIMELINE
    {
      const int N2 = 768;
      float* X; cudaMalloc(&X, size_t(M) * N2 * 4); cudaMemset(X, 0, size_t(M) * N2 * 4);
      long long* dbg; cudaMalloc(&dbg, 512 * 8); cudaMemset(dbg, 0, 512 * 8);
      GemmOperands op{A, K, W, K};
      GemmParams p{};
      p.M = M; p.N = N2; p.K = K; p.out_scale = 1.f; p.out_f32 = X; p.ldc = N2; p.residual = X; p.ldr = N2; p.dbg = dbg;
      GemmPlan plan = make_gemm_plan(op, p, 256);
      run_gemm(plan, 0); cudaDeviceSynchronize();
      cudaMemset(dbg, 0, 512 * 8);
      run_gemm(plan, 0); cudaDeviceSynchronize();
      std::vector<long long> h(512);
      cudaMemcpy(h.data(), dbg, 512 * 8, cudaMemcpyDeviceToHost);
      long long t0 = h[0];
      printf("GEMM timeline CTA0 (out-proj shape, cycles): it | epi: wait_start tfull_got chunk0 chunk2 chunk4 chunk6 | mma: first_kb last_kb\n");
      for (int it = 0; it < 7; ++it)
        printf("%d | %7lld %7lld %7lld %7lld %7lld %7lld | %7lld %7lld\n", it, h[it * 8] - t0, h[it * 8 + 1] - t0, h[it * 8 + 2] - t0, h[it * 8 + 3] - t0,
               h[it * 8 + 4] - t0, h[it * 8 + 5] - t0, h[256 + it * 2] - t0, h[256 + it * 2 + 1] - t0);
    }
#endif
    {  // fp32 residual epilogue (out-proj shape): where does the time go?
      const int N2 = 768;
      float* X; cudaMalloc(&X, size_t(M) * N2 * 4); cudaMemset(X, 0, size_t(M) * N2 * 4);
      float* bias; cudaMalloc(&bias, N2 * 4); cudaMemset(bias, 0, N2 * 4);
      for (int mode = 0; mode < 5; ++mode) {
        GemmOperands op{A, K, W, K};
        GemmParams p{};
        p.M = M; p.N = N2; p.K = K; p.out_scale = 1.f; p.bias = bias; p.out_f32 = X; p.ldc = N2;
        if (mode != 3) { p.residual = X; p.ldr = N2; }
        p.dbg_mode = mode == 3 ? 0 : (mode == 4 ? 6 : mode);
        GemmPlan plan = make_gemm_plan(op, p, 256);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) run_gemm(plan, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) run_gemm(plan, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
        printf("outproj dbg_mode=%d (0 full st.cs, 1 drain only, 2 no stores, 3 no residual, 4 full plain st)  %.3f ms  %.1f TFLOP/s\n", mode, ms,
               2.0 * M * N2 * K / ms * 1e-9);
      }
    }
    return 0;
  }
  if (argc > 6 && !strcmp(argv[1], "prof")) {  // prof M N K bn mode  (single shape, for ncu)
    run_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), true);
    return g_fail ? 1 : 0;
  }
  // correctness: tails in M, N, K; every BLOCK_N; every epilogue
  run_case(128, 64, 64, 64, 0, false);
  run_case(128, 128, 64, 128, 0, false);
  run_case(128, 256, 64, 256, 0, false);
  run_case(128, 256, 256, 256, 0, false);
  run_case(300, 200, 192, 64, 0, false);
  run_case(300, 200, 192, 128, 0, false);
  run_case(300, 520, 192, 256, 0, false);
  run_case(1000, 768, 768, 256, 1, false);
  run_case(1000, 768, 768, 128, 2, false);
  run_case(1000, 768, 768, 256, 3, false);
  run_case(777, 512, 320, 256, 4, false);
  run_case(20000, 512, 256, 256, 2, false);  // many tiles per CTA: exercises ring + accumulator phases
  run_case(1000, 768, 768, 192, 2, false);   // CTA-pair 256x192 tiles (fp32-residual epilogue)
  run_case(20000, 384, 256, 192, 2, false);
  run_case(777, 576, 320, 192, 4, false);
  run_split_case(200, 256, 256);
  if (!quick) {
    const int M = 16 * 2049;
    run_case(M, 2304, 768, 256, 0, true);
    run_case(M, 2304, 768, 128, 0, true);
    run_case(M, 768, 768, 256, 2, true);
    run_case(M, 768, 768, 128, 2, true);
    run_case(M, 768, 768, 192, 2, true);
    run_case(M, 3072, 768, 256, 1, true);
    run_case(M, 768, 3072, 256, 2, true);
    run_case(M, 768, 3072, 128, 2, true);
    run_case(M, 768, 3072, 192, 2, true);
  }
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
