// Standalone GPU sanity/timing binary for the tcgen05 flash-attention kernel + the few-query kernel.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_utils.cuh"

using namespace dclip;

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t x = uint32_t(i) * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  float u = (x & 0xffffff) / float(0x1000000) - 0.5f;
  p[i] = __float2bfloat16(u * scale);
}

// one thread per (b, h, query): plain fp32 softmax attention
__global__ void ref_attn_kernel(const __nv_bfloat16* qkv, float* out, int B, int H, int N, float scale) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H * N) return;
  int qi = idx % N, h = (idx / N) % H, b = idx / (N * H);
  int D = H * 64, ld = 3 * D;
  const __nv_bfloat16* q = qkv + (size_t(b) * N + qi) * ld + h * 64;
  float qf[64];
  for (int d = 0; d < 64; ++d) qf[d] = __bfloat162float(q[d]);
  float m = -INFINITY, l = 0.f, o[64];
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
  for (int j = 0; j < N; ++j) {
    const __nv_bfloat16* k = qkv + (size_t(b) * N + j) * ld + D + h * 64;
    const __nv_bfloat16* v = qkv + (size_t(b) * N + j) * ld + 2 * D + h * 64;
    float s = 0.f;
    for (int d = 0; d < 64; ++d) s += qf[d] * __bfloat162float(k[d]);
    s *= scale;
    float mn = fmaxf(m, s), a = expf(m - mn), pj = expf(s - mn);
    for (int d = 0; d < 64; ++d) o[d] = o[d] * a + pj * __bfloat162float(v[d]);
    l = l * a + pj;
    m = mn;
  }
  for (int d = 0; d < 64; ++d) out[(size_t(b) * N + qi) * D + h * 64 + d] = o[d] / l;
}

static int g_fail = 0;

static void run_case(int B, int H, int N, float amp, bool use_cls_split, bool timing) {
  const int D = H * 64, ld = 3 * D;
  __nv_bfloat16 *qkv, *out;
  float* ref;
  cudaMalloc(&qkv, size_t(B) * N * ld * 2);
  cudaMalloc(&out, size_t(B) * N * D * 2);
  cudaMalloc(&ref, size_t(B) * N * D * 4);
  fill_bf16<<<(size_t(B) * N * ld + 255) / 256, 256>>>(qkv, size_t(B) * N * ld, 11u, amp);
  cudaMemset(out, 0xff, size_t(B) * N * D * 2);
  const float scale = 0.125f;
  AttnOperands op{qkv, qkv, qkv, ld, ld, ld, (long long)N * ld, (long long)N * ld, (long long)N * ld, N};
  AttnParams p{};
  p.B = B; p.H = H; p.Nq_total = N; p.q_start = use_cls_split ? 1 : 0; p.Nk = N;
  p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out; p.out_batch_stride = (long long)N * D; p.ldo = D;
  SmallAttnParams sp{};
  sp.q = sp.k = sp.v = qkv; sp.is_f32 = 0; sp.B = B; sp.H = H; sp.Nk = N; sp.q_first = 0; sp.q_count = 1;
  sp.ldq = sp.ldk = sp.ldv = ld; sp.q_bs = sp.k_bs = sp.v_bs = (long long)N * ld;
  sp.q_col0 = 0; sp.k_col0 = D; sp.v_col0 = 2 * D; sp.scale = scale; sp.causal = 0;
  sp.out = out; sp.out_f32 = 0; sp.ldo = D; sp.out_bs = (long long)N * D;
  AttnPlan plan;
  try {
    plan = make_attn_plan(op, p);
    run_attn(plan, 0);
    if (use_cls_split) run_attn_small(sp, 0);
  } catch (Error& e) {
    printf("FAIL launch: %s\n", e.msg.c_str());
    g_fail++;
    return;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL B=%d H=%d N=%d: CUDA error %s\n", B, H, N, cudaGetErrorString(e));
    exit(2);
  }
  ref_attn_kernel<<<(B * H * N + 63) / 64, 64>>>(qkv, ref, B, H, N, scale);
  cudaDeviceSynchronize();
  std::vector<float> h_ref(size_t(B) * N * D);
  std::vector<__nv_bfloat16> h_out(size_t(B) * N * D);
  cudaMemcpy(h_ref.data(), ref, h_ref.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_out.data(), out, h_out.size() * 2, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0;
  size_t bad = 0, first_bad = size_t(-1);
  for (size_t i = 0; i < h_ref.size(); ++i) {
    float got = __bfloat162float(h_out[i]), r = h_ref[i];
    double err = fabs(double(got) - r);
    if (!(err <= fabs(r) * 0.02 + 2e-3 * amp)) { bad++; if (first_bad == size_t(-1)) first_bad = i; }
    if (err > max_err) max_err = err;
    if (fabs(r) > max_ref) max_ref = fabs(r);
  }
  printf("%s attn B=%d H=%d N=%d amp=%.1f cls_split=%d max_abs_err=%.3e max_ref=%.3e bad=%zu", bad ? "FAIL" : "ok  ", B, H, N, amp,
         int(use_cls_split), max_err, max_ref, bad);
  if (bad) {
    g_fail++;
    size_t t = first_bad / D;
    printf(" first_bad=(b=%zu,row=%zu,col=%zu got=%f ref=%f)", t / N, t % N, first_bad % D, __bfloat162float(h_out[first_bad]), h_ref[first_bad]);
  }
  if (timing && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) { run_attn(plan, 0); if (use_cls_split) run_attn_small(sp, 0); }
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; ++i) { run_attn(plan, 0); if (use_cls_split) run_attn_small(sp, 0); }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("  time=%.3f ms  %.1f TFLOP/s", ms, 4.0 * B * H * double(N) * N * 64 / ms * 1e-9);
  }
  printf("\n");
  fflush(stdout);
  cudaFree(qkv); cudaFree(out); cudaFree(ref);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no CUDA device\n"); return 3; }
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (argc > 4 && !strcmp(argv[1], "timeline")) {
    const int B = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[4]);
    const int D = H * 64, ld = 3 * D;
    __nv_bfloat16 *qkv, *out;
    long long* dbg;
    cudaMalloc(&qkv, size_t(B) * N * ld * 2); cudaMalloc(&out, size_t(B) * N * D * 2); cudaMalloc(&dbg, 1024 * 8);
    fill_bf16<<<(size_t(B) * N * ld + 255) / 256, 256>>>(qkv, size_t(B) * N * ld, 11u, 2.0f);
    AttnOperands op{qkv, qkv, qkv, ld, ld, ld, (long long)N * ld, (long long)N * ld, (long long)N * ld, N};
    AttnParams p{};
    p.B = B; p.H = H; p.Nq_total = N; p.q_start = getenv("DCLIP_TL_QSTART") ? atoi(getenv("DCLIP_TL_QSTART")) : 1; p.Nk = N; p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D;
    p.scale_log2 = 0.125f * 1.4426950408889634f; p.out = out; p.out_batch_stride = (long long)N * D; p.ldo = D;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(dbg, 0, 1024 * 8);
      p.dbg = dbg; p.dbg_cta = rep == 0 ? 0 : (getenv("DCLIP_TL_CTA") ? atoi(getenv("DCLIP_TL_CTA")) : 700);
      AttnPlan plan = make_attn_plan(op, p);
      run_attn(plan, 0);
      cudaDeviceSynchronize();
      std::vector<long long> h(1024);
      cudaMemcpy(h.data(), dbg, 1024 * 8, cudaMemcpyDeviceToHost);
      long long t0 = h[6];
      printf("timeline CTA %d (cycles rel. to first wait)  cols: wait_start s_full_got ld_done max_done exps_done o_done_got arrived | mma: pv_issue qk_issue\n", p.dbg_cta);
      const int nrows = getenv("DCLIP_TL_ROWS") ? atoi(getenv("DCLIP_TL_ROWS")) : 17;
      for (int j = 0; j < nrows; ++j)
        for (int i = 0; i < 2; ++i) {
          const long long* d = &h[(i * 32 + j) * 8];
          printf("wg%d j=%2d  %7lld %7lld %7lld %7lld %7lld %7lld %7lld | %7lld %7lld\n", i, j, d[6] - t0, d[0] - t0, d[1] - t0, d[2] - t0,
                 d[3] - t0, d[4] ? d[4] - t0 : 0, d[5] - t0, h[512 + (i * 32 + j) * 2] - t0, h[512 + (i * 32 + j) * 2 + 1] ? h[512 + (i * 32 + j) * 2 + 1] - t0 : 0);
        }
    }
    return 0;
  }
  if (argc > 4 && !strcmp(argv[1], "timeline4")) {  // P4 kernel: per-warpgroup sub-tile stamps of one CTA
    const int B = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[4]);
    const int D = H * 64, ld = 3 * D;
    __nv_bfloat16 *qkv, *out;
    long long* dbg;
    cudaMalloc(&qkv, size_t(B) * N * ld * 2); cudaMalloc(&out, size_t(B) * N * D * 2); cudaMalloc(&dbg, 2048 * 8);
    fill_bf16<<<(size_t(B) * N * ld + 255) / 256, 256>>>(qkv, size_t(B) * N * ld, 11u, 2.0f);
    AttnOperands op{qkv, qkv, qkv, ld, ld, ld, (long long)N * ld, (long long)N * ld, (long long)N * ld, N};
    AttnParams p{};
    p.B = B; p.H = H; p.Nq_total = N; p.q_start = 0; p.Nk = N; p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D;
    p.scale_log2 = 0.125f * 1.4426950408889634f; p.out = out; p.out_batch_stride = (long long)N * D; p.ldo = D;
    cudaMemset(dbg, 0, 2048 * 8);
    p.dbg = dbg; p.dbg_cta = getenv("DCLIP_TL_CTA") ? atoi(getenv("DCLIP_TL_CTA")) : 70;
    AttnPlan plan = make_attn_plan(op, p);
    run_attn(plan, 0);
    cudaDeviceSynchronize();
    std::vector<long long> h(2048);
    cudaMemcpy(h.data(), dbg, 2048 * 8, cudaMemcpyDeviceToHost);
    long long t0 = h[0];
    for (int w = 1; w < 4; ++w) if (h[w * 40 * 8] && h[w * 40 * 8] < t0) t0 = h[w * 40 * 8];
    printf("P4 timeline CTA %d (cycles rel. to first wait)  cols: wait_start s_full_got softmax_done wg_barrier_done pv_issued next_qk_issued\n", p.dbg_cta);
    const int nrows = getenv("DCLIP_TL_ROWS") ? atoi(getenv("DCLIP_TL_ROWS")) : 36;
    for (int c = 0; c < nrows; ++c)
      for (int w = 0; w < 4; ++w) {
        const long long* d = &h[(w * 40 + c) * 8];
        printf("wg%d c=%2d  %7lld %7lld %7lld %7lld %7lld %7lld   (wait %5lld softmax %5lld barrier %4lld pv_issue %4lld qk_issue %4lld | qk->s_full %5lld)\n", w, c, d[0] - t0, d[1] - t0,
               d[2] - t0, d[3] - t0, d[4] - t0, d[5] - t0, d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[3], d[5] - d[4],
               (c + 1 < 40 && d[8 + 1] && d[5]) ? d[8 + 1] - d[5] : 0);
      }
    return 0;
  }
  if (argc > 5 && !strcmp(argv[1], "qstart")) {  // time the launch restricted to query rows >= q_start (no check)
    const int B = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[4]), qs = atoi(argv[5]);
    const int D = H * 64, ld = 3 * D;
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, size_t(B) * N * ld * 2); cudaMalloc(&out, size_t(B) * N * D * 2);
    fill_bf16<<<(size_t(B) * N * ld + 255) / 256, 256>>>(qkv, size_t(B) * N * ld, 11u, 2.0f);
    AttnOperands op{qkv, qkv, qkv, ld, ld, ld, (long long)N * ld, (long long)N * ld, (long long)N * ld, N};
    AttnParams p{};
    p.B = B; p.H = H; p.Nq_total = N; p.q_start = qs; p.Nk = N; p.q_col0 = 0; p.k_col0 = D; p.v_col0 = 2 * D;
    p.scale_log2 = 0.125f * 1.4426950408889634f; p.out = out; p.out_batch_stride = (long long)N * D; p.ldo = D;
    AttnPlan plan = make_attn_plan(op, p);
    for (int i = 0; i < 3; ++i) run_attn(plan, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) run_attn(plan, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("q_start=%d grid=%d time=%.4f ms per launch (%s)\n", qs, plan.grid, ms / 20, cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (argc > 4 && !strcmp(argv[1], "prof2")) {  // production blocking (q_start = 0), timing only
    run_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), 2.0f, false, true);
    return g_fail ? 1 : 0;
  }
  if (argc > 4 && !strcmp(argv[1], "prof")) {
    run_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), 2.0f, true, true);
    return g_fail ? 1 : 0;
  }
  run_case(1, 1, 128, 2.0f, false, false);
  run_case(1, 1, 256, 2.0f, false, false);
  run_case(1, 2, 300, 2.0f, false, false);
  run_case(2, 3, 513, 2.0f, false, false);
  run_case(2, 3, 513, 2.0f, true, false);
  run_case(1, 2, 1025, 8.0f, true, false);   // large logits: exercises the lazy O rescale
  run_case(1, 4, 2049, 4.0f, true, false);
  run_case(1, 2, 2629, 2.0f, false, false);  // ViT-L/14 token count
  run_case(16, 12, 2049, 2.0f, true, true);
  run_case(16, 12, 2049, 2.0f, false, true);
  printf(g_fail ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", g_fail);
  return g_fail ? 1 : 0;
}
