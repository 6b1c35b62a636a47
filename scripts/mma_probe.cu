// tcgen05.mma issue / execution timing on sm_100a (one CTA): how long does a group of K-step MMAs (kind::f16, M = 128,
// K = 16 each) take as a function of N, of the operand source (SS: A in smem, TS: A in TMEM), of whether consecutive MMAs
// accumulate into the SAME TMEM accumulator (dependent chain) or rotate over several, and of how many warps issue at once?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I denseclip_vit_multimodal_b200/csrc -o build/mma_probe scripts/mma_probe.cu
// Prints, per configuration, cycles per MMA measured (a) over the ISSUE loop alone and (b) until the commit barrier fires.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace dclip;

struct Cfg {
  int N;          // MMA N
  int ts;         // 1: A operand from TMEM
  int n_acc;      // accumulators rotated over (1 = fully dependent chain)
  int group;      // consecutive MMAs on one accumulator before rotating
  int issuers;    // warps issuing concurrently (each with its own accumulators)
  int total;      // MMAs per issuer
};

__global__ void __launch_bounds__(256, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp < c.issuers) {
    // A tile 128 x 64 bf16 (16 KB) at 0 + warp * 32 KB, B tile 256 x 64 bf16 (32 KB max) shared at 64 KB.. (zeros)
    const uint64_t dA = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t dB = make_smem_desc_sw128(smem_u32(smem + 65536), 16, 1024);
    const uint32_t idesc = make_idesc_bf16(128, c.N);
    // accumulators: issuer w owns TMEM columns [w * (512 / issuers), ...): n_acc accumulators of N columns, and (TS) an A
    // operand region of 32 columns behind them
    const int span = 512 / c.issuers;
    const uint32_t d0 = tb + warp * span;
    const uint32_t ta = d0 + c.n_acc * c.N;   // (may alias when it does not fit; timing only)
    __syncwarp();
    long long t0 = clock64();
    if (elect_one_sync()) {
      int acc = 0, in_group = 0;
      for (int i = 0; i < c.total; ++i) {
        const uint32_t d = d0 + (acc * c.N) % (span - 32 > 0 ? span : 512);
        const int k = i & 3;
        if (c.ts) umma_ts_f16(d, (ta & 0xffff01ff) + k * 8, dB + k * 2, idesc, 1u);
        else umma_ss_f16(d, dA + k * 2, dB + k * 2, idesc, 1u);
        if (++in_group == c.group) { in_group = 0; if (++acc == c.n_acc) acc = 0; }
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (elect_one_sync()) umma_commit(&bars[warp]);
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    long long t2 = clock64();
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const Cfg cfgs[] = {
      // N, ts, n_acc, group, issuers, total
      {256, 0, 1, 1, 1, 64}, {128, 0, 1, 1, 1, 64}, {64, 0, 1, 1, 1, 64}, {32, 0, 1, 1, 1, 64},
      {128, 0, 2, 1, 1, 64}, {64, 0, 2, 1, 1, 64}, {64, 0, 4, 1, 1, 64}, {64, 0, 2, 4, 1, 64}, {64, 0, 4, 4, 1, 64},
      {64, 1, 1, 1, 1, 64}, {64, 1, 2, 1, 1, 64}, {64, 1, 4, 1, 1, 64}, {64, 1, 2, 4, 1, 64}, {128, 1, 1, 1, 1, 64},
      {64, 0, 1, 1, 2, 64}, {64, 0, 1, 1, 4, 64}, {64, 1, 1, 1, 4, 64}, {128, 0, 1, 1, 2, 64}, {128, 0, 1, 1, 4, 64},
      {64, 0, 1, 1, 1, 4}, {64, 0, 1, 1, 1, 8}, {64, 1, 1, 1, 1, 4}, {128, 0, 1, 1, 1, 4}, {64, 0, 1, 1, 4, 4}, {64, 1, 1, 1, 4, 4},
  };
  printf("%4s %3s %5s %5s %7s %5s | %12s %12s   (cycles per MMA; issuer 0)\n", "N", "ts", "n_acc", "group", "issuers", "total", "issue only", "to commit");
  for (const Cfg& c : cfgs) {
    long long best[2] = {1LL << 60, 1LL << 60};
    for (int rep = 0; rep < 5; ++rep) {
      probe<<<1, 256, 100 * 1024>>>(c, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      for (int k = 0; k < 2; ++k) if (h[k] < best[k]) best[k] = h[k];
    }
    printf("%4d %3d %5d %5d %7d %5d | %12.1f %12.1f   (total %lld / %lld cycles)\n", c.N, c.ts, c.n_acc, c.group, c.issuers, c.total,
           double(best[0]) / c.total, double(best[1]) / c.total, best[0], best[1]);
  }
  return 0;
}
