// TMEM read/write throughput on sm_100a: how many cycles does a tcgen05.ld.32x32b.x32 (4 KB per warp) cost when
// 1, 4 (one per sub-partition) or 8 (two per sub-partition) warps issue them back to back?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I denseclip_vit_multimodal_b200/csrc -o /tmp/tmem_probe scripts/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "attn_tcgen05.cuh"
using namespace dclip;

// MODE 0: ld only; 1: st only; 2: ld + 32 ex2 per load (the softmax ratio: one ex2 per loaded element)
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cyc, int iters, int active_warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = lane + i;
  tmem_st_32x32b_x32(base, r);
  tmem_st_32x32b_x32(base + 32, r);
  tmem_st_32x32b_x32(base + 64, r);
  tmem_st_32x32b_x32(base + 96, r);
  tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  if (warp < active_warps || (active_warps == 4 && warp < 4) ) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (MODE == 0 || MODE == 2) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(base + c * 32, v);
          tmem_wait_ld();
          if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += ex2_approx(__uint_as_float(v[i]));
          } else {
            acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
          }
        } else {
          r[0] = it;
          tmem_st_32x32b_x32(base + c * 32, r);
        }
      }
      if (MODE == 1) tmem_wait_st();
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE>
void run(const char* name, int active) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8 * 8);
  const int iters = 2048;
  probe<MODE><<<148, 256>>>(out, cyc, iters, active);
  probe<MODE><<<148, 256>>>(out, cyc, iters, active);
  cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s warps=%d  %7.1f cycles per x32 op per warp (4 KB)  -> %6.1f B/cycle/SM\n", name, active, (double)h[0] / (iters * 4),
         4096.0 * active / ((double)h[0] / (iters * 4)));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int a : {1, 4, 8}) run<0>("tcgen05.ld x32", a);
  for (int a : {1, 4, 8}) run<1>("tcgen05.st x32", a);
  for (int a : {1, 4, 8}) run<2>("tcgen05.ld x32 + 32 ex2", a);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
