// Stand-alone timing of the flash-attention exponential phase (the straight-line code of attn_tcgen05.cuh: 128 or 64
// scores per thread in registers -> FFMA2, MUFU.EX2, row-sum FADD2, bf16 pack), with 1, 2 or 4 warps per SM
// sub-partition and no other activity on the SM.  Answers: is ~11.3 cycles per MUFU a property of the compiled
// instruction stream, or of interference inside the attention kernel?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I denseclip_vit_multimodal_b200/csrc -o /tmp/exp_probe scripts/exp_phase_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace dclip;

// NEIGHBOUR workloads for the second warp of every sub-partition (threads >= 128) while warps 0-3 run the exponential phase:
// 1 = dependent FMNMX/FADD chain (ALU issue pressure), 2 = mbarrier.try_wait spin on a barrier that never completes,
// 3 = tcgen05.ld.x32 + wait::ld loop, 4 = 64 FMNMX3 then a 32-thread named barrier (bursty, like the load/max phase)
template <int NEIGHBOUR>
__device__ __forceinline__ float neighbour_work(volatile int* stop, uint64_t* bar, uint32_t taddr, int lane) {
  float a = lane * 0.001f, b = 1.0f + lane;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = lane + i;
  while (!*stop) {
    if (NEIGHBOUR == 1) {
#pragma unroll
      for (int i = 0; i < 64; ++i) { a = fmaxf(a, b + i); b = fminf(b, a - i); }
    } else if (NEIGHBOUR == 2) {
      a += mbar_try_wait(bar, 0) ? 1.f : 0.f;
    } else if (NEIGHBOUR == 3) {
      tmem_ld_32x32b_x32(taddr, v);
      tmem_wait_ld();
      a += __uint_as_float(v[0]);
    } else if (NEIGHBOUR == 5) {
      named_bar_sync(1 + (threadIdx.x >> 5) % 4, 32);
    } else if (NEIGHBOUR == 6) {  // 3-input max (FMNMX3), four independent chains, no barrier
      float m0 = a, m1 = b, m2 = a, m3 = b;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]) + b, __uint_as_float(v[i + 1]) - b));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]) + b, __uint_as_float(v[i + 3]) - b));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]) + b, __uint_as_float(v[i + 5]) - b));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]) + b, __uint_as_float(v[i + 7]) - b));
      }
      a = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      b += a;
    } else if (NEIGHBOUR == 7) {  // __syncwarp + predicated single-lane mbarrier arrive (the s_free / p_ready pattern)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    } else if (NEIGHBOUR == 4) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) a = fmaxf(a, fmaxf(__uint_as_float(v[i]) + b, __uint_as_float(v[i + 1]) - b));
      b += a;
      named_bar_sync(1 + (threadIdx.x >> 5) % 4, 32);
    }
  }
  return a + b;
}

template <int NC, int VARIANT, int NEIGHBOUR = 0>
__global__ void __launch_bounds__(256, 1) probe(const float* in, uint32_t* out, long long* cyc, int iters, float sc, int z) {
  __shared__ __align__(8) uint64_t nbar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  if (NEIGHBOUR) {
    if (threadIdx.x == 0) { mbar_init(&nbar, 1); fence_barrier_init(); stop = 0; }
    if (threadIdx.x < 32) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x >= 128) {
      const float r = neighbour_work<NEIGHBOUR>(&stop, &nbar, tslot + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16), threadIdx.x & 31);
      out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(r);
      tc_fence_before();
      __syncthreads();
      return;
    }
  }
  uint32_t su[NC];
  const float* src = in + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * NC;
#pragma unroll
  for (int e = 0; e < NC; ++e) su[e] = __float_as_uint(src[e]);
  float l = 0.f, m_used = 1.0f;
  uint32_t sink = 0;
  if (NEIGHBOUR) named_bar_sync(9, 128); else __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint64_t sc2 = pack_f32x2(sc, sc);
    const float nmc = -m_used * sc;
    const uint64_t nmc2 = pack_f32x2(nmc, nmc);
    uint64_t acc0 = pack_f32x2(0.f, 0.f), acc1 = acc0;
    if constexpr (VARIANT == 0) {
      uint32_t pk[16];
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float a, b;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(su[g * 8 + e]), __uint_as_float(su[g * 8 + e + 1])), sc2, nmc2), a, b);
          pv[e] = ex2_approx(a);
          pv[e + 1] = ex2_approx(b);
        }
        acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(pv[0], pv[1]), pack_f32x2(pv[2], pv[3])));
        acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(pv[4], pv[5]), pack_f32x2(pv[6], pv[7])));
        pk[(g & 3) * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
        pk[(g & 3) * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
        pk[(g & 3) * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
        pk[(g & 3) * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
        if ((g & 3) == 3) {
#pragma unroll
          for (int k = 0; k < 16; ++k) sink ^= pk[k];  // stands in for the tcgen05.st of 32 columns of P
        }
      }
    } else {
      // results are written to a second register array (the real kernel overwrites su in place; here su must survive)
      uint32_t ex[NC];
      auto issue = [&](int c) {
#pragma unroll
        for (int e = c * 32; e < c * 32 + 32; e += 2) {
          float a, b;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(su[e]), __uint_as_float(su[e + 1])), sc2, nmc2), a, b);
          ex[e] = __float_as_uint(ex2_approx(a));
          ex[e + 1] = __float_as_uint(ex2_approx(b));
        }
      };
      auto consume = [&](int c) {
        uint32_t pk[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) pv[k] = __uint_as_float(ex[c * 32 + g * 8 + k]);
          acc0 = add_f32x2(acc0, add_f32x2(pack_f32x2(pv[0], pv[1]), pack_f32x2(pv[2], pv[3])));
          acc1 = add_f32x2(acc1, add_f32x2(pack_f32x2(pv[4], pv[5]), pack_f32x2(pv[6], pv[7])));
          pk[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]);
          pk[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
          pk[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]);
          pk[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) sink ^= pk[k];
      };
      static_assert(VARIANT == 0 || NC == 128, "pipelined variant: 4 chunks");
#pragma unroll 1
      for (int k = z; k < 1; ++k) issue(0);
#pragma unroll 1
      for (int k = z; k < 1; ++k) { issue(1); consume(0); }
#pragma unroll 1
      for (int k = z; k < 1; ++k) { issue(2); consume(1); }
#pragma unroll 1
      for (int k = z; k < 1; ++k) { issue(3); consume(2); }
#pragma unroll 1
      for (int k = z; k < 1; ++k) consume(3);
    }
    float a0, a1, a2, a3;
    unpack_f32x2(acc0, a0, a1);
    unpack_f32x2(acc1, a2, a3);
    l += (a0 + a1) + (a2 + a3);
    m_used += 1e-6f * l;  // loop-carried so iterations cannot be merged
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink + __float_as_uint(l);
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
  if (NEIGHBOUR) {
    named_bar_sync(9, 128);
    if (threadIdx.x == 0) stop = 1;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tslot, 512); }
  }
}

template <int NC, int VARIANT, int NEIGHBOUR = 0>
void run(int warps) {
  const int threads = warps * 32, iters = 2000;
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, (size_t)148 * 256 * NC * 4);
  cudaMemset(in, 0, (size_t)148 * 256 * NC * 4);
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 16 * 8);
  probe<NC, VARIANT, NEIGHBOUR><<<148, NEIGHBOUR ? 256 : threads>>>(in, out, cyc, iters, 0.18f, 0);
  probe<NC, VARIANT, NEIGHBOUR><<<148, NEIGHBOUR ? 256 : threads>>>(in, out, cyc, iters, 0.18f, 0);
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_tile = (double)h[0] / iters;
  const double mufu_per_smsp = (double)NC * (warps / 4.0);  // MUFU warp-instructions per sub-partition per iteration
  printf("variant %d neighbour %d NC=%3d  %d warp(s)/sub-partition: %7.1f cycles per tile-row pass  -> %5.2f cycles per MUFU instr per sub-partition\n", VARIANT, NEIGHBOUR, NC,
         warps / 4, per_tile, per_tile / mufu_per_smsp);
  cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<128, 0>(4); run<128, 0>(8);
  run<64, 0>(4); run<64, 0>(8);
  run<128, 1>(4); run<128, 1>(8);
  run<128, 0, 1>(4); run<128, 0, 2>(4); run<128, 0, 3>(4); run<128, 0, 4>(4); run<128, 0, 5>(4); run<128, 0, 6>(4); run<128, 0, 7>(4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
