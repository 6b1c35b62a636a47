// How long do the "control" operations of the softmax warps take when the other warp of the same SM sub-partition is
// streaming MUFU.EX2?  Warps 0-3 (one per sub-partition) optionally run a MUFU-saturating loop; warps 4-7 time a chain of
// dependent operations of one kind: mbarrier.try_wait on a completed phase, tcgen05.ld.x32 + wait::ld, a 32-thread named
// barrier, shared-memory store+load, vote.  Prints cycles per operation without / with the MUFU stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I denseclip_vit_multimodal_b200/csrc -o /tmp/mio_probe scripts/mio_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "attn_tcgen05.cuh"
using namespace dclip;

template <int op>
__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cyc, int iters, int mufu_on) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ float sm[256];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&bar);  // phase 0 complete
  __syncthreads();
  float acc = 0.f;
  if (warp < 4) {
    if (mufu_on) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = -0.5f + 0.001f * (lane + i);
      while (!stop) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a = fmaf(x[i], 0.5f, -0.25f), b = fmaf(x[i + 1], 0.5f, -0.25f);
            a = ex2_approx(a);
            b = ex2_approx(b);
            acc += __uint_as_float(pack_bf16x2(a, b));
            x[i] += a * 1e-3f;
            x[i + 1] += b * 1e-3f;
          }
        }
      }
    }
  } else {
    const uint32_t taddr = slot + ((uint32_t)((warp & 3) * 32) << 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (op == 0) {
        mbar_wait(&bar, 0);
      } else if (op == 1) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_wait_ld();
        acc += __uint_as_float(v[it & 31]);
      } else if (op == 2) {
        named_bar_sync(1 + (warp & 3), 32);
      } else if (op == 3) {
        sm[threadIdx.x] = acc + it;
        __syncwarp();
        acc += sm[threadIdx.x ^ 1];
        __syncwarp();
      } else if (op == 4) {
        acc += __any_sync(0xffffffffu, acc > (float)it) ? 1.f : 0.f;
      } else if (op == 7) {  // one lane waits, the warp reconverges
        if (lane == 0) mbar_wait(&bar, 0);
        __syncwarp();
      } else if (op == 8) {  // test_wait (non-blocking probe), all lanes
        uint32_t ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!ok);
      } else if (op == 9) {  // elect one + test_wait
        if (elect_one_sync()) {
          uint32_t ok;
          do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
          } while (!ok);
        }
        __syncwarp();
      } else if (op == 6) {
        acc += 1.0f;
        asm volatile("" : "+f"(acc));
      } else if (op == 5) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = it + i;
        tmem_st_32x32b_x16(taddr, v);
        tmem_wait_st();
      }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 4 + (warp & 3)] = t1 - t0;
    __syncwarp();
    if (warp == 4 && lane == 0) {}
  }
  // the timing warps finish first, then release the MUFU warps
  if (warp >= 4) {
    named_bar_sync(8, 128);
    if (threadIdx.x == 128) stop = 1;
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 4 * 8);
  const char* names[] = {"mbarrier.try_wait (completed phase)", "tcgen05.ld.x32 + wait::ld", "bar.sync (32 threads)", "st.shared + ld.shared",
                         "vote.any", "tcgen05.st.x16 + wait::st", "(empty loop: fadd)", "try_wait by lane 0 + syncwarp", "test_wait, all lanes", "test_wait by one elected lane + syncwarp"};
  const int iters = 2000;
  for (int op = 0; op < 10; ++op) {
    if (op >= 1 && op <= 5) continue;
    double r[2];
    for (int on = 0; on < 2; ++on) {
      switch (op) {
        case 0: probe<0><<<148, 256>>>(out, cyc, iters, on); break;
        case 1: probe<1><<<148, 256>>>(out, cyc, iters, on); break;
        case 2: probe<2><<<148, 256>>>(out, cyc, iters, on); break;
        case 3: probe<3><<<148, 256>>>(out, cyc, iters, on); break;
        case 4: probe<4><<<148, 256>>>(out, cyc, iters, on); break;
        case 5: probe<5><<<148, 256>>>(out, cyc, iters, on); break;
        case 7: probe<7><<<148, 256>>>(out, cyc, iters, on); break;
        case 8: probe<8><<<148, 256>>>(out, cyc, iters, on); break;
        case 9: probe<9><<<148, 256>>>(out, cyc, iters, on); break;
        default: probe<6><<<148, 256>>>(out, cyc, iters, on); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[4];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      r[on] = (double)h[0] / iters;
    }
    printf("%-38s %7.1f cycles alone   %7.1f cycles next to a MUFU-streaming warp\n", names[op], r[0], r[1]);
  }
  return 0;
}
