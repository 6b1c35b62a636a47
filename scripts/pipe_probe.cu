// Which issue pipes do the softmax inner-loop instructions share on sm_100a?
// Runs 2 warps per SM sub-partition (8 warps/CTA, 1 CTA/SM) of pure register work and reports cycles per warp-instruction
// per sub-partition for: ex2.approx alone, cvt.rn.bf16x2.f32 alone, both interleaved 2:1 (the flash-attention ratio),
// ex2 + an integer round-and-pack (add 0x8000 + prmt), and ex2 + fma.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe scripts/pipe_probe.cu && /tmp/pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t cvt2(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_int(float lo, float hi) {
  uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u, r;
  asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(float* out, long long* cyc, int iters, float seed) {
  float x[16];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) {  // ex2 only
        x[i] = ex2(x[i]);
        x[i + 1] = ex2(x[i + 1]);
      } else if (MODE == 1) {  // cvt only (2 per pair so the instruction count matches mode 0)
        acc += cvt2(x[i], x[i + 1]);
        acc ^= cvt2(x[i + 1], x[i]);
        x[i] += 1.0f;
      } else if (MODE == 2) {  // 2 ex2 + 1 cvt
        x[i] = ex2(x[i]);
        x[i + 1] = ex2(x[i + 1]);
        acc ^= cvt2(x[i], x[i + 1]);
      } else if (MODE == 3) {  // 2 ex2 + integer round/pack
        x[i] = ex2(x[i]);
        x[i + 1] = ex2(x[i + 1]);
        acc ^= pack_int(x[i], x[i + 1]);
      } else if (MODE == 4) {  // 2 ex2 + 2 fma
        x[i] = ex2(x[i]);
        x[i + 1] = ex2(x[i + 1]);
        x[i] = fmaf(x[i], 0.5f, seed);
        x[i + 1] = fmaf(x[i + 1], 0.5f, seed);
      } else if (MODE == 5) {  // 2 ex2 + 2 fma + 1 cvt + 1 packed add (the full attention inner loop)
        float a = fmaf(x[i], 0.5f, seed), b = fmaf(x[i + 1], 0.5f, seed);
        a = ex2(a);
        b = ex2(b);
        acc ^= cvt2(a, b);
        x[i] += a;
        x[i + 1] += b;
      } else if (MODE == 7) {  // packed: fma.f32x2 + 2 ex2 + cvt + add.f32x2
        const unsigned long long h2 = pk2(0.5f, 0.5f), s2 = pk2(seed, seed);
        float a, b;
        upk2(fma2(pk2(x[i], x[i + 1]), h2, s2), a, b);
        a = ex2(a);
        b = ex2(b);
        acc ^= cvt2(a, b);
        upk2(add2(pk2(x[i], x[i + 1]), pk2(a, b)), x[i], x[i + 1]);
      } else if (MODE == 8) {  // fma.f32x2 only (x4, independent)
        const unsigned long long h2 = pk2(0.5f, 0.5f), s2 = pk2(seed, seed);
        upk2(fma2(pk2(x[i], x[i + 1]), h2, s2), x[i], x[i + 1]);
        upk2(fma2(pk2(x[i], x[i + 1]), h2, s2), x[i], x[i + 1]);
      } else if (MODE == 9) {  // ex2.approx.ftz.bf16x2: two MUFU.EX2.BF16 + PRMT per instruction -- is the half-precision SFU op any faster?
        unsigned int v = __float_as_uint(x[i]), w;
        asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(w) : "r"(v));
        x[i] = __uint_as_float(w & 0xbfffbfffu);
        unsigned int v2 = __float_as_uint(x[i + 1]), w2;
        asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(w2) : "r"(v2));
        x[i + 1] = __uint_as_float(w2 & 0xbfffbfffu);
      } else if (MODE == 6) {  // as 5 with the integer pack
        float a = fmaf(x[i], 0.5f, seed), b = fmaf(x[i + 1], 0.5f, seed);
        a = ex2(a);
        b = ex2(b);
        acc ^= pack_int(a, b);
        x[i] += a;
        x[i + 1] += b;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ex2_per_iter, int threads = 256) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  probe<MODE><<<148, threads>>>(out, cyc, iters, -0.5f);
  probe<MODE><<<148, threads>>>(out, cyc, iters, -0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  // per sub-partition: 2 warps, each iters*8 pairs
  double pairs = (threads / 128.0) * iters * 8;
  printf("[%d warp/SMSP] %-44s %8.2f cycles per (pair of elements) per sub-partition", threads / 128, name, avg / pairs);
  if (ex2_per_iter) printf("  -> %.2f cycles per ex2 warp-instruction", avg / pairs / 2);
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2 x2", 1);
  run<1>("cvt.rn.bf16x2 x2", 0);
  run<2>("ex2 x2 + cvt.bf16x2", 1);
  run<3>("ex2 x2 + int round/pack (2 iadd + prmt)", 1);
  run<4>("ex2 x2 + fma x2", 1);
  run<5>("fma x2 + ex2 x2 + cvt + add x2", 1);
  run<6>("fma x2 + ex2 x2 + int pack + add x2", 1);
  run<9>("ex2.bf16x2 x2 (= 4 MUFU.EX2.BF16)", 0);
  run<0>("ex2 x2", 1, 128);
  run<2>("ex2 x2 + cvt.bf16x2", 1, 128);
  run<4>("ex2 x2 + fma x2", 1, 128);
  run<5>("fma x2 + ex2 x2 + cvt + add x2", 1, 128);
  run<7>("fma.f32x2 + ex2 x2 + cvt + add.f32x2", 1, 128);
  run<7>("fma.f32x2 + ex2 x2 + cvt + add.f32x2", 1, 256);
  run<8>("fma.f32x2 x2 (dependent pair)", 0, 128);
  run<8>("fma.f32x2 x2 (dependent pair)", 0, 256);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
