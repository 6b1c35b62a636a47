// Bandwidth-bound kernels of the DenseCLIP forward path: LayerNorm, patch im2col, positional-embedding bilinear
// interpolation, CLS row, feature-tap transpose, token mean, L2-normalise + pixel-text score map, bilinear upsample.
// All are coalesced / 128-bit vectorised with warp-shuffle reductions; fp32 math throughout.
#pragma once
#include "ptx.cuh"
#include "host_utils.cuh"

namespace dclip {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (models.py:243-249: nn.LayerNorm, eps 1e-5, fp32 statistics, two-pass variance).
// One warp per row, the row lives in registers.  D % 128 == 0, D <= 1024.
// Outputs (any subset): fp32 [M, ldo], bf16 [M, ldb] (+ lo half at column split_off when split != 0).
// ---------------------------------------------------------------------------------------------------------
struct LayerNormParams {
  const float* x; long long ldx;
  const float* gamma; const float* beta;
  float eps;
  int M, D;
  float* out_f32; long long ldo;
  __nv_bfloat16* out_bf16; long long ldb;
  int split; int split_off;
  // optional (LayerNorm folding of the ViT blocks): per-row (sum, sum of squares) of the OUTPUT in slot 0 of the slot-major
  // stats_out[slot * stats_ld + row], slots 1 .. stats_n - 1 zeroed (the layout the residual GEMM epilogues fill per 32-column chunk)
  float2* stats_out = nullptr; int stats_ld = 0; int stats_n = 0;
};

template <int VEC>  // VEC = D / 128 float4 chunks per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const LayerNormParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.M) return;
  const float4* xr = reinterpret_cast<const float4*>(p.x + (long long)row * p.ldx);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / float(p.D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) / float(p.D) + p.eps);
  float o1 = 0.f, o2 = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c0 = 4 * (lane + 32 * i);
    const float4 g = *reinterpret_cast<const float4*>(p.gamma + c0);
    const float4 b = *reinterpret_cast<const float4*>(p.beta + c0);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    o1 += (y.x + y.y) + (y.z + y.w);
    o2 += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
    if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (long long)row * p.ldo + c0) = y;
    if (p.out_bf16) {
      __nv_bfloat16* ob = p.out_bf16 + (long long)row * p.ldb + c0;
      const uint32_t h0 = pack_bf16x2(y.x, y.y), h1 = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(ob) = make_uint2(h0, h1);
      if (p.split) {
        const uint32_t l0 = pack_bf16x2(y.x - __uint_as_float(h0 << 16), y.y - __uint_as_float(h0 & 0xffff0000u));
        const uint32_t l1 = pack_bf16x2(y.z - __uint_as_float(h1 << 16), y.w - __uint_as_float(h1 & 0xffff0000u));
        *reinterpret_cast<uint2*>(ob + p.split_off) = make_uint2(l0, l1);
      }
    }
  }
  if (p.stats_out) {
    o1 = warp_sum(o1);
    o2 = warp_sum(o2);
    for (int j = lane; j < p.stats_n; j += 32) p.stats_out[(long long)j * p.stats_ld + row] = j == 0 ? make_float2(o1, o2) : make_float2(0.f, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (optionally hi|lo split), row-wise with leading dims.  Used to pack weights and activations.
// ---------------------------------------------------------------------------------------------------------
struct CastParams {
  const float* x; long long ldx;
  __nv_bfloat16* out; long long ldo;
  int rows, cols;   // cols % 2 == 0
  int split; int split_off;
  float scale;
};

__global__ void __launch_bounds__(256) cast_bf16_kernel(const CastParams p) {
  const long long half_cols = p.cols >> 1;
  const long long total = (long long)p.rows * half_cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / half_cols;
    const int c = int(i - r * half_cols) * 2;
    const float2 v = *reinterpret_cast<const float2*>(p.x + r * p.ldx + c);
    const float a = v.x * p.scale, b = v.y * p.scale;
    const uint32_t hi = pack_bf16x2(a, b);
    *reinterpret_cast<uint32_t*>(p.out + r * p.ldo + c) = hi;
    if (p.split)
      *reinterpret_cast<uint32_t*>(p.out + r * p.ldo + p.split_off + c) =
          pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Patch im2col for the stride = kernel patch-embed conv (models.py:407,546-548): img fp32 [B,3,H,W] ->
// A bf16 [B*gh*gw, lda] with K index = c*ps*ps + ky*ps + kx (the flattening of conv1.weight [D,3,ps,ps]).
// Pixels right/below the last full patch are dropped, as the conv does.
// ---------------------------------------------------------------------------------------------------------
struct Im2colParams {
  const float* img;
  __nv_bfloat16* out; long long lda;
  int B, H, W, ps, gh, gw;
  int split; int split_off;
};

__global__ void __launch_bounds__(256) im2col_patch_kernel(const Im2colParams p) {
  const int K = 3 * p.ps * p.ps, K2 = K >> 1;
  const long long total = (long long)p.B * p.gh * p.gw * K2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K2;
    const int k = int(i - row * K2) * 2;
    const int c = k / (p.ps * p.ps), rem = k - c * p.ps * p.ps, ky = rem / p.ps, kx = rem - ky * p.ps;
    const int b = int(row / (p.gh * p.gw)), pr = int(row - (long long)b * p.gh * p.gw), py = pr / p.gw, px = pr - py * p.gw;
    const float* src = p.img + (((long long)b * 3 + c) * p.H + (py * p.ps + ky)) * p.W + px * p.ps + kx;
    const float a = src[0], bb = src[1];  // ps is even, so (kx, kx+1) stay inside one patch row
    const uint32_t hi = pack_bf16x2(a, bb);
    *reinterpret_cast<uint32_t*>(p.out + row * p.lda + k) = hi;
    if (p.split)
      *reinterpret_cast<uint32_t*>(p.out + row * p.lda + p.split_off + k) =
          pack_bf16x2(a - __uint_as_float(hi << 16), bb - __uint_as_float(hi & 0xffff0000u));
  }
}

// Vector path (ps % 8 == 0, W % 4 == 0, 16-byte aligned image): a thread converts 8 consecutive pixels of one patch row
// (two 128-bit loads, one 128-bit store per output half) -- 4x fewer threads and index decodes than the 2-pixel kernel.
__global__ void __launch_bounds__(256) im2col_patch_vec8_kernel(const Im2colParams p) {
  const int K = 3 * p.ps * p.ps, K8 = K >> 3, ps8 = p.ps >> 3;
  const long long total = (long long)p.B * p.gh * p.gw * K8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K8;
    const int k8 = int(i - row * K8);
    const int c = k8 / (p.ps * ps8), rem = k8 - c * p.ps * ps8, ky = rem / ps8, kx = (rem - ky * ps8) * 8;
    const int b = int(row / (p.gh * p.gw)), pr = int(row - (long long)b * p.gh * p.gw), py = pr / p.gw, px = pr - py * p.gw;
    const float4* src = reinterpret_cast<const float4*>(p.img + (((long long)b * 3 + c) * p.H + (py * p.ps + ky)) * p.W + px * p.ps + kx);
    const float4 a = __ldcs(src), bb = __ldcs(src + 1);
    const uint32_t h0 = pack_bf16x2(a.x, a.y), h1 = pack_bf16x2(a.z, a.w), h2 = pack_bf16x2(bb.x, bb.y), h3 = pack_bf16x2(bb.z, bb.w);
    const int k = c * p.ps * p.ps + ky * p.ps + kx;
    *reinterpret_cast<uint4*>(p.out + row * p.lda + k) = make_uint4(h0, h1, h2, h3);
    if (p.split) {
      auto lo = [](float x, float y, uint32_t h) { return pack_bf16x2(x - __uint_as_float(h << 16), y - __uint_as_float(h & 0xffff0000u)); };
      *reinterpret_cast<uint4*>(p.out + row * p.lda + p.split_off + k) =
          make_uint4(lo(a.x, a.y, h0), lo(a.z, a.w, h1), lo(bb.x, bb.y, h2), lo(bb.z, bb.w, h3));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Positional-embedding interpolation (models.py:514-540): pos [1+g0*g0, D] -> out [1+gh*gw, D];
// row 0 copied, the g0 x g0 grid resized bilinearly with align_corners=False (ATen upsample_bilinear2d index rule:
// src = scale*(dst+0.5)-0.5 clamped at 0; i1 = i0 + (i0 < in-1)).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (float(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = int(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - float(i0);
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) posemb_interp_kernel(const float* __restrict__ pos, float* __restrict__ out, int g0,
                                                           int gh, int gw, int D) {
  const int tok = blockIdx.x;  // 0 .. gh*gw
  if (tok == 0) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = pos[d];
    return;
  }
  const int y = (tok - 1) / gw, x = (tok - 1) % gw;
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(y, float(g0) / float(gh), g0, y0, y1, ly0, ly1);
  bilinear_src(x, float(g0) / float(gw), g0, x0, x1, lx0, lx1);
  const float* p00 = pos + (long long)(1 + y0 * g0 + x0) * D;
  const float* p01 = pos + (long long)(1 + y0 * g0 + x1) * D;
  const float* p10 = pos + (long long)(1 + y1 * g0 + x0) * D;
  const float* p11 = pos + (long long)(1 + y1 * g0 + x1) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[(long long)tok * D + d] = ly0 * (lx0 * p00[d] + lx1 * p01[d]) + ly1 * (lx0 * p10[d] + lx1 * p11[d]);
}

// x[b, 0, :] = class_embedding + pos[0]   (models.py:551-556)
__global__ void cls_row_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B,
                               int Ntok, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[(long long)b * Ntok * D + d] = cls[d] + pos[d];
}

// ---------------------------------------------------------------------------------------------------------
// Feature tap (models.py:568-582): tokens fp32 [B, Ntok, D] (CLS at 0) -> NCHW fp32 [B, D, P]; 32x32 smem transpose.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tap_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int Ntok, int P,
                                                      int D) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xb = x + (long long)b * Ntok * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pp = p0 + ty + 8 * i;
    tile[ty + 8 * i][tx] = (pp < P && d0 + tx < D) ? xb[(long long)(1 + pp) * D + d0 + tx] : 0.f;
  }
  __syncthreads();
  float* ob = out + (long long)b * D * P;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = d0 + ty + 8 * i;
    if (d < D && p0 + tx < P) ob[(long long)d * P + p0 + tx] = tile[tx][ty + 8 * i];
  }
}

// NCHW fp32 [B, C, P] -> token-major [B, P(+row0 offset), ld] fp32 and/or bf16 (the inverse of the tap; used when a
// caller hands `_process_features` externally produced NCHW features).
__global__ void __launch_bounds__(256) nchw_to_tokens_kernel(const float* __restrict__ in, float* out_f32,
                                                            __nv_bfloat16* out_bf16, int C, int P, long long ld,
                                                            long long out_bs, int row_off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* ib = in + (long long)b * C * P;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    tile[ty + 8 * i][tx] = (c < C && p0 + tx < P) ? ib[(long long)c * P + p0 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pp = p0 + ty + 8 * i;
    if (pp < P && c0 + tx < C) {
      const float v = tile[tx][ty + 8 * i];
      const long long o = (long long)b * out_bs + (long long)(row_off + pp) * ld + c0 + tx;
      if (out_f32) out_f32[o] = v;
      if (out_bf16) out_bf16[o] = __float2bfloat16(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Token mean over the P patch tokens (adaptive_avg_pool2d(.,(1,1)), denseclip.py:596): x [B, rows, ld] -> [B, D].
// grid (D/32, B); each warp strides over tokens, lanes over channels.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int row0,
                                                        int P, long long ld, long long bs, int D) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane, b = blockIdx.y;
  float acc = 0.f;
  if (d < D)
    for (int t = w; t < P; t += 8) acc += x[(long long)b * bs + (long long)(row0 + t) * ld + d];
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && d < D) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    out[(long long)b * D + d] = s / float(P);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pixel-text score map (denseclip.py:672-675): score[b,k,p] = <v/max(|v|,eps), t_k/max(|t_k|,eps)>,
// vis fp32 token-major [B, P rows from row0, ld] (C channels), text fp32 [B, K, C]; eps = 1e-12 (F.normalize).
// One warp per pixel; the image's K normalised text rows are staged in smem (K*C*4 bytes, 38 KB for 19 x 512).
// Also optionally writes the un-normalised vis features back as NCHW fp32 (the reference's `visual_embeddings`).
// ---------------------------------------------------------------------------------------------------------
struct ScoreParams {
  const float* vis; long long ld; long long bs; int row0;
  const float* text;    // [B, K, C]
  float* score;         // [B, K, P]
  int B, K, C, P;
  float eps;
};

// Sum of 32 per-lane values across the warp, transposed: on return lane L holds the warp total of v[L] in v[0]
// (31 shuffles for 32 values instead of 5 per value: at step `o` the lanes with bit `o` set keep the upper half).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// PIX pixels per warp and trip: the K normalised text rows (smem) are read once per PIX pixels (the one-pixel version read
// 38 KB of smem per pixel and did 19 five-step warp reductions per pixel: 0.89 TB/s; this one is HBM-bound).
// Lane l holds channels 4 (l + 32 i) .. + 3 of each pixel (128-bit coalesced loads); C % 128 == 0.
template <int NC4, int PIX = 4>  // NC4 = C / 128 float4 chunks per lane
__global__ void __launch_bounds__(256) score_map_kernel(const ScoreParams p) {
  extern __shared__ float st[];  // [K][C] normalised text of this image
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int C4 = p.C >> 2;
  float4* st4w = reinterpret_cast<float4*>(st);
  {
    // stage the raw text rows with ONE round of coalesced 128-bit loads (all in flight together), then normalise in smem
    const float4* t4 = reinterpret_cast<const float4*>(p.text + (long long)b * p.K * p.C);
    for (int i = threadIdx.x; i < p.K * C4; i += 256) st4w[i] = __ldg(t4 + i);
    __syncthreads();
    for (int k = w; k < p.K; k += 8) {
      float ss = 0.f;
      for (int c = lane; c < C4; c += 32) {
        const float4 t = st4w[k * C4 + c];
        ss += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
      }
      const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), p.eps);
      for (int c = lane; c < C4; c += 32) {
        float4 t = st4w[k * C4 + c];
        t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
        st4w[k * C4 + c] = t;
      }
    }
  }
  __syncthreads();
  const float4* st4 = reinterpret_cast<const float4*>(st);
  for (int px0 = (blockIdx.x * 8 + w) * PIX; px0 < p.P; px0 += gridDim.x * 8 * PIX) {
    float4 vv[PIX][NC4];
    float inv[PIX];
#pragma unroll
    for (int q = 0; q < PIX; ++q) {
      const int px = min(px0 + q, p.P - 1);   // (a ragged last group recomputes the last pixel; its stores are masked)
      const float4* v = reinterpret_cast<const float4*>(p.vis + (long long)b * p.bs + (long long)(p.row0 + px) * p.ld);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NC4; ++i) {
        vv[q][i] = __ldcs(v + lane + 32 * i);
        ss += (vv[q][i].x * vv[q][i].x + vv[q][i].y * vv[q][i].y) + (vv[q][i].z * vv[q][i].z + vv[q][i].w * vv[q][i].w);
      }
      inv[q] = 1.f / fmaxf(sqrtf(warp_sum(ss)), p.eps);
    }
    // classes in rounds of 8: 8 x PIX = 32 partial dot products per lane, one transpose-reduce per round
    for (int k0 = 0; k0 < p.K; k0 += 32 / PIX) {
      float acc[32];
#pragma unroll
      for (int kk = 0; kk < 32 / PIX; ++kk) {
        const int k = min(k0 + kk, p.K - 1);
        // packed fp32x2 FMAs (two lanes per issue slot): the kernel is issue-bound on the 19 x C multiply-adds per pixel
        uint64_t a01[PIX], a23[PIX];
#pragma unroll
        for (int q = 0; q < PIX; ++q) a01[q] = a23[q] = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NC4; ++i) {
          const float4 t = st4[k * C4 + lane + 32 * i];
          const uint64_t t01 = pack_f32x2(t.x, t.y), t23 = pack_f32x2(t.z, t.w);
#pragma unroll
          for (int q = 0; q < PIX; ++q) {
            a01[q] = fma_f32x2(pack_f32x2(vv[q][i].x, vv[q][i].y), t01, a01[q]);
            a23[q] = fma_f32x2(pack_f32x2(vv[q][i].z, vv[q][i].w), t23, a23[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < PIX; ++q) {
          float s0, s1, s2, s3;
          unpack_f32x2(a01[q], s0, s1);
          unpack_f32x2(a23[q], s2, s3);
          acc[kk * PIX + q] = (s0 + s1) + (s2 + s3);
        }
      }
      const float tot = warp_transpose_reduce32(acc, lane);   // lane = kk * PIX + q
      const int k = k0 + lane / PIX, q = lane % PIX;
      float invq = inv[0];
#pragma unroll
      for (int qq = 1; qq < PIX; ++qq) invq = (q == qq) ? inv[qq] : invq;
      if (k < p.K && px0 + q < p.P) p.score[((long long)b * p.K + k) * p.P + px0 + q] = tot * invq;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Bilinear upsample, align_corners=False (denseclip.py:894-916): in fp32 token-major [B, h*w, ldi] (C channels used)
// or NCHW -> out NCHW fp32 [B, C, H, W].  One thread per 4 output pixels along W (128-bit stores).
// ---------------------------------------------------------------------------------------------------------
struct UpsampleParams {
  const float* in; int in_nchw; long long ldi; long long in_bs;
  float* out;
  int B, C, h, w, H, W;
};

__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const UpsampleParams p) {
  const int W4 = p.W >> 2;
  const long long total = (long long)p.B * p.C * p.H * W4;
  const float sy = float(p.h) / float(p.H), sx = float(p.w) / float(p.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = int(i % W4);
    const int y = int((i / W4) % p.H);
    const int c = int((i / ((long long)W4 * p.H)) % p.C);
    const int b = int(i / ((long long)W4 * p.H * p.C));
    int y0, y1;
    float ly0, ly1;
    bilinear_src(y, sy, p.h, y0, y1, ly0, ly1);
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int x0, x1;
      float lx0, lx1;
      bilinear_src(x4 * 4 + e, sx, p.w, x0, x1, lx0, lx1);
      float v00, v01, v10, v11;
      if (p.in_nchw) {
        const float* base = p.in + ((long long)b * p.C + c) * p.h * p.w;
        v00 = base[y0 * p.w + x0]; v01 = base[y0 * p.w + x1]; v10 = base[y1 * p.w + x0]; v11 = base[y1 * p.w + x1];
      } else {
        const float* base = p.in + (long long)b * p.in_bs + c;
        v00 = base[(long long)(y0 * p.w + x0) * p.ldi]; v01 = base[(long long)(y0 * p.w + x1) * p.ldi];
        v10 = base[(long long)(y1 * p.w + x0) * p.ldi]; v11 = base[(long long)(y1 * p.w + x1) * p.ldi];
      }
      r[e] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
    *reinterpret_cast<float4*>(p.out + (((long long)b * p.C + c) * p.H + y) * p.W + x4 * 4) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 3x3 / pad 1 conv operand gather (generic fallback; the aligned case uses the GEMM's implicit-conv TMA mode):
// out[b*h*w + y*w + x][(ky*3+kx)*C + c] = in[b][row0 + (y+ky-1)*w + (x+kx-1)][c]  (0 outside the image)
// ---------------------------------------------------------------------------------------------------------
struct Conv3x3GatherParams {
  const void* in; int in_f32; long long ld; long long bs; int row0;
  int B, h, w, C;
  __nv_bfloat16* out; long long ldo;
};

__global__ void __launch_bounds__(256) conv3x3_gather_kernel(const Conv3x3GatherParams p) {
  const int C2 = p.C >> 1;
  const long long total = (long long)p.B * p.h * p.w * 9 * C2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C2) * 2;
    const int tap = int((i / C2) % 9);
    const long long pix = i / (9LL * C2);
    const int x = int(pix % p.w), y = int((pix / p.w) % p.h), b = int(pix / ((long long)p.w * p.h));
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint32_t v = 0;
    if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
      const long long off = (long long)b * p.bs + (long long)(p.row0 + yy * p.w + xx) * p.ld + c;
      if (p.in_f32) {
        const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.in) + off);
        v = pack_bf16x2(f.x, f.y);
      } else {
        v = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p.in) + off);
      }
    }
    *reinterpret_cast<uint32_t*>(p.out + pix * p.ldo + tap * p.C + c) = v;
  }
}

// Token-major fast path: one thread produces 4 horizontally adjacent output pixels x UP_ROWS output rows for ALL channels.
// The horizontal interpolation of the two source rows is kept in registers and only recomputed when (y0, y1) changes
// (every 16 output rows at the x16 scale of the heads), so a pixel costs 2 FMAs + its share of a 128-bit store instead of
// 4 loads + 7 flops: the kernel went from issue-bound (3.9 TB/s) to write-bound.  Same arithmetic order as before.
constexpr int UP_ROWS = 16;
__global__ void __launch_bounds__(256) upsample_bilinear_tok_kernel(const UpsampleParams p) {
  const int W4 = p.W >> 2;
  const int nstrip = (p.H + UP_ROWS - 1) / UP_ROWS;
  const long long total = (long long)p.B * nstrip * W4;
  const float sy = float(p.h) / float(p.H), sx = float(p.w) / float(p.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = int(i % W4);
    const int strip = int((i / W4) % nstrip);
    const int b = int(i / ((long long)W4 * nstrip));
    long long xo0[4], xo1[4];
    float lx0[4], lx1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int x0, x1;
      bilinear_src(x4 * 4 + e, sx, p.w, x0, x1, lx0[e], lx1[e]);
      xo0[e] = (long long)x0 * p.ldi;
      xo1[e] = (long long)x1 * p.ldi;
    }
    const float* base = p.in + (long long)b * p.in_bs;
    float* obase = p.out + (long long)b * p.C * p.H * p.W + x4 * 4;
    const int yend = min(p.H, (strip + 1) * UP_ROWS);
    for (int c = 0; c < p.C; c += 4) {
      int py0 = -1, py1 = -1;
      float h0[4][4], h1[4][4];  // [channel][pixel]: horizontally interpolated source rows y0 / y1
      for (int y = strip * UP_ROWS; y < yend; ++y) {
        int y0, y1;
        float ly0, ly1;
        bilinear_src(y, sy, p.h, y0, y1, ly0, ly1);
        if (y0 != py0 || y1 != py1) {
          const float* ra = base + (long long)y0 * p.w * p.ldi + c;
          const float* rb = base + (long long)y1 * p.w * p.ldi + c;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 a = *reinterpret_cast<const float4*>(ra + xo0[e]), bb = *reinterpret_cast<const float4*>(ra + xo1[e]);
            const float4 cc = *reinterpret_cast<const float4*>(rb + xo0[e]), d = *reinterpret_cast<const float4*>(rb + xo1[e]);
            h0[0][e] = lx0[e] * a.x + lx1[e] * bb.x;  h1[0][e] = lx0[e] * cc.x + lx1[e] * d.x;
            h0[1][e] = lx0[e] * a.y + lx1[e] * bb.y;  h1[1][e] = lx0[e] * cc.y + lx1[e] * d.y;
            h0[2][e] = lx0[e] * a.z + lx1[e] * bb.z;  h1[2][e] = lx0[e] * cc.z + lx1[e] * d.z;
            h0[3][e] = lx0[e] * a.w + lx1[e] * bb.w;  h1[3][e] = lx0[e] * cc.w + lx1[e] * d.w;
          }
          py0 = y0;
          py1 = y1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < p.C)
            __stcs(reinterpret_cast<float4*>(obase + ((long long)(c + k) * p.H + y) * p.W),
                   make_float4(ly0 * h0[k][0] + ly1 * h1[k][0], ly0 * h0[k][1] + ly1 * h1[k][1], ly0 * h0[k][2] + ly1 * h1[k][2],
                               ly0 * h0[k][3] + ly1 * h1[k][3]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused bilinear upsample + channel argmax (simple_test's `seg_logit.argmax(dim=1)`, denseclip.py:987-1000) without
// materialising the [B, K, H, W] logits: in = token-major fp32 [B, h*w, ldi] low-res logits, out = uint8 [B, H, W].
// Ties resolve to the lowest class index, like torch.argmax.  One thread per 4 output pixels along W.
// ---------------------------------------------------------------------------------------------------------
struct UpsampleArgmaxParams {
  const float* in; long long ldi; long long in_bs;
  uint8_t* out;
  int B, K, h, w, H, W;
};

// Strip variant for K <= 4*K4 classes whose token rows hold 4*K4 readable floats: 4 pixels x UP_ROWS rows per thread,
// horizontally interpolated source rows for all classes cached in registers (recomputed when (y0, y1) changes).
template <int K4>
__global__ void __launch_bounds__(256) upsample_argmax_strip_kernel(const UpsampleArgmaxParams p) {
  const int W4 = p.W >> 2;
  const int nstrip = (p.H + UP_ROWS - 1) / UP_ROWS;
  const long long total = (long long)p.B * nstrip * W4;
  const float sy = float(p.h) / float(p.H), sx = float(p.w) / float(p.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = int(i % W4);
    const int strip = int((i / W4) % nstrip);
    const int b = int(i / ((long long)W4 * nstrip));
    long long xo0[4], xo1[4];
    float lx0[4], lx1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int x0, x1;
      bilinear_src(x4 * 4 + e, sx, p.w, x0, x1, lx0[e], lx1[e]);
      xo0[e] = (long long)x0 * p.ldi;
      xo1[e] = (long long)x1 * p.ldi;
    }
    const float* base = p.in + (long long)b * p.in_bs;
    const int yend = min(p.H, (strip + 1) * UP_ROWS);
    int py0 = -1, py1 = -1;
    float h0[K4 * 4][4], h1[K4 * 4][4];  // [class][pixel]
    for (int y = strip * UP_ROWS; y < yend; ++y) {
      int y0, y1;
      float ly0, ly1;
      bilinear_src(y, sy, p.h, y0, y1, ly0, ly1);
      if (y0 != py0 || y1 != py1) {
        const float* ra = base + (long long)y0 * p.w * p.ldi;
        const float* rb = base + (long long)y1 * p.w * p.ldi;
#pragma unroll
        for (int c = 0; c < K4; ++c) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 a = *reinterpret_cast<const float4*>(ra + xo0[e] + 4 * c), bb = *reinterpret_cast<const float4*>(ra + xo1[e] + 4 * c);
            const float4 cc = *reinterpret_cast<const float4*>(rb + xo0[e] + 4 * c), d = *reinterpret_cast<const float4*>(rb + xo1[e] + 4 * c);
            h0[4 * c + 0][e] = lx0[e] * a.x + lx1[e] * bb.x;  h1[4 * c + 0][e] = lx0[e] * cc.x + lx1[e] * d.x;
            h0[4 * c + 1][e] = lx0[e] * a.y + lx1[e] * bb.y;  h1[4 * c + 1][e] = lx0[e] * cc.y + lx1[e] * d.y;
            h0[4 * c + 2][e] = lx0[e] * a.z + lx1[e] * bb.z;  h1[4 * c + 2][e] = lx0[e] * cc.z + lx1[e] * d.z;
            h0[4 * c + 3][e] = lx0[e] * a.w + lx1[e] * bb.w;  h1[4 * c + 3][e] = lx0[e] * cc.w + lx1[e] * d.w;
          }
        }
        py0 = y0;
        py1 = y1;
      }
      uint32_t packed = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < K4 * 4; ++k) {
          const float v = ly0 * h0[k][e] + ly1 * h1[k][e];
          if (k < p.K && v > best) { best = v; arg = k; }
        }
        packed |= uint32_t(arg) << (8 * e);
      }
      *reinterpret_cast<uint32_t*>(p.out + ((long long)b * p.H + y) * p.W + x4 * 4) = packed;
    }
  }
}

__global__ void __launch_bounds__(256) upsample_argmax_kernel(const UpsampleArgmaxParams p) {
  const int W4 = p.W >> 2;
  const long long total = (long long)p.B * p.H * W4;
  const float sy = float(p.h) / float(p.H), sx = float(p.w) / float(p.W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = int(i % W4);
    const int y = int((i / W4) % p.H);
    const int b = int(i / ((long long)W4 * p.H));
    int y0, y1;
    float ly0, ly1;
    bilinear_src(y, sy, p.h, y0, y1, ly0, ly1);
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int x0, x1;
      float lx0, lx1;
      bilinear_src(x4 * 4 + e, sx, p.w, x0, x1, lx0, lx1);
      const float* r00 = p.in + (long long)b * p.in_bs + (long long)(y0 * p.w + x0) * p.ldi;
      const float* r01 = p.in + (long long)b * p.in_bs + (long long)(y0 * p.w + x1) * p.ldi;
      const float* r10 = p.in + (long long)b * p.in_bs + (long long)(y1 * p.w + x0) * p.ldi;
      const float* r11 = p.in + (long long)b * p.in_bs + (long long)(y1 * p.w + x1) * p.ldi;
      float best = -INFINITY;
      int arg = 0;
      for (int k = 0; k < p.K; ++k) {
        const float v = ly0 * (lx0 * r00[k] + lx1 * r01[k]) + ly1 * (lx0 * r10[k] + lx1 * r11[k]);
        if (v > best) { best = v; arg = k; }
      }
      packed |= uint32_t(arg) << (8 * e);
    }
    *reinterpret_cast<uint32_t*>(p.out + ((long long)b * p.H + y) * p.W + x4 * 4) = packed;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Evaluation statistics on the device (reference: torchmetrics JaccardIndex / Accuracy with ignore_index and
// MeanSquaredError(squared=False) over the masked depth pixels, train_denseclip.py:351-355, 582-593).
//   conf[t*K + p] += #pixels with target t (!= ignore_index, < K) predicted as p       (int64, accumulated)
//   depth_stats[0] += sum (pred - gt)^2 over pixels with mask != 0,  depth_stats[1] += their count   (double, accumulated)
// Block-private shared-memory histogram, one global atomic per non-empty bin per block; integer sums are exact and
// order-independent, so shards can be reduced with one NCCL all-reduce.
// ---------------------------------------------------------------------------------------------------------
struct EvalStatsParams {
  const uint8_t* pred; const void* target; int target_i64; long long n; int K; int ignore_index;
  const float* depth_pred; const float* depth_gt; const uint8_t* depth_mask; long long n_depth;
  unsigned long long* conf; double* depth_stats;
};

__global__ void __launch_bounds__(256) eval_stats_kernel(const EvalStatsParams p) {
  extern __shared__ unsigned int hist[];  // [K*K]
  const int KK = p.K * p.K;
  for (int i = threadIdx.x; i < KK; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x, first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p.pred) {
    for (long long i = first; i < p.n; i += stride) {
      const long long t = p.target_i64 ? reinterpret_cast<const long long*>(p.target)[i] : (long long)reinterpret_cast<const uint8_t*>(p.target)[i];
      const int pr = p.pred[i];
      if (t != p.ignore_index && t >= 0 && t < p.K && pr < p.K) atomicAdd(&hist[int(t) * p.K + pr], 1u);
    }
  }
  double se = 0.0, cnt = 0.0;
  if (p.depth_pred) {
    for (long long i = first; i < p.n_depth; i += stride) {
      if (!p.depth_mask || p.depth_mask[i]) {
        const double d = double(p.depth_pred[i]) - double(p.depth_gt[i]);
        se += d * d;
        cnt += 1.0;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KK; i += blockDim.x)
    if (hist[i]) atomicAdd(&p.conf[i], (unsigned long long)hist[i]);
  if (p.depth_pred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
      atomicAdd(&p.depth_stats[0], se);
      atomicAdd(&p.depth_stats[1], cnt);
    }
  }
}

// out[i] = a[i] + gamma[i % C] * d[i]     (text + gamma * text_diff, denseclip.py:665)
__global__ void gamma_residual_kernel(const float* __restrict__ a, const float* __restrict__ g, const float* __restrict__ d,
                                      float* __restrict__ out, long long n, int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + g[i % C] * d[i];
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
inline void launch_layernorm(const LayerNormParams& p, cudaStream_t stream) {
  if (p.M <= 0) return;
  const int grid = (p.M + 7) / 8;
  switch (p.D) {
    case 128: layernorm_kernel<1><<<grid, 256, 0, stream>>>(p); break;
    case 256: layernorm_kernel<2><<<grid, 256, 0, stream>>>(p); break;
    case 384: layernorm_kernel<3><<<grid, 256, 0, stream>>>(p); break;
    case 512: layernorm_kernel<4><<<grid, 256, 0, stream>>>(p); break;
    case 640: layernorm_kernel<5><<<grid, 256, 0, stream>>>(p); break;
    case 768: layernorm_kernel<6><<<grid, 256, 0, stream>>>(p); break;
    case 896: layernorm_kernel<7><<<grid, 256, 0, stream>>>(p); break;
    case 1024: layernorm_kernel<8><<<grid, 256, 0, stream>>>(p); break;
    default: throw Error{"layernorm: D must be a multiple of 128 and <= 1024"};
  }
}

inline void launch_im2col(const Im2colParams& p, cudaStream_t stream) {
  if (p.ps % 8 == 0 && p.W % 4 == 0 && p.lda % 8 == 0 && p.split_off % 8 == 0 && (reinterpret_cast<uintptr_t>(p.img) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) {
    const long long total8 = (long long)p.B * p.gh * p.gw * (3 * p.ps * p.ps / 8);
    const int grid8 = int(total8 / 256 < 148 * 16 ? (total8 + 255) / 256 : 148 * 16);
    im2col_patch_vec8_kernel<<<grid8 > 0 ? grid8 : 1, 256, 0, stream>>>(p);
    return;
  }
  const long long total = (long long)p.B * p.gh * p.gw * (3 * p.ps * p.ps / 2);
  const int grid = int(total / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  im2col_patch_kernel<<<grid > 0 ? grid : 1, 256, 0, stream>>>(p);
}

inline void launch_tap_nchw(const float* tokens, float* out, int B, int Ntok, int D, cudaStream_t stream) {
  const int P = Ntok - 1;
  dim3 grid((P + 31) / 32, (D + 31) / 32, B);
  tap_nchw_kernel<<<grid, 256, 0, stream>>>(tokens, out, Ntok, P, D);
}

}  // namespace dclip
