// Training-mode kernels of the trainable tail: what `loss.backward()` touches in the reference's
// training step (train_denseclip.py:1226-1330).  The backbone and the text tower are frozen (train_denseclip.py:1040-1044) and
// the heads consume the neck output of the ORIGINAL backbone features (denseclip.py:755-812), so gradients reach exactly
//   ViTFeatureFusionNeck (models.py:717-782): 12 x [conv3x3 -> BatchNorm(batch statistics) -> ReLU], concat, conv1x1 -> BN -> ReLU
//   FCNHead x 2 (denseclip.py:305-349):       conv3x3 -> BN -> ReLU -> Dropout(0.1) -> conv1x1 -> classifier conv1x1
//   F.interpolate(bilinear, align_corners=False) to the ground-truth size (denseclip.py:822-858)
//   CrossEntropyLoss(ignore_index) + SILogLoss (train_denseclip.py:1086-1096, losses.py:21-79)
// The matrix products (forward convs, weight gradients, input gradients) run on the tcgen05 GEMM (gemm_tcgen05.cuh: implicit
// conv mode and the shifted-K weight-gradient mode); this file holds everything around them.  All activations are token-major
// fp32 [M = B*gh*gw, N channels]; every reduction is a fixed two-stage tree in fp64 (deterministic, no atomics).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace dclip {

// ---------------------------------------------------------------------------------------------------------
// Column reductions over the rows of token-major matrices.
//   mode 0 (BatchNorm batch statistics, nn.BatchNorm2d in training mode): out0 = mean, out1 = biased variance,
//           out2 = rstd = 1/sqrt(var + eps); optionally the running statistics are updated in place
//           (running = (1 - momentum) * running + momentum * new, the variance one unbiased: M / (M - 1)).
//   mode 1 (backward reductions): g = a (upstream gradient) [* dropout mask * mask_scale] [* (BN output > 0) for ReLU];
//           out0 = sum_rows g (= dbeta, or the bias gradient of a plain conv when x == nullptr),
//           out1 = sum_rows g * xhat (= dgamma), xhat = (x - mean) * rstd.
// ---------------------------------------------------------------------------------------------------------
struct ColReduceParams {
  const float* a; long long lda;
  const float* x; long long ldx;
  const float *mean, *rstd, *gamma, *beta;
  const uint8_t* mask; long long ldm; float mask_scale;
  int relu, M, N, mode;
  double* part;   // [nblk][2][N]
  int nblk, rows_per_blk;
  float *out0, *out1, *out2;
  float eps;
  float *run_mean, *run_var; float momentum;
};

// g of one element for the backward modes (shared by the reduce and the apply kernels so they agree bit for bit)
__device__ __forceinline__ float tail_grad_elem(float a, float xhat, float gamma, float beta, int relu, const uint8_t* mask,
                                                long long midx, float mask_scale) {
  float g = a;
  if (mask) g = mask[midx] ? g * mask_scale : 0.f;
  if (relu && !(fmaf(gamma, xhat, beta) > 0.f)) g = 0.f;
  return g;
}

__global__ void __launch_bounds__(256) col_reduce_partial_kernel(const ColReduceParams p) {
  // block = 32 columns x 8 row lanes; grid = (ceil(N/32), nblk)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * p.rows_per_blk, r1 = min(p.M, r0 + p.rows_per_blk);
  double s0 = 0.0, s1 = 0.0;
  if (c < p.N) {
    float mean = 0.f, rstd = 1.f, gamma = 1.f, beta = 0.f;
    if (p.mode == 1 && p.x) { mean = p.mean[c]; rstd = p.rstd[c]; gamma = p.gamma[c]; beta = p.beta[c]; }
    float f0 = 0.f, f1 = 0.f;
    int n = 0;
    for (int r = r0 + ty; r < r1; r += 8) {
      const float a = p.a[(long long)r * p.lda + c];
      if (p.mode == 0) {
        f0 += a;
        f1 = fmaf(a, a, f1);
      } else {
        const float xhat = p.x ? (p.x[(long long)r * p.ldx + c] - mean) * rstd : 0.f;
        const float g = tail_grad_elem(a, xhat, gamma, beta, p.x ? p.relu : 0, p.mask, (long long)r * p.ldm + c, p.mask_scale);
        f0 += g;
        f1 = fmaf(g, xhat, f1);
      }
      if (++n == 16) { s0 += f0; s1 += f1; f0 = f1 = 0.f; n = 0; }   // short fp32 runs, fp64 across them
    }
    s0 += f0;
    s1 += f1;
  }
  __shared__ double sh[2][8][32];
  sh[0][ty][tx] = s0;
  sh[1][ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && c < p.N) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t0 += sh[0][k][tx]; t1 += sh[1][k][tx]; }
    p.part[((long long)blockIdx.y * 2 + 0) * p.N + c] = t0;
    p.part[((long long)blockIdx.y * 2 + 1) * p.N + c] = t1;
  }
}

__global__ void __launch_bounds__(256) col_reduce_final_kernel(const ColReduceParams p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.N) return;
  double t0 = 0.0, t1 = 0.0;
  for (int b = 0; b < p.nblk; ++b) {
    t0 += p.part[((long long)b * 2 + 0) * p.N + c];
    t1 += p.part[((long long)b * 2 + 1) * p.N + c];
  }
  if (p.mode == 0) {
    const double mean = t0 / p.M;
    double var = t1 / p.M - mean * mean;
    if (var < 0.0) var = 0.0;
    p.out0[c] = float(mean);
    p.out1[c] = float(var);
    if (p.out2) p.out2[c] = float(1.0 / sqrt(var + double(p.eps)));
    if (p.run_mean) p.run_mean[c] = (1.f - p.momentum) * p.run_mean[c] + p.momentum * float(mean);
    if (p.run_var) p.run_var[c] = (1.f - p.momentum) * p.run_var[c] + p.momentum * float(p.M > 1 ? var * p.M / (p.M - 1) : var);
  } else {
    p.out0[c] = float(t0);
    if (p.out1) p.out1[c] = float(t1);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Elementwise BatchNorm(+ReLU)(+Dropout) forward and backward on token-major matrices.
//   mode 0: y = gamma * (x - mean) * rstd + beta; ReLU; dropout (mask * mask_scale) -> out_f32 and/or out_bf16
//   mode 1: dx = gamma * rstd * (g - sum_g / M - xhat * sum_gx / M), g as in col_reduce mode 1 -> out_f32 and/or out_bf16
//   mode 2: dx = g only (dropout / ReLU masks without BatchNorm; x may be null) -> out_f32 and/or out_bf16
// ---------------------------------------------------------------------------------------------------------
struct BnApplyParams {
  const float* a; long long lda;     // mode 0: unused; modes 1, 2: upstream gradient
  const float* x; long long ldx;     // pre-BatchNorm activations
  const float *mean, *rstd, *gamma, *beta, *sum_g, *sum_gx;
  const uint8_t* mask; long long ldm; float mask_scale;
  int relu, M, N, mode;
  float* out_f32; long long ldo;
  __nv_bfloat16* out_bf16; long long ldb;
};

__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyParams p) {
  const long long total = (long long)p.M * p.N;
  const float inv_m = 1.f / float(p.M);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % p.N);
    const long long r = i / p.N;
    const bool bn = p.mean != nullptr;
    const float mean = bn ? p.mean[c] : 0.f, rstd = bn ? p.rstd[c] : 1.f, gamma = bn ? p.gamma[c] : 1.f, beta = bn ? p.beta[c] : 0.f;
    const float xv = p.x ? p.x[r * p.ldx + c] : 0.f;
    const float xhat = (xv - mean) * rstd;
    float y;
    if (p.mode == 0) {
      y = bn ? fmaf(gamma, xhat, beta) : xv;
      if (p.relu) y = fmaxf(y, 0.f);
      if (p.mask) y = p.mask[r * p.ldm + c] ? y * p.mask_scale : 0.f;
    } else {
      const float g = tail_grad_elem(p.a[r * p.lda + c], xhat, gamma, beta, p.x ? p.relu : 0, p.mask, r * p.ldm + c, p.mask_scale);
      y = p.mode == 1 ? gamma * rstd * (g - p.sum_g[c] * inv_m - xhat * p.sum_gx[c] * inv_m) : g;
    }
    if (p.out_f32) p.out_f32[r * p.ldo + c] = y;
    if (p.out_bf16) p.out_bf16[r * p.ldb + c] = __float2bfloat16(y);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Operand layout of the weight-gradient GEMMs: token-major [B][gh*gw][C] (fp32 or bf16, row pitch ld, image pitch bs) ->
// channel-major bf16 [C][ldk] over ZERO-PADDED images: pixel (b, y, x) sits at k0 = (b * (gh + pad) + y) * pitch + x with
// pitch >= gw + pad (pad = 1: at least one zero column per row and one zero row per image; pad = 0 and pitch = gw: a plain
// transpose, the 1x1-conv case).  The output column is k = k0 + lead - shift, i.e. out[c][k] = padded[c][k - lead + shift]:
// `shift` in {-1, 0, +1} produces the three horizontally shifted copies of X^T (TMA box starts must be 16-byte aligned in the
// innermost dimension, so the kx - 1 element shift of a 3x3 tap cannot be a TMA coordinate; the (ky - 1) * pitch row shift can,
// with pitch % 8 == 0), `lead` zero columns in front keep every coordinate non-negative.  Everything that is not a pixel is
// written as zero, up to ldk.  32x32 tiles through shared memory: reads coalesced along C, writes along k.
// ---------------------------------------------------------------------------------------------------------
struct TransposePadParams {
  const void* in; int in_f32; long long ld, bs;   // row pitch / image pitch in elements
  int B, gh, gw, C, pad, pitch, lead, shift;
  __nv_bfloat16* out; long long ldk;
  int nshift; long long plane;   // nshift = 3: planes out + j * plane (j = 0, 1, 2) receive shift = -1, 0, +1 from ONE read of the input
};

// One block = 32 output columns (k) x 32 channels.  The pixel decode (two integer divisions) is done once per input row of the
// tile by 34 threads and shared, not per element (the per-element 64-bit divisions made the first version issue-bound at 0.5 TB/s).
__global__ void __launch_bounds__(256) transpose_pad_kernel(const TransposePadParams p) {
  __shared__ float tile[34][33];
  __shared__ long long src_off[34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int ph = p.gh + p.pad, pw = p.pitch;
  const long long K = (long long)p.B * ph * pw;
  const int lo = p.nshift == 3 ? -1 : p.shift;          // padded index of tile row i: k0 - lead + lo + i
  const int nrows = p.nshift == 3 ? 34 : 32;
  if (threadIdx.x < nrows) {
    const long long k = (long long)k0 - p.lead + lo + threadIdx.x;
    long long off = -1;
    if (k >= 0 && k < K) {
      const unsigned ku = unsigned(k);                   // K < 2^31 (checked by the launcher)
      const unsigned row = ku / unsigned(pw), x = ku - row * unsigned(pw);
      const unsigned b = row / unsigned(ph), y = row - b * unsigned(ph);
      if (int(x) < p.gw && int(y) < p.gh) off = (long long)b * p.bs + ((long long)y * p.gw + x) * p.ld;
    }
    src_off[threadIdx.x] = off;
  }
  __syncthreads();
  const int c = c0 + tx;
  for (int i = ty; i < nrows; i += 8) {
    const long long off = src_off[i];
    float v = 0.f;
    if (off >= 0 && c < p.C)
      v = p.in_f32 ? static_cast<const float*>(p.in)[off + c] : __bfloat162float(static_cast<const __nv_bfloat16*>(p.in)[off + c]);
    tile[i][tx] = v;
  }
  __syncthreads();
  const long long k = k0 + tx;
  if (k >= p.ldk) return;
  for (int j = 0; j < (p.nshift == 3 ? 3 : 1); ++j) {
    __nv_bfloat16* o = p.out + j * p.plane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int cc = c0 + ty + 8 * q;
      if (cc < p.C) o[(long long)cc * p.ldk + k] = __float2bfloat16(tile[tx + j][ty + 8 * q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of F.interpolate(mode='bilinear', align_corners=False) (denseclip.py:838, 849): the adjoint of the forward gather,
// written as a gather itself (every low-resolution cell sums the output pixels it contributed to: deterministic).
//   dout NCHW fp32 [B, K, H, W] -> dtok token-major fp32 [B*gh*gw, ldc] (column k)
// One block per (b, k, low-res row y): pass 1 folds the output rows that touch row y into a W-wide line in shared memory
// (coalesced along W), pass 2 folds that line into the gw cells.  Same index rule as the forward kernels (bilinear_src).
// ---------------------------------------------------------------------------------------------------------
struct UpsampleBwdParams {
  const float* dout; int B, K, H, W, gh, gw;
  float* dtok; long long ldc;
};

__device__ __forceinline__ void bilinear_src_tt(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (float(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = int(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - float(i0);
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample_bilinear_bwd_kernel(const UpsampleBwdParams p) {
  extern __shared__ float line[];   // [W]
  const int y = blockIdx.x % p.gh;
  const int k = (blockIdx.x / p.gh) % p.K;
  const int b = blockIdx.x / (p.gh * p.K);
  const float sy = float(p.gh) / float(p.H), sx = float(p.gw) / float(p.W);
  const float ry = float(p.H) / float(p.gh), rx = float(p.W) / float(p.gw);
  const float* src = p.dout + ((long long)b * p.K + k) * p.H * p.W;
  const int Y0 = max(0, int(floorf((float(y) - 1.f) * ry)) - 1), Y1 = min(p.H, int(ceilf((float(y) + 2.f) * ry)) + 1);
  for (int X = threadIdx.x; X < p.W; X += blockDim.x) line[X] = 0.f;
  __syncthreads();
  for (int Y = Y0; Y < Y1; ++Y) {
    int y0, y1;
    float l0, l1;
    bilinear_src_tt(Y, sy, p.gh, y0, y1, l0, l1);
    const float w = (y0 == y ? l0 : 0.f) + (y1 == y ? l1 : 0.f);
    if (w == 0.f) continue;   // block-uniform
    for (int X = threadIdx.x; X < p.W; X += blockDim.x) line[X] = fmaf(w, src[(long long)Y * p.W + X], line[X]);
  }
  __syncthreads();
  for (int x = threadIdx.x; x < p.gw; x += blockDim.x) {
    const int X0 = max(0, int(floorf((float(x) - 1.f) * rx)) - 1), X1 = min(p.W, int(ceilf((float(x) + 2.f) * rx)) + 1);
    float acc = 0.f;
    for (int X = X0; X < X1; ++X) {
      int x0, x1;
      float l0, l1;
      bilinear_src_tt(X, sx, p.gw, x0, x1, l0, l1);
      const float w = (x0 == x ? l0 : 0.f) + (x1 == x ? l1 : 0.f);
      acc = fmaf(w, line[X], acc);
    }
    p.dtok[(((long long)b * p.gh + y) * p.gw + x) * p.ldc + k] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Losses.  Block partials in fp64, one finishing block: deterministic.
//   cross entropy, mean over the non-ignored pixels (torch.nn.CrossEntropyLoss(ignore_index), train_denseclip.py:1086):
//     logits NCHW fp32 [B, K, HW], target int64 [B, HW]; stats[0] = loss, stats[1] = number of counted pixels
//   SILog (losses.py:21-79): d = log(max(pred, eps)) - log(max(target, eps)) on the masked pixels,
//     loss = sum d^2 / T - lambda * (sum d)^2 / T^2; stats = {loss, T, sum d}
// ---------------------------------------------------------------------------------------------------------
struct LossParams {
  const float* pred;        // CE: logits; SILog: prediction
  const long long* target;  // CE
  const float* ftarget;     // SILog
  const uint8_t* mask;      // SILog (bool, optional)
  int B, K; long long HW;
  int ignore_index; float lambd, eps;
  double* part; int nblk;   // [nblk][3]
  float* stats;             // [4]
  const float* gout;        // backward: upstream scalar gradient (device)
  float* grad;              // backward: same shape as pred
};

__device__ __forceinline__ void block_sum3(double& a, double& b, double& c) {
  __shared__ double sh[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  if (w == 0) {
    a = l < 8 ? sh[0][l] : 0.0; b = l < 8 ? sh[1][l] : 0.0; c = l < 8 ? sh[2][l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
  }
}

__global__ void __launch_bounds__(256) ce_loss_partial_kernel(const LossParams p) {
  const long long total = (long long)p.B * p.HW;
  double loss = 0.0, cnt = 0.0, unused = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = p.target[i];
    if (t == p.ignore_index) continue;
    const long long b = i / p.HW, px = i - b * p.HW;
    const float* lg = p.pred + b * p.K * p.HW + px;
    float m = -INFINITY;
    for (int k = 0; k < p.K; ++k) m = fmaxf(m, lg[(long long)k * p.HW]);
    float s = 0.f;
    for (int k = 0; k < p.K; ++k) s += expf(lg[(long long)k * p.HW] - m);
    const float lt = (t >= 0 && t < p.K) ? lg[t * p.HW] : 0.f;
    loss += double(m + logf(s) - lt);
    cnt += 1.0;
  }
  block_sum3(loss, cnt, unused);
  if (threadIdx.x == 0) { p.part[blockIdx.x * 3 + 0] = loss; p.part[blockIdx.x * 3 + 1] = cnt; p.part[blockIdx.x * 3 + 2] = 0.0; }
}

__global__ void __launch_bounds__(256) silog_loss_partial_kernel(const LossParams p) {
  const long long total = (long long)p.B * p.HW;
  double s1 = 0.0, s2 = 0.0, cnt = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (p.mask && !p.mask[i]) continue;
    const float d = logf(fmaxf(p.pred[i], p.eps)) - logf(fmaxf(p.ftarget[i], p.eps));
    s1 += double(d);
    s2 += double(d) * double(d);
    cnt += 1.0;
  }
  block_sum3(s1, s2, cnt);
  if (threadIdx.x == 0) { p.part[blockIdx.x * 3 + 0] = s1; p.part[blockIdx.x * 3 + 1] = s2; p.part[blockIdx.x * 3 + 2] = cnt; }
}

// kind 0: cross entropy, kind 1: SILog
__global__ void __launch_bounds__(256) loss_final_kernel(const LossParams p, int kind) {
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < p.nblk; i += blockDim.x) { a += p.part[i * 3]; b += p.part[i * 3 + 1]; c += p.part[i * 3 + 2]; }
  block_sum3(a, b, c);
  if (threadIdx.x == 0) {
    if (kind == 0) {
      p.stats[0] = b > 0.0 ? float(a / b) : NAN;   // torch: mean over zero counted pixels is nan
      p.stats[1] = float(b);
      p.stats[2] = 0.f;
    } else {
      p.stats[0] = c > 0.0 ? float(b / c - double(p.lambd) * a * a / (c * c)) : 0.f;   // losses.py:51-53: 0 when nothing is valid
      p.stats[1] = float(c);
      p.stats[2] = float(a);
    }
    p.stats[3] = 0.f;
  }
}

__global__ void __launch_bounds__(256) ce_loss_bwd_kernel(const LossParams p) {
  const long long total = (long long)p.B * p.HW;
  const float scale = p.stats[1] > 0.f ? p.gout[0] / p.stats[1] : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = p.target[i];
    const long long b = i / p.HW, px = i - b * p.HW;
    const float* lg = p.pred + b * p.K * p.HW + px;
    float* gr = p.grad + b * p.K * p.HW + px;
    if (t == p.ignore_index) {
      for (int k = 0; k < p.K; ++k) gr[(long long)k * p.HW] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int k = 0; k < p.K; ++k) m = fmaxf(m, lg[(long long)k * p.HW]);
    float s = 0.f;
    for (int k = 0; k < p.K; ++k) s += expf(lg[(long long)k * p.HW] - m);
    const float inv = 1.f / s;
    for (int k = 0; k < p.K; ++k)
      gr[(long long)k * p.HW] = scale * (expf(lg[(long long)k * p.HW] - m) * inv - (k == t ? 1.f : 0.f));
  }
}

__global__ void __launch_bounds__(256) silog_loss_bwd_kernel(const LossParams p) {
  const long long total = (long long)p.B * p.HW;
  const float T = p.stats[1], S1 = p.stats[2], g = p.gout[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float out = 0.f;
    if (T > 0.f && (!p.mask || p.mask[i])) {
      const float pr = p.pred[i];
      if (pr >= p.eps) {   // torch.clamp(min=eps) passes the gradient where pred >= eps
        const float d = logf(pr) - logf(fmaxf(p.ftarget[i], p.eps));
        out = g * (2.f * d / T - 2.f * p.lambd * S1 / (T * T)) / pr;
      }
    }
    p.grad[i] = out;
  }
}

}  // namespace dclip
