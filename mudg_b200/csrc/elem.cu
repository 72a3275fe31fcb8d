// HBM-bound kernels of the path: GroupNorm (stats / finalize / apply+SiLU), LayerNorm, channel concat,
// nearest 2x upsample, stride-2 im2col, row softmax, weight packing, embedding MLP pieces.
// All activations are channels-last fp16 with C % 8 == 0 -> every access is a 128-bit vector.
#include "ops.h"

#include <algorithm>

namespace mudg {

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// The same with the reciprocal on the FMA pipe (magic-number seed + 3 Newton steps, ~1e-7 relative): ONE MUFU op per element
// instead of two.  gn_apply is co-limited by the MUFU unit (ncu: XU 66 %) -- see knob gn_silu.
__device__ __forceinline__ float silu_fma(float x) {
  const float d = 1.f + __expf(-fmaxf(x, -80.f));             // <= 5.6e34: the seed below needs d < 2^126
  float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  return x * r;
}

// The default: one MUFU op and three instructions, silu(x) = h + h tanh(h), h = x / 2 (MUFU.TANH).  PTX only promises
// 2^-11 relative for tanh.approx, which the cancellation in 1 + tanh(h) could turn into 5e-4 |h| absolute for negative
// inputs; MEASURED against float64 on a level-0 tensor (tests/gpu_probe_silu.py, profiles/r2_silu_tanh.log) the fp16 output
// differs from the exact SiLU by 7.372e-5 on average against 7.371e-5 for fp16 rounding alone, and by at most 7.1e-5 for
// inputs below -2: the unit is far better than its bound near saturation.  gn_apply was co-limited by the MUFU unit (XU 66 %
// with ex2 + rcp); forward -0.9 %.
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ---------------------------------------------------------------- GroupNorm statistics
// x: [S samples][R rows][C]; sums: [S][32][2] doubles (sum, sum of squares), pre-zeroed.
// block = (C/8) x rows_per_iter threads; a thread owns one 8-channel vector column and strides over rows,
// so partial sums stay in registers until one flush at the end.
__global__ void gn_stats_kernel(const __half* __restrict__ x, double* __restrict__ sums, int64_t R, int C, int cpg,
                                int rows_per_block) {
  // fp64 accumulators: the ORDER of the shared / global atomics varies from run to run, and in fp32 that moved the statistics
  // in the 7th digit -- enough to flip fp16 roundings downstream and make two identical forwards differ by a few fp16 ulps
  // (6e-3 max on the UNet output).  With fp64 sums of per-thread fp32 partials (fixed order) the result is repeatable.
  __shared__ double acc[32][2];
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs, r0 = threadIdx.x / vecs, rstep = blockDim.x / vecs;
  if (threadIdx.x < 64) (&acc[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  const int s = blockIdx.y;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(R, row_begin + rows_per_block);
  float sm[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; i++) sm[i] = sq[i] = 0.f;
  const __half* base = x + ((int64_t)s * R) * C + v * 8;
  if (r0 < rstep) {
    int64_t r = row_begin + r0;
    for (; r + 3 * (int64_t)rstep < row_end; r += 4 * (int64_t)rstep) {   // 4 independent 128-bit loads in flight
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; u++) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + (r + u * (int64_t)rstep) * C));
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float f[8];
        unpack8(raw[u], f);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          sm[i] += f[i];
          sq[i] += f[i] * f[i];
        }
      }
    }
    for (; r < row_end; r += rstep) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + r * C));
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sm[i] += f[i];
        sq[i] += f[i] * f[i];
      }
    }
    // flush: merge channels of the same group first
    int g_cur = (v * 8) / cpg;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int g = (v * 8 + i) / cpg;
      if (g != g_cur) {
        atomicAdd(&acc[g_cur][0], (double)a);
        atomicAdd(&acc[g_cur][1], (double)b);
        a = b = 0.f;
        g_cur = g;
      }
      a += sm[i];
      b += sq[i];
    }
    atomicAdd(&acc[g_cur][0], (double)a);
    atomicAdd(&acc[g_cur][1], (double)b);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    atomicAdd(&sums[((int64_t)s * 32 + g) * 2 + k], acc[g][k]);
  }
}

// Normalise + affine (+ SiLU) in ONE pass straight from the (sum, sum of squares) pairs: block = (C/8) x rows_per_iter
// threads working on rows [row_begin, row_end) of sample blockIdx.y.  A thread owns one 8-channel vector column, so its
// 8 scales / shifts  y = x * scale + shift == (x - mean) * rstd * gamma + beta  are computed once and stay in registers;
// the row loop keeps 4 independent 128-bit loads in flight per thread.
__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y, const double* __restrict__ sums,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int64_t R, int C, int cpg,
                                double count, float eps, int rows_per_block, int act) {
  const int vecs = C >> 3;
  const int v = threadIdx.x % vecs, r0 = threadIdx.x / vecs, rstep = blockDim.x / vecs;
  const int s = blockIdx.y;
  if (r0 >= rstep) return;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int c = v * 8 + i, g = c / cpg;
    const double mean = sums[((int64_t)s * 32 + g) * 2] / count;
    double var = sums[((int64_t)s * 32 + g) * 2 + 1] / count - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf((float)var + eps);
    sc[i] = __ldg(gamma + c) * rstd;
    sh[i] = __ldg(beta + c) - (float)mean * sc[i];
  }
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(R, row_begin + rows_per_block);
  const __half* xb = x + ((int64_t)s * R) * C + v * 8;
  __half* yb = y + ((int64_t)s * R) * C + v * 8;
  auto one = [&](const uint4& raw) {
    float f[8];
    unpack8(raw, f);
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = fmaf(f[i], sc[i], sh[i]);
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu(f[i]);
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu_fma(f[i]);
    } else if (act == 3) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu_tanh(f[i]);
    }
    return pack8(f);
  };
  int64_t r = row_begin + r0;
  for (; r + 3 * (int64_t)rstep < row_end; r += 4 * (int64_t)rstep) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; u++) raw[u] = __ldg(reinterpret_cast<const uint4*>(xb + (r + u * (int64_t)rstep) * C));
#pragma unroll
    for (int u = 0; u < 4; u++) *reinterpret_cast<uint4*>(yb + (r + u * (int64_t)rstep) * C) = one(raw[u]);
  }
  for (; r < row_end; r += rstep) *reinterpret_cast<uint4*>(yb + r * C) = one(__ldg(reinterpret_cast<const uint4*>(xb + r * C)));
}

// ---------------------------------------------------------------- GroupNorm in ONE kernel for samples that fit shared memory
// One CTA per (sample, slab of Cs channels = whole groups, multiple of 8): the [R][Cs] slab is read once into shared
// memory with its per-channel sums, reduced in a FIXED order (no atomics, no pre-zeroed buffer, no statistics pass), then
// normalised (+ SiLU) from shared memory.  2 passes over HBM instead of 3 and 1 launch instead of 3 (memset, statistics,
// apply): the levels 1-3 of MDM512 and 2-3 of MDM1024, where the tensors are a few MB and launches dominate.
constexpr int GNS_THREADS = 256;
__global__ void __launch_bounds__(GNS_THREADS) gn_small_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               int R, int C, int cpg, int Cs, float eps, int act) {
  extern __shared__ __align__(16) unsigned char gns_smem[];
  const int vpr = Cs >> 3, rstep = GNS_THREADS / vpr;
  uint4* slab = reinterpret_cast<uint4*>(gns_smem);                        // [R][vpr]
  float* part = reinterpret_cast<float*>(slab + (size_t)R * vpr);          // [rstep][Cs][2]
  double* csum = reinterpret_cast<double*>(part + (size_t)rstep * Cs * 2); // [Cs][2]
  float* stat = reinterpret_cast<float*>(csum + Cs * 2);                   // [groups in slab][2]: mean, rstd
  const int s = blockIdx.y, c0 = blockIdx.x * Cs;
  const int v = threadIdx.x % vpr, r0 = threadIdx.x / vpr;
  const __half* xb = x + ((int64_t)s * R) * C + c0 + v * 8;
  if (r0 < rstep) {
    float sm[8], sq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) sm[i] = sq[i] = 0.f;
    for (int r = r0; r < R; r += rstep) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)r * C));
      slab[r * vpr + v] = raw;
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int i = 0; i < 8; i++) { sm[i] += f[i]; sq[i] = fmaf(f[i], f[i], sq[i]); }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      part[((size_t)r0 * Cs + v * 8 + i) * 2] = sm[i];
      part[((size_t)r0 * Cs + v * 8 + i) * 2 + 1] = sq[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < Cs) {                      // per-channel totals, rows partitions summed in order
    double a = 0.0, b = 0.0;
    for (int k = 0; k < rstep; k++) {
      a += (double)part[((size_t)k * Cs + threadIdx.x) * 2];
      b += (double)part[((size_t)k * Cs + threadIdx.x) * 2 + 1];
    }
    csum[threadIdx.x * 2] = a;
    csum[threadIdx.x * 2 + 1] = b;
  }
  __syncthreads();
  if (threadIdx.x < Cs / cpg) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < cpg; k++) {
      a += csum[(threadIdx.x * cpg + k) * 2];
      b += csum[(threadIdx.x * cpg + k) * 2 + 1];
    }
    const double count = (double)R * cpg, mean = a / count;
    double var = b / count - mean * mean;
    if (var < 0) var = 0;
    stat[threadIdx.x * 2] = (float)mean;
    stat[threadIdx.x * 2 + 1] = rsqrtf((float)var + eps);
  }
  __syncthreads();
  if (r0 >= rstep) return;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int cl = v * 8 + i, g = cl / cpg;
    sc[i] = __ldg(gamma + c0 + cl) * stat[g * 2 + 1];
    sh[i] = __ldg(beta + c0 + cl) - stat[g * 2] * sc[i];
  }
  __half* yb = y + ((int64_t)s * R) * C + c0 + v * 8;
  for (int r = r0; r < R; r += rstep) {
    float f[8];
    unpack8(slab[r * vpr + v], f);
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = fmaf(f[i], sc[i], sh[i]);
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu(f[i]);
    } else if (act == 2) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu_fma(f[i]);
    } else if (act == 3) {
#pragma unroll
      for (int i = 0; i < 8; i++) f[i] = silu_tanh(f[i]);
    }
    *reinterpret_cast<uint4*>(yb + (int64_t)r * C) = pack8(f);
  }
}

// ---------------------------------------------------------------- LayerNorm (warp per row, row kept in registers)
template <int MAXV>
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int vecs = C >> 3;
  float f[MAXV][8];
  float sum = 0.f;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
#pragma unroll
  for (int k = 0; k < MAXV; k++) {
    const int v = lane + 32 * k;
    if (v < vecs) {
      unpack8(__ldg(xr + v), f[k]);
#pragma unroll
      for (int i = 0; i < 8; i++) sum += f[k][i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; k++) {
    const int v = lane + 32 * k;
    if (v < vecs) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float d = f[k][i] - mean;
        var += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / C + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int k = 0; k < MAXV; k++) {
    const int v = lane + 32 * k;
    if (v < vecs) {
      const float4* g4 = reinterpret_cast<const float4*>(gamma + v * 8);
      const float4* b4 = reinterpret_cast<const float4*>(beta + v * 8);
      const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
      float o[8];
      o[0] = (f[k][0] - mean) * rstd * g0.x + b0.x; o[1] = (f[k][1] - mean) * rstd * g0.y + b0.y;
      o[2] = (f[k][2] - mean) * rstd * g0.z + b0.z; o[3] = (f[k][3] - mean) * rstd * g0.w + b0.w;
      o[4] = (f[k][4] - mean) * rstd * g1.x + b1.x; o[5] = (f[k][5] - mean) * rstd * g1.y + b1.y;
      o[6] = (f[k][6] - mean) * rstd * g1.z + b1.z; o[7] = (f[k][7] - mean) * rstd * g1.w + b1.w;
      yr[v] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------- layout helpers
__global__ void concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o,
                              int64_t rows, int64_t rows_b, int va, int vb) {
  const int vo = va + vb;
  const int64_t total = rows * vo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vo;
    const int v = i - r * vo;
    const int64_t rb = r < rows_b ? r : r % rows_b;
    o[i] = v < va ? __ldg(a + r * va + v) : __ldg(b + rb * vb + (v - va));
  }
}

// Channel concat of one sample's rows [row_begin, row_end) that also accumulates the GroupNorm statistics of its output
// (the ResBlock that consumes a skip concat starts with a GroupNorm): same thread layout and accumulation discipline as
// gn_stats_kernel -- a thread owns one 8-channel vector column of the OUTPUT, fp32 partials, fp64 atomics.
__global__ void concat_stats_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o,
                                    double* __restrict__ sums, int64_t R, int64_t rows_b, int va, int vb, int cpg,
                                    int rows_per_block) {
  __shared__ double acc[32][2];
  const int vecs = va + vb;
  const int v = threadIdx.x % vecs, r0 = threadIdx.x / vecs, rstep = blockDim.x / vecs;
  if (threadIdx.x < 64) (&acc[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  const int s = blockIdx.y;
  const int64_t row_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t row_end = min(R, row_begin + rows_per_block);
  float sm[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; i++) sm[i] = sq[i] = 0.f;
  if (r0 < rstep) {
    for (int64_t r = row_begin + r0; r < row_end; r += rstep) {
      const int64_t gr = (int64_t)s * R + r;                 // global row
      const int64_t rb = gr < rows_b ? gr : gr % rows_b;
      const uint4 raw = v < va ? __ldg(a + gr * va + v) : __ldg(b + rb * vb + (v - va));
      o[gr * vecs + v] = raw;
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sm[i] += f[i];
        sq[i] += f[i] * f[i];
      }
    }
    int g_cur = (v * 8) / cpg;
    float x = 0.f, y = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int g = (v * 8 + i) / cpg;
      if (g != g_cur) {
        atomicAdd(&acc[g_cur][0], (double)x);
        atomicAdd(&acc[g_cur][1], (double)y);
        x = y = 0.f;
        g_cur = g;
      }
      x += sm[i];
      y += sq[i];
    }
    atomicAdd(&acc[g_cur][0], (double)x);
    atomicAdd(&acc[g_cur][1], (double)y);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    atomicAdd(&sums[((int64_t)s * 32 + g) * 2 + k], acc[g][k]);
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int F, int H, int W, int vecs) {
  const int64_t total = (int64_t)F * 2 * H * 2 * W * vecs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int v = t % vecs; t /= vecs;
    const int w = t % (2 * W); t /= 2 * W;
    const int h = t % (2 * H);
    const int f = t / (2 * H);
    y[i] = __ldg(x + (((int64_t)f * H + (h >> 1)) * W + (w >> 1)) * vecs + v);
  }
}

// 3x3 stride-2 patches: out[f, ho, wo, (kh*3+kw)*C + c] = x[f, 2ho+kh-pad, 2wo+kw-pad, c] (0 outside).
// pad = 1: symmetric padding (UNet Downsample); pad = 0: zeros only at the bottom/right (VAE Downsample's (0,1,0,1) pad)
__global__ void im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int F, int H, int W, int Ho, int Wo,
                                 int vecs, int pad) {
  const int64_t total = (int64_t)F * Ho * Wo * 9 * vecs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int v = t % vecs; t /= vecs;
    const int tap = t % 9; t /= 9;
    const int wo = t % Wo; t /= Wo;
    const int ho = t % Ho;
    const int f = t / Ho;
    const int h = 2 * ho + tap / 3 - pad, w = 2 * wo + tap % 3 - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W) val = __ldg(x + (((int64_t)f * H + h) * W + w) * vecs + v);
    y[i] = val;
  }
}

// in-place softmax over rows of length n (fp16 storage, fp32 math); one block per row
__global__ void softmax_rows_kernel(__half* __restrict__ x, int64_t rows, int n) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  __half* xr = x + row * n;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, __half2float(xr[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (blockDim.x >> 5); i++) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += __expf(__half2float(xr[i]) - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); i++) s += red[i];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xr[i] = __float2half_rn(__expf(__half2float(xr[i]) - m) * inv);
}

// ---------------------------------------------------------------- weights
// src [O][I][taps] (fp32 or fp16, PyTorch conv layout flattened) -> dst [O][taps][Ipad] fp16, zero padded
template <typename TS>
__global__ void pack_weight_kernel(const TS* __restrict__ src, __half* __restrict__ dst, int O, int I, int taps,
                                   int Ipad) {
  const int64_t total = (int64_t)O * taps * Ipad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c = t % Ipad; t /= Ipad;
    const int tap = t % taps;
    const int o = t / taps;
    float v = 0.f;
    if (c < I) v = static_cast<float>(src[((int64_t)o * I + c) * taps + tap]);
    dst[i] = __float2half_rn(v);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = static_cast<TD>(static_cast<float>(s[i]));
}

// rows of `src` [rows][cols] gathered into dst in blocks: used to interleave GEGLU value/gate rows.
// dst row (g*128 + j) = src row (g*64 + j) for j<64 (value), src row (half + g*64 + j-64) for j>=64 (gate)
__global__ void geglu_interleave_kernel(const __half* __restrict__ w, const float* __restrict__ b,
                                        __half* __restrict__ wo, float* __restrict__ bo, int half_rows, int cols) {
  const int64_t total = (int64_t)2 * half_rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = i - r * cols;
    const int g = r / 128, j = r % 128;
    const int64_t sr = j < 64 ? (int64_t)g * 64 + j : (int64_t)half_rows + g * 64 + (j - 64);
    wo[i] = w[sr * cols + c];
    if (c == 0 && b != nullptr) bo[r] = b[sr];
  }
}

// ---------------------------------------------------------------- embeddings
// [cos | sin] sinusoid of an int64 index (utils_diffusion.py:8-28)
// blockIdx.y picks the index vector (timestep / class label / fps): out [3][B][dim]
__global__ void sinusoid_kernel(const int64_t* __restrict__ t0, const int64_t* __restrict__ t1, const int64_t* __restrict__ t2,
                                float* __restrict__ out, int B, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int64_t* t = blockIdx.y == 0 ? t0 : (blockIdx.y == 1 ? t1 : t2);
  out += (size_t)blockIdx.y * B * dim;
  const int b = i / half, k = i % half;
  const float freq = expf(-logf(10000.f) * (float)k / (float)half);
  const float arg = (float)t[b] * freq;
  out[b * dim + k] = cosf(arg);
  out[b * dim + half + k] = sinf(arg);
}

// Up to 24 small Linear layers in ONE launch (the three embedding MLPs, the 22 ResBlock emb_layers): warp per output,
// fp32 activations, fp16 weights read as 128-bit vectors.  Separate outputs (job j: y_j[b][n], n < n_out_j) or, with
// `sum`, one output y_0[b][n] = sum_j (x_j[b] . W_j[n] + bias_j[n]); `silu_out` applies SiLU to what is stored.
__global__ void __launch_bounds__(256) batched_linear_kernel(const LinearBatch lb) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int per_row = lb.sum ? lb.off[1] : lb.off[lb.count];
  if (warp >= (int64_t)lb.Bn * per_row) return;
  const int b = (int)(warp / per_row), gn = (int)(warp % per_row);
  int j0 = 0, j1 = lb.count, n = gn;
  if (!lb.sum) {
    while (gn >= lb.off[j0 + 1]) j0++;
    j1 = j0 + 1;
    n = gn - lb.off[j0];
  }
  float acc = 0.f;
  for (int j = j0; j < j1; j++) {
    const float* xr = lb.x[j] + (int64_t)b * lb.K;
    const __half* wr = lb.W[j] + (int64_t)n * lb.K;
    if ((lb.K & 7) == 0) {
      for (int k = lane * 8; k < lb.K; k += 256) {
        const uint4 w8 = __ldg(reinterpret_cast<const uint4*>(wr + k));
        const float4 x0 = *reinterpret_cast<const float4*>(xr + k), x1 = *reinterpret_cast<const float4*>(xr + k + 4);
        float w[8];
        unpack8(w8, w);
        acc = fmaf(x0.x, w[0], acc); acc = fmaf(x0.y, w[1], acc); acc = fmaf(x0.z, w[2], acc); acc = fmaf(x0.w, w[3], acc);
        acc = fmaf(x1.x, w[4], acc); acc = fmaf(x1.y, w[5], acc); acc = fmaf(x1.z, w[6], acc); acc = fmaf(x1.w, w[7], acc);
      }
    } else {
      for (int k = lane; k < lb.K; k += 32) acc = fmaf(xr[k], __half2float(wr[k]), acc);
    }
    if (lane == 0 && lb.bias[j] != nullptr) acc += lb.bias[j][n];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (lb.silu_out) acc = acc / (1.f + expf(-acc));
    lb.y[lb.sum ? 0 : j0][(int64_t)b * (lb.sum ? lb.off[1] : lb.off[j0 + 1] - lb.off[j0]) + n] = acc;
  }
}

// x [B][C][R] (fp32 or fp16; R = T*H*W rows, NCTHW) -> y [B][R][Cpad] fp16 channels-last, zero padded
template <typename TS>
__global__ void to_channels_last_kernel(const TS* __restrict__ x, __half* __restrict__ y, int B, int C, int64_t R,
                                        int Cpad) {
  const int64_t total = (int64_t)B * R * Cpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = i % Cpad;
    const int64_t r = (i / Cpad) % R;
    const int b = i / (Cpad * R);
    y[i] = __float2half_rn(c < C ? static_cast<float>(x[((int64_t)b * C + c) * R + r]) : 0.f);
  }
}

// dst[b][r][c] = (fp16) src[b][r0 + r][c]; src rows have pitch `cols`, batch pitch `src_rows * cols`
template <typename TS>
__global__ void gather_rows_kernel(const TS* __restrict__ src, __half* __restrict__ dst, int B, int src_rows, int r0,
                                   int nrows, int cols) {
  const int64_t total = (int64_t)B * nrows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = i % cols;
    const int r = (i / cols) % nrows;
    const int b = i / ((int64_t)cols * nrows);
    dst[i] = __float2half_rn(static_cast<float>(src[((int64_t)b * src_rows + r0 + r) * cols + c]));
  }
}

// y [B][R][Cp] channels-last fp16 -> out [B][C][R] fp16 (first C channels; NCTHW with R = T*H*W)
__global__ void from_channels_last_kernel(const __half* __restrict__ y, __half* __restrict__ out, int B, int C, int64_t R,
                                          int Cp) {
  const int64_t total = (int64_t)B * C * R;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i % R;
    const int c = (i / R) % C;
    const int b = i / (R * C);
    out[i] = y[((int64_t)b * R + r) * Cp + c];
  }
}

// LayerNorm statistics only (the normalisation itself is folded into the consuming GEMM): (mean, rstd) per row.
// A warp handles ROWS consecutive rows at once: all their 128-bit loads are issued before the first shuffle, so a warp
// keeps ROWS x MAXV x 16 B in flight per lane instead of one row's worth (the one-row version ran at 39 % of the HBM
// roofline inside the clip: too few bytes in flight per SM while the warps sat in their dependent shuffle chains), and
// the ROWS butterfly reductions interleave.
template <int MAXV, int ROWS>
__global__ void __launch_bounds__(256) ln_stats_kernel(const __half* __restrict__ x, float2* __restrict__ out, int64_t rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= rows) return;
  const int vecs = C >> 3;
  uint4 raw[ROWS][MAXV];
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (row0 + r) * C);
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
      const int v = lane + 32 * k;
      raw[r][k] = (v < vecs && row0 + r < rows) ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
    }
  }
  // one sweep over the registers: sum and sum of squares together (var = E[x^2] - mean^2 in fp32: for rows of 320..1280
  // O(1) activations the cancellation costs ~1e-7 (1 + mean^2 / var) relative, far below the fp16 data); the two-sweep form
  // spent more issue slots on the second unpack than the loads took (measured 36 % of the HBM roofline inside the clip)
  float sum[ROWS], sq[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; k++) {
      float f[8];
      unpack8(raw[r][k], f);              // padding vectors are zeros: they add nothing to either sum
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        s0 += f[i]; q0 = fmaf(f[i], f[i], q0);
        s1 += f[i + 1]; q1 = fmaf(f[i + 1], f[i + 1], q1);
      }
    }
    sum[r] = s0 + s1;
    sq[r] = q0 + q1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
      sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], o);
    }
  float var[ROWS], mean[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; r++) {
    mean[r] = sum[r] / C;
    var[r] = fmaxf(sq[r] - mean[r] * sum[r], 0.f);     // = C * variance
  }
  if (lane < ROWS && row0 + lane < rows) {
    float m = mean[0], v = var[0];
#pragma unroll
    for (int r = 1; r < ROWS; r++)
      if (lane == r) { m = mean[r]; v = var[r]; }
    out[row0 + lane] = make_float2(m, rsqrtf(v / C + eps));
  }
}

// (mean, rstd) per row from the per-64-column partial (sum, sum of squares) planes a producing GEMM's epilogue stored
// (TapGemm::ln_out): thread per row, the planes are read coalesced (consecutive rows are consecutive float2), all loads of a
// row issued before the first add.  Same single-sweep variance as ln_stats_kernel.
template <int PARTS>
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float2* __restrict__ parts, float2* __restrict__ out, int64_t rows,
                                                          int nparts, float inv_c, float eps) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f, q = 0.f;
  for (int c0 = 0; c0 < nparts; c0 += PARTS) {
    float2 v[PARTS];
#pragma unroll
    for (int j = 0; j < PARTS; j++) v[j] = c0 + j < nparts ? __ldg(parts + (int64_t)(c0 + j) * rows + r) : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < PARTS; j++) { s += v[j].x; q += v[j].y; }
  }
  const float mean = s * inv_c;
  const float var = fmaxf(q - mean * s, 0.f) * inv_c;
  out[r] = make_float2(mean, rsqrtf(var + eps));
}

// Fold a LayerNorm's affine part into the Linear that consumes it: W[n][k] *= gamma[k] (re-rounded to fp16, in place),
// c1[n] = sum_k fp16(W gamma), c2[n] = sum_k W[n][k] beta[k] + bias[n].  Warp per output row.
__global__ void ln_fold_kernel(__half* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ bias, float* __restrict__ c1, float* __restrict__ c2, int N, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  __half* wr = W + (int64_t)warp * K;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = __half2float(wr[k]);
    const __half wf = __float2half_rn(w * gamma[k]);
    wr[k] = wf;
    a += __half2float(wf);
    b += w * beta[k];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    c1[warp] = a;
    c2[warp] = b + (bias ? bias[warp] : 0.f);
  }
}

// GroupNorm (no activation) folded into the Linear that consumes it: per sample s
//   Ws[s][n][k] = fp16( W[n][k] * scale_s[k] ),   cs[s][n] = sum_k W[n][k] * shift_s[k]
// with scale = gamma * rstd(group), shift = beta - mean * scale straight from the fp64 (sum, sumsq) pairs.
// grid (N / 8, S), 8 warps per block = 8 output rows; the sample's scale / shift vectors are built once per block in smem.
__global__ void __launch_bounds__(256) gn_fold_weights_kernel(const __half* __restrict__ W, const double* __restrict__ sums,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              __half* __restrict__ Ws, float* __restrict__ cs, int N, int K,
                                                              int cpg, double count, float eps) {
  extern __shared__ float fold_sm[];       // [2][K]
  float* sc = fold_sm;
  float* sh = fold_sm + K;
  const int s = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int g = k / cpg;
    const double mean = sums[((int64_t)s * 32 + g) * 2] / count;
    double var = sums[((int64_t)s * 32 + g) * 2 + 1] / count - mean * mean;
    if (var < 0) var = 0;
    const float rstd = rsqrtf((float)var + eps);
    const float a = gamma[k] * rstd;
    sc[k] = a;
    sh[k] = beta[k] - (float)mean * a;
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const __half2* wr = reinterpret_cast<const __half2*>(W + (int64_t)n * K);
  __half2* wo = reinterpret_cast<__half2*>(Ws + ((int64_t)s * N + n) * K);
  float c = 0.f;
  for (int k2 = lane; k2 < K / 2; k2 += 32) {
    const float2 w = __half22float2(wr[k2]);
    wo[k2] = __floats2half2_rn(w.x * sc[2 * k2], w.y * sc[2 * k2 + 1]);
    c = fmaf(w.x, sh[2 * k2], c);
    c = fmaf(w.y, sh[2 * k2 + 1], c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cs[(int64_t)s * N + n] = c;
}

inline int grid_for(int64_t work, int threads) {
  int64_t b = (work + threads - 1) / threads;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sm_count() * 16));
}

}  // namespace

// ================================================================ launchers
static void gn_geometry(int S, int64_t rows_per_sample, int C, int& threads, int& rpb, int& chunks) {
  const int vecs = C / 8;
  const int rpi = std::max(1, 256 / vecs);
  threads = vecs * rpi;
  // enough blocks to fill the machine, at least 32 rows per block
  const int64_t want_blocks = std::max<int64_t>(1, (int64_t)sm_count() * 8 / std::max(1, S));
  int64_t r = std::max<int64_t>(32, (rows_per_sample + want_blocks - 1) / want_blocks);
  r = (r + rpi - 1) / rpi * rpi;
  rpb = (int)r;
  chunks = (int)((rows_per_sample + r - 1) / r);
}

void gn_stats(const __half* x, int S, int64_t rows_per_sample, int C, double* sums /*[S*64], zeroed*/, cudaStream_t st) {
  MUDG_REQUIRE(C % 32 == 0 && C / 8 <= 1024, "GroupNorm needs C %% 32 == 0 and C <= 8192 (C=%d)", C);
  int threads, rpb, chunks;
  gn_geometry(S, rows_per_sample, C, threads, rpb, chunks);
  gn_stats_kernel<<<dim3(chunks, S), threads, 0, st>>>(x, sums, rows_per_sample, C, C / 32, rpb);
  MUDG_CUDA(cudaGetLastError());
}

void gn_apply(const __half* x, __half* y, const double* sums, int S, int64_t rows_per_sample, int C, const float* gamma,
              const float* beta, float eps, bool silu_act, cudaStream_t st) {
  MUDG_REQUIRE(C % 32 == 0 && C / 8 <= 1024, "GroupNorm needs C %% 32 == 0 and C <= 8192 (C=%d)", C);
  int threads, rpb, chunks;
  gn_geometry(S, rows_per_sample, C, threads, rpb, chunks);
  gn_apply_kernel<<<dim3(chunks, S), threads, 0, st>>>(x, y, sums, gamma, beta, rows_per_sample, C, C / 32,
                                                     (double)rows_per_sample * (C / 32), eps, rpb, silu_act ? knobs().gn_silu : 0);
  MUDG_CUDA(cudaGetLastError());
}

// slab width for gn_small_kernel: whole groups and whole 128-bit vectors
static int gn_small_slab(int C) {
  const int cpg = C / 32;
  int cs = cpg;
  while (cs % 8) cs += cpg;
  return cs;
}
static size_t gn_small_smem(int64_t rows_per_sample, int Cs) {
  const int rstep = GNS_THREADS / (Cs / 8);
  return (size_t)rows_per_sample * Cs * 2 + (size_t)rstep * Cs * 2 * sizeof(float) + (size_t)Cs * 2 * sizeof(double) + 64 * sizeof(float);
}
bool gn_small_ok(int S, int64_t rows_per_sample, int C) {
  if (C % 32 != 0 || rows_per_sample < 1 || rows_per_sample > 65535) return false;
  const int Cs = gn_small_slab(C);
  if (Cs > GNS_THREADS || Cs / 8 > GNS_THREADS || C % Cs != 0) return false;
  // measured (tests/gpu_ab_knob.py gn_small): -3.1 % on the MDM512 forward; at MDM1024 level 2 (47 MB tensors, bandwidth-bound,
  // the two-pass form re-reads from L2) the one-kernel form is slightly slower, so it is kept to tensors of a few MB
  // -- and to at most 4 slabs per SM (MDM1024 level 3: 1 024 slabs of 11 KB, +0.3 % on that forward).
  if ((int64_t)S * rows_per_sample * C * 2 > (int64_t)16 << 20) return false;
  const int64_t ctas = (int64_t)S * (C / Cs);
  return gn_small_smem(rows_per_sample, Cs) <= 160 * 1024 && ctas >= 16 && ctas <= 4 * (int64_t)sm_count();
}
void gn_small(const __half* x, __half* y, int S, int64_t rows_per_sample, int C, const float* gamma, const float* beta,
              float eps, bool silu_act, cudaStream_t st) {
  MUDG_REQUIRE(gn_small_ok(S, rows_per_sample, C), "gn_small: [%d][%lld][%d] does not fit", S, (long long)rows_per_sample, C);
  const int Cs = gn_small_slab(C);
  const size_t smem = gn_small_smem(rows_per_sample, Cs);
  MUDG_CUDA(cudaFuncSetAttribute(gn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  gn_small_kernel<<<dim3(C / Cs, S), GNS_THREADS, smem, st>>>(x, y, gamma, beta, (int)rows_per_sample, C, C / 32, Cs, eps,
                                                              silu_act ? knobs().gn_silu : 0);
  MUDG_CUDA(cudaGetLastError());
}

void gn_fold_weights(const __half* W, const double* sums, int S, int64_t rows_per_sample, const float* gamma,
                     const float* beta, float eps, __half* Ws, float* cs, int N, int K, cudaStream_t st) {
  MUDG_REQUIRE(K % 32 == 0 && K <= 4096, "gn_fold_weights: K = %d", K);
  gn_fold_weights_kernel<<<dim3((N + 7) / 8, S), 256, 2 * K * sizeof(float), st>>>(W, sums, gamma, beta, Ws, cs, N, K, K / 32,
                                                                                (double)rows_per_sample * (K / 32), eps);
  MUDG_CUDA(cudaGetLastError());
}

void layernorm(const __half* x, __half* y, const float* gamma, const float* beta, int64_t rows, int C, float eps,
               cudaStream_t st) {
  MUDG_REQUIRE(C % 8 == 0 && C <= 2560, "LayerNorm width %d unsupported", C);
  const int wpb = 8;
  const int64_t blocks = (rows + wpb - 1) / wpb;
  const int vecs = C / 8;
  if (vecs <= 64) layernorm_kernel<2><<<(unsigned)blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  else if (vecs <= 160) layernorm_kernel<5><<<(unsigned)blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  else layernorm_kernel<10><<<(unsigned)blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  MUDG_CUDA(cudaGetLastError());
}

void concat_channels(const __half* a, int Ca, const __half* b, int Cb, __half* out, int64_t rows, int64_t rows_b,
                     cudaStream_t st) {
  MUDG_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "concat needs C %% 8 == 0");
  const int64_t total = rows * ((Ca + Cb) / 8);
  concat_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                      reinterpret_cast<uint4*>(out), rows, rows_b, Ca / 8, Cb / 8);
  MUDG_CUDA(cudaGetLastError());
}

void concat_channels_stats(const __half* a, int Ca, const __half* b, int Cb, __half* out, int S, int64_t rows_per_sample,
                           int64_t rows_b, double* sums, cudaStream_t st) {
  const int C = Ca + Cb;
  MUDG_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0 && C % 32 == 0 && C / 8 <= 1024, "concat + GroupNorm statistics: channels %d + %d", Ca, Cb);
  int threads, rpb, chunks;
  gn_geometry(S, rows_per_sample, C, threads, rpb, chunks);
  concat_stats_kernel<<<dim3(chunks, S), threads, 0, st>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                        reinterpret_cast<uint4*>(out), sums, rows_per_sample, rows_b, Ca / 8,
                                                        Cb / 8, C / 32, rpb);
  MUDG_CUDA(cudaGetLastError());
}

void upsample2x(const __half* x, __half* y, int F, int H, int W, int C, cudaStream_t st) {
  const int64_t total = (int64_t)F * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), F,
                                                          H, W, C / 8);
  MUDG_CUDA(cudaGetLastError());
}

void im2col_s2(const __half* x, __half* y, int F, int H, int W, int C, int pad, cudaStream_t st) {
  const int Ho = (H + pad - 2) / 2 + 1, Wo = (W + pad - 2) / 2 + 1;
  const int64_t total = (int64_t)F * Ho * Wo * 9 * (C / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), F,
                                                         H, W, Ho, Wo, C / 8, pad);
  MUDG_CUDA(cudaGetLastError());
}

// exact (erf) GELU in place, fp16 (the Resampler's FeedForward, resampler.py:31-37: nn.GELU())
__global__ void gelu_inplace_kernel(__half* __restrict__ x, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 v = reinterpret_cast<uint4*>(x)[i];
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float2 f = __half22float2(h[k]);
      f.x = 0.5f * f.x * (1.f + erff(f.x * 0.70710678118654752f));
      f.y = 0.5f * f.y * (1.f + erff(f.y * 0.70710678118654752f));
      h[k] = __floats2half2_rn(f.x, f.y);
    }
    reinterpret_cast<uint4*>(x)[i] = v;
  }
}
void gelu_inplace(__half* x, int64_t n, cudaStream_t st) {
  MUDG_REQUIRE(n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "gelu_inplace: alignment");
  const int64_t n8 = n / 8;
  const int blocks = (int)std::min<int64_t>((n8 + 255) / 256, (int64_t)sm_count() * 8);
  gelu_inplace_kernel<<<blocks, 256, 0, st>>>(x, n8);
  MUDG_CUDA(cudaGetLastError());
}

void softmax_rows(__half* x, int64_t rows, int n, cudaStream_t st) {
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(x, rows, n);
  MUDG_CUDA(cudaGetLastError());
}

void pack_weight(const void* src, bool src_fp32, __half* dst, int O, int I, int taps, int Ipad, cudaStream_t st) {
  const int64_t total = (int64_t)O * taps * Ipad;
  if (src_fp32)
    pack_weight_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const float*>(src), dst, O, I, taps, Ipad);
  else
    pack_weight_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, O, I, taps, Ipad);
  MUDG_CUDA(cudaGetLastError());
}

void cast_to_f32(const void* src, bool src_fp32, float* dst, int64_t n, cudaStream_t st) {
  if (src_fp32) cast_kernel<float, float><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const float*>(src), dst, n);
  else cast_kernel<__half, float><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, n);
  MUDG_CUDA(cudaGetLastError());
}

void cast_to_f16(const void* src, bool src_fp32, __half* dst, int64_t n, cudaStream_t st) {
  if (src_fp32) cast_kernel<float, __half><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const float*>(src), dst, n);
  else cast_kernel<__half, __half><<<grid_for(n, 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, n);
  MUDG_CUDA(cudaGetLastError());
}

void geglu_interleave(const __half* w, const float* b, __half* wo, float* bo, int half_rows, int cols, cudaStream_t st) {
  MUDG_REQUIRE(half_rows % 64 == 0, "GEGLU inner dim must be a multiple of 64");
  const int64_t total = (int64_t)2 * half_rows * cols;
  geglu_interleave_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, b, wo, bo, half_rows, cols);
  MUDG_CUDA(cudaGetLastError());
}

void sinusoid3(const int64_t* t, const int64_t* label, const int64_t* fs, float* out, int B, int dim, cudaStream_t st) {
  const int n = B * (dim / 2);
  sinusoid_kernel<<<dim3((n + 127) / 128, 3), 128, 0, st>>>(t, label, fs, out, B, dim);
  MUDG_CUDA(cudaGetLastError());
}
void batched_linear(const LinearBatch& lb, cudaStream_t st) {
  MUDG_REQUIRE(lb.count >= 1 && lb.count <= LinearBatch::MAX_JOBS, "batched_linear: %d jobs", lb.count);
  const int64_t warps = (int64_t)lb.Bn * (lb.sum ? lb.off[1] : lb.off[lb.count]);
  batched_linear_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(lb);
  MUDG_CUDA(cudaGetLastError());
}
void to_channels_last(const void* x, bool x_fp32, __half* y, int B, int C, int64_t R, int Cpad, cudaStream_t st) {
  const int64_t total = (int64_t)B * R * Cpad;
  if (x_fp32)
    to_channels_last_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const float*>(x), y, B, C, R, Cpad);
  else
    to_channels_last_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(x), y, B, C, R, Cpad);
  MUDG_CUDA(cudaGetLastError());
}

void gather_rows_f16(const void* src, bool src_fp32, __half* dst, int B, int src_rows, int r0, int nrows, int cols,
                     cudaStream_t st) {
  const int64_t total = (int64_t)B * nrows * cols;
  if (src_fp32)
    gather_rows_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const float*>(src), dst, B, src_rows, r0, nrows, cols);
  else
    gather_rows_kernel<__half><<<grid_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, B, src_rows, r0, nrows, cols);
  MUDG_CUDA(cudaGetLastError());
}

void from_channels_last(const __half* y, __half* out, int B, int C, int64_t R, int Cp, cudaStream_t st) {
  const int64_t total = (int64_t)B * C * R;
  from_channels_last_kernel<<<grid_for(total, 256), 256, 0, st>>>(y, out, B, C, R, Cp);
  MUDG_CUDA(cudaGetLastError());
}

void ln_stats(const __half* x, float2* out, int64_t rows, int C, float eps, cudaStream_t st) {
  MUDG_REQUIRE(C % 8 == 0 && C <= 2560, "LayerNorm width %d unsupported", C);
  const int wpb = 8;
  const int vecs = C / 8;
  auto blocks = [&](int rpw) { return (unsigned)((rows + (int64_t)wpb * rpw - 1) / ((int64_t)wpb * rpw)); };
  if (vecs <= 64) ln_stats_kernel<2, 4><<<blocks(4), wpb * 32, 0, st>>>(x, out, rows, C, eps);
  else if (vecs <= 160) ln_stats_kernel<5, 2><<<blocks(2), wpb * 32, 0, st>>>(x, out, rows, C, eps);
  else ln_stats_kernel<10, 1><<<blocks(1), wpb * 32, 0, st>>>(x, out, rows, C, eps);
  MUDG_CUDA(cudaGetLastError());
}

void ln_finalize(const float2* parts, int nparts, float2* out, int64_t rows, int C, float eps, cudaStream_t st) {
  MUDG_REQUIRE(nparts > 0 && nparts * 64 == C, "ln_finalize: %d partial planes for width %d", nparts, C);
  const unsigned blocks = (unsigned)((rows + 255) / 256);
  if (nparts % 5 == 0) ln_finalize_kernel<5><<<blocks, 256, 0, st>>>(parts, out, rows, nparts, 1.f / (float)C, eps);
  else ln_finalize_kernel<4><<<blocks, 256, 0, st>>>(parts, out, rows, nparts, 1.f / (float)C, eps);
  MUDG_CUDA(cudaGetLastError());
}

void ln_fold(__half* W, const float* gamma, const float* beta, const float* bias, float* c1, float* c2, int N, int K,
             cudaStream_t st) {
  ln_fold_kernel<<<(N * 32 + 255) / 256, 256, 0, st>>>(W, gamma, beta, bias, c1, c2, N, K);
  MUDG_CUDA(cudaGetLastError());
}

}  // namespace mudg
