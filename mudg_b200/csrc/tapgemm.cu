// Tap GEMM kernels (see gemm.h): tcgen05 / TMA implementations of the one contraction every Linear / Conv2d 3x3 /
// Conv3d (3,1,1) / 1x1 conv of the path maps to.
//
//   tapgemm_tc2_kernel<SUB>  persistent single-CTA kernel (128 x <=128 tiles, SUB M sub-tiles per weight tile)
//   tapgemm_tc3_kernel<EPI>  CTA pair (cta_group::2), 256 x <=256 tiles: the large problems
//   tapgemm_generic_kernel   CUDA cores, arbitrary strides / dtypes: the three irregular layers (< 0.1 % of the FLOPs)
//
// Common structure: warp 0 = TMA producer (per k-step one 5-D box of A: 64 channels x 128 pixels, shifted by the tap
// offset, OOB -> zeros == conv padding; and one box of W), warp 1 = tcgen05.mma issuer (fp16 -> fp32 in TMEM), the other
// warps = epilogue (tcgen05.ld with lane == tile row; bias / per-sample bias / residual / folded LayerNorm / GEGLU /
// GroupNorm partial sums; fp16 pack into swizzled smem; TMA store, which clips partial tiles).
// The CUDA-core checker of the same contract lives in csrc/test/testhooks.cu (tests only).
#include "gemm.h"
#include "ptx.cuh"

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace mudg {

namespace {

constexpr int BM = 128, BK = 64;

// Division by a runtime constant without the ~40-instruction IDIV sequence (Granlund-Montgomery, dividends < 2^31):
// the tile decode runs once per tile in EVERY role of the persistent kernels, on their critical path.
struct FastDiv {
  uint32_t mul, shr;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, d < 1 ? 1 : d};
  if (f.d > 1) {
    int lg = 0;
    while ((int64_t(1) << lg) < f.d) lg++;
    const int pw = 31 + lg;
    f.mul = (uint32_t)(((uint64_t(1) << pw) + (uint64_t)f.d - 1) / (uint64_t)f.d);
    f.shr = (uint32_t)(pw - 32);
  }
  return f;
}
__device__ __forceinline__ int fd_div(const FastDiv& f, int n) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}
__device__ __forceinline__ void fd_divmod(const FastDiv& f, int n, int& q, int& r) {
  q = fd_div(f, n);
  r = n - q * f.d;
}

struct TcParams {
  FastDiv fd_tw, fd_th, fd_tt, fd_nt, fd_b2;   // tiles_w, tiles_h, tiles_t, N tiles, bias2_div
  int ntaps, kchunks, cin;
  int tiles_n, tiles_w, tiles_h, tiles_t;
  int bw, bh, bt, bb;
  int N, n_out;
  int geglu, has_res;
  float alpha;
  const float* bias;
  const float* bias2;
  int bias2_div, nb2, dimT;
  const float2* ln_stats;   // folded LayerNorm: per-row (mean, rstd), or null
  const float* ln_c1;       // [N]
  float2* ln_out;           // per-row, per-64-column-chunk (sum, sum of squares) of the output [n_out / 64][ln_rows], or null
  int64_t ln_rows;
  int8_t taps[9][4];
};

// ================================================================ tap GEMM v2: persistent, double-buffered accumulators
// Two persistent CTAs per SM loop over output tiles (N tiles fastest so co-resident CTAs share A in L2):
//   warp 0   TMA producer  -- 3-stage ring of {A 128x64, B 128x64} per CTA (6 stages in flight per SM)
//   warp 1   MMA issuer    -- accumulates tile t into TMEM buffer t&1 (2 x 128 columns); the MMA N is the tile's real
//                             width (128, or 64 for the last tile of N = 64 mod 128: N = 320 costs 128+128+64)
//   warps 2-5 epilogue     -- drains buffer t&1 while the issuer already works on tile t+1.  Residual rows are
//                             prefetched from global one 32-column slice ahead; 64-column chunks go through a 16 KB
//                             swizzled staging tile and are written with TMA (clips partial tiles).
// TMEM alloc, barrier init and descriptor prefetch are paid once per CTA instead of once per tile.
// SUB = M sub-tiles (128 pixels each) a CTA computes per tile against ONE shared weight tile:
//   SUB 1: 192 threads, 3 stages x 32 KB, 2 CTAs/SM                (small problems, short K)
//   SUB 2: 320 threads, 4 stages x 48 KB, 1 CTA/SM, 8 epilogue warps: 25 % less L2->SM operand traffic per MAC
constexpr int G2_BN_MAX = 128;
constexpr int G2_A_BYTES = BM * BK * 2;                 // 16 KB per sub-tile
constexpr int G2_B_BYTES = G2_BN_MAX * BK * 2;          // 16 KB
// DEEP (SUB 1 only): a 6-stage ring and one CTA per SM for launches with at most one tile per SM -- there the k-loop is bound
// by the bytes in flight (3 stages x 32 KB against ~1 us of L2 / HBM latency is ~100 GB/s per SM: 680 clk per 256-clk k-step).
template <int SUB, int DEEP = 0> struct G2Cfg {
  static constexpr int STAGES = DEEP ? 6 : (SUB == 1 ? 3 : 4);
  static constexpr int STAGE_BYTES = SUB * G2_A_BYTES + G2_B_BYTES;
  static constexpr int STG_BYTES = SUB * 16384;
  static constexpr int SMEM = STAGES * STAGE_BYTES + STG_BYTES + 256;   // barriers (2 STAGES + 4) x 8 B + the TMEM slot; base declared 1024-aligned
  static constexpr int THREADS = 64 + SUB * 128;
  static constexpr int TMEM_COLS = SUB * 256;            // 2 accumulator buffers x SUB x 128 columns
  static constexpr int CTAS_PER_SM = (SUB == 1 && !DEEP) ? 2 : 1;
};
static_assert(2 * (G2Cfg<1>::SMEM + 1024) <= 233472, "two CTAs of the SUB=1 persistent GEMM must fit one SM");
static_assert(G2Cfg<2>::SMEM + 1024 <= 233472, "SUB=2 persistent GEMM must fit one SM");
static_assert(G2Cfg<1, 1>::SMEM + 1024 <= 233472, "deep SUB=1 persistent GEMM must fit one SM");

struct G2Params {
  TcParams b;            // shared fields (taps, bias, ...)
  int nt;                // number of N tiles (all 128 wide except possibly the last)
  int total_tiles;       // N tiles x ceil(M tiles / SUB)
  int m_tiles;
  int tiles_b;
  int dimW, dimH, dimB;  // extents for the residual bounds check
  const __half* R;       // residual (read straight from global / L2)
  int alpha_is_one;
};

// gelu(g) = 0.5 g (1 + erf(g / sqrt 2)) with ONE MUFU op: erfc(a / sqrt 2) = 2^(-a Q(a)) for a = |g|, Q a degree-4 polynomial
// fitted (weighted by a erfc, a in [0, 5.5], increasing beyond) to -log2(erfc(a / sqrt 2)) / a, so
//   gelu(g) = max(g, 0) - 0.5 |g| 2^(-|g| Q(|g|)):  1 MUFU + 9 FMA-pipe instructions, max |err| 1.3e-6 over all g
// (tests/test_oracle_golden.py::test_gelu_q4_polynomial holds the fit against erf in float64).  The GEGLU epilogues are
// co-limited by the MUFU unit and the issue slots: the Abramowitz-Stegun 7.1.25 form used before (t = 1 / (1 + p a), cubic in
// t, times exp(-a^2 / 2): 2 MUFU ops, |err| 2.5e-5) made the level-0 GEGLU GEMM 6 % slower (profiles/r2_gelu_q4_ab.log).
__device__ __forceinline__ float gelu_q4(float g) {
  const float a = fabsf(g);
  float q = fmaf(0.0005244618f, a, -0.0074173124f);
  q = fmaf(q, a, 0.052593093f);
  q = fmaf(q, a, 0.45923585f);
  q = fmaf(q, a, 1.1510944f);
  const float e = ex2_approx(-a * q);
  return fmaf(-0.5f * a, e, fmaxf(g, 0.f));
}

struct G2Tile {
  int n0, bn, w0, h0, t0, b0;
};
// tile = n-tile + nt * m-super-tile; sub-tile `sub` of a super tile is M tile (msuper*SUB + sub).  An M tile index past
// the end decodes to a batch coordinate outside the tensor: its loads are zero-filled and its stores clipped away.
template <int SUB>
__device__ __forceinline__ G2Tile g2_decode(const G2Params& p, int tile, int sub) {
  G2Tile t;
  int nt_i, m, tw, th, tt, tb;
  fd_divmod(p.b.fd_nt, tile, m, nt_i);
  m = m * SUB + sub;
  t.n0 = nt_i * G2_BN_MAX;
  t.bn = min(G2_BN_MAX, p.b.N - t.n0);
  fd_divmod(p.b.fd_tw, m, m, tw);
  fd_divmod(p.b.fd_th, m, m, th);
  fd_divmod(p.b.fd_tt, m, tb, tt);
  t.w0 = tw * p.b.bw; t.h0 = th * p.b.bh; t.t0 = tt * p.b.bt; t.b0 = tb * p.b.bb;
  return t;
}

template <int SUB, int DEEP = 0>
__global__ void __launch_bounds__(G2Cfg<SUB, DEEP>::THREADS, G2Cfg<SUB, DEEP>::CTAS_PER_SM)
tapgemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmD, const G2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();          // SWIZZLE_128B tiles need a 1024 B aligned base
  using Cfg = G2Cfg<SUB, DEEP>;
  constexpr int G2_STAGES = Cfg::STAGES, G2_STAGE_BYTES = Cfg::STAGE_BYTES;
  uint8_t* stg_all = smem + G2_STAGES * G2_STAGE_BYTES;  // 16 KB per epilogue group
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + Cfg::STG_BYTES);
  uint64_t* full = bars;                        // [STAGES]
  uint64_t* empty = bars + G2_STAGES;           // [STAGES]
  uint64_t* tmem_full = bars + 2 * G2_STAGES;   // [2]
  uint64_t* tmem_empty = bars + 2 * G2_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < G2_STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], SUB * 128); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ktotal = p.b.ntaps * p.b.kchunks;

  if (warp == 0) {
    // whole warp runs the loop (uniform registers), one elected lane issues; no divisions / modulo per k-step
    uint32_t s = 0, ph = 1;
    const int ntaps = p.b.ntaps, kchunks = p.b.kchunks;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      G2Tile tl[SUB];
#pragma unroll
      for (int u = 0; u < SUB; u++) tl[u] = g2_decode<SUB>(p, tile, u);
      for (int tap = 0; tap < ntaps; tap++) {
        const int dw = p.b.taps[tap][0], dh = p.b.taps[tap][1], dt = p.b.taps[tap][2];
        int kb = tap * p.b.cin;
        for (int kc = 0; kc < kchunks; kc++, kb += BK) {
          mbar_wait(&empty[s], ph);
          if (elect_one()) {
            mbar_expect_tx(&full[s], G2_STAGE_BYTES);
            uint8_t* a_s = smem + s * G2_STAGE_BYTES;
#pragma unroll
            for (int u = 0; u < SUB; u++)
              tma_load_5d(a_s + u * G2_A_BYTES, &tmA, &full[s], kc * BK, tl[u].w0 + dw, tl[u].h0 + dh, tl[u].t0 + dt,
                          tl[u].b0);
            // rows past N (last tile of N = 64 mod 128) are zero-filled and never multiplied (MMA N = bn)
            tma_load_5d(a_s + SUB * G2_A_BYTES, &tmB, &full[s], kb, tl[0].n0, 0, 0, 0);
          }
          __syncwarp();
          if (++s == G2_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      uint32_t s = 0, ph = 0, lt = 0;
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
        const G2Tile tl = g2_decode<SUB>(p, tile, 0);
        const uint32_t acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t idesc = umma_idesc_f16(BM, tl.bn, 0, 0);
        const uint32_t d_tmem = tmem_base + acc * (SUB * G2_BN_MAX);
        for (int k = 0; k < ktotal; k++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da_s = da0 + (uint64_t)(s * (G2_STAGE_BYTES >> 4));   // start address field: 16 B units
            const uint64_t db = da_s + (uint64_t)((SUB * G2_A_BYTES) >> 4);
#pragma unroll
            for (int u = 0; u < SUB; u++) {
              const uint64_t da = da_s + (uint64_t)((u * G2_A_BYTES) >> 4);
#pragma unroll
              for (int kk = 0; kk < BK / 16; kk++)
                umma_f16(d_tmem + u * G2_BN_MAX, da + 2 * kk, db + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[s]);
            if (k == ktotal - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++s == G2_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    const int grp = (warp - 2) >> 2;        // which M sub-tile this epilogue group drains
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    uint8_t* stg = stg_all + grp * 16384;
    uint8_t* srow = stg + row * 128;
    auto group_sync = [&]() {               // 128 threads of this group only
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    // row -> (iw, ih, it, ib) inside the pixel box
    int r = row;
    const int iw = r % p.b.bw; r /= p.b.bw;
    const int ih = r % p.b.bh; r /= p.b.bh;
    const int it_ = r % p.b.bt;
    const int ib_ = r / p.b.bt;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, lt++) {
      const G2Tile tl = g2_decode<SUB>(p, tile, grp);
      const uint32_t acc = lt & 1;
      const uint32_t t_row = tmem_base + acc * (SUB * G2_BN_MAX) + grp * G2_BN_MAX + lane_off;
      const int pw = tl.w0 + iw, ph = tl.h0 + ih, pt = tl.t0 + it_, pb = tl.b0 + ib_;
      const bool row_ok = pw < p.dimW && ph < p.dimH && pt < p.b.dimT && pb < p.dimB;
      const int64_t pix = (((int64_t)pb * p.b.dimT + pt) * p.dimH + ph) * p.dimW + pw;
      float2 ln_ms = make_float2(0.f, 1.f);
      if (p.b.ln_stats != nullptr && row_ok) ln_ms = __ldg(p.b.ln_stats + pix);
      if (p.b.geglu) {
        mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
        __syncwarp();
        tc_fence_after();
        // weight rows interleaved 64 value / 64 gate: TMEM columns [0,64) value, [64,128) gate -> 64 output columns
        if (issuer) tma_store_wait_read0();
        __syncwarp();
        group_sync();
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
          uint32_t v[32], g[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld32(t_row + 64 + c * 32, g);
          tmem_ld_wait();
          if (c == 1) {                     // all TMEM reads of this accumulator are done
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          const float4* bv = reinterpret_cast<const float4*>(p.b.bias + tl.n0 + c * 32);
          const float4* cv = reinterpret_cast<const float4*>(p.b.ln_c1 + tl.n0 + c * 32);
          uint32_t o[16];
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            float4 b_v = make_float4(0.f, 0.f, 0.f, 0.f), b_g = b_v;
            if (p.b.bias != nullptr) { b_v = __ldg(bv + i4); b_g = __ldg(bv + 16 + i4); }
            // alpha (plain scale) or the folded LayerNorm: rstd * (acc - mean * c1[n]) + c2[n]
            float al = p.b.alpha;
            float4 s_v = make_float4(0.f, 0.f, 0.f, 0.f), s_g = s_v;
            if (p.b.ln_stats != nullptr) {
              al = ln_ms.y;
              const float4 c_v = __ldg(cv + i4), c_g = __ldg(cv + 16 + i4);
              const float k = -ln_ms.x * ln_ms.y;
              s_v = make_float4(k * c_v.x, k * c_v.y, k * c_v.z, k * c_v.w);
              s_g = make_float4(k * c_g.x, k * c_g.y, k * c_g.z, k * c_g.w);
            }
            const float v0 = fmaf(__uint_as_float(v[4 * i4]), al, b_v.x + s_v.x), v1 = fmaf(__uint_as_float(v[4 * i4 + 1]), al, b_v.y + s_v.y);
            const float v2 = fmaf(__uint_as_float(v[4 * i4 + 2]), al, b_v.z + s_v.z), v3 = fmaf(__uint_as_float(v[4 * i4 + 3]), al, b_v.w + s_v.w);
            const float g0 = fmaf(__uint_as_float(g[4 * i4]), al, b_g.x + s_g.x), g1 = fmaf(__uint_as_float(g[4 * i4 + 1]), al, b_g.y + s_g.y);
            const float g2 = fmaf(__uint_as_float(g[4 * i4 + 2]), al, b_g.z + s_g.z), g3 = fmaf(__uint_as_float(g[4 * i4 + 3]), al, b_g.w + s_g.w);
            o[2 * i4] = pack_half2(v0 * gelu_q4(g0), v1 * gelu_q4(g1));
            o[2 * i4 + 1] = pack_half2(v2 * gelu_q4(g2), v3 * gelu_q4(g3));
          }
#pragma unroll
          for (int j = 0; j < 4; j++)
            *reinterpret_cast<uint4*>(srow + (((c * 4 + j) ^ (row & 7)) << 4)) =
                make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
        fence_proxy_async_smem();
        group_sync();
        if (issuer) {
          tma_store_5d(&tmD, stg, tl.n0 / 2, tl.w0, tl.h0, tl.t0, tl.b0);
          tma_store_commit();
        }
        __syncwarp();
        continue;
      }
      int sample = 0;
      if (p.b.bias2 != nullptr) {
        sample = fd_div(p.b.fd_b2, pb * p.b.dimT + pt);
        if (sample >= p.b.nb2) sample = p.b.nb2 - 1;
      }
      const uint4* rp = (p.R != nullptr && row_ok) ? reinterpret_cast<const uint4*>(p.R + pix * p.b.n_out + tl.n0) : nullptr;
      if (p.R != nullptr && tile + (int)gridDim.x < p.total_tiles) {
        // residual rows of this CTA's NEXT tile: pull them from HBM into L2 now, a whole mainloop ahead of their use
        const G2Tile nx = g2_decode<SUB>(p, tile + gridDim.x, grp);
        const int nw = nx.w0 + iw, nh_ = nx.h0 + ih, nt_ = nx.t0 + it_, nb_ = nx.b0 + ib_;
        if (nw < p.dimW && nh_ < p.dimH && nt_ < p.b.dimT && nb_ < p.dimB) {
          const int64_t npix = (((int64_t)nb_ * p.b.dimT + nt_) * p.dimH + nh_) * p.dimW + nw;
          const __half* np_ = p.R + npix * p.b.n_out + nx.n0;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(np_));
          if (nx.bn > 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + 64));
        }
      }
      uint4 res_nxt[4];
      if (rp != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; j++) res_nxt[j] = __ldg(rp + j);
      }
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const int nh = tl.bn >> 5;             // 32-column slices (2 or 4)
#pragma unroll 1
      for (int hf = 0; hf < nh; hf++) {
        uint4 res[4];
        if (rp != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; j++) res[j] = res_nxt[j];
          if (hf + 1 < nh) {
#pragma unroll
            for (int j = 0; j < 4; j++) res_nxt[j] = __ldg(rp + (hf + 1) * 4 + j);
          }
        }
        uint32_t v[32];
        tmem_ld32(t_row + hf * 32, v);
        if ((hf & 1) == 0) {                 // staging tile must be free: previous TMA store has read it
          if (issuer) tma_store_wait_read0();
          __syncwarp();
          group_sync();
        }
        tmem_ld_wait();
        if (hf == nh - 1) {                  // all TMEM reads of this accumulator are done
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        const int col0 = tl.n0 + hf * 32;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; i++) f[i] = __uint_as_float(v[i]);
        if (!p.alpha_is_one) {
#pragma unroll
          for (int i = 0; i < 32; i++) f[i] *= p.b.alpha;
        }
        if (p.b.ln_stats != nullptr) {      // folded LayerNorm: rstd * (acc - mean * c1[n]); c2 arrives as the bias
          const float4* cp = reinterpret_cast<const float4*>(p.b.ln_c1 + col0);
          const float k = -ln_ms.x * ln_ms.y;
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 c4 = __ldg(cp + i4);
            f[4 * i4] = fmaf(f[4 * i4], ln_ms.y, k * c4.x); f[4 * i4 + 1] = fmaf(f[4 * i4 + 1], ln_ms.y, k * c4.y);
            f[4 * i4 + 2] = fmaf(f[4 * i4 + 2], ln_ms.y, k * c4.z); f[4 * i4 + 3] = fmaf(f[4 * i4 + 3], ln_ms.y, k * c4.w);
          }
        }
        if (p.b.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(p.b.bias + col0);
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 b4 = __ldg(bp + i4);
            f[4 * i4] += b4.x; f[4 * i4 + 1] += b4.y; f[4 * i4 + 2] += b4.z; f[4 * i4 + 3] += b4.w;
          }
        }
        if (p.b.bias2 != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(p.b.bias2 + (size_t)sample * p.b.N + col0);
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 b4 = __ldg(bp + i4);
            f[4 * i4] += b4.x; f[4 * i4 + 1] += b4.y; f[4 * i4 + 2] += b4.z; f[4 * i4 + 3] += b4.w;
          }
        }
        if (rp != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float2 r0 = unpack_half2(res[j].x), r1 = unpack_half2(res[j].y), r2 = unpack_half2(res[j].z),
                         r3 = unpack_half2(res[j].w);
            f[8 * j + 0] += r0.x; f[8 * j + 1] += r0.y; f[8 * j + 2] += r1.x; f[8 * j + 3] += r1.y;
            f[8 * j + 4] += r2.x; f[8 * j + 5] += r2.y; f[8 * j + 6] += r3.x; f[8 * j + 7] += r3.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<uint4*>(srow + ((((hf & 1) * 4 + j) ^ (row & 7)) << 4)) =
              make_uint4(pack_half2(f[8 * j + 0], f[8 * j + 1]), pack_half2(f[8 * j + 2], f[8 * j + 3]),
                         pack_half2(f[8 * j + 4], f[8 * j + 5]), pack_half2(f[8 * j + 6], f[8 * j + 7]));
        if (hf & 1) {                        // a 64-column chunk is complete
          fence_proxy_async_smem();
          group_sync();
          if (issuer) {
            tma_store_5d(&tmD, stg, tl.n0 + (hf >> 1) * 64, tl.w0, tl.h0, tl.t0, tl.b0);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (issuer) tma_store_wait_read0();
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ================================================================ tap GEMM v3: CTA pair, 256 x (<=256) tiles
// A cluster of two CTAs (one TPC) works on one output tile with tcgen05.mma.cta_group::2 (M = 256):
//   CTA r owns M sub-tile r (128 pixels): it TMA-loads its A box and HALF of the weight tile (bn/2 rows) per k-step,
//   so the pair moves 16 KB of operands per 1M MACs from L2 -- 2/3 of the 256x128 single-CTA tile, half of 128x128;
//   the leader's elected thread issues the MMAs for both SMs; each CTA's 128 x bn fp32 accumulator lives in its own TMEM
//   (2 buffers x 256 columns); commits are multicast to both CTAs' barriers.
//   warp 0 producer (both CTAs; bytes are signalled on the LEADER's full barrier), warp 1 MMA issuer (leader only),
//   warps 2-5 / 6-9: epilogue groups draining columns [0,128) / [128,256) of the CTA's rows (same fused epilogue as v2).
// N is split into balanced tiles of whole 64-column units (128 for GEGLU): 320 -> 192+128, 640 -> 256+192+192.
constexpr int G3_STAGES = 5;                              // NG 2 (see below); NG 3 runs 4 stages
constexpr int G3_A_BYTES = BM * BK * 2;                   // 16 KB: this CTA's 128 pixels x 64 channels
constexpr int G3_B_BYTES = 128 * BK * 2;                  // 16 KB: up to 128 weight rows (half of the N tile)
constexpr int G3_STAGE_BYTES = G3_A_BYTES + G3_B_BYTES;
// NG = epilogue groups of 4 warps, each with two 16 KB staging tiles (ping-pong).
//   NG 2: 320 threads, 5-stage operand ring: the long-K layers, whose epilogue hides behind the MMA main loop.
//   NG 3: 512 threads -- warps 0-3 = {producer, issuer, 2 idle} shrink to 40 registers (setmaxnreg), the 12 epilogue warps grow
//         to 152 --, 4-stage ring: the short-K Linear layers (K = 320 / 640), whose latency-bound epilogue (IPC 0.25) is the
//         bottleneck at 4 700 clk per 256 x 256 tile against 1 300-2 600 clk of MMA.  Measured INSIDE the power-capped clip
//         (SM clock 1.58 GHz): QKV 320->960 260 -> 210 us, GEGLU 320->2560 783 -> 686 us, 320->320+res 169 -> 156 us; in
//         isolation at boost clocks the third group is a LOSS (profiles/r2_gemm_groups_ab.log), and so it is for the 3-tap
//         temporal convs and the no-residual K = 640 layers in the clip -- hence the narrow selection rule in tapgemm_tc3().
template <int NG> struct G3Cfg {
  static constexpr int STAGES = NG == 2 ? G3_STAGES : 4;
  static constexpr int STG_BYTES = NG * 2 * 16384;
  static constexpr int SMEM = STAGES * G3_STAGE_BYTES + STG_BYTES + 256;
  static constexpr int THREADS = NG == 2 ? 64 + 2 * 128 : 128 + 3 * 128;
  static constexpr int EW0 = NG == 2 ? 2 : 4;          // first epilogue warp
};
static_assert(G3Cfg<2>::SMEM <= 232448 && G3Cfg<3>::SMEM <= 232448, "pair GEMM must fit one SM");

struct G3Params {
  TcParams b;
  int nt;                       // N tiles
  int n_unit, n_q, n_rem;       // tile i spans (n_q + (i < n_rem)) units of n_unit columns
  FastDiv rot_div;              // d = cluster count when it is a multiple of nt (N-tile rotation), else d = 0
  int bn_first;                 // width of tile 0 (tmB0's box is bn_first/2 rows; narrower tiles use tmB1)
  int total_tiles;              // nt x ceil(M tiles / 2)
  int tiles_b;
  int dimW, dimH, dimB;
  const __half* R;
  int res_mma;                  // the residual is added by the tensor core (tmR x identity, see the producer): R is null then
  int alpha_is_one;
  double* gn_sums;              // GroupNorm statistics of the output ([S][32][2] fp64, pre-zeroed), or null
  FastDiv fd_cpg, fd_gn;        // channels per group; frames per GroupNorm sample (sample = (b*T + t) / d)
  FastDiv fd_ws;                // per-sample weights: weight matrix of a tile = (b0*T + t0) / d; d = 0: one matrix
  int dbg;                      // debug experiments (knob gemm_dbg): 1 = do not issue the output TMA stores
  long long* trace;             // debug (tests/gpu_trace_gemm.py): clock64 time line of cluster 0 / rank 0, [4 roles][64 tiles][8]
};

#define G3_TRACE(role, lt_, ev)                                                                       \
  do {                                                                                                 \
    if (p.trace != nullptr && blockIdx.x == 0 && (lt_) < 64) p.trace[((role) * 64 + (lt_)) * 8 + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ G2Tile g3_decode(const G3Params& p, int tile, int rank) {
  G2Tile t;
  int nt_i, m, tw, th, tt, tb;
  fd_divmod(p.b.fd_nt, tile, m, nt_i);
  if (p.rot_div.d > 0) {           // cluster count is a multiple of nt: rotate so a cluster does not keep one N tile width
    nt_i += fd_div(p.rot_div, tile);
    nt_i -= fd_div(p.b.fd_nt, nt_i) * p.nt;
  }
  m = m * 2 + rank;
  t.bn = (p.n_q + (nt_i < p.n_rem ? 1 : 0)) * p.n_unit;
  t.n0 = (nt_i * p.n_q + min(nt_i, p.n_rem)) * p.n_unit;
  fd_divmod(p.b.fd_tw, m, m, tw);
  fd_divmod(p.b.fd_th, m, m, th);
  fd_divmod(p.b.fd_tt, m, tb, tt);
  t.w0 = tw * p.b.bw; t.h0 = th * p.b.bh; t.t0 = tt * p.b.bt; t.b0 = tb * p.b.bb;
  return t;
}

// GroupNorm statistics of one staged 64-column output chunk (fp16, SWIZZLE_128B, 128 rows x 128 B), fused into the
// producing GEMM so the consumer norm needs no statistics pass over HBM.  Warp q of an epilogue group sums the 32 rows it
// staged itself: lane l owns columns (2l, 2l+1) -- a row is one conflict-free 128 B shared-memory wavefront --, fp32 partials
// over <= 32 rows, a fixed-order segmented shuffle reduction over the lanes of a group (channels per group is even, so a
// column pair never straddles groups), then one fp64 RED per (group, moment): the same accumulation discipline as
// gn_stats_kernel (fp64 sums of fixed-order fp32 partials), hence the same run-to-run repeatability.
__device__ __forceinline__ void gn_chunk_stats(const uint8_t* stg, int q, int lane, uint32_t valid, int col_base,
                                               const FastDiv& fd_cpg, double* sums_sample) {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const uint32_t coff = (uint32_t)(lane & 3) * 4u, chunk = (uint32_t)lane >> 2;
  const uint8_t* base = stg + q * 32 * 128;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    if (valid & (1u << i)) {                 // warp-uniform: rows past the tensor edge hold bias-only garbage
      const uint32_t w = *reinterpret_cast<const uint32_t*>(base + i * 128 + ((chunk ^ (uint32_t)(i & 7)) << 4) + coff);
      const float2 f = unpack_half2(w);
      s0 += f.x; q0 = fmaf(f.x, f.x, q0);
      s1 += f.y; q1 = fmaf(f.y, f.y, q1);
    }
  }
  float sm = s0 + s1, sq = q0 + q1;
  const int g = fd_div(fd_cpg, col_base + 2 * lane);
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float so = __shfl_down_sync(0xffffffffu, sm, d), qo = __shfl_down_sync(0xffffffffu, sq, d);
    const int go = __shfl_down_sync(0xffffffffu, g, d);
    if (lane + d < 32 && go == g) { sm += so; sq += qo; }
  }
  const int gprev = __shfl_up_sync(0xffffffffu, g, 1);
  if (lane == 0 || gprev != g) {
    atomicAdd(sums_sample + 2 * g, (double)sm);
    atomicAdd(sums_sample + 2 * g + 1, (double)sq);
  }
}

// EPI: epilogue specialisation.  -1 = everything decided at run time (and the only variant with the GEGLU path);
// >= 0: bit 0 bias, bit 1 per-sample bias, bit 2 residual, bit 3 folded LayerNorm, alpha == 1, no GEGLU.  Tested from the
// kernel-parameter bank, every `if (p.b.bias ...)` in the 32-column slice loop is a dependent LDCU -> UISETP -> BRA.U chain of
// 60-80 clocks that a lone, latency-bound epilogue warp cannot overlap with anything (clock64 trace: ~500 clk per slice).
template <int EPI, int NG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G3Cfg<NG>::THREADS, 1)
tapgemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                   const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmD,
                   const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmJ, const G3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  constexpr int NST = G3Cfg<NG>::STAGES, EW0 = G3Cfg<NG>::EW0;
  uint8_t* stg_all = smem + NST * G3_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + G3Cfg<NG>::STG_BYTES);
  uint64_t* full = bars;                        // [STAGES]  used in the leader CTA only
  uint64_t* empty = bars + NST;                 // [STAGES]  one per CTA, signalled by the multicast commit
  uint64_t* tmem_full = bars + 2 * NST;         // [2]       one per CTA, multicast commit
  uint64_t* tmem_empty = bars + 2 * NST + 2;    // [2]  leader only: 2 CTAs x NG epilogue groups arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * NG); }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB0);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmD);
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ktotal = p.b.ntaps * p.b.kchunks;
  // NG 3: register re-partition between the warpgroups.  Each warpgroup executes its setmaxnreg at the top of a region
  // that does not merge with the other's before the end of the kernel, so ptxas budgets the two regions separately.
  if (warp < EW0) {
  if (NG == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // The whole warp runs the loop so that addresses and coordinates stay in uniform registers; one elected lane
    // issues.  The body runs once per k-step (every 256-512 tensor clocks): no divisions, no modulo.
    uint32_t s = 0, ph = 1;                 // stage, parity to wait for on its "empty" barrier
    const uint32_t full0 = mapa_u32(&full[0], 0);
    const int ntaps = p.b.ntaps, kchunks = p.b.kchunks;
    uint32_t plt = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, plt++) {
      const G2Tile tl = g3_decode(p, tile, (int)rank);
      if (lane == 0) G3_TRACE(0, plt, 0);                 // producer starts the tile
      const CUtensorMap* mb = (tl.bn == p.bn_first) ? &tmB0 : &tmB1;
      const int brow = tl.n0 + (int)rank * (tl.bn >> 1);
      const int wsmp = p.fd_ws.d > 0 ? fd_div(p.fd_ws, tl.b0 * p.b.dimT + tl.t0) : 0;   // per-sample weight matrix
      const uint32_t stage_tx = 2u * (uint32_t)(G3_A_BYTES + (tl.bn >> 1) * (BK * 2));
      for (int tap = 0; tap < ntaps; tap++) {
        const int cw = tl.w0 + p.b.taps[tap][0], ch = tl.h0 + p.b.taps[tap][1], ct = tl.t0 + p.b.taps[tap][2];
        int kb = tap * p.b.cin;
        for (int kc = 0; kc < kchunks; kc++, kb += BK) {
          mbar_wait(&empty[s], ph);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full[s], stage_tx);     // bytes of BOTH CTAs' boxes
            uint8_t* a_s = smem + s * G3_STAGE_BYTES;
            tma_load_5d_pair(a_s, &tmA, full0 + 8u * s, kc * BK, cw, ch, ct, tl.b0);
            tma_load_5d_pair(a_s + G3_A_BYTES, mb, full0 + 8u * s, kb, brow, wsmp, 0, 0);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
      if (p.res_mma) {
        // Residual through the tensor core: per 64-column chunk one more k-step whose A operand is the residual tile
        // R[rows, n0 + 64 c .. + 64) and whose B operand is a 64 x 64 identity, accumulated into accumulator columns
        // [64 c, 64 c + 64).  The residual then rides the TMA ring (requested stages ahead) instead of being fetched by the
        // epilogue threads, whose loads took 2 000-3 000 clk under load (clock64 trace, profiles/r2_experiments.md).
        const uint32_t rtx = 2u * (uint32_t)(G3_A_BYTES + 32 * (BK * 2));
        for (int c = 0; c < (tl.bn >> 6); c++) {
          mbar_wait(&empty[s], ph);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full[s], rtx);
            uint8_t* a_s = smem + s * G3_STAGE_BYTES;
            tma_load_5d_pair(a_s, &tmR, full0 + 8u * s, tl.n0 + c * 64, tl.w0, tl.h0, tl.t0, tl.b0);
            tma_load_5d_pair(a_s + G3_A_BYTES, &tmJ, full0 + 8u * s, 0, (int)rank * 32, 0, 0, 0);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
      }
      if (lane == 0) G3_TRACE(0, plt, 1);                 // all loads of the tile issued
    }
  } else if (warp == 1) {
    if (rank == 0) {                         // whole warp loops (uniform registers), one elected lane issues
      uint32_t s = 0, ph = 0, lt = 0;
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + G3_A_BYTES, 16, 1024);
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, lt++) {
        const G2Tile tl = g3_decode(p, tile, 0);
        const uint32_t acc = lt & 1;
        if (lane == 0) G3_TRACE(1, lt, 0);                     // issuer ready for the tile
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);      // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        if (lane == 0) G3_TRACE(1, lt, 1);                     // accumulator free
        const uint32_t idesc = umma_idesc_f16(256, tl.bn, 0, 0);
        const uint32_t idesc_res = umma_idesc_f16(256, 64, 0, 0);     // residual k-steps: 64 accumulator columns at a time
        const uint32_t d_tmem = tmem_base + acc * 256;
        const int ksteps = ktotal + (p.res_mma ? (tl.bn >> 6) : 0);
        for (int k = 0; k < ksteps; k++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (k == 0 && lane == 0) G3_TRACE(1, lt, 2);         // first operands landed
          if (elect_one()) {
            const uint64_t so = (uint64_t)(s * (G3_STAGE_BYTES >> 4));     // descriptor start address: 16 B units
            if (k < ktotal) {
#pragma unroll
              for (int kk = 0; kk < BK / 16; kk++)
                umma_f16_pair(d_tmem, da0 + so + 2 * kk, db0 + so + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
            } else {
              const uint32_t d_res = d_tmem + 64u * (uint32_t)(k - ktotal);
#pragma unroll
              for (int kk = 0; kk < BK / 16; kk++) umma_f16_pair(d_res, da0 + so + 2 * kk, db0 + so + 2 * kk, idesc_res, 1u);
            }
            umma_commit_pair(&empty[s], 3);
            if (k == ksteps - 1) umma_commit_pair(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++s == NST) { s = 0; ph ^= 1u; }
        }
        if (lane == 0) G3_TRACE(1, lt, 3);                     // last MMA of the tile issued
      }
    }
  }
  } else {
    if (NG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // NG epilogue groups of 4 warps (a warp may only touch TMEM lanes 32*(warp%4)..+32, so each group spans all 128
    // rows).  The tile's 64-column output chunks are dealt round-robin to the groups across tiles (a running chunk
    // counter), so chunk counts that are no multiple of NG (bn = 192) do not leave one group with more work.
    const int grp = (warp - EW0) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool issuer = (((warp - EW0) & 3) == 0 && lane == 0);
    uint8_t* stg_grp = stg_all + grp * 32768;   // two staging tiles: this group's n-th chunk uses tile n & 1
    uint32_t chunk_no = 0;                      // chunks this group has stored
    uint32_t cc = 0;                            // chunks of all tiles so far (all groups), modulo NG
    auto group_sync = [&]() {
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else if (grp == 1) asm volatile("bar.sync 2, 128;" ::: "memory");
      else asm volatile("bar.sync 3, 128;" ::: "memory");
    };
    int r = row;
    const int iw = r % p.b.bw; r /= p.b.bw;
    const int ih = r % p.b.bh; r /= p.b.bh;
    const int it_ = r % p.b.bt;
    const int ib_ = r / p.b.bt;
    constexpr bool kGeneric = EPI < 0;
    const bool f_geglu = kGeneric && p.b.geglu;
    const bool f_bias = kGeneric ? (p.b.bias != nullptr) : ((EPI & 1) != 0);
    const bool f_bias2 = kGeneric ? (p.b.bias2 != nullptr) : ((EPI & 2) != 0);
    const bool f_res = kGeneric ? (p.R != nullptr) : ((EPI & 4) != 0);
    const bool f_ln = kGeneric ? (p.b.ln_stats != nullptr) : ((EPI & 8) != 0);
    const bool f_alpha = kGeneric && !p.alpha_is_one;
    const int csh = f_geglu ? 7 : 6;            // accumulator columns per chunk: 64, or 128 (64 value + 64 gate)
    auto arrive_empty = [&](uint32_t bar) {     // MUDG_GEMM_DBG bit 1: fall back to the (slow) release arrive
      if (p.dbg & 2) mbar_arrive_cluster(bar);
      else mbar_arrive_cluster_relaxed(bar);
    };
    uint32_t lt = 0;
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, lt++) {
      const G2Tile tl = g3_decode(p, tile, (int)rank);
      const uint32_t acc = lt & 1;
      const uint32_t empty_bar = mapa_u32(&tmem_empty[acc], 0);
      const int nchunks = tl.bn >> csh;
      const int first = (grp + NG - (int)cc) % NG;             // my chunks: first, first + NG, ...
      const int mine = first < nchunks ? (nchunks - first + NG - 1) / NG : 0;
      cc = (cc + (uint32_t)nchunks) % NG;
      if (mine == 0) {
        mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
        if (issuer) arrive_empty(empty_bar);
        __syncwarp();
        continue;
      }
      const uint32_t t_row = tmem_base + acc * 256 + lane_off;
      const int pw = tl.w0 + iw, ph = tl.h0 + ih, pt = tl.t0 + it_, pb = tl.b0 + ib_;
      const bool row_ok = pw < p.dimW && ph < p.dimH && pt < p.b.dimT && pb < p.dimB;
      const int64_t pix = (((int64_t)pb * p.b.dimT + pt) * p.dimH + ph) * p.dimW + pw;
      float2 ln_ms = make_float2(0.f, 1.f);
      if (f_ln && row_ok) ln_ms = __ldg(p.b.ln_stats + pix);
      if (f_geglu) {
        mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int ci = 0; ci < mine; ci++) {
          const int ch = first + NG * ci;                      // chunk: accumulator columns [128 ch, 128 ch + 128)
          const int nbase = tl.n0 + ch * 128;
          uint8_t* stg = stg_grp + (chunk_no & 1) * 16384;
          uint8_t* srow = stg + row * 128;
          chunk_no++;
#pragma unroll 1
          for (int c = 0; c < 2; c++) {
            uint32_t v[32], g[32];
            tmem_ld32(t_row + ch * 128 + c * 32, v);
            tmem_ld32(t_row + ch * 128 + 64 + c * 32, g);
            tmem_ld_wait();
            if (c == 1 && ci == mine - 1) tc_fence_before();
            const float4* bv = reinterpret_cast<const float4*>(p.b.bias + nbase + c * 32);
            const float4* cv = reinterpret_cast<const float4*>(p.b.ln_c1 + nbase + c * 32);
            uint32_t o[16];
            if (p.b.ln_stats != nullptr) {     // folded LayerNorm: rstd * acc + (k * c1[n] + c2[n]),  k = -mean * rstd
              const float al = ln_ms.y, k = -ln_ms.x * ln_ms.y;
#pragma unroll
              for (int i4 = 0; i4 < 8; i4++) {
                const float4 b_v = __ldg(bv + i4), b_g = __ldg(bv + 16 + i4);
                const float4 c_v = __ldg(cv + i4), c_g = __ldg(cv + 16 + i4);
                const float v0 = fmaf(__uint_as_float(v[4 * i4]), al, fmaf(k, c_v.x, b_v.x));
                const float v1 = fmaf(__uint_as_float(v[4 * i4 + 1]), al, fmaf(k, c_v.y, b_v.y));
                const float v2 = fmaf(__uint_as_float(v[4 * i4 + 2]), al, fmaf(k, c_v.z, b_v.z));
                const float v3 = fmaf(__uint_as_float(v[4 * i4 + 3]), al, fmaf(k, c_v.w, b_v.w));
                const float g0 = fmaf(__uint_as_float(g[4 * i4]), al, fmaf(k, c_g.x, b_g.x));
                const float g1 = fmaf(__uint_as_float(g[4 * i4 + 1]), al, fmaf(k, c_g.y, b_g.y));
                const float g2 = fmaf(__uint_as_float(g[4 * i4 + 2]), al, fmaf(k, c_g.z, b_g.z));
                const float g3 = fmaf(__uint_as_float(g[4 * i4 + 3]), al, fmaf(k, c_g.w, b_g.w));
                o[2 * i4] = pack_half2(v0 * gelu_q4(g0), v1 * gelu_q4(g1));
                o[2 * i4 + 1] = pack_half2(v2 * gelu_q4(g2), v3 * gelu_q4(g3));
              }
            } else {
              const float al = p.b.alpha;
#pragma unroll
              for (int i4 = 0; i4 < 8; i4++) {
                float4 b_v = make_float4(0.f, 0.f, 0.f, 0.f), b_g = b_v;
                if (p.b.bias != nullptr) { b_v = __ldg(bv + i4); b_g = __ldg(bv + 16 + i4); }
                const float v0 = fmaf(__uint_as_float(v[4 * i4]), al, b_v.x), v1 = fmaf(__uint_as_float(v[4 * i4 + 1]), al, b_v.y);
                const float v2 = fmaf(__uint_as_float(v[4 * i4 + 2]), al, b_v.z), v3 = fmaf(__uint_as_float(v[4 * i4 + 3]), al, b_v.w);
                const float g0 = fmaf(__uint_as_float(g[4 * i4]), al, b_g.x), g1 = fmaf(__uint_as_float(g[4 * i4 + 1]), al, b_g.y);
                const float g2 = fmaf(__uint_as_float(g[4 * i4 + 2]), al, b_g.z), g3 = fmaf(__uint_as_float(g[4 * i4 + 3]), al, b_g.w);
                o[2 * i4] = pack_half2(v0 * gelu_q4(g0), v1 * gelu_q4(g1));
                o[2 * i4 + 1] = pack_half2(v2 * gelu_q4(g2), v3 * gelu_q4(g3));
              }
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
              *reinterpret_cast<uint4*>(srow + (((c * 4 + j) ^ (row & 7)) << 4)) =
                  make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read0();   // my previous chunk's store has read the OTHER staging tile
          __syncwarp();
          group_sync();                         // every thread of the group is past its TMEM reads and smem writes
          if (issuer) {
            if (ci == mine - 1) arrive_empty(empty_bar);
            tma_store_5d(&tmD, stg, nbase / 2, tl.w0, tl.h0, tl.t0, tl.b0);
            tma_store_commit();
          }
          __syncwarp();
        }
        continue;
      }
      int sample = 0;
      if (f_bias2) {
        sample = fd_div(p.b.fd_b2, pb * p.b.dimT + pt);
        if (sample >= p.b.nb2) sample = p.b.nb2 - 1;
      }
      // residual row of this thread: chunk ch covers 8 uint4 (64 columns) starting at rp + 8 ch
      const uint4* rp = (f_res && row_ok) ? reinterpret_cast<const uint4*>(p.R + pix * p.b.n_out + tl.n0) : nullptr;
      if (f_res && tile + n_clusters < p.total_tiles) {
        // residual rows of this CTA's NEXT tile: pull them from HBM into L2 a whole mainloop ahead of their use
        // (128 B line = one chunk; the two groups take alternate lines)
        const G2Tile nx = g3_decode(p, tile + n_clusters, (int)rank);
        const int nw = nx.w0 + iw, nh_ = nx.h0 + ih, nt_ = nx.t0 + it_, nb_ = nx.b0 + ib_;
        if (nw < p.dimW && nh_ < p.dimH && nt_ < p.b.dimT && nb_ < p.dimB) {
          const int64_t npix = (((int64_t)nb_ * p.b.dimT + nt_) * p.dimH + nh_) * p.dimW + nw;
          const __half* np_ = p.R + npix * p.b.n_out + nx.n0;
          for (int c = grp; c < (nx.bn >> 6); c += NG) asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + c * 64));
        }
      }
      // Residual rows come straight from global memory (L2 after the prefetch above); a slice's 64 B per thread must be in
      // flight long before it is used.  kResDeep (NG 3, usually ONE chunk per group and tile): the whole 128 B chunk row
      // is requested before the wait for the accumulator and each half is re-requested for the next chunk as soon as it has
      // been consumed -- with a one-slice look-ahead the second slice waited 1 500+ clk on L2 (clock64 trace: 4 000-5 800
      // clk per chunk against 1 300-1 800 for the same chunk without residual).
      constexpr bool kResDeep = NG == 3;
      uint4 res_nxt[kResDeep ? 8 : 4];
      if (rp != nullptr) {
#pragma unroll
        for (int j = 0; j < (kResDeep ? 8 : 4); j++) res_nxt[j] = __ldg(rp + first * 8 + j);
      }
      if (grp < 2 && q == 2 && lane == 0) G3_TRACE(2 + grp, lt, 0);       // epilogue group ready for the tile
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      if (grp < 2 && q == 2 && lane == 0) G3_TRACE(2 + grp, lt, 1);       // accumulator complete
      const int nslices = mine * 2;          // my 32-column slices, two per chunk
      // Two register sets for the accumulator slices: the tcgen05.ld of slice s+1 is in flight while slice s is converted
      // and staged (tcgen05.wait::ld covers every outstanding load, so the next one is issued right after the wait).
      // Only where it fits the 168-register cap without spilling (measured: with the residual registers, or in the
      // run-time variant that also carries the GEGLU path, the spills cost more than the overlap gains).
      constexpr bool kPrefetchLd = NG == 2 && EPI >= 0 && (EPI & 4) == 0;      // (NG 3 runs at 152 registers: no room)
      uint32_t va[32], vb[32];
      if (kPrefetchLd) tmem_ld32(t_row + first * 64, va);
      float ln_s = 0.f, ln_q = 0.f;          // LayerNorm partials of my row over the first half of the current chunk
      auto do_slice = [&](int sl, uint32_t (&v)[32], uint32_t (&vn)[32]) {
        const int hf = sl & 1;
        const int ch = first + (sl >> 1) * NG;                 // chunk: accumulator / output columns [64 ch, 64 ch + 64)
        const int coff = ch * 64 + hf * 32;
        uint4 res[4];
        if (rp != nullptr) {
          if (kResDeep) {
#pragma unroll
            for (int j = 0; j < 4; j++) res[j] = hf ? res_nxt[4 + j] : res_nxt[j];
            if (sl + 2 < nslices) {                       // the same half of my next chunk
              const uint4* np4 = rp + (ch + NG) * 8 + hf * 4;
              if (hf) {
#pragma unroll
                for (int j = 0; j < 4; j++) res_nxt[4 + j] = __ldg(np4 + j);
              } else {
#pragma unroll
                for (int j = 0; j < 4; j++) res_nxt[j] = __ldg(np4 + j);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) res[j] = res_nxt[j];
            if (sl + 1 < nslices) {
              const int nch = first + ((sl + 1) >> 1) * NG;
              const uint4* np4 = rp + nch * 8 + ((sl + 1) & 1) * 4;
#pragma unroll
              for (int j = 0; j < 4; j++) res_nxt[j] = __ldg(np4 + j);
            }
          }
        }
        uint8_t* stg = stg_grp + (chunk_no & 1) * 16384;
        uint8_t* srow = stg + row * 128;
        if (kPrefetchLd) {
          tmem_ld_wait_regs(v);
          if (sl + 1 < nslices) tmem_ld32(t_row + (first + ((sl + 1) >> 1) * NG) * 64 + ((sl + 1) & 1) * 32, vn);
          else tc_fence_before();            // last TMEM read of this accumulator (released at the next group_sync)
        } else {
          tmem_ld32(t_row + coff, v);
          tmem_ld_wait();
          if (sl == nslices - 1) tc_fence_before();
        }
        const int col0 = tl.n0 + coff;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; i++) f[i] = __uint_as_float(v[i]);
        if (f_alpha) {
#pragma unroll
          for (int i = 0; i < 32; i++) f[i] *= p.b.alpha;
        }
        if (f_ln && f_bias) {
          // folded LayerNorm + its bias (c2) as two FMAs per element: rstd * acc + (k * c1[n] + c2[n]),  k = -mean * rstd
          const float4* cp = reinterpret_cast<const float4*>(p.b.ln_c1 + col0);
          const float4* bp = reinterpret_cast<const float4*>(p.b.bias + col0);
          const float k = -ln_ms.x * ln_ms.y;
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 c4 = __ldg(cp + i4), b4 = __ldg(bp + i4);
            f[4 * i4] = fmaf(f[4 * i4], ln_ms.y, fmaf(k, c4.x, b4.x)); f[4 * i4 + 1] = fmaf(f[4 * i4 + 1], ln_ms.y, fmaf(k, c4.y, b4.y));
            f[4 * i4 + 2] = fmaf(f[4 * i4 + 2], ln_ms.y, fmaf(k, c4.z, b4.z)); f[4 * i4 + 3] = fmaf(f[4 * i4 + 3], ln_ms.y, fmaf(k, c4.w, b4.w));
          }
        } else if (f_ln) {
          const float4* cp = reinterpret_cast<const float4*>(p.b.ln_c1 + col0);
          const float k = -ln_ms.x * ln_ms.y;
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 c4 = __ldg(cp + i4);
            f[4 * i4] = fmaf(f[4 * i4], ln_ms.y, k * c4.x); f[4 * i4 + 1] = fmaf(f[4 * i4 + 1], ln_ms.y, k * c4.y);
            f[4 * i4 + 2] = fmaf(f[4 * i4 + 2], ln_ms.y, k * c4.z); f[4 * i4 + 3] = fmaf(f[4 * i4 + 3], ln_ms.y, k * c4.w);
          }
        }
        if (f_bias && !f_ln) {
          const float4* bp = reinterpret_cast<const float4*>(p.b.bias + col0);
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 b4 = __ldg(bp + i4);
            f[4 * i4] += b4.x; f[4 * i4 + 1] += b4.y; f[4 * i4 + 2] += b4.z; f[4 * i4 + 3] += b4.w;
          }
        }
        if (f_bias2) {
          const float4* bp = reinterpret_cast<const float4*>(p.b.bias2 + (size_t)sample * p.b.N + col0);
#pragma unroll
          for (int i4 = 0; i4 < 8; i4++) {
            const float4 b4 = __ldg(bp + i4);
            f[4 * i4] += b4.x; f[4 * i4 + 1] += b4.y; f[4 * i4 + 2] += b4.z; f[4 * i4 + 3] += b4.w;
          }
        }
        if (rp != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const float2 r0 = unpack_half2(res[j].x), r1 = unpack_half2(res[j].y), r2 = unpack_half2(res[j].z),
                         r3 = unpack_half2(res[j].w);
            f[8 * j + 0] += r0.x; f[8 * j + 1] += r0.y; f[8 * j + 2] += r1.x; f[8 * j + 3] += r1.y;
            f[8 * j + 4] += r2.x; f[8 * j + 5] += r2.y; f[8 * j + 6] += r3.x; f[8 * j + 7] += r3.y;
          }
        }
        if (p.b.ln_out != nullptr) {
          // LayerNorm statistics for the consumer of this output (TapGemm::ln_out): thread == row, so a row's (sum, sum of
          // squares) over the chunk's 64 columns are two private registers -- no shuffles, one 8-byte store per chunk
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            s0 += f[i]; q0 = fmaf(f[i], f[i], q0);
            s1 += f[i + 1]; q1 = fmaf(f[i + 1], f[i + 1], q1);
          }
          if (!hf) { ln_s = s0 + s1; ln_q = q0 + q1; }
          else if (row_ok)
            p.b.ln_out[(int64_t)((tl.n0 >> 6) + ch) * p.b.ln_rows + pix] = make_float2(ln_s + (s0 + s1), ln_q + (q0 + q1));
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<uint4*>(srow + (((hf * 4 + j) ^ (row & 7)) << 4)) =
              make_uint4(pack_half2(f[8 * j + 0], f[8 * j + 1]), pack_half2(f[8 * j + 2], f[8 * j + 3]),
                         pack_half2(f[8 * j + 4], f[8 * j + 5]), pack_half2(f[8 * j + 6], f[8 * j + 7]));
        if (hf) {                            // a 64-column chunk is complete
          fence_proxy_async_smem();
          if (issuer) tma_store_wait_read0();  // my previous chunk's store has read the OTHER staging tile
          __syncwarp();
          group_sync();
          if (issuer) {
            if (sl == nslices - 1) arrive_empty(empty_bar);   // the whole group is past its TMEM reads
            if (!(p.dbg & 1)) {
              tma_store_5d(&tmD, stg, tl.n0 + ch * 64, tl.w0, tl.h0, tl.t0, tl.b0);
              tma_store_commit();
            }
          }
          __syncwarp();
          if (p.gn_sums != nullptr) {          // GroupNorm statistics of this chunk for the consumer norm (see gn_chunk_stats)
            const uint32_t valid = __ballot_sync(0xffffffffu, row_ok);
            if (valid != 0u) {
              const int smp = __shfl_sync(0xffffffffu, fd_div(p.fd_gn, pb * p.b.dimT + pt), __ffs((int)valid) - 1);
              gn_chunk_stats(stg, q, lane, valid, tl.n0 + ch * 64, p.fd_cpg, p.gn_sums + (size_t)smp * 64);
            }
          }
          chunk_no++;
          if (grp < 2 && q == 2 && lane == 0) G3_TRACE(2 + grp, lt, 2 + (sl >> 1));   // chunk stored
        }
      };
#pragma unroll 1
      for (int ci = 0; ci < mine; ci++) {
        if (kPrefetchLd) {
          do_slice(2 * ci, va, vb);
          do_slice(2 * ci + 1, vb, va);
        } else {
          do_slice(2 * ci, va, va);
          do_slice(2 * ci + 1, va, va);
        }
      }
    }
    if (issuer) tma_store_wait_read0();
    __syncwarp();
    tc_fence_before();
  }
  cluster_sync_all();      // the peer's smem / barriers / TMEM stay alive until both CTAs are done
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <typename TA, typename TD>
__global__ void tapgemm_generic_kernel(TapGemmGeneric g) {
  const int64_t total = (int64_t)g.B * g.T * g.H * g.W * g.N;
  const TA* A = static_cast<const TA*>(g.A);
  TD* D = static_cast<TD*>(g.D);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = idx % g.N;
    int64_t m = idx / g.N;
    const int w = m % g.W; m /= g.W;
    const int h = m % g.H; m /= g.H;
    const int t = m % g.T;
    const int b = m / g.T;
    float acc = 0.f;
    for (int tap = 0; tap < g.ntaps; tap++) {
      const int ww = w + g.taps[tap][0], hh = h + g.taps[tap][1], tt = t + g.taps[tap][2];
      if (ww < 0 || ww >= g.W || hh < 0 || hh >= g.H || tt < 0 || tt >= g.T) continue;
      const TA* a = A + b * g.a_sb + tt * g.a_st + hh * g.a_sh + ww * g.a_sw;
      const __half* wr = g.Wt + ((int64_t)n * g.ntaps + tap) * g.CinW;
      for (int c = 0; c < g.Cin; c++) acc += static_cast<float>(a[c * g.a_sc]) * __half2float(wr[c]);
    }
    float out = acc * g.alpha;
    if (g.bias) out += g.bias[n];
    D[b * g.d_sb + t * g.d_st + h * g.d_sh + w * g.d_sw + n * g.d_sn] = static_cast<TD>(out);
  }
}

// Largest power-of-two box edge (<= budget) that overshoots the extent by at most 1/16.
int pick_box(int extent, int budget) {
  int best = 1;
  for (int pw = 2; pw <= budget; pw *= 2) {
    const int64_t cov = (int64_t)((extent + pw - 1) / pw) * pw;
    if (cov * 16 <= (int64_t)extent * 17) best = pw;
  }
  return best;
}

TcParams make_params(const TapGemm& g) {
  TcParams p{};
  p.ntaps = g.ntaps;
  p.kchunks = (g.Cin + BK - 1) / BK;
  p.cin = g.Cin;
  p.N = g.N;
  p.n_out = g.geglu ? g.N / 2 : g.N;
  p.geglu = g.geglu ? 1 : 0;
  p.has_res = g.R ? 1 : 0;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.bias2 = g.bias2;
  p.bias2_div = g.bias2_div > 0 ? g.bias2_div : 1;
  p.fd_b2 = make_fastdiv(p.bias2_div);
  p.nb2 = g.nb2;
  p.dimT = g.T;
  p.ln_stats = g.ln_stats;
  p.ln_c1 = g.ln_c1;
  p.ln_rows = (int64_t)g.B * g.T * g.H * g.W;
  p.ln_out = nullptr;       // only the pair kernel produces them (tapgemm_tc3)
  for (int i = 0; i < g.ntaps; i++)
    for (int j = 0; j < 3; j++) p.taps[i][j] = g.taps[i][j];
  return p;
}

}  // namespace

void set_taps_3x3(int8_t taps[9][3]) {
  for (int kh = 0; kh < 3; kh++)
    for (int kw = 0; kw < 3; kw++) {
      taps[kh * 3 + kw][0] = (int8_t)(kw - 1);
      taps[kh * 3 + kw][1] = (int8_t)(kh - 1);
      taps[kh * 3 + kw][2] = 0;
    }
}
void set_taps_t3(int8_t taps[9][3]) {
  for (int kt = 0; kt < 3; kt++) {
    taps[kt][0] = 0;
    taps[kt][1] = 0;
    taps[kt][2] = (int8_t)(kt - 1);
  }
}

bool tapgemm_tc_eligible(const TapGemm& g) {
  if (g.Cin % 8 != 0 || g.N % 64 != 0) return false;
  if (g.geglu && g.N % 128 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.Wt) | reinterpret_cast<uintptr_t>(g.D) |
       reinterpret_cast<uintptr_t>(g.R)) & 15)
    return false;
  return true;
}

// CTA-pair kernel (tapgemm_tc3_kernel): large problems only (several waves of 256 x 256 tiles over the 74 TPCs)
static long long* g_gemm_trace = nullptr;
void gemm_set_trace(long long* buf) { g_gemm_trace = buf; }

bool tapgemm_pair_wanted(const TapGemm& g, int64_t m_tiles, int nt128) {
  const int mode = knobs().gemm_pair;
  if (mode == 0) return false;
  if (mode == 1) return true;
  const int ktot_steps = g.ntaps * ((g.Cin + BK - 1) / BK);
  // switch-over measured on the MDM512 shapes (tests/gpu_bench_gemm.py mdm512): the pair kernel wins from about one
  // 128 x 128 tile per SM upward (conv 3x3 1280 at 2 560 rows: 64.6 vs 77.9 us), the single-CTA kernel below (640 rows)
  return m_tiles * nt128 >= (int64_t)sm_count() && ktot_steps >= 4;
}

static int64_t tapgemm_boxes(const TapGemm& g, int& bw, int& bh, int& bt, int& bb) {
  int budget = BM;
  bw = pick_box(g.W, budget); budget /= bw;
  bh = pick_box(g.H, budget); budget /= bh;
  bt = pick_box(g.T, budget); budget /= bt;
  bb = budget;
  return (int64_t)((g.W + bw - 1) / bw) * ((g.H + bh - 1) / bh) * ((g.T + bt - 1) / bt) * ((g.B + bb - 1) / bb);
}

bool tapgemm_ln_out_ok(const TapGemm& g) {
  if (!tapgemm_tc_eligible(g) || g.geglu || knobs().ln_fuse == 0) return false;
  int bw, bh, bt, bb;
  const int64_t m_tiles = tapgemm_boxes(g, bw, bh, bt, bb);
  return tapgemm_pair_wanted(g, m_tiles, (g.N + G2_BN_MAX - 1) / G2_BN_MAX);
}

bool tapgemm_per_sample_ok(const TapGemm& g) {
  if (!tapgemm_tc_eligible(g) || g.wt_samples <= 0) return false;
  int bw, bh, bt, bb;
  const int64_t m_tiles = tapgemm_boxes(g, bw, bh, bt, bb);
  const int div = g.wt_div > 0 ? g.wt_div : 1;
  return tapgemm_pair_wanted(g, m_tiles, (g.N + G2_BN_MAX - 1) / G2_BN_MAX) && bb == 1 && div % bt == 0;
}

// 64 x 64 fp16 identity, one per device: the B operand of the residual k-steps
static const __half* identity64() {
  static __half* dev_ptr[64] = {};
  int d = 0;
  MUDG_CUDA(cudaGetDevice(&d));
  __half*& ptr = dev_ptr[d & 63];
  if (ptr == nullptr) {
    std::vector<__half> h(64 * 64, __float2half(0.f));
    for (int i = 0; i < 64; i++) h[i * 64 + i] = __float2half(1.f);
    MUDG_CUDA(cudaMalloc(&ptr, h.size() * sizeof(__half)));
    MUDG_CUDA(cudaMemcpy(ptr, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  return ptr;
}

bool tapgemm_tc3(const TapGemm& g, cudaStream_t st) {
  G3Params p{};
  p.b = make_params(g);
  int budget = BM;
  p.b.bw = pick_box(g.W, budget); budget /= p.b.bw;
  p.b.bh = pick_box(g.H, budget); budget /= p.b.bh;
  p.b.bt = pick_box(g.T, budget); budget /= p.b.bt;
  p.b.bb = budget;
  p.b.tiles_w = (g.W + p.b.bw - 1) / p.b.bw;
  p.b.tiles_h = (g.H + p.b.bh - 1) / p.b.bh;
  p.b.tiles_t = (g.T + p.b.bt - 1) / p.b.bt;
  p.tiles_b = (g.B + p.b.bb - 1) / p.b.bb;
  const int64_t m_tiles = (int64_t)p.b.tiles_w * p.b.tiles_h * p.b.tiles_t * p.tiles_b;
  // balanced N tiles of whole units (64 columns; 128 for GEGLU so value/gate pairs stay inside one epilogue group)
  p.n_unit = g.geglu ? 128 : 64;
  const int units = g.N / p.n_unit, max_units = 256 / p.n_unit;
  p.nt = (units + max_units - 1) / max_units;
  {
    // Wave balance: the tiles of one launch are equal work, dealt to sm_count / 2 CTA pairs, so the launch takes
    // ceil(tiles / pairs) waves of one tile width each.  Narrower N tiles (down to half the maximum) are taken when they cut
    // that product by 10 % or more: 2 560 x 1 280 is one wave of 70 192-wide tiles instead of one of 50 256-wide ones, 4 608 x
    // 1 280 two waves of 192 instead of two of 256.  The constant stands for the per-tile pipeline fill / epilogue tail and the
    // lower MMA efficiency of narrow tiles (with 32, 10 240 x 640 went to 128-wide tiles and lost 6 %).
    const int64_t pairs = std::max(1, sm_count() / 2), m_pairs = (m_tiles + 1) / 2;
    auto cost = [&](int nt) {
      const int width = ((units + nt - 1) / nt) * p.n_unit;
      return ((int64_t)nt * m_pairs + pairs - 1) / pairs * (width + 96);
    };
    int best = p.nt;
    if (knobs().gemm_balance)
      for (int nt = p.nt + 1; nt <= std::min(units, 2 * p.nt); nt++)
        if (cost(nt) * 10 <= cost(best) * 9) best = nt;
    p.nt = best;
  }
  p.n_q = units / p.nt;
  p.n_rem = units % p.nt;
  p.bn_first = (p.n_q + (p.n_rem > 0 ? 1 : 0)) * p.n_unit;
  const int bn_last = p.n_q * p.n_unit;
  p.b.tiles_n = p.nt;
  p.b.fd_nt = make_fastdiv(p.nt);
  p.b.fd_tw = make_fastdiv(p.b.tiles_w);
  p.b.fd_th = make_fastdiv(p.b.tiles_h);
  p.b.fd_tt = make_fastdiv(p.b.tiles_t);
  const int64_t total = (int64_t)p.nt * ((m_tiles + 1) / 2);
  MUDG_REQUIRE(total < (int64_t(1) << 30), "grid too large");
  p.total_tiles = (int)total;
  p.dimW = g.W; p.dimH = g.H; p.dimB = g.B;
  p.R = g.R;
  p.alpha_is_one = g.alpha == 1.f ? 1 : 0;
  // fused GroupNorm statistics: a 128-pixel box must lie inside one sample, and a column pair inside one group
  const int gn_div = g.gn_div > 0 ? g.gn_div : 1;
  // (only where the MMA main loop leaves the epilogue slack: measured at level 0, the 3-tap temporal conv -- K = 960, already
  // epilogue-bound -- got 95 us slower with the statistics, twice what the separate statistics pass costs; the 9-tap convs
  // absorb them for free)
  const int gn_min_taps = knobs().gn_fuse == 2 ? 1 : 9;
  const bool gn_fuse = g.gn_sums != nullptr && knobs().gn_fuse != 0 && g.ntaps >= gn_min_taps && !g.geglu && p.b.bb == 1 &&
                       gn_div % p.b.bt == 0 && p.b.n_out % 64 == 0 && (p.b.n_out / 32) % 2 == 0;
  p.gn_sums = gn_fuse ? g.gn_sums : nullptr;
  p.fd_cpg = make_fastdiv(std::max(1, p.b.n_out / 32));
  p.fd_gn = make_fastdiv(gn_div);
  p.trace = g_gemm_trace;
  p.dbg = knobs().gemm_dbg;
  MUDG_REQUIRE(g.ln_out == nullptr || !g.geglu, "LayerNorm partials of a GEGLU output are not implemented");
  p.b.ln_out = g.ln_out;

  const uint64_t C = g.Cin, No = p.b.n_out, Ktot = (uint64_t)g.ntaps * g.Cin;
  const uint64_t adims[5] = {C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, (uint64_t)g.B};
  const uint64_t astr[4] = {C * 2, C * 2 * g.W, C * 2 * g.W * g.H, C * 2 * g.W * g.H * g.T};
  const uint32_t abox[5] = {BK, (uint32_t)p.b.bw, (uint32_t)p.b.bh, (uint32_t)p.b.bt, (uint32_t)p.b.bb};
  const uint64_t ddims[5] = {No, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, (uint64_t)g.B};
  const uint64_t dstr[4] = {No * 2, No * 2 * g.W, No * 2 * g.W * g.H, No * 2 * g.W * g.H * g.T};
  // weights [samples][N][K]: one matrix, or one per GroupNorm sample (see TapGemm::wt_samples)
  const uint64_t nws = g.wt_samples > 0 ? (uint64_t)g.wt_samples : 1;
  p.fd_ws = FastDiv{0u, 0u, 0};
  if (g.wt_samples > 0) {
    MUDG_REQUIRE(p.b.bb == 1 && (g.wt_div > 0 ? g.wt_div : 1) % p.b.bt == 0, "per-sample weights: a tile would straddle samples");
    p.fd_ws = make_fastdiv(g.wt_div > 0 ? g.wt_div : 1);
  }
  const uint64_t bdims[5] = {Ktot, (uint64_t)g.N, nws, 1, 1};
  const uint64_t bstr[4] = {Ktot * 2, Ktot * 2 * g.N, Ktot * 2 * g.N * nws, Ktot * 2 * g.N * nws};
  const uint32_t bbox0[5] = {BK, (uint32_t)(p.bn_first / 2), 1, 1, 1};
  const uint32_t bbox1[5] = {BK, (uint32_t)(bn_last / 2), 1, 1, 1};
  const CUtensorMap* ma = get_tmap(g.A, adims, astr, abox);
  const CUtensorMap* mb0 = get_tmap(g.Wt, bdims, bstr, bbox0);
  const CUtensorMap* mb1 = get_tmap(g.Wt, bdims, bstr, bbox1);
  const CUtensorMap* md = get_tmap(g.D, ddims, dstr, abox);
  const int ktot_steps = g.ntaps * p.b.kchunks;
  // Residual through the tensor core (see the producer).  One 64-column MMA step per chunk costs about one extra full-width
  // k-step per tile; measured (tests/gpu_bench_gemm.py resmma) it wins on every shape: 320->320+res 178 -> 160 us, the
  // MDM512-sized ones -14..16 %, and even the 9-tap convs and K = 5120 layers by 1-4 % (their epilogue loses its only
  // global loads).  Knob gemm_resmma = 0 keeps the residual in the epilogue.
  const bool res_ok = g.R != nullptr && !g.geglu && g.alpha == 1.f;
  const bool res_mma = res_ok && knobs().gemm_resmma != 0;
  p.res_mma = res_mma ? 1 : 0;
  const CUtensorMap* mr = md;
  const CUtensorMap* mj = md;
  if (res_mma) {
    p.R = nullptr;                                   // the epilogue does not see a residual
    mr = get_tmap(g.R, ddims, dstr, abox);
    const uint64_t jdims[5] = {64, 64, 1, 1, 1};
    const uint64_t jstr[4] = {128, 128 * 64, 128 * 64, 128 * 64};
    const uint32_t jbox[5] = {BK, 32, 1, 1, 1};
    mj = get_tmap(identity64(), jdims, jstr, jbox);
  }
  const int clusters = (int)std::min<int64_t>(total, sm_count() / 2);
  p.rot_div = FastDiv{0u, 0u, 0};
  if (p.nt > 1 && clusters % p.nt == 0) p.rot_div = make_fastdiv(clusters);
  // epilogue specialisation (see the kernel's EPI comment); knob gemm_epi = 0 forces the run-time variant
  const bool epi_on = knobs().gemm_epi != 0;
  int epi = -1;
  if (epi_on && !g.geglu && g.alpha == 1.f)
    epi = (g.bias ? 1 : 0) | (g.bias2 ? 2 : 0) | ((g.R && !res_mma) ? 4 : 0) | (g.ln_stats ? 8 : 0);
  if (epi != 0 && epi != 1 && epi != 3 && epi != 5 && epi != 9) epi = -1;
  // three epilogue groups (see G3Cfg) for the short-K Linear layers: K = 320 always, K = 640 when the epilogue carries a
  // residual or GEGLU (measured per shape inside the clip); knob gemm_groups = 2 | 3 forces
  int ng = (g.ntaps == 1 && (ktot_steps <= 5 || (ktot_steps <= 10 && (g.geglu || g.R != nullptr)))) ? 3 : 2;
  if (knobs().gemm_groups == 2 || knobs().gemm_groups == 3) ng = knobs().gemm_groups;
  const dim3 grid(2 * clusters);
#define MUDG_TC3_LAUNCH(E, G)                                                                                         \
  do {                                                                                                                 \
    static OncePerDevice attr_done;                                                                                    \
    if (attr_done.first())                                                                                             \
      MUDG_CUDA(cudaFuncSetAttribute(tapgemm_tc3_kernel<E, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, G3Cfg<G>::SMEM)); \
    tapgemm_tc3_kernel<E, G><<<grid, G3Cfg<G>::THREADS, G3Cfg<G>::SMEM, st>>>(*ma, *mb0, *mb1, *md, *mr, *mj, p);       \
  } while (0)
#define MUDG_TC3_EPI(E)                 \
  do {                                  \
    if (ng == 3) MUDG_TC3_LAUNCH(E, 3); \
    else MUDG_TC3_LAUNCH(E, 2);         \
  } while (0)
  switch (epi) {
    case 0: MUDG_TC3_EPI(0); break;
    case 1: MUDG_TC3_EPI(1); break;
    case 3: MUDG_TC3_EPI(3); break;
    case 5: MUDG_TC3_EPI(5); break;
    case 9: MUDG_TC3_EPI(9); break;
    default: MUDG_TC3_EPI(-1); break;
  }
#undef MUDG_TC3_EPI
#undef MUDG_TC3_LAUNCH
  knobs().last_gemm_path = 4 | ((epi + 1) << 8) | (gn_fuse ? 1 << 16 : 0) | (res_mma ? 1 << 17 : 0) | (g.ln_out ? 1 << 18 : 0) | (ng << 20);
  MUDG_CUDA(cudaGetLastError());
  return gn_fuse;
}

bool tapgemm_tc2(const TapGemm& g, cudaStream_t st) {
  MUDG_REQUIRE(tapgemm_tc_eligible(g), "layer not eligible for the tcgen05 path (Cin=%d N=%d)", g.Cin, g.N);
  MUDG_REQUIRE(g.wt_samples == 0 || tapgemm_per_sample_ok(g), "per-sample weights need the pair kernel (check tapgemm_per_sample_ok first)");
  {
    int budget = BM;
    const int bw = pick_box(g.W, budget); budget /= bw;
    const int bh = pick_box(g.H, budget); budget /= bh;
    const int bt = pick_box(g.T, budget); budget /= bt;
    const int64_t m_tiles = (int64_t)((g.W + bw - 1) / bw) * ((g.H + bh - 1) / bh) * ((g.T + bt - 1) / bt) *
                            ((g.B + budget - 1) / budget);
    if (tapgemm_pair_wanted(g, m_tiles, (g.N + G2_BN_MAX - 1) / G2_BN_MAX)) return tapgemm_tc3(g, st);
  }
  MUDG_REQUIRE(g.ln_out == nullptr, "LayerNorm partials need the pair kernel (check tapgemm_ln_out_ok first)");
  G2Params p{};
  p.b = make_params(g);
  int budget = BM;
  p.b.bw = pick_box(g.W, budget); budget /= p.b.bw;
  p.b.bh = pick_box(g.H, budget); budget /= p.b.bh;
  p.b.bt = pick_box(g.T, budget); budget /= p.b.bt;
  p.b.bb = budget;
  p.b.tiles_w = (g.W + p.b.bw - 1) / p.b.bw;
  p.b.tiles_h = (g.H + p.b.bh - 1) / p.b.bh;
  p.b.tiles_t = (g.T + p.b.bt - 1) / p.b.bt;
  p.tiles_b = (g.B + p.b.bb - 1) / p.b.bb;
  p.nt = (g.N + G2_BN_MAX - 1) / G2_BN_MAX;
  p.b.tiles_n = p.nt;
  p.b.fd_nt = make_fastdiv(p.nt);
  p.b.fd_tw = make_fastdiv(p.b.tiles_w);
  p.b.fd_th = make_fastdiv(p.b.tiles_h);
  p.b.fd_tt = make_fastdiv(p.b.tiles_t);
  const int64_t m_tiles = (int64_t)p.b.tiles_w * p.b.tiles_h * p.b.tiles_t * p.tiles_b;
  MUDG_REQUIRE(m_tiles * p.nt < (int64_t(1) << 30), "grid too large");
  p.m_tiles = (int)m_tiles;
  // 256-row super tiles when there is enough work to fill the machine several times over (knob gemm_sub = 1|2 forces)
  const int force_sub = knobs().gemm_sub;
  const int ktot_steps = g.ntaps * ((g.Cin + BK - 1) / BK);
  int sub = (m_tiles * p.nt >= 8 * (int64_t)sm_count() && ktot_steps >= 8) ? 2 : 1;
  if (force_sub == 1 || force_sub == 2) sub = force_sub;
  const int64_t total = (int64_t)p.nt * ((m_tiles + sub - 1) / sub);
  p.total_tiles = (int)total;
  p.dimW = g.W; p.dimH = g.H; p.dimB = g.B;
  p.R = g.R;
  p.alpha_is_one = g.alpha == 1.f ? 1 : 0;

  const uint64_t C = g.Cin, No = p.b.n_out, Ktot = (uint64_t)g.ntaps * g.Cin;
  const uint64_t adims[5] = {C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, (uint64_t)g.B};
  const uint64_t astr[4] = {C * 2, C * 2 * g.W, C * 2 * g.W * g.H, C * 2 * g.W * g.H * g.T};
  const uint32_t abox[5] = {BK, (uint32_t)p.b.bw, (uint32_t)p.b.bh, (uint32_t)p.b.bt, (uint32_t)p.b.bb};
  const uint64_t ddims[5] = {No, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T, (uint64_t)g.B};
  const uint64_t dstr[4] = {No * 2, No * 2 * g.W, No * 2 * g.W * g.H, No * 2 * g.W * g.H * g.T};
  const uint64_t bdims[5] = {Ktot, (uint64_t)g.N, 1, 1, 1};
  const uint64_t bstr[4] = {Ktot * 2, Ktot * 2 * g.N, Ktot * 2 * g.N, Ktot * 2 * g.N};
  const uint32_t bbox[5] = {BK, G2_BN_MAX, 1, 1, 1};
  const CUtensorMap* ma = get_tmap(g.A, adims, astr, abox);
  const CUtensorMap* mb = get_tmap(g.Wt, bdims, bstr, bbox);
  const CUtensorMap* md = get_tmap(g.D, ddims, dstr, abox);
  static OncePerDevice attr_set;
  if (attr_set.first()) {
    MUDG_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<1>::SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<2>::SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(tapgemm_tc2_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<1, 1>::SMEM));
  }
  const bool deep = sub == 1 && total <= sm_count() && ktot_steps >= 8 && knobs().gemm_deep != 0;
  if (deep) {
    tapgemm_tc2_kernel<1, 1><<<(int)total, G2Cfg<1, 1>::THREADS, G2Cfg<1, 1>::SMEM, st>>>(*ma, *mb, *md, p);
  } else if (sub == 1) {
    const int grid = (int)std::min<int64_t>(total, 2 * sm_count());
    tapgemm_tc2_kernel<1><<<grid, G2Cfg<1>::THREADS, G2Cfg<1>::SMEM, st>>>(*ma, *mb, *md, p);
  } else {
    const int grid = (int)std::min<int64_t>(total, sm_count());
    tapgemm_tc2_kernel<2><<<grid, G2Cfg<2>::THREADS, G2Cfg<2>::SMEM, st>>>(*ma, *mb, *md, p);
  }
  knobs().last_gemm_path = sub == 1 ? 2 : 3;
  MUDG_CUDA(cudaGetLastError());
  return false;
}

void tapgemm_generic(const TapGemmGeneric& g, cudaStream_t st) {
  const int64_t total = (int64_t)g.B * g.T * g.H * g.W * g.N;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 64);
  if (g.a_fp32 && g.d_fp32) tapgemm_generic_kernel<float, float><<<blocks, 256, 0, st>>>(g);
  else if (g.a_fp32) tapgemm_generic_kernel<float, __half><<<blocks, 256, 0, st>>>(g);
  else if (g.d_fp32) tapgemm_generic_kernel<__half, float><<<blocks, 256, 0, st>>>(g);
  else tapgemm_generic_kernel<__half, __half><<<blocks, 256, 0, st>>>(g);
  MUDG_CUDA(cudaGetLastError());
}

// The product entry.  With the profiler on (bench.py's roofline leg) the launch is bracketed by a CUDA-event pair and
// booked with its algorithmic work: 2 * rows * N * taps * Cin FLOPs (padding and tile overhang not counted) and the HBM
// bytes an ideal kernel moves (A once, W once, D once, residual once, LayerNorm statistics).
bool tapgemm(const TapGemm& g, cudaStream_t st) {
  if (!tapgemm_tc_eligible(g)) {
    // Irregular plain GEMMs (N not a multiple of 64: the VAE mid-block attention at latent sizes whose token count is
    // not, e.g. 8 x 12) run on the arbitrary-stride CUDA-core kernel like the other irregular layers; anything with a
    // fused epilogue must be eligible.
    MUDG_REQUIRE(!g.geglu && g.ln_stats == nullptr && g.bias2 == nullptr && g.R == nullptr,
                 "layer not eligible for the tcgen05 path (Cin=%d N=%d, 16 B alignment) and it needs a fused epilogue", g.Cin, g.N);
    TapGemmGeneric q;
    q.A = g.A; q.B = g.B; q.T = g.T; q.H = g.H; q.W = g.W; q.Cin = g.Cin;
    q.a_sc = 1; q.a_sw = g.Cin; q.a_sh = (int64_t)g.W * g.Cin; q.a_st = q.a_sh * g.H; q.a_sb = q.a_st * g.T;
    q.ntaps = g.ntaps;
    memcpy(q.taps, g.taps, sizeof(q.taps));
    q.Wt = g.Wt; q.CinW = g.Cin; q.N = g.N;
    q.D = g.D; q.d_sn = 1; q.d_sw = g.N; q.d_sh = (int64_t)g.W * g.N; q.d_st = q.d_sh * g.H; q.d_sb = q.d_st * g.T;
    q.bias = g.bias; q.alpha = g.alpha;
    ProfScope ps(PF_GENERIC_CONV, 0.0, 0.0, st, "irregular gemm");
    tapgemm_generic(q, st);
    return false;
  }
  if (!prof_active()) return tapgemm_tc2(g, st);
  const double rows = (double)g.B * g.T * g.H * g.W;
  const int n_out = g.geglu ? g.N / 2 : g.N;
  const double flops = 2.0 * rows * (double)g.N * (double)g.ntaps * g.Cin;
  const double bytes = 2.0 * (rows * g.Cin + (double)g.N * g.ntaps * g.Cin + rows * n_out * (g.R ? 2.0 : 1.0)) +
                       (g.ln_stats ? 8.0 * rows : 0.0);
  char buf[96];
  snprintf(buf, sizeof buf, "%.0fx%dx%dx%d:g%dr%dl%db%d", rows, g.N, g.ntaps, g.Cin, g.geglu ? 1 : 0, g.R ? 1 : 0,
           g.ln_stats ? 1 : 0, g.bias2 ? 1 : 0);
  ProfScope ps(PF_GEMM, flops, bytes, st, buf);
  return tapgemm_tc2(g, st);
}

}  // namespace mudg
