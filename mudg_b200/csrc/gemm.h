// "Tap GEMM": the one contraction every Linear / Conv2d 3x3 / Conv3d (3,1,1) / 1x1 conv of the path maps to.
//
//   D[b,t,h,w,n] = epi( alpha * sum_{tap, c} A[b, t+dt, h+dh, w+dw, c] * Wt[n, tap*Cin + c] )
//
// A, D, R are channels-last fp16 (see Act); out-of-range taps read zeros (== conv zero padding).
// epi: + bias[n] + bias2[sample(b,t)][n] + R[b,t,h,w,n]; or GEGLU (weights pre-interleaved 64 value rows /
// 64 gate rows): D[.., j] = (v + bv) * gelu_erf(g + bg).
#pragma once
#include "common.h"

namespace mudg {

struct TapGemm {
  // A geometry (channels-last, contiguous)
  const __half* A = nullptr;
  int B = 1, T = 1, H = 1, W = 1, Cin = 0;
  int ntaps = 1;
  int8_t taps[9][3] = {{0, 0, 0}};   // (dw, dh, dt)
  // weights [N][ntaps*Cin] fp16, K-major
  const __half* Wt = nullptr;
  int N = 0;
  // output [B,T,H,W,n_out] (n_out == N, or N/2 with geglu)
  __half* D = nullptr;
  const __half* R = nullptr;       // optional residual, same shape as D
  const float* bias = nullptr;     // [N]
  const float* bias2 = nullptr;    // [nb2][N]; row = (b*T + t) / bias2_div
  int bias2_div = 1, nb2 = 0;
  float alpha = 1.f;
  bool geglu = false;
  // LayerNorm folded into the GEMM (A is the RAW activation, Wt already carries gamma):
  //   D = rstd[m] * (acc - mean[m] * ln_c1[n]) + bias[n]      with bias = W beta (+ the layer's own bias)
  const float2* ln_stats = nullptr;   // [rows] (mean, rstd)
  const float* ln_c1 = nullptr;       // [N] row sums of the gamma-scaled fp16 weight
  // LayerNorm statistics of the OUTPUT for the LayerNorm that consumes it: REQUEST to store, per row and 64-column chunk of
  // the output, (sum, sum of squares) of the fp32 epilogue values into ln_out [n_out / 64][rows] float2 (plain stores, every
  // element written exactly once: deterministic, no zeroing); ln_finalize (ops.h) turns the planes into the (mean, rstd)
  // rows a consumer takes as ln_stats -- an 8 % pass over the partials instead of ln_stats_kernel's full read of the
  // activation.  Only the CTA-pair kernel does this: ask tapgemm_ln_out_ok() first.
  float2* ln_out = nullptr;
  // GroupNorm statistics of the OUTPUT for the norm that consumes it (32 groups over the n_out channels): REQUEST to
  // accumulate (sum, sum of squares) per (sample, group) into gn_sums [S][32][2] fp64 (pre-zeroed) from the epilogue;
  // sample = (b*T + t) / gn_div.  tapgemm() returns whether it did (only the CTA-pair kernel does, and only when a
  // 128-pixel tile cannot straddle samples); otherwise the caller runs gn_stats on the output.
  double* gn_sums = nullptr;
  int gn_div = 1;
  // Per-sample weights: Wt is [wt_samples][N][ntaps*Cin] and the rows of sample (b*T + t) / wt_div use matrix number that.
  // This is how a GroupNorm WITHOUT activation is folded into the Linear that consumes it (SpatialTransformer /
  // TemporalTransformer norm -> proj_in): y = (x * scale_s + shift_s) W^T + b = x (W diag(scale_s))^T + (W shift_s + b), the
  // second term arriving as bias2.  CTA-pair kernel only, and only when a 128-pixel tile cannot straddle samples:
  // ask tapgemm_per_sample_ok() first.
  int wt_samples = 0, wt_div = 1;
};

// generic-stride variant for the irregular layers (tiny Cin / tiny N, fp32 NCTHW in/out)
struct TapGemmGeneric {
  const void* A = nullptr;
  bool a_fp32 = false;
  int B = 1, T = 1, H = 1, W = 1, Cin = 0;
  int64_t a_sb = 0, a_st = 0, a_sh = 0, a_sw = 0, a_sc = 1;   // element strides
  int ntaps = 1;
  int8_t taps[9][3] = {{0, 0, 0}};
  const __half* Wt = nullptr;     // [N][ntaps*CinW]
  int CinW = 0;                   // weight row pitch per tap (>= Cin, padded)
  int N = 0;
  void* D = nullptr;
  bool d_fp32 = false;
  int64_t d_sb = 0, d_st = 0, d_sh = 0, d_sw = 0, d_sn = 1;
  const float* bias = nullptr;
  float alpha = 1.f;
};

bool tapgemm_tc_eligible(const TapGemm& g);
bool tapgemm_ln_out_ok(const TapGemm& g);       // would tapgemm() honour g.ln_out (CTA-pair kernel, plain epilogue)?
bool tapgemm_per_sample_ok(const TapGemm& g);   // would tapgemm() run g (with wt_samples / wt_div set) on the pair kernel, tiles inside one sample?
bool tapgemm_tc2(const TapGemm& g, cudaStream_t st);       // persistent single-CTA kernel, or the CTA-pair kernel for large problems
void tapgemm_generic(const TapGemmGeneric& g, cudaStream_t st);
// The product entry: tcgen05 path (+ optional per-launch timing).  Returns true when the GroupNorm statistics requested
// through g.gn_sums were accumulated by the epilogue.
bool tapgemm(const TapGemm& g, cudaStream_t st);

void gemm_set_trace(long long* buf);   // debug: clock64 time line of the pair GEMM's CTA 0 ([4][64][8] int64), null = off

void set_taps_3x3(int8_t taps[9][3]);      // (dw,dh) in {-1,0,1}^2, tap index = kh*3+kw (weight layout [N][kh][kw][Cin])
void set_taps_t3(int8_t taps[9][3]);       // dt in {-1,0,1}

}  // namespace mudg
