// Launchers of the non-GEMM kernels (elem.cu, attn.cu, sampler.cu).  All take raw device pointers + a stream.
#pragma once
#include "common.h"

namespace mudg {

// GroupNorm over [S samples][rows_per_sample][C] (32 groups, fp64 (sum, sum of squares) pairs [S][32][2]):
// gn_stats accumulates into pre-zeroed sums (or the producing GEMM's epilogue does: TapGemm::gn_sums); gn_apply
// normalises straight from the sums in one pass (+ SiLU).
void gn_stats(const __half* x, int S, int64_t rows_per_sample, int C, double* sums, cudaStream_t st);
void gn_apply(const __half* x, __half* y, const double* sums, int S, int64_t rows_per_sample, int C, const float* gamma,
              const float* beta, float eps, bool silu_act, cudaStream_t st);
// GroupNorm (+ SiLU) in one kernel, statistics included, when a (sample, channel slab) fits shared memory (gn_small_ok)
bool gn_small_ok(int S, int64_t rows_per_sample, int C);
void gn_small(const __half* x, __half* y, int S, int64_t rows_per_sample, int C, const float* gamma, const float* beta,
              float eps, bool silu_act, cudaStream_t st);
// GroupNorm without activation folded into the Linear W [N][K] that consumes it: per-sample weights Ws [S][N][K] (fp16)
// and bias rows cs [S][N] (fp32, WITHOUT the layer's own bias); see TapGemm::wt_samples
void gn_fold_weights(const __half* W, const double* sums, int S, int64_t rows_per_sample, const float* gamma,
                     const float* beta, float eps, __half* Ws, float* cs, int N, int K, cudaStream_t st);
void layernorm(const __half* x, __half* y, const float* gamma, const float* beta, int64_t rows, int C, float eps,
               cudaStream_t st);
void ln_stats(const __half* x, float2* out, int64_t rows, int C, float eps, cudaStream_t st);
// the same (mean, rstd) rows from the [nparts = C / 64][rows] partial (sum, sum of squares) planes of TapGemm::ln_out
void ln_finalize(const float2* parts, int nparts, float2* out, int64_t rows, int C, float eps, cudaStream_t st);
void ln_fold(__half* W, const float* gamma, const float* beta, const float* bias, float* c1, float* c2, int N, int K,
             cudaStream_t st);
// out[r] = [a[r] | b[r % rows_b]] (rows_b divides rows: a skip tensor shared by the copies of a CFG batch)
void concat_channels(const __half* a, int Ca, const __half* b, int Cb, __half* out, int64_t rows, int64_t rows_b,
                     cudaStream_t st);
// the same concat, also accumulating the GroupNorm statistics of its output into pre-zeroed sums [S][32][2] (S samples of
// rows_per_sample rows each)
void concat_channels_stats(const __half* a, int Ca, const __half* b, int Cb, __half* out, int S, int64_t rows_per_sample,
                           int64_t rows_b, double* sums, cudaStream_t st);
void upsample2x(const __half* x, __half* y, int F, int H, int W, int C, cudaStream_t st);
void im2col_s2(const __half* x, __half* y, int F, int H, int W, int C, int pad, cudaStream_t st);   // Ho = (H + pad - 2) / 2 + 1
void softmax_rows(__half* x, int64_t rows, int n, cudaStream_t st);
void gelu_inplace(__half* x, int64_t n, cudaStream_t st);   // exact erf GELU, fp16, n % 8 == 0
void pack_weight(const void* src, bool src_fp32, __half* dst, int O, int I, int taps, int Ipad, cudaStream_t st);
void cast_to_f32(const void* src, bool src_fp32, float* dst, int64_t n, cudaStream_t st);
void cast_to_f16(const void* src, bool src_fp32, __half* dst, int64_t n, cudaStream_t st);
void geglu_interleave(const __half* w, const float* b, __half* wo, float* bo, int half_rows, int cols, cudaStream_t st);
void to_channels_last(const void* x, bool x_fp32, __half* y, int B, int C, int64_t R, int Cpad, cudaStream_t st);
void from_channels_last(const __half* y, __half* out, int B, int C, int64_t R, int Cp, cudaStream_t st);
void gather_rows_f16(const void* src, bool src_fp32, __half* dst, int B, int src_rows, int r0, int nrows, int cols,
                     cudaStream_t st);
void sinusoid3(const int64_t* t, const int64_t* label, const int64_t* fs, float* out /*[3][B][dim]*/, int B, int dim,
               cudaStream_t st);
// Small Linear layers batched into one launch (embedding MLPs, ResBlock emb_layers): y_j[b][:] = x_j[b] W_j^T + bias_j for
// job j (n_out_j = off[j+1] - off[j]); with `sum` all jobs share n_out = off[1] and y[0] receives their sum.  fp32 in / out.
struct LinearBatch {
  static constexpr int MAX_JOBS = 24;
  const float* x[MAX_JOBS];
  const __half* W[MAX_JOBS];
  const float* bias[MAX_JOBS];
  float* y[MAX_JOBS];
  int off[MAX_JOBS + 1];
  int count = 0, K = 0, Bn = 0, sum = 0, silu_out = 0;
};
void batched_linear(const LinearBatch& lb, cudaStream_t st);

// ---- attention (attn.cu)
// Flash attention, head dim 64, fp16 in/out, tcgen05.  Q rows: [F frames][Nq tokens], row pitch q_pitch elements,
// head h at columns [h*64, h*64+64).  Up to two K/V segments with SEPARATE softmaxes whose outputs are summed
// (text + image cross-attention, attention.py:129-142).  KV batch index of frame f = f / kv_div.
struct FlashSeg {
  const __half* K = nullptr;
  const __half* V = nullptr;
  const __half* VT = nullptr;   // V transposed by transpose_v: [nbatch][heads*64][vt_pitch], kv contiguous (the tcgen05 kernel reads this)
  int vt_pitch = 0;             // >= len, multiple of 8
  int pitch = 0;       // row pitch (elements)
  int len = 0;         // tokens per kv batch
  int nbatch = 0;      // number of kv batches
  int kv_div = 1;
};
struct FlashArgs {
  const __half* Q = nullptr;
  int q_pitch = 0;
  __half* O = nullptr;
  int o_pitch = 0;
  int F = 0, Nq = 0, heads = 0;
  int nseg = 1;
  FlashSeg seg[2];
  float scale = 0.125f;
};
void flash_attention(const FlashArgs& a, cudaStream_t st);
void transpose_v(const __half* V, int pitch, int len, int nbatch, int heads, __half* VT, int len_pad, cudaStream_t st);
void flash_set_trace(long long* buf);   // debug: clock64 time line of CTA 0 ([3][96][8] int64), null = off

// ---- cross-attention to the per-frame context (xattn.cu): 77 text + 16 image keys merged into one 96-key block.
// xattn_pack builds the merged operands from the projected K|V rows (text_kv [N][77][2C], img_kv [F][16][2C], K in the first
// C columns, V in the last): K [F][96][C] and V^T [F][C][128]; xattn_per_frame is the attention itself (separate text /
// image softmaxes, outputs summed), Q / O rows [F][Nq] with head h at columns [h*64, h*64+64).
struct XattnArgs {
  const __half* Q = nullptr;
  int q_pitch = 0;
  __half* O = nullptr;
  int o_pitch = 0;
  int F = 0, Nq = 0, heads = 0;
  const __half* K = nullptr;     // [F][96][heads*64]
  const __half* VT = nullptr;    // [F][heads*64][128]
  float scale = 0.125f;
};
void xattn_pack(const __half* text_kv, const __half* img_kv, __half* K, __half* VT, int F, int T, int C, cudaStream_t st);
void xattn_per_frame(const XattnArgs& a, cudaStream_t st);

// Temporal self-attention over T for every (b, h, w, head): qkv rows [B*T*HW][3*inner] (q | k | v), out [rows][inner]
void temporal_attention(const __half* qkv, __half* out, int B, int T, int HW, int heads, float scale, cudaStream_t st);

// ---- sampler (sampler.cu): fused CFG + guidance rescale + v-pred DDIM update (ddim.py:226-277)
struct DdimStepArgs {
  const float* x = nullptr;        // [B][n] fp32
  const __half* v_cond = nullptr;  // [B][n] fp16 (UNet output)
  const __half* v_uncond = nullptr;  // or null
  const float* noise = nullptr;    // [B][n] fp32
  float* x_prev = nullptr;
  float* pred_x0 = nullptr;
  int B = 0;
  int64_t n = 0;
  float cfg_scale = 1.f, guidance_rescale = 0.f;
  float sqrt_ac = 0.f, sqrt_1mac = 0.f, rescale = 1.f, sqrt_a_prev = 0.f, dir_coef = 0.f, sigma = 0.f;
};
void ddim_step(const DdimStepArgs& a, cudaStream_t st);

// ---- post-decode frame pipeline (post.cu): decoded clip -> uint8 RGB (+ fp32 depth / class index), bit-exact against
// the reference's CPU code (eval_tools.py).  frames [B][3][T][HW]; modes_host[b]: 0 colour, 1 depth, 2 semantic.
constexpr int POSTDECODE_MAX_SAMPLES = 64;
void postdecode(const void* frames, int dtype /*0 fp32, 1 fp16, 2 uint8*/, uint8_t* rgb /*[B][T][3][HW]*/, float* depth /*[B][T][HW] or null*/,
                uint8_t* cls /*[B][T][HW] or null*/, int B, int T, int64_t HW, const int* modes_host, cudaStream_t st);
// Spectral colour map of a [0,1] fp32 map -> [n][3] uint8 (HWC), colormap(..., "Spectral", bytes=True) of the reference
void spectral_colormap(const float* map, uint8_t* out_hwc, int64_t n, cudaStream_t st);

}  // namespace mudg
