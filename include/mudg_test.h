/* libmudg_sm100_test.so -- TEST-ONLY companion of libmudg_sm100.so (tests/, bench diagnostics).
 *
 * The test library links its own copy of the product objects plus csrc/test/testhooks.cu: single-kernel entry points,
 * CUDA-core checkers of the tcgen05 kernels' contracts, the tcgen05.mma issue-rate probe, clock64 traces and the tuning
 * knobs.  None of these symbols exist in the product library, and the product library reads no tuning switches from
 * the environment.  It also exports the whole product ABI (include/mudg.h), so a test can run the product path with a
 * knob changed.
 */
#ifndef MUDG_TEST_H_
#define MUDG_TEST_H_
#include "mudg.h"
#ifdef __cplusplus
extern "C" {
#endif

/* knobs (csrc/common.h: struct Knobs): gemm_pair, gemm_sub, gemm_epi, gemm_dbg, flash_stagger, flash_poly, tattn_generic,
 * gn_fuse; "reset" restores the shipped defaults */
MUDG_EXPORT int mudg_test_set_knob(const char* name, int value);
/* kernel the last tap-GEMM launch took: 2 = tapgemm_tc2<1>, 3 = tapgemm_tc2<2>, 4 = tapgemm_tc3; | (EPI + 1) << 8 */
MUDG_EXPORT int mudg_test_last_gemm_path(void);   /* bit 16: GroupNorm statistics were accumulated by the epilogue */
/* request GroupNorm statistics of the output from the NEXT mudg_test_tapgemm(backend 0) call: sums [S][32][2] fp64,
 * pre-zeroed; sample = (b*T + t) / gn_div */
MUDG_EXPORT int mudg_test_next_gemm_gn(void* sums_f64, int gn_div);
/* per-sample weight matrices for the NEXT mudg_test_tapgemm(backend 0) call: Wt is [samples][N][K], rows of sample
 * (b*T + t) / div use matrix number that (TapGemm::wt_samples; the GroupNorm-into-proj_in fold) */
MUDG_EXPORT int mudg_test_next_gemm_per_sample(int samples, int div);

/* LayerNorm partial sums of the output of the NEXT mudg_test_tapgemm(backend 0) call: ln_out = [n_out / 64][rows] float2 the
 * epilogue fills with each row's (sum, sum of squares) per 64-column chunk (TapGemm::ln_out; pair kernel only);
 * mudg_test_ln_finalize reduces such planes to the [rows] (mean, rstd) a folded-LayerNorm consumer takes (eps 1e-5) */
MUDG_EXPORT int mudg_test_next_gemm_ln(void* ln_out);
MUDG_EXPORT int mudg_test_ln_finalize(const void* parts, int nparts, void* mean_rstd, int64_t rows, int C, void* stream);

/* backend 0 = product dispatch (tcgen05), 1 = CUDA-core checker.  mode 0 linear, 1 conv 3x3, 2 temporal conv (3,1,1).
 * ln_stats ([rows] float2 mean,rstd) / ln_c1 ([N]): folded-LayerNorm epilogue (bias then carries W beta + bias), or NULL */
MUDG_EXPORT int mudg_test_tapgemm(const void* A, int B, int T, int H, int W, int Cin, int mode, const void* Wt, int N,
                                  void* D, const void* R, const float* bias, const float* bias2, int bias2_div, int nb2,
                                  float alpha, int geglu, const void* ln_stats, const float* ln_c1, int backend,
                                  void* stream);
MUDG_EXPORT int mudg_test_flash(const void* Q, int q_pitch, void* O, int o_pitch, int F, int Nq, int heads,
                                const void* K0, const void* V0, int pitch0, int len0, int nbatch0, int div0,
                                const void* K1, const void* V1, int pitch1, int len1, int nbatch1, int div1, float scale,
                                int backend, void* stream);
/* cross-attention to a per-frame context through the merged 96-key kernel (xattn.cu): Q / O [F][Nq][heads*64],
 * text_kv [F/T][77][2*heads*64], img_kv [F][16][2*heads*64] (K in the first half of a row, V in the second) */
MUDG_EXPORT int mudg_test_xattn(const void* Q, void* O, int F, int T, int Nq, int heads, const void* text_kv, const void* img_kv,
                                float scale, void* stream);
/* debug: device buffer [3][96][8] int64 receiving the clock64 time line of CTA 0 of the next flash launches (NULL = off) */
MUDG_EXPORT int mudg_test_flash_trace(void* buf);
/* debug: device buffer [4][64][8] int64 receiving the clock64 time line of CTA 0 of the next pair-GEMM launches */
MUDG_EXPORT int mudg_test_gemm_trace(void* buf);
/* debug: tcgen05.mma issue-rate probe; out = device int64 [ctas][2] (clocks until issued, until complete) */
MUDG_EXPORT int mudg_test_mma_probe(int variant, int reps, int ctas, int mode, void* out, void* stream);
/* debug: special-function-unit throughput (mode 0 ex2.f32, 1 ex2.f16x2, 2 rcp, 3 FFMA reference); 8 ops per thread and iteration */
MUDG_EXPORT int mudg_test_mufu_probe(int mode, int iters, int ctas, int threads, void* out, void* clocks, void* stream);
MUDG_EXPORT int mudg_test_temporal_attn(const void* qkv, void* out, int B, int T, int HW, int heads, float scale,
                                        void* stream);
MUDG_EXPORT int mudg_test_groupnorm(const void* x, void* y, int S, int64_t rows_per_sample, int C, const float* gamma,
                                    const float* beta, float eps, int silu, void* stream);
/* the one-kernel GroupNorm (statistics + apply from shared memory); fails when the sample does not fit */
MUDG_EXPORT int mudg_test_groupnorm_small(const void* x, void* y, int S, int64_t rows_per_sample, int C, const float* gamma,
                                          const float* beta, float eps, int silu, void* stream);
MUDG_EXPORT int mudg_test_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int C,
                                    void* stream);
MUDG_EXPORT int mudg_test_ln_stats(const void* x, void* mean_rstd, int64_t rows, int C, void* stream);
MUDG_EXPORT int mudg_test_ln_fold(void* W_f16, const float* gamma, const float* beta, const float* bias, float* c1,
                                  float* c2, int N, int K, void* stream);
#ifdef __cplusplus
}
#endif
#endif
