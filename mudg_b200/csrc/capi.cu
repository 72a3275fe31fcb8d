// extern "C" surface of libmudg_sm100.so (see include/mudg.h).  No exceptions cross the ABI.
#include <cmath>

#include "model.h"
#include "mudg.h"

namespace mudg {
const char* last_error_cstr();
}
using namespace mudg;

struct MudgCtx {
  Model model;
  MudgCtx(int device, const MudgUNetConfig& u, const MudgVaeConfig& v) : model(device, u, v) {}
};

#define MUDG_API_BEGIN try {
#define MUDG_API_END                          \
  return 0;                                   \
  }                                           \
  catch (const std::exception& e) {           \
    mudg::set_last_error(e.what());           \
    return -1;                                \
  }                                           \
  catch (...) {                               \
    mudg::set_last_error("unknown exception"); \
    return -2;                                \
  }

static cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

MUDG_EXPORT const char* mudg_last_error(void) { return last_error_cstr(); }

MUDG_EXPORT int mudg_create(int device, const MudgUNetConfig* unet, const MudgVaeConfig* vae, MudgCtx** out) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(unet && vae && out, "null argument");
  MUDG_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  MUDG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  MUDG_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  MUDG_REQUIRE(major == 10, "libmudg_sm100 needs a Blackwell sm_100 device (got sm_%d%d); there is no fallback", major, minor);
  *out = new MudgCtx(device, *unet, *vae);
  MUDG_API_END
}

MUDG_EXPORT void mudg_destroy(MudgCtx* ctx) {
  if (!ctx) return;
  cudaDeviceSynchronize();
  delete ctx;
}

MUDG_EXPORT int mudg_load_weight(MudgCtx* ctx, int which, const char* key, const void* dev_ptr, int dtype,
                                 const int64_t* shape, int ndim, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && key && dev_ptr && shape, "null argument");
  MUDG_REQUIRE(which == MUDG_UNET || which == MUDG_VAE || which == MUDG_RESAMPLER, "unknown weight set %d", which);
  WeightStore& ws = which == MUDG_VAE ? ctx->model.vae_w : (which == MUDG_RESAMPLER ? ctx->model.res_w : ctx->model.unet_w);
  ws.load(key, dev_ptr, dtype, shape, ndim, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_finalize_weights(MudgCtx* ctx, int which, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx, "null ctx");
  ctx->model.finalize(which, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_set_context(MudgCtx* ctx, const void* context, int dtype, int N, int L, int T, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && context, "null argument");
  ctx->model.set_context(context, dtype, N, L, T, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_unet_forward(MudgCtx* ctx, const void* x, const int64_t* t, const int64_t* c_label,
                                  const int64_t* fs, int N, int T, int h, int w, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && t && c_label && fs && out, "null argument");
  ctx->model.unet_forward(x, t, c_label, fs, N, T, h, w, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_ddim_step(const void* x, const void* v_cond, const void* v_uncond, const void* noise,
                               void* x_prev, void* pred_x0, int B, int64_t n, float cfg_scale, float guidance_rescale,
                               float sqrt_ac, float sqrt_1mac, float rescale, float a_prev, float sigma_t, void* stream) {
  MUDG_API_BEGIN
  DdimStepArgs a;
  a.x = static_cast<const float*>(x);
  a.v_cond = static_cast<const __half*>(v_cond);
  a.v_uncond = static_cast<const __half*>(v_uncond);
  a.noise = static_cast<const float*>(noise);
  a.x_prev = static_cast<float*>(x_prev);
  a.pred_x0 = static_cast<float*>(pred_x0);
  a.B = B; a.n = n; a.cfg_scale = cfg_scale; a.guidance_rescale = guidance_rescale;
  a.sqrt_ac = sqrt_ac; a.sqrt_1mac = sqrt_1mac; a.rescale = rescale;
  a.sqrt_a_prev = sqrtf(a_prev);
  a.dir_coef = sqrtf(1.f - a_prev - sigma_t * sigma_t);
  a.sigma = sigma_t;
  ddim_step(a, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_vae_decode(MudgCtx* ctx, const void* z, int F, int h, int w, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && z && out, "null argument");
  ctx->model.vae_decode(z, F, h, w, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_vae_encode(MudgCtx* ctx, const void* x, int F, int H, int W, void* moments, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && moments, "null argument");
  ctx->model.vae_encode(x, F, H, W, moments, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_postdecode(const void* frames, int dtype, int B, int T, int H, int W, const int* modes,
                                void* rgb_u8, void* depth_f32, void* class_u8, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(frames && modes && rgb_u8, "null argument");
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16 || dtype == MUDG_U8, "postdecode: dtype %d", dtype);
  MUDG_REQUIRE(H >= 1 && W >= 1, "postdecode: empty frame");
  postdecode(frames, dtype, static_cast<uint8_t*>(rgb_u8), static_cast<float*>(depth_f32),
             static_cast<uint8_t*>(class_u8), B, T, (int64_t)H * W, modes, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_resampler_forward(MudgCtx* ctx, const void* x, int dtype, int B, int L, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && out, "null argument");
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16, "resampler: dtype %d", dtype);
  ctx->model.resampler_forward(x, dtype, B, L, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_colormap_spectral(const void* map_f32, int64_t n, void* out_u8, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(map_f32 && out_u8, "null argument");
  spectral_colormap(static_cast<const float*>(map_f32), static_cast<uint8_t*>(out_u8), n, S(stream));
  MUDG_API_END
}

MUDG_EXPORT size_t mudg_workspace_bytes(MudgCtx* ctx, int N, int T, int h, int w) {
  try {
    return ctx->model.plan_unet(N, T, h, w);
  } catch (const std::exception& e) {
    mudg::set_last_error(e.what());
    return 0;
  }
}

MUDG_EXPORT int64_t mudg_launch_count(MudgCtx* ctx) { return ctx ? ctx->model.launches : 0; }

MUDG_EXPORT int mudg_profile_gemm(int enable) {
  MUDG_API_BEGIN
  gemm_profile_enable(enable != 0);
  MUDG_API_END
}

MUDG_EXPORT int mudg_profile_gemm_read(double* ms_total, double* flops_total, int64_t* launches) {
  MUDG_API_BEGIN
  gemm_profile_read(ms_total, flops_total, launches);
  MUDG_API_END
}

// ------------------------------------------------------------------ test hooks
MUDG_EXPORT int mudg_test_tapgemm(const void* A, int B, int T, int H, int W, int Cin, int mode, const void* Wt, int N,
                                  void* D, const void* R, const float* bias, const float* bias2, int bias2_div, int nb2,
                                  float alpha, int geglu, int backend, void* stream) {
  MUDG_API_BEGIN
  TapGemm g;
  g.A = static_cast<const __half*>(A);
  g.B = B; g.T = T; g.H = H; g.W = W; g.Cin = Cin;
  if (mode == 0) { g.ntaps = 1; g.taps[0][0] = g.taps[0][1] = g.taps[0][2] = 0; }
  else if (mode == 1) { g.ntaps = 9; set_taps_3x3(g.taps); }
  else { g.ntaps = 3; set_taps_t3(g.taps); }
  g.Wt = static_cast<const __half*>(Wt);
  g.N = N;
  g.D = static_cast<__half*>(D);
  g.R = static_cast<const __half*>(R);
  g.bias = bias; g.bias2 = bias2; g.bias2_div = bias2_div; g.nb2 = nb2;
  g.alpha = alpha; g.geglu = geglu != 0;
  if (backend == 0) tapgemm_tc2(g, S(stream));
  else if (backend == 2) tapgemm_tc(g, S(stream));
  else tapgemm_simt(g, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_flash(const void* Q, int q_pitch, void* O, int o_pitch, int F, int Nq, int heads,
                                const void* K0, const void* V0, int pitch0, int len0, int nbatch0, int div0,
                                const void* K1, const void* V1, int pitch1, int len1, int nbatch1, int div1, float scale,
                                int backend, void* stream) {
  MUDG_API_BEGIN
  FlashArgs a;
  a.Q = static_cast<const __half*>(Q); a.q_pitch = q_pitch;
  a.O = static_cast<__half*>(O); a.o_pitch = o_pitch;
  a.F = F; a.Nq = Nq; a.heads = heads; a.scale = scale;
  a.nseg = K1 ? 2 : 1;
  a.seg[0].K = static_cast<const __half*>(K0); a.seg[0].V = static_cast<const __half*>(V0);
  a.seg[0].pitch = pitch0; a.seg[0].len = len0; a.seg[0].nbatch = nbatch0; a.seg[0].kv_div = div0;
  a.seg[1].K = static_cast<const __half*>(K1); a.seg[1].V = static_cast<const __half*>(V1);
  a.seg[1].pitch = pitch1; a.seg[1].len = len1; a.seg[1].nbatch = nbatch1; a.seg[1].kv_div = div1;
  if (backend == 0) {
    // the tcgen05 kernel reads V transposed: build V^T of each segment in a (grow-only) scratch buffer of the test hook
    static __half* scratch[2] = {nullptr, nullptr};
    static size_t scratch_bytes[2] = {0, 0};
    for (int i = 0; i < a.nseg; i++) {
      FlashSeg& sg = a.seg[i];
      const int pad = (sg.len + 7) / 8 * 8;
      const size_t need = sizeof(__half) * (size_t)sg.nbatch * heads * 64 * pad;
      if (scratch_bytes[i] < need) {
        MUDG_CUDA(cudaDeviceSynchronize());
        cudaFree(scratch[i]);
        MUDG_CUDA(cudaMalloc(&scratch[i], need));
        scratch_bytes[i] = need;
      }
      transpose_v(sg.V, sg.pitch, sg.len, sg.nbatch, heads, scratch[i], pad, S(stream));
      sg.VT = scratch[i];
      sg.vt_pitch = pad;
    }
    flash_attention(a, S(stream));
  } else {
    flash_attention_simt(a, S(stream));
  }
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_gemm_trace(void* buf) {
  MUDG_API_BEGIN
  gemm_set_trace(static_cast<long long*>(buf));
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_mma_probe(int variant, int reps, int ctas, int mode, void* out, void* stream) {
  MUDG_API_BEGIN
  mma_probe(variant, reps, ctas, mode, static_cast<long long*>(out), S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_flash_trace(void* buf) {
  MUDG_API_BEGIN
  flash_set_trace(static_cast<long long*>(buf));
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_temporal_attn(const void* qkv, void* out, int B, int T, int HW, int heads, float scale,
                                        void* stream) {
  MUDG_API_BEGIN
  temporal_attention(static_cast<const __half*>(qkv), static_cast<__half*>(out), B, T, HW, heads, scale, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_groupnorm(const void* x, void* y, int Sn, int64_t rows_per_sample, int C, const float* gamma,
                                    const float* beta, float eps, int silu, void* stream) {
  MUDG_API_BEGIN
  double* sums = nullptr;
  float* ss = nullptr;
  MUDG_CUDA(cudaMalloc(&sums, sizeof(double) * Sn * 64));
  MUDG_CUDA(cudaMalloc(&ss, sizeof(float) * Sn * C * 2));
  gn_scale_shift(static_cast<const __half*>(x), Sn, rows_per_sample, C, gamma, beta, eps, sums, ss, ss + (size_t)Sn * C,
                 S(stream));
  gn_apply(static_cast<const __half*>(x), static_cast<__half*>(y), ss, ss + (size_t)Sn * C, (int64_t)Sn * rows_per_sample, C,
           rows_per_sample, silu != 0, S(stream));
  MUDG_CUDA(cudaStreamSynchronize(S(stream)));
  cudaFree(sums);
  cudaFree(ss);
  MUDG_API_END
}

MUDG_EXPORT int mudg_test_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int C,
                                    void* stream) {
  MUDG_API_BEGIN
  layernorm(static_cast<const __half*>(x), static_cast<__half*>(y), gamma, beta, rows, C, 1e-5f, S(stream));
  MUDG_API_END
}

}  // extern "C"
