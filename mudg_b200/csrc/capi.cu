// extern "C" surface of libmudg_sm100.so (see include/mudg.h).  No exceptions cross the ABI.
#include <cmath>
#include <cstring>

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

// Every entry point that takes a context runs on the context's device, whatever the calling thread's current device is,
// and leaves the caller's current device unchanged.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) MUDG_CUDA(cudaSetDevice(dev));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

extern "C" {

MUDG_EXPORT const char* mudg_last_error(void) { return last_error_cstr(); }

MUDG_EXPORT int mudg_create(int device, const MudgUNetConfig* unet, const MudgVaeConfig* vae, MudgCtx** out) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(unet && vae && out, "null argument");
  DeviceGuard guard(device);
  int major = 0, minor = 0;
  MUDG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  MUDG_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  MUDG_REQUIRE(major == 10, "libmudg_sm100 needs a Blackwell sm_100 device (got sm_%d%d); there is no fallback", major, minor);
  *out = new MudgCtx(device, *unet, *vae);
  MUDG_API_END
}

MUDG_EXPORT void mudg_destroy(MudgCtx* ctx) {
  if (!ctx) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->model.device());
  cudaDeviceSynchronize();
  delete ctx;
  if (prev >= 0) cudaSetDevice(prev);
}

MUDG_EXPORT int mudg_load_weight(MudgCtx* ctx, int which, const char* key, const void* dev_ptr, int dtype,
                                 const int64_t* shape, int ndim, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && key && dev_ptr && shape, "null argument");
  MUDG_REQUIRE(which >= MUDG_UNET && which <= MUDG_CLIP_TEXT, "unknown weight set %d", which);
  DeviceGuard guard(ctx->model.device());
  ctx->model.begin_load(which);
  ctx->model.store(which).load(key, dev_ptr, dtype, shape, ndim, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_finalize_weights(MudgCtx* ctx, int which, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx, "null ctx");
  DeviceGuard guard(ctx->model.device());
  ctx->model.finalize(which, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_set_context(MudgCtx* ctx, const void* context, int dtype, int N, int L, int T, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && context, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.set_context(context, dtype, N, L, T, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_unet_forward(MudgCtx* ctx, const void* x, const int64_t* t, const int64_t* c_label,
                                  const int64_t* fs, int N, int T, int h, int w, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && t && c_label && fs && out, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.unet_forward(x, t, c_label, fs, N, 1, T, h, w, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_unet_forward_shared(MudgCtx* ctx, const void* x, const int64_t* t, const int64_t* c_label,
                                         const int64_t* fs, int N, int dup, int T, int h, int w, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && t && c_label && fs && out, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.unet_forward(x, t, c_label, fs, N, dup, T, h, w, out, S(stream));
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
  ProfScope ps(PF_SAMPLER, 0.0, (double)B * (double)n * (v_uncond ? 20.0 : 18.0), S(stream), "ddim_step");
  ddim_step(a, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_vae_decode(MudgCtx* ctx, const void* z, int F, int h, int w, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && z && out, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.vae_decode(z, F, h, w, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_vae_encode(MudgCtx* ctx, const void* x, int F, int H, int W, void* moments, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && moments, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.vae_encode(x, F, H, W, moments, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_postdecode(const void* frames, int dtype, int B, int T, int H, int W, const int* modes,
                                void* rgb_u8, void* depth_f32, void* class_u8, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(frames && modes && rgb_u8, "null argument");
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16 || dtype == MUDG_U8, "postdecode: dtype %d", dtype);
  MUDG_REQUIRE(H >= 1 && W >= 1, "postdecode: empty frame");
  ProfScope ps(PF_POST, 0.0, (double)B * T * H * W * 3.0 * ((dtype == MUDG_F32 ? 4.0 : dtype == MUDG_F16 ? 2.0 : 1.0) + 1.0), S(stream), "postdecode");
  postdecode(frames, dtype, static_cast<uint8_t*>(rgb_u8), static_cast<float*>(depth_f32),
             static_cast<uint8_t*>(class_u8), B, T, (int64_t)H * W, modes, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_clip_image_forward(MudgCtx* ctx, const void* img, int dtype, int B, int H, int W, int resize, int heads,
                                        void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && img && out, "null argument");
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16, "clip image: dtype %d", dtype);
  DeviceGuard guard(ctx->model.device());
  ctx->model.clip_image_forward(img, dtype, B, H, W, resize, heads, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_clip_text_forward(MudgCtx* ctx, const int64_t* tokens, int B, int L, int heads, int skip_last, void* out,
                                       void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && tokens && out, "null argument");
  DeviceGuard guard(ctx->model.device());
  ctx->model.clip_text_forward(tokens, B, L, heads, skip_last, out, S(stream));
  MUDG_API_END
}

MUDG_EXPORT int mudg_resampler_forward(MudgCtx* ctx, const void* x, int dtype, int B, int L, void* out, void* stream) {
  MUDG_API_BEGIN
  MUDG_REQUIRE(ctx && x && out, "null argument");
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16, "resampler: dtype %d", dtype);
  DeviceGuard guard(ctx->model.device());
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
    return ctx->model.plan_unet(N, 1, T, h, w);
  } catch (const std::exception& e) {
    mudg::set_last_error(e.what());
    return 0;
  }
}

MUDG_EXPORT int64_t mudg_launch_count(MudgCtx* ctx) { return ctx ? ctx->model.launches : 0; }

MUDG_EXPORT int mudg_profile(int enable) {
  MUDG_API_BEGIN
  prof_enable(enable != 0);
  MUDG_API_END
}

MUDG_EXPORT size_t mudg_profile_report(char* buf, size_t cap) {
  try {
    static thread_local std::string last;
    if (buf == nullptr || cap == 0) {      // first call: build the report, return the size needed (incl. the NUL)
      last = prof_report();
      return last.size() + 1;
    }
    const size_t n = last.size() < cap - 1 ? last.size() : cap - 1;
    memcpy(buf, last.data(), n);
    buf[n] = 0;
    return n + 1;
  } catch (const std::exception& e) {
    mudg::set_last_error(e.what());
    return 0;
  }
}

}  // extern "C"
