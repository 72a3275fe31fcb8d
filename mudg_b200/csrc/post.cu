// Post-decode frame pipeline ("next" row f.2 of SURVEY.md section 8): what the reference driver does on the CPU with the
// decoded clip before it writes PNG / npy files (virtual_render/virtual_pose_render.py:243, eval_tools.py:13-134, 206-347):
//   colour   : clamp(x,-1,1) -> (x+1)/2*255 -> uint8 (truncation)                        eval_tools.py:22-27
//   depth    : that uint8 frame -> mean over RGB / 255 (fp32, saved as npy) -> Spectral colour map -> uint8 RGB
//              eval_tools.py:70-74, colormap/method_custom :205-236, visualize_depth :282-289
//   semantic : that uint8 frame -> nearest of the 19 palette colours (L2, first minimum wins) -> class index + palette
//              colour                                                                       eval_tools.py:297-347
// One pass over the frame: HBM-bound byte work (6 B/pixel in as fp16, 3 B/pixel out + 4 B depth or 1 B class index).
// Every float operation is written with the explicitly rounded intrinsics in the reference's operation order (no FMA
// contraction), so the uint8 / fp32 / index outputs are bit-exact against the reference's CPU arithmetic.
#include "ops.h"

#include <algorithm>

namespace mudg {

namespace {

__constant__ float c_spectral[11][3] = {   // matplotlib "Spectral" anchors as the reference lists them (eval_tools.py:170-182)
    {(float)0.61960784313725492, (float)0.003921568627450980, (float)0.25882352941176473},
    {(float)0.83529411764705885, (float)0.24313725490196078, (float)0.30980392156862746},
    {(float)0.95686274509803926, (float)0.42745098039215684, (float)0.2627450980392157},
    {(float)0.99215686274509807, (float)0.68235294117647061, (float)0.38039215686274508},
    {(float)0.99607843137254903, (float)0.8784313725490196, (float)0.54509803921568623},
    {(float)1.0, (float)1.0, (float)0.74901960784313726},
    {(float)0.90196078431372551, (float)0.96078431372549022, (float)0.59607843137254901},
    {(float)0.6705882352941176, (float)0.8666666666666667, (float)0.64313725490196083},
    {(float)0.4, (float)0.76078431372549016, (float)0.6470588235294118},
    {(float)0.19607843137254902, (float)0.53333333333333333, (float)0.74117647058823533},
    {(float)0.36862745098039218, (float)0.30980392156862746, (float)0.63529411764705879}};

__constant__ int c_palette[19][3] = {   // eval_tools.py:312-332
    {255, 120, 50}, {255, 192, 203}, {255, 255, 0},  {0, 150, 245},   {0, 255, 255},  {255, 127, 0}, {255, 0, 0},
    {255, 240, 150}, {135, 60, 0},   {160, 32, 240}, {255, 0, 255},   {139, 137, 137}, {75, 0, 75},  {150, 240, 80},
    {230, 230, 250}, {0, 175, 0},    {0, 255, 127},  {222, 155, 161}, {140, 62, 69}};

// torch.clamp(x.float(), -1, 1) -> (x + 1.0) / 2.0 -> * 255 -> .to(uint8)   (NaN propagates through clamp; the cast of a
// NaN is undefined in the reference as well -- mapped to 0 here)
__device__ __forceinline__ uint32_t to_u8(float x) {
  x = (x != x) ? x : fminf(fmaxf(x, -1.f), 1.f);
  const float g = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.f), 2.f), 255.f);
  return (g != g) ? 0u : (uint32_t)(int)g;
}

__device__ __forceinline__ uint32_t to_u8(__half x) { return to_u8(__half2float(x)); }
__device__ __forceinline__ uint32_t to_u8(uint8_t x) { return x; }      // already converted (visualize_* on uint8 frames)

struct Px {
  uint32_t r, g, b;
};
struct PostModes {
  int8_t m[POSTDECODE_MAX_SAMPLES];
};

// colormap(..., "Spectral", bytes=True), method_custom (eval_tools.py:205-236): K = 11 anchors
__device__ __forceinline__ Px spectral_px(float d01) {
  const float pos = __fmul_rn(fminf(fmaxf(d01, 0.f), 1.f), 10.f);
  const int left = (int)pos;
  const int right = min(left + 1, 10);
  const float d = __fsub_rn(pos, (float)left);
  const float omd = __fsub_rn(1.f, d);
  Px o;
  uint32_t c[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float v = __fadd_rn(__fmul_rn(omd, c_spectral[left][k]), __fmul_rn(d, c_spectral[right][k]));
    c[k] = (uint32_t)(int)__fmul_rn(v, 255.f);
  }
  o.r = c[0]; o.g = c[1]; o.b = c[2];
  return o;
}

__device__ __forceinline__ Px depth_px(Px in, float* depth_out) {
  // torch.mean(frame.float(), dim=0) / 255: sum (exact), / 3, / 255 in fp32
  const float mean = __fdiv_rn((float)(in.r + in.g + in.b), 3.f);
  const float d01 = __fdiv_rn(mean, 255.f);
  if (depth_out) *depth_out = d01;
  return spectral_px(d01);
}

__global__ void spectral_kernel(const float* __restrict__ map, uint8_t* __restrict__ out_hwc, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const Px p = spectral_px(__ldg(map + i));
    out_hwc[3 * i] = (uint8_t)p.r; out_hwc[3 * i + 1] = (uint8_t)p.g; out_hwc[3 * i + 2] = (uint8_t)p.b;
  }
}

__device__ __forceinline__ Px semantic_px(Px in, uint8_t* cls_out) {
  // argmin_k || rgb - palette[k] ||_2, first minimum wins (np.argmin); sqrt is monotone and exact integers up to
  // 3*255^2 stay distinct in fp64, so the squared integer distance gives the same index
  int best = 0, bestd = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < 19; k++) {
    const int dr = (int)in.r - c_palette[k][0], dg = (int)in.g - c_palette[k][1], db = (int)in.b - c_palette[k][2];
    const int dd = dr * dr + dg * dg + db * db;
    if (dd < bestd) { bestd = dd; best = k; }
  }
  if (cls_out) *cls_out = (uint8_t)best;
  Px o;
  o.r = (uint32_t)c_palette[best][0]; o.g = (uint32_t)c_palette[best][1]; o.b = (uint32_t)c_palette[best][2];
  return o;
}

// frames: [B][3][T][H*W] (decode_first_stage layout, fp16 or fp32); rgb out: [B][T][3][H*W] uint8 (CHW per frame, what
// write_png takes); depth out [B][T][H*W] fp32; cls out [B][T][H*W] uint8.  mode[b]: 0 colour, 1 depth, 2 semantic.
// A thread converts VEC consecutive pixels of one frame: 128-bit loads per channel for fp16, 64-bit uint8 stores.
template <typename TIn, int VEC>
__global__ void __launch_bounds__(256)
postdecode_kernel(const TIn* __restrict__ frames, uint8_t* __restrict__ rgb, float* __restrict__ depth,
                  uint8_t* __restrict__ cls, int B, int T, int64_t HW, const __grid_constant__ PostModes modes) {
  const int64_t vec_per_frame = HW / VEC;
  const int64_t total = (int64_t)B * T * vec_per_frame;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i % vec_per_frame;
    const int64_t bt = i / vec_per_frame;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const int mode = modes.m[b];
    const int64_t px0 = v * VEC;
    const TIn* src = frames + (((int64_t)b * 3) * T + t) * HW + px0;       // channel stride T*HW
    const int64_t cs = (int64_t)T * HW;
    alignas(16) TIn ch[3][VEC];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (VEC * sizeof(TIn) == 16) {
        *reinterpret_cast<uint4*>(ch[c]) = __ldg(reinterpret_cast<const uint4*>(src + c * cs));
      } else if (VEC * sizeof(TIn) == 8) {
        *reinterpret_cast<uint2*>(ch[c]) = __ldg(reinterpret_cast<const uint2*>(src + c * cs));
      } else if (VEC * sizeof(TIn) == 32) {
        reinterpret_cast<uint4*>(ch[c])[0] = __ldg(reinterpret_cast<const uint4*>(src + c * cs));
        reinterpret_cast<uint4*>(ch[c])[1] = __ldg(reinterpret_cast<const uint4*>(src + c * cs) + 1);
      } else {
#pragma unroll
        for (int k = 0; k < VEC; k++) ch[c][k] = __ldg(src + c * cs + k);
      }
    }
    // outputs are packed into 32-bit words in registers (little endian: pixel k -> byte k & 3 of word k >> 2)
    constexpr int NW = (VEC + 3) / 4;
    uint32_t ow[3][NW], cw[NW];
    float dv[VEC];
#pragma unroll
    for (int w = 0; w < NW; w++) ow[0][w] = ow[1][w] = ow[2][w] = cw[w] = 0u;
#pragma unroll
    for (int k = 0; k < VEC; k++) {
      Px p;
      p.r = to_u8(ch[0][k]); p.g = to_u8(ch[1][k]); p.b = to_u8(ch[2][k]);
      uint8_t ck = 0;
      dv[k] = 0.f;
      if (mode == 1) p = depth_px(p, &dv[k]);
      else if (mode == 2) p = semantic_px(p, &ck);
      const int sh = 8 * (k & 3);
      ow[0][k >> 2] |= p.r << sh; ow[1][k >> 2] |= p.g << sh; ow[2][k >> 2] |= p.b << sh;
      cw[k >> 2] |= (uint32_t)ck << sh;
    }
    uint8_t* dst = rgb + (((int64_t)b * T + t) * 3) * HW + px0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (VEC == 8) *reinterpret_cast<uint2*>(dst + c * HW) = make_uint2(ow[c][0], ow[c][1]);
      else dst[c * HW] = (uint8_t)ow[c][0];
    }
    if (mode == 1 && depth) {
      float* dd = depth + ((int64_t)b * T + t) * HW + px0;
      if (VEC == 8) {
        reinterpret_cast<float4*>(dd)[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
        reinterpret_cast<float4*>(dd)[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
      } else {
        dd[0] = dv[0];
      }
    }
    if (mode == 2 && cls) {
      uint8_t* cd = cls + ((int64_t)b * T + t) * HW + px0;
      if (VEC == 8) *reinterpret_cast<uint2*>(cd) = make_uint2(cw[0], cw[1]);
      else cd[0] = (uint8_t)cw[0];
    }
  }
}

}  // namespace

void spectral_colormap(const float* map, uint8_t* out_hwc, int64_t n, cudaStream_t st) {
  MUDG_REQUIRE(n >= 1, "spectral_colormap: empty map");
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  spectral_kernel<<<blocks, 256, 0, st>>>(map, out_hwc, n);
  MUDG_CUDA(cudaGetLastError());
}

void postdecode(const void* frames, int dtype, uint8_t* rgb, float* depth, uint8_t* cls, int B, int T, int64_t HW,
                const int* modes_host, cudaStream_t st) {
  MUDG_REQUIRE(B >= 1 && T >= 1 && HW >= 1, "postdecode: empty clip");
  MUDG_REQUIRE(dtype >= 0 && dtype <= 2, "postdecode: dtype %d (0 fp32, 1 fp16, 2 uint8)", dtype);
  MUDG_REQUIRE(B <= POSTDECODE_MAX_SAMPLES, "postdecode: at most %d samples per call, got %d", POSTDECODE_MAX_SAMPLES, B);
  PostModes m{};
  for (int b = 0; b < B; b++) {
    MUDG_REQUIRE(modes_host[b] >= 0 && modes_host[b] <= 2, "postdecode: mode %d (0 colour, 1 depth, 2 semantic)", modes_host[b]);
    m.m[b] = (int8_t)modes_host[b];
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(rgb) |
                         reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(cls)) & 15) == 0;
  const bool vec8 = aligned && HW % 8 == 0;
  const int64_t work = (int64_t)B * T * (vec8 ? HW / 8 : HW);
  const int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 8);
#define MUDG_POST_LAUNCH(TIN, V) \
  postdecode_kernel<TIN, V><<<blocks, 256, 0, st>>>(static_cast<const TIN*>(frames), rgb, depth, cls, B, T, HW, m)
  if (vec8) {
    if (dtype == 0) MUDG_POST_LAUNCH(float, 8);
    else if (dtype == 1) MUDG_POST_LAUNCH(__half, 8);
    else MUDG_POST_LAUNCH(uint8_t, 8);
  } else {
    if (dtype == 0) MUDG_POST_LAUNCH(float, 1);
    else if (dtype == 1) MUDG_POST_LAUNCH(__half, 1);
    else MUDG_POST_LAUNCH(uint8_t, 1);
  }
#undef MUDG_POST_LAUNCH
  MUDG_CUDA(cudaGetLastError());
}

}  // namespace mudg
