// OpenCLIP ViT-H/14 towers as the reference drives them once per clip ("next" row f.3):
//   FrozenOpenCLIPImageEmbedderV2.encode_with_vision_transformer  (lvdm/modules/encoders/condition.py:339-372)
//   FrozenOpenCLIPImageEmbedderV2.preprocess                      (condition.py:318-326: kornia resize + CLIP normalise)
//   FrozenOpenCLIPEmbedder.encode_with_transformer                (condition.py:214-232, layer = "penultimate")
// The arithmetic behind those call sites lives in open_clip (VisionTransformer / Transformer / ResidualAttentionBlock over
// torch.nn.MultiheadAttention), which the reference imports and does not vendor; weights use open_clip's state-dict names
// below `model.visual.` / `model.`.  Every Linear and the 14 x 14 patch convolution (as an im2col GEMM) run on the tcgen05
// tap-GEMM with bias / residual fused; LayerNorm and erf GELU are the UNet's kernels.  The attention is new: 257 tokens at
// head dim 80 (image) and 77 causal tokens at head dim 64 (text) fit no tile of the d = 64 flash kernel, and the whole
// tower is ~0.3 TFLOP once per clip, so it is one CUDA-core kernel with K / V of a head staged in shared memory and fp32
// scores -- weight bandwidth (1.26 GB + 0.6 GB of fp16) bounds the call, not this kernel.
#include "model.h"

namespace mudg {
namespace {

// ---------------------------------------------------------------- image -> patch rows
struct ClipPrepArgs {
  int B, H, W, S, P;            // source H x W, tower input S x S, patch P
  int resize;                   // 1: kornia resize (+ anti-alias blur) + (x + 1) / 2 + normalise; 0: source is S x S, ready
  int ky, kx;                   // gaussian taps per axis (1 = no blur)
  float gy[65], gx[65];
  float mean[3], inv_std[3];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {       // F.pad(mode="reflect"): no edge repeat
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

template <typename TS>
__device__ __forceinline__ float blurred(const TS* __restrict__ pl, int y, int x, const ClipPrepArgs& a) {
  if (a.ky == 1 && a.kx == 1) return (float)pl[(size_t)y * a.W + x];
  float acc = 0.f;
  for (int dy = 0; dy < a.ky; dy++) {
    const TS* row = pl + (size_t)reflect_idx(y + dy - a.ky / 2, a.H) * a.W;
    float r = 0.f;
    for (int dx = 0; dx < a.kx; dx++) r += a.gx[dx] * (float)row[reflect_idx(x + dx - a.kx / 2, a.W)];
    acc += a.gy[dy] * r;
  }
  return acc;
}

// One thread per pixel of the S x S tower input: A[(b, py, px)][(iy, ix)][c of 8] (fp16; the weight is packed [O][P*P][8]).
template <typename TS>
__global__ void clip_patches_kernel(const TS* __restrict__ img, __half* __restrict__ A, ClipPrepArgs a) {
  const int64_t n = (int64_t)a.B * 3 * a.S * a.S;
  const int G = a.S / a.P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.S), y = (int)((i / a.S) % a.S), c = (int)((i / ((int64_t)a.S * a.S)) % 3);
    const int b = (int)(i / ((int64_t)3 * a.S * a.S));
    const TS* pl = img + ((size_t)b * 3 + c) * a.H * a.W;
    float v;
    if (!a.resize) {
      v = (float)pl[(size_t)y * a.W + x];
    } else {
      // F.interpolate(mode="bicubic", align_corners=True): src = dst * (in - 1) / (out - 1), A = -0.75, clamped taps
      const float sy = a.S > 1 ? (float)(a.H - 1) / (float)(a.S - 1) : 0.f, sx = a.S > 1 ? (float)(a.W - 1) / (float)(a.S - 1) : 0.f;
      const float ry = sy * y, rx = sx * x;
      const int iy = (int)floorf(ry), ix = (int)floorf(rx);
      const float ty = ry - iy, tx = rx - ix;
      const float A_ = -0.75f;
      const float wy[4] = {cubic2(ty + 1.f, A_), cubic1(ty, A_), cubic1(1.f - ty, A_), cubic2(2.f - ty, A_)};
      const float wx[4] = {cubic2(tx + 1.f, A_), cubic1(tx, A_), cubic1(1.f - tx, A_), cubic2(2.f - tx, A_)};
      v = 0.f;
      for (int j = 0; j < 4; j++) {
        const int yy = min(max(iy - 1 + j, 0), a.H - 1);
        float r = 0.f;
        for (int k = 0; k < 4; k++) r += wx[k] * blurred(pl, yy, min(max(ix - 1 + k, 0), a.W - 1), a);
        v += wy[j] * r;
      }
      v = ((v + 1.f) * 0.5f - a.mean[c]) * a.inv_std[c];
    }
    const int64_t row = ((int64_t)b * G + y / a.P) * G + x / a.P;
    A[(row * a.P * a.P + (y % a.P) * a.P + x % a.P) * 8 + c] = __float2half(v);
  }
}

// x[b][0] = class_embedding + pos[0];  x[b][1 + i] = patch[b][i] + pos[1 + i]   (condition.py:355-359)
__global__ void clip_tokens_kernel(const __half* __restrict__ patches, const float* __restrict__ cls, const __half* __restrict__ pos,
                                   __half* __restrict__ x, int B, int tokens, int width) {
  const int64_t n = (int64_t)B * tokens * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % width), t = (int)((i / width) % tokens), b = (int)(i / ((int64_t)width * tokens));
    const float base = t == 0 ? cls[c] : __half2float(patches[((int64_t)b * (tokens - 1) + t - 1) * width + c]);
    x[i] = __float2half(base + __half2float(pos[(int64_t)t * width + c]));
  }
}

// x[b][l] = token_embedding[tokens[b][l]] + positional_embedding[l]   (condition.py:215-216)
__global__ void clip_text_embed_kernel(const int64_t* __restrict__ tok, const __half* __restrict__ table, const __half* __restrict__ pos,
                                       __half* __restrict__ x, int B, int L, int width, int vocab) {
  const int64_t n = (int64_t)B * L * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % width), l = (int)((i / width) % L);
    int64_t id = tok[i / width];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);       // the host wrapper rejects out-of-range ids before the call
    x[i] = __float2half(__half2float(table[id * width + c]) + __half2float(pos[(int64_t)l * width + c]));
  }
}

// ---------------------------------------------------------------- multi-head attention, any head dim <= 128
// torch.nn.MultiheadAttention on the packed projection qkv [B * L][3 * heads * d] (q | k | v, head h at columns h * d):
// softmax(q k^T / sqrt(d) (+ causal mask)) v.  One CTA per (32-query tile, head, sample); K rows padded to an odd word
// stride so that the 32 lanes (one key each) read conflict-free; scores, softmax and the P V sum in fp32.
constexpr int MHA_WARPS = 4, MHA_QT = 32;

__global__ void __launch_bounds__(MHA_WARPS * 32) mha_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L,
                                                              int heads, int d, float scale, int causal) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int Wd = heads * d, pitch = 3 * Wd, ks = d + 2;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * MHA_QT;
  const int q1 = min(q0 + MHA_QT, L);
  const int nkeys = causal ? q1 : L;                         // keys any query of this tile can see
  __half* Vs = reinterpret_cast<__half*>(smem);              // [L][d]       (16 B aligned rows)
  __half* Ks = Vs + (size_t)L * d;                           // [L][d + 2]
  float* qs = reinterpret_cast<float*>(Ks + (size_t)L * ks); // [MHA_WARPS][d]
  float* ps = qs + MHA_WARPS * d;                            // [MHA_WARPS][L]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = d / 8;
  for (int i = threadIdx.x; i < nkeys * chunks; i += blockDim.x) {
    const int k = i / chunks, c = i % chunks;
    const __half* src = qkv + ((size_t)b * L + k) * pitch + (size_t)h * d + c * 8;
    const uint4 kv = *reinterpret_cast<const uint4*>(src + Wd);
    *reinterpret_cast<uint4*>(Vs + (size_t)k * d + c * 8) = *reinterpret_cast<const uint4*>(src + 2 * Wd);
    uint32_t* kd = reinterpret_cast<uint32_t*>(Ks + (size_t)k * ks + c * 8);
    kd[0] = kv.x; kd[1] = kv.y; kd[2] = kv.z; kd[3] = kv.w;
  }
  __syncthreads();
  float* q = qs + warp * d;
  float* p = ps + (size_t)warp * L;
  for (int qi = q0 + warp; qi < q1; qi += MHA_WARPS) {
    const __half* qsrc = qkv + ((size_t)b * L + qi) * pitch + (size_t)h * d;
    for (int i = lane; i < d; i += 32) q[i] = __half2float(qsrc[i]) * scale;
    __syncwarp();
    const int nk = causal ? qi + 1 : L;
    float m = -INFINITY;
    for (int k = lane; k < nk; k += 32) {
      const __half2* kr = reinterpret_cast<const __half2*>(Ks + (size_t)k * ks);
      float s = 0.f;
      for (int i = 0; i < d / 2; i++) {
        const float2 kk = __half22float2(kr[i]);
        s = fmaf(q[2 * i], kk.x, s);
        s = fmaf(q[2 * i + 1], kk.y, s);
      }
      p[k] = s;
      m = fmaxf(m, s);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int k = lane; k < nk; k += 32) {
      const float e = expf(p[k] - m);
      p[k] = e;
      sum += e;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nk; k++) {
      const float pk = p[k];
      const __half* vr = Vs + (size_t)k * d;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (lane + 32 * j < d) acc[j] = fmaf(pk, __half2float(vr[lane + 32 * j]), acc[j]);
    }
    const float inv = 1.f / sum;
    __half* o = out + ((size_t)b * L + qi) * Wd + (size_t)h * d;
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (lane + 32 * j < d) o[lane + 32 * j] = __float2half(acc[j] * inv);
    __syncwarp();                                            // q / p are rewritten by the next query
  }
}

void mha(const __half* qkv, __half* out, int B, int L, int heads, int d, bool causal, cudaStream_t st) {
  MUDG_REQUIRE(d % 8 == 0 && d <= 128, "attention head dim %d (multiple of 8, at most 128)", d);
  const size_t smem = sizeof(__half) * ((size_t)L * d + (size_t)L * (d + 2)) + sizeof(float) * MHA_WARPS * ((size_t)d + L);
  MUDG_REQUIRE(smem <= 200 * 1024, "attention over %d tokens at head dim %d needs %zu B of shared memory", L, d, smem);
  MUDG_CUDA(cudaFuncSetAttribute(mha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));   // per device
  dim3 grid((L + MHA_QT - 1) / MHA_QT, heads, B);
  mha_kernel<<<grid, MHA_WARPS * 32, smem, st>>>(qkv, out, L, heads, d, 1.f / sqrtf((float)d), causal ? 1 : 0);
  MUDG_CUDA(cudaGetLastError());
}

int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 16); }

}  // namespace

// ---------------------------------------------------------------- weights
void Model::finalize_clip(int which) {
  const bool vision = which == MUDG_CLIP_IMAGE;
  WeightStore& w = vision ? clipv_w : clipt_w;
  ClipDims c;
  if (vision) {
    const Weight& conv = w.W("conv1.weight");
    MUDG_REQUIRE(conv.I == 3 && conv.Ipad == 8, "CLIP conv1: %d input channels", conv.I);
    c.width = conv.O;
    c.patch = (int)lround(sqrt((double)conv.taps));
    MUDG_REQUIRE(c.patch * c.patch == conv.taps, "CLIP conv1: %d taps is not a square patch", conv.taps);
    const Weight& pos = w.W("positional_embedding");
    c.tokens = pos.O;
    c.grid = (int)lround(sqrt((double)(c.tokens - 1)));
    MUDG_REQUIRE(c.grid * c.grid + 1 == c.tokens && pos.I == c.width, "CLIP positional_embedding [%d][%d]", pos.O, pos.I);
    MUDG_REQUIRE(w.V("class_embedding").n == c.width, "CLIP class_embedding width");
    w.V("ln_pre.weight"); w.V("ln_pre.bias");
  } else {
    const Weight& tab = w.W("token_embedding.weight");
    c.vocab = tab.O; c.width = tab.I;
    const Weight& pos = w.W("positional_embedding");
    c.tokens = pos.O;
    MUDG_REQUIRE(pos.I == c.width, "CLIP text positional_embedding width %d vs %d", pos.I, c.width);
    w.V("ln_final.weight"); w.V("ln_final.bias");
  }
  MUDG_REQUIRE(c.width % 64 == 0, "CLIP width %d must be a multiple of 64 (tcgen05 GEMM tiles)", c.width);
  while (w.hasW("transformer.resblocks." + std::to_string(c.layers) + ".attn.in_proj_weight")) {
    const std::string p = "transformer.resblocks." + std::to_string(c.layers);
    MUDG_REQUIRE(w.W(p + ".attn.in_proj_weight").O == 3 * c.width && w.W(p + ".attn.in_proj_weight").I == c.width &&
                     w.V(p + ".attn.in_proj_bias").n == 3 * c.width && w.W(p + ".attn.out_proj.weight").O == c.width,
                 "CLIP block %d: attention projection shapes", c.layers);
    const int mlp = w.W(p + ".mlp.c_fc.weight").O;
    MUDG_REQUIRE(mlp % 64 == 0 && w.W(p + ".mlp.c_proj.weight").I == mlp && (c.layers == 0 || mlp == c.mlp), "CLIP block %d: mlp width", c.layers);
    c.mlp = mlp;
    for (const char* n : {".ln_1", ".ln_2"}) { w.V(p + n + ".weight"); w.V(p + n + ".bias"); }
    w.V(p + ".attn.out_proj.bias"); w.V(p + ".mlp.c_fc.bias"); w.V(p + ".mlp.c_proj.bias");
    c.layers++;
  }
  MUDG_REQUIRE(c.layers > 0, "CLIP tower: no transformer.resblocks.* loaded");
  (vision ? cv_ : ct_) = c;
  (vision ? clipv_ready_ : clipt_ready_) = true;
}

// ---------------------------------------------------------------- graph
// open_clip ResidualAttentionBlock (pre-LN): x += out_proj(MHA(ln_1 x));  x += c_proj(gelu(c_fc(ln_2 x)))
Act Model::clip_block(Act x, const std::string& p, int B, int L, int heads, bool causal) {
  const int width = x.C;
  Act h = layer_norm(x, p + ".ln_1");
  Act qkv = linear(h, p + ".attn.in_proj_weight", p + ".attn.in_proj_bias", nullptr);
  release(h);
  Act o = alloc(1, 1, 1, B * L, width);
  if (live()) {
    mha(qkv.p, o.p, B, L, heads, width / heads, causal, st_);
    launches++;
  }
  release(qkv);
  Act x2 = linear(o, p + ".attn.out_proj.weight", p + ".attn.out_proj.bias", &x);
  release(o);
  release(x);
  Act h2 = layer_norm(x2, p + ".ln_2");
  Act f = linear(h2, p + ".mlp.c_fc.weight", p + ".mlp.c_fc.bias", nullptr);
  release(h2);
  if (live()) {
    gelu_inplace(f.p, f.numel(), st_);
    launches++;
  }
  Act y = linear(f, p + ".mlp.c_proj.weight", p + ".mlp.c_proj.bias", &x2);
  release(f);
  release(x2);
  return y;
}

void Model::clip_image_body(const void* img, int dtype, int B, int H, int W, int resize, int heads, void* out) {
  ws_ = &clipv_w;
  arena_.reset();
  const ClipDims& c = cv_;
  const int S = c.grid * c.patch, K = c.patch * c.patch * 8;
  Act A = alloc(1, 1, 1, B * c.grid * c.grid, K);
  if (live()) {
    ClipPrepArgs a{};
    a.B = B; a.H = H; a.W = W; a.S = S; a.P = c.patch; a.resize = resize; a.ky = a.kx = 1;
    a.gy[0] = a.gx[0] = 1.f;
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, sd[3] = {0.26862954f, 0.26130258f, 0.27577711f};   // condition.py:311-312
    for (int i = 0; i < 3; i++) { a.mean[i] = mean[i]; a.inv_std[i] = 1.f / sd[i]; }
    if (resize) {
      // kornia.geometry.resize(antialias=True): blur only when down-scaling; per axis sigma = max((factor - 1) / 2, 0.001),
      // taps = int(max(4 sigma, 3)) made odd
      const double fy = (double)H / S, fx = (double)W / S;
      if (std::max(fy, fx) > 1.0) {
        auto taps = [](double f, float* g) {
          const double sigma = std::max((f - 1.0) / 2.0, 0.001);
          int k = (int)std::max(2.0 * 2 * sigma, 3.0);
          k += 1 - k % 2;
          MUDG_REQUIRE(k <= 65, "CLIP preprocess: down-scale factor %.1f needs %d blur taps (at most 65)", f, k);
          double sum = 0.0, v[65];
          for (int i = 0; i < k; i++) { const double x = i - k / 2; v[i] = exp(-x * x / (2.0 * sigma * sigma)); sum += v[i]; }
          for (int i = 0; i < k; i++) g[i] = (float)(v[i] / sum);
          return k;
        };
        a.ky = taps(fy, a.gy);
        a.kx = taps(fx, a.gx);
      }
    }
    MUDG_CUDA(cudaMemsetAsync(A.p, 0, A.bytes(), st_));       // channels 3..7 of every tap stay zero
    const int64_t n = (int64_t)B * 3 * S * S;
    if (dtype == MUDG_F32) clip_patches_kernel<float><<<grid_for(n), 256, 0, st_>>>(static_cast<const float*>(img), A.p, a);
    else clip_patches_kernel<__half><<<grid_for(n), 256, 0, st_>>>(static_cast<const __half*>(img), A.p, a);
    MUDG_CUDA(cudaGetLastError());
    launches += 2;
  }
  const Weight* conv = live() ? &clipv_w.W("conv1.weight") : nullptr;
  Act patches = gemm_raw(A.p, B * c.grid * c.grid, K, conv ? conv->w : nullptr, c.width, nullptr, nullptr, 1.f);   // conv1 has no bias
  release(A);
  Act x = alloc(1, 1, 1, B * c.tokens, c.width);
  if (live()) {
    clip_tokens_kernel<<<grid_for(x.numel()), 256, 0, st_>>>(patches.p, clipv_w.V("class_embedding").p,
                                                              clipv_w.W("positional_embedding").w, x.p, B, c.tokens, c.width);
    MUDG_CUDA(cudaGetLastError());
    launches++;
  }
  release(patches);
  Act y = layer_norm(x, "ln_pre");
  release(x);
  for (int i = 0; i < c.layers; i++) y = clip_block(y, "transformer.resblocks." + std::to_string(i), B, c.tokens, heads, false);
  if (live()) {
    cast_to_f32(y.p, false, static_cast<float*>(out), y.numel(), st_);
    launches++;
  }
  release(y);
}

void Model::clip_text_body(const int64_t* tokens, int B, int L, int heads, int skip_last, void* out) {
  ws_ = &clipt_w;
  arena_.reset();
  const ClipDims& c = ct_;
  Act x = alloc(1, 1, 1, B * L, c.width);
  if (live()) {
    clip_text_embed_kernel<<<grid_for(x.numel()), 256, 0, st_>>>(tokens, clipt_w.W("token_embedding.weight").w,
                                                                  clipt_w.W("positional_embedding").w, x.p, B, L, c.width, c.vocab);
    MUDG_CUDA(cudaGetLastError());
    launches++;
  }
  for (int i = 0; i < c.layers - skip_last; i++)              // text_transformer_forward stops `layer_idx` blocks early
    x = clip_block(x, "transformer.resblocks." + std::to_string(i), B, L, heads, true);
  Act y = layer_norm(x, "ln_final");
  release(x);
  if (live()) {
    cast_to_f32(y.p, false, static_cast<float*>(out), y.numel(), st_);
    launches++;
  }
  release(y);
}

void Model::clip_image_forward(const void* img, int dtype, int B, int H, int W, int resize, int heads, void* out, cudaStream_t st) {
  MUDG_REQUIRE(clipv_ready_, "CLIP image tower weights not finalized");
  const ClipDims& c = cv_;
  MUDG_REQUIRE(B >= 1 && H >= 1 && W >= 1, "CLIP image: empty input");
  MUDG_REQUIRE(heads >= 1 && c.width % heads == 0, "CLIP image: %d heads do not divide width %d", heads, c.width);
  MUDG_REQUIRE(resize || (H == c.grid * c.patch && W == c.grid * c.patch), "CLIP image: %d x %d input without resize (tower takes %d x %d)",
               H, W, c.grid * c.patch, c.grid * c.patch);
  MUDG_REQUIRE(!resize || (H >= 2 && W >= 2), "CLIP image: resize needs at least 2 x 2 pixels");
  run_planned([&](bool plan) { clip_image_body(plan ? nullptr : img, dtype, B, H, W, resize, heads, plan ? nullptr : out); }, st);
}

void Model::clip_text_forward(const int64_t* tokens, int B, int L, int heads, int skip_last, void* out, cudaStream_t st) {
  MUDG_REQUIRE(clipt_ready_, "CLIP text tower weights not finalized");
  const ClipDims& c = ct_;
  MUDG_REQUIRE(B >= 1 && L >= 1 && L <= c.tokens, "CLIP text: %d tokens per prompt (context length %d)", L, c.tokens);
  MUDG_REQUIRE(heads >= 1 && c.width % heads == 0, "CLIP text: %d heads do not divide width %d", heads, c.width);
  MUDG_REQUIRE(skip_last >= 0 && skip_last < c.layers, "CLIP text: cannot skip %d of %d blocks", skip_last, c.layers);
  run_planned([&](bool plan) { clip_text_body(plan ? nullptr : tokens, B, L, heads, skip_last, plan ? nullptr : out); }, st);
}

}  // namespace mudg
