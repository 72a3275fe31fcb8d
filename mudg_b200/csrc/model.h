// Host-side graph walker: holds the packed weights and sequences the kernels of UNetModel.forward
// (openaimodel3d.py:567-628) and Decoder.forward (ae_modules.py:539-578) on the caller's stream.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"
#include "gemm.h"
#include "mudg.h"
#include "ops.h"

namespace mudg {

struct Weight {          // fp16 [O][taps][Ipad], K-major rows
  __half* w = nullptr;
  int O = 0, I = 0, Ipad = 0, taps = 1;
  int K() const { return taps * Ipad; }
};
struct Vec {             // fp32 [n]
  float* p = nullptr;
  int n = 0;
};

struct Layer {
  std::string kind;      // conv | res | spatial | temporal | down | up
  std::string prefix;
  int cin = 0, cout = 0, ch = 0, heads = 0, inner = 0;
  bool linear_proj = true;
  int res_index = -1;    // index into the per-forward emb_out table
};
struct Block {
  std::vector<Layer> layers;
};

class WeightStore {
 public:
  ~WeightStore();
  void load(const std::string& key, const void* dev_ptr, int dtype, const int64_t* shape, int ndim, cudaStream_t st);
  const Weight& W(const std::string& key) const;
  const Vec& V(const std::string& key) const;
  bool hasW(const std::string& key) const { return w_.count(key) != 0; }
  bool hasV(const std::string& key) const { return v_.count(key) != 0; }
  // out = rows of keys stacked ([sum O][K]); sources are released
  void stack_rows(const std::string& out_key, const std::vector<std::string>& keys, cudaStream_t st);
  void pad_rows(const std::string& wkey, const std::string& bkey, int O_new, cudaStream_t st);
  void make_geglu(const std::string& proj_prefix, cudaStream_t st);
  void fold_ln(const std::string& wkey, const std::string& bkey, const std::string& ln, cudaStream_t st);   // "<p>.weight"/".bias" -> "<p>.geglu.weight"/".bias"
  void drop(const std::string& key);
  void clear();                      // free every tensor (raw and derived)
  bool finalized = false;            // set by Model::finalize; the next load() then starts a fresh set (Model::begin_load)
  size_t bytes() const { return bytes_; }

 private:
  std::unordered_map<std::string, Weight> w_;
  std::unordered_map<std::string, Vec> v_;
  size_t bytes_ = 0;
};

struct KvCache {          // per SpatialTransformer: [N][77][2C] and [Nimg_batches][Limg][2C]
  __half* text = nullptr;
  __half* img = nullptr;
  size_t text_bytes = 0, img_bytes = 0;
  __half* text_vt = nullptr;   // V halves transposed for the tcgen05 kernel: [N][C][80], [Nimg_batches][C][Limg padded to 8]
  __half* img_vt = nullptr;
  size_t text_vt_bytes = 0, img_vt_bytes = 0;
  // per-frame context (L == 77 + 16 T): text + image keys merged into one 96-key block per frame for xattn_per_frame
  __half* kx = nullptr;        // [F][96][C]
  __half* vtx = nullptr;       // [F][C][128]
  size_t kx_bytes = 0, vtx_bytes = 0;
};

class Model {
 public:
  Model(int device, const MudgUNetConfig& u, const MudgVaeConfig& v);
  ~Model();

  WeightStore unet_w, vae_w, res_w;   // res_w: Resampler (image_proj_model), "next" row f.3
  WeightStore clipv_w, clipt_w;       // OpenCLIP image / text towers (clip.cu), "next" row f.3
  WeightStore& store(int which);      // MUDG_UNET / MUDG_VAE / MUDG_RESAMPLER / MUDG_CLIP_IMAGE / MUDG_CLIP_TEXT
  // A load after finalize() replaces the WHOLE weight set: the fused / folded tensors (qkv, kv_text, kv_img, GEGLU
  // interleave, LayerNorm folds, padded rows) are derived from raw tensors that finalize() released, so nothing of the
  // old set may survive.  Also drops the captured CUDA graphs and the cross-attention K/V cache (both hold weight data).
  void begin_load(int which);
  void finalize(int which, cudaStream_t st);
  int device() const { return device_; }
  void set_context(const void* ctx, int dtype, int N, int L, int T, cudaStream_t st);
  // dup > 1: x / t / label / fs hold N / dup distinct samples tiled dup times (sample n == sample n % (N / dup)); only the
  // context differs (classifier-free guidance).  The layers before the first cross-attention then run once per distinct sample.
  void unet_forward(const void* x, const int64_t* t, const int64_t* label, const int64_t* fs, int N, int dup, int T, int h,
                    int w, void* out, cudaStream_t st);
  void vae_decode(const void* z, int F, int h, int w, void* out, cudaStream_t st);
  void vae_encode(const void* x, int F, int H, int W, void* moments, cudaStream_t st);
  // Resampler.forward (resampler.py:131-144): x [B, L, embedding_dim] -> out [B, num_queries*video_length, output_dim] fp32
  void resampler_forward(const void* x, int dtype, int B, int L, void* out, cudaStream_t st);
  // FrozenOpenCLIPImageEmbedderV2.encode_with_vision_transformer (condition.py:339-372): img [B, 3, H, W] -> out
  // [B, tokens, width] fp32.  resize != 0: img is in [-1, 1] at any size and goes through the reference's preprocess
  // (condition.py:318-326); resize == 0: img is already the normalised tower input.
  void clip_image_forward(const void* img, int dtype, int B, int H, int W, int resize, int heads, void* out, cudaStream_t st);
  // FrozenOpenCLIPEmbedder.encode_with_transformer (condition.py:214-232): tokens [B, L] int64 -> out [B, L, width] fp32;
  // skip_last = the reference's layer_idx (1 for layer = "penultimate")
  void clip_text_forward(const int64_t* tokens, int B, int L, int heads, int skip_last, void* out, cudaStream_t st);
  size_t plan_unet(int N, int dup, int T, int h, int w);
  size_t plan_vae(int h, int w);
  int64_t launches = 0;

 private:
  // ---- graph
  void build_plan();
  MudgUNetConfig ucfg_;
  MudgVaeConfig vcfg_;
  int device_;
  std::vector<Block> in_blocks_, out_blocks_;
  Block mid_;
  int n_res_ = 0;
  bool unet_ready_ = false, vae_ready_ = false, vae_enc_ready_ = false, res_ready_ = false;
  struct ResamplerDims { int nq = 0, dim = 0, emb = 0, outd = 0, depth = 0, heads = 0, inner = 0, ff = 0; } rs_;
  void resampler_body(const void* x, int dtype, int B, int L, void* out);
  struct ClipDims { int width = 0, layers = 0, mlp = 0, grid = 0, patch = 0, tokens = 0, vocab = 0; } cv_, ct_;
  bool clipv_ready_ = false, clipt_ready_ = false;
  void finalize_clip(int which);
  Act clip_block(Act x, const std::string& p, int B, int L, int heads, bool causal);
  void clip_image_body(const void* img, int dtype, int B, int H, int W, int resize, int heads, void* out);
  void clip_text_body(const int64_t* tokens, int B, int L, int heads, int skip_last, void* out);
  // body(true) walks the graph allocating nothing (arena high-water mark), then body(false) runs it on `st`
  template <class F>
  void run_planned(F&& body, cudaStream_t st) {
    struct Reset {
      Model& m;
      ~Reset() { m.arena_.planning = false; m.planning_ = false; prof_pause(false); }
    };
    {
      Reset guard{*this};
      arena_.planning = true; planning_ = true;
      arena_.reset_high();
      prof_pause(true);
      body(true);
    }
    const size_t need = arena_.high_water();
    arena_.reset();
    ensure_arena(need);
    st_ = st;
    body(false);
  }

  // ---- per-call state
  Arena arena_;
  cudaStream_t st_ = nullptr;
  bool planning_ = false;
  const WeightStore* ws_ = nullptr;
  int ctx_N_ = 0, ctx_L_ = 0, ctx_T_ = 0, ctx_Limg_ = 0;
  bool ctx_per_frame_ = false;
  std::unordered_map<std::string, KvCache> kv_;
  __half *ctx_text_stage_ = nullptr, *ctx_img_stage_ = nullptr;   // fp16 staging of the context rows (grow-only)
  size_t ctx_text_stage_bytes_ = 0, ctx_img_stage_bytes_ = 0;
  std::vector<float*> emb_out_;   // per ResBlock [N][Cout]
  int T_real_ = 1;                // frames per sample of the current forward
  int N_ = 1;
  struct GraphSlot {
    // Two executable graphs instantiated from ONE capture, launched alternately.  (Measured: cudaGraphLaunch of the
    // 1 126-node forward blocks the host until the previous launch has drained -- 144 ms inside every unet_forward call --
    // and a ring of 8 did not change that, so the launch queue, not the exec object, is the limit; two are kept so that a
    // relaunch never has to wait for its own previous instance.)
    static constexpr int GRAPH_RING = 2;
    cudaGraphExec_t exec = nullptr;          // == ring[0] once captured (non-null <=> captured)
    cudaGraphExec_t ring[GRAPH_RING] = {};
    int next = 0;
    float* in_x = nullptr;
    int64_t* in_idx = nullptr;
    __half* out = nullptr;
    int runs = 0;
    int64_t launches = 0;
    int64_t ctx_version = -1;
  };
  std::map<std::array<int, 5>, GraphSlot> graphs_;
  int64_t ctx_version_ = 0;
  cudaStream_t own_stream_ = nullptr;
  void drop_graphs();
  std::map<std::array<int, 5>, size_t> unet_plans_;
  std::map<std::array<int, 2>, size_t> vae_plans_;

  void ensure_arena(size_t bytes);
  Act alloc(int B, int T, int H, int W, int C);
  void release(Act& a);
  void* alloc_bytes(size_t n);
  void release_bytes(void* p);
  bool live() const { return !planning_; }

  // ---- ops (all no-ops apart from allocation when planning)
  // GroupNorm statistics handed from a producing conv to the norm that consumes its output
  // LayerNorm folded into its consumer GEMM: the consumer takes (mean, rstd) per row.  They come from ln_stats (a full read
  // of the activation), or -- where the GEMM that PRODUCES the activation runs on the pair kernel -- from per-64-column
  // partial sums its epilogue stored (LnReq; TapGemm::ln_out), reduced by ln_finalize.
  struct LnReq {               // handed to the linear() that PRODUCES the LayerNorm's input
    float2* parts = nullptr;   // [nparts][rows], allocated by linear() when its GEMM will store them (pair kernel), else null
    int nparts = 0;
  };
  struct GnReq {
    bool over_time = false;   // statistics per sample over (C/32, T, H, W) instead of per frame
    double* sums = nullptr;   // [S][32][2] fp64 in the arena (allocated + zeroed by gn_request, released by group_norm)
    bool fused = false;       // the producer's epilogue accumulated them
  };
  void gn_request(GnReq& r, int S);
  Act group_norm(const Act& x, const std::string& p, float eps, bool silu, bool over_time, GnReq* pre = nullptr);
  Act linear_gn(const Act& x, const std::string& norm, float eps, bool over_time, const std::string& wkey,
                const std::string& bkey, GnReq* pre, LnReq* ln_out = nullptr);
  // planning walks use fake pointers: the fold decision must not depend on pointer alignment
  bool fold_plan_ok(TapGemm g) { g.A = g.D = nullptr; g.Wt = nullptr; return tapgemm_per_sample_ok(g); }
  Act layer_norm(const Act& x, const std::string& p);
  bool ln_plan_ok(TapGemm g) { g.A = g.R = nullptr; g.D = nullptr; g.Wt = nullptr; return tapgemm_ln_out_ok(g); }
  Act linear(const Act& x, const std::string& wkey, const std::string& bkey, const Act* residual, bool geglu = false,
             float alpha = 1.f, const float2* ln = nullptr, LnReq* ln_out = nullptr);
  float2* layer_norm_stats(const Act& x, LnReq* pre = nullptr);
  Act conv3x3(const Act& x, const std::string& p, const Act* residual, const float* bias2, GnReq* gn = nullptr);
  Act conv_t3(const Act& x, const std::string& p, const Act* residual, GnReq* gn = nullptr);
  Act gemm_raw(const __half* A, int M, int K, const __half* Wt, int N, const float* bias, const Act* residual, float alpha);
  Act concat(const Act& a, const Act& b, GnReq* gn = nullptr);
  Act tile_batch(const Act& x, int n);
  Act upsample(const Act& x);
  Act downsample(const Act& x, const std::string& p, int pad);

  Act res_block(const Act& x, const Layer& l, GnReq* out_gn = nullptr, GnReq* in_gn = nullptr);
  Act transformer_block_tail(Act x, const std::string& p, LnReq* pre = nullptr);   // LN3 + GEGLU FF + residual (consumes x)
  Act spatial_transformer(const Act& x, const Layer& l, GnReq* in_gn = nullptr);
  Act temporal_transformer(const Act& x, const Layer& l);
  Act run_block(Act h, const Block& b, bool owns_input, GnReq* in_gn = nullptr);
  void compute_embeddings(const int64_t* t, const int64_t* label, const int64_t* fs, int N);
  void unet_body(const void* x, const int64_t* t, const int64_t* label, const int64_t* fs, int N, int dup, int T, int h, int w,
                 void* out);
  Act vae_res(const Act& x, const std::string& p);
  Act vae_attn(const Act& x, const std::string& p);
  void vae_body(const void* z, int h, int w, void* out);
  void vae_encode_body(const void* x, int H, int W, void* moments);
};

}  // namespace mudg
