#include "model.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <ctime>

namespace mudg {

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ================================================================ WeightStore
WeightStore::~WeightStore() {
  for (auto& kv : w_) cudaFree(kv.second.w);
  for (auto& kv : v_) cudaFree(kv.second.p);
}

void WeightStore::load(const std::string& key, const void* dev_ptr, int dtype, const int64_t* shape, int ndim,
                       cudaStream_t st) {
  MUDG_REQUIRE(ndim >= 1 && ndim <= 5, "weight %s: ndim %d", key.c_str(), ndim);
  MUDG_REQUIRE(dtype == MUDG_F32 || dtype == MUDG_F16, "weight %s: dtype %d", key.c_str(), dtype);
  drop(key);
  if (ndim == 1) {
    Vec v;
    v.n = (int)shape[0];
    MUDG_CUDA(cudaMalloc(&v.p, sizeof(float) * v.n));
    cast_to_f32(dev_ptr, dtype == MUDG_F32, v.p, v.n, st);
    v_[key] = v;
    bytes_ += sizeof(float) * v.n;
    return;
  }
  Weight w;
  w.O = (int)shape[0];
  w.I = (int)shape[1];
  w.taps = 1;
  for (int i = 2; i < ndim; i++) w.taps *= (int)shape[i];
  w.Ipad = round_up(w.I, 8);
  const size_t n = (size_t)w.O * w.taps * w.Ipad;
  MUDG_CUDA(cudaMalloc(&w.w, n * sizeof(__half)));
  pack_weight(dev_ptr, dtype == MUDG_F32, w.w, w.O, w.I, w.taps, w.Ipad, st);
  w_[key] = w;
  bytes_ += n * sizeof(__half);
}

const Weight& WeightStore::W(const std::string& key) const {
  auto it = w_.find(key);
  if (it == w_.end()) throw Error("missing weight: " + key);
  return it->second;
}
const Vec& WeightStore::V(const std::string& key) const {
  auto it = v_.find(key);
  if (it == v_.end()) throw Error("missing weight: " + key);
  return it->second;
}
void WeightStore::drop(const std::string& key) {
  auto a = w_.find(key);
  if (a != w_.end()) {
    bytes_ -= (size_t)a->second.O * a->second.K() * sizeof(__half);
    cudaFree(a->second.w);
    w_.erase(a);
  }
  auto b = v_.find(key);
  if (b != v_.end()) {
    bytes_ -= sizeof(float) * b->second.n;
    cudaFree(b->second.p);
    v_.erase(b);
  }
}

void WeightStore::clear() {
  for (auto& kv : w_) cudaFree(kv.second.w);
  for (auto& kv : v_) cudaFree(kv.second.p);
  w_.clear();
  v_.clear();
  bytes_ = 0;
  finalized = false;
}

void WeightStore::stack_rows(const std::string& out_key, const std::vector<std::string>& keys, cudaStream_t st) {
  if (hasW(out_key)) return;   // already fused
  Weight o;
  const Weight& f = W(keys[0]);
  o.I = f.I; o.Ipad = f.Ipad; o.taps = f.taps;
  for (auto& k : keys) {
    const Weight& w = W(k);
    MUDG_REQUIRE(w.K() == f.K(), "stack_rows: K mismatch for %s", k.c_str());
    o.O += w.O;
  }
  MUDG_CUDA(cudaMalloc(&o.w, (size_t)o.O * o.K() * sizeof(__half)));
  size_t off = 0;
  for (auto& k : keys) {
    const Weight& w = W(k);
    const size_t n = (size_t)w.O * w.K();
    MUDG_CUDA(cudaMemcpyAsync(o.w + off, w.w, n * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    off += n;
  }
  MUDG_CUDA(cudaStreamSynchronize(st));   // weight-load time only; the sources are freed right below
  for (auto& k : keys) drop(k);
  w_[out_key] = o;
  bytes_ += off * sizeof(__half);
}

// zero-padded copy with O_new rows: lets a tiny-N layer (N = 4 / 3) run on the tensor-core path
void WeightStore::pad_rows(const std::string& wkey, const std::string& bkey, int O_new, cudaStream_t st) {
  if (hasW(wkey + ".pad")) return;
  const Weight& w = W(wkey);
  const Vec& b = V(bkey);
  MUDG_REQUIRE(O_new >= w.O, "pad_rows");
  Weight o = w;
  o.O = O_new;
  Vec ob;
  ob.n = O_new;
  const size_t nb = (size_t)O_new * w.K() * sizeof(__half);
  MUDG_CUDA(cudaMalloc(&o.w, nb));
  MUDG_CUDA(cudaMalloc(&ob.p, sizeof(float) * O_new));
  MUDG_CUDA(cudaMemsetAsync(o.w, 0, nb, st));
  MUDG_CUDA(cudaMemsetAsync(ob.p, 0, sizeof(float) * O_new, st));
  MUDG_CUDA(cudaMemcpyAsync(o.w, w.w, (size_t)w.O * w.K() * sizeof(__half), cudaMemcpyDeviceToDevice, st));
  MUDG_CUDA(cudaMemcpyAsync(ob.p, b.p, sizeof(float) * b.n, cudaMemcpyDeviceToDevice, st));
  w_[wkey + ".pad"] = o;
  v_[bkey + ".pad"] = ob;
  bytes_ += nb + sizeof(float) * O_new;
}

// Fold LayerNorm `ln` (gamma, beta) into Linear `wkey` (+ optional bias `bkey`): the weight is scaled in place and two
// vectors are added: "<wkey>.ln_c1" (row sums) and "<wkey>.ln_c2" (W beta + bias), consumed by the GEMM epilogue.
void WeightStore::fold_ln(const std::string& wkey, const std::string& bkey, const std::string& ln, cudaStream_t st) {
  if (hasV(wkey + ".ln_c1")) return;
  const Weight& w = W(wkey);
  MUDG_REQUIRE(w.taps == 1 && w.Ipad == w.I, "fold_ln: %s is not a plain Linear", wkey.c_str());
  const Vec& g = V(ln + ".weight");
  const Vec& b = V(ln + ".bias");
  MUDG_REQUIRE(g.n == w.I, "fold_ln: LayerNorm width %d vs Linear K %d", g.n, w.I);
  Vec c1, c2;
  c1.n = c2.n = w.O;
  MUDG_CUDA(cudaMalloc(&c1.p, sizeof(float) * w.O));
  MUDG_CUDA(cudaMalloc(&c2.p, sizeof(float) * w.O));
  ln_fold(w.w, g.p, b.p, bkey.empty() ? nullptr : V(bkey).p, c1.p, c2.p, w.O, w.I, st);
  v_[wkey + ".ln_c1"] = c1;
  v_[wkey + ".ln_c2"] = c2;
  bytes_ += 2 * sizeof(float) * w.O;
}

void WeightStore::make_geglu(const std::string& p, cudaStream_t st) {
  if (hasW(p + ".geglu.weight")) return;
  const Weight& w = W(p + ".weight");
  const Vec& b = V(p + ".bias");
  MUDG_REQUIRE(w.O % 128 == 0 && w.taps == 1, "GEGLU proj rows %d", w.O);
  Weight o = w;
  Vec ob;
  ob.n = b.n;
  MUDG_CUDA(cudaMalloc(&o.w, (size_t)w.O * w.K() * sizeof(__half)));
  MUDG_CUDA(cudaMalloc(&ob.p, sizeof(float) * b.n));
  geglu_interleave(w.w, b.p, o.w, ob.p, w.O / 2, w.K(), st);
  MUDG_CUDA(cudaStreamSynchronize(st));
  drop(p + ".weight");
  drop(p + ".bias");
  w_[p + ".geglu.weight"] = o;
  v_[p + ".geglu.bias"] = ob;
  bytes_ += (size_t)o.O * o.K() * sizeof(__half) + sizeof(float) * ob.n;
}

// ================================================================ Model: construction / plan
Model::Model(int device, const MudgUNetConfig& u, const MudgVaeConfig& v) : ucfg_(u), vcfg_(v), device_(device) {
  MUDG_REQUIRE(u.num_head_channels == 64, "kernels are specialised for num_head_channels == 64 (got %d)",
               u.num_head_channels);
  MUDG_REQUIRE(u.model_channels % 32 == 0, "model_channels must be a multiple of 32");
  build_plan();
}
Model::~Model() {
  drop_graphs();
  if (own_stream_) cudaStreamDestroy(own_stream_);
  for (auto& kv : graphs_) {
    cudaFree(kv.second.in_x);
    cudaFree(kv.second.in_idx);
    cudaFree(kv.second.out);
  }
  for (auto& kv : kv_) {
    cudaFree(kv.second.text);
    cudaFree(kv.second.img);
    cudaFree(kv.second.text_vt);
    cudaFree(kv.second.img_vt);
    cudaFree(kv.second.kx);
    cudaFree(kv.second.vtx);
  }
  cudaFree(ctx_text_stage_);
  cudaFree(ctx_img_stage_);
}

// mirrors UNetModel.__init__ (openaimodel3d.py:398-565)
void Model::build_plan() {
  const MudgUNetConfig& c = ucfg_;
  const int mc = c.model_channels;
  auto has_attn = [&](int ds) {
    for (int i = 0; i < c.n_attention_resolutions; i++)
      if (c.attention_resolutions[i] == ds) return true;
    return false;
  };
  auto pfx = [](const char* root, int a, int b) { return std::string(root) + "." + std::to_string(a) + "." + std::to_string(b); };
  int res_count = 0;
  auto res = [&](const std::string& p, int cin, int cout) {
    Layer l; l.kind = "res"; l.prefix = p; l.cin = cin; l.cout = cout; l.res_index = res_count++;
    return l;
  };
  auto spatial = [&](const std::string& p, int ch) {
    Layer l; l.kind = "spatial"; l.prefix = p; l.ch = ch; l.heads = ch / 64; l.inner = ch;
    return l;
  };
  auto temporal = [&](const std::string& p, int ch) {
    Layer l; l.kind = "temporal"; l.prefix = p; l.ch = ch; l.heads = ch / 64; l.inner = ch; l.linear_proj = true;
    return l;
  };
  in_blocks_.clear(); out_blocks_.clear();
  {
    Block b; Layer l; l.kind = "conv"; l.prefix = "input_blocks.0.0"; l.cin = c.in_channels; l.cout = mc;
    b.layers.push_back(l); in_blocks_.push_back(b);
  }
  std::vector<int> chans{mc};
  int ch = mc, ds = 1, idx = 1;
  for (int level = 0; level < c.n_channel_mult; level++) {
    const int mult = c.channel_mult[level];
    for (int r = 0; r < c.num_res_blocks; r++) {
      Block b;
      b.layers.push_back(res(pfx("input_blocks", idx, 0), ch, mult * mc));
      ch = mult * mc;
      if (has_attn(ds)) {
        b.layers.push_back(spatial(pfx("input_blocks", idx, 1), ch));
        b.layers.push_back(temporal(pfx("input_blocks", idx, 2), ch));
      }
      in_blocks_.push_back(b); chans.push_back(ch); idx++;
    }
    if (level != c.n_channel_mult - 1) {
      Block b; Layer l; l.kind = "down"; l.prefix = pfx("input_blocks", idx, 0); l.ch = ch;
      b.layers.push_back(l); in_blocks_.push_back(b); chans.push_back(ch); idx++; ds *= 2;
    }
  }
  mid_.layers.clear();
  mid_.layers.push_back(res("middle_block.0", ch, ch));
  mid_.layers.push_back(spatial("middle_block.1", ch));
  mid_.layers.push_back(temporal("middle_block.2", ch));
  mid_.layers.push_back(res("middle_block.3", ch, ch));
  int oidx = 0;
  for (int level = c.n_channel_mult - 1; level >= 0; level--) {
    const int mult = c.channel_mult[level];
    for (int i = 0; i <= c.num_res_blocks; i++) {
      const int ich = chans.back(); chans.pop_back();
      Block b; int li = 0;
      b.layers.push_back(res(pfx("output_blocks", oidx, li++), ch + ich, mc * mult));
      ch = mc * mult;
      if (has_attn(ds)) {
        b.layers.push_back(spatial(pfx("output_blocks", oidx, li++), ch));
        b.layers.push_back(temporal(pfx("output_blocks", oidx, li++), ch));
      }
      if (level && i == c.num_res_blocks) {
        Layer l; l.kind = "up"; l.prefix = pfx("output_blocks", oidx, li++); l.ch = ch;
        b.layers.push_back(l); ds /= 2;
      }
      out_blocks_.push_back(b); oidx++;
    }
  }
  n_res_ = res_count;
}

WeightStore& Model::store(int which) {
  switch (which) {
    case MUDG_UNET: return unet_w;
    case MUDG_VAE: return vae_w;
    case MUDG_RESAMPLER: return res_w;
    case MUDG_CLIP_IMAGE: return clipv_w;
    case MUDG_CLIP_TEXT: return clipt_w;
  }
  throw Error(fmt("unknown weight set %d", which));
}

void Model::begin_load(int which) {
  WeightStore& ws = store(which);
  if (!ws.finalized) return;
  MUDG_CUDA(cudaDeviceSynchronize());      // kernels of earlier forwards may still read the old set
  ws.clear();
  clear_tmap_cache();                      // tensor maps are keyed by weight pointers
  if (which == MUDG_UNET) {
    unet_ready_ = false;
    drop_graphs();                         // captured kernel arguments / tensor maps point at the freed weights
    ctx_N_ = ctx_L_ = ctx_T_ = 0;          // the K/V cache was projected with the old weights: set_context must run again
    ctx_version_++;
  } else if (which == MUDG_VAE) {
    vae_ready_ = vae_enc_ready_ = false;
  } else if (which == MUDG_RESAMPLER) {
    res_ready_ = false;
  } else if (which == MUDG_CLIP_IMAGE) {
    clipv_ready_ = false;
  } else {
    clipt_ready_ = false;
  }
}

void Model::finalize(int which, cudaStream_t st) {
  store(which).finalized = true;
  if (which == MUDG_CLIP_IMAGE || which == MUDG_CLIP_TEXT) {
    finalize_clip(which);
    return;
  }
  if (which == MUDG_RESAMPLER) {
    // touch every key Resampler.forward needs and derive the dimensions from the shapes (resampler.py:104-129)
    WeightStore& w = res_w;
    rs_ = ResamplerDims{};
    rs_.nq = w.W("latents").O; rs_.dim = w.W("latents").I;
    rs_.emb = w.W("proj_in.weight").I; rs_.outd = w.W("proj_out.weight").O;
    MUDG_REQUIRE(w.W("proj_in.weight").O == rs_.dim && w.W("proj_out.weight").I == rs_.dim, "Resampler: proj_in / proj_out width");
    w.V("proj_in.bias"); w.V("proj_out.bias"); w.V("norm_out.weight"); w.V("norm_out.bias");
    while (w.hasW("layers." + std::to_string(rs_.depth) + ".0.to_q.weight")) {
      const std::string a = "layers." + std::to_string(rs_.depth) + ".0", f = "layers." + std::to_string(rs_.depth) + ".1";
      const int inner = w.W(a + ".to_q.weight").O;
      MUDG_REQUIRE(inner % 64 == 0 && w.W(a + ".to_kv.weight").O == 2 * inner && w.W(a + ".to_out.weight").I == inner,
                   "Resampler layer %d: dim_head must be 64 (inner %d)", rs_.depth, inner);
      MUDG_REQUIRE(rs_.depth == 0 || inner == rs_.inner, "Resampler: layers differ");
      rs_.inner = inner; rs_.heads = inner / 64;
      rs_.ff = w.W(f + ".1.weight").O;
      MUDG_REQUIRE(w.W(f + ".3.weight").I == rs_.ff, "Resampler FF width");
      for (const char* n : {".norm1", ".norm2"}) { w.V(a + n + ".weight"); w.V(a + n + ".bias"); }
      w.V(f + ".0.weight"); w.V(f + ".0.bias");
      rs_.depth++;
    }
    MUDG_REQUIRE(rs_.depth > 0, "Resampler: no layers loaded");
    res_ready_ = true;
    return;
  }
  if (which == MUDG_VAE) {
    // touch every key Decoder.forward needs (throws on a missing one)
    const MudgVaeConfig& v = vcfg_;
    auto need_res = [&](const std::string& p, int ci, int co) {
      vae_w.V(p + ".norm1.weight"); vae_w.W(p + ".conv1.weight"); vae_w.V(p + ".norm2.bias"); vae_w.W(p + ".conv2.weight");
      if (ci != co) vae_w.W(p + ".nin_shortcut.weight");
    };
    int block_in = v.ch * v.ch_mult[v.n_ch_mult - 1];
    vae_w.W("post_quant_conv.weight"); vae_w.W("decoder.conv_in.weight");
    need_res("decoder.mid.block_1", block_in, block_in);
    for (const char* n : {"q", "k", "v", "proj_out"}) vae_w.W(std::string("decoder.mid.attn_1.") + n + ".weight");
    need_res("decoder.mid.block_2", block_in, block_in);
    for (int lvl = v.n_ch_mult - 1; lvl >= 0; lvl--) {
      const int block_out = v.ch * v.ch_mult[lvl];
      for (int ib = 0; ib <= v.num_res_blocks; ib++) {
        need_res("decoder.up." + std::to_string(lvl) + ".block." + std::to_string(ib), block_in, block_out);
        block_in = block_out;
      }
      if (lvl != 0) vae_w.W("decoder.up." + std::to_string(lvl) + ".upsample.conv.weight");
    }
    vae_w.V("decoder.norm_out.weight"); vae_w.W("decoder.conv_out.weight");
    vae_w.pad_rows("decoder.conv_out.weight", "decoder.conv_out.bias", 64, st);
    vae_ready_ = true;
    if (vae_w.hasW("encoder.conv_in.weight")) {       // the encoder is optional (decode-only deployments)
      int bi = v.ch;
      vae_w.W("encoder.conv_in.weight");
      for (int lvl = 0; lvl < v.n_ch_mult; lvl++) {
        const int bo = v.ch * v.ch_mult[lvl];
        for (int ib = 0; ib < v.num_res_blocks; ib++) {
          need_res("encoder.down." + std::to_string(lvl) + ".block." + std::to_string(ib), bi, bo);
          bi = bo;
        }
        if (lvl != v.n_ch_mult - 1) vae_w.W("encoder.down." + std::to_string(lvl) + ".downsample.conv.weight");
      }
      need_res("encoder.mid.block_1", bi, bi);
      for (const char* n : {"q", "k", "v", "proj_out"}) vae_w.W(std::string("encoder.mid.attn_1.") + n + ".weight");
      need_res("encoder.mid.block_2", bi, bi);
      vae_w.V("encoder.norm_out.weight"); vae_w.W("encoder.conv_out.weight"); vae_w.W("quant_conv.weight");
      vae_w.pad_rows("encoder.conv_out.weight", "encoder.conv_out.bias", 64, st);
      vae_enc_ready_ = true;
    }
    return;
  }
  WeightStore& w = unet_w;
  auto fuse_block = [&](const std::string& tb, bool cross) {
    w.stack_rows(tb + ".attn1.qkv.weight", {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, st);
    if (cross) {
      w.stack_rows(tb + ".attn2.kv_text.weight", {tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, st);
      w.stack_rows(tb + ".attn2.kv_img.weight", {tb + ".attn2.to_k_ip.weight", tb + ".attn2.to_v_ip.weight"}, st);
      w.W(tb + ".attn2.to_q.weight");
    } else {
      w.stack_rows(tb + ".attn2.qkv.weight", {tb + ".attn2.to_q.weight", tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, st);
    }
    w.make_geglu(tb + ".ff.net.0.proj", st);
    // LayerNorms are folded into the Linear they feed (raw activations go straight into the GEMM)
    w.fold_ln(tb + ".attn1.qkv.weight", "", tb + ".norm1", st);
    w.fold_ln(tb + (cross ? ".attn2.to_q.weight" : ".attn2.qkv.weight"), "", tb + ".norm2", st);
    w.fold_ln(tb + ".ff.net.0.proj.geglu.weight", tb + ".ff.net.0.proj.geglu.bias", tb + ".norm3", st);
    for (const char* n : {".attn1.to_out.0", ".attn2.to_out.0", ".ff.net.2"}) { w.W(tb + n + ".weight"); w.V(tb + n + ".bias"); }
    for (const char* n : {".norm1", ".norm2", ".norm3"}) { w.V(tb + n + ".weight"); w.V(tb + n + ".bias"); }
  };
  auto visit = [&](const Layer& l) {
    if (l.kind == "spatial" || l.kind == "temporal") {
      fuse_block(l.prefix + ".transformer_blocks.0", l.kind == "spatial");
      w.V(l.prefix + ".norm.weight"); w.W(l.prefix + ".proj_in.weight"); w.W(l.prefix + ".proj_out.weight");
    } else if (l.kind == "res") {
      w.V(l.prefix + ".in_layers.0.weight"); w.W(l.prefix + ".in_layers.2.weight"); w.W(l.prefix + ".emb_layers.1.weight");
      w.V(l.prefix + ".out_layers.0.weight"); w.W(l.prefix + ".out_layers.3.weight");
      if (l.cin != l.cout) w.W(l.prefix + ".skip_connection.weight");
      for (int j = 1; j <= 4; j++) {
        const std::string q = l.prefix + ".temopral_conv.conv" + std::to_string(j);
        w.V(q + ".0.weight"); w.W(q + (j == 1 ? ".2" : ".3") + ".weight");
      }
    } else if (l.kind == "down") w.W(l.prefix + ".op.weight");
    else if (l.kind == "up") w.W(l.prefix + ".conv.weight");
    else if (l.kind == "conv") w.W(l.prefix + ".weight");
  };
  for (auto& b : in_blocks_) for (auto& l : b.layers) visit(l);
  for (auto& l : mid_.layers) visit(l);
  for (auto& b : out_blocks_) for (auto& l : b.layers) visit(l);
  Layer ia; ia.kind = "temporal"; ia.prefix = "init_attn.0";
  visit(ia);
  for (const char* n : {"time_embed", "class_embed", "fps_embedding"})
    for (const char* k : {".0", ".2"}) { w.W(std::string(n) + k + ".weight"); w.V(std::string(n) + k + ".bias"); }
  w.V("out.0.weight"); w.W("out.2.weight");
  w.pad_rows("out.2.weight", "out.2.bias", 64, st);
  unet_ready_ = true;
}

// ================================================================ allocation helpers
void Model::ensure_arena(size_t bytes) {
  if (arena_.capacity() < bytes) {
    MUDG_CUDA(cudaDeviceSynchronize());
    drop_graphs();                     // the slab moves: pointers captured in CUDA graphs are stale
    arena_.reserve(bytes + (bytes >> 4));
  }
}
Act Model::alloc(int B, int T, int H, int W, int C) {
  Act a; a.B = B; a.T = T; a.H = H; a.W = W; a.C = C;
  a.p = static_cast<__half*>(arena_.alloc(a.bytes()));
  return a;
}
void Model::release(Act& a) {
  if (a.p) arena_.free(a.p);
  a.p = nullptr;
}
void* Model::alloc_bytes(size_t n) { return arena_.alloc(n); }
void Model::release_bytes(void* p) { arena_.free(p); }

// ================================================================ ops
// Statistics buffer of a GroupNorm whose input is about to be produced: [S][32][2] fp64, zeroed on the stream.
void Model::gn_request(GnReq& r, int S) {
  r.sums = static_cast<double*>(alloc_bytes(sizeof(double) * S * 64));
  r.fused = false;
  if (live()) {
    MUDG_CUDA(cudaMemsetAsync(r.sums, 0, sizeof(double) * S * 64, st_));
    launches++;
  }
}

// GroupNorm (+ SiLU).  `pre`: statistics requested from the op that produced x (GnReq filled by conv3x3 / conv_t3); when
// the producer's epilogue accumulated them (pre->fused) the norm is ONE pass over x, otherwise a statistics pass runs first.
Act Model::group_norm(const Act& x, const std::string& p, float eps, bool silu, bool over_time, GnReq* pre) {
  const int S = over_time ? x.B : x.B * x.T;
  const int64_t rps = over_time ? (int64_t)x.T * x.H * x.W : (int64_t)x.H * x.W;
  Act y = alloc(x.B, x.T, x.H, x.W, x.C);
  GnReq own;
  if (pre == nullptr || pre->sums == nullptr) {
    gn_request(own, S);
    pre = &own;
  } else {
    MUDG_REQUIRE(pre->over_time == over_time, "GroupNorm %s: statistics were requested for the other sample partition", p.c_str());
  }
  if (live()) {
    const Vec& g = ws_->V(p + ".weight");
    const Vec& b = ws_->V(p + ".bias");
    MUDG_REQUIRE(g.n == x.C, "GroupNorm %s: %d channels vs activation %d", p.c_str(), g.n, x.C);
    // ideal traffic: 1 read + 1 write of the activation
    ProfScope ps(PF_GN, 0.0, 4.0 * (double)x.numel(), st_, fmt("%lldx%d%s", (long long)x.rows(), x.C, pre->fused ? ":fused" : "").c_str());
    if (!pre->fused && knobs().gn_small != 0 && gn_small_ok(S, rps, x.C)) {
      gn_small(x.p, y.p, S, rps, x.C, g.p, b.p, eps, silu, st_);          // statistics + apply in one kernel
    } else {
      if (!pre->fused) {
        gn_stats(x.p, S, rps, x.C, pre->sums, st_);
        launches++;
      }
      gn_apply(x.p, y.p, pre->sums, S, rps, x.C, g.p, b.p, eps, silu, st_);
    }
    launches++;
  }
  release_bytes(pre->sums);
  pre->sums = nullptr;
  return y;
}

// proj_in(GroupNorm(x)) of the Spatial / Temporal transformers (attention.py:454-456, 532-540).  The norm has no activation,
// so it is affine per (sample, channel) and folds into per-sample weights: the Linear reads the RAW activation and the
// normalised tensor is never written (saves one read + one write of the activation per transformer).  Used when the GEMM runs
// on the pair kernel with tiles inside one sample and the per-sample weights are small next to the activation; otherwise
// the norm runs as its own pass.
Act Model::linear_gn(const Act& x, const std::string& norm, float eps, bool over_time, const std::string& wkey,
                     const std::string& bkey, GnReq* pre, LnReq* ln_out) {
  const Weight& w = ws_->W(wkey);
  const int S = over_time ? x.B : x.B * x.T;
  const int64_t rps = over_time ? (int64_t)x.T * x.H * x.W : (int64_t)x.H * x.W;
  TapGemm g;
  g.A = x.p; g.B = x.B; g.T = x.T; g.H = x.H; g.W = x.W; g.Cin = x.C;
  g.ntaps = 1;
  g.Wt = w.w; g.N = w.O;                      // (pointers are placeholders until the buffers below exist: geometry check only)
  g.D = x.p;
  g.wt_samples = S; g.wt_div = over_time ? x.T : 1;
  const bool small_weights = (int64_t)S * w.O * 2 <= x.rows();      // per-sample weights <= 1/2 of the activation bytes
  const bool fold = knobs().gn_fold != 0 && w.taps == 1 && w.K() == x.C && x.C % 32 == 0 && small_weights &&
                    (planning_ ? fold_plan_ok(g) : tapgemm_per_sample_ok(g));
  if (!fold) {
    Act n = group_norm(x, norm, eps, false, over_time, pre);
    Act y = linear(n, wkey, bkey, nullptr, false, 1.f, nullptr, ln_out);
    release(n);
    return y;
  }
  GnReq own;
  if (pre == nullptr || pre->sums == nullptr) {
    gn_request(own, S);
    pre = &own;
  }
  Act y = alloc(x.B, x.T, x.H, x.W, w.O);
  if (ln_out && ln_plan_ok(g)) {             // LayerNorm partials of the output for the transformer block's norm1
    ln_out->nparts = w.O / 64;
    ln_out->parts = static_cast<float2*>(alloc_bytes(sizeof(float2) * (size_t)ln_out->nparts * x.rows()));
    g.ln_out = ln_out->parts;
  }
  __half* Ws = static_cast<__half*>(alloc_bytes(sizeof(__half) * (size_t)S * w.O * w.K()));
  float* cs = static_cast<float*>(alloc_bytes(sizeof(float) * (size_t)S * w.O));
  if (live()) {
    if (!pre->fused) {
      ProfScope ps(PF_GN, 0.0, 2.0 * (double)x.numel(), st_, fmt("%lldx%d:stats", (long long)x.rows(), x.C).c_str());
      gn_stats(x.p, S, rps, x.C, pre->sums, st_);
      launches++;
    }
    {
      ProfScope ps(PF_GN, 0.0, 0.0, st_, "fold_weights");
      gn_fold_weights(w.w, pre->sums, S, rps, ws_->V(norm + ".weight").p, ws_->V(norm + ".bias").p, eps, Ws, cs, w.O, w.K(), st_);
      launches++;
    }
    g.Wt = Ws; g.D = y.p;
    g.bias = bkey.empty() ? nullptr : ws_->V(bkey).p;
    g.bias2 = cs; g.bias2_div = g.wt_div; g.nb2 = S;
    tapgemm(g, st_);
    launches++;
  }
  release_bytes(Ws);
  release_bytes(cs);
  release_bytes(pre->sums);
  pre->sums = nullptr;
  return y;
}

Act Model::layer_norm(const Act& x, const std::string& p) {
  Act y = alloc(x.B, x.T, x.H, x.W, x.C);
  if (live()) {
    ProfScope ps(PF_LAYERNORM, 0.0, 4.0 * (double)x.numel(), st_, fmt("%lldx%d", (long long)x.rows(), x.C).c_str());
    layernorm(x.p, y.p, ws_->V(p + ".weight").p, ws_->V(p + ".bias").p, x.rows(), x.C, 1e-5f, st_);
    launches++;
  }
  return y;
}

// (mean, rstd) per row for a LayerNorm that has been folded into its consumer GEMM
float2* Model::layer_norm_stats(const Act& x, LnReq* pre) {
  float2* st = static_cast<float2*>(alloc_bytes(sizeof(float2) * (size_t)x.rows()));
  if (pre && pre->parts) {                   // the producer's epilogue has stored per-chunk partial sums: reduce those
    MUDG_REQUIRE(pre->nparts * 64 == x.C, "LayerNorm partials: %d chunks for width %d", pre->nparts, x.C);
    if (live()) {
      ProfScope ps(PF_LN_STATS, 0.0, 8.0 * (double)x.rows() * (pre->nparts + 1), st_, fmt("%lldx%d:fused", (long long)x.rows(), x.C).c_str());
      ln_finalize(pre->parts, pre->nparts, st, x.rows(), x.C, 1e-5f, st_);
      launches++;
    }
    release_bytes(pre->parts);
    pre->parts = nullptr;
    return st;
  }
  if (live()) {
    ProfScope ps(PF_LN_STATS, 0.0, 2.0 * (double)x.numel() + 8.0 * (double)x.rows(), st_, fmt("%lldx%d", (long long)x.rows(), x.C).c_str());
    ln_stats(x.p, st, x.rows(), x.C, 1e-5f, st_);
    launches++;
  }
  return st;
}

Act Model::linear(const Act& x, const std::string& wkey, const std::string& bkey, const Act* residual, bool geglu,
                  float alpha, const float2* ln, LnReq* ln_out) {
  const Weight& w0 = ws_->W(wkey);
  MUDG_REQUIRE(w0.K() == x.C, "linear %s: K %d vs activation width %d", wkey.c_str(), w0.K(), x.C);
  const int n_out = geglu ? w0.O / 2 : w0.O;
  Act y = alloc(x.B, x.T, x.H, x.W, n_out);
  MUDG_REQUIRE(x.rows() < (int64_t(1) << 31), "too many rows");
  TapGemm g;
  g.B = 1; g.T = 1; g.H = 1; g.W = (int)x.rows(); g.Cin = x.C;
  g.ntaps = 1;
  g.N = w0.O;
  g.alpha = alpha; g.geglu = geglu;
  if (ln_out && !geglu && n_out % 64 == 0 && ln_plan_ok(g)) {   // same decision in the planning walk and the live one
    ln_out->nparts = n_out / 64;
    ln_out->parts = static_cast<float2*>(alloc_bytes(sizeof(float2) * (size_t)ln_out->nparts * x.rows()));
    g.ln_out = ln_out->parts;
  }
  if (live()) {
    const Weight& w = ws_->W(wkey);
    g.A = x.p;
    g.Wt = w.w;
    g.D = y.p;
    g.R = residual ? residual->p : nullptr;
    g.bias = bkey.empty() ? nullptr : ws_->V(bkey).p;
    if (ln) {                                  // folded LayerNorm: the weight already carries gamma
      g.ln_stats = ln;
      g.ln_c1 = ws_->V(wkey + ".ln_c1").p;
      g.bias = ws_->V(wkey + ".ln_c2").p;
    }
    if (residual) MUDG_REQUIRE(residual->C == n_out && residual->rows() == x.rows(), "linear %s: residual shape", wkey.c_str());
    tapgemm(g, st_);
    launches++;
  }
  return y;
}

Act Model::conv3x3(const Act& x, const std::string& p, const Act* residual, const float* bias2, GnReq* gn) {
  const Weight& w = ws_->W(p + ".weight");
  Act y = alloc(x.B, x.T, x.H, x.W, w.O);
  if (gn) gn_request(*gn, gn->over_time ? x.B : x.B * x.T);
  if (live()) {
    MUDG_REQUIRE(w.taps == 9 && w.Ipad == x.C, "conv3x3 %s: weight [%d][%d][%d] vs activation C=%d", p.c_str(), w.O, w.taps,
                 w.Ipad, x.C);
    TapGemm g;
    g.A = x.p; g.B = 1; g.T = x.B * x.T; g.H = x.H; g.W = x.W; g.Cin = x.C;   // frames are independent
    g.ntaps = 9; set_taps_3x3(g.taps);
    g.Wt = w.w; g.N = w.O; g.D = y.p;
    g.R = residual ? residual->p : nullptr;
    g.bias = ws_->V(p + ".bias").p;
    if (bias2) { g.bias2 = bias2; g.bias2_div = T_real_; g.nb2 = N_; }
    if (gn) { g.gn_sums = gn->sums; g.gn_div = gn->over_time ? x.T : 1; }      // frames are flattened into the T dimension
    const bool fused = tapgemm(g, st_);
    if (gn) gn->fused = fused;
    launches++;
  }
  return y;
}

Act Model::conv_t3(const Act& x, const std::string& p, const Act* residual, GnReq* gn) {
  const Weight& w = ws_->W(p + ".weight");
  Act y = alloc(x.B, x.T, x.H, x.W, w.O);
  if (gn) gn_request(*gn, gn->over_time ? x.B : x.B * x.T);
  if (live()) {
    MUDG_REQUIRE(w.taps == 3 && w.Ipad == x.C, "temporal conv %s: weight shape", p.c_str());
    TapGemm g;
    g.A = x.p; g.B = x.B; g.T = x.T; g.H = x.H; g.W = x.W; g.Cin = x.C;
    g.ntaps = 3; set_taps_t3(g.taps);
    g.Wt = w.w; g.N = w.O; g.D = y.p;
    g.R = residual ? residual->p : nullptr;
    g.bias = ws_->V(p + ".bias").p;
    if (gn) { g.gn_sums = gn->sums; g.gn_div = gn->over_time ? x.T : 1; }
    const bool fused = tapgemm(g, st_);
    if (gn) gn->fused = fused;
    launches++;
  }
  return y;
}

Act Model::concat(const Act& a, const Act& b, GnReq* gn) {
  // b may hold fewer samples than a (a skip tensor from the shared CFG prefix): its rows then repeat with period b.rows().
  // gn: per-frame GroupNorm statistics of the result, accumulated in the same pass (the consumer is a ResBlock).
  Act y = alloc(a.B, a.T, a.H, a.W, a.C + b.C);
  if (gn) gn_request(*gn, a.B * a.T);
  if (live()) {
    MUDG_REQUIRE(a.rows() % b.rows() == 0, "concat: %lld rows vs %lld", (long long)a.rows(), (long long)b.rows());
    ProfScope ps(PF_CONCAT, 0.0, 4.0 * (double)y.numel(), st_, fmt("%lldx%d", (long long)y.rows(), y.C).c_str());
    if (gn && !gn->over_time && y.C % 32 == 0 && knobs().gn_fuse != 0) {
      concat_channels_stats(a.p, a.C, b.p, b.C, y.p, a.B * a.T, (int64_t)a.H * a.W, b.rows(), gn->sums, st_);
      gn->fused = true;
    } else {
      concat_channels(a.p, a.C, b.p, b.C, y.p, a.rows(), b.rows(), st_);
    }
    launches++;
  }
  return y;
}

// x [B0, ...] -> [n, ...]: the samples tiled n / B0 times (sample i of the result = sample i % B0 of x)
Act Model::tile_batch(const Act& x, int n) {
  Act y = alloc(n, x.T, x.H, x.W, x.C);
  if (live()) {
    MUDG_REQUIRE(n % x.B == 0, "tile_batch: %d samples from %d", n, x.B);
    ProfScope ps(PF_CONCAT, 0.0, 2.0 * (double)(x.numel() + y.numel()), st_, "tile_batch");
    for (int i = 0; i < n / x.B; i++)
      MUDG_CUDA(cudaMemcpyAsync(y.p + (size_t)i * x.numel(), x.p, x.bytes(), cudaMemcpyDeviceToDevice, st_));
    launches += n / x.B;
  }
  return y;
}

Act Model::upsample(const Act& x) {
  Act y = alloc(x.B, x.T, 2 * x.H, 2 * x.W, x.C);
  if (live()) {
    ProfScope ps(PF_RESAMPLE, 0.0, 2.0 * (double)(x.numel() + y.numel()), st_, "upsample2x");
    upsample2x(x.p, y.p, x.B * x.T, x.H, x.W, x.C, st_);
    launches++;
  }
  return y;
}

Act Model::downsample(const Act& x, const std::string& p, int pad) {
  // pad 1: Conv2d 3x3 stride 2 padding 1 (UNet Downsample, openaimodel3d.py:66-70)
  // pad 0: F.pad (0,1,0,1) + Conv2d 3x3 stride 2 padding 0 (VAE Downsample, ae_modules.py:103-106)
  const int Ho = (x.H + pad - 2) / 2 + 1, Wo = (x.W + pad - 2) / 2 + 1;
  Act col = alloc(x.B, x.T, Ho, Wo, 9 * x.C);
  if (live()) {
    ProfScope ps(PF_RESAMPLE, 0.0, 2.0 * (double)(x.numel() + col.numel()), st_, "im2col_s2");
    im2col_s2(x.p, col.p, x.B * x.T, x.H, x.W, x.C, pad, st_);
    launches++;
  }
  Act y = linear(col, p + ".weight", p + ".bias", nullptr);
  release(col);
  return y;
}

// ================================================================ UNet blocks
// ResBlock._forward (openaimodel3d.py:210-236).  Every GroupNorm whose input comes out of a conv takes its statistics from
// that conv's epilogue (GnReq): out_layers.0 <- in_layers.2, temporal conv1..4 <- out_layers.3 / conv1..3, and -- when
// `out_gn` is given -- the norm of the SpatialTransformer that follows <- conv4.
Act Model::res_block(const Act& x, const Layer& l, GnReq* out_gn, GnReq* in_gn) {
  const std::string& p = l.prefix;
  Act a = group_norm(x, p + ".in_layers.0", 1e-5f, true, false, in_gn);
  GnReq r1;
  Act h = conv3x3(a, p + ".in_layers.2", nullptr, emb_out_[l.res_index], &r1);
  release(a);
  Act a2 = group_norm(h, p + ".out_layers.0", 1e-5f, true, false, &r1);
  release(h);
  Act skip;
  const bool has_skip = l.cin != l.cout;
  if (has_skip) skip = linear(x, p + ".skip_connection.weight", p + ".skip_connection.bias", nullptr);
  GnReq rt;
  rt.over_time = true;
  Act h2 = conv3x3(a2, p + ".out_layers.3", has_skip ? &skip : &x, nullptr, &rt);
  release(a2);
  if (has_skip) release(skip);
  // TemporalConvBlock (openaimodel3d.py:272-279): GroupNorm statistics span (C/32, T, H, W)
  Act y = h2;
  for (int j = 1; j <= 4; j++) {
    const std::string q = p + ".temopral_conv.conv" + std::to_string(j);
    Act n = group_norm(y, q + ".0", 1e-5f, true, true, &rt);
    if (j > 1) release(y);
    rt = GnReq{};
    rt.over_time = true;
    y = conv_t3(n, q + (j == 1 ? ".2" : ".3"), j == 4 ? &h2 : nullptr, j < 4 ? &rt : out_gn);
    release(n);
  }
  release(h2);
  return y;
}

// LN3 -> GEGLU FF -> +x   (BasicTransformerBlock._forward last line, attention.py:399); consumes x
Act Model::transformer_block_tail(Act x, const std::string& tb, LnReq* pre) {
  float2* s3 = layer_norm_stats(x, pre);
  Act hid = linear(x, tb + ".ff.net.0.proj.geglu.weight", tb + ".ff.net.0.proj.geglu.bias", nullptr, true, 1.f, s3);
  release_bytes(s3);
  Act y = linear(hid, tb + ".ff.net.2.weight", tb + ".ff.net.2.bias", &x);
  release(hid);
  release(x);
  return y;
}

Act Model::spatial_transformer(const Act& xin, const Layer& l, GnReq* in_gn) {   // attention.py:451-467, use_linear=True
  const std::string& p = l.prefix;
  const std::string tb = p + ".transformer_blocks.0";
  const int C = l.ch, HW = xin.H * xin.W;
  // Each LayerNorm's statistics come out of the epilogue of the GEMM that produces its input (LnReq) where that GEMM runs on
  // the pair kernel (layer_norm_stats then only reduces the partial sums); it reads the activation otherwise.
  LnReq r1, r2, r3;
  Act x = linear_gn(xin, p + ".norm", 1e-6f, false, p + ".proj_in.weight", p + ".proj_in.bias", in_gn, &r1);
  // attn1: self-attention over the H*W tokens of each frame
  float2* s1 = layer_norm_stats(x, &r1);
  Act qkv = linear(x, tb + ".attn1.qkv.weight", "", nullptr, false, 1.f, s1);
  release_bytes(s1);
  Act a1 = alloc(xin.B, xin.T, xin.H, xin.W, C);
  const int hw_pad = round_up(HW, 8);
  const int F1 = xin.B * xin.T;
  __half* vt = static_cast<__half*>(alloc_bytes(sizeof(__half) * (size_t)F1 * C * hw_pad));   // V^T [F][C][HW]
  if (live()) {
    {
      ProfScope ps(PF_TRANSPOSE_V, 0.0, 4.0 * (double)F1 * HW * C, st_, fmt("%dx%dx%d", F1, HW, C).c_str());
      transpose_v(qkv.p + 2 * C, 3 * C, HW, F1, l.heads, vt, hw_pad, st_);
    }
    // algorithmic: 4 * Nq * Nkv * 64 FLOPs per (frame, head); Q, K, V read and O written once
    ProfScope ps(PF_FLASH_SELF, 4.0 * (double)HW * HW * 64.0 * l.heads * F1, 8.0 * (double)F1 * HW * C, st_,
                 fmt("%dx%dx%d", F1, HW, l.heads).c_str());
    FlashArgs fa;
    fa.Q = qkv.p; fa.q_pitch = 3 * C; fa.O = a1.p; fa.o_pitch = C; fa.F = F1; fa.Nq = HW; fa.heads = l.heads;
    fa.nseg = 1;
    fa.seg[0].K = qkv.p + C; fa.seg[0].V = qkv.p + 2 * C; fa.seg[0].pitch = 3 * C; fa.seg[0].len = HW;
    fa.seg[0].VT = vt; fa.seg[0].vt_pitch = hw_pad;
    fa.seg[0].nbatch = F1; fa.seg[0].kv_div = 1;
    fa.scale = 0.125f;
    flash_attention(fa, st_);
    launches += 2;
  }
  release_bytes(vt);
  release(qkv);
  const bool split = xin.B < N_;          // (the partials of a tensor that is tiled over the CFG copies below are not)
  Act x1 = linear(a1, tb + ".attn1.to_out.0.weight", tb + ".attn1.to_out.0.bias", &x, false, 1.f, nullptr, split ? nullptr : &r2);
  release(a1);
  release(x);
  // Shared CFG prefix ends here: up to this point the `dup` copies of a sample (same latent, same t / label / fs) are
  // identical, from the first cross-attention on they differ.  Tile x1 and the block input (residual of proj_out).
  Act xin_full = xin;
  if (split) {
    Act t1 = tile_batch(x1, N_);
    release(x1);
    x1 = t1;
    xin_full = tile_batch(xin, N_);
  }
  const int F = xin_full.B * xin_full.T;
  // attn2: text (77 tokens, shared by the frames of a sample) + image tokens, separate softmaxes summed
  float2* s2 = layer_norm_stats(x1, &r2);
  Act q = linear(x1, tb + ".attn2.to_q.weight", "", nullptr, false, 1.f, s2);
  release_bytes(s2);
  Act a2 = alloc(xin_full.B, xin.T, xin.H, xin.W, C);
  if (live()) {
    auto it = kv_.find(p);
    MUDG_REQUIRE(it != kv_.end() && ctx_N_ == N_ && ctx_T_ == T_real_, "mudg_set_context(N=%d, T=%d) must precede the forward",
                 N_, T_real_);
    const KvCache& kc = it->second;
    const int lkv = ucfg_.text_context_len + (ctx_per_frame_ ? 16 : ctx_Limg_);
    ProfScope ps(PF_FLASH_CROSS, 4.0 * (double)HW * lkv * 64.0 * l.heads * F, 4.0 * (double)F * HW * C, st_,
                 fmt("%dx%dx%dx%d", F, HW, lkv, l.heads).c_str());
    if (ctx_per_frame_) {
      XattnArgs xa;
      xa.Q = q.p; xa.q_pitch = C; xa.O = a2.p; xa.o_pitch = C; xa.F = F; xa.Nq = HW; xa.heads = l.heads;
      xa.K = kc.kx; xa.VT = kc.vtx; xa.scale = 0.125f;
      xattn_per_frame(xa, st_);
      launches++;
    } else {
    FlashArgs fa;
    fa.Q = q.p; fa.q_pitch = C; fa.O = a2.p; fa.o_pitch = C; fa.F = F; fa.Nq = HW; fa.heads = l.heads;
    fa.nseg = 2;
    fa.seg[0].K = kc.text; fa.seg[0].V = kc.text + C; fa.seg[0].pitch = 2 * C; fa.seg[0].len = ucfg_.text_context_len;
    fa.seg[0].VT = kc.text_vt; fa.seg[0].vt_pitch = round_up(ucfg_.text_context_len, 8);
    fa.seg[0].nbatch = N_; fa.seg[0].kv_div = T_real_;
    fa.seg[1].K = kc.img; fa.seg[1].V = kc.img + C; fa.seg[1].pitch = 2 * C;
    fa.seg[1].VT = kc.img_vt; fa.seg[1].vt_pitch = round_up(ctx_per_frame_ ? 16 : ctx_Limg_, 8);
    if (ctx_per_frame_) { fa.seg[1].len = 16; fa.seg[1].nbatch = N_ * T_real_; fa.seg[1].kv_div = 1; }
    else { fa.seg[1].len = ctx_Limg_; fa.seg[1].nbatch = N_; fa.seg[1].kv_div = T_real_; }
    fa.scale = 0.125f;
    flash_attention(fa, st_);
    launches++;
    }
  }
  release(q);
  Act x2 = linear(a2, tb + ".attn2.to_out.0.weight", tb + ".attn2.to_out.0.bias", &x1, false, 1.f, nullptr, &r3);
  release(a2);
  release(x1);
  Act x3 = transformer_block_tail(x2, tb, &r3);
  Act y = linear(x3, p + ".proj_out.weight", p + ".proj_out.bias", &xin_full);
  release(x3);
  if (split) release(xin_full);
  return y;
}

Act Model::temporal_transformer(const Act& xin, const Layer& l) {   // attention.py:529-576
  const std::string& p = l.prefix;
  const std::string tb = p + ".transformer_blocks.0";
  const int HW = xin.H * xin.W;
  LnReq rq;                        // statistics of the next LayerNorm's input, from the GEMM that produces it
  Act x = linear_gn(xin, p + ".norm", 1e-6f, true, p + ".proj_in.weight", p + ".proj_in.bias", nullptr, &rq);
  for (int k = 1; k <= 2; k++) {   // attn1 and attn2 are both self-attention over T (only_self_att)
    const std::string an = tb + ".attn" + std::to_string(k);
    float2* sk = layer_norm_stats(x, &rq);
    Act qkv = linear(x, an + ".qkv.weight", "", nullptr, false, 1.f, sk);
    release_bytes(sk);
    rq = LnReq{};
    Act a = alloc(xin.B, xin.T, xin.H, xin.W, l.inner);
    if (live()) {
      ProfScope ps(PF_TATTN, 4.0 * (double)xin.T * xin.T * 64.0 * l.heads * xin.B * HW, 8.0 * (double)a.numel(), st_,
                   fmt("%dx%dx%dx%d", xin.B, xin.T, HW, l.heads).c_str());
      temporal_attention(qkv.p, a.p, xin.B, xin.T, HW, l.heads, 0.125f, st_);
      launches++;
    }
    release(qkv);
    Act xn = linear(a, an + ".to_out.0.weight", an + ".to_out.0.bias", &x, false, 1.f, nullptr, &rq);
    release(a);
    release(x);
    x = xn;
  }
  Act x3 = transformer_block_tail(x, tb, &rq);
  Act y = linear(x3, p + ".proj_out.weight", p + ".proj_out.bias", &xin);
  release(x3);
  return y;
}

// TimestepEmbedSequential dispatch (openaimodel3d.py:36-48).  Frees the intermediate tensors; the block input
// is freed only when `owns_input` (skip tensors stay alive until the output path consumes them).
Act Model::run_block(Act h, const Block& b, bool owns_input, GnReq* in_gn) {
  bool owned = owns_input;
  GnReq pend;                      // per-frame statistics of a ResBlock output for the SpatialTransformer norm that follows
  for (size_t li = 0; li < b.layers.size(); li++) {
    const Layer& l = b.layers[li];
    const bool next_spatial = li + 1 < b.layers.size() && b.layers[li + 1].kind == "spatial";
    Act y;
    if (l.kind == "res") {
      pend = GnReq{};
      y = res_block(h, l, next_spatial ? &pend : nullptr, (li == 0 && in_gn && in_gn->sums) ? in_gn : nullptr);
    }
    else if (l.kind == "spatial") y = spatial_transformer(h, l, pend.sums ? &pend : nullptr);
    else if (l.kind == "temporal") y = temporal_transformer(h, l);
    else if (l.kind == "down") y = downsample(h, l.prefix + ".op", 1);
    else if (l.kind == "up") { Act u = upsample(h); y = conv3x3(u, l.prefix + ".conv", nullptr, nullptr); release(u); }
    else if (l.kind == "conv") y = conv3x3(h, l.prefix, nullptr, nullptr);
    else throw Error("unknown layer kind " + l.kind);
    if (owned) release(h);
    owned = true;
    h = y;
  }
  return h;
}

// time_embed(t) + class_embed(label) + fps_embedding(fs), then every ResBlock's Linear(SiLU(emb))
// (openaimodel3d.py:569-576,594-602 and :219).  Computed on N rows, not N*T (emb is repeated over frames).
void Model::compute_embeddings(const int64_t* t, const int64_t* label, const int64_t* fs, int N) {
  // emb = time_embed(sin(t)) + class_embed(sin(label)) + fps_embedding(sin(fs)) (openaimodel3d.py:567-579); every ResBlock
  // then applies emb_layers = SiLU -> Linear (:169-176).  Four launches: the three sinusoids, the three first MLP layers
  // (+ SiLU), the three second layers summed (+ the emb_layers' SiLU), and all emb_layers Linears of the graph at once.
  const int mc = ucfg_.model_channels, ted = 4 * mc;
  float* sin_buf = static_cast<float*>(alloc_bytes(sizeof(float) * 3 * N * mc));
  float* hid = static_cast<float*>(alloc_bytes(sizeof(float) * 3 * N * ted));
  float* emb = static_cast<float*>(alloc_bytes(sizeof(float) * N * ted));       // SiLU(emb): the only form the graph uses
  const char* names[3] = {"time_embed", "class_embed", "fps_embedding"};
  ProfScope ps_embed(PF_EMBED, 0.0, 0.0, st_, "time/class/fps MLPs + emb_layers");
  if (live()) {
    sinusoid3(t, label, fs, sin_buf, N, mc, st_);
    LinearBatch l0, l2;
    l0.count = l2.count = 3;
    l0.K = mc; l2.K = ted;
    l0.Bn = l2.Bn = N;
    l0.silu_out = 1;                      // MLP = Linear -> SiLU -> Linear
    l2.sum = 1; l2.silu_out = 1;
    l0.off[0] = l2.off[0] = 0;
    for (int i = 0; i < 3; i++) {
      const std::string n = names[i];
      l0.x[i] = sin_buf + (size_t)i * N * mc; l0.W[i] = ws_->W(n + ".0.weight").w; l0.bias[i] = ws_->V(n + ".0.bias").p;
      l0.y[i] = hid + (size_t)i * N * ted; l0.off[i + 1] = (i + 1) * ted;
      l2.x[i] = hid + (size_t)i * N * ted; l2.W[i] = ws_->W(n + ".2.weight").w; l2.bias[i] = ws_->V(n + ".2.bias").p;
      l2.y[i] = emb; l2.off[i + 1] = ted;
    }
    batched_linear(l0, st_);
    batched_linear(l2, st_);
    launches += 3;
  }
  emb_out_.assign(n_res_, nullptr);
  LinearBatch le;
  le.K = ted; le.Bn = N; le.off[0] = 0;
  auto flush = [&]() {
    if (le.count && live()) {
      batched_linear(le, st_);
      launches++;
    }
    le.count = 0;
  };
  auto visit = [&](const Layer& l) {
    if (l.kind != "res") return;
    float* e = static_cast<float*>(alloc_bytes(sizeof(float) * N * l.cout));
    emb_out_[l.res_index] = e;
    if (live()) {
      const int j = le.count++;
      le.x[j] = emb; le.W[j] = ws_->W(l.prefix + ".emb_layers.1.weight").w; le.bias[j] = ws_->V(l.prefix + ".emb_layers.1.bias").p;
      le.y[j] = e; le.off[j + 1] = le.off[j] + l.cout;
      if (le.count == LinearBatch::MAX_JOBS) flush();
    }
  };
  for (auto& b : in_blocks_) for (auto& l : b.layers) visit(l);
  for (auto& l : mid_.layers) visit(l);
  for (auto& b : out_blocks_) for (auto& l : b.layers) visit(l);
  flush();
  release_bytes(sin_buf);
  release_bytes(hid);
  release_bytes(emb);
}

void Model::unet_body(const void* x, const int64_t* t, const int64_t* label, const int64_t* fs, int N, int dup, int T, int h,
                      int w, void* out) {
  ws_ = &unet_w;
  N_ = N; T_real_ = T;
  arena_.reset();
  compute_embeddings(t, label, fs, N);
  const int cpad = round_up(ucfg_.in_channels, 8);
  // Shared prefix (classifier-free guidance batches): x / t / label / fs hold N / dup distinct samples tiled dup times and
  // only the context differs, so everything before the first cross-attention runs on the N / dup distinct samples
  // (spatial_transformer tiles the batch back to N where the copies start to differ).
  bool has_spatial = false;
  for (auto& b : in_blocks_) for (auto& l : b.layers) has_spatial |= (l.kind == "spatial");
  const int B0 = (dup > 1 && has_spatial) ? N / dup : N;
  Act xin = alloc(B0, T, h, w, cpad);
  if (live()) {
    ProfScope ps(PF_LAYOUT, 0.0, 0.0, st_, "to_channels_last");
    to_channels_last(x, true, xin.p, B0, ucfg_.in_channels, (int64_t)T * h * w, cpad, st_);
    launches++;
  }
  std::vector<Act> hs;
  Act cur = run_block(xin, in_blocks_[0], true);
  {
    Layer ia; ia.kind = "temporal"; ia.prefix = "init_attn.0"; ia.ch = ucfg_.model_channels;
    ia.heads = ucfg_.init_attn_heads; ia.inner = ucfg_.init_attn_heads * 64; ia.linear_proj = false;
    Act y = temporal_transformer(cur, ia);
    release(cur);
    cur = y;
  }
  hs.push_back(cur);
  for (size_t i = 1; i < in_blocks_.size(); i++) {
    cur = run_block(cur, in_blocks_[i], false);   // the input is a skip tensor: keep it
    hs.push_back(cur);
  }
  cur = run_block(cur, mid_, false);              // hs.back() is still needed by output block 0
  for (const Block& b : out_blocks_) {
    Act skip = hs.back(); hs.pop_back();
    GnReq cat_gn;                                   // statistics for the first GroupNorm of the block, from the concat pass
    const bool res_first = !b.layers.empty() && b.layers[0].kind == "res";
    Act cat = concat(cur, skip, res_first ? &cat_gn : nullptr);
    release(cur);
    release(skip);
    cur = run_block(cat, b, true, &cat_gn);
  }
  // out: GroupNorm32 + SiLU + Conv3x3(model_channels -> out_channels), written as [N, Cout, T, h, w] fp16
  Act o = group_norm(cur, "out.0", 1e-5f, true, false);
  release(cur);
  // 4 output channels: run the conv on the tensor cores with the weight zero-padded to 64 rows, keep 4 columns
  Act y64 = alloc(N, T, h, w, 64);
  if (live()) {
    const Weight& wt = ws_->W("out.2.weight.pad");
    TapGemm g;
    g.A = o.p; g.B = 1; g.T = N * T; g.H = h; g.W = w; g.Cin = o.C;
    g.ntaps = 9; set_taps_3x3(g.taps);
    g.Wt = wt.w; g.N = wt.O; g.D = y64.p;
    g.bias = ws_->V("out.2.bias.pad").p;
    tapgemm(g, st_);
    ProfScope ps(PF_LAYOUT, 0.0, 0.0, st_, "from_channels_last");
    from_channels_last(y64.p, static_cast<__half*>(out), N, ucfg_.out_channels, (int64_t)T * h * w, 64, st_);
    launches += 2;
  }
  release(y64);
  release(o);
}

// ================================================================ context (cross-attention K/V cache)
static bool dbg_timing() {
  static const bool on = [] { const char* e = getenv("MUDG_DEBUG_TIMING"); return e && e[0] == '1'; }();
  return on;
}
static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void Model::set_context(const void* ctx, int dtype, int N, int L, int T, cudaStream_t st) {
  MUDG_REQUIRE(unet_ready_, "weights not finalized");
  const double t_begin = dbg_timing() ? now_ms() : 0.0;
  const int tl = ucfg_.text_context_len, D = ucfg_.context_dim;
  MUDG_REQUIRE(L > tl, "context needs image tokens after the %d text tokens (L=%d)", tl, L);
  const int Limg = L - tl;
  const bool per_frame = (L == tl + 16 * T);     // openaimodel3d.py:581 hard-coded split
  bool realloc = (N != ctx_N_ || L != ctx_L_ || T != ctx_T_);
  // fp16 copies of the text / image rows: persistent, grow-only staging (no cudaMalloc / cudaFree / host sync per call:
  // a caller that re-creates the context tensor every DDIM step would otherwise stall the whole pipeline on them)
  const size_t text_b = sizeof(__half) * (size_t)N * tl * D, img_b = sizeof(__half) * (size_t)N * Limg * D;
  if (ctx_text_stage_bytes_ < text_b) {
    MUDG_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx_text_stage_);
    MUDG_CUDA(cudaMalloc(&ctx_text_stage_, text_b));
    ctx_text_stage_bytes_ = text_b;
  }
  if (ctx_img_stage_bytes_ < img_b) {
    MUDG_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx_img_stage_);
    MUDG_CUDA(cudaMalloc(&ctx_img_stage_, img_b));
    ctx_img_stage_bytes_ = img_b;
  }
  __half *text = ctx_text_stage_, *img = ctx_img_stage_;
  gather_rows_f16(ctx, dtype == MUDG_F32, text, N, L, 0, tl, D, st);
  gather_rows_f16(ctx, dtype == MUDG_F32, img, N, L, tl, Limg, D, st);
  auto visit = [&](const Layer& l) {
    if (l.kind != "spatial") return;
    const int C = l.ch;
    KvCache& kc = kv_[l.prefix];
    const size_t tb = sizeof(__half) * (size_t)N * tl * 2 * C, ib = sizeof(__half) * (size_t)N * Limg * 2 * C;
    if (kc.text_bytes < tb) { MUDG_CUDA(cudaDeviceSynchronize()); cudaFree(kc.text); MUDG_CUDA(cudaMalloc(&kc.text, tb)); kc.text_bytes = tb; realloc = true; }
    if (kc.img_bytes < ib) { MUDG_CUDA(cudaDeviceSynchronize()); cudaFree(kc.img); MUDG_CUDA(cudaMalloc(&kc.img, ib)); kc.img_bytes = ib; realloc = true; }
    const std::string tbp = l.prefix + ".transformer_blocks.0.attn2.";
    for (int k = 0; k < 2; k++) {
      const Weight& w = unet_w.W(tbp + (k ? "kv_img.weight" : "kv_text.weight"));
      MUDG_REQUIRE(w.K() == D, "context_dim mismatch");
      TapGemm g;
      g.A = k ? img : text; g.B = 1; g.T = 1; g.H = 1; g.W = N * (k ? Limg : tl); g.Cin = D;
      g.ntaps = 1; g.Wt = w.w; g.N = w.O; g.D = k ? kc.img : kc.text;
      tapgemm(g, st);
      launches++;
    }
    // V^T of both segments for the tcgen05 attention kernel ([batches][C][len padded to 8], kv contiguous)
    const int heads = C / 64;
    const int tl_pad = round_up(tl, 8);
    const int li = per_frame ? 16 : Limg, li_pad = round_up(li, 8), nbi = per_frame ? N * T : N;
    const size_t tvb = sizeof(__half) * (size_t)N * C * tl_pad, ivb = sizeof(__half) * (size_t)nbi * C * li_pad;
    if (kc.text_vt_bytes < tvb) { cudaFree(kc.text_vt); MUDG_CUDA(cudaMalloc(&kc.text_vt, tvb)); kc.text_vt_bytes = tvb; realloc = true; }
    if (kc.img_vt_bytes < ivb) { cudaFree(kc.img_vt); MUDG_CUDA(cudaMalloc(&kc.img_vt, ivb)); kc.img_vt_bytes = ivb; realloc = true; }
    if (per_frame) {
      // one merged 96-key block per frame (xattn.cu); the two-segment flash kernel is not used for this context layout
      const size_t kb = sizeof(__half) * (size_t)N * T * 96 * C, vb = sizeof(__half) * (size_t)N * T * C * 128;
      if (kc.kx_bytes < kb) { cudaFree(kc.kx); MUDG_CUDA(cudaMalloc(&kc.kx, kb)); kc.kx_bytes = kb; realloc = true; }
      if (kc.vtx_bytes < vb) { cudaFree(kc.vtx); MUDG_CUDA(cudaMalloc(&kc.vtx, vb)); kc.vtx_bytes = vb; realloc = true; }
      MUDG_REQUIRE(tl == 77, "xattn: the merged block is laid out for 77 text tokens (got %d)", tl);
      xattn_pack(kc.text, kc.img, kc.kx, kc.vtx, N * T, T, C, st);
      launches++;
    } else {
      transpose_v(kc.text + C, 2 * C, tl, N, heads, kc.text_vt, tl_pad, st);
      transpose_v(kc.img + C, 2 * C, li, nbi, heads, kc.img_vt, li_pad, st);
      launches += 2;
    }
  };
  for (auto& b : in_blocks_) for (auto& l : b.layers) visit(l);
  for (auto& l : mid_.layers) visit(l);
  for (auto& b : out_blocks_) for (auto& l : b.layers) visit(l);
  if (dbg_timing()) fprintf(stderr, "[mudg] set_context: issued in %.1f ms, realloc %d\n", now_ms() - t_begin, (int)realloc);
  ctx_N_ = N; ctx_L_ = L; ctx_T_ = T; ctx_Limg_ = Limg; ctx_per_frame_ = per_frame;
  ctx_version_ += realloc ? 1 : 0;     // K/V buffers moved (or the token layout changed): graphs must be re-captured
}

// ================================================================ entry points
size_t Model::plan_unet(int N, int dup, int T, int h, int w) {
  MUDG_REQUIRE(unet_ready_, "weights not finalized");
  arena_.planning = true;
  planning_ = true;
  prof_pause(true);
  arena_.reset_high();
  unet_body(nullptr, nullptr, nullptr, nullptr, N, dup, T, h, w, nullptr);
  const size_t need = arena_.high_water();
  arena_.planning = false;
  planning_ = false;
  prof_pause(false);
  arena_.reset();
  return need;
}

// The ~1300 launches of one forward are captured once per shape into a CUDA graph (inputs/outputs go through fixed
// staging buffers so the captured pointers stay valid); later steps replay it.  MUDG_GRAPH=0 disables the capture.
void Model::unet_forward(const void* x, const int64_t* t, const int64_t* label, const int64_t* fs, int N, int dup, int T, int h,
                         int w, void* out, cudaStream_t st) {
  MUDG_REQUIRE(unet_ready_, "weights not finalized");
  MUDG_REQUIRE(dup >= 1 && N % dup == 0, "unet_forward: N = %d is not a multiple of dup = %d", N, dup);
  std::array<int, 5> key{N, dup, T, h, w};
  auto it = unet_plans_.find(key);
  if (it == unet_plans_.end()) it = unet_plans_.emplace(key, plan_unet(N, dup, T, h, w)).first;
  ensure_arena(it->second);
  st_ = st;
  static const bool graphs_on = [] {
    const char* e = getenv("MUDG_GRAPH");
    return !(e && e[0] == '0');
  }();
  if (!graphs_on || prof_active()) {
    unet_body(x, t, label, fs, N, dup, T, h, w, out);
    return;
  }
  // Capture is illegal on the legacy default stream (what PyTorch uses unless told otherwise): run on an own
  // *blocking* stream there, which the legacy stream implicitly orders against on both sides.
  if (st == nullptr || st == cudaStreamLegacy) {
    if (!own_stream_) MUDG_CUDA(cudaStreamCreate(&own_stream_));
    st = own_stream_;
    st_ = st;
  }
  GraphSlot& g = graphs_[key];
  const size_t xin = sizeof(float) * (size_t)N * ucfg_.in_channels * T * h * w;
  const size_t xout = sizeof(__half) * (size_t)N * ucfg_.out_channels * T * h * w;
  if (dbg_timing() && g.ctx_version != ctx_version_) fprintf(stderr, "[mudg] unet_forward: graph invalidated (context version)\n");
  if (g.ctx_version != ctx_version_) {                     // set_context may have re-allocated the K/V caches
    for (auto& e : g.ring) { if (e) cudaGraphExecDestroy(e); e = nullptr; }
    g.exec = nullptr;
    g.runs = 0;
    g.ctx_version = ctx_version_;
  }
  if (!g.in_x) {
    MUDG_CUDA(cudaMalloc(&g.in_x, xin));
    MUDG_CUDA(cudaMalloc(&g.in_idx, sizeof(int64_t) * 3 * N));
    MUDG_CUDA(cudaMalloc(&g.out, xout));
  }
  MUDG_CUDA(cudaMemcpyAsync(g.in_x, x, xin, cudaMemcpyDeviceToDevice, st));
  MUDG_CUDA(cudaMemcpyAsync(g.in_idx, t, sizeof(int64_t) * N, cudaMemcpyDeviceToDevice, st));
  MUDG_CUDA(cudaMemcpyAsync(g.in_idx + N, label, sizeof(int64_t) * N, cudaMemcpyDeviceToDevice, st));
  MUDG_CUDA(cudaMemcpyAsync(g.in_idx + 2 * N, fs, sizeof(int64_t) * N, cudaMemcpyDeviceToDevice, st));
  if (g.exec) {
    MUDG_CUDA(cudaGraphLaunch(g.ring[g.next], st));
    g.next = (g.next + 1) % GraphSlot::GRAPH_RING;
    launches += g.launches;
  } else if (g.runs == 0) {
    // first call: eager (sets kernel attributes, fills the tensor-map cache)
    const int64_t l0 = launches;
    unet_body(g.in_x, g.in_idx, g.in_idx + N, g.in_idx + 2 * N, N, dup, T, h, w, g.out);
    g.launches = launches - l0;
  } else {
    cudaGraph_t graph = nullptr;
    MUDG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      unet_body(g.in_x, g.in_idx, g.in_idx + N, g.in_idx + 2 * N, N, dup, T, h, w, g.out);
    } catch (...) {
      cudaStreamEndCapture(st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    MUDG_CUDA(cudaStreamEndCapture(st, &graph));
    for (auto& e : g.ring) MUDG_CUDA(cudaGraphInstantiate(&e, graph, 0));
    cudaGraphDestroy(graph);
    g.exec = g.ring[0];
    g.next = 1;
    MUDG_CUDA(cudaGraphLaunch(g.ring[0], st));
  }
  g.runs++;
  MUDG_CUDA(cudaMemcpyAsync(out, g.out, xout, cudaMemcpyDeviceToDevice, st));
}

void Model::drop_graphs() {
  for (auto& kv : graphs_) {
    for (auto& e : kv.second.ring) { if (e) cudaGraphExecDestroy(e); e = nullptr; }
    kv.second.exec = nullptr;
    kv.second.runs = 0;
  }
}

// ================================================================ VAE decoder (ae_modules.py:466-578)
Act Model::gemm_raw(const __half* A, int M, int K, const __half* Wt, int N, const float* bias, const Act* residual,
                    float alpha) {
  Act y = alloc(1, 1, 1, M, N);
  if (live()) {
    TapGemm g;
    g.A = A; g.B = 1; g.T = 1; g.H = 1; g.W = M; g.Cin = K; g.ntaps = 1;
    g.Wt = Wt; g.N = N; g.D = y.p; g.R = residual ? residual->p : nullptr; g.bias = bias; g.alpha = alpha;
    tapgemm(g, st_);
    launches++;
  }
  return y;
}

Act Model::vae_res(const Act& x, const std::string& p) {   // ResnetBlock.forward, temb=None (ae_modules.py:190-210)
  Act a = group_norm(x, p + ".norm1", 1e-6f, true, false);
  Act h = conv3x3(a, p + ".conv1", nullptr, nullptr);
  release(a);
  Act a2 = group_norm(h, p + ".norm2", 1e-6f, true, false);
  release(h);
  const bool has_skip = ws_->hasW(p + ".nin_shortcut.weight");
  Act skip;
  if (has_skip) skip = linear(x, p + ".nin_shortcut.weight", p + ".nin_shortcut.bias", nullptr);
  Act y = conv3x3(a2, p + ".conv2", has_skip ? &skip : &x, nullptr);
  release(a2);
  if (has_skip) release(skip);
  return y;
}

// AttnBlock.forward (ae_modules.py:53-78): one head, d = C.  Scores are materialised once per frame
// (N x N fp16, 170 MB at 72x128) -- 0.05 % of a clip's FLOPs, so three plain GEMMs + a row softmax.
Act Model::vae_attn(const Act& x, const std::string& p) {
  const int M = (int)x.rows(), C = x.C;
  Act y = group_norm(x, p + ".norm", 1e-6f, false, false);
  Act q = linear(y, p + ".q.weight", p + ".q.bias", nullptr);
  Act k = linear(y, p + ".k.weight", p + ".k.bias", nullptr);
  // V^T[c, token] = Wv[c, :] . y[token, :]  (bias added after P V: softmax rows sum to 1)
  const Weight& wv = ws_->W(p + ".v.weight");
  Act vt = gemm_raw(wv.w, C, C, y.p, M, nullptr, nullptr, 1.f);
  release(y);
  Act s = gemm_raw(q.p, M, C, k.p, M, nullptr, nullptr, 1.f / sqrtf((float)C));
  release(q);
  release(k);
  if (live()) {
    ProfScope ps(PF_SOFTMAX, 0.0, 4.0 * (double)M * M, st_, "vae_attn");
    softmax_rows(s.p, M, M, st_);
    launches++;
  }
  Act o = gemm_raw(s.p, M, M, vt.p, C, ws_->V(p + ".v.bias").p, nullptr, 1.f);
  release(s);
  release(vt);
  o.B = x.B; o.T = x.T; o.H = x.H; o.W = x.W;
  Act out = linear(o, p + ".proj_out.weight", p + ".proj_out.bias", &x);
  release(o);
  return out;
}

void Model::vae_body(const void* z, int h, int w, void* out) {
  ws_ = &vae_w;
  N_ = 1; T_real_ = 1;
  arena_.reset();
  const MudgVaeConfig& v = vcfg_;
  const int cpad = round_up(v.z_channels, 8);
  Act zin = alloc(1, 1, h, w, cpad);
  Act zq = alloc(1, 1, h, w, cpad);
  if (live()) {
    to_channels_last(z, true, zin.p, 1, v.z_channels, (int64_t)h * w, cpad, st_);
    MUDG_CUDA(cudaMemsetAsync(zq.p, 0, zq.bytes(), st_));
    const Weight& wt = ws_->W("post_quant_conv.weight");       // 1x1 conv (autoencoder.py:105)
    TapGemmGeneric g;
    g.A = zin.p; g.B = 1; g.T = 1; g.H = h; g.W = w; g.Cin = wt.Ipad;
    g.a_sc = 1; g.a_sw = cpad; g.a_sh = (int64_t)w * cpad;
    g.ntaps = 1; g.Wt = wt.w; g.CinW = wt.Ipad; g.N = wt.O;
    g.D = zq.p; g.d_sn = 1; g.d_sw = cpad; g.d_sh = (int64_t)w * cpad;
    g.bias = ws_->V("post_quant_conv.bias").p;
    tapgemm_generic(g, st_);
    launches += 3;
  }
  release(zin);
  Act cur = conv3x3(zq, "decoder.conv_in", nullptr, nullptr);
  release(zq);
  auto step = [&](Act y) { release(cur); cur = y; };
  step(vae_res(cur, "decoder.mid.block_1"));
  step(vae_attn(cur, "decoder.mid.attn_1"));
  step(vae_res(cur, "decoder.mid.block_2"));
  for (int lvl = v.n_ch_mult - 1; lvl >= 0; lvl--) {
    for (int ib = 0; ib <= v.num_res_blocks; ib++)
      step(vae_res(cur, "decoder.up." + std::to_string(lvl) + ".block." + std::to_string(ib)));
    if (lvl != 0) {
      Act u = upsample(cur);
      release(cur);
      cur = conv3x3(u, "decoder.up." + std::to_string(lvl) + ".upsample.conv", nullptr, nullptr);
      release(u);
    }
  }
  Act o = group_norm(cur, "decoder.norm_out", 1e-6f, true, false);
  release(cur);
  // 3 output channels: tensor-core path with the weight zero-padded to 64 rows, then keep 3 columns as NCHW
  Act y64 = alloc(1, 1, o.H, o.W, 64);
  if (live()) {
    const Weight& wt = ws_->W("decoder.conv_out.weight.pad");
    TapGemm g;
    g.A = o.p; g.B = 1; g.T = 1; g.H = o.H; g.W = o.W; g.Cin = o.C;
    g.ntaps = 9; set_taps_3x3(g.taps);
    g.Wt = wt.w; g.N = wt.O; g.D = y64.p;
    g.bias = ws_->V("decoder.conv_out.bias.pad").p;
    tapgemm(g, st_);
    from_channels_last(y64.p, static_cast<__half*>(out), 1, vcfg_.out_ch, (int64_t)o.H * o.W, 64, st_);
    launches += 2;
  }
  release(y64);
  release(o);
}

size_t Model::plan_vae(int h, int w) {
  MUDG_REQUIRE(vae_ready_, "VAE weights not finalized");
  arena_.planning = true;
  planning_ = true;
  prof_pause(true);
  arena_.reset_high();
  vae_body(nullptr, h, w, nullptr);
  const size_t need = arena_.high_water();
  arena_.planning = false;
  planning_ = false;
  prof_pause(false);
  arena_.reset();
  return need;
}

void Model::vae_decode(const void* z, int F, int h, int w, void* out, cudaStream_t st) {
  MUDG_REQUIRE(vae_ready_, "VAE weights not finalized");
  std::array<int, 2> key{h, w};
  auto it = vae_plans_.find(key);
  if (it == vae_plans_.end()) it = vae_plans_.emplace(key, plan_vae(h, w)).first;
  ensure_arena(it->second);
  st_ = st;
  const MudgVaeConfig& v = vcfg_;
  for (int f = 0; f < F; f++)   // perframe_ae loop (ddpm3d.py:659-664)
    vae_body(static_cast<const float*>(z) + (size_t)f * v.z_channels * h * w, h, w,
             static_cast<__half*>(out) + (size_t)f * v.out_ch * 64 * h * w);
}

// ================================================================ VAE encoder (ae_modules.py:364-463) -- "next" row (f)1
void Model::vae_encode_body(const void* x, int H, int W, void* moments) {
  ws_ = &vae_w;
  N_ = 1; T_real_ = 1;
  arena_.reset();
  const MudgVaeConfig& v = vcfg_;
  Act xin = alloc(1, 1, H, W, 8);
  if (live()) {
    to_channels_last(x, true, xin.p, 1, 3, (int64_t)H * W, 8, st_);
    launches++;
  }
  Act cur = conv3x3(xin, "encoder.conv_in", nullptr, nullptr);
  release(xin);
  auto step = [&](Act y) { release(cur); cur = y; };
  for (int lvl = 0; lvl < v.n_ch_mult; lvl++) {
    for (int ib = 0; ib < v.num_res_blocks; ib++)
      step(vae_res(cur, "encoder.down." + std::to_string(lvl) + ".block." + std::to_string(ib)));
    if (lvl != v.n_ch_mult - 1) step(downsample(cur, "encoder.down." + std::to_string(lvl) + ".downsample.conv", 0));
  }
  step(vae_res(cur, "encoder.mid.block_1"));
  step(vae_attn(cur, "encoder.mid.attn_1"));
  step(vae_res(cur, "encoder.mid.block_2"));
  Act o = group_norm(cur, "encoder.norm_out", 1e-6f, true, false);
  release(cur);
  // conv_out (-> 2z = 8 channels) on the tensor cores with the weight zero-padded to 64 rows, then quant_conv 1x1
  Act y64 = alloc(1, 1, o.H, o.W, 64);
  if (live()) {
    const Weight& wt = ws_->W("encoder.conv_out.weight.pad");
    TapGemm g;
    g.A = o.p; g.B = 1; g.T = 1; g.H = o.H; g.W = o.W; g.Cin = o.C;
    g.ntaps = 9; set_taps_3x3(g.taps);
    g.Wt = wt.w; g.N = wt.O; g.D = y64.p;
    g.bias = ws_->V("encoder.conv_out.bias.pad").p;
    tapgemm(g, st_);
    const Weight& wq = ws_->W("quant_conv.weight");
    const int h = o.H, w = o.W;
    TapGemmGeneric q;
    q.A = y64.p; q.B = 1; q.T = 1; q.H = h; q.W = w; q.Cin = wq.I;
    q.a_sc = 1; q.a_sw = 64; q.a_sh = (int64_t)w * 64;
    q.ntaps = 1; q.Wt = wq.w; q.CinW = wq.Ipad; q.N = wq.O;
    q.D = moments; q.d_fp32 = true; q.d_sw = 1; q.d_sh = w; q.d_sn = (int64_t)h * w;
    q.bias = ws_->V("quant_conv.bias").p;
    tapgemm_generic(q, st_);
    launches += 2;
  }
  release(y64);
  release(o);
}

void Model::vae_encode(const void* x, int F, int H, int W, void* moments, cudaStream_t st) {
  MUDG_REQUIRE(vae_enc_ready_, "VAE encoder weights not loaded");
  MUDG_REQUIRE(H % 8 == 0 && W % 8 == 0, "image size must be a multiple of 8");
  std::array<int, 2> key{-H, -W};
  auto it = vae_plans_.find(key);
  if (it == vae_plans_.end()) {
    arena_.planning = true; planning_ = true; arena_.reset_high();
    prof_pause(true);
    vae_encode_body(nullptr, H, W, nullptr);
    prof_pause(false);
    const size_t need = arena_.high_water();
    arena_.planning = false; planning_ = false; arena_.reset();
    it = vae_plans_.emplace(key, need).first;
  }
  ensure_arena(it->second);
  st_ = st;
  const MudgVaeConfig& v = vcfg_;
  for (int f = 0; f < F; f++)
    vae_encode_body(static_cast<const float*>(x) + (size_t)f * 3 * H * W, H, W,
                    static_cast<float*>(moments) + (size_t)f * 2 * v.z_channels * (H / 8) * (W / 8));
}

}  // namespace mudg

// ================================================================ Resampler ("next" row f.3; resampler.py:48-144)
// Once per clip: [B, 257, 1280] image-encoder tokens -> [B, 16*T, 1024] image context.  Same kernels as the UNet: tcgen05
// tap-GEMM for every Linear (bias / residual fused), flash attention (d = 64) over the [x ; latents] keys, warp-per-row
// LayerNorm, plus an in-place erf GELU.
namespace mudg {

void Model::resampler_body(const void* x, int dtype, int B, int L, void* out) {
  ws_ = &res_w;
  arena_.reset();
  const ResamplerDims& r = rs_;
  const int Lkv = L + r.nq;
  auto rows = [&](int n_rows, int width) { return alloc(1, 1, 1, n_rows, width); };
  Act xin = rows(B * L, r.emb);
  Act lat = rows(B * r.nq, r.dim);
  if (live()) {
    cast_to_f16(x, dtype == MUDG_F32, xin.p, (int64_t)B * L * r.emb, st_);
    for (int b = 0; b < B; b++)      // latents.repeat(B, 1, 1)
      MUDG_CUDA(cudaMemcpyAsync(lat.p + (size_t)b * r.nq * r.dim, res_w.W("latents").w, sizeof(__half) * (size_t)r.nq * r.dim,
                                cudaMemcpyDeviceToDevice, st_));
    launches += 1 + B;
  }
  Act xp = linear(xin, "proj_in.weight", "proj_in.bias", nullptr);
  release(xin);
  for (int i = 0; i < r.depth; i++) {
    const std::string a = "layers." + std::to_string(i) + ".0", f = "layers." + std::to_string(i) + ".1";
    // PerceiverAttention.forward (resampler.py:66-101)
    Act xn = layer_norm(xp, a + ".norm1");
    Act ln = layer_norm(lat, a + ".norm2");
    Act q = linear(ln, a + ".to_q.weight", "", nullptr);
    Act kvin = rows(B * Lkv, r.dim);               // torch.cat((x, latents), dim=-2)
    if (live()) {
      const size_t rowb = sizeof(__half) * (size_t)r.dim;
      MUDG_CUDA(cudaMemcpy2DAsync(kvin.p, rowb * Lkv, xn.p, rowb * L, rowb * L, B, cudaMemcpyDeviceToDevice, st_));
      MUDG_CUDA(cudaMemcpy2DAsync(kvin.p + (size_t)L * r.dim, rowb * Lkv, ln.p, rowb * r.nq, rowb * r.nq, B,
                                  cudaMemcpyDeviceToDevice, st_));
      launches += 2;
    }
    release(xn);
    release(ln);
    Act kv = linear(kvin, a + ".to_kv.weight", "", nullptr);      // [B*Lkv][k | v]
    release(kvin);
    const int pad = round_up(Lkv, 8);
    __half* vt = static_cast<__half*>(alloc_bytes(sizeof(__half) * (size_t)B * r.inner * pad));
    Act o = rows(B * r.nq, r.inner);
    if (live()) {
      transpose_v(kv.p + r.inner, 2 * r.inner, Lkv, B, r.heads, vt, pad, st_);
      FlashArgs fa;
      fa.Q = q.p; fa.q_pitch = r.inner; fa.O = o.p; fa.o_pitch = r.inner; fa.F = B; fa.Nq = r.nq; fa.heads = r.heads;
      fa.nseg = 1;
      fa.seg[0].K = kv.p; fa.seg[0].V = kv.p + r.inner; fa.seg[0].pitch = 2 * r.inner; fa.seg[0].len = Lkv;
      fa.seg[0].VT = vt; fa.seg[0].vt_pitch = pad; fa.seg[0].nbatch = B; fa.seg[0].kv_div = 1;
      fa.scale = 0.125f;                            // (q * 64^-1/4) . (k * 64^-1/4)
      flash_attention(fa, st_);
      launches += 2;
    }
    release_bytes(vt);
    release(kv);
    release(q);
    Act lat2 = linear(o, a + ".to_out.weight", "", &lat);         // attn(x, latents) + latents
    release(o);
    release(lat);
    // FeedForward (resampler.py:31-37): LayerNorm -> Linear -> GELU -> Linear, + latents
    Act hn = layer_norm(lat2, f + ".0");
    Act h1 = linear(hn, f + ".1.weight", "", nullptr);
    release(hn);
    if (live()) {
      gelu_inplace(h1.p, (int64_t)h1.rows() * h1.C, st_);
      launches++;
    }
    lat = linear(h1, f + ".3.weight", "", &lat2);
    release(h1);
    release(lat2);
  }
  release(xp);
  Act po = linear(lat, "proj_out.weight", "proj_out.bias", nullptr);
  release(lat);
  Act y = layer_norm(po, "norm_out");
  release(po);
  if (live()) {
    cast_to_f32(y.p, false, static_cast<float*>(out), (int64_t)y.rows() * y.C, st_);
    launches++;
  }
  release(y);
}

void Model::resampler_forward(const void* x, int dtype, int B, int L, void* out, cudaStream_t st) {
  MUDG_REQUIRE(res_ready_, "Resampler weights not finalized");
  MUDG_REQUIRE(B >= 1 && L >= 1, "Resampler: empty input");
  arena_.planning = true; planning_ = true;
  arena_.reset_high();
  prof_pause(true);
  resampler_body(nullptr, dtype, B, L, nullptr);
  prof_pause(false);
  const size_t need = arena_.high_water();
  arena_.planning = false; planning_ = false;
  arena_.reset();
  ensure_arena(need);
  st_ = st;
  resampler_body(x, dtype, B, L, out);
}

}  // namespace mudg
