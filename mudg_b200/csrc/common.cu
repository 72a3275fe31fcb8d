#include "common.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace mudg {

static thread_local std::string g_last_error;

std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}
void set_last_error(const std::string& s) { g_last_error = s; }
const char* last_error_cstr() { return g_last_error.c_str(); }

Knobs& knobs() {
  static Knobs k;
  return k;
}

// ---------------------------------------------------------------- profiler
namespace {
struct ProfRec {
  int fam;
  double flops, bytes;
  std::string shape;
};
struct Profiler {
  bool on = false, paused = false;
  std::vector<cudaEvent_t> ev;      // pairs, grow-only
  std::vector<ProfRec> rec;
} g_profiler;
const char* kFamNames[PF_COUNT] = {"gemm", "flash_self", "flash_cross", "temporal_attn", "groupnorm", "ln_stats", "layernorm",
                                   "transpose_v", "concat", "resample", "layout", "embed", "softmax", "generic_conv",
                                   "ddim_step", "postdecode"};
}  // namespace

void prof_enable(bool on) {
  g_profiler.on = on;
  g_profiler.rec.clear();
}
bool prof_active() { return g_profiler.on && !g_profiler.paused; }
void prof_pause(bool paused) { g_profiler.paused = paused; }

ProfScope::ProfScope(int fam, double flops, double bytes, cudaStream_t s, const char* shape) : st(s) {
  Profiler& p = g_profiler;
  if (!p.on || p.paused) return;
  idx = (int)p.rec.size();
  while (p.ev.size() < 2 * (size_t)(idx + 1)) {
    cudaEvent_t e;
    MUDG_CUDA(cudaEventCreate(&e));
    p.ev.push_back(e);
  }
  p.rec.push_back(ProfRec{fam, flops, bytes, shape ? shape : ""});
  cudaEventRecord(p.ev[2 * idx], st);
}
ProfScope::~ProfScope() {
  if (idx >= 0) cudaEventRecord(g_profiler.ev[2 * idx + 1], st);
}

std::string prof_report() {
  Profiler& p = g_profiler;
  MUDG_CUDA(cudaDeviceSynchronize());
  struct Agg { double n = 0, ms = 0, flops = 0, bytes = 0; };
  std::map<std::pair<int, std::string>, Agg> agg;
  for (size_t i = 0; i < p.rec.size(); i++) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.ev[2 * i], p.ev[2 * i + 1]) != cudaSuccess) continue;
    Agg& a = agg[{p.rec[i].fam, p.rec[i].shape}];
    a.n += 1; a.ms += t; a.flops += p.rec[i].flops; a.bytes += p.rec[i].bytes;
  }
  std::string out = "family,shape,launches,ms,flops,bytes\n";
  for (auto& kv : agg)
    out += fmt("%s,%s,%.0f,%.5f,%.6e,%.6e\n", kFamNames[kv.first.first], kv.first.second.c_str(), kv.second.n, kv.second.ms,
               kv.second.flops, kv.second.bytes);
  p.rec.clear();
  return out;
}

int sm_count() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

// ---------------------------------------------------------------- Arena
Arena::~Arena() {
  if (base_) cudaFree(base_);
}
void Arena::reserve(size_t bytes) {
  if (base_) {
    cudaFree(base_);
    base_ = nullptr;
  }
  cap_ = 0;
  if (bytes) {
    MUDG_CUDA(cudaMalloc(&base_, bytes));
    cap_ = bytes;
  }
  reset();
  clear_tmap_cache();
}
void Arena::reset() {
  top_ = 0;
  free_.clear();
  live_.clear();
}
void* Arena::alloc(size_t bytes) {
  size_t sz = (bytes + 1023) & ~size_t(1023);
  if (sz == 0) sz = 1024;
  size_t off;
  auto it = free_.find(sz);
  if (it != free_.end() && !it->second.empty()) {
    off = it->second.back();
    it->second.pop_back();
  } else {
    off = top_;
    top_ += sz;
    if (top_ > high_) high_ = top_;
    if (!planning && top_ > cap_)
      throw Error(fmt("arena exhausted: need %zu bytes, capacity %zu", top_, cap_));
  }
  live_[off] = sz;
  // planning mode hands out fake (never dereferenced) addresses with the same offsets
  return (planning ? reinterpret_cast<char*>(uintptr_t(1) << 40) : base_) + off;
}
void Arena::free(void* p) {
  if (!p) return;
  char* b = planning ? reinterpret_cast<char*>(uintptr_t(1) << 40) : base_;
  size_t off = static_cast<char*>(p) - b;
  auto it = live_.find(off);
  if (it == live_.end()) throw Error("arena: free of unknown block");
  free_[it->second].push_back(off);
  live_.erase(it);
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MUDG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    MUDG_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

using TmapKey = std::array<uint64_t, 15>;
static std::map<TmapKey, CUtensorMap>& tmap_cache() {
  static std::map<TmapKey, CUtensorMap> c;
  return c;
}
void clear_tmap_cache() { tmap_cache().clear(); }

const CUtensorMap* get_tmap(const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                            const uint32_t box[5]) {
  TmapKey key;
  key[0] = reinterpret_cast<uint64_t>(base);
  for (int i = 0; i < 5; i++) key[1 + i] = dims[i];
  for (int i = 0; i < 4; i++) key[6 + i] = strides_bytes[i];
  for (int i = 0; i < 5; i++) key[10 + i] = box[i];
  auto& cache = tmap_cache();
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  MUDG_REQUIRE((reinterpret_cast<uint64_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < 5; i++) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    MUDG_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i < 4; i++) {
    gstr[i] = strides_bytes[i];
    MUDG_REQUIRE((strides_bytes[i] & 15) == 0, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                 (unsigned long long)strides_bytes[i]);
  }
  MUDG_REQUIRE(box[0] * 2 <= 128, "inner box exceeds the 128B swizzle span");
  CUtensorMap m;
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), gdim, gstr, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(fmt("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu %llu box %u %u %u %u %u", (int)r,
                    (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                    (unsigned long long)dims[3], (unsigned long long)dims[4], box[0], box[1], box[2], box[3], box[4]));
  auto ins = cache.emplace(key, m);
  return &ins.first->second;
}

}  // namespace mudg
