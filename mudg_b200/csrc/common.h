// Host-side plumbing shared by every translation unit of libmudg_sm100.so:
// error convention, device arena, tensor views, TMA tensor-map factory.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace mudg {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
std::string fmt(const char* f, ...);
void set_last_error(const std::string& s);

#define MUDG_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::mudg::Error(::mudg::fmt("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e))); \
  } while (0)
#define MUDG_REQUIRE(cond, ...)                                                                      \
  do {                                                                                               \
    if (!(cond)) throw ::mudg::Error(::mudg::fmt("%s:%d require(%s): ", __FILE__, __LINE__, #cond) + ::mudg::fmt(__VA_ARGS__)); \
  } while (0)

// ---------------------------------------------------------------- activations
// Channels-last activation: element (b,t,h,w,c) at ((((b*T+t)*H+h)*W+w)*C + c).  Spatial ops see
// frames F=B*T; temporal ops see the (B,T) split; linear ops see rows M = B*T*H*W.
struct Act {
  __half* p = nullptr;
  int B = 1, T = 1, H = 1, W = 1, C = 1;
  int64_t rows() const { return (int64_t)B * T * H * W; }
  int64_t numel() const { return rows() * C; }
  size_t bytes() const { return (size_t)numel() * sizeof(__half); }
  int frames() const { return B * T; }
};

// ---------------------------------------------------------------- arena
// Size-bucketed caching sub-allocator over one cudaMalloc'd slab.  Deterministic (same call sequence ->
// same addresses), so tensor maps can be cached and the step can be captured in a CUDA graph.
// In planning mode nothing is allocated; only the high-water mark is recorded.
class Arena {
 public:
  ~Arena();
  void reserve(size_t bytes);          // (re)allocate the slab; invalidates everything
  void reset();                        // drop all live blocks (start of a forward)
  void* alloc(size_t bytes);
  void free(void* p);
  size_t capacity() const { return cap_; }
  size_t high_water() const { return high_; }
  void reset_high() { high_ = 0; }
  bool planning = false;

 private:
  char* base_ = nullptr;
  size_t cap_ = 0, top_ = 0, high_ = 0;
  std::map<size_t, std::vector<size_t>> free_;    // size -> offsets
  std::unordered_map<size_t, size_t> live_;       // offset -> size
};

// ---------------------------------------------------------------- TMA tensor maps
// rank-5 fp16 tiled map, SWIZZLE_128B, zero OOB fill.  dims/box innermost first; strides in BYTES for dims 1..4.
const CUtensorMap* get_tmap(const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                            const uint32_t box[5]);
void clear_tmap_cache();

struct Stream {
  cudaStream_t s = nullptr;
};

int sm_count();
// first() is true once per CUDA device: function attributes (dynamic shared memory limits) are per device, a process may hold
// contexts on several
struct OncePerDevice {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Tuning / diagnostic knobs.  The product library never reads the environment for these: they hold the shipped defaults
// and can only be changed through libmudg_sm100_test.so (csrc/test/testhooks.cu, tests/ only), which links its own
// copy of these objects.
struct Knobs {
  int gemm_pair = -1;       // -1 heuristic, 0 never, 1 always use the CTA-pair kernel (tapgemm_tc3)
  int gemm_sub = 0;         // 0 heuristic, 1 | 2: force the M sub-tile count of the single-CTA kernel (tapgemm_tc2)
  int gemm_epi = 1;         // compile-time epilogue variants of tapgemm_tc3 (0 = run-time variant only)
  int gemm_groups = 0;      // epilogue groups of tapgemm_tc3: 0 per-shape rule, 2 | 3 forced
  int gemm_resmma = 1;      // tapgemm_tc3: residual added by the tensor core (identity k-steps); 0 = by the epilogue threads
  int gemm_deep = 1;        // tapgemm_tc2: 6-stage, one-CTA-per-SM variant for launches of at most one tile per SM
  int gemm_balance = 1;     // tapgemm_tc3: narrower N tiles when they balance the waves over the CTA pairs
  int gemm_dbg = 0;         // bit 0: skip the output TMA stores, bit 1: release (not relaxed) accumulator hand-back
  int flash_stagger = -1;   // MUFU hand-over point of the two softmax groups (-1 per-split default, 0 off, 1 end, 2 / 3 / 4 after 3/4, 1/2, 1/4)
  int flash_poly = 0;       // polynomial exp2 on 1/4 (1) or 1/2 (2) of the exponentials
  int flash_split = 0;      // softmax warps per TMEM lane quarter: 0 heuristic (2 for self-attention over >= 512 keys), 1 | 2 forced
  int tattn_generic = 0;    // force the generic-T temporal attention kernel for T == 16
  int gn_fuse = 1;          // GroupNorm statistics from the producing GEMM's epilogue
  int gn_silu = 3;          // SiLU of the GroupNorm apply: 3 = h + h tanh(h) through MUFU.TANH (one MUFU op), 1 = ex2 + rcp (two), 2 = ex2 + Newton reciprocal on the FMA pipe
  int gn_small = 1;         // one-kernel GroupNorm (statistics + apply from shared memory) for samples that fit
  int ln_fuse = 1;          // LayerNorm statistics as per-chunk partial sums from the producing GEMM's epilogue (pair kernel)
  int gn_fold = 1;          // GroupNorm (no activation) folded into per-sample weights of the consuming Linear
  int last_gemm_path = 0;   // written by tapgemm(): 2 = tc2<1>, 3 = tc2<2>, 4 = tc3; | (epi + 1) << 8 | gn fused << 16 | residual by MMA << 17 | ln partials << 18 | groups << 20
};
Knobs& knobs();

// ---------------------------------------------------------------- per-launch profiler (bench.py's roofline leg)
// When enabled every kernel family of the path is bracketed by a CUDA-event pair on its launching stream (forwards then
// run eagerly, not as a graph replay); the report gives launches, device time, algorithmic FLOPs and algorithmic HBM
// bytes per (family, shape).  Off (the default) it costs one branch per launch.
enum ProfFam {
  PF_GEMM = 0, PF_FLASH_SELF, PF_FLASH_CROSS, PF_TATTN, PF_GN, PF_LN_STATS, PF_LAYERNORM, PF_TRANSPOSE_V, PF_CONCAT,
  PF_RESAMPLE, PF_LAYOUT, PF_EMBED, PF_SOFTMAX, PF_GENERIC_CONV, PF_SAMPLER, PF_POST, PF_COUNT
};
void prof_enable(bool on);
bool prof_active();
void prof_pause(bool paused);  // planning passes walk the graph without launching: nothing may be booked
std::string prof_report();     // CSV: family,shape,launches,ms,flops,bytes  (synchronises the device, clears the records)
struct ProfScope {
  ProfScope(int fam, double flops, double bytes, cudaStream_t st, const char* shape = "");
  ~ProfScope();
  int idx = -1;
  cudaStream_t st = nullptr;
};

}  // namespace mudg
