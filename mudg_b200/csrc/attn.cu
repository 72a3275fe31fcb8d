// Attention kernels of the UNet (head dim 64 everywhere on the path: num_head_channels 64).
//
// flash2_kernel: spatial self-attention (up to 9216 tokens / frame) and text + image cross-attention on tcgen05
//   (one CTA per SM, 256 queries = two 128-row tiles of one (frame, head); see the kernel's own comment).
//   Two K/V segments keep separate softmax statistics and are summed at the end (attention.py:129-142).
// temporal_attn16_kernel / temporal_attn_kernel: T <= 64 tokens per (pixel, head): HBM-bound, mma.sync / CUDA cores.
// transpose_v_kernel: V -> V^T once per layer (the P V MMA wants a K-major B operand).
// The CUDA-core checker and the tcgen05 issue-rate probes live in csrc/test/testhooks.cu (tests only).
#include "ops.h"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace mudg {

namespace {

constexpr int TILE = 128 * 64 * 2;                       // 16 KB: 128 rows x 64 fp16

struct FaParams {
  long long* trace;   // debug (tests/gpu_trace_flash.py): clock64 time line of CTA 0, [3 roles][96 blocks][8 events]
  int stagger;        // flash2: offset the two softmax groups by half a period (MUDG_FLASH_STAGGER=0 disables)
  int nseg;
  int len[2];
  int kv_div[2];
  float scale_log2;   // scale * log2(e)
};

// ================================================================ flash attention v2 (FA4-style schedule)
// One CTA per SM, 256 queries (two 128-row tiles) of one (frame, head):
//   warp 0      TMA producer: Q0,Q1 once; K/V 128-token blocks through a 3-stage ring (shared by both tiles)
//   warp 1      tcgen05 issuer: per kv block S_i = Q_i K^T (N128 K64) and O_i += P_i V (N64 K128): P is read from
//               TMEM (A-in-TMEM MMA), V MN-major from its row-major TMA tile -- P never touches shared memory
//   warps 2-5   softmax group 0 (tile 0);  warps 6-9 softmax group 1 (tile 1): thread == row == TMEM lane.
//               While group 0 works on S_0 the tensor core runs the other tile's GEMMs (ping-pong).
//   The S row is pulled into registers in one TMEM round trip (the issuer then already computes the next S),
//   P = 2^(s - m) is written back to TMEM as packed fp16, O stays in TMEM across kv blocks; the running max is only
//   raised when it grows by > 2^8 (lazy rescale: P <= 256 fits fp16), in which case the warp rescales its O rows
//   in place (tcgen05.ld/st); the row sum l is a register.
constexpr int F2_THREADS = 320;                     // SPLIT 1; SPLIT 2 runs 64 + 2 x 256 = 576 threads
constexpr int F2_KV_STAGES = 5;
// Q0,Q1 (reused as the two output staging tiles once every S MMA of the tile has retired) + the K / V^T ring
constexpr int F2_XCH = 2 * 2 * 2 * 128 * 4;         // SPLIT 2: row max / row sum exchange [parity][group][half][row] fp32
constexpr int F2_SMEM = 2 * TILE + F2_KV_STAGES * 2 * TILE + 1024 + 256 + F2_XCH;
constexpr float F2_LAZY = 8.0f;

#define F2_TRACE(role, blk, ev)                                                                        \
  do {                                                                                                  \
    if (p.trace != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (blk) < 96)                \
      p.trace[((role) * 96 + (blk)) * 8 + (ev)] = clock64();                                            \
  } while (0)

// SPLIT = softmax warps per TMEM lane quarter and group.  SPLIT 2 (long self-attention): a row's 128 scores of a block are
// halved between two warps (they may both touch the quarter's 32 TMEM lanes), 8 warps per group, 576 threads, so FOUR
// softmax warps sit on every SM sub-partition instead of two.  Why: the MUFU unit needs 2-4 resident issuing warps per
// sub-partition to reach its 16 ex2/clk/SM (tests/gpu_probe_mufu.py: 1 warp 10.5, 2 warps 14.4, 4 warps 16.0) and the
// exponent phase of SPLIT 1 runs ONE warp per sub-partition at 10.8 clk per MUFU op.  The halves agree on the row max
// through shared memory (one group barrier per block), keep partial row sums (added once per segment), and each rescales /
// reads out its 32 of the 64 O columns.
template <int NSEG, int F2_POLY, int SPLIT>
__global__ void __launch_bounds__(64 + SPLIT * 256, 1)
flash2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmO,
              const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
              const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // 2 tiles
  uint8_t* sKV = smem + 2 * TILE;                      // stage s: K at + s*2*TILE, V at + TILE
  uint8_t* sOut = sQ;                                  // tile i: 16 KB output staging = its own (dead) Q tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + F2_KV_STAGES * 2 * TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                        // [F2_KV_STAGES]
  uint64_t* kv_empty = kv_full + F2_KV_STAGES;         // [F2_KV_STAGES]
  uint64_t* s_full = kv_empty + F2_KV_STAGES;          // [2]
  uint64_t* p_ready = s_full + 2;                      // [2]
  uint64_t* o_done = p_ready + 2;                      // [2]
  uint64_t* s_free = o_done + 2;                       // [2]  S tile copied to registers -> TMEM tile reusable
  uint64_t* exp_done = s_free + 2;                     // [2]  group i has finished the exponentials of a block (MUFU turn)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(exp_done + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2 parity][2 groups][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, head = blockIdx.y, f = blockIdx.z;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < F2_KV_STAGES; i++) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128 * SPLIT); mbar_init(&o_done[i], 1); mbar_init(&s_free[i], 128 * SPLIT);
    }
    mbar_init(&exp_done[0], 128 * SPLIT); mbar_init(&exp_done[1], 128 * SPLIT);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) | O0 [256,320) O1 [320,384) | P0 [384,448) P1 [448,512) (fp16 pairs)

  int nblk[2];
  nblk[0] = (p.len[0] + 127) >> 7;
  nblk[1] = NSEG > 1 ? (p.len[1] + 127) >> 7 : 0;
  const int nb = nblk[0] + nblk[1];
  // The two softmax groups take turns on the MUFU unit: g0 exp(0), g1 exp(0), g0 exp(1), g1 exp(1), ...  Left alone they
  // start in phase and stay there (or drift), so the unit idles whenever BOTH are in their TMEM-load / max / P-store
  // phase and is oversubscribed when both exponentiate (ncu: XU pipe 61 % busy, clock64 trace: 2500 clk per exp phase
  // instead of ~1100).  With the hand-over each group does its other work while its partner exponentiates.
  const bool do_stagger = p.stagger != 0 && nb >= 4;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE);
      tma_load_5d(sQ, &tmQ, q_full, head * 64, q0, f, 0, 0);
      tma_load_5d(sQ + TILE, &tmQ, q_full, head * 64, q0 + 128, f, 0, 0);
      int it = 0;
      for (int sg = 0; sg < NSEG; sg++) {
        const CUtensorMap* mk = sg ? &tmK1 : &tmK0;
        const CUtensorMap* mv = sg ? &tmV1 : &tmV0;
        const int kb = f / p.kv_div[sg];
        for (int j = 0; j < nblk[sg]; j++, it++) {
          const int s = it % F2_KV_STAGES;
          mbar_wait(&kv_empty[s], ((it / F2_KV_STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], 2 * TILE);
          uint8_t* k_s = sKV + s * 2 * TILE;
          tma_load_5d(k_s, mk, &kv_full[s], head * 64, j * 128, kb, 0, 0);
          // V arrives TRANSPOSED ([d][kv], kv contiguous): two K-major SWIZZLE_128B atoms of 64 kv columns each
          tma_load_5d(k_s + TILE, mv, &kv_full[s], j * 128, head * 64, kb, 0, 0);
          tma_load_5d(k_s + TILE + TILE / 2, mv, &kv_full[s], j * 128 + 64, head * 64, kb, 0, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // The WHOLE warp runs the issue loop (uniform control flow, descriptors in uniform registers) and one elected lane
    // issues: the same code under `if (lane == 0)` spent ~140 clocks of dependent R2UR / ELECT / waterfall instructions
    // per tcgen05.mma (clock64 trace), which made the issue thread -- not the tensor pipe or the MUFU -- the bottleneck.
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0);
    // P V with V^T K-major in shared memory: an MN-major B operand (V as it lies in the QKV rows) feeds the tensor core
    // at a quarter of the rate (tests/gpu_probe_mma.py: 118 clk per N=64 step instead of 33)
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 0);
    const uint64_t dq0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t dq1 = umma_desc_sw128(smem_u32(sQ + TILE), 16, 1024);
    const uint64_t dkv0 = umma_desc_sw128(smem_u32(sKV), 16, 1024);
    mbar_wait(q_full, 0);
    // Event loop: each softmax group is served on its own -- S_i(b) as soon as K(b) has landed and group i holds S_i(b-1)
    // in registers, P_i V(b) as soon as group i has stored P_i(b).  Nothing here makes one group wait for the other, so
    // the half-period offset set up at the start (see `do_stagger`) survives: one group exponentiates (MUFU) while the
    // other loads / reduces / stores.  Neither event is latency critical (both results are needed a whole block later),
    // so an idle round sleeps instead of stealing issue slots from the softmax warps on this SM sub-partition.
    const uint32_t tD[2] = {tmem_base, tmem_base + 128};
    const uint32_t tOo[2] = {tmem_base + 256, tmem_base + 320};
    const uint32_t tPp[2] = {tmem_base + 384, tmem_base + 448};
    const uint64_t dq[2] = {dq0, dq1};
    int nS[2] = {0, 0}, nP[2] = {0, 0};            // next block whose S / P V is to be issued, per group
    uint32_t sS[2] = {0, 0}, phS[2] = {0, 0};      // kv ring stage / kv_full parity of block nS[i]
    uint32_t sP[2] = {0, 0};                       // kv ring stage of block nP[i]
    int seg_first = 0;
    while (nP[0] < nb || nP[1] < nb) {
      bool progress = false;
#pragma unroll
      for (int i = 0; i < 2; i++) {
        if (nS[i] < nb) {
          bool ok = mbar_test(&kv_full[sS[i]], phS[i]);
          if (nS[i] > 0) ok = ok && mbar_test(&s_free[i], (uint32_t)(nS[i] - 1) & 1u);
          if (__all_sync(0xffffffffu, ok)) {
            tc_fence_after();
            if (elect_one()) {
              const uint64_t dk = dkv0 + (uint64_t)(sS[i] * ((2 * TILE) >> 4));
#pragma unroll
              for (int k = 0; k < 4; k++) umma_f16(tD[i], dq[i] + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
              umma_commit(&s_full[i]);
            }
            __syncwarp();
            if (lane == 0) F2_TRACE(2, nS[i], i);                 // S_i(b) issued
            nS[i]++;
            if (++sS[i] == F2_KV_STAGES) { sS[i] = 0; phS[i] ^= 1u; }
            progress = true;
          }
        }
        if (nP[i] < nb) {
          const bool ok = mbar_test(&p_ready[i], (uint32_t)nP[i] & 1u);
          if (__all_sync(0xffffffffu, ok)) {
            tc_fence_after();
            if (NSEG > 1) seg_first = nP[i] >= nblk[0] ? nblk[0] : 0;
            if (elect_one()) {
              const uint64_t dv = dkv0 + (uint64_t)(sP[i] * ((2 * TILE) >> 4) + (TILE >> 4));
              const uint32_t acc0 = nP[i] != seg_first ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < 8; k++) {   // K = 16 kv tokens = 8 TMEM columns of P = 32 B inside a V^T atom (4 per atom)
                const uint64_t dvk = dv + (uint64_t)((k >> 2) * ((TILE / 2) >> 4) + (k & 3) * 2);
                umma_f16_ts(tOo[i], tPp[i] + k * 8, dvk, idesc_o, (k != 0) ? 1u : acc0);
              }
              umma_commit(&o_done[i]);
              if (nP[i ^ 1] > nP[i]) umma_commit(&kv_empty[sP[i]]);   // the other tile's P V of this block is already in
            }
            __syncwarp();
            if (lane == 0) F2_TRACE(2, nP[i], 2 + i);             // P_i V(b) issued
            nP[i]++;
            if (++sP[i] == F2_KV_STAGES) sP[i] = 0;
            progress = true;
          }
        }
      }
      if (!progress) __nanosleep(64);
    }
  } else if (SPLIT == 2) {
    // ---------------- SPLIT 2: 8 warps per group; warp (q, half) owns columns [64 half, 64 half + 64) of rows 32 q .. 32 q + 31
    const int grp = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + grp * 128 + half * 64 + lane_off;
    const uint32_t tO = tmem_base + 256 + grp * 64 + half * 32 + lane_off;
    const uint32_t tP = tmem_base + 384 + grp * 64 + half * 32 + lane_off;
    uint8_t* myOut = sOut + grp * TILE;
    auto group_sync = [&]() {                        // the 256 threads of this group
      if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    auto xslot = [&](int parity, int h) { return xch + ((parity * 2 + grp) * 2 + h) * 128 + row; };
    float m_used = -INFINITY, l_run = 0.f;
    const int nb1 = nblk[0];
#pragma unroll 1
    for (int it = 0; it < nb1; it++) {
      const int valid = p.len[0] - it * 128 - half * 64;      // my columns >= valid are padding (TMA zero fill)
      const bool tracer = q == 2 && half == 0 && lane == 0;
      if (tracer) F2_TRACE(grp, it, 0);                      // ready for S(it)
      mbar_wait(&s_full[grp], it & 1);
      __syncwarp();
      tc_fence_after();
      if (tracer) F2_TRACE(grp, it, 1);                      // S(it) available
      uint32_t sr[2][32];
      tmem_ld32(tS, sr[0]);
      tmem_ld32(tS + 32, sr[1]);
      tmem_ld_wait();
      if (tracer) F2_TRACE(grp, it, 2);                      // S(it) in registers
      tc_fence_before();
      mbar_arrive(&s_free[grp]);
      float mx = -INFINITY;
      if (valid >= 64) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(sr[c][i]));
      }
      // the two halves of a row agree on the block's row max (slots double-buffered by block parity)
      *xslot(it & 1, half) = mx;
      group_sync();
      mx = fmaxf(mx, *xslot(it & 1, half ^ 1));
      if (do_stagger) {
        if (grp == 1) mbar_wait(&exp_done[0], (uint32_t)it & 1u);
        else if (it > 0) mbar_wait(&exp_done[1], (uint32_t)(it - 1) & 1u);
        __syncwarp();
      }
      if (tracer) F2_TRACE(grp, it, 5);                      // row max agreed, MUFU turn acquired
      mx *= p.scale_log2;
      const bool grow = (it > 0) && (mx > m_used + F2_LAZY);
      const float m_old = m_used;
      if (it == 0 || grow) m_used = mx;
      float lsum = 0.f;
      if (valid >= 64) {
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = fmaf(__uint_as_float(sr[c][i]), p.scale_log2, -m_used);
            const float a1 = fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, -m_used);
            const float p0 = (F2_POLY == 3 && (i & 2)) ? ex2_poly(a0) : ex2_approx(a0);
            const float p1 = ((F2_POLY == 1 && (i & 2)) || F2_POLY >= 2) ? ex2_poly(a1) : ex2_approx(a1);
            l4[(i >> 1) & 3] += p0 + p1;
            sr[c][i >> 1] = pack_half2(p0, p1);
            if (i == 30 && c == 0 && do_stagger && p.stagger >= 2 && p.stagger <= 4) mbar_arrive_after(&exp_done[grp], sr[c][15]);
          }
        lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < valid) ? ex2_approx(fmaf(__uint_as_float(sr[c][i]), p.scale_log2, -m_used)) : 0.f;
            const float p1 = (c * 32 + i + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, -m_used)) : 0.f;
            lsum += p0 + p1;
            sr[c][i >> 1] = pack_half2(p0, p1);
          }
      }
      if (do_stagger && (p.stagger == 1 || p.stagger > 4 || valid < 64)) mbar_arrive(&exp_done[grp]);
      if (tracer) F2_TRACE(grp, it, 3);                      // exponentials done
      if (it > 0) {                                  // the previous P V of this tile has retired: P and O may be touched
        mbar_wait(&o_done[grp], (it - 1) & 1);
        __syncwarp();
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {         // both halves see the same row maxima, hence the same decision
          const float factor = ex2_approx(m_old - m_used);
          l_run *= factor;
          uint32_t v[32];
          tmem_ld32(tO, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
          tmem_st32(tO, v);
        }
      }
      l_run += lsum;
      {
        uint32_t w[32];                              // my 64 keys = 32 packed words = P columns [32 half, 32 half + 32)
#pragma unroll
        for (int i = 0; i < 16; i++) { w[i] = sr[0][i]; w[16 + i] = sr[1][i]; }
        tmem_st32(tP, w);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[grp]);
      if (tracer) F2_TRACE(grp, it, 4);                      // P(it) stored
    }
    // ---- O / l: the halves add their partial row sums, each reads out its 32 of the 64 O columns
    mbar_wait(&o_done[grp], (nb1 - 1) & 1);
    __syncwarp();
    tc_fence_after();
    *xslot(nb1 & 1, half) = l_run;
    group_sync();
    const float inv = 1.f / (l_run + *xslot(nb1 & 1, half ^ 1));
    uint32_t v[32];
    tmem_ld32(tO, v);
    tmem_ld_wait();
    tc_fence_before();
    uint8_t* stg = myOut + row * 128;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      const uint4 val = make_uint4(pack_half2(__uint_as_float(v[8 * jj]) * inv, __uint_as_float(v[8 * jj + 1]) * inv),
                                   pack_half2(__uint_as_float(v[8 * jj + 2]) * inv, __uint_as_float(v[8 * jj + 3]) * inv),
                                   pack_half2(__uint_as_float(v[8 * jj + 4]) * inv, __uint_as_float(v[8 * jj + 5]) * inv),
                                   pack_half2(__uint_as_float(v[8 * jj + 6]) * inv, __uint_as_float(v[8 * jj + 7]) * inv));
      *reinterpret_cast<uint4*>(stg + (((half * 4 + jj) ^ (row & 7)) << 4)) = val;
    }
    fence_proxy_async_smem();
    group_sync();
    if (((warp - 2) & 7) == 0 && lane == 0) {        // first warp of each group
      tma_store_5d(&tmO, myOut, head * 64, q0 + grp * 128, f, 0, 0);
      tma_store_commit();
      tma_store_wait_read0();
    }
    __syncwarp();
    tc_fence_before();
  } else {
    const int grp = (warp - 2) >> 2;                 // query tile handled by this softmax group
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + grp * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + grp * 64 + lane_off;
    const uint32_t tP = tmem_base + 384 + grp * 64 + lane_off;
    uint8_t* myOut = sOut + grp * TILE;
    uint32_t acc[NSEG > 1 ? 32 : 1];
    if (NSEG > 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) acc[i] = 0u;
    }
    float o_final[64];
    int it = 0;
#pragma unroll 1
    for (int sg = 0; sg < NSEG; sg++) {
      float m_used = -INFINITY, l_run = 0.f;
#pragma unroll 1
      for (int j = 0; j < nblk[sg]; j++, it++) {
        const int valid = p.len[sg] - j * 128;
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 0);     // ready for S(it)
        mbar_wait(&s_full[grp], it & 1);
        __syncwarp();
        tc_fence_after();
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 1);     // S(it) available
        // ---- the whole 128-column S row goes to registers in one TMEM round trip; the TMEM tile is then free
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; c++) tmem_ld32(tS + c * 32, sr[c]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[grp]);
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 2);     // S(it) in registers
        float mx = -INFINITY;
        if (valid >= 128) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains: the max is latency-bound
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              m4[(i >> 1) & 3] = fmaxf(m4[(i >> 1) & 3], fmaxf(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])));
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(sr[c][i]));
        }
        if (do_stagger) {
          if (grp == 1) mbar_wait(&exp_done[0], (uint32_t)it & 1u);
          else if (it > 0) mbar_wait(&exp_done[1], (uint32_t)(it - 1) & 1u);
          __syncwarp();
        }
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 5);     // MUFU turn acquired
        mx *= p.scale_log2;
        const bool grow = (j > 0) && (mx > m_used + F2_LAZY);
        const float m_old = m_used;
        if (j == 0 || grow) m_used = mx;
        // ---- P = 2^(s*scale - m_used): fp32 exp2, row sum in a register, packed to fp16 in place over the S registers
        float lsum = 0.f;
        if (valid >= 128) {
          float l4[4] = {0.f, 0.f, 0.f, 0.f};                            // independent chains for the row sum too
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              // (optional) every 4th exponential on the FMA pipe (polynomial), the rest on the MUFU unit
              const float a1 = fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, -m_used);
              // F2_POLY: 0 all on the MUFU, 1 every 4th / 2 every 2nd / 3 three of four exponentials on the FMA pipe
              const float a0 = fmaf(__uint_as_float(sr[c][i]), p.scale_log2, -m_used);
              const float p0 = (F2_POLY == 3 && (i & 2)) ? ex2_poly(a0) : ex2_approx(a0);
              const float p1 = ((F2_POLY == 1 && (i & 2)) || F2_POLY >= 2) ? ex2_poly(a1) : ex2_approx(a1);
              l4[(i >> 1) & 3] += p0 + p1;
              sr[c][i >> 1] = pack_half2(p0, p1);
              // early hand-over of the MUFU turn: the partner group may start its exponentials when 3/4 (or 1/2) of
              // mine are done, so the unit has a second warp to pick from during my (single-warp, latency-limited) tail
              if (i == 30 && do_stagger && ((c == 2 && p.stagger == 2) || (c == 1 && p.stagger == 3) || (c == 0 && p.stagger == 4)))
                mbar_arrive_after(&exp_done[grp], sr[c][15]);
            }
          lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = (c * 32 + i < valid) ? ex2_approx(fmaf(__uint_as_float(sr[c][i]), p.scale_log2, -m_used)) : 0.f;
              const float p1 = (c * 32 + i + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, -m_used)) : 0.f;
              lsum += p0 + p1;
              sr[c][i >> 1] = pack_half2(p0, p1);
            }
        }
        if (do_stagger && (p.stagger == 1 || p.stagger > 4 || valid < 128)) mbar_arrive(&exp_done[grp]);
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 3);     // exponentials done
        // the previous P V of this tile must have retired before P (and possibly O) are overwritten
        if (j > 0) {
          mbar_wait(&o_done[grp], (it - 1) & 1);
          __syncwarp();
          tc_fence_after();
          if (__any_sync(0xffffffffu, grow)) {       // warp-uniform: tcgen05.ld/st are warp-collective
            const float factor = ex2_approx(m_old - m_used);   // 1 for the lanes whose max did not move
            l_run *= factor;
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
              uint32_t v[32];
              tmem_ld32(tO + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * factor);
              tmem_st32(tO + c * 32, v);
            }
          }
        }
        l_run += lsum;
        // P -> TMEM: 128 fp16 = 64 columns; chunk c of the S row became 16 packed words sr[c][0..15]
        {
          uint32_t w[32];
#pragma unroll
          for (int i = 0; i < 16; i++) { w[i] = sr[0][i]; w[16 + i] = sr[1][i]; }
          tmem_st32(tP, w);
#pragma unroll
          for (int i = 0; i < 16; i++) { w[i] = sr[2][i]; w[16 + i] = sr[3][i]; }
          tmem_st32(tP + 32, w);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[grp]);
        if ((threadIdx.x & 127) == 64) F2_TRACE(grp, it, 4);     // P(it) stored
      }
      // ---- segment done: O / l
      mbar_wait(&o_done[grp], (it - 1) & 1);
      __syncwarp();
      tc_fence_after();
      const float inv = 1.f / l_run;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t v[32];
        tmem_ld32(tO + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) o_final[c * 32 + i] = __uint_as_float(v[i]) * inv;
      }
      if (NSEG > 1) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float2 a = unpack_half2(acc[i]);
          acc[i] = pack_half2(a.x + o_final[2 * i], a.y + o_final[2 * i + 1]);
        }
      }
      tc_fence_before();
    }
    // ---- epilogue: fp16 tile -> swizzled staging smem -> TMA store
    uint8_t* stg = myOut + row * 128;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      uint4 val;
      if (NSEG > 1) val = make_uint4(acc[4 * jj], acc[4 * jj + 1], acc[4 * jj + 2], acc[4 * jj + 3]);
      else val = make_uint4(pack_half2(o_final[8 * jj], o_final[8 * jj + 1]), pack_half2(o_final[8 * jj + 2], o_final[8 * jj + 3]),
                            pack_half2(o_final[8 * jj + 4], o_final[8 * jj + 5]), pack_half2(o_final[8 * jj + 6], o_final[8 * jj + 7]));
      *reinterpret_cast<uint4*>(stg + ((jj ^ (row & 7)) << 4)) = val;
    }
    fence_proxy_async_smem();
    if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
    if (q == 2 && lane == 0) {       // warps 2 and 6 (first warp of each group)
      tma_store_5d(&tmO, myOut, head * 64, q0 + grp * 128, f, 0, 0);
      tma_store_commit();
      tma_store_wait_read0();
    }
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------- temporal attention
// qkv rows are ordered (b, t, pixel); a warp handles `pairs` (pixel, head) problems at once: lane -> (pair, t)
// for T <= 32, and lane -> t, t+32 for T in (32, 64].  K/V of the warp's pairs are staged in shared memory.
__global__ void temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int T, int HW,
                                     int heads, float scale_log2, int group, int pairs_per_warp) {
  extern __shared__ uint4 tsm[];
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int inner = heads * 64;
  const int pitch = 3 * inner;
  const int64_t npairs = (int64_t)B * HW * heads;
  const int64_t first = ((int64_t)blockIdx.x * warps_per_block + warp_in_block) * pairs_per_warp;
  // smem per warp: pairs * T rows * (8 K chunks + 8 V chunks)
  uint4* wsm = tsm + (size_t)warp_in_block * pairs_per_warp * T * 16;

  for (int i = lane; i < pairs_per_warp * T * 16; i += 32) {
    const int ch = i & 7;
    const int kv = (i >> 3) & 1;
    const int t = (i >> 4) % T;
    const int pp = (i >> 4) / T;
    const int64_t pr = first + pp;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (pr < npairs) {
      const int head = pr % heads;
      const int64_t bp = pr / heads;
      const int px = bp % HW;
      const int b = bp / HW;
      const __half* src = qkv + (((int64_t)b * T + t) * HW + px) * pitch + (1 + kv) * inner + head * 64;
      val = __ldg(reinterpret_cast<const uint4*>(src) + ch);
    }
    wsm[((pp * T + t) * 2 + kv) * 8 + ch] = val;
  }
  __syncwarp();

  const int pp = (group < 32) ? lane / group : 0;
  const int tl = (group < 32) ? lane % group : lane;
  const int64_t pr = first + pp;
  if (pp >= pairs_per_warp || pr >= npairs) return;
  const int head = pr % heads;
  const int64_t bp = pr / heads;
  const int px = bp % HW;
  const int b = bp / HW;
  for (int t = tl; t < T; t += 32) {
    if (group < 32 && t != tl) break;
    const int64_t rowi = ((int64_t)b * T + t) * HW + px;
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + rowi * pitch + head * 64);
    __half2 qh[32];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint4 v = __ldg(qp + c);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      qh[4 * c] = h[0]; qh[4 * c + 1] = h[1]; qh[4 * c + 2] = h[2]; qh[4 * c + 3] = h[3];
    }
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; i++) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < T; j++) {
      const uint4* kr = wsm + ((pp * T + j) * 2 + 0) * 8;
      const uint4* vr = wsm + ((pp * T + j) * 2 + 1) * 8;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const uint4 kv4 = kr[c];
        const __half2* kh = reinterpret_cast<const __half2*>(&kv4);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 kf = __half22float2(kh[e]);
          const float2 qf = __half22float2(qh[4 * c + e]);
          d += qf.x * kf.x + qf.y * kf.y;
        }
      }
      d *= scale_log2;
      const float mn = fmaxf(m, d);
      const float al = exp2f(m - mn), pj = exp2f(d - mn);
      l = l * al + pj;
      m = mn;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const uint4 vv4 = vr[c];
        const __half2* vh = reinterpret_cast<const __half2*>(&vv4);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float2 vf = __half22float2(vh[e]);
          o[8 * c + 2 * e] = o[8 * c + 2 * e] * al + pj * vf.x;
          o[8 * c + 2 * e + 1] = o[8 * c + 2 * e + 1] * al + pj * vf.y;
        }
      }
    }
    const float inv = 1.f / l;
    uint4* op = reinterpret_cast<uint4*>(out + rowi * inner + head * 64);
#pragma unroll
    for (int c = 0; c < 8; c++)
      op[c] = make_uint4(pack_half2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_half2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                         pack_half2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_half2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
  }
}

// ---------------------------------------------------------------- temporal attention, T == 16: tensor-core path
// One warp per (b, pixel, head) problem: S = Q K^T (16x16x64) and O = P V (16x64x16) are 16 mma.sync.m16n8k16 in
// total, so the kernel is purely HBM-bound.  Q/K/V tiles (16 tokens x 128 B, gathered with stride H*W*3C) are
// brought in with cp.async into a 2-deep per-warp ring of XOR-swizzled smem tiles and read back with ldmatrix
// (.trans for V); the softmax lives in the accumulator fragments (quad shuffles); P is re-packed as the A operand.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int TA_WARPS = 4;
constexpr int TA_TILE = 16 * 128;                 // one 16-token x 64-dim fp16 tile
constexpr int TA_SMEM = TA_WARPS * 2 * 3 * TA_TILE;   // 48 KB

__global__ void __launch_bounds__(TA_WARPS * 32)
temporal_attn16_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int HW, int heads, float scale_log2) {
  extern __shared__ __align__(128) uint8_t ta_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = heads * 64;
  const int64_t pitch = 3 * (int64_t)inner;
  const int64_t npairs = (int64_t)B * HW * heads;
  const int64_t wid = (int64_t)blockIdx.x * TA_WARPS + warp;
  const int64_t wstride = (int64_t)gridDim.x * TA_WARPS;
  uint8_t* wbase = ta_smem + warp * (2 * 3 * TA_TILE);
  const uint32_t wbase_u = smem_u32(wbase);

  // lane -> (token row, 16 B chunk) pairs for the gathers: 16 rows x 8 chunks = 128 pieces per tile, 4 per lane
  auto issue_loads = [&](int64_t pr, int buf) {
    const int head = pr % heads;
    const int64_t bp = pr / heads;
    const int px = bp % HW;
    const int b = bp / HW;
    const __half* base = qkv + ((int64_t)b * 16 * HW + px) * pitch + head * 64;
    const uint32_t sb = wbase_u + buf * (3 * TA_TILE);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int piece = lane + 32 * i;
      const int r = piece >> 3, c = piece & 7;
      const __half* src = base + (int64_t)r * HW * pitch + c * 8;
      const uint32_t dst = sb + r * 128 + ((c ^ (r & 7)) << 4);
      cp_async16(dst, src);                                  // Q
      cp_async16(dst + TA_TILE, src + inner);                // K
      cp_async16(dst + 2 * TA_TILE, src + 2 * inner);        // V
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (wid < npairs) issue_loads(wid, 0);
  for (int64_t pr = wid; pr < npairs; pr += wstride, buf ^= 1) {
    const int64_t nxt = pr + wstride;
    if (nxt < npairs) {
      issue_loads(nxt, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const uint32_t sq = wbase_u + buf * (3 * TA_TILE), sk = sq + TA_TILE, sv = sq + 2 * TA_TILE;
    // ---- S = Q K^T
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
      {
        const int r = (lane & 7) + 8 * ((lane >> 3) & 1), c = 2 * ks + (lane >> 4);
        ldsm_x4(sq + r * 128 + ((c ^ (r & 7)) << 4), a0, a1, a2, a3);
      }
      {
        const int r = (lane & 7) + 8 * (lane >> 4), c = 2 * ks + ((lane >> 3) & 1);
        ldsm_x4(sk + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
      }
      mma16816(s0, a0, a1, a2, a3, b0, b1);     // keys 0..7
      mma16816(s1, a0, a1, a2, a3, b2, b3);     // keys 8..15
    }
    // ---- softmax over the 16 keys of rows g (c0,c1) and g+8 (c2,c3); a row lives in one quad of lanes
    float mA = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
    float mB = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1)); mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1)); mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
    const float oA = mA * scale_log2, oB = mB * scale_log2;
    const float p00 = exp2f(fmaf(s0[0], scale_log2, -oA)), p01 = exp2f(fmaf(s0[1], scale_log2, -oA));
    const float p10 = exp2f(fmaf(s1[0], scale_log2, -oA)), p11 = exp2f(fmaf(s1[1], scale_log2, -oA));
    const float p02 = exp2f(fmaf(s0[2], scale_log2, -oB)), p03 = exp2f(fmaf(s0[3], scale_log2, -oB));
    const float p12 = exp2f(fmaf(s1[2], scale_log2, -oB)), p13 = exp2f(fmaf(s1[3], scale_log2, -oB));
    float lA = p00 + p01 + p10 + p11, lB = p02 + p03 + p12 + p13;
    lA += __shfl_xor_sync(0xffffffffu, lA, 1); lA += __shfl_xor_sync(0xffffffffu, lA, 2);
    lB += __shfl_xor_sync(0xffffffffu, lB, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    const uint32_t pa0 = pack_half2(p00, p01), pa1 = pack_half2(p02, p03), pa2 = pack_half2(p10, p11),
                   pa3 = pack_half2(p12, p13);
    // ---- O = P V
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int nj = 0; nj < 8; nj += 2) {
      uint32_t b0, b1, b2, b3;
      const int r = (lane & 7) + 8 * ((lane >> 3) & 1), c = nj + (lane >> 4);
      ldsm_x4_t(sv + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
      mma16816(o[nj], pa0, pa1, pa2, pa3, b0, b1);
      mma16816(o[nj + 1], pa0, pa1, pa2, pa3, b2, b3);
    }
    const float iA = 1.f / lA, iB = 1.f / lB;
    // ---- stage O (fp16) in the Q tile, then 128-bit coalesced stores
    __syncwarp();
    {
      const int g = lane >> 2, t = lane & 3;
      uint8_t* so = wbase + buf * (3 * TA_TILE);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        *reinterpret_cast<uint32_t*>(so + g * 128 + ((j ^ (g & 7)) << 4) + t * 4) = pack_half2(o[j][0] * iA, o[j][1] * iA);
        *reinterpret_cast<uint32_t*>(so + (g + 8) * 128 + ((j ^ ((g + 8) & 7)) << 4) + t * 4) = pack_half2(o[j][2] * iB, o[j][3] * iB);
      }
    }
    __syncwarp();
    {
      const int head = pr % heads;
      const int64_t bp = pr / heads;
      const int px = bp % HW;
      const int b = bp / HW;
      __half* ob = out + ((int64_t)b * 16 * HW + px) * inner + head * 64;
      const uint8_t* so = wbase + buf * (3 * TA_TILE);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int piece = lane + 32 * i;
        const int r = piece >> 3, c = piece & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(so + r * 128 + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(ob + (int64_t)r * HW * inner + c * 8) = val;
      }
    }
    __syncwarp();     // the tile is reused by the prefetch issued two iterations from now
  }
}

// ---------------------------------------------------------------- temporal attention, T == 32 | 64: the same tensor-core scheme
// One warp (T = 32) or two (T = 64) per (b, pixel, head) problem; Q / K / V tiles of T tokens x 128 B in a 2-deep cp.async
// ring.  Per 16-query block:
// S = Q K^T (T / 8 n-tiles x 4 k-steps), softmax in the accumulator fragments, O = P V (8 n-tiles x T / 16 k-steps), O staged
// over the block's own Q rows (dead by then) and stored with 128-bit coalesced writes.  Algorithmic traffic is the same
// 4 x T x 128 B per problem as T == 16, so the kernel is HBM-bound (T = 64: 1 MFLOP and 4 096 exponentials per 32 KB).
template <int T>
struct TaCfg {
  static constexpr int TILE = T * 128;
  static constexpr int PROBS = 4;                       // problems in flight per CTA (each with a 2-deep tile ring)
  static constexpr int WPP = T >= 64 ? 2 : 1;           // warps per problem: T = 64 has 4 query blocks, two per warp (one warp
                                                        // alone ran its 1 000-instruction dependent stream at IPC 0.11: 4.0 TB/s)
  static constexpr int WARPS = PROBS * WPP;
  static constexpr int SMEM = PROBS * 2 * 3 * TILE;     // T = 64: 192 KB (one CTA per SM), T = 32: 96 KB
};

template <int T>
__global__ void __launch_bounds__(TaCfg<T>::WARPS * 32)
temporal_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int HW, int heads, float scale_log2) {
  extern __shared__ __align__(128) uint8_t ta_smem[];
  constexpr int TILE = TaCfg<T>::TILE, NT = T / 8, KT = T / 16, WPP = TaCfg<T>::WPP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPP, sub = warp % WPP;          // problem slot of the CTA, my share of its query blocks
  const int inner = heads * 64;
  const int64_t pitch = 3 * (int64_t)inner;
  const int64_t npairs = (int64_t)B * HW * heads;
  const int64_t wid = (int64_t)blockIdx.x * TaCfg<T>::PROBS + slot;
  const int64_t wstride = (int64_t)gridDim.x * TaCfg<T>::PROBS;
  uint8_t* wbase = ta_smem + slot * (2 * 3 * TILE);
  const uint32_t wbase_u = smem_u32(wbase);
  auto slot_sync = [&]() {                                // the WPP warps of one problem
    if (WPP > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(32 * WPP) : "memory");
    else __syncwarp();
  };

  auto issue_loads = [&](int64_t pr, int buf) {
    const int head = pr % heads;
    const int64_t bp = pr / heads;
    const int px = bp % HW;
    const int b = bp / HW;
    const __half* base = qkv + ((int64_t)b * T * HW + px) * pitch + head * 64;
    const uint32_t sb = wbase_u + buf * (3 * TILE);
#pragma unroll
    for (int i = 0; i < T / 4 / WPP; i++) {               // T rows x 8 chunks of 16 B, shared between the problem's warps
      const int piece = (sub * (T / 4 / WPP) + i) * 32 + lane;
      const int r = piece >> 3, c = piece & 7;
      const __half* src = base + (int64_t)r * HW * pitch + c * 8;
      const uint32_t dst = sb + r * 128 + ((c ^ (r & 7)) << 4);
      cp_async16(dst, src);                                  // Q
      cp_async16(dst + TILE, src + inner);                   // K
      cp_async16(dst + 2 * TILE, src + 2 * inner);           // V
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if (wid < npairs) issue_loads(wid, 0);
  for (int64_t pr = wid; pr < npairs; pr += wstride, buf ^= 1) {
    // my share of this problem's tiles has landed; after the sync so has my partner's, and both of us are past every read
    // of the other buffer (previous iteration), which the prefetch below overwrites
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    slot_sync();
    const int64_t nxt = pr + wstride;
    if (nxt < npairs) issue_loads(nxt, buf ^ 1);
    const uint32_t sq = wbase_u + buf * (3 * TILE), sk = sq + TILE, sv = sq + 2 * TILE;
    uint8_t* so = wbase + buf * (3 * TILE);
    const int head = pr % heads;
    const int64_t bp = pr / heads;
    const int px = bp % HW;
    const int b = bp / HW;
    __half* ob = out + ((int64_t)b * T * HW + px) * inner + head * 64;
#pragma unroll 1
    for (int mt = sub; mt < KT; mt += WPP) {
      // ---- S = Q[mt] K^T : 16 queries x T keys
      float s[NT][4];
#pragma unroll
      for (int j = 0; j < NT; j++) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        uint32_t a0, a1, a2, a3;
        {
          const int r = mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), c = 2 * ks + (lane >> 4);
          ldsm_x4(sq + r * 128 + ((c ^ (r & 7)) << 4), a0, a1, a2, a3);
        }
#pragma unroll
        for (int np = 0; np < KT; np++) {                    // 16 keys per ldmatrix.x4
          uint32_t b0, b1, b2, b3;
          const int r = np * 16 + (lane & 7) + 8 * (lane >> 4), c = 2 * ks + ((lane >> 3) & 1);
          ldsm_x4(sk + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
          mma16816(s[2 * np], a0, a1, a2, a3, b0, b1);
          mma16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
      }
      // ---- softmax over the T keys of rows g (c0, c1) and g + 8 (c2, c3); a row lives in one quad of lanes
      float mA = -INFINITY, mB = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT; j++) {
        mA = fmaxf(mA, fmaxf(s[j][0], s[j][1]));
        mB = fmaxf(mB, fmaxf(s[j][2], s[j][3]));
      }
      mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1)); mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
      mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1)); mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
      const float oA = mA * scale_log2, oB = mB * scale_log2;
      float lA = 0.f, lB = 0.f;
      uint32_t pa[KT][4];
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const float p0 = exp2f(fmaf(s[j][0], scale_log2, -oA)), p1 = exp2f(fmaf(s[j][1], scale_log2, -oA));
        const float p2 = exp2f(fmaf(s[j][2], scale_log2, -oB)), p3 = exp2f(fmaf(s[j][3], scale_log2, -oB));
        lA += p0 + p1;
        lB += p2 + p3;
        pa[j >> 1][(j & 1) * 2] = pack_half2(p0, p1);        // A fragment of the P V k-step j / 2: keys 8 (j & 1) + ...
        pa[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
      }
      lA += __shfl_xor_sync(0xffffffffu, lA, 1); lA += __shfl_xor_sync(0xffffffffu, lA, 2);
      lB += __shfl_xor_sync(0xffffffffu, lB, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
      // ---- O = P V
      float o[8][4];
#pragma unroll
      for (int j = 0; j < 8; j++) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KT; kk++) {
#pragma unroll
        for (int nj = 0; nj < 8; nj += 2) {
          uint32_t b0, b1, b2, b3;
          const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), c = nj + (lane >> 4);
          ldsm_x4_t(sv + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
          mma16816(o[nj], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b0, b1);
          mma16816(o[nj + 1], pa[kk][0], pa[kk][1], pa[kk][2], pa[kk][3], b2, b3);
        }
      }
      const float iA = 1.f / lA, iB = 1.f / lB;
      // ---- stage O (fp16) over this block's own Q rows (only this warp reads them, and it is done), then store the block's
      // 16 rows with 128-bit coalesced writes
      __syncwarp();
      {
        const int g = mt * 16 + (lane >> 2), t = lane & 3;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          *reinterpret_cast<uint32_t*>(so + g * 128 + ((j ^ (g & 7)) << 4) + t * 4) = pack_half2(o[j][0] * iA, o[j][1] * iA);
          *reinterpret_cast<uint32_t*>(so + (g + 8) * 128 + ((j ^ ((g + 8) & 7)) << 4) + t * 4) = pack_half2(o[j][2] * iB, o[j][3] * iB);
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int piece = lane + 32 * i;
        const int r = mt * 16 + (piece >> 3), c = piece & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(so + r * 128 + ((c ^ (r & 7)) << 4));
        *reinterpret_cast<uint4*>(ob + (int64_t)r * HW * inner + c * 8) = val;
      }
    }
  }
}

template <int T>
static void launch_temporal_mma(const __half* qkv, __half* out, int B, int HW, int heads, float scale, cudaStream_t st) {
  MUDG_CUDA(cudaFuncSetAttribute(temporal_attn_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TaCfg<T>::SMEM));
  const int64_t np = (int64_t)B * HW * heads;
  const int64_t want = (np + TaCfg<T>::PROBS - 1) / TaCfg<T>::PROBS;
  const int per_sm = TaCfg<T>::SMEM > 100 * 1024 ? 1 : 2;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count() * per_sm);
  temporal_attn_mma_kernel<T><<<grid, TaCfg<T>::WARPS * 32, TaCfg<T>::SMEM, st>>>(qkv, out, B, HW, heads, scale * 1.4426950408889634f);
  MUDG_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- V -> V^T (the P V MMA wants its B operand K-major)
// V rows [nbatch][len] of pitch elements, head h at columns [h*64, h*64+64)  ->  VT [nbatch][heads*64][len_pad], kv contiguous.
__global__ void __launch_bounds__(256) transpose_v_kernel(const __half* __restrict__ V, int pitch, int len, int heads,
                                                          __half* __restrict__ VT, int len_pad) {
  __shared__ __align__(16) __half tile[64][72];
  const int kv0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int t = threadIdx.x;
  {
    const int r = t >> 2;
    const __half* src = V + ((int64_t)b * len + kv0 + r) * pitch + h * 64;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int c = (t & 3) + 4 * u;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (kv0 + r < len) v = __ldg(reinterpret_cast<const uint4*>(src) + c);
      *reinterpret_cast<uint4*>(&tile[r][c * 8]) = v;
    }
  }
  __syncthreads();
  {
    const int d = t >> 2;
    __half* dst = VT + (((int64_t)b * heads + h) * 64 + d) * len_pad + kv0;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int c = (t & 3) + 4 * u;
      if (kv0 + c * 8 >= len_pad) continue;
      __align__(16) __half o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = tile[c * 8 + i][d];
      *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

const CUtensorMap* vt_map(const __half* base, int len, int len_pad, int width, int batches) {
  const uint64_t dims[5] = {(uint64_t)len, (uint64_t)width, (uint64_t)batches, 1, 1};
  const uint64_t pb = (uint64_t)len_pad * 2;
  const uint64_t str[4] = {pb, pb * width, pb * width * batches, pb * width * batches};
  const uint32_t box[5] = {64, 64, 1, 1, 1};
  return get_tmap(base, dims, str, box);
}

const CUtensorMap* rows_map(const __half* base, int width, int pitch, int rows, int batches) {
  const uint64_t dims[5] = {(uint64_t)width, (uint64_t)rows, (uint64_t)batches, 1, 1};
  const uint64_t pb = (uint64_t)pitch * 2;
  const uint64_t str[4] = {pb, pb * rows, pb * rows * batches, pb * rows * batches};
  const uint32_t box[5] = {64, 128, 1, 1, 1};
  return get_tmap(base, dims, str, box);
}

}  // namespace

void transpose_v(const __half* V, int pitch, int len, int nbatch, int heads, __half* VT, int len_pad, cudaStream_t st) {
  MUDG_REQUIRE(pitch % 8 == 0 && len_pad % 8 == 0 && len_pad >= len && len > 0, "transpose_v: pitch %d len %d pad %d", pitch, len, len_pad);
  MUDG_REQUIRE(((reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(VT)) & 15) == 0, "transpose_v: alignment");
  dim3 grid((len_pad + 63) / 64, heads, nbatch);
  transpose_v_kernel<<<grid, 256, 0, st>>>(V, pitch, len, heads, VT, len_pad);
  MUDG_CUDA(cudaGetLastError());
}

static long long* g_flash_trace = nullptr;
void flash_set_trace(long long* buf) { g_flash_trace = buf; }

void flash_attention(const FlashArgs& a, cudaStream_t st) {
  MUDG_REQUIRE(a.nseg == 1 || a.nseg == 2, "nseg");
  MUDG_REQUIRE(a.q_pitch % 8 == 0 && a.o_pitch % 8 == 0, "pitch alignment");
  const int width = a.heads * 64;
  const CUtensorMap* mq = rows_map(a.Q, width, a.q_pitch, a.Nq, a.F);
  const CUtensorMap* mo = rows_map(a.O, width, a.o_pitch, a.Nq, a.F);
  const CUtensorMap* mk[2];
  const CUtensorMap* mv[2];
  FaParams p{};
  p.nseg = a.nseg;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.stagger = knobs().flash_stagger;      // (-1: 3 for SPLIT 1, 1 for SPLIT 2 -- set below once the split is known)
  p.trace = g_flash_trace;
  for (int i = 0; i < 2; i++) {
    const FlashSeg& s = a.seg[i < a.nseg ? i : 0];
    MUDG_REQUIRE(s.pitch % 8 == 0 && s.len > 0, "kv segment");
    mk[i] = rows_map(s.K, width, s.pitch, s.len, s.nbatch);
    MUDG_REQUIRE(s.VT != nullptr && s.vt_pitch % 8 == 0 && s.vt_pitch >= s.len, "flash attention needs V^T (transpose_v) of every kv segment");
    mv[i] = vt_map(s.VT, s.len, s.vt_pitch, width, s.nbatch);
    p.len[i] = s.len;
    p.kv_div[i] = s.kv_div > 0 ? s.kv_div : 1;
  }
  static OncePerDevice attr;
  if (attr.first()) {
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<1, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<1, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<1, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    MUDG_CUDA(cudaFuncSetAttribute(flash2_kernel<2, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
  }
  // Polynomial exp2 offload (FA4 trick).  Measured on B200 (tests/gpu_bench_flash.py): 5.49 ms vs 4.99 ms without it at
  // 9216 tokens -- the exponent phase is latency-bound at this occupancy, so it stays off (knob flash_poly).
  const int poly = a.nseg == 1 ? knobs().flash_poly : 0;
  dim3 grid((a.Nq + 255) / 256, a.heads, a.F);
  // two softmax warps per TMEM lane quarter (SPLIT 2) for the long self-attentions; knob flash_split = 1 | 2 forces
  const int nb_total = (p.len[0] + 127) / 128;
  int split = (a.nseg == 1 && nb_total >= 4) ? 2 : 1;
  if (knobs().flash_split == 1 || (knobs().flash_split == 2 && a.nseg == 1)) split = knobs().flash_split;
  // MUFU hand-over point: measured best (level 0, 9216 keys) is "after half of the exponentials" with one softmax warp
  // per sub-partition (4.93 ms) and "after all of them" with two (4.77 ms; 5.50 ms with the early hand-over)
  if (p.stagger < 0) p.stagger = split == 2 ? 1 : 3;
  if (a.nseg == 2) flash2_kernel<2, 0, 1><<<grid, F2_THREADS, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  else if (split == 2 && poly == 1) flash2_kernel<1, 1, 2><<<grid, 576, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  else if (split == 2 && poly == 2) flash2_kernel<1, 2, 2><<<grid, 576, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  else if (split == 2) flash2_kernel<1, 0, 2><<<grid, 576, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  else if (poly == 1) flash2_kernel<1, 1, 1><<<grid, F2_THREADS, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  else flash2_kernel<1, 0, 1><<<grid, F2_THREADS, F2_SMEM, st>>>(*mq, *mo, *mk[0], *mv[0], *mk[1], *mv[1], p);
  MUDG_CUDA(cudaGetLastError());
}

void temporal_attention(const __half* qkv, __half* out, int B, int T, int HW, int heads, float scale, cudaStream_t st) {
  MUDG_REQUIRE(T >= 1 && T <= 64, "temporal attention supports T <= 64 (T=%d)", T);
  const bool generic_only = knobs().tattn_generic != 0;
  if (T == 16 && !generic_only) {
    static OncePerDevice attr16;
    if (attr16.first()) MUDG_CUDA(cudaFuncSetAttribute(temporal_attn16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    const int64_t np = (int64_t)B * HW * heads;
    const int64_t want = (np + TA_WARPS - 1) / TA_WARPS;
    const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count() * 4);
    temporal_attn16_kernel<<<grid, TA_WARPS * 32, TA_SMEM, st>>>(qkv, out, B, HW, heads, scale * 1.4426950408889634f);
    MUDG_CUDA(cudaGetLastError());
    return;
  }
  if ((T == 32 || T == 64) && !generic_only) {
    if (T == 32) launch_temporal_mma<32>(qkv, out, B, HW, heads, scale, st);
    else launch_temporal_mma<64>(qkv, out, B, HW, heads, scale, st);
    return;
  }
  int group = 1;
  while (group < T && group < 32) group *= 2;
  const int ppw = group < 32 ? 32 / group : 1;
  const int warps = 4;
  const size_t smem = (size_t)warps * ppw * T * 16 * sizeof(uint4);
  const int64_t npairs = (int64_t)B * HW * heads;
  const int64_t blocks = (npairs + (int64_t)warps * ppw - 1) / ((int64_t)warps * ppw);
  if (smem > 48 * 1024)      // per device and growing with T: set every time (host-side only, microseconds)
    MUDG_CUDA(cudaFuncSetAttribute(temporal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  temporal_attn_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(qkv, out, B, T, HW, heads,
                                                                  scale * 1.4426950408889634f, group, ppw);
  MUDG_CUDA(cudaGetLastError());
}

}  // namespace mudg
