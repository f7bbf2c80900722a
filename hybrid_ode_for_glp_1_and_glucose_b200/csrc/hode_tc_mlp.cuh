// hode_tc_mlp.cuh — the tile-level tensor-core MLP shared by the rollout (hode_rollout_tc.cu) and
// the adjoint (hode_adjoint_tc.cu): TMEM column map, per-tile context, MMA issue for one layer,
// the ReLU/TF32-split epilogue, and the main / helper halves of one MLP evaluation.
#pragma once
#include "hode_common.cuh"
#include "hode_tcgen05.cuh"

namespace hode {

// Optional cycle-level timeline of one main thread (debug builds only: -DHODE_TIMELINE), read back
// with the hode_debug_timeline export; compiles to nothing otherwise.
#ifdef HODE_TIMELINE
__device__ long long g_tl[2 * 16384];
__device__ int g_tl_n;
#define HODE_TL(id)                                                                         \
  do {                                                                                      \
    if (blockIdx.x == 3 && threadIdx.x == 0) {                                              \
      const int n_ = g_tl_n;                                                                \
      if (n_ < 16384) { g_tl[2 * n_] = (id); g_tl[2 * n_ + 1] = clock64(); g_tl_n = n_ + 1; } \
    }                                                                                       \
  } while (0)
#else
#define HODE_TL(id) do { } while (0)
#endif

namespace {
constexpr int TILE = 128;
constexpr int H = 64;
// TMEM columns of one tile:
//   [0,64) accumulator D | [64,128) A_hi | [128,192) A_lo | [192,200) constant [1,1,0..] (bias step)
constexpr uint32_t TM_D0 = 0, TM_AHI = 64, TM_ALO = 128, TM_ONES = 192, TM_TILE_STRIDE = 256;

// Activation stash of the adjoint (one block per stage and hidden layer, per CTA): the tile's
// a_l = relu(z_l) as the BF16 operand image the weight-gradient MMAs read (csrc/probe/bf16_probe.cu),
//   element (trajectory t, feature f) at byte (f / 8) * ST_GRP + t * 16 + (f % 8) * 2,
// a "hi" part (BF16 round-to-nearest) and a "mid" part (BF16 of the remainder; a ~= hi + mid to
// 2^-17), followed by the ReLU masks: word [half][t] = bit j set iff a[32 half + j] > 0.
constexpr int ST_GRP = 2048;                 // one 8-feature group: 128 trajectories x 16 B
constexpr int ST_PART = 8 * ST_GRP;          // 64 features
constexpr int ST_BLK = 2 * ST_PART + 1024;   // hi, mid, masks
}  // namespace

// x[0..7] -> 8 BF16 hi (round to nearest) and 8 BF16 mid = bf16(x - hi); feature 0 in the low half of word 0
__device__ __forceinline__ void bf16_split8(const float* v, uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[q]) : "f"(v[2 * q + 1]), "f"(v[2 * q]));
    const float r0 = v[2 * q] - __uint_as_float(h[q] << 16);
    const float r1 = v[2 * q + 1] - __uint_as_float(h[q] & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m[q]) : "f"(r1), "f"(r0));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

// this thread's 32 activations a = relu(z) of columns [32 half, 32 half + 32), given as the TF32
// hi / lo parts the epilogue produced (a = hi + lo exactly) -> stash block `blk`
__device__ __forceinline__ void stash_store32(uint8_t* blk, int row, int half, const uint32_t* hi_, const uint32_t* lo_) {
  uint32_t mask = 0u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = __uint_as_float(hi_[8 * g + j]) + __uint_as_float(lo_[8 * g + j]);
      mask |= (a[j] > 0.f ? 1u : 0u) << (8 * g + j);
    }
    uint4 hi, mid;
    bf16_split8(a, hi, mid);
    uint8_t* p = blk + (half * 4 + g) * ST_GRP + row * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + ST_PART) = mid;
  }
  reinterpret_cast<uint32_t*>(blk + 2 * ST_PART)[half * TILE + row] = mask;
}

// ---- per-tile context -------------------------------------------------------------------------------
struct TileCtx {
  const float* img;     // shared-memory weight image
  uint64_t* mma_bar;    // this tile's MMA-complete mbarrier
  uint32_t tmem;        // this tile's TMEM column base (lane field 0)
  uint32_t lane_base;   // (warp%4)*32 << 16
  uint32_t parity;      // mbarrier phase to wait for next
  int bar_id;           // named barrier of the tile's 4 main warps (128 threads)
  int bar_all;          // named barrier of the tile's 4 main + 4 helper warps (256 threads)
  int wq;               // warp index inside the tile (warp-uniform)
  int L;                // hidden layer count
};

__device__ __forceinline__ void tile_sync_all(const TileCtx& c) {
  asm volatile("bar.sync %0, 256;" ::"r"(c.bar_all) : "memory");
}
// External-issue mode (the adjoint): a dedicated warp issues every MMA chain, so that no epilogue
// warp is held in a blocking tcgen05.mma issue while it still has stores to do.  The 256 epilogue
// threads ARRIVE on this named barrier, the issuer warp waits on it (mlp_fwd_issue).
constexpr int EXT_ISSUE_BAR = 3, EXT_ISSUE_THREADS = 2 * TILE + 32;
__device__ __forceinline__ void ext_issue_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(EXT_ISSUE_BAR), "n"(EXT_ISSUE_THREADS) : "memory");
}
__device__ __forceinline__ void ext_issue_wait() {
  asm volatile("bar.sync %0, %1;" ::"n"(EXT_ISSUE_BAR), "n"(EXT_ISSUE_THREADS) : "memory");
}

// Issue the MMAs of one layer (one elected thread): D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.
// The two correction products are accumulated FIRST, into a still-small accumulator: the tensor
// core truncates when it adds into D, and measured on B200 (csrc/probe/tc_probe.cu) this order
// gives max err/sum|a*b| = 1.7e-7 (a plain fp32 FMA chain gives 2.2e-7), against 4.5e-7 when the
// products are interleaved per k-step and 6.9e-7 when the large product goes first.
template <bool X3, int N, int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t d, uint32_t ahi, uint32_t alo, uint32_t aones,
                                            uint32_t b_hi, uint32_t b_lo, uint32_t b_bias) {
  constexpr uint32_t idesc = tc::make_idesc_tf32(TILE, N);
  constexpr uint32_t lbo16 = (uint32_t)N;                 // (N*16 bytes) >> 4
  constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
  const uint32_t lo_hi = ((b_hi >> 4) & 0x3FFFu) | (lbo16 << 16);
  const uint32_t lo_lo = ((b_lo >> 4) & 0x3FFFu) | (lbo16 << 16);
  const uint32_t lo_bias = ((b_bias >> 4) & 0x3FFFu) | (lbo16 << 16);
  // D = bias (hi + lo through the constant [1,1,0..] block): initialises the accumulator
  tc::mma_tf32_ts(d, aones, ((uint64_t)desc_hi << 32) | (uint64_t)lo_bias, idesc, 0u);
  if (X3) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      tc::mma_tf32_ts(d, alo + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hi + (uint32_t)ks * 2u * lbo16),
                      idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      tc::mma_tf32_ts(d, ahi + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_lo + (uint32_t)ks * 2u * lbo16),
                      idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    tc::mma_tf32_ts(d, ahi + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hi + (uint32_t)ks * 2u * lbo16),
                    idesc, 1u);
}

// ReLU + TF32 split of 16 accumulator columns (the bias is already in the accumulator), in
// place: v -> hi bits, lo -> lo bits.  hi is rounded to nearest TF32; lo = a - hi is exact in
// fp32 and is truncated to TF32 by the tensor core: a ~= hi + lo to 2^-22.
template <bool X3>
__device__ __forceinline__ void epilogue16(uint32_t* v, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float a = fmaxf(__uint_as_float(v[j]), 0.f);
    const uint32_t h = (__float_as_uint(a) + 0x1000u) & 0xFFFFE000u;
    v[j] = h;
    if (X3) lo[j] = __float_as_uint(a - __uint_as_float(h));
  }
}

// The residual MLP for the 128 trajectories of a tile (reference models/nn_residual.py:136-146).
// Every thread of the tile must call this converged.  x: the 9 input features of this thread's
// trajectory; r: the 6 residuals.  `overlap` runs right after the layer-0 MMAs have been issued: per-thread
// work that does not depend on the network (the mechanistic RHS) hides behind their latency.
template <bool X3, bool EXT = false, class F>
__device__ __forceinline__ void mlp_tile(TileCtx& c, const float* x, float* r, uint8_t* stash, int stash_row,
                                         F&& overlap) {
  const uint32_t t_d = c.tmem + c.lane_base + TM_D0;
  const uint32_t t_ahi = c.tmem + c.lane_base + TM_AHI;
  const uint32_t t_alo = c.tmem + c.lane_base + TM_ALO;
  const uint32_t m_d = c.tmem + TM_D0, m_ahi = c.tmem + TM_AHI, m_alo = c.tmem + TM_ALO;
  const uint32_t m_ones = c.tmem + TM_ONES;
  const float* img = c.img;
  const uint32_t img_s = tc::smem_u32(img);
  const uint32_t bias_s = img_s + (uint32_t)(2 * 1024 + (c.L - 1) * 2 * 4096 + 2 * 1024) * 4u;
  HODE_TL(0);
  // ---- layer 0 operand: 9 features zero-padded to K = 16 ----------------------------------------
  {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k < HODE_NN_IN) tc::split_tf32(x[k], hi[k], lo[k]);
      else { hi[k] = 0u; lo[k] = 0u; }
    }
    HODE_TMEM_ST_X16(t_ahi, hi);
    if (X3) HODE_TMEM_ST_X16(t_alo, lo);
  }
  tc::wait_st();
  tc::fence_before_sync();
  HODE_TL(1);
  // main AND helper warps: the helpers arrive here only after they have observed the previous
  // call's last mbarrier phase, so the layer-0 commit below cannot flip the barrier a second time
  // under a helper that is still busy (it would then wait for a phase that has already passed)
  if (EXT) {
    ext_issue_arrive();
  } else {
    tile_sync_all(c);
    HODE_TL(2);
    if (c.wq == 0) {
      if (tc::elect_one()) {
        tc::fence_after_sync();
        issue_layer<X3, H, 2>(m_d, m_ahi, m_alo, m_ones, img_s, img_s + 1024 * 4, bias_s);
        tc::mma_commit(c.mma_bar);
      }
      __syncwarp();
    }
  }
  HODE_TL(3);
  overlap();
  __syncwarp();   // per-thread code may leave the warp diverged (division slow paths); the tcgen05 .sync.aligned
                  // instructions below need it converged
  uint32_t w_off = 2 * 1024;  // float offset of the next layer's weights inside the image
  // ---- hidden layers: epilogue of layer l feeds the MMAs of layer l+1 ---------------------------
#pragma unroll 1
  for (int l = 0; l < c.L; ++l) {
    const bool last = (l + 1 == c.L);
    const uint32_t b_hi = img_s + w_off * 4;
    const uint32_t b_lo = b_hi + (last ? 1024u : 4096u) * 4u;
    tc::mbar_wait(c.mma_bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
    HODE_TL(10 + 10 * l);
    // the main warp owns accumulator columns [0,32) of its 32 lanes, the helper warp of the same
    // lane quarter columns [32,64) (mlp_tile_helper): the epilogue latency per layer is halved
    uint32_t v0[32], lo[32];
    HODE_TMEM_LD_X32(t_d, v0);
    tc::wait_ld();
    HODE_TL(11 + 10 * l);
    epilogue16<X3>(v0, lo);
    HODE_TMEM_ST_X16(t_ahi, v0);
    if (X3) HODE_TMEM_ST_X16(t_alo, lo);
    epilogue16<X3>(v0 + 16, lo + 16);
    HODE_TMEM_ST_X16(t_ahi + 16, (v0 + 16));
    if (X3) HODE_TMEM_ST_X16(t_alo + 16, (lo + 16));
    tc::wait_st();
    tc::fence_before_sync();
    HODE_TL(12 + 10 * l);
    if (EXT) {
      ext_issue_arrive();
    } else {
      tile_sync_all(c);
      HODE_TL(13 + 10 * l);
      if (c.wq == 0) {
        if (tc::elect_one()) {
          tc::fence_after_sync();
          const uint32_t b_bias = bias_s + (uint32_t)(l + 1) * 512u * 4u;
          if (!last) issue_layer<X3, H, 8>(m_d, m_ahi, m_alo, m_ones, b_hi, b_lo, b_bias);
          else issue_layer<X3, 16, 8>(m_d, m_ahi, m_alo, m_ones, b_hi, b_lo, b_bias);
          tc::mma_commit(c.mma_bar);
        }
        __syncwarp();
      }
    }
    // adjoint: a_l = relu(z_l) of this thread's trajectory, columns [0,32), goes to the stash AFTER
    // the next layer's MMAs have been issued (off the critical path)
    if (X3 && stash) stash_store32(stash + (size_t)l * ST_BLK, stash_row, 0, v0, lo);
    HODE_TL(14 + 10 * l);
    w_off += 2 * 4096;
  }
  // ---- output layer epilogue: 6 of the 16 accumulator columns ------------------------------------
  tc::mbar_wait(c.mma_bar, c.parity);
  c.parity ^= 1u;
  tc::fence_after_sync();
  HODE_TL(90);
  {
    uint32_t v[8];
    HODE_TMEM_LD_X8(t_d, v);
    tc::wait_ld();
#pragma unroll
    for (int i = 0; i < NS; ++i) r[i] = __uint_as_float(v[i]);
  }
  HODE_TL(91);
  // The next call overwrites A (tcgen05.st: every MMA has completed) and its layer-0 MMAs write
  // D only after a tile barrier that every thread reaches after its wait::ld above.
}

template <bool X3>
__device__ __forceinline__ void mlp_tile(TileCtx& c, const float* x, float* r, uint8_t* stash = nullptr,
                                         int stash_row = 0) {
  mlp_tile<X3, false>(c, x, r, stash, stash_row, [] {});
}

// External-issue mode: the MMA chains of one mlp_tile<X3, true>() call, issued by a dedicated (converged)
// warp.  Every chain is committed to the tile's mbarrier; c.parity tracks the phase of the last commit.
template <bool X3>
__device__ __forceinline__ void mlp_fwd_issue(TileCtx& c) {
  const uint32_t m_d = c.tmem + TM_D0, m_ahi = c.tmem + TM_AHI, m_alo = c.tmem + TM_ALO;
  const uint32_t m_ones = c.tmem + TM_ONES;
  const uint32_t img_s = tc::smem_u32(c.img);
  const uint32_t bias_s = img_s + (uint32_t)(2 * 1024 + (c.L - 1) * 2 * 4096 + 2 * 1024) * 4u;
  ext_issue_wait();
  if (tc::elect_one()) {
    tc::fence_after_sync();
    issue_layer<X3, H, 2>(m_d, m_ahi, m_alo, m_ones, img_s, img_s + 1024 * 4, bias_s);
    tc::mma_commit(c.mma_bar);
  }
  __syncwarp();
  c.parity ^= 1u;
  uint32_t w_off = 2 * 1024;
#pragma unroll 1
  for (int l = 0; l < c.L; ++l) {
    const bool last = (l + 1 == c.L);
    const uint32_t b_hi = img_s + w_off * 4;
    const uint32_t b_lo = b_hi + (last ? 1024u : 4096u) * 4u;
    ext_issue_wait();
    if (tc::elect_one()) {
      tc::fence_after_sync();
      const uint32_t b_bias = bias_s + (uint32_t)(l + 1) * 512u * 4u;
      if (!last) issue_layer<X3, H, 8>(m_d, m_ahi, m_alo, m_ones, b_hi, b_lo, b_bias);
      else issue_layer<X3, 16, 8>(m_d, m_ahi, m_alo, m_ones, b_hi, b_lo, b_bias);
      tc::mma_commit(c.mma_bar);
    }
    __syncwarp();
    c.parity ^= 1u;
    w_off += 2 * 4096;
  }
}

// Helper warps: the other half of every hidden-layer epilogue.  Must be called once per
// mlp_tile() call of the tile's main warps (same number of tile-wide barriers and mbarrier phases).
template <bool X3, bool EXT = false>
__device__ __forceinline__ void mlp_tile_helper(TileCtx& c, uint8_t* stash = nullptr, int stash_row = 0) {
  const uint32_t t_d = c.tmem + c.lane_base + TM_D0 + 32;
  const uint32_t t_ahi = c.tmem + c.lane_base + TM_AHI + 32;
  const uint32_t t_alo = c.tmem + c.lane_base + TM_ALO + 32;
  if (EXT) ext_issue_arrive();
  else tile_sync_all(c);   // pairs with the main warps' barrier before the layer-0 MMAs (see mlp_tile)
#pragma unroll 1
  for (int l = 0; l < c.L; ++l) {
    tc::mbar_wait(c.mma_bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
    uint32_t v[32], lo[32];
    HODE_TMEM_LD_X32(t_d, v);
    tc::wait_ld();
    epilogue16<X3>(v, lo);
    HODE_TMEM_ST_X16(t_ahi, v);
    if (X3) HODE_TMEM_ST_X16(t_alo, lo);
    epilogue16<X3>(v + 16, lo + 16);
    HODE_TMEM_ST_X16(t_ahi + 16, (v + 16));
    if (X3) HODE_TMEM_ST_X16(t_alo + 16, (lo + 16));
    tc::wait_st();
    tc::fence_before_sync();
    if (EXT) ext_issue_arrive();
    else tile_sync_all(c);
    if (X3 && stash) stash_store32(stash + (size_t)l * ST_BLK, stash_row, 1, v, lo);   // columns [32,64)
  }
  // the output layer's phase: nothing to read, but the phase must be observed so that the next
  // call's first wait cannot be satisfied by a stale parity
  tc::mbar_wait(c.mma_bar, c.parity);
  c.parity ^= 1u;
}

}  // namespace hode

#ifdef HODE_TIMELINE
// exactly one translation unit is compiled with -DHODE_TIMELINE (tools/timeline.py)
extern "C" int hode_debug_timeline(long long* out_host, int max_events) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, hode::g_tl_n, sizeof(int));
  if (n > max_events) n = max_events;
  cudaMemcpyFromSymbol(out_host, hode::g_tl, (size_t)n * 2 * sizeof(long long));
  int zero = 0;
  cudaMemcpyToSymbol(hode::g_tl_n, &zero, sizeof(int));
  return n;
}
#endif
