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
#ifndef HODE_TL_TID
#define HODE_TL_TID 0   // the thread whose marks are recorded (-DHODE_TL_TID=...: another role's warp)
#endif
#define HODE_TL(id)                                                                         \
  do {                                                                                      \
    if (blockIdx.x == 3 && threadIdx.x == HODE_TL_TID) {                                              \
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
// Arithmetic of the tile MLP (template parameter MODE; `true` / `false` of older call sites = 3xTF32 / TF32):
//   MLP_TF32   one kind::tf32 pass (~1e-3 on the residual: "fast" mode, never the default)
//   MLP_X3     3xTF32: D = A_lo B_hi + A_hi B_lo + A_hi B_hi, three kind::tf32 passes (the default: max error
//              1.7e-7 of sum|a b|, better than an FP32 FMA chain)
//   MLP_MIXED  D = bf16(A_lo) bf16(B_hi) + bf16(A_hi) bf16(B_lo) + A_hi B_hi: one kind::tf32 pass + two
//              kind::f16 BF16 passes at twice its rate = 2 pass-equivalents (opt-in: max error 4.5e-7 of
//              sum|a b|, 2x an FP32 FMA chain — csrc/probe/mix_probe.cu)
//   MLP_MIX3   D = bf16(A_lo) bf16(B_hi) + A_hi B_lo + A_hi B_hi: two kind::tf32 passes + one BF16 pass = 2.5
//              pass-equivalents; its A operand needs only 96 TMEM columns (no second copy of A_hi), so THREE
//              tiles fit the 512 columns of an SM — the rollout's default (hode_rollout_tc.cu); at least as
//              accurate as MLP_MIXED (the A_hi B_lo term is exact to TF32 instead of BF16)
//   MLP_H16    D = bf16(A_lo) bf16(B_hi) + f16(A_hi / 64) f16(64 B_lo) + f16(A_hi) f16(B_hi) with A_hi = fp16(a) (round to
//              nearest, saturating at +-65504) and A_lo = a - A_hi, B likewise: THREE kind::f16 passes = 1.5
//              pass-equivalents.  FP16 carries the same 11 significant bits as TF32, so the main term is as exact as a
//              TF32 one at twice the rate.  The A_hi B_lo term runs in FP16 too, with the power-of-two scale moved from
//              one operand to the other so that the remainder B_lo (<= 2^-11 |B|) sits in FP16's normal range: weights are
//              then represented to 22 bits, as in MLP_MIX3.  (With bf16(A_hi) bf16(B_lo) instead, the 2^-20 rounding of
//              B_lo acts like a fixed perturbation of the weights, because post-ReLU A_hi never changes sign: measured
//              1.7e-5 instead of < 1e-5 on the parameter-sweep parity case; csrc/probe/f16_probe.cu mode 2.)
//              A operand = 96 TMEM columns (f16(A_hi), f16(A_hi / 64), bf16(A_lo), two features per column), so three
//              tiles fit like MLP_MIX3's; weights = 1.5 float-sized parts per layer.  (Mixed-format instructions —
//              a_format != b_format in one kind::f16 descriptor — raise an illegal-instruction fault on B200.)
//              Contract: float32-equivalent products while |activations|, |weights| <= 65504 (beyond that the hi part
//              saturates and the remainder is carried with BF16's 8 bits); activations below 2^-8 and weight remainders
//              below 2^-20 keep absolute precisions of 2^-19 / 2^-31 in the A_hi B_lo term (FP16 subnormals).
//   MLP_H16_2T the arithmetic of MLP_H16 (bit-identical results) in the two-tile shape with helper warps: every hidden-layer
//              epilogue is split over two warps per lane quarter, which shortens the chain of a tile by ~18 %.  Used when the
//              cohort fits two tiles per SM anyway (small batches / config 3's per-GPU shard): latency instead of occupancy.
constexpr int MLP_TF32 = 0, MLP_X3 = 1, MLP_MIXED = 2, MLP_MIX3 = 3, MLP_H16 = 4, MLP_H16_2T = 5;
// modes that run three tiles per CTA without helper warps
template <int MODE> __host__ __device__ constexpr bool three_tiles() { return MODE == MLP_MIX3 || MODE == MLP_H16; }
template <int MODE> __host__ __device__ constexpr bool is_h16() { return MODE == MLP_H16 || MODE == MLP_H16_2T; }
// TMEM columns of one tile:
//   [0,64) accumulator D | [64,128) A_hi (TF32) | [192,200) constant [1,1,0..] (bias step) |
//   MLP_X3:    [128,192) A_lo = A - A_hi (TF32)
//   MLP_MIXED: [128,160) bf16(A_hi), two features per column | [160,192) bf16(A - A_hi)
//   MLP_MIX3:  [128,160) bf16(A - A_hi); tile stride 160, ONE constant block for the whole CTA at column 480
constexpr uint32_t TM_D0 = 0, TM_AHI = 64, TM_ALO = 128, TM_AHB = 128, TM_ALB = 160, TM_ONES = 192, TM_TILE_STRIDE = 256;
constexpr uint32_t TM3_ALB = 128, TM3_TILE_STRIDE = 160, TM3_ONES_ABS = 480;
//   MLP_H16:   [64,96) f16(A_hi), two features per column | [96,128) f16(A_hi / 64) | [128,160) bf16(A - A_hi); tile stride 160,
//              constant block at 480 (the MLP_MIX3 map with a different A block)
constexpr uint32_t TMH_A16 = 64, TMH_AHS = 96, TMH_ALB = 128, TMH_TILE_STRIDE = 160;
template <int MODE> __host__ __device__ constexpr uint32_t tm_tile_stride() {
  return MODE == MLP_MIX3 ? TM3_TILE_STRIDE : (MODE == MLP_H16 ? TMH_TILE_STRIDE : TM_TILE_STRIDE);
}
// float offsets inside the shared-memory weight image (prep_tc_image_kernel, hode_rollout_tc.cu):
//   layer 0 (K = 16: 9 features, feature 9 = constant 1 whose weight column is the bias, zero padding):
//     [B_hi tf32 1024][second half 1024]
//   hidden layer l = 1..L-1 (K = 64, N = 64): [B_hi 4096][second half 4096]
//   output layer (K = 64, N = 16):            [B_hi 1024][second half 1024]
//   second half = B_lo (TF32 of B - B_hi; MLP_X3) or [bf16(B_hi)][bf16(B - B_hi)] (half as many floats each; MLP_MIXED)
//   bias blocks (one K = 8 TF32 step, columns 0/1 = hi/lo): L x 512 (block 0 unused) + 128
//   MLP_MIX3: [B_hi tf32][B_lo tf32][bf16(B_hi)] = 2.5 x the floats of B_hi per layer
constexpr uint32_t IMG_L0 = 2048, IMG_HID = 8192, IMG_OUT = 2048;
//   MLP_H16:  [f16(B)][f16(64 (B - f16(B)))][bf16(f16(B))], 2-byte elements = 1.5 x the floats of B_hi per layer
template <int MODE> __host__ __device__ constexpr uint32_t img_l0() { return MODE == MLP_MIX3 ? 2560u : (is_h16<MODE>() ? 1536u : IMG_L0); }
template <int MODE> __host__ __device__ constexpr uint32_t img_hid() { return MODE == MLP_MIX3 ? 10240u : (is_h16<MODE>() ? 6144u : IMG_HID); }
// (Measured dead end: the 64 -> 6 output layer of MLP_MIX3 on the CUDA cores — 384 FFMA per thread straight from the last
// hidden accumulator, no N = 16 MMAs and one issue -> commit -> wake-up phase less — was 6 % SLOWER, 968 M against
// 1 034 M trajectory-steps/s: the tile's chain is bound by its threads' instruction latency, not by the tensor phases.)
template <int MODE> __host__ __device__ constexpr uint32_t img_out() { return MODE == MLP_MIX3 ? 2560u : (is_h16<MODE>() ? 1536u : IMG_OUT); }
// What bounds MLP_MIX3 (round 2 measurements, tools/probe_sched.py + tools/timeline.py on 1 / 2 / 3 resident tiles):
// the DP5(4) round of a tile takes 34.3 us whether the tile is alone on its SM or not, and the issue of a hidden
// layer's 21 N = 64 MMAs takes 861 cycles alone (41 per MMA: tcgen05.mma issue is paced by execution, 32.5 cycles,
// plus ~8) and 1 040 with three tiles.  Per evaluation a tile holds the tensor pipe for 3 x 21 x 41 + 5 x 41 + 21 x 17
// ~ 3.1 k cycles, three tiles for 9.3 k of the 11.2 k cycles an evaluation lasts: the pipe is 83 % occupied in issue
// terms (ncu counts 62 % "active": the gaps between MMAs and the N = 16 output layer are not active cycles).
// Two measured dead ends follow from that.  (1) A split pipeline — every layer's accumulator produced as two N = 32
// halves and consumed as two K = 32 halves, the next layer's K-half-0 MMAs running under the second half of the
// epilogue — doubles the MMA count at 24.5 cycles each, i.e. trades chain latency for tensor time one to one:
// correct (rk4 parity 1.6e-6) and 5 % SLOWER (969 M against 1 020 M trajectory-steps/s), both with warp 0 issuing
// between the halves of its own epilogue and (2) with a dedicated issuer warp per tile (a fourth warpgroup,
// setmaxnreg 160 / 24, 314 bytes of spills).  What is left is the 17 % between the chain and the pipe.

// Activation stash of the adjoint (one block per hidden layer, per CTA): the tile's
// a_l = relu(z_l) as the BF16 operand image the weight-gradient MMAs read (csrc/probe/bf16_probe.cu),
//   element (trajectory t, feature f) at byte (f / 8) * ST_GRP + t * 16 + (f % 8) * 2,
// a "hi" part (BF16 round-to-nearest) and a "mid" part (BF16 of the remainder; a ~= hi + mid to 2^-17).
constexpr int ST_GRP = 2048;                 // one 8-feature group: 128 trajectories x 16 B
constexpr int ST_PART = 8 * ST_GRP;          // 64 features
constexpr int ST_BLK = 2 * ST_PART;          // hi, mid
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
// hi / lo parts the epilogue produced (a = hi + lo exactly) -> stash block `blk`; returns the ReLU mask
// (bit j set iff a[32 half + j] > 0)
__device__ __forceinline__ uint32_t stash_store32(uint8_t* blk, int row, int half, const uint32_t* hi_, const uint32_t* lo_) {
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
  return mask;
}

// ---- per-tile context -------------------------------------------------------------------------------
struct TileCtx {
  const float* img;     // shared-memory weight image
  uint64_t* mma_bar;    // this tile's MMA-complete mbarrier
  uint32_t tmem;        // this tile's TMEM column base (lane field 0)
  uint32_t t_ones;      // TMEM address (lane field 0) of the constant [1,1,0..] block of the bias step
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
__device__ __forceinline__ void tile_sync_main(const TileCtx& c) {
  asm volatile("bar.sync %0, 128;" ::"r"(c.bar_id) : "memory");
}
// kind::f16 instruction descriptor, BF16 operands, FP32 accumulation, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 instruction descriptor with separate operand formats (0 = F16, 1 = BF16), FP32 accumulation, both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int fmt_a, int fmt_b, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt_a << 7) | ((uint32_t)fmt_b << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Issue the MMAs of one layer (one elected thread): split-precision products with FP32-equivalent accuracy.
// The correction products are accumulated FIRST, into a still-small accumulator: the tensor core truncates when
// it adds into D, and measured on B200 (csrc/probe/tc_probe.cu, mix_probe.cu) this order gives max err / sum|a b| =
// 1.7e-7 for 3xTF32 (a plain FP32 FMA chain: 2.2e-7; 6.9e-7 with the large product first) and 4.5e-7 for the mixed
// split (8.0e-7 with the large product first).
// The bias rides on the tensor pipe: one K = 8 TF32 step against a constant [1,1,0..] block (hidden and output
// layers) or, in layer 0, the weight column of the constant-1 input feature.
// b_hi: TF32 image of the weights; b_2: the second half of the layer's image (IMG_* above).
template <int MODE, int N, int K, bool BIAS>
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t b_hi, uint32_t b_2, uint32_t b_bias, uint32_t t_ones) {
  constexpr uint32_t idesc = tc::make_idesc_tf32(TILE, N), idesc_b = make_idesc_bf16(TILE, N);
  constexpr uint32_t lbo16 = (uint32_t)N;                 // (N*16 bytes) >> 4: stride between 16-byte K chunks
  constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
  const uint32_t d = tmem + TM_D0, ahi = tmem + TM_AHI;
  const uint32_t lo_hi = ((b_hi >> 4) & 0x3FFFu) | (lbo16 << 16);
  const uint32_t lo_2 = ((b_2 >> 4) & 0x3FFFu) | (lbo16 << 16);
  const uint32_t lo_bias = ((b_bias >> 4) & 0x3FFFu) | (lbo16 << 16);
  uint32_t acc = 0u;
  if (BIAS) {   // D = bias (hi + lo through the constant [1,1,0..] block): initialises the accumulator
    tc::mma_tf32_ts(d, t_ones, ((uint64_t)desc_hi << 32) | (uint64_t)lo_bias, idesc, 0u);
    acc = 1u;
  }
  if (MODE == MLP_X3) {
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {   // A_lo B_hi
      tc::mma_tf32_ts(d, tmem + TM_ALO + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hi + (uint32_t)ks * 2u * lbo16), idesc, acc);
      acc = 1u;
    }
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks)     // A_hi B_lo
      tc::mma_tf32_ts(d, ahi + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_2 + (uint32_t)ks * 2u * lbo16), idesc, 1u);
  }
  if (MODE == MLP_MIXED) {
    // one BF16 MMA contracts K = 16 = 8 TMEM columns of A and two 16-byte K chunks of B; bf16(B - B_hi) follows
    // bf16(B_hi) after K * N * 2 bytes
    const uint32_t lo_lob = lo_2 + ((uint32_t)(K * N * 2) >> 4);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {  // bf16(A_lo) bf16(B_hi)
      mma_bf16_ts(d, tmem + TM_ALB + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_2 + (uint32_t)ks * 2u * lbo16), idesc_b, acc);
      acc = 1u;
    }
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks)    // bf16(A_hi) bf16(B_lo)
      mma_bf16_ts(d, tmem + TM_AHB + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_lob + (uint32_t)ks * 2u * lbo16), idesc_b, 1u);
  }
  if (MODE == MLP_MIX3) {
    // b_2 = B_lo (TF32), then bf16(B_hi) after K * N * 4 bytes
    const uint32_t lo_hib = lo_2 + ((uint32_t)(K * N * 4) >> 4);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {  // bf16(A_lo) bf16(B_hi)
      mma_bf16_ts(d, tmem + TM3_ALB + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hib + (uint32_t)ks * 2u * lbo16), idesc_b, acc);
      acc = 1u;
    }
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks)     // A_hi B_lo
      tc::mma_tf32_ts(d, ahi + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_2 + (uint32_t)ks * 2u * lbo16), idesc, 1u);
  }
  if constexpr (is_h16<MODE>()) {
    // b_hi = f16(B) (2-byte elements, K chunks of 8), then f16(64 (B - f16(B))) and bf16(f16(B)), K * N * 2 bytes each;
    // one MMA contracts K = 16 = 8 TMEM columns of A and two 16-byte K chunks of B
    constexpr uint32_t id_hh = make_idesc_f16(0, 0, TILE, N);
    const uint32_t lo_los = lo_hi + ((uint32_t)(K * N * 2) >> 4), lo_hib = lo_los + ((uint32_t)(K * N * 2) >> 4);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {  // bf16(A_lo) bf16(B_hi)
      mma_bf16_ts(d, tmem + TMH_ALB + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hib + (uint32_t)ks * 2u * lbo16), idesc_b, acc);
      acc = 1u;
    }
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks)    // f16(A_hi / 64) f16(64 B_lo)
      mma_bf16_ts(d, tmem + TMH_AHS + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_los + (uint32_t)ks * 2u * lbo16), id_hh, 1u);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks)    // f16(A_hi) f16(B_hi)
      mma_bf16_ts(d, tmem + TMH_A16 + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hi + (uint32_t)ks * 2u * lbo16), id_hh, 1u);
  } else {
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {     // A_hi B_hi
      tc::mma_tf32_ts(d, ahi + ks * 8, ((uint64_t)desc_hi << 32) | (uint64_t)(lo_hi + (uint32_t)ks * 2u * lbo16), idesc, acc);
      acc = 1u;
    }
  }
}

// the MMAs of layer `next` (0 = input layer .. L = output layer) of a tile, from its shared-memory weight image
template <int MODE>
__device__ __forceinline__ void issue_mlp_layer(uint32_t tmem, uint32_t t_ones, uint32_t img_s, int L, int next) {
  const uint32_t bias_s = img_s + (img_l0<MODE>() + (uint32_t)(L - 1) * img_hid<MODE>() + img_out<MODE>()) * 4u;
  if (next == 0) {
    issue_layer<MODE, H, 16, false>(tmem, img_s, img_s + 1024u * 4u, 0u, t_ones);
  } else if (next < L) {
    const uint32_t b = img_s + (img_l0<MODE>() + (uint32_t)(next - 1) * img_hid<MODE>()) * 4u;
    issue_layer<MODE, H, 64, true>(tmem, b, b + 4096u * 4u, bias_s + (uint32_t)next * 512u * 4u, t_ones);
  } else {
    const uint32_t b = img_s + (img_l0<MODE>() + (uint32_t)(L - 1) * img_hid<MODE>()) * 4u;
    issue_layer<MODE, 16, 64, true>(tmem, b, b + 1024u * 4u, bias_s + (uint32_t)L * 512u * 4u, t_ones);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // feature 2c in the low half (mix_probe test 1)
  return r;
}

// ReLU + split of 16 accumulator columns (the bias is already in the accumulator): v -> TF32 hi bits (in place,
// round to nearest), lo = a - hi (exact in FP32; the tensor core truncates it to TF32: a ~= hi + lo to 2^-22),
// hb / lb = the BF16 operands of the mixed split's two cross terms (8 columns each).
template <int MODE>
__device__ __forceinline__ void epilogue16(uint32_t* v, uint32_t* lo, uint32_t* hb, uint32_t* lb) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float a0 = fmaxf(__uint_as_float(v[j]), 0.f), a1 = fmaxf(__uint_as_float(v[j + 1]), 0.f);
    const uint32_t h0 = (__float_as_uint(a0) + 0x1000u) & 0xFFFFE000u, h1 = (__float_as_uint(a1) + 0x1000u) & 0xFFFFE000u;
    v[j] = h0;
    v[j + 1] = h1;
    if (MODE != MLP_TF32) {
      const float l0 = a0 - __uint_as_float(h0), l1 = a1 - __uint_as_float(h1);
      lo[j] = __float_as_uint(l0);
      lo[j + 1] = __float_as_uint(l1);
      if (MODE == MLP_MIXED) {
        hb[j >> 1] = pack_bf16x2(__uint_as_float(h0), __uint_as_float(h1));
        lb[j >> 1] = pack_bf16x2(l0, l1);
      }
    }
  }
}
// MLP_MIX3: v -> a = relu(v) in place — the tensor core reads the upper 19 bits of a kind::tf32 operand, i.e. it
// truncates a to A_hi by itself — and lb = bf16 pairs of the remainder a - trunc(a) (8 columns).  3 instructions per
// column: FMNMX, LOP3, half a packed FADD2 (sm_100 f32x2), half a F2FP.  Truncation instead of rounding makes the
// remainder at most 2^-10 |a| instead of 2^-11 |a|; its BF16 rounding error (2^-19 |a|) stays far below the 1e-5
// parity bar (tools/tc_mode_errors.py: 1.0e-6 against the oracle, the same as 3xTF32).
__device__ __forceinline__ void epilogue16_mix3(uint32_t* v, uint32_t* lb) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float a0 = fmaxf(__uint_as_float(v[j]), 0.f), a1 = fmaxf(__uint_as_float(v[j + 1]), 0.f);
#ifdef HODE_MIX3_TRUNC
    const uint32_t h0 = __float_as_uint(a0) & 0xFFFFE000u, h1 = __float_as_uint(a1) & 0xFFFFE000u;
    v[j] = __float_as_uint(a0);
    v[j + 1] = __float_as_uint(a1);
#else
    const uint32_t h0 = (__float_as_uint(a0) + 0x1000u) & 0xFFFFE000u, h1 = (__float_as_uint(a1) + 0x1000u) & 0xFFFFE000u;
    v[j] = h0;
    v[j + 1] = h1;
#endif
    uint64_t ap, hp, lp;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ap) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(hp) : "r"(h0), "r"(h1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(lp) : "l"(ap), "l"(hp));
    float l0, l1;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(l0), "=f"(l1) : "l"(lp));
    lb[j >> 1] = pack_bf16x2(l0, l1);
  }
}
// MLP_H16: 16 accumulator columns -> 8 packed f16 pairs of a = relu(v) (round to nearest, saturating), the same scaled by
// 2^-6, and 8 packed BF16 pairs of the remainder a - f16(a).  Per pair: 2 FMNMX, F2FP.F16, HMUL2, 2 HADD2.F32 (unpack), one
// packed FADD2, F2FP.BF16.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // feature 2c in the low half
  return r;
}
constexpr float H16_LO_SCALE = 64.f;   // 2^6: B_lo is stored as f16(64 B_lo), A_hi a second time as f16(A_hi / 64)
__device__ __forceinline__ void split_h16_pair(float a0, float a1, uint32_t& h16, uint32_t& hs, uint32_t& lb) {
  h16 = pack_f16x2_sat(a0, a1);
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(hs) : "r"(h16), "r"(0x24002400u));   // x 2^-6 (exact unless subnormal)
  float f0, f1;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(f0), "=f"(f1) : "r"(h16));
  uint64_t ap, hp, lp;
  asm("mov.b64 %0, {%1,%2};" : "=l"(ap) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(hp) : "f"(f0), "f"(f1));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(lp) : "l"(ap), "l"(hp));
  float l0, l1;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(l0), "=f"(l1) : "l"(lp));
  lb = pack_bf16x2(l0, l1);
}
// weight-side split of one element: f16(w), f16(64 (w - f16(w))), bf16(f16(w))
__device__ __forceinline__ void split_h16_weight(float w, uint16_t& h16, uint16_t& los, uint16_t& hib) {
  const uint32_t p = pack_f16x2_sat(w, 0.f);
  float f0;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f0) : "r"(p));
  h16 = (uint16_t)(p & 0xFFFFu);
  los = (uint16_t)(pack_f16x2_sat((w - f0) * H16_LO_SCALE, 0.f) & 0xFFFFu);
  hib = (uint16_t)(pack_bf16x2(f0, 0.f) & 0xFFFFu);
}
__device__ __forceinline__ void epilogue16_h16(const uint32_t* v, uint32_t* h16, uint32_t* hs, uint32_t* lb) {
#pragma unroll
  for (int j = 0; j < 16; j += 2)
    split_h16_pair(fmaxf(__uint_as_float(v[j]), 0.f), fmaxf(__uint_as_float(v[j + 1]), 0.f), h16[j >> 1], hs[j >> 1], lb[j >> 1]);
}
// The hidden-layer epilogue of one MLP_H16 thread: all 64 accumulator columns of its lane, pipelined like MLP_MIX3's.
__device__ __forceinline__ void epilogue64_h16(uint32_t t_lane) {
  uint32_t c0[16], c1[16], c2[16], c3[16], h16[8], hs[8], lb[8];
  HODE_TMEM_LD_X16(t_lane + TM_D0, c0);
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 16, c1);
  tc::wait_ld();
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 32, c2);
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 48, c3);
#define HODE_H16_CHUNK(c, q)                          \
  epilogue16_h16(c, h16, hs, lb);                     \
  HODE_TMEM_ST_X8(t_lane + TMH_A16 + 8 * (q), h16);   \
  HODE_TMEM_ST_X8(t_lane + TMH_AHS + 8 * (q), hs);    \
  HODE_TMEM_ST_X8(t_lane + TMH_ALB + 8 * (q), lb)
  HODE_H16_CHUNK(c0, 0);
  HODE_H16_CHUNK(c1, 1);
  tc::wait_ld();
  HODE_H16_CHUNK(c2, 2);
  HODE_H16_CHUNK(c3, 3);
#undef HODE_H16_CHUNK
  tc::wait_st();
  tc::fence_before_sync();
}
// The hidden-layer epilogue of one MLP_MIX3 thread (no helper warps): all 64 accumulator columns of its lane.
__device__ __forceinline__ void epilogue64_mix3(uint32_t t_lane) {
  // four 16-column chunks, software-pipelined: tcgen05.wait::ld drains every outstanding load, so the loads of the
  // next two chunks are issued before the arithmetic of the current ones and their latency hides behind it
  uint32_t c0[16], c1[16], c2[16], c3[16], lb[8];
  HODE_TMEM_LD_X16(t_lane + TM_D0, c0);
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 16, c1);
  tc::wait_ld();
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 32, c2);
  HODE_TMEM_LD_X16(t_lane + TM_D0 + 48, c3);
  epilogue16_mix3(c0, lb);
  HODE_TMEM_ST_X16(t_lane + TM_AHI, c0);
  HODE_TMEM_ST_X8(t_lane + TM3_ALB, lb);
  epilogue16_mix3(c1, lb);
  HODE_TMEM_ST_X16(t_lane + TM_AHI + 16, c1);
  HODE_TMEM_ST_X8(t_lane + TM3_ALB + 8, lb);
  tc::wait_ld();
  epilogue16_mix3(c2, lb);
  HODE_TMEM_ST_X16(t_lane + TM_AHI + 32, c2);
  HODE_TMEM_ST_X8(t_lane + TM3_ALB + 16, lb);
  epilogue16_mix3(c3, lb);
  HODE_TMEM_ST_X16(t_lane + TM_AHI + 48, c3);
  HODE_TMEM_ST_X8(t_lane + TM3_ALB + 24, lb);
  tc::wait_st();
  tc::fence_before_sync();
}

// MLP_H16_2T: 32 accumulator columns [col0, col0 + 32) of this thread's lane (main warp: col0 = 0, helper warp: 32)
__device__ __forceinline__ void epilogue32_h16(uint32_t t_lane, uint32_t col0) {
  uint32_t c0[16], c1[16], h16[8], hs[8], lb[8];
  const uint32_t p0 = col0 >> 1;   // two features per operand column
  HODE_TMEM_LD_X16(t_lane + TM_D0 + col0, c0);
  HODE_TMEM_LD_X16(t_lane + TM_D0 + col0 + 16, c1);
  tc::wait_ld();
  epilogue16_h16(c0, h16, hs, lb);
  HODE_TMEM_ST_X8(t_lane + TMH_A16 + p0, h16);
  HODE_TMEM_ST_X8(t_lane + TMH_AHS + p0, hs);
  HODE_TMEM_ST_X8(t_lane + TMH_ALB + p0, lb);
  epilogue16_h16(c1, h16, hs, lb);
  HODE_TMEM_ST_X8(t_lane + TMH_A16 + p0 + 8, h16);
  HODE_TMEM_ST_X8(t_lane + TMH_AHS + p0 + 8, hs);
  HODE_TMEM_ST_X8(t_lane + TMH_ALB + p0 + 8, lb);
  tc::wait_st();
  tc::fence_before_sync();
}
// The hidden-layer epilogue of one thread: 32 accumulator columns [col0, col0 + 32) of its TMEM lane -> the next
// layer's A operand.  v / lo keep a = v + lo for the adjoint's stash.  Ends with wait::st + fence: the caller
// signals the MMA issuer next.  store = false: only v / lo are produced (the last hidden layer of the adjoint's
// recomputation feeds no further product).
template <int MODE>
__device__ __forceinline__ void epilogue32_to_tmem(uint32_t t_lane, uint32_t col0, uint32_t* v, uint32_t* lo, bool store = true) {
  const uint32_t t_ahi = t_lane + TM_AHI + col0, t_alo = t_lane + TM_ALO + col0;
  const uint32_t t_ahb = t_lane + TM_AHB + (col0 >> 1), t_alb = t_lane + TM_ALB + (col0 >> 1);
  HODE_TMEM_LD_X32(t_lane + TM_D0 + col0, v);
  tc::wait_ld();
  uint32_t hb[16], lb[16];
  epilogue16<MODE>(v, lo, hb, lb);
  if (store) {
    HODE_TMEM_ST_X16(t_ahi, v);
    if (MODE == MLP_X3) HODE_TMEM_ST_X16(t_alo, lo);
    if (MODE == MLP_MIXED) { HODE_TMEM_ST_X8(t_ahb, hb); HODE_TMEM_ST_X8(t_alb, lb); }
  }
  epilogue16<MODE>(v + 16, lo + 16, hb + 8, lb + 8);
  if (store) {
    HODE_TMEM_ST_X16(t_ahi + 16, (v + 16));
    if (MODE == MLP_X3) HODE_TMEM_ST_X16(t_alo + 16, (lo + 16));
    if (MODE == MLP_MIXED) { HODE_TMEM_ST_X8(t_ahb + 8, (hb + 8)); HODE_TMEM_ST_X8(t_alb + 8, (lb + 8)); }
    tc::wait_st();
  }
  tc::fence_before_sync();
}

// layer-0 operand of one thread: its 9 input features, feature 9 = 1 (its weight column is the layer's bias),
// zero padding to K = 16
template <int MODE>
__device__ __forceinline__ void store_input_operand(uint32_t t_lane, const float* x) {
  if constexpr (is_h16<MODE>()) {
    uint32_t h16[8], hs[8], lb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float x0 = (2 * c < HODE_NN_IN) ? x[2 * c] : (2 * c == HODE_NN_IN ? 1.0f : 0.0f);
      const float x1 = (2 * c + 1 < HODE_NN_IN) ? x[2 * c + 1] : (2 * c + 1 == HODE_NN_IN ? 1.0f : 0.0f);
      split_h16_pair(x0, x1, h16[c], hs[c], lb[c]);
    }
    HODE_TMEM_ST_X8(t_lane + TMH_A16, h16);
    HODE_TMEM_ST_X8(t_lane + TMH_AHS, hs);
    HODE_TMEM_ST_X8(t_lane + TMH_ALB, lb);
    tc::wait_st();
    tc::fence_before_sync();
  } else {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float xv = (k < HODE_NN_IN) ? x[k] : (k == HODE_NN_IN ? 1.0f : 0.0f);
    hi[k] = (__float_as_uint(xv) + 0x1000u) & 0xFFFFE000u;
    lo[k] = __float_as_uint(xv - __uint_as_float(hi[k]));
  }
  HODE_TMEM_ST_X16(t_lane + TM_AHI, hi);
  if (MODE == MLP_X3) HODE_TMEM_ST_X16(t_lane + TM_ALO, lo);
  if (MODE == MLP_MIXED) {
    uint32_t hb[8], lb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      hb[c] = pack_bf16x2(__uint_as_float(hi[2 * c]), __uint_as_float(hi[2 * c + 1]));
      lb[c] = pack_bf16x2(__uint_as_float(lo[2 * c]), __uint_as_float(lo[2 * c + 1]));
    }
    HODE_TMEM_ST_X8(t_lane + TM_AHB, hb);
    HODE_TMEM_ST_X8(t_lane + TM_ALB, lb);
  }
  if (MODE == MLP_MIX3) {
    uint32_t lb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) lb[c] = pack_bf16x2(__uint_as_float(lo[2 * c]), __uint_as_float(lo[2 * c + 1]));
    HODE_TMEM_ST_X8(t_lane + TM3_ALB, lb);
  }
  tc::wait_st();
  tc::fence_before_sync();
  }
}

// The residual MLP for the 128 trajectories of a tile (reference models/nn_residual.py:136-146).
// Every thread of the tile must call this converged.  x: the 9 input features of this thread's
// trajectory; r: the 6 residuals.  `overlap` runs right after the layer-0 MMAs have been issued: per-thread
// work that does not depend on the network (the mechanistic RHS) hides behind their latency.
// (The adjoint, hode_adjoint_tc.cu, builds its own recomputation from the same pieces: store_input_operand,
// issue_layer, epilogue32_to_tmem.)
template <int X3, class F>
__device__ __forceinline__ void mlp_tile(TileCtx& c, const float* x, float* r, F&& overlap) {
  const uint32_t t_lane = c.tmem + c.lane_base;
  const uint32_t img_s = tc::smem_u32(c.img);
  HODE_TL(0);
  store_input_operand<X3>(t_lane, x);
  HODE_TL(1);
  // main AND helper warps: the helpers arrive here only after they have observed the previous
  // call's last mbarrier phase, so the layer-0 commit below cannot flip the barrier a second time
  // under a helper that is still busy (it would then wait for a phase that has already passed)
  // (MLP_MIX3 has no helper warps: every barrier of the tile is over its 128 main threads)
  if (three_tiles<X3>()) tile_sync_main(c); else tile_sync_all(c);
  HODE_TL(2);
  if (c.wq == 0) {
    if (tc::elect_one()) {
      tc::fence_after_sync();
      issue_mlp_layer<X3>(c.tmem, c.t_ones, img_s, c.L, 0);
      tc::mma_commit(c.mma_bar);
    }
    __syncwarp();
  }
  HODE_TL(3);
  overlap();
  __syncwarp();   // per-thread code may leave the warp diverged (division slow paths); the tcgen05 .sync.aligned
                  // instructions below need it converged
  // ---- hidden layers: epilogue of layer l feeds the MMAs of layer l+1 ---------------------------
#pragma unroll 1
  for (int l = 0; l < c.L; ++l) {
    tc::mbar_wait(c.mma_bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
    HODE_TL(10 + 10 * l);
    // the main warp owns accumulator columns [0,32) of its 32 lanes, the helper warp of the same
    // lane quarter columns [32,64) (mlp_tile_helper): the epilogue latency per layer is halved
    if (three_tiles<X3>()) {
      if (X3 == MLP_H16) epilogue64_h16(t_lane); else epilogue64_mix3(t_lane);
      HODE_TL(12 + 10 * l);
      tile_sync_main(c);
    } else if constexpr (X3 == MLP_H16_2T) {
      epilogue32_h16(t_lane, 0u);
      HODE_TL(12 + 10 * l);
      tile_sync_all(c);
    } else {
      uint32_t v0[32], lo[32];
      epilogue32_to_tmem<X3>(t_lane, 0u, v0, lo);
      HODE_TL(12 + 10 * l);
      tile_sync_all(c);
    }
    HODE_TL(13 + 10 * l);
    if (c.wq == 0) {
      if (tc::elect_one()) {
        tc::fence_after_sync();
        issue_mlp_layer<X3>(c.tmem, c.t_ones, img_s, c.L, l + 1);
        tc::mma_commit(c.mma_bar);
      }
      __syncwarp();
    }
    HODE_TL(14 + 10 * l);
  }
  // ---- output layer epilogue: 6 of the 16 accumulator columns ------------------------------------
  tc::mbar_wait(c.mma_bar, c.parity);
  c.parity ^= 1u;
  tc::fence_after_sync();
  HODE_TL(90);
  {
    uint32_t v[8];
    HODE_TMEM_LD_X8(t_lane + TM_D0, v);
    tc::wait_ld();
#pragma unroll
    for (int i = 0; i < NS; ++i) r[i] = __uint_as_float(v[i]);
  }
  HODE_TL(91);
  // The next call overwrites A (tcgen05.st: every MMA has completed) and its layer-0 MMAs write
  // D only after a tile barrier that every thread reaches after its wait::ld above.
}

// Helper warps: the other half of every hidden-layer epilogue.  Must be called once per
// mlp_tile() call of the tile's main warps (same number of tile-wide barriers and mbarrier phases).
template <int X3>
__device__ __forceinline__ void mlp_tile_helper(TileCtx& c) {
  const uint32_t t_lane = c.tmem + c.lane_base;
  tile_sync_all(c);   // pairs with the main warps' barrier before the layer-0 MMAs (see mlp_tile)
#pragma unroll 1
  for (int l = 0; l < c.L; ++l) {
    tc::mbar_wait(c.mma_bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
    if constexpr (X3 == MLP_H16_2T) {
      epilogue32_h16(t_lane, 32u);
    } else {
      uint32_t v[32], lo[32];
      epilogue32_to_tmem<X3>(t_lane, 32u, v, lo);
    }
    tile_sync_all(c);
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
