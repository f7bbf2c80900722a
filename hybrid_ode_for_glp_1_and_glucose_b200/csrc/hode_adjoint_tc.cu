// hode_adjoint_tc.cu — discrete adjoint of the rollout with the MLP forward recomputation, the delta
// back-propagation and the weight gradients on the tcgen05 tensor cores (BASELINE.json north_star item 3).
//
// Same mathematics as hode_adjoint_simt.cu (autograd through the unrolled RK steps with the step
// sizes frozen); what changes is where the 13 248-MAC network products run and how the work is laid out:
//   * schedule (device side, deterministic): trajectories are radix-sorted by accepted-step count,
//     cut into 128-trajectory tiles, and the tiles are dealt to the CTAs longest-first — a tile runs
//     for the maximum step count of its members, and adaptive step counts differ several-fold;
//   * one tile per CTA at a time, three warpgroups with their own register budgets (setmaxnreg):
//     4 main warps (one trajectory per thread: integrator state, mechanistic VJP, stage recurrences),
//     4 helper warps (the other half of every epilogue), 1 MMA-issuer warp;
//   * per accepted step, in reverse:  (1) the forward weight image is bulk-copied into shared
//     memory and the stages are recomputed with hode_tc_mlp.cuh's mlp_tile (3xTF32); every hidden
//     activation goes to a per-CTA stash in global memory, already in the BF16 operand layout of the
//     weight-gradient MMAs.  DP5(4) is first-same-as-last: the rollout saved k1 of every step, so 6
//     stages are recomputed and pulled back per step instead of 7;
//     (2) every stage is pulled back through L + 1 issue phases; the issuer warp issues every MMA chain
//     and runs the operand pipeline (W_l^T and the stashed activations arrive by TMA two phases ahead);
//     delta_{l-1} = u_{l-1} * relu'(a_{l-1}) in the epilogue (stashed ReLU bit masks);
//   * delta is written ONCE per phase, by its owner threads, as a two-term BF16 image in shared memory
//     (16-byte vectors) that serves both products of the phase (csrc/probe/bf16_probe.cu), 3 passes each:
//       u_{l-1} = delta_l W_l            reads it as a K-major A operand against the BF16 image of W_l^T,
//       dW_l += delta_l^T [a_{l-1} | 1]  reads it MN-major (contraction over the tile's 128 trajectories)
//     against the stashed activations; the weight-gradient accumulators stay in TMEM for the whole kernel, so
//     every gradient element is summed in one fixed order (no atomics, bit-reproducible).  The recomputation
//     stays at 3xTF32: two-term BF16 there put the weight gradient 2.6e-4 off float64 autograd;
//   * per-CTA partial gradients -> workspace -> reduce_partials (hode_adjoint_simt.cu), in CTA order.
// Restrictions: nn_hidden == 64, nn_layers <= 4 (the TMEM accumulator map is compiled for them);
// other shapes use the FP32 adjoint.
#include <math.h>

#include <cub/device/device_radix_sort.cuh>

#include "hode_common.cuh"
#include "hode_kernels.h"
#include "hode_tc_mlp.cuh"
#include "hode_tcgen05.cuh"

namespace hode {

namespace {

constexpr int MAXL = 4;       // hidden layers supported by the register accumulators
constexpr int NSTAGE_MAX = 7;

// Butcher tableaux, [solver][..]: 0 = classical RK4, 1 = Dormand-Prince 5(4) (+ stage 7 = FSAL)
__constant__ float kA[2][7][7] = {
    {{0}, {0.5f}, {0.f, 0.5f}, {0.f, 0.f, 1.f}},
    {{0},
     {dp::a21},
     {dp::a31, dp::a32},
     {dp::a41, dp::a42, dp::a43},
     {dp::a51, dp::a52, dp::a53, dp::a54},
     {dp::a61, dp::a62, dp::a63, dp::a64, dp::a65},
     {dp::b1, 0.f, dp::b3, dp::b4, dp::b5, dp::b6}}};
__constant__ float kB[2][7] = {{1.f / 6, 1.f / 3, 1.f / 3, 1.f / 6}, {dp::b1, 0.f, dp::b3, dp::b4, dp::b5, dp::b6, 0.f}};
__constant__ float kC[2][7] = {{0.f, 0.5f, 0.5f, 1.f}, {0.f, dp::c2, dp::c3, dp::c4, dp::c5, 1.f, 1.f}};
__constant__ float kP[7][4] = {{dp::p11, dp::p12, dp::p13, dp::p14}, {0.f, 0.f, 0.f, 0.f},
                               {0.f, dp::p32, dp::p33, dp::p34},     {0.f, dp::p42, dp::p43, dp::p44},
                               {0.f, dp::p52, dp::p53, dp::p54},     {0.f, dp::p62, dp::p63, dp::p64},
                               {0.f, dp::p72, dp::p73, dp::p74}};

// x -> BF16 hi (round to nearest) and BF16 mid = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& mid) {
  uint32_t h, m;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(0.f), "f"(x));
  const float r = x - __uint_as_float(h << 16);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(0.f), "f"(r));
  hi = (uint16_t)(h & 0xFFFFu);
  mid = (uint16_t)(m & 0xFFFFu);
}

// ---- weight gradients on the tensor cores ---------------------------------------------------------
// dW_l += delta_l^T [a_{l-1} | 1] contracts over the 128 trajectories of the tile: an SS-form MMA
// with K = trajectory.  kind::tf32 takes K-major operands only, which would force a transposed
// staging (4-byte scattered stores); kind::f16 takes MN-major operands, so "thread t owns
// trajectory t" writes 8 consecutive features as ONE 16-byte vector (csrc/probe/bf16_probe.cu):
//   element (trajectory t, feature f) at byte (f / 8) * ST_GRP + t * 16 + (f % 8) * 2
//   descriptor: LBO = 128 B (between 8-trajectory groups), SBO = ST_GRP (between 8-feature groups).
// Operands are split in two BF16 terms, x ~= hi + mid (2^-17), and the product takes 3 passes
// mid*hi + hi*mid + hi*hi at the BF16 rate (twice the TF32 rate): measured max error 2.7e-6 of
// sum|ab| per 128-term product.  The activations a_{l-1} are already stored in this form by the
// forward recomputation (stash, hode_tc_mlp.cuh) and come back by bulk copy; only delta is written
// by the threads.  Inputs carry one extra 8-feature group whose first feature is the constant 1
// (its accumulator column is the bias gradient).
// Accumulators live in TMEM for the whole kernel (fp32):
//   hidden layer l = 1..3 : columns DW_H0 + 80 (l-1) .. +72   D[j][k] = dW_l[j][k], column 64 = db_l[j]
//   layer 0               : columns DW_0  .. +16               D[j][k] = dW_0[j][k] (k < 9), column 15 = db_0[j]
//   output layer (transp.): columns DW_O  .. +16               D[k][n] = dW_out[n][k] (n < 6), row 64 = db_out[n]
constexpr uint32_t DW_H0 = 208, DW_0 = 448, DW_O = 464;
// shared memory of the reverse sweep (bytes; everything double-buffered by issue phase):
//   [input operands 2 x (hi, mid) x 9 groups][delta operands 2 x (hi, mid) x 8 groups][W_l^T slots 2 x (hi, lo)]
constexpr int AB_PART = 9 * ST_GRP, AB_BYTES = 2 * AB_PART;
constexpr int DB_BYTES = 2 * ST_PART;
constexpr int WS_BYTES = 2 * 64 * 64 * 2;   // one layer's W^T: BF16 hi, mid
constexpr int N_AB = 3;   // input operands are fetched TWO phases ahead (they stream from HBM: the stash of 148 CTAs exceeds L2)
constexpr int OFF_AB = 0, OFF_DB = OFF_AB + N_AB * AB_BYTES, OFF_WS = OFF_DB + 2 * DB_BYTES;
constexpr int BWD_BYTES = OFF_WS + 2 * WS_BYTES;

// kind::f16 instruction descriptors: D = f32, A = B = BF16; both operands MN-major / both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_k(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= R^T C over the 128 trajectories: 3 passes x 8 k-steps of 16 trajectories.
// r_*: shared-memory byte address of the operand whose features become the ROWS of D, c_*: the COLUMNS.
template <int M, int N>
__device__ __forceinline__ void issue_dw(uint32_t d, uint32_t r_hi, uint32_t r_mid, uint32_t c_hi, uint32_t c_mid,
                                         uint32_t init) {
  constexpr uint32_t idesc = make_idesc_bf16_mn(M, N);
  const uint64_t rh = tc::make_desc(r_hi, 128u, ST_GRP), rm = tc::make_desc(r_mid, 128u, ST_GRP);
  const uint64_t ch = tc::make_desc(c_hi, 128u, ST_GRP), cm = tc::make_desc(c_mid, 128u, ST_GRP);
  // one k-step = 16 trajectories = 256 B: the start-address field (16-byte units) advances by 16
#pragma unroll
  for (int ks = 0; ks < TILE / 16; ++ks) mma_bf16_ss(d, rm + 16u * ks, ch + 16u * ks, idesc, (ks == 0 && init) ? 0u : 1u);
#pragma unroll
  for (int ks = 0; ks < TILE / 16; ++ks) mma_bf16_ss(d, rh + 16u * ks, cm + 16u * ks, idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < TILE / 16; ++ks) mma_bf16_ss(d, rh + 16u * ks, ch + 16u * ks, idesc, 1u);
}

// u[128 x N] = delta[128 x 16 KSTEPS] W: the delta image read as a K-major A operand (K-chunk = one
// 8-feature group, LBO = ST_GRP, 8-trajectory row groups SBO = 128 B) against the K-major BF16 image of
// W^T: element (n, k) at byte (k / 8) * (N * 16) + n * 16 + (k % 8) * 2 (csrc/probe/bf16_probe.cu test 3);
// same two-term split and pass order as the weight gradients.  The first MMA initialises D.
template <int N, int KSTEPS>
__device__ __forceinline__ void issue_u(uint32_t d, uint32_t a_hi, uint32_t a_mid, uint32_t b_hi, uint32_t b_mid) {
  constexpr uint32_t idesc = make_idesc_bf16_k(TILE, N);
  const uint64_t ah = tc::make_desc(a_hi, ST_GRP, 128u), am = tc::make_desc(a_mid, ST_GRP, 128u);
  const uint64_t bh = tc::make_desc(b_hi, (uint32_t)N * 16u, 128u), bm = tc::make_desc(b_mid, (uint32_t)N * 16u, 128u);
  // one k-step = 16 features = 2 K-chunks: A advances 2 ST_GRP bytes, B 2 * N * 16 bytes (16-byte units)
  constexpr uint32_t sa = 2u * ST_GRP / 16u, sb = 2u * (uint32_t)N;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, am + sa * ks, bh + sb * ks, idesc, ks == 0 ? 0u : 1u);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, ah + sa * ks, bm + sb * ks, idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, ah + sa * ks, bh + sb * ks, idesc, 1u);
}

// Issue phases of the reverse sweep.  One stage has L + 1 of them, q = L .. 0:
//   phase q:  u_{q-1} = delta_q W_q  (delta image x W_q^T slot)   and   dW_q += delta_q^T [a_{q-1} | 1]
// (a_{-1} = the stage's input features x, u_{-1} = the cotangent of x).  Phases are numbered through
// the whole kernel (`ph`); phase ph uses buffer ph & 1 of every double-buffered region, and its
// mbarriers complete once per phase, so the parity to wait for is (ph >> 1) & 1.
struct BwdCtx {
  uint8_t* smem;              // reverse-sweep layout (OFF_*)
  const float* wsrc;          // global: this parameter set's transposed weight image
  uint8_t* stash_cta;         // global: this CTA's activation stash
  uint8_t* stage_blk;         // stash block of the current stage, layer 0
  uint64_t* wload_bar;        // [2] W^T slot filled
  uint64_t* aload_bar;        // [N_AB] input operand filled (phase ph uses buffer ph % N_AB, parity (ph / N_AB) & 1)
  uint64_t* gemm_bar;         // weight-gradient MMAs of a phase complete
  uint32_t ph;                // phases issued so far
  int k, k_end, stage_top;    // phase index inside the current sweep, phases in the sweep, stage of phase 0
  uint32_t first;             // 1 until the accumulators have been initialised
  int row;                    // trajectory slot 0..127
};

// wait until every weight-gradient MMA issued so far has completed (at most the last phase can be
// pending: its u-chain, which the epilogue threads have waited for, was issued after all earlier ones)
__device__ __forceinline__ void wait_gemm(const BwdCtx& b) {
  if (b.ph > 0u) tc::mbar_wait(b.gemm_bar, (b.ph - 1u) & 1u);
}

// input operand of sweep phase k1 (global phase ph1): bulk copy of the stashed a_{q-1}, hi and mid
__device__ __forceinline__ void prefetch_A(const BwdCtx& b, int L, int k1, uint32_t ph1) {
  if (k1 >= b.k_end) return;
  const int st = b.stage_top - k1 / (L + 1), q = L - k1 % (L + 1);
  uint64_t* bar = b.aload_bar + (ph1 % N_AB);
  if (q >= 1) {
    const uint8_t* src = b.stash_cta + ((size_t)st * L + (q - 1)) * ST_BLK;
    uint8_t* dst = b.smem + OFF_AB + (ph1 % N_AB) * AB_BYTES;
    tc::mbar_expect_tx(bar, 2u * ST_PART);
    tc::bulk_g2s(dst, src, ST_PART, bar);
    tc::bulk_g2s(dst + AB_PART, src + ST_PART, ST_PART, bar);
  } else {
    tc::mbar_arrive(bar);   // phase 0: the main threads write x themselves
  }
}
// W_q^T of sweep phase k2 (global phase ph2)
__device__ __forceinline__ void prefetch_W(const BwdCtx& b, int L, int k2, uint32_t ph2) {
  if (k2 >= b.k_end) return;
  const int q = L - k2 % (L + 1);
  int off, floats;
  if (q == L) { off = 0; floats = 1024; }
  else if (q >= 1) { off = 1024 + (L - 1 - q) * 4096; floats = 4096; }
  else { off = 1024 + (L - 1) * 4096; floats = 1024; }
  uint64_t* bar = b.wload_bar + (ph2 & 1u);
  tc::mbar_expect_tx(bar, (uint32_t)floats * 4u);
  tc::bulk_g2s(b.smem + OFF_WS + (ph2 & 1u) * WS_BYTES, b.wsrc + off, (uint32_t)floats * 4u, bar);
}

// The MMAs of the reverse sweep are issued by a dedicated warp: tcgen05.mma issue blocks while the
// tensor pipe is busy, and a warp that also runs an epilogue would hold the whole tile back for
// that long.  The 256 epilogue threads only ARRIVE on the named barrier; the issuer warp waits on it.
// (One barrier per phase is race-free: the u-chain whose completion lets a thread move on to its next
// arrival is only issued after the barrier has completed, so arrivals of two phases never mix.)
constexpr int ISSUE_BAR = 3, ISSUE_BAR_THREADS = 2 * TILE + 32;
__device__ __forceinline__ void issue_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(ISSUE_BAR), "n"(ISSUE_BAR_THREADS) : "memory");
}
__device__ __forceinline__ void issue_wait() {
  asm volatile("bar.sync %0, %1;" ::"n"(ISSUE_BAR), "n"(ISSUE_BAR_THREADS) : "memory");
}

// Issuer warp: the MMA chains of one stage of the reverse sweep, mirroring mlp_bwd_tile's phases, and the
// operand pipeline: once the u-chain of phase ph has completed, everything phase ph - 1 read is free
// (its MMAs precede that chain), so the input operand of phase ph + 1 and W^T of phase ph + 2 are fetched.
__device__ __forceinline__ void mlp_bwd_issue(TileCtx& c, BwdCtx& b) {
  const int L = c.L;
  const uint32_t m_d = c.tmem + TM_D0;
  const uint32_t base = tc::smem_u32(b.smem);
  // (runtime loops here and in mlp_bwd_tile: unrolled over the phases the kernel outgrows the
  // instruction cache — three roles run three different code streams on one SM)
#pragma unroll 1
  for (int q = L; q >= 0; --q) {
    const uint32_t buf = b.ph & 1u, par = (b.ph >> 1) & 1u;
    const uint32_t ws = base + OFF_WS + buf * WS_BYTES;
    const uint32_t abuf = b.ph % N_AB, apar = (b.ph / N_AB) & 1u;
    const uint32_t a_hi = base + OFF_AB + abuf * AB_BYTES, a_mid = a_hi + AB_PART;
    const uint32_t d_hi = base + OFF_DB + buf * DB_BYTES, d_mid = d_hi + ST_PART;
    issue_wait();
    tc::mbar_wait(b.wload_bar + buf, par);
    if (tc::elect_one()) {
      tc::fence_after_sync();
      if (q == L) issue_u<H, 1>(m_d, d_hi, d_mid, ws, ws + 2048u);          // u_{L-1} = delta_L W_out   (K = 16)
      else if (q >= 1) issue_u<H, 4>(m_d, d_hi, d_mid, ws, ws + 8192u);     // u_{q-1} = delta_q W_q
      else issue_u<16, 4>(m_d, d_hi, d_mid, ws, ws + 2048u);                // g_x = delta_0 W_0
      tc::mma_commit(c.mma_bar);
    }
    __syncwarp();
    const uint32_t u_parity = c.parity;
    c.parity ^= 1u;
    tc::mbar_wait(b.aload_bar + abuf, apar);
    if (tc::elect_one()) {
      // dW_out^T [in k][out n] = [a_{L-1} | 1]^T delta_L;  dW_q += delta_q^T [a_{q-1} | 1];  dW_0 += delta_0^T [x | 1]
      if (q == L) issue_dw<128, 16>(c.tmem + DW_O, a_hi, a_mid, d_hi, d_mid, b.first);
      else if (q >= 1) issue_dw<64, 72>(c.tmem + DW_H0 + 80u * (uint32_t)(q - 1), d_hi, d_mid, a_hi, a_mid, b.first);
      else issue_dw<64, 16>(c.tmem + DW_0, d_hi, d_mid, a_hi, a_mid, b.first);
      tc::mma_commit(b.gemm_bar);
    }
    __syncwarp();
    tc::mbar_wait(c.mma_bar, u_parity);
    if ((threadIdx.x & 31) == 0) {
      prefetch_A(b, L, b.k + 2, b.ph + 2u);   // its buffer was last read by phase ph - 1, complete before this u-chain
      prefetch_W(b, L, b.k + 2, b.ph + 2u);
    }
    __syncwarp();
    b.ph += 1u;
    b.k += 1;
  }
  b.first = 0u;
}

// ---- MLP backward for one stage (tile-collective: all 256 epilogue threads) ---------------------------
// MAIN threads own accumulator columns [0,32) of their trajectory, helpers [32,64).
// g6 (main): cotangent of the 6 network outputs; x9 (main): the stage's input features;
// gx (main, out): cotangent of the 9 input features.
// Notation: delta_l = cotangent of the pre-activation of layer l (l = 0..L-1), delta_L = g.
// `overlap` runs after phase L has been handed to the issuer: per-thread work that does not depend on
// the network's cotangents (the mechanistic VJP) hides behind the first MMA chain.
template <bool MAIN, class F>
__device__ __forceinline__ void mlp_bwd_tile(TileCtx& c, BwdCtx& b, const float* x9, const float* g6, float* gx,
                                             F&& overlap) {
  constexpr int half = MAIN ? 0 : 32;
  constexpr int hidx = MAIN ? 0 : 1;
  const int L = c.L;
  const uint32_t t_d = c.tmem + c.lane_base + TM_D0 + half;

  // ---- prologue: delta_L = g; phase L (its input operand and W^T slot are already on their way) ------
  HODE_TL(220);
  if (MAIN) {
    float d[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = (j < NS) ? g6[j] : 0.f;
    uint8_t* db = b.smem + OFF_DB + (b.ph & 1u) * DB_BYTES + b.row * 16;
    uint4 vh, vm;
    bf16_split8(d, vh, vm);
    *reinterpret_cast<uint4*>(db) = vh;
    *reinterpret_cast<uint4*>(db + ST_PART) = vm;
    *reinterpret_cast<uint4*>(db + ST_GRP) = make_uint4(0u, 0u, 0u, 0u);   // N = 16: features 8..15 are zero
    *reinterpret_cast<uint4*>(db + ST_GRP + ST_PART) = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async();
  }
  tc::fence_before_sync();
  HODE_TL(221);
  issue_arrive();   // the issuer warp launches phase L (mlp_bwd_issue) once all 256 threads are here
  b.ph += 1u;
  b.k += 1;
  HODE_TL(222);
  overlap();
  __syncwarp();   // reconverge after per-thread code: tcgen05 .sync.aligned instructions follow

  // ---- p = L .. 1: u_{p-1} arrives, delta_{p-1} = u_{p-1} * relu'(a_{p-1}) goes out for phase p-1 ------
#pragma unroll 1
  for (int p = L; p >= 1; --p) {
    const uint32_t mask = reinterpret_cast<const uint32_t*>(b.stage_blk + (size_t)(p - 1) * ST_BLK + 2 * ST_PART)[hidx * TILE + b.row];
    tc::mbar_wait(c.mma_bar, c.parity);
    c.parity ^= 1u;
    tc::fence_after_sync();
    HODE_TL(230 + 10 * p);
    uint32_t u[32];
    HODE_TMEM_LD_X32(t_d, u);
    tc::wait_ld();
    float d[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) d[j] = ((mask >> j) & 1u) ? __uint_as_float(u[j]) : 0.f;
    HODE_TL(231 + 10 * p);
    {
      uint8_t* db = b.smem + OFF_DB + (b.ph & 1u) * DB_BYTES + (hidx * 4) * ST_GRP + b.row * 16;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 vh, vm;
        bf16_split8(d + 8 * g, vh, vm);
        *reinterpret_cast<uint4*>(db + g * ST_GRP) = vh;
        *reinterpret_cast<uint4*>(db + g * ST_GRP + ST_PART) = vm;
      }
    }
    if (p == 1 && MAIN) {   // inputs of layer 0: the 9 stage features, zero padding, feature 15 = 1 (bias column)
      uint8_t* ab = b.smem + OFF_AB + (b.ph % N_AB) * AB_BYTES + b.row * 16;
      const float xb[8] = {x9[8], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
      uint4 vh, vm;
      bf16_split8(x9, vh, vm);
      *reinterpret_cast<uint4*>(ab) = vh;
      *reinterpret_cast<uint4*>(ab + AB_PART) = vm;
      bf16_split8(xb, vh, vm);
      *reinterpret_cast<uint4*>(ab + ST_GRP) = vh;
      *reinterpret_cast<uint4*>(ab + ST_GRP + AB_PART) = vm;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    HODE_TL(233 + 10 * p);
    issue_arrive();
    b.ph += 1u;
    b.k += 1;
    HODE_TL(234 + 10 * p);
  }
  // ---- final phase: g_x -------------------------------------------------------------------------
  tc::mbar_wait(c.mma_bar, c.parity);
  c.parity ^= 1u;
  if (MAIN) {
    tc::fence_after_sync();
    uint32_t v[16];
    HODE_TMEM_LD_X16(c.tmem + c.lane_base + TM_D0, v);
    tc::wait_ld();
#pragma unroll
    for (int k = 0; k < HODE_NN_IN; ++k) gx[k] = __uint_as_float(v[k]);
  }
  b.first = 0u;
}

// closed-form VJP of f_physio — same formulas as hode_adjoint_simt.cu::rhs_mech_vjp
__device__ __forceinline__ void mech_vjp(const Theta& p, const float* y, float GD, bool gd_present,
                                         const float* c, float* gy, float* gth) {
  const float G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const float Pi = 1.0f + p.rho * GLP1;
  const float dG = G - p.G_b, dI = I - p.I_b, dGlu = Glu - p.Glu_b;
  const float inv_e = 1.0f / (p.EC_50 + GLP1);
  const float frac_e = GLP1 * inv_e;
  const float ge = p.E_max * frac_e;
  const float inv_m = 1.0f / (p.K_m + G);
  float r = 0.f, dr_du = 0.f, dr_dv = 0.f, u = 0.f;
  const float v = p.igd_pow;
  if (gd_present) {
    u = powf(GD, p.g);
    const float inv = 1.0f / (v + u);
    r = u * inv;
    dr_du = v * inv * inv;
    dr_dv = -u * inv * inv;
  }
  const float k_GE = p.k_GE0 * (1.0f - r);
  const float lin5 = -p.p_7 - p.p_8 * I + p.p_9 * G;
  gy[0] += c[1] * Pi * p.a_GI + c[3] * p.V_max * p.K_m * inv_m * inv_m + c[5] * FFA * p.p_9 - c[0] * k_GE;
  gy[1] += -c[1] * p.k_I - c[5] * FFA * p.p_8 - 0.01f * c[0];
  gy[2] += -c[2] * ge + 0.005f * c[0];
  gy[3] += c[1] * p.rho * p.a_GI * dG - c[2] * dGlu * p.E_max * p.EC_50 * inv_e * inv_e - c[3] * p.k_L;
  gy[5] += c[5] * lin5;
  gth[0] += c[1] * Pi * dG;
  gth[1] += -c[1] * dI;
  gth[2] += c[1] * GLP1 * p.a_GI * dG;
  gth[3] += -c[1] * Pi * p.a_GI;
  gth[4] += c[1] * p.k_I + 0.01f * c[0];
  gth[5] += -c[2] * dGlu * frac_e;
  gth[6] += c[2] * dGlu * p.E_max * GLP1 * inv_e * inv_e;
  gth[7] += c[2] * ge - 0.005f * c[0];
  gth[8] += c[3] * G * inv_m;
  gth[9] += -c[3] * p.V_max * G * inv_m * inv_m;
  gth[10] += -c[3] * GLP1;
  gth[11] += -c[0] * G * (1.0f - r);
  if (gd_present) {
    const float g_r = c[0] * p.k_GE0 * G;
    gth[12] += g_r * dr_dv * p.g * powf(p.IGD_50, p.g - 1.0f);
    float dg = dr_dv * v * logf(p.IGD_50);
    if (GD > 0.f) dg += dr_du * u * logf(GD);
    gth[13] += g_r * dg;
  }
  gth[14] += -c[5] * FFA;
  gth[15] += -c[5] * FFA * I;
  gth[16] += c[5] * FFA * G;
}

}  // namespace

struct AdjTcArgs {
  RolloutArgs R;
  const float* grad_traj;  // [S,B,T,6]
  float* grad_y0;          // [S,B,6] or nullptr
  float* partials;         // [gridDim.y * gridDim.x][P + 17]
  float* stash;            // [7][L][64][NT] activation stash
  const float* img_fwd;    // [S][fwd_floats]
  const float* img_bwd;    // [S][bwd_floats]
  int fwd_floats, bwd_floats;
  // schedule (adj_schedule_kernel): trajectories sorted by accepted-step count, tiles handed to CTAs
  const int32_t* perm;        // [S*B] unit index of sorted slot q (within its parameter set)
  const int32_t* sched_off;   // [S][grid_x + 1] range of sched_tiles owned by CTA (s, x)
  const int32_t* sched_tiles; // [S][n_tiles] tile indices grouped by owner
  int n_tiles;
};

// ---------------------------------------------------------------------------------------------------
// grid = (ctas per parameter set, S), block = 256 (4 main + 4 helper warps), 1 CTA / SM
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(3 * TILE, 1) rollout_bwd_tc_kernel(const AdjTcArgs G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ __align__(8) uint64_t wload_bar[2];
  __shared__ __align__(8) uint64_t aload_bar[N_AB];
  __shared__ __align__(8) uint64_t gemm_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_nmax;

  const RolloutArgs& A = G.R;
  const int tid = threadIdx.x, lane_id = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // warps 0-3: main (one trajectory per thread), 4-7: helpers (other half of every epilogue), 8: MMA issuer
  const bool main_role = warp < 4, helper = warp >= 4 && warp < 8, issuer = warp == 8;
  const int wq = warp & 3, row = tid & 127;
  const int s = blockIdx.y, T = A.T, L = A.L;
  const int solver = A.solver == HODE_SOLVER_RK4 ? 0 : 1;
  const int N = solver == 0 ? 4 : 7;
  // DP5(4) is first-same-as-last: stage 1 of step n+1 IS stage 7 of step n.  The rollout saved its
  // value (save_k), so the recomputation starts at stage 2, and the reverse sweep pulls the shared
  // evaluation back once, as stage 7 of the earlier step, with both cotangents added (`carry`).
  // Stage 1 of the very first step is pulled back by one extra iteration: a zero-length step at
  // (t0, y0) whose stage 7 receives the carry.
  const bool fsal = solver == 1 && A.save_k1 != 0;
  const int i0 = fsal ? 1 : 0;

  // shared memory: the forward weight image during the recomputation; during the reverse sweep the
  // same bytes hold the double-buffered operand images and W^T slots (OFF_AB / OFF_DB / OFF_WS)
  float* img = reinterpret_cast<float*>(smem_raw);
  const int bwd_cap = BWD_BYTES / 4;
  const int img_cap = ((G.fwd_floats > bwd_cap ? G.fwd_floats : bwd_cap) + 255) & ~255;
  // next step's record of every main thread, fetched while the current reverse sweep runs (the streaming
  // stash traffic evicts an L2 prefetch long before it is used: the plain load cost 4.8 k cycles per step)
  float* rec_sh = img + img_cap;                       // [128][HODE_REC_FLOATS]
  float* t_sh_buf = rec_sh + TILE * HODE_REC_FLOATS;
  float* red = img;   // [17][128] theta-gradient reduction scratch at the very end
  if (tid == 0) {
    tc::mbar_init(&mma_bar, 1);
    tc::mbar_init(&load_bar, 1);
    tc::mbar_init(&wload_bar[0], 1);
    tc::mbar_init(&wload_bar[1], 1);
    for (int i = 0; i < N_AB; ++i) tc::mbar_init(&aload_bar[i], 1);
    tc::mbar_init(&gemm_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  const float* t_shared = nullptr;
  if (!A.t_per_traj && T <= HODE_SIMT_MAX_SHARED_T) {
    for (int i = tid; i < T; i += blockDim.x) t_sh_buf[i] = A.t_obs[i];
    t_shared = t_sh_buf;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  TileCtx c;
  c.img = img;
  c.mma_bar = &mma_bar;
  c.tmem = tmem_base_s;
  c.lane_base = (uint32_t)(wq * 32) << 16;
  c.parity = 0;
  c.bar_id = 1;
  c.bar_all = 2;
  c.wq = wq;
  c.L = L;
  {
    uint32_t ones[8] = {0x3F800000u, 0x3F800000u, 0u, 0u, 0u, 0u, 0u, 0u};
    HODE_TMEM_ST_X8(c.tmem + c.lane_base + TM_ONES, ones);
    tc::wait_st();
  }
  uint32_t load_parity = 0;
  // this CTA's activation stash: [stage][layer] blocks of ST_BLK bytes (hode_tc_mlp.cuh)
  const size_t stage_stride = (size_t)L * ST_BLK;
  uint8_t* stash0 = reinterpret_cast<uint8_t*>(G.stash) +
                    ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)NSTAGE_MAX * stage_stride;

  BwdCtx bc;
  bc.smem = smem_raw;
  bc.wsrc = G.img_bwd + (size_t)s * G.bwd_floats;
  bc.stash_cta = stash0;
  bc.stage_blk = stash0;
  bc.wload_bar = wload_bar;
  bc.aload_bar = aload_bar;
  bc.gemm_bar = &gemm_bar;
  bc.ph = 0u;
  bc.k = 0; bc.k_end = 0; bc.stage_top = 0;
  bc.first = 1u;
  bc.row = row;

  const Theta th = load_theta(A.theta + (size_t)s * HODE_N_THETA);
  const bool gd_present = A.in_mode[HODE_CH_GD] != HODE_IN_ABSENT;
  const int nsub = A.n_substeps > 0 ? A.n_substeps : 1;
  float gth[HODE_N_THETA];
#pragma unroll
  for (int i = 0; i < HODE_N_THETA; ++i) gth[i] = 0.f;

  // swap the shared-memory weight image (forward <-> transposed); every thread calls it
  auto load_image = [&](const float* src, int floats) {
    // the weight-gradient MMAs of the previous reverse sweep still read the staging arrays
    wait_gemm(bc);
    tc::fence_before_sync();
    __syncthreads();   // nobody still reads the old contents (all MMAs that did have been waited for)
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)floats * 4u;
      tc::mbar_expect_tx(&load_bar, bytes);
      tc::bulk_g2s(img, src, bytes, &load_bar);
    }
    tc::mbar_wait(&load_bar, load_parity);
    load_parity ^= 1u;
  };
  const float* fwd_src = G.img_fwd + (size_t)s * G.fwd_floats;
  // start of a reverse sweep: the forward image is dead; (re)write the constant features of the
  // input staging (feature 64 = 1 -> bias-gradient column, 65..79 = 0)
  // input-operand group 8 = [1, 0 x 7]: its accumulator column is the bias gradient), and start
  // the operand pipeline of the sweep's first phases.  The stash was written with ordinary global
  // stores and is read back by bulk copies: cross-proxy fence before the barrier.
  auto begin_reverse = [&]() {
    tc::fence_proxy_async_all();
    tc::fence_before_sync();
    __syncthreads();
    if (main_role) {
#pragma unroll
      for (int bf = 0; bf < N_AB; ++bf) {
        uint8_t* g8 = smem_raw + OFF_AB + bf * AB_BYTES + 8 * ST_GRP + row * 16;
        *reinterpret_cast<uint4*>(g8) = make_uint4(0x00003F80u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(g8 + AB_PART) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    bc.k = 0;
    bc.k_end = (N - i0) * (L + 1);
    bc.stage_top = N - 1;
    if (tid == 0) {
      prefetch_W(bc, L, 0, bc.ph);
      prefetch_A(bc, L, 0, bc.ph);
      prefetch_A(bc, L, 1, bc.ph + 1u);
      prefetch_W(bc, L, 1, bc.ph + 1u);
    }
  };

  const int32_t* my_tiles = G.sched_tiles + (size_t)s * G.n_tiles;
  const int tile_beg = G.sched_off[(size_t)s * (gridDim.x + 1) + blockIdx.x];
  const int tile_end = G.sched_off[(size_t)s * (gridDim.x + 1) + blockIdx.x + 1];
  // nmax of a tile: every thread of the CTA passes here (two CTA-wide barriers)
  auto tile_nmax = [&](int n_) -> int {
    if (tid == 0) s_nmax = 0;
    __syncthreads();
    if (n_ > 0) atomicMax(&s_nmax, n_);
    __syncthreads();
    return s_nmax;
  };

  // Three roles, each entirely inside its own branch so that ptxas allocates registers against
  // the role's budget: the main warpgroup carries the per-trajectory integrator and adjoint state
  // (it takes the registers the other two warpgroups give back), the helper warpgroup only runs
  // epilogue halves, the third warpgroup holds the MMA-issuer warp (its other three warps idle
  // through the CTA-wide barriers: setmaxnreg works on whole warpgroups).
  if (helper) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;" ::: "memory");
  for (int tk = tile_beg; tk < tile_end; ++tk) {
    const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
    for (int it = 0; it < n_iter; ++it) {
      load_image(fwd_src, G.fwd_floats);
#pragma unroll 1
      for (int i = i0; i < N; ++i) mlp_tile_helper<true, true>(c, stash0 + (size_t)i * stage_stride, row);
      begin_reverse();
#pragma unroll 1
      for (int i = N - 1; i >= i0; --i) {
        bc.stage_blk = stash0 + (size_t)i * stage_stride;
        mlp_bwd_tile<false>(c, bc, nullptr, nullptr, nullptr, [] {});
      }
    }
  }
  wait_gemm(bc);
  tc::fence_before_sync();
  __syncthreads();   // accumulators complete
  __syncthreads();   // theta-gradient scratch written
  tc::fence_before_sync();
  __syncthreads();   // TMEM may be released
  } else if (!main_role) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
  for (int tk = tile_beg; tk < tile_end; ++tk) {
    const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
    for (int it = 0; it < n_iter; ++it) {
      load_image(fwd_src, G.fwd_floats);
      if (issuer) {
#pragma unroll 1
        for (int i = i0; i < N; ++i) mlp_fwd_issue<true>(c);
      }
      begin_reverse();
      if (issuer) {
#pragma unroll 1
        for (int i = N - 1; i >= i0; --i) mlp_bwd_issue(c, bc);
      }
    }
  }
  wait_gemm(bc);
  tc::fence_before_sync();
  __syncthreads();   // accumulators complete
  __syncthreads();   // theta-gradient scratch written
  tc::fence_before_sync();
  __syncthreads();   // TMEM may be released
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory");
  for (int tk = tile_beg; tk < tile_end; ++tk) {
    const long q = (long)my_tiles[tk] * TILE + row;   // slot in the step-count-sorted order
    const bool valid = q < A.B;
    const long unit = valid ? (long)G.perm[(size_t)s * A.B + q] : (long)s * A.B;
    const long bs = unit - (long)s * A.B;
    int n = valid ? A.save_n[unit] : 0;
    const bool ok = valid && n >= 0;
    if (n < 0) n = 0;
    const int nmax = tile_nmax(n);

    TrajInputs in;
    in.T = T; in.cur = 0;
    in.t_obs = A.t_per_traj ? A.t_obs + bs * T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      in.mode[ch] = A.in_mode[ch];
      in.u[ch] = in.mode[ch] == HODE_IN_SERIES ? A.u[ch] + bs * T
               : in.mode[ch] == HODE_IN_CONST ? A.u[ch] + bs : nullptr;
    }
    const float* gtraj = G.grad_traj + (size_t)unit * T * NS;
    const double t_bound = (double)in.t_obs[T - 1];
    float lam[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) lam[i] = 0.f;
    int ei = T - 1;

    float carry[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) carry[i] = 0.f;
    double t_next = 0.0;
    const int n_iter = nmax + (fsal ? 1 : 0);
    for (int it = 0; it < n_iter; ++it) {
      // =========================== forward recomputation =========================================
      HODE_TL(200);
      load_image(fwd_src, G.fwd_floats);
      HODE_TL(201);
      const int sidx = n - 1 - it;
      const bool real = ok && sidx >= 0;
      const bool act = real || (fsal && ok && sidx == -1);   // sidx == -1: the zero-length step at (t0, y0)
      double t = (double)in.t_obs[0], t_new = t, h = 0.0;
      float y[NS], k1[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) { y[i] = 0.f; k1[i] = 0.f; }
      if (act && !real && n > 0) {   // the zero-length step: y0 = the state the first record starts from
        double t_;
        float h_;
        step_rec_load(step_rec(A, unit, 0), t_, h_, y, nullptr);
      }
      if (real) {
        const float* rec = step_rec(A, unit, sidx);
        if (it > 0) {   // fetched into shared memory during the previous iteration (it was `real` there too)
          tc::cp_async_wait_all();
          rec = rec_sh + row * HODE_REC_FLOATS;
        }
        float h_rec;
        step_rec_load(rec, t, h_rec, y, fsal ? k1 : nullptr);
        if (solver == 0) {
          h = (double)h_rec;
          t_new = t + h;
        } else {
          t_new = (sidx + 1 < n) ? t_next : t_bound;   // the next record's start time, seen one iteration ago
          h = t_new - t;
        }
        t_next = t;
      }
      const float hf = (float)h;
#ifdef HODE_TIMELINE
      if (hf + y[0] + k1[5] == 123456.f) HODE_TL(299);   // (forces the loads to complete before the next mark)
#endif
      HODE_TL(205);
      // Inputs of this step as one linear piece per channel (value = c_v1 + alpha * c_dv): valid when a
      // step cannot cross an input kink (RK4, kink clipping, or no series input), see
      // hode_rollout_tc.cu::lane_cache_inputs.  Otherwise every stage looks its interval up.
      const bool cached = solver == 0 || A.kink_mode == HODE_KINK_CLIP || !any_series(in);
      float c_t1 = 0.f, c_dt = 1.f, c_v1[3], c_dv[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        c_v1[ch] = in.mode[ch] == HODE_IN_CONST ? in.u[ch][0] : 0.f;
        c_dv[ch] = 0.f;
      }
      {
        int lo = 0, hi = T;
        const float t32 = (float)t;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (in.t_obs[mid] < t32) lo = mid + 1; else hi = mid; }
        in.cur = lo > 0 ? lo - 1 : 0;
        if (cached && any_series(in) && T >= 2) {
          int i0 = (lo < T && in.t_obs[lo] == t32) ? lo : lo - 1;
          i0 = i0 < 0 ? 0 : (i0 > T - 2 ? T - 2 : i0);
          c_t1 = in.t_obs[i0];
          c_dt = in.t_obs[i0 + 1] - c_t1;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            if (in.mode[ch] != HODE_IN_SERIES) continue;
            const float v1 = in.u[ch][i0], v2 = in.u[ch][i0 + 1];
            c_v1[ch] = v1;
            c_dv[ch] = v2 - v1;
          }
        }
      }
#ifdef HODE_TIMELINE
      if (c_v1[0] + c_dv[1] + c_dt == 123456.f) HODE_TL(299);
#endif
      HODE_TL(206);
      // stage derivatives / cotangents are indexed statically (predicated selects) so that they
      // live in registers rather than in local memory
      float k[NSTAGE_MAX][NS], tv[NSTAGE_MAX], gdv[NSTAGE_MAX];
#pragma unroll
      for (int i = 0; i < NSTAGE_MAX; ++i) {
        tv[i] = 0.f; gdv[i] = 0.f;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) k[i][cc] = 0.f;
      }
#pragma unroll
      for (int cc = 0; cc < NS; ++cc) k[0][cc] = k1[cc];
#pragma unroll 1
      for (int i = i0; i < N; ++i) {
        float ys[NS];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          float a_ = 0.f;
#pragma unroll
          for (int j = 0; j < NSTAGE_MAX - 1; ++j) a_ = fmaf(kA[solver][i][j], k[j][cc], a_);   // zero for j >= i
          ys[cc] = fmaf(hf, a_, y[cc]);
        }
        const float ci = kC[solver][i];
        const double te = (i == 0) ? t : (ci == 1.0f ? t_new : t + (double)ci * h);
        const float t32 = (float)te;
        float meal, tvns, gd;
        if (cached) {
          const float alpha = __fdiv_rn(t32 - c_t1, c_dt);
          meal = __fadd_rn(c_v1[HODE_CH_MEAL], __fmul_rn(alpha, c_dv[HODE_CH_MEAL]));
          tvns = __fadd_rn(c_v1[HODE_CH_TVNS], __fmul_rn(alpha, c_dv[HODE_CH_TVNS]));
          gd = __fadd_rn(c_v1[HODE_CH_GD], __fmul_rn(alpha, c_dv[HODE_CH_GD]));
        } else {
          int idx = 0;
          if (any_series(in)) idx = grid_index_from(in, t32, in.cur);
          meal = input_channel(in, HODE_CH_MEAL, t32, idx);
          tvns = input_channel(in, HODE_CH_TVNS, t32, idx);
          gd = input_channel(in, HODE_CH_GD, t32, idx);
        }
        float x[HODE_NN_IN], r[NS], d[NS];
        x[0] = t32;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) x[1 + cc] = ys[cc];
        x[7] = ys[3];
        x[8] = tvns;
        __syncwarp();
        HODE_TL(207);
        mlp_tile<true, true>(c, x, r, stash0 + (size_t)i * stage_stride, row,
                             [&] { rhs_mech(th, ys, meal, gd, gd_present, d); });
#pragma unroll
        for (int jj = 0; jj < NSTAGE_MAX; ++jj) {
          if (jj == i) {
            tv[jj] = tvns; gdv[jj] = gd;
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) k[jj][cc] = __fadd_rn(d[cc], r[cc]);
          }
        }
      }
      // =========================== reverse sweep ===================================================
      HODE_TL(202);
      // the next iteration's step record: fetch it into shared memory while this sweep runs
      if (real && sidx > 0) {
        const float* nxt = step_rec(A, unit, sidx - 1);
#pragma unroll
        for (int q4 = 0; q4 < HODE_REC_FLOATS / 4; ++q4) tc::cp_async16(rec_sh + row * HODE_REC_FLOATS + 4 * q4, nxt + 4 * q4);
        tc::cp_async_commit();
      }
      begin_reverse();
      HODE_TL(203);
      float gy[NS], gk[NSTAGE_MAX][NS];
#pragma unroll
      for (int i = 0; i < NSTAGE_MAX; ++i)
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) gk[i][cc] = 0.f;
      if (solver == 0) {
        if (act && (sidx + 1) % nsub == 0) {
          const float* g = gtraj + (size_t)((sidx + 1) / nsub) * NS;
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) lam[cc] += g[cc];
        }
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) gy[cc] = act ? lam[cc] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) gk[j][cc] = hf * kB[0][j] * gy[cc];
      } else {
        float gnew[NS];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) { gnew[cc] = act ? lam[cc] : 0.f; gy[cc] = 0.f; }
        if (act) {
          while (ei >= 0 && (double)in.t_obs[ei] > t) {
            const double te = (double)in.t_obs[ei];
            const float* g = gtraj + (size_t)ei * NS;
            if (te >= t_new) {
              if (te == t_new) {
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) gnew[cc] += g[cc];
              }
            } else {
              const float xq = (float)((te - t) / h);
#pragma unroll
              for (int cc = 0; cc < NS; ++cc) gy[cc] += g[cc];
#pragma unroll
              for (int i = 0; i < 7; ++i) {
                const float wgt = hf * xq * fmaf(xq, fmaf(xq, fmaf(xq, kP[i][3], kP[i][2]), kP[i][1]), kP[i][0]);
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) gk[i][cc] = fmaf(wgt, g[cc], gk[i][cc]);
              }
            }
            --ei;
          }
        }
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          gy[cc] += gnew[cc];
#pragma unroll
          for (int j = 0; j < 6; ++j) gk[j][cc] = fmaf(hf * kB[1][j], gnew[cc], gk[j][cc]);
          gk[6][cc] += carry[cc];   // stage 1 of the next step = this step's stage 7 (zero unless fsal)
        }
      }
#pragma unroll 1
      for (int i = N - 1; i >= i0; --i) {
        float ys[NS], gys[NS], gki[NS];
        float tvi = 0.f, gdi = 0.f;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          float a_ = 0.f;
#pragma unroll
          for (int j = 0; j < NSTAGE_MAX - 1; ++j) a_ = fmaf(kA[solver][i][j], k[j][cc], a_);
          ys[cc] = fmaf(hf, a_, y[cc]);
          gys[cc] = 0.f;
          gki[cc] = 0.f;
        }
#pragma unroll
        for (int jj = 0; jj < NSTAGE_MAX; ++jj) {
          if (jj == i) {
            tvi = tv[jj]; gdi = gdv[jj];
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) gki[cc] = gk[jj][cc];
          }
        }
        const float ci = kC[solver][i];
        const double te = (i == 0) ? t : (ci == 1.0f ? t_new : t + (double)ci * h);
        float x[HODE_NN_IN], gx[HODE_NN_IN];
        x[0] = (float)te;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) x[1 + cc] = ys[cc];
        x[7] = ys[3];
        x[8] = tvi;
        bc.stage_blk = stash0 + (size_t)i * stage_stride;
        __syncwarp();
        HODE_TL(210);
        mlp_bwd_tile<true>(c, bc, x, gki, gx, [&] { mech_vjp(th, ys, gdi, gd_present, gki, gys, gth); });
        HODE_TL(211);
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) gys[cc] += gx[1 + cc];
        gys[3] += gx[7];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          gy[cc] += gys[cc];
#pragma unroll
          for (int j = 0; j < NSTAGE_MAX - 1; ++j) gk[j][cc] = fmaf(hf * kA[solver][i][j], gys[cc], gk[j][cc]);
        }
      }
      if (act) {
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) lam[cc] = gy[cc];
      }
      if (fsal) {
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) carry[cc] = real ? gk[0][cc] : 0.f;
      }
      HODE_TL(204);
    }
    {
      if (ok) {
        if (solver == 0) {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) lam[cc] += gtraj[cc];
        } else {
          const double t0 = (double)in.t_obs[0];
          for (; ei >= 0; --ei) {
            if ((double)in.t_obs[ei] > t0) continue;
            const float* g = gtraj + (size_t)ei * NS;
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) lam[cc] += g[cc];
          }
        }
      }
      if (G.grad_y0 && valid) {
        float* o = G.grad_y0 + (size_t)unit * NS;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) o[cc] = ok ? lam[cc] : 0.f;
      }
    }
  }

  // ---- per-CTA partial gradients: TMEM accumulators -> workspace ---------------------------------
  wait_gemm(bc);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  float* out = G.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(A.P + HODE_N_THETA);
  const bool have = bc.first == 0u;   // false: this CTA processed no step, the accumulators were never written
  const int offo = 640 + (L - 1) * 4160;
  // M = 64 accumulators (layer 0 and the hidden layers): row j of D lives in TMEM lane 32 (j / 16) + j % 16
  // (csrc/probe/adj_probe.cu), i.e. in the first 16 lanes of every main warp.  The loads are warp-wide.
  if (main_role) {
    const int j = 16 * wq + lane_id;
    const bool owner = lane_id < 16;
    uint32_t v[16];
    // layer 0: D[j][k] (k < 9), column 15 = db_0[j]
    HODE_TMEM_LD_X16(c.tmem + c.lane_base + DW_0, v);
    tc::wait_ld();
    if (owner) {
#pragma unroll
      for (int k = 0; k < HODE_NN_IN; ++k) out[j * HODE_NN_IN + k] = have ? __uint_as_float(v[k]) : 0.f;
      out[576 + j] = have ? __uint_as_float(v[15]) : 0.f;
    }
#pragma unroll
    for (int l = 1; l < MAXL; ++l) {
      if (l >= L) continue;
      const int off = 640 + (l - 1) * 4160;
#pragma unroll
      for (int cidx = 0; cidx < 5; ++cidx) {
        HODE_TMEM_LD_X16(c.tmem + c.lane_base + DW_H0 + 80u * (uint32_t)(l - 1) + 16u * (uint32_t)cidx, v);
        tc::wait_ld();
        if (owner) {
          if (cidx < 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) out[off + j * H + cidx * 16 + i] = have ? __uint_as_float(v[i]) : 0.f;
          } else {
            out[off + 4096 + j] = have ? __uint_as_float(v[0]) : 0.f;
          }
        }
      }
    }
  }
  // output layer (transposed, M = 128: lane = row): D[k][n] = dW_out[n][k] for the input features k < 64
  if (main_role && wq < 2) {
    const int j = row;
    uint32_t v[16];
    HODE_TMEM_LD_X16(c.tmem + c.lane_base + DW_O, v);
    tc::wait_ld();
#pragma unroll
    for (int nn = 0; nn < NS; ++nn) out[offo + nn * H + j] = have ? __uint_as_float(v[nn]) : 0.f;
  }
  if (main_role && wq == 2) {   // TMEM lane 64: the constant-1 input feature -> db_out (warp-wide load)
    uint32_t v[16];
    HODE_TMEM_LD_X16(c.tmem + c.lane_base + DW_O, v);
    tc::wait_ld();
    if (lane_id == 0) {
#pragma unroll
      for (int nn = 0; nn < NS; ++nn) out[offo + 384 + nn] = have ? __uint_as_float(v[nn]) : 0.f;
    }
  }
  if (main_role) {
    // deterministic reduction of the theta gradients over the 128 trajectory slots
#pragma unroll
    for (int i = 0; i < HODE_N_THETA; ++i) red[i * TILE + row] = gth[i];
  }
  __syncthreads();
  if (main_role && row < HODE_N_THETA) {
    float sacc = 0.f;
    for (int r = 0; r < TILE; ++r) sacc += red[row * TILE + r];
    out[A.P + row] = sacc;
  }
  tc::fence_before_sync();
  __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
  }
}

// ---- transposed weight image --------------------------------------------------------------------------
// BF16 hi / mid, K-major: element (n, k) of a block at byte (k / 8) * (N * 16) + n * 16 + (k % 8) * 2.  Bytes:
// [W_out^T: (n = in 64, k = out 16) hi 2048, mid 2048][W_l^T, l = L-1..1: hi 8192, mid 8192][W_0^T: (n = in 16, k = out 64) hi 2048, mid 2048]
int tc_bwd_image_floats(int L) { return 1024 + (L - 1) * 4096 + 1024; }

__global__ void prep_tc_bwd_image_kernel(const float* __restrict__ W, float* __restrict__ img, int L, int P,
                                         int img_floats) {
  const float* w = W + (size_t)blockIdx.x * P;
  float* out = img + (size_t)blockIdx.x * img_floats;
  for (int i = threadIdx.x; i < img_floats; i += blockDim.x) out[i] = 0.f;
  __syncthreads();
  // packed offsets: layer 0 at 0 (576 + 64), hidden l at 640 + (l-1)*4160, output at 640 + (L-1)*4160
  uint8_t* dst = reinterpret_cast<uint8_t*>(out);
  for (int l = L; l >= 0; --l) {
    const int n_out = (l == L) ? NS : H;          // = K of the transposed operand
    const int n_in = (l == 0) ? HODE_NN_IN : H;   // = N of the transposed operand
    const int Npad = (l == 0) ? 16 : H;
    const int Kpad = (l == L) ? 16 : H;
    const int part = (Kpad / 8) * Npad * 16;      // bytes of the hi (or mid) part
    const float* wl = w + (l == 0 ? 0 : 640 + (l - 1) * 4160);
    for (int i = threadIdx.x; i < n_out * n_in; i += blockDim.x) {
      const int o = i / n_in, in = i - o * n_in;   // W_l[o][in]  ->  (n = in, k = o)
      uint16_t hi, mid;
      split_bf16(wl[i], hi, mid);
      const int off = (o >> 3) * (Npad * 16) + in * 16 + (o & 7) * 2;
      *reinterpret_cast<uint16_t*>(dst + off) = hi;
      *reinterpret_cast<uint16_t*>(dst + part + off) = mid;
    }
    dst += 2 * part;
  }
}

// ---- schedule: sort by accepted-step count, tiles to CTAs by longest-processing-time-first ----------
// A tile runs for max(accepted steps) iterations over its 128 trajectories, and adaptive step counts
// differ several-fold inside a cohort (bench cohort: mean 33, mean tile maximum 55).  Trajectories
// are therefore sorted by step count (stable radix sort: the order, and with it every gradient
// sum, is a pure function of the inputs), cut into tiles, and the tiles are dealt to the CTAs of
// their parameter set greedily, longest first, each to the least-loaded CTA.
constexpr uint32_t SORT_N_BITS = 20, SORT_N_MASK = (1u << SORT_N_BITS) - 1u;

__global__ void adj_sort_keys_kernel(const int32_t* __restrict__ save_n, uint32_t* __restrict__ keys,
                                     int32_t* __restrict__ vals, long n_units, int B, int sorted) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units) return;
  int n = save_n[i];
  n = n < 0 ? 0 : (n > (int)SORT_N_MASK ? (int)SORT_N_MASK : n);
  const uint32_t s = (uint32_t)(i / B);
  keys[i] = sorted ? ((s << SORT_N_BITS) | (SORT_N_MASK - (uint32_t)n)) : (uint32_t)n;   // descending n inside a set
  vals[i] = (int32_t)i;
}

// one CTA per parameter set; warp 0 runs the greedy assignment, then thread x gathers CTA x's list
__global__ void __launch_bounds__(256) adj_schedule_kernel(const uint32_t* __restrict__ keys, int B, int n_tiles, int gx,
                                                           int sorted, int32_t* __restrict__ owner,
                                                           int32_t* __restrict__ sched_off, int32_t* __restrict__ sched_tiles) {
  constexpr int COST_CACHE = 4096;
  __shared__ int cnt[256];
  __shared__ int off[257];
  __shared__ unsigned cost_sh[COST_CACHE];   // tile costs, fetched in parallel (the greedy loop is serial)
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const uint32_t* k = keys + (size_t)s * B;
  int32_t* own = owner + (size_t)s * n_tiles;
  cnt[tid] = 0;
  if (sorted)
    for (int t = tid; t < n_tiles && t < COST_CACHE; t += blockDim.x)
      cost_sh[t] = 1u + SORT_N_MASK - (k[(size_t)t * TILE] & SORT_N_MASK);
  __syncthreads();
  if (tid < 32) {
    unsigned load[8];   // CTA x = lane + 32 * slot
#pragma unroll
    for (int j = 0; j < 8; ++j) load[j] = (lane + 32 * j) < gx ? 0u : 0xFFFFFFFFu;
    for (int t = 0; t < n_tiles; ++t) {
      // cost = iterations of the tile (its largest step count; the slots are sorted descending) + 1
      unsigned cost = 1u;
      if (sorted) {
        cost = t < COST_CACHE ? cost_sh[t] : 1u + SORT_N_MASK - (k[(size_t)t * TILE] & SORT_N_MASK);
      } else {   // unsorted fallback: scan the tile
        unsigned m = 0;
        for (int r = lane; r < TILE && (size_t)t * TILE + r < (size_t)B; r += 32) m = max(m, k[(size_t)t * TILE + r]);
#pragma unroll
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        cost += m;
      }
      unsigned long long best = ~0ull;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned long long cand = ((unsigned long long)load[j] << 32) | (unsigned)(lane + 32 * j);
        best = cand < best ? cand : best;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
      }
      const int x = (int)(best & 0xFFFFFFFFull);
      if ((x & 31) == lane) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j == (x >> 5)) load[j] += cost;
        own[t] = x;
        cnt[x] += 1;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int x = 0; x < gx; ++x) { off[x] = acc; acc += cnt[x]; }
    off[gx] = acc;
  }
  __syncthreads();
  if (tid <= gx) sched_off[(size_t)s * (gx + 1) + tid] = off[tid];
  if (tid < gx) {
    int o = off[tid];
    for (int t = 0; t < n_tiles; ++t)
      if (own[t] == tid) sched_tiles[(size_t)s * n_tiles + o++] = t;
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
bool adj_tc_supported(int H_, int L_) { return H_ == 64 && L_ >= 1 && L_ <= MAXL; }

AdjTcPlan adj_tc_plan(int B, int S, int L, int P, int T, int t_per_traj) {
  AdjTcPlan p{};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms < 1) sms = 148;
  const long blocks = ((long)B + TILE - 1) / TILE;
  long gx = sms / (S > 0 ? S : 1);
  if (gx > 256) gx = 256;   // adj_schedule_kernel keeps 8 CTA loads per lane
  if (gx > blocks) gx = blocks;
  if (gx < 1) gx = 1;
  p.grid_x = (int)gx;
  p.grid_y = S;
  p.n_tiles = (int)(blocks > 0 ? blocks : 1);
  p.fwd_floats = tc_image_floats(L);
  p.bwd_floats = tc_bwd_image_floats(L);
  const int bwd_cap = BWD_BYTES / 4;
  size_t floats = (size_t)(((p.fwd_floats > bwd_cap ? p.fwd_floats : bwd_cap) + 255) & ~255);
  floats += TILE * HODE_REC_FLOATS;
  if (!t_per_traj && T <= HODE_SIMT_MAX_SHARED_T) floats += T;
  p.smem = (floats * sizeof(float) + 1023) & ~(size_t)1023;
  p.partial_floats = (size_t)gx * S * (size_t)(P + HODE_N_THETA);
  p.stash_floats = (size_t)gx * S * (size_t)NSTAGE_MAX * L * (ST_BLK / 4);
  p.img_floats = (size_t)S * (p.fwd_floats + p.bwd_floats);
  // schedule scratch: sort keys / values (in, out), tile owners, per-CTA tile lists, cub temporaries
  const size_t units = (size_t)S * (size_t)(B > 0 ? B : 0);
  p.sched_ints = 4 * units + (size_t)S * (2 * (size_t)p.n_tiles + gx + 1);
  p.sort_bytes = 0;
  if (units > 0)
    cub::DeviceRadixSort::SortPairs(nullptr, p.sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (long long)units, 0, 32, (cudaStream_t)0);
  return p;
}

size_t adj_tc_workspace_bytes(const AdjTcPlan& p) {
  auto al = [](size_t x) { return (x + 63) & ~(size_t)63; };
  return (al(p.partial_floats) + al(p.stash_floats) + al(p.img_floats) + al(p.sched_ints)) * sizeof(float) +
         ((p.sort_bytes + 255) & ~(size_t)255);
}

cudaError_t launch_rollout_bwd_tc(const RolloutArgs& A, const float* grad_traj, float* grad_y0,
                                  float* grad_theta, float* grad_W, void* workspace, cudaStream_t stream) {
  const AdjTcPlan p = adj_tc_plan(A.B, A.S, A.L, A.P, A.T, A.t_per_traj);
  if (p.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  auto al = [](size_t x) { return (x + 63) & ~(size_t)63; };
  AdjTcArgs G{};
  G.R = A;
  G.grad_traj = grad_traj;
  G.grad_y0 = grad_y0;
  G.partials = reinterpret_cast<float*>(workspace);
  G.stash = G.partials + al(p.partial_floats);
  float* imgs = G.stash + al(p.stash_floats);
  G.img_fwd = imgs;
  G.img_bwd = imgs + (size_t)A.S * p.fwd_floats;
  G.fwd_floats = p.fwd_floats;
  G.bwd_floats = p.bwd_floats;
  // schedule scratch
  const size_t units = (size_t)A.S * (size_t)A.B;
  uint32_t* keys_in = reinterpret_cast<uint32_t*>(imgs + al(p.img_floats));
  uint32_t* keys_out = keys_in + units;
  int32_t* vals_in = reinterpret_cast<int32_t*>(keys_out + units);
  int32_t* perm = vals_in + units;
  int32_t* owner = perm + units;
  int32_t* sched_tiles = owner + (size_t)A.S * p.n_tiles;
  int32_t* sched_off = sched_tiles + (size_t)A.S * p.n_tiles;
  void* sort_tmp = reinterpret_cast<float*>(keys_in) + al(p.sched_ints);
  G.perm = perm;
  G.sched_off = sched_off;
  G.sched_tiles = sched_tiles;
  G.n_tiles = p.n_tiles;
  cudaError_t e = tc_prepare_fwd_images(A.W, imgs, A.S, A.L, A.P, stream);
  if (e != cudaSuccess) return e;
  prep_tc_bwd_image_kernel<<<A.S, 256, 0, stream>>>(A.W, imgs + (size_t)A.S * p.fwd_floats, A.L, A.P, p.bwd_floats);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  {
    // the composite key holds the parameter-set index above SORT_N_BITS bits of step count
    int s_bits = 0;
    while ((1L << s_bits) < (long)A.S) ++s_bits;
    const int sorted = (s_bits + (int)SORT_N_BITS <= 32) ? 1 : 0;
    adj_sort_keys_kernel<<<(unsigned)((units + 255) / 256), 256, 0, stream>>>(A.save_n, keys_in, vals_in, (long)units,
                                                                              A.B, sorted);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t* sched_keys = keys_in;
    if (sorted) {
      size_t tmp_bytes = p.sort_bytes;
      e = cub::DeviceRadixSort::SortPairs(sort_tmp, tmp_bytes, (const uint32_t*)keys_in, keys_out, (const int32_t*)vals_in,
                                          perm, (long long)units, 0, s_bits + (int)SORT_N_BITS, stream);
      if (e != cudaSuccess) return e;
      sched_keys = keys_out;
    } else {
      e = cudaMemcpyAsync(perm, vals_in, units * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return e;
    }
    adj_schedule_kernel<<<A.S, 256, 0, stream>>>(sched_keys, A.B, p.n_tiles, p.grid_x, sorted, owner, sched_off,
                                                 sched_tiles);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  e = cudaFuncSetAttribute(rollout_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  if (e != cudaSuccess) return e;
  rollout_bwd_tc_kernel<<<dim3(p.grid_x, p.grid_y), 3 * TILE, p.smem, stream>>>(G);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_reduce_partials(G.partials, p.grid_x, A.S, A.P, grad_W, grad_theta, stream);
}

}  // namespace hode
