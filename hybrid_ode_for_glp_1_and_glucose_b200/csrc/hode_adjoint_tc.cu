// hode_adjoint_tc.cu — discrete adjoint of the rollout with the MLP forward recomputation, the delta
// back-propagation and the weight gradients on the tcgen05 tensor cores (BASELINE.json north_star item 3).
//
// Same mathematics as hode_adjoint_simt.cu (autograd through the unrolled RK steps with the step sizes
// frozen).  Round-2 data flow — no activation ever travels through HBM:
//   * the tensor-core rollout records, with every accepted step, its stage derivatives k1..k6 (192-byte
//     record, hode_common.cuh).  Every stage INPUT of a step is then a linear combination of recorded values,
//     so the stages are independent of one another: the adjoint treats every (step, stage) pair as one ITEM,
//     walks the items in reverse order and, per item, recomputes the hidden activations of that ONE stage
//     (the output layer is not needed) and pulls the stage's cotangent back through them.  The activations of
//     one stage (4 x 32 KB as BF16 operand images) live in a per-CTA scratch that is written and read back
//     within microseconds: 38 MB for the whole grid, L2-resident (round 1 stashed all 6 stages of a step and
//     streamed 78 GB through HBM per launch);
//   * the recomputation F of item e+1 runs CONCURRENTLY with the pull-back B of item e on the same 128
//     trajectories: the two are independent (F needs no cotangent).  Each chain has its own epilogue warps and its
//     own MMA-issuer warp, so neither waits for the other; the tensor pipe interleaves their MMAs;
//   * roles (warp-specialised, setmaxnreg), 768 threads: 4 main warps (one trajectory per thread: records, stage
//     inputs, cotangent of the network outputs, mechanistic VJP, RK recurrences), 8 recomputation-epilogue warps
//     (ReLU / operand split / activation scratch, 32 columns each), 8 pull-back-epilogue warps (ReLU mask / delta
//     image, 32 columns each), and one warp each for: F issue + forward-weight loads, B issue, W^T loads (TMA),
//     activation loads (TMA); loaders run ahead as far as the buffers allow (full / free mbarrier pairs,
//     free = tcgen05.commit);
//   * schedule (device side, deterministic): trajectories are radix-sorted by accepted-step count, cut into
//     128-trajectory tiles, and the tiles are dealt to the CTAs longest-first;
//   * delta is written ONCE per phase, by its owner threads, as a two-term BF16 image in shared memory
//     (16-byte vectors) that serves both products of the phase (csrc/probe/bf16_probe.cu), 3 passes each:
//       u_{l-1} = delta_l W_l            reads it as a K-major A operand against the BF16 image of W_l^T,
//       dW_l += delta_l^T [a_{l-1} | 1]  reads it MN-major (contraction over the tile's 128 trajectories)
//     against the stashed activations; the weight-gradient accumulators stay in TMEM for the whole kernel, so
//     every gradient element is summed in one fixed order (no atomics, bit-reproducible).  The recomputation
//     uses the rollout's own arithmetic on the rollout's own stage inputs: its activations (and ReLU masks)
//     are the forward pass's, bit for bit;
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

constexpr int MAXL = 4;       // hidden layers supported by the TMEM accumulator map
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

// ---- TMEM column map (all 512 columns; csrc/probe/mix_probe.cu test 3: accumulators at any multiple of 8) --------
//   [0,200)    the recomputation's tile (hode_tc_mlp.cuh: D, A_hi, A_lo / BF16 operands, constant block)
//   [200,264)  D_B: accumulator of the delta chain u_{q-1} = delta_q W_q
//   weight-gradient accumulators, resident for the whole kernel (fp32):
//   hidden layer l = 1..3 : columns DW_H0 + 72 (l-1) .. +72   D[j][k] = dW_l[j][k], column 64 = db_l[j]      (M = 64)
//   layer 0               : columns DW_0  .. +16               D[j][k] = dW_0[j][k] (k < 9), column 15 = db_0[j] (M = 64)
//   output layer (transp.): columns DW_O  .. +16               D[k][n] = dW_out[n][k] (n < 6), row 64 = db_out[n] (M = 128)
constexpr uint32_t TM_DB = 200, DW_H0 = 264, DW_HS = 72, DW_0 = 480, DW_O = 496;

// ---- shared memory (bytes) ------------------------------------------------------------------------------------
//   AB  2 x [hi: 8 feature groups + the constant-1 group][mid: likewise]: input operand a_{q-1} of the weight gradients
//   DB  2 x [hi][mid]: the delta image of a phase (buffered by phase parity)
//   XB  2 x [hi: 2 groups][mid: 2 groups]: the stage's input features x, input operand of dW_0 (buffered by item parity)
//   WT  2 x one layer's W^T (BF16 hi, mid)         WF  one layer's forward image + its bias block
//   MK  2 x [L][2 halves][128] ReLU masks of the item being recomputed / pulled back (word: bit j <=> a[32 half + j] > 0)
constexpr int AB_PART = 9 * ST_GRP, AB_BYTES = 2 * AB_PART;
constexpr int DB_BYTES = 2 * ST_PART;
constexpr int XB_PART = 2 * ST_GRP, XB_BYTES = 2 * XB_PART;
constexpr int WT_BYTES = 2 * 64 * 64 * 2;
constexpr int WF_FLOATS = 8192 + 512, WF_BYTES = WF_FLOATS * 4;
constexpr int MK_BYTES = MAXL * 2 * TILE * 4;
constexpr int OFF_AB = 0, OFF_DB = OFF_AB + 2 * AB_BYTES, OFF_XB = OFF_DB + 2 * DB_BYTES, OFF_WT = OFF_XB + 2 * XB_BYTES,
              OFF_WF = OFF_WT + 2 * WT_BYTES, OFF_MK = OFF_WF + WF_BYTES, OFF_T = OFF_MK + 2 * MK_BYTES;
constexpr int ADJ_SMEM_BASE = OFF_T;            // + 4 T bytes when the shared time grid fits
constexpr int ADJ_SMEM_MAX = 227 * 1024 - 256;  // (static shared memory: the mbarriers, padded to the dynamic part's alignment)
constexpr int ADJ_THREADS = 6 * TILE;

struct Bars {
  uint64_t f_bar;        // recomputation: a layer's MMAs complete (observed by the F-epilogue warps)
  uint64_t fl_bar;       // recomputation: the LAST layer's MMAs of an item complete (observed by the main warps)
  uint64_t b_bar;        // pull-back: a phase's delta chain complete (observed by the B-epilogue warps)
  uint64_t gx_bar;       // pull-back: the final phase (g_x) of an item complete (observed by the main warps)
  uint64_t wf_full, wf_free;
  uint64_t wt_full[2], wt_free[2];
  uint64_t ab_full[2], ab_free[2];
  // 256 arrivals: a_l of the item being recomputed is in scratch set (item & 1).  One barrier per SET and layer: the
  // recomputation runs up to one item ahead of the pull-back, so a per-layer barrier could complete twice before
  // its waiter looks (parity aliasing); a per-set one completes once per two items
  uint64_t st_done[2][MAXL];
  uint64_t done;
};

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

// ---- weight gradients on the tensor cores ---------------------------------------------------------
// dW_l += delta_l^T [a_{l-1} | 1] contracts over the 128 trajectories of the tile: an SS-form MMA
// with K = trajectory.  kind::f16 takes MN-major operands, so "thread t owns trajectory t" writes 8
// consecutive features as ONE 16-byte vector (csrc/probe/bf16_probe.cu):
//   element (trajectory t, feature f) at byte (f / 8) * ST_GRP + t * 16 + (f % 8) * 2
//   descriptor: LBO = 128 B (between 8-trajectory groups), SBO = ST_GRP (between 8-feature groups).
// Operands are split in two BF16 terms, x ~= hi + mid (2^-17), and the product takes 3 passes
// mid*hi + hi*mid + hi*hi at the BF16 rate: measured max error 2.7e-6 of sum|ab| per 128-term product.
// D[tmem] (+)= R^T C: r_*: the operand whose features become the ROWS of D, c_*: the COLUMNS.
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
  constexpr uint32_t sa = 2u * ST_GRP / 16u, sb = 2u * (uint32_t)N;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, am + sa * ks, bh + sb * ks, idesc, ks == 0 ? 0u : 1u);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, ah + sa * ks, bm + sb * ks, idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) mma_bf16_ss(d, ah + sa * ks, bh + sb * ks, idesc, 1u);
}

// ---- named barriers: producers ARRIVE, the issuer warp of the chain waits (bar.sync) ---------------------------------
//   BAR_X  main -> F issuer: the item's input operand is in TMEM          BAR_F  F-epilogue warps -> F issuer
//   BAR_BX main -> B issuer: delta_L and the dW_0 input operand are written BAR_B  B-epilogue warps -> B issuer
// Race-free: the MMA chain whose completion lets a thread move on to its next arrival on a barrier is only issued
// after the barrier's previous generation has completed.
constexpr int BAR_X = 1, BAR_F = 2, BAR_BX = 3, BAR_B = 4, BAR_MAIN = 5;
template <int ID, int THREADS>
__device__ __forceinline__ void bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(THREADS) : "memory"); }
template <int ID, int THREADS>
__device__ __forceinline__ void bar_wait() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(THREADS) : "memory"); }
constexpr int N_MAIN = TILE + 32, N_EPI = 2 * TILE + 32;

// closed-form VJP of f_physio — same formulas as hode_adjoint_simt.cu::rhs_mech_vjp
__device__ __forceinline__ void mech_vjp(const Theta& p, const float* y, float GD, bool gd_present,
                                         const float* c, float* gy, float* gth) {
  const float G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const float Pi = 1.0f + p.rho * GLP1;
  const float dG = G - p.G_b, dI = I - p.I_b, dGlu = Glu - p.Glu_b;
  const float inv_e = __frcp_rn(p.EC_50 + GLP1);
  const float frac_e = GLP1 * inv_e;
  const float ge = p.E_max * frac_e;
  const float inv_m = __frcp_rn(p.K_m + G);
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

// ---- recomputation epilogue (F-epilogue warps; half 0 = columns [0,32), half 1 = [32,64)) ---------------------------
// hidden layer l of an item: a_l = relu(z_l) -> next layer's A operand (TMEM), -> the scratch as the weight gradients'
// BF16 operand image (global, L2-resident) and its ReLU mask -> shared memory.  The scratch is read back by bulk copies
// (async proxy), so its generic-proxy stores need a cross-proxy fence before the "stashed" signal.  That fence costs
// ~1 k cycles whenever it is issued (measured: profiles/r02_timeline_*), and the pull-back needs the item's operands
// only one slot later: ONE fence per item, after the last layer, then all L signals.
template <int MODE>
__device__ __forceinline__ void fwd_epilogue(Bars* bars, uint8_t* smem, uint32_t t_lane, int half, int row, uint8_t* stash_set,
                                             int set, int l, int L) {
  const uint32_t col0 = (uint32_t)half * 32u;
  const bool last = (l + 1 == L);
  uint32_t v[32], lo[32];
  HODE_TL(300);
  epilogue32_to_tmem<MODE>(t_lane, col0, v, lo, !last);
  bar_arrive<BAR_F, N_EPI>();
  HODE_TL(301);
  const uint32_t mask = stash_store32(stash_set + (size_t)l * ST_BLK, row, half, v, lo);
  reinterpret_cast<uint32_t*>(smem + OFF_MK + set * MK_BYTES)[(l * 2 + half) * TILE + row] = mask;
  HODE_TL(302);
  if (last) {
    tc::fence_proxy_async_all();
    for (int ll = 0; ll < L; ++ll) tc::mbar_arrive(&bars->st_done[set][ll]);
  }
  HODE_TL(303);
}

// items of one tile: iteration it = 0..n_iter-1 (one accepted step each, last step first), stages i = N-1 .. i_lo(it).
// DP5(4) is first-same-as-last: stage 1 of step n+1 IS stage 7 of step n, pulled back once, as stage 7 of the
// earlier step, with both cotangents added; stage 1 of the very first step is pulled back by one extra iteration (a
// zero-length step at (t0, y0)) of which only stage 7 carries a cotangent.
__device__ __forceinline__ int item_i_lo(int it, int n_iter, int N, int i0, bool fsal) {
  return (fsal && it == n_iter - 1) ? N - 1 : i0;
}

}  // namespace

struct AdjTcArgs {
  RolloutArgs R;
  const float* grad_traj;  // [S,B,T,6]
  float* grad_y0;          // [S,B,6] or nullptr
  float* partials;         // [gridDim.y * gridDim.x][P + 17]
  uint8_t* stash;          // [ctas][2][L][ST_BLK] activation scratch (L2-resident)
  const float* img_fwd;    // [S][fwd_floats] (hode_rollout_tc.cu prep_tc_image_kernel)
  const float* img_bwd;    // [S][bwd_floats]
  int fwd_floats, bwd_floats;
  int t_in_smem;           // the shared time grid is staged in shared memory
  // schedule (adj_schedule_kernel): trajectories sorted by accepted-step count, tiles handed to CTAs
  const int32_t* perm;        // [S*B] unit index of sorted slot q (within its parameter set)
  const int32_t* sched_off;   // [S][grid_x + 1] range of sched_tiles owned by CTA (s, x)
  const int32_t* sched_tiles; // [S][n_tiles] tile indices grouped by owner
  int n_tiles;
};

// ---------------------------------------------------------------------------------------------------
// grid = (ctas per parameter set, S), block = 768 = 6 warpgroups, 1 CTA / SM:
//   WG0 main | WG1, WG2 recomputation epilogues (columns [0,32) / [32,64)) | WG3, WG4 pull-back epilogues |
//   WG5: warp 20 F issuer + forward-weight loads, 21 B issuer, 22 W^T loader, 23 activation loader
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(ADJ_THREADS, 1) rollout_bwd_tc_kernel(const AdjTcArgs G) {
  extern __shared__ __align__(128) uint8_t smem_raw[];   // (operand descriptors without swizzle need 16-byte alignment)
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_nmax;

  const RolloutArgs& A = G.R;
  const int tid = threadIdx.x, lane_id = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int wg = warp >> 2;
  const int wq = warp & 3, row = tid & 127;
  const int s = blockIdx.y, T = A.T, L = A.L;
  const int solver = A.solver == HODE_SOLVER_RK4 ? 0 : 1;
  const int N = solver == 0 ? 4 : 7;
  const bool fsal = solver == 1;   // (the records of this path always carry the stage derivatives)
  const int i0 = fsal ? 1 : 0;

  float* t_sh_buf = reinterpret_cast<float*>(smem_raw + OFF_T);
  if (tid == 0) {
    tc::mbar_init(&bars.f_bar, 1);
    tc::mbar_init(&bars.fl_bar, 1);
    tc::mbar_init(&bars.b_bar, 1);
    tc::mbar_init(&bars.gx_bar, 1);
    tc::mbar_init(&bars.wf_full, 1);
    tc::mbar_init(&bars.wf_free, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars.wt_full[i], 1);
      tc::mbar_init(&bars.wt_free[i], 1);
      tc::mbar_init(&bars.ab_full[i], 1);
      tc::mbar_init(&bars.ab_free[i], 1);
    }
    for (int i = 0; i < MAXL; ++i) { tc::mbar_init(&bars.st_done[0][i], 2 * TILE); tc::mbar_init(&bars.st_done[1][i], 2 * TILE); }
    tc::mbar_init(&bars.done, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  const float* t_shared = nullptr;
  if (G.t_in_smem) {
    for (int i = tid; i < T; i += blockDim.x) t_sh_buf[i] = A.t_obs[i];
    t_shared = t_sh_buf;
  }
  if (wg == 0) {
    // the constant-1 input feature of the weight gradients (its accumulator column is the bias gradient):
    // group 8 of both input-operand buffers = [1, 0 x 7] (hi) / 0 (mid); the bulk copies only touch groups 0..7
#pragma unroll
    for (int bf = 0; bf < 2; ++bf) {
      uint8_t* g8 = smem_raw + OFF_AB + bf * AB_BYTES + 8 * ST_GRP + row * 16;
      *reinterpret_cast<uint4*>(g8) = make_uint4(0x00003F80u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(g8 + AB_PART) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
  if (wg == 0) {
    uint32_t ones[8] = {0x3F800000u, 0x3F800000u, 0u, 0u, 0u, 0u, 0u, 0u};
    HODE_TMEM_ST_X8(tmem + lane_base + TM_ONES, ones);
    tc::wait_st();
  }
  uint8_t* stash0 = G.stash + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(2 * L) * ST_BLK;
  const float* fwd_src = G.img_fwd + (size_t)s * G.fwd_floats;
  const float* bwd_src = G.img_bwd + (size_t)s * G.bwd_floats;

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
  auto tile_items = [&](int n_iter) -> int {
    int items = 0;
    for (int it = 0; it < n_iter; ++it) items += N - item_i_lo(it, n_iter, N, i0, fsal);
    return items;
  };

  // Roles: each entirely inside its own branch so that ptxas allocates registers against the role's budget.
  if (wg == 1 || wg == 2) {
    // ================================ recomputation epilogues ==================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    const int half = wg - 1;
    uint32_t f_cnt = 0u, m_f = 0u;
    for (int tk = tile_beg; tk < tile_end; ++tk) {
      const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
      const int items = tile_items(n_iter);
#pragma unroll 1
      for (int m = 0; m < items; ++m) {
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          tc::mbar_wait(&bars.f_bar, f_cnt & 1u);
          f_cnt += 1u;
          tc::fence_after_sync();
          fwd_epilogue<MODE>(&bars, smem_raw, tmem + lane_base, half, row, stash0 + (size_t)(m_f & 1u) * L * ST_BLK, (int)(m_f & 1u), l, L);
        }
        m_f += 1u;
      }
    }
  } else if (wg == 3 || wg == 4) {
    // ================================ pull-back epilogues ======================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    const int half = wg - 3;
    uint32_t b_cnt = 0u, ph = 0u, m_b = 0u;
    for (int tk = tile_beg; tk < tile_end; ++tk) {
      const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
      const int items = tile_items(n_iter);
#pragma unroll 1
      for (int m = 0; m < items; ++m, ++m_b) {
        // phase p = L..1 has delivered u_{p-1} in D_B: delta_{p-1} = u_{p-1} * relu'(a_{p-1}) -> delta image of phase p-1
#pragma unroll 1
        for (int p = L; p >= 1; --p) {
          tc::mbar_wait(&bars.b_bar, b_cnt & 1u);
          b_cnt += 1u;
          tc::fence_after_sync();
          HODE_TL(320);
          // ReLU mask of a_{p-1}: written to shared memory by the F-epilogue warps when the item was recomputed
          tc::mbar_wait(&bars.st_done[m_b & 1u][p - 1], (m_b >> 1) & 1u);
          HODE_TL(321);
          const uint32_t mask = reinterpret_cast<const uint32_t*>(smem_raw + OFF_MK + (m_b & 1u) * MK_BYTES)[((p - 1) * 2 + half) * TILE + row];
          uint8_t* db = smem_raw + OFF_DB + ((ph + 1u) & 1u) * DB_BYTES + (half * 4) * ST_GRP + row * 16;
#pragma unroll
          for (int c16 = 0; c16 < 2; ++c16) {
            uint32_t u[16];
            HODE_TMEM_LD_X16(tmem + lane_base + TM_DB + (uint32_t)(half * 32 + c16 * 16), u);
            tc::wait_ld();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float d[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] = ((mask >> (c16 * 16 + g * 8 + j)) & 1u) ? __uint_as_float(u[g * 8 + j]) : 0.f;
              uint4 vh, vm;
              bf16_split8(d, vh, vm);
              *reinterpret_cast<uint4*>(db + (c16 * 2 + g) * ST_GRP) = vh;
              *reinterpret_cast<uint4*>(db + (c16 * 2 + g) * ST_GRP + ST_PART) = vm;
            }
          }
          HODE_TL(322);
          tc::fence_proxy_async();
          tc::fence_before_sync();
          bar_arrive<BAR_B, N_EPI>();
          HODE_TL(323);
          ph += 1u;
        }
        tc::mbar_wait(&bars.b_bar, b_cnt & 1u);   // g_x (read by the main warps): the phase must be observed
        HODE_TL(324);
        b_cnt += 1u;
        ph += 1u;
      }
    }
  } else if (wg == 5) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
    const uint32_t base = tc::smem_u32(smem_raw);
    if (warp == 20) {
      // ================================ F issuer (+ forward weights, one layer at a time) ========================
      const uint32_t wf_s = base + OFF_WF;
      const int bias_off = (int)(IMG_L0 + (uint32_t)(L - 1) * IMG_HID + IMG_OUT);
      uint32_t n_f = 0u;
      auto load_layer = [&](int l) {   // the slot is free: its previous user's MMAs have completed
        if (lane_id == 0) {
          uint8_t* dst = smem_raw + OFF_WF;
          if (l == 0) {
            tc::mbar_expect_tx(&bars.wf_full, IMG_L0 * 4u);
            tc::bulk_g2s(dst, fwd_src, IMG_L0 * 4u, &bars.wf_full);
          } else {
            tc::mbar_expect_tx(&bars.wf_full, (IMG_HID + 512u) * 4u);
            tc::bulk_g2s(dst, fwd_src + IMG_L0 + (size_t)(l - 1) * IMG_HID, IMG_HID * 4u, &bars.wf_full);
            tc::bulk_g2s(dst + IMG_HID * 4u, fwd_src + bias_off + l * 512, 512u * 4u, &bars.wf_full);
          }
        }
        __syncwarp();
      };
      bool any = false;
      for (int tk = tile_beg; tk < tile_end; ++tk) {
        const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
        const int items = tile_items(n_iter);
#pragma unroll 1
        for (int m = 0; m < items; ++m) {
#pragma unroll 1
          for (int l = 0; l < L; ++l) {
            if (n_f == 0u) load_layer(0);   // (every later load is started right after its predecessor's MMAs, below)
            if (l == 0) {
              bar_wait<BAR_X, N_MAIN>();               // the item's input operand (main warps)
              if (any) bar_wait<BAR_F, N_EPI>();       // the previous item's last accumulator has been read
            } else {
              bar_wait<BAR_F, N_EPI>();
            }
            tc::mbar_wait(&bars.wf_full, n_f & 1u);
            if (tc::elect_one()) {
              tc::fence_after_sync();
              if (l == 0) issue_layer<MODE, H, 16, false>(tmem, wf_s, wf_s + 1024u * 4u, 0u, tmem + TM_ONES);
              else issue_layer<MODE, H, 64, true>(tmem, wf_s, wf_s + 4096u * 4u, wf_s + 8192u * 4u, tmem + TM_ONES);
              tc::mma_commit(&bars.f_bar);
              tc::mma_commit(&bars.wf_free);
              if (l + 1 == L) tc::mma_commit(&bars.fl_bar);
            }
            __syncwarp();
            any = true;
            // next layer's weights as soon as this layer's MMAs have read theirs (the epilogue runs meanwhile)
            const bool more = !(tk == tile_end - 1 && m == items - 1 && l == L - 1);
            tc::mbar_wait(&bars.wf_free, n_f & 1u);
            if (more) load_layer(l + 1 == L ? 0 : l + 1);
            n_f += 1u;
          }
        }
      }
      if (any) bar_wait<BAR_F, N_EPI>();   // the last item's last arrival
    } else if (warp == 21) {
      // ================================ B issuer ===================================================================
      // phase q:  u_{q-1} = delta_q W_q (delta image x W_q^T slot)  and  dW_q += delta_q^T [a_{q-1} | 1]
      // (a_{-1} = the stage's input features x, u_{-1} = the cotangent of x)
      uint32_t ph = 0u, first = 1u, m_b = 0u;
      for (int tk = tile_beg; tk < tile_end; ++tk) {
        const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
        const int items = tile_items(n_iter);
#pragma unroll 1
        for (int m = 0; m < items; ++m) {
#pragma unroll 1
          for (int q = L; q >= 0; --q) {
            const uint32_t buf = ph & 1u, par = (ph >> 1) & 1u;
            const uint32_t ws = base + OFF_WT + buf * WT_BYTES;
            const uint32_t a_hi = base + OFF_AB + buf * AB_BYTES, a_mid = a_hi + AB_PART;
            const uint32_t x_hi = base + OFF_XB + (m_b & 1u) * XB_BYTES, x_mid = x_hi + XB_PART;
            const uint32_t d_hi = base + OFF_DB + buf * DB_BYTES, d_mid = d_hi + ST_PART;
            if (q == L) bar_wait<BAR_BX, N_MAIN>();
            else bar_wait<BAR_B, N_EPI>();
            tc::mbar_wait(&bars.wt_full[buf], par);
            if (tc::elect_one()) {
              tc::fence_after_sync();
              if (q == L) issue_u<H, 1>(tmem + TM_DB, d_hi, d_mid, ws, ws + 2048u);          // u_{L-1} = delta_L W_out (K = 16)
              else if (q >= 1) issue_u<H, 4>(tmem + TM_DB, d_hi, d_mid, ws, ws + 8192u);     // u_{q-1} = delta_q W_q
              else issue_u<16, 4>(tmem + TM_DB, d_hi, d_mid, ws, ws + 2048u);                // g_x = delta_0 W_0
              tc::mma_commit(&bars.b_bar);
              tc::mma_commit(&bars.wt_free[buf]);
              if (q == 0) tc::mma_commit(&bars.gx_bar);
            }
            __syncwarp();
            tc::mbar_wait(&bars.ab_full[buf], par);
            if (tc::elect_one()) {
              tc::fence_after_sync();
              // dW_out^T [in k][out n] = [a_{L-1} | 1]^T delta_L;  dW_q += delta_q^T [a_{q-1} | 1];  dW_0 += delta_0^T [x | 1]
              if (q == L) issue_dw<128, 16>(tmem + DW_O, a_hi, a_mid, d_hi, d_mid, first);
              else if (q >= 1) issue_dw<64, 72>(tmem + DW_H0 + DW_HS * (uint32_t)(q - 1), d_hi, d_mid, a_hi, a_mid, first);
              else issue_dw<64, 16>(tmem + DW_0, d_hi, d_mid, x_hi, x_mid, first);
              tc::mma_commit(&bars.ab_free[buf]);
            }
            __syncwarp();
            ph += 1u;
          }
          first = 0u;
          m_b += 1u;
        }
      }
      if (tc::elect_one()) tc::mma_commit(&bars.done);   // every pull-back MMA of this CTA has completed when this arrives
      __syncwarp();
    } else if (warp == 22) {
      // ================================ loader: W_q^T of every pull-back phase ====================================
      uint32_t ph = 0u;
      for (int tk = tile_beg; tk < tile_end; ++tk) {
        const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
        const int items = tile_items(n_iter);
#pragma unroll 1
        for (int m = 0; m < items; ++m) {
#pragma unroll 1
          for (int q = L; q >= 0; --q) {
            const uint32_t buf = ph & 1u, k = ph >> 1;
            if (k >= 1u) tc::mbar_wait(&bars.wt_free[buf], (k - 1u) & 1u);
            if (lane_id == 0) {
              int off, floats;
              if (q == L) { off = 0; floats = 1024; }
              else if (q >= 1) { off = 1024 + (L - 1 - q) * 4096; floats = 4096; }
              else { off = 1024 + (L - 1) * 4096; floats = 1024; }
              tc::mbar_expect_tx(&bars.wt_full[buf], (uint32_t)floats * 4u);
              tc::bulk_g2s(smem_raw + OFF_WT + buf * WT_BYTES, bwd_src + off, (uint32_t)floats * 4u, &bars.wt_full[buf]);
            }
            __syncwarp();
            ph += 1u;
          }
        }
      }
    } else {
      // ================================ loader: stashed activations a_{q-1} of every pull-back phase ==================
      uint32_t ph = 0u, m_all = 0u;
      for (int tk = tile_beg; tk < tile_end; ++tk) {
        const int n_iter = tile_nmax(0) + (fsal ? 1 : 0);
        const int items = tile_items(n_iter);
#pragma unroll 1
        for (int m = 0; m < items; ++m) {
#pragma unroll 1
          for (int q = L; q >= 0; --q) {
            const uint32_t buf = ph & 1u, k = ph >> 1;
            // written by the F-epilogue warps (generic proxy + fence); item m_all is the (m_all >> 1)-th user of its set
            if (q >= 1) tc::mbar_wait(&bars.st_done[m_all & 1u][q - 1], (m_all >> 1) & 1u);
            if (k >= 1u) tc::mbar_wait(&bars.ab_free[buf], (k - 1u) & 1u);   // the buffer's previous weight-gradient MMAs are done
            if (lane_id == 0) {
              if (q >= 1) {
                const uint8_t* src = stash0 + ((size_t)(m_all & 1u) * L + (q - 1)) * ST_BLK;
                uint8_t* dst = smem_raw + OFF_AB + buf * AB_BYTES;
                tc::mbar_expect_tx(&bars.ab_full[buf], 2u * ST_PART);
                tc::bulk_g2s(dst, src, ST_PART, &bars.ab_full[buf]);
                tc::bulk_g2s(dst + AB_PART, src + ST_PART, ST_PART, &bars.ab_full[buf]);
              } else {
                tc::mbar_arrive(&bars.ab_full[buf]);   // phase 0: its input operand is x (XB, written by the main warps)
              }
            }
            __syncwarp();
            ph += 1u;
          }
          m_all += 1u;
        }
      }
    }
  } else {
    // ================================ main warps: one trajectory per thread ====================================
    // register pool of the CTA = 768 threads x 80 registers at launch; after the role split it must not grow:
    // 128 x 168 (main) + 256 x 88 (F epilogues) + 256 x 56 (B epilogues) + 128 x 24 (issuers, loaders) = 61 440
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;" ::: "memory");
    uint32_t m_all = 0u;   // items handed to the chains so far (recomputation runs one item ahead of the pull-back)
    const Theta th = load_theta(A.theta + (size_t)s * HODE_N_THETA);
    const bool gd_present = A.in_mode[HODE_CH_GD] != HODE_IN_ABSENT;
    const int nsub = A.n_substeps > 0 ? A.n_substeps : 1;
    bool have = false;
    float gth[HODE_N_THETA];
#pragma unroll
    for (int i = 0; i < HODE_N_THETA; ++i) gth[i] = 0.f;

    for (int tk = tile_beg; tk < tile_end; ++tk) {
      const long q = (long)my_tiles[tk] * TILE + row;   // slot in the step-count-sorted order
      const bool valid = q < A.B;
      const long unit = valid ? (long)G.perm[(size_t)s * A.B + q] : (long)s * A.B;
      const long bs = unit - (long)s * A.B;
      int n = valid ? A.save_n[unit] : 0;
      const bool ok = valid && n >= 0;
      if (n < 0) n = 0;
      const int nmax = tile_nmax(n);
      const int n_iter = nmax + (fsal ? 1 : 0);
      if (n_iter == 0) {
        if (G.grad_y0 && valid) {   // no steps at all (T == 1): the gradient of y0 is the cotangent of the first row
          float* o = G.grad_y0 + (size_t)unit * NS;
          const float* g = G.grad_traj + (size_t)unit * T * NS;
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) o[cc] = ok ? g[cc] : 0.f;
        }
        continue;
      }
      have = true;

      TrajInputs in;
      in.T = T; in.cur = 0;
      in.t_obs = A.t_per_traj ? A.t_obs + bs * T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        in.mode[ch] = A.in_mode[ch];
        in.u[ch] = in.mode[ch] == HODE_IN_SERIES ? A.u[ch] + bs * T
                 : in.mode[ch] == HODE_IN_CONST ? A.u[ch] + bs : nullptr;
      }
      // Inputs of a step as one linear piece per channel (value = v1 + alpha * dv, the rollout's own formula,
      // hode_rollout_tc.cu::lane_eval): valid when a step cannot cross an input kink (RK4, kink clipping, or no
      // series input).  Otherwise every stage looks its interval up.  Channels: [0] = tVNS, [1] = GD (the meal
      // input enters f additively and has no parameter: the adjoint does not need it).
      const bool cached = solver == 0 || A.kink_mode == HODE_KINK_CLIP || !any_series(in);
      const float* gtraj = G.grad_traj + (size_t)unit * T * NS;
      const double t_bound = (double)in.t_obs[T - 1];
      const double t_first = (double)in.t_obs[0];
      float lam[NS], carry[NS], y_later[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) { lam[i] = 0.f; carry[i] = 0.f; y_later[i] = 0.f; }
      double t_later = 0.0;
      bool later_valid = false;   // y_later / t_later hold the start of the step pulled back one iteration ago
      int ei = T - 1;

      // ---- the step being pulled back ("context") and the input piece of it / of the next (earlier) step --------
      double t = t_first, t_new = t_first, h = 0.0;
      float hf = 0.f, y[NS], k[6][NS];
      bool real = false, act = false;
      float pc_t1 = 0.f, pc_inv = 1.f, pc_v1[2], pc_dv[2];   // current piece
      float pn_t1 = 0.f, pn_inv = 1.f, pn_v1[2], pn_dv[2];   // piece of the step of the next iteration
      auto piece_for = [&](double tt, float& t1, float& inv, float* v1, float* dv) {
        t1 = 0.f; inv = 1.f;
        v1[0] = in.mode[HODE_CH_TVNS] == HODE_IN_CONST ? in.u[HODE_CH_TVNS][0] : 0.f;
        v1[1] = in.mode[HODE_CH_GD] == HODE_IN_CONST ? in.u[HODE_CH_GD][0] : 0.f;
        dv[0] = 0.f; dv[1] = 0.f;
        if (cached && any_series(in) && T >= 2) {
          int lo = 0, hi = T;
          const float t32 = (float)tt;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (in.t_obs[mid] < t32) lo = mid + 1; else hi = mid; }
          int j0 = (lo < T && in.t_obs[lo] == t32) ? lo : lo - 1;
          j0 = j0 < 0 ? 0 : (j0 > T - 2 ? T - 2 : j0);
          t1 = in.t_obs[j0];
          inv = 1.0f / (in.t_obs[j0 + 1] - t1);
          if (in.mode[HODE_CH_TVNS] == HODE_IN_SERIES) {
            const float a_ = in.u[HODE_CH_TVNS][j0], b_ = in.u[HODE_CH_TVNS][j0 + 1];
            v1[0] = a_; dv[0] = b_ - a_;
          }
          if (in.mode[HODE_CH_GD] == HODE_IN_SERIES) {
            const float a_ = in.u[HODE_CH_GD][j0], b_ = in.u[HODE_CH_GD][j0 + 1];
            v1[1] = a_; dv[1] = b_ - a_;
          }
        }
      };
      // tVNS and GD at time t32 (piece = the step's cached piece)
      auto inputs_at = [&](float t32, float t1, float inv, const float* v1, const float* dv, float& tvns, float& gd) {
        if (cached) {
          const float alpha = (t32 - t1) * inv;
          tvns = fmaf(alpha, dv[0], v1[0]);
          gd = fmaf(alpha, dv[1], v1[1]);
        } else {
          int lo = 0, hi = T;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (in.t_obs[mid] < t32) lo = mid + 1; else hi = mid; }
          tvns = input_channel(in, HODE_CH_TVNS, t32, lo);
          gd = input_channel(in, HODE_CH_GD, t32, lo);
        }
      };
      // record of step sidx -> context (the records come from L2: the next one is prefetched a step ahead)
      auto set_ctx = [&](int it) {
        const int sidx = n - 1 - it;
        real = ok && sidx >= 0;
        act = real || (fsal && ok && sidx == -1);   // sidx == -1: the zero-length step at (t0, y0)
        t = t_first; t_new = t_first; h = 0.0;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          y[cc] = (act && !real) ? y_later[cc] : 0.f;   // the zero-length step starts from the state the first record holds
#pragma unroll
          for (int j = 0; j < 6; ++j) k[j][cc] = 0.f;
        }
        if (real) {
          const float4* r4 = reinterpret_cast<const float4*>(step_rec(A, unit, sidx));
          float r[HODE_REC_FLOATS_K];
#pragma unroll
          for (int c4 = 0; c4 < HODE_REC_FLOATS_K / 4; ++c4) {
            const float4 v = r4[c4];
            r[4 * c4] = v.x; r[4 * c4 + 1] = v.y; r[4 * c4 + 2] = v.z; r[4 * c4 + 3] = v.w;
          }
          t = __longlong_as_double(((long long)__float_as_int(r[1]) << 32) | (long long)(unsigned)__float_as_int(r[0]));
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) {
            y[cc] = r[4 + cc];
#pragma unroll
            for (int j = 0; j < 6; ++j) k[j][cc] = r[10 + 6 * j + cc];
          }
          if (solver == 0) {
            h = (double)r[2];
            t_new = t + h;
          } else {
            t_new = (sidx + 1 < n) ? t_later : t_bound;   // the later record's start time, seen one iteration ago
            h = t_new - t;
          }
        }
        hf = (float)h;
        if (ok && sidx >= 1) {   // the next iteration's record: pull it into L2 now (192 bytes: two or three 64-byte halves of lines)
          const char* nxt = reinterpret_cast<const char*>(step_rec(A, unit, sidx - 1));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + 128));
        }
      };
      // input features of stage i of the current step: x = [t_i, ys_i, ys_i[3], tVNS(t_i)]
      auto stage_x = [&](int i, float* ys, float* x, float& gd) {
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          float a_ = 0.f;
#pragma unroll
          for (int j = 0; j < 6; ++j) a_ = fmaf(kA[solver][i][j], k[j][cc], a_);   // zero for j >= i
          ys[cc] = fmaf(hf, a_, y[cc]);
        }
        if (fsal && i == N - 1 && real && later_valid) {
          // stage 7 is evaluated at the step's result = the state the later step started from (its record), bit for bit
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) ys[cc] = y_later[cc];
        }
        const float ci = kC[solver][i];
        const double te = (i == 0) ? t : (ci == 1.0f ? t_new : t + (double)ci * h);
        const float t32 = (float)te;
        float tvns;
        inputs_at(t32, pc_t1, pc_inv, pc_v1, pc_dv, tvns, gd);
        if (!act) { tvns = 0.f; gd = 0.f; }
        x[0] = t32;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) x[1 + cc] = ys[cc];
        x[7] = ys[3];
        x[8] = tvns;
      };
      // hand the recomputation of an item to the F chain: its input operand -> TMEM (the A operand region is free
      // once the previous item's last layer has completed)
      auto start_F = [&](const float* xf) {
        if (m_all >= 1u) tc::mbar_wait(&bars.fl_bar, (m_all - 1u) & 1u);
        tc::fence_after_sync();
        __syncwarp();
        store_input_operand<MODE>(tmem + lane_base, xf);
        bar_arrive<BAR_X, N_MAIN>();
        m_all += 1u;
      };

      // ---- first iteration's context; its top stage goes to the recomputation chain right away --------------------
      set_ctx(0);
      piece_for(t, pc_t1, pc_inv, pc_v1, pc_dv);
      // stage inputs of the item whose pull-back starts next: computed ONCE, when the item is handed to the
      // recomputation chain (one slot earlier), and kept in registers
      float ys[NS], x[HODE_NN_IN], gdi;
      stage_x(N - 1, ys, x, gdi);
      start_F(x);
      uint32_t m_b = m_all - 1u;   // index of the item whose pull-back starts next

      for (int it = 0; it < n_iter; ++it) {
        HODE_TL(200);
        const int sidx = n - 1 - it;
        const int i_lo = item_i_lo(it, n_iter, N, i0, fsal);
        // ---- cotangents of this step's result and stage derivatives -------------------------------------------
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
                const float xq = __fdividef((float)(te - t), hf);   // (the rollout's own formula)
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
            gk[6][cc] += carry[cc];   // stage 1 of the later step = this step's stage 7
          }
        }

#pragma unroll 1
        for (int i = N - 1; i >= i_lo; --i) {
          HODE_TL(201);
          const bool has_F = !(it == n_iter - 1 && i == i_lo);
          float gys[NS], gki[NS];
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) { gys[cc] = 0.f; gki[cc] = 0.f; }
#pragma unroll
          for (int jj = 0; jj < NSTAGE_MAX; ++jj) {
            if (jj == i) {
#pragma unroll
              for (int cc = 0; cc < NS; ++cc) gki[cc] = gk[jj][cc];
            }
          }
          // ---- pull-back of this item: delta_L = the cotangent of the 6 network outputs, and x as the input operand
          // of dW_0 -> B chain (the chain's previous item is complete: its g_x has been read below)
          {
            float d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = (j < NS) ? gki[j] : 0.f;
            const uint32_t ph0 = m_b * (uint32_t)(L + 1);   // first phase of this item (buffers alternate by phase)
            uint8_t* db = smem_raw + OFF_DB + (ph0 & 1u) * DB_BYTES + row * 16;
            uint4 vh, vm;
            bf16_split8(d, vh, vm);
            *reinterpret_cast<uint4*>(db) = vh;
            *reinterpret_cast<uint4*>(db + ST_PART) = vm;
            *reinterpret_cast<uint4*>(db + ST_GRP) = make_uint4(0u, 0u, 0u, 0u);   // N = 16: features 8..15 are zero
            *reinterpret_cast<uint4*>(db + ST_GRP + ST_PART) = make_uint4(0u, 0u, 0u, 0u);
            // inputs of layer 0: the 9 stage features, zero padding, feature 15 = 1 (bias column)
            uint8_t* xb = smem_raw + OFF_XB + (m_b & 1u) * XB_BYTES + row * 16;
            const float xt[8] = {x[8], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
            bf16_split8(x, vh, vm);
            *reinterpret_cast<uint4*>(xb) = vh;
            *reinterpret_cast<uint4*>(xb + XB_PART) = vm;
            bf16_split8(xt, vh, vm);
            *reinterpret_cast<uint4*>(xb + ST_GRP) = vh;
            *reinterpret_cast<uint4*>(xb + ST_GRP + XB_PART) = vm;
            tc::fence_proxy_async();
            tc::fence_before_sync();
            bar_arrive<BAR_BX, N_MAIN>();
          }
          HODE_TL(202);
          // ---- recomputation of the NEXT item: its input operand ---------------------------------------------------
          float xf[HODE_NN_IN], ysf[NS], gdf = 0.f;
          if (has_F) {
            if (i > i_lo) {
              stage_x(i - 1, ysf, xf, gdf);
            } else {
              // top stage of the next iteration's step (one step earlier in time)
              const int sn = sidx - 1;
              const bool act_n = ok && (sn >= 0 || (fsal && sn == -1));
              double te;
              if (solver == 1) {
                // DP5(4): its stage 7 is evaluated at ITS result = the state and time this step starts from
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) ysf[cc] = y[cc];
                te = t;
              } else {
                // RK4: stage 4 at y' + h' k3', t' + h' from the (prefetched) next record
                double tp = t_first;
                float hp = 0.f, yp[NS], k3[NS];
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) { yp[cc] = 0.f; k3[cc] = 0.f; }
                if (ok && sn >= 0) {
                  const float4* r4 = reinterpret_cast<const float4*>(step_rec(A, unit, sn));
                  const float4 c0 = r4[0], c1 = r4[1], c2 = r4[2], c5 = r4[5], c6 = r4[6];
                  tp = __longlong_as_double(((long long)__float_as_int(c0.y) << 32) | (long long)(unsigned)__float_as_int(c0.x));
                  hp = c0.z;
                  yp[0] = c1.x; yp[1] = c1.y; yp[2] = c1.z; yp[3] = c1.w; yp[4] = c2.x; yp[5] = c2.y;
                  k3[0] = c5.z; k3[1] = c5.w; k3[2] = c6.x; k3[3] = c6.y; k3[4] = c6.z; k3[5] = c6.w;
                }
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) ysf[cc] = fmaf(hp, k3[cc], yp[cc]);
                te = tp + (double)hp;
              }
              const float t32 = (float)te;
              float tvns;
              inputs_at(t32, pn_t1, pn_inv, pn_v1, pn_dv, tvns, gdf);
              if (!act_n) {
                tvns = 0.f; gdf = 0.f;
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) ysf[cc] = 0.f;
              }
              xf[0] = t32;
#pragma unroll
              for (int cc = 0; cc < NS; ++cc) xf[1 + cc] = ysf[cc];
              xf[7] = ysf[3];
              xf[8] = tvns;
            }
            start_F(xf);
          }
          HODE_TL(203);
          // ---- while the chains run: the mechanistic VJP, and (once per step) the next step's input piece ------------
          mech_vjp(th, ys, gdi, gd_present, gki, gys, gth);
          if (i == N - 2 && it + 1 < n_iter) {
            const int sn = sidx - 1;
            double tn = t_first;
            if (ok && sn >= 0) tn = step_rec_t(step_rec(A, unit, sn));
            piece_for(tn, pn_t1, pn_inv, pn_v1, pn_dv);
          }
          __syncwarp();   // reconverge after per-thread code: tcgen05 .sync.aligned instructions follow
          HODE_TL(204);
          // ---- the pull-back's result: g_x ---------------------------------------------------------------------------
          tc::mbar_wait(&bars.gx_bar, m_b & 1u);
          tc::fence_after_sync();
          HODE_TL(205);
          {
            uint32_t v[16];
            HODE_TMEM_LD_X16(tmem + lane_base + TM_DB, v);
            tc::wait_ld();
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) gys[cc] += __uint_as_float(v[1 + cc]);
            gys[3] += __uint_as_float(v[7]);
          }
          tc::fence_before_sync();
          m_b += 1u;
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) {
            gy[cc] += gys[cc];
#pragma unroll
            for (int j = 0; j < NSTAGE_MAX - 1; ++j) gk[j][cc] = fmaf(hf * kA[solver][i][j], gys[cc], gk[j][cc]);
          }
          if (has_F) {   // the next item's stage inputs become the current ones
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) ys[cc] = ysf[cc];
#pragma unroll
            for (int kk = 0; kk < HODE_NN_IN; ++kk) x[kk] = xf[kk];
            gdi = gdf;
          }
          HODE_TL(206);
        }
        if (act) {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) lam[cc] = gy[cc];
        }
        if (fsal) {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) carry[cc] = real ? gk[0][cc] : 0.f;
        }
        if (real) {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) y_later[cc] = y[cc];
          t_later = t;
          later_valid = true;
        }
        // ---- next iteration's context ------------------------------------------------------------------------------
        if (it + 1 < n_iter) {
          set_ctx(it + 1);
          pc_t1 = pn_t1; pc_inv = pn_inv;
          pc_v1[0] = pn_v1[0]; pc_v1[1] = pn_v1[1]; pc_dv[0] = pn_dv[0]; pc_dv[1] = pn_dv[1];
        }
        HODE_TL(207);
      }
      {
        if (ok) {
          if (solver == 0) {
#pragma unroll
            for (int cc = 0; cc < NS; ++cc) lam[cc] += gtraj[cc];
          } else {
            for (; ei >= 0; --ei) {
              if ((double)in.t_obs[ei] > t_first) continue;
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
    tc::mbar_wait(&bars.done, 0u);
    tc::fence_after_sync();
    float* out = G.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(A.P + HODE_N_THETA);
    const int offo = 640 + (L - 1) * 4160;
    // M = 64 accumulators (layer 0 and the hidden layers): row j of D lives in TMEM lane 32 (j / 16) + j % 16
    // (csrc/probe/adj_probe.cu), i.e. in the first 16 lanes of every main warp.  The loads are warp-wide.
    {
      const int j = 16 * wq + lane_id;
      const bool owner = lane_id < 16;
      uint32_t v[16];
      // layer 0: D[j][k] (k < 9), column 15 = db_0[j]
      HODE_TMEM_LD_X16(tmem + lane_base + DW_0, v);
      tc::wait_ld();
      if (owner) {
#pragma unroll
        for (int kk = 0; kk < HODE_NN_IN; ++kk) out[j * HODE_NN_IN + kk] = have ? __uint_as_float(v[kk]) : 0.f;
        out[576 + j] = have ? __uint_as_float(v[15]) : 0.f;
      }
#pragma unroll
      for (int l = 1; l < MAXL; ++l) {
        if (l >= L) continue;
        const int off = 640 + (l - 1) * 4160;
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
          HODE_TMEM_LD_X16(tmem + lane_base + DW_H0 + DW_HS * (uint32_t)(l - 1) + 16u * (uint32_t)cidx, v);
          tc::wait_ld();
          if (owner) {
#pragma unroll
            for (int i = 0; i < 16; ++i) out[off + j * H + cidx * 16 + i] = have ? __uint_as_float(v[i]) : 0.f;
          }
        }
        uint32_t v8[8];
        HODE_TMEM_LD_X8(tmem + lane_base + DW_H0 + DW_HS * (uint32_t)(l - 1) + 64u, v8);
        tc::wait_ld();
        if (owner) out[off + 4096 + j] = have ? __uint_as_float(v8[0]) : 0.f;
      }
    }
    // output layer (transposed, M = 128: lane = row): D[k][n] = dW_out[n][k] for the input features k < 64
    if (wq < 2) {
      const int j = row;
      uint32_t v[16];
      HODE_TMEM_LD_X16(tmem + lane_base + DW_O, v);
      tc::wait_ld();
#pragma unroll
      for (int nn = 0; nn < NS; ++nn) out[offo + nn * H + j] = have ? __uint_as_float(v[nn]) : 0.f;
    }
    if (wq == 2) {   // TMEM lane 64: the constant-1 input feature -> db_out (warp-wide load)
      uint32_t v[16];
      HODE_TMEM_LD_X16(tmem + lane_base + DW_O, v);
      tc::wait_ld();
      if (lane_id == 0) {
#pragma unroll
        for (int nn = 0; nn < NS; ++nn) out[offo + 384 + nn] = have ? __uint_as_float(v[nn]) : 0.f;
      }
    }
    // deterministic reduction of the theta gradients over the 128 trajectory slots (the operand buffers are dead:
    // every MMA has completed)
    float* red = reinterpret_cast<float*>(smem_raw);
#pragma unroll
    for (int i = 0; i < HODE_N_THETA; ++i) red[i * TILE + row] = gth[i];
    asm volatile("bar.sync %0, 128;" ::"n"(BAR_MAIN) : "memory");
    if (row < HODE_N_THETA) {
      float sacc = 0.f;
      for (int r = 0; r < TILE; ++r) sacc += red[row * TILE + r];
      out[A.P + row] = sacc;
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
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
  // (the plan does not depend on the MLP arithmetic: both split-precision images have the same size)
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
  p.t_in_smem = (!t_per_traj && ADJ_SMEM_BASE + 4 * T <= ADJ_SMEM_MAX) ? 1 : 0;
  p.smem = (size_t)ADJ_SMEM_BASE + (p.t_in_smem ? (size_t)4 * T : 0);
  p.partial_floats = (size_t)gx * S * (size_t)(P + HODE_N_THETA);
  // activation scratch: per CTA two sets (the item being recomputed / the item being pulled back) of L operand images
  p.stash_floats = (size_t)gx * S * (size_t)2 * L * (ST_BLK / 4);
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

cudaError_t launch_rollout_bwd_tc(const RolloutArgs& A, int mlp_mode, const float* grad_traj, float* grad_y0,
                                  float* grad_theta, float* grad_W, void* workspace, cudaStream_t stream) {
  if (A.rec_floats < HODE_REC_FLOATS_K || !A.save_k1) return cudaErrorInvalidValue;   // needs the stage derivatives
  const AdjTcPlan p = adj_tc_plan(A.B, A.S, A.L, A.P, A.T, A.t_per_traj);
  if (p.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  auto al = [](size_t x) { return (x + 63) & ~(size_t)63; };
  AdjTcArgs G{};
  G.R = A;
  G.grad_traj = grad_traj;
  G.grad_y0 = grad_y0;
  G.partials = reinterpret_cast<float*>(workspace);
  G.stash = reinterpret_cast<uint8_t*>(G.partials + al(p.partial_floats));
  float* imgs = G.partials + al(p.partial_floats) + al(p.stash_floats);
  G.t_in_smem = p.t_in_smem;
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
  // the recomputation uses the rollout's arithmetic (its activations are then the forward pass's, bit for bit); a
  // single-pass TF32 rollout is differentiated with the 3xTF32 recomputation
  const int fwd_mode = mlp_mode == HODE_MLP_TF32BF16 ? HODE_MLP_TF32BF16 : HODE_MLP_TF32X3;
  cudaError_t e = tc_prepare_fwd_images(A.W, imgs, A.S, A.L, A.P, fwd_mode, stream);
  if (e != cudaSuccess) return e;
  count_launch();
  prep_tc_bwd_image_kernel<<<A.S, 256, 0, stream>>>(A.W, imgs + (size_t)A.S * p.fwd_floats, A.L, A.P, p.bwd_floats);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  {
    // the composite key holds the parameter-set index above SORT_N_BITS bits of step count
    int s_bits = 0;
    while ((1L << s_bits) < (long)A.S) ++s_bits;
    const int sorted = (s_bits + (int)SORT_N_BITS <= 32) ? 1 : 0;
    count_launch();
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
    count_launch();
    adj_schedule_kernel<<<A.S, 256, 0, stream>>>(sched_keys, A.B, p.n_tiles, p.grid_x, sorted, owner, sched_off,
                                                 sched_tiles);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (fwd_mode == HODE_MLP_TF32BF16) {
    e = cudaFuncSetAttribute(rollout_bwd_tc_kernel<MLP_MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    count_launch();
    rollout_bwd_tc_kernel<MLP_MIXED><<<dim3(p.grid_x, p.grid_y), ADJ_THREADS, p.smem, stream>>>(G);
  } else {
    e = cudaFuncSetAttribute(rollout_bwd_tc_kernel<MLP_X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    count_launch();
    rollout_bwd_tc_kernel<MLP_X3><<<dim3(p.grid_x, p.grid_y), ADJ_THREADS, p.smem, stream>>>(G);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return launch_reduce_partials(G.partials, p.grid_x, A.S, A.P, grad_W, grad_theta, stream);
}

}  // namespace hode

#ifdef HODE_DEBUG_WAIT
// debug build only (tools/debug_wait_adj.py): where stuck mbarrier waits are recorded (host-mapped memory)
extern "C" int hode_debug_set_buffer(int* buf) {
  return (int)cudaMemcpyToSymbol(hode::tc::g_dbg_buf, &buf, sizeof(int*));
}
#endif
