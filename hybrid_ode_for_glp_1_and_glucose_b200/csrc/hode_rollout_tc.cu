// hode_rollout_tc.cu — the tensor-core rollout: RK stages of 128 trajectories in lock-step,
// residual MLP as a chain of tcgen05 TF32 MMAs with activations and accumulators in TMEM.
//
// BASELINE.json north_star items (1)+(2): the Dormand-Prince / RK4 stepper state (6-state
// mechanistic RHS, stage vectors, error norm, per-trajectory step control) lives in registers,
// one trajectory per thread; the 9->64->..->64->6 MLP of every RK stage is evaluated for a
// TILE of 128 trajectories as dense [128 x K] x [K x N] contractions on the 5th-gen tensor
// cores.  Replaces reference models/hybrid_ode_nn.py:184-256 + models/nn_residual.py:136-146.
//
// CTA = 2 tiles x 128 threads (persistent, one CTA per SM).  Per tile, TMEM columns:
//   [0,64) accumulator D | [64,128) A_hi (TF32) | [128,160) bf16(A_hi) | [160,192) bf16(A - A_hi) |
//   [192,200) constant [1,1,0..] (bias step)
// Per layer:  epilogue threads read D (tcgen05.ld; the bias is already in it), ReLU, split into a TF32
// high part and the BF16 operands of the two cross terms, and write the next layer's A operand straight
// back to TMEM (tcgen05.st); one elected thread issues
//   D = bias + bf16(A_lo)*bf16(B_hi) + bf16(A_hi)*bf16(B_lo) + A_hi*B_hi
// (one kind::tf32 pass + two kind::f16 BF16 passes at twice its rate: fp32-equivalent accuracy for 2
// pass-equivalents, hode_tc_mlp.cuh; single TF32 pass in HODE_MLP_TF32 mode) with B = pre-split weights in
// shared memory (K-major, no swizzle), then tcgen05.commit -> mbarrier.  The two tiles of a
// CTA run out of phase, so one tile's epilogue overlaps the other tile's MMAs.
// Weights are staged once per (CTA, parameter set) with one bulk async copy (TMA engine) of a
// pre-split image built by prep_tc_image_kernel.
// Trajectories are pulled lane by lane from a global queue, so a lane that finishes early
// (adaptive steps diverge 5x between trajectories) is refilled instead of idling.
#include <math.h>
#include <stdlib.h>

#include "hode_common.cuh"
#include "hode_kernels.h"
#include "hode_tcgen05.cuh"
#include "hode_tc_mlp.cuh"

namespace hode {

namespace {
// MLP_MIX3: three tiles of 128 trajectories per CTA and no helper warps (384 threads x 168 registers; 3 x 160 + 8 TMEM
// columns); the other modes: two tiles, each with a helper warpgroup (512 threads, setmaxnreg 200 / 56).
// MLP_H16: the same shape and TMEM footprint.
template <int MODE> constexpr int tiles_per_cta() { return three_tiles<MODE>() ? 3 : 2; }
template <int MODE> constexpr bool has_helpers() { return !three_tiles<MODE>(); }
template <int MODE> constexpr int cta_threads() { return tiles_per_cta<MODE>() * TILE * (has_helpers<MODE>() ? 2 : 1); }
constexpr int N_KSTAGE = 7 * NS;   // floats of the stage-derivative store per trajectory (k1..k7)
}  // namespace

// ---- weight image ----------------------------------------------------------------------------------
// Per layer (hode_tc_mlp.cuh IMG_*): [B_hi: TF32 rounding of W, fp32 words][second half], K-major without swizzle.
// Element (n, k) of an [N, K] weight:
//   TF32 parts at float  ((k / 4) * N + n) * 4 + k % 4         (16-byte K chunks of 4 floats)
//   BF16 parts at 2-byte ((k / 8) * N + n) * 8 + k % 8         (16-byte K chunks of 8 halves)
// second half: 3xTF32 -> B_lo = TF32 of (W - B_hi); mixed -> bf16(B_hi) then bf16(W - B_hi).
// Layer 0 is K = 16: the 9 input features, then feature 9 = the constant 1 whose weight column is the layer's
// bias, then zero padding.  The other layers get a bias block: one K = 8 TF32 step whose columns 0/1 hold the
// TF32 hi/lo parts of the bias; it multiplies a constant A block [1, 1, 0, ...] kept in TMEM, so the bias rides
// on the tensor pipe instead of costing one FADD per accumulator element in the epilogue.
int tc_image_floats_mode(int L, int mlp_mode);
int tc_image_floats(int L) { return 2 * 1024 + (L - 1) * 2 * 4096 + 2 * 1024 + L * 512 + 128; }
// MLP_MIX3 image: [B_hi tf32][B_lo tf32][bf16(B_hi)] per layer
static int tc_image_floats_max(int L) {   // the workspace layout does not depend on the mode
  const int a = tc_image_floats(L), b = tc_image_floats_mode(L, HODE_MLP_TF32X2BF16);
  return a > b ? a : b;
}
int tc_image_floats_mode(int L, int mlp_mode) {
  if (mlp_mode == HODE_MLP_F16BF16X2)
    return (int)(img_l0<MLP_H16>() + (uint32_t)(L - 1) * img_hid<MLP_H16>() + img_out<MLP_H16>()) + L * 512 + 128;
  if (mlp_mode != HODE_MLP_TF32X2BF16) return tc_image_floats(L);
  return (int)(img_l0<MLP_MIX3>() + (uint32_t)(L - 1) * img_hid<MLP_MIX3>() + img_out<MLP_MIX3>()) + L * 512 + 128;
}

// mixed: 0 = 3xTF32 (second half B_lo), 1 = MLP_MIXED (second half bf16(B_hi), bf16(B_lo)),
//        2 = MLP_MIX3 (B_lo, then bf16(B_hi): 2.5 parts per layer),
//        3 = MLP_H16 ([f16(W)][f16(64 (W - f16(W)))][bf16(f16(W))], 2-byte elements: 1.5 float-sized parts per layer)
__global__ void prep_tc_image_kernel(const float* __restrict__ W, float* __restrict__ img, int L, int P,
                                     int img_floats, int mixed) {
  const float* w = W + (size_t)blockIdx.x * P;
  float* out = img + (size_t)blockIdx.x * img_floats;
  for (int i = threadIdx.x; i < img_floats; i += blockDim.x) out[i] = 0.f;
  __syncthreads();
  int n_in = HODE_NN_IN;
  float* dst = out;
  float* bias_dst = out + img_floats - (L * 512 + 128);
  for (int l = 0; l <= L; ++l) {
    const int n_out = (l == L) ? NS : H;
    const int Npad = (l == L) ? 16 : H;
    const int Kpad = (l == 0) ? 16 : H;
    const int part = Kpad * Npad;   // floats of each half
    uint16_t* hib = reinterpret_cast<uint16_t*>(dst + (mixed == 2 ? 2 * part : part));
    uint16_t* lob = reinterpret_cast<uint16_t*>(dst + part + part / 2);
    const int n_cols = n_in + (l == 0 ? 1 : 0);   // layer 0: column n_in is the bias
    for (int i = threadIdx.x; i < n_out * n_cols; i += blockDim.x) {
      const int n = i / n_cols, k = i - n * n_cols;
      const float wv = (k < n_in) ? w[n * n_in + k] : w[n_out * n_in + n];
      if (mixed == 3) {
        uint16_t* h16 = reinterpret_cast<uint16_t*>(dst);
        const int ob16 = ((k >> 3) * Npad + n) * 8 + (k & 7);
        split_h16_weight(wv, h16[ob16], h16[part + ob16], h16[2 * part + ob16]);
        continue;
      }
      uint32_t hi, lo;
      tc::split_tf32(wv, hi, lo);
      const int o = ((k >> 2) * Npad + n) * 4 + (k & 3);
      dst[o] = __uint_as_float(hi);
      const int ob = ((k >> 3) * Npad + n) * 8 + (k & 7);
      if (mixed == 1) {
        hib[ob] = (uint16_t)(pack_bf16x2(__uint_as_float(hi), 0.f) & 0xFFFFu);
        lob[ob] = (uint16_t)(pack_bf16x2(wv - __uint_as_float(hi), 0.f) & 0xFFFFu);
      } else {
        dst[part + o] = __uint_as_float(lo);
        if (mixed == 2) hib[ob] = (uint16_t)(pack_bf16x2(__uint_as_float(hi), 0.f) & 0xFFFFu);
      }
    }
    if (l > 0) {
      for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
        uint32_t hi, lo;
        tc::split_tf32(w[n_out * n_in + j], hi, lo);
        bias_dst[j * 4 + 0] = __uint_as_float(hi);   // (n = j, k = 0)
        bias_dst[j * 4 + 1] = __uint_as_float(lo);   // (n = j, k = 1)
      }
    }
    w += n_out * n_in + n_out;
    dst += (mixed == 3) ? part + part / 2 : ((mixed == 2) ? 2 * part + part / 2 : 2 * part);
    bias_dst += (l == L) ? 128 : 512;
    n_in = n_out;
  }
}

cudaError_t tc_prepare_fwd_images(const float* W, float* img, int S, int L, int P, int mlp_mode, cudaStream_t stream) {
  count_launch();
  prep_tc_image_kernel<<<S, 256, 0, stream>>>(W, img, L, P, tc_image_floats_mode(L, mlp_mode),
                                              mlp_mode == HODE_MLP_TF32BF16 ? 1 : (mlp_mode == HODE_MLP_TF32X2BF16 ? 2 :
                                              (mlp_mode == HODE_MLP_F16BF16X2 ? 3 : 0)));
  return cudaGetLastError();
}

// ---- per-lane trajectory state ---------------------------------------------------------------------
struct Lane {
  TrajInputs in;
  float y[NS], cmp[NS];
  double t;
  long unit;          // flat (sample, trajectory) index, -1 = idle
  long b;             // trajectory index
  float* out;         // this trajectory's [T,6] output rows (or nullptr)
  int ei;             // next observation index to emit
  int status, n_acc, n_rej, n_saved;
  bool has;
  // Inputs of the current step as one linear piece per channel: value(t) = c_v1 + alpha * c_dv
  // with alpha = (t - c_t1) / c_dt.  Valid whenever a step cannot cross an input kink (RK4, kink
  // clipping, or no series inputs): the loads happen once per accepted step instead of once per
  // RK stage, which takes the L2 latency of the input rows off the stage critical path.
  bool cached;
  float c_t1, c_inv_dt, c_v1[3], c_dv[3];
  bool c_stale;         // the step start has crossed a kink since the cache was loaded
};

// (i0, i0+1): the grid interval the step starts in.  Flat stretches that a step may cross have
// c_dv == 0, so the value does not depend on which of their intervals is cached.
__device__ __forceinline__ void lane_cache_inputs(Lane& ln, int i0) {
  const float t1 = ln.in.t_obs[i0], t2 = ln.in.t_obs[i0 + 1];
  ln.c_t1 = t1;
  ln.c_inv_dt = 1.0f / (t2 - t1);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (ln.in.mode[ch] != HODE_IN_SERIES) continue;
    const float v1 = ln.in.u[ch][i0], v2 = ln.in.u[ch][i0 + 1];
    ln.c_v1[ch] = v1;
    ln.c_dv[ch] = v2 - v1;
  }
}

__device__ __forceinline__ float rms6v(const float* v) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) s = fmaf(v[i], v[i], s);
  return sqrtf(s * (1.0f / NS));
}

// f_physio for the tensor-core path: same formulas as rhs_mech (hode_common.cuh), but with
// reciprocal-based division and free FMA contraction — this path is accurate to float32 round-off,
// not bit-identical to the CPU evaluation (the 3xTF32 network products are not either), and the
// three IEEE divisions of rhs_mech sit on the critical path of every RK stage.
__device__ __forceinline__ void rhs_mech_fast(const Theta& p, const float* y, float meal, float GD,
                                              bool gd_present, float* d) {
  const float G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const float Pi = fmaf(p.rho, GLP1, 1.0f);
  d[1] = Pi * p.a_GI * (G - p.G_b) - p.k_I * (I - p.I_b);
  const float glp1_effect = p.E_max * __fdividef(GLP1, p.EC_50 + GLP1);
  d[2] = -glp1_effect * (Glu - p.Glu_b);
  d[3] = p.V_max * __fdividef(G, p.K_m + G) - p.k_L * GLP1;
  float k_GE = p.kge0;
  if (gd_present) {
    const float gdg = __powf(GD, p.g);
    k_GE = p.k_GE0 * (1.0f - __fdividef(gdg, p.igd_pow + gdg));
  }
  d[5] = FFA * (-p.p_7 - p.p_8 * I + p.p_9 * G);
  d[0] = meal - 0.01f * (I - p.I_b) + 0.005f * (Glu - p.Glu_b) - k_GE * G;
  d[4] = 0.0f;
}

// the 19 floats stage_theta() wrote: no powf here
__device__ __forceinline__ Theta load_theta_staged(const float* __restrict__ th) {
  Theta p;
  p.a_GI = th[0]; p.k_I = th[1]; p.rho = th[2]; p.G_b = th[3]; p.I_b = th[4];
  p.E_max = th[5]; p.EC_50 = th[6]; p.Glu_b = th[7]; p.V_max = th[8]; p.K_m = th[9];
  p.k_L = th[10]; p.k_GE0 = th[11]; p.IGD_50 = th[12]; p.g = th[13];
  p.p_7 = th[14]; p.p_8 = th[15]; p.p_9 = th[16];
  p.igd_pow = th[17];
  p.kge0 = th[18];
  return p;
}

// nextafter(t, +inf) - t for a finite t (SciPy's min_step = 10 ulp, rk.py:121), by stepping the bit pattern
__device__ __forceinline__ double ulp_above(double t) {
  if (t == 0.0) return 4.9406564584124654e-324;
  const long long b = __double_as_longlong(t);
  return __longlong_as_double(t > 0.0 ? b + 1 : b - 1) - t;
}

// barrier + OR-reduction of a predicate over the n threads of named barrier `bar` in one instruction
__device__ __forceinline__ bool tile_or(int bar, int n, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 q, %3, 0;\n\t"
      "bar.red.or.pred p, %1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r) : "r"(bar), "r"(n), "r"((uint32_t)pred) : "memory");
  return r != 0;
}

// One evaluation of f_physio + g_NN for this lane (tile-collective).
// th_sm: the parameter set's 19 floats (17 parameters, IGD_50^g, k_GE at GD = 0) in shared memory; th_row: this
// trajectory's own 17 parameters in global memory (theta_per_traj mode) or nullptr.  The parameters are fetched
// where they are used, behind the layer-0 MMAs, instead of living in 19 registers across the whole evaluation.
template <int X3>
__device__ __forceinline__ void lane_eval(TileCtx& c, const float* th_sm, const float* th_row, Lane& ln, double te,
                                          const float* ys, float* d) {
  const float t32 = (float)te;
  float meal = 0.f, tvns = 0.f, gd = 0.f;
  if (ln.has) {
    if (ln.cached) {
      // same arithmetic as input_channel(): v1 + ((t - t1) / (t2 - t1)) * (v2 - v1)
      const float alpha = (t32 - ln.c_t1) * ln.c_inv_dt;
      meal = fmaf(alpha, ln.c_dv[HODE_CH_MEAL], ln.c_v1[HODE_CH_MEAL]);
      tvns = fmaf(alpha, ln.c_dv[HODE_CH_TVNS], ln.c_v1[HODE_CH_TVNS]);
      gd = fmaf(alpha, ln.c_dv[HODE_CH_GD], ln.c_v1[HODE_CH_GD]);
    } else {
      int idx = 0;
      if (any_series(ln.in)) idx = grid_index_from(ln.in, t32, ln.in.cur);
      meal = input_channel(ln.in, HODE_CH_MEAL, t32, idx);
      tvns = input_channel(ln.in, HODE_CH_TVNS, t32, idx);
      gd = input_channel(ln.in, HODE_CH_GD, t32, idx);
    }
  }
  float x[HODE_NN_IN], r[NS];
  x[0] = t32;
#pragma unroll
  for (int i = 0; i < NS; ++i) x[1 + i] = ys[i];
  x[7] = ys[3];
  x[8] = tvns;
  __syncwarp();
  mlp_tile<X3>(c, x, r, [&] {
    const Theta th = th_row ? load_theta(th_row) : load_theta_staged(th_sm);
    rhs_mech_fast(th, ys, meal, gd, ln.in.mode[HODE_CH_GD] != HODE_IN_ABSENT, d);
  });
#pragma unroll
  for (int i = 0; i < NS; ++i) d[i] = __fadd_rn(d[i], r[i]);
}

__device__ __forceinline__ void lane_bind(Lane& ln, const RolloutArgs& A, const float* t_shared,
                                          int s, long b) {
  const long unit = (long)s * A.B + b;
  ln.unit = unit;
  ln.b = b;
  ln.has = true;
  ln.in.T = A.T;
  ln.in.cur = 0;
  ln.in.t_obs = A.t_per_traj ? A.t_obs + b * A.T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    ln.in.mode[ch] = A.in_mode[ch];
    ln.in.u[ch] = A.in_mode[ch] == HODE_IN_SERIES ? A.u[ch] + b * A.T
                : A.in_mode[ch] == HODE_IN_CONST ? A.u[ch] + b : nullptr;
  }
  ln.out = A.traj ? A.traj + (size_t)unit * A.T * (A.out_nc ? A.out_nc : NS) : nullptr;
  ln.cached = A.solver == HODE_SOLVER_RK4 || A.kink_mode == HODE_KINK_CLIP || !any_series(ln.in);
  ln.c_t1 = 0.f;
  ln.c_inv_dt = 1.f;
  ln.c_stale = true;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    ln.c_v1[ch] = ln.in.mode[ch] == HODE_IN_CONST ? ln.in.u[ch][0] : 0.f;
    ln.c_dv[ch] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) { ln.y[i] = A.y0[b * NS + i]; ln.cmp[i] = 0.f; }
  ln.ei = 0;
  ln.status = HODE_ST_OK;
  ln.n_acc = ln.n_rej = ln.n_saved = 0;
}

__device__ __forceinline__ void lane_finish(Lane& ln, const RolloutArgs& A, int vi_n) {
  const long n_units = (long)A.S * A.B;
  if (ln.out || vi_n) {
    const float z[NS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (; ln.ei < A.T; ++ln.ei) emit_row(A, ln.out, ln.b, ln.ei, z, vi_n);
  }
  if (A.status) A.status[ln.unit] = ln.status;
  if (A.counters) {
    A.counters[ln.unit] = ln.n_acc;
    A.counters[n_units + ln.unit] = ln.n_rej;
  }
  if (A.save_n) A.save_n[ln.unit] = ln.status == HODE_ST_OK ? ln.n_saved : -1 - ln.n_saved;  // < 0: no gradient
  if (A.done_count) {
    __threadfence();   // this trajectory's outputs are visible device-wide before it is counted
    const int blk = (int)(ln.b / A.done_block);
    const long first = (long)blk * A.done_block;
    const int size = (int)((first + A.done_block <= (long)A.B) ? A.done_block : (long)A.B - first);
    if (atomicAdd(&A.done_count[blk], 1) + 1 == size) {
      __threadfence_system();
      A.done_flag[blk] = 1;
    }
  }
  ln.has = false;
  ln.unit = -1;
}

// The step's record: start time, size, start state and — for the tensor-core adjoint — its stage derivatives
// k1 .. k(1 + n_extra) (DP5(4): k1..k6; RK4: k1..k3): with them every stage INPUT of the step is a linear
// combination the adjoint evaluates directly, so it recomputes the stages one at a time, in reverse order.
__device__ __forceinline__ void lane_save_step(Lane& ln, const RolloutArgs& A, double t, float hf, const float (*k)[NS],
                                               int n_extra) {
  float* r = step_rec(A, ln.unit, ln.n_saved);
  step_rec_store(r, t, hf, ln.y, A.save_k1 ? k[0] : nullptr);
  if (A.rec_floats >= HODE_REC_FLOATS_K) step_rec_store_stages(r, k, n_extra);
  ++ln.n_saved;
}

// Bit i set <=> grid point i is a kink of some series input (T <= 64).  Built once per
// trajectory with independent loads, so the first-touch latency of the input rows is paid with
// memory-level parallelism instead of one dependent miss per grid point during stepping.
__device__ __forceinline__ unsigned long long build_kink_mask(const TrajInputs& in) {
  unsigned long long diff = 0ull;   // bit i: v[i] != v[i+1] in some series channel
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (in.mode[ch] != HODE_IN_SERIES) continue;
    const float* v = in.u[ch];
#pragma unroll 8
    for (int i = 0; i + 1 < in.T; ++i)
      if (v[i] != v[i + 1]) diff |= 1ull << i;
  }
  unsigned long long m = (diff | (diff << 1));          // kink(i) = diff(i-1) | diff(i)
  m &= ~1ull;                                             // i = 0 is never a kink
  if (in.T >= 1) m &= ~(1ull << (in.T - 1));             // nor is the last point
  return m;
}

// ---------------------------------------------------------------------------------------------------
// The kernel.  grid = number of SMs (persistent), block = 256 threads = 2 tiles.
// queue[s]: next unclaimed trajectory of parameter set s (zeroed before launch).
// One "round" = one RK4 step (4 evaluations) or one DP5(4) attempt (6 evaluations) for every lane
// of the tile; the evaluations of a round go through ONE call site of the tile MLP (slot loop),
// which keeps the kernel small enough for the instruction cache.
// ---------------------------------------------------------------------------------------------------
template <int X3, int SOLVER>
__global__ void __launch_bounds__(cta_threads<X3>(), 1)
rollout_tc_kernel(const RolloutArgs A, const float* __restrict__ img_g, int img_floats, int* queue) {
  constexpr int NT = tiles_per_cta<X3>();
  constexpr int NMAIN = NT * TILE;
  constexpr bool HELP = has_helpers<X3>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* img = reinterpret_cast<float*>(smem_raw);
  // stage derivatives k1..k7 of every trajectory of the CTA: component (j, i) of main thread m at
  // k_sm[(6 j + i) * NMAIN + m] (conflict-free); they are touched a few times per evaluation, so they
  // live here instead of 42 registers per thread
  float* k_sm = img + ((img_floats + 3) & ~3);
  float* t_sh_buf = k_sm + N_KSTAGE * NMAIN;
  __shared__ __align__(8) uint64_t mma_bar[NT];
  __shared__ __align__(8) uint64_t load_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int cta_queue;   // fused posterior-predictive mode: next trajectory of this CTA's range
  __shared__ float th_sm[20];  // the current parameter set's mechanistic parameters + derived values

  const int tid = threadIdx.x, lane_id = tid & 31;
  // warp-uniform by construction; the shuffle lets ptxas keep everything derived from it
  // (tile, TMEM base, barrier ids) in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // warpgroups: 0..NT-1 = main warps of the tiles (one trajectory per thread), NT.. = their helpers
  const int wg = warp >> 2, wq = warp & 3;
  const bool helper = HELP && wg >= NT;
  const int tile = helper ? wg - NT : wg;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NT; ++i) tc::mbar_init(&mma_bar[i], 1);
    tc::mbar_init(&load_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  const float* t_shared = nullptr;
  if (!A.t_per_traj && A.T <= HODE_SIMT_MAX_SHARED_T) {
    for (int i = tid; i < A.T; i += blockDim.x) t_sh_buf[i] = A.t_obs[i];
    t_shared = t_sh_buf;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();

  TileCtx c;
  c.img = img;
  c.mma_bar = &mma_bar[tile];
  c.tmem = tmem_base_s + (uint32_t)tile * tm_tile_stride<X3>();
  c.t_ones = three_tiles<X3>() ? tmem_base_s + TM3_ONES_ABS : c.tmem + TM_ONES;
  c.lane_base = (uint32_t)(wq * 32) << 16;
  c.parity = 0;
  c.bar_id = 1 + tile;
  c.bar_all = 1 + NT + tile;
  c.wq = wq;
  c.L = A.L;
  {  // constant A block of the bias k-step: this thread's TMEM lane gets [1, 1, 0, 0, 0, 0, 0, 0]
     // (MLP_MIX3: one block for the CTA; every tile's thread of a lane writes the same values)
    uint32_t ones[8] = {0x3F800000u, 0x3F800000u, 0u, 0u, 0u, 0u, 0u, 0u};
    HODE_TMEM_ST_X8(c.t_ones + c.lane_base, ones);
    tc::wait_st();
  }
  // all barriers of a MLP_MIX3 tile are over its 128 main threads
  float* const kp = k_sm + (tid < NMAIN ? tid : 0);
#define KS(j, i) kp[((j) * NS + (i)) * NMAIN]

  constexpr int NSLOT = (SOLVER == HODE_SOLVER_RK4) ? 4 : 6;
  const int T = A.T;
  const float rtol = A.rtol, atol = A.atol;
  const int max_steps = A.max_steps > 0 ? A.max_steps : 100000;
  const int nsub = A.n_substeps > 0 ? A.n_substeps : 1;
  const bool clip = (A.kink_mode == HODE_KINK_CLIP);
  uint32_t load_parity = 0;

  // Fused posterior-predictive mode: every trajectory belongs to ONE CTA for all S parameter sets
  // (the running mean/std of a trajectory is then updated in a fixed order without atomics).
  const bool vi = A.vi_mean != nullptr;
  const long per_cta = ((long)A.B + gridDim.x - 1) / gridDim.x;
  const long b_lo = vi ? (long)blockIdx.x * per_cta : 0;
  const long b_hi = vi ? (b_lo + per_cta < (long)A.B ? b_lo + per_cta : (long)A.B) : (long)A.B;

  // 512 threads x 128 registers fill the register file; the helper warpgroups hand most of theirs
  // to the main warpgroups, which carry the per-trajectory integrator state.  Each role's code is
  // entirely inside its branch so that ptxas allocates registers against the role's own budget.
  if (helper) {
    if (HELP) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
    for (int si = 0; si < A.S; ++si) {
      const int s = (int)((blockIdx.x + (unsigned)si) % (unsigned)A.S);
      __syncthreads();
      if (tid == 0) cta_queue = 0;   // (tid 0 is a main thread; kept for symmetry of the barriers)
      __syncthreads();
      (void)s;
      tc::mbar_wait(&load_bar, load_parity);
      load_parity ^= 1u;
      // mirror of the main warps' round structure: same tile-wide barriers, NSLOT MLP calls per round
      for (;;) {
        if (!tile_or(c.bar_all, 2 * TILE, false)) break;
#pragma unroll 1
        for (int slot = 0; slot < NSLOT; ++slot) mlp_tile_helper<X3>(c);
      }
    }
  } else {
    if (HELP) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
    for (int si = 0; si < A.S; ++si) {
      const int s = (int)((blockIdx.x + (unsigned)si) % (unsigned)A.S);
      const int vi_n = vi ? si + 1 : 0;
      // ---- stage this parameter set's weight image (one bulk copy) ------------------------------
      __syncthreads();  // both tiles are done with the previous image (and, in vi mode, sample)
      if (tid == 0) cta_queue = 0;
      __syncthreads();
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)img_floats * 4u;
        tc::mbar_expect_tx(&load_bar, bytes);
        tc::bulk_g2s(img, img_g + (size_t)s * img_floats, bytes, &load_bar);
      }
      tc::mbar_wait(&load_bar, load_parity);
      load_parity ^= 1u;
      if (tid == 0 && !A.theta_per_traj) {   // (the previous set's readers are behind the barriers above)
        const Theta th0 = load_theta(A.theta + (size_t)s * HODE_N_THETA);
#pragma unroll
        for (int i = 0; i < HODE_N_THETA; ++i) th_sm[i] = A.theta[(size_t)s * HODE_N_THETA + i];
        th_sm[17] = th0.igd_pow;
        th_sm[18] = th0.kge0;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(15), "r"(NMAIN) : "memory");   // main threads only: th_sm is staged

      Lane ln;
      ln.has = false;
      ln.unit = -1;
      ln.b = 0;
      ln.out = nullptr;
      ln.ei = 0;
      ln.status = 0; ln.n_acc = 0; ln.n_rej = 0; ln.n_saved = 0;
      ln.t = 0.0;
      ln.cached = true; ln.c_t1 = 0.f; ln.c_inv_dt = 1.f; ln.c_stale = true;
  #pragma unroll
      for (int ch = 0; ch < 3; ++ch) { ln.c_v1[ch] = 0.f; ln.c_dv[ch] = 0.f; }
      ln.in.T = T; ln.in.cur = 0; ln.in.t_obs = A.t_obs;
  #pragma unroll
      for (int ch = 0; ch < 3; ++ch) { ln.in.mode[ch] = HODE_IN_ABSENT; ln.in.u[ch] = nullptr; }
  #pragma unroll
      for (int i = 0; i < NS; ++i) { ln.y[i] = 0.f; ln.cmp[i] = 0.f; }
      bool queue_dry = false;
      int q_next = 0, q_end = 0;   // this warp's current chunk of the trajectory queue (warp-uniform)
      // Static first hand-out (plain rollout): main warp p = (tile * gridDim.x + blockIdx.x) * 4 + wq of the grid takes
      // queue positions [32 p, 32 p + 32) without touching the queue counter, which then serves positions from
      // n_static on.  Tile-major over the SMs: a cohort smaller than the grid's lanes (config 3's 32 768-trajectory
      // shard against 148 x 384 lanes) fills tile 0 of EVERY SM before any SM runs a second tile, instead of
      // whichever CTAs reach the counter first running three full tiles while the rest of the chip idles
      // (32 768 trajectories, longest first: 3.45 -> 3.28 ms).
      bool first_pull = !vi;
      const long n_static = vi ? 0L : min((long)A.B, (long)gridDim.x * NT * 4 * 32);
      const long guided_div = max(1L, (vi ? NT * 4 : (long)gridDim.x * NT * 4) / 2);   // half the warps that share the queue
      // stage vectors: KS(0, .) = k1 (FSAL) ... KS(6, .) = k7;  RK4 uses k1..k4
  #pragma unroll
      for (int j = 0; j < 7; ++j)
  #pragma unroll
        for (int i = 0; i < NS; ++i) KS(j, i) = 0.f;
      double h_abs = 0.0, t_stop = 0.0, t_bound = 0.0;
      unsigned long long kink_mask = 0ull;
      int attempts = 0, kink_cur = 1;
      bool need_init = false, prev_rejected = false, need_stop = true;
      int rk_n = 0, rk_ss = 0;

      for (;;) {
        HODE_TL(100);
        // ---- refill idle lanes from the global queue (warp-aggregated) --------------------------
        {
          const bool want_any = !ln.has && !queue_dry;
          const unsigned m_any = __ballot_sync(0xffffffffu, want_any);
          bool want = false;
          long b = 0;
          if (m_any) {
            // The warp draws trajectories from the queue in chunks (one atomic per chunk instead of one per refill)
            // and hands them to its idle lanes in lane order.  Guided self-scheduling: chunks of 32 while the queue is
            // long, shrinking to exactly what the idle lanes need near its end — a warp that sits on an unserved chunk
            // while other warps have run dry was the tail of the kernel (8.1 ms with all lanes refilling at once,
            // 11.0 ms with fixed chunks of 32, on the 262 144-trajectory bench cohort).
            const int n_want = __popc(m_any);
            const int rank = __popc(m_any & ((1u << lane_id) - 1u));
            int served = min(q_end - q_next, n_want);      // from what is left of the previous chunk
            if (want_any && rank < served) { want = true; b = b_lo + (long)q_next + rank; }
            q_next += served;
            if (served < n_want) {
              const int need = n_want - served;
              const long rem = (b_hi - b_lo) - (long)q_end;   // (stale: the queue head as this warp last saw it)
              int chunk = (int)min(32L, max((long)need, rem / guided_div));
              if (first_pull) {
                first_pull = false;
                chunk = 32;
                q_next = (((int)tile * (int)gridDim.x + (int)blockIdx.x) * 4 + wq) * 32;
              } else {
                int base = 0;
                if (lane_id == 0) base = vi ? atomicAdd(&cta_queue, chunk) : (int)n_static + atomicAdd(&queue[s], chunk);
                q_next = __shfl_sync(0xffffffffu, base, 0);
              }
              q_end = q_next + chunk;
              if (want_any && rank >= served) { want = true; b = b_lo + (long)q_next + (rank - served); }
              q_next += need;
            }
          }
          {
            if (want) {
              if (b < b_hi) {
                if (A.order) b = (long)A.order[b];   // launch order hint: queue position -> trajectory
                if (A.in_ready && b >= (long)A.in_ready_first) {
                  // streaming host entry: this trajectory's inputs may still be in flight (the copies were queued before
                  // the kernel; the bound below only keeps a failed copy from hanging the device)
                  const volatile int* f = A.in_ready + (b - (long)A.in_ready_first) / A.in_ready_block;
                  const long long t0 = clock64();
                  while (*f == 0 && clock64() - t0 < 4000000000LL) { }
                  __threadfence();
                }
                lane_bind(ln, A, t_shared, s, b);
                ln.t = (double)ln.in.t_obs[0];
                t_bound = (double)ln.in.t_obs[T - 1];
                need_init = true; prev_rejected = false; need_stop = true;
                attempts = 0; kink_cur = 1; h_abs = 0.0;
                rk_n = 0; rk_ss = 0;
                if (SOLVER == HODE_SOLVER_RK4) {
                  emit_row(A, ln.out, ln.b, 0, ln.y, vi_n);
                  ln.ei = 1;
                  if (T < 2) lane_finish(ln, A, vi_n);
                } else {
                  if (clip && T <= 64 && any_series(ln.in))
                    kink_mask = (A.kink_masks && !(A.in_ready && ln.b >= (long)A.in_ready_first)) ? A.kink_masks[ln.b]
                                                                                                   : build_kink_mask(ln.in);
                  // the two evaluations of select_initial_step (at t0 and t0 + h0) read the inputs through the
                  // cached piece: load the first grid interval now (the first real attempt reloads it)
                  if (ln.cached && any_series(ln.in) && T >= 2) lane_cache_inputs(ln, 0);
                }
              } else {
                queue_dry = true;
              }
            }
          }
        }
        HODE_TL(101);
        // ---- does this tile still have work? ------------------------------------------------------
        // (a lane without a trajectory that has not yet seen the end of the queue may still be served)
        if (!tile_or(HELP ? c.bar_all : c.bar_id, HELP ? 2 * TILE : TILE, ln.has || !queue_dry)) break;

        HODE_TL(102);
        // ---- round set-up ---------------------------------------------------------------------------
        const bool init = (SOLVER == HODE_SOLVER_DOPRI5) && ln.has && need_init;
        bool run = ln.has && !init;   // lane performs a real step / attempt this round
        double t = ln.t, h = 0.0, t_new = ln.t, h0 = 0.0;
        float hf = 0.f, h0f = 0.f, d1 = 0.f;
        // y_new = y + h sum b_j k_j of a DP5(4) attempt, from the stage store (evaluated twice with the same operands:
        // as the input of the 7th stage and again when the round is closed, instead of 12 registers held across it)
        auto dp_increment = [&](float* incr, float* ynew) {
  #pragma unroll
          for (int i = 0; i < NS; ++i) {
            incr[i] = hf * fmaf(dp::b6, KS(5, i), fmaf(dp::b5, KS(4, i), fmaf(dp::b4, KS(3, i), fmaf(dp::b3, KS(2, i), dp::b1 * KS(0, i)))));
            ynew[i] = ln.y[i] + (incr[i] - ln.cmp[i]);
          }
        };
        if (SOLVER == HODE_SOLVER_RK4) {
          if (run) {
            const double ta = (double)ln.in.t_obs[rk_n];
            h = ((double)ln.in.t_obs[rk_n + 1] - ta) / nsub;
            t = ta + rk_ss * h;
            t_new = t + h;
            hf = (float)h;
            ln.in.cur = rk_n;
            if (rk_ss == 0 && any_series(ln.in)) lane_cache_inputs(ln, rk_n);
          }
        } else if (run) {
          if (need_stop) {
            t_stop = t_bound;
            // the grid is float32 and t is float64: f > t  <=>  f > rd(t), the largest float <= t, so the searches
            // below compare in float32 (the FP64 pipe is narrow, and these loops run every accepted step)
            const float t_rd = __double2float_rd(t);
            if (clip && any_series(ln.in)) {
              if (T <= 64) {
                unsigned long long m = kink_mask & ~((1ull << kink_cur) - 1ull);
                while (m) {
                  const int i = __ffsll((long long)m) - 1;
                  if (ln.in.t_obs[i] > t_rd) { t_stop = (double)ln.in.t_obs[i]; kink_cur = i; break; }
                  m &= m - 1ull;
                  kink_cur = i + 1;
                }
              } else {
                while (kink_cur < T - 1 && !(ln.in.t_obs[kink_cur] > t_rd && is_kink(ln.in, kink_cur)))
                  ++kink_cur;
                if (kink_cur < T - 1) t_stop = (double)ln.in.t_obs[kink_cur];
              }
            }
            {
              const int j = grid_index_from(ln.in, (float)t, ln.in.cur);   // grid points < float(t)
              ln.in.cur = j > 0 ? j - 1 : 0;
              // Between two kinks the inputs are one linear piece (flat stretches have c_dv == 0), so the
              // cached piece stays valid until a step has ended on a kink: reload only then.
              if (ln.cached && any_series(ln.in) && (ln.c_stale || !clip)) {
                int i0 = (j < T && ln.in.t_obs[j] == (float)t) ? j : j - 1;
                i0 = i0 < 0 ? 0 : (i0 > T - 2 ? T - 2 : i0);
                lane_cache_inputs(ln, i0);
                ln.c_stale = false;
              }
            }
            need_stop = false;
          }
          const double min_step = 10.0 * ulp_above(t);
          if (!prev_rejected && h_abs < min_step) h_abs = min_step;
          if (h_abs < min_step) { ln.status = HODE_ST_STEP_TOO_SMALL; run = false; }
          else if (attempts >= max_steps) { ln.status = HODE_ST_MAX_STEPS; run = false; }
          else {
            ++attempts;
            t_new = t + h_abs;
            if (t_new - t_stop > 0) t_new = t_stop;
            h = t_new - t;
            h_abs = h;
            hf = (float)h;
          }
        }

        HODE_TL(103);
        // ---- the evaluations of this round: ONE call site of the tile MLP ------------------------
  #pragma unroll 1
        for (int slot = 0; slot < NSLOT; ++slot) {
          float ys[NS], d[NS];
          double te;
          if (SOLVER == HODE_SOLVER_RK4) {
            const float a = (slot == 0) ? 0.f : (slot == 3 ? hf : 0.5f * hf);
  #pragma unroll
            for (int i = 0; i < NS; ++i) {
              const float kv = (slot == 0) ? 0.f : KS(slot - 1, i);
              ys[i] = (slot == 0) ? ln.y[i] : fmaf(a, kv, ln.y[i]);
            }
            te = (slot == 0) ? t : (slot == 3 ? t + h : t + 0.5 * h);
          } else {
            if (slot == 0) {
  #pragma unroll
              for (int i = 0; i < NS; ++i) ys[i] = init ? ln.y[i] : fmaf(hf, dp::a21 * KS(0, i), ln.y[i]);
              te = init ? t : t + (double)dp::c2 * h;
            } else if (slot == 1) {
  #pragma unroll
              for (int i = 0; i < NS; ++i)
                ys[i] = init ? fmaf(h0f, KS(0, i), ln.y[i])
                             : fmaf(hf, fmaf(dp::a32, KS(1, i), dp::a31 * KS(0, i)), ln.y[i]);
              te = init ? t + h0 : t + (double)dp::c3 * h;
            } else if (slot == 2) {
  #pragma unroll
              for (int i = 0; i < NS; ++i)
                ys[i] = fmaf(hf, fmaf(dp::a43, KS(2, i), fmaf(dp::a42, KS(1, i), dp::a41 * KS(0, i))), ln.y[i]);
              te = t + (double)dp::c4 * h;
            } else if (slot == 3) {
  #pragma unroll
              for (int i = 0; i < NS; ++i)
                ys[i] = fmaf(hf, fmaf(dp::a54, KS(3, i), fmaf(dp::a53, KS(2, i), fmaf(dp::a52, KS(1, i), dp::a51 * KS(0, i)))), ln.y[i]);
              te = t + (double)dp::c5 * h;
            } else if (slot == 4) {
  #pragma unroll
              for (int i = 0; i < NS; ++i)
                ys[i] = fmaf(hf, fmaf(dp::a65, KS(4, i), fmaf(dp::a64, KS(3, i), fmaf(dp::a63, KS(2, i), fmaf(dp::a62, KS(1, i), dp::a61 * KS(0, i))))), ln.y[i]);
              te = t_new;
            } else {
              float incr[NS];
              dp_increment(incr, ys);
              te = t_new;
            }
          }
          lane_eval<X3>(c, th_sm, A.theta_per_traj ? A.theta + (size_t)ln.b * HODE_N_THETA : nullptr, ln, te, ys, d);
          if (SOLVER == HODE_SOLVER_RK4) {
  #pragma unroll
            for (int i = 0; i < NS; ++i) KS(slot, i) = d[i];
          } else {
  #pragma unroll
            for (int i = 0; i < NS; ++i) KS(slot + 1, i) = d[i];
            if (init && slot == 0) {
              // f(t0, y0) -> k1; outputs at t_eval <= t0; first part of select_initial_step
  #pragma unroll
              for (int i = 0; i < NS; ++i) KS(0, i) = d[i];
              while (ln.ei < T && (double)ln.in.t_obs[ln.ei] <= t) {
                emit_row(A, ln.out, ln.b, ln.ei, ln.y, vi_n);
                ++ln.ei;
              }
              float v0[NS], v1[NS];
  #pragma unroll
              for (int i = 0; i < NS; ++i) {
                const float sc = atol + fabsf(ln.y[i]) * rtol;
                v0[i] = ln.y[i] / sc;
                v1[i] = d[i] / sc;
              }
              const float d0 = rms6v(v0);
              d1 = rms6v(v1);
              h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6 : 0.01 * (double)d0 / (double)d1;
              const double interval = t_bound - t;
              if (h0 > interval) h0 = interval;
              h0f = (float)h0;
            } else if (init && slot == 1) {
              const double interval = t_bound - t;
              if (interval > 0) {
                float v0[NS];
  #pragma unroll
                for (int i = 0; i < NS; ++i) v0[i] = (d[i] - KS(0, i)) / (atol + fabsf(ln.y[i]) * rtol);
                const float d2 = rms6v(v0) / h0f;
                double h1;
                if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmax(1e-6, h0 * 1e-3);
                else h1 = (double)powf(0.01f / fmaxf(d1, d2), 0.2f);
                h_abs = fmin(fmin(100.0 * h0, h1), interval);
              }
              need_init = false;
            }
          }
        }

        HODE_TL(104);
        // ---- close the round ----------------------------------------------------------------------------
        if (SOLVER == HODE_SOLVER_RK4) {
          if (run) {
            if (A.save_n && ln.n_saved < A.max_saved) {
              float kl[3][NS];
  #pragma unroll
              for (int j = 0; j < 3; ++j)
  #pragma unroll
                for (int i = 0; i < NS; ++i) kl[j][i] = KS(j, i);
              lane_save_step(ln, A, t, hf, kl, 2);
            }
  #pragma unroll
            for (int i = 0; i < NS; ++i) {
              const float inc = (hf * (1.0f / 6.0f)) * (KS(0, i) + 2.0f * KS(1, i) + 2.0f * KS(2, i) + KS(3, i));
              const float yk = inc - ln.cmp[i];
              const float tn = ln.y[i] + yk;
              ln.cmp[i] = (tn - ln.y[i]) - yk;
              ln.y[i] = tn;
            }
            ++ln.n_acc;
            if (++rk_ss == nsub) {
              rk_ss = 0;
              ++rk_n;
              emit_row(A, ln.out, ln.b, rk_n, ln.y, vi_n);
              ln.ei = rk_n + 1;
              if (rk_n + 1 >= T) lane_finish(ln, A, vi_n);
            }
          }
          continue;
        }
        if (init) {
          if (!(t < t_bound)) lane_finish(ln, A, vi_n);   // T == 1 or zero-length span
        } else if (ln.has && !run) {
          lane_finish(ln, A, vi_n);                         // step too small / budget exhausted
        } else if (run) {
          float ynew[NS], incr[NS];
          dp_increment(incr, ynew);
          float e2 = 0.f;
          bool finite = true;
  #pragma unroll
          for (int i = 0; i < NS; ++i) {
            const float scale = fmaf(fmaxf(fabsf(ln.y[i]), fabsf(ynew[i])), rtol, atol);
            const float ee = hf * fmaf(dp::e7, KS(6, i), fmaf(dp::e6, KS(5, i), fmaf(dp::e5, KS(4, i), fmaf(dp::e4, KS(3, i), fmaf(dp::e3, KS(2, i), dp::e1 * KS(0, i))))));
            const float q = __fdividef(ee, scale);
            e2 = fmaf(q, q, e2);
            finite = finite && isfinite(ynew[i]);
          }
          const float err = sqrtf(e2 * (1.0f / NS));
          if (err < 1.0f) {
            float factor = (err == 0.f) ? 10.f : fminf(10.f, 0.9f * __powf(err, -0.2f));
            if (prev_rejected) factor = fminf(1.f, factor);
            prev_rejected = false;
            ++ln.n_acc;
            bool ok = true;
            if (A.save_n) {
              if (ln.n_saved < A.max_saved) {
                float kl[6][NS];
  #pragma unroll
                for (int j = 0; j < 6; ++j)
  #pragma unroll
                  for (int i = 0; i < NS; ++i) kl[j][i] = KS(j, i);
                lane_save_step(ln, A, t, hf, kl, 5);
              } else { ln.status = HODE_ST_REC_OVERFLOW; ok = false; }
            }
            if (ok) {
              // f <= t_new  <=>  f <= rd(t_new);  f == t_new  <=>  t_new is a float and f == rd(t_new)
              const float tn_rd = __double2float_rd(t_new);
              const bool tn_is_float = (double)tn_rd == t_new;
              if (ln.ei < T && ln.in.t_obs[ln.ei] <= tn_rd) {
                float Q[NS][4];
  #pragma unroll
                for (int i = 0; i < NS; ++i) dp::dense_q(KS(0, i), KS(2, i), KS(3, i), KS(4, i), KS(5, i), KS(6, i), Q[i]);
                while (ln.ei < T && ln.in.t_obs[ln.ei] <= tn_rd) {
                  const float tef = ln.in.t_obs[ln.ei];
                  const double te = (double)tef;
                  float yo[NS];
                  if (tn_is_float && tef == tn_rd) {
  #pragma unroll
                    for (int i = 0; i < NS; ++i) yo[i] = ynew[i];
                  } else {
                    const float xq = __fdividef((float)(te - t), hf);
  #pragma unroll
                    for (int i = 0; i < NS; ++i) {
                      const float poly = xq * fmaf(xq, fmaf(xq, fmaf(xq, Q[i][3], Q[i][2]), Q[i][1]), Q[i][0]);
                      yo[i] = fmaf(hf, poly, ln.y[i]);
                    }
                  }
                  emit_row(A, ln.out, ln.b, ln.ei, yo, vi_n);
                  ++ln.ei;
                }
              }
  #pragma unroll
              for (int i = 0; i < NS; ++i) {
                const float yk = incr[i] - ln.cmp[i];
                ln.cmp[i] = (ynew[i] - ln.y[i]) - yk;
                ln.y[i] = ynew[i];
                KS(0, i) = KS(6, i);
              }
              ln.t = t_new;
              h_abs *= (double)factor;
              need_stop = true;
              if (t_new - t_stop >= 0) ln.c_stale = true;   // the next step starts on a kink
              if (t_new - t_bound >= 0) lane_finish(ln, A, vi_n);
            } else {
              lane_finish(ln, A, vi_n);
            }
          } else {
            ++ln.n_rej;
            if (!finite || !(err == err)) {
              ln.status = HODE_ST_STEP_TOO_SMALL;
              lane_finish(ln, A, vi_n);
            } else {
              h_abs *= (double)fmaxf(0.2f, 0.9f * __powf(err, -0.2f));
              prev_rejected = true;
            }
          }
        }
      }
    }

  }

#undef KS
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// One warp per trajectory: bit i of the mask <=> some series input is not flat across (i-1, i, i+1).
__global__ void kink_mask_kernel(const RolloutArgs A, unsigned long long* __restrict__ masks) {
  const long b = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= A.B) return;
  unsigned d0 = 0u, d1 = 0u;   // lane i: v[i] != v[i+1], v[i+32] != v[i+33]
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (A.in_mode[ch] != HODE_IN_SERIES) continue;
    const float* v = A.u[ch] + b * A.T;
    const int i0 = lane, i1 = lane + 32;
    if (i0 + 1 < A.T && v[i0] != v[i0 + 1]) d0 = 1u;
    if (i1 + 1 < A.T && v[i1] != v[i1 + 1]) d1 = 1u;
  }
  const unsigned long long diff = (unsigned long long)__ballot_sync(0xffffffffu, d0) |
                                  ((unsigned long long)__ballot_sync(0xffffffffu, d1) << 32);
  unsigned long long m = diff | (diff << 1);
  m &= ~1ull;
  if (A.T >= 1) m &= ~(1ull << (A.T - 1));
  if (lane == 0) masks[b] = m;
}

size_t tc_workspace_bytes(int S, int L, int B) {
  return (size_t)S * tc_image_floats_max(L) * sizeof(float) + 256 + (((size_t)S * sizeof(int) + 255) & ~(size_t)255) +
         (size_t)B * sizeof(unsigned long long);
}

cudaError_t launch_rollout_tc(const RolloutArgs& A_in, int mlp_mode, void* workspace, cudaStream_t stream) {
  RolloutArgs A = A_in;
  const int img_floats = tc_image_floats_mode(A.L, mlp_mode);
  const int max_img_floats = tc_image_floats_max(A.L);   // the workspace layout is mode-independent
  float* img = reinterpret_cast<float*>(workspace);
  int* queue = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) +
                                      (((size_t)A.S * max_img_floats * sizeof(float) + 255) / 256) * 256);
  cudaError_t e = cudaMemsetAsync(queue, 0, (size_t)A.S * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  if (A.solver == HODE_SOLVER_DOPRI5 && A.kink_mode == HODE_KINK_CLIP && A.T <= 64 &&
      (A.in_mode[0] == HODE_IN_SERIES || A.in_mode[1] == HODE_IN_SERIES || A.in_mode[2] == HODE_IN_SERIES)) {
    unsigned long long* masks = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<char*>(queue) + (((size_t)A.S * sizeof(int) + 255) & ~(size_t)255));
    count_launch();
    // (streaming host entry: only the first in_ready_first trajectories are resident yet; the others build their mask
    // when they are bound)
    const long n_masks = A.in_ready ? (long)A.in_ready_first : (long)A.B;
    RolloutArgs Am = A;
    Am.B = (int32_t)n_masks;
    kink_mask_kernel<<<(unsigned)((n_masks + 7) / 8), 256, 0, stream>>>(Am, masks);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    A.kink_masks = masks;
  }
  e = tc_prepare_fwd_images(A.W, img, A.S, A.L, A.P, mlp_mode, stream);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool h16 = mlp_mode == HODE_MLP_F16BF16X2;
  // HODE_MLP_F16BF16X2 has two launch shapes with bit-identical results: three tiles per SM (throughput) and two tiles
  // with helper warps (a shorter round).  A cohort that fits two tiles per SM anyway runs the second one: small
  // batches, config 3's 32 768-trajectory per-GPU shard.  HODE_H16_TILES=2|3 forces a shape (tests, measurements).
  bool h16_2t = h16 && (long)A.B <= (long)sms * tiles_per_cta<MLP_H16_2T>() * TILE;
  if (h16) {
    const char* force = getenv("HODE_H16_TILES");
    if (force && force[0] == '2') h16_2t = true;
    if (force && force[0] == '3') h16_2t = false;
  }
  const bool mix3 = mlp_mode == HODE_MLP_TF32X2BF16 || (h16 && !h16_2t);   // the three-tile launch shape
  const int n_main = (mix3 ? tiles_per_cta<MLP_MIX3>() : tiles_per_cta<MLP_X3>()) * TILE;
  const int n_thr = mix3 ? cta_threads<MLP_MIX3>() : cta_threads<MLP_X3>();
  size_t smem = (size_t)(((img_floats + 3) & ~3) + N_KSTAGE * n_main) * sizeof(float);
  if (!A.t_per_traj && A.T <= HODE_SIMT_MAX_SHARED_T) smem += (size_t)A.T * sizeof(float);
  smem = (smem + 1023) & ~(size_t)1023;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  const long units = (long)A.B;
  int grid = sms;
  const long need = (units + n_main - 1) / n_main;
  if (need < grid) grid = (int)(need > 0 ? need : 1);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    count_launch();
    kern<<<grid, n_thr, smem, stream>>>(A, img, img_floats, queue);
    return cudaSuccess;
  };
  const bool rk4 = A.solver == HODE_SOLVER_RK4;
  if (mlp_mode == HODE_MLP_TF32X3)
    e = rk4 ? launch(rollout_tc_kernel<MLP_X3, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_X3, HODE_SOLVER_DOPRI5>);
  else if (h16_2t)
    e = rk4 ? launch(rollout_tc_kernel<MLP_H16_2T, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_H16_2T, HODE_SOLVER_DOPRI5>);
  else if (h16)
    e = rk4 ? launch(rollout_tc_kernel<MLP_H16, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_H16, HODE_SOLVER_DOPRI5>);
  else if (mix3)
    e = rk4 ? launch(rollout_tc_kernel<MLP_MIX3, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_MIX3, HODE_SOLVER_DOPRI5>);
  else if (mlp_mode == HODE_MLP_TF32BF16)
    e = rk4 ? launch(rollout_tc_kernel<MLP_MIXED, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_MIXED, HODE_SOLVER_DOPRI5>);
  else
    e = rk4 ? launch(rollout_tc_kernel<MLP_TF32, HODE_SOLVER_RK4>) : launch(rollout_tc_kernel<MLP_TF32, HODE_SOLVER_DOPRI5>);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace hode

