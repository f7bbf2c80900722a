// hode_common.cuh — device-side building blocks shared by every libhode kernel (sm_100a).
//
// Replaces, per trajectory and per RK stage, what the reference does in Python:
//   mechanistic RHS ........ reference models/ode_core.py:122-161
//   input interpolation .... reference models/hybrid_ode_nn.py:217-231
//   Dormand-Prince tableau . scipy/integrate/_ivp/rk.py:538-565 (third-party; published tableau)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hode.h"

namespace hode {

constexpr int NS = HODE_N_STATE;

// ---- launch-constant description of one rollout (passed by value to kernels) -----------
struct RolloutArgs {
  const float* y0;      // [B,6]
  const float* t_obs;   // [T] or [B,T]
  const float* u[3];    // per in_mode
  const float* theta;   // [S,17]
  const float* W;       // [S,P] or nullptr
  float* traj;          // [S,B,T,6] (or nullptr in fused-statistics mode)
  int32_t* status;      // [S,B] or nullptr
  int32_t* counters;    // [2,S,B] or nullptr
  // saved accepted steps (discrete adjoint): one record of rec_floats floats per (unit, step), unit-major
  // [S*B][max_saved][rec_floats] — a trajectory's steps are contiguous (the adjoint walks them backwards).
  //   rec_floats = 16: { t (f64), h, pad, y[6], k1[6] }                       (FP32 adjoint: recomputes the stages)
  //   rec_floats = 48: the same + the stage derivatives k2..k6 (DP5(4)) / k2, k3 (RK4) at floats 16.. :
  //                    the tensor-core adjoint then evaluates every stage INPUT directly from the record, so the
  //                    stages of a step can be recomputed one at a time, in reverse, next to their own pull-back
  float* save_rec;
  int32_t rec_floats;
  int32_t save_k1;      // records carry the stage derivatives (tensor-core rollout)
  int32_t* save_n;      // [S*B] number of saved steps
  int32_t max_saved;
  int32_t B, T, S;
  int32_t t_per_traj;
  int32_t in_mode[3];
  int32_t H, L;         // MLP hidden width, hidden layer count
  int32_t P;            // floats per packed MLP parameter set
  int32_t solver, n_substeps, max_steps, kink_mode, rhs_part;
  float rtol, atol;
  // fused posterior-predictive mode (hode_vi_predictive): running mean and sum of squared
  // deviations over the S parameter sets, [B,T,6] each; traj is nullptr in this mode
  float* vi_mean;
  float* vi_m2;
  // hode_rollout_fwd_ex options:
  //   theta_per_traj: theta is [B,17], one mechanistic parameter set PER TRAJECTORY (S == 1; the network stays shared) —
  //                   parameter sweeps such as the Sobol analysis of the reference's plots/plot_all.py:124-224
  //   order:          optional [B] launch order of the trajectories (tensor-core rollout: longest first, from the
  //                   previous pass's attempt counters, removes the tail of small cohorts)
  int32_t theta_per_traj;
  const int32_t* order;
  // host entry with an output-state mask: traj rows hold only the out_nc selected columns (0 = all six)
  uint32_t out_mask;
  int32_t out_nc;
  // optional: bit i of kink_masks[b] set <=> grid point i is a kink of some series input of
  // trajectory b (T <= 64), precomputed by kink_mask_kernel so that lane refill loads one word
  const unsigned long long* kink_masks;
  // optional completion tracking for the streaming host entry (S == 1): done_count[b / done_block]
  // counts finished trajectories of a block (device memory); the lane that completes a block raises
  // done_flag[block] (host-mapped pinned memory) so that the host can start copying that block's
  // outputs back while the kernel is still integrating the rest
  int* done_count;
  volatile int* done_flag;
  int done_block;
  // optional input gating for the streaming host entry (S == 1): the inputs of trajectories b >= in_ready_first are
  // still on their way from the host when the kernel starts; block k = (b - in_ready_first) / in_ready_block may be
  // bound to a lane once in_ready[k] != 0 (device memory, written by the copy stream right after that block's inputs)
  const int* in_ready;
  int in_ready_first;
  int in_ready_block;
};

// ---- step records ---------------------------------------------------------------------------------------
constexpr int HODE_REC_FLOATS = 16;        // base record (and the FP32 path's whole record)
constexpr int HODE_REC_FLOATS_K = 48;      // with the stage derivatives k2..k6 (tensor-core path)
__device__ __forceinline__ float* step_rec(const RolloutArgs& A, long unit, int step) {
  return A.save_rec + ((size_t)unit * A.max_saved + step) * A.rec_floats;
}
__device__ __forceinline__ void step_rec_store(float* r, double t, float h, const float* y, const float* k1) {
  float4* r4 = reinterpret_cast<float4*>(r);
  const long long tb = __double_as_longlong(t);
  r4[0] = make_float4(__int_as_float((int)(tb & 0xFFFFFFFFll)), __int_as_float((int)(tb >> 32)), h, 0.f);
  r4[1] = make_float4(y[0], y[1], y[2], y[3]);
  if (k1) {
    r4[2] = make_float4(y[4], y[5], k1[0], k1[1]);
    r4[3] = make_float4(k1[2], k1[3], k1[4], k1[5]);
  } else {
    r4[2] = make_float4(y[4], y[5], 0.f, 0.f);
  }
}
// stage derivatives k2..k(1+n_extra) -> floats 16.. of a 48-float record (k_j component c at float 10 + 6 (j-1) + c)
__device__ __forceinline__ void step_rec_store_stages(float* r, const float (*k)[NS], int n_extra) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int c = 0; c < NS; ++c) v[6 * j + c] = (j < n_extra) ? k[1 + j][c] : 0.f;
  v[30] = 0.f; v[31] = 0.f;
  float4* r4 = reinterpret_cast<float4*>(r) + 4;
#pragma unroll
  for (int q = 0; q < 8; ++q) r4[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ double step_rec_t(const float* r) { return *reinterpret_cast<const double*>(r); }
__device__ __forceinline__ void step_rec_load(const float* r, double& t, float& h, float* y, float* k1) {
  const float4* r4 = reinterpret_cast<const float4*>(r);
  const float4 a = r4[0], b = r4[1], c = r4[2];
  t = __longlong_as_double(((long long)__float_as_int(a.y) << 32) | (long long)(unsigned)__float_as_int(a.x));
  h = a.z;
  y[0] = b.x; y[1] = b.y; y[2] = b.z; y[3] = b.w; y[4] = c.x; y[5] = c.y;
  if (k1) {
    const float4 d = r4[3];
    k1[0] = c.z; k1[1] = c.w; k1[2] = d.x; k1[3] = d.y; k1[4] = d.z; k1[5] = d.w;
  }
}

// ---- Dormand-Prince 5(4) coefficients (float) --------------------------------------------
namespace dp {
constexpr float c2 = 1.f / 5, c3 = 3.f / 10, c4 = 4.f / 5, c5 = 8.f / 9;
constexpr float a21 = 1.f / 5;
constexpr float a31 = 3.f / 40, a32 = 9.f / 40;
constexpr float a41 = 44.f / 45, a42 = -56.f / 15, a43 = 32.f / 9;
constexpr float a51 = 19372.f / 6561, a52 = -25360.f / 2187, a53 = 64448.f / 6561,
                a54 = -212.f / 729;
constexpr float a61 = 9017.f / 3168, a62 = -355.f / 33, a63 = 46732.f / 5247, a64 = 49.f / 176,
                a65 = -5103.f / 18656;
constexpr float b1 = 35.f / 384, b3 = 500.f / 1113, b4 = 125.f / 192, b5 = -2187.f / 6784,
                b6 = 11.f / 84;
constexpr float e1 = -71.f / 57600, e3 = 71.f / 16695, e4 = -71.f / 1920, e5 = 17253.f / 339200,
                e6 = -22.f / 525, e7 = 1.f / 40;
// dense-output matrix P (7x4), rk.py:554-565; row index 1 is zero.  pJK = P[J-1][K-1].
constexpr float p11 = 1.f;
constexpr float p12 = (float)(-8048581381.0 / 2820520608), p13 = (float)(8663915743.0 / 2820520608),
                p14 = (float)(-12715105075.0 / 11282082432);
constexpr float p32 = (float)(131558114200.0 / 32700410799), p33 = (float)(-68118460800.0 / 10900136933),
                p34 = (float)(87487479700.0 / 32700410799);
constexpr float p42 = (float)(-1754552775.0 / 470086768), p43 = (float)(14199869525.0 / 1410260304),
                p44 = (float)(-10690763975.0 / 1880347072);
constexpr float p52 = (float)(127303824393.0 / 49829197408), p53 = (float)(-318862633887.0 / 49829197408),
                p54 = (float)(701980252875.0 / 199316789632);
constexpr float p62 = (float)(-282668133.0 / 205662961), p63 = (float)(2019193451.0 / 616988883),
                p64 = (float)(-1453857185.0 / 822651844);
constexpr float p72 = (float)(40617522.0 / 29380423), p73 = (float)(-110615467.0 / 29380423),
                p74 = (float)(69997945.0 / 29380423);

// Q = K^T P for one state component (rk.py:178-180); k1..k7 are that component's stages.
__device__ __forceinline__ void dense_q(float k1, float k3, float k4, float k5, float k6, float k7,
                                        float* q) {
  q[0] = p11 * k1;
  q[1] = fmaf(p72, k7, fmaf(p62, k6, fmaf(p52, k5, fmaf(p42, k4, fmaf(p32, k3, p12 * k1)))));
  q[2] = fmaf(p73, k7, fmaf(p63, k6, fmaf(p53, k5, fmaf(p43, k4, fmaf(p33, k3, p13 * k1)))));
  q[3] = fmaf(p74, k7, fmaf(p64, k6, fmaf(p54, k5, fmaf(p44, k4, fmaf(p34, k3, p14 * k1)))));
}
}  // namespace dp

// ---- per-thread view of one trajectory's time grid and inputs ------------------------------
struct TrajInputs {
  const float* t_obs;  // this trajectory's grid (global, or the shared copy)
  const float* u[3];   // series: this trajectory's [T] row; const: pointer to its value
  int mode[3];
  int T;
  int cur;             // monotone cursor: count of grid points known to be < the step start
};

// np.searchsorted(t_eval, t, side='left') restricted to a forward walk from the cursor.
__device__ __forceinline__ int grid_index_from(const TrajInputs& in, float t32, int start) {
  int i = start;
  while (i < in.T && in.t_obs[i] < t32) ++i;
  return i;
}

// Reference models/hybrid_ode_nn.py:217-231 for one channel, float32 arithmetic.
__device__ __forceinline__ float input_channel(const TrajInputs& in, int ch, float t32, int idx) {
  if (in.mode[ch] == HODE_IN_ABSENT) return 0.f;
  if (in.mode[ch] == HODE_IN_CONST) return in.u[ch][0];
  const float* v = in.u[ch];
  if (idx == 0) return v[0];
  if (idx >= in.T) return v[in.T - 1];
  const float t1 = in.t_obs[idx - 1], t2 = in.t_obs[idx];
  const float alpha = __fdiv_rn(t32 - t1, t2 - t1);
  const float v1 = v[idx - 1];
  return __fadd_rn(v1, __fmul_rn(alpha, v[idx] - v1));
}

__device__ __forceinline__ bool any_series(const TrajInputs& in) {
  return in.mode[0] == HODE_IN_SERIES || in.mode[1] == HODE_IN_SERIES ||
         in.mode[2] == HODE_IN_SERIES;
}

// Grid point i is a kink when some series input is not flat across (i-1, i, i+1).
__device__ __forceinline__ bool is_kink(const TrajInputs& in, int i) {
  if (i <= 0 || i >= in.T - 1) return false;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (in.mode[ch] != HODE_IN_SERIES) continue;
    const float a = in.u[ch][i - 1], b = in.u[ch][i], c = in.u[ch][i + 1];
    if (a != b || b != c) return true;
  }
  return false;
}

// ---- mechanistic parameters held in registers -------------------------------------------------
struct Theta {
  float a_GI, k_I, rho, G_b, I_b, E_max, EC_50, Glu_b, V_max, K_m, k_L, k_GE0, IGD_50, g, p_7,
      p_8, p_9;
  float igd_pow;  // IGD_50^g (theta-only; hoisted out of the stage loop)
  float kge0;     // k_GE when GD == 0 (no GD channel): k_GE0 * (1 - 0^g / (IGD_50^g + 0^g))
};

__device__ __forceinline__ Theta load_theta(const float* __restrict__ th) {
  Theta p;
  p.a_GI = th[0]; p.k_I = th[1]; p.rho = th[2]; p.G_b = th[3]; p.I_b = th[4];
  p.E_max = th[5]; p.EC_50 = th[6]; p.Glu_b = th[7]; p.V_max = th[8]; p.K_m = th[9];
  p.k_L = th[10]; p.k_GE0 = th[11]; p.IGD_50 = th[12]; p.g = th[13];
  p.p_7 = th[14]; p.p_8 = th[15]; p.p_9 = th[16];
  p.igd_pow = powf(p.IGD_50, p.g);
  const float z = powf(0.0f, p.g);
  p.kge0 = __fmul_rn(p.k_GE0, 1.0f - __fdiv_rn(z, p.igd_pow + z));
  return p;
}

// f_physio: reference models/ode_core.py:122-161, same operation order, float32, IEEE division.
// (FMA contraction is disabled inside so the result matches an unfused CPU evaluation bit for
// bit wherever powf agrees.)
__device__ __forceinline__ void rhs_mech(const Theta& p, const float* y, float meal, float GD,
                                         bool gd_present, float* d) {
  const float G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const float Pi = __fadd_rn(1.0f, __fmul_rn(p.rho, GLP1));
  d[1] = __fsub_rn(__fmul_rn(__fmul_rn(Pi, p.a_GI), G - p.G_b), __fmul_rn(p.k_I, I - p.I_b));
  const float glp1_effect = __fmul_rn(p.E_max, __fdiv_rn(GLP1, p.EC_50 + GLP1));
  d[2] = __fmul_rn(-glp1_effect, Glu - p.Glu_b);
  d[3] = __fsub_rn(__fmul_rn(p.V_max, __fdiv_rn(G, p.K_m + G)), __fmul_rn(p.k_L, GLP1));
  float k_GE = p.kge0;
  if (gd_present) {
    const float gdg = powf(GD, p.g);
    k_GE = __fmul_rn(p.k_GE0, 1.0f - __fdiv_rn(gdg, p.igd_pow + gdg));
  }
  d[5] = __fadd_rn(__fsub_rn(__fmul_rn(-p.p_7, FFA), __fmul_rn(__fmul_rn(p.p_8, I), FFA)),
                   __fmul_rn(__fmul_rn(p.p_9, G), FFA));
  const float insulin_effect = __fmul_rn(0.01f, I - p.I_b);
  const float glucagon_effect = __fmul_rn(0.005f, Glu - p.Glu_b);
  d[0] = __fsub_rn(__fadd_rn(__fsub_rn(meal, insulin_effect), glucagon_effect),
                   __fmul_rn(k_GE, G));
  d[4] = 0.0f;
}

// float offset of layer l's transposed weight block inside the shared-memory MLP image.
// Image layout per layer: Wt[n_in][ldo] (ldo = n_out rounded up to 8) then bias[ldo].
__host__ __device__ inline int mlp_ldo(int n_out) { return (n_out + 7) & ~7; }
__host__ __device__ inline int mlp_image_floats(int H, int L) {
  int n = HODE_NN_IN * mlp_ldo(H) + mlp_ldo(H);
  for (int l = 1; l < L; ++l) n += H * mlp_ldo(H) + mlp_ldo(H);
  n += H * mlp_ldo(NS) + mlp_ldo(NS);
  return n;
}

__device__ __forceinline__ void store_row6c(float* p, const float* y) {
  float2* q = reinterpret_cast<float2*>(p);
  q[0] = make_float2(y[0], y[1]);
  q[1] = make_float2(y[2], y[3]);
  q[2] = make_float2(y[4], y[5]);
}

// Emit one observation row.  vi_n == 0: store into traj.  vi_n > 0 (fused posterior-predictive
// mode, reference inference/vi.py:306-310): this is the vi_n-th sample of trajectory b; update the
// running mean / sum of squared deviations (Welford) in A.vi_mean / A.vi_m2, and turn the latter
// into the unbiased std on the last sample.  Each (b, ei) element is owned by one thread at a
// time and the samples of a trajectory are visited in a fixed order: deterministic.
__device__ __forceinline__ void emit_row(const RolloutArgs& A, float* out, long b, int ei, const float* y,
                                         int vi_n) {
  if (vi_n == 0) {
    if (out) {
      if (A.out_nc == 0) {
        store_row6c(out + (size_t)ei * NS, y);
      } else {
        float* o = out + (size_t)ei * A.out_nc;
        int j = 0;
#pragma unroll
        for (int i = 0; i < NS; ++i)
          if ((A.out_mask >> i) & 1u) o[j++] = y[i];
      }
    }
    return;
  }
  float* pm = A.vi_mean + ((size_t)b * A.T + ei) * NS;
  float* ps = A.vi_m2 + ((size_t)b * A.T + ei) * NS;
  float m[NS], q[NS];
#ifdef HODE_DBG_VI_NO_RMW   // measurement build only (tools/build_variants.py): the update without its loads — wrong statistics
  if (true) {
#else
  if (vi_n == 1) {
#endif
#pragma unroll
    for (int i = 0; i < NS; ++i) { m[i] = y[i]; q[i] = 0.f; }
  } else {
    const float inv = 1.0f / (float)vi_n;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const float d = y[i] - pm[i];
      m[i] = fmaf(d, inv, pm[i]);
      q[i] = fmaf(d, y[i] - m[i], ps[i]);
    }
  }
  if (vi_n == A.S) {
    const float invs = 1.0f / (float)(A.S - 1);   // S == 1: 0 * inf = NaN, as torch.std of one sample
#pragma unroll
    for (int i = 0; i < NS; ++i) q[i] = sqrtf(q[i] * invs);
  }
  store_row6c(pm, m);
  store_row6c(ps, q);
}

// Stage the packed reference-layout parameters (weight [out,in] row-major, bias) of one
// parameter set into the transposed shared-memory image.
__device__ inline void stage_mlp_image(float* img, const float* __restrict__ Wg, int H, int L) {
  int n_in = HODE_NN_IN;
  const float* src = Wg;
  float* dst = img;
  for (int l = 0; l <= L; ++l) {
    const int n_out = (l == L) ? NS : H;
    const int ldo = mlp_ldo(n_out);
    for (int i = threadIdx.x; i < n_in * ldo; i += blockDim.x) {
      const int k = i / ldo, j = i - k * ldo;
      dst[i] = (j < n_out) ? src[j * n_in + k] : 0.f;
    }
    for (int j = threadIdx.x; j < ldo; j += blockDim.x)
      dst[n_in * ldo + j] = (j < n_out) ? src[n_out * n_in + j] : 0.f;
    src += n_out * n_in + n_out;
    dst += n_in * ldo + ldo;
    n_in = n_out;
  }
}

}  // namespace hode
