// hode_rollout_simt.cu — one trajectory per thread, everything in registers / shared memory.
//
// This is the FP32 CUDA-core rollout: the mechanistic-only path (BASELINE config
// "ablation_no_nn") and the parity-mode hybrid path (HODE_MLP_FP32).  It replaces the
// reference's per-trajectory Python loop around scipy.integrate.solve_ivp
// (reference models/hybrid_ode_nn.py:184-256): RK stages, error norm, step controller,
// dense output to the observation times and the zero-padding-on-failure contract all
// happen inside one kernel; HBM sees only y0, the inputs and the [B,T,6] outputs.
#include <math.h>

#include "hode_common.cuh"
#include "hode_kernels.h"

#include "hode_dop853_coef.cuh"

namespace hode {

// ------------------------------------------------------------------------------------------
// Residual MLP on CUDA cores.  Weights live in shared memory as a transposed image
// (Wt[k][j], j contiguous, see mlp_image_floats) so that a warp's read of 4 consecutive
// output-neuron weights is one broadcast LDS.128.  Activations between layers live in a
// per-thread shared-memory column act[k * stride] (bank-conflict free).
// Reference: models/nn_residual.py:60-78 (layers), :138-146 (feature order).
// ------------------------------------------------------------------------------------------
struct MlpSmem {
  const float* img;  // weight image
  float* actA;       // this thread's column, buffer A (element k at actA[k*stride])
  float* actB;       // buffer B
  int stride;
  int H, L;
};

// Generic width (H <= 128), runtime loop bounds.
struct Vec9 { float v[HODE_NN_IN]; };
struct Vec6 { float v[NS]; };

__device__ __noinline__ Vec6 mlp_eval_generic(const MlpSmem m, const Vec9 xin) {
  const float* x9 = xin.v;
  Vec6 res;
  float* out6 = res.v;
  float* in = m.actA;
  float* outb = m.actB;
#pragma unroll
  for (int k = 0; k < HODE_NN_IN; ++k) in[k * m.stride] = x9[k];
  const float* w = m.img;
  int n_in = HODE_NN_IN;
  for (int l = 0; l <= m.L; ++l) {
    const int n_out = (l == m.L) ? NS : m.H;
    const int ldo = mlp_ldo(n_out);
    const float* bias = w + n_in * ldo;
    for (int j0 = 0; j0 < ldo; j0 += 8) {
      float acc[8];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + j0);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + j0 + 4);
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
        acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
      }
#pragma unroll 4
      for (int k = 0; k < n_in; ++k) {
        const float a = in[k * m.stride];
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * ldo + j0);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * ldo + j0 + 4);
        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
        acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
        acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
      }
      if (l == m.L) {
#pragma unroll
        for (int i = 0; i < NS; ++i) out6[i] = acc[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (j0 + i < n_out) outb[(j0 + i) * m.stride] = fmaxf(acc[i], 0.f);
      }
    }
    float* tmp = in; in = outb; outb = tmp;
    w = bias + ldo;
    n_in = n_out;
  }
  return res;
}

// Width 64: the layer input is held in 64 registers, so the inner loop is 8 FFMA per two
// broadcast LDS.128 with no per-thread activation loads.
__device__ __noinline__ Vec6 mlp_eval_h64(const MlpSmem m, const Vec9 xin) {
  constexpr int H = 64;
  const float* x9 = xin.v;
  Vec6 res;
  float* out6 = res.v;
  float a[H];
  float* col = m.actA;
  const float* w = m.img;
  // layer 0: 9 -> 64
  {
    const float* bias = w + HODE_NN_IN * H;
#pragma unroll 1
    for (int j0 = 0; j0 < H; j0 += 8) {
      float acc[8];
      const float4 b0 = *reinterpret_cast<const float4*>(bias + j0);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + j0 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
      for (int k = 0; k < HODE_NN_IN; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * H + j0);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * H + j0 + 4);
        acc[0] = fmaf(x9[k], w0.x, acc[0]); acc[1] = fmaf(x9[k], w0.y, acc[1]);
        acc[2] = fmaf(x9[k], w0.z, acc[2]); acc[3] = fmaf(x9[k], w0.w, acc[3]);
        acc[4] = fmaf(x9[k], w1.x, acc[4]); acc[5] = fmaf(x9[k], w1.y, acc[5]);
        acc[6] = fmaf(x9[k], w1.z, acc[6]); acc[7] = fmaf(x9[k], w1.w, acc[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) col[(j0 + i) * m.stride] = fmaxf(acc[i], 0.f);
    }
    w = bias + H;
  }
  // hidden layers 1..L-1: 64 -> 64
#pragma unroll 1
  for (int l = 1; l < m.L; ++l) {
#pragma unroll
    for (int k = 0; k < H; ++k) a[k] = col[k * m.stride];
    const float* bias = w + H * H;
#pragma unroll 1
    for (int j0 = 0; j0 < H; j0 += 8) {
      float acc[8];
      const float4 b0 = *reinterpret_cast<const float4*>(bias + j0);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + j0 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * H + j0);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * H + j0 + 4);
        acc[0] = fmaf(a[k], w0.x, acc[0]); acc[1] = fmaf(a[k], w0.y, acc[1]);
        acc[2] = fmaf(a[k], w0.z, acc[2]); acc[3] = fmaf(a[k], w0.w, acc[3]);
        acc[4] = fmaf(a[k], w1.x, acc[4]); acc[5] = fmaf(a[k], w1.y, acc[5]);
        acc[6] = fmaf(a[k], w1.z, acc[6]); acc[7] = fmaf(a[k], w1.w, acc[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) col[(j0 + i) * m.stride] = fmaxf(acc[i], 0.f);
    }
    w = bias + H;
  }
  // output layer: 64 -> 6 (image padded to 8 columns)
  {
#pragma unroll
    for (int k = 0; k < H; ++k) a[k] = col[k * m.stride];
    const float* bias = w + H * 8;
    float acc[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + 4);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(w + k * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w + k * 8 + 4);
      acc[0] = fmaf(a[k], w0.x, acc[0]); acc[1] = fmaf(a[k], w0.y, acc[1]);
      acc[2] = fmaf(a[k], w0.z, acc[2]); acc[3] = fmaf(a[k], w0.w, acc[3]);
      acc[4] = fmaf(a[k], w1.x, acc[4]); acc[5] = fmaf(a[k], w1.y, acc[5]);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) out6[i] = acc[i];
  }
  return res;
}

// ------------------------------------------------------------------------------------------
// f_physio + g_NN for one trajectory (reference models/hybrid_ode_nn.py:108-134).
// MLP_KIND: 0 none, 1 generic, 2 width-64 register path.
// ------------------------------------------------------------------------------------------
template <int MLP_KIND>
__device__ __forceinline__ void rhs_full(const Theta& th, const MlpSmem& m, TrajInputs& in,
                                         double t, const float* y, float* d) {
  const float t32 = (float)t;
  float meal = 0.f, tvns = 0.f, gd = 0.f;
  if (any_series(in)) {
    const int idx = grid_index_from(in, t32, in.cur);
    meal = input_channel(in, HODE_CH_MEAL, t32, idx);
    tvns = input_channel(in, HODE_CH_TVNS, t32, idx);
    gd = input_channel(in, HODE_CH_GD, t32, idx);
  } else {
    meal = input_channel(in, HODE_CH_MEAL, t32, 0);
    tvns = input_channel(in, HODE_CH_TVNS, t32, 0);
    gd = input_channel(in, HODE_CH_GD, t32, 0);
  }
  rhs_mech(th, y, meal, gd, in.mode[HODE_CH_GD] != HODE_IN_ABSENT, d);
  if (MLP_KIND != 0) {
    Vec9 x;
    x.v[0] = t32;
#pragma unroll
    for (int i = 0; i < NS; ++i) x.v[1 + i] = y[i];
    x.v[7] = y[3];
    x.v[8] = tvns;
    const Vec6 r = (MLP_KIND == 2) ? mlp_eval_h64(m, x) : mlp_eval_generic(m, x);
#pragma unroll
    for (int i = 0; i < NS; ++i) d[i] = __fadd_rn(d[i], r.v[i]);
  }
}

__device__ __forceinline__ void store_row(float* traj_row, const float* y) {
  float2* p = reinterpret_cast<float2*>(traj_row);
  p[0] = make_float2(y[0], y[1]);
  p[1] = make_float2(y[2], y[3]);
  p[2] = make_float2(y[4], y[5]);
}

__device__ __forceinline__ float rms6(const float* v) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) s = fmaf(v[i], v[i], s);
  return sqrtf(s * (1.0f / NS));
}

// ------------------------------------------------------------------------------------------
// One (parameter set s, trajectory b) unit, integrated by one thread.
// ------------------------------------------------------------------------------------------
template <int MLP_KIND>
__device__ __forceinline__ void simt_integrate(const RolloutArgs& A, const MlpSmem& mlp, const float* t_shared,
                                               int s, long b, int vi_n) {
  const long unit = (long)s * A.B + b;
  const long n_units = (long)A.S * A.B;
  const Theta th = load_theta(A.theta + (A.theta_per_traj ? (size_t)b : (size_t)s) * HODE_N_THETA);
  TrajInputs in;
  in.T = A.T;
  in.cur = 0;
  in.t_obs = A.t_per_traj ? A.t_obs + b * A.T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    in.mode[ch] = A.in_mode[ch];
    in.u[ch] = in.mode[ch] == HODE_IN_SERIES ? A.u[ch] + b * A.T
             : in.mode[ch] == HODE_IN_CONST ? A.u[ch] + b : nullptr;
  }
  float* out = A.traj ? A.traj + (size_t)unit * A.T * (A.out_nc ? A.out_nc : NS) : nullptr;
  const int T = A.T;

  float y[NS], cmp[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) { y[i] = A.y0[b * NS + i]; cmp[i] = 0.f; }

  int status = HODE_ST_OK, n_acc = 0, n_rej = 0, n_saved = 0;
  int ei = 0;  // next observation index to emit

  if (A.solver == HODE_SOLVER_RK4) {
    // -------- classical RK4, n_substeps equal steps per observation interval -------------
    emit_row(A, out, b, 0, y, vi_n);
    ei = 1;
    const int nsub = A.n_substeps > 0 ? A.n_substeps : 1;
    for (int n = 0; n + 1 < T; ++n) {
      const double ta = (double)in.t_obs[n], tb = (double)in.t_obs[n + 1];
      const double h = (tb - ta) / nsub;
      const float hf = (float)h;
      in.cur = n;  // grid points 0..n-1 are < any stage time of this interval
      for (int ss = 0; ss < nsub; ++ss) {
        const double t = ta + ss * h;
        if (A.save_n && n_saved < A.max_saved) {
          step_rec_store(step_rec(A, unit, n_saved), t, hf, y, nullptr);
          ++n_saved;
        }
        float k1[NS], k2[NS], k3[NS], k4[NS], ys[NS];
        rhs_full<MLP_KIND>(th, mlp, in, t, y, k1);
#pragma unroll
        for (int i = 0; i < NS; ++i) ys[i] = fmaf(0.5f * hf, k1[i], y[i]);
        rhs_full<MLP_KIND>(th, mlp, in, t + 0.5 * h, ys, k2);
#pragma unroll
        for (int i = 0; i < NS; ++i) ys[i] = fmaf(0.5f * hf, k2[i], y[i]);
        rhs_full<MLP_KIND>(th, mlp, in, t + 0.5 * h, ys, k3);
#pragma unroll
        for (int i = 0; i < NS; ++i) ys[i] = fmaf(hf, k3[i], y[i]);
        rhs_full<MLP_KIND>(th, mlp, in, t + h, ys, k4);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          // compensated (Kahan) accumulation of the state across steps
          const float incr = (hf * (1.0f / 6.0f)) * (k1[i] + 2.0f * k2[i] + 2.0f * k3[i] + k4[i]);
          const float yk = incr - cmp[i];
          const float tn = y[i] + yk;
          cmp[i] = (tn - y[i]) - yk;
          y[i] = tn;
        }
        ++n_acc;
      }
      emit_row(A, out, b, n + 1, y, vi_n);
      ei = n + 2;
    }
  } else {
    // -------- Dormand-Prince 5(4) with SciPy's controller (rk.py:111-176) ------------------
    const double t0 = (double)in.t_obs[0], t_bound = (double)in.t_obs[T - 1];
    const float rtol = A.rtol, atol = A.atol;
    const int max_steps = A.max_steps > 0 ? A.max_steps : 100000;
    double t = t0;
    float k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS], ynew[NS], ys[NS];
    rhs_full<MLP_KIND>(th, mlp, in, t, y, k1);
    // outputs at t_eval <= t0 (ivp.py:701-718 emits t_eval[0] == t0 from the first step)
    while (ei < T && (double)in.t_obs[ei] <= t) {
      emit_row(A, out, b, ei, y, vi_n);
      ++ei;
    }
    double h_abs = 0.0;
    bool alive = (t < t_bound);
    if (alive) {
      // select_initial_step, common.py:68-134 (direction +1, max_step inf, order 4)
      float sc[NS], v0[NS], v1[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        sc[i] = atol + fabsf(y[i]) * rtol;
        v0[i] = y[i] / sc[i];
        v1[i] = k1[i] / sc[i];
      }
      const float d0 = rms6(v0), d1 = rms6(v1);
      double h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6 : 0.01 * (double)d0 / (double)d1;
      const double interval = t_bound - t0;
      if (h0 > interval) h0 = interval;
      const float h0f = (float)h0;
#pragma unroll
      for (int i = 0; i < NS; ++i) ys[i] = fmaf(h0f, k1[i], y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t0 + h0, ys, k2);
#pragma unroll
      for (int i = 0; i < NS; ++i) v0[i] = (k2[i] - k1[i]) / sc[i];
      const float d2 = rms6(v0) / h0f;
      double h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = (double)powf(0.01f / fmaxf(d1, d2), 0.2f);
      h_abs = fmin(fmin(100.0 * h0, h1), interval);
    }
    int attempts = 0, kink_cur = 1;
    bool prev_rejected = false;
    double t_stop = t_bound;
    bool need_stop = true;
    while (alive) {
      if (need_stop) {
        // next step end-point: first kink strictly after t (HODE_KINK_CLIP) else t_bound
        t_stop = t_bound;
        if (A.kink_mode == HODE_KINK_CLIP && any_series(in)) {
          while (kink_cur < T - 1 && !((double)in.t_obs[kink_cur] > t && is_kink(in, kink_cur)))
            ++kink_cur;
          if (kink_cur < T - 1) t_stop = (double)in.t_obs[kink_cur];
        }
        in.cur = grid_index_from(in, (float)t, in.cur);
        if (in.cur > 0) --in.cur;  // float(t) may round below t: keep one point of slack
        need_stop = false;
      }
      const double min_step = 10.0 * (nextafter(t, (double)INFINITY) - t);
      if (!prev_rejected && h_abs < min_step) h_abs = min_step;
      if (h_abs < min_step) { status = HODE_ST_STEP_TOO_SMALL; break; }
      if (attempts >= max_steps) { status = HODE_ST_MAX_STEPS; break; }
      ++attempts;
      double t_new = t + h_abs;
      if (t_new - t_stop > 0) t_new = t_stop;
      const double h = t_new - t;
      h_abs = h;
      const float hf = (float)h;
      // ---- stages (rk_step, rk.py:14-71) -----------------------------------------------------
#pragma unroll
      for (int i = 0; i < NS; ++i) ys[i] = fmaf(hf, dp::a21 * k1[i], y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t + (double)dp::c2 * h, ys, k2);
#pragma unroll
      for (int i = 0; i < NS; ++i) ys[i] = fmaf(hf, fmaf(dp::a32, k2[i], dp::a31 * k1[i]), y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t + (double)dp::c3 * h, ys, k3);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        ys[i] = fmaf(hf, fmaf(dp::a43, k3[i], fmaf(dp::a42, k2[i], dp::a41 * k1[i])), y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t + (double)dp::c4 * h, ys, k4);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        ys[i] = fmaf(hf, fmaf(dp::a54, k4[i], fmaf(dp::a53, k3[i], fmaf(dp::a52, k2[i], dp::a51 * k1[i]))), y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t + (double)dp::c5 * h, ys, k5);
#pragma unroll
      for (int i = 0; i < NS; ++i)
        ys[i] = fmaf(hf, fmaf(dp::a65, k5[i], fmaf(dp::a64, k4[i], fmaf(dp::a63, k3[i], fmaf(dp::a62, k2[i], dp::a61 * k1[i])))), y[i]);
      rhs_full<MLP_KIND>(th, mlp, in, t_new, ys, k6);
      float incr[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        incr[i] = hf * fmaf(dp::b6, k6[i], fmaf(dp::b5, k5[i], fmaf(dp::b4, k4[i], fmaf(dp::b3, k3[i], dp::b1 * k1[i]))));
        ynew[i] = y[i] + (incr[i] - cmp[i]);
      }
      rhs_full<MLP_KIND>(th, mlp, in, t_new, ynew, k7);
      // ---- error norm (rk.py:103-107, common.py:63-65) --------------------------------------
      float e2 = 0.f;
      bool finite = true;
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const float scale = fmaf(fmaxf(fabsf(y[i]), fabsf(ynew[i])), rtol, atol);
        const float ee = hf * fmaf(dp::e7, k7[i], fmaf(dp::e6, k6[i], fmaf(dp::e5, k5[i], fmaf(dp::e4, k4[i], fmaf(dp::e3, k3[i], dp::e1 * k1[i])))));
        const float r = ee / scale;
        e2 = fmaf(r, r, e2);
        finite = finite && isfinite(ynew[i]);
      }
      const float err = sqrtf(e2 * (1.0f / NS));
      if (err < 1.0f) {
        // ---- accepted ----------------------------------------------------------------------
        float factor = (err == 0.f) ? 10.f : fminf(10.f, 0.9f * powf(err, -0.2f));
        if (prev_rejected) factor = fminf(1.f, factor);
        prev_rejected = false;
        ++n_acc;
        if (A.save_n) {
          if (n_saved < A.max_saved) {
            step_rec_store(step_rec(A, unit, n_saved), t, hf, y, nullptr);
            ++n_saved;
          } else {
            status = HODE_ST_REC_OVERFLOW;
            break;
          }
        }
        // dense output at observation times in (t, t_new] (rk.py:178-180, ivp.py:701-718)
        if (ei < T && (double)in.t_obs[ei] <= t_new) {
          float Q[NS][4];
#pragma unroll
          for (int i = 0; i < NS; ++i) dp::dense_q(k1[i], k3[i], k4[i], k5[i], k6[i], k7[i], Q[i]);
          while (ei < T && (double)in.t_obs[ei] <= t_new) {
            const double te = (double)in.t_obs[ei];
            float yo[NS];
            if (te == t_new) {
#pragma unroll
              for (int i = 0; i < NS; ++i) yo[i] = ynew[i];
            } else {
              const float x = (float)((te - t) / h);
#pragma unroll
              for (int i = 0; i < NS; ++i) {
                const float poly = x * fmaf(x, fmaf(x, fmaf(x, Q[i][3], Q[i][2]), Q[i][1]), Q[i][0]);
                yo[i] = fmaf(hf, poly, y[i]);
              }
            }
            emit_row(A, out, b, ei, yo, vi_n);
            ++ei;
          }
        }
        // commit the step (Kahan-compensated state)
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const float yk = incr[i] - cmp[i];
          cmp[i] = (ynew[i] - y[i]) - yk;
          y[i] = ynew[i];
          k1[i] = k7[i];
        }
        t = t_new;
        h_abs *= (double)factor;
        need_stop = true;
        if (t - t_bound >= 0) alive = false;
      } else {
        // ---- rejected (also the NaN case: `err < 1` is false) ---------------------------------
        ++n_rej;
        if (!finite || !(err == err)) { status = HODE_ST_STEP_TOO_SMALL; break; }
        h_abs *= (double)fmaxf(0.2f, 0.9f * powf(err, -0.2f));
        prev_rejected = true;
      }
    }
  }

  // failure: zero-pad the remaining observation rows (reference models/hybrid_ode_nn.py:252-254)
  if (out || vi_n) {
    const float z[NS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (; ei < T; ++ei) emit_row(A, out, b, ei, z, vi_n);
  }
  if (A.status) A.status[unit] = status;
  if (A.counters) {
    A.counters[unit] = n_acc;
    A.counters[n_units + unit] = n_rej;
  }
  if (A.save_n) A.save_n[unit] = status == HODE_ST_OK ? n_saved : -1 - n_saved;  // < 0: no gradient
}

// ------------------------------------------------------------------------------------------
// DOP853 — what the reference's solver='dopri5' and 'dop853' really run (models/hybrid_ode_nn.py:174-181 ->
// scipy rk.py:568-720).  One unit per thread, the reference's own arithmetic: float32 RHS, float64 stepping.
// 12 stages + the 8(5,3) error norm (rk.py:683-691), the RK45 controller with exponent -1/8, and the
// 7th-order dense output from three extra stages (rk.py:693-712, 739-765), evaluated only for steps that
// contain observation times (ivp.py:701-718).  The 16 x 6 stage derivatives live in local memory: this is
// the compatibility solver, not the throughput path (HODE_SOLVER_DOPRI5 on the tensor cores is).
// ------------------------------------------------------------------------------------------
template <int MLP_KIND>
__device__ __noinline__ void dop853_integrate(const RolloutArgs& A, const MlpSmem& mlp, const float* t_shared,
                                              int s, long b, int vi_n) {
  const long unit = (long)s * A.B + b;
  const long n_units = (long)A.S * A.B;
  const Theta th = load_theta(A.theta + (A.theta_per_traj ? (size_t)b : (size_t)s) * HODE_N_THETA);
  TrajInputs in;
  in.T = A.T;
  in.cur = 0;
  in.t_obs = A.t_per_traj ? A.t_obs + b * A.T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    in.mode[ch] = A.in_mode[ch];
    in.u[ch] = in.mode[ch] == HODE_IN_SERIES ? A.u[ch] + b * A.T
             : in.mode[ch] == HODE_IN_CONST ? A.u[ch] + b : nullptr;
  }
  float* out = A.traj ? A.traj + (size_t)unit * A.T * (A.out_nc ? A.out_nc : NS) : nullptr;
  const int T = A.T;
  const double rtol = (double)A.rtol, atol = (double)A.atol;
  const int max_steps = A.max_steps > 0 ? A.max_steps : 100000;
  const double t0 = (double)in.t_obs[0], t_bound = (double)in.t_obs[T - 1];

  double y[NS], K[DOP853_N_STAGES_EXT][NS];
  float yf[NS], df[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) { yf[i] = A.y0[b * NS + i]; y[i] = (double)yf[i]; }
  int status = HODE_ST_OK, n_acc = 0, n_rej = 0, ei = 0;
  double t = t0;
  rhs_full<MLP_KIND>(th, mlp, in, t, yf, df);
#pragma unroll
  for (int i = 0; i < NS; ++i) K[0][i] = (double)df[i];
  while (ei < T && (double)in.t_obs[ei] <= t) {
    emit_row(A, out, b, ei, yf, vi_n);
    ++ei;
  }
  bool alive = t < t_bound;
  double h_abs = 0.0;
  if (alive) {
    // select_initial_step, common.py:68-134 (order = error_estimator_order = 7)
    double sc[NS], s0 = 0, s1 = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      sc[i] = atol + fabs(y[i]) * rtol;
      const double a = y[i] / sc[i], c = K[0][i] / sc[i];
      s0 += a * a; s1 += c * c;
    }
    const double d0 = sqrt(s0) / sqrt((double)NS), d1 = sqrt(s1) / sqrt((double)NS);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    const double interval = t_bound - t0;
    if (h0 > interval) h0 = interval;
    float y1[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) y1[i] = (float)(y[i] + h0 * K[0][i]);
    rhs_full<MLP_KIND>(th, mlp, in, t0 + h0, y1, df);
    double s2 = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) { const double a = ((double)df[i] - K[0][i]) / sc[i]; s2 += a * a; }
    const double d2 = sqrt(s2) / sqrt((double)NS) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 8.0);
    h_abs = fmin(fmin(100.0 * h0, h1), interval);
  }
  int attempts = 0, kink_cur = 1;
  bool prev_rejected = false, need_stop = true;
  double t_stop = t_bound;
  while (alive) {
    if (need_stop) {
      t_stop = t_bound;
      if (A.kink_mode == HODE_KINK_CLIP && any_series(in)) {
        while (kink_cur < T - 1 && !((double)in.t_obs[kink_cur] > t && is_kink(in, kink_cur))) ++kink_cur;
        if (kink_cur < T - 1) t_stop = (double)in.t_obs[kink_cur];
      }
      in.cur = grid_index_from(in, (float)t, in.cur);
      if (in.cur > 0) --in.cur;
      need_stop = false;
    }
    const double min_step = 10.0 * (nextafter(t, (double)INFINITY) - t);
    if (!prev_rejected && h_abs < min_step) h_abs = min_step;
    if (h_abs < min_step) { status = HODE_ST_STEP_TOO_SMALL; break; }
    if (attempts >= max_steps) { status = HODE_ST_MAX_STEPS; break; }
    ++attempts;
    double t_new = t + h_abs;
    if (t_new - t_stop > 0) t_new = t_stop;
    const double h = t_new - t;
    h_abs = h;
    // ---- rk_step, rk.py:14-71 ---------------------------------------------------------------
#pragma unroll 1
    for (int st = 1; st < DOP853_N_STAGES; ++st) {
      float ysf[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        double dy = 0;
        for (int j = 0; j < st; ++j) dy += K[j][i] * DOP853_A[st][j];
        ysf[i] = (float)(y[i] + dy * h);
      }
      rhs_full<MLP_KIND>(th, mlp, in, t + DOP853_C[st] * h, ysf, df);
#pragma unroll
      for (int i = 0; i < NS; ++i) K[st][i] = (double)df[i];
    }
    double y_new[NS];
    float ynf[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double acc = 0;
      for (int j = 0; j < DOP853_N_STAGES; ++j) acc += K[j][i] * DOP853_B[j];
      y_new[i] = y[i] + h * acc;
      ynf[i] = (float)y_new[i];
    }
    rhs_full<MLP_KIND>(th, mlp, in, t + h, ynf, df);
#pragma unroll
    for (int i = 0; i < NS; ++i) K[DOP853_N_STAGES][i] = (double)df[i];
    // ---- _estimate_error_norm, rk.py:683-691 -----------------------------------------------------
    double e5n = 0, e3n = 0;
    bool finite = true;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
      double a5 = 0, a3 = 0;
      for (int j = 0; j <= DOP853_N_STAGES; ++j) { a5 += K[j][i] * DOP853_E5[j]; a3 += K[j][i] * DOP853_E3[j]; }
      a5 /= scale; a3 /= scale;
      e5n += a5 * a5; e3n += a3 * a3;
      finite = finite && isfinite(y_new[i]);
    }
    const double err = (e5n == 0 && e3n == 0) ? 0.0 : fabs(h) * e5n / sqrt((e5n + 0.01 * e3n) * NS);
    if (err < 1.0) {
      double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -1.0 / 8.0));
      if (prev_rejected) factor = fmin(1.0, factor);
      prev_rejected = false;
      ++n_acc;
      if (ei < T && (double)in.t_obs[ei] <= t_new) {
        // ---- _dense_output_impl, rk.py:693-712 -------------------------------------------------------
#pragma unroll 1
        for (int st = DOP853_N_STAGES + 1; st < DOP853_N_STAGES_EXT; ++st) {
          float ysf[NS];
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double dy = 0;
            for (int j = 0; j < st; ++j) dy += K[j][i] * DOP853_A[st][j];
            ysf[i] = (float)(y[i] + dy * h);
          }
          rhs_full<MLP_KIND>(th, mlp, in, t + DOP853_C[st] * h, ysf, df);
#pragma unroll
          for (int i = 0; i < NS; ++i) K[st][i] = (double)df[i];
        }
        double F[7][NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          const double dy = y_new[i] - y[i];
          F[0][i] = dy;
          F[1][i] = h * K[0][i] - dy;
          F[2][i] = 2 * dy - h * (K[DOP853_N_STAGES][i] + K[0][i]);
          for (int q = 0; q < 4; ++q) {
            double acc = 0;
            for (int j = 0; j < DOP853_N_STAGES_EXT; ++j) acc += DOP853_D[q][j] * K[j][i];
            F[3 + q][i] = h * acc;
          }
        }
        while (ei < T && (double)in.t_obs[ei] <= t_new) {
          // Dop853DenseOutput._call_impl, rk.py:746-765
          const double x = ((double)in.t_obs[ei] - t) / h;
          float yo[NS];
#pragma unroll
          for (int i = 0; i < NS; ++i) {
            double v = 0;
#pragma unroll
            for (int q = 0; q < 7; ++q) {
              v += F[6 - q][i];
              v *= (q % 2 == 0) ? x : 1 - x;
            }
            yo[i] = (float)(v + y[i]);
          }
          emit_row(A, out, b, ei, yo, vi_n);
          ++ei;
        }
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) { y[i] = y_new[i]; K[0][i] = K[DOP853_N_STAGES][i]; }
      t = t_new;
      h_abs *= factor;
      need_stop = true;
      if (t - t_bound >= 0) alive = false;
    } else {
      ++n_rej;
      if (!finite || !(err == err)) { status = HODE_ST_STEP_TOO_SMALL; break; }
      h_abs *= fmax(0.2, 0.9 * pow(err, -1.0 / 8.0));
      prev_rejected = true;
    }
  }
  if (out || vi_n) {
    const float z[NS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (; ei < T; ++ei) emit_row(A, out, b, ei, z, vi_n);
  }
  if (A.status) A.status[unit] = status;
  if (A.counters) {
    A.counters[unit] = n_acc;
    A.counters[n_units + unit] = n_rej;
  }
}

// ------------------------------------------------------------------------------------------
// The rollout kernel.  Normal mode: grid = (ceil(B / blockDim.x), S), thread = one (sample,
// trajectory) unit.  Fused posterior-predictive mode (A.vi_mean != nullptr): grid =
// (ceil(B / blockDim.x), 1); the CTA walks through all S parameter sets for its trajectories,
// restaging the weight image for each, and reduces mean / std on the fly.
// ------------------------------------------------------------------------------------------
// SOLVER853: the DOP853 instantiation (its own kernel, so that its local-memory stage store does not touch the
// register allocation of the RK4 / DP5(4) kernels)
template <int MLP_KIND, bool SOLVER853 = false>
__global__ void __launch_bounds__(128, (MLP_KIND == 0 && !SOLVER853) ? 4 : 1) rollout_simt_kernel(const RolloutArgs A) {   // mechanistic: 4 CTAs per SM (<= 128 registers)
  extern __shared__ __align__(16) float smem[];
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vi = A.vi_mean != nullptr;
  const int n_iter = vi ? A.S : 1;

  // ---- shared memory carve-up: [MLP image][activation columns][shared time grid] ----------
  float* sm = smem;
  MlpSmem mlp;
  mlp.img = nullptr; mlp.actA = nullptr; mlp.actB = nullptr;
  mlp.stride = blockDim.x; mlp.H = A.H; mlp.L = A.L;
  if (MLP_KIND != 0) {
    const int img_floats = (mlp_image_floats(A.H, A.L) + 3) & ~3;
    mlp.img = sm;
    sm += img_floats;
    const int act_rows = A.H > 16 ? A.H : 16;
    mlp.actA = sm + threadIdx.x;
    sm += act_rows * blockDim.x;
    if (MLP_KIND == 1) {
      mlp.actB = sm + threadIdx.x;
      sm += act_rows * blockDim.x;
    }
  }
  const float* t_shared = nullptr;
  if (!A.t_per_traj && A.T <= HODE_SIMT_MAX_SHARED_T) {
    for (int i = threadIdx.x; i < A.T; i += blockDim.x) sm[i] = A.t_obs[i];
    t_shared = sm;
  }
  for (int si = 0; si < n_iter; ++si) {
    const int s = vi ? (int)((blockIdx.x + (unsigned)si) % (unsigned)A.S) : (int)blockIdx.y;
    if (si > 0) __syncthreads();  // every thread is done with the previous parameter set's image
    if (MLP_KIND != 0) stage_mlp_image(smem, A.W + (size_t)s * A.P, A.H, A.L);
    __syncthreads();
    if (b < A.B) {
      if (SOLVER853) dop853_integrate<MLP_KIND>(A, mlp, t_shared, s, b, vi ? si + 1 : 0);
      else simt_integrate<MLP_KIND>(A, mlp, t_shared, s, b, vi ? si + 1 : 0);
    }
  }
}


// ------------------------------------------------------------------------------------------
// Batched single evaluation of f_physio + g_NN (reference models/hybrid_ode_nn.py:108-134,
// as used by the physics-residual loss at :327).  Inputs are per-trajectory constants.
// ------------------------------------------------------------------------------------------
template <int MLP_KIND>
__global__ void __launch_bounds__(128) rhs_kernel(const RolloutArgs A, const float* __restrict__ tt,
                                                   const float* __restrict__ state,
                                                   float* __restrict__ outp) {
  extern __shared__ __align__(16) float smem[];
  const int s = blockIdx.y;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float* sm = smem;
  MlpSmem mlp;
  mlp.img = nullptr; mlp.actA = nullptr; mlp.actB = nullptr;
  mlp.stride = blockDim.x; mlp.H = A.H; mlp.L = A.L;
  if (MLP_KIND != 0) {
    const int img_floats = (mlp_image_floats(A.H, A.L) + 3) & ~3;
    stage_mlp_image(sm, A.W + (size_t)s * A.P, A.H, A.L);
    mlp.img = sm;
    sm += img_floats;
    const int act_rows = A.H > 16 ? A.H : 16;
    mlp.actA = sm + threadIdx.x;
    sm += act_rows * blockDim.x;
    if (MLP_KIND == 1) mlp.actB = sm + threadIdx.x;
  }
  __syncthreads();
  if (b >= A.B) return;
  const Theta th = load_theta(A.theta + (size_t)s * HODE_N_THETA);
  TrajInputs in;
  in.T = 1; in.cur = 0; in.t_obs = tt + b;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    in.mode[ch] = A.in_mode[ch] == HODE_IN_ABSENT ? HODE_IN_ABSENT : HODE_IN_CONST;
    in.u[ch] = in.mode[ch] == HODE_IN_CONST ? A.u[ch] + b : nullptr;
  }
  float y[NS], d[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) y[i] = state[b * NS + i];
  rhs_full<MLP_KIND>(th, mlp, in, (double)tt[b], y, d);
  if (MLP_KIND != 0 && A.rhs_part == HODE_RHS_NN_ONLY) {
    // g_NN alone: subtracting f_physio again would not be exact, so evaluate the net directly
    Vec9 x;
    x.v[0] = tt[b];
#pragma unroll
    for (int i = 0; i < NS; ++i) x.v[1 + i] = y[i];
    x.v[7] = y[3];
    x.v[8] = input_channel(in, HODE_CH_TVNS, tt[b], 0);
    const Vec6 r = (MLP_KIND == 2) ? mlp_eval_h64(mlp, x) : mlp_eval_generic(mlp, x);
#pragma unroll
    for (int i = 0; i < NS; ++i) d[i] = r.v[i];
  }
  store_row(outp + ((size_t)s * A.B + b) * NS, d);
}

// ------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------
size_t simt_smem_bytes(const RolloutArgs& A, int mlp_kind, int block) {
  size_t floats = 0;
  if (mlp_kind != 0) {
    floats += (mlp_image_floats(A.H, A.L) + 3) & ~3;
    const int act_rows = A.H > 16 ? A.H : 16;
    floats += (size_t)act_rows * block * (mlp_kind == 1 ? 2 : 1);
  }
  if (!A.t_per_traj && A.T <= HODE_SIMT_MAX_SHARED_T) floats += A.T;
  return floats * sizeof(float);
}

// smallest dynamic shared memory any FP32 kernel launch of this network shape needs (32-thread CTAs, per-row time
// grid): the API rejects shapes whose weight image + activation columns cannot fit (HODE_E_UNSUPPORTED)
size_t simt_min_smem_bytes(int H, int L) {
  RolloutArgs A{};
  A.H = H; A.L = L; A.t_per_traj = 1;
  const int kind = (H == 64 && L >= 1) ? 2 : 1;
  return simt_smem_bytes(A, kind, 32);
}

cudaError_t launch_rollout_simt(const RolloutArgs& A, int mlp_mode, cudaStream_t stream) {
  int kind = 0;
  if (mlp_mode != HODE_MLP_NONE) kind = (A.H == 64 && A.L >= 1) ? 2 : 1;
  // block size: 128 threads unless the activation columns of a wide generic net do not fit
  int block = 128;
  while (block > 32 && simt_smem_bytes(A, kind, block) > 220 * 1024) block >>= 1;
  const size_t smem = simt_smem_bytes(A, kind, block);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((A.B + block - 1) / block), A.vi_mean ? 1u : (unsigned)A.S);
  cudaError_t e;
  if (A.solver == HODE_SOLVER_DOP853) {
    auto launch853 = [&](auto kern) -> cudaError_t {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
      count_launch();
      kern<<<grid, block, smem, stream>>>(A);
      return cudaGetLastError();
    };
    return kind == 0 ? launch853(rollout_simt_kernel<0, true>)
         : kind == 1 ? launch853(rollout_simt_kernel<1, true>) : launch853(rollout_simt_kernel<2, true>);
  }
  switch (kind) {
    case 0:
      count_launch();
      rollout_simt_kernel<0><<<grid, block, smem, stream>>>(A);
      break;
    case 1:
      e = cudaFuncSetAttribute(rollout_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      count_launch();
      rollout_simt_kernel<1><<<grid, block, smem, stream>>>(A);
      break;
    default:
      e = cudaFuncSetAttribute(rollout_simt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      count_launch();
      rollout_simt_kernel<2><<<grid, block, smem, stream>>>(A);
      break;
  }
  return cudaGetLastError();
}

cudaError_t launch_rhs(const RolloutArgs& A, int mlp_mode, const float* t, const float* state,
                       float* out, cudaStream_t stream) {
  int kind = 0;
  if (mlp_mode != HODE_MLP_NONE) kind = (A.H == 64 && A.L >= 1) ? 2 : 1;
  RolloutArgs R = A;
  R.t_per_traj = 1;  // no shared time grid in this kernel
  int block = 128;
  while (block > 32 && simt_smem_bytes(R, kind, block) > 220 * 1024) block >>= 1;
  const size_t smem = simt_smem_bytes(R, kind, block);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  dim3 grid((unsigned)((A.B + block - 1) / block), (unsigned)A.S);
  cudaError_t e;
  switch (kind) {
    case 0:
      count_launch();
      rhs_kernel<0><<<grid, block, smem, stream>>>(R, t, state, out);
      break;
    case 1:
      e = cudaFuncSetAttribute(rhs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      count_launch();
      rhs_kernel<1><<<grid, block, smem, stream>>>(R, t, state, out);
      break;
    default:
      e = cudaFuncSetAttribute(rhs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      count_launch();
      rhs_kernel<2><<<grid, block, smem, stream>>>(R, t, state, out);
      break;
  }
  return cudaGetLastError();
}

}  // namespace hode
