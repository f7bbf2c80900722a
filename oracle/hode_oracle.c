/*
 * hode_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's trajectory-rollout path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (libhode.so) never does.
 *
 * What is restated, with the reference lines each piece follows
 * (paths into OliverDOU776/Hybrid-ODE-for-GLP-1-and-Glucose):
 *   - mechanistic RHS ............ models/ode_core.py:122-161
 *   - residual MLP ............... models/nn_residual.py:60-78,136-146
 *   - f_physio + g_NN ............ models/hybrid_ode_nn.py:108-134
 *   - input interpolation ........ models/hybrid_ode_nn.py:210-231
 *   - per-trajectory solve loop .. models/hybrid_ode_nn.py:184-256 (incl. zero padding)
 * The integrator arithmetic lives in a third-party dependency that is NOT in the
 * reference tree: SciPy (requirements.txt:3 "scipy>=1.10.0", 1.18.1 installed where the
 * golden vectors were made).  Its published algorithm is restated here:
 *   - RK45 = Dormand-Prince 5(4) . scipy/integrate/_ivp/rk.py:14-71,111-176,538-565
 *   - initial step ............... scipy/integrate/_ivp/common.py:68-134
 *   - t_eval dense-output loop ... scipy/integrate/_ivp/ivp.py:701-728, rk.py:178-180
 *   - DOP853 (what the reference's solver='dopri5' / 'dop853' really runs,
 *     models/hybrid_ode_nn.py:174-181) ... rk.py:568-720, tableau dop853_coefficients.py
 *     (dop853_coef.h, generated from SciPy by gen_dop853_coef.py), dense output rk.py:739-765
 * 'rk4' (fixed step) has no counterpart in the reference; it is the classical tableau.
 *
 * Parity pin: tests/golden/ holds inputs/outputs produced by importing the reference
 * itself (tests/golden/make_golden.py); tests/test_oracle.py checks this file against
 * them.  Mode HODE_ORACLE_RHS_F32 mimics the reference exactly (float32 RHS, float64
 * stepping); HODE_ORACLE_RHS_F64 computes everything in double (the "truth" used for
 * tolerance checks).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/hode.h"
#include "dop853_coef.h"

#define NS HODE_N_STATE
#define ORACLE_RHS_F32 0
#define ORACLE_RHS_F64 1

typedef struct {
  const hode_cfg* cfg;
  int rhs_mode;
  /* per-trajectory views */
  const float* t_obs; /* [T] */
  const float* u[3];  /* NULL | 1 value | [T] */
  const float* theta; /* [17] */
  const float* W;     /* packed MLP or NULL */
  long nfev;
} traj_ctx;

/* ---- input interpolation: models/hybrid_ode_nn.py:217-231 ------------------------ */
/* np.searchsorted(t_eval, t) with side='left' = number of grid points strictly < t. */
static int searchsorted_left_f32(const float* a, int n, float t) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static float input_at_f32(const traj_ctx* c, int ch, float t32) {
  const int mode = c->cfg->in_mode[ch];
  if (mode == HODE_IN_ABSENT) return 0.0f;
  if (mode == HODE_IN_CONST) return c->u[ch][0];
  const int T = c->cfg->n_obs;
  const float* v = c->u[ch];
  int idx = searchsorted_left_f32(c->t_obs, T, t32);
  if (idx == 0) return v[0];
  if (idx >= T) return v[T - 1];
  float t1 = c->t_obs[idx - 1], t2 = c->t_obs[idx];
  float alpha = (t32 - t1) / (t2 - t1);
  return v[idx - 1] + alpha * (v[idx] - v[idx - 1]);
}

static double input_at_f64(const traj_ctx* c, int ch, double t) {
  const int mode = c->cfg->in_mode[ch];
  if (mode == HODE_IN_ABSENT) return 0.0;
  if (mode == HODE_IN_CONST) return (double)c->u[ch][0];
  const int T = c->cfg->n_obs;
  const float* v = c->u[ch];
  int lo = 0, hi = T;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((double)c->t_obs[mid] < t) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return (double)v[0];
  if (lo >= T) return (double)v[T - 1];
  double t1 = c->t_obs[lo - 1], t2 = c->t_obs[lo];
  double alpha = (t - t1) / (t2 - t1);
  return (double)v[lo - 1] + alpha * ((double)v[lo] - (double)v[lo - 1]);
}

/* ---- RHS in float32: the reference evaluates ode_residual on float32 tensors ------ */
static void rhs_f32(traj_ctx* c, double t, const double* y, double* out) {
  const float* th = c->theta;
  const float a_GI = th[0], k_I = th[1], rho = th[2], G_b = th[3], I_b = th[4];
  const float E_max = th[5], EC_50 = th[6], Glu_b = th[7], V_max = th[8], K_m = th[9];
  const float k_L = th[10], k_GE0 = th[11], IGD_50 = th[12], g = th[13];
  const float p_7 = th[14], p_8 = th[15], p_9 = th[16];
  const float t32 = (float)t;
  float s[NS];
  for (int i = 0; i < NS; ++i) s[i] = (float)y[i];
  const float G = s[0], I = s[1], Glu = s[2], GLP1 = s[3], FFA = s[5];
  const float meal = input_at_f32(c, HODE_CH_MEAL, t32);
  const float tvns = input_at_f32(c, HODE_CH_TVNS, t32);
  const float GD = input_at_f32(c, HODE_CH_GD, t32);
  float d[NS];
  /* models/ode_core.py:124-125 */
  float Pi = 1.0f + rho * GLP1;
  d[1] = Pi * a_GI * (G - G_b) - k_I * (I - I_b);
  /* :129-130 */
  float glp1_effect = E_max * (GLP1 / (EC_50 + GLP1));
  d[2] = -glp1_effect * (Glu - Glu_b);
  /* :134-135 */
  d[3] = V_max * (G / (K_m + G)) - k_L * GLP1;
  /* :139-140 */
  float gdg = powf(GD, g);
  float GD_effect = gdg / (powf(IGD_50, g) + gdg);
  float k_GE = k_GE0 * (1.0f - GD_effect);
  /* :144 */
  d[5] = -p_7 * FFA - p_8 * I * FFA + p_9 * G * FFA;
  /* :148-150 */
  float insulin_effect = 0.01f * (I - I_b);
  float glucagon_effect = 0.005f * (Glu - Glu_b);
  d[0] = meal - insulin_effect + glucagon_effect - k_GE * G;
  d[4] = 0.0f; /* :153 */
  if (c->cfg->mlp != HODE_MLP_NONE && c->W) {
    /* models/nn_residual.py:138-146: x = [t, state(6), glp1, tvns] */
    const int H = c->cfg->nn_hidden, L = c->cfg->nn_layers;
    float a[HODE_MAX_HIDDEN], b[HODE_MAX_HIDDEN];
    a[0] = t32;
    for (int i = 0; i < NS; ++i) a[1 + i] = s[i];
    a[7] = GLP1;
    a[8] = tvns;
    const float* w = c->W;
    int n_in = HODE_NN_IN;
    for (int l = 0; l <= L; ++l) {
      const int n_out = (l == L) ? NS : H;
      const float* bias = w + (size_t)n_out * n_in;
      for (int j = 0; j < n_out; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < n_in; ++k) acc += w[(size_t)j * n_in + k] * a[k];
        acc += bias[j];
        b[j] = (l == L) ? acc : (acc > 0.0f ? acc : 0.0f);
      }
      memcpy(a, b, sizeof(float) * n_out);
      w = bias + n_out;
      n_in = n_out;
    }
    for (int i = 0; i < NS; ++i) d[i] = d[i] + a[i]; /* models/hybrid_ode_nn.py:132 */
  }
  for (int i = 0; i < NS; ++i) out[i] = (double)d[i];
  c->nfev++;
}

/* ---- RHS in float64: same formulas, every operation in double ("truth") ----------- */
static void rhs_f64(traj_ctx* c, double t, const double* y, double* out) {
  const float* th = c->theta;
  const double a_GI = th[0], k_I = th[1], rho = th[2], G_b = th[3], I_b = th[4];
  const double E_max = th[5], EC_50 = th[6], Glu_b = th[7], V_max = th[8], K_m = th[9];
  const double k_L = th[10], k_GE0 = th[11], IGD_50 = th[12], g = th[13];
  const double p_7 = th[14], p_8 = th[15], p_9 = th[16];
  const double G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const double meal = input_at_f64(c, HODE_CH_MEAL, t);
  const double tvns = input_at_f64(c, HODE_CH_TVNS, t);
  const double GD = input_at_f64(c, HODE_CH_GD, t);
  double d[NS];
  double Pi = 1.0 + rho * GLP1;
  d[1] = Pi * a_GI * (G - G_b) - k_I * (I - I_b);
  d[2] = -(E_max * (GLP1 / (EC_50 + GLP1))) * (Glu - Glu_b);
  d[3] = V_max * (G / (K_m + G)) - k_L * GLP1;
  double gdg = pow(GD, g);
  double k_GE = k_GE0 * (1.0 - gdg / (pow(IGD_50, g) + gdg));
  d[5] = -p_7 * FFA - p_8 * I * FFA + p_9 * G * FFA;
  d[0] = meal - 0.01 * (I - I_b) + 0.005 * (Glu - Glu_b) - k_GE * G;
  d[4] = 0.0;
  if (c->cfg->mlp != HODE_MLP_NONE && c->W) {
    const int H = c->cfg->nn_hidden, L = c->cfg->nn_layers;
    double a[HODE_MAX_HIDDEN], b[HODE_MAX_HIDDEN];
    a[0] = t;
    for (int i = 0; i < NS; ++i) a[1 + i] = y[i];
    a[7] = GLP1;
    a[8] = tvns;
    const float* w = c->W;
    int n_in = HODE_NN_IN;
    for (int l = 0; l <= L; ++l) {
      const int n_out = (l == L) ? NS : H;
      const float* bias = w + (size_t)n_out * n_in;
      for (int j = 0; j < n_out; ++j) {
        double acc = 0.0;
        for (int k = 0; k < n_in; ++k) acc += (double)w[(size_t)j * n_in + k] * a[k];
        acc += (double)bias[j];
        b[j] = (l == L) ? acc : (acc > 0.0 ? acc : 0.0);
      }
      memcpy(a, b, sizeof(double) * n_out);
      w = bias + n_out;
      n_in = n_out;
    }
    for (int i = 0; i < NS; ++i) d[i] += a[i];
  }
  for (int i = 0; i < NS; ++i) out[i] = d[i];
  c->nfev++;
}

static void rhs(traj_ctx* c, double t, const double* y, double* out) {
  if (c->rhs_mode == ORACLE_RHS_F64) rhs_f64(c, t, y, out); else rhs_f32(c, t, y, out);
}

/* ---- Dormand-Prince 5(4): scipy rk.py:538-565 -------------------------------------- */
static const double DP_C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
static const double DP_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double DP_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double DP_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200,
                               -22.0 / 525, 1.0 / 40};
static const double DP_P[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408,
     701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

static double rms_norm(const double* x, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += x[i] * x[i];
  return sqrt(s) / sqrt((double)n);
}

/* scipy common.py:68-134 (direction = +1, max_step = inf, order = 4) */
static double select_initial_step(traj_ctx* c, double t0, const double* y0, double t_bound,
                                  const double* f0, double rtol, double atol, int order) {
  double interval = fabs(t_bound - t0);
  if (interval == 0.0) return 0.0;
  double scale[NS], a[NS], b[NS];
  for (int i = 0; i < NS; ++i) {
    scale[i] = atol + fabs(y0[i]) * rtol;
    a[i] = y0[i] / scale[i];
    b[i] = f0[i] / scale[i];
  }
  double d0 = rms_norm(a, NS), d1 = rms_norm(b, NS);
  double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  if (h0 > interval) h0 = interval;
  double y1[NS], f1[NS];
  for (int i = 0; i < NS; ++i) y1[i] = y0[i] + h0 * f0[i];
  rhs(c, t0 + h0, y1, f1);
  for (int i = 0; i < NS; ++i) a[i] = (f1[i] - f0[i]) / scale[i];
  double d2 = rms_norm(a, NS) / h0;
  double h1;
  if (d1 <= 1e-15 && d2 <= 1e-15) {
    h1 = fmax(1e-6, h0 * 1e-3);
  } else {
    h1 = pow(0.01 / fmax(d1, d2), 1.0 / (order + 1));
  }
  double h = 100 * h0;
  if (h1 < h) h = h1;
  if (interval < h) h = interval;
  return h;
}

/* Optional recording of accepted steps (for replay by the gradient oracle). */
typedef struct {
  double* t;   /* [cap] start time of accepted step */
  double* h;   /* [cap] */
  int cap, n;
} step_log;

/* Grid point i is a kink when some series input is not flat across (i-1, i, i+1). */
static int is_kink(const traj_ctx* c, int i) {
  const int T = c->cfg->n_obs;
  if (i <= 0 || i >= T - 1) return 0;
  for (int ch = 0; ch < 3; ++ch) {
    if (c->cfg->in_mode[ch] != HODE_IN_SERIES) continue;
    const float* v = c->u[ch];
    if (v[i - 1] != v[i] || v[i] != v[i + 1]) return 1;
  }
  return 0;
}

/* Next step end-point: the first kink strictly after t (HODE_KINK_CLIP), else t_bound. */
static double next_stop(const traj_ctx* c, double t, double t_bound, int* cursor) {
  if (c->cfg->kink_mode != HODE_KINK_CLIP) return t_bound;
  const int T = c->cfg->n_obs;
  int i = *cursor;
  while (i < T - 1 && !((double)c->t_obs[i] > t && is_kink(c, i))) ++i;
  *cursor = i;
  return i < T - 1 ? (double)c->t_obs[i] : t_bound;
}

/* One trajectory, DP5(4) exactly as scipy.solve_ivp(method='RK45', t_eval=...) does it. */
static int solve_dopri5(traj_ctx* c, const float* y0f, float* traj_out, int32_t* n_acc,
                        int32_t* n_rej, step_log* log) {
  const hode_cfg* cfg = c->cfg;
  const int T = cfg->n_obs;
  const double rtol = (double)cfg->rtol, atol = (double)cfg->atol;
  const int max_steps = cfg->max_steps > 0 ? cfg->max_steps : 100000;
  const double t0 = (double)c->t_obs[0], t_bound = (double)c->t_obs[T - 1];
  double t = t0, y[NS], f[NS], K[7][NS];
  for (int i = 0; i < NS; ++i) y[i] = (double)y0f[i];
  int accepted = 0, rejected = 0, status = HODE_ST_OK;
  int ei = 0; /* t_eval_i */
  memset(traj_out, 0, sizeof(float) * (size_t)T * NS);
  rhs(c, t, y, f);
  double h_abs = select_initial_step(c, t0, y, t_bound, f, rtol, atol, 4);
  if (t == t_bound) { /* OdeSolver.step(): already finished; solve_ivp still emits t_eval==t */
    for (; ei < T && (double)c->t_obs[ei] <= t; ++ei)
      for (int i = 0; i < NS; ++i) traj_out[ei * NS + i] = (float)y[i];
    *n_acc = 0; *n_rej = 0;
    return status;
  }
  int attempts = 0, kink_cursor = 1;
  while (1) {
    const double t_stop = next_stop(c, t, t_bound, &kink_cursor);
    double min_step = 10.0 * fabs(nextafter(t, INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    int step_rejected = 0;
    double h, t_new, y_new[NS], f_new[NS];
    while (1) {
      if (h_abs < min_step) { status = HODE_ST_STEP_TOO_SMALL; goto done; }
      if (attempts >= max_steps) { status = HODE_ST_MAX_STEPS; goto done; }
      ++attempts;
      h = h_abs;
      t_new = t + h;
      if (t_new - t_stop > 0) t_new = t_stop; /* scipy clips to t_bound only (rk.py:140-141) */
      h = t_new - t;
      h_abs = fabs(h);
      /* rk_step, rk.py:14-71 */
      memcpy(K[0], f, sizeof f);
      for (int s = 1; s < 6; ++s) {
        double ys[NS];
        for (int i = 0; i < NS; ++i) {
          double dy = 0;
          for (int j = 0; j < s; ++j) dy += K[j][i] * DP_A[s][j];
          ys[i] = y[i] + dy * h;
        }
        rhs(c, t + DP_C[s] * h, ys, K[s]);
      }
      for (int i = 0; i < NS; ++i) {
        double acc = 0;
        for (int j = 0; j < 6; ++j) acc += K[j][i] * DP_B[j];
        y_new[i] = y[i] + h * acc;
      }
      rhs(c, t + h, y_new, f_new);
      memcpy(K[6], f_new, sizeof f_new);
      double e[NS];
      int finite = 1;
      for (int i = 0; i < NS; ++i) {
        double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
        double acc = 0;
        for (int j = 0; j < 7; ++j) acc += K[j][i] * DP_E[j];
        e[i] = acc * h / scale;
        if (!isfinite(y_new[i])) finite = 0;
      }
      double err = rms_norm(e, NS);
      if (err < 1.0) {
        double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
        if (step_rejected && factor > 1.0) factor = 1.0;
        h_abs *= factor;
        break;
      }
      if (!finite || !(err == err)) {
        /* NaN error norm: `error_norm < 1` is False so SciPy keeps shrinking until
           h_abs < min_step; the outcome is TOO_SMALL_STEP. Short-cut it. */
        status = HODE_ST_STEP_TOO_SMALL;
        ++rejected;
        goto done;
      }
      h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
      step_rejected = 1;
      ++rejected;
    }
    ++accepted;
    if (log && log->n < log->cap) { log->t[log->n] = t; log->h[log->n] = h; log->n++; }
    /* dense output at the t_eval points inside (t_old, t_new], ivp.py:701-718 */
    {
      int ei_new = ei;
      while (ei_new < T && (double)c->t_obs[ei_new] <= t_new) ++ei_new;
      if (ei_new > ei) {
        double Q[NS][4];
        for (int i = 0; i < NS; ++i)
          for (int p = 0; p < 4; ++p) {
            double acc = 0;
            for (int j = 0; j < 7; ++j) acc += K[j][i] * DP_P[j][p];
            Q[i][p] = acc;
          }
        for (int k = ei; k < ei_new; ++k) {
          double x = ((double)c->t_obs[k] - t) / h;
          double p1 = x, p2 = x * x, p3 = p2 * x, p4 = p3 * x;
          for (int i = 0; i < NS; ++i) {
            double v = h * (Q[i][0] * p1 + Q[i][1] * p2 + Q[i][2] * p3 + Q[i][3] * p4) + y[i];
            traj_out[k * NS + i] = (float)v;
          }
        }
        ei = ei_new;
      }
    }
    t = t_new;
    memcpy(y, y_new, sizeof y);
    memcpy(f, f_new, sizeof f);
    if (t - t_bound >= 0) break;
  }
done:
  *n_acc = accepted;
  *n_rej = rejected;
  return status;
}

/* One trajectory, DOP853 exactly as scipy.solve_ivp(method='DOP853', t_eval=...) does it: 12 stages with an
   8(5,3) error norm (rk.py:683-691), the RK45 controller with exponent -1/8 (rk.py:111-176), and the 7th-order
   dense output from three extra stages (rk.py:693-712, 739-765), evaluated only for steps that contain t_eval
   points (ivp.py:701-718). */
static int solve_dop853(traj_ctx* c, const float* y0f, float* traj_out, int32_t* n_acc,
                        int32_t* n_rej, step_log* log) {
  const hode_cfg* cfg = c->cfg;
  const int T = cfg->n_obs;
  const double rtol = (double)cfg->rtol, atol = (double)cfg->atol;
  const int max_steps = cfg->max_steps > 0 ? cfg->max_steps : 100000;
  const double t0 = (double)c->t_obs[0], t_bound = (double)c->t_obs[T - 1];
  double t = t0, y[NS], f[NS], K[DOP853_N_STAGES_EXT][NS];
  for (int i = 0; i < NS; ++i) y[i] = (double)y0f[i];
  int accepted = 0, rejected = 0, status = HODE_ST_OK;
  int ei = 0;
  memset(traj_out, 0, sizeof(float) * (size_t)T * NS);
  rhs(c, t, y, f);
  double h_abs = select_initial_step(c, t0, y, t_bound, f, rtol, atol, 7);
  if (t == t_bound) {
    for (; ei < T && (double)c->t_obs[ei] <= t; ++ei)
      for (int i = 0; i < NS; ++i) traj_out[ei * NS + i] = (float)y[i];
    *n_acc = 0; *n_rej = 0;
    return status;
  }
  int attempts = 0, kink_cursor = 1;
  while (1) {
    const double t_stop = next_stop(c, t, t_bound, &kink_cursor);
    double min_step = 10.0 * fabs(nextafter(t, INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    int step_rejected = 0;
    double h, t_new, y_new[NS], f_new[NS];
    while (1) {
      if (h_abs < min_step) { status = HODE_ST_STEP_TOO_SMALL; goto done; }
      if (attempts >= max_steps) { status = HODE_ST_MAX_STEPS; goto done; }
      ++attempts;
      h = h_abs;
      t_new = t + h;
      if (t_new - t_stop > 0) t_new = t_stop;
      h = t_new - t;
      h_abs = fabs(h);
      memcpy(K[0], f, sizeof f);
      for (int s = 1; s < DOP853_N_STAGES; ++s) {
        double ys[NS];
        for (int i = 0; i < NS; ++i) {
          double dy = 0;
          for (int j = 0; j < s; ++j) dy += K[j][i] * DOP853_A[s][j];
          ys[i] = y[i] + dy * h;
        }
        rhs(c, t + DOP853_C[s] * h, ys, K[s]);
      }
      for (int i = 0; i < NS; ++i) {
        double acc = 0;
        for (int j = 0; j < DOP853_N_STAGES; ++j) acc += K[j][i] * DOP853_B[j];
        y_new[i] = y[i] + h * acc;
      }
      rhs(c, t + h, y_new, f_new);
      memcpy(K[DOP853_N_STAGES], f_new, sizeof f_new);
      /* _estimate_error_norm, rk.py:683-691 */
      double e5n = 0, e3n = 0;
      int finite = 1;
      for (int i = 0; i < NS; ++i) {
        const double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
        double a5 = 0, a3 = 0;
        for (int j = 0; j <= DOP853_N_STAGES; ++j) { a5 += K[j][i] * DOP853_E5[j]; a3 += K[j][i] * DOP853_E3[j]; }
        a5 /= scale; a3 /= scale;
        e5n += a5 * a5; e3n += a3 * a3;
        if (!isfinite(y_new[i])) finite = 0;
      }
      double err;
      if (e5n == 0 && e3n == 0) err = 0.0;
      else err = fabs(h) * e5n / sqrt((e5n + 0.01 * e3n) * NS);
      if (err < 1.0) {
        double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -1.0 / 8.0));
        if (step_rejected && factor > 1.0) factor = 1.0;
        h_abs *= factor;
        break;
      }
      if (!finite || !(err == err)) { status = HODE_ST_STEP_TOO_SMALL; ++rejected; goto done; }
      h_abs *= fmax(0.2, 0.9 * pow(err, -1.0 / 8.0));
      step_rejected = 1;
      ++rejected;
    }
    ++accepted;
    if (log && log->n < log->cap) { log->t[log->n] = t; log->h[log->n] = h; log->n++; }
    {
      int ei_new = ei;
      while (ei_new < T && (double)c->t_obs[ei_new] <= t_new) ++ei_new;
      if (ei_new > ei) {
        /* _dense_output_impl, rk.py:693-712 */
        for (int s = DOP853_N_STAGES + 1; s < DOP853_N_STAGES_EXT; ++s) {
          double ys[NS];
          for (int i = 0; i < NS; ++i) {
            double dy = 0;
            for (int j = 0; j < s; ++j) dy += K[j][i] * DOP853_A[s][j];
            ys[i] = y[i] + dy * h;
          }
          rhs(c, t + DOP853_C[s] * h, ys, K[s]);
        }
        double F[7][NS];
        for (int i = 0; i < NS; ++i) {
          const double dy = y_new[i] - y[i];
          F[0][i] = dy;
          F[1][i] = h * K[0][i] - dy;
          F[2][i] = 2 * dy - h * (f_new[i] + K[0][i]);
          for (int q = 0; q < 4; ++q) {
            double acc = 0;
            for (int j = 0; j < DOP853_N_STAGES_EXT; ++j) acc += DOP853_D[q][j] * K[j][i];
            F[3 + q][i] = h * acc;
          }
        }
        for (int k = ei; k < ei_new; ++k) {
          /* Dop853DenseOutput._call_impl, rk.py:746-765 */
          const double x = ((double)c->t_obs[k] - t) / h;
          for (int i = 0; i < NS; ++i) {
            double v = 0;
            for (int q = 0; q < 7; ++q) {
              v += F[6 - q][i];
              v *= (q % 2 == 0) ? x : 1 - x;
            }
            traj_out[k * NS + i] = (float)(v + y[i]);
          }
        }
        ei = ei_new;
      }
    }
    t = t_new;
    memcpy(y, y_new, sizeof y);
    memcpy(f, f_new, sizeof f);
    if (t - t_bound >= 0) break;
  }
done:
  *n_acc = accepted;
  *n_rej = rejected;
  return status;
}

/* One trajectory, classical RK4 with n_substeps equal steps per observation interval. */
static int solve_rk4(traj_ctx* c, const float* y0f, float* traj_out, int32_t* n_acc) {
  const hode_cfg* cfg = c->cfg;
  const int T = cfg->n_obs, nsub = cfg->n_substeps > 0 ? cfg->n_substeps : 1;
  double y[NS], k1[NS], k2[NS], k3[NS], k4[NS], ys[NS];
  for (int i = 0; i < NS; ++i) { y[i] = (double)y0f[i]; traj_out[i] = y0f[i]; }
  int steps = 0;
  for (int n = 0; n + 1 < T; ++n) {
    const double ta = (double)c->t_obs[n], tb = (double)c->t_obs[n + 1];
    const double h = (tb - ta) / nsub;
    for (int s = 0; s < nsub; ++s) {
      const double t = ta + s * h;
      rhs(c, t, y, k1);
      for (int i = 0; i < NS; ++i) ys[i] = y[i] + 0.5 * h * k1[i];
      rhs(c, t + 0.5 * h, ys, k2);
      for (int i = 0; i < NS; ++i) ys[i] = y[i] + 0.5 * h * k2[i];
      rhs(c, t + 0.5 * h, ys, k3);
      for (int i = 0; i < NS; ++i) ys[i] = y[i] + h * k3[i];
      rhs(c, t + h, ys, k4);
      for (int i = 0; i < NS; ++i) y[i] += (h / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
      ++steps;
    }
    for (int i = 0; i < NS; ++i) traj_out[(n + 1) * NS + i] = (float)y[i];
  }
  *n_acc = steps;
  return HODE_ST_OK;
}

int64_t hode_oracle_mlp_param_count(int32_t H, int32_t L) {
  int64_t n = (int64_t)HODE_NN_IN * H + H;
  for (int l = 1; l < L; ++l) n += (int64_t)H * H + H;
  n += (int64_t)H * NS + NS;
  return n;
}

static void bind_traj(traj_ctx* c, const hode_cfg* cfg, int rhs_mode, long b, long s,
                      const float* t_obs, const float* const u[3], const float* theta,
                      const float* W) {
  const int T = cfg->n_obs;
  c->cfg = cfg;
  c->rhs_mode = rhs_mode;
  c->t_obs = cfg->t_per_traj ? t_obs + b * T : t_obs;
  for (int ch = 0; ch < 3; ++ch) {
    if (cfg->in_mode[ch] == HODE_IN_SERIES) c->u[ch] = u[ch] + b * T;
    else if (cfg->in_mode[ch] == HODE_IN_CONST) c->u[ch] = u[ch] + b;
    else c->u[ch] = NULL;
  }
  c->theta = theta + s * HODE_N_THETA;
  c->W = (cfg->mlp != HODE_MLP_NONE && W) ? W + s * hode_oracle_mlp_param_count(cfg->nn_hidden, cfg->nn_layers) : NULL;
  c->nfev = 0;
}

/* Work-sharing over trajectories with plain pthreads (chunks of 16 units). */
typedef struct {
  const hode_cfg* cfg;
  int rhs_mode;
  const float *y0, *t_obs, *u[3], *theta, *W;
  float* traj;
  int32_t *status, *counters;
  long total, B;
  volatile long next;
  volatile long long nfev;
} job;

static void* worker(void* arg) {
  job* j = (job*)arg;
  const hode_cfg* cfg = j->cfg;
  const int T = cfg->n_obs;
  long long nfev = 0;
  for (;;) {
    long lo = __sync_fetch_and_add(&j->next, 16L);
    if (lo >= j->total) break;
    long hi = lo + 16 < j->total ? lo + 16 : j->total;
    for (long unit = lo; unit < hi; ++unit) {
      const long s = unit / j->B, b = unit % j->B;
      traj_ctx c;
      bind_traj(&c, cfg, j->rhs_mode, b, s, j->t_obs, j->u, j->theta, j->W);
      int32_t na = 0, nr = 0;
      int st;
      float* out = j->traj + (size_t)unit * T * NS;
      if (cfg->solver == HODE_SOLVER_RK4) st = solve_rk4(&c, j->y0 + b * NS, out, &na);
      else if (cfg->solver == HODE_SOLVER_DOP853) st = solve_dop853(&c, j->y0 + b * NS, out, &na, &nr, NULL);
      else st = solve_dopri5(&c, j->y0 + b * NS, out, &na, &nr, NULL);
      if (j->status) j->status[unit] = st;
      if (j->counters) { j->counters[unit] = na; j->counters[j->total + unit] = nr; }
      nfev += c.nfev;
    }
  }
  __sync_fetch_and_add(&j->nfev, nfev);
  return NULL;
}

/*
 * Host-memory mirror of hode_rollout_fwd (same argument meaning, every pointer on the
 * host).  rhs_mode: 0 = float32 RHS / float64 stepping (the reference's arithmetic),
 * 1 = all float64.  n_threads <= 1 runs serially like the reference's `for b` loop
 * (models/hybrid_ode_nn.py:184); >1 shards trajectories over pthreads.
 * nfev_total (optional) receives the number of RHS evaluations.
 */
int hode_oracle_rollout(const hode_cfg* cfg, int rhs_mode, int n_threads, const float* y0,
                        const float* t_obs, const float* u_meal, const float* u_tvns,
                        const float* u_gd, const float* theta, const float* W, float* traj,
                        int32_t* status, int32_t* counters, int64_t* nfev_total) {
  if (!cfg || !y0 || !t_obs || !theta || !traj) return HODE_E_NULL;
  const long B = cfg->n_traj, S = cfg->n_samples > 0 ? cfg->n_samples : 1;
  const int T = cfg->n_obs;
  if (B < 0 || T < 1) return HODE_E_SIZE;
  if (cfg->mlp != HODE_MLP_NONE && (cfg->nn_hidden > HODE_MAX_HIDDEN || cfg->nn_hidden < 1))
    return HODE_E_UNSUPPORTED;
  job j;
  j.cfg = cfg; j.rhs_mode = rhs_mode; j.y0 = y0; j.t_obs = t_obs;
  j.u[0] = u_meal; j.u[1] = u_tvns; j.u[2] = u_gd;
  j.theta = theta; j.W = W; j.traj = traj; j.status = status; j.counters = counters;
  j.total = S * B; j.B = B; j.next = 0; j.nfev = 0;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  if (n_threads == 1) {
    worker(&j);
  } else {
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < n_threads; ++i)
      if (pthread_create(&th[started], NULL, worker, &j) == 0) ++started;
    if (started == 0) worker(&j);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  }
  if (nfev_total) *nfev_total = j.nfev;
  return 0;
}

/* Same, for ONE trajectory, also returning the accepted-step log (start time, h). */
int hode_oracle_rollout_log(const hode_cfg* cfg, int rhs_mode, long b, const float* y0,
                            const float* t_obs, const float* u_meal, const float* u_tvns,
                            const float* u_gd, const float* theta, const float* W,
                            float* traj_row, double* step_t, double* step_h, int32_t cap,
                            int32_t* n_steps) {
  const float* u[3] = {u_meal, u_tvns, u_gd};
  traj_ctx c;
  bind_traj(&c, cfg, rhs_mode, b, 0, t_obs, u, theta, W);
  step_log log = {step_t, step_h, cap, 0};
  int32_t na = 0, nr = 0;
  int st = cfg->solver == HODE_SOLVER_DOP853 ? solve_dop853(&c, y0 + b * NS, traj_row, &na, &nr, &log)
                                             : solve_dopri5(&c, y0 + b * NS, traj_row, &na, &nr, &log);
  *n_steps = log.n;
  return st;
}

/* Host mirror of hode_rhs: one batched evaluation of f_physio + g_NN (out [B,6]). */
int hode_oracle_rhs(const hode_cfg* cfg, int rhs_mode, const float* t, const float* state,
                    const float* u_meal, const float* u_tvns, const float* u_gd,
                    const float* theta, const float* W, double* out) {
  hode_cfg one = *cfg;
  one.n_obs = 1;
  for (int ch = 0; ch < 3; ++ch)
    if (one.in_mode[ch] == HODE_IN_SERIES) one.in_mode[ch] = HODE_IN_CONST;
  const float* u[3] = {u_meal, u_tvns, u_gd};
  for (long b = 0; b < cfg->n_traj; ++b) {
    traj_ctx c;
    one.t_per_traj = 1;
    bind_traj(&c, &one, rhs_mode, b, 0, t, u, theta, W);
    double y[NS];
    for (int i = 0; i < NS; ++i) y[i] = (double)state[b * NS + i];
    rhs(&c, (double)t[b], y, out + b * NS);
  }
  return 0;
}
