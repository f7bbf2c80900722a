// hode_train.cu — one training step of the reference in ONE library call.
//
// Replaces HybridODENN.loss (reference models/hybrid_ode_nn.py:263-351) + the body of train_epoch's batch loop
// (train/train_hybrid.py:244-261: loss.backward(), clip_grad_norm_, Adam.step()):
//   data term     MSE(rollout, observations)                               (:284-288)
//   physics term  mean_k MSE((Phi_0.1(x_k) - x_k) / 0.1, f(t_k, x_k)) over the drawn grid indices k, x_k the predicted
//                 states at index k, Phi_0.1 a re-solve over local time [0, 0.1] with the inputs frozen at index k
//                 (:297-333) — all (index, trajectory) pairs stacked into one rollout, one RHS and one RHS-VJP launch
//   L2 term       lambda2 * (lambda2 * sum ||weight||^2) over the Linear weights, not the biases (:336-345 with
//                 models/nn_residual.py:198-223: the reference applies lambda2 twice)
//   gradient      as in the reference only the physics residual (through f, not through the graph-free solves) and the
//                 L2 term carry gradient; with data_gradient = 1 the discrete adjoint of the data term is added
//                 (hode_rollout_bwd) — the through-solver gradient BASELINE.json's north star adds
//   clip + Adam   torch.nn.utils.clip_grad_norm_ (scale = clip / (norm + 1e-6) when norm > clip) and torch.optim.Adam's
//                 update (bias-corrected, weight_decay 0) on the packed network parameters, in place
// Everything is stream-ordered on the caller's stream, allocation-free (one workspace) and free of host
// synchronisation, so the call can be captured in a CUDA graph.  Reductions are block partials summed in a fixed
// order: bit-reproducible.
#include <stdio.h>

#include "hode_kernels.h"

namespace hode {
namespace {

constexpr int TB = 256;

// rows of the stacked physics problem: row r = k * B + b
__global__ void __launch_bounds__(TB) physics_gather_kernel(const float* __restrict__ traj, const float* __restrict__ t_obs,
                                                            const float* __restrict__ u0, const float* __restrict__ u1,
                                                            const float* __restrict__ u2, const int32_t* __restrict__ idx, int B,
                                                            int T, int n_pts, int t_per_traj, int m0, int m1, int m2,
                                                            float* __restrict__ state, float* __restrict__ t_rows,
                                                            float* __restrict__ r0, float* __restrict__ r1, float* __restrict__ r2) {
  const long r = (long)blockIdx.x * TB + threadIdx.x;
  if (r >= (long)n_pts * B) return;
  const int k = (int)(r / B);
  const long b = r - (long)k * B;
  int i = idx[k];
  i = i < 0 ? 0 : (i >= T ? T - 1 : i);
#pragma unroll
  for (int c = 0; c < NS; ++c) state[r * NS + c] = traj[((size_t)b * T + i) * NS + c];
  t_rows[r] = t_per_traj ? t_obs[b * T + i] : t_obs[i];
  if (m0) r0[r] = m0 == HODE_IN_SERIES ? u0[b * T + i] : u0[b];
  if (m1) r1[r] = m1 == HODE_IN_SERIES ? u1[b * T + i] : u1[b];
  if (m2) r2[r] = m2 == HODE_IN_SERIES ? u2[b * T + i] : u2[b];
}

// resid = f(t_k, x_k) - (Phi(x_k) - x_k) / dt;  grad_out = scale * resid;  block partial sums of resid^2
__global__ void __launch_bounds__(TB) physics_resid_kernel(const float* __restrict__ nxt2, const float* __restrict__ state,
                                                           const float* __restrict__ dx_ode, float inv_dt, float scale, long n,
                                                           float* __restrict__ grad_out, float* __restrict__ partial) {
  __shared__ float red[TB];
  float acc = 0.f;
  for (int j = 0; j < 8; ++j) {
    const long e = ((long)blockIdx.x * 8 + j) * TB + threadIdx.x;   // element of [rows, 6]
    if (e < n) {
      const long row = e / NS;
      const int c = (int)(e - row * NS);
      const float fd = (nxt2[(row * 2 + 1) * NS + c] - state[e]) * inv_dt;   // row 1 of the [rows, 2, 6] re-solve
      const float d = dx_ode[e] - fd;
      grad_out[e] = scale * d;
      acc = fmaf(d, d, acc);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = TB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// data term without gradient: block partial sums of (traj - obs)^2
__global__ void __launch_bounds__(TB) sq_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, long n,
                                                     float* __restrict__ partial) {
  __shared__ float red[TB];
  float acc = 0.f;
  for (int j = 0; j < 16; ++j) {
    const long e = ((long)blockIdx.x * 16 + j) * TB + threadIdx.x;
    if (e < n) { const float d = a[e] - b[e]; acc = fmaf(d, d, acc); }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = TB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// 1 if packed parameter i is a Linear WEIGHT (not a bias), reference models/nn_residual.py:214-216
__device__ __forceinline__ bool is_weight(int i, int H, int L) {
  int off = 0, n_in = HODE_NN_IN;
  for (int l = 0; l <= L; ++l) {
    const int n_out = l == L ? NS : H;
    if (i < off + n_out * n_in) return true;
    off += n_out * n_in;
    if (i < off + n_out) return false;
    off += n_out;
    n_in = n_out;
  }
  return false;
}

// g = g_data + g_physics + 2 lambda2^2 w (weights only);  block partials of ||g||^2 and of sum w^2 (weights only)
__global__ void __launch_bounds__(TB) combine_grad_kernel(const float* __restrict__ W, const float* __restrict__ g_data,
                                                          const float* __restrict__ g_phys, float reg_coeff, int P, int H, int L,
                                                          float* __restrict__ g, float* __restrict__ part_g2,
                                                          float* __restrict__ part_w2) {
  __shared__ float r1[TB], r2[TB];
  const int i = blockIdx.x * TB + threadIdx.x;
  float gg = 0.f, ww = 0.f;
  if (i < P) {
    const float w = W[i];
    const bool wt = is_weight(i, H, L);
    float v = (g_data ? g_data[i] : 0.f) + (g_phys ? g_phys[i] : 0.f);
    if (wt) { v = fmaf(2.0f * reg_coeff, w, v); ww = w * w; }
    g[i] = v;
    gg = v * v;
  }
  r1[threadIdx.x] = gg; r2[threadIdx.x] = ww;
  __syncthreads();
  for (int o = TB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_g2[blockIdx.x] = r1[0]; part_w2[blockIdx.x] = r2[0]; }
}

// scalars: [0] total [1] data [2] physics [3] reg [4] grad norm (before clipping) [5] clip scale
__global__ void finalize_kernel(const float* __restrict__ p_data, int n_data, double inv_n_data,
                                const float* __restrict__ p_phys, int n_phys, double inv_n_phys,
                                const float* __restrict__ p_g2, const float* __restrict__ p_w2, int n_blk, float lambda1,
                                float lambda2, float clip, int dev_step, float b1, float b2, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double d = 0, ph = 0, g2 = 0, w2 = 0;
  for (int i = 0; i < n_data; ++i) d += (double)p_data[i];
  for (int i = 0; i < n_phys; ++i) ph += (double)p_phys[i];
  for (int i = 0; i < n_blk; ++i) { g2 += (double)p_g2[i]; w2 += (double)p_w2[i]; }
  const double data = d * inv_n_data, phys = ph * inv_n_phys;
  const double reg = (double)lambda2 * w2;   // regularization_loss(l2_weight = lambda2); the total multiplies by lambda2 again
  const double norm = sqrt(g2);
  out[0] = (float)(data + (double)lambda1 * phys + (double)lambda2 * reg);
  out[1] = (float)data;
  out[2] = (float)phys;
  out[3] = (float)reg;
  out[4] = (float)norm;
  // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
  float sc = 1.0f;
  if (clip > 0.f) { const double c = (double)clip / (norm + 1e-6); sc = c < 1.0 ? (float)c : 1.0f; }
  out[5] = sc;
  if (dev_step) {   // CUDA-graph replays: the update count lives on the device ([6]); [7], [8] = Adam's bias corrections
    const float t = out[6] + 1.0f;
    out[6] = t;
    out[7] = 1.0f - powf(b1, t);
    out[8] = sqrtf(1.0f - powf(b2, t));
  }
}

// torch.optim.Adam (amsgrad off, weight_decay 0, maximize off), on the clipped gradient; g is rewritten with the
// clipped gradient (what p.grad holds after clip_grad_norm_)
__global__ void __launch_bounds__(TB) adam_kernel(float* __restrict__ W, float* __restrict__ g, float* __restrict__ m,
                                                  float* __restrict__ v, const float* __restrict__ scalars, int P, float lr,
                                                  float b1, float b2, float eps, float bc1, float bc2_sqrt, int update) {
  const int i = blockIdx.x * TB + threadIdx.x;
  if (i >= P) return;
  if (update == 2) { bc1 = scalars[7]; bc2_sqrt = scalars[8]; }   // device-side step counter (finalize_kernel)
  const float gi = g[i] * scalars[5];
  g[i] = gi;
  if (!update) return;
  const float mi = fmaf(1.0f - b1, gi - m[i], m[i]);        // lerp(m, g, 1 - beta1)
  const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  W[i] = W[i] - (lr / bc1) * (mi / denom);
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TrainPlan {
  hode_cfg c2, c3;          // physics re-solve (rows x [0, dt]), RHS / RHS-VJP (rows x 1)
  long rows;
  size_t fwd1, bwd1, fwd2, bwd3;
  size_t o_fwd1, o_bwd1, o_gtraj, o_gdata, o_gphys, o_state, o_trow, o_u[3], o_tloc, o_traj2, o_st2, o_dx, o_gout, o_fwd2, o_bwd3,
      o_part, total;
  int nb_data, nb_phys, nb_par;
};

int make_plan(const hode_cfg* cfg, const hode_train_cfg* tc, TrainPlan& p) {
  const long B = cfg->n_traj, T = cfg->n_obs;
  const int P = (int)hode_mlp_param_count(cfg->nn_hidden, cfg->nn_layers);
  p.rows = (long)tc->n_physics * B;
  hode_cfg c1 = *cfg;
  c1.save_steps = tc->data_gradient ? 1 : 0;
  int rc = hode_workspace_bytes(&c1, &p.fwd1, &p.bwd1);
  if (rc) return rc;
  if (!tc->data_gradient) p.bwd1 = 0;
  p.c2 = *cfg;
  p.c2.n_traj = (int32_t)p.rows; p.c2.n_obs = 2; p.c2.t_per_traj = 0; p.c2.save_steps = 0; p.c2.max_saved_steps = 0;
  for (int ch = 0; ch < 3; ++ch) p.c2.in_mode[ch] = cfg->in_mode[ch] == HODE_IN_ABSENT ? HODE_IN_ABSENT : HODE_IN_CONST;
  p.c3 = p.c2;
  p.c3.n_obs = 1; p.c3.t_per_traj = 1; p.c3.rhs_part = HODE_RHS_FULL;
  p.fwd2 = p.bwd3 = 0;
  if (p.rows > 0) {
    size_t dummy = 0;
    rc = hode_workspace_bytes(&p.c2, &p.fwd2, &dummy);
    if (rc) return rc;
    rc = hode_workspace_bytes(&p.c3, &dummy, &p.bwd3);
    if (rc) return rc;
  }
  const long n_data = B * T * NS, n_phys = p.rows * NS;
  p.nb_data = (int)((n_data + 16 * TB - 1) / (16 * TB));
  p.nb_phys = (int)((n_phys + 8 * TB - 1) / (8 * TB));
  p.nb_par = (P + TB - 1) / TB;
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = al256(off + bytes); return o; };
  p.o_fwd1 = carve(p.fwd1);
  p.o_bwd1 = carve(p.bwd1);
  p.o_gtraj = carve(tc->data_gradient ? (size_t)n_data * 4 : 0);
  p.o_gdata = carve((size_t)(P + HODE_N_THETA) * 4);
  p.o_gphys = carve((size_t)(P + HODE_N_THETA) * 4);
  p.o_state = carve((size_t)n_phys * 4);
  p.o_trow = carve((size_t)p.rows * 4);
  for (int ch = 0; ch < 3; ++ch) p.o_u[ch] = carve(cfg->in_mode[ch] != HODE_IN_ABSENT ? (size_t)p.rows * 4 : 0);
  p.o_tloc = carve(2 * 4);
  p.o_traj2 = carve((size_t)n_phys * 2 * 4);
  p.o_st2 = carve((size_t)p.rows * 4);
  p.o_dx = carve((size_t)n_phys * 4);
  p.o_gout = carve((size_t)n_phys * 4);
  p.o_fwd2 = carve(p.fwd2);
  p.o_bwd3 = carve(p.bwd3);
  p.o_part = carve((size_t)(p.nb_data + p.nb_phys + 2 * p.nb_par + 16) * 4);
  p.total = off;
  return 0;
}

__global__ void set2_kernel(float* p, float a, float b) { p[0] = a; p[1] = b; }

}  // namespace
}  // namespace hode

using namespace hode;

extern "C" {

static thread_local char g_terr[256] = "ok";

int hode_train_step_workspace_bytes(const hode_cfg* cfg, const hode_train_cfg* tc, size_t* bytes) {
  if (!cfg || !tc || !bytes) return HODE_E_NULL;
  if (tc->struct_bytes != (int32_t)sizeof(hode_train_cfg)) return HODE_E_SIZE;
  if (cfg->mlp == HODE_MLP_NONE || cfg->n_samples != 1 || tc->n_physics < 0 || tc->n_physics > 64) return HODE_E_UNSUPPORTED;
  TrainPlan p;
  const int rc = make_plan(cfg, tc, p);
  if (rc) return rc;
  *bytes = p.total;
  return 0;
}

int hode_train_step(const hode_cfg* cfg, const hode_train_cfg* tc, const float* y0, const float* t_obs, const float* u_meal,
                    const float* u_tvns, const float* u_gd, const float* theta, float* W, const float* obs,
                    const int32_t* physics_idx, float* adam_m, float* adam_v, float* grad_W, float* scalars, float* traj,
                    int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  if (!cfg || !tc) return HODE_E_NULL;
  if (tc->struct_bytes != (int32_t)sizeof(hode_train_cfg)) return HODE_E_SIZE;
  if (cfg->mlp == HODE_MLP_NONE || cfg->n_samples != 1 || tc->n_physics < 0 || tc->n_physics > 64) return HODE_E_UNSUPPORTED;
  if (!y0 || !t_obs || !theta || !W || !obs || !grad_W || !scalars || !traj) return HODE_E_NULL;
  if (tc->n_physics > 0 && !physics_idx) return HODE_E_NULL;
  if (tc->adam_step != 0 && (!adam_m || !adam_v)) return HODE_E_NULL;
  TrainPlan p;
  int rc = make_plan(cfg, tc, p);
  if (rc) return rc;
  if (!workspace || workspace_bytes < p.total) return HODE_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const long B = cfg->n_traj, T = cfg->n_obs;
  const int P = (int)hode_mlp_param_count(cfg->nn_hidden, cfg->nn_layers);
  const long n_data = B * T * NS, n_phys = p.rows * NS;
  float* part = (float*)(ws + p.o_part);
  float *p_data = part, *p_phys = part + p.nb_data, *p_g2 = p_phys + p.nb_phys, *p_w2 = p_g2 + p.nb_par;
  float* g_data = (float*)(ws + p.o_gdata);
  float* g_phys = (float*)(ws + p.o_gphys);
  if (B == 0) {
    cudaMemsetAsync(scalars, 0, 6 * sizeof(float), st);   // (the device step counter [6] is left alone)
    cudaMemsetAsync(grad_W, 0, (size_t)P * sizeof(float), st);
    return 0;
  }

  // ---- data term: rollout (+ discrete adjoint) ---------------------------------------------------------------------
  hode_cfg c1 = *cfg;
  c1.save_steps = tc->data_gradient ? 1 : 0;
  if (tc->data_gradient) {
    float loss_dummy_unused;
    (void)loss_dummy_unused;
    rc = hode_loss_fused_fwd_bwd(&c1, y0, t_obs, u_meal, u_tvns, u_gd, theta, W, obs, traj, status, nullptr, scalars + 1,
                                 (float*)(ws + p.o_gtraj), nullptr, g_data + P, g_data, ws + p.o_fwd1, p.fwd1, ws + p.o_bwd1,
                                 p.bwd1, stream);
    if (rc) return rc;
  } else {
    rc = hode_rollout_fwd(&c1, y0, t_obs, u_meal, u_tvns, u_gd, theta, W, traj, status, nullptr, ws + p.o_fwd1, p.fwd1, stream);
    if (rc) return rc;
  }
  count_launch();
  sq_diff_kernel<<<p.nb_data, TB, 0, st>>>(traj, obs, n_data, p_data);

  // ---- physics term ----------------------------------------------------------------------------------------------------
  if (p.rows > 0) {
    float* state = (float*)(ws + p.o_state);
    float* t_rows = (float*)(ws + p.o_trow);
    float* ur[3];
    for (int ch = 0; ch < 3; ++ch) ur[ch] = cfg->in_mode[ch] != HODE_IN_ABSENT ? (float*)(ws + p.o_u[ch]) : nullptr;
    float* t_loc = (float*)(ws + p.o_tloc);
    float* traj2 = (float*)(ws + p.o_traj2);
    float* dx = (float*)(ws + p.o_dx);
    float* gout = (float*)(ws + p.o_gout);
    const float dt = tc->physics_dt > 0.f ? tc->physics_dt : 0.1f;
    count_launch();
    set2_kernel<<<1, 1, 0, st>>>(t_loc, 0.0f, dt);
    count_launch();
    physics_gather_kernel<<<(unsigned)((p.rows + TB - 1) / TB), TB, 0, st>>>(
        traj, t_obs, u_meal, u_tvns, u_gd, physics_idx, (int)B, (int)T, tc->n_physics, cfg->t_per_traj, cfg->in_mode[0],
        cfg->in_mode[1], cfg->in_mode[2], state, t_rows, ur[0], ur[1], ur[2]);
    rc = hode_rollout_fwd(&p.c2, state, t_loc, ur[0], ur[1], ur[2], theta, W, traj2, (int32_t*)(ws + p.o_st2), nullptr,
                          ws + p.o_fwd2, p.fwd2, stream);
    if (rc) return rc;
    rc = hode_rhs(&p.c3, t_rows, state, ur[0], ur[1], ur[2], theta, W, dx, stream);
    if (rc) return rc;
    // d physics / d dx_ode = 2 (dx_ode - dx_fd) / (rows * 6), times lambda1
    count_launch();
    physics_resid_kernel<<<p.nb_phys, TB, 0, st>>>(traj2, state, dx, 1.0f / dt, tc->lambda1 * 2.0f / (float)n_phys, n_phys, gout,
                                                  p_phys);
    rc = hode_rhs_vjp(&p.c3, t_rows, state, ur[0], ur[1], ur[2], theta, W, gout, nullptr, g_phys + P, g_phys, ws + p.o_bwd3,
                      p.bwd3, stream);
    if (rc) return rc;
  }

  // ---- gradient, norms, losses, clip, Adam -----------------------------------------------------------------------------
  count_launch();
  combine_grad_kernel<<<p.nb_par, TB, 0, st>>>(W, tc->data_gradient ? g_data : nullptr, p.rows > 0 ? g_phys : nullptr,
                                              tc->lambda2 * tc->lambda2, P, cfg->nn_hidden, cfg->nn_layers, grad_W, p_g2, p_w2);
  count_launch();
  finalize_kernel<<<1, 32, 0, st>>>(p_data, p.nb_data, 1.0 / (double)n_data, p_phys, p.rows > 0 ? p.nb_phys : 0,
                                    p.rows > 0 ? 1.0 / (double)n_phys : 0.0, p_g2, p_w2, p.nb_par, tc->lambda1, tc->lambda2,
                                    tc->grad_clip, tc->adam_step < 0 ? 1 : 0, tc->beta1, tc->beta2, scalars);
  const int t_ = tc->adam_step;
  const float bc1 = t_ > 0 ? 1.0f - powf(tc->beta1, (float)t_) : 1.0f;
  const float bc2 = t_ > 0 ? sqrtf(1.0f - powf(tc->beta2, (float)t_)) : 1.0f;
  count_launch();
  adam_kernel<<<p.nb_par, TB, 0, st>>>(W, grad_W, adam_m, adam_v, scalars, P, tc->lr, tc->beta1, tc->beta2, tc->eps, bc1, bc2,
                                       t_ > 0 ? 1 : (t_ < 0 ? 2 : 0));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_terr, sizeof g_terr, "hode_train_step launch: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // extern "C"
