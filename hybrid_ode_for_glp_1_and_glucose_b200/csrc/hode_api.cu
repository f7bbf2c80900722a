// hode_api.cu — the extern "C" surface of libhode.so (see include/hode.h).
//
// Argument validation, workspace carving and kernel selection.  No torch types, no
// allocation (except in the *_host convenience entry, which uses the stream-ordered
// allocator for its staging buffers), no CPU fallback: a configuration the CUDA kernels do
// not cover returns HODE_E_UNSUPPORTED.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include <cub/device/device_radix_sort.cuh>

#include "hode_kernels.h"

namespace hode {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace hode
using hode::count_launch;

namespace {

thread_local char g_err[512] = "ok";

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof g_err, fmt, detail);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  size_t off_n, off_rec, off_tc, total;
  int max_saved;
};

bool uses_tensor_cores(const hode_cfg* c) {
  return c->mlp == HODE_MLP_TF32X3 || c->mlp == HODE_MLP_TF32 || c->mlp == HODE_MLP_TF32BF16 ||
         c->mlp == HODE_MLP_TF32X2BF16 || c->mlp == HODE_MLP_F16BF16X2;
}

int max_saved_steps(const hode_cfg* c) {
  if (c->solver == HODE_SOLVER_RK4) {
    const int nsub = c->n_substeps > 0 ? c->n_substeps : 1;
    return (c->n_obs - 1) * nsub;
  }
  if (c->max_saved_steps > 0) return c->max_saved_steps;
  // default: derived from the problem.  With kink clipping a series input that changes at every grid point
  // forces at least T - 1 accepted steps; 2 (T - 1) leaves room for the controller's own steps in between.
  const int by_grid = 2 * (c->n_obs - 1);
  return by_grid > 128 ? by_grid : 128;
}

bool adjoint_on_tensor_cores(const hode_cfg* c) {
  return uses_tensor_cores(c) && hode::adj_tc_supported(c->nn_hidden, c->nn_layers);
}

// floats per accepted-step record: the tensor-core adjoint reads the stage derivatives from the record
int rec_floats(const hode_cfg* c) { return adjoint_on_tensor_cores(c) ? hode::HODE_REC_FLOATS_K : hode::HODE_REC_FLOATS; }

Workspace fwd_workspace(const hode_cfg* c) {
  Workspace w{};
  const size_t units = (size_t)(c->n_samples > 0 ? c->n_samples : 1) * (size_t)c->n_traj;
  w.max_saved = max_saved_steps(c);
  size_t off = 0;
  if (c->save_steps) {
    w.off_n = off; off = align_up(off + units * sizeof(int32_t), 256);
    // one record per (unit, step): t, h, y[6], k1[6] (+ k2..k6 for the tensor-core adjoint), hode_common.cuh step_rec
    w.off_rec = off; off = align_up(off + units * w.max_saved * (size_t)rec_floats(c) * sizeof(float), 256);
  }
  if (uses_tensor_cores(c)) {
    // pre-split weight images (one per parameter set) + the per-set trajectory queue counters
    w.off_tc = off;
    off = align_up(off + hode::tc_workspace_bytes(c->n_samples, c->nn_layers, c->n_traj), 256);
  }
  w.total = off;
  return w;
}

int validate(const hode_cfg* c) {
  if (!c) return fail(HODE_E_NULL, "cfg is NULL");
  if (c->struct_bytes != (int32_t)sizeof(hode_cfg))
    return fail(HODE_E_SIZE, "cfg.struct_bytes != sizeof(hode_cfg): header/library mismatch");
  if (c->n_traj < 0 || c->n_obs < 1) return fail(HODE_E_SIZE, "n_traj < 0 or n_obs < 1");
  if (c->n_samples < 1) return fail(HODE_E_SIZE, "n_samples < 1");
  for (int ch = 0; ch < 3; ++ch)
    if (c->in_mode[ch] < HODE_IN_ABSENT || c->in_mode[ch] > HODE_IN_SERIES)
      return fail(HODE_E_SHAPE, "in_mode out of range");
  if (c->solver != HODE_SOLVER_RK4 && c->solver != HODE_SOLVER_DOPRI5 && c->solver != HODE_SOLVER_DOP853)
    return fail(HODE_E_UNSUPPORTED, "unknown solver");
  if (c->solver == HODE_SOLVER_DOP853) {
    // the compatibility solver: FP32 CUDA-core kernels, forward only
    if (c->mlp != HODE_MLP_NONE && c->mlp != HODE_MLP_FP32)
      return fail(HODE_E_UNSUPPORTED, "HODE_SOLVER_DOP853 runs on the FP32 kernels: mlp must be HODE_MLP_NONE or HODE_MLP_FP32");
    if (c->save_steps)
      return fail(HODE_E_UNSUPPORTED, "HODE_SOLVER_DOP853 is forward only (no step records / adjoint)");
  }
  if (c->mlp < HODE_MLP_NONE || c->mlp > HODE_MLP_F16BF16X2)
    return fail(HODE_E_UNSUPPORTED, "unknown mlp arithmetic");
  if (c->mlp != HODE_MLP_NONE) {
    if (c->nn_hidden < 1 || c->nn_hidden > HODE_MAX_HIDDEN || c->nn_layers < 1 ||
        c->nn_layers > HODE_MAX_LAYERS)
      return fail(HODE_E_UNSUPPORTED, "nn_hidden must be in [1,128] and nn_layers in [1,8]");
    // the FP32 kernels (rollout, RHS, RHS-VJP: reachable from every mode) keep the weight image and one activation
    // column per thread in shared memory
    if (hode::simt_min_smem_bytes(c->nn_hidden, c->nn_layers) > 227 * 1024)
      return fail(HODE_E_UNSUPPORTED, "network too large: its weight image and activation columns exceed the 227 KB of "
                                      "shared memory per CTA (e.g. 128 x 4 needs 239 KB)");
    if (uses_tensor_cores(c) && (c->nn_hidden != 64 || c->nn_layers > 5))
      return fail(HODE_E_UNSUPPORTED, "tensor-core MLP requires nn_hidden == 64 and nn_layers <= 5 (the weight image of "
                                      "deeper networks and the stage store do not fit the 227 KB of shared memory)");
    if ((c->mlp == HODE_MLP_TF32X2BF16 || c->mlp == HODE_MLP_F16BF16X2) && c->nn_layers > 4)
      return fail(HODE_E_UNSUPPORTED, "HODE_MLP_TF32X2BF16 / HODE_MLP_F16BF16X2 (three tiles per SM) require nn_layers <= 4");
  }
  if (c->solver != HODE_SOLVER_RK4 && (!(c->rtol > 0) || !(c->atol >= 0)))
    return fail(HODE_E_SIZE, "rtol must be > 0 and atol >= 0");
  if (c->kink_mode != HODE_KINK_SCIPY && c->kink_mode != HODE_KINK_CLIP)
    return fail(HODE_E_UNSUPPORTED, "unknown kink_mode");
  return 0;
}

hode::RolloutArgs make_args(const hode_cfg* c, const float* y0, const float* t_obs,
                            const float* u_meal, const float* u_tvns, const float* u_gd,
                            const float* theta, const float* W) {
  hode::RolloutArgs A{};
  A.y0 = y0; A.t_obs = t_obs;
  A.u[0] = u_meal; A.u[1] = u_tvns; A.u[2] = u_gd;
  A.theta = theta; A.W = W;
  A.B = c->n_traj; A.T = c->n_obs; A.S = c->n_samples;
  A.t_per_traj = c->t_per_traj;
  for (int ch = 0; ch < 3; ++ch) A.in_mode[ch] = c->in_mode[ch];
  A.H = c->nn_hidden; A.L = c->nn_layers;
  A.P = c->mlp != HODE_MLP_NONE ? (int)hode_mlp_param_count(c->nn_hidden, c->nn_layers) : 0;
  A.solver = c->solver; A.n_substeps = c->n_substeps; A.max_steps = c->max_steps;
  A.kink_mode = c->kink_mode;
  A.rhs_part = c->rhs_part;
  A.rtol = (float)c->rtol; A.atol = (float)c->atol;
  return A;
}

int check_inputs(const hode_cfg* c, const float* u_meal, const float* u_tvns, const float* u_gd,
                 const float* W) {
  const float* u[3] = {u_meal, u_tvns, u_gd};
  for (int ch = 0; ch < 3; ++ch)
    if (c->in_mode[ch] != HODE_IN_ABSENT && !u[ch])
      return fail(HODE_E_NULL, "input channel declared present but its pointer is NULL");
  if (c->mlp != HODE_MLP_NONE && !W) return fail(HODE_E_NULL, "W is NULL but cfg.mlp != NONE");
  return 0;
}

}  // namespace

extern "C" {

int hode_version(void) { return HODE_ABI_VERSION; }

int64_t hode_launch_count(void) { return (int64_t)hode::g_launches.load(std::memory_order_relaxed); }

const char* hode_last_error_string(void) { return g_err; }

int64_t hode_mlp_param_count(int32_t H, int32_t L) {
  if (H < 1 || L < 1) return 0;
  int64_t n = (int64_t)HODE_NN_IN * H + H;
  for (int l = 1; l < L; ++l) n += (int64_t)H * H + H;
  n += (int64_t)H * HODE_N_STATE + HODE_N_STATE;
  return n;
}

static hode::AdjPlan bwd_plan(const hode_cfg* c) {
  const bool has_nn = c->mlp != HODE_MLP_NONE;
  const int P = has_nn ? (int)hode_mlp_param_count(c->nn_hidden, c->nn_layers) : 0;
  return hode::adj_plan(c->n_traj, c->n_samples, has_nn ? c->nn_hidden : 0, has_nn ? c->nn_layers : 0,
                        P, has_nn, c->n_obs, c->t_per_traj, 7);
}

static bool bwd_uses_tensor_cores(const hode_cfg* c) { return adjoint_on_tensor_cores(c); }

int32_t hode_step_record_floats(const hode_cfg* cfg) {
  if (validate(cfg) || !cfg->save_steps) return 0;
  return rec_floats(cfg);
}

int32_t hode_step_record_capacity(const hode_cfg* cfg) {
  if (validate(cfg) || !cfg->save_steps) return 0;
  return max_saved_steps(cfg);
}

static hode::AdjTcPlan bwd_tc_plan(const hode_cfg* c) {
  return hode::adj_tc_plan(c->n_traj, c->n_samples, c->nn_layers,
                           (int)hode_mlp_param_count(c->nn_hidden, c->nn_layers), c->n_obs, c->t_per_traj);
}

int hode_workspace_bytes(const hode_cfg* cfg, size_t* fwd_bytes, size_t* bwd_bytes) {
  int rc = validate(cfg);
  if (rc) return rc;
  const Workspace w = fwd_workspace(cfg);
  if (fwd_bytes) *fwd_bytes = w.total;
  // gradient scratch (per-CTA partial gradients + activation stash [+ weight images of the
  // tensor-core adjoint]); also covers hode_rhs_vjp
  if (bwd_bytes) {
    size_t n = hode::adj_workspace_bytes(bwd_plan(cfg));
    if (bwd_uses_tensor_cores(cfg)) {
      const size_t m = hode::adj_tc_workspace_bytes(bwd_tc_plan(cfg));
      if (m > n) n = m;
    }
    *bwd_bytes = n;
  }
  return 0;
}

static int check_opts(const hode_cfg* cfg, const hode_fwd_opts* o) {
  if (!o) return 0;
  if (o->struct_bytes != (int32_t)sizeof(hode_fwd_opts))
    return fail(HODE_E_SIZE, "opts.struct_bytes != sizeof(hode_fwd_opts): header/library mismatch");
  if (o->theta_per_traj && cfg->n_samples != 1)
    return fail(HODE_E_UNSUPPORTED, "theta_per_traj needs n_samples == 1 (theta is [B,17], W is shared)");
  if (o->theta_per_traj && cfg->save_steps)
    return fail(HODE_E_UNSUPPORTED, "theta_per_traj is forward-only (no gradient with respect to a per-trajectory theta)");
  if (o->out_state_mask >> HODE_N_STATE) return fail(HODE_E_SHAPE, "out_state_mask has bits beyond the 6 states");
  return 0;
}

int hode_rollout_fwd(const hode_cfg* cfg, const float* y0, const float* t_obs,
                     const float* u_meal, const float* u_tvns, const float* u_gd,
                     const float* theta, const float* W, float* traj, int32_t* status,
                     int32_t* counters, void* workspace, size_t workspace_bytes, void* stream) {
  return hode_rollout_fwd_ex(cfg, nullptr, y0, t_obs, u_meal, u_tvns, u_gd, theta, W, traj, status, counters, workspace,
                             workspace_bytes, stream);
}

int hode_rollout_fwd_ex(const hode_cfg* cfg, const hode_fwd_opts* opts, const float* y0, const float* t_obs,
                        const float* u_meal, const float* u_tvns, const float* u_gd,
                        const float* theta, const float* W, float* traj, int32_t* status,
                        int32_t* counters, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  rc = check_opts(cfg, opts);
  if (rc) return rc;
  if (!y0 || !t_obs || !theta || !traj) return fail(HODE_E_NULL, "y0/t_obs/theta/traj is NULL");
  rc = check_inputs(cfg, u_meal, u_tvns, u_gd, W);
  if (rc) return rc;
  if (cfg->n_traj == 0) return 0;
  hode::RolloutArgs A = make_args(cfg, y0, t_obs, u_meal, u_tvns, u_gd, theta, W);
  A.traj = traj; A.status = status; A.counters = counters;
  if (opts) {
    A.theta_per_traj = opts->theta_per_traj ? 1 : 0;
    A.order = opts->order;
    if (opts->out_state_mask && opts->out_state_mask != 0x3Fu) {
      A.out_mask = opts->out_state_mask;
      A.out_nc = __builtin_popcount(opts->out_state_mask);
    }
  }
  const Workspace w = fwd_workspace(cfg);
  if (w.total > 0 && (!workspace || workspace_bytes < w.total))
    return fail(HODE_E_WORKSPACE, "workspace missing or smaller than hode_workspace_bytes()");
  if (cfg->save_steps) {
    char* base = (char*)workspace;
    A.save_n = (int32_t*)(base + w.off_n);
    A.save_rec = (float*)(base + w.off_rec);
    A.save_k1 = uses_tensor_cores(cfg) ? 1 : 0;   // the tensor-core rollout records its stage derivatives
    A.rec_floats = rec_floats(cfg);
    A.max_saved = w.max_saved;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (cfg->mlp == HODE_MLP_NONE || cfg->mlp == HODE_MLP_FP32) {
    e = hode::launch_rollout_simt(A, cfg->mlp, st);
  } else {
    e = hode::launch_rollout_tc(A, cfg->mlp, (char*)workspace + w.off_tc, st);
  }
  if (e != cudaSuccess) return cuda_fail(e, "hode_rollout_fwd launch");
  return 0;
}

int hode_rollout_bwd(const hode_cfg* cfg, const float* y0, const float* t_obs, const float* u_meal,
                     const float* u_tvns, const float* u_gd, const float* theta, const float* W,
                     const float* grad_traj, float* grad_y0, float* grad_theta, float* grad_W,
                     const void* fwd_workspace_ptr, size_t fwd_workspace_bytes, void* bwd_workspace,
                     size_t bwd_workspace_bytes, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!t_obs || !theta || !grad_traj) return fail(HODE_E_NULL, "t_obs/theta/grad_traj is NULL");
  (void)y0;
  rc = check_inputs(cfg, u_meal, u_tvns, u_gd, W);
  if (rc) return rc;
  if (!cfg->save_steps)
    return fail(HODE_E_UNSUPPORTED, "hode_rollout_bwd needs the forward pass to run with save_steps = 1");
  const Workspace w = fwd_workspace(cfg);
  if (!fwd_workspace_ptr || fwd_workspace_bytes < w.total)
    return fail(HODE_E_WORKSPACE, "forward workspace missing or smaller than hode_workspace_bytes()");
  const bool tc_adj = bwd_uses_tensor_cores(cfg);
  const hode::AdjPlan plan = bwd_plan(cfg);
  if (!tc_adj && plan.smem > 227 * 1024)
    return fail(HODE_E_UNSUPPORTED, "network too large for the shared-memory gradient accumulators");
  size_t need = 0;
  hode_workspace_bytes(cfg, nullptr, &need);
  if (!bwd_workspace || bwd_workspace_bytes < need)
    return fail(HODE_E_WORKSPACE, "backward workspace missing or smaller than hode_workspace_bytes()");
  if (cfg->n_traj == 0) return 0;
  hode::RolloutArgs A = make_args(cfg, y0, t_obs, u_meal, u_tvns, u_gd, theta, W);
  char* base = (char*)fwd_workspace_ptr;
  A.save_n = (int32_t*)(base + w.off_n);
  A.save_rec = (float*)(base + w.off_rec);
  A.save_k1 = uses_tensor_cores(cfg) ? 1 : 0;
  A.rec_floats = rec_floats(cfg);
  A.max_saved = w.max_saved;
  // forward on the tensor cores -> adjoint on the tensor cores (3xTF32); FP32 forward -> FP32 adjoint
  cudaError_t e = tc_adj ? hode::launch_rollout_bwd_tc(A, cfg->mlp, grad_traj, grad_y0, grad_theta, grad_W, bwd_workspace,
                                                       (cudaStream_t)stream)
                         : hode::launch_rollout_bwd(A, cfg->mlp != HODE_MLP_NONE, grad_traj, grad_y0, grad_theta,
                                                    grad_W, bwd_workspace, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "hode_rollout_bwd launch");
  return 0;
}

int hode_generate_4gi(int32_t n_subjects, int32_t n_obs, double interval_hours, int32_t patient_type, double rtol,
                      double atol, const float* baselines, const float* meal_rate, float* out, int32_t* status,
                      void* stream) {
  if (n_subjects < 0 || n_obs < 1) return fail(HODE_E_SIZE, "n_subjects < 0 or n_obs < 1");
  if (!(interval_hours > 0.0)) return fail(HODE_E_SIZE, "interval_hours must be positive");
  if (patient_type != 0 && patient_type != 1) return fail(HODE_E_UNSUPPORTED, "patient_type must be 0 (T2DM) or 1 (HV)");
  if (n_subjects == 0) return 0;
  if (!baselines || !out) return fail(HODE_E_NULL, "baselines/out is NULL");
  cudaError_t e = hode::launch_gen4gi(n_subjects, n_obs, interval_hours, patient_type, rtol > 0.0 ? rtol : 1e-9,
                                      atol > 0.0 ? atol : 1e-12, baselines, meal_rate, out, status,
                                      (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "hode_generate_4gi launch");
  return 0;
}

// ---- fused data-loss step ---------------------------------------------------------------------------
namespace {
constexpr int MSE_BLOCK = 256, MSE_ELEMS_PER_BLOCK = 256 * 16;

// grad[s][i] = scale * (traj[s][i] - obs[i]);  partial[s][block] = sum of squared residuals of the block's
// elements, reduced in a fixed tree order
__global__ void __launch_bounds__(MSE_BLOCK) mse_grad_kernel(const float* __restrict__ traj, const float* __restrict__ obs,
                                                             float* __restrict__ grad, float* __restrict__ partial,
                                                             long n_per_set, float scale) {
  __shared__ float red[MSE_BLOCK];
  const int s = blockIdx.y;
  const long base = (long)blockIdx.x * MSE_ELEMS_PER_BLOCK;
  const float* tr = traj + (size_t)s * n_per_set;
  float* g = grad + (size_t)s * n_per_set;
  float acc = 0.f;
#pragma unroll 4
  for (int j = 0; j < MSE_ELEMS_PER_BLOCK / MSE_BLOCK; ++j) {
    const long i = base + (long)j * MSE_BLOCK + threadIdx.x;
    if (i < n_per_set) {
      const float d = tr[i] - obs[i];
      g[i] = scale * d;
      acc = fmaf(d, d, acc);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = MSE_BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)s * gridDim.x + blockIdx.x] = red[0];
}

// loss[s] = inv_n * sum of the set's block partials, accumulated in double in block order
__global__ void __launch_bounds__(MSE_BLOCK) mse_final_kernel(const float* __restrict__ partial, int n_blocks, double inv_n,
                                                              float* __restrict__ loss) {
  __shared__ double red[MSE_BLOCK];
  const int s = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_blocks; i += MSE_BLOCK) acc += (double)partial[(size_t)s * n_blocks + i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = MSE_BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[s] = (float)(red[0] * inv_n);
}
}  // namespace

int hode_loss_fused_fwd_bwd(const hode_cfg* cfg, const float* y0, const float* t_obs, const float* u_meal,
                            const float* u_tvns, const float* u_gd, const float* theta, const float* W,
                            const float* obs, float* traj, int32_t* status, int32_t* counters, float* loss,
                            float* grad_traj_scratch, float* grad_y0, float* grad_theta, float* grad_W,
                            void* fwd_workspace, size_t fwd_workspace_bytes, void* bwd_workspace,
                            size_t bwd_workspace_bytes, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!obs || !loss || !grad_traj_scratch) return fail(HODE_E_NULL, "obs/loss/grad_traj_scratch is NULL");
  if (!cfg->save_steps)
    return fail(HODE_E_UNSUPPORTED, "hode_loss_fused_fwd_bwd needs cfg.save_steps = 1");
  const int S = cfg->n_samples > 0 ? cfg->n_samples : 1;
  const long n_per_set = (long)cfg->n_traj * cfg->n_obs * HODE_N_STATE;
  const int n_blocks = (int)((n_per_set + MSE_ELEMS_PER_BLOCK - 1) / MSE_ELEMS_PER_BLOCK);
  size_t need_bwd = 0;
  hode_workspace_bytes(cfg, nullptr, &need_bwd);
  // the block partials live at the head of the backward workspace: consumed (stream order) before
  // hode_rollout_bwd overwrites it
  if (!bwd_workspace || bwd_workspace_bytes < need_bwd || bwd_workspace_bytes < (size_t)S * n_blocks * sizeof(float))
    return fail(HODE_E_WORKSPACE, "backward workspace missing or too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->n_traj == 0) {
    // no trajectories: the loss and every gradient are zero (grad_y0 has no rows)
    cudaError_t e = cudaMemsetAsync(loss, 0, (size_t)S * sizeof(float), st);
    if (e == cudaSuccess && grad_theta) e = cudaMemsetAsync(grad_theta, 0, (size_t)S * HODE_N_THETA * sizeof(float), st);
    if (e == cudaSuccess && grad_W && cfg->mlp != HODE_MLP_NONE)
      e = cudaMemsetAsync(grad_W, 0, (size_t)S * (size_t)hode_mlp_param_count(cfg->nn_hidden, cfg->nn_layers) * sizeof(float), st);
    if (e != cudaSuccess) return cuda_fail(e, "hode_loss_fused_fwd_bwd memset");
    return 0;
  }
  rc = hode_rollout_fwd(cfg, y0, t_obs, u_meal, u_tvns, u_gd, theta, W, traj, status, counters, fwd_workspace,
                        fwd_workspace_bytes, stream);
  if (rc) return rc;
  float* partial = reinterpret_cast<float*>(bwd_workspace);
  count_launch();
  mse_grad_kernel<<<dim3((unsigned)n_blocks, (unsigned)S), MSE_BLOCK, 0, st>>>(traj, obs, grad_traj_scratch, partial,
                                                                             n_per_set, (float)(2.0 / (double)n_per_set));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "hode_loss_fused_fwd_bwd residual launch");
  count_launch();
  mse_final_kernel<<<S, MSE_BLOCK, 0, st>>>(partial, n_blocks, 1.0 / (double)n_per_set, loss);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "hode_loss_fused_fwd_bwd reduction launch");
  return hode_rollout_bwd(cfg, y0, t_obs, u_meal, u_tvns, u_gd, theta, W, grad_traj_scratch, grad_y0, grad_theta,
                          grad_W, fwd_workspace, fwd_workspace_bytes, bwd_workspace, bwd_workspace_bytes, stream);
}

int hode_rhs_vjp(const hode_cfg* cfg, const float* t, const float* state, const float* u_meal,
                 const float* u_tvns, const float* u_gd, const float* theta, const float* W,
                 const float* grad_out, float* grad_state, float* grad_theta, float* grad_W,
                 void* bwd_workspace, size_t bwd_workspace_bytes, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!t || !state || !theta || !grad_out) return fail(HODE_E_NULL, "t/state/theta/grad_out is NULL");
  rc = check_inputs(cfg, u_meal, u_tvns, u_gd, W);
  if (rc) return rc;
  if (cfg->n_samples != 1) return fail(HODE_E_UNSUPPORTED, "hode_rhs_vjp takes one parameter set");
  const hode::AdjPlan plan = bwd_plan(cfg);
  if (plan.smem > 227 * 1024)
    return fail(HODE_E_UNSUPPORTED, "network too large for the shared-memory gradient accumulators");
  if (!bwd_workspace || bwd_workspace_bytes < hode::adj_workspace_bytes(plan))
    return fail(HODE_E_WORKSPACE, "backward workspace missing or smaller than hode_workspace_bytes()");
  hode::RolloutArgs A = make_args(cfg, state, t, u_meal, u_tvns, u_gd, theta, W);
  if (cfg->n_traj == 0) {
    // no rows: the gradients are zero
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_theta) cudaMemsetAsync(grad_theta, 0, HODE_N_THETA * sizeof(float), st);
    if (grad_W && A.P) cudaMemsetAsync(grad_W, 0, (size_t)A.P * sizeof(float), st);
    return 0;
  }
  cudaError_t e = hode::launch_rhs_vjp(A, cfg->mlp != HODE_MLP_NONE, t, state, grad_out, grad_state,
                                       grad_theta, grad_W, bwd_workspace, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "hode_rhs_vjp launch");
  return 0;
}

int hode_vi_predictive(const hode_cfg* cfg, const float* y0, const float* t_obs, const float* u_meal,
                       const float* u_tvns, const float* u_gd, const float* theta, const float* W,
                       float* mean, float* std_out, int32_t* status, int32_t* counters,
                       void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!y0 || !t_obs || !theta || !mean || !std_out)
    return fail(HODE_E_NULL, "y0/t_obs/theta/mean/std is NULL");
  rc = check_inputs(cfg, u_meal, u_tvns, u_gd, W);
  if (rc) return rc;
  if (cfg->save_steps) return fail(HODE_E_UNSUPPORTED, "save_steps is not available in the fused sweep");
  if (cfg->n_traj == 0) return 0;
  hode::RolloutArgs A = make_args(cfg, y0, t_obs, u_meal, u_tvns, u_gd, theta, W);
  A.traj = nullptr; A.status = status; A.counters = counters;
  A.vi_mean = mean; A.vi_m2 = std_out;
  const Workspace w = fwd_workspace(cfg);
  if (w.total > 0 && (!workspace || workspace_bytes < w.total))
    return fail(HODE_E_WORKSPACE, "workspace missing or smaller than hode_workspace_bytes()");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (cfg->mlp == HODE_MLP_NONE || cfg->mlp == HODE_MLP_FP32) {
    e = hode::launch_rollout_simt(A, cfg->mlp, st);
  } else {
    e = hode::launch_rollout_tc(A, cfg->mlp, (char*)workspace + w.off_tc, st);
  }
  if (e != cudaSuccess) return cuda_fail(e, "hode_vi_predictive launch");
  return 0;
}

int hode_rhs(const hode_cfg* cfg, const float* t, const float* state, const float* u_meal,
             const float* u_tvns, const float* u_gd, const float* theta, const float* W,
             float* out, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!t || !state || !theta || !out) return fail(HODE_E_NULL, "t/state/theta/out is NULL");
  rc = check_inputs(cfg, u_meal, u_tvns, u_gd, W);
  if (rc) return rc;
  if (cfg->n_traj == 0) return 0;
  hode::RolloutArgs A = make_args(cfg, state, t, u_meal, u_tvns, u_gd, theta, W);
  const int mode = cfg->mlp == HODE_MLP_NONE ? HODE_MLP_NONE : HODE_MLP_FP32;
  cudaError_t e = hode::launch_rhs(A, mode, t, state, out, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "hode_rhs launch");
  return 0;
}

namespace {

// Keep the host entries' staging memory in the stream-ordered pool between calls: without a release
// threshold the pool hands it back to the driver at every synchronisation and the next call pays a
// fresh (hundreds of MB) allocation.
void tune_default_pool() {
  static thread_local int pool_tuned_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (pool_tuned_dev != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool_tuned_dev = dev;
  }
}

// Host entry, tensor-core path, one parameter set: ONE persistent launch over all trajectories (the
// lane-refill queue keeps every SM busy to the end; chunked launches lose ~25 % to their tails),
// with the device->host copy of the results overlapped block by block: the kernel counts finished
// trajectories per block of STREAM_BLOCK and raises a flag in host-mapped memory when a block is
// complete; this thread polls the flags and queues that block's copy on a second stream.
constexpr int STREAM_MAX_BLOCKS = 1024;
// trajectories per result block (HODE_STREAM_BLOCK overrides it: measurement knob)
static int stream_block() {
  static int v = 0;
  if (!v) {
    const char* s = getenv("HODE_STREAM_BLOCK");
    int x = s ? atoi(s) : 0;
    v = (x >= 256 && x <= (1 << 20)) ? x : 8192;
  }
  return v;
}
#define STREAM_BLOCK (stream_block())
// ... and the host->device copy of the inputs overlapped the other way round: the kernel starts as soon as the first
// IN_FIRST trajectories (more than one wave of lanes: 148 SMs x 384) are resident; the rest follows in blocks of
// IN_BLOCK on a second copy stream, each block followed by a 4-byte copy that raises its ready flag in device memory;
// a lane that draws a trajectory of a block still in flight waits for the flag (hode_rollout_tc.cu).  Every block's
// copies run 8 rows into the next block: a 128-byte line that straddles a block boundary is then already complete
// when an SM first caches it.
constexpr int IN_FIRST = 65536;
constexpr int IN_BLOCK = 32768;
constexpr int IN_OVERLAP_ROWS = 8;

// Launch-order hint of the host entry (opts.prev_counters): sort key of trajectory b = (its input block, descending
// attempt count of the previous pass); a stable radix sort of (key, b) gives the order — longest first within blocks
// of IN_BLOCK trajectories, blocks in index order, so that inputs can still arrive, and result blocks complete, block
// by block.
__global__ void order_keys_kernel(const int32_t* __restrict__ counters, int B, int block, uint32_t* __restrict__ keys,
                                  int32_t* __restrict__ vals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long att = (long)counters[b] + (long)counters[B + b];
  att = att < 0 ? 0 : (att > 0xFFFFF ? 0xFFFFF : att);
  keys[b] = ((uint32_t)(b / block) << 20) | (uint32_t)(0xFFFFF - att);
  vals[b] = b;
}

int rollout_fwd_host_streamed(const hode_cfg* cfg, const hode_fwd_opts* opts, const float* y0_h, const float* t_obs_h,
                              const float* const uh[3], const float* theta_h, const float* W_h,
                              float* traj_h, int32_t* status_h, int32_t* counters_h, cudaStream_t st) {
  const size_t B = cfg->n_traj, T = cfg->n_obs;
  const uint32_t mask = (opts && opts->out_state_mask && opts->out_state_mask != 0x3Fu) ? opts->out_state_mask : 0u;
  const size_t nc = mask ? (size_t)__builtin_popcount(mask) : 6;   // state columns copied back
  const size_t n_theta = (opts && opts->theta_per_traj) ? B : 1;
  const size_t P = hode_mlp_param_count(cfg->nn_hidden, cfg->nn_layers);
  const int n_blk = (int)((B + STREAM_BLOCK - 1) / STREAM_BLOCK);
  size_t ub[3];
  for (int ch = 0; ch < 3; ++ch)
    ub[ch] = cfg->in_mode[ch] == HODE_IN_SERIES ? B * T * 4 : cfg->in_mode[ch] == HODE_IN_CONST ? B * 4 : 0;
  const Workspace wsp = fwd_workspace(cfg);
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_y0 = carve(B * 24), o_t = carve((cfg->t_per_traj ? B * T : T) * 4);
  size_t o_u[3];
  for (int ch = 0; ch < 3; ++ch) o_u[ch] = carve(ub[ch]);
  const bool gated = B >= (size_t)IN_FIRST + IN_BLOCK;   // inputs of trajectories >= IN_FIRST follow the kernel launch
  const size_t n_first = gated ? (size_t)IN_FIRST : B;
  const int n_in_blk = gated ? (int)((B - n_first + IN_BLOCK - 1) / IN_BLOCK) : 0;
  const size_t o_th = carve(n_theta * 17 * 4), o_W = carve(P * 4), o_traj = carve(B * T * nc * 4), o_st = carve(B * 4),
               o_cn = carve(2 * B * 4), o_done = carve((size_t)n_blk * 4), o_ready = carve((size_t)(n_in_blk + 1) * 4),
               o_ws = carve(wsp.total);
  // launch-order hint: previous counters, sort keys / values (double-buffered) and cub's scratch
  const int32_t* prev_h = opts ? opts->prev_counters : nullptr;
  size_t sort_tmp = 0;
  if (prev_h) cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                              (const int32_t*)nullptr, (int32_t*)nullptr, (int)B, 0, 32, st);
  const size_t o_pc = carve(prev_h ? 2 * B * 4 : 0), o_k0 = carve(prev_h ? B * 4 : 0), o_k1 = carve(prev_h ? B * 4 : 0),
               o_v0 = carve(prev_h ? B * 4 : 0), o_ord = carve(prev_h ? B * 4 : 0), o_tmp = carve(sort_tmp);

  static thread_local int* flags_h = nullptr;     // host-mapped completion flags (allocated once); [STREAM_MAX_BLOCKS] = 1
  static thread_local cudaStream_t copy_stream = nullptr, in_stream = nullptr;
  cudaError_t e = cudaSuccess;
  if (!flags_h) {
    e = cudaHostAlloc((void**)&flags_h, (STREAM_MAX_BLOCKS + 1) * sizeof(int), cudaHostAllocMapped);
    if (e != cudaSuccess) { flags_h = nullptr; return cuda_fail(e, "cudaHostAlloc(flags)"); }
    flags_h[STREAM_MAX_BLOCKS] = 1;   // source of the ready-flag copies
  }
  if (!copy_stream) {
    e = cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { copy_stream = nullptr; return cuda_fail(e, "cudaStreamCreate"); }
  }
  if (!in_stream) {
    e = cudaStreamCreateWithFlags(&in_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { in_stream = nullptr; return cuda_fail(e, "cudaStreamCreate"); }
  }
  int* flags_d = nullptr;
  e = cudaHostGetDevicePointer((void**)&flags_d, flags_h, 0);
  if (e != cudaSuccess) return cuda_fail(e, "cudaHostGetDevicePointer");
  for (int i = 0; i < n_blk; ++i) flags_h[i] = 0;

  char* d = nullptr;
  e = cudaMallocAsync((void**)&d, off, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  int lib_rc = 0;
  int copied = 0;
  cudaEvent_t ev_alloc = nullptr;
  // per-trajectory inputs of rows [lo, hi) (+ the overlap) on stream `cs`
  auto copy_rows = [&](size_t lo, size_t hi, cudaStream_t cs) -> cudaError_t {
    const size_t hi2 = hi + IN_OVERLAP_ROWS < B ? hi + IN_OVERLAP_ROWS : B;
    cudaError_t ce = cudaMemcpyAsync(d + o_y0 + lo * 24, y0_h + lo * 6, (hi2 - lo) * 24, cudaMemcpyHostToDevice, cs);
    if (ce == cudaSuccess && cfg->t_per_traj)
      ce = cudaMemcpyAsync(d + o_t + lo * T * 4, t_obs_h + lo * T, (hi2 - lo) * T * 4, cudaMemcpyHostToDevice, cs);
    for (int ch = 0; ch < 3 && ce == cudaSuccess; ++ch) {
      const size_t row = ub[ch] / B;   // bytes per trajectory of this channel (0: absent)
      if (row) ce = cudaMemcpyAsync(d + o_u[ch] + lo * row, (const char*)uh[ch] + lo * row, (hi2 - lo) * row, cudaMemcpyHostToDevice, cs);
    }
    return ce;
  };
#define CK(call) if ((e = (call)) != cudaSuccess) goto done
  if (prev_h) {   // first, so that the sort runs under the input copies that follow
    CK(cudaMemcpyAsync(d + o_pc, prev_h, 2 * B * 4, cudaMemcpyHostToDevice, st));
    hode::count_launch();
    order_keys_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>((const int32_t*)(d + o_pc), (int)B, IN_BLOCK,
                                                                (uint32_t*)(d + o_k0), (int32_t*)(d + o_v0));
    CK(cudaGetLastError());
    CK(cub::DeviceRadixSort::SortPairs(d + o_tmp, sort_tmp, (const uint32_t*)(d + o_k0), (uint32_t*)(d + o_k1),
                                       (const int32_t*)(d + o_v0), (int32_t*)(d + o_ord), (int)B, 0, 32, st));
  }
  CK(copy_rows(0, n_first, st));
  if (!cfg->t_per_traj) CK(cudaMemcpyAsync(d + o_t, t_obs_h, T * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_th, theta_h, n_theta * 17 * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_W, W_h, P * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d + o_done, 0, (size_t)n_blk * 4 , st));
  CK(cudaMemsetAsync(d + o_ready, 0, (size_t)(n_in_blk + 1) * 4, st));
  if (gated) {
    // the staging buffer is a stream-ordered allocation of `st`: the second copy stream may touch it (and the cleared
    // flags) only behind this event
    CK(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
    CK(cudaEventRecord(ev_alloc, st));
    CK(cudaStreamWaitEvent(in_stream, ev_alloc, 0));
  }
  {
    hode::RolloutArgs A = make_args(cfg, (float*)(d + o_y0), (float*)(d + o_t),
                                    ub[0] ? (float*)(d + o_u[0]) : nullptr, ub[1] ? (float*)(d + o_u[1]) : nullptr,
                                    ub[2] ? (float*)(d + o_u[2]) : nullptr, (float*)(d + o_th), (float*)(d + o_W));
    A.traj = (float*)(d + o_traj);
    A.status = (int32_t*)(d + o_st);
    A.counters = (int32_t*)(d + o_cn);
    A.done_count = (int*)(d + o_done);
    A.done_flag = flags_d;
    A.done_block = STREAM_BLOCK;
    A.out_mask = mask;
    A.out_nc = mask ? (int)nc : 0;
    A.theta_per_traj = n_theta > 1 ? 1 : 0;
    if (prev_h) A.order = (const int32_t*)(d + o_ord);
    if (gated) {
      A.in_ready = (const int*)(d + o_ready);
      A.in_ready_first = (int)n_first;
      A.in_ready_block = IN_BLOCK;
    }
    CK(hode::launch_rollout_tc(A, cfg->mlp, d + o_ws + wsp.off_tc, st));
  }
  // the rest of the inputs, block by block, behind the running kernel
  for (int k = 0; k < n_in_blk; ++k) {
    const size_t lo = n_first + (size_t)k * IN_BLOCK, hi = lo + IN_BLOCK < B ? lo + IN_BLOCK : B;
    CK(copy_rows(lo, hi, in_stream));
    CK(cudaMemcpyAsync(d + o_ready + (size_t)k * 4, flags_h + STREAM_MAX_BLOCKS, 4, cudaMemcpyHostToDevice, in_stream));
  }
  // copy every block back as soon as the kernel reports it complete
  {
    volatile int* vf = flags_h;
    cudaEvent_t ev_kernel;
    CK(cudaEventCreateWithFlags(&ev_kernel, cudaEventDisableTiming));
    e = cudaEventRecord(ev_kernel, st);
    // blocks complete out of order (a block is as late as its longest trajectory): every sweep over the flags queues the
    // copies of all blocks that have completed since the last one; flags_h[i] = 2 marks "copy queued"
    const size_t blk = (size_t)STREAM_BLOCK;
    bool kernel_over = false;
    while (e == cudaSuccess && copied < n_blk && !kernel_over) {
      kernel_over = cudaEventQuery(ev_kernel) != cudaErrorNotReady;   // (checked BEFORE the sweep: flags raised by then are seen)
      for (int i = 0; i < n_blk && e == cudaSuccess; ++i) {
        if (vf[i] != 1) continue;
        const size_t lo = (size_t)i * blk, hi = lo + blk < B ? lo + blk : B;
        e = cudaMemcpyAsync(traj_h + lo * T * nc, d + o_traj + lo * T * nc * 4, (hi - lo) * T * nc * 4, cudaMemcpyDeviceToHost,
                            copy_stream);
        vf[i] = 2;
        ++copied;
      }
    }
    cudaEventDestroy(ev_kernel);
    if (e != cudaSuccess) goto done;
  }
  CK(cudaStreamSynchronize(st));
  // whatever is left (normally nothing; everything after a launch failure)
  for (int i = 0; i < n_blk; ++i) {
    if (flags_h[i] == 2) continue;
    const size_t lo = (size_t)i * STREAM_BLOCK, hi = lo + STREAM_BLOCK < B ? lo + STREAM_BLOCK : B;
    CK(cudaMemcpyAsync(traj_h + lo * T * nc, d + o_traj + lo * T * nc * 4, (hi - lo) * T * nc * 4, cudaMemcpyDeviceToHost,
                       copy_stream));
  }
  if (status_h) CK(cudaMemcpyAsync(status_h, d + o_st, B * 4, cudaMemcpyDeviceToHost, st));
  if (counters_h) CK(cudaMemcpyAsync(counters_h, d + o_cn, 2 * B * 4, cudaMemcpyDeviceToHost, st));
#undef CK
done:
  {
    cudaError_t e1 = cudaStreamSynchronize(in_stream);
    cudaError_t e2 = cudaStreamSynchronize(copy_stream);
    cudaError_t e3 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
  }
  if (ev_alloc) cudaEventDestroy(ev_alloc);
  cudaFreeAsync(d, st);
  cudaStreamSynchronize(st);
  if (lib_rc) return lib_rc;
  if (e != cudaSuccess) return cuda_fail(e, "hode_rollout_fwd_host (streamed)");
  return 0;
}

}  // namespace

int hode_rollout_fwd_host(const hode_cfg* cfg, const float* y0_h, const float* t_obs_h,
                          const float* u_meal_h, const float* u_tvns_h, const float* u_gd_h,
                          const float* theta_h, const float* W_h, float* traj_h,
                          int32_t* status_h, int32_t* counters_h, void* stream) {
  return hode_rollout_fwd_host_ex(cfg, nullptr, y0_h, t_obs_h, u_meal_h, u_tvns_h, u_gd_h, theta_h, W_h, traj_h, status_h,
                                  counters_h, stream);
}

int hode_rollout_fwd_host_ex(const hode_cfg* cfg, const hode_fwd_opts* opts, const float* y0_h, const float* t_obs_h,
                             const float* u_meal_h, const float* u_tvns_h, const float* u_gd_h,
                             const float* theta_h, const float* W_h, float* traj_h,
                             int32_t* status_h, int32_t* counters_h, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  rc = check_opts(cfg, opts);
  if (rc) return rc;
  const uint32_t mask = (opts && opts->out_state_mask && opts->out_state_mask != 0x3Fu) ? opts->out_state_mask : 0u;
  const size_t nc = mask ? (size_t)__builtin_popcount(mask) : 6;
  const bool th_per_traj = opts && opts->theta_per_traj;
  hode_fwd_opts dev_opts{};
  dev_opts.struct_bytes = (int32_t)sizeof(hode_fwd_opts);
  dev_opts.theta_per_traj = th_per_traj ? 1 : 0;
  dev_opts.out_state_mask = mask;
  if (!y0_h || !t_obs_h || !theta_h || !traj_h)
    return fail(HODE_E_NULL, "y0/t_obs/theta/traj host pointer is NULL");
  rc = check_inputs(cfg, u_meal_h, u_tvns_h, u_gd_h, W_h);
  if (rc) return rc;
  if (cfg->save_steps) return fail(HODE_E_UNSUPPORTED, "save_steps is not available on the host entry");
  if (cfg->n_traj == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = cfg->n_traj, T = cfg->n_obs, S = cfg->n_samples;
  const size_t P = cfg->mlp != HODE_MLP_NONE ? hode_mlp_param_count(cfg->nn_hidden, cfg->nn_layers) : 0;
  const float* uh[3] = {u_meal_h, u_tvns_h, u_gd_h};
  size_t urow[3];  // bytes per trajectory of each input channel
  for (int ch = 0; ch < 3; ++ch)
    urow[ch] = cfg->in_mode[ch] == HODE_IN_SERIES ? T * 4 : cfg->in_mode[ch] == HODE_IN_CONST ? 4 : 0;

  if (S == 1 && uses_tensor_cores(cfg) && B >= 2 * (size_t)STREAM_BLOCK &&
      B <= (size_t)STREAM_BLOCK * STREAM_MAX_BLOCKS) {
    tune_default_pool();
    return rollout_fwd_host_streamed(cfg, opts, y0_h, t_obs_h, uh, theta_h, W_h, traj_h, status_h, counters_h, st);
  }

  // Trajectory chunks are pipelined over a few streams: while chunk c integrates, chunk c+1 is
  // on its way in and chunk c-1 on its way out (the two copy engines and the SMs overlap, and
  // the next chunk's CTAs fill SMs that the previous chunk's stragglers have left).  Parameter
  // sweeps (S > 1) keep the [S,B,...] layouts contiguous and run as one chunk.
  constexpr int MAX_STREAMS = 3;
  int n_chunks = 1;
  if (S == 1) {
    n_chunks = (int)(B / 49152);
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > 8) n_chunks = 8;
  }
  const int n_streams = n_chunks < MAX_STREAMS ? n_chunks : MAX_STREAMS;
  const size_t rows_max = (B + n_chunks - 1) / n_chunks;
  hode_cfg sub = *cfg;
  sub.n_traj = (int32_t)rows_max;
  const Workspace wsp = fwd_workspace(&sub);

  const size_t sz_t_shared = cfg->t_per_traj ? 0 : T * 4, sz_th = (th_per_traj ? B : S) * 17 * 4, sz_W = S * P * 4;
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_t = carve(sz_t_shared), o_th = carve(sz_th), o_W = carve(sz_W);
  const size_t o_y0 = carve(B * 6 * 4), o_trow = carve(cfg->t_per_traj ? B * T * 4 : 0);
  size_t o_u[3];
  for (int ch = 0; ch < 3; ++ch) o_u[ch] = carve(B * urow[ch]);
  const size_t o_traj = carve(S * B * T * nc * 4), o_st = carve(S * B * 4), o_cn = carve(2 * S * B * 4);
  size_t o_ws[MAX_STREAMS];
  for (int i = 0; i < n_streams; ++i) o_ws[i] = carve(wsp.total);

  tune_default_pool();
  char* d = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&d, off ? off : 256, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
  cudaStream_t xs[MAX_STREAMS] = {st, nullptr, nullptr};
  cudaEvent_t ev_ready = nullptr, ev_done[MAX_STREAMS] = {nullptr, nullptr, nullptr};
  int lib_rc = 0;
#define CK(call) if ((e = (call)) != cudaSuccess) goto done
  for (int i = 1; i < n_streams; ++i) CK(cudaStreamCreateWithFlags(&xs[i], cudaStreamNonBlocking));
  // shared small inputs on the caller's stream, then fork
  if (sz_t_shared) CK(cudaMemcpyAsync(d + o_t, t_obs_h, sz_t_shared, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + o_th, theta_h, sz_th, cudaMemcpyHostToDevice, st));
  if (sz_W) CK(cudaMemcpyAsync(d + o_W, W_h, sz_W, cudaMemcpyHostToDevice, st));
  if (n_streams > 1) {
    CK(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
    CK(cudaEventRecord(ev_ready, st));
    for (int i = 1; i < n_streams; ++i) CK(cudaStreamWaitEvent(xs[i], ev_ready, 0));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const size_t lo = (size_t)c * rows_max, hi = lo + rows_max < B ? lo + rows_max : B;
    if (lo >= hi) break;
    const size_t rows = hi - lo;
    cudaStream_t cs = xs[c % n_streams];
    CK(cudaMemcpyAsync(d + o_y0 + lo * 24, y0_h + lo * 6, rows * 24, cudaMemcpyHostToDevice, cs));
    if (cfg->t_per_traj)
      CK(cudaMemcpyAsync(d + o_trow + lo * T * 4, t_obs_h + lo * T, rows * T * 4, cudaMemcpyHostToDevice, cs));
    for (int ch = 0; ch < 3; ++ch)
      if (urow[ch])
        CK(cudaMemcpyAsync(d + o_u[ch] + lo * urow[ch], (const char*)uh[ch] + lo * urow[ch], rows * urow[ch],
                           cudaMemcpyHostToDevice, cs));
    sub.n_traj = (int32_t)rows;
    // S == 1 when chunked: unit index == trajectory index, so every output is a contiguous slice
    lib_rc = hode_rollout_fwd_ex(
        &sub, &dev_opts, (float*)(d + o_y0 + lo * 24),
        cfg->t_per_traj ? (float*)(d + o_trow + lo * T * 4) : (float*)(d + o_t),
        urow[0] ? (float*)(d + o_u[0] + lo * urow[0]) : nullptr,
        urow[1] ? (float*)(d + o_u[1] + lo * urow[1]) : nullptr,
        urow[2] ? (float*)(d + o_u[2] + lo * urow[2]) : nullptr, (float*)(d + o_th + (th_per_traj ? lo * 17 * 4 : 0)),
        sz_W ? (float*)(d + o_W) : nullptr, (float*)(d + o_traj + lo * T * nc * 4), (int32_t*)(d + o_st + lo * 4),
        // chunk c keeps its [2, rows] counters at byte offset 2 * lo * 4 of the [2, B] buffer
        (int32_t*)(d + o_cn + lo * 8), wsp.total ? d + o_ws[c % n_streams] : nullptr,
        wsp.total, (void*)cs);
    if (lib_rc) goto done;
    CK(cudaMemcpyAsync(traj_h + lo * T * nc, d + o_traj + lo * T * nc * 4, rows * T * nc * 4 * (n_chunks == 1 ? S : 1),
                       cudaMemcpyDeviceToHost, cs));
    if (status_h)
      CK(cudaMemcpyAsync(status_h + lo, d + o_st + lo * 4, rows * 4 * (n_chunks == 1 ? S : 1),
                         cudaMemcpyDeviceToHost, cs));
    if (counters_h) {
      if (n_chunks == 1) {
        CK(cudaMemcpyAsync(counters_h, d + o_cn, 2 * S * B * 4, cudaMemcpyDeviceToHost, cs));
      } else {
        CK(cudaMemcpyAsync(counters_h + lo, d + o_cn + lo * 8, rows * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(counters_h + B + lo, d + o_cn + lo * 8 + rows * 4, rows * 4, cudaMemcpyDeviceToHost, cs));
      }
    }
  }
  // join
  for (int i = 1; i < n_streams; ++i) {
    CK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    CK(cudaEventRecord(ev_done[i], xs[i]));
    CK(cudaStreamWaitEvent(st, ev_done[i], 0));
  }
#undef CK
done:
  cudaFreeAsync(d, st);
  {
    cudaError_t e2 = cudaSuccess;
    for (int i = 1; i < n_streams; ++i)
      if (xs[i]) { cudaError_t e3 = cudaStreamSynchronize(xs[i]); if (e2 == cudaSuccess) e2 = e3; }
    cudaError_t e3 = cudaStreamSynchronize(st);
    if (e2 == cudaSuccess) e2 = e3;
    if (e == cudaSuccess) e = e2;
  }
  for (int i = 1; i < n_streams; ++i) {
    if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    if (xs[i]) cudaStreamDestroy(xs[i]);
  }
  if (ev_ready) cudaEventDestroy(ev_ready);
  if (lib_rc) return lib_rc;
  if (e != cudaSuccess) return cuda_fail(e, "hode_rollout_fwd_host");
  return 0;
}

}  // extern "C"
