/*
 * hode.h — C ABI of libhode.so, the B200-native batched hybrid ODE-NN integrator.
 *
 * This is the drop-in boundary for the trajectory-rollout and gradient path of
 * OliverDOU776/Hybrid-ODE-for-GLP-1-and-Glucose.  The reference has no FFI of
 * its own (it is pure Python); each entry point below names the reference
 * interface it replaces (file:line into the reference tree).
 *
 * Conventions
 *   - every function is extern "C", returns int: 0 = ok, <0 = argument error
 *     (HODE_E_*), >0 = a cudaError_t raised by a launch.  Nothing throws.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - all launches are asynchronous on the caller's stream (a cudaStream_t
 *     passed as void*; 0 = the legacy default stream).
 *   - the library allocates nothing behind the caller's back: workspace sizes
 *     come from hode_workspace_bytes().
 *   - there is no CPU fallback anywhere in this library.
 *
 * State layout (reference models/ode_core.py:103-108): 6 floats per trajectory,
 *   [G, I, Glu, GLP1, GE, FFA], row-major [B,6].
 * theta layout (reference models/ode_core.py:44-71, buffer registration order):
 *   [a_GI, k_I, rho, G_b, I_b, E_max, EC_50, Glu_b, V_max, K_m, k_L,
 *    k_GE0, IGD_50, g, p_7, p_8, p_9]                      -> 17 floats / set.
 * W layout (reference models/nn_residual.py:60-78, named_parameters() order):
 *   for l in 0..L: weight_l [out_l, in_l] row-major, then bias_l [out_l];
 *   in_0 = 9, out_L = 6, all other widths = nn_hidden.  (13 510 floats at 64x4.)
 */
#ifndef HODE_H_
#define HODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HODE_ABI_VERSION 4
#define HODE_N_STATE 6
#define HODE_N_THETA 17
#define HODE_NN_IN 9
#define HODE_MAX_HIDDEN 128
#define HODE_MAX_LAYERS 8

/* solver keyword (reference models/hybrid_ode_nn.py:137,174-181) */
enum {
  HODE_SOLVER_RK4 = 0,    /* classical fixed-step RK4, n_substeps per observation interval */
  HODE_SOLVER_DOPRI5 = 1, /* Dormand-Prince 5(4), SciPy RK45 controller + dense output     */
  HODE_SOLVER_DOP853 = 2  /* Hairer's DOP853 as SciPy runs it (rk.py:568-720): what the reference's solver='dopri5'
                             and 'dop853' reach (models/hybrid_ode_nn.py:174-181).  float32 RHS, float64 stepping
                             (the reference's arithmetic); FP32 CUDA-core kernels only, forward only */
};

/* how one external input channel is supplied (reference models/hybrid_ode_nn.py:217-231) */
enum {
  HODE_IN_ABSENT = 0, /* channel is identically zero (models/ode_core.py:115-117)        */
  HODE_IN_CONST = 1,  /* one value per trajectory, [B]   (values.dim()==1 branch, :230)  */
  HODE_IN_SERIES = 2  /* one value per observation time, [B,T], linearly interpolated    */
};
#define HODE_CH_MEAL 0
#define HODE_CH_TVNS 1
#define HODE_CH_GD 2

/* arithmetic used for the residual MLP inside each RK stage */
enum {
  HODE_MLP_NONE = 0,   /* no residual network (ablation no_nn, train/train_hybrid.py:423-427) */
  HODE_MLP_FP32 = 1,   /* FP32 FMA on CUDA cores: parity mode                                */
  HODE_MLP_TF32X3 = 2, /* tcgen05 tensor cores, 3xTF32 split (fp32-equivalent accuracy)      */
  HODE_MLP_TF32 = 3,   /* tcgen05 tensor cores, single TF32 pass (fast, ~1e-3 on residual)   */
  HODE_MLP_TF32BF16 = 4, /* tcgen05: one TF32 pass + two BF16 cross-term passes (2 pass-equivalents; max
                            error 4.5e-7 of sum|ab| per product against 1.7e-7 for TF32X3: opt-in)   */
  HODE_MLP_TF32X2BF16 = 5, /* tcgen05: A_hi*B_hi and A_hi*B_lo in TF32, A_lo*B_hi in BF16 (2.5 pass-equivalents,
                             error <= TF32BF16's).  Its TMEM footprint lets the rollout run THREE 128-trajectory
                             tiles per SM instead of two: the fastest float32-equivalent mode (nn_layers <= 4).
                             Gradients of a rollout made in this mode are computed with the TF32X3 adjoint. */
  HODE_MLP_F16BF16X2 = 6  /* tcgen05: hi parts in FP16 (11 significant bits, like TF32), remainders in BF16; three
                             kind::f16 passes f16(A_hi)*f16(B_hi) + f16(A_hi/64)*f16(64 B_lo) + bf16(A_lo)*bf16(B_hi) = 1.5
                             pass-equivalents, three tiles per SM (nn_layers <= 4).  Float32-equivalent products while
                             |activations| and |weights| stay below 65504 (FP16's range; the hi part saturates beyond
                             it and precision falls to BF16's).  Gradients: TF32X3 adjoint, as above.
                             Two launch shapes with bit-identical results: three tiles per SM, or — when the cohort fits
                             two tiles per SM anyway (n_traj <= 256 x SMs) — two tiles with helper warps (shorter round).
                             The environment variable HODE_H16_TILES=2|3 forces one (tests and measurements). */
};

/*
 * Treatment of the kinks of the piece-wise-linear inputs by the adaptive solver.
 * SCIPY reproduces the reference: steps ignore the observation grid and the controller
 * finds (or steps over!) each kink by rejection (scipy ivp.py:701-728 never clips to t_eval).
 * CLIP ends a step at every grid point where a series input changes slope, so every step
 * integrates a smooth RHS and the result converges to the exact solution.
 */
enum { HODE_KINK_SCIPY = 0, HODE_KINK_CLIP = 1 };

/* which terms hode_rhs returns */
enum { HODE_RHS_FULL = 0, HODE_RHS_NN_ONLY = 1 };

/* per-trajectory status codes written by the rollout */
enum {
  HODE_ST_OK = 0,
  HODE_ST_STEP_TOO_SMALL = 1, /* SciPy: "Required step size is less than spacing between numbers." */
  HODE_ST_MAX_STEPS = 2,      /* attempt budget (cfg.max_steps) exhausted                        */
  HODE_ST_NONFINITE = 3,      /* state became NaN/Inf                                            */
  HODE_ST_REC_OVERFLOW = 4    /* save_steps: more accepted steps than cfg.max_saved_steps records (the
                                 trajectory is zero-padded and gets no gradient; raise the capacity)  */
};

/* argument errors */
enum {
  HODE_E_NULL = -1,
  HODE_E_SIZE = -2,
  HODE_E_SHAPE = -3,
  HODE_E_UNSUPPORTED = -4,
  HODE_E_WORKSPACE = -5,
  HODE_E_NO_DEVICE = -6
};

typedef struct hode_cfg {
  int32_t struct_bytes;  /* = sizeof(hode_cfg); checked                                        */
  int32_t n_traj;        /* B: trajectories (rows of y0)                                       */
  int32_t n_obs;         /* T: observation times per trajectory (>= 1)                         */
  int32_t t_per_traj;    /* 0: t_obs is [T] shared; 1: t_obs is [B,T]                          */
  int32_t in_mode[3];    /* HODE_IN_* for meal, tVNS, GD                                       */
  int32_t nn_hidden;     /* H (ignored when mlp == HODE_MLP_NONE)                              */
  int32_t nn_layers;     /* L hidden layers (reference n_layers); Linear count = L+1           */
  int32_t mlp;           /* HODE_MLP_*                                                         */
  int32_t n_samples;     /* S >= 1 parameter sets (VI sweep); unit (s,b) uses theta[s], W[s]   */
  int32_t solver;        /* HODE_SOLVER_*                                                      */
  int32_t n_substeps;    /* RK4: steps per observation interval (>= 1)                         */
  int32_t max_steps;     /* DOPRI5: attempted-step budget per trajectory (0 -> 100000)         */
  double rtol;           /* DOPRI5 (reference default 1e-6, models/hybrid_ode_nn.py:138)       */
  double atol;           /* DOPRI5 (reference default 1e-8); both at byte offset 56/64         */
  int32_t save_steps;    /* 1: record accepted steps in the workspace for hode_rollout_bwd     */
  int32_t kink_mode;     /* DOPRI5: HODE_KINK_* (how input-slope discontinuities are handled)  */
  int32_t max_saved_steps; /* DOPRI5 + save_steps: accepted-step record capacity per trajectory;
                              0 -> max(128, 2 (T - 1)): with kink clipping a series input that changes at
                              every grid point forces T - 1 accepted steps                          */
  int32_t rhs_part;      /* hode_rhs only: HODE_RHS_FULL or HODE_RHS_NN_ONLY                    */
} hode_cfg;

/* ABI version of the loaded library. */
int hode_version(void);

/* Number of CUDA kernels this library has launched in this process so far (library kernels such as cub's radix
 * sort are not counted): bench.py reports the difference over its timed regions as gpu_launches. */
int64_t hode_launch_count(void);

/* Human-readable text for the last non-zero return on this host thread. */
const char* hode_last_error_string(void);

/* Number of floats in one packed MLP parameter set (13 510 for hidden=64, layers=4). */
int64_t hode_mlp_param_count(int32_t nn_hidden, int32_t nn_layers);

/* Floats per accepted-step record in the forward workspace for this cfg (16, or 48 when the tensor-core
 * adjoint will consume them: the record then carries the step's stage derivatives); 0 without save_steps. */
int32_t hode_step_record_floats(const hode_cfg* cfg);

/* Accepted-step record capacity per trajectory hode_rollout_fwd will use for this cfg (see max_saved_steps). */
int32_t hode_step_record_capacity(const hode_cfg* cfg);

/* Bytes of device workspace hode_rollout_fwd needs (saved steps when save_steps = 1, tensor-core
 * weight images) and of gradient scratch hode_rollout_bwd / hode_rhs_vjp need. */
int hode_workspace_bytes(const hode_cfg* cfg, size_t* fwd_bytes, size_t* bwd_bytes);

/*
 * Batched IVP solve: replaces the per-trajectory loop + scipy.integrate.solve_ivp
 * call of HybridODENN.forward (reference models/hybrid_ode_nn.py:184-256) and, with
 * n_samples > 1, the forward_with_params loop of the VI sweeps
 * (reference inference/vi.py:294-304, models/bayes.py:199-206).
 *
 *   y0     [B,6]                 t_obs [T] or [B,T] (strictly increasing per row)
 *   u[c]   NULL | [B] | [B,T]    per cfg.in_mode[c]
 *   theta  [S,17]                W [S, hode_mlp_param_count] (NULL when mlp == NONE)
 *   traj   [S,B,T,6]  out        (rows after a solver failure are zero, as the
 *                                 reference zero-pads, models/hybrid_ode_nn.py:252-254)
 *   status [S,B] out             counters [2,S,B] out (accepted, rejected attempts); may be NULL
 */
int hode_rollout_fwd(const hode_cfg* cfg, const float* y0, const float* t_obs,
                     const float* u_meal, const float* u_tvns, const float* u_gd,
                     const float* theta, const float* W, float* traj, int32_t* status,
                     int32_t* counters, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Options of the extended rollout entry (all optional; a zeroed struct = hode_rollout_fwd).
 *   theta_per_traj  1: theta is [B,17], one mechanistic parameter set per TRAJECTORY (n_samples must be 1; the
 *                   network parameters W [1,P] stay shared).  Replaces the loop of the reference's Sobol sweep, which
 *                   overwrites the ode_core buffers and calls forward() once per parameter set
 *                   (plots/plot_all.py:168-187: 16 384 sets x 1 trajectory).  Forward only.
 *   order           device pointer to [B] trajectory indices (a permutation of 0..B-1), or NULL: the order in which
 *                   the tensor-core rollout hands trajectories to its lanes.  Longest first (descending attempt
 *                   counters of the previous pass over the same cohort — training re-integrates a cohort every
 *                   epoch, train/train_hybrid.py:225-275) removes the tail that bounds small cohorts.  Results do not
 *                   depend on it.  Ignored by the FP32 kernels and by the fused posterior-predictive sweep.
 *   out_state_mask  hode_rollout_fwd_host only: bit i set = state column i is copied back to the host; traj_host is
 *                   then [B,T,popcount(mask)] (0 = all six columns).  The loss of the reference consumes all columns,
 *                   but its figures and metrics read glucose-centric ones (plots/plot_all.py:183-187).
 *   prev_counters   hode_rollout_fwd_host(_ex) only: HOST pointer to the [2,B] attempt counters a previous call returned
 *                   for this cohort (counters_host of that call; it may be the very buffer this call writes), or NULL.
 *                   The library derives the launch order from them on the device: longest first within blocks of
 *                   32 768 trajectories, blocks in index order — the tail of the launch disappears as with `order`, and
 *                   result blocks still complete (and travel back) progressively.  A scheduling hint only.
 */
typedef struct hode_fwd_opts {
  int32_t struct_bytes;    /* = sizeof(hode_fwd_opts); checked */
  int32_t theta_per_traj;
  const int32_t* order;
  uint32_t out_state_mask;
  uint32_t reserved;
  const int32_t* prev_counters;
} hode_fwd_opts;

/* hode_rollout_fwd with options (opts == NULL: identical to hode_rollout_fwd). */
int hode_rollout_fwd_ex(const hode_cfg* cfg, const hode_fwd_opts* opts, const float* y0, const float* t_obs,
                        const float* u_meal, const float* u_tvns, const float* u_gd,
                        const float* theta, const float* W, float* traj, int32_t* status,
                        int32_t* counters, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Discrete adjoint of hode_rollout_fwd over the recorded accepted steps (step sizes
 * frozen): the gradient autograd would give through the unrolled RK steps.  The reference
 * has no through-solver gradient (models/hybrid_ode_nn.py:248 returns a graph-free tensor);
 * this is the gradient path BASELINE.json's north_star item (3) asks for.
 *   cfg must be the forward's cfg (save_steps = 1); fwd_workspace is the buffer that
 *   forward call filled; bwd_workspace has hode_workspace_bytes().bwd_bytes bytes.
 *   grad_traj [S,B,T,6] in;  grad_y0 [S,B,6] out (may be NULL);
 *   grad_theta [S,17] out (may be NULL);  grad_W [S,P] out (may be NULL).
 * Trajectories whose forward status is not HODE_ST_OK contribute zero gradient.
 * The summation order is fixed: results are bit-reproducible run to run.
 */
int hode_rollout_bwd(const hode_cfg* cfg, const float* y0, const float* t_obs,
                     const float* u_meal, const float* u_tvns, const float* u_gd,
                     const float* theta, const float* W, const float* grad_traj,
                     float* grad_y0, float* grad_theta, float* grad_W,
                     const void* fwd_workspace, size_t fwd_workspace_bytes,
                     void* bwd_workspace, size_t bwd_workspace_bytes, void* stream);

/*
 * One training-shaped step of the data term in a single call: rollout with recorded steps,
 * loss[s] = mean((traj[s] - obs)^2) over [B,T,6] and its discrete-adjoint gradients.  Replaces
 * `predictions = self.forward(...)`, `F.mse_loss(predictions, observations)` (reference
 * models/hybrid_ode_nn.py:284-288) and the backward of that term in train_epoch
 * (train/train_hybrid.py:244-252) — which in the reference carries no gradient because forward
 * returns a graph-free tensor (:248).  Same contracts as hode_rollout_fwd / hode_rollout_bwd
 * (cfg.save_steps must be 1); between them the residual 2 (traj - obs) / (B T 6) is formed on
 * the device and the loss is reduced in a fixed order (bit-reproducible).
 *   obs [B,T,6] in (shared by the S parameter sets);  traj [S,B,T,6] out;  loss [S] out (device);
 *   grad_traj_scratch [S,B,T,6] (caller-owned scratch, holds dloss/dtraj on return);
 *   grad_y0 [S,B,6] / grad_theta [S,17] / grad_W [S,P] out (each may be NULL).
 * The physics-residual and L2 / KL terms of the reference's loss are parameter-side sums the
 * caller adds (hode_rhs + hode_rhs_vjp; models/hybrid_ode_nn.py:327-345).
 */
int hode_loss_fused_fwd_bwd(const hode_cfg* cfg, const float* y0, const float* t_obs,
                            const float* u_meal, const float* u_tvns, const float* u_gd,
                            const float* theta, const float* W, const float* obs, float* traj,
                            int32_t* status, int32_t* counters, float* loss,
                            float* grad_traj_scratch, float* grad_y0, float* grad_theta,
                            float* grad_W, void* fwd_workspace, size_t fwd_workspace_bytes,
                            void* bwd_workspace, size_t bwd_workspace_bytes, void* stream);

/*
 * One training step of the reference in a single call: HybridODENN.loss (reference models/hybrid_ode_nn.py:263-351) plus the
 * batch body of train_epoch (train/train_hybrid.py:244-261: backward, clip_grad_norm_, Adam.step) on the packed network
 * parameters.  Stream-ordered, allocation-free (one workspace), no host synchronisation: capturable in a CUDA graph.
 *   loss = MSE(rollout, obs) + lambda1 * physics + lambda2 * (lambda2 * sum ||Linear.weight||^2)       (the reference
 *   applies lambda2 twice, :336-345), physics = mean over the n_physics drawn grid indices k of
 *   MSE((Phi_dt(x_k) - x_k) / dt, f(t_k, x_k)) with the inputs frozen at index k (:297-333); the (index, trajectory) pairs
 *   are stacked into ONE re-solve, ONE RHS and ONE RHS-VJP launch.
 *   Gradient: as in the reference only the physics residual (through f) and the L2 term carry gradient (the solves are
 *   graph-free, :248); data_gradient = 1 adds the discrete adjoint of the data term (hode_rollout_bwd).
 *   Then ||g|| clipping (torch.nn.utils.clip_grad_norm_) and torch.optim.Adam's update (weight_decay 0), in place.
 */
typedef struct hode_train_cfg {
  int32_t struct_bytes;   /* = sizeof(hode_train_cfg); checked                                                    */
  int32_t n_physics;      /* number of physics grid indices (the reference: min(20, len(time_points))); 0 = no term */
  int32_t data_gradient;  /* 0: reference semantics (no gradient through the solver); 1: + adjoint of the data term */
  int32_t adam_step;      /* 1-based count of this update; 0 = loss and gradient only (validate()); < 0 = the count
                             lives on the device (scalars[6], incremented by the call): CUDA-graph replays          */
  float lambda1, lambda2;
  float grad_clip;        /* max gradient norm (config training.gradient_clip); <= 0: no clipping                  */
  float lr, beta1, beta2, eps;
  float physics_dt;       /* local-time horizon of the re-solve (0.1 in the reference; <= 0 selects it)             */
} hode_train_cfg;

int hode_train_step_workspace_bytes(const hode_cfg* cfg, const hode_train_cfg* tc, size_t* bytes);

/*
 *   W [P] in/out (updated when adam_step > 0); obs [B,T,6]; physics_idx DEVICE int32 [n_physics] (grid indices, drawn by
 *   the caller the way the reference draws them: torch.randperm(len(time_points))[:n]); adam_m / adam_v [P] in/out;
 *   grad_W [P] out (the clipped gradient, what p.grad holds after clip_grad_norm_); scalars [9] (device): out [0..5] =
 *   total loss, data, physics, reg (= lambda2 * sum w^2), gradient norm before clipping, clip coefficient; [6] = update
 *   count (in/out when adam_step < 0), [7], [8] scratch;
 *   traj [B,T,6] out (the predictions); status [B] out (may be NULL).  cfg.n_samples must be 1, cfg.mlp != NONE.
 */
int hode_train_step(const hode_cfg* cfg, const hode_train_cfg* tc, const float* y0, const float* t_obs,
                    const float* u_meal, const float* u_tvns, const float* u_gd, const float* theta, float* W,
                    const float* obs, const int32_t* physics_idx, float* adam_m, float* adam_v, float* grad_W,
                    float* scalars, float* traj, int32_t* status, void* workspace, size_t workspace_bytes,
                    void* stream);

/*
 * Vector-Jacobian product of hode_rhs: the backward of HybridODENN.ode_residual, the only
 * differentiable model call of the reference's loss (models/hybrid_ode_nn.py:327 -> :330,
 * loss.backward() at train/train_hybrid.py:252).
 *   grad_out [B,6] in;  grad_state [B,6] out (may be NULL);  grad_theta [17] out (may be
 *   NULL);  grad_W [P] out (may be NULL).  cfg.rhs_part selects FULL or NN_ONLY as in hode_rhs.
 */
int hode_rhs_vjp(const hode_cfg* cfg, const float* t, const float* state, const float* u_meal,
                 const float* u_tvns, const float* u_gd, const float* theta, const float* W,
                 const float* grad_out, float* grad_state, float* grad_theta, float* grad_W,
                 void* bwd_workspace, size_t bwd_workspace_bytes, void* stream);

/*
 * Posterior-predictive sweep with the mean / unbiased std over the S parameter sets
 * reduced on the fly (reference inference/vi.py:291-310, models/bayes.py:196-212):
 * the [S,B,T,6] stack is never materialised.  Every trajectory is owned by one CTA for all S
 * sets and updated in a fixed order (Welford): no atomics, bit-reproducible.  Failed
 * (s, b) units enter the statistics as the zero-padded rows the reference would stack.
 *   mean [B,T,6] out, std [B,T,6] out (NaN when S == 1, like torch.std), status [S,B] out
 *   (may be NULL), counters [2,S,B] out (may be NULL); workspace as for hode_rollout_fwd.
 */
int hode_vi_predictive(const hode_cfg* cfg, const float* y0, const float* t_obs,
                       const float* u_meal, const float* u_tvns, const float* u_gd,
                       const float* theta, const float* W, float* mean, float* std_out,
                       int32_t* status, int32_t* counters, void* workspace,
                       size_t workspace_bytes, void* stream);

/*
 * One batched evaluation of f_physio + g_NN: replaces HybridODENN.ode_residual
 * (reference models/hybrid_ode_nn.py:108-134) for the physics-residual loss term
 * (:327).  t [B], state [B,6], u[c] NULL|[B], out [B,6].
 */
int hode_rhs(const hode_cfg* cfg, const float* t, const float* state, const float* u_meal,
             const float* u_tvns, const float* u_gd, const float* theta, const float* W,
             float* out, void* stream);

/*
 * Cohort generation on the device (the step before the path): the reference's
 * 8-state 4GI simulator FourGIModel.simulate (data/generate4GI.py:72-211 — equations :72-160, one
 * scipy.integrate.odeint call per sampling interval with the meal spread over that interval
 * :188-200) for n_subjects subjects at once, one thread per subject, adaptive DP5(4) in float64.
 *   patient_type  0 = 'T2DM', 1 = 'HV' (:19-24, :103-106)
 *   baselines [n,5]       glucose, insulin, GLP-1, glucagon, GIP baselines (BSL*, :64-70; generate_dataset
 *                         perturbs them per subject, :231-235)
 *   meal_rate [n,T-1]     glucose input (mmol/h) during sampling interval k, or NULL (no meals);
 *                         the reference's value is meal_size / interval_hours for the interval holding meal_time
 *   out [n,T,5]           concentrations at the sampling times: glucose (mmol/L), insulin, GLP-1,
 *                         glucagon, GIP (pmol/L) — simulate()'s return values (:204-210)
 *   status [n] (may be NULL): HODE_ST_*; rows after a failure are zero
 * rtol / atol <= 0 select 1e-9 / 1e-12.
 */
int hode_generate_4gi(int32_t n_subjects, int32_t n_obs, double interval_hours, int32_t patient_type,
                      double rtol, double atol, const float* baselines, const float* meal_rate,
                      float* out, int32_t* status, void* stream);

/*
 * The step before the path on the device: GlucoseDataset's sliding windows and z-scoring (reference
 * train/train_hybrid.py:43-155) for a cohort of n_subjects equally long records (what hode_generate_4gi produces).
 * Windows start at 0, stride, 2 stride, ... <= n_t - sequence_length per subject (:104-115); the normalisation statistics
 * are the mean and population std (+ 1e-6) over ALL rows of ALL windows, overlapping rows counted once per window, in
 * float64 (:117-120); initial_state is the first row of each normalised window (:141).
 *   states [n_subjects, n_t, 6], inputs [n_subjects, n_t, n_inputs] (may be NULL when n_inputs == 0), time [n_t] shared
 *   or [n_subjects, n_t] (time_per_subject);  with W = n_subjects * hode_window_count(...):
 *   obs [W, L, 6], initial_state [W, 6], win_inputs [W, L, n_inputs], win_time [W, L] out; mean_std [12] DOUBLE out
 *   (device: 6 means then 6 stds; mean 0 / std 1 when normalize == 0); workspace >= 32 KiB.
 */
int hode_window_count(int32_t n_t, int32_t sequence_length, int32_t stride);
int hode_window_dataset(int32_t n_subjects, int32_t n_t, int32_t n_inputs, int32_t sequence_length, int32_t stride,
                        int32_t normalize, int32_t time_per_subject, const float* states, const float* inputs,
                        const float* time, float* obs, float* initial_state, float* win_inputs, float* win_time,
                        double* mean_std, void* workspace, size_t workspace_bytes, void* stream);

/*
 * The reductions of the batch consumers after the path: compute_rmse / compute_mae per state and
 * overall, the calibration sums of compute_calibration_error and the target moments behind evaluate_model's normalised
 * RMSE (reference eval/evaluate.py:26-181, :262-286), in one pass over pred / target / unc [n_rows, 6].
 *   unc NULL: unc_const is used for every element (the reference's fixed 0.1 for non-Bayesian models, :239); unc NULL
 *   and unc_const <= 0: no calibration sums.  thresholds [n_bins] (n_bins <= 32): the normalised-error thresholds of the
 *   calibration curve (host-computed the reference's way, :139-143).
 *   out [60] DOUBLE (device): [0,6) sum (p-t)^2 per state, [6,12) sum |p-t|, [12,18) sum t, [18,24) sum t^2, [24] sum of
 *   interval width + penalty (MSIS numerator), [25] sum unc, [26] count inside the 95 % interval, [27] sum of normalised
 *   errors, [28 + i] count(normalised error <= thresholds[i]).  workspace >= 160 KiB.
 */
int hode_eval_metrics(int64_t n_rows, const float* pred, const float* target, const float* unc, float unc_const,
                      const float* thresholds, int32_t n_bins, double* out, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * Host-buffer convenience entry (the e2e path timed by bench.py): same contract as
 * hode_rollout_fwd but every pointer is HOST memory (pinned or pageable); the call
 * stages H2D copies, runs the rollout and copies traj/status/counters back, all on
 * `stream`, and synchronises that stream before returning.
 */
int hode_rollout_fwd_host(const hode_cfg* cfg, const float* y0_host, const float* t_obs_host,
                          const float* u_meal_host, const float* u_tvns_host,
                          const float* u_gd_host, const float* theta_host, const float* W_host,
                          float* traj_host, int32_t* status_host, int32_t* counters_host,
                          void* stream);

/* hode_rollout_fwd_host with options: opts->out_state_mask selects the state columns copied back (traj_host is
 * [B,T,popcount(mask)]); theta_per_traj as in hode_rollout_fwd_ex (theta_host is then [B,17]); prev_counters gives the
 * launch-order hint of a cohort integrated before; order is ignored. */
int hode_rollout_fwd_host_ex(const hode_cfg* cfg, const hode_fwd_opts* opts, const float* y0_host,
                             const float* t_obs_host, const float* u_meal_host, const float* u_tvns_host,
                             const float* u_gd_host, const float* theta_host, const float* W_host,
                             float* traj_host, int32_t* status_host, int32_t* counters_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HODE_H_ */
