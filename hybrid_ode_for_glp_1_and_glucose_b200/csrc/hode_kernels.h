// hode_kernels.h — internal launcher declarations shared by the libhode translation units.
#pragma once
#include <cuda_runtime.h>

#include "hode_common.cuh"

#define HODE_SIMT_MAX_SHARED_T 2048

namespace hode {

// every kernel launch of the library is counted (hode_launch_count: what bench.py reports as gpu_launches)
void count_launch();

// one trajectory per thread, FP32 CUDA cores (hode_rollout_simt.cu)
cudaError_t launch_rollout_simt(const RolloutArgs& A, int mlp_mode, cudaStream_t stream);
size_t simt_min_smem_bytes(int H, int L);

// batched single RHS evaluation (hode_rollout_simt.cu)
cudaError_t launch_rhs(const RolloutArgs& A, int mlp_mode, const float* t, const float* state,
                       float* out, cudaStream_t stream);

// 128-trajectory tiles, MLP on tcgen05 tensor cores (hode_rollout_tc.cu)
size_t tc_workspace_bytes(int S, int L, int B);
cudaError_t launch_rollout_tc(const RolloutArgs& A, int mlp_mode, void* workspace, cudaStream_t stream);

// FP32 CUDA-core gradients (hode_adjoint_simt.cu)
struct AdjPlan {
  int grid_x, grid_y;
  size_t smem;            // dynamic shared memory per CTA
  size_t partial_floats;  // per-CTA partial gradients
  size_t scratch_floats;  // activation stash
};
AdjPlan adj_plan(int B, int S, int H, int L, int P, bool has_nn, int T, int t_per_traj, int n_stages);
size_t adj_workspace_bytes(const AdjPlan& p);
cudaError_t launch_rollout_bwd(const RolloutArgs& A, bool has_nn, const float* grad_traj, float* grad_y0,
                               float* grad_theta, float* grad_W, void* workspace, cudaStream_t stream);
cudaError_t launch_rhs_vjp(const RolloutArgs& A, bool has_nn, const float* t, const float* state,
                           const float* grad_out, float* grad_state, float* grad_theta, float* grad_W,
                           void* workspace, cudaStream_t stream);

cudaError_t launch_reduce_partials(const float* partials, int ctas_per_set, int S, int P, float* grad_W,
                                   float* grad_theta, cudaStream_t stream);

// tensor-core discrete adjoint (hode_adjoint_tc.cu)
int tc_image_floats(int L);                      // 3xTF32 / TF32+BF16 images (what the adjoint recomputes from)
int tc_image_floats_mode(int L, int mlp_mode);   // ... and the larger HODE_MLP_TF32X2BF16 image
int tc_bwd_image_floats(int L);
cudaError_t tc_prepare_fwd_images(const float* W, float* img, int S, int L, int P, int mlp_mode, cudaStream_t stream);
struct AdjTcPlan {
  int grid_x, grid_y, fwd_floats, bwd_floats, n_tiles, t_in_smem;
  size_t smem, partial_floats, stash_floats, img_floats;
  size_t sched_ints;   // sort keys / values, tile owners and per-CTA tile lists (int32 words)
  size_t sort_bytes;   // cub::DeviceRadixSort temporaries
};
bool adj_tc_supported(int H, int L);
AdjTcPlan adj_tc_plan(int B, int S, int L, int P, int T, int t_per_traj);
size_t adj_tc_workspace_bytes(const AdjTcPlan& p);
cudaError_t launch_rollout_bwd_tc(const RolloutArgs& A, int mlp_mode, const float* grad_traj, float* grad_y0,
                                  float* grad_theta, float* grad_W, void* workspace, cudaStream_t stream);

// on-device cohort generation (hode_gen4gi.cu)
cudaError_t launch_gen4gi(int n, int T, double dt_hours, int patient_type, double rtol, double atol,
                          const float* baselines, const float* meal_rate, float* out, int32_t* status,
                          cudaStream_t stream);

}  // namespace hode
