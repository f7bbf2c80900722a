// hode_kernels.h — internal launcher declarations shared by the libhode translation units.
#pragma once
#include <cuda_runtime.h>

#include "hode_common.cuh"

#define HODE_SIMT_MAX_SHARED_T 2048

namespace hode {

// one trajectory per thread, FP32 CUDA cores (hode_rollout_simt.cu)
cudaError_t launch_rollout_simt(const RolloutArgs& A, int mlp_mode, cudaStream_t stream);

// batched single RHS evaluation (hode_rollout_simt.cu)
cudaError_t launch_rhs(const RolloutArgs& A, int mlp_mode, const float* t, const float* state,
                       float* out, cudaStream_t stream);

// 128-trajectory tiles, MLP on tcgen05 tensor cores (hode_rollout_tc.cu)
size_t tc_workspace_bytes(int S, int L);
cudaError_t launch_rollout_tc(const RolloutArgs& A, int mlp_mode, void* workspace, cudaStream_t stream);

}  // namespace hode
