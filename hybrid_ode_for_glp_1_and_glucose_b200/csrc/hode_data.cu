// hode_data.cu — the data formats either side of the path, on the device:
//   hode_window_dataset  GlucoseDataset's sliding windows + z-scoring (reference train/train_hybrid.py:43-155) for a
//                        cohort of equally long subject records (what hode_generate_4gi produces)
//   hode_eval_metrics    the reductions behind compute_rmse / compute_mae / compute_calibration_error and the
//                        normalised RMSE of evaluate_model (reference eval/evaluate.py:26-181, :262-286)
// Both are HBM-bound streaming passes (one read of every element); sums are accumulated in double, as block partials
// added in block order: bit-reproducible.
#include <stdio.h>

#include "hode_kernels.h"

namespace hode {
namespace {

constexpr int DB = 256;
constexpr int N_MET = 28;   // metric slots before the calibration counts

__device__ __forceinline__ void block_sum_store(double* red, double v, double* dst) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = DB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *dst = red[0];
  __syncthreads();
}

// number of windows (start = 0, stride, 2 stride, ... <= n_t - L) that contain time index j
__device__ __forceinline__ int window_multiplicity(int j, int n_t, int L, int stride) {
  if (n_t < L) return 0;
  const int last = (n_t - L) / stride;                       // index of the last window
  int lo = j - L + 1;                                        // starts s = k stride with lo <= s <= j
  lo = lo < 0 ? 0 : lo;
  const int k_lo = (lo + stride - 1) / stride;
  int k_hi = j / stride;
  k_hi = k_hi > last ? last : k_hi;
  return k_hi >= k_lo ? k_hi - k_lo + 1 : 0;
}

// partial[blk][0..5] = sum w x, [6..11] = sum w x^2, [12] = sum w   over the block's (subject, time) rows
__global__ void __launch_bounds__(DB) window_stats_kernel(const float* __restrict__ states, long n_rows, int n_t, int L, int stride,
                                                          double* __restrict__ partial) {
  __shared__ double red[DB];
  double sx[NS], sxx[NS], sw = 0.0;
#pragma unroll
  for (int c = 0; c < NS; ++c) { sx[c] = 0.0; sxx[c] = 0.0; }
  for (long r = (long)blockIdx.x * DB + threadIdx.x; r < n_rows; r += (long)gridDim.x * DB) {
    const int j = (int)(r % n_t);
    const double w = (double)window_multiplicity(j, n_t, L, stride);
    if (w == 0.0) continue;
    sw += w;
#pragma unroll
    for (int c = 0; c < NS; ++c) {
      const double x = (double)states[r * NS + c];
      sx[c] += w * x;
      sxx[c] += w * x * x;
    }
  }
  double* dst = partial + (size_t)blockIdx.x * 13;
#pragma unroll
  for (int c = 0; c < NS; ++c) { block_sum_store(red, sx[c], dst + c); block_sum_store(red, sxx[c], dst + 6 + c); }
  block_sum_store(red, sw, dst + 12);
}

// mean[c], std[c] = population std + 1e-6 (np.std, reference train_hybrid.py:118-119); identity when !normalize
__global__ void window_stats_final_kernel(const double* __restrict__ partial, int n_blocks, int normalize, double* __restrict__ mean_std) {
  const int c = threadIdx.x;
  if (c >= NS) return;
  double sx = 0, sxx = 0, sw = 0;
  for (int b = 0; b < n_blocks; ++b) { sx += partial[(size_t)b * 13 + c]; sxx += partial[(size_t)b * 13 + 6 + c]; sw += partial[(size_t)b * 13 + 12]; }
  double m = 0.0, sd = 1.0;
  if (normalize && sw > 0) {
    m = sx / sw;
    double var = sxx / sw - m * m;
    var = var > 0 ? var : 0;
    sd = sqrt(var) + 1e-6;
  }
  mean_std[c] = m;
  mean_std[NS + c] = sd;
}

// window w = subject * n_win + k: obs[w][l][c] = (states[subject][k stride + l][c] - mean) / std, etc.
__global__ void __launch_bounds__(DB) window_gather_kernel(const float* __restrict__ states, const float* __restrict__ inputs,
                                                           const float* __restrict__ time, int time_per_subject, int n_in, int n_t, int L,
                                                           int stride, int n_win, long n_total, const double* __restrict__ mean_std,
                                                           float* __restrict__ obs, float* __restrict__ init, float* __restrict__ win_in,
                                                           float* __restrict__ win_t) {
  const long e = (long)blockIdx.x * DB + threadIdx.x;   // (window, l)
  if (e >= n_total) return;
  const long w = e / L;
  const int l = (int)(e - w * L);
  const long subj = w / n_win;
  const int k = (int)(w - subj * n_win);
  const long row = subj * n_t + (long)k * stride + l;
#pragma unroll
  for (int c = 0; c < NS; ++c) {
    const float v = (float)(((double)states[row * NS + c] - mean_std[c]) / mean_std[NS + c]);
    obs[e * NS + c] = v;
    if (l == 0) init[w * NS + c] = v;
  }
  for (int c = 0; c < n_in; ++c) win_in[e * n_in + c] = inputs[row * n_in + c];
  win_t[e] = time_per_subject ? time[row] : time[(long)k * stride + l];
}

// ---- evaluation metrics --------------------------------------------------------------------------------------------
// slots: [0,6) sum (p - t)^2 per state | [6,12) sum |p - t| | [12,18) sum t | [18,24) sum t^2 | 24 sum (width + penalty)
// | 25 sum unc | 26 count inside the 95 % interval | 27 sum |p - t| / (unc + 1e-6) | 28 + i: count(normalised error <= thr[i])
__global__ void __launch_bounds__(DB) eval_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          const float* __restrict__ unc, float unc_const, const float* __restrict__ thr,
                                                          int n_bins, long n_rows, double* __restrict__ partial) {
  __shared__ double red[DB];
  double acc[N_MET];
#pragma unroll
  for (int i = 0; i < N_MET; ++i) acc[i] = 0.0;
  double cnt[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) cnt[i] = 0.0;
  const bool calib = unc != nullptr || unc_const > 0.f;
  for (long r = (long)blockIdx.x * DB + threadIdx.x; r < n_rows; r += (long)gridDim.x * DB) {
#pragma unroll
    for (int c = 0; c < NS; ++c) {
      const double p = (double)pred[r * NS + c], t = (double)target[r * NS + c];
      const double d = p - t;
      acc[c] += d * d;
      acc[6 + c] += fabs(d);
      acc[12 + c] += t;
      acc[18 + c] += t * t;
      if (calib) {
        const double u = unc ? (double)unc[r * NS + c] : (double)unc_const;
        const double lower = p - 1.96 * u, upper = p + 1.96 * u;
        const double pen = (2.0 / 0.05) * ((t < lower ? lower - t : 0.0) + (t > upper ? t - upper : 0.0));
        acc[24] += (upper - lower) + pen;
        acc[25] += u;
        acc[26] += (t >= lower && t <= upper) ? 1.0 : 0.0;
        const double ne = fabs(d) / (u + 1e-6);
        acc[27] += ne;
        for (int i = 0; i < n_bins; ++i) cnt[i] += ne <= (double)thr[i] ? 1.0 : 0.0;
      }
    }
  }
  double* dst = partial + (size_t)blockIdx.x * (N_MET + 32);
#pragma unroll
  for (int i = 0; i < N_MET; ++i) block_sum_store(red, acc[i], dst + i);
  for (int i = 0; i < 32; ++i) block_sum_store(red, i < n_bins ? cnt[i] : 0.0, dst + N_MET + i);
}

__global__ void eval_metrics_final_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ out) {
  const int i = threadIdx.x;
  if (i >= N_MET + 32) return;
  double s = 0;
  for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * (N_MET + 32) + i];
  out[i] = s;
}

}  // namespace
}  // namespace hode

using namespace hode;

extern "C" {

int hode_window_count(int32_t n_t, int32_t sequence_length, int32_t stride) {
  if (sequence_length < 1 || stride < 1 || n_t < sequence_length) return 0;
  return (n_t - sequence_length) / stride + 1;
}

int hode_window_dataset(int32_t n_subjects, int32_t n_t, int32_t n_inputs, int32_t sequence_length, int32_t stride,
                        int32_t normalize, int32_t time_per_subject, const float* states, const float* inputs, const float* time,
                        float* obs, float* initial_state, float* win_inputs, float* win_time, double* mean_std, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (n_subjects < 0 || n_t < 1 || n_inputs < 0 || sequence_length < 1 || stride < 1) return HODE_E_SIZE;
  if (!states || !time || !obs || !initial_state || !win_time || !mean_std || (n_inputs > 0 && (!inputs || !win_inputs))) return HODE_E_NULL;
  const int n_win = hode_window_count(n_t, sequence_length, stride);
  cudaStream_t st = (cudaStream_t)stream;
  const int n_blocks = 296;
  if (!workspace || workspace_bytes < (size_t)n_blocks * 13 * sizeof(double)) return HODE_E_WORKSPACE;
  double* partial = (double*)workspace;
  const long n_rows = (long)n_subjects * n_t;
  count_launch();
  window_stats_kernel<<<n_blocks, DB, 0, st>>>(states, n_rows, n_t, sequence_length, stride, partial);
  count_launch();
  window_stats_final_kernel<<<1, 32, 0, st>>>(partial, n_blocks, normalize, mean_std);
  const long n_total = (long)n_subjects * n_win * sequence_length;
  if (n_total > 0) {
    count_launch();
    window_gather_kernel<<<(unsigned)((n_total + DB - 1) / DB), DB, 0, st>>>(states, inputs, time, time_per_subject, n_inputs, n_t,
                                                                           sequence_length, stride, n_win, n_total, mean_std, obs,
                                                                           initial_state, win_inputs, win_time);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

int hode_eval_metrics(int64_t n_rows, const float* pred, const float* target, const float* unc, float unc_const,
                      const float* thresholds, int32_t n_bins, double* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (n_rows < 0 || n_bins < 0 || n_bins > 32) return HODE_E_SIZE;
  if (!pred || !target || !out || (n_bins > 0 && !thresholds)) return HODE_E_NULL;
  const int n_blocks = 296;
  if (!workspace || workspace_bytes < (size_t)n_blocks * (N_MET + 32) * sizeof(double)) return HODE_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  count_launch();
  eval_metrics_kernel<<<n_blocks, DB, 0, st>>>(pred, target, unc, unc_const, thresholds, n_bins, (long)n_rows, (double*)workspace);
  count_launch();
  eval_metrics_final_kernel<<<1, 64, 0, st>>>((const double*)workspace, n_blocks, out);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // extern "C"
