// adj_probe.cu — hardware probe for the tensor-core ADJOINT (not part of libhode.so).
//
// Validates on a real B200 the two tcgen05 operand forms the backward pass needs:
//   test 1  delta-backprop:  D[m][n] = sum_k A[m][k] * W[k][n]   (A from TMEM, M=128, K=N=64)
//           with B = the FORWARD weight image of W[out=k][in=n] read as an MN-major operand
//           (no second, transposed copy of the weights in shared memory);
//   test 2  weight gradient: D[j][n] = sum_t Dl[t][j] * Ac[t][n] (contraction over the 128
//           trajectories of a tile): both operands from shared memory (SS form), both
//           MN-major, written by "thread t owns row t" in the interleaved canonical layout;
//           M = 128 (rows 64..127 unused) and M = 64 (to discover its TMEM lane layout).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I.. -o build/adj_probe adj_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../hode_tcgen05.cuh"

using namespace hode;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// instruction descriptor with major-ness bits: a_major bit 15, b_major bit 16 (1 = MN-major)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int M = 128, N = 64, K = 64;

// ---- test 1 -------------------------------------------------------------------------------------
// variant 0: LBO = 128 B (K groups of 8), SBO = N*16 B (MN groups of 4);  variant 1: swapped.
__global__ void __launch_bounds__(128) bwd_probe(const float* __restrict__ A, const float* __restrict__ W,
                                                 float* __restrict__ D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* img = reinterpret_cast<float*>(smem);  // forward image of W[out][in]: (n=out,k=in) at ((k/4)*64+n)*4+k%4
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 64; i += blockDim.x) {
    const int o = i / 64, in = i % 64;
    img[((in >> 2) * 64 + o) * 4 + (in & 3)] = __uint_as_float((__float_as_uint(W[i]) + 0x1000u) & 0xFFFFE000u);
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (__float_as_uint(A[tid * K + half * 32 + j]) + 0x1000u) & 0xFFFFE000u;
    HODE_TMEM_ST_X32(tb + 64 + lane_base + half * 32, v);
  }
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = idesc_tf32(M, N, 0, 1);
    const uint32_t s = tc::smem_u32(img);
    const uint32_t lbo = variant == 0 ? 128u : 1024u, sbo = variant == 0 ? 1024u : 128u;
    // k (= out index) advances by 8 per MMA: 8 rows of 16 B = 128 B in the forward image
    for (int ks = 0; ks < K / 8; ++ks)
      tc::mma_tf32_ts(tb, tb + 64 + ks * 8, tc::make_desc(s + ks * 128, lbo, sbo), idesc, ks ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int half = 0; half < 2; ++half) {
    HODE_TMEM_LD_X32(tb + lane_base + half * 32, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[tid * N + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 128);
}

// ---- test 2 -------------------------------------------------------------------------------------
// Operand staging (both operands): element (f, t) [f = feature 0..63, t = trajectory 0..127] at float
//   (f % 4) + 4 * (t % 8) + SBO_F * (f / 4) + LBO_F * (t / 8),   SBO_F = 32, LBO_F = 16 * 32 + pad
// i.e. core matrix = 8 trajectories x 4 features (128 B); thread t writes its 64 features as 16 float4.
__global__ void __launch_bounds__(128) dw_probe(const float* __restrict__ Dl, const float* __restrict__ Ac,
                                                float* __restrict__ out /*[128 lanes][64]*/, int m_rows, int pad,
                                                int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int LBO_F = 16 * 32 + pad;
  float* sd = reinterpret_cast<float*>(smem);
  float* sa = sd + 16 * LBO_F + 1024;   // slack: M = 128 reads feature groups 16..31 past the real ones
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * (16 * LBO_F + 1024); i += blockDim.x) sd[i] = 0.f;
  __syncthreads();
  for (int g = 0; g < 16; ++g) {
    float4 d, a;
    d.x = Dl[tid * 64 + g * 4 + 0]; d.y = Dl[tid * 64 + g * 4 + 1]; d.z = Dl[tid * 64 + g * 4 + 2]; d.w = Dl[tid * 64 + g * 4 + 3];
    a.x = Ac[tid * 64 + g * 4 + 0]; a.y = Ac[tid * 64 + g * 4 + 1]; a.z = Ac[tid * 64 + g * 4 + 2]; a.w = Ac[tid * 64 + g * 4 + 3];
    auto rnd = [](float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); };
    d.x = rnd(d.x); d.y = rnd(d.y); d.z = rnd(d.z); d.w = rnd(d.w);
    a.x = rnd(a.x); a.y = rnd(a.y); a.z = rnd(a.z); a.w = rnd(a.w);
    const int o = 4 * (tid & 7) + 32 * g + LBO_F * (tid >> 3);
    *reinterpret_cast<float4*>(sd + o) = d;
    *reinterpret_cast<float4*>(sa + o) = a;
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(m_rows, 64, 1, 1);
    const uint32_t a0 = tc::smem_u32(sd), b0 = tc::smem_u32(sa);
    const uint32_t lbo = variant == 0 ? (uint32_t)LBO_F * 4u : 128u, sbo = variant == 0 ? 128u : (uint32_t)LBO_F * 4u;
    for (int ks = 0; ks < 128 / 8; ++ks)   // one K group of 8 trajectories per MMA
      mma_tf32_ss(tb, tc::make_desc(a0 + ks * LBO_F * 4, lbo, sbo), tc::make_desc(b0 + ks * LBO_F * 4, lbo, sbo), idesc,
                  ks ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    HODE_TMEM_LD_X32(tb + lane_base + half * 32, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 64);
}

// ---- test 3: SS form, both operands K-major, written transposed by "thread t owns trajectory t":
// element (f, t) at (t % 4) + 4 * (f % 8) + 32 * (f / 8) + LBO_K * (t / 4), LBO_K = 8 * 32 + pad
__global__ void __launch_bounds__(128) dw_probe_kmajor(const float* __restrict__ Dl, const float* __restrict__ Ac,
                                                       float* __restrict__ out, int m_rows, int pad) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int LBO_K = 8 * 32 + pad;
  float* sd = reinterpret_cast<float*>(smem);
  float* sa = sd + 32 * LBO_K + 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * (32 * LBO_K + 1024); i += blockDim.x) sd[i] = 0.f;
  __syncthreads();
  auto rnd = [](float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); };
  for (int f = 0; f < 64; ++f) {
    const int o = (tid & 3) + 4 * (f & 7) + 32 * (f >> 3) + LBO_K * (tid >> 2);
    sd[o] = rnd(Dl[tid * 64 + f]);
    sa[o] = rnd(Ac[tid * 64 + f]);
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    const uint32_t idesc = idesc_tf32(m_rows, 64, 0, 0);
    const uint32_t a0 = tc::smem_u32(sd), b0 = tc::smem_u32(sa);
    for (int ks = 0; ks < 128 / 8; ++ks)
      mma_tf32_ss(tb, tc::make_desc(a0 + ks * 2 * LBO_K * 4, LBO_K * 4, 128), tc::make_desc(b0 + ks * 2 * LBO_K * 4, LBO_K * 4, 128),
                  idesc, ks ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    HODE_TMEM_LD_X32(tb + lane_base + half * 32, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 64);
}

static float tf32r(float x) {
  uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x;
}

int main() {
  std::vector<float> A(M * K), W(64 * 64), Dl(128 * 64), Ac(128 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : A) x = rnd();
  for (auto& x : W) x = rnd();
  for (auto& x : Dl) x = rnd();
  for (auto& x : Ac) x = rnd();
  float *dA, *dW, *dD, *dDl, *dAc, *dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dDl, Dl.size() * 4)); CK(cudaMalloc(&dAc, Ac.size() * 4)); CK(cudaMalloc(&dO, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dDl, Dl.data(), Dl.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAc, Ac.data(), Ac.size() * 4, cudaMemcpyHostToDevice));

  // ---- test 1
  std::vector<double> ref1(M * N, 0.0);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)tf32r(A[m * K + k]) * (double)tf32r(W[k * 64 + n]);
      ref1[m * N + n] = s;
    }
  for (int variant = 0; variant < 2; ++variant) {
    CK(cudaMemset(dD, 0, M * N * 4));
    bwd_probe<<<1, 128, 64 * 64 * 4>>>(dA, dW, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test1 variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < M * N; ++i) err = fmax(err, fabs(D[i] - ref1[i]));
    printf("test1 (MN-major B = forward image) variant %d (LBO=%s): max abs err %.3e  %s   D[0..3]=%g %g %g %g ref=%g %g %g %g\n", variant,
           variant == 0 ? "128B,SBO=1024B" : "1024B,SBO=128B", err, err < 1e-4 ? "OK" : "MISMATCH", D[0], D[1], D[2], D[3],
           ref1[0], ref1[1], ref1[2], ref1[3]);
  }

  // ---- test 2
  std::vector<double> ref2(64 * 64, 0.0);
  for (int j = 0; j < 64; ++j)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int t = 0; t < 128; ++t) s += (double)tf32r(Dl[t * 64 + j]) * (double)tf32r(Ac[t * 64 + n]);
      ref2[j * 64 + n] = s;
    }
  for (int m_rows : {128, 64})
    for (int pad : {0, 4})
      for (int variant = 0; variant < 2; ++variant) {
        const int LBO_F = 16 * 32 + pad;
        const size_t smem = 2 * (16 * LBO_F + 1024) * 4;
        CK(cudaFuncSetAttribute(dw_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(dO, 0, 128 * 64 * 4));
        dw_probe<<<1, 128, smem>>>(dDl, dAc, dO, m_rows, pad, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test2 M=%d pad=%d variant %d: CUDA error %s\n", m_rows, pad, variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> O(128 * 64);
        CK(cudaMemcpy(O.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        // which TMEM lane holds output row j?
        double err_id = 0;
        for (int j = 0; j < 64; ++j)
          for (int n = 0; n < 64; ++n) err_id = fmax(err_id, fabs(O[j * 64 + n] - ref2[j * 64 + n]));
        printf("test2 (SS, MN-major A and B) M=%d pad=%d variant %d: rows 0..63 in lanes 0..63: max abs err %.3e %s O[0..1]=%g %g ref %g %g\n",
               m_rows, pad, variant, err_id, err_id < 2e-4 ? "OK" : "MISMATCH", O[0], O[1], ref2[0], ref2[1]);
        if (false) {
          // search the lane of a few rows
          for (int j : {0, 1, 16, 17, 32, 48, 63}) {
            int best = -1; double be = 1e30;
            for (int lane = 0; lane < 128; ++lane) {
              double e2 = 0;
              for (int n = 0; n < 64; ++n) e2 = fmax(e2, fabs(O[lane * 64 + n] - ref2[j * 64 + n]));
              if (e2 < be) { be = e2; best = lane; }
            }
            printf("    row %2d best matches lane %3d (err %.2e)\n", j, best, be);
          }
        }
      }
  for (int m_rows : {128, 64})
    for (int pad : {0, 4}) {
      const int LBO_K = 8 * 32 + pad;
      const size_t smem = 2 * (32 * LBO_K + 1024) * 4;
      CK(cudaFuncSetAttribute(dw_probe_kmajor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaMemset(dO, 0, 128 * 64 * 4));
      dw_probe_kmajor<<<1, 128, smem>>>(dDl, dAc, dO, m_rows, pad);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("test3 M=%d pad=%d: CUDA error %s\n", m_rows, pad, cudaGetErrorString(e)); return 1; }
      std::vector<float> O(128 * 64);
      CK(cudaMemcpy(O.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost));
      double err_id = 0;
      for (int j = 0; j < 64; ++j)
        for (int n = 0; n < 64; ++n) err_id = fmax(err_id, fabs(O[j * 64 + n] - ref2[j * 64 + n]));
      printf("test3 (SS, K-major A and B, transposed staging) M=%d pad=%d: rows 0..63 in lanes 0..63: max abs err %.3e %s  O[0..1]=%g %g ref %g %g\n",
             m_rows, pad, err_id, err_id < 2e-4 ? "OK" : "MISMATCH", O[0], O[1], ref2[0], ref2[1]);
      if (err_id >= 2e-4) {
        for (int j : {0, 1, 16, 17, 32, 48, 63}) {
          int best = -1; double be = 1e30;
          for (int lane = 0; lane < 128; ++lane) {
            double e2 = 0;
            for (int n = 0; n < 64; ++n) e2 = fmax(e2, fabs(O[lane * 64 + n] - ref2[j * 64 + n]));
            if (e2 < be) { be = e2; best = lane; }
          }
          printf("    row %2d best matches lane %3d (err %.2e)\n", j, best, be);
        }
      }
    }
  return 0;
}
