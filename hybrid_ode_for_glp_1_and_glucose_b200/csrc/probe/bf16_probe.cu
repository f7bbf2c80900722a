// bf16_probe.cu — hardware probe for the BF16 weight-gradient path of the tensor-core adjoint
// (not part of libhode.so).
//
// dW = delta^T [a | 1] contracts over the 128 trajectories of a tile.  kind::f16 (BF16 operands,
// FP32 accumulation) accepts MN-major shared-memory operands, so "thread t owns trajectory t"
// writes 8 consecutive features as ONE 16-byte vector and no transposed staging is needed:
//   element (t, f) at byte  (f / 8) * 2048 + t * 16 + (f % 8) * 2        (core = 8 t x 8 f = 128 B)
// This probe checks on a real B200
//   test 1  layout: single pass on BF16-exact inputs, M = 64 / N = 72 and M = 128 / N = 16, both
//           LBO/SBO assignments, against a float64 reference;
//   test 2  precision of the two-term split x ~= hi + mid (3 passes: mid*hi + hi*mid + hi*hi) on
//           FP32 inputs: max |err| / sum|a b|.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bf16_probe bf16_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../hode_tcgen05.cuh"

using namespace hode;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// kind::f16 instruction descriptor: D = f32 [4,6) = 1, A = B = BF16 ([7,10) = [10,13) = 1),
// a_major bit 15, b_major bit 16 (1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// x[0..7] -> 8 BF16 hi (round to nearest) and 8 BF16 mid = bf16(x - hi), feature 0 in the low half of word 0
__device__ __forceinline__ void bf16_split8(const float* v, uint4& hi, uint4& mid) {
  uint32_t h[4], m[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[q]) : "f"(v[2 * q + 1]), "f"(v[2 * q]));
    const float r0 = v[2 * q] - __uint_as_float(h[q] << 16);
    const float r1 = v[2 * q + 1] - __uint_as_float(h[q] & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m[q]) : "f"(r1), "f"(r0));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  mid = make_uint4(m[0], m[1], m[2], m[3]);
}

constexpr int GRP = 2048;                 // bytes per 8-feature group: 128 trajectories x 16 B
constexpr int D_GROUPS = 8, A_GROUPS = 9; // delta: 64 features; inputs: 64 + the constant-1 group
constexpr int PART_D = 16 * GRP, PART_A = 16 * GRP;   // 16 groups of room: M = 128 reads 16 groups

// Dl, Ac: [128][64] fp32.  out: [128 lanes][80] accumulator columns.  passes: 1 (hi only) or 3.
// mode 0: M = m_rows rows = delta features, N = 72 columns = input features (+ const);  mode 1: M = 128 rows = input
// features (transposed product), N = 16 columns = delta features 0..15.
__global__ void __launch_bounds__(128) dw_bf16_probe(const float* __restrict__ Dl, const float* __restrict__ Ac,
                                                     float* __restrict__ out, int mode, int m_rows, int passes, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sd_hi = smem;
  uint8_t* sd_mid = sd_hi + PART_D;
  uint8_t* sa_hi = sd_mid + PART_D;
  uint8_t* sa_mid = sa_hi + PART_A;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (2 * PART_D + 2 * PART_A) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int g = 0; g < 8; ++g) {
    uint4 hi, mid;
    bf16_split8(Dl + tid * 64 + g * 8, hi, mid);
    *reinterpret_cast<uint4*>(sd_hi + g * GRP + tid * 16) = hi;
    *reinterpret_cast<uint4*>(sd_mid + g * GRP + tid * 16) = mid;
    bf16_split8(Ac + tid * 64 + g * 8, hi, mid);
    *reinterpret_cast<uint4*>(sa_hi + g * GRP + tid * 16) = hi;
    *reinterpret_cast<uint4*>(sa_mid + g * GRP + tid * 16) = mid;
  }
  *reinterpret_cast<uint4*>(sa_hi + 8 * GRP + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);   // feature 64 = 1
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    const uint32_t lbo = variant == 0 ? 128u : (uint32_t)GRP, sbo = variant == 0 ? (uint32_t)GRP : 128u;
    const uint32_t idesc = mode == 0 ? idesc_bf16(m_rows, 72, 1, 1) : idesc_bf16(128, 16, 1, 1);
    const uint32_t r_hi = tc::smem_u32(mode == 0 ? sd_hi : sa_hi), r_mid = tc::smem_u32(mode == 0 ? sd_mid : sa_mid);
    const uint32_t c_hi = tc::smem_u32(mode == 0 ? sa_hi : sd_hi), c_mid = tc::smem_u32(mode == 0 ? sa_mid : sd_mid);
    uint32_t acc = 0u;
    // one MMA contracts 16 trajectories = 2 K-groups of 128 B
    if (passes == 3) {
      for (int ks = 0; ks < 8; ++ks, acc = 1u)
        mma_bf16_ss(tb, tc::make_desc(r_mid + ks * 256, lbo, sbo), tc::make_desc(c_hi + ks * 256, lbo, sbo), idesc, acc);
      for (int ks = 0; ks < 8; ++ks)
        mma_bf16_ss(tb, tc::make_desc(r_hi + ks * 256, lbo, sbo), tc::make_desc(c_mid + ks * 256, lbo, sbo), idesc, 1u);
    }
    for (int ks = 0; ks < 8; ++ks, acc = 1u)
      mma_bf16_ss(tb, tc::make_desc(r_hi + ks * 256, lbo, sbo), tc::make_desc(c_hi + ks * 256, lbo, sbo), idesc, acc);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[16];
  for (int c16 = 0; c16 < 5; ++c16) {
    HODE_TMEM_LD_X16(tb + lane_base + c16 * 16, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[tid * 80 + c16 * 16 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 128);
}

// ---- test 3: the SAME delta image read as a K-major A operand (M = 128 trajectories, K = 64 features):
// u[t][n] = sum_f delta[t][f] * Wt[n][f] with B = W^T image [N][K] K-major: element (n, k) at byte
// (k / 8) * (N * 16) + n * 16 + (k % 8) * 2.  variant 0: LBO = K-chunk stride, SBO = 128 B; variant 1: swapped.
__global__ void __launch_bounds__(128) u_bf16_probe(const float* __restrict__ Dl, const float* __restrict__ Wt /*[N=64][K=64]*/,
                                                    float* __restrict__ out /*[128][64]*/, int passes, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sd_hi = smem;
  uint8_t* sd_mid = sd_hi + 8 * GRP;
  uint8_t* sw_hi = sd_mid + 8 * GRP;     // 8 K-chunks x (64 n x 16 B) = 8 KB
  uint8_t* sw_mid = sw_hi + 8 * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int g = 0; g < 8; ++g) {
    uint4 hi, mid;
    bf16_split8(Dl + tid * 64 + g * 8, hi, mid);
    *reinterpret_cast<uint4*>(sd_hi + g * GRP + tid * 16) = hi;
    *reinterpret_cast<uint4*>(sd_mid + g * GRP + tid * 16) = mid;
  }
  if (tid < 64) {
    for (int kc = 0; kc < 8; ++kc) {
      uint4 hi, mid;
      bf16_split8(Wt + tid * 64 + kc * 8, hi, mid);
      *reinterpret_cast<uint4*>(sw_hi + kc * 1024 + tid * 16) = hi;
      *reinterpret_cast<uint4*>(sw_mid + kc * 1024 + tid * 16) = mid;
    }
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
    const uint32_t a_lbo = variant == 0 ? (uint32_t)GRP : 128u, a_sbo = variant == 0 ? 128u : (uint32_t)GRP;
    const uint32_t b_lbo = variant == 0 ? 1024u : 128u, b_sbo = variant == 0 ? 128u : 1024u;
    const uint32_t a_hi = tc::smem_u32(sd_hi), a_mid = tc::smem_u32(sd_mid), b_hi = tc::smem_u32(sw_hi), b_mid = tc::smem_u32(sw_mid);
    uint32_t acc = 0u;
    // one MMA = K 16 = 2 K-chunks: A advances 2 * GRP, B advances 2 * 1024
    if (passes == 3) {
      for (int ks = 0; ks < 4; ++ks, acc = 1u)
        mma_bf16_ss(tb, tc::make_desc(a_mid + ks * 2 * GRP, a_lbo, a_sbo), tc::make_desc(b_hi + ks * 2048, b_lbo, b_sbo), idesc, acc);
      for (int ks = 0; ks < 4; ++ks)
        mma_bf16_ss(tb, tc::make_desc(a_hi + ks * 2 * GRP, a_lbo, a_sbo), tc::make_desc(b_mid + ks * 2048, b_lbo, b_sbo), idesc, 1u);
    }
    for (int ks = 0; ks < 4; ++ks, acc = 1u)
      mma_bf16_ss(tb, tc::make_desc(a_hi + ks * 2 * GRP, a_lbo, a_sbo), tc::make_desc(b_hi + ks * 2048, b_lbo, b_sbo), idesc, acc);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[16];
  for (int c16 = 0; c16 < 4; ++c16) {
    HODE_TMEM_LD_X16(tb + lane_base + c16 * 16, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[tid * 64 + c16 * 16 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 64);
}

static float bf16r(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  memcpy(&x, &u, 4); return x;
}

int main() {
  std::vector<float> Dl(128 * 64), Ac(128 * 64), Dr(128 * 64), Ar(128 * 64);
  srand(1);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& x : Dl) x = rnd() * 1e-3f;
  for (auto& x : Ac) x = fmaxf(rnd(), 0.f) * 37.f;
  for (size_t i = 0; i < Dl.size(); ++i) { Dr[i] = bf16r(Dl[i]); Ar[i] = bf16r(Ac[i]); }
  float *dDl, *dAc, *dO;
  CK(cudaMalloc(&dDl, Dl.size() * 4)); CK(cudaMalloc(&dAc, Ac.size() * 4)); CK(cudaMalloc(&dO, 128 * 80 * 4));
  const size_t smem = 2 * PART_D + 2 * PART_A;
  CK(cudaFuncSetAttribute(dw_bf16_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> O(128 * 80);

  auto ref = [&](const std::vector<float>& D_, const std::vector<float>& A_, int j, int k, double* l1) {
    double s = 0, a1 = 0;   // sum_t D[t][j] * (k < 64 ? A[t][k] : k == 64 ? 1 : 0)
    for (int t = 0; t < 128; ++t) {
      const double a = k < 64 ? A_[t * 64 + k] : (k == 64 ? 1.0 : 0.0);
      s += (double)D_[t * 64 + j] * a;
      a1 += fabs((double)D_[t * 64 + j] * a);
    }
    if (l1) *l1 = a1;
    return s;
  };
  // row j of an M = 64 accumulator lives in TMEM lane 32 (j / 16) + j % 16; M = 128: lane = row
  auto lane_of = [](int m_rows, int j) { return m_rows == 64 ? 32 * (j / 16) + j % 16 : j; };

  for (int passes : {1, 3}) {
    const std::vector<float>& Din = passes == 1 ? Dr : Dl;
    const std::vector<float>& Ain = passes == 1 ? Ar : Ac;
    CK(cudaMemcpy(dDl, Din.data(), Din.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAc, Ain.data(), Ain.size() * 4, cudaMemcpyHostToDevice));
    for (int variant = 0; variant < 2; ++variant) {
      for (int m_rows : {64, 128}) {
        CK(cudaMemset(dO, 0, 128 * 80 * 4));
        dw_bf16_probe<<<1, 128, smem>>>(dDl, dAc, dO, 0, m_rows, passes, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode 0 M=%d variant %d: CUDA error %s\n", m_rows, variant, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0, worst_abs = 0;
        for (int j = 0; j < 64; ++j)
          for (int k = 0; k < 72; ++k) {
            double l1;
            const double r = ref(Din, Ain, j, k, &l1);
            const double err = fabs(O[lane_of(m_rows, j) * 80 + k] - r);
            worst_abs = fmax(worst_abs, err);
            if (l1 > 0) worst = fmax(worst, err / l1);
          }
        printf("passes %d  D[64 x 72] = delta^T [a|1]  M=%3d  variant %d (LBO=%4d SBO=%4d): max err/sum|ab| %.3e  max abs %.3e\n", passes,
               m_rows, variant, variant == 0 ? 128 : GRP, variant == 0 ? GRP : 128, worst, worst_abs);
      }
      {   // transposed product: rows = input features (M = 128, 65 used), columns = delta features 0..15
        CK(cudaMemset(dO, 0, 128 * 80 * 4));
        dw_bf16_probe<<<1, 128, smem>>>(dDl, dAc, dO, 1, 128, passes, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode 1 variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int k = 0; k < 65; ++k)
          for (int j = 0; j < 16; ++j) {
            double l1;
            const double r = ref(Din, Ain, j, k, &l1);
            if (l1 > 0) worst = fmax(worst, fabs(O[k * 80 + j] - r) / l1);
          }
        printf("passes %d  D[128 x 16] = [a|1]^T delta  variant %d: max err/sum|ab| %.3e\n", passes, variant, worst);
      }
    }
  }
  // ---- test 3: K-major A (the delta image) x K-major B (W^T image)
  {
    std::vector<float> Wt(64 * 64), Wr(64 * 64);
    for (auto& x : Wt) x = rnd() * 0.3f;
    for (size_t i = 0; i < Wt.size(); ++i) Wr[i] = bf16r(Wt[i]);
    float *dWt, *dU;
    CK(cudaMalloc(&dWt, Wt.size() * 4)); CK(cudaMalloc(&dU, 128 * 64 * 4));
    const size_t smem3 = 2 * 8 * GRP + 2 * 8 * 1024;
    CK(cudaFuncSetAttribute(u_bf16_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    std::vector<float> U(128 * 64);
    for (int passes : {1, 3}) {
      const std::vector<float>& Din = passes == 1 ? Dr : Dl;
      const std::vector<float>& Win = passes == 1 ? Wr : Wt;
      CK(cudaMemcpy(dDl, Din.data(), Din.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dWt, Win.data(), Win.size() * 4, cudaMemcpyHostToDevice));
      for (int variant = 0; variant < 2; ++variant) {
        CK(cudaMemset(dU, 0, 128 * 64 * 4));
        u_bf16_probe<<<1, 128, smem3>>>(dDl, dWt, dU, passes, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test3 variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(U.data(), dU, U.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int t = 0; t < 128; ++t)
          for (int n = 0; n < 64; ++n) {
            double s = 0, l1 = 0;
            for (int f = 0; f < 64; ++f) { const double p = (double)Din[t * 64 + f] * (double)Win[n * 64 + f]; s += p; l1 += fabs(p); }
            if (l1 > 0) worst = fmax(worst, fabs(U[t * 64 + n] - s) / l1);
          }
        printf("passes %d  u[128 x 64] = delta W (K-major A and B)  variant %d: max err/sum|ab| %.3e\n", passes, variant, worst);
      }
    }
  }
  return 0;
}
