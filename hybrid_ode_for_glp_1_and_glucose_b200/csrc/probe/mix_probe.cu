// mix_probe.cu — round-2 hardware probe (not part of libhode.so).  Answers, on a real B200:
//   test 1  accuracy of the MIXED split-precision product the rollout's MLP uses from round 2 on:
//             D = A_hi(tf32) B_hi(tf32)  +  bf16(A_lo) bf16(B_hi)  +  bf16(A_hi) bf16(B_lo)
//           (one kind::tf32 pass + two kind::f16 BF16 passes at twice the TF32 rate, all three into the same
//           FP32 accumulator, A operands in TMEM) against float64, next to the 3xTF32 product of round 1;
//           also pins the packing of BF16 A operands in TMEM (element 2c in the low half of column c);
//   test 2  cycles per tcgen05.mma in long chains for the shapes the kernels issue (TS/SS, tf32/bf16,
//           N = 64/32/16, M = 64 N = 72 MN-major) -> the tensor-time model in DESIGN.md;
//   test 3  accumulators at TMEM column offsets that are multiples of 8 but not of 16/32 (the adjoint's
//           round-2 column map packs M = 64 / N = 72 accumulators at a 72-column stride).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mix_probe mix_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../hode_tcgen05.cuh"

using namespace hode;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// fmt: 0 = f16, 1 = bf16, 2 = tf32
__host__ __device__ constexpr uint32_t idesc_of(int fmt, int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

constexpr int M = 128, N = 64, K = 64;

// ---------------------------------------------------------------------------------------------------
// test 1.  mode 0: 3xTF32 (lo blocks first);  mode 1: mixed, corrections first;  mode 2: mixed, hi*hi first;
// mode 3: mixed with the BF16 A operand packed the other way round (must be WRONG: pins the packing).
// d_col: accumulator column offset inside the allocation.
// shared memory: [B_hi tf32: 16 chunks x 64 n x 4][B_hib bf16: 8 chunks x 64 n x 8][B_lob bf16][B_lo tf32]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mix_gemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                int mode, int d_col) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* Bhi = reinterpret_cast<float*>(smem);                    // 16 KB
  uint16_t* Bhib = reinterpret_cast<uint16_t*>(smem + 16384);     // 8 KB
  uint16_t* Blob = reinterpret_cast<uint16_t*>(smem + 24576);     // 8 KB
  float* Blo = reinterpret_cast<float*>(smem + 32768);            // 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const float w = B[i];
    uint32_t h, l;
    tc::split_tf32(w, h, l);
    const int o = ((k >> 2) * N + n) * 4 + (k & 3);
    Bhi[o] = __uint_as_float(h);
    Blo[o] = __uint_as_float(l);
    const float wl = w - __uint_as_float(h);
    // bf16 K-major image: element (n, k) at 2-byte index (k / 8) * (N * 8) + n * 8 + k % 8
    const int ob = (k >> 3) * (N * 8) + n * 8 + (k & 7);
    Bhib[ob] = (uint16_t)(pack_bf16(__uint_as_float(h), 0.f) & 0xFFFFu);
    Blob[ob] = (uint16_t)(pack_bf16(wl, 0.f) & 0xFFFFu);
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t tD = tb + (uint32_t)d_col, tAhi = tb + 320, tAlo = tb + 384, tAhib = tb + 448, tAlob = tb + 480;
  {
    uint32_t hi[64], lo[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) tc::split_tf32(A[tid * K + j], hi[j], lo[j]);
    HODE_TMEM_ST_X32(tAhi + lane_base, hi);
    HODE_TMEM_ST_X32(tAhi + lane_base + 32, (hi + 32));
    HODE_TMEM_ST_X32(tAlo + lane_base, lo);
    HODE_TMEM_ST_X32(tAlo + lane_base + 32, (lo + 32));
    uint32_t hb[32], lb[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float a0 = A[tid * K + 2 * c], a1 = A[tid * K + 2 * c + 1];
      const float h0 = __uint_as_float(hi[2 * c]), h1 = __uint_as_float(hi[2 * c + 1]);
      if (mode == 3) { hb[c] = pack_bf16(h1, h0); lb[c] = pack_bf16(a1 - h1, a0 - h0); }
      else { hb[c] = pack_bf16(h0, h1); lb[c] = pack_bf16(a0 - h0, a1 - h1); }
    }
    HODE_TMEM_ST_X32(tAhib + lane_base, hb);
    HODE_TMEM_ST_X32(tAlob + lane_base, lb);
  }
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t it = idesc_of(2, M, N, 0, 0), ib = idesc_of(1, M, N, 0, 0);
    const uint32_t bhi = tc::smem_u32(Bhi), blo = tc::smem_u32(Blo), bhib = tc::smem_u32(Bhib), blob = tc::smem_u32(Blob);
    const uint32_t LBO = N * 16, SBO = 128;
    uint32_t acc = 0;
    auto tf = [&](uint32_t a, uint32_t b) {
      for (int ks = 0; ks < K / 8; ++ks) { tc::mma_tf32_ts(tD, a + ks * 8, tc::make_desc(b + ks * 2 * LBO, LBO, SBO), it, acc); acc = 1; }
    };
    auto bf = [&](uint32_t a, uint32_t b) {   // one MMA = K 16 = 8 TMEM columns of A, 2 K-chunks of B
      for (int ks = 0; ks < K / 16; ++ks) { mma_f16_ts(tD, a + ks * 8, tc::make_desc(b + ks * 2 * LBO, LBO, SBO), ib, acc); acc = 1; }
    };
    if (mode == 0) { tf(tAlo, bhi); tf(tAhi, blo); tf(tAhi, bhi); }
    else if (mode == 2) { tf(tAhi, bhi); bf(tAlob, bhib); bf(tAhib, blob); }
    else { bf(tAlob, bhib); bf(tAhib, blob); tf(tAhi, bhi); }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {
    HODE_TMEM_LD_X32(tD + lane_base + half * 32, r);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[tid * N + half * 32 + j] = __uint_as_float(r[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------
// test 2: cycles per MMA.  cfg: fmt (1 bf16 / 2 tf32), a_src (0 TMEM, 1 smem K-major, 2 smem MN-major (B too)),
// Mm, Nn, chain length, number of accumulators the chain rotates over.
// ---------------------------------------------------------------------------------------------------
struct ChainCfg { int fmt, a_src, Mm, Nn, n, n_d, alt; };

// templated + unrolled: the first version of this probe looped over runtime configuration and measured its own
// scalar overhead (146 cycles per MMA for every shape)
template <int FMT, int ASRC, int MM, int NN, int ND, int ALT>
__global__ void __launch_bounds__(128) chain_time(int n_outer, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0x3C003C00u;
    for (int cc = 256; cc < 512; cc += 32) HODE_TMEM_ST_X32(tb + lane_base + cc, v);
    tc::wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    if (tc::elect_one()) {
      tc::fence_after_sync();
      const uint32_t s0 = tc::smem_u32(smem);
      constexpr uint32_t id = idesc_of(FMT, MM, NN, ASRC == 2, ASRC == 2);
      constexpr uint32_t id_alt = idesc_of(FMT == 2 ? 1 : 2, MM, NN, 0, 0);
      const uint64_t bdesc = ASRC == 2 ? tc::make_desc(s0 + 32768, 128u, 2048u) : tc::make_desc(s0 + 32768, (uint32_t)NN * 16u, 128u);
      const uint64_t adesc = ASRC == 2 ? tc::make_desc(s0, 128u, 2048u) : tc::make_desc(s0, 2048u, 128u);
      const long long t0 = clock64();
      for (int o = 0; o < n_outer; ++o) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t d = tb + (uint32_t)((i % ND) * 80);
          const uint32_t acc = (o > 0 || i >= ND) ? 1u : 0u;
          const bool alt = ALT && (i & 1);
          if (ASRC == 0) {
            if ((FMT == 2) != alt) tc::mma_tf32_ts(d, tb + 256 + (i & 7) * 8, bdesc, alt ? id_alt : id, acc);
            else mma_f16_ts(d, tb + 256 + (i & 7) * 8, bdesc, alt ? id_alt : id, acc);
          } else {
            if (FMT == 2) mma_tf32_ss(d, adesc, bdesc, id, acc);
            else mma_f16_ss(d, adesc, bdesc, id, acc);
          }
        }
      }
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0);
  } else {
    tc::mbar_wait(&bar, 0);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int FMT, int ASRC, int MM, int NN, int ND, int ALT>
void run_chain(const char* name, long long* dT) {
  CK(cudaFuncSetAttribute(chain_time<FMT, ASRC, MM, NN, ND, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int n_outer = 16, n = 32 * n_outer;
  for (int grid : {1, 148}) {
    chain_time<FMT, ASRC, MM, NN, ND, ALT><<<grid, 128, 65536>>>(n_outer, dT);
    CK(cudaDeviceSynchronize());
    chain_time<FMT, ASRC, MM, NN, ND, ALT><<<grid, 128, 65536>>>(n_outer, dT);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), dT, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    printf("test2 %-32s grid %3d: %.1f cycles/MMA (median CTA; min %.1f max %.1f)\n", name, grid,
           (double)h[grid / 2] / n, (double)h[0] / n, (double)h[grid - 1] / n);
  }
}

// ---------------------------------------------------------------------------------------------------
// test 3: M = 64, N = 72 MN-major SS accumulator (the weight-gradient form) at column offset d_col
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) dcol_probe(const float* __restrict__ Dl, const float* __restrict__ Ac, float* __restrict__ out,
                                                  int d_col) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int GRP = 2048;
  uint8_t* sd = smem;                 // delta, 16 groups of room
  uint8_t* sa = smem + 16 * GRP;      // inputs
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32 * GRP) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int g = 0; g < 8; ++g) {
    uint32_t h[4], a[4];
    for (int q = 0; q < 4; ++q) {
      h[q] = pack_bf16(Dl[tid * 64 + g * 8 + 2 * q], Dl[tid * 64 + g * 8 + 2 * q + 1]);
      a[q] = pack_bf16(Ac[tid * 64 + g * 8 + 2 * q], Ac[tid * 64 + g * 8 + 2 * q + 1]);
    }
    *reinterpret_cast<uint4*>(sd + g * GRP + tid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sa + g * GRP + tid * 16) = make_uint4(a[0], a[1], a[2], a[3]);
  }
  *reinterpret_cast<uint4*>(sa + 8 * GRP + tid * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  {  // poison the neighbourhood so that a misplaced accumulator shows
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0x7FC00000u;
    for (int cc = 0; cc < 512; cc += 32) HODE_TMEM_ST_X32(tb + lane_base + cc, v);
    tc::wait_st();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t id = idesc_of(1, 64, 72, 1, 1);
    const uint32_t r = tc::smem_u32(sd), cc = tc::smem_u32(sa);
    for (int ks = 0; ks < 8; ++ks)
      mma_f16_ss(tb + (uint32_t)d_col, tc::make_desc(r + ks * 256, 128u, GRP), tc::make_desc(cc + ks * 256, 128u, GRP), id, ks ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[8];
  for (int c8 = 0; c8 < 9; ++c8) {
    HODE_TMEM_LD_X8(tb + lane_base + (uint32_t)d_col + c8 * 8, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) out[tid * 72 + c8 * 8 + j] = __uint_as_float(v[j]);
  }
  // the 8 columns right after the accumulator must still be poison
  HODE_TMEM_LD_X8(tb + lane_base + (uint32_t)d_col + 72, v);
  tc::wait_ld();
  out[128 * 72 + tid] = __uint_as_float(v[0]);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tb, 512);
}

int main() {
  std::vector<float> hA(M * K), hB(N * K), hD(M * N);
  srand(1);
  for (auto& x : hA) x = fabsf((float)rand() / RAND_MAX * 200.f - 50.f) * 0.01f;
  for (auto& x : hB) x = ((float)rand() / RAND_MAX - 0.5f) * 0.5f;
  float *dA, *dB, *dD, *dA2;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, (128 * 72 + 128) * 4));
  {
    std::vector<float> hA2(M * K);
    for (auto& x : hA2) x = ((float)rand() / RAND_MAX - 0.5f) * 2.f;
    CK(cudaMalloc(&dA2, hA2.size() * 4));
    CK(cudaMemcpy(dA2, hA2.data(), hA2.size() * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(mix_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
  const char* names[4] = {"3xTF32 (round 1)", "mixed, corrections first", "mixed, hi*hi first", "mixed, A pairs swapped (must fail)"};
  for (int dcol : {0, 200}) {
    for (int mode = 0; mode < 4; ++mode) {
      CK(cudaMemset(dD, 0, hD.size() * 4));
      mix_gemm<<<1, 128, 49152>>>(dA, dB, dD, mode, dcol);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double max_rel = 0, sum_abs = 0, max_f32 = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0, mag = 0;
          float r32 = 0.f;
          for (int k = 0; k < K; ++k) {
            ref += (double)hA[m * K + k] * hB[n * K + k];
            mag += fabs((double)hA[m * K + k] * hB[n * K + k]);
            r32 = fmaf(hA[m * K + k], hB[n * K + k], r32);
          }
          const double err = fabs(hD[m * N + n] - ref);
          max_rel = fmax(max_rel, err / mag);
          sum_abs += err / mag;
          max_f32 = fmax(max_f32, fabs(r32 - ref) / mag);
        }
      printf("test1 d_col %3d mode %d (%-34s): max err/sum|ab| %.3e  mean %.3e  (fp32 fma chain max %.3e)\n", dcol, mode,
             names[mode], max_rel, sum_abs / (M * N), max_f32);
    }
  }

  // ---- test 2 ----
  long long* dT;
  CK(cudaMalloc(&dT, 148 * sizeof(long long)));
  run_chain<2, 0, 128, 64, 1, 0>("tf32 TS M128 N64 K8  same D", dT);
  run_chain<2, 0, 128, 64, 2, 0>("tf32 TS M128 N64 K8  2 D", dT);
  run_chain<1, 0, 128, 64, 1, 0>("bf16 TS M128 N64 K16 same D", dT);
  run_chain<2, 0, 128, 64, 1, 1>("tf32/bf16 alternating TS N64", dT);
  run_chain<2, 0, 128, 32, 1, 0>("tf32 TS M128 N32", dT);
  run_chain<1, 0, 128, 32, 1, 0>("bf16 TS M128 N32", dT);
  run_chain<2, 0, 128, 16, 1, 0>("tf32 TS M128 N16", dT);
  run_chain<1, 0, 128, 16, 1, 0>("bf16 TS M128 N16", dT);
  run_chain<2, 0, 128, 128, 1, 0>("tf32 TS M128 N128", dT);
  run_chain<1, 0, 128, 128, 1, 0>("bf16 TS M128 N128", dT);
  run_chain<2, 1, 128, 64, 1, 0>("tf32 SS K-major M128 N64", dT);
  run_chain<1, 1, 128, 64, 1, 0>("bf16 SS K-major M128 N64", dT);
  run_chain<1, 1, 128, 16, 1, 0>("bf16 SS K-major M128 N16", dT);
  run_chain<1, 2, 64, 72, 1, 0>("bf16 SS MN-major M64 N72", dT);
  run_chain<1, 2, 64, 72, 3, 0>("bf16 SS MN-major M64 N72 3 D", dT);
  run_chain<1, 2, 64, 16, 1, 0>("bf16 SS MN-major M64 N16", dT);
  run_chain<1, 2, 128, 16, 1, 0>("bf16 SS MN-major M128 N16", dT);
  run_chain<1, 2, 128, 64, 1, 0>("bf16 SS MN-major M128 N64", dT);
  run_chain<1, 2, 128, 80, 1, 0>("bf16 SS MN-major M128 N80", dT);

  // ---- test 3 ----
  CK(cudaFuncSetAttribute(dcol_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  std::vector<float> ref(128 * 72 + 128), got(128 * 72 + 128);
  for (int dcol : {0, 256, 264, 328, 400, 432}) {
    CK(cudaMemset(dD, 0, (128 * 72 + 128) * 4));
    dcol_probe<<<1, 128, 65536>>>(dA, dA2, dD, dcol);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
    if (dcol == 0) { ref = got; }
    int bad = 0, nan_after = 0;
    for (int t = 0; t < 128; ++t) {
      const bool owner = (t & 31) < 16;   // M = 64: rows live in lanes 32 (j / 16) + j % 16
      if (!owner) continue;
      for (int cidx = 0; cidx < 72; ++cidx)
        if (memcmp(&ref[t * 72 + cidx], &got[t * 72 + cidx], 4) != 0) ++bad;
      if (got[128 * 72 + t] != got[128 * 72 + t]) ++nan_after;
    }
    printf("test3 M64 N72 accumulator at column %3d: %d mismatching elements vs column 0; poison intact after it in %d/64 rows\n",
           dcol, bad, nan_after);
  }
  printf("probe done\n");
  return 0;
}
