// f16_probe.cu — round-2 hardware probe (not part of libhode.so).  Answers, on a real B200, whether the residual
// MLP's split-precision product can run entirely on kind::f16 MMAs (twice the TF32 rate):
//   D = bf16(A_lo) X(B_hi) + f16(A_hi) bf16(B_lo) + f16(A_hi) f16(B_hi),   A_hi = fp16(a) (round to nearest,
//   saturating), A_lo = a - A_hi (exact in FP32), B likewise
//   mode 0  the current default (MLP_MIX3): bf16(A_lo) bf16(B_hi) + A_hi B_lo + A_hi B_hi with TF32 main terms
//   mode 1  what HODE_MLP_F16BF16X2 issues: bf16(A_lo) bf16(B_hi) + f16(A_hi / 64) f16(64 B_lo) + f16(A_hi) f16(B_hi)
//   mode 2  FP16 main term, both cross terms as BF16 x BF16 (B_lo rounded to 8 bits: a fixed 2^-20 weight perturbation)
//   mode 3  FP16 main term, cross terms as MIXED-FORMAT instructions (a_format != b_format in one kind::f16
//           descriptor): [A = bf16(A_lo)] x [B = f16(B_hi)] and [A = f16(A_hi)] x [B = bf16(B_lo)] — no copies.
//           MEASURED: raises "an illegal instruction was encountered" on B200 -> only run with `f16_probe mixed`
// on four operand ranges (normal, FP16-subnormal activations, large activations, activations past 65504), plus the
// cycles per MMA of FP16 and mixed-format chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f16_probe f16_probe.cu
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

// formats: 0 = f16, 1 = bf16, 2 = tf32
__host__ __device__ constexpr uint32_t idesc_ab(int fa, int fb, int M, int N) {
  return (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t p) {
  float2 r;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(p));
  return r;
}

constexpr int M = 128, N = 64, K = 64;

// shared memory (bytes): [B_hi tf32 16K][B_lo tf32 16K][bf16(B_hi tf32) 8K][f16(B) 8K][bf16(f16(B)) 8K][bf16(B - f16(B)) 8K]
//                        [f16(64 (B - f16(B))) 8K]
__global__ void __launch_bounds__(128) f16_gemm(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* Bhi = reinterpret_cast<float*>(smem);
  float* Blo = reinterpret_cast<float*>(smem + 16384);
  uint16_t* Bhib = reinterpret_cast<uint16_t*>(smem + 32768);
  uint16_t* Bh16 = reinterpret_cast<uint16_t*>(smem + 40960);
  uint16_t* Bh16b = reinterpret_cast<uint16_t*>(smem + 49152);
  uint16_t* Bl16b = reinterpret_cast<uint16_t*>(smem + 57344);
  uint16_t* Bl16s = reinterpret_cast<uint16_t*>(smem + 65536);
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
    const int ob = (k >> 3) * (N * 8) + n * 8 + (k & 7);
    Bhib[ob] = (uint16_t)(pack_bf16(__uint_as_float(h), 0.f) & 0xFFFFu);
    const uint32_t h16 = pack_f16(w, 0.f) & 0xFFFFu;
    const float wh = unpack_f16(h16).x;
    Bh16[ob] = (uint16_t)h16;
    Bh16b[ob] = (uint16_t)(pack_bf16(wh, 0.f) & 0xFFFFu);
    Bl16b[ob] = (uint16_t)(pack_bf16(w - wh, 0.f) & 0xFFFFu);
    Bl16s[ob] = (uint16_t)(pack_f16((w - wh) * 64.f, 0.f) & 0xFFFFu);
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tb = tmem_base_s, lane_base = (uint32_t)(warp * 32) << 16;
  // columns: D [0,64) | A_hi tf32 [64,128) | bf16(a - A_hi tf32) [128,160) | f16(a) [160,192) | bf16(f16(a)) [192,224) | bf16(a - f16(a)) [224,256)
  //          | f16(a) / 64 [256,288)
  const uint32_t tD = tb, tAhi = tb + 64, tAlb = tb + 128, tA16 = tb + 160, tA16b = tb + 192, tAl16b = tb + 224, tA16s = tb + 256;
  {
    uint32_t hi[64], lb[32], a16[32], a16b[32], al16b[32], a16s[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float a0 = A[tid * K + 2 * c], a1 = A[tid * K + 2 * c + 1];
      uint32_t h0, h1, l0, l1;
      tc::split_tf32(a0, h0, l0);
      tc::split_tf32(a1, h1, l1);
      hi[2 * c] = h0; hi[2 * c + 1] = h1;
      lb[c] = pack_bf16(a0 - __uint_as_float(h0), a1 - __uint_as_float(h1));
      a16[c] = pack_f16(a0, a1);
      const float2 f = unpack_f16(a16[c]);
      a16b[c] = pack_bf16(f.x, f.y);
      al16b[c] = pack_bf16(a0 - f.x, a1 - f.y);
      asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(a16s[c]) : "r"(a16[c]), "r"(0x24002400u));
    }
    HODE_TMEM_ST_X32(tAhi + lane_base, hi);
    HODE_TMEM_ST_X32(tAhi + lane_base + 32, (hi + 32));
    HODE_TMEM_ST_X32(tAlb + lane_base, lb);
    HODE_TMEM_ST_X32(tA16 + lane_base, a16);
    HODE_TMEM_ST_X32(tA16b + lane_base, a16b);
    HODE_TMEM_ST_X32(tAl16b + lane_base, al16b);
    HODE_TMEM_ST_X32(tA16s + lane_base, a16s);
  }
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t LBO = N * 16, SBO = 128;
    uint32_t acc = 0;
    auto tf = [&](uint32_t a, uint32_t b) {
      const uint32_t it = idesc_ab(2, 2, M, N);
      for (int ks = 0; ks < K / 8; ++ks) { tc::mma_tf32_ts(tD, a + ks * 8, tc::make_desc(b + ks * 2 * LBO, LBO, SBO), it, acc); acc = 1; }
    };
    auto hf = [&](uint32_t a, uint32_t b, int fa, int fb) {   // one MMA = K 16 = 8 TMEM columns of A, 2 K-chunks of B
      const uint32_t id = idesc_ab(fa, fb, M, N);
      for (int ks = 0; ks < K / 16; ++ks) { mma_f16_ts(tD, a + ks * 8, tc::make_desc(b + ks * 2 * LBO, LBO, SBO), id, acc); acc = 1; }
    };
    const uint32_t bhi = tc::smem_u32(Bhi), blo = tc::smem_u32(Blo), bhib = tc::smem_u32(Bhib), bh16 = tc::smem_u32(Bh16),
                   bh16b = tc::smem_u32(Bh16b), bl16b = tc::smem_u32(Bl16b), bl16s = tc::smem_u32(Bl16s);
    if (mode == 0) { hf(tAlb, bhib, 1, 1); tf(tAhi, blo); tf(tAhi, bhi); }
    else if (mode == 1) { hf(tAl16b, bh16b, 1, 1); hf(tA16s, bl16s, 0, 0); hf(tA16, bh16, 0, 0); }
    else if (mode == 2) { hf(tAl16b, bh16b, 1, 1); hf(tA16b, bl16b, 1, 1); hf(tA16, bh16, 0, 0); }
    else { hf(tAl16b, bh16, 1, 0); hf(tA16, bl16b, 0, 1); hf(tA16, bh16, 0, 0); }
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

// cycles per MMA of a long chain into one accumulator: formats (FA, FB) alternate with (FA2, FB2)
template <int FA, int FB, int FA2, int FB2, int NN>
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
      constexpr uint32_t id = idesc_ab(FA, FB, 128, NN), id2 = idesc_ab(FA2, FB2, 128, NN);
      const uint64_t bdesc = tc::make_desc(s0 + 32768, (uint32_t)NN * 16u, 128u);
      const long long t0 = clock64();
      for (int o = 0; o < n_outer; ++o) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          mma_f16_ts(tb, tb + 256 + (i & 7) * 8, bdesc, (i & 1) ? id2 : id, (o > 0 || i > 0) ? 1u : 0u);
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

template <int FA, int FB, int FA2, int FB2, int NN>
void run_chain(const char* name, long long* dT) {
  CK(cudaFuncSetAttribute(chain_time<FA, FB, FA2, FB2, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int n_outer = 16, n = 32 * n_outer;
  for (int grid : {1, 148}) {
    chain_time<FA, FB, FA2, FB2, NN><<<grid, 128, 65536>>>(n_outer, dT);
    CK(cudaDeviceSynchronize());
    chain_time<FA, FB, FA2, FB2, NN><<<grid, 128, 65536>>>(n_outer, dT);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), dT, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    printf("cycles %-40s grid %3d: %.1f cycles/MMA (median CTA; min %.1f max %.1f)\n", name, grid,
           (double)h[grid / 2] / n, (double)h[0] / n, (double)h[grid - 1] / n);
  }
}

int main(int argc, char** argv) {
  const int n_modes = (argc > 1 && !strcmp(argv[1], "mixed")) ? 4 : 3;
  std::vector<float> hA(M * K), hB(N * K), hD(M * N);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dB, hB.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaFuncSetAttribute(f16_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728));
  const char* names[4] = {"tf32x2 + bf16 (MLP_MIX3)", "f16 main + scaled f16 + bf16 (MLP_H16)", "f16 main + 2 bf16 x bf16 cross", "f16 main + 2 mixed-format cross"};
  struct Range { const char* what; float a_scale, b_scale; };
  const Range ranges[] = {{"activations 0..1.5, weights +-0.25", 0.01f, 0.5f},
                          {"activations 0..1.5e-4 (fp16 subnormal), weights +-0.25", 1e-6f, 0.5f},
                          {"activations 0..4500, weights +-0.25", 30.f, 0.5f},
                          {"activations 0..150000 (past 65504), weights +-0.25", 1000.f, 0.5f},
                          {"activations 0..1.5, weights +-2.5e-4 (lo parts fp16-subnormal)", 0.01f, 5e-4f},
                          {"activations 0..300 (raw glucose-sized inputs), weights +-25", 2.f, 50.f}};
  for (const Range& rg : ranges) {
    srand(1);
    for (auto& x : hA) x = fabsf((float)rand() / RAND_MAX * 200.f - 50.f) * rg.a_scale;
    for (auto& x : hB) x = ((float)rand() / RAND_MAX - 0.5f) * rg.b_scale;
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    printf("range: %s\n", rg.what);
    for (int mode = 0; mode < n_modes; ++mode) {
      CK(cudaMemset(dD, 0, hD.size() * 4));
      f16_gemm<<<1, 128, 73728>>>(dA, dB, dD, mode);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
      double max_rel = 0, sum_rel = 0, max_f32 = 0;
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
          sum_rel += err / mag;
          max_f32 = fmax(max_f32, fabs(r32 - ref) / mag);
        }
      printf("  mode %d (%-38s): max err/sum|ab| %.3e  mean %.3e  (fp32 fma chain max %.3e)\n", mode, names[mode], max_rel,
             sum_rel / (M * N), max_f32);
    }
  }
  long long* dT;
  CK(cudaMalloc(&dT, 148 * sizeof(long long)));
  run_chain<0, 0, 0, 0, 64>("f16 x f16 TS M128 N64 K16", dT);
  run_chain<1, 1, 1, 1, 64>("bf16 x bf16 TS M128 N64 K16", dT);
  run_chain<0, 0, 1, 1, 64>("f16xf16 / bf16xbf16 alternating N64", dT);
  if (n_modes == 4) run_chain<0, 1, 0, 1, 64>("A f16 x B bf16 TS M128 N64 K16", dT);
  run_chain<0, 0, 0, 0, 16>("f16 x f16 TS M128 N16 K16", dT);
  printf("probe done\n");
  return 0;
}
