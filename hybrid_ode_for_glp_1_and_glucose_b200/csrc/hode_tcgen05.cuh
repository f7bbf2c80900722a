// hode_tcgen05.cuh — thin inline-PTX wrappers for the Blackwell tensor-core path (sm_100a):
// TMEM allocation, tcgen05.ld/st/mma/commit, mbarrier, bulk async copy.  The recipe
// (kind::tf32, A operand in TMEM, B operand K-major/no-swizzle in shared memory) was
// validated on hardware by csrc/probe/tc_probe.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace hode {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Warp-uniform election of one lane (PTX elect.sync).  tcgen05.mma must be issued from
// warp-uniform control flow with warp-uniform operands: from a per-thread branch such as
// `if (threadIdx.x == 0)` ptxas wraps EVERY UTCHMMA in an ELECT / R2UR.BROADCAST loop, which
// was measured (ncu, profiles/) to cost ~100 cycles per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}

// ---- TMEM -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x N consecutive 32-bit columns; thread i of the warp owns TMEM lane (lane_base + i)
#define HODE_TMEM_LD_X32(taddr, r)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                            \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22," \
               "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                 "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                 "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),       \
                 "=r"(r[31])                                                                        \
               : "r"(taddr) : "memory")

#define HODE_TMEM_LD_X16(taddr, r)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                            \
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), \
                 "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                 "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                              \
               : "r"(taddr) : "memory")

#define HODE_TMEM_LD_X8(taddr, r)                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"            \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), \
                 "=r"(r[7])                                                                         \
               : "r"(taddr) : "memory")

#define HODE_TMEM_ST_X32(taddr, r)                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                       \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23," \
               "%24,%25,%26,%27,%28,%29,%30,%31,%32};"                                               \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),       \
                 "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),     \
                 "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), \
                 "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), \
                 "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

#define HODE_TMEM_ST_X8(taddr, r)                                                                    \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"              \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),       \
                 "r"(r[6]), "r"(r[7]) : "memory")

#define HODE_TMEM_ST_X16(taddr, r)                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                       \
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                           \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),       \
                 "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),     \
                 "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
#ifdef HODE_DEBUG_WAIT
// debug build: a wait that does not complete within ~0.5 s records where it is stuck in a host-mapped buffer
// (hode_debug_set_buffer; it survives the trap) and traps
__device__ int* g_dbg_buf = nullptr;   // [0] = count, then {block x, block y, thread, line, parity} per record
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int line) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 1000000000LL) {
      if (g_dbg_buf) {
        const int i = atomicAdd(g_dbg_buf, 1);
        if (i < 4000) {
          int* r = g_dbg_buf + 1 + 5 * i;
          r[0] = blockIdx.x; r[1] = blockIdx.y; r[2] = threadIdx.x; r[3] = line; r[4] = (int)parity;
        }
        __threadfence_system();
      }
      const long long t1 = clock64();
      while (clock64() - t1 < 200000000LL) { }   // let the other stuck threads report too
      __trap();
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
#elif defined(HODE_SPIN_WAIT)
// measurement build (tools/build_variants.py): poll with test_wait instead of suspending in try_wait
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "HODE_SPIN_LOOP:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra HODE_SPIN_DONE;\n\t"
      "bra HODE_SPIN_LOOP;\n\t"
      "HODE_SPIN_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "HODE_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra HODE_WAIT_DONE;\n\t"
      "bra HODE_WAIT_LOOP;\n\t"
      "HODE_WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
#endif
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all state spaces: generic-proxy GLOBAL stores that a later bulk copy (async proxy) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// per-thread 16-byte asynchronous copy global -> shared (LDGSTS) and its completion wait
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- MMA ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T   (kind::tf32; A: lane = row, column = k)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (bit layout of cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64).
// LBO = byte stride between the two 16-byte K chunks of one MMA, SBO = between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, both K-major,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest split of an fp32 value into a TF32-representable high part and the (also
// TF32-rounded) remainder: x ~= hi + lo with |x - hi - lo| <= 2^-22 |x|
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  hi = h;
  lo = (__float_as_uint(x - __uint_as_float(h)) + 0x1000u) & 0xFFFFE000u;
}

}  // namespace tc
}  // namespace hode
