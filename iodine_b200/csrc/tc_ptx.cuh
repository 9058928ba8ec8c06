// tc_ptx.cuh -- PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, refine_tc.cu): mbarriers, TMA / bulk
// copies, tcgen05.mma / commit / ld, shared-memory matrix descriptors.  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace iod {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// suspend-time hint of try_wait (ns): the thread sleeps in hardware until the phase completes or the hint expires, instead
// of coming back every few hundred cycles to spin -- waiting warps then stop taking issue slots from the working ones
#ifndef IOD_TRYWAIT_HINT_NS
#define IOD_TRYWAIT_HINT_NS 20000u
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(IOD_TRYWAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
// A deadlock must become a launch failure, never a hung GPU: bounded wait, then trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_tc: mbarrier wait timed out (tag %d, block %d, thread %d)\n", tag, (int)blockIdx.x,
             (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
      : "memory");
}
// pull `bytes` (multiple of 16) of global memory into L2 ahead of a later ordinary load; no destination, no barrier
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// kind::tf32: fp32 containers, 10-bit mantissa used (the low 13 bits are ignored), K = 8 per instruction
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
template <bool TF>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF) tc_mma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
  else tc_mma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
}
// round an fp32 value to tf32 (round to nearest, ties away): what the tensor core would otherwise TRUNCATE
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, layout type 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr16, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)(addr16 & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) |
         ((uint64_t)(sbo16 & 0x3FFFu) << 32) | (1ull << 46);
}

#define IOD_TMEM_LD16(r, addr)                                                                  \
  asm volatile(                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                 \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),      \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), \
        "=r"(r[14]), "=r"(r[15])                                                                \
      : "r"(addr))

// zero 16 accumulator columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z) : "memory");
}

// ELU with the hardware exponential: |error| <= ~2e-7 absolute, far below bf16 rounding.
// Branch-free form with the flush-to-zero ex2 (the non-ftz __expf carries extra scaling instructions for
// subnormals, and the ELU-heavy kernels are instruction-issue bound): max(v,0) + (2^(min(v,0) log2 e) - 1).
__device__ __forceinline__ float elu_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 0.f) * 1.4426950408889634f));
  return fmaxf(v, 0.f) + (e - 1.f);
}


}  // namespace iod
