// PTX helpers shared by the tcgen05 blend kernels: mbarriers, TMA bulk copies, cp.async, UMMA shared-memory
// descriptors and the MMA itself.  sm_100a only.
#pragma once
#include <stdint.h>

namespace b2s {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarriers and TMA bulk copies (global -> shared, completion on an mbarrier) ----------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 20); ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();   // a copy or MMA that never completes is a bug: fail the launch instead of hanging the GPU
}

// ---- cp.async (per-thread staging: a thread later reads only the slots it copied itself) --------
__device__ __forceinline__ void cp_async16_b(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4_b(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_b() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_b() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- UMMA ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE.  The canonical layouts (cute/atom/mma_traits_sm100.hpp):
//   K-major : 8-row x 16-byte core matrices; the K chunks of 8 elements LBO apart, 8-row groups SBO apart
//   MN-major: core matrix = 8 K-rows x 16 bytes (8 MN elements); K groups of 8 LBO apart, MN groups of 8 SBO apart
// (validated by profiles/microbench/umma_probe.cu and umma_probe_mn.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;       // leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
  return d;                                               // base offset 0, SWIZZLE_NONE
}
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return umma_desc(saddr, lbo_bytes, sbo_bytes);
}

// instruction descriptor of kind::f16: f32 accumulate, f16 x f16 operands
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace b2s
