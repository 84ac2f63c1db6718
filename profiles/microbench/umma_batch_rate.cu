// How long does one small batch of tcgen05.mma take from issue to the mbarrier arrival of its commit, and how does that
// scale with the number of co-resident CTAs?  Shapes of the blend kernels:
//   bwd: 4 x (M=128, N=64, K=16), K-major operands      (one step of blend_wsum_bwd_umma*_kernel)
//   fwd: 8 x (M=64, N=32, K=16) + 8 x (M=64, N=16, K=16), MN-major operands  (one step of blend_wsum_fwd_umma_kernel)
// Each CTA (128 threads) loops: thread 0 issues the batch + commit, ALL threads wait on the mbarrier (as the kernels do).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_batch_rate umma_batch_rate.cu && ./umma_batch_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

template <int MODE>   // 0: bwd shape, 1: fwd shape, 2: bwd shape x 4 batches per commit
__global__ void __launch_bounds__(128) rate(int iters, long long* cyc) {
  extern __shared__ __align__(128) unsigned char smem[];   // 40 KB of operands (zeros)
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s, sb = smem_u32(smem);
  uint32_t phase = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (tid == 0) {
      if (MODE == 0 || MODE == 2) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dax = make_desc(sb, 2048, 128), day = make_desc(sb + 4096, 2048, 128);
        for (int rep = 0; rep < (MODE == 2 ? 4 : 1); ++rep) {
          umma(tmem, dax, make_desc(sb + 8192, 1024, 128), idesc, 0);
          umma(tmem, dax, make_desc(sb + 8192 + 2048, 1024, 128), idesc, 1);
          umma(tmem + 64, day, make_desc(sb + 8192 + 4096, 1024, 128), idesc, 0);
          umma(tmem + 64, day, make_desc(sb + 8192 + 6144, 1024, 128), idesc, 1);
        }
      } else {
        const uint32_t id32 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
        const uint32_t id16 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
        const uint64_t dah = make_desc(sb, 128, 2048), dal = make_desc(sb + 16384, 128, 2048), db = make_desc(sb + 32768, 128, 2048);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          umma(tmem, dah + 16 * s, db + 16 * s, id32, s > 0);
          umma(tmem, dal + 16 * s, db + 16 * s, id16, 1);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128) : "memory");
}

template <int MODE>
void run(int cps, long long* cyc) {
  const int iters = 2000, grid = 148 * cps;
  cudaFuncSetAttribute(rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  rate<MODE><<<grid, 128, 40 * 1024>>>(10, cyc);
  cudaDeviceSynchronize();
  rate<MODE><<<grid, 128, 40 * 1024>>>(iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
  long long h[148 * 4]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
  const char* names[3] = {"bwd 4x(128x64x16)", "fwd 8x(64x32x16)+8x(64x16x16)", "bwd 16x(128x64x16) per commit"};
  printf("%-32s %d CTA/SM: %.0f cycles per batch per CTA, %.0f cycles per batch per SM\n", names[MODE], cps, mean / iters, mean / iters / cps);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, sizeof(long long) * 148 * 4);
  for (int c = 1; c <= 4; ++c) { run<0>(c, cyc); run<1>(c, cyc); run<2>(c, cyc); }
  return 0;
}
