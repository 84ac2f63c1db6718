// TMEM read throughput of tcgen05.ld on B200: CTAs of 4 warps (each warp reads its own 32-lane quarter of a 128-column
// allocation) loop over 32x32b loads of shape x4 / x16 / x32 / x64; reports bytes per clock per SM with 1..4 CTAs per SM.
// The tcgen05 backward blend reads 128 fp32 accumulators per (Gaussian, tile): this is the ceiling of that read-back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_ld_rate tmem_ld_rate.cu && ./tmem_ld_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
  if constexpr (X == 4) {
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
    sink ^= a ^ b ^ c ^ d;
  } else if constexpr (X == 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= r[i];
  } else {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
  }
}

// every iteration reads all 128 columns of the warp's lane quarter: 128 / X loads, then ONE wait (WAIT_EACH = false)
// or a wait after every load (WAIT_EACH = true, the dependent pattern)
template <int X, bool WAIT_EACH>
__global__ void __launch_bounds__(128) rate(int iters, uint32_t* out, long long* cyc) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t taddr = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t sink = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += X) {
      ld<X>(taddr + c, sink);
      if (WAIT_EACH) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    }
    if (!WAIT_EACH) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  }
  const long long t1 = clock64();
  if (sink == 0x12345678u) out[threadIdx.x] = sink;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base_s), "r"(128) : "memory");
}

template <int X, bool W>
void run(int ctas_per_sm, uint32_t* out, long long* cyc) {
  const int iters = 2000, grid = 148 * ctas_per_sm;
  rate<X, W><<<grid, 128>>>(10, out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate<X, W><<<grid, 128>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 4]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
  const double bytes_per_cta = (double)iters * 128 /*lanes*/ * 128 /*cols*/ * 4;
  printf("x%-2d %s  %d CTA/SM: %.1f B/clk/SM (CTA-cycles %.0f, kernel %.3f ms)\n", X, W ? "wait-each" : "wait-once", ctas_per_sm,
         bytes_per_cta * ctas_per_sm / mean, mean, ms);
}

int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, sizeof(long long) * 148 * 4);
  for (int c = 1; c <= 4; c *= 2) {
    run<4, false>(c, out, cyc); run<16, false>(c, out, cyc); run<32, false>(c, out, cyc);
    run<4, true>(c, out, cyc); run<16, true>(c, out, cyc); run<32, true>(c, out, cyc);
  }
  return 0;
}
