// Probe: one tcgen05.mma (kind::f16, M=128, N=64, K=16, A and B K-major in shared memory without swizzle),
// accumulator read back with tcgen05.ld 32x32b -- validates the smem-descriptor / TMEM layout assumptions of the
// tensor-core blend backward before they are built into the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int M = 128, N = 64, K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE canonical layout: 8-row x 16-byte core matrices, rows of a core matrix contiguous (16 B apart),
// row groups SBO bytes apart, the two 8-element K chunks LBO bytes apart.
__host__ __device__ inline int canon_off(int row, int k, int rows) {   // in halves
  const int lbo = rows / 8 * 64;                                       // halves: one K chunk = rows * 8 halves
  return (k / 8) * lbo + (row / 8) * 64 + (row % 8) * 8 + (k % 8);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int accumulate_twice) {
  __shared__ __align__(128) __half sA[M * K];
  __shared__ __align__(128) __half sB[N * K];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) sA[canon_off(i / K, i % K, M)] = A[i];
  for (int i = tid; i < N * K; i += 128) sB[canon_off(i / K, i % K, N)] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic-proxy smem writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint64_t da = make_desc(smem_u32(sA), M * 16, 128), db = make_desc(smem_u32(sB), N * 16, 128);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // F32 accum, F16 x F16, K-major
    for (int rep = 0; rep <= accumulate_twice; ++rep) {
      const uint32_t acc = rep;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMA
  {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (!done) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // each warp reads its 32 lanes; 64 columns in 4 loads of 16
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int q = 0; q < 16; ++q) D[(size_t)tid * N + c0 + q] = __uint_as_float(r[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(64) : "memory");
  (void)lane;
}

int main() {
  __half hA[M * K], hB[N * K];
  float fA[M * K], fB[N * K];
  srand(1);
  for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int twice = 0; twice < 2; ++twice) {
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128>>>(dA, dB, dD, twice);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float hD[M * N];
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k];
        ref *= (twice + 1);
        worst = fmax(worst, fabs(ref - hD[m * N + n]));
      }
    printf("accumulate x%d: max |D - A.B^T| = %.3e  (D[0][0]=%f D[5][7]=%f D[127][63]=%f)\n", twice + 1, worst, hD[0], hD[5 * N + 7], hD[127 * N + 63]);
  }
  return 0;
}
