// Probe for the tensor-core blend FORWARD: tcgen05.mma kind::f16 with M = 64, A and B both MN-major in shared memory
// (SWIZZLE_NONE), K = 128 consumed as 8 instructions of K = 16, N = 32 then N = 16 accumulated into the first 16
// columns; accumulator read back with tcgen05.ld 32x32b.  Validates, before they are built into the kernel:
//   * the MN-major canonical layout: core matrix = 8 K-rows x 16 bytes (8 MN elements); K groups LBO apart,
//     MN groups SBO apart   (cute/atom/mma_traits_sm100.hpp: ((1,n),(8,k)):((X,SBO),(1,LBO)) in uint128 units)
//   * the operand placement "thread i owns K index i": offset(mn, i) = ((mn/8)*128 + i)*16 B + (mn%8)*2 B, i.e.
//     LBO = 128 B, SBO = 2048 B, K slice s starts at +256*s B
//   * the M = 64 accumulator placement: row m -> TMEM lane (m/16)*32 + m%16
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe_mn umma_probe_mn.cu && ./umma_probe_mn
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int M = 64, N = 32, N2 = 16, K = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int mn_off(int mn, int i) { return ((mn / 8) * K + i) * 8 + (mn % 8); }   // in halves

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

// A [M][K], A2 [M][K], B [N][K] row-major halves in global memory; D out: [128 lanes][N]; cycles out
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* A2, const __half* B, float* D, long long* cyc, int reps) {
  __shared__ __align__(128) __half sA[M * K];
  __shared__ __align__(128) __half sA2[M * K];
  __shared__ __align__(128) __half sB[N * K];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * K; i += 128) { sA[mn_off(i / K, i % K)] = A[i]; sA2[mn_off(i / K, i % K)] = A2[i]; }
  for (int i = tid; i < N * K; i += 128) sB[mn_off(i / K, i % K)] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  uint32_t phase = 0;
  long long t0 = 0, t1 = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      const uint32_t id32 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint32_t id16 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N2 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      if (rep == 1) t0 = clock64();
      for (int s = 0; s < K / 16; ++s) {
        const uint64_t da = make_desc(smem_u32(sA) + 256 * s, 128, 2048), da2 = make_desc(smem_u32(sA2) + 256 * s, 128, 2048);
        const uint64_t db = make_desc(smem_u32(sB) + 256 * s, 128, 2048);
        umma(tmem, da, db, id32, s > 0);          // D[:, 0:32]  (+)= A . B^T
        umma(tmem, da2, db, id16, 1);             // D[:, 0:16]  += A2 . B[0:16]^T
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
      uint32_t done = 0;
      for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      if (!done) __trap();
      phase ^= 1u;
    }
    if (tid == 0) t1 = clock64();
  }
  if (tid == 0 && reps > 1) *cyc = (t1 - t0) / (reps - 1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(32) : "memory");
}

int main() {
  static __half hA[M * K], hA2[M * K], hB[N * K];
  static float fA[M * K], fA2[M * K], fB[N * K];
  srand(1);
  for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < M * K; ++i) { hA2[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fA2[i] = __half2float(hA2[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fB[i] = __half2float(hB[i]); }
  __half *dA, *dA2, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dA2, sizeof(hA2)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dA2, hA2, sizeof(hA2), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int reps = 1; reps <= 65; reps += 64) {
    cudaMemset(dD, 0, 128 * N * 4); cudaMemset(dC, 0, 8);
    probe<<<1, 128>>>(dA, dA2, dB, dD, dC, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float hD[128 * N]; long long cyc = 0;
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double worst = 0, mag = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)fA[m * K + k] * fB[n * K + k] + (n < N2 ? (double)fA2[m * K + k] * fB[n * K + k] : 0.0);
        const int lane = (m / 16) * 32 + m % 16;
        worst = fmax(worst, fabs(ref - hD[lane * N + n]));
        mag = fmax(mag, fabs(ref));
      }
    printf("reps %d: max |D - ref| = %.3e (max |ref| %.1f)  D[0][0]=%f  cycles per 16-MMA batch (issue+commit+wait) = %lld\n", reps, worst, mag, hD[0], cyc);
  }
  return 0;
}
