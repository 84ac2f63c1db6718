// Micro-benchmark: warp-level tensor-core (mma.sync -> HMMA) issue rates on B200 next to the
// FP32 pipe, to decide whether the separable blend (rank-1 updates = small GEMMs) should run
// its inner products on the tensor cores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 2048;
constexpr int NACC = 8;   // independent accumulator tiles per warp

__device__ __forceinline__ void mma_tf32_k8(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32_k4(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}
__device__ __forceinline__ void mma_bf16_k16(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float x) {
  float d[NACC][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(x + threadIdx.x * 1e-3f + i);
  b[0] = __float_as_uint(x * 0.5f);
  b[1] = __float_as_uint(x * 0.25f);
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) d[i][q] = i + q;
  float f[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) f[i] = x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) mma_tf32_k8(d[i], a, b);
      if (MODE == 1) mma_tf32_k4(d[i], a, b);
      if (MODE == 2) mma_bf16_k16(d[i], a, b);
      if (MODE == 3) {   // 1 MMA : 4 FFMA  (tensor + FP32 pipes together)
        mma_tf32_k8(d[i], a, b);
#pragma unroll
        for (int q = 0; q < 4; ++q) f[(i + q) % NACC] = fmaf(f[(i + q) % NACC], x, 0.5f);
      }
      if (MODE == 4) {   // fresh accumulator every MMA (C = 0), result folded with FADDs -- the blend-backward shape
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        mma_tf32_k8(t, a, b);
        d[i][0] += t[0]; d[i][1] += t[1]; d[i][2] += t[2]; d[i][3] += t[3];
      }
      if (MODE == 5) {   // SHFL.BFLY
        f[i] += __shfl_xor_sync(0xffffffffu, f[i], 1);
      }
      if (MODE == 6) {   // cvt.rna.tf32 + FADD (the hi/lo split of 3xTF32)
        uint32_t h;
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(f[i]));
        f[i] = f[i] - __uint_as_float(h) + 1.0f;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + f[i];
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, double macs_per_instr, double instr_per_iter) {
  float* d; cudaMalloc(&d, 4);
  int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(d, 1.0001f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(d, 1.0001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double warps = (double)blocks * 8;
  const double instr = warps * ITERS * NACC * instr_per_iter;
  const double per_clk_sm = instr / (ms * 1e-3) / ((double)khz * 1e3) / sms;
  printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM  %8.1f MAC/clk/SM  %7.1f TFLOP/s (nominal %d MHz)\n", name, ms,
         per_clk_sm, per_clk_sm * macs_per_instr, 2.0 * instr * macs_per_instr / (ms * 1e-3) / 1e12, khz / 1000);
  cudaFree(d);
}

int main() {
  run<0>("HMMA m16n8k8 tf32", 16 * 8 * 8, 1);
  run<1>("HMMA m16n8k4 tf32", 16 * 8 * 4, 1);
  run<2>("HMMA m16n8k16 bf16", 16 * 8 * 16, 1);
  run<3>("1 HMMA tf32 k8 + 4 FFMA", 16 * 8 * 8, 1);
  run<4>("HMMA tf32 k8 (C=0) + 4 FADD", 16 * 8 * 8, 1);
  run<5>("SHFL.BFLY", 0, 1);
  run<6>("cvt.rna.tf32 + 2 FADD", 0, 1);
  return 0;
}
