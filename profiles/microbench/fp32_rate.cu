// Micro-benchmark: issue rates of the pipes the blend kernels live on (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu && ./fp32_rate
// Prints lane-ops per clock per SM for FFMA, FFMA2 (fma.rn.f32x2), FADD, MUFU.EX2, and mixes.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int NACC = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float acc[NACC];
  float2 acc2[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i] = threadIdx.x * 1e-3f + i; acc2[i] = make_float2(acc[i], acc[i] + 1.f); }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) acc[i] = fmaf(acc[i], a, b);                       // FFMA
      if (MODE == 1) acc2[i] = __ffma2_rn(acc2[i], a2, b2);             // FFMA2
      if (MODE == 2) acc[i] = acc[i] + a;                               // FADD
      if (MODE == 3) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i])); acc[i] = y; }  // MUFU
      if (MODE == 4) {                                                   // 8 FFMA : 1 MUFU (blend fwd mix)
        float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc[i]));
        acc[i] = fmaf(y, a, b);
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[(i + q + 1) % NACC] = fmaf(acc[(i + q + 1) % NACC], a, y);
      }
      if (MODE == 5) {                                                   // 4 FFMA2 : 1 MUFU
        float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(acc2[i].x));
        const float2 y2 = make_float2(y, y);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc2[(i + q) % NACC] = __ffma2_rn(acc2[(i + q) % NACC], a2, y2);
      }
      if (MODE == 6) acc2[i] = __fadd2_rn(acc2[i], a2);                 // FADD2
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i] + acc2[i].x + acc2[i].y;
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, double lane_ops_per_thread_iter) {
  float* d; cudaMalloc(&d, 4);
  int dev = 0, sms = 0, khz = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double ops = (double)blocks * 256 * ITERS * NACC * lane_ops_per_thread_iter;
  const double per_clk_sm = ops / (ms * 1e-3) / ((double)khz * 1e3) / sms;
  printf("%-28s %8.3f ms  %7.1f lane-instr/clk/SM (at nominal %d MHz)  %.2f T lane-instr/s\n", name, ms, per_clk_sm, khz / 1000,
         ops / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  run<0>("FFMA (3-reg)", 1);
  run<1>("FFMA2 (instr)", 1);
  run<2>("FADD", 1);
  run<6>("FADD2 (instr)", 1);
  run<3>("MUFU.EX2", 1);
  run<4>("8 FFMA + 1 MUFU (instr)", 9);
  run<5>("4 FFMA2 + 1 MUFU (instr)", 5);
  return 0;
}
