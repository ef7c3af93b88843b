// MUFU throughput probe (sm_100a): exp2 of fp32 (ex2.approx.ftz.f32, one value per lane-op) against the packed
// ex2.approx.ftz.bf16x2 (two values per lane-op) and ex2.approx.f16x2, full occupancy, register-resident dependent chains.
// Prints values per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(float* out, int iters) {
  unsigned a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;   // small positive halves / bf16s
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(a[i]);
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
static void run(const char* name, int vals_per_op) {
  float* d;
  cudaMalloc(&d, 4);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  probe<MODE><<<sms * 4, 512>>>(d, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE><<<sms * 4, 512>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ops = (double)sms * 4 * 512 * iters * 8;          // lane-ops
  const double clk = ms * 1e-3 * khz * 1e3;                      // at the nominal max clock (the probe runs unthrottled)
  printf("%-28s %8.3f ms  %6.2f lane-ops/clk/SM  %6.2f values/clk/SM\n", name, ms, ops / clk / sms, ops * vals_per_op / clk / sms);
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  return 0;
}
