// MUFU throughput probe: ex2.approx.ftz.f32 vs ex2.approx.ftz.f16x2 vs ex2.approx.ftz.bf16x2 (elements per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/exp2_rate scripts/micro/exp2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int elems_per_op) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 2, 1024>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = double(148) * 2 * 1024 * 8.0 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %.3f ms  %.2f Gop/s  %.2f Gelem/s  (~%.1f elem/clk/SM at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
         ops * elems_per_op / ms / 1e6, ops * elems_per_op / ms / 1e6 / 148 / (clk / 1e6) , clk / 1000);
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ex2.approx.f16", 1);
  return 0;
}
