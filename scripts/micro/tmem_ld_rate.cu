// TMEM read bandwidth probe: bytes per clock per SM for tcgen05.ld shapes and warp counts (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_ld_rate scripts/micro/tmem_ld_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define R32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"
#define O32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
template <int SHAPE>
__global__ void __launch_bounds__(512, 1) k(long long* cycles, unsigned* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t r[32], acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + uint32_t((it & 7) * 32 + (warp >> 2) * 32) % 480;
    if (SHAPE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " R32 ", [%32];" : O32(r) : "r"(a));
    if (SHAPE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " R32 ", [%32];" : O32(r) : "r"(a));
    if (SHAPE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " R32 ", [%32];" : O32(r) : "r"(a));
    if (SHAPE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " R32 ", [%32];" : O32(r) : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= r[i];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}
template <int SHAPE>
void run(const char* name) {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  for (int warps = 4; warps <= 16; warps += 4) {
    const int iters = 4000;
    k<SHAPE><<<148, warps * 32, 0>>>(cyc, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-22s warps %2d: %s\n", name, warps, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < 148; ++i) mean += h[i]; mean /= 148;
    const double bytes = double(warps) * 32 * 128 * iters;      // 32 registers = 128 B per thread per load
    printf("%-22s warps %2d: %8.0f cycles  %6.1f B/clk/SM\n", name, warps, mean, bytes / mean);
  }
  cudaFree(cyc); cudaFree(sink);
}
int main() {
  run<0>("32x32b.x32");
  run<1>("16x256b.x8");
  run<2>("16x128b.x16");
  run<3>("16x64b.x32");
  return 0;
}
