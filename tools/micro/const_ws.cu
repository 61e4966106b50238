// Microbenchmark: throughput of warp-uniform loads from the kernel-parameter constant bank (LDCU) as a
// function of the table working set.  Each warp walks its own pseudo-random positions inside the first
// `ws` bytes of a 28 KB parameter table; the loaded value feeds the next LDS address like the hop
// lists of hubbard_eng.cuh.  Build: nvcc -arch=sm_100a -O3 -o const_ws const_ws.cu ; run: ./const_ws
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Tab { uint16_t e[14336]; };
__global__ void __launch_bounds__(1024, 1) walk(const __grid_constant__ Tab tab, int ws_entries, int iters, double* out) {
  extern __shared__ double xs[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) xs[i] = i;
  __syncthreads();
  const uint32_t xa = (uint32_t)__cvta_generic_to_shared(xs) + lane * 8;
  double acc = 0.0;
  uint32_t pos = (warp * 977u) % ws_entries;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t e = tab.e[pos + j];
      double v;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(xa + (e & 0x1ff8u)));
      acc += v;
    }
    pos = (pos * 5u + 131u + warp) % (ws_entries - 4);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  Tab h;
  for (int i = 0; i < 14336; ++i) h.e[i] = (uint16_t)((i * 2654435761u) >> 16);
  double* out; cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int ws_kb : {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20, 24, 28}) {
    const int ws_entries = ws_kb * 512;
    walk<<<148, 1024, 65536>>>(h, ws_entries, 200, out);
    cudaEventRecord(a);
    walk<<<148, 1024, 65536>>>(h, ws_entries, iters, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double loads = 4.0 * iters * 32;   // per SM (warp-level LDCU)
    printf("ws %2d KB: %.3f ms  %.2f clk per warp-level const load per SM (at 1.965 GHz)  err=%s\n", ws_kb, ms,
           ms * 1e-3 * 1.965e9 / loads, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
