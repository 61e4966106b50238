// Microbenchmark: posted remote stores vs remote red.add.f64 into a peer's memory over NVLink.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libpeer_red.so peer_red.cu
// mode 0: dst[i] = src[i];  mode 1: red.relaxed.sys.add.f64 dst[i] += src[i];  mode 2: red.gpu scope;
// mode 3: transposed-tile pattern of the push-style way back: 32 rows x 128 columns per tile, remote rows of 1 KB at
// stride ld doubles, via red.
#include <cstdint>
#include <cuda_runtime.h>
extern "C" __global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, long long n, int mode) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = src[i];
    if (mode == 0) dst[i] = v;
    else if (mode == 1) asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(dst + i), "d"(v) : "memory");
    else asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" :: "l"(dst + i), "d"(v) : "memory");
  }
}
extern "C" __global__ void k_tile(double* __restrict__ dst, const double* __restrict__ src, long long rows, long long ld, int mode) {
  // dst: rows x ld (row-major); every warp instruction touches 32 consecutive doubles of one row; consecutive
  // instructions of a warp walk 4 column segments, then the next row of the tile
  const long long tiles_c = ld / 128, ntiles = (rows / 32) * tiles_c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long tr = t / tiles_c, tc = t - tr * tiles_c;
    for (int rr = ty; rr < 32; rr += 8)
#pragma unroll
      for (int cs = 0; cs < 4; ++cs) {
        const long long i = (tr * 32 + rr) * ld + tc * 128 + cs * 32 + tx;
        const double v = src[i];
        if (mode == 0) dst[i] = v;
        else asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(dst + i), "d"(v) : "memory");
      }
  }
}
extern "C" int run_copy(double* dst, const double* src, long long n, int mode, int grid, void* stream) {
  k_copy<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, src, n, mode);
  return (int)cudaGetLastError();
}
extern "C" int run_tile(double* dst, const double* src, long long rows, long long ld, int mode, int grid, void* stream) {
  k_tile<<<grid, 256, 0, (cudaStream_t)stream>>>(dst, src, rows, ld, mode);
  return (int)cudaGetLastError();
}
