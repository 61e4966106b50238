// Developer micro-benchmarks for the K4 design (not product code): measures, on the 16-site
// half-filling sector (12870 x 12870 fp64),
//   copy        y = x (16-byte lanes)                          -> HBM ceiling of this access shape
//   gather      y[u,:] = sum_k +-x[u'_k,:] with the real 4x4 up-hop table, rows in natural order
//   gather-chunk  the same, traversed column-chunk-major so that every gather is an L2 hit
//   smem        LDS.64 throughput: contiguous / odd-stride / random
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef long long i64;
#define MAXK 24

struct Tab { int cnt; int tgt[MAXK]; float sgn[MAXK]; };

__global__ void __launch_bounds__(1024, 1) copy_kernel(const double2* __restrict__ x, double2* __restrict__ y, i64 n2) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n2; i += (i64)gridDim.x * blockDim.x) y[i] = x[i];
}

// task = (row, chunk); chunk-major when chunk_major != 0
template <int NT, int UPG>
__global__ void __launch_bounds__(NT, 1) gather_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                       const Tab* __restrict__ tab, int nu, int nd, int wc,
                                                       int chunk_major) {
  __shared__ Tab st;
  const int nchunk = (nd + wc - 1) / wc;
  const i64 ntask = (i64)nu * nchunk;
  for (i64 t = blockIdx.x; t < ntask; t += gridDim.x) {
    int u, c;
    if (chunk_major) { c = (int)(t / nu); u = (int)(t - (i64)c * nu); }
    else { u = (int)(t / nchunk); c = (int)(t - (i64)u * nchunk); }
    __syncthreads();
    if (threadIdx.x < sizeof(Tab) / 4) ((int*)&st)[threadIdx.x] = ((const int*)&tab[u])[threadIdx.x];
    __syncthreads();
    const int c0 = c * wc, c1 = min(nd, c0 + wc);
    const int cu = st.cnt;
    for (int d = c0 + 2 * threadIdx.x; d < c1; d += 2 * NT) {
      double a0 = 0, a1 = 0;
      for (int k0 = 0; k0 < cu; k0 += UPG) {
        double2 g[UPG];
#pragma unroll
        for (int k = 0; k < UPG; ++k)
          if (k0 + k < cu) g[k] = __ldg(reinterpret_cast<const double2*>(x + (i64)st.tgt[k0 + k] * nd + d));
#pragma unroll
        for (int k = 0; k < UPG; ++k)
          if (k0 + k < cu) { a0 += st.sgn[k0 + k] * g[k].x; a1 += st.sgn[k0 + k] * g[k].y; }
      }
      *reinterpret_cast<double2*>(y + (i64)u * nd + d) = make_double2(a0, a1);
    }
  }
}

// gather with cache hints: MODE 1: y stores streaming (.cs); MODE 2: + x loads with L2 evict_last policy
template <int NT, int UPG, int MODE>
__global__ void __launch_bounds__(NT, 1) gather_hint_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                            const Tab* __restrict__ tab, int nu, int nd) {
  __shared__ Tab st;
  unsigned long long pol = 0;
  if (MODE >= 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  for (int u = blockIdx.x; u < nu; u += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < sizeof(Tab) / 4) ((int*)&st)[threadIdx.x] = ((const int*)&tab[u])[threadIdx.x];
    __syncthreads();
    const int cu = st.cnt;
    for (int d = 2 * threadIdx.x; d < nd; d += 2 * NT) {
      double a0 = 0, a1 = 0;
      for (int k0 = 0; k0 < cu; k0 += UPG) {
        double2 g[UPG];
#pragma unroll
        for (int k = 0; k < UPG; ++k)
          if (k0 + k < cu) {
            const double* ptr = x + (i64)st.tgt[k0 + k] * nd + d;
            if (MODE >= 2) asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(g[k].x), "=d"(g[k].y) : "l"(ptr), "l"(pol));
            else g[k] = __ldg(reinterpret_cast<const double2*>(ptr));
          }
#pragma unroll
        for (int k = 0; k < UPG; ++k)
          if (k0 + k < cu) { a0 += st.sgn[k0 + k] * g[k].x; a1 += st.sgn[k0 + k] * g[k].y; }
      }
      __stcs(reinterpret_cast<double2*>(y + (i64)u * nd + d), make_double2(a0, a1));
    }
  }
}

// TMA-fed gather: the neighbour rows are pulled into a shared-memory ring by 1-D bulk copies
// (cp.async.bulk, mbarrier complete_tx), NS stages of CH doubles; consumers add them from smem.
// Question: does a deep, register-free pipeline beat the 8 x 16-byte LDG.128 in flight per thread?
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
               :: "r"(s_addr(bar)), "r"(parity) : "memory");
}
template <int NS, int CH>
__global__ void __launch_bounds__(1024, 1) tma_gather_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                            const Tab* __restrict__ tab, int nu, int nd) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double* ring = reinterpret_cast<double*>(smraw);
  __shared__ uint64_t full[NS], empty[NS];
  __shared__ Tab st;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s_addr(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s_addr(&empty[s])), "r"(32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunk = (nd + CH - 1) / CH;
  long long issued = 0, consumed = 0;   // item counters (identical sequence in every thread)
  for (int u = blockIdx.x; u < nu; u += gridDim.x) {
    __syncthreads();
    if (tid < sizeof(Tab) / 4) ((int*)&st)[tid] = ((const int*)&tab[u])[tid];
    __syncthreads();
    const int cu = st.cnt;
    const long long row_items = (long long)nchunk * cu;
    long long prod = 0;   // items of this row already issued (thread 0)
    for (int c = 0; c < nchunk; ++c) {
      const int c0 = c * CH, len = min(CH, nd - c0);
      double a0 = 0, a1 = 0;
      for (int k = 0; k < cu; ++k) {
        if (tid == 0) {  // keep the ring full: issue items of this row up to NS ahead of consumption
          while (prod < row_items && issued - consumed < NS) {
            const int pc = (int)(prod / cu), pk = (int)(prod % cu);
            const int s = (int)(issued % NS);
            const uint32_t ph = (uint32_t)((issued / NS) & 1);
            if (issued >= NS) mbar_wait(&empty[s], ph ^ 1u);
            const int pc0 = pc * CH, plen = min(CH, nd - pc0);
            const uint32_t bytes = (uint32_t)plen * 8u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s_addr(&full[s])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(s_addr(ring + (size_t)s * CH)), "l"(x + (i64)st.tgt[pk] * nd + pc0), "r"(bytes), "r"(s_addr(&full[s])) : "memory");
            ++issued; ++prod;
          }
        }
        const int s = (int)(consumed % NS);
        const uint32_t ph = (uint32_t)((consumed / NS) & 1);
        mbar_wait(&full[s], ph);
        const int d = 2 * tid;
        if (d < len) {
          const double2 v = *reinterpret_cast<const double2*>(ring + (size_t)s * CH + d);
          a0 += st.sgn[k] * v.x; a1 += st.sgn[k] * v.y;
        }
        __syncwarp();
        if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s_addr(&empty[s])) : "memory");
        ++consumed;
        if (tid != 0) { if (issued < consumed) issued = consumed; }  // only thread 0 tracks `issued` exactly
      }
      const int d = 2 * tid;
      if (d < len) *reinterpret_cast<double2*>(y + (i64)u * nd + c0 + d) = make_double2(a0, a1);
    }
  }
}

// smem LDS.64 throughput: mode 0 contiguous, 1 odd stride (71), 2 random, 3 contiguous LDS.128
__global__ void __launch_bounds__(1024, 1) smem_kernel(double* out, int mode, int iters, const int* __restrict__ perm) {
  extern __shared__ double xs[];
  const int n = 13000;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = (double)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc = 0;
  int base;
  if (mode == 0) base = w * 32 + lane;
  else if (mode == 1) base = lane * 71 + w;
  else if (mode == 2) base = perm[threadIdx.x];
  else base = (w * 32 + lane) * 2;
  int idx[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) idx[q] = mode == 3 ? ((base + q * 2048) % 12800) : ((base + q * 1100) % 12800);
  if (mode == 3) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double2 v = *reinterpret_cast<const double2*>(&xs[idx[q] + 2 * (it & 31)]);
        acc += v.x + v.y;
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int q = 0; q < 8; ++q) acc += xs[idx[q] + (it & 63)];
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

int main(int argc, char** argv) {
  const int L = 16, n = 8;
  std::vector<int> states;
  for (int s = 0; s < (1 << L); ++s) if (__builtin_popcount(s) == n) states.push_back(s);
  const int nu = (int)states.size(), nd = nu;
  std::vector<int> rank(1 << L, -1);
  for (int i = 0; i < nu; ++i) rank[states[i]] = i;
  std::vector<std::pair<int, int>> bonds;
  const bool chain = argc > 1 && atoi(argv[1]) == 1;
  if (chain) for (int i = 0; i + 1 < L; ++i) bonds.push_back({i, i + 1});
  else for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
    int i = 4 * r + c;
    if (c + 1 < 4) bonds.push_back({i, i + 1});
    if (r + 1 < 4) bonds.push_back({i, i + 4});
  }
  std::vector<Tab> tab(nu);
  double avg = 0;
  for (int i = 0; i < nu; ++i) {
    Tab t; t.cnt = 0;
    for (auto b : bonds) {
      int s = states[i];
      int b1 = (s >> b.first) & 1, b2 = (s >> b.second) & 1;
      if (b1 == b2) continue;
      int s2 = s ^ (1 << b.first) ^ (1 << b.second);
      int mask = ((1 << b.second) - 1) & ~((1 << (b.first + 1)) - 1);
      t.tgt[t.cnt] = rank[s2];
      t.sgn[t.cnt] = (__builtin_popcount(s & mask) & 1) ? -1.f : 1.f;
      ++t.cnt;
    }
    avg += t.cnt;
    tab[i] = t;
  }
  printf("lattice %s: nu=%d avg up hops %.2f\n", chain ? "chain16" : "4x4", nu, avg / nu);
  const i64 dim = (i64)nu * nd;
  double *x, *y; Tab* dtab;
  CK(cudaMalloc(&x, dim * 8)); CK(cudaMalloc(&y, dim * 8)); CK(cudaMalloc(&dtab, sizeof(Tab) * nu));
  CK(cudaMemset(x, 0, dim * 8)); CK(cudaMemset(y, 0, dim * 8));
  CK(cudaMemcpy(dtab, tab.data(), sizeof(Tab) * nu, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](auto fn, int reps) {
    fn(); fn(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) fn();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
  };
  {
    float ms = timeit([&] { copy_kernel<<<148 * 2, 1024>>>((const double2*)x, (double2*)y, dim / 2); }, 10);
    printf("copy: %.3f ms  %.0f GB/s\n", ms, 16.0 * dim / ms / 1e6);
  }
  const double hops = avg / nu;
  auto report = [&](const char* name, float ms) {
    printf("%s: %.3f ms | algorithmic(16B/state) %.0f GB/s | gather traffic (%.1f+1)x8B/state = %.0f GB/s\n",
           name, ms, 16.0 * dim / ms / 1e6, hops, (hops + 1) * 8.0 * dim / ms / 1e6);
  };
  for (int grid : {148, 296}) {
    char nm[128];
    snprintf(nm, sizeof nm, "gather natural rows, grid %d x1024 UPG8", grid);
    report(nm, timeit([&] { gather_kernel<1024, 8><<<grid, 1024>>>(x, y, dtab, nu, nd, nd, 0); }, 5));
  }
  report("gather natural rows, grid 296 x512 UPG8", timeit([&] { gather_kernel<512, 8><<<296, 512>>>(x, y, dtab, nu, nd, nd, 0); }, 5));
  report("gather natural rows, grid 592 x512 UPG13", timeit([&] { gather_kernel<512, 13><<<592, 512>>>(x, y, dtab, nu, nd, nd, 0); }, 5));
  report("gather natural rows + st.cs, 148x1024 UPG8", timeit([&] { gather_hint_kernel<1024, 8, 1><<<148, 1024>>>(x, y, dtab, nu, nd); }, 5));
  report("gather natural rows + st.cs + ld evict_last, 148x1024 UPG8", timeit([&] { gather_hint_kernel<1024, 8, 2><<<148, 1024>>>(x, y, dtab, nu, nd); }, 5));
  report("gather natural rows + st.cs, 148x512 UPG16", timeit([&] { gather_hint_kernel<512, 16, 1><<<148, 512>>>(x, y, dtab, nu, nd); }, 5));
  report("gather natural rows + st.cs, 148x768 UPG12", timeit([&] { gather_hint_kernel<768, 12, 1><<<148, 768>>>(x, y, dtab, nu, nd); }, 5));
  report("gather chunk-major wc=2048, 148x1024 UPG8", timeit([&] { gather_kernel<1024, 8><<<148, 1024>>>(x, y, dtab, nu, nd, 2048, 1); }, 5));
  report("gather chunk-major wc=4096, 148x1024 UPG8", timeit([&] { gather_kernel<1024, 8><<<148, 1024>>>(x, y, dtab, nu, nd, 4096, 1); }, 5));
  {
    constexpr int NS = 12, CH = 2048;
    const size_t smem = (size_t)NS * CH * 8;
    CK(cudaFuncSetAttribute(tma_gather_kernel<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    report("gather via TMA bulk copies, 12 x 16 KB ring, 148x1024", timeit([&] { tma_gather_kernel<NS, CH><<<148, 1024, smem>>>(x, y, dtab, nu, nd); }, 5));
  }
  {
    constexpr int NS = 24, CH = 1024;
    const size_t smem = (size_t)NS * CH * 8;
    CK(cudaFuncSetAttribute(tma_gather_kernel<NS, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    report("gather via TMA bulk copies, 24 x 8 KB ring, 148x1024 (512 active)", timeit([&] { tma_gather_kernel<NS, CH><<<148, 1024, smem>>>(x, y, dtab, nu, nd); }, 5));
  }
  for (int wc : {512}) {
    char nm[128];
    snprintf(nm, sizeof nm, "gather chunk-major wc=%d (L2 set %.0f MB), grid 1184 x128 UPG13", wc, 8.0 * nu * wc / 1e6);
    report(nm, timeit([&] { gather_kernel<128, 13><<<1184, 128>>>(x, y, dtab, nu, nd, wc, 1); }, 5));
    snprintf(nm, sizeof nm, "gather chunk-major wc=%d, grid 592 x256 UPG13", wc);
    report(nm, timeit([&] { gather_kernel<256, 13><<<592, 256>>>(x, y, dtab, nu, nd, wc, 1); }, 5));
  }
  // smem
  {
    std::vector<int> perm(1024);
    srand(1);
    for (int i = 0; i < 1024; ++i) perm[i] = rand() % 12800;
    int* dperm; CK(cudaMalloc(&dperm, 4096)); CK(cudaMemcpy(dperm, perm.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000));
    const int iters = 4000;
    const char* names[] = {"contiguous LDS.64", "odd-stride(71) LDS.64", "random LDS.64", "contiguous LDS.128"};
    for (int mode = 0; mode < 4; ++mode) {
      float ms = timeit([&] { smem_kernel<<<148, 1024, 110000>>>(y, mode, iters, dperm); }, 3);
      double elems = 148.0 * 1024 * iters * 8 * (mode == 3 ? 2 : 1);
      printf("smem %s: %.3f ms -> %.1f doubles/clk/SM (at 1.965 GHz)\n", names[mode], ms,
             elems / 148 / (ms * 1e-3 * 1.965e9));
    }
  }
  return 0;
}
