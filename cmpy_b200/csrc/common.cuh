// common.cuh -- shared host/device helpers of libcmpy_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <atomic>
#include "../../include/cmpy_b200.h"

typedef long long i64;
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------
// (single translation unit: cmpy_b200.cu includes every .cuh exactly once)
static thread_local std::string g_cmpy_err;
static std::atomic<long long> g_cmpy_launches{0};

static inline int cmpy_fail(int code, const std::string& msg) {
  g_cmpy_err = msg;
  return code;
}

#define CU_CHECK(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return cmpy_fail(CMPY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                          " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                    \
  } while (0)

// NVTX range around an ABI call / a phase (visible to ncu --nvtx and Nsight Systems; a function-pointer test
// when no tool is attached).  The reference has no tracing at all.
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

#define KERNEL_CHECK()                                   \
  do {                                                   \
    g_cmpy_launches.fetch_add(1);                        \
    CU_CHECK(cudaGetLastError());                        \
  } while (0)

#define ARG_CHECK(cond, msg)                             \
  do {                                                   \
    if (!(cond)) return cmpy_fail(CMPY_ERR_ARG, msg);    \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------
// binomial table (host + __constant__ copy)
// ---------------------------------------------------------------------------------
#define BINOM_N 65
__constant__ u64 c_binom[BINOM_N][BINOM_N];

static const u64* host_binom() {  // [BINOM_N*BINOM_N], saturating at 2^63-1
  static u64 tab[BINOM_N * BINOM_N];
  static bool init = false;
  if (!init) {
    const u64 SAT = 0x7fffffffffffffffull;
    for (int n = 0; n < BINOM_N; ++n)
      for (int k = 0; k < BINOM_N; ++k) {
        u64 v;
        if (k == 0) v = 1;
        else if (n == 0) v = 0;
        else {
          u64 a = tab[(n - 1) * BINOM_N + k - 1], b = tab[(n - 1) * BINOM_N + k];
          v = (a >= SAT - b) ? SAT : a + b;
        }
        tab[n * BINOM_N + k] = v;
      }
    init = true;
  }
  return tab;
}

static int ensure_binom_uploaded() {  // once per device
  static bool done[64] = {false};
  int dev = 0;
  CU_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return CMPY_OK;
  CU_CHECK(cudaMemcpyToSymbol(c_binom, host_binom(), sizeof(u64) * BINOM_N * BINOM_N));
  if (dev < 64) done[dev] = true;
  return CMPY_OK;
}

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
// (__host__ __device__: tests/emu runs the same functions on the CPU; the device code reads the
// __constant__ table, the host code the table it was copied from.)
__host__ __device__ __forceinline__ u64 binom_at(int n, int k) {
#ifdef __CUDA_ARCH__
  return c_binom[n][k];
#else
  return host_binom()[n * BINOM_N + k];
#endif
}
__host__ __device__ __forceinline__ int popc64(u64 v) {
#ifdef __CUDA_ARCH__
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}
__host__ __device__ __forceinline__ int ffs64(u64 v) {
#ifdef __CUDA_ARCH__
  return __ffsll((long long)v);
#else
  return __builtin_ffsll((long long)v);
#endif
}

// colex (combinadic) rank of s among integers of the same popcount, ascending order.
__host__ __device__ __forceinline__ i64 colex_rank(u64 s) {
  i64 r = 0;
  int k = 0;
  while (s) {
    int p = ffs64(s) - 1;
    s &= s - 1;
    ++k;
    r += (i64)binom_at(p, k);
  }
  return r;
}

// inverse: the idx-th (0-based) integer with popcount n (any width up to 64 bits)
__host__ __device__ __forceinline__ u64 colex_unrank(i64 idx, int n, int num_sites) {
  u64 s = 0;
  u64 r = (u64)idx;
  int p = num_sites;
  for (int k = n; k >= 1; --k) {
    do { --p; } while (binom_at(p, k) > r);
    s |= (1ull << p);
    r -= binom_at(p, k);
  }
  return s;
}

__host__ __device__ __forceinline__ i64 bsearch_left(const i64* __restrict__ a, i64 n, i64 x) {
  i64 lo = 0, hi = n;
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum, fixed order (deterministic). `red` = shared double[32]. Result valid
// in every thread of warp 0 (thread 0 is what callers use).
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` reuse
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0) t = warp_sum(t);
  return t;
}

// Two-stage deterministic grid reduction: every block deposits its partial; the block
// that draws the last ticket sums all partials in a fixed order and returns true (in
// thread 0, with *total valid).  `partials` has >= gridDim.x entries; *ticket must be 0
// on entry and is reset to 0 by the last block.
__device__ __forceinline__ bool grid_sum_last(double block_partial, double* partials,
                                              unsigned* ticket, double* red, double* total) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_partial;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x)
    acc += ((volatile double*)partials)[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    *total = acc;
    *ticket = 0u;
    __threadfence();
  }
  return threadIdx.x == 0;
}

// Raise a kernel's dynamic shared memory limit to the device maximum (idempotent; never
// lowers it, so operators of different sizes can coexist).
template <typename K>
static int raise_smem_limit(K kernel, i64 optin) {
  cudaFuncAttributes a;
  CU_CHECK(cudaFuncGetAttributes(&a, kernel));
  const int maxdyn = (int)(optin - (i64)a.sharedSizeBytes);
  CU_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxdyn));
  return CMPY_OK;
}

// Lanczos fusion context (device scalars live in one small buffer owned by the handle)
struct LzCtx {
  int enabled;        // 0: plain y = Hx
  const int* iter;    // current iteration index j (device)
  const double* beta; // beta[0..]  (beta[j] = |r_j|)
  double* alpha;      // alpha[j] written by the last block
  double* partials;   // >= gridDim.x
  unsigned* ticket;
};

// ---------------------------------------------------------------------------------
// operator base class
// ---------------------------------------------------------------------------------
struct cmpy_op_s {
  i64 size = 0;
  int device = 0;
  int sm_count = 148;
  i64 smem_optin = 0;
  // reduction workspace
  double* d_partials = nullptr;  // [max_blocks]
  unsigned* d_ticket = nullptr;
  int max_blocks = 0;
  // lanczos device scalars
  double* d_alpha = nullptr;
  double* d_beta = nullptr;
  int* d_iter = nullptr;
  double* d_ritz = nullptr;
  int lz_cap = 0;
  int variant = 0;
  virtual ~cmpy_op_s() {
    cudaFree(d_partials); cudaFree(d_ticket); cudaFree(d_alpha); cudaFree(d_beta);
    cudaFree(d_iter); cudaFree(d_ritz);
  }
  // y = Hx (lz.enabled==0) or fused Lanczos step on (x, y=w) (lz.enabled==1)
  virtual int apply(const double* x, double* y, const LzCtx& lz, cudaStream_t st) = 0;
  virtual int diagonal(double* d_diag, cudaStream_t st) = 0;
  int init_workspace() {
    CU_CHECK(cudaGetDevice(&device));
    cudaDeviceProp prop;
    CU_CHECK(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    smem_optin = (i64)prop.sharedMemPerBlockOptin;
    max_blocks = sm_count * 32;
    CU_CHECK(cudaMalloc(&d_partials, sizeof(double) * max_blocks));
    CU_CHECK(cudaMalloc(&d_ticket, sizeof(unsigned) * 4));
    CU_CHECK(cudaMemset(d_ticket, 0, sizeof(unsigned) * 4));
    return CMPY_OK;
  }
  int ensure_lz_capacity(int maxit) {
    if (maxit + 2 <= lz_cap) return CMPY_OK;
    cudaFree(d_alpha); cudaFree(d_beta); cudaFree(d_iter); cudaFree(d_ritz);
    d_alpha = d_beta = d_ritz = nullptr; d_iter = nullptr;
    lz_cap = maxit + 2;
    CU_CHECK(cudaMalloc(&d_alpha, sizeof(double) * lz_cap));
    CU_CHECK(cudaMalloc(&d_beta, sizeof(double) * lz_cap));
    CU_CHECK(cudaMalloc(&d_ritz, sizeof(double) * lz_cap));
    CU_CHECK(cudaMalloc(&d_iter, sizeof(int) * 4));
    return CMPY_OK;
  }
};
