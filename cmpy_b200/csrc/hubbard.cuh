// hubbard.cuh -- K4: matrix-free fp64 H.v for Hubbard / Anderson sectors.
//
// Amplitudes are a (num_up x num_dn) row-major matrix X (idx = up_idx*num_dn + dn_idx,
// ref: cmpy/operators.py:33-90).  H.X = D o X + X T_dn^T + T_up X:
//   D      diagonal  E_up[u] + E_dn[d] + sum_i u_i [up&dn]_i   (ref: operators.py:305-422)
//   T_dn   dn hops: gathers inside one row                     (ref: operators.py:522-527)
//   T_up   up hops: whole-row gathers from other rows          (ref: operators.py:515-520)
// Matrix element of a hop = sign*hop (ref: operators.py:425-460).
//
// Kernels:
//   hub_flat_kernel  one thread per amplitude, every source read from global (L1/L2);
//                    any sector shape (num_dn may be 1).
//   hub_row_kernel   persistent CTAs, one row at a time: the row of x is staged in shared
//                    memory (all dn-hop gathers hit smem), up-hop sources are coalesced
//                    reads of whole remote rows.
// Both optionally fuse the Lanczos step  w = Hx/beta_j - (beta_j/beta_{j-1}) w_old,
// alpha_j = <x/beta_j, w>  (ref recurrence: cmpy/exactdiag.py:324-347).
#pragma once
#include "common.cuh"
#include "sector.cuh"

struct HubParams {
  i64 num_up, num_dn;
  const uint32_t* up_states;  // 32-bit copies of the strings
  const uint32_t* dn_states;
  const uint32_t* ell_up; const uint8_t* cnt_up;
  const uint32_t* ell_dn; const uint8_t* cnt_dn;
  const double* e_up; const double* e_dn;
  const double* hop;  // [nbonds]
  const double* u;    // [num_sites]
  int num_sites;
  double u0, hop0;
  i64 row0, nrows;    // slab of rows handled by this launch (x, y point at the slab)
  int with_up;        // include up hops (needs the whole vector: row0 == 0, nrows == num_up)
  int accumulate;     // y += instead of y =
  const double* acc_scale;  // device {c1, c2} (or null): y = c1 * (H x) + c2 * y  (row engine only; used by the
                            // sharded Lanczos recurrence, cmpy/exactdiag.py:324-347 in the unnormalised basis)
  const double* x;
  double* y;
  LzCtx lz;
};

template <bool UNI>
__device__ __forceinline__ double hub_diag(const HubParams& p, uint32_t ups, uint32_t dns,
                                           double eu, double ed) {
  if (UNI) return eu + ed + p.u0 * (double)__popc(ups & dns);
  uint32_t both = ups & dns;
  double w = 0.0;
  while (both) {
    int i = __ffs(both) - 1;
    both &= both - 1;
    w += p.u[i];
  }
  return eu + ed + w;
}

template <bool LZ>
__device__ __forceinline__ void hub_store(const HubParams& p, i64 i, double acc, double xi,
                                          double s1, double s2, bool has_prev, double& dot) {
  if (LZ) {
    double w = s1 * acc;
    if (has_prev) w -= s2 * p.y[i];
    p.y[i] = w;
    dot += (s1 * xi) * w;
  } else {
    p.y[i] = p.accumulate ? p.y[i] + acc : acc;
  }
}

template <bool LZ>
__device__ __forceinline__ void lz_scalars(const LzCtx& lz, int& j, double& s1, double& s2,
                                           bool& has_prev) {
  j = 0; s1 = 1.0; s2 = 0.0; has_prev = false;
  if (LZ) {
    j = *lz.iter;
    double bj = lz.beta[j];
    s1 = 1.0 / bj;
    has_prev = j > 0;
    if (has_prev) s2 = bj / lz.beta[j - 1];
  }
}

template <bool LZ>
__device__ __forceinline__ void lz_finish(const LzCtx& lz, int j, double dot, double* red) {
  if (LZ) {
    double bsum = block_sum(dot, red);
    double total;
    // enabled == 2: a later launch of the same H.v (the operator is split over several kernels)
    if (grid_sum_last(bsum, lz.partials, lz.ticket, red, &total)) lz.alpha[j] = (lz.enabled == 2 ? lz.alpha[j] : 0.0) + total;
  }
}

template <bool UNI, bool LZ>
__global__ void __launch_bounds__(256) hub_flat_kernel(HubParams p) {
  __shared__ double red[32];
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  const i64 nd = p.num_dn, nu = p.num_up;
  const i64 total = p.nrows * nd;
  double dot = 0.0;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    const i64 r = i / nd, d = i - r * nd, u = p.row0 + r;
    const uint32_t ups = p.up_states[u], dns = p.dn_states[d];
    const double xi = p.x[i];
    double acc = hub_diag<UNI>(p, ups, dns, p.e_up[u], p.e_dn[d]) * xi;
    double h = 0.0;
    const double* xr = p.x + r * nd;
    int c = p.cnt_dn[d];
    for (int k = 0; k < c; ++k) {
      uint32_t e = p.ell_dn[(i64)k * nd + d];
      double v = xr[e & ELL_TGT_MASK];
      if (UNI) h += (e >> 31) ? -v : v;
      else acc += ((e >> 31) ? -v : v) * p.hop[(e >> ELL_TGT_BITS) & 63u];
    }
    if (p.with_up) {
      c = p.cnt_up[u];
      for (int k = 0; k < c; ++k) {
        uint32_t e = p.ell_up[(i64)k * nu + u];
        double v = p.x[(i64)(e & ELL_TGT_MASK) * nd + d];
        if (UNI) h += (e >> 31) ? -v : v;
        else acc += ((e >> 31) ? -v : v) * p.hop[(e >> ELL_TGT_BITS) & 63u];
      }
    }
    if (UNI) acc += p.hop0 * h;
    hub_store<LZ>(p, i, acc, xi, s1, s2, has_prev, dot);
  }
  lz_finish<LZ>(p.lz, j, dot, red);
}

// Row kernel: dynamic smem = num_dn doubles (the staged row).
template <bool UNI, bool LZ>
__global__ void hub_row_kernel(HubParams p) {
  extern __shared__ __align__(16) double xs[];
  __shared__ double red[32];
  __shared__ uint32_t s_up[ELL_MAX_BONDS];
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  const i64 nd = p.num_dn, nu = p.num_up;
  const int tid = threadIdx.x, nt = blockDim.x;
  double dot = 0.0;
  for (i64 r = blockIdx.x; r < p.nrows; r += gridDim.x) {
    const i64 u = p.row0 + r;
    const double* __restrict__ xr = p.x + r * nd;
    // stage the row
    for (i64 d = tid; d < nd; d += nt) xs[d] = xr[d];
    const int cu = p.with_up ? (int)p.cnt_up[u] : 0;
    if (tid < cu) s_up[tid] = p.ell_up[(i64)tid * nu + u];
    __syncthreads();
    const uint32_t ups = p.up_states[u];
    const double eu = p.e_up[u];
    for (i64 d = tid; d < nd; d += nt) {
      const uint32_t dns = p.dn_states[d];
      const double xi = xs[d];
      double acc = hub_diag<UNI>(p, ups, dns, eu, p.e_dn[d]) * xi;
      double h = 0.0;
      const int c = p.cnt_dn[d];
      for (int k = 0; k < c; ++k) {
        uint32_t e = p.ell_dn[(i64)k * nd + d];
        double v = xs[e & ELL_TGT_MASK];
        if (UNI) h += (e >> 31) ? -v : v;
        else acc += ((e >> 31) ? -v : v) * p.hop[(e >> ELL_TGT_BITS) & 63u];
      }
      for (int k = 0; k < cu; ++k) {
        uint32_t e = s_up[k];
        double v = p.x[(i64)(e & ELL_TGT_MASK) * nd + d];
        if (UNI) h += (e >> 31) ? -v : v;
        else acc += ((e >> 31) ? -v : v) * p.hop[(e >> ELL_TGT_BITS) & 63u];
      }
      if (UNI) acc += p.hop0 * h;
      hub_store<LZ>(p, r * nd + d, acc, xi, s1, s2, has_prev, dot);
    }
    __syncthreads();
  }
  lz_finish<LZ>(p.lz, j, dot, red);
}

template <bool UNI>
__global__ void hub_diag_kernel(HubParams p, double* __restrict__ diag) {
  const i64 nd = p.num_dn;
  const i64 total = p.num_up * nd;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    const i64 u = i / nd, d = i - u * nd;
    diag[i] = hub_diag<UNI>(p, p.up_states[u], p.dn_states[d], p.e_up[u], p.e_dn[d]);
  }
}

