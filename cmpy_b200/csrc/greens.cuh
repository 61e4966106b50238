// greens.cuh -- K6 (ladder operators), K8 (continued fraction / pole sums), COO operator,
// slab transpose.
#pragma once
#include "common.cuh"
#include "sector.cuh"
#include "lanczos.cuh"

// ---- K8 --------------------------------------------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx cdiv_real(double num, cplx d) {  // num / d
  double s = 1.0 / (d.re * d.re + d.im * d.im);
  return cplx{num * d.re * s, -num * d.im * s};
}

// coefficients in device memory: a[n], b2[n] (b2[k] = beta_k^2 linking k-1 and k, b2[0] unused)
__global__ void __launch_bounds__(128) cf_eval_kernel(const double* __restrict__ a,
                                                     const double* __restrict__ b2, int n,
                                                     double norm2, double e0, double sgn,
                                                     const double* __restrict__ z, i64 nz,
                                                     double* __restrict__ g, int accumulate) {
  extern __shared__ double sh[];  // chunk of coefficients: a | b2
  const int CH = 512;
  double* sa = sh; double* sb = sh + CH;
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  const bool active = i < nz;
  cplx w = active ? cplx{z[2 * i], z[2 * i + 1]} : cplx{0.0, 1.0};
  // backward recurrence: t_{n-1} = w - s*(a_{n-1}-e0); t_k = w - s*(a_k-e0) - b2_{k+1}/t_{k+1}
  cplx t{0.0, 0.0};
  bool first = true;
  for (int hi = n; hi > 0; hi -= CH) {
    const int lo = hi - CH > 0 ? hi - CH : 0;
    __syncthreads();
    for (int k = threadIdx.x; k < hi - lo; k += blockDim.x) {
      sa[k] = a[lo + k];
      sb[k] = (lo + k + 1 < n) ? b2[lo + k + 1] : 0.0;  // b2_{k+1}
    }
    __syncthreads();
    for (int k = hi - lo - 1; k >= 0; --k) {
      cplx d{w.re - sgn * (sa[k] - e0), w.im};
      if (!first) {
        cplx q = cdiv_real(sb[k], t);
        d.re -= q.re; d.im -= q.im;
      }
      t = d;
      first = false;
    }
  }
  if (active) {
    cplx r = cdiv_real(norm2, t);
    if (accumulate) { g[2 * i] += r.re; g[2 * i + 1] += r.im; }
    else { g[2 * i] = r.re; g[2 * i + 1] = r.im; }
  }
}

// g[i] (+)= sum_k w[k] / (z[i] - p[k]) ; one thread per frequency, poles tiled via smem
__global__ void __launch_bounds__(128) pole_sum_kernel(const double* __restrict__ wts,
                                                      const double* __restrict__ poles, i64 np,
                                                      const double* __restrict__ z, i64 nz,
                                                      double* __restrict__ g, int accumulate) {
  __shared__ double sw[512], sp[512];
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  const bool active = i < nz;
  const double zr = active ? z[2 * i] : 0.0, zi = active ? z[2 * i + 1] : 1.0;
  double ar = 0.0, ai = 0.0;
  for (i64 base = 0; base < np; base += 512) {
    const int cnt = (int)((np - base) < 512 ? (np - base) : 512);
    __syncthreads();
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) { sw[k] = wts[base + k]; sp[k] = poles[base + k]; }
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      double dr = zr - sp[k];
      double s = sw[k] / (dr * dr + zi * zi);
      ar += s * dr; ai -= s * zi;
    }
  }
  if (active) {
    if (accumulate) { g[2 * i] += ar; g[2 * i + 1] += ai; }
    else { g[2 * i] = ar; g[2 * i + 1] = ai; }
  }
}

// ---- K6 --------------------------------------------------------------------------
struct LadderParams {
  const i64* up; i64 num_up; const i64* dn; i64 num_dn;        // origin sector
  const i64* up_t; i64 num_up_t; const i64* dn_t; i64 num_dn_t; // target sector
  int pos, sigma, dagger, signed_mode, ncomp;
  const double* x; double* y;
};

// gather formulation: one thread per target amplitude, every target written exactly once
__global__ void __launch_bounds__(256) ladder_kernel(LadderParams p) {
  const i64 total = p.num_up_t * p.num_dn_t;
  const u64 bit = 1ull << p.pos;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    const i64 ut = i / p.num_dn_t, dt = i - ut * p.num_dn_t;
    i64 src = -1;
    double sg = 1.0;
    if (p.sigma == 1) {
      const u64 st = (u64)p.up_t[ut];
      const bool has = (st & bit) != 0;
      if (has == (p.dagger != 0)) {
        const u64 ss = st ^ bit;  // origin string
        const i64 su = bsearch_left(p.up, p.num_up, (i64)ss);
        if (su < p.num_up && (u64)p.up[su] == ss) {
          src = su * p.num_dn + dt;
          if (p.signed_mode && (__popcll(ss & (bit - 1)) & 1)) sg = -1.0;
        }
      }
    } else {
      const u64 st = (u64)p.dn_t[dt];
      const bool has = (st & bit) != 0;
      if (has == (p.dagger != 0)) {
        const u64 ss = st ^ bit;
        const i64 sd = bsearch_left(p.dn, p.num_dn, (i64)ss);
        if (sd < p.num_dn && (u64)p.dn[sd] == ss) {
          src = ut * p.num_dn + sd;
          if (p.signed_mode &&
              ((__popcll((u64)p.up_t[ut]) + __popcll(ss & (bit - 1))) & 1)) sg = -1.0;
        }
      }
    }
    for (int c = 0; c < p.ncomp; ++c)
      p.y[i * p.ncomp + c] = (src >= 0) ? sg * p.x[src * p.ncomp + c] : 0.0;
  }
}

// ---- COO operator ----------------------------------------------------------------
__global__ void __launch_bounds__(256) coo_matvec_kernel(const i64* __restrict__ rows,
                                                        const i64* __restrict__ cols,
                                                        const double* __restrict__ vals, i64 nnz,
                                                        const double* __restrict__ x,
                                                        double* __restrict__ y) {
  for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < nnz; k += (i64)gridDim.x * blockDim.x)
    atomicAdd(&y[cols[k]], vals[k] * x[rows[k]]);  // y[col] += val * x[row]
}

__global__ void __launch_bounds__(256) coo_diag_kernel(const i64* __restrict__ rows,
                                                      const i64* __restrict__ cols,
                                                      const double* __restrict__ vals, i64 nnz,
                                                      double* __restrict__ diag) {
  for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < nnz; k += (i64)gridDim.x * blockDim.x)
    if (rows[k] == cols[k]) atomicAdd(&diag[rows[k]], vals[k]);
}

struct CooOp : cmpy_op_s {
  i64 nnz = 0;
  i64* d_rows = nullptr; i64* d_cols = nullptr; double* d_vals = nullptr;
  double* d_tmp = nullptr;
  ~CooOp() override { cudaFree(d_rows); cudaFree(d_cols); cudaFree(d_vals); cudaFree(d_tmp); }
  int apply(const double* x, double* y, const LzCtx& lz, cudaStream_t st) override {
    double* out = lz.enabled ? d_tmp : y;
    CU_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * size, st));
    if (nnz > 0) {
      coo_matvec_kernel<<<grid_for(nnz, 256, sm_count * 8), 256, 0, st>>>(d_rows, d_cols, d_vals,
                                                                          nnz, x, out);
      KERNEL_CHECK();
    }
    if (lz.enabled) {
      LzCtx c = lz; c.partials = d_partials; c.ticket = d_ticket;
      lz_epilogue_kernel<<<grid_for(size, 256, sm_count * 8), 256, 0, st>>>(d_tmp, x, y, size, c);
      KERNEL_CHECK();
    }
    return CMPY_OK;
  }
  int diagonal(double* d_diag, cudaStream_t st) override {
    CU_CHECK(cudaMemsetAsync(d_diag, 0, sizeof(double) * size, st));
    if (nnz > 0) {
      coo_diag_kernel<<<grid_for(nnz, 256, sm_count * 8), 256, 0, st>>>(d_rows, d_cols, d_vals, nnz, d_diag);
      KERNEL_CHECK();
    }
    return CMPY_OK;
  }
};

// ---- slab transpose / placement ---------------------------------------------------
// out[c*ld_out + r] (+)= in[r*ld_in + c]; 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, i64 nrows,
                                                       i64 ncols, i64 ld_in,
                                                       double* __restrict__ out, i64 ld_out,
                                                       int accumulate) {
  __shared__ double tile[32][33];
  const i64 tiles_c = (ncols + 31) / 32, tiles_r = (nrows + 31) / 32;
  const i64 ntiles = tiles_c * tiles_r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const i64 tr = t / tiles_c, tc = t - tr * tiles_c;
    const i64 r0 = tr * 32, c0 = tc * 32;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      i64 r = r0 + ty + k, c = c0 + tx;
      if (r < nrows && c < ncols) tile[ty + k][tx] = in[r * ld_in + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      i64 c = c0 + ty + k, r = r0 + tx;
      if (r < nrows && c < ncols) {
        double v = tile[tx][ty + k];
        i64 o = c * ld_out + r;
        out[o] = accumulate ? out[o] + v : v;
      }
    }
  }
}

// out[r*ld_out + c] (+)= in[r*ld_in + c]  (pitched block copy / accumulate)
__global__ void __launch_bounds__(256) copy2d_kernel(const double* __restrict__ in, i64 nrows,
                                                    i64 ncols, i64 ld_in, double* __restrict__ out,
                                                    i64 ld_out, int accumulate) {
  const i64 total = nrows * ncols;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    const i64 r = i / ncols, c = i - r * ncols;
    const double v = in[r * ld_in + c];
    const i64 o = r * ld_out + c;
    out[o] = accumulate ? out[o] + v : v;
  }
}
