// lanczos.cuh -- K7: fused Lanczos vector kernels + host driver + tridiagonal solver.
//
// Two-vector Lanczos with lazily normalised vectors.  Buffers hold the UNNORMALISED
// residuals r_j (q_j = r_j / beta_j, beta_j = |r_j| kept on the device):
//   step j, kernel 1 (fused into the operator's H.v kernel):
//       w = (1/beta_j) H r_j - (beta_j/beta_{j-1}) r_{j-1}      (w overwrites r_{j-1})
//       alpha_j = <r_j/beta_j, w>
//   step j, kernel 2 (lz_update_kernel):
//       r_{j+1} = w - (alpha_j/beta_j) r_j ;  beta_{j+1} = |r_{j+1}| ;  ++j
// i.e. 48 B/state per iteration, the recurrence of cmpy/exactdiag.py:324-347 in the
// normalised basis.  All reductions are deterministic (fixed-order two-stage sums), so a
// second pass regenerates bit-identical vectors for the Ritz-vector accumulation.
#pragma once
#include "common.cuh"
#include <math.h>
#include <algorithm>

// out[0] = sum x[i]*y[i]; optionally sqrt
__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ x,
                                                 const double* __restrict__ y, i64 n,
                                                 double* partials, unsigned* ticket,
                                                 double* out, int take_sqrt) {
  __shared__ double red[32];
  double acc = 0.0;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    acc += x[i] * y[i];
  double b = block_sum(acc, red);
  double total;
  if (grid_sum_last(b, partials, ticket, red, &total)) out[0] = take_sqrt ? sqrt(total) : total;
}

// kernel 2 of the Lanczos step (see header comment). X = r_j, W = w -> r_{j+1}.
__global__ void __launch_bounds__(256) lz_update_kernel(const double* __restrict__ X,
                                                       double* __restrict__ W, i64 n,
                                                       const double* alpha, double* beta,
                                                       int* iter, double* partials,
                                                       unsigned* ticket) {
  __shared__ double red[32];
  const int j = *iter;
  const double c = alpha[j] / beta[j];
  double acc = 0.0;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if ((n & 1) == 0) {
    const double2* X2 = reinterpret_cast<const double2*>(X);
    double2* W2 = reinterpret_cast<double2*>(W);
    for (; i < n / 2; i += stride) {
      double2 w = W2[i], x = X2[i];
      w.x -= c * x.x; w.y -= c * x.y;
      W2[i] = w;
      acc += w.x * w.x + w.y * w.y;
    }
  } else {
    for (; i < n; i += stride) {
      double w = W[i] - c * X[i];
      W[i] = w;
      acc += w * w;
    }
  }
  double b = block_sum(acc, red);
  double total;
  if (grid_sum_last(b, partials, ticket, red, &total)) {
    beta[j + 1] = sqrt(total);
    *iter = j + 1;
  }
}

// generic (unfused) kernel 1 epilogue for operators that produce Hx in a temp buffer
__global__ void __launch_bounds__(256) lz_epilogue_kernel(const double* __restrict__ hx,
                                                         const double* __restrict__ X,
                                                         double* __restrict__ W, i64 n, LzCtx lz) {
  __shared__ double red[32];
  const int j = *lz.iter;
  const double bj = lz.beta[j];
  const double s1 = 1.0 / bj;
  const bool has_prev = j > 0;
  const double s2 = has_prev ? bj / lz.beta[j - 1] : 0.0;
  double dot = 0.0;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    double w = s1 * hx[i];
    if (has_prev) w -= s2 * W[i];
    W[i] = w;
    dot += (s1 * X[i]) * w;
  }
  double b = block_sum(dot, red);
  double total;
  if (grid_sum_last(b, lz.partials, lz.ticket, red, &total)) lz.alpha[j] = total;
}

// psi += (ritz[j]/beta[j]) * X   (second pass, before kernel 1 of step j)
__global__ void __launch_bounds__(256) lz_accum_kernel(const double* __restrict__ X,
                                                      double* __restrict__ psi, i64 n,
                                                      const double* ritz, const double* beta,
                                                      const int* iter) {
  const int j = *iter;
  const double c = ritz[j] / beta[j];
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    psi[i] += c * X[i];
}

// ---------------------------------------------------------------------------------
// symmetric tridiagonal helpers (host)
// ---------------------------------------------------------------------------------
// number of eigenvalues of T(alpha, beta) strictly below x (Sturm count)
static int sturm_count(const double* a, const double* b, int n, double x) {
  int cnt = 0;
  double d = 1.0;
  for (int i = 0; i < n; ++i) {
    double off = (i == 0) ? 0.0 : b[i - 1] * b[i - 1];
    d = a[i] - x - (i == 0 ? 0.0 : off / d);
    if (d == 0.0) d = -1e-300;
    if (d < 0.0) ++cnt;
  }
  return cnt;
}

// k-th smallest eigenvalue (0-based) by bisection to full double precision
static double tridiag_kth(const double* a, const double* b, int n, int k) {
  double lo = a[0], hi = a[0];
  for (int i = 0; i < n; ++i) {
    double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < n - 1 ? fabs(b[i]) : 0.0);
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
  }
  double span = hi - lo;
  lo -= 1e-12 * (fabs(lo) + span) + 1e-300;
  hi += 1e-12 * (fabs(hi) + span) + 1e-300;
  for (int it = 0; it < 200; ++it) {
    double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (sturm_count(a, b, n, mid) > k) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

// eigenvector of T for eigenvalue theta by inverse iteration (banded Gaussian
// elimination with partial pivoting); returns the normalised vector in s[n].
static void tridiag_eigvec(const double* a, const double* b, int n, double theta, double* s) {
  if (n == 1) { s[0] = 1.0; return; }
  std::vector<double> u0(n), u1(n), u2(n), y(n), rhs(n);
  double nrm = 0.0;
  for (int i = 0; i < n; ++i) nrm = std::max(nrm, fabs(a[i]) + (i < n - 1 ? fabs(b[i]) : 0.0));
  const double tiny = std::max(nrm, 1.0) * 2.3e-16;
  for (int i = 0; i < n; ++i) rhs[i] = ((i & 1) ? 0.9 : 1.1) / sqrt((double)n);
  for (int rep = 0; rep < 4; ++rep) {
    // current (partially eliminated) row i: p0 at col i, p1 at col i+1, p2 at col i+2
    double p0 = a[0] - theta, p1 = b[0], p2 = 0.0, pr = rhs[0];
    for (int i = 0; i < n - 1; ++i) {
      const double q0 = b[i];                              // row i+1, col i
      const double q1 = a[i + 1] - theta;                  //          col i+1
      const double q2 = (i + 1 < n - 1) ? b[i + 1] : 0.0;  //          col i+2
      const double qr = rhs[i + 1];
      if (fabs(p0) >= fabs(q0)) {
        if (fabs(p0) < tiny) p0 = (p0 < 0 ? -tiny : tiny);
        const double f = q0 / p0;
        u0[i] = p0; u1[i] = p1; u2[i] = p2; y[i] = pr;
        p0 = q1 - f * p1; p1 = q2 - f * p2; p2 = 0.0; pr = qr - f * pr;
      } else {
        const double f = p0 / q0;
        u0[i] = q0; u1[i] = q1; u2[i] = q2; y[i] = qr;
        p0 = p1 - f * q1; p1 = p2 - f * q2; p2 = 0.0; pr = pr - f * qr;
      }
    }
    if (fabs(p0) < tiny) p0 = (p0 < 0 ? -tiny : tiny);
    u0[n - 1] = p0; u1[n - 1] = 0.0; u2[n - 1] = 0.0; y[n - 1] = pr;
    rhs[n - 1] = y[n - 1] / u0[n - 1];
    rhs[n - 2] = (y[n - 2] - u1[n - 2] * rhs[n - 1]) / u0[n - 2];
    for (int i = n - 3; i >= 0; --i)
      rhs[i] = (y[i] - u1[i] * rhs[i + 1] - u2[i] * rhs[i + 2]) / u0[i];
    double nn = 0.0, mx = 0.0;
    for (int i = 0; i < n; ++i) mx = std::max(mx, fabs(rhs[i]));
    if (!(mx > 0.0) || !std::isfinite(mx)) {
      for (int i = 0; i < n; ++i) rhs[i] = (i == 0) ? 1.0 : 0.0;
      mx = 1.0;
    }
    for (int i = 0; i < n; ++i) { rhs[i] /= mx; nn += rhs[i] * rhs[i]; }
    nn = sqrt(nn);
    for (int i = 0; i < n; ++i) rhs[i] /= nn;
  }
  double sg = 1.0;
  for (int i = 0; i < n; ++i) if (fabs(rhs[i]) > 1e-8) { sg = rhs[i] < 0 ? -1.0 : 1.0; break; }
  for (int i = 0; i < n; ++i) s[i] = sg * rhs[i];
}

// ---------------------------------------------------------------------------------
// driver
// ---------------------------------------------------------------------------------
struct LanczosRun {
  cmpy_op_s* op;
  cudaStream_t st;
  i64 n;
  LzCtx ctx;
  int upd_grid;

  int step(const double* X, double* W) {
    int rc = op->apply(X, W, ctx, st);
    if (rc) return rc;
    lz_update_kernel<<<upd_grid, 256, 0, st>>>(X, W, n, op->d_alpha, op->d_beta, op->d_iter,
                                               op->d_partials, op->d_ticket);
    KERNEL_CHECK();
    return CMPY_OK;
  }
};

static int lanczos_run_impl(cmpy_op_s* op, const double* d_v0, double* d_w0, double* d_w1,
                            int maxit, double tol, double resid_tol, int check_every,
                            int use_graph, double* h_alpha, double* h_beta, int* h_nit, double* h_e0,
                            double* h_resid, double* d_eigvec, cudaStream_t st) {
  ARG_CHECK(op && d_v0 && d_w0 && d_w1 && h_alpha && h_beta && h_nit && h_e0, "null argument");
  ARG_CHECK(maxit >= 1, "maxit must be >= 1");
  const i64 n = op->size;
  if (check_every < 2) check_every = 2;
  if (check_every & 1) ++check_every;
  int rc = op->ensure_lz_capacity(maxit + check_every);
  if (rc) return rc;
  LanczosRun run;
  run.op = op; run.st = st; run.n = n;
  run.ctx.enabled = 1; run.ctx.iter = op->d_iter; run.ctx.beta = op->d_beta;
  run.ctx.alpha = op->d_alpha; run.ctx.partials = op->d_partials; run.ctx.ticket = op->d_ticket;
  run.upd_grid = grid_for((n + 1) / 2, 256, op->sm_count * 8);

  std::vector<double> alpha(maxit + check_every + 2), beta(maxit + check_every + 2);
  std::vector<double> svec;
  double e0 = 0.0, e0_prev = 0.0, resid = 1e300;
  int m_final = 0;
  bool converged = false;

  auto init_pass = [&]() -> int {
    CU_CHECK(cudaMemsetAsync(op->d_iter, 0, sizeof(int) * 4, st));
    CU_CHECK(cudaMemsetAsync(op->d_ticket, 0, sizeof(unsigned) * 4, st));
    CU_CHECK(cudaMemcpyAsync(d_w0, d_v0, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    dot_kernel<<<grid_for(n, 256, op->sm_count * 8), 256, 0, st>>>(
        d_w0, d_w0, n, op->d_partials, op->d_ticket, op->d_beta, 1);
    KERNEL_CHECK();
    return CMPY_OK;
  };

  // graph of two steps (A->B, B->A)
  cudaGraphExec_t gexec = nullptr;
  auto two_steps = [&]() -> int {
    int r = run.step(d_w0, d_w1);
    if (r) return r;
    return run.step(d_w1, d_w0);
  };
  auto build_graph = [&]() -> int {
    if (!use_graph) return CMPY_OK;
    cudaGraph_t g = nullptr;
    cudaStream_t cs;
    CU_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaStream_t saved = run.st;
    run.st = cs;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    int r = CMPY_OK;
    if (e == cudaSuccess) {
      r = two_steps();
      e = cudaStreamEndCapture(cs, &g);
    }
    run.st = saved;
    if (e != cudaSuccess || r != CMPY_OK || !g) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      cudaStreamDestroy(cs);
      gexec = nullptr;  // fall back to plain launches
      return CMPY_OK;
    }
    e = cudaGraphInstantiate(&gexec, g, 0);
    cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (e != cudaSuccess) { cudaGetLastError(); gexec = nullptr; }
    return CMPY_OK;
  };

  rc = init_pass();
  if (rc) return rc;
  rc = build_graph();
  if (rc) return rc;

  int done = 0;  // completed iterations
  while (done < maxit && !converged) {
    int todo = std::min(check_every, ((maxit - done) + 1) & ~1);
    for (int k = 0; k < todo; k += 2) {
      if (gexec) {
        CU_CHECK(cudaGraphLaunch(gexec, st));
        g_cmpy_launches.fetch_add(4);
      } else {
        rc = two_steps();
        if (rc) { if (gexec) cudaGraphExecDestroy(gexec); return rc; }
      }
    }
    done += todo;
    CU_CHECK(cudaMemcpyAsync(alpha.data(), op->d_alpha, sizeof(double) * done,
                             cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(beta.data(), op->d_beta, sizeof(double) * (done + 1),
                             cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    // usable length: stop at a breakdown (invariant subspace) or non-finite value
    int m = std::min(done, maxit);
    double scale = 0.0;
    for (int i = 0; i < m; ++i) {
      if (!std::isfinite(alpha[i]) || !std::isfinite(beta[i + 1])) { m = i; converged = true; break; }
      scale = std::max(scale, fabs(alpha[i]) + fabs(beta[i + 1]));
      if (beta[i + 1] <= 1e-13 * std::max(scale, 1.0)) { m = i + 1; converged = true; break; }
    }
    if (m < 1) { m = 1; }
    // T_m = tridiag(alpha[0..m), beta[1..m))
    e0 = tridiag_kth(alpha.data(), beta.data() + 1, m, 0);
    svec.assign(m, 0.0);
    tridiag_eigvec(alpha.data(), beta.data() + 1, m, e0, svec.data());
    resid = fabs(beta[m] * svec[m - 1]);
    double gap = 1.0;
    if (m >= 2) gap = std::max(tridiag_kth(alpha.data(), beta.data() + 1, m, 1) - e0, 1e-8);
    m_final = m;
    const bool res_ok = (resid_tol > 0.0) ? (resid < resid_tol) : (resid * resid / gap < tol);
    if (done > check_every && fabs(e0 - e0_prev) < tol && res_ok) converged = true;
    if (tol > 0.0 && resid < 1e-14 * std::max(1.0, fabs(e0))) converged = true;
    e0_prev = e0;
  }
  if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }

  for (int i = 0; i < m_final; ++i) h_alpha[i] = alpha[i];
  for (int i = 0; i <= m_final; ++i) h_beta[i] = beta[i];
  *h_nit = m_final;
  *h_e0 = e0;
  if (h_resid) *h_resid = resid;

  if (d_eigvec) {
    // second pass: regenerate r_j (bit-identical) and accumulate psi = sum_j s_j r_j/beta_j
    CU_CHECK(cudaMemcpyAsync(op->d_ritz, svec.data(), sizeof(double) * m_final,
                             cudaMemcpyHostToDevice, st));
    CU_CHECK(cudaMemsetAsync(d_eigvec, 0, sizeof(double) * n, st));
    rc = init_pass();
    if (rc) return rc;
    const int ag = grid_for(n, 256, op->sm_count * 8);
    double* X = d_w0; double* W = d_w1;
    for (int j = 0; j < m_final; ++j) {
      lz_accum_kernel<<<ag, 256, 0, st>>>(X, d_eigvec, n, op->d_ritz, op->d_beta, op->d_iter);
      KERNEL_CHECK();
      if (j == m_final - 1) break;
      rc = run.step(X, W);
      if (rc) return rc;
      std::swap(X, W);
    }
    CU_CHECK(cudaStreamSynchronize(st));
  }
  return converged ? CMPY_OK : CMPY_ERR_NOT_CONVERGED;
}
