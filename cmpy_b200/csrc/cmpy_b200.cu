// cmpy_b200.cu -- C ABI of libcmpy_b200.so (see include/cmpy_b200.h).
// Single translation unit; build: cmpy_b200/csrc/Makefile (nvcc, sm_100a).
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "sector.cuh"
#include "hubbard.cuh"
#include "hubbard_seg.cuh"
#include "hubbard_cls.cuh"
#include "hubbard_op.cuh"
#include "lanczos.cuh"
#include "greens.cuh"
#include "heisenberg.cuh"
#include "peer.cuh"
#include "dist.cuh"

#define API extern "C" __attribute__((visibility("default")))

API const char* cmpy_last_error(void) { return g_cmpy_err.c_str(); }
API int cmpy_version(void) { return 100; }
API int64_t cmpy_launch_count(void) { return g_cmpy_launches.load(); }
API void cmpy_reset_launch_count(void) { g_cmpy_launches.store(0); }

API int cmpy_device_info(int* sm_count, int64_t* l2_bytes, int64_t* smem_per_block_optin) {
  int dev = 0;
  CU_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (l2_bytes) *l2_bytes = prop.l2CacheSize;
  if (smem_per_block_optin) *smem_per_block_optin = (int64_t)prop.sharedMemPerBlockOptin;
  return CMPY_OK;
}

// ---- K1 --------------------------------------------------------------------------
API int cmpy_binomial(int n, int k, int64_t* h_out) {
  ARG_CHECK(h_out, "null output");
  ARG_CHECK(n >= 0 && n < BINOM_N, "binomial: 0 <= n <= 64");
  *h_out = (k < 0 || k > n) ? 0 : (int64_t)host_binom()[n * BINOM_N + k];
  return CMPY_OK;
}

API int cmpy_sector_enumerate(int num_sites, int n, int64_t* d_states, void* stream) {
  ARG_CHECK(num_sites >= 0 && num_sites <= 62, "enumerate: 0 <= num_sites <= 62");
  ARG_CHECK(n >= 0 && n <= num_sites, "enumerate: 0 <= n <= num_sites");
  ARG_CHECK(d_states, "null output");
  int rc = ensure_binom_uploaded();
  if (rc) return rc;
  i64 count = (i64)host_binom()[num_sites * BINOM_N + n];
  sector_enumerate_kernel<<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(
      num_sites, n, count, (i64*)d_states);
  KERNEL_CHECK();
  return CMPY_OK;
}

API int cmpy_sector_rank(const int64_t* d_states, int64_t m, int64_t* d_idx, void* stream) {
  ARG_CHECK(m >= 0, "rank: negative count");
  if (m == 0) return CMPY_OK;
  ARG_CHECK(d_states && d_idx, "null pointer");
  int rc = ensure_binom_uploaded();
  if (rc) return rc;
  sector_rank_kernel<<<grid_for(m, 256), 256, 0, as_stream(stream)>>>((const i64*)d_states, m,
                                                                      (i64*)d_idx);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- K2 / K3 -----------------------------------------------------------------------
API int cmpy_species_hops(const int64_t* d_states, int64_t num, int fixed_popcount, int width,
                          int site1, int site2, int32_t* d_target, int8_t* d_sign, void* stream) {
  ARG_CHECK(num >= 0, "negative count");
  // the reference asserts site1 < site2 (cmpy/operators.py:438)
  ARG_CHECK(site1 >= 0 && site1 < site2 && site2 < 63, "species_hops: need 0 <= site1 < site2");
  if (num == 0) return CMPY_OK;
  ARG_CHECK(d_states && d_target && d_sign, "null pointer");
  ARG_CHECK(num < (1ll << 31), "string list too long");
  int rc = ensure_binom_uploaded();
  if (rc) return rc;
  species_hops_kernel<<<grid_for(num, 128), 128, 0, as_stream(stream)>>>(
      (const i64*)d_states, num, fixed_popcount, width, site1, site2, d_target, d_sign);
  KERNEL_CHECK();
  return CMPY_OK;
}

static int fill_site_values(const double* h_values, int nvalues, SiteValues& sv) {
  ARG_CHECK(nvalues >= 0 && nvalues <= 64, "site values: 0 <= n <= 64");
  ARG_CHECK(nvalues == 0 || h_values, "null values");
  sv.n = nvalues;
  for (int i = 0; i < 64; ++i) sv.v[i] = i < nvalues ? h_values[i] : 0.0;
  return CMPY_OK;
}

API int cmpy_weighted_elements(const int64_t* d_states, int64_t num, const double* h_values,
                               int nvalues, double* d_out, void* stream) {
  SiteValues sv;
  int rc = fill_site_values(h_values, nvalues, sv);
  if (rc) return rc;
  if (num == 0) return CMPY_OK;
  ARG_CHECK(d_states && d_out && num > 0, "bad arguments");
  weighted_elements_kernel<<<grid_for(num, 256), 256, 0, as_stream(stream)>>>(
      (const i64*)d_states, num, sv, d_out);
  KERNEL_CHECK();
  return CMPY_OK;
}

API int cmpy_inter_elements(const int64_t* d_up, int64_t num_up, const int64_t* d_dn,
                            int64_t num_dn, const double* h_u, int nvalues, double* d_out,
                            void* stream) {
  SiteValues sv;
  int rc = fill_site_values(h_u, nvalues, sv);
  if (rc) return rc;
  if (num_up * num_dn == 0) return CMPY_OK;
  ARG_CHECK(d_up && d_dn && d_out && num_up > 0 && num_dn > 0, "bad arguments");
  inter_elements_kernel<<<grid_for(num_up * num_dn, 256), 256, 0, as_stream(stream)>>>(
      (const i64*)d_up, num_up, (const i64*)d_dn, num_dn, sv, d_out);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- K4 --------------------------------------------------------------------------
API int cmpy_hubbard_create(int num_sites, const int64_t* h_up_states, int64_t num_up,
                            const int64_t* h_dn_states, int64_t num_dn, int fixed_popcount,
                            int nbonds, const int32_t* h_bonds, const double* h_hop,
                            const double* h_eps, const double* h_u, int sign_width,
                            cmpy_op_t* out) {
  ARG_CHECK(out, "null output handle");
  *out = nullptr;
  ARG_CHECK(num_sites >= 1 && num_sites <= 32, "hubbard: 1 <= num_sites <= 32");
  ARG_CHECK(h_up_states && h_dn_states && num_up >= 1 && num_dn >= 1, "hubbard: empty sector");
  ARG_CHECK(nbonds >= 0 && nbonds <= ELL_MAX_BONDS, "hubbard: 0 <= nbonds <= 64");
  ARG_CHECK(nbonds == 0 || (h_bonds && h_hop), "hubbard: null bond arrays");
  ARG_CHECK(h_eps && h_u, "hubbard: null eps/u");
  int rc = ensure_binom_uploaded();
  if (rc) return rc;
  BondList bl;
  bl.n = nbonds;
  for (int b = 0; b < nbonds; ++b) {
    bl.s1[b] = h_bonds[2 * b];
    bl.s2[b] = h_bonds[2 * b + 1];
    // the reference asserts site1 < site2 (cmpy/operators.py:438)
    ARG_CHECK(bl.s1[b] >= 0 && bl.s1[b] < bl.s2[b] && bl.s2[b] < num_sites,
              "hubbard: bond must satisfy 0 <= site1 < site2 < num_sites");
  }
  SiteValues eps;
  rc = fill_site_values(h_eps, num_sites, eps);
  if (rc) return rc;
  HubbardOp* op = new HubbardOp();
  rc = op->init_workspace();
  if (rc) { delete op; return rc; }
  op->num_sites = num_sites; op->nbonds = nbonds; op->sign_width = sign_width;
  op->size = num_up * num_dn;
  rc = op->build_species(op->up, (const i64*)h_up_states, num_up, fixed_popcount, bl, eps);
  if (!rc) rc = op->build_species(op->dn, (const i64*)h_dn_states, num_dn, fixed_popcount, bl, eps);
  if (rc) { delete op; return rc; }
  bool uni = true, eps_uni = true;
  for (int i = 1; i < num_sites; ++i) eps_uni = eps_uni && (h_eps[i] == h_eps[0]);
  op->eps_uniform = eps_uni;
  for (int i = 1; i < num_sites; ++i) uni = uni && (h_u[i] == h_u[0]);
  for (int b = 1; b < nbonds; ++b) uni = uni && (h_hop[b] == h_hop[0]);
  op->uniform = uni;
  op->u0 = h_u[0];
  op->hop0 = nbonds > 0 ? h_hop[0] : 0.0;
  cudaError_t e = cudaMalloc(&op->d_hop, sizeof(double) * ELL_MAX_BONDS);  // zero padded
  if (e == cudaSuccess) e = cudaMemset(op->d_hop, 0, sizeof(double) * ELL_MAX_BONDS);
  if (e == cudaSuccess) e = cudaMalloc(&op->d_u, sizeof(double) * num_sites);
  if (e == cudaSuccess && nbonds > 0)
    e = cudaMemcpy(op->d_hop, h_hop, sizeof(double) * nbonds, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(op->d_u, h_u, sizeof(double) * num_sites, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete op;
    return cmpy_fail(CMPY_ERR_CUDA, std::string("hubbard_create: ") + cudaGetErrorString(e));
  }
  rc = uni ? op->configure_row<true>() : op->configure_row<false>();
  if (!rc && fixed_popcount)
    rc = op->configure_seg(__builtin_popcountll((unsigned long long)h_dn_states[0]), bl.s1, bl.s2, h_eps);
  if (!rc && fixed_popcount)
    rc = op->configure_cls(__builtin_popcountll((unsigned long long)h_dn_states[0]), bl.s1, bl.s2, h_eps);
  if (!rc && fixed_popcount)
    rc = op->configure_eng(__builtin_popcountll((unsigned long long)h_dn_states[0]), bl.s1, bl.s2, h_eps);
  if (!rc && fixed_popcount)
    rc = op->configure_long(__builtin_popcountll((unsigned long long)h_dn_states[0]), bl.s1, bl.s2, h_eps);
  if (rc) { delete op; return rc; }
  *out = op;
  return CMPY_OK;
}

API int cmpy_heisenberg_create(int num_sites, int n_up, int npairs, const int32_t* h_pairs,
                               double j, double jz, cmpy_op_t* out) {
  ARG_CHECK(out, "null output handle");
  *out = nullptr;
  ARG_CHECK(npairs >= 0 && (npairs == 0 || h_pairs), "heisenberg: bad pair list");
  HeisenbergOp* op = new HeisenbergOp();
  int rc = op->init_workspace();
  if (!rc) rc = op->build(num_sites, n_up, npairs, h_pairs, j, jz);
  if (rc) { delete op; return rc; }
  *out = op;
  return CMPY_OK;
}

API int cmpy_coo_create(int64_t size, int64_t nnz, const int64_t* h_rows, const int64_t* h_cols,
                        const double* h_vals, cmpy_op_t* out) {
  ARG_CHECK(out, "null output handle");
  *out = nullptr;
  ARG_CHECK(size >= 1 && nnz >= 0, "coo: bad size");
  ARG_CHECK(nnz == 0 || (h_rows && h_cols && h_vals), "coo: null arrays");
  for (i64 k = 0; k < nnz; ++k)
    ARG_CHECK(h_rows[k] >= 0 && h_rows[k] < size && h_cols[k] >= 0 && h_cols[k] < size,
              "coo: index out of range");
  CooOp* op = new CooOp();
  int rc = op->init_workspace();
  if (rc) { delete op; return rc; }
  op->size = size; op->nnz = nnz;
  size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  cudaError_t e = cudaMalloc(&op->d_rows, sizeof(i64) * nn);
  if (e == cudaSuccess) e = cudaMalloc(&op->d_cols, sizeof(i64) * nn);
  if (e == cudaSuccess) e = cudaMalloc(&op->d_vals, sizeof(double) * nn);
  if (e == cudaSuccess) e = cudaMalloc(&op->d_tmp, sizeof(double) * size);
  if (e == cudaSuccess && nnz > 0) {
    e = cudaMemcpy(op->d_rows, h_rows, sizeof(i64) * nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(op->d_cols, h_cols, sizeof(i64) * nnz, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(op->d_vals, h_vals, sizeof(double) * nnz, cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    delete op;
    return cmpy_fail(CMPY_ERR_CUDA, std::string("coo_create: ") + cudaGetErrorString(e));
  }
  *out = op;
  return CMPY_OK;
}

// ---- operator calls ----------------------------------------------------------------
API int cmpy_op_destroy(cmpy_op_t op) {
  if (op) delete op;
  return CMPY_OK;
}

API int cmpy_op_size(cmpy_op_t op, int64_t* h_size) {
  ARG_CHECK(op && h_size, "null argument");
  *h_size = op->size;
  return CMPY_OK;
}

static LzCtx no_lz() {
  LzCtx c;
  c.enabled = 0; c.iter = nullptr; c.beta = nullptr; c.alpha = nullptr; c.partials = nullptr;
  c.ticket = nullptr;
  return c;
}

API int cmpy_hv_apply(cmpy_op_t op, const double* d_x, double* d_y, void* stream) {
  NvtxRange nvtx_("cmpy_hv_apply");
  ARG_CHECK(op && d_x && d_y, "null argument");
  ARG_CHECK(d_x != d_y, "hv_apply: x and y must be distinct buffers");
  return op->apply(d_x, d_y, no_lz(), as_stream(stream));
}

API int cmpy_hubbard_apply_rows(cmpy_op_t op, const double* d_x_slab, double* d_y_slab,
                                int64_t row0, int64_t nrows, int accumulate, void* stream) {
  NvtxRange nvtx_("cmpy_hubbard_apply_rows");
  ARG_CHECK(op && d_x_slab && d_y_slab, "null argument");
  HubbardOp* h = dynamic_cast<HubbardOp*>(op);
  ARG_CHECK(h, "apply_rows: not a Hubbard operator");
  ARG_CHECK(row0 >= 0 && nrows >= 0 && row0 + nrows <= h->up.num, "apply_rows: bad row range");
  return h->apply_slab(d_x_slab, d_y_slab, row0, nrows, 0, accumulate, no_lz(), as_stream(stream));
}

API int cmpy_hv_set_variant(cmpy_op_t op, int variant) {
  ARG_CHECK(op && variant >= 0 && variant <= 11, "bad variant");
  op->variant = variant;
  return CMPY_OK;
}

API int cmpy_hubbard_set_grid_limit(cmpy_op_t op, int max_ctas) {
  ARG_CHECK(op && max_ctas >= 0, "bad argument");
  HubbardOp* h = dynamic_cast<HubbardOp*>(op);
  ARG_CHECK(h, "set_grid_limit: not a Hubbard operator");
  h->grid_limit = max_ctas;
  return CMPY_OK;
}

API int cmpy_op_diagonal(cmpy_op_t op, double* d_diag, void* stream) {
  ARG_CHECK(op && d_diag, "null argument");
  return op->diagonal(d_diag, as_stream(stream));
}

__global__ void __launch_bounds__(256) sum_kernel(const double* __restrict__ x, i64 n,
                                                 double* partials, unsigned* ticket, double* out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    acc += x[i];
  double b = block_sum(acc, red);
  double total;
  if (grid_sum_last(b, partials, ticket, red, &total)) out[0] = total;
}

API int cmpy_op_trace(cmpy_op_t op, double* h_trace) {
  ARG_CHECK(op && h_trace, "null argument");
  double* d_diag = nullptr;
  CU_CHECK(cudaMalloc(&d_diag, sizeof(double) * (op->size + 1)));
  int rc = op->diagonal(d_diag, 0);
  if (!rc) {
    sum_kernel<<<grid_for(op->size, 256, op->sm_count * 4), 256>>>(d_diag, op->size, op->d_partials,
                                                                 op->d_ticket, d_diag + op->size);
    g_cmpy_launches.fetch_add(1);
    cudaError_t e = cudaMemcpy(h_trace, d_diag + op->size, sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cmpy_fail(CMPY_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(d_diag);
  return rc;
}

// ---- K6 --------------------------------------------------------------------------
API int cmpy_ladder_apply(const int64_t* d_up, int64_t num_up, const int64_t* d_dn, int64_t num_dn,
                          const int64_t* d_up_t, int64_t num_up_t, const int64_t* d_dn_t,
                          int64_t num_dn_t, int pos, int sigma, int dagger, int signed_mode,
                          int ncomp, const double* d_x, double* d_y, void* stream) {
  NvtxRange nvtx_("cmpy_ladder_apply");
  ARG_CHECK(d_up && d_dn && d_up_t && d_dn_t && d_x && d_y, "null argument");
  ARG_CHECK(sigma == 1 || sigma == 2, "sigma must be UP=1 or DN=2");
  ARG_CHECK(pos >= 0 && pos < 63, "bad site");
  ARG_CHECK(ncomp == 1 || ncomp == 2, "ncomp must be 1 (real) or 2 (complex)");
  ARG_CHECK(num_up >= 1 && num_dn >= 1 && num_up_t >= 1 && num_dn_t >= 1, "empty sector");
  if (sigma == 1) ARG_CHECK(num_dn == num_dn_t, "sigma=UP: dn lists must coincide");
  if (sigma == 2) ARG_CHECK(num_up == num_up_t, "sigma=DN: up lists must coincide");
  LadderParams p;
  p.up = (const i64*)d_up; p.num_up = num_up; p.dn = (const i64*)d_dn; p.num_dn = num_dn;
  p.up_t = (const i64*)d_up_t; p.num_up_t = num_up_t; p.dn_t = (const i64*)d_dn_t; p.num_dn_t = num_dn_t;
  p.pos = pos; p.sigma = sigma; p.dagger = dagger; p.signed_mode = signed_mode; p.ncomp = ncomp;
  p.x = d_x; p.y = d_y;
  ladder_kernel<<<grid_for(num_up_t * num_dn_t, 256), 256, 0, as_stream(stream)>>>(p);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- K7 --------------------------------------------------------------------------
API int cmpy_dot(cmpy_op_t op, const double* d_x, const double* d_y, int64_t n, double* d_out,
                 void* stream) {
  ARG_CHECK(op && d_x && d_y && d_out && n >= 0, "bad argument");
  dot_kernel<<<grid_for(n, 256, op->sm_count * 8), 256, 0, as_stream(stream)>>>(
      d_x, d_y, n, op->d_partials, op->d_ticket, d_out, 0);
  KERNEL_CHECK();
  return CMPY_OK;
}

API int cmpy_lanczos_run(cmpy_op_t op, const double* d_v0, double* d_w0, double* d_w1, int maxit,
                         double tol, double resid_tol, int check_every, int use_graph,
                         double* h_alpha, double* h_beta, int* h_nit, double* h_e0,
                         double* h_resid, double* d_eigvec, void* stream) {
  NvtxRange nvtx_("cmpy_lanczos_run");
  return lanczos_run_impl(op, d_v0, d_w0, d_w1, maxit, tol, resid_tol, check_every, use_graph,
                          h_alpha, h_beta, h_nit, h_e0, h_resid, d_eigvec, as_stream(stream));
}

API int cmpy_tridiag_lowest(const double* h_alpha, const double* h_beta, int n, int k,
                            double* h_evals, double* h_evec0) {
  ARG_CHECK(h_alpha && n >= 1 && k >= 1 && h_evals, "bad argument");
  ARG_CHECK(n == 1 || h_beta, "null beta");
  if (k > n) k = n;
  for (int i = 0; i < k; ++i) h_evals[i] = tridiag_kth(h_alpha, h_beta, n, i);
  if (h_evec0) tridiag_eigvec(h_alpha, h_beta, n, h_evals[0], h_evec0);
  return CMPY_OK;
}

// ---- K8 --------------------------------------------------------------------------
API int cmpy_cf_eval(const double* h_alpha, const double* h_beta, int n, double norm2, double e0,
                     int s, const double* d_z, int64_t nz, double* d_g, int accumulate,
                     void* stream) {
  NvtxRange nvtx_("cmpy_cf_eval");
  ARG_CHECK(h_alpha && n >= 1 && d_z && d_g && nz >= 0, "bad argument");
  ARG_CHECK(n == 1 || h_beta, "null beta");
  ARG_CHECK(s == 1 || s == -1, "s must be +1 or -1");
  if (nz == 0) return CMPY_OK;
  cudaStream_t st = as_stream(stream);
  std::vector<double> buf(2 * (size_t)n);
  for (int k = 0; k < n; ++k) {
    buf[k] = h_alpha[k];
    buf[n + k] = (k == 0) ? 0.0 : h_beta[k - 1] * h_beta[k - 1];
  }
  double* d_coef = nullptr;
  CU_CHECK(cudaMallocAsync(&d_coef, sizeof(double) * 2 * n, st));
  CU_CHECK(cudaMemcpyAsync(d_coef, buf.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st));
  CU_CHECK(cudaStreamSynchronize(st));  // buf is a local
  cf_eval_kernel<<<(int)((nz + 127) / 128), 128, sizeof(double) * 1024, st>>>(
      d_coef, d_coef + n, n, norm2, e0, (double)s, d_z, nz, d_g, accumulate);
  KERNEL_CHECK();
  CU_CHECK(cudaFreeAsync(d_coef, st));
  return CMPY_OK;
}

API int cmpy_pole_sum(const double* d_weights, const double* d_poles, int64_t npoles,
                      const double* d_z, int64_t nz, double* d_g, int accumulate, void* stream) {
  ARG_CHECK(d_z && d_g && nz >= 0 && npoles >= 0, "bad argument");
  ARG_CHECK(npoles == 0 || (d_weights && d_poles), "null poles");
  if (nz == 0) return CMPY_OK;
  pole_sum_kernel<<<(int)((nz + 127) / 128), 128, 0, as_stream(stream)>>>(
      d_weights, d_poles, npoles, d_z, nz, d_g, accumulate);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- K9 building blocks -------------------------------------------------------------
API int cmpy_transpose(const double* d_in, int64_t nrows, int64_t ncols, int64_t ld_in,
                       double* d_out, int64_t ld_out, int accumulate, void* stream) {
  ARG_CHECK(d_in && d_out && nrows >= 0 && ncols >= 0 && ld_in >= ncols && ld_out >= nrows,
            "bad argument");
  if (nrows == 0 || ncols == 0) return CMPY_OK;
  i64 ntiles = ((nrows + 31) / 32) * ((ncols + 31) / 32);
  int g = (int)(ntiles < 148 * 16 ? ntiles : 148 * 16);
  transpose_kernel<<<g, 256, 0, as_stream(stream)>>>(d_in, nrows, ncols, ld_in, d_out, ld_out,
                                                     accumulate);
  KERNEL_CHECK();
  return CMPY_OK;
}

API int cmpy_copy2d(const double* d_in, int64_t nrows, int64_t ncols, int64_t ld_in,
                    double* d_out, int64_t ld_out, int accumulate, void* stream) {
  ARG_CHECK(d_in && d_out && nrows >= 0 && ncols >= 0 && ld_in >= ncols && ld_out >= ncols,
            "bad argument");
  if (nrows == 0 || ncols == 0) return CMPY_OK;
  copy2d_kernel<<<grid_for(nrows * ncols, 256, 148 * 16), 256, 0, as_stream(stream)>>>(
      d_in, nrows, ncols, ld_in, d_out, ld_out, accumulate);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- K9 over peer memory -----------------------------------------------------------------
static int fill_peer_table(PeerTable& pt, int world, const int64_t* h_col_bounds, void* const* h_peer_ptrs) {
  ARG_CHECK(world >= 1 && world <= PEER_MAX, "peer transpose: 1 <= world <= 16");
  ARG_CHECK(h_col_bounds && h_peer_ptrs, "peer transpose: null tables");
  pt.world = world;
  for (int q = 0; q < world; ++q) {
    ARG_CHECK(h_peer_ptrs[q], "peer transpose: null peer pointer");
    ARG_CHECK(h_col_bounds[q] <= h_col_bounds[q + 1], "peer transpose: column bounds must ascend");
    pt.base[q] = (double*)h_peer_ptrs[q];
    pt.cb[q] = h_col_bounds[q];
  }
  pt.cb[world] = h_col_bounds[world];
  return CMPY_OK;
}

API int cmpy_transpose_push(const double* d_x_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                            int64_t ld_t, int world, const int64_t* h_col_bounds,
                            void* const* h_peer_ptrs, void* stream) {
  return cmpy_transpose_push_capped(d_x_slab, nrows, num_dn, row0, ld_t, world, h_col_bounds, h_peer_ptrs, 0,
                                    stream);
}

API int cmpy_transpose_push_capped(const double* d_x_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                                   int64_t ld_t, int world, const int64_t* h_col_bounds,
                                   void* const* h_peer_ptrs, int max_ctas, void* stream) {
  NvtxRange nvtx_("cmpy_transpose_push");
  ARG_CHECK(max_ctas >= 0, "bad argument");
  ARG_CHECK(d_x_slab && nrows >= 0 && num_dn >= 0 && row0 >= 0 && ld_t >= row0 + nrows, "bad argument");
  PeerTable pt;
  int rc = fill_peer_table(pt, world, h_col_bounds, h_peer_ptrs);
  if (rc) return rc;
  ARG_CHECK(pt.cb[0] == 0 && pt.cb[world] == num_dn, "peer transpose: bounds must cover the columns");
  if (nrows == 0 || num_dn == 0) return CMPY_OK;
  const i64 ntiles = ((nrows + 127) / 128) * ((num_dn + 31) / 32);
  int g = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
  if (max_ctas > 0 && g > max_ctas) g = max_ctas;
  peer_transpose_kernel<false, 128><<<g, 256, 0, as_stream(stream)>>>(const_cast<double*>(d_x_slab), nrows,
                                                                      num_dn, row0, ld_t, pt);
  KERNEL_CHECK();
  return CMPY_OK;
}

API int cmpy_transpose_pull_acc(double* d_y_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                                int64_t ld_t, int world, const int64_t* h_col_bounds,
                                void* const* h_peer_ptrs, void* stream) {
  NvtxRange nvtx_("cmpy_transpose_pull_acc");
  ARG_CHECK(d_y_slab && nrows >= 0 && num_dn >= 0 && row0 >= 0 && ld_t >= row0 + nrows, "bad argument");
  PeerTable pt;
  int rc = fill_peer_table(pt, world, h_col_bounds, h_peer_ptrs);
  if (rc) return rc;
  ARG_CHECK(pt.cb[0] == 0 && pt.cb[world] == num_dn, "peer transpose: bounds must cover the columns");
  if (nrows == 0 || num_dn == 0) return CMPY_OK;
  const i64 ntiles = ((nrows + 31) / 32) * ((num_dn + 31) / 32);
  const int g = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
  peer_transpose_kernel<true, 32><<<g, 256, 0, as_stream(stream)>>>(d_y_slab, nrows, num_dn, row0, ld_t, pt);
  KERNEL_CHECK();
  return CMPY_OK;
}

// ---- sharded H.v / sharded Lanczos as single C calls (dist.cuh) ----------------------------------
API int cmpy_dist_ctl_bytes(void) { return (int)((sizeof(DistCtl) + 255) & ~(size_t)255); }

API int cmpy_dist_create(cmpy_op_t op_main, cmpy_op_t op_t, int world, int rank, void* const* h_peer_xt,
                         void* const* h_peer_yt, void* const* h_peer_ctl, cmpy_dist_t* out) {
  return dist_create_impl(op_main, op_t, world, rank, h_peer_xt, h_peer_yt, h_peer_ctl, out);
}

API int cmpy_dist_destroy(cmpy_dist_t d) {
  delete d;
  return CMPY_OK;
}

API int cmpy_hv_apply_sharded(cmpy_dist_t d, const double* d_x_slab, double* d_y_slab, int accumulate,
                              void* stream) {
  NvtxRange nvtx_("cmpy_hv_apply_sharded");
  ARG_CHECK(d && d_x_slab && d_y_slab, "null argument");
  return d->apply(d_x_slab, d_y_slab, accumulate ? 1 : 0, false, as_stream(stream));
}

API int cmpy_dist_allreduce_sum(cmpy_dist_t d, const double* d_in2, double* d_out2, void* stream) {
  ARG_CHECK(d && d_in2 && d_out2, "null argument");
  CU_CHECK(cudaMemcpyAsync(d->d_part, d_in2, sizeof(double) * DIST_NVAL, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return d->allreduce(0, d_out2, as_stream(stream));
}

API int cmpy_dist_barrier(cmpy_dist_t d, void* stream) {
  ARG_CHECK(d, "null argument");
  return d->barrier(as_stream(stream));
}

API int cmpy_lanczos_sharded(cmpy_dist_t d, double* d_r_slab, double* d_w_slab, int maxit, double tol,
                             int check_every, double* h_alpha, double* h_beta, int* h_nit, double* h_e0,
                             void* stream) {
  NvtxRange nvtx_("cmpy_lanczos_sharded");
  return dist_lanczos_impl(d, d_r_slab, d_w_slab, maxit, tol, check_every, h_alpha, h_beta, h_nit, h_e0,
                           as_stream(stream));
}
