// dist.cuh -- K9 / K7 on several GPUs as C entry points: the up-string-sharded H.v and the sharded
// two-vector Lanczos recurrence (one process per GPU; slabs and a small control block in symmetric,
// peer-mapped memory).  The reference has no distributed path; the layout being
// sharded is cmpy/operators.py:33-90, the recurrence cmpy/exactdiag.py:324-347.
//
//   cmpy_hv_apply_sharded : barrier | push X -> XT_q (side stream) || local dn pass | barrier |
//                           up pass on the dn-major slab | barrier | pull-accumulate YT_q -> y
//   cmpy_lanczos_sharded  : per iteration the same H.v with the recurrence folded in
//                             w = (1/b_j) H r_j - (b_j/b_{j-1}) r_{j-1}     (dn pass and pull take the device scalars)
//                             a_j = <r_j, w> / b_j                           (local dot + all-reduce)
//                             r_{j+1} = w - (a_j/b_j) r_j,  b_{j+1} = |r_{j+1}|   (local update + all-reduce)
//                           i.e. the unnormalised two-vector form of lanczos.cuh; alpha / beta / j live on the
//                           device, the host looks at them every `check_every` iterations.
// Cross-rank synchronisation is done by this library's own kernels over the control blocks: a barrier is
// one 32-thread kernel (thread q: st.release.sys of the epoch into rank q's flag[my rank], then
// ld.acquire.sys spin on my flag[q]); the scalar all-reduce writes this rank's partials into every peer's
// slot, runs the same handshake and sums the slots in rank order (deterministic, identical on all ranks).
#pragma once
#include "common.cuh"
#include "peer.cuh"
#include "lanczos.cuh"
#include "hubbard_op.cuh"

#define DIST_NVAL 2
struct DistCtl {                           // one per rank, in symmetric memory, zero-initialised by the caller
  unsigned long long flag[PEER_MAX];       // flag[q]: last epoch rank q has signalled to this rank
  double slot[2][PEER_MAX][DIST_NVAL];     // slot[parity][q]: partials of rank q (double-buffered by epoch parity)
};

struct CtlTable {
  DistCtl* ctl[PEER_MAX];
  int world, rank;
};

__device__ __forceinline__ void dist_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long dist_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// thread q < world: tell rank q that this rank reached `ep`, wait until rank q has told us the same.
// The wait gives up after ~20 s of SM clocks: a peer that died must not hang this GPU for good.  After a
// give-up the epoch counter stops advancing and later results are meaningless (the process is expected to
// fail on its dead peer anyway; the parity checks of bench.py / the tests would flag a silent case).
#define DIST_SPIN_LIMIT 40000000000ll
__device__ __forceinline__ bool dist_handshake(const CtlTable& ct, unsigned long long ep, int q) {
  dist_st_release(&ct.ctl[q]->flag[ct.rank], ep);
  const unsigned long long* mine = &ct.ctl[ct.rank]->flag[q];
  const long long t0 = clock64();
  unsigned spins = 0;
  while (dist_ld_acquire(mine) < ep) {
    if ((++spins & 1023u) == 0u && clock64() - t0 > DIST_SPIN_LIMIT) return false;
  }
  return true;
}

// epoch: device counter of this rank (every rank runs the same sequence of barrier / all-reduce kernels)
__global__ void __launch_bounds__(32) dist_barrier_kernel(CtlTable ct, unsigned long long* epoch) {
  const unsigned long long ep = *epoch + 1;
  __syncwarp();
  bool ok = true;
  if ((int)threadIdx.x < ct.world) ok = dist_handshake(ct, ep, threadIdx.x);
  ok = __all_sync(0xffffffffu, ok);
  if (threadIdx.x == 0 && ok) *epoch = ep;
}

// mode 0: out[v] = sum_q partial_q[v]
// mode 1: alpha[j] = sum / beta[j]                        (Lanczos: a_j = <r_j, w> / b_j)
// mode 2: beta[j+1] = sqrt(sum); coef = {1/b_{j+1}, -b_{j+1}/b_j}; iter = j + 1
__global__ void __launch_bounds__(32) dist_allreduce_kernel(CtlTable ct, unsigned long long* epoch,
                                                            const double* __restrict__ partial, int mode,
                                                            double* out, double* alpha, double* beta, int* iter,
                                                            double* coef) {
  const unsigned long long ep = *epoch + 1;
  const int par = (int)(ep & 1ull), q = threadIdx.x;
  __syncwarp();
  if (q < ct.world) {
#pragma unroll
    for (int v = 0; v < DIST_NVAL; ++v) ct.ctl[q]->slot[par][ct.rank][v] = partial[v];
    __threadfence_system();
    if (!dist_handshake(ct, ep, q)) partial = nullptr;   // timed out: flagged below
  }
  const bool ok = __all_sync(0xffffffffu, partial != nullptr);
  if (threadIdx.x == 0 && ok) {
    double s[DIST_NVAL];
#pragma unroll
    for (int v = 0; v < DIST_NVAL; ++v) s[v] = 0.0;
    const volatile DistCtl* me = ct.ctl[ct.rank];
    for (int r = 0; r < ct.world; ++r)
#pragma unroll
      for (int v = 0; v < DIST_NVAL; ++v) s[v] += me->slot[par][r][v];
    if (mode == 0) {
#pragma unroll
      for (int v = 0; v < DIST_NVAL; ++v) out[v] = s[v];
    } else if (mode == 1) {
      const int j = *iter;
      alpha[j] = s[0] / beta[j];
    } else {
      const int j = *iter;
      const double b1 = sqrt(s[0]), b0 = beta[j];
      beta[j + 1] = b1;
      coef[0] = 1.0 / b1;
      coef[1] = -b1 / b0;
      *iter = j + 1;
    }
    *epoch = ep;
  }
}

// local partial of <x, y> (mode 0) -- fixed-order two-stage sum, result in part[0] (part[1] = 0)
__global__ void __launch_bounds__(256) dist_dot_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                      i64 n, double* partials, unsigned* ticket, double* part) {
  __shared__ double red[32];
  double acc = 0.0;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i] * y[i];
  double b = block_sum(acc, red);
  double total;
  if (grid_sum_last(b, partials, ticket, red, &total)) { part[0] = total; part[1] = 0.0; }
}

// r_{j+1} = w - (alpha_j / beta_j) r_j in place of w; part[0] = local |r_{j+1}|^2
__global__ void __launch_bounds__(256) dist_update_kernel(const double* __restrict__ X, double* __restrict__ W, i64 n,
                                                         const double* alpha, const double* beta, const int* iter,
                                                         double* partials, unsigned* ticket, double* part) {
  __shared__ double red[32];
  const int j = *iter;
  const double c = alpha[j] / beta[j];
  double acc = 0.0;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double w = W[i] - c * X[i];
    W[i] = w;
    acc += w * w;
  }
  double b = block_sum(acc, red);
  double total;
  if (grid_sum_last(b, partials, ticket, red, &total)) { part[0] = total; part[1] = 0.0; }
}

// beta[0] = sqrt(sum); coef = {1/beta0, 0}; iter = 0   (start of the recurrence; after an all-reduce in mode 0)
__global__ void dist_start_kernel(const double* sum, double* beta, int* iter, double* coef) {
  const double b0 = sqrt(sum[0]);
  beta[0] = b0;
  coef[0] = 1.0 / b0;
  coef[1] = 0.0;
  *iter = 0;
}

struct cmpy_dist_s {
  HubbardOp* op_main = nullptr;   // rows = up strings (diagonal + dn hops)
  HubbardOp* op_t = nullptr;      // rows = dn strings, zero eps / u: its "dn hops" are the up hops of H
  int world = 1, rank = 0;
  i64 num_up = 0, num_dn = 0;
  i64 rb[PEER_MAX + 1], cb[PEER_MAX + 1];
  PeerTable xt, yt;
  CtlTable ct;
  unsigned long long* d_epoch = nullptr;
  double* d_part = nullptr;       // [DIST_NVAL] local partials
  double* d_coef = nullptr;       // {c1, c2} of the scaled accumulation
  double* d_sum = nullptr;        // [DIST_NVAL]
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int push_sms = 32;
  ~cmpy_dist_s() {
    cudaFree(d_epoch); cudaFree(d_part); cudaFree(d_coef); cudaFree(d_sum);
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
  i64 nrows() const { return rb[rank + 1] - rb[rank]; }
  i64 ncols() const { return cb[rank + 1] - cb[rank]; }
  double* my_xt() const { return xt.base[rank]; }
  double* my_yt() const { return yt.base[rank]; }

  int barrier(cudaStream_t st) {
    dist_barrier_kernel<<<1, 32, 0, st>>>(ct, d_epoch);
    KERNEL_CHECK();
    return CMPY_OK;
  }

  // y = (H x)_slab, y += (H x)_slab, or (scale) y = c1 (H x)_slab + c2 y with {c1, c2} = d_coef
  int apply(const double* x, double* y, int accumulate, bool scaled, cudaStream_t st) {
    const i64 r0 = rb[rank], nr = nrows(), c0 = cb[rank], nc = ncols();
    int rc = barrier(st);   // every rank is done with the XT / YT slabs of the previous call
    if (rc) return rc;
    // The local dn pass (one CTA and all shared memory per SM) is enqueued FIRST, on push_sms fewer SMs;
    // the push of the transposed tiles into the owners' XT slabs follows on the side stream with a grid
    // capped to what fits the SMs left free (6 CTAs each).  Measured on 2 x B200, 4x4 sector: 2.86 ms per
    // H.v against 2.96 ms with the push enqueued first (its 8 persistent CTAs per SM then occupy every SM
    // before the dn pass arrives).
    nvtxRangePushA("sharded H.v: dn pass || push");
    struct PopAll { int n = 1; ~PopAll() { while (n-- > 0) nvtxRangePop(); } } nvtx_pop;   // early returns pop too
    const bool reserve = world > 1 && push_sms > 0 && push_sms < op_main->sm_count;
    LzCtx nolz; nolz.enabled = 0; nolz.iter = nullptr; nolz.beta = nullptr; nolz.alpha = nullptr;
    nolz.partials = nullptr; nolz.ticket = nullptr;
    CU_CHECK(cudaEventRecord(ev_fork, st));    // fork point: after the barrier, before the dn pass
    const int saved = op_main->grid_limit;
    if (reserve) op_main->grid_limit = op_main->sm_count - push_sms;
    rc = nr > 0 ? op_main->apply_slab(x, y, r0, nr, 0, accumulate, nolz, st, scaled ? d_coef : nullptr) : CMPY_OK;
    op_main->grid_limit = saved;
    if (rc) return rc;
    CU_CHECK(cudaStreamWaitEvent(side, ev_fork, 0));
    if (nr > 0 && num_dn > 0) {
      const i64 ntiles = ((nr + 127) / 128) * ((num_dn + 31) / 32);
      int g = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
      if (reserve && g > 6 * push_sms) g = 6 * push_sms;
      peer_transpose_kernel<false, 128><<<g, 256, 0, side>>>(const_cast<double*>(x), nr, num_dn, r0, num_up, xt);
      KERNEL_CHECK();
    }
    CU_CHECK(cudaEventRecord(ev_join, side));
    CU_CHECK(cudaStreamWaitEvent(st, ev_join, 0));
    rc = barrier(st);       // all pushes have landed
    if (rc) return rc;
    nvtxRangePop(); nvtxRangePushA("sharded H.v: up pass");
    if (nc > 0) {
      rc = op_t->apply_slab(my_xt(), my_yt(), c0, nc, 0, 0, nolz, st);   // up hops, row-local in the dn-major slab
      if (rc) return rc;
    }
    rc = barrier(st);       // every YT slab is complete
    if (rc) return rc;
    nvtxRangePop(); nvtxRangePushA("sharded H.v: pull-accumulate");
    if (nr > 0 && num_dn > 0) {
      const double* sc = scaled ? d_coef : nullptr;
      // 64-row tiles (remote runs of 512 bytes): measured on 2 x B200, 4x4 sector, 32 rows 2.72 ms per H.v,
      // 64: 2.61, 128: 2.65 (profiles/r2_dist_check_n2_pull*.log)
      const i64 ntiles = ((nr + 63) / 64) * ((num_dn + 31) / 32);
      const int g = (int)(ntiles < 148 * 8 ? ntiles : 148 * 8);
      peer_transpose_kernel<true, 64><<<g, 256, 0, st>>>(y, nr, num_dn, r0, num_up, yt, sc);
      KERNEL_CHECK();
    }
    return CMPY_OK;
  }

  int allreduce(int mode, double* out, cudaStream_t st) {
    dist_allreduce_kernel<<<1, 32, 0, st>>>(ct, d_epoch, d_part, mode, out, op_main->d_alpha, op_main->d_beta,
                                            op_main->d_iter, d_coef);
    KERNEL_CHECK();
    return CMPY_OK;
  }
};

static int dist_create_impl(cmpy_op_s* op_main, cmpy_op_s* op_t, int world, int rank, void* const* h_peer_xt,
                            void* const* h_peer_yt, void* const* h_peer_ctl, cmpy_dist_s** out) {
  ARG_CHECK(out, "null output handle");
  *out = nullptr;
  HubbardOp* a = dynamic_cast<HubbardOp*>(op_main);
  HubbardOp* b = dynamic_cast<HubbardOp*>(op_t);
  ARG_CHECK(a && b, "dist_create: Hubbard operators expected");
  ARG_CHECK(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "dist_create: bad world / rank");
  ARG_CHECK(h_peer_xt && h_peer_yt && h_peer_ctl, "dist_create: null pointer tables");
  ARG_CHECK(a->up.num == b->dn.num && a->dn.num == b->up.num, "dist_create: op_t must be the transposed-role operator");
  cmpy_dist_s* d = new cmpy_dist_s();
  d->op_main = a; d->op_t = b; d->world = world; d->rank = rank;
  d->num_up = a->up.num; d->num_dn = a->dn.num;
  for (int k = 0; k <= world; ++k) {   // the balanced contiguous partition of cmpy_b200/dist.py::ShardPlan
    d->rb[k] = (d->num_up * k) / world;
    d->cb[k] = (d->num_dn * k) / world;
  }
  d->xt.world = d->yt.world = world;
  d->ct.world = world; d->ct.rank = rank;
  for (int q = 0; q < world; ++q) {
    if (!h_peer_xt[q] || !h_peer_yt[q] || !h_peer_ctl[q]) { delete d; return cmpy_fail(CMPY_ERR_ARG, "dist_create: null peer pointer"); }
    d->xt.base[q] = (double*)h_peer_xt[q]; d->yt.base[q] = (double*)h_peer_yt[q];
    d->ct.ctl[q] = (DistCtl*)h_peer_ctl[q];
  }
  for (int q = 0; q <= world; ++q) { d->xt.cb[q] = d->cb[q]; d->yt.cb[q] = d->cb[q]; }
  cudaError_t e = cudaMalloc(&d->d_epoch, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(d->d_epoch, 0, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&d->d_part, sizeof(double) * DIST_NVAL);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_coef, sizeof(double) * 2);
  if (e == cudaSuccess) e = cudaMalloc(&d->d_sum, sizeof(double) * DIST_NVAL);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete d; return cmpy_fail(CMPY_ERR_CUDA, std::string("dist_create: ") + cudaGetErrorString(e)); }
  *out = d;
  return CMPY_OK;
}

// Two-vector sharded Lanczos.  d_r: slab of the (unnormalised) start vector, overwritten; d_w: second slab.
static int dist_lanczos_impl(cmpy_dist_s* d, double* d_r, double* d_w, int maxit, double tol, int check_every,
                             double* h_alpha, double* h_beta, int* h_nit, double* h_e0, cudaStream_t st) {
  ARG_CHECK(d && d_r && d_w && h_alpha && h_beta && h_nit && h_e0 && maxit >= 1, "bad argument");
  HubbardOp* op = d->op_main;
  if (check_every < 1) check_every = 1;
  int rc = op->ensure_lz_capacity(maxit + check_every + 2);
  if (rc) return rc;
  const i64 n = d->nrows() * d->num_dn;
  const int g = grid_for(n > 0 ? n : 1, 256, op->sm_count * 8);
  CU_CHECK(cudaMemsetAsync(op->d_ticket, 0, sizeof(unsigned) * 4, st));
  // b_0 = |r_0| (all ranks), w = 0
  dist_dot_kernel<<<g, 256, 0, st>>>(d_r, d_r, n, op->d_partials, op->d_ticket, d->d_part);
  KERNEL_CHECK();
  rc = d->allreduce(0, d->d_sum, st);
  if (rc) return rc;
  dist_start_kernel<<<1, 1, 0, st>>>(d->d_sum, op->d_beta, op->d_iter, d->d_coef);
  KERNEL_CHECK();
  CU_CHECK(cudaMemsetAsync(d_w, 0, sizeof(double) * (size_t)(n > 0 ? n : 1), st));
  std::vector<double> alpha(maxit + check_every + 2), beta(maxit + check_every + 3);
  double* X = d_r; double* W = d_w;
  double e0 = 0.0, e0_prev = 0.0;
  int done = 0, m_final = 0;
  bool converged = false;
  while (done < maxit && !converged) {
    const int todo = std::min(check_every, maxit - done);
    for (int k = 0; k < todo; ++k) {
      rc = d->apply(X, W, 0, true, st);   // W = (1/b_j) H X - (b_j/b_{j-1}) W
      if (rc) return rc;
      dist_dot_kernel<<<g, 256, 0, st>>>(X, W, n, op->d_partials, op->d_ticket, d->d_part);
      KERNEL_CHECK();
      rc = d->allreduce(1, nullptr, st);  // alpha_j
      if (rc) return rc;
      dist_update_kernel<<<g, 256, 0, st>>>(X, W, n, op->d_alpha, op->d_beta, op->d_iter, op->d_partials,
                                            op->d_ticket, d->d_part);
      KERNEL_CHECK();
      rc = d->allreduce(2, nullptr, st);  // beta_{j+1}, coefficients of the next step, ++j
      if (rc) return rc;
      std::swap(X, W);
    }
    done += todo;
    CU_CHECK(cudaMemcpyAsync(alpha.data(), op->d_alpha, sizeof(double) * done, cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaMemcpyAsync(beta.data(), op->d_beta, sizeof(double) * (done + 1), cudaMemcpyDeviceToHost, st));
    CU_CHECK(cudaStreamSynchronize(st));
    int m = done;
    double scale = 0.0;
    for (int i = 0; i < m; ++i) {   // usable length: stop at a breakdown (invariant subspace) or non-finite value
      if (!std::isfinite(alpha[i]) || !std::isfinite(beta[i + 1])) { m = i; converged = true; break; }
      scale = std::max(scale, fabs(alpha[i]) + fabs(beta[i + 1]));
      if (beta[i + 1] <= 1e-13 * std::max(scale, 1.0)) { m = i + 1; converged = true; break; }
    }
    if (m < 1) m = 1;
    e0 = tridiag_kth(alpha.data(), beta.data() + 1, m, 0);
    m_final = m;
    if (done > check_every && fabs(e0 - e0_prev) < tol) converged = true;
    e0_prev = e0;
  }
  for (int i = 0; i < m_final; ++i) h_alpha[i] = alpha[i];
  for (int i = 0; i <= m_final; ++i) h_beta[i] = beta[i];
  *h_nit = m_final;
  *h_e0 = e0;
  return converged ? CMPY_OK : CMPY_ERR_NOT_CONVERGED;
}
