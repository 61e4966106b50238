/* cmpy_b200.h -- C ABI of libcmpy_b200.so (B200 / sm_100a exact-diagonalisation engine).
 *
 * The reference (dylanljones/cmpy) is pure Python and has NO FFI.  Its only seams are
 * Python protocols: scipy's LinearOperator._matvec, the
 * projector generators and the Basis/Sector containers.  Each entry point below is the
 * call a ctypes/cffi binding would make from the reference function it replaces; that
 * function is cited as `ref: file:line` (paths relative to the reference root).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed
 *    as void* (NULL = default stream).  Calls are asynchronous on that stream unless the
 *    doc says "synchronous".
 *  - pointers named d_* are DEVICE pointers borrowed from the caller (the library never
 *    frees them); pointers named h_* are HOST pointers.
 *  - every function returns 0 on success or a negative cmpy_status; the message is in
 *    cmpy_last_error() (thread local).  No exceptions or exit() cross the ABI.
 *  - a handle owns its device tables + reduction workspace, sized at create time; it is
 *    bound to the device current at creation and is not thread-safe.
 *  - index layout: idx = up_idx * num_dn + dn_idx (ref: cmpy/operators.py:33-90).
 */
#ifndef CMPY_B200_H
#define CMPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CMPY_OK = 0,
  CMPY_ERR_ARG = -1,
  CMPY_ERR_CUDA = -2,
  CMPY_ERR_NOMEM = -3,
  CMPY_ERR_UNSUPPORTED = -4,
  CMPY_ERR_NOT_CONVERGED = -5
} cmpy_status;

typedef struct cmpy_op_s* cmpy_op_t; /* opaque linear-operator handle */

const char* cmpy_last_error(void);
int cmpy_version(void);
/* number of kernels launched by this library in this process (for bench.py's
 * `gpu_launches`); cmpy_reset_launch_count() zeroes it. */
int64_t cmpy_launch_count(void);
void cmpy_reset_launch_count(void);
/* CUDA device properties needed by the host side (SM count, L2 bytes, smem/CTA). */
int cmpy_device_info(int* sm_count, int64_t* l2_bytes, int64_t* smem_per_block_optin);

/* ---- K1: sector enumeration / ranking ------------------------------------------ */
/* C(n, k) (exact, n <= 64).  ref: sizes of Basis.get_states, cmpy/basis.py:655-666 */
int cmpy_binomial(int n, int k, int64_t* h_out);
/* All `num_sites`-bit integers with popcount n, ascending, into d_states[C(num_sites,n)]
 * (combinadic unranking, one thread per index).
 * ref: Basis.generate_states, cmpy/basis.py:655-666; SpinBasis.generate_states, 748-764 */
int cmpy_sector_enumerate(int num_sites, int n, int64_t* d_states, void* stream);
/* Colex rank of each state among states of its own popcount: d_idx[i] = position of
 * d_states[i] in the ascending fixed-popcount list.
 * ref: bisect_left(states, x), cmpy/operators.py:276-299 */
int cmpy_sector_rank(const int64_t* d_states, int64_t m, int64_t* d_idx, void* stream);

/* ---- K2/K3: per-string hop tables and energies (projector building blocks) ----- */
/* For one species and one bond (site1 < site2): d_target[i] = index of
 * states[i] ^ (1<<site1 | 1<<site2) in `d_states` when the two bits differ, else -1;
 * d_sign[i] = (-1)^popcount(states[i] & bits(site1+1..site2-1) & (2^width - 1)).
 * `fixed_popcount` != 0: states are a full fixed-popcount sector (combinadic rank);
 * 0: any ascending list (binary search).
 * ref: _compute_hopping_term / _hopping_sign / bit_count, cmpy/operators.py:253-273,425-460 */
int cmpy_species_hops(const int64_t* d_states, int64_t num, int fixed_popcount, int width,
                      int site1, int site2, int32_t* d_target, int8_t* d_sign, void* stream);
/* d_out[i] = sum_{k < nvalues, bit k of states[i]} values[k], ascending k, plain adds.
 * ref: weighted_element, cmpy/operators.py:226-250 */
int cmpy_weighted_elements(const int64_t* d_states, int64_t num, const double* h_values,
                           int nvalues, double* d_out, void* stream);
/* d_out[u*num_dn + d] = weighted_element(up[u] & dn[d], u_values).
 * ref: project_hubbard_inter, cmpy/operators.py:305-356 */
int cmpy_inter_elements(const int64_t* d_up, int64_t num_up, const int64_t* d_dn, int64_t num_dn,
                        const double* h_u, int nvalues, double* d_out, void* stream);

/* ---- K4: matrix-free Hubbard / Anderson H.v ------------------------------------ */
/* Builds the operator H = sum_i eps_i n_i + sum_i u_i n_iup n_idn + sum_b hop_b (c+c + h.c.)
 * on the sector spanned by the given ascending up/dn string lists (host arrays).
 *   h_bonds[2*b], h_bonds[2*b+1] = site1 < site2 of bond b; h_hop[b] its amplitude
 *   (matrix element is +hop*sign, ref: cmpy/operators.py:454,460);
 *   sign_width = the `num_sites` argument of project_hopping (Anderson passes 0).
 *   fixed_popcount: see cmpy_species_hops.
 * ref: HubbardModel._hamiltonian_data cmpy/models/hubbard.py:13-22,75-81;
 *      SingleImpurityAndersonModel._hamiltonian_data cmpy/models/anderson.py:147-158;
 *      AbstractManyBodyModel.hamilton_operator cmpy/models/abc.py:250-256.  Synchronous. */
int cmpy_hubbard_create(int num_sites, const int64_t* h_up_states, int64_t num_up,
                        const int64_t* h_dn_states, int64_t num_dn, int fixed_popcount,
                        int nbonds, const int32_t* h_bonds, const double* h_hop,
                        const double* h_eps, const double* h_u, int sign_width,
                        cmpy_op_t* out);

/* ---- K5: matrix-free Heisenberg / XXZ H.v -------------------------------------- */
/* H on the ascending fixed-popcount list of `num_sites`-bit states with `n_up` set bits
 * (n_up < 0: all 2^num_sites states).  Directed neighbor pairs (pos1, pos2) as the
 * reference visits them (each undirected bond normally appears twice):
 * diagonal += (+-)0.25*jz per visit, off-diagonal 0.125*j per visit.
 * ref: HeisenbergModel._hamiltonian_data cmpy/models/heisenberg.py:19-40 */
int cmpy_heisenberg_create(int num_sites, int n_up, int npairs, const int32_t* h_pairs,
                           double j, double jz, cmpy_op_t* out);

/* ---- COO operator (source-compatible HamiltonOperator(size, data, indices)) ----- */
/* y[col] += val * x[row]; host COO arrays are copied to the device.
 * ref: HamiltonOperator.__init__/_matvec cmpy/operators.py:617-630 */
int cmpy_coo_create(int64_t size, int64_t nnz, const int64_t* h_rows, const int64_t* h_cols,
                    const double* h_vals, cmpy_op_t* out);

/* ---- operator calls ------------------------------------------------------------ */
int cmpy_op_destroy(cmpy_op_t op);
int cmpy_op_size(cmpy_op_t op, int64_t* h_size);
/* y = H x  (x, y: device fp64[size], distinct buffers).
 * ref: HamiltonOperator._matvec cmpy/operators.py:626-630 */
int cmpy_hv_apply(cmpy_op_t op, const double* d_x, double* d_y, void* stream);
/* Same operator applied to a contiguous slab of rows [row0, row0+nrows) of the
 * (num_up x num_dn) amplitude matrix: y_slab = (D + dn-hops)(x_slab) when
 * with_up_hops == 0.  Building block of the up-string-sharded H.v. */
int cmpy_hubbard_apply_rows(cmpy_op_t op, const double* d_x_slab, double* d_y_slab,
                            int64_t row0, int64_t nrows, int accumulate, void* stream);
/* Which H.v kernel variant cmpy_hv_apply uses: 0 = auto, 1 = global-gather,
 * 2 = shared-memory row staging with per-string tables, 3 / 4 = two-level segment kernel
 * with 512 / 1024 threads per CTA, 5-7 = class-major kernel (1024 x 8, 512 x 16, 768 x 12
 * threads x gathers in flight), 8 = long-row (more than 16 sites) class-major sub-row launches,
 * 11 = generation-3 row engine (hubbard_eng.cuh; the default of the row-slab entry point).
 * (Profiling / tests only.) */
int cmpy_hv_set_variant(cmpy_op_t op, int variant);
/* trace(H) = sum of the diagonal.  ref: HamiltonOperator._trace cmpy/operators.py:641-646.
 * Synchronous. */
int cmpy_op_trace(cmpy_op_t op, double* h_trace);
/* d_diag[i] = H[i,i] */
int cmpy_op_diagonal(cmpy_op_t op, double* d_diag, void* stream);

/* ---- K6: ladder operators between sectors -------------------------------------- */
/* y (target sector, fp64 or complex128 viewed as `ncomp` doubles per amplitude) =
 * c^dagger_{pos,sigma} x (dagger != 0) or c_{pos,sigma} x.  sigma: 1 = UP, 2 = DN
 * (ref: cmpy/basis.py:40).  signed_mode == 0 reproduces the reference (plain copy, no
 * fermionic sign); != 0 applies (-1)^(particles before (pos,sigma)), up species first.
 * Target strings: d_up_t / d_dn_t (for sigma=UP the dn lists coincide and vice versa).
 * ref: _apply_creation_up/dn, _apply_annihilation_up/dn cmpy/operators.py:652-703 */
int cmpy_ladder_apply(const int64_t* d_up, int64_t num_up, const int64_t* d_dn, int64_t num_dn,
                      const int64_t* d_up_t, int64_t num_up_t, const int64_t* d_dn_t,
                      int64_t num_dn_t, int pos, int sigma, int dagger, int signed_mode,
                      int ncomp, const double* d_x, double* d_y, void* stream);

/* ---- K7: Lanczos ---------------------------------------------------------------- */
/* Fused vector kernels exposed individually (tests / custom drivers):
 * dot = sum x[i]*y[i]; deterministic two-stage reduction; result in d_out[0]. */
int cmpy_dot(cmpy_op_t op, const double* d_x, const double* d_y, int64_t n, double* d_out,
             void* stream);
/* Runs plain (no re-orthogonalisation) Lanczos from the start vector d_v0 (not
 * modified) using two work vectors d_w0, d_w1 (fp64[size]).  Every `check_every`
 * iterations the lowest Ritz value of the tridiagonal matrix is computed on the host
 * and the run stops when |dE0| < tol and the Ritz residual estimate beta_m*|s_m| is
 * below resid_tol (resid_tol <= 0: residual^2/gap < tol), or at `maxit`
 * (CMPY_ERR_NOT_CONVERGED; outputs still valid).  Outputs: h_alpha[maxit], h_beta[maxit+1] (beta[0] = |v0|, beta[k] = norm of
 * the k-th residual), *h_nit, *h_e0, *h_resid (residual estimate of the Ritz pair).
 * If d_eigvec != NULL a second pass accumulates the normalised Ritz vector there.
 * ref: sla.eigsh(hamop, k=1, which="SA") cmpy/exactdiag.py:37 (result parity);
 *      iter_lanczos_coeffs cmpy/exactdiag.py:324-347 (recurrence).  Synchronous. */
int cmpy_lanczos_run(cmpy_op_t op, const double* d_v0, double* d_w0, double* d_w1,
                     int maxit, double tol, double resid_tol, int check_every, int use_graph,
                     double* h_alpha, double* h_beta, int* h_nit, double* h_e0,
                     double* h_resid, double* d_eigvec, void* stream);
/* Lowest eigenvalue(s) of the symmetric tridiagonal matrix (alpha[n], beta[n-1]) by
 * bisection; host-only helper.  ref: lanczos_ground_state cmpy/exactdiag.py:368-375 */
int cmpy_tridiag_lowest(const double* h_alpha, const double* h_beta, int n, int k,
                        double* h_evals, double* h_evec0);

/* ---- K8: continued fraction / pole sums ----------------------------------------- */
/* g[i] (+)= norm2 / (w - s*(a0-e0) - b1^2/(w - s*(a1-e0) - ...)), w = z[i],
 * s = +1 (particle part, poles at E_m - E0) or -1 (hole part, poles at E0 - E_n).
 * z, g: device complex128 arrays (interleaved re,im).  One thread per frequency.
 * ref (result parity): _accumulate_sum / gf_lehmann cmpy/exactdiag.py:110-129,215-245 */
int cmpy_cf_eval(const double* h_alpha, const double* h_beta, int n, double norm2, double e0,
                 int s, const double* d_z, int64_t nz, double* d_g, int accumulate, void* stream);
/* g[i] (+)= sum_k weights[k] / (z[i] - poles[k])  (device arrays).
 * ref: gf0_lehmann cmpy/greens.py:18-64; _accumulate_sum cmpy/exactdiag.py:110-129 */
int cmpy_pole_sum(const double* d_weights, const double* d_poles, int64_t npoles,
                  const double* d_z, int64_t nz, double* d_g, int accumulate, void* stream);

/* ---- K9 building blocks: slab transposes for the up-string-sharded H.v ---------- */
/* out[c*ld_out + r] (+)= in[r*ld_in + c] for r < nrows, c < ncols (tiled smem transpose).
 * Packs a (rows x cols) block of the row-major amplitude slab into the dn-major layout that
 * the all-to-all exchanges. */
int cmpy_transpose(const double* d_in, int64_t nrows, int64_t ncols, int64_t ld_in,
                   double* d_out, int64_t ld_out, int accumulate, void* stream);
/* out[r*ld_out + c] (+)= in[r*ld_in + c]: places / accumulates a received block. */
int cmpy_copy2d(const double* d_in, int64_t nrows, int64_t ncols, int64_t ld_in,
                double* d_out, int64_t ld_out, int accumulate, void* stream);

/* Caps the number of CTAs of the persistent one-CTA-per-SM row kernels behind
 * cmpy_hubbard_apply_rows (0 = no cap).  The sharded operator leaves a few SMs to the peer
 * transpose that runs concurrently with the local dn pass (cmpy_b200/dist.py). */
int cmpy_hubbard_set_grid_limit(cmpy_op_t op, int max_ctas);

/* ---- K9 over NVLink peer memory (one process per GPU, slabs in symmetric memory) ----
 * push: for the local slab of up-rows [row0, row0+nrows) (row-major nrows x num_dn) and every
 *       rank q owning the dn-columns [col_bounds[q], col_bounds[q+1]):
 *         XT_q[(c - col_bounds[q]) * ld_t + row0 + r] = x[r * num_dn + c]
 *       XT_q = h_peer_ptrs[q], a device pointer of THIS process that maps rank q's dn-major slab
 *       (remote stores over NVLink for q != rank).  ld_t = num_up.
 * pull_acc: y[r * num_dn + c] += YT_q[(c - col_bounds[q]) * ld_t + row0 + r]  (remote loads).
 * The caller orders the phases with cross-rank barriers (cmpy_b200/dist.py).  No reference
 * counterpart: cmpy has no distributed path; layout cmpy/operators.py:33-90. */
int cmpy_transpose_push(const double* d_x_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                        int64_t ld_t, int world, const int64_t* h_col_bounds,
                        void* const* h_peer_ptrs, void* stream);
/* cmpy_transpose_push with at most max_ctas persistent CTAs (0 = default grid): lets the push run
 * on the few SMs the caller keeps free of the local dn pass instead of spreading over all of them. */
int cmpy_transpose_push_capped(const double* d_x_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                               int64_t ld_t, int world, const int64_t* h_col_bounds,
                               void* const* h_peer_ptrs, int max_ctas, void* stream);
int cmpy_transpose_pull_acc(double* d_y_slab, int64_t nrows, int64_t num_dn, int64_t row0,
                            int64_t ld_t, int world, const int64_t* h_col_bounds,
                            void* const* h_peer_ptrs, void* stream);

/* ---- the sharded H.v and the sharded Lanczos recurrence as single C calls ----
 * One process per GPU (last row / 8(e)).  The caller owns three symmetric
 * (peer-mapped) allocations per rank -- the dn-major slabs XT and YT (max_q ncols_q * num_up doubles) and a
 * zero-initialised control block of cmpy_dist_ctl_bytes() bytes -- and hands over, for each of them, the
 * table of device pointers through which THIS process addresses every rank's copy (how they are mapped is
 * the caller's business: torch.distributed._symmetric_memory, cuMem* fabric handles or cudaIpc*).
 * op_main: Hubbard operator of the sector (rows = up strings); op_t: the operator with the roles of the
 * species swapped and eps = u = 0 (its dn hops are the up hops of H).  Rows / columns are partitioned as
 * [floor(n k / world), floor(n (k + 1) / world)).  Cross-rank ordering uses the library's own barrier
 * kernel over the control blocks; every rank must issue the same sequence of cmpy_dist_* calls.
 * No reference counterpart: cmpy has no distributed path; layout cmpy/operators.py:33-90, Lanczos
 * recurrence cmpy/exactdiag.py:324-347. */
typedef struct cmpy_dist_s* cmpy_dist_t;
int cmpy_dist_ctl_bytes(void);
int cmpy_dist_create(cmpy_op_t op_main, cmpy_op_t op_t, int world, int rank, void* const* h_peer_xt,
                     void* const* h_peer_yt, void* const* h_peer_ctl, cmpy_dist_t* out);
int cmpy_dist_destroy(cmpy_dist_t d);
/* y_slab = (H x)_slab (accumulate: y_slab += ...): barrier, push || local dn pass, barrier, up pass on the
 * dn-major slab, barrier, pull-accumulate. */
int cmpy_hv_apply_sharded(cmpy_dist_t d, const double* d_x_slab, double* d_y_slab, int accumulate,
                          void* stream);
/* out[0..1] = sum over ranks of in[0..1] (fixed rank order, bit-identical on every rank). */
int cmpy_dist_allreduce_sum(cmpy_dist_t d, const double* d_in2, double* d_out2, void* stream);
int cmpy_dist_barrier(cmpy_dist_t d, void* stream);
/* Two-vector sharded Lanczos from the (unnormalised) start slab d_r_slab (overwritten; d_w_slab is the
 * second slab).  h_alpha[maxit], h_beta[maxit + 1] (beta[0] = |r_0|); stops when the lowest Ritz value moves
 * by less than tol between two checks (every check_every iterations) or at a breakdown.  Returns CMPY_OK or
 * CMPY_ERR_NOT_CONVERGED; needs the row engine (uniform model, rows of at most 16 sites), otherwise
 * CMPY_ERR_UNSUPPORTED. */
int cmpy_lanczos_sharded(cmpy_dist_t d, double* d_r_slab, double* d_w_slab, int maxit, double tol,
                         int check_every, double* h_alpha, double* h_beta, int* h_nit, double* h_e0,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMPY_B200_H */
