/* hv_oracle.c -- CPU ORACLE / BASELINE, TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement ("port") of the reference's hot path for sizes where the Python
 * reference is impractical: sector enumeration, per-string hop
 * tables, and the Hubbard H.v.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs load this library; nothing under cmpy_b200/ links or calls it.
 *
 * Parity: PINNED via tests/test_oracle_c.py, which checks this file against
 * oracle/oracle_np.py and the golden fixtures produced by the unmodified reference.
 *
 * Each function cites the reference file:line it follows.  OpenMP over up-rows is the
 * only liberty taken (the reference is single-threaded Python; bench.py reports `cores`).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ascending integers with popcount n (Gosper's hack) == Basis.generate_states,
 * cmpy/basis.py:655-666 */
int64_t orc_enumerate(int num_sites, int n, int64_t* out, int64_t cap) {
  if (n < 0 || n > num_sites) return 0;
  if (n == 0) { if (cap > 0) out[0] = 0; return 1; }
  uint64_t s = (1ull << n) - 1ull, limit = 1ull << num_sites;
  int64_t k = 0;
  while (s < limit) {
    if (k < cap) out[k] = (int64_t)s;
    ++k;
    uint64_t c = s & (~s + 1ull), r = s + c;
    s = (((r ^ s) >> 2) / c) | r;
  }
  return k;
}

/* bisect_left, cmpy/operators.py:276-299 */
static int64_t bisect_left(const int64_t* a, int64_t n, int64_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* bit_count(number, width), cmpy/operators.py:253-273 */
static int bit_count_width(int64_t number, int width) {
  int c = 0;
  for (int i = 0; i < width; ++i) if (number & (1ll << i)) ++c;
  return c;
}

/* _hopping_sign, cmpy/operators.py:425-433 */
static int hopping_sign(int64_t state, int width, int s1, int s2) {
  int64_t mask = 0;
  for (int i = s1 + 1; i < s2; ++i) mask += 1ll << i;
  return (bit_count_width(state & mask, width) & 1) ? -1 : 1;
}

/* weighted_element, cmpy/operators.py:226-250 */
static double weighted_element(int64_t state, const double* values, int n) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) if (state & (1ll << i)) v += values[i];
  return v;
}

/* _compute_hopping_term for one species and one bond, cmpy/operators.py:436-460:
 * target[i] = index of the hopped string or -1, sign[i] = +-1 */
void orc_species_hops(const int64_t* states, int64_t num, int width, int s1, int s2,
                      int32_t* target, int8_t* sign) {
  const int64_t op1 = 1ll << s1, op2 = 1ll << s2;
  for (int64_t i = 0; i < num; ++i) {
    int64_t ini = states[i];
    int occ1 = (ini & op1) != 0, occ2 = (ini & op2) != 0;
    target[i] = -1;
    sign[i] = (int8_t)hopping_sign(ini, width, s1, s2);
    if (occ1 != occ2) target[i] = (int32_t)bisect_left(states, num, ini ^ op1 ^ op2);
  }
}

/* H.v on the up-rows [row0, row0+nrows): the operator of _ham_data
 * (cmpy/models/hubbard.py:13-22) / SIAM (cmpy/models/anderson.py:147-158) applied
 * matrix-free:  y[row] = sum_col H[col][row] x[col]  (HamiltonOperator._matvec,
 * cmpy/operators.py:626-630; H is symmetric so gather == scatter).
 * tgt_up/sgn_up: [nbonds][num_up] hop tables, tgt_dn/sgn_dn: [nbonds][num_dn].
 * x: full vector (num_up*num_dn); y: rows*num_dn outputs. */
void orc_hubbard_hv_rows(int num_sites, const int64_t* up, int64_t num_up, const int64_t* dn,
                         int64_t num_dn, int nbonds, const double* hop, const double* eps,
                         const double* u, const int32_t* tgt_up, const int8_t* sgn_up,
                         const int32_t* tgt_dn, const int8_t* sgn_dn, const double* x,
                         double* y, int64_t row0, int64_t nrows, int nthreads) {
  double* e_dn = (double*)malloc(sizeof(double) * (size_t)num_dn);
  for (int64_t d = 0; d < num_dn; ++d) e_dn[d] = weighted_element(dn[d], eps, num_sites);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int64_t r = 0; r < nrows; ++r) {
    const int64_t a = row0 + r;
    const double e_up = weighted_element(up[a], eps, num_sites);
    const double* xr = x + a * num_dn;
    double* yr = y + r * num_dn;
    for (int64_t d = 0; d < num_dn; ++d) {
      double acc = 0.0;
      /* project_onsite_energy: up entry, then dn entry (zero skipped) */
      if (e_up != 0.0) acc += e_up * xr[d];
      if (e_dn[d] != 0.0) acc += e_dn[d] * xr[d];
      /* project_hubbard_inter */
      double w = weighted_element(up[a] & dn[d], u, num_sites);
      if (w != 0.0) acc += w * xr[d];
      yr[d] = acc;
    }
    for (int b = 0; b < nbonds; ++b) {
      /* project_hopping: up block */
      int32_t t = tgt_up[(int64_t)b * num_up + a];
      if (t >= 0) {
        const double val = sgn_up[(int64_t)b * num_up + a] * hop[b];
        const double* xt = x + (int64_t)t * num_dn;
        for (int64_t d = 0; d < num_dn; ++d) yr[d] += val * xt[d];
      }
      /* dn block */
      const int32_t* td = tgt_dn + (int64_t)b * num_dn;
      const int8_t* sd = sgn_dn + (int64_t)b * num_dn;
      for (int64_t d = 0; d < num_dn; ++d)
        if (td[d] >= 0) yr[d] += (sd[d] * hop[b]) * xr[td[d]];
    }
  }
  free(e_dn);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- Heisenberg (spin) sector ------------------------------------------------------------------ */
/* binomials for the combinadic rank: position of a state in the ascending list of N-bit integers of
 * fixed popcount == states.index(s2) in HeisenbergModel._hamiltonian_data
 * (cmpy/models/heisenberg.py:38, states from SpinBasis.get_states, cmpy/basis.py:748-774) */
static int64_t g_binom[65][65];
static void binom_init(void) {
  static int done = 0;
  if (done) return;
  for (int n = 0; n < 65; ++n)
    for (int k = 0; k < 65; ++k)
      g_binom[n][k] = (k == 0) ? 1 : (n == 0 ? 0 : g_binom[n - 1][k - 1] + g_binom[n - 1][k]);
  done = 1;
}
static int64_t colex_rank64(uint64_t s) {
  int64_t r = 0;
  int k = 0;
  while (s) {
    int p = __builtin_ctzll(s);
    s &= s - 1;
    r += g_binom[p][++k];
  }
  return r;
}
static uint64_t colex_unrank64(int64_t idx, int n, int num_sites) {
  uint64_t s = 0;
  int p = num_sites;
  for (int k = n; k >= 1; --k) {
    do { --p; } while (g_binom[p][k] > idx);
    s |= 1ull << p;
    idx -= g_binom[p][k];
  }
  return s;
}

/* y[r] = (H x)[i0 + r], r < count, for the sector of popcount n_up of an N-site Heisenberg model:
 * HeisenbergModel._hamiltonian_data (cmpy/models/heisenberg.py:19-40) applied matrix-free, the loops
 * pos1 in range(N), pos2 in neighbors(pos1) (directed pairs, nbr_ptr / nbr_idx) in the reference's
 * order; H is symmetric, so the gather below equals HamiltonOperator._matvec (cmpy/operators.py:626-630). */
void orc_heisenberg_hv_range(int num_sites, int n_up, const int32_t* nbr_ptr, const int32_t* nbr_idx,
                             double j, double jz, const double* x, double* y, int64_t i0, int64_t count,
                             int nthreads) {
  binom_init();
  const double factor = 0.25;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static, 4096)
#endif
  for (int64_t r = 0; r < count; ++r) {
    const int64_t i = i0 + r;
    const uint64_t s1 = colex_unrank64(i, n_up, num_sites);
    const double xi = x[i];
    double acc = 0.0;
    for (int pos1 = 0; pos1 < num_sites; ++pos1)
      for (int32_t q = nbr_ptr[pos1]; q < nbr_ptr[pos1 + 1]; ++q) {
        const int pos2 = nbr_idx[q];
        const int b1 = (int)((s1 >> pos1) & 1ull), b2 = (int)((s1 >> pos2) & 1ull);
        const double sign = (b1 == b2) ? 1.0 : -1.0;
        acc += (sign * factor * jz) * xi;
        if (b1 != b2) acc += (factor * j / 2) * x[colex_rank64(s1 ^ (1ull << pos1) ^ (1ull << pos2))];
      }
    y[r] = acc;
  }
}

int64_t orc_binomial(int n, int k) {
  binom_init();
  return (n < 0 || n > 64 || k < 0 || k > 64) ? 0 : g_binom[n][k];
}
