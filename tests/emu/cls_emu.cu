// cls_emu.cu -- TEST INFRASTRUCTURE (not part of the product): runs the shared-memory phases of
// the class-major H.v kernel (cmpy_b200/csrc/hubbard_cls.cuh) lane by lane on the CPU.
//
// The phase bodies (cls_phase_a/b) are __host__ __device__;
// inside one phase every (warp, lane) only reads what the previous phase wrote and writes its
// own slots, so executing the lanes one after the other is exactly what the barrier-separated
// kernel computes.  Staging and un-staging follow the kernel's column-pair -> slot map
// (pair_seg + seg_delta).  tests/test_cls_emulation.py compares the result with a direct
// evaluation of (D + T_dn) x on one row (ref: cmpy/operators.py:305-527).
//
// Built by the test itself:  nvcc -O1 -std=c++17 -shared -Xcompiler -fPIC cls_emu.cu
#include "../../cmpy_b200/csrc/hubbard_cls.cuh"
#include "../../cmpy_b200/csrc/peer.cuh"
#include <cmath>

static int g_emu_shift = 0;
extern "C" void emu_set_shift(int shift) { g_emu_shift = shift ? 1 : 0; }

template <bool SPIN>
static void run_phases(const ClsHost& H, const SpinDiag& sd, const std::vector<double>& xs,
                       std::vector<double>& ys, uint32_t ups, double eu, double u0, double hop0,
                       int nwarps) {
  const ClsLayout& L = H.lay;
  const unsigned char* tab = H.blob.data();
  // engine 0: the item headers of hub_cls_kernel, verbatim
  const uint16_t* item_a = reinterpret_cast<const uint16_t*>(tab + L.off_item_a);
  const uint16_t* item_b = reinterpret_cast<const uint16_t*>(tab + L.off_item_b);
  const uint8_t* k_of_q = tab + L.off_k_of_q;
  const uint16_t* dl_of_q = reinterpret_cast<const uint16_t*>(tab + L.off_dl_of_q);
  const uint32_t* ll_ptr = reinterpret_cast<const uint32_t*>(tab + L.off_ll_ptr);
  const uint8_t* ll_ent = tab + L.off_ll_ent;
  const uint16_t* dh_list = reinterpret_cast<const uint16_t*>(tab + L.off_dh_list);
  const uint16_t* hi_sbase = reinterpret_cast<const uint16_t*>(tab + L.off_hi_sbase);
  const uint8_t* hi_k = tab + L.off_hi_k;
  const uint32_t* hh_ptr = reinterpret_cast<const uint32_t*>(tab + L.off_hh_ptr);
  const uint16_t* hh_ent = reinterpret_cast<const uint16_t*>(tab + L.off_hh_ent);
  const uint16_t* lh_hi = reinterpret_cast<const uint16_t*>(tab + L.off_lh_hi);
  const uint8_t* lh_lo = tab + L.off_lh_lo;
  for (int it = 0; it < L.na; ++it)
    for (int lane = 0; lane < 32; ++lane) {
      const int q = item_a[it];
      const int k = k_of_q[q];
      const uint32_t pp = ll_ptr[q];
      const uint32_t dlbits = dl_of_q[q];
      const int hk = L.H[k], r = q - L.qoff[k];
      if (hk <= 32) cls_phase_a<1, SPIN>(L, sd, xs.data(), ys.data(), ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
      else if (hk <= 64) cls_phase_a<2, SPIN>(L, sd, xs.data(), ys.data(), ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
      else cls_phase_a<3, SPIN>(L, sd, xs.data(), ys.data(), ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
    }
  for (int it = 0; it < L.nb; ++it)
    for (int lane = 0; lane < 32; ++lane) {
      const int dh = item_b[it];
      const int k = hi_k[dh];
      const uint32_t pp = hh_ptr[dh];
      const int sb = hi_sbase[dh], sk = L.S[k];
      if (sk <= 32) cls_phase_b<1>(L, xs.data(), ys.data(), hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
      else if (sk <= 64) cls_phase_b<2>(L, xs.data(), ys.data(), hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
      else cls_phase_b<3>(L, xs.data(), ys.data(), hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
    }
}

// One row of y = (D + T_dn) x through the emulated phases.
//   spin = 0: Hubbard flavour, diag = eu + u0 * popc(ups & dn)
//   spin = 1: XXZ flavour, diag = sd_e0 + sd_escale * #antiparallel bonds (bonds grouped by site
//             distance as in heisenberg.cuh), hops without sign (pass sign_width = 0)
// returns 0 ok, 1 sector not supported by the table builder, 2 bad argument, 3 a slot was not written
extern "C" int emu_cls_row(int num_sites, int n_dn, int nbonds, const int* s1, const int* s2,
                           int sign_width, double eps0, double u0, double hop0, unsigned ups, double eu,
                           int eng, int spin, double sd_e0, double sd_escale, int nwarps,
                           const double* x_row, double* y_row, int* info) {
  const u64* B = host_binom();
  if (num_sites < 2 || num_sites > 16 || n_dn < 0 || n_dn > num_sites || nwarps < 1) return 2;
  const i64 num_dn = (i64)B[num_sites * BINOM_N + n_dn];
  ClsHost H;
  const double eps[1] = {eps0};
  if (eng != 0) return 2;   // (the round-1 chunked-task engine was removed)
  if (build_cls_host(H, num_sites, n_dn, num_dn, nbonds, s1, s2, sign_width, eps, 232448)) return 2;
  if (!H.ok) return 1;
  const ClsLayout& L = H.lay;
  SpinDiag sd;
  memset(&sd, 0, sizeof(sd));
  if (spin) {
    for (int k = 0; k < nbonds; ++k) {
      const int delta = s2[k] - s1[k];
      int i = 0;
      for (; i < sd.ndelta; ++i) if (sd.delta[i] == delta) break;
      if (i == sd.ndelta) {
        if (sd.ndelta == 4) return 1;
        sd.delta[sd.ndelta++] = delta;
      }
      sd.dmask[i] |= 1u << s1[k];
    }
    sd.e0 = sd_e0; sd.escale = sd_escale;
  }
  const int xs_total = (L.xs_elems + CLS_ZREG + 1) & ~1;
  std::vector<double> xs(xs_total, 0.0), ys(L.xs_elems, std::nan(""));
  const int16_t* seg_delta = reinterpret_cast<const int16_t*>(H.blob.data() + L.off_seg_delta);
  // the column-pair -> slot map of hub_cls_kernel: shift = 0 pairs (2i, 2i+1); shift = 1 (sub-rows of
  // long rows that start at an odd element of the vector) pairs (2i-1, 2i) with one-column pairs at
  // both ends
  const int shift = g_emu_shift, ndi = (int)num_dn;
  const int npairs = (ndi + shift + 1) / 2;
  const std::vector<uint16_t>& pseg = shift ? H.pair_seg1 : H.pair_seg;
  if ((int)pseg.size() < npairs) return 3;
  std::vector<int> slot(num_dn, -1);
  for (int pi = 0; pi < npairs; ++pi) {
    const uint32_t ps = pseg[pi];
    const int si = (int)(ps & 0x7fffu), d = 2 * pi - shift;
    int slot0 = d + seg_delta[si];
    int slot1 = d + 1 + seg_delta[si + (int)(ps >> 15)];
    if (d < 0) slot0 = slot1;
    if (d + 1 >= ndi) slot1 = slot0;
    if (d >= 0) slot[d] = slot0;
    if (d + 1 < ndi) slot[d + 1] = slot1;
  }
  for (i64 d = 0; d < num_dn; ++d) {
    if (slot[d] < 0 || slot[d] >= L.xs_elems) return 3;
    xs[slot[d]] = x_row[d];
  }
  if (spin) run_phases<true>(H, sd, xs, ys, ups, eu, u0, hop0, nwarps);
  else run_phases<false>(H, sd, xs, ys, ups, eu, u0, hop0, nwarps);
  for (i64 d = 0; d < num_dn; ++d) {
    y_row[d] = ys[slot[d]];
    if (std::isnan(y_row[d])) return 3;
  }
  if (info) {
    info[0] = L.bytes; info[1] = (int)H.smem; info[2] = L.na;
    info[3] = L.nb; info[4] = L.nlh; info[5] = L.xs_elems;
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// Long rows (more than 16 sites): one dn row through the sub-row launches of hub_cls_kernel<..., LONG>.
// Tables come from build_long_tables in host-mirror mode (the builder the library uses); staging and
// the long-row part of phase C (top-bond sub-row gathers, straddling-bond index maps) restate the
// kernel lines.  `covered` marks the columns whose sub-row class the kernel takes (odd-length classes
// are left to another kernel and stay 0).
static int g_long_spin = 0;           // 1: XXZ flavour (hub_cls_kernel<..., LONG, SPIN> as launched by heisenberg.cuh)
static double g_long_sd[2] = {0, 0};  // e0, escale of the spin diagonal
extern "C" void emu_long_set_spin(int spin, double e0, double escale) {
  g_long_spin = spin; g_long_sd[0] = e0; g_long_sd[1] = escale;
}

extern "C" int emu_long_row(int num_sites, int n_dn, int nbonds, const int* s1, const int* s2,
                            int sign_width, double u0, double hop0, unsigned ups, double eu, int eng,
                            const double* x_row, double* y_row, unsigned char* covered) {
  const u64* B = host_binom();
  if (num_sites <= LONG_RBITS || num_sites > 32 || n_dn < 0 || n_dn > num_sites) return 2;
  const i64 num_dn = (i64)B[num_sites * BINOM_N + n_dn];
  LongTables dummy;
  LongHost LH;
  std::vector<int> skipped;
  const double eps[1] = {0.0};
  if (build_long_tables(dummy, num_sites, n_dn, num_dn, nbonds, s1, s2, sign_width, eps, 232448, &skipped, &LH))
    return 2;
  if (LH.sets.empty()) return 1;
  SpinDiag sd;
  memset(&sd, 0, sizeof(sd));
  if (g_long_spin) {   // bonds grouped by site distance, as in HeisenbergOp::configure_fast
    for (int k = 0; k < nbonds; ++k) {
      const int delta = s2[k] - s1[k];
      int i = 0;
      for (; i < sd.ndelta; ++i) if (sd.delta[i] == delta) break;
      if (i == sd.ndelta) {
        if (sd.ndelta == 4) return 1;
        sd.delta[sd.ndelta++] = delta;
      }
      sd.dmask[i] |= 1u << s1[k];
    }
    sd.e0 = g_long_sd[0]; sd.escale = g_long_sd[1];
  }
  for (i64 d = 0; d < num_dn; ++d) { y_row[d] = 0.0; covered[d] = 0; }
  for (const LongSetHost& S : LH.sets) {
    const ClsLayout& L = S.cls.lay;
    const int ndi = S.row_len, shift = S.shift;
    const int xs_total = (L.xs_elems + CLS_ZREG + 1) & ~1;
    const int16_t* seg_delta = reinterpret_cast<const int16_t*>(S.cls.blob.data() + L.off_seg_delta);
    const std::vector<uint16_t>& pseg = shift ? S.cls.pair_seg1 : S.cls.pair_seg;
    const int npairs = (ndi + shift + 1) / 2;
    if ((int)pseg.size() < npairs) return 3;
    std::vector<int> slot(ndi, -1);
    for (int pi = 0; pi < npairs; ++pi) {
      const uint32_t ps = pseg[pi];
      const int si = (int)(ps & 0x7fffu), d = 2 * pi - shift;
      int slot0 = d + seg_delta[si];
      int slot1 = d + 1 + seg_delta[si + (int)(ps >> 15)];
      if (d < 0) slot0 = slot1;
      if (d + 1 >= ndi) slot1 = slot0;
      if (d >= 0) slot[d] = slot0;
      if (d + 1 < ndi) slot[d + 1] = slot1;
    }
    for (int ti = 0; ti < S.ntop; ++ti) {
      const uint32_t dtop = S.top_val[ti];
      const i64 base = S.sub_off[ti];
      if ((int)(base & 1) != shift) return 4;   // the 16-byte alignment rule of the launch
      const double* xr = x_row + base;
      std::vector<double> xs(xs_total, 0.0), ys(L.xs_elems, std::nan(""));
      for (int d = 0; d < ndi; ++d) {
        if (slot[d] < 0 || slot[d] >= L.xs_elems) return 3;
        xs[slot[d]] = xr[d];
      }
      const double eu_sub = eu + LH.e_dn_const + u0 * (double)__builtin_popcount((ups >> 16) & dtop);
      if (g_long_spin) run_phases<true>(S.cls, sd, xs, ys, dtop << 16, 0.0, 0.0, hop0, 32);
      else run_phases<false>(S.cls, sd, xs, ys, ups, eu_sub, u0, hop0, 32);
      for (int d = 0; d < ndi; ++d) {
        double a = ys[slot[d]];
        if (std::isnan(a)) return 3;
        for (int q = S.tb_ptr[ti]; q < S.tb_ptr[ti + 1]; ++q)   // hops inside dtop: whole sub-row gathers
          a += (S.tb_ent[q].y ? -hop0 : hop0) * xr[d + S.tb_ent[q].x];
        for (int sb = 0; sb < LH.nsb; ++sb) {                   // bonds straddling site 15 | 16
          const int2 src = S.sb_src[(size_t)sb * S.ntop + ti];
          if (src.y & 1) {
            const uint32_t e0 = S.sb_map[((size_t)(2 * sb + ((src.y >> 1) & 1))) * ndi + d];
            const uint32_t ptop = (uint32_t)(src.y >> 2) & 1u;
            if (e0 >> 31) {
              const double v = hop0 * xr[src.x + (i64)(e0 & 0x3fffffffu)];
              a += (((e0 >> 30) & 1u) ^ ptop) ? -v : v;
            }
          }
        }
        y_row[base + d] = a;
        covered[base + d] = 1;
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// Segment kernel (hub_seg_kernel, the default full H.v): its per-amplitude device function
// seg_dn_part<UNI> is __host__ __device__; here it runs for every column of one staged row against
// the tables of build_seg_tables (host-mirror mode).  uniform = 1: one hop amplitude / U / eps
// (template UNI); 0: per-bond hop, per-site U and eps.
extern "C" int emu_seg_row(int num_sites, int n_dn, int nbonds, const int* s1, const int* s2,
                           int sign_width, const double* hop, const double* u, const double* eps,
                           int uniform, unsigned ups, double eu, const double* x_row, double* y_row) {
  const u64* B = host_binom();
  if (num_sites < 1 || num_sites > 20 || n_dn < 0 || n_dn > num_sites || nbonds > ELL_MAX_BONDS) return 2;
  const i64 num_dn = (i64)B[num_sites * BINOM_N + n_dn];
  SegTables T;
  std::vector<unsigned char> blob;
  if (build_seg_tables(T, num_sites, n_dn, num_dn, nbonds, s1, s2, sign_width, eps, uniform != 0, &blob)) return 2;
  if (!T.ok) return 1;
  std::vector<uint32_t> dn;   // ascending strings of popcount n_dn (Gosper)
  if (n_dn == 0) dn.push_back(0);
  else {
    uint64_t v = (1ull << n_dn) - 1;
    while (v < (1ull << num_sites)) {
      dn.push_back((uint32_t)v);
      const uint64_t c = v & (~v + 1), r = v + c;
      v = (((r ^ v) >> 2) / c) | r;
    }
  }
  if ((i64)dn.size() != num_dn) return 3;
  SegParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.lay = T.lay; sp.e_dn_const = T.e_dn_const;
  sp.hp.u0 = u[0]; sp.hp.hop0 = hop[0];
  double s_u[32];
  for (int i = 0; i < 32; ++i) s_u[i] = i < num_sites ? u[i] : 0.0;
  std::vector<double> xs(x_row, x_row + num_dn);
  const seg_addr xs_s = (seg_addr)xs.data();
  for (i64 d = 0; d < num_dn; ++d) {
    if (uniform) y_row[d] = seg_dn_part<true>(sp, blob.data(), xs_s, hop, s_u, ups, eu, (int)d, dn[d], xs[d]);
    else y_row[d] = seg_dn_part<false>(sp, blob.data(), xs_s, hop, s_u, ups, eu, (int)d, dn[d], xs[d]);
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// K1 / K3 helpers (sector.cuh, common.cuh): the combinadic unrank / rank functions behind
// cmpy_sector_enumerate / cmpy_sector_rank and the ascending-site energy sum behind
// cmpy_weighted_elements, run on the CPU.
extern "C" int emu_sector_enumerate(int num_sites, int n, long long first, long long count, long long* out) {
  if (num_sites < 0 || num_sites > 64 || n < 0 || n > num_sites) return 2;
  for (long long i = 0; i < count; ++i) out[i] = (long long)colex_unrank(first + i, n, num_sites);
  return 0;
}
extern "C" int emu_sector_rank(const long long* states, long long count, long long* out) {
  for (long long i = 0; i < count; ++i) out[i] = (long long)colex_rank((u64)states[i]);
  return 0;
}
extern "C" int emu_weighted_elements(const long long* states, long long count, int nvals, const double* vals,
                                     double* out) {
  if (nvals < 0 || nvals > 64) return 2;
  SiteValues sv;
  sv.n = nvals;
  for (int i = 0; i < nvals; ++i) sv.v[i] = vals[i];
  for (long long i = 0; i < count; ++i) out[i] = weighted_element_dev((u64)states[i], sv);
  return 0;
}


// ---------------------------------------------------------------------------------
// Partial pull transpose (peer.cuh, opt-in chunked second half of the sharded H.v): the tile loops of

