// hubbard_seg.cuh -- K4, production kernel: two-level ("segment") Hubbard H.v.
//
// A dn string is split into (dh, dl): dl = low m = min(L, 8) bits, dh = the rest.  In the
// ascending fixed-popcount list the strings sharing dh are contiguous ("segment" of
// C(m, n_dn - popc(dh)) entries), so a row of the amplitude matrix is a ragged 2-D array
// [dh][rank(dl)].  Every dn-hop table then factorises and fits in shared memory:
//   LL bonds (both sites < m)  : per-dl compact list of (target rank, sign, bond)
//   HH bonds (both sites >= m) : per-dh list of (source segment offset, sign, bond);
//                                the gather is a conflict-free shift of a whole segment
//   LH bonds (straddling)      : per-(bond, dl) target rank + parity, per-(bond, dh)
//                                segment offset + parity
// No per-string table is ever read from global memory; the only HBM/L2 traffic is the
// vector itself: own row once (staged in smem), the up-hop neighbour rows (coalesced),
// and y.  One warp owns one segment at a time (lanes = rank of dl); all hop loops have
// warp-uniform or short, fully unrolled trip counts, so each lane keeps many independent
// loads in flight.
//
// ref for the matrix elements: cmpy/operators.py:305-527 (see hubbard.cuh).
#pragma once
#include "common.cuh"
#include "sector.cuh"
#include "hubbard.cuh"

#define SEG_MAX_LH 16

struct SegLayout {          // byte offsets into the table blob (all 16-byte aligned)
  int m, nlo, nhi, nseg, wll, nlh, n_dn;
  int off_seg_dh;           // u16 [nseg]
  int nitems, off_items;    // u32 [nitems]  work items: dh | (first rank << 16)
  int off_hi_off;           // u16 [nhi]   offset of segment dh inside a row
  int off_hi_kl;            // i8  [nhi]   popcount class of the segment, -1 invalid
  int off_cls_off;          // u16 [m+2]
  int off_lo_list;          // u8  [nlo]   class-major list of dl
  int off_ll_cnt;           // u8  [nlo]
  int off_ll_ent;           // u16 [nlo][wll]   (rank' | neg<<7 | bond<<8)
  int off_hh_ptr;           // u16 [nhi+1]
  int off_hh_ent;           // u32 [...]        (src segment offset | bond<<16 | neg<<31)
  int off_lh_lo;            // u8  [nlh][nlo]   (rank' | parity<<7)
  int off_lh_hi;            // u32 [nlh][nhi]   (offset of dh^bit | parity<<16 | bit<<17)
  int off_e_lo;             // f64 [nlo]  (only when !uniform)
  int off_e_hi;             // f64 [nhi]
  int bytes;
  int lh_lobit[SEG_MAX_LH]; // site index (inside dl) of the low end of each LH bond
  int lh_bond[SEG_MAX_LH];
};

struct SegParams {
  HubParams hp;             // vector pointers, up tables, slab, lz context
  SegLayout lay;
  const unsigned char* blob;
  double e_dn_const;        // uniform eps: eps * n_dn
};

struct __align__(16) UpEnt { i64 off; double coef; };

// WLL = compile-time padded width of the LL entry rows (8 or 16 u16 entries)
template <bool UNI, bool LZ, int WLL>
__global__ void __launch_bounds__(1024, 1) hub_seg_kernel(SegParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ UpEnt s_up[ELL_MAX_BONDS];
  __shared__ double s_hop[ELL_MAX_BONDS];
  __shared__ double s_u[32];
  const HubParams& p = sp.hp;
  const SegLayout& L = sp.lay;
  unsigned char* tab = smem_raw;
  double* xs = reinterpret_cast<double*>(smem_raw + L.bytes);
  const int tid = threadIdx.x, nt = blockDim.x;
  {  // table blob -> smem
    const uint4* src = reinterpret_cast<const uint4*>(sp.blob);
    uint4* dst = reinterpret_cast<uint4*>(tab);
    for (int k = tid; k < L.bytes / 16; k += nt) dst[k] = src[k];
    if (tid < ELL_MAX_BONDS) s_hop[tid] = p.hop[tid];
    if (tid < 32) s_u[tid] = tid < p.num_sites ? p.u[tid] : 0.0;
  }
  const uint32_t* items = reinterpret_cast<const uint32_t*>(tab + L.off_items);
  const uint16_t* hi_off = reinterpret_cast<const uint16_t*>(tab + L.off_hi_off);
  const int8_t* hi_kl = reinterpret_cast<const int8_t*>(tab + L.off_hi_kl);
  const uint16_t* cls_off = reinterpret_cast<const uint16_t*>(tab + L.off_cls_off);
  const uint8_t* lo_list = tab + L.off_lo_list;
  const uint8_t* ll_cnt = tab + L.off_ll_cnt;
  const uint16_t* ll_ent = reinterpret_cast<const uint16_t*>(tab + L.off_ll_ent);
  const uint16_t* hh_ptr = reinterpret_cast<const uint16_t*>(tab + L.off_hh_ptr);
  const uint32_t* hh_ent = reinterpret_cast<const uint32_t*>(tab + L.off_hh_ent);
  const uint8_t* lh_lo = tab + L.off_lh_lo;
  const uint32_t* lh_hi = reinterpret_cast<const uint32_t*>(tab + L.off_lh_hi);
  const double* e_lo = reinterpret_cast<const double*>(tab + L.off_e_lo);
  const double* e_hi = reinterpret_cast<const double*>(tab + L.off_e_hi);

  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  const i64 nd = p.num_dn, nu = p.num_up;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const double hop0 = p.hop0;
  double dot = 0.0;

  for (i64 r_row = blockIdx.x; r_row < p.nrows; r_row += gridDim.x) {
    const i64 u = p.row0 + r_row;
    const double* __restrict__ xr = p.x + r_row * nd;
    __syncthreads();  // previous row fully consumed (and tables loaded on first pass)
    // ---- stage the row (vectorised when 16-byte aligned) ----
    if (((nd & 1) == 0) && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0)) {
      const double2* x2 = reinterpret_cast<const double2*>(xr);
      double2* s2p = reinterpret_cast<double2*>(xs);
      for (i64 d = tid; d < nd / 2; d += nt) s2p[d] = x2[d];
    } else {
      for (i64 d = tid; d < nd; d += nt) xs[d] = xr[d];
    }
    const int cu = p.with_up ? (int)p.cnt_up[u] : 0;
    if (tid < cu) {
      const uint32_t e = p.ell_up[(i64)tid * nu + u];
      UpEnt ue;
      ue.off = ((i64)(e & ELL_TGT_MASK) - u) * nd;  // relative to the current row
      const double hv = UNI ? hop0 : s_hop[(e >> ELL_TGT_BITS) & 63u];
      ue.coef = (e >> 31) ? -hv : hv;
      s_up[tid] = ue;
    }
    __syncthreads();
    const uint32_t ups = p.up_states[u];
    const double eu = p.e_up[u];
    double* __restrict__ yr = p.y + r_row * nd;

    for (int it = warp; it < L.nitems; it += nwarps) {
      const uint32_t item = items[it];
      const int dh = (int)(item & 0xffffu);
      const int r = (int)(item >> 16) + lane;
      const int kl = hi_kl[dh];
      const int base = hi_off[dh];
      const int cb = cls_off[kl];
      const int len = cls_off[kl + 1] - cb;
      const int hp0 = hh_ptr[dh], hp1 = hh_ptr[dh + 1];
      const double ehi = UNI ? sp.e_dn_const : e_hi[dh];
      if (r < len) {
        const int dl = lo_list[cb + r];
        const double xi = xs[base + r];
        const uint32_t dns = ((uint32_t)dh << L.m) | (uint32_t)dl;
        double diag;
        if (UNI) {
          diag = eu + ehi + p.u0 * (double)__popc(ups & dns);
        } else {
          double w = 0.0;
          uint32_t both = ups & dns;
          while (both) { const int i = __ffs(both) - 1; both &= both - 1; w += s_u[i]; }
          diag = eu + (ehi + e_lo[dl]) + w;
        }
        double acc = diag * xi;
        double h = 0.0;  // UNI: signed sum of neighbours, scaled by hop0 at the end
        // ---- LL hops: compact per-dl list, entries fetched with 128-bit loads ----
        {
          const int c = ll_cnt[dl];
          const uint4* ep = reinterpret_cast<const uint4*>(ll_ent + dl * WLL);
          uint32_t w32[WLL / 2];
#pragma unroll
          for (int q = 0; q < WLL / 8; ++q) {
            const uint4 t4 = ep[q];
            w32[4 * q] = t4.x; w32[4 * q + 1] = t4.y; w32[4 * q + 2] = t4.z; w32[4 * q + 3] = t4.w;
          }
#pragma unroll
          for (int q = 0; q < WLL; ++q) {
            if (q < c) {
              const uint32_t e = (w32[q >> 1] >> ((q & 1) * 16)) & 0xffffu;
              const double v = xs[base + (int)(e & 127u)];
              const double sv = (e & 128u) ? -v : v;
              if (UNI) h += sv; else acc += sv * s_hop[e >> 8];
            }
          }
        }
        // ---- HH hops: warp-uniform list, whole-segment shifts ----
        for (int q = hp0; q < hp1; ++q) {
          const uint32_t e = hh_ent[q];
          const double v = xs[(int)(e & 0xffffu) + r];
          const double sv = (e >> 31) ? -v : v;
          if (UNI) h += sv; else acc += sv * s_hop[(e >> 16) & 63u];
        }
        // ---- LH hops ----
        for (int b = 0; b < L.nlh; ++b) {
          const uint32_t hi = lh_hi[b * L.nhi + dh];
          const uint32_t lo = lh_lo[b * L.nlo + dl];
          const uint32_t bit_lo = ((uint32_t)dl >> L.lh_lobit[b]) & 1u;
          const uint32_t bit_hi = (hi >> 17) & 1u;
          if (bit_lo != bit_hi) {
            const double v = xs[(int)(hi & 0xffffu) + (int)(lo & 127u)];
            const double sv = (((lo >> 7) ^ (hi >> 16)) & 1u) ? -v : v;
            if (UNI) h += sv; else acc += sv * s_hop[L.lh_bond[b]];
          }
        }
        if (UNI) acc += hop0 * h;
        // ---- up hops: coalesced gathers from neighbour rows (L2 / HBM) ----
        {
          const double* __restrict__ xg = xr + base + r;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          int k = 0;
          for (; k + 4 <= cu; k += 4) {
            const UpEnt e0 = s_up[k], e1 = s_up[k + 1], e2 = s_up[k + 2], e3 = s_up[k + 3];
            const double v0 = __ldg(xg + e0.off), v1 = __ldg(xg + e1.off);
            const double v2 = __ldg(xg + e2.off), v3 = __ldg(xg + e3.off);
            a0 += e0.coef * v0; a1 += e1.coef * v1; a2 += e2.coef * v2; a3 += e3.coef * v3;
          }
          for (; k < cu; ++k) {
            const UpEnt e0 = s_up[k];
            a0 += e0.coef * __ldg(xg + e0.off);
          }
          acc += (a0 + a1) + (a2 + a3);
        }
        if (LZ) {
          double w = s1 * acc;
          if (has_prev) w -= s2 * yr[base + r];
          yr[base + r] = w;
          dot += (s1 * xi) * w;
        } else {
          yr[base + r] = p.accumulate ? yr[base + r] + acc : acc;
        }
      }
    }
  }
  lz_finish<LZ>(p.lz, j, dot, red);
}

// ---------------------------------------------------------------------------------
// host: table construction
// ---------------------------------------------------------------------------------
struct SegTables {
  SegLayout lay;
  unsigned char* d_blob = nullptr;
  bool ok = false;
  int wll_pad = 8;
  double e_dn_const = 0.0;
  void release() { cudaFree(d_blob); d_blob = nullptr; ok = false; }
};

static inline int align16(int x) { return (x + 15) & ~15; }

// Builds the two-level tables for the dn species. Returns ok=false (no error) when the
// configuration is outside what the segment kernel supports (the caller falls back).
static int build_seg_tables(SegTables& T, int num_sites, int n_dn, i64 num_dn, int nbonds,
                            const int* s1, const int* s2, int sign_width, const double* eps,
                            bool uniform) {
  T.ok = false;
  const u64* B = host_binom();
  if (n_dn < 0 || n_dn > num_sites) return CMPY_OK;
  if ((i64)B[num_sites * BINOM_N + n_dn] != num_dn) return CMPY_OK;
  if (num_dn >= 65536) return CMPY_OK;
  const int m = num_sites < 8 ? num_sites : 8;
  const int hb = num_sites - m;
  if (hb > 12) return CMPY_OK;
  const int nlo = 1 << m, nhi = 1 << hb;
  SegLayout& L = T.lay;
  memset(&L, 0, sizeof(L));
  L.m = m; L.nlo = nlo; L.nhi = nhi; L.n_dn = n_dn;
  // classes of dl
  std::vector<int> cls_off(m + 2, 0);
  for (int k = 0; k <= m; ++k) cls_off[k + 1] = cls_off[k] + (int)B[m * BINOM_N + k];
  std::vector<int> lo_rank(nlo);
  std::vector<uint8_t> lo_list(nlo);
  {
    std::vector<int> fill(m + 1, 0);
    for (int v = 0; v < nlo; ++v) {
      int k = __builtin_popcount(v);
      lo_rank[v] = fill[k];
      lo_list[cls_off[k] + fill[k]] = (uint8_t)v;
      ++fill[k];
    }
  }
  // segments
  std::vector<int> hi_off(nhi, 0), hi_kl(nhi, -1);
  std::vector<uint16_t> seg_dh;
  {
    i64 off = 0;
    for (int dh = 0; dh < nhi; ++dh) {
      int kl = n_dn - __builtin_popcount(dh);
      if (kl < 0 || kl > m) continue;
      hi_kl[dh] = kl; hi_off[dh] = (int)off;
      off += (i64)B[m * BINOM_N + kl];
      seg_dh.push_back((uint16_t)dh);
    }
    if (off != num_dn) return cmpy_fail(CMPY_ERR_ARG, "segment tables: size mismatch");
  }
  L.nseg = (int)seg_dh.size();
  std::vector<uint32_t> items;
  for (uint16_t dh : seg_dh) {
    const int len = (int)B[m * BINOM_N + hi_kl[dh]];
    for (int r0 = 0; r0 < len; r0 += 32) items.push_back((uint32_t)dh | ((uint32_t)r0 << 16));
  }
  L.nitems = (int)items.size();
  // classify bonds
  std::vector<int> ll, hh, lh;
  for (int b = 0; b < nbonds; ++b) {
    if (s2[b] < m) ll.push_back(b);
    else if (s1[b] >= m) hh.push_back(b);
    else lh.push_back(b);
  }
  if ((int)lh.size() > SEG_MAX_LH) return CMPY_OK;
  L.nlh = (int)lh.size();
  auto parity = [&](u64 state, int a, int b2) {
    return __builtin_popcountll(state & between_mask(a, b2, sign_width)) & 1;
  };
  // LL: per dl compact list
  std::vector<uint8_t> ll_cnt(nlo, 0);
  std::vector<std::vector<uint16_t>> ll_rows(nlo);
  int wll = 0;
  for (int dl = 0; dl < nlo; ++dl) {
    for (int b : ll) {
      const int b1 = (dl >> s1[b]) & 1, b2 = (dl >> s2[b]) & 1;
      if (b1 == b2) continue;
      const int nl = dl ^ (1 << s1[b]) ^ (1 << s2[b]);
      const int neg = parity((u64)dl, s1[b], s2[b]);  // bits between lie inside dl
      ll_rows[dl].push_back((uint16_t)(lo_rank[nl] | (neg << 7) | (b << 8)));
    }
    ll_cnt[dl] = (uint8_t)ll_rows[dl].size();
    wll = std::max(wll, (int)ll_rows[dl].size());
  }
  if (wll > 16) return CMPY_OK;
  T.wll_pad = wll <= 8 ? 8 : 16;
  L.wll = T.wll_pad;
  // HH: per dh list
  std::vector<uint16_t> hh_ptr(nhi + 1, 0);
  std::vector<uint32_t> hh_ent;
  for (int dh = 0; dh < nhi; ++dh) {
    hh_ptr[dh] = (uint16_t)hh_ent.size();
    if (hi_kl[dh] < 0) continue;
    for (int b : hh) {
      const int a = s1[b] - m, c = s2[b] - m;
      const int b1 = (dh >> a) & 1, b2 = (dh >> c) & 1;
      if (b1 == b2) continue;
      const int nh = dh ^ (1 << a) ^ (1 << c);
      const int neg = parity((u64)dh << m, s1[b], s2[b]);
      hh_ent.push_back((uint32_t)hi_off[nh] | ((uint32_t)b << 16) | ((uint32_t)neg << 31));
    }
    if (hh_ent.size() >= 65535) return CMPY_OK;
  }
  hh_ptr[nhi] = (uint16_t)hh_ent.size();
  // LH
  std::vector<uint8_t> lh_lo((size_t)std::max(1, L.nlh) * nlo, 0);
  std::vector<uint32_t> lh_hi((size_t)std::max(1, L.nlh) * nhi, 0);
  for (int q = 0; q < L.nlh; ++q) {
    const int b = lh[q];
    const int a = s1[b], c = s2[b] - m;  // a in dl, c in dh
    L.lh_lobit[q] = a; L.lh_bond[q] = b;
    for (int dl = 0; dl < nlo; ++dl) {
      const int nl = dl ^ (1 << a);
      const int par = parity((u64)dl, a, m);  // bits of dl strictly above a (below site m)
      lh_lo[(size_t)q * nlo + dl] = (uint8_t)(lo_rank[nl] | (par << 7));
    }
    for (int dh = 0; dh < nhi; ++dh) {
      const int nh = dh ^ (1 << c);
      const int bit = (dh >> c) & 1;
      // bits of dh strictly below c, i.e. sites m .. s2-1, limited to the sign width
      const int par = parity((u64)dh << m, m - 1, s2[b]);
      const int off = (nh < nhi && hi_kl[nh] >= 0) ? hi_off[nh] : 0;
      lh_hi[(size_t)q * nhi + dh] = (uint32_t)off | ((uint32_t)par << 16) | ((uint32_t)bit << 17);
    }
  }
  // energies
  std::vector<double> e_lo(nlo, 0.0), e_hi(nhi, 0.0);
  bool eps_uniform = true;
  for (int i = 1; i < num_sites; ++i) eps_uniform = eps_uniform && (eps[i] == eps[0]);
  for (int dl = 0; dl < nlo; ++dl) { double v = 0; for (int i = 0; i < m; ++i) if (dl >> i & 1) v += eps[i]; e_lo[dl] = v; }
  for (int dh = 0; dh < nhi; ++dh) { double v = 0; for (int i = 0; i < hb; ++i) if (dh >> i & 1) v += eps[m + i]; e_hi[dh] = v; }
  T.e_dn_const = 0.0;
  { double v = 0; for (int i = 0; i < n_dn; ++i) v += eps[0]; T.e_dn_const = v; }
  (void)eps_uniform; (void)uniform;
  // layout
  int o = 0;
  L.off_seg_dh = o; o = align16(o + 2 * L.nseg);
  L.off_items = o; o = align16(o + 4 * L.nitems);
  L.off_hi_off = o; o = align16(o + 2 * nhi);
  L.off_hi_kl = o; o = align16(o + nhi);
  L.off_cls_off = o; o = align16(o + 2 * (m + 2));
  L.off_lo_list = o; o = align16(o + nlo);
  L.off_ll_cnt = o; o = align16(o + nlo);
  L.off_ll_ent = o; o = align16(o + 2 * nlo * L.wll);
  L.off_hh_ptr = o; o = align16(o + 2 * (nhi + 1));
  L.off_hh_ent = o; o = align16(o + 4 * (int)std::max<size_t>(1, hh_ent.size()));
  L.off_lh_lo = o; o = align16(o + std::max(1, L.nlh) * nlo);
  L.off_lh_hi = o; o = align16(o + 4 * std::max(1, L.nlh) * nhi);
  L.off_e_lo = o; o = align16(o + 8 * nlo);
  L.off_e_hi = o; o = align16(o + 8 * nhi);
  L.bytes = o;
  std::vector<unsigned char> blob(o, 0);
  memcpy(&blob[L.off_seg_dh], seg_dh.data(), 2 * L.nseg);
  memcpy(&blob[L.off_items], items.data(), 4 * items.size());
  { std::vector<uint16_t> t(nhi); for (int i = 0; i < nhi; ++i) t[i] = (uint16_t)hi_off[i]; memcpy(&blob[L.off_hi_off], t.data(), 2 * nhi); }
  { std::vector<int8_t> t(nhi); for (int i = 0; i < nhi; ++i) t[i] = (int8_t)hi_kl[i]; memcpy(&blob[L.off_hi_kl], t.data(), nhi); }
  { std::vector<uint16_t> t(m + 2); for (int i = 0; i < m + 2; ++i) t[i] = (uint16_t)cls_off[i]; memcpy(&blob[L.off_cls_off], t.data(), 2 * (m + 2)); }
  memcpy(&blob[L.off_lo_list], lo_list.data(), nlo);
  memcpy(&blob[L.off_ll_cnt], ll_cnt.data(), nlo);
  {
    uint16_t* dst = reinterpret_cast<uint16_t*>(&blob[L.off_ll_ent]);
    for (int dl = 0; dl < nlo; ++dl)
      for (size_t q = 0; q < ll_rows[dl].size(); ++q) dst[dl * L.wll + q] = ll_rows[dl][q];
  }
  memcpy(&blob[L.off_hh_ptr], hh_ptr.data(), 2 * (nhi + 1));
  if (!hh_ent.empty()) memcpy(&blob[L.off_hh_ent], hh_ent.data(), 4 * hh_ent.size());
  memcpy(&blob[L.off_lh_lo], lh_lo.data(), lh_lo.size());
  memcpy(&blob[L.off_lh_hi], lh_hi.data(), 4 * lh_hi.size());
  memcpy(&blob[L.off_e_lo], e_lo.data(), 8 * nlo);
  memcpy(&blob[L.off_e_hi], e_hi.data(), 8 * nhi);
  CU_CHECK(cudaMalloc(&T.d_blob, o));
  CU_CHECK(cudaMemcpy(T.d_blob, blob.data(), o, cudaMemcpyHostToDevice));
  T.ok = true;
  return CMPY_OK;
}
