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
// and y.  Each lane owns two adjacent columns: it first issues a batch of 16-byte up-hop
// gathers, does the whole dn part out of shared memory while they are in flight, then
// consumes them (memory-level parallelism without relying on occupancy: one 512-thread CTA
// per SM, up to 128 registers per thread).
//
// ref for the matrix elements: cmpy/operators.py:305-527 (see hubbard.cuh).
#pragma once
#include <string.h>
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
  int off_lo_rank;          // u8  [nlo]   rank of dl inside its popcount class
  int off_ll_cnt;           // u8  [nlo]   npos | ntot<<4
  int off_ll_ent;           // u16 [wll][nlo]   (rank'*8 | bond<<10), '+' entries first
  int off_hh_ptr;           // u32 [nhi+1]      start | npos<<16 | ntot<<24
  int off_hh_ent;           // u32 [...]        (segment byte offset | bond<<24)
  int off_lh_lo;            // u16 [nlh][nlo]   (rank'*8 | parity<<14 | bit<<15)
  int off_lh_hi;            // u32 [nlh][nhi]   (byte offset of segment dh^bit | parity<<30 | bit<<31)
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

// flip the sign of v when `neg` (0/1) is set, via the IEEE sign bit
__host__ __device__ __forceinline__ double flip_sign(double v, uint32_t neg) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(v) ^ (int)(neg << 31), __double2loint(v));
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
  b ^= (unsigned long long)(neg & 1u) << 63;
  memcpy(&v, &b, 8);
  return v;
#endif
}

// seg_dn_part is __host__ __device__ so that tests/emu/cls_emu.cu can run it on the CPU against the
// tables of build_seg_tables: shared-memory byte addresses are 32-bit shared-space addresses on the
// device (identical code to the plain uint32_t version) and plain pointers on the host.
#ifdef __CUDA_ARCH__
typedef uint32_t seg_addr;
__device__ __forceinline__ double lds_f64(uint32_t saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ int seg_popc(uint32_t v) { return __popc(v); }
__device__ __forceinline__ int seg_ffs(uint32_t v) { return __ffs(v); }
#else
typedef uintptr_t seg_addr;
inline double lds_f64(uintptr_t saddr) { return *reinterpret_cast<const double*>(saddr); }
inline int seg_popc(uint32_t v) { return __builtin_popcount(v); }
inline int seg_ffs(uint32_t v) { return __builtin_ffs((int)v); }
#endif

// All dn hops + diagonal of one amplitude (column d of the staged row).
// Table entries hold BYTE offsets; every list is ordered "+ entries, then - entries" so no
// per-hop sign arithmetic is needed (UNI). `xs_s` = shared-space address of xs[0].
template <bool UNI>
__host__ __device__ __forceinline__ double seg_dn_part(
    const SegParams& sp, const unsigned char* tab, seg_addr xs_s, const double* s_hop,
    const double* s_u, uint32_t ups, double eu, int d, uint32_t dns, double xi) {
  const SegLayout& L = sp.lay;
  const uint8_t* lo_rank = tab + L.off_lo_rank;
  const uint8_t* ll_cnt = tab + L.off_ll_cnt;
  const uint16_t* ll_ent = reinterpret_cast<const uint16_t*>(tab + L.off_ll_ent);
  const uint32_t* hh_ptr = reinterpret_cast<const uint32_t*>(tab + L.off_hh_ptr);
  const uint32_t* hh_ent = reinterpret_cast<const uint32_t*>(tab + L.off_hh_ent);
  const uint16_t* lh_lo = reinterpret_cast<const uint16_t*>(tab + L.off_lh_lo);
  const uint32_t* lh_hi = reinterpret_cast<const uint32_t*>(tab + L.off_lh_hi);
  const int dl = (int)(dns & (uint32_t)(L.nlo - 1));
  const int dh = (int)(dns >> L.m);
  const int r = lo_rank[dl];
  const seg_addr seg_s = xs_s + (uint32_t)(d - r) * 8u;  // address of the segment start
  const seg_addr r_s = xs_s + (uint32_t)r * 8u;          // xs + r
  double diag;
  if (UNI) {
    diag = eu + sp.e_dn_const + sp.hp.u0 * (double)seg_popc(ups & dns);
  } else {
    const double* e_lo = reinterpret_cast<const double*>(tab + L.off_e_lo);
    const double* e_hi = reinterpret_cast<const double*>(tab + L.off_e_hi);
    double w = 0.0;
    uint32_t both = ups & dns;
    while (both) { const int i = seg_ffs(both) - 1; both &= both - 1; w += s_u[i]; }
    diag = eu + (e_hi[dh] + e_lo[dl]) + w;
  }
  double acc = diag * xi;
  double hp = 0.0, hn = 0.0;  // UNI: sums of '+' and '-' neighbours
  {  // LL hops: per-dl list, layout [q][dl] (conflict-free), entry = rank'*8 | bond<<10
    const uint32_t c = ll_cnt[dl];
    const int cpos = (int)(c & 15u), ctot = (int)(c >> 4);
    const uint16_t* ep = ll_ent + dl;
    int q = 0;
#pragma unroll 1
    for (; q < cpos; ++q) {
      const uint32_t e = ep[q * L.nlo];
      const double v = lds_f64(seg_s + (e & 0x3ffu));
      if (UNI) hp += v; else acc += v * s_hop[e >> 10];
    }
#pragma unroll 1
    for (; q < ctot; ++q) {
      const uint32_t e = ep[q * L.nlo];
      const double v = lds_f64(seg_s + (e & 0x3ffu));
      if (UNI) hn += v; else acc -= v * s_hop[e >> 10];
    }
  }
  {  // HH hops: per-dh list (segment byte offset | bond<<24), '+' entries then '-'
    const uint32_t pp = hh_ptr[dh];
    int q = (int)(pp & 0xffffu);
    const int qpos = q + (int)((pp >> 16) & 0xffu), qtot = q + (int)(pp >> 24);
#pragma unroll 1
    for (; q < qpos; ++q) {
      const uint32_t e = hh_ent[q];
      const double v = lds_f64(r_s + (e & 0xffffffu));
      if (UNI) hp += v; else acc += v * s_hop[e >> 24];
    }
#pragma unroll 1
    for (; q < qtot; ++q) {
      const uint32_t e = hh_ent[q];
      const double v = lds_f64(r_s + (e & 0xffffffu));
      if (UNI) hn += v; else acc -= v * s_hop[e >> 24];
    }
  }
#pragma unroll 1
  for (int b = 0; b < L.nlh; ++b) {  // LH hops
    // hi: segment byte offset (bits 0-23) | parity<<30 | bit<<31 ; lo: rank'*8 | parity<<14 | bit<<15
    const uint32_t hi = lh_hi[b * L.nhi + dh];
    const uint32_t lo = lh_lo[b * L.nlo + dl];
    const uint32_t t = hi ^ (lo << 16);  // bit31 = hop allowed, bit30 = negative
    if (t >> 31) {
      const double v = lds_f64(xs_s + (hi & 0xffffffu) + (lo & 0x3ffu));
      const double sv = flip_sign(v, (t >> 30) & 1u);
      if (UNI) hp += sv; else acc += sv * s_hop[L.lh_bond[b]];
    }
  }
  if (UNI) acc += sp.hp.hop0 * (hp - hn);
  return acc;
}

// NT = threads per CTA (512: <=128 regs, 8 gathers in flight per lane; 896: <=73 regs, 5; 1024: <=64 regs, 4)

// smem: [table blob][dn strings u32[nd_pad]][row of x: double[nd]]
template <bool UNI, bool LZ, bool VEC2, int NT>
__global__ void __launch_bounds__(NT, 1) hub_seg_kernel(SegParams sp) {
  constexpr int SEG_UPG = NT <= 512 ? 8 : (NT <= 896 ? 5 : 4);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ UpEnt s_up[ELL_MAX_BONDS];
  __shared__ double s_hop[ELL_MAX_BONDS];
  __shared__ double s_u[32];
  const HubParams& p = sp.hp;
  const SegLayout& L = sp.lay;
  const i64 nd = p.num_dn, nu = p.num_up;
  const int ndi = (int)nd;
  unsigned char* tab = smem_raw;
  uint32_t* s_dn = reinterpret_cast<uint32_t*>(smem_raw + L.bytes);
  double* xs = reinterpret_cast<double*>(smem_raw + L.bytes + (((size_t)nd * 4 + 15) & ~(size_t)15));
  const int tid = threadIdx.x, nt = blockDim.x;
  {  // tables + dn strings -> smem (once per CTA)
    const uint4* src = reinterpret_cast<const uint4*>(sp.blob);
    uint4* dst = reinterpret_cast<uint4*>(tab);
    for (int k = tid; k < L.bytes / 16; k += nt) dst[k] = src[k];
    for (int k = tid; k < ndi; k += nt) s_dn[k] = p.dn_states[k];
    for (int k = tid; k < ELL_MAX_BONDS; k += nt) s_hop[k] = p.hop[k];   // CTAs may have only 32 threads
    if (tid < 32) s_u[tid] = tid < p.num_sites ? p.u[tid] : 0.0;
  }
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  double dot = 0.0;
  const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(xs);

  for (i64 r_row = blockIdx.x; r_row < p.nrows; r_row += gridDim.x) {
    const i64 u = p.row0 + r_row;
    const double* __restrict__ xr = p.x + r_row * nd;
    __syncthreads();  // previous row fully consumed (and tables loaded on the first pass)
    if (VEC2) {
      const double2* x2 = reinterpret_cast<const double2*>(xr);
      double2* s2p = reinterpret_cast<double2*>(xs);
      for (int d = tid; d < ndi / 2; d += nt) s2p[d] = x2[d];
    } else {
      for (int d = tid; d < ndi; d += nt) xs[d] = xr[d];
    }
    const int cu = p.with_up ? (int)p.cnt_up[u] : 0;
    for (int k = tid; k < cu; k += nt) {   // strided: cu can exceed a 32-thread CTA (> 32 bonds)
      const uint32_t e = p.ell_up[(i64)k * nu + u];
      UpEnt ue;
      ue.off = ((i64)(e & ELL_TGT_MASK) - u) * nd;  // relative to the current row
      const double hv = UNI ? p.hop0 : s_hop[(e >> ELL_TGT_BITS) & 63u];
      ue.coef = (e >> 31) ? -hv : hv;
      s_up[k] = ue;
    }
    __syncthreads();
    const uint32_t ups = p.up_states[u];
    const double eu = p.e_up[u];
    double* __restrict__ yr = p.y + r_row * nd;

    if (VEC2) {
      for (int d = 2 * tid; d < ndi; d += 2 * nt) {
        const double* __restrict__ xg = xr + d;
        // ---- issue the first group of up-hop gathers (coalesced 16 B per lane) ----
        double2 g[SEG_UPG];
#pragma unroll
        for (int k = 0; k < SEG_UPG; ++k)
          if (k < cu) g[k] = __ldg(reinterpret_cast<const double2*>(xg + s_up[k].off));
        // ---- dn part for both columns (shared memory only) ----
        const double2 xi = *reinterpret_cast<const double2*>(xs + d);
        const uint2 dn2 = *reinterpret_cast<const uint2*>(s_dn + d);
        double a0 = seg_dn_part<UNI>(sp, tab, xs_s, s_hop, s_u, ups, eu, d, dn2.x, xi.x);
        double a1 = seg_dn_part<UNI>(sp, tab, xs_s, s_hop, s_u, ups, eu, d + 1, dn2.y, xi.y);
        // ---- consume the gathers, then the remaining up hops ----
#pragma unroll
        for (int k = 0; k < SEG_UPG; ++k)
          if (k < cu) { const double c = s_up[k].coef; a0 += c * g[k].x; a1 += c * g[k].y; }
        for (int k0 = SEG_UPG; k0 < cu; k0 += SEG_UPG) {
#pragma unroll
          for (int k = 0; k < SEG_UPG; ++k)
            if (k0 + k < cu) g[k] = __ldg(reinterpret_cast<const double2*>(xg + s_up[k0 + k].off));
#pragma unroll
          for (int k = 0; k < SEG_UPG; ++k)
            if (k0 + k < cu) { const double c = s_up[k0 + k].coef; a0 += c * g[k].x; a1 += c * g[k].y; }
        }
        double2* yp = reinterpret_cast<double2*>(yr + d);
        // streaming stores: y is not read again before it has left L2 (microbenchmark: 2.09 vs
        // 2.18 ms for the gather part with st.cs), the cache stays with the gathered rows of x
        if (LZ) {
          double w0 = s1 * a0, w1 = s1 * a1;
          if (has_prev) { const double2 yo = *yp; w0 -= s2 * yo.x; w1 -= s2 * yo.y; }
          __stcs(yp, make_double2(w0, w1));
          dot += (s1 * xi.x) * w0 + (s1 * xi.y) * w1;
        } else if (p.accumulate) {
          const double2 yo = *yp;
          __stcs(yp, make_double2(yo.x + a0, yo.y + a1));
        } else {
          __stcs(yp, make_double2(a0, a1));
        }
      }
    } else {
      for (int d = tid; d < ndi; d += nt) {
        const double* __restrict__ xg = xr + d;
        const double xi = xs[d];
        double a0 = seg_dn_part<UNI>(sp, tab, xs_s, s_hop, s_u, ups, eu, d, s_dn[d], xi);
        double b0 = 0.0, b1 = 0.0;
        int k = 0;
        for (; k + 2 <= cu; k += 2) {
          const UpEnt e0 = s_up[k], e1 = s_up[k + 1];
          b0 += e0.coef * __ldg(xg + e0.off); b1 += e1.coef * __ldg(xg + e1.off);
        }
        if (k < cu) { const UpEnt e0 = s_up[k]; b0 += e0.coef * __ldg(xg + e0.off); }
        a0 += b0 + b1;
        if (LZ) {
          double w = s1 * a0;
          if (has_prev) w -= s2 * yr[d];
          yr[d] = w;
          dot += (s1 * xi) * w;
        } else {
          yr[d] = p.accumulate ? yr[d] + a0 : a0;
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
                            bool uniform, std::vector<unsigned char>* host_blob = nullptr) {
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
  // LL: per dl compact list, '+' entries first; entry = rank'*8 | bond<<10
  std::vector<uint8_t> ll_cnt(nlo, 0);
  std::vector<std::vector<uint16_t>> ll_rows(nlo);
  int wll = 0;
  for (int dl = 0; dl < nlo; ++dl) {
    std::vector<uint16_t> pos, negl;
    for (int b : ll) {
      const int b1 = (dl >> s1[b]) & 1, b2 = (dl >> s2[b]) & 1;
      if (b1 == b2) continue;
      const int nl = dl ^ (1 << s1[b]) ^ (1 << s2[b]);
      const int neg = parity((u64)dl, s1[b], s2[b]);  // bits between lie inside dl
      (neg ? negl : pos).push_back((uint16_t)((lo_rank[nl] * 8) | (b << 10)));
    }
    ll_rows[dl] = pos;
    ll_rows[dl].insert(ll_rows[dl].end(), negl.begin(), negl.end());
    if (ll_rows[dl].size() > 15) return CMPY_OK;
    ll_cnt[dl] = (uint8_t)(pos.size() | (ll_rows[dl].size() << 4));
    wll = std::max(wll, (int)ll_rows[dl].size());
  }
  T.wll_pad = wll < 1 ? 1 : wll;
  L.wll = T.wll_pad;
  // HH: per dh list, '+' entries first; entry = segment byte offset | bond<<24;
  // hh_ptr[dh] = start | npos<<16 | ntot<<24
  std::vector<uint32_t> hh_ptr(nhi + 1, 0);
  std::vector<uint32_t> hh_ent;
  for (int dh = 0; dh < nhi; ++dh) {
    const size_t start = hh_ent.size();
    if (start >= 65535) return CMPY_OK;
    std::vector<uint32_t> pos, negl;
    if (hi_kl[dh] >= 0) {
      for (int b : hh) {
        const int a = s1[b] - m, c = s2[b] - m;
        const int b1 = (dh >> a) & 1, b2 = (dh >> c) & 1;
        if (b1 == b2) continue;
        const int nh = dh ^ (1 << a) ^ (1 << c);
        const int neg = parity((u64)dh << m, s1[b], s2[b]);
        (neg ? negl : pos).push_back((uint32_t)(hi_off[nh] * 8) | ((uint32_t)b << 24));
      }
    }
    if (pos.size() + negl.size() > 255) return CMPY_OK;
    hh_ptr[dh] = (uint32_t)start | ((uint32_t)pos.size() << 16) | ((uint32_t)(pos.size() + negl.size()) << 24);
    hh_ent.insert(hh_ent.end(), pos.begin(), pos.end());
    hh_ent.insert(hh_ent.end(), negl.begin(), negl.end());
  }
  // LH
  std::vector<uint16_t> lh_lo((size_t)std::max(1, L.nlh) * nlo, 0);
  std::vector<uint32_t> lh_hi((size_t)std::max(1, L.nlh) * nhi, 0);
  for (int q = 0; q < L.nlh; ++q) {
    const int b = lh[q];
    const int a = s1[b], c = s2[b] - m;  // a in dl, c in dh
    L.lh_lobit[q] = a; L.lh_bond[q] = b;
    for (int dl = 0; dl < nlo; ++dl) {
      const int nl = dl ^ (1 << a);
      const int par = parity((u64)dl, a, m);  // bits of dl strictly above a (below site m)
      const int bit_lo = (dl >> a) & 1;
      lh_lo[(size_t)q * nlo + dl] = (uint16_t)((lo_rank[nl] * 8) | (par << 14) | (bit_lo << 15));
    }
    for (int dh = 0; dh < nhi; ++dh) {
      const int nh = dh ^ (1 << c);
      const int bit = (dh >> c) & 1;
      // bits of dh strictly below c, i.e. sites m .. s2-1, limited to the sign width
      const int par = parity((u64)dh << m, m - 1, s2[b]);
      const int off = (nh < nhi && hi_kl[nh] >= 0) ? hi_off[nh] : 0;
      lh_hi[(size_t)q * nhi + dh] = (uint32_t)(off * 8) | ((uint32_t)par << 30) | ((uint32_t)bit << 31);
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
  L.off_lo_rank = o; o = align16(o + nlo);
  L.off_ll_cnt = o; o = align16(o + nlo);
  L.off_ll_ent = o; o = align16(o + 2 * nlo * L.wll);
  L.off_hh_ptr = o; o = align16(o + 4 * (nhi + 1));
  L.off_hh_ent = o; o = align16(o + 4 * (int)std::max<size_t>(1, hh_ent.size()));
  L.off_lh_lo = o; o = align16(o + 2 * std::max(1, L.nlh) * nlo);
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
  { std::vector<uint8_t> t(nlo); for (int i = 0; i < nlo; ++i) t[i] = (uint8_t)lo_rank[i]; memcpy(&blob[L.off_lo_rank], t.data(), nlo); }
  memcpy(&blob[L.off_ll_cnt], ll_cnt.data(), nlo);
  {
    uint16_t* dst = reinterpret_cast<uint16_t*>(&blob[L.off_ll_ent]);  // layout [q][dl]
    for (int dl = 0; dl < nlo; ++dl)
      for (size_t q = 0; q < ll_rows[dl].size(); ++q) dst[q * nlo + dl] = ll_rows[dl][q];
  }
  memcpy(&blob[L.off_hh_ptr], hh_ptr.data(), 4 * (nhi + 1));
  if (!hh_ent.empty()) memcpy(&blob[L.off_hh_ent], hh_ent.data(), 4 * hh_ent.size());
  memcpy(&blob[L.off_lh_lo], lh_lo.data(), 2 * lh_lo.size());
  memcpy(&blob[L.off_lh_hi], lh_hi.data(), 4 * lh_hi.size());
  memcpy(&blob[L.off_e_lo], e_lo.data(), 8 * nlo);
  memcpy(&blob[L.off_e_hi], e_hi.data(), 8 * nhi);
  if (host_blob) {  // tests/emu: keep the tables on the host, no CUDA call
    *host_blob = blob;
    T.ok = true;
    return CMPY_OK;
  }
  CU_CHECK(cudaMalloc(&T.d_blob, o));
  CU_CHECK(cudaMemcpy(T.d_blob, blob.data(), o, cudaMemcpyHostToDevice));
  T.ok = true;
  return CMPY_OK;
}
