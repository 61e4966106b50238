// hubbard_cls.cuh -- K4, generation 2: class-major three-phase Hubbard H.v (uniform hop / U / eps).
//
// Same matrix elements as hubbard.cuh (ref: cmpy/operators.py:305-527); what changes is how
// one row of the amplitude matrix X[u, :] is laid out and walked in shared memory so that
// the dn-hop reads are warp-uniform, bank-conflict-free shared-memory accesses driven by ONE
// hop list per warp (no per-lane table look-ups, no divergence), and the up-hop gathers are
// 16-byte coalesced loads in a flat pass.
//
// A dn string is (dh, dl), dl = low m bits.  Strings with the same dh are contiguous in the
// ascending list (a "segment"); its length S_k = C(m, k) depends only on the class
// k = popc(dl) = n_dn - popc(dh).  Inside a class the row is a dense H_k x S_k matrix
// [jj = rank of dh][r = rank of dl].  In shared memory the row is stored class-major with an
// ODD pitch P_k > S_k (the slack slot stays 0.0 and is the target of list padding):
//   phase A  lanes along jj at fixed (k, r):  LL hops (both sites < m) map r -> r' with a hop
//            list that depends on (k, r) only; reads xs[base_k + jj*P_k + r'] (odd stride =
//            16 distinct 8-byte banks per half warp).  ys = diag * x + hop * sum_LL.
//   phase B  lanes along r at fixed dh:  HH hops (both sites >= m) map the whole segment onto
//            another segment of the same class (one list per warp, contiguous reads); LH hops
//            (one site on each side) change the class: per-lane target rank from a rank-indexed
//            table whose not-allowed entries point at the zero slack slot.  ys += hop * sums.
//   phase C  flat over the natural row, two columns per lane: y = ys + up hops, the up-hop
//            sources being 16-byte coalesced loads of the remote rows (batched for
//            memory-level parallelism).  Optional fused Lanczos epilogue.
// Each thread handles up to 3 blocks of 32 lanes of one item so that every list entry is
// decoded once per 3 amplitudes.
// Measured on B200 (tools/microbench.cu): odd-stride LDS.64 = contiguous LDS.64 = 15.8
// doubles/clk/SM, random LDS.64 = 5.2; whole-row gathers top out at ~8.5-9 TB/s of L2->SM traffic.
#pragma once
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include "hubbard_seg.cuh"

#define CLS_MAX_CLS 12
#define CLS_ZREG 96  // zero region appended to xs (target of HH list padding)

struct ClsLayout {
  int m, hb, nlo, nhi, ncls, nq, nlh, n_dn, nseg;
  int xs_elems;   // class-major padded row, without the zero region
  int na, nb;     // phase A / phase B item counts
  int S[CLS_MAX_CLS], H[CLS_MAX_CLS], P[CLS_MAX_CLS];
  int xbase[CLS_MAX_CLS], qoff[CLS_MAX_CLS], hoff[CLS_MAX_CLS];
  // byte offsets into the table blob (16-byte aligned)
  int off_item_a;    // u16 [na]   q = (k, r) pair
  int off_item_b;    // u16 [nb]   dh
  int off_k_of_q;    // u8  [nq]
  int off_dl_of_q;   // u16 [nq]   bit pattern of dl
  int off_ll_ptr;    // u32 [nq]   start | (# '+' pairs) << 16 | (# pairs) << 24
  int off_ll_ent;    // u8  [..]   r'; '+' entries first; both parts padded to pairs with S_k
  int off_dh_list;   // u16 [nseg] class-major list of dh
  int off_hi_goff;   // u16 [nhi]  offset of the segment inside a row (global layout)
  int off_hi_sbase;  // u16 [nhi]  offset of the segment inside xs
  int off_hi_k;      // u8  [nhi]  class of the segment, 0xff = none
  int off_hh_ptr;    // u32 [nhi]  start | (# '+' pairs) << 16 | (# pairs) << 24
  int off_hh_ent;    // u16 [..]   sbase of the source segment; padded with the zero region
  int off_lh_hi;     // u16 [nlh][nhi]     sbase of segment dh^bit | par << 14 | bit << 15
  int off_lh_lo;     // u8  [nlh][2][nq]   per value of the dh bit: r' | par << 7, or the slack slot
  int off_seg_delta; // i16 [nseg + 1]     (sbase - goff) of the segments in natural order
  int bytes;
};


// Long rows (more than 16 sites): the dn string is (dtop, drest), drest = low 16 bits.  For a
// fixed dtop the strings are contiguous in the row (a "sub-row" of C(16, n_dn - popc(dtop))
// amplitudes) and the class-major machinery runs on the sub-row as if it were a 16-site row.
// Bonds inside dtop map a sub-row onto a sub-row of the same length (handled like up hops:
// coalesced 16-byte gathers in phase C); bonds with one site on each side change the sub-row
// AND the rank (index map from global memory, per-lane gathers in phase C).  One launch per
// popcount of dtop (the tables depend on it).
struct LongCtx {
  int ntop;                 // sub-rows per up-row in this launch
  int row_len;              // amplitudes per sub-row
  int nsb;                  // bonds straddling site 15|16
  int shift;                // 1: every sub-row of this launch starts at an ODD element of the vector;
                            // the 16-byte column pairs are then (2i-1, 2i) instead of (2i, 2i+1)
  const uint32_t* top_val;  // [ntop]     dtop
  const int* sub_off;       // [ntop]     offset of the sub-row inside the dn row
  const int* tb_ptr;        // [ntop + 1] top-bond hop lists
  const int2* tb_ent;       //            (offset relative to this sub-row, 1 = negative)
  const int2* sb_src;       // [nsb][ntop] (relative offset of the source sub-row, flags: bit0 valid,
                            //             bit1 = dtop bit set, bit2 = parity of the dtop part)
  const uint32_t* sb_map;   // [nsb][2][row_len] rank in the source sub-row | parity << 30 | valid << 31
};

// Heisenberg / XXZ flavour of the same machinery (an up spin is a hard-core boson: hops without
// sign, uniform amplitude): only the diagonal differs,
//   E(s) = dz * (n_bonds - 2 * #antiparallel bonds),  #antiparallel = sum_i popc((s ^ (s >> delta_i)) & mask_i)
// with the bonds grouped by their site distance delta_i (ref: cmpy/models/heisenberg.py:19-40).
struct SpinDiag {
  int ndelta;
  int delta[4];
  uint32_t dmask[4];
  double e0, escale;   // E = e0 + escale * #antiparallel
};

struct ClsParams {
  HubParams hp;
  ClsLayout lay;
  LongCtx lg;
  SpinDiag sd;
  const unsigned char* blob;
  const uint16_t* pair_seg;  // global: natural segment ordinal of the first valid column of pair i |
                             // (second column starts the next segment) << 15; one table per shift
  double e_dn_const;
};

struct __align__(16) UpEnt2 { int off; int pad; double coef; };  // element offset relative to the row

// Hop lists are stored as pairs (e0, e1).  The kernels keep two accumulators (p, n) whose
// difference is the signed sum: a pair of the '+' part does p += x[e0]; n -= x[e1], a pair of
// the '-' part does n += x[e0]; p -= x[e1]; odd-length parts are padded with a slot holding 0.

// The phase bodies are __host__ __device__ so that tests/emu/cls_emu.cu can run them lane by lane
// on the CPU against the oracle (no CUDA-only intrinsics outside these two helpers).
#define CLS_HD __host__ __device__ __forceinline__
CLS_HD int cls_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
// v with its IEEE sign bit flipped when signmask = 0x80000000 (0: unchanged)
CLS_HD double cls_flip(double v, uint32_t signmask) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(v) ^ (int)signmask, __double2loint(v));
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
  b ^= (unsigned long long)signmask << 32;
  memcpy(&v, &b, 8);
  return v;
#endif
}
CLS_HD int cls_min(int a, int b) { return a < b ? a : b; }
// ---- phase A body: one (k, r) pair, T blocks of 32 dh-segments (lanes along jj) ----
template <int T, bool SPIN>
CLS_HD void cls_phase_a(const ClsLayout& L, const SpinDiag& sd,
                                            const double* __restrict__ xs,
                                            double* __restrict__ ys, const uint8_t* __restrict__ ll_ent,
                                            const uint16_t* __restrict__ dh_list, uint32_t pp, int k,
                                            int r, uint32_t dlbits, uint32_t ups, double eu, double u0,
                                            double hop0, int lane) {
  const int hk = L.H[k], pk = L.P[k], xb = L.xbase[k];
  int o[T];  // element offset of (jj, r = 0); lanes past the class recompute its last segment
#pragma unroll
  for (int t = 0; t < T; ++t) o[t] = xb + cls_min(lane + 32 * t, hk - 1) * pk;
  double ap[T], an[T];
#pragma unroll
  for (int t = 0; t < T; ++t) { ap[t] = 0.0; an[t] = 0.0; }
  const uint16_t* ent2 = reinterpret_cast<const uint16_t*>(ll_ent + (pp & 0xffffu));
  const int npos = (int)((pp >> 16) & 0xffu), ntot = (int)(pp >> 24);
  int i = 0;
#pragma unroll 1
  for (; i < npos; ++i) {
    const uint32_t e = ent2[i];
    const double* __restrict__ q0 = xs + (e & 0xffu);
    const double* __restrict__ q1 = xs + (e >> 8);
#pragma unroll
    for (int t = 0; t < T; ++t) { ap[t] += q0[o[t]]; an[t] -= q1[o[t]]; }
  }
#pragma unroll 1
  for (; i < ntot; ++i) {
    const uint32_t e = ent2[i];
    const double* __restrict__ q0 = xs + (e & 0xffu);
    const double* __restrict__ q1 = xs + (e >> 8);
#pragma unroll
    for (int t = 0; t < T; ++t) { an[t] += q0[o[t]]; ap[t] -= q1[o[t]]; }
  }
  const uint16_t* __restrict__ dhl = dh_list + L.hoff[k] + lane;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    if (lane + 32 * t < hk) {
      const uint32_t dns = ((uint32_t)dhl[32 * t] << L.m) | dlbits;
      double diag;
      if (SPIN) {  // ups carries the top bits of the spin string (dtop << 16)
        const uint32_t sfull = ups | dns;
        int cnt = 0;
        for (int i = 0; i < sd.ndelta; ++i) cnt += cls_popc((sfull ^ (sfull >> sd.delta[i])) & sd.dmask[i]);
        diag = sd.e0 + sd.escale * (double)cnt;
      } else {
        diag = eu + u0 * (double)cls_popc(ups & dns);
      }
      ys[o[t] + r] = diag * xs[o[t] + r] + hop0 * (ap[t] - an[t]);
    }
  }
}

// ---- phase B body: one segment dh, T blocks of 32 ranks (lanes along r) ----
// Lanes past the end of the segment read the following slots (inside xs) and are never stored.
template <int T>
CLS_HD void cls_phase_b(const ClsLayout& L, const double* __restrict__ xs,
                                            double* __restrict__ ys, const uint16_t* __restrict__ hh_ent,
                                            const uint16_t* __restrict__ lh_hi,
                                            const uint8_t* __restrict__ lh_lo, uint32_t pp, int k,
                                            int dh, int sb, double hop0, int lane) {
  const int sk = L.S[k];
  const double* __restrict__ xp0 = xs + lane;
  double hp[T], hn[T];
#pragma unroll
  for (int t = 0; t < T; ++t) { hp[t] = 0.0; hn[t] = 0.0; }
  const uint32_t* ent2 = reinterpret_cast<const uint32_t*>(hh_ent + (pp & 0xffffu));
  const int npos = (int)((pp >> 16) & 0xffu), ntot = (int)(pp >> 24);
  int i = 0;
#pragma unroll 1
  for (; i < npos; ++i) {
    const uint32_t e = ent2[i];
    const double* __restrict__ q0 = xp0 + (e & 0xffffu);
    const double* __restrict__ q1 = xp0 + (e >> 16);
#pragma unroll
    for (int t = 0; t < T; ++t) { hp[t] += q0[32 * t]; hn[t] -= q1[32 * t]; }
  }
#pragma unroll 1
  for (; i < ntot; ++i) {
    const uint32_t e = ent2[i];
    const double* __restrict__ q0 = xp0 + (e & 0xffffu);
    const double* __restrict__ q1 = xp0 + (e >> 16);
#pragma unroll
    for (int t = 0; t < T; ++t) { hn[t] += q0[32 * t]; hp[t] -= q1[32 * t]; }
  }
  const uint8_t* __restrict__ lo0 = lh_lo + L.qoff[k] + lane;
#pragma unroll 1
  for (int b = 0; b < L.nlh; ++b) {
    const uint32_t hi = lh_hi[b * L.nhi + dh];
    const double* __restrict__ xh = xs + (hi & 0x3fffu);
    const uint8_t* __restrict__ lo_tab = lo0 + (2 * b + (int)(hi >> 15)) * L.nq;
    const uint32_t par_hi = (hi >> 14) << 31;  // bit 14 of hi -> sign bit position
#pragma unroll
    for (int t = 0; t < T; ++t) {
      // table rows are nq long and r < 96 <= slack after the last class: in bounds for every lane
      const uint32_t lo = lo_tab[32 * t];
      const double v = xh[lo & 0x7fu];
      hp[t] += cls_flip(v, ((lo << 24) ^ par_hi) & 0x80000000u);
    }
  }
  double* __restrict__ yp0 = ys + sb + lane;
#pragma unroll
  for (int t = 0; t < T; ++t)
    if (lane + 32 * t < sk) yp0[32 * t] += hop0 * (hp[t] - hn[t]);
}

__device__ __forceinline__ int seg_delta_g(const ClsParams& cp, int si) {
  return (int)reinterpret_cast<const int16_t*>(cp.blob + cp.lay.off_seg_delta)[si];
}

// smem: [table blob][xs: xs_elems + CLS_ZREG doubles][ys: xs_elems doubles]
// UPG = 0 compiles the up-hop gathers out (row-slab launches of the sharded operator).
template <bool LZ, int NT, int UPG_, bool LONG = false, bool SPIN = false>
__global__ void __launch_bounds__(NT, 1) hub_cls_kernel(ClsParams cp) {
  constexpr bool WITH_UP = UPG_ > 0;
  constexpr int UPG = WITH_UP ? UPG_ : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ UpEnt2 s_up[ELL_MAX_BONDS + 32];
  const HubParams& p = cp.hp;
  const ClsLayout& L = cp.lay;
  const i64 nd = p.num_dn, nu = p.num_up;
  const int ndi = LONG ? cp.lg.row_len : (int)nd;   // amplitudes handled per work item
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  unsigned char* tab = smem_raw;
  double* xs = reinterpret_cast<double*>(smem_raw + L.bytes);
  const int xs_total = (L.xs_elems + CLS_ZREG + 1) & ~1;
  double* ys = xs + xs_total;
  {
    const uint4* src = reinterpret_cast<const uint4*>(cp.blob);
    uint4* dst = reinterpret_cast<uint4*>(tab);
    for (int k = tid; k < L.bytes / 16; k += NT) dst[k] = src[k];
    for (int k = tid; k < xs_total; k += NT) xs[k] = 0.0;  // slack slots and the zero region stay 0
  }
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

  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  double dot = 0.0;
  const double ediag0 = cp.e_dn_const;
  const double u0 = p.u0, hop0 = p.hop0;

  // Row-invariant phase-C bookkeeping: ys slots of the column pairs this thread owns.
  constexpr int MAXP = (13312 / 2 + NT - 1) / NT;   // rows are limited by shared memory (< 13.2 K)
  const int shift = LONG ? cp.lg.shift : 0;
  const int npairs = (ndi + shift + 1) / 2;
  uint32_t slots[MAXP];
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pi = tid + i * NT;
    slots[i] = 0;
    if (pi < npairs) {
      const uint32_t ps = cp.pair_seg[pi];
      const int si = (int)(ps & 0x7fffu), d = 2 * pi - shift;
      int slot0 = d + seg_delta_g(cp, si);
      int slot1 = d + 1 + seg_delta_g(cp, si + (int)(ps >> 15));
      if (d < 0) slot0 = slot1;          // column -1 / ndi belong to the neighbouring sub-rows
      if (d + 1 >= ndi) slot1 = slot0;
      slots[i] = (uint32_t)slot0 | ((uint32_t)slot1 << 16);
    }
  }
#ifdef CLS_TIMING
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
#define CLS_TICK(i) { const long long tn = clock64(); tph[i] += tn - tlast; tlast = tn; }
#else
#define CLS_TICK(i)
#endif
  const i64 nitems = LONG ? p.nrows * cp.lg.ntop : p.nrows;
  for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
    const i64 r_row = LONG ? item / cp.lg.ntop : item;
    const int ti = LONG ? (int)(item - r_row * cp.lg.ntop) : 0;
    const i64 u = p.row0 + r_row;
    const i64 base = r_row * nd + (LONG ? (i64)cp.lg.sub_off[ti] : 0);
    const double* __restrict__ xr = p.x + base;
    double* __restrict__ yr = p.y + base;
    const uint32_t dtop = LONG ? cp.lg.top_val[ti] : 0u;
    __syncthreads();  // previous row fully consumed (first pass: tables loaded, xs zeroed)
    CLS_TICK(0)
    // ---- stage the row: natural order -> class-major padded layout (16-byte loads, the same
    //      column-pair -> slot map as phase C) ----
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int pi = tid + i * NT;
      if (pi < npairs) {
        const int d = 2 * pi - shift;
        const bool v0 = d >= 0, v1 = d + 1 < ndi;
        if (v0 && v1) {
          const double2 v = __ldg(reinterpret_cast<const double2*>(xr + d));
          xs[slots[i] & 0xffffu] = v.x;
          xs[slots[i] >> 16] = v.y;
        } else if (v0) {
          xs[slots[i] & 0xffffu] = __ldg(xr + d);
        } else if (v1) {
          xs[slots[i] >> 16] = __ldg(xr + d + 1);
        }
      }
    }
    const int tb0 = LONG ? cp.lg.tb_ptr[ti] : 0;
    const int cu = LONG ? cp.lg.tb_ptr[ti + 1] - tb0 : ((WITH_UP && p.with_up) ? (int)p.cnt_up[u] : 0);
    if (WITH_UP && tid < ELL_MAX_BONDS + 32) {
      UpEnt2 ue;
      ue.off = 0; ue.pad = 0; ue.coef = 0.0;  // padding: never loaded, coefficient 0
      if (tid < cu) {
        if (LONG) {  // hops inside dtop: whole sub-row onto a sub-row of the same up-row
          const int2 e = cp.lg.tb_ent[tb0 + tid];
          ue.off = e.x;
          ue.coef = e.y ? -hop0 : hop0;
        } else {
          const uint32_t e = p.ell_up[(i64)tid * nu + u];
          ue.off = (int)(((i64)(e & ELL_TGT_MASK) - u) * nd);  // relative to the current row
          ue.coef = (e >> 31) ? -hop0 : hop0;
        }
      }
      s_up[tid] = ue;
    }
    const uint32_t ups = SPIN ? (dtop << 16) : p.up_states[u];
    const double eu = SPIN ? 0.0 : p.e_up[u] + ediag0 + (LONG ? u0 * (double)__popc((ups >> 16) & dtop) : 0.0);
    CLS_TICK(1)
    __syncthreads();
    CLS_TICK(2)

    if (!LONG && r_row + gridDim.x < p.nrows) {  // pull the next row of this CTA into L2 while we compute
      const char* nxt = reinterpret_cast<const char*>(xr + (i64)gridDim.x * nd);
      for (int b = tid * 128; b < ndi * 8; b += NT * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + b));
    }
    // ---- phase A: lanes along jj, LL hops + diagonal -> ys ----
    for (int it = warp; it < L.na; it += NW) {
      const int q = item_a[it];
      const int k = k_of_q[q];
      const uint32_t pp = ll_ptr[q];
      const uint32_t dlbits = dl_of_q[q];
      const int hk = L.H[k], r = q - L.qoff[k];
      if (hk <= 32) cls_phase_a<1, SPIN>(L, cp.sd, xs, ys, ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
      else if (hk <= 64) cls_phase_a<2, SPIN>(L, cp.sd, xs, ys, ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
      else cls_phase_a<3, SPIN>(L, cp.sd, xs, ys, ll_ent, dh_list, pp, k, r, dlbits, ups, eu, u0, hop0, lane);
    }
    CLS_TICK(3)
    __syncthreads();
    CLS_TICK(2)
    // ---- phase B: lanes along r, HH + LH hops accumulated into ys ----
    for (int it = warp; it < L.nb; it += NW) {
      const int dh = item_b[it];
      const int k = hi_k[dh];
      const uint32_t pp = hh_ptr[dh];
      const int sb = hi_sbase[dh], sk = L.S[k];
      if (sk <= 32) cls_phase_b<1>(L, xs, ys, hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
      else if (sk <= 64) cls_phase_b<2>(L, xs, ys, hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
      else cls_phase_b<3>(L, xs, ys, hh_ent, lh_hi, lh_lo, pp, k, dh, sb, hop0, lane);
    }
    CLS_TICK(4)
    __syncthreads();
    CLS_TICK(2)
    // ---- phase C: flat natural order, two columns per lane: up hops + store ----
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int pi = tid + i * NT;
      if (pi >= npairs) break;
      const int d = 2 * pi - shift;
      const bool v0 = d >= 0, v1 = d + 1 < ndi, full = v0 && v1;
      const double* __restrict__ xg = xr + d;
      const int slot0 = (int)(slots[i] & 0xffffu), slot1 = (int)(slots[i] >> 16);
      double a0, a1;
      auto load2 = [&](int off) -> double2 {   // 16-byte load; edge pairs (one valid column) scalar
        if (full) return __ldg(reinterpret_cast<const double2*>(xg + off));
        return make_double2(v0 ? __ldg(xg + off) : 0.0, v1 ? __ldg(xg + off + 1) : 0.0);
      };
      if (WITH_UP) {
        double2 gth[UPG];
#pragma unroll
        for (int q = 0; q < UPG; ++q) {
          const int off = s_up[q].off;
          gth[q] = (q < cu) ? load2(off) : make_double2(0.0, 0.0);
        }
        a0 = ys[slot0]; a1 = ys[slot1];
#pragma unroll
        for (int q = 0; q < UPG; ++q) {
          const double c = s_up[q].coef;
          a0 += c * gth[q].x; a1 += c * gth[q].y;
        }
#pragma unroll 1
        for (int q0 = UPG; q0 < cu; q0 += UPG) {
#pragma unroll
          for (int q = 0; q < UPG; ++q) {
            const int off = s_up[q0 + q].off;
            gth[q] = (q0 + q < cu) ? load2(off) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int q = 0; q < UPG; ++q) {
            const double c = s_up[q0 + q].coef;
            a0 += c * gth[q].x; a1 += c * gth[q].y;
          }
        }
      } else {
        a0 = ys[slot0]; a1 = ys[slot1];
      }
      if (LONG) {
#pragma unroll 1
        for (int sb = 0; sb < cp.lg.nsb; ++sb) {
          const int2 src = cp.lg.sb_src[sb * cp.lg.ntop + ti];
          if (src.y & 1) {
            const uint32_t* map = cp.lg.sb_map + ((size_t)(2 * sb + ((src.y >> 1) & 1)) * ndi + d);
            const uint32_t e0 = v0 ? map[0] : 0u, e1 = v1 ? map[1] : 0u;
            const uint32_t ptop = (uint32_t)(src.y >> 2) & 1u;
            const double* __restrict__ xsrc = xr + src.x;
            if (e0 >> 31) a0 += flip_sign(hop0 * __ldg(xsrc + (e0 & 0x3fffffffu)), ((e0 >> 30) & 1u) ^ ptop);
            if (e1 >> 31) a1 += flip_sign(hop0 * __ldg(xsrc + (e1 & 0x3fffffffu)), ((e1 >> 30) & 1u) ^ ptop);
          }
        }
      }
      double w0 = a0, w1 = a1;
      if (LZ) {
        const double x0 = xs[slot0], x1 = xs[slot1];
        w0 = s1 * a0; w1 = s1 * a1;
        if (has_prev) {
          if (v0) w0 -= s2 * yr[d];
          if (v1) w1 -= s2 * yr[d + 1];
        }
        if (v0) dot += (s1 * x0) * w0;
        if (v1) dot += (s1 * x1) * w1;
      } else if (p.acc_scale) {   // y = c1 * (H x) + c2 * y (sharded Lanczos recurrence, device scalars)
        const double c1 = p.acc_scale[0], c2 = p.acc_scale[1];
        w0 = c1 * w0 + (v0 ? c2 * yr[d] : 0.0);
        w1 = c1 * w1 + (v1 ? c2 * yr[d + 1] : 0.0);
      } else if (p.accumulate) {
        if (v0) w0 += yr[d];
        if (v1) w1 += yr[d + 1];
      }
      if (full) *reinterpret_cast<double2*>(yr + d) = make_double2(w0, w1);
      else if (v0) yr[d] = w0;
      else if (v1) yr[d + 1] = w1;
    }
    CLS_TICK(5)
  }
#ifdef CLS_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 992))
    printf("cls timing tid %d: wait_prev %lld copyin %lld barriers %lld A %lld B %lld C %lld (cycles)\n", tid,
           tph[0], tph[1], tph[2], tph[3], tph[4], tph[5]);
#endif
  lz_finish<LZ>(p.lz, j, dot, red);
}

// ---------------------------------------------------------------------------------
// host: table construction
// ---------------------------------------------------------------------------------
struct ClsTables {
  ClsLayout lay;
  unsigned char* d_blob = nullptr;
  uint16_t* d_pair_seg = nullptr;    // pairs (2i, 2i+1)
  uint16_t* d_pair_seg1 = nullptr;   // pairs (2i-1, 2i): sub-rows that start at an odd element
  bool ok = false;
  double e_dn_const = 0.0;
  size_t smem = 0;
  void release() {
    cudaFree(d_blob); cudaFree(d_pair_seg); cudaFree(d_pair_seg1);
    d_blob = nullptr; d_pair_seg = nullptr; d_pair_seg1 = nullptr; ok = false;
  }
};

// Builds the class-major tables for the dn species.  ok=false (no error) when the sector is
// outside what the kernel supports (non-sector string list, odd row length, row too long for
// shared memory, classes longer than 96).
// Host-only part (no CUDA calls): layout, table blob and the column-pair maps.
struct ClsHost {
  ClsLayout lay;
  std::vector<unsigned char> blob;
  std::vector<uint16_t> pair_seg, pair_seg1;
  bool ok = false;
  double e_dn_const = 0.0;
  size_t smem = 0;
};

static int build_cls_host(ClsHost& T, int num_sites, int n_dn, i64 num_dn, int nbonds,
                          const int* s1, const int* s2, int sign_width, const double* eps,
                          i64 smem_optin) {
  T.ok = false;
  const u64* B = host_binom();
  if (n_dn < 0 || n_dn > num_sites || num_sites < 2) return CMPY_OK;
  if ((i64)B[num_sites * BINOM_N + n_dn] != num_dn) return CMPY_OK;
  if (num_dn >= 16384 || (num_dn & 1)) return CMPY_OK;
  const int m = (num_sites + 1) / 2;
  if (m > 10) return CMPY_OK;
  const int hb = num_sites - m;
  const int nlo = 1 << m, nhi = 1 << hb;
  ClsLayout& L = T.lay;
  memset(&L, 0, sizeof(L));
  L.m = m; L.hb = hb; L.nlo = nlo; L.nhi = nhi; L.ncls = m + 1; L.nq = nlo; L.n_dn = n_dn;
  if (L.ncls > CLS_MAX_CLS) return CMPY_OK;
  int xoff = 0, qo = 0, ho = 0;
  for (int k = 0; k <= m; ++k) {
    const int hk = n_dn - k;
    L.S[k] = (int)B[m * BINOM_N + k];
    L.H[k] = (hk >= 0 && hk <= hb) ? (int)B[hb * BINOM_N + hk] : 0;
    L.P[k] = (L.S[k] + 1) | 1;  // odd and > S_k: at least one slack slot per segment
    L.xbase[k] = xoff; xoff += L.H[k] * L.P[k];
    L.qoff[k] = qo; qo += L.S[k];
    L.hoff[k] = ho; ho += L.H[k];
    if (L.H[k] > 0 && (L.S[k] > 96 || L.H[k] > 96)) return CMPY_OK;  // bodies cover 3 blocks of 32
  }
  L.xs_elems = xoff;
  if (xoff + CLS_ZREG >= 16384) return CMPY_OK;
  const int nseg = ho;
  L.nseg = nseg;
  const int zbase = xoff;  // start of the zero region
  // ranks of dl inside its class, class-major lists
  std::vector<int> lo_rank(nlo), dl_of_q(nlo), k_of_q(nlo);
  {
    std::vector<int> fill(m + 1, 0);
    for (int v = 0; v < nlo; ++v) {
      const int k = __builtin_popcount(v);
      lo_rank[v] = fill[k];
      dl_of_q[L.qoff[k] + fill[k]] = v;
      k_of_q[L.qoff[k] + fill[k]] = k;
      ++fill[k];
    }
  }
  std::vector<int> hi_k(nhi, 0xff), hi_goff(nhi, 0), hi_sbase(nhi, 0), dh_list(std::max(nseg, 1), 0);
  std::vector<int> seg_delta;  // natural order
  std::vector<uint16_t>& pair_seg = T.pair_seg;
  std::vector<uint16_t>& pair_seg1 = T.pair_seg1;
  pair_seg.assign((size_t)num_dn / 2, 0);
  pair_seg1.assign((size_t)num_dn / 2 + 1, 0);
  std::vector<int> seg_of((size_t)num_dn, 0);
  {
    std::vector<int> fill(m + 1, 0);
    i64 off = 0;
    int ordinal = 0;
    for (int dh = 0; dh < nhi; ++dh) {
      const int k = n_dn - __builtin_popcount(dh);
      if (k < 0 || k > m) continue;
      hi_k[dh] = k; hi_goff[dh] = (int)off;
      hi_sbase[dh] = L.xbase[k] + fill[k] * L.P[k];
      dh_list[L.hoff[k] + fill[k]] = dh;
      ++fill[k];
      seg_delta.push_back(hi_sbase[dh] - hi_goff[dh]);
      for (i64 d = off; d < off + L.S[k]; ++d) {
        seg_of[d] = ordinal;
        if ((d & 1) == 0) pair_seg[d / 2] = (uint16_t)(ordinal | ((d + 1 == off + L.S[k]) ? 0x8000 : 0));
      }
      off += L.S[k];
      ++ordinal;
    }
    if (off != num_dn) return cmpy_fail(CMPY_ERR_ARG, "class tables: size mismatch");
    seg_delta.push_back(0);
    if (ordinal >= 0x8000) return CMPY_OK;
    // shifted pairs (2i-1, 2i): ordinal of the first valid column, flag = second column in the next segment
    for (i64 pi = 0; pi <= num_dn / 2; ++pi) {
      const i64 d0 = 2 * pi - 1, d1 = 2 * pi;
      if (d0 < 0) pair_seg1[pi] = (uint16_t)seg_of[d1];
      else if (d1 >= num_dn) pair_seg1[pi] = (uint16_t)seg_of[d0];
      else pair_seg1[pi] = (uint16_t)(seg_of[d0] | (seg_of[d1] != seg_of[d0] ? 0x8000 : 0));
    }
  }
  // bonds
  std::vector<int> ll, hh, lh;
  for (int b = 0; b < nbonds; ++b) {
    if (s2[b] < m) ll.push_back(b);
    else if (s1[b] >= m) hh.push_back(b);
    else lh.push_back(b);
  }
  L.nlh = (int)lh.size();
  auto parity = [&](u64 state, int a, int b2) {
    return __builtin_popcountll(state & between_mask(a, b2, sign_width)) & 1;
  };
  // pack a ('+' list, '-' list) into pairs, each part padded to an even length with `dummy`
  auto pack = [&](const std::vector<uint16_t>& pos, const std::vector<uint16_t>& neg, uint16_t dummy,
                  auto& out, uint32_t& ptr) -> bool {
    if (out.size() & 1) out.push_back(dummy);
    const size_t start = out.size();
    const size_t np2 = (pos.size() + 1) / 2, nn2 = (neg.size() + 1) / 2;
    if (start >= 65536 || np2 + nn2 > 255) return false;
    typedef typename std::remove_reference<decltype(out)>::type::value_type E;
    for (size_t i = 0; i < 2 * np2; ++i) out.push_back((E)(i < pos.size() ? pos[i] : dummy));
    for (size_t i = 0; i < 2 * nn2; ++i) out.push_back((E)(i < neg.size() ? neg[i] : dummy));
    ptr = (uint32_t)start | ((uint32_t)np2 << 16) | ((uint32_t)(np2 + nn2) << 24);
    return true;
  };
  std::vector<uint32_t> ll_ptr(nlo, 0), hh_ptr(nhi, 0);
  std::vector<uint8_t> ll_ent;   // ranks and the slack slot are <= 96
  std::vector<uint16_t> hh_ent;
  for (int q = 0; q < nlo; ++q) {
    const int dl = dl_of_q[q], k = k_of_q[q];
    std::vector<uint16_t> pos, negl;
    for (int b : ll) {
      const int b1 = (dl >> s1[b]) & 1, b2 = (dl >> s2[b]) & 1;
      if (b1 == b2) continue;
      const int nl = dl ^ (1 << s1[b]) ^ (1 << s2[b]);
      (parity((u64)dl, s1[b], s2[b]) ? negl : pos).push_back((uint16_t)lo_rank[nl]);
    }
    if (!pack(pos, negl, (uint16_t)L.S[k], ll_ent, ll_ptr[q])) return CMPY_OK;
  }
  for (int dh = 0; dh < nhi; ++dh) {
    std::vector<uint16_t> pos, negl;
    if (hi_k[dh] != 0xff) {
      for (int b : hh) {
        const int a = s1[b] - m, c = s2[b] - m;
        const int b1 = (dh >> a) & 1, b2 = (dh >> c) & 1;
        if (b1 == b2) continue;
        const int nh = dh ^ (1 << a) ^ (1 << c);
        (parity((u64)dh << m, s1[b], s2[b]) ? negl : pos).push_back((uint16_t)hi_sbase[nh]);
      }
    }
    if (!pack(pos, negl, (uint16_t)zbase, hh_ent, hh_ptr[dh])) return CMPY_OK;
  }
  // LH tables
  std::vector<uint16_t> lh_hi((size_t)std::max(1, L.nlh) * nhi, 0);
  std::vector<uint8_t> lh_lo((size_t)std::max(1, L.nlh) * 2 * nlo + 96, 0);  // + slack for lanes past a segment
  for (int qb = 0; qb < L.nlh; ++qb) {
    const int b = lh[qb];
    const int a = s1[b], c = s2[b] - m;  // a inside dl, c inside dh
    for (int dh = 0; dh < nhi; ++dh) {
      const int nh = dh ^ (1 << c);
      const int bit = (dh >> c) & 1;
      const int par = parity((u64)dh << m, m - 1, s2[b]);  // bits of dh strictly below c
      const bool valid = hi_k[dh] != 0xff && hi_k[nh] != 0xff;
      lh_hi[(size_t)qb * nhi + dh] = (uint16_t)((valid ? hi_sbase[nh] : zbase) | (par << 14) | (bit << 15));
    }
    for (int beta = 0; beta < 2; ++beta) {  // beta = value of the dh bit
      for (int q = 0; q < nlo; ++q) {
        const int dl = dl_of_q[q], k = k_of_q[q];
        const int bit_lo = (dl >> a) & 1;
        const int kp = k + (beta ? 1 : -1);  // class of the source segment
        uint8_t e;
        if (kp < 0 || kp > m || L.H[kp] == 0) {
          e = 0;  // the hi entry points at the zero region
        } else if (bit_lo == beta) {
          e = (uint8_t)L.S[kp];  // not allowed: slack slot of the source segment (0.0)
        } else {
          const int nl = dl ^ (1 << a);
          const int par = parity((u64)dl, a, m);  // bits of dl strictly above a
          e = (uint8_t)(lo_rank[nl] | (par << 7));
        }
        lh_lo[((size_t)qb * 2 + beta) * nlo + q] = e;
      }
    }
  }
  // work items
  std::vector<uint16_t> item_a, item_b;
  for (int k = 0; k <= m; ++k)
    if (L.H[k] > 0)
      for (int r = 0; r < L.S[k]; ++r) item_a.push_back((uint16_t)(L.qoff[k] + r));
  for (int dh = 0; dh < nhi; ++dh)
    if (hi_k[dh] != 0xff) item_b.push_back((uint16_t)dh);
  L.na = (int)item_a.size();
  L.nb = (int)item_b.size();
  // energies: eps uniform -> eps * n_dn summed like weighted_element (ascending adds)
  { double v = 0; for (int i = 0; i < n_dn; ++i) v += eps[0]; T.e_dn_const = v; }
  // blob layout
  int o = 0;
  auto place = [&](int& off, size_t bytes) { off = o; o = align16(o + (int)std::max<size_t>(bytes, 16)); };
  place(L.off_item_a, 2 * item_a.size());
  place(L.off_item_b, 2 * item_b.size());
  place(L.off_k_of_q, nlo);
  place(L.off_dl_of_q, 2 * nlo);
  place(L.off_ll_ptr, 4 * nlo);
  place(L.off_ll_ent, ll_ent.size());
  place(L.off_dh_list, 2 * dh_list.size());
  place(L.off_hi_goff, 2 * nhi);
  place(L.off_hi_sbase, 2 * nhi);
  place(L.off_hi_k, nhi);
  place(L.off_hh_ptr, 4 * nhi);
  place(L.off_hh_ent, 2 * hh_ent.size());
  place(L.off_lh_hi, 2 * lh_hi.size());
  place(L.off_lh_lo, lh_lo.size());
  place(L.off_seg_delta, 2 * seg_delta.size());
  L.bytes = o;
  const int xs_total = (L.xs_elems + CLS_ZREG + 1) & ~1;
  T.smem = (size_t)L.bytes + sizeof(double) * ((size_t)xs_total + (size_t)L.xs_elems);
  if (T.smem + 2560 > (size_t)smem_optin) return CMPY_OK;  // + static smem of the kernel
  std::vector<unsigned char>& blob = T.blob;
  blob.assign(o, 0);
  auto put = [&](int off, const void* src, size_t bytes) { if (bytes) memcpy(&blob[off], src, bytes); };
  auto narrow16 = [](const std::vector<int>& v) { std::vector<uint16_t> t(v.size()); for (size_t i = 0; i < v.size(); ++i) t[i] = (uint16_t)v[i]; return t; };
  auto narrow8 = [](const std::vector<int>& v) { std::vector<uint8_t> t(v.size()); for (size_t i = 0; i < v.size(); ++i) t[i] = (uint8_t)v[i]; return t; };
  { auto t = narrow16(dl_of_q); put(L.off_dl_of_q, t.data(), 2 * t.size()); }
  put(L.off_ll_ptr, ll_ptr.data(), 4 * ll_ptr.size());
  put(L.off_ll_ent, ll_ent.data(), ll_ent.size());
  { auto t = narrow16(dh_list); put(L.off_dh_list, t.data(), 2 * t.size()); }
  put(L.off_hh_ent, hh_ent.data(), 2 * hh_ent.size());
  { std::vector<int16_t> t(seg_delta.size()); for (size_t i = 0; i < t.size(); ++i) t[i] = (int16_t)seg_delta[i]; put(L.off_seg_delta, t.data(), 2 * t.size()); }
  put(L.off_item_a, item_a.data(), 2 * item_a.size());
  put(L.off_item_b, item_b.data(), 2 * item_b.size());
  { auto t = narrow8(k_of_q); put(L.off_k_of_q, t.data(), t.size()); }
  { auto t = narrow16(hi_goff); put(L.off_hi_goff, t.data(), 2 * t.size()); }
  { auto t = narrow16(hi_sbase); put(L.off_hi_sbase, t.data(), 2 * t.size()); }
  { auto t = narrow8(hi_k); put(L.off_hi_k, t.data(), t.size()); }
  put(L.off_hh_ptr, hh_ptr.data(), 4 * hh_ptr.size());
  put(L.off_lh_hi, lh_hi.data(), 2 * lh_hi.size());
  put(L.off_lh_lo, lh_lo.data(), lh_lo.size());
  T.ok = true;
  return CMPY_OK;
}

// Builds the class-major tables for the dn species and uploads them.  ok=false (no error) when
// the sector is outside what the kernel supports.
static int build_cls_tables(ClsTables& T, int num_sites, int n_dn, i64 num_dn, int nbonds,
                            const int* s1, const int* s2, int sign_width, const double* eps,
                            i64 smem_optin) {
  T.ok = false;
  ClsHost H;
  int rc = build_cls_host(H, num_sites, n_dn, num_dn, nbonds, s1, s2, sign_width, eps, smem_optin);
  if (rc || !H.ok) return rc;
  T.lay = H.lay; T.e_dn_const = H.e_dn_const; T.smem = H.smem;
  CU_CHECK(cudaMalloc(&T.d_blob, H.blob.size()));
  CU_CHECK(cudaMemcpy(T.d_blob, H.blob.data(), H.blob.size(), cudaMemcpyHostToDevice));
  CU_CHECK(cudaMalloc(&T.d_pair_seg, sizeof(uint16_t) * std::max<size_t>(H.pair_seg.size(), 1)));
  CU_CHECK(cudaMemcpy(T.d_pair_seg, H.pair_seg.data(), sizeof(uint16_t) * H.pair_seg.size(), cudaMemcpyHostToDevice));
  CU_CHECK(cudaMalloc(&T.d_pair_seg1, sizeof(uint16_t) * H.pair_seg1.size()));
  CU_CHECK(cudaMemcpy(T.d_pair_seg1, H.pair_seg1.data(), sizeof(uint16_t) * H.pair_seg1.size(), cudaMemcpyHostToDevice));
  T.ok = true;
  return CMPY_OK;
}

// ---------------------------------------------------------------------------------
// host: long rows (more than 16 sites), see LongCtx
// ---------------------------------------------------------------------------------
#define LONG_RBITS 16

struct LongSet {            // all sub-rows whose dtop has the same popcount
  ClsTables cls;            // class-major tables of the 16-site sub-row sector
  int pt = 0, ntop = 0, row_len = 0, shift = 0;
  uint32_t* d_top_val = nullptr;
  int* d_sub_off = nullptr;
  int* d_tb_ptr = nullptr;
  int2* d_tb_ent = nullptr;
  int2* d_sb_src = nullptr;
  uint32_t* d_sb_map = nullptr;
  void release() {
    cls.release();
    cudaFree(d_top_val); cudaFree(d_sub_off); cudaFree(d_tb_ptr); cudaFree(d_tb_ent);
    cudaFree(d_sb_src); cudaFree(d_sb_map);
    d_top_val = nullptr; d_sub_off = nullptr; d_tb_ptr = nullptr; d_tb_ent = nullptr;
    d_sb_src = nullptr; d_sb_map = nullptr;
  }
};

struct LongTables {
  std::vector<LongSet> sets;
  bool ok = false;
  int nsb = 0;
  double e_dn_const = 0.0;
  void release() { for (auto& s : sets) s.release(); sets.clear(); ok = false; }
};

template <typename T>
static int upload_vec(T*& dptr, const std::vector<T>& v) {
  CU_CHECK(cudaMalloc(&dptr, sizeof(T) * std::max<size_t>(v.size(), 1)));
  if (!v.empty()) CU_CHECK(cudaMemcpy(dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return CMPY_OK;
}

// `skipped` (optional): popcounts of the top bits whose sub-rows the class-major kernel cannot
// take (odd length ...); without it such a sector makes the whole table set unavailable.
// Host mirror of the long-row tables (tests/emu only): with `mirror` set, build_long_tables makes no
// CUDA call and leaves every table in host vectors instead of uploading it.
struct LongSetHost {
  ClsHost cls;
  int pt = 0, ntop = 0, row_len = 0, shift = 0;
  std::vector<uint32_t> top_val, sb_map;
  std::vector<int> sub_off, tb_ptr;
  std::vector<int2> tb_ent, sb_src;
};
struct LongHost {
  std::vector<LongSetHost> sets;
  int nsb = 0;
  double e_dn_const = 0.0;
};

static int build_long_tables(LongTables& T, int num_sites, int n_dn, i64 num_dn, int nbonds,
                             const int* s1, const int* s2, int sign_width, const double* eps,
                             i64 smem_optin, std::vector<int>* skipped = nullptr,
                             LongHost* mirror = nullptr) {
  if (!mirror) T.release();
  const u64* B = host_binom();
  const int R = LONG_RBITS, tb = num_sites - R;
  if (tb < 1 || tb > 16 || n_dn < 0 || n_dn > num_sites) return CMPY_OK;
  if ((i64)B[num_sites * BINOM_N + n_dn] != num_dn || num_dn >= (1ll << 31)) return CMPY_OK;
  const int ntopall = 1 << tb;
  // sub-row offsets in the ascending (natural) order of the dn strings
  std::vector<i64> sub_off(ntopall, -1);
  {
    i64 off = 0;
    for (int dt = 0; dt < ntopall; ++dt) {
      const int nr = n_dn - __builtin_popcount(dt);
      if (nr < 0 || nr > R) continue;
      sub_off[dt] = off;
      off += (i64)B[R * BINOM_N + nr];
    }
    if (off != num_dn) return cmpy_fail(CMPY_ERR_ARG, "long tables: size mismatch");
  }
  std::vector<int> lo1, lo2, top, strad;
  for (int b = 0; b < nbonds; ++b) {
    if (s2[b] < R) { lo1.push_back(s1[b]); lo2.push_back(s2[b]); }
    else if (s1[b] >= R) top.push_back(b);
    else strad.push_back(b);
  }
  T.nsb = (int)strad.size();
  auto parity = [&](u64 state, int a, int b2) {
    return __builtin_popcountll(state & between_mask(a, b2, sign_width)) & 1;
  };
  { double v = 0; for (int i = 0; i < n_dn; ++i) v += eps[0]; T.e_dn_const = v; }
  for (int pt = 0; pt <= tb; ++pt) {
    const int nr = n_dn - pt;
    if (nr < 0 || nr > R) continue;
    LongSet S;
    S.pt = pt;
    S.row_len = (int)B[R * BINOM_N + nr];
    LongSetHost HS;
    int rc;
    bool cls_ok;
    if (mirror) {
      rc = build_cls_host(HS.cls, R, nr, S.row_len, (int)lo1.size(), lo1.data(), lo2.data(), sign_width,
                          eps, smem_optin);
      cls_ok = HS.cls.ok;
    } else {
      rc = build_cls_tables(S.cls, R, nr, S.row_len, (int)lo1.size(), lo1.data(), lo2.data(), sign_width,
                            eps, smem_optin);
      cls_ok = S.cls.ok;
    }
    if (rc) { S.release(); return rc; }
    if (!cls_ok) {
      S.release();
      if (skipped) { skipped->push_back(pt); continue; }
      T.release();
      return CMPY_OK;
    }
    std::vector<uint32_t> top_val;
    for (int dt = 0; dt < ntopall; ++dt)
      if (__builtin_popcount(dt) == pt) top_val.push_back((uint32_t)dt);
    S.ntop = (int)top_val.size();
    {  // 16-byte alignment of the sub-rows: all even (shift 0) or all odd (shift 1) element offsets
      int n_odd = 0;
      for (uint32_t dt : top_val) n_odd += (int)(sub_off[dt] & 1);
      if (n_odd != 0 && n_odd != S.ntop) {
        S.release();
        if (skipped) { skipped->push_back(pt); continue; }
        T.release();
        return CMPY_OK;
      }
      S.shift = n_odd ? 1 : 0;
    }
    std::vector<int> so(S.ntop), tb_ptr(S.ntop + 1, 0);
    std::vector<int2> tb_ent, sb_src((size_t)std::max(1, T.nsb) * S.ntop, make_int2(0, 0));
    for (int ti = 0; ti < S.ntop; ++ti) {
      const int dt = (int)top_val[ti];
      so[ti] = (int)sub_off[dt];
      tb_ptr[ti] = (int)tb_ent.size();
      for (int b : top) {
        const int a = s1[b] - R, c = s2[b] - R;
        if (((dt >> a) & 1) == ((dt >> c) & 1)) continue;
        const int nt = dt ^ (1 << a) ^ (1 << c);
        const int neg = parity((u64)dt << R, s1[b], s2[b]);
        tb_ent.push_back(make_int2((int)(sub_off[nt] - sub_off[dt]), neg));
      }
      if ((int)tb_ent.size() - tb_ptr[ti] > ELL_MAX_BONDS) { S.release(); T.release(); return CMPY_OK; }
      for (int q = 0; q < T.nsb; ++q) {
        const int b = strad[q];
        const int c = s2[b] - R;
        const int bit = (dt >> c) & 1;
        const int nt = dt ^ (1 << c);
        const int nrs = nr + (bit ? 1 : -1);
        int flags = 0, rel = 0;
        if (nrs >= 0 && nrs <= R && sub_off[nt] >= 0) {
          const int ptop = parity((u64)dt << R, R - 1, s2[b]);  // dtop bits strictly below c
          flags = 1 | (bit << 1) | (ptop << 2);
          rel = (int)(sub_off[nt] - sub_off[dt]);
        }
        sb_src[(size_t)q * S.ntop + ti] = make_int2(rel, flags);
      }
    }
    tb_ptr[S.ntop] = (int)tb_ent.size();
    // index maps of the straddling bonds for this n_rest (both directions)
    std::vector<uint32_t> sb_map((size_t)std::max(1, T.nsb) * 2 * S.row_len, 0);
    if (T.nsb > 0) {
      std::vector<uint32_t> strings;
      for (uint32_t v = 0; v < (1u << R); ++v)
        if (__builtin_popcount(v) == nr) strings.push_back(v);
      auto rank = [&](uint32_t v) {
        i64 r = 0; int k = 0;
        while (v) { const int pos = __builtin_ctz(v); v &= v - 1; ++k; r += (i64)B[pos * BINOM_N + k]; }
        return (uint32_t)r;
      };
      for (int q = 0; q < T.nsb; ++q) {
        const int a = s1[strad[q]];
        for (int dir = 0; dir < 2; ++dir)
          for (int r = 0; r < S.row_len; ++r) {
            const uint32_t v = strings[r];
            const int bit_a = (v >> a) & 1;
            uint32_t e = 0;
            // dir = 1: the dtop bit is set, the particle comes from dtop into site a (must be empty);
            // dir = 0: the particle leaves site a (must be occupied) towards dtop
            if (bit_a != dir) {
              const uint32_t nv = v ^ (1u << a);
              const uint32_t par = (uint32_t)parity((u64)v, a, R);  // drest bits strictly above a
              e = rank(nv) | (par << 30) | (1u << 31);
            }
            sb_map[((size_t)q * 2 + dir) * S.row_len + r] = e;
          }
      }
    }
    if (mirror) {
      HS.pt = pt; HS.ntop = S.ntop; HS.row_len = S.row_len; HS.shift = S.shift;
      HS.top_val = top_val; HS.sub_off = so; HS.tb_ptr = tb_ptr; HS.tb_ent = tb_ent;
      HS.sb_src = sb_src; HS.sb_map = sb_map;
      mirror->nsb = T.nsb; mirror->e_dn_const = T.e_dn_const;
      mirror->sets.push_back(std::move(HS));
      continue;
    }
    rc = upload_vec(S.d_top_val, top_val);
    if (!rc) rc = upload_vec(S.d_sub_off, so);
    if (!rc) rc = upload_vec(S.d_tb_ptr, tb_ptr);
    if (!rc) rc = upload_vec(S.d_tb_ent, tb_ent);
    if (!rc) rc = upload_vec(S.d_sb_src, sb_src);
    if (!rc) rc = upload_vec(S.d_sb_map, sb_map);
    if (rc) { S.release(); T.release(); return rc; }
    T.sets.push_back(S);
  }
  T.ok = !T.sets.empty();
  return CMPY_OK;
}
