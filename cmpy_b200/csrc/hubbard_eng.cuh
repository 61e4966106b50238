// hubbard_eng.cuh -- K4, generation 3: the class-major row engine with warp-uniform hop lists in the
// CONSTANT bank (uniform datapath) -- Hubbard H.v for uniform hop / U / eps, rows of <= 16 sites.
//
// Same matrix elements as hubbard.cuh (ref: cmpy/operators.py:305-527, cmpy/models/hubbard.py:13-22)
// and the same class-major view of a row as hubbard_cls.cuh: a dn string is (dh, dl), dl = low m bits,
// class k = popc(dl); inside a class the row is a dense H_k x S_k matrix [jj = rank of dh][r = rank of dl],
// stored in shared memory with an ODD pitch P_k >= S_k (lanes along jj: odd stride, lanes along r:
// contiguous -- both conflict-free).  What is new:
//
//   * every table a warp walks (hop lists, per-column / per-segment descriptors, task lists) lives in
//     the kernel-parameter constant bank (`__grid_constant__`, <= 32 KB).  The warp index is made
//     warp-uniform with one SHFL, so ptxas keeps the whole list traversal on the uniform datapath:
//     one `LDCU.U16 URx, c[0x0][URy+..]` per list entry and `LDS.64 R, [Rlane + URx (+ imm)]` per
//     32 hop terms -- no address arithmetic, no table look-up through the LSU, no list decode in the
//     vector pipes.  Inner loops: 1 + 2T instructions per list entry and T blocks of 32 lanes.
//   * phase A (lanes along jj): LL hops + LH hops + diagonal -> ys (in units of hop);
//     phase B (lanes along r):  HH hops + ys -> y, stored STRAIGHT to global memory: a segment is
//     contiguous in the row, so the store is coalesced.  There is no third "flat" phase, no
//     column-pair -> slot map, and the optional up-hop row gathers / Lanczos epilogue run in phase B's
//     epilogue with the gathers issued ahead of the hop loop.
//   * the row is staged with cp.async (LDGSTS, 8 bytes: segments are only 8-byte aligned in the row),
//     one segment per warp pass, no register staging.
//   * work is cut on the host into cost-balanced pieces per warp (longest-processing-time first).
//
// The phase bodies are __host__ __device__ so that tests/emu/eng_emu.cu runs them lane by lane on the
// CPU against a direct evaluation of (D + T_dn) x.
#pragma once
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include "hubbard.cuh"

#define ENG_MAX_CLS 9        // m <= 8 low bits -> classes 0..8
#define ENG_MAX_LH 4         // LH bonds whose per-lane state phase A keeps in registers
#define ENG_MAX_TASKS 160
#define ENG_MAX_ENT 2304     // u16 entries per hop-list table
#define ENG_MAX_Q 256        // dl values (2^m)
#define ENG_MAX_SEG 256      // dh values with a non-empty class
#define ENG_ZREG 96          // zeros behind xs: target of inactive LH lanes and of tail-lane over-reads
#define ENG_MAX_WARPS 32
#define ENG_UPB 4            // up-hop row gathers in flight per lane and block

// ---- constant-bank tables (kernel parameter) -----------------------------------------------------
struct EngConst {
  int m, hb, n_dn, nq, nseg, nlh;
  int xs_elems, zoff;                 // padded class-major row (doubles); zoff = first zero slot
  int row_len;                        // amplitudes per row
  int S[ENG_MAX_CLS], H[ENG_MAX_CLS];
  int P8[ENG_MAX_CLS];                // pitch of a segment in BYTES (odd number of doubles)
  int xb8[ENG_MAX_CLS];               // byte offset of the class inside xs
  int qoff[ENG_MAX_CLS], hoff[ENG_MAX_CLS];
  uint16_t aptr[ENG_MAX_WARPS + 1], bptr[ENG_MAX_WARPS + 1], sptr[ENG_MAX_WARPS + 1];
  uint32_t task_a[ENG_MAX_TASKS];     // k | blk << 4 | T << 6 | r0 << 8 | nr << 16     (jj0 = 64 * blk)
  uint32_t task_b[ENG_MAX_TASKS];     // k | T << 6 | jj0 << 8 | njj << 16
  uint32_t task_s[ENG_MAX_TASKS];     // first natural segment | count << 16
  uint32_t ll_desc[ENG_MAX_Q];        // per (k, r): start | (# '+') << 16 | (# '-') << 24
  uint32_t hh_desc[ENG_MAX_SEG];      // per class-major segment: the same
  uint16_t ll_ent[ENG_MAX_ENT];       // 8 * r'
  uint16_t hh_ent[ENG_MAX_ENT];       // P8[k] * jj'   (relative to the class)
  uint16_t lhq[ENG_MAX_LH][ENG_MAX_Q];  // per (bond, (k, r)): 8 * r' | dl bit << 10 | parity(dl part) << 11 | valid << 12
  uint16_t dl_of_q[ENG_MAX_Q];        // bit pattern of dl
  uint32_t seg_nat[ENG_MAX_SEG];      // natural order: slot (doubles) | offset in the row << 14 | k << 28
  uint16_t goff_cm[ENG_MAX_SEG];      // class-major segment -> offset in the row
};

// ---- per-lane tables (global memory, read once per task) ----------------------------------------
struct EngLane {
  const uint16_t* dh_cm;     // [nseg]       dh bits of the class-major segment
  const uint32_t* lh_lane;   // [nlh][nseg]  slot of the segment dh ^ bit (doubles; the zero region when its
                             //              class is empty) | dh bit << 14 | parity(dh part) << 15
};

struct EngArgs {
  HubParams hp;
  EngLane ln;
  double e_dn_const;
};

#define ENG_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
typedef uint32_t eng_addr;
__device__ __forceinline__ double eng_ld(eng_addr a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void eng_st(eng_addr a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ int eng_popc(uint32_t v) { return __popc(v); }
__device__ __forceinline__ double eng_flip(double v, uint32_t signmask) {
  return __hiloint2double(__double2hiint(v) ^ (int)signmask, __double2loint(v));
}
template <typename T> __device__ __forceinline__ T eng_ldg(const T* p) { return __ldg(p); }
#else
typedef uintptr_t eng_addr;
inline double eng_ld(eng_addr a) { return *reinterpret_cast<const double*>(a); }
inline void eng_st(eng_addr a, double v) { *reinterpret_cast<double*>(a) = v; }
inline int eng_popc(uint32_t v) { return __builtin_popcount(v); }
inline double eng_flip(double v, uint32_t signmask) {
  unsigned long long b;
  memcpy(&b, &v, 8);
  b ^= (unsigned long long)signmask << 32;
  memcpy(&v, &b, 8);
  return v;
}
template <typename T> inline T eng_ldg(const T* p) { return *p; }
#endif

// ---- phase A: one task = nr consecutive columns r of class k, T blocks of 32 segments (lanes along jj) ----
// ys[(jj, r)] = (diag / hop) * x + sum_LL +- x[(jj, r')] + sum_LH +- x[(jj', r')]
template <int T>
ENG_HD void eng_task_a(const EngConst& C, const EngLane& ln, uint32_t task, eng_addr xs_a, eng_addr ydelta,
                       uint32_t ups, double eu_s, double u0_s, int lane) {
  const int k = (int)(task & 15u), jj0 = (int)((task >> 4) & 3u) * 64;
  const int r0 = (int)((task >> 8) & 255u), nr = (int)((task >> 16) & 255u);
  const int hk = C.H[k], nlh = C.nlh;
  const eng_addr pk8 = (eng_addr)C.P8[k], cb = xs_a + (eng_addr)C.xb8[k];
  const eng_addr zaddr = xs_a + (eng_addr)C.zoff * 8u;
  const int sgb = C.hoff[k], q0 = C.qoff[k];
  eng_addr xa[T];
  double dgh[T];
  bool live[T];
  eng_addr a0[ENG_MAX_LH][T], a1[ENG_MAX_LH][T];
  uint32_t sg[ENG_MAX_LH][T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int jj = jj0 + lane + 32 * t;
    live[t] = jj < hk;
    const int jc = live[t] ? jj : hk - 1;   // lanes past the class recompute its last segment, never store
    xa[t] = cb + (eng_addr)jc * pk8;
    const uint32_t dhb = (uint32_t)eng_ldg(ln.dh_cm + sgb + jc) << C.m;
    dgh[t] = eu_s + u0_s * (double)eng_popc(ups & dhb);
#pragma unroll
    for (int b = 0; b < ENG_MAX_LH; ++b) {
      a0[b][t] = zaddr; a1[b][t] = zaddr; sg[b][t] = 0u;
      if (b < nlh) {
        const uint32_t w = eng_ldg(ln.lh_lane + (size_t)b * C.nseg + sgb + jc);
        const eng_addr src = xs_a + (eng_addr)(w & 0x3fffu) * 8u;
        if (w & 0x4000u) a0[b][t] = src; else a1[b][t] = src;   // dl bit clear needs the dh bit set, and v.v.
        sg[b][t] = (w >> 15) << 31;
      }
    }
  }
#pragma unroll 1
  for (int r = r0; r < r0 + nr; ++r) {
    const int q = q0 + r;
    const uint32_t d = C.ll_desc[q];
    int i = (int)(d & 0xffffu);
    const int ie = i + (int)((d >> 16) & 0xffu), je = ie + (int)(d >> 24);
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
#pragma unroll 4
    for (; i < ie; ++i) {
      const eng_addr e = (eng_addr)C.ll_ent[i];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] += eng_ld(xa[t] + e);
    }
#pragma unroll 4
    for (; i < je; ++i) {
      const eng_addr e = (eng_addr)C.ll_ent[i];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] -= eng_ld(xa[t] + e);
    }
#pragma unroll
    for (int b = 0; b < ENG_MAX_LH; ++b) {
      if (b < nlh) {
        const uint32_t w = C.lhq[b][q];
        if (w & 0x1000u) {
          const eng_addr off = (eng_addr)(w & 0x3ffu);
          const uint32_t sl = (w & 0x800u) << 20;
          if (w & 0x400u) {
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] += eng_flip(eng_ld(a1[b][t] + off), sg[b][t] ^ sl);
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] += eng_flip(eng_ld(a0[b][t] + off), sg[b][t] ^ sl);
          }
        }
      }
    }
    const double dgl = u0_s * (double)eng_popc(ups & (uint32_t)C.dl_of_q[q]);
    const eng_addr r8 = (eng_addr)r * 8u;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const double xv = eng_ld(xa[t] + r8);
      const double y = (dgh[t] + dgl) * xv + acc[t];
      if (live[t]) eng_st(xa[t] + r8 + ydelta, y);
    }
  }
}

ENG_HD void eng_run_a(const EngConst& C, const EngLane& ln, int warp, eng_addr xs_a, eng_addr ydelta,
                      uint32_t ups, double eu_s, double u0_s, int lane) {
  for (int it = C.aptr[warp]; it < C.aptr[warp + 1]; ++it) {
    const uint32_t task = C.task_a[it];
    if (((task >> 6) & 3u) == 1u) eng_task_a<1>(C, ln, task, xs_a, ydelta, ups, eu_s, u0_s, lane);
    else eng_task_a<2>(C, ln, task, xs_a, ydelta, ups, eu_s, u0_s, lane);
  }
}

// row-uniform data of phase B's epilogue
struct EngEpi {
  const double* xr;     // row of x in global memory (up-hop gathers are relative to it)
  double* yr;           // row of y
  double hop0;
  int accumulate;
  int cu;               // up-hop gathers of this row
  const i64* up_off;    // [cu] element offset of the source row relative to xr (shared memory)
  const double* up_coef;
  double s1, s2;        // Lanczos scalars
  bool has_prev;
};

// ---- phase B: one task = njj consecutive segments of class k, T blocks of 32 ranks (lanes along r) ----
// y[goff + r] = hop * (ys[(jj, r)] + sum_HH +- x[(jj', r)]) (+ up-hop row gathers), stored to global memory
template <int T, bool LZ, bool WITH_UP>
ENG_HD void eng_task_b(const EngConst& C, uint32_t task, eng_addr xs_a, eng_addr ydelta, const EngEpi& E,
                       double& dot, int lane) {
  const int k = (int)(task & 15u), jj0 = (int)((task >> 8) & 255u), njj = (int)((task >> 16) & 255u);
  const int sk = C.S[k], sgb = C.hoff[k];
  const eng_addr pk8 = (eng_addr)C.P8[k];
  const eng_addr xl = xs_a + (eng_addr)C.xb8[k] + (eng_addr)lane * 8u;
  bool live[T];
#pragma unroll
  for (int t = 0; t < T; ++t) live[t] = lane + 32 * t < sk;
#pragma unroll 1
  for (int jj = jj0; jj < jj0 + njj; ++jj) {
    const int sgi = sgb + jj;
    const uint32_t d = C.hh_desc[sgi];
    const int goff = (int)C.goff_cm[sgi];
    double up[T];
#pragma unroll
    for (int t = 0; t < T; ++t) up[t] = 0.0;
    double g0[ENG_UPB][T];
    if (WITH_UP) {   // first batch of up-hop gathers: issued here, consumed behind the hop loops
#pragma unroll
      for (int q = 0; q < ENG_UPB; ++q)
#pragma unroll
        for (int t = 0; t < T; ++t)
          g0[q][t] = (q < E.cu && live[t]) ? eng_ldg(E.xr + E.up_off[q] + goff + lane + 32 * t) : 0.0;
    }
    int i = (int)(d & 0xffffu);
    const int ie = i + (int)((d >> 16) & 0xffu), je = ie + (int)(d >> 24);
    double acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
#pragma unroll 4
    for (; i < ie; ++i) {
      const eng_addr e = (eng_addr)C.hh_ent[i];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] += eng_ld(xl + e + 256u * t);
    }
#pragma unroll 4
    for (; i < je; ++i) {
      const eng_addr e = (eng_addr)C.hh_ent[i];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] -= eng_ld(xl + e + 256u * t);
    }
    if (WITH_UP) {
#pragma unroll
      for (int q = 0; q < ENG_UPB; ++q)
        if (q < E.cu) {
          const double c = E.up_coef[q];
#pragma unroll
          for (int t = 0; t < T; ++t) up[t] += c * g0[q][t];
        }
#pragma unroll 1
      for (int qb = ENG_UPB; qb < E.cu; qb += ENG_UPB) {
        double g[ENG_UPB][T];
#pragma unroll
        for (int q = 0; q < ENG_UPB; ++q)
#pragma unroll
          for (int t = 0; t < T; ++t)
            g[q][t] = (qb + q < E.cu && live[t]) ? eng_ldg(E.xr + E.up_off[qb + q] + goff + lane + 32 * t) : 0.0;
#pragma unroll
        for (int q = 0; q < ENG_UPB; ++q)
          if (qb + q < E.cu) {
            const double c = E.up_coef[qb + q];
#pragma unroll
            for (int t = 0; t < T; ++t) up[t] += c * g[q][t];
          }
      }
    }
    const eng_addr own = xl + (eng_addr)jj * pk8;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const double a = E.hop0 * (eng_ld(own + ydelta + 256u * t) + acc[t]) + up[t];
      if (live[t]) {
        double* yp = E.yr + goff + lane + 32 * t;
        if (LZ) {
          double w = E.s1 * a;
          if (E.has_prev) w -= E.s2 * *yp;
          dot += (E.s1 * eng_ld(own + 256u * t)) * w;
          *yp = w;
        } else {
          *yp = E.accumulate ? *yp + a : a;
        }
      }
    }
  }
}

template <bool LZ, bool WITH_UP>
ENG_HD void eng_run_b(const EngConst& C, int warp, eng_addr xs_a, eng_addr ydelta, const EngEpi& E, double& dot,
                      int lane) {
  for (int it = C.bptr[warp]; it < C.bptr[warp + 1]; ++it) {
    const uint32_t task = C.task_b[it];
    const uint32_t T = (task >> 6) & 3u;
    if (T == 1u) eng_task_b<1, LZ, WITH_UP>(C, task, xs_a, ydelta, E, dot, lane);
    else if (T == 2u) eng_task_b<2, LZ, WITH_UP>(C, task, xs_a, ydelta, E, dot, lane);
    else eng_task_b<3, LZ, WITH_UP>(C, task, xs_a, ydelta, E, dot, lane);
  }
}

#ifdef __CUDACC__
__device__ __forceinline__ void eng_cp_async8(uint32_t dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// smem: [xs: xs_elems + ENG_ZREG doubles][ys: xs_elems + ENG_ZREG doubles]
template <bool LZ, bool WITH_UP, int NT>
__global__ void __launch_bounds__(NT, 1) hub_eng_kernel(const __grid_constant__ EngConst C, const EngArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ i64 s_up_off[ELL_MAX_BONDS];
  __shared__ double s_up_coef[ELL_MAX_BONDS];
  const HubParams& p = A.hp;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for ptxas: tables walk the uniform datapath
  double* xs = reinterpret_cast<double*>(smem_raw);
  const int xs_total = C.xs_elems + ENG_ZREG;
  const uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(xs);
  const uint32_t ydelta = (uint32_t)xs_total * 8u;
  for (int i = tid; i < 2 * xs_total; i += NT) xs[i] = 0.0;   // slack slots and the zero region stay 0
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  double dot = 0.0;
  const i64 nd = p.num_dn, nu = p.num_up;
  const double inv_hop = 1.0 / p.hop0;
  const double u0_s = p.u0 * inv_hop;
  EngEpi E;
  E.hop0 = p.hop0; E.accumulate = p.accumulate; E.s1 = s1; E.s2 = s2; E.has_prev = has_prev;
  E.up_off = s_up_off; E.up_coef = s_up_coef; E.cu = 0;
  __syncthreads();
  for (i64 row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    const i64 u = p.row0 + row;
    const double* __restrict__ xr = p.x + row * nd;
    E.xr = xr; E.yr = p.y + row * nd;
    // ---- stage the row: natural order -> class-major padded layout, one segment per warp pass ----
    for (int it = C.sptr[warp]; it < C.sptr[warp + 1]; ++it) {
      const uint32_t ts = C.task_s[it];
      const int sA = (int)(ts & 0xffffu), sB = sA + (int)(ts >> 16);
#pragma unroll 1
      for (int s = sA; s < sB; ++s) {
        const uint32_t w = C.seg_nat[s];
        const uint32_t dst = xs_a + (w & 0x3fffu) * 8u + (uint32_t)lane * 8u;
        const double* src = xr + ((w >> 14) & 0x3fffu) + lane;
        const int sk = C.S[w >> 28];
        if (lane < sk) eng_cp_async8(dst, src);
        if (lane + 32 < sk) eng_cp_async8(dst + 256u, src + 32);
        if (lane + 64 < sk) eng_cp_async8(dst + 512u, src + 64);
      }
    }
    if (WITH_UP) {
      const int cu = p.with_up ? (int)p.cnt_up[u] : 0;
      E.cu = cu;
      for (int q = tid; q < cu; q += NT) {
        const uint32_t e = p.ell_up[(i64)q * nu + u];
        s_up_off[q] = ((i64)(e & ELL_TGT_MASK) - u) * nd;   // relative to the current row
        s_up_coef[q] = (e >> 31) ? -p.hop0 : p.hop0;
      }
    }
    const uint32_t ups = p.up_states[u];
    const double eu_s = (p.e_up[u] + A.e_dn_const) * inv_hop;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    eng_run_a(C, A.ln, warp, xs_a, ydelta, ups, eu_s, u0_s, lane);
    __syncthreads();
    eng_run_b<LZ, WITH_UP>(C, warp, xs_a, ydelta, E, dot, lane);
    __syncthreads();   // xs / ys of this row fully consumed
  }
  lz_finish<LZ>(p.lz, j, dot, red);
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------------
// host: table construction (no CUDA calls: shared with tests/emu/eng_emu.cu)
// ---------------------------------------------------------------------------------------------------
struct EngHost {
  EngConst C;
  std::vector<uint16_t> dh_cm;
  std::vector<uint32_t> lh_lane;
  bool ok = false;
  double e_dn_const = 0.0;
  size_t smem = 0;
  int nwarps = 32;
};

// Longest-processing-time-first assignment of cost-weighted pieces to warps; fills ptr / tasks.
static bool eng_assign(const std::vector<std::pair<double, uint32_t>>& pieces, int nwarps, uint16_t* ptr,
                       uint32_t* tasks, int& ntasks_total, int cap) {
  std::vector<size_t> order(pieces.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](size_t a, size_t b) { return pieces[a].first > pieces[b].first; });
  std::vector<double> load(nwarps, 0.0);
  std::vector<std::vector<uint32_t>> mine(nwarps);
  for (size_t oi : order) {
    int best = 0;
    for (int w = 1; w < nwarps; ++w)
      if (load[w] < load[best]) best = w;
    load[best] += pieces[oi].first;
    mine[best].push_back(pieces[oi].second);
  }
  int n = 0;
  for (int w = 0; w < ENG_MAX_WARPS + 1; ++w) ptr[w] = 0;
  for (int w = 0; w < nwarps; ++w) {
    ptr[w] = (uint16_t)n;
    for (uint32_t t : mine[w]) {
      if (n >= cap) return false;
      tasks[n++] = t;
    }
  }
  for (int w = nwarps; w <= ENG_MAX_WARPS; ++w) ptr[w] = (uint16_t)n;
  ntasks_total = n;
  return true;
}

// ok=false (no error) when the sector is outside what the engine supports.
static int build_eng_host(EngHost& T, int num_sites, int n_dn, i64 num_dn, int nbonds, const int* s1,
                          const int* s2, int sign_width, const double* eps, i64 smem_optin, int nwarps,
                          bool with_up_cost) {
  T.ok = false;
  T.nwarps = nwarps;
  const u64* B = host_binom();
  if (n_dn < 0 || n_dn > num_sites || num_sites < 2 || num_sites > 16) return CMPY_OK;
  if ((i64)B[num_sites * BINOM_N + n_dn] != num_dn) return CMPY_OK;
  if (num_dn >= 16384 - ENG_ZREG || nwarps < 1 || nwarps > ENG_MAX_WARPS) return CMPY_OK;
  const int m = (num_sites + 1) / 2;
  if (m > 8) return CMPY_OK;
  const int hb = num_sites - m;
  const int nlo = 1 << m, nhi = 1 << hb;
  EngConst& C = T.C;
  memset(&C, 0, sizeof(C));
  C.m = m; C.hb = hb; C.n_dn = n_dn; C.nq = nlo; C.row_len = (int)num_dn;
  int xoff = 0, qo = 0, ho = 0;
  for (int k = 0; k <= m; ++k) {
    const int hk = n_dn - k;
    C.S[k] = (int)B[m * BINOM_N + k];
    C.H[k] = (hk >= 0 && hk <= hb) ? (int)B[hb * BINOM_N + hk] : 0;
    const int P = C.S[k] | 1;   // odd pitch >= S_k
    C.P8[k] = 8 * P;
    C.xb8[k] = 8 * xoff; xoff += C.H[k] * P;
    C.qoff[k] = qo; qo += C.S[k];
    C.hoff[k] = ho; ho += C.H[k];
    if (C.H[k] > 0 && (C.S[k] > 96 || C.H[k] > 255 || C.S[k] > 255)) return CMPY_OK;
    if (C.H[k] * P * 8 > 65535 + 8) return CMPY_OK;   // HH entries are u16 byte offsets inside the class
  }
  C.xs_elems = xoff; C.zoff = xoff;
  const int nseg = ho;
  C.nseg = nseg;
  if (xoff + ENG_ZREG >= 16384 || nseg > ENG_MAX_SEG || nseg < 1) return CMPY_OK;
  T.smem = sizeof(double) * 2 * ((size_t)xoff + ENG_ZREG);
  if ((i64)T.smem + 2048 > smem_optin) return CMPY_OK;
  // ranks of dl / dh inside their classes
  std::vector<int> lo_rank(nlo), dl_of_q(nlo), k_of_q(nlo);
  {
    std::vector<int> fill(m + 1, 0);
    for (int v = 0; v < nlo; ++v) {
      const int k = __builtin_popcount(v);
      lo_rank[v] = fill[k];
      dl_of_q[C.qoff[k] + fill[k]] = v;
      k_of_q[C.qoff[k] + fill[k]] = k;
      ++fill[k];
    }
  }
  std::vector<int> hi_k(nhi, -1), hi_goff(nhi, 0), hi_cm(nhi, -1), hi_jj(nhi, 0), dh_cm(nseg, 0);
  {
    std::vector<int> fill(m + 1, 0);
    i64 off = 0;
    int ordinal = 0;
    for (int dh = 0; dh < nhi; ++dh) {
      const int k = n_dn - __builtin_popcount(dh);
      if (k < 0 || k > m) continue;
      hi_k[dh] = k; hi_goff[dh] = (int)off;
      hi_jj[dh] = fill[k];
      hi_cm[dh] = C.hoff[k] + fill[k];
      dh_cm[hi_cm[dh]] = dh;
      const int slot = C.xb8[k] / 8 + fill[k] * (C.P8[k] / 8);
      C.seg_nat[ordinal] = (uint32_t)slot | ((uint32_t)off << 14) | ((uint32_t)k << 28);
      C.goff_cm[hi_cm[dh]] = (uint16_t)off;
      ++fill[k];
      off += C.S[k];
      ++ordinal;
    }
    if (off != num_dn || ordinal != nseg) return cmpy_fail(CMPY_ERR_ARG, "engine tables: size mismatch");
  }
  auto slot_of = [&](int dh) { return C.xb8[hi_k[dh]] / 8 + hi_jj[dh] * (C.P8[hi_k[dh]] / 8); };
  std::vector<int> ll, hh, lh;
  for (int b = 0; b < nbonds; ++b) {
    if (s1[b] >= s2[b]) return CMPY_OK;
    if (s2[b] < m) ll.push_back(b);
    else if (s1[b] >= m) hh.push_back(b);
    else lh.push_back(b);
  }
  C.nlh = (int)lh.size();
  if (C.nlh > ENG_MAX_LH) return CMPY_OK;
  auto parity = [&](u64 state, int a, int b2) {
    return __builtin_popcountll(state & between_mask(a, b2, sign_width)) & 1;
  };
  // hop lists: '+' entries, then '-' entries
  int nle = 0, nhe = 0;
  std::vector<int> n_ll(nlo, 0), n_hh(nseg, 0);
  for (int q = 0; q < nlo; ++q) {
    const int dl = dl_of_q[q];
    std::vector<uint16_t> pos, neg;
    for (int b : ll) {
      const int b1 = (dl >> s1[b]) & 1, b2 = (dl >> s2[b]) & 1;
      if (b1 == b2) continue;
      const int nl = dl ^ (1 << s1[b]) ^ (1 << s2[b]);
      (parity((u64)dl, s1[b], s2[b]) ? neg : pos).push_back((uint16_t)(8 * lo_rank[nl]));
    }
    if (nle + pos.size() + neg.size() > ENG_MAX_ENT || pos.size() > 255 || neg.size() > 255) return CMPY_OK;
    C.ll_desc[q] = (uint32_t)nle | ((uint32_t)pos.size() << 16) | ((uint32_t)neg.size() << 24);
    for (uint16_t v : pos) C.ll_ent[nle++] = v;
    for (uint16_t v : neg) C.ll_ent[nle++] = v;
    n_ll[q] = (int)(pos.size() + neg.size());
  }
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const int dh = dh_cm[sgi], k = hi_k[dh];
    std::vector<uint16_t> pos, neg;
    for (int b : hh) {
      const int a = s1[b] - m, c = s2[b] - m;
      const int b1 = (dh >> a) & 1, b2 = (dh >> c) & 1;
      if (b1 == b2) continue;
      const int nh = dh ^ (1 << a) ^ (1 << c);
      (parity((u64)dh << m, s1[b], s2[b]) ? neg : pos).push_back((uint16_t)(hi_jj[nh] * C.P8[k]));
    }
    if (nhe + pos.size() + neg.size() > ENG_MAX_ENT || pos.size() > 255 || neg.size() > 255) return CMPY_OK;
    C.hh_desc[sgi] = (uint32_t)nhe | ((uint32_t)pos.size() << 16) | ((uint32_t)neg.size() << 24);
    for (uint16_t v : pos) C.hh_ent[nhe++] = v;
    for (uint16_t v : neg) C.hh_ent[nhe++] = v;
    n_hh[sgi] = (int)(pos.size() + neg.size());
  }
  for (int q = 0; q < nlo; ++q) C.dl_of_q[q] = (uint16_t)dl_of_q[q];
  // LH tables: warp-uniform part per (bond, (k, r)), per-lane part per (bond, class-major segment)
  T.dh_cm.assign(nseg, 0);
  for (int sgi = 0; sgi < nseg; ++sgi) T.dh_cm[sgi] = (uint16_t)dh_cm[sgi];
  T.lh_lane.assign((size_t)std::max(1, C.nlh) * nseg, (uint32_t)C.zoff);
  std::vector<int> n_lh(nlo, 0);
  for (int qb = 0; qb < C.nlh; ++qb) {
    const int b = lh[qb];
    const int a = s1[b], c = s2[b] - m;   // a inside dl, c inside dh
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const int dh = dh_cm[sgi];
      const int nh = dh ^ (1 << c);
      const uint32_t bit = (uint32_t)((dh >> c) & 1);
      const uint32_t par = (uint32_t)parity((u64)dh << m, m - 1, s2[b]);   // bits of dh strictly below c
      const uint32_t src = (hi_k[nh] >= 0) ? (uint32_t)slot_of(nh) : (uint32_t)C.zoff;
      T.lh_lane[(size_t)qb * nseg + sgi] = src | (bit << 14) | (par << 15);
    }
    for (int q = 0; q < nlo; ++q) {
      const int dl = dl_of_q[q], k = k_of_q[q];
      const int bit_lo = (dl >> a) & 1;
      const int kp = k + (bit_lo ? -1 : 1);   // class of the source segment
      uint16_t e = 0;
      if (kp >= 0 && kp <= m && C.H[kp] > 0 && C.H[k] > 0) {
        const int nl = dl ^ (1 << a);
        const int par = parity((u64)dl, a, m);   // bits of dl strictly above a
        e = (uint16_t)((8 * lo_rank[nl]) | (bit_lo << 10) | (par << 11) | (1 << 12));
        n_lh[q] += 1;
      }
      C.lhq[qb][q] = e;
    }
  }
  // ---- tasks: cost-balanced pieces, longest first.  Costs = issued instructions of the compiled loops ----
  {
    std::vector<std::pair<double, uint32_t>> pa, pb, ps;
    double tot_a = 0.0, tot_b = 0.0;
    auto cost_a = [&](int k, int r, int Tt) {
      const int q = C.qoff[k] + r;
      return 6.0 + Tt * (6.0 + 2.0 * n_ll[q] + 3.0 * n_lh[q]) + n_ll[q] + 2.0 * C.nlh;
    };
    auto cost_b = [&](int k, int jj) {
      const int Tt = (C.S[k] + 31) / 32;
      const int sgi = C.hoff[k] + jj;
      return 8.0 + Tt * (7.0 + 2.0 * n_hh[sgi] + (with_up_cost ? 40.0 : 0.0)) + n_hh[sgi];
    };
    for (int k = 0; k <= m; ++k) {
      if (C.H[k] <= 0) continue;
      for (int blk = 0; blk * 64 < C.H[k]; ++blk) {
        const int Tt = (std::min(C.H[k] - blk * 64, 64) + 31) / 32;
        for (int r = 0; r < C.S[k]; ++r) tot_a += cost_a(k, r, Tt);
      }
      for (int jj = 0; jj < C.H[k]; ++jj) tot_b += cost_b(k, jj);
    }
    const double tgt_a = tot_a / (3.0 * nwarps) + 40.0, tgt_b = tot_b / (3.0 * nwarps) + 20.0;
    for (int k = 0; k <= m; ++k) {
      if (C.H[k] <= 0) continue;
      for (int blk = 0; blk * 64 < C.H[k]; ++blk) {
        const int Tt = (std::min(C.H[k] - blk * 64, 64) + 31) / 32;
        int r0 = 0;
        double acc = 40.0 * Tt;   // per-task set-up (hoisted per-lane state)
        for (int r = 0; r < C.S[k]; ++r) {
          acc += cost_a(k, r, Tt);
          if (acc >= tgt_a || r + 1 == C.S[k]) {
            pa.push_back({acc, (uint32_t)k | ((uint32_t)blk << 4) | ((uint32_t)Tt << 6) | ((uint32_t)r0 << 8) |
                                   ((uint32_t)(r + 1 - r0) << 16)});
            r0 = r + 1; acc = 40.0 * Tt;
          }
        }
      }
      {
        const int Tt = (C.S[k] + 31) / 32;
        int j0 = 0;
        double acc = 12.0;
        for (int jj = 0; jj < C.H[k]; ++jj) {
          acc += cost_b(k, jj);
          if (acc >= tgt_b || jj + 1 == C.H[k]) {
            pb.push_back({acc, (uint32_t)k | ((uint32_t)Tt << 6) | ((uint32_t)j0 << 8) | ((uint32_t)(jj + 1 - j0) << 16)});
            j0 = jj + 1; acc = 12.0;
          }
        }
      }
    }
    {  // staging: runs of natural segments, cost = 32-lane passes
      double tot = 0.0;
      for (int s = 0; s < nseg; ++s) tot += 3.0 + (C.S[C.seg_nat[s] >> 28] + 31) / 32;
      const double tgt = tot / (2.0 * nwarps) + 4.0;
      int s0 = 0;
      double acc = 0.0;
      for (int s = 0; s < nseg; ++s) {
        acc += 3.0 + (C.S[C.seg_nat[s] >> 28] + 31) / 32;
        if (acc >= tgt || s + 1 == nseg) {
          ps.push_back({acc, (uint32_t)s0 | ((uint32_t)(s + 1 - s0) << 16)});
          s0 = s + 1; acc = 0.0;
        }
      }
    }
    int na = 0, nb = 0, ns = 0;
    if (!eng_assign(pa, nwarps, C.aptr, C.task_a, na, ENG_MAX_TASKS)) return CMPY_OK;
    if (!eng_assign(pb, nwarps, C.bptr, C.task_b, nb, ENG_MAX_TASKS)) return CMPY_OK;
    if (!eng_assign(ps, nwarps, C.sptr, C.task_s, ns, ENG_MAX_TASKS)) return CMPY_OK;
  }
  // energies: eps uniform -> eps * n_dn summed like weighted_element (ascending adds)
  { double v = 0; for (int i = 0; i < n_dn; ++i) v += eps[0]; T.e_dn_const = v; }
  T.ok = true;
  return CMPY_OK;
}
