// heisenberg.cuh -- K5: matrix-free fp64 H.v for the Heisenberg / XXZ model on a fixed
// magnetisation sector (or the full 2^N space).
// ref: HeisenbergModel._hamiltonian_data cmpy/models/heisenberg.py:19-40 -- every directed
// neighbor pair contributes sign*0.25*jz to the diagonal and, for anti-parallel bits,
// 0.125*j to the flipped state.
//
// Two-level ("Lin table") layout: state = (hi, lo), lo = low `lo_bits` bits.
//   index(state) = off[hi] + lo_rank[lo]       (ascending integer order == reference order)
// so the vector is a ragged matrix: row `hi` holds the C(lo_bits, n_up - popc(hi)) states
// sharing the high bits.  Bonds inside the low bits are gathers inside one row (staged in
// shared memory), bonds inside the high bits are coalesced reads of whole remote rows,
// bonds straddling the split are per-element gathers.
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "sector.cuh"
#include "hubbard.cuh"
#include "hubbard_cls.cuh"

#define HEIS_MAX_BONDS 64

struct HeisBonds {
  int n_lo, n_hi, n_mix;
  uint32_t lo_mask[HEIS_MAX_BONDS];   // both bits in lo
  double lo_w[HEIS_MAX_BONDS], lo_dz[HEIS_MAX_BONDS];
  uint32_t hi_mask[HEIS_MAX_BONDS];   // both bits in hi (shifted down by lo_bits)
  double hi_w[HEIS_MAX_BONDS], hi_dz[HEIS_MAX_BONDS];
  uint32_t mix_lo[HEIS_MAX_BONDS], mix_hi[HEIS_MAX_BONDS];  // single-bit masks
  double mix_w[HEIS_MAX_BONDS], mix_dz[HEIS_MAX_BONDS];
};

struct HeisParams {
  int pt_filter;            // >= 0: only rows whose hi part has this popcount (rest: class-major path)
  int lo_bits, n_up, all_states;
  i64 nhi;
  const i64* off;           // [nhi + 1]
  const uint32_t* lo_list;  // class-major list of lo strings
  const i64* cls_off;       // [lo_bits + 2]
  const uint16_t* lo_rank;  // [2^lo_bits]
  const HeisBonds* bonds;   // device copy
  const double* x; double* y;
  LzCtx lz;
};

__device__ __forceinline__ int heis_cls(const HeisParams& p, uint32_t hi) {
  return p.all_states ? 0 : p.n_up - __popc(hi);
}

template <bool LZ>
__global__ void heis_row_kernel(HeisParams p) {
  extern __shared__ __align__(16) double xs[];
  __shared__ double red[32];
  __shared__ HeisBonds sb;
  __shared__ i64 s_hi_base[HEIS_MAX_BONDS];
  __shared__ double s_hi_w[HEIS_MAX_BONDS];
  __shared__ int s_hi_cnt;
  __shared__ double s_hi_diag;
  {
    const int nwords = sizeof(HeisBonds) / 4;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.bonds);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sb);
    for (int k = threadIdx.x; k < nwords; k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(p.lz, j, s1, s2, has_prev);
  const int tid = threadIdx.x, nt = blockDim.x;
  double dot = 0.0;
  for (i64 hi64 = blockIdx.x; hi64 < p.nhi; hi64 += gridDim.x) {
    const uint32_t hi = (uint32_t)hi64;
    const i64 base = p.off[hi64];
    const i64 len = p.off[hi64 + 1] - base;
    if (len == 0 || (p.pt_filter >= 0 && __popc(hi) != p.pt_filter)) continue;
    const i64 cbase = p.cls_off[heis_cls(p, hi)];
    const double* __restrict__ xr = p.x + base;
    for (i64 r = tid; r < len; r += nt) xs[r] = xr[r];
    if (tid == 0) {
      int c = 0;
      double dz = 0.0;
      for (int b = 0; b < sb.n_hi; ++b) {
        const uint32_t m = sb.hi_mask[b];
        const uint32_t v = hi & m;
        const bool differ = (v != 0) && (v != m);
        dz += differ ? -sb.hi_dz[b] : sb.hi_dz[b];
        if (differ) { s_hi_base[c] = p.off[hi ^ m]; s_hi_w[c] = sb.hi_w[b]; ++c; }
      }
      s_hi_cnt = c; s_hi_diag = dz;
    }
    __syncthreads();
    const int chi = s_hi_cnt;
    const double dhi = s_hi_diag;
    for (i64 r = tid; r < len; r += nt) {
      const uint32_t lo = p.lo_list[cbase + r];
      const double xi = xs[r];
      double dg = dhi;
      double acc = 0.0;
      for (int b = 0; b < sb.n_lo; ++b) {
        const uint32_t m = sb.lo_mask[b];
        const uint32_t v = lo & m;
        const bool differ = (v != 0) && (v != m);
        dg += differ ? -sb.lo_dz[b] : sb.lo_dz[b];
        if (differ) acc += sb.lo_w[b] * xs[p.lo_rank[lo ^ m]];
      }
      for (int k = 0; k < chi; ++k) acc += s_hi_w[k] * p.x[s_hi_base[k] + r];
      for (int b = 0; b < sb.n_mix; ++b) {
        const bool bl = (lo & sb.mix_lo[b]) != 0, bh = (hi & sb.mix_hi[b]) != 0;
        const bool differ = bl != bh;
        dg += differ ? -sb.mix_dz[b] : sb.mix_dz[b];
        if (differ) {
          const i64 t = p.off[hi ^ sb.mix_hi[b]] + (i64)p.lo_rank[lo ^ sb.mix_lo[b]];
          acc += sb.mix_w[b] * p.x[t];
        }
      }
      acc += dg * xi;
      const i64 i = base + r;
      if (LZ) {
        double w = s1 * acc;
        if (has_prev) w -= s2 * p.y[i];
        p.y[i] = w;
        dot += (s1 * xi) * w;
      } else {
        p.y[i] = acc;
      }
    }
    __syncthreads();
  }
  lz_finish<LZ>(p.lz, j, dot, red);
}

__global__ void heis_diag_kernel(HeisParams p, double* __restrict__ diag) {
  const HeisBonds& sb = *p.bonds;
  for (i64 hi64 = blockIdx.x; hi64 < p.nhi; hi64 += gridDim.x) {
    const uint32_t hi = (uint32_t)hi64;
    const i64 base = p.off[hi64];
    const i64 len = p.off[hi64 + 1] - base;
    if (len == 0) continue;
    const i64 cbase = p.cls_off[heis_cls(p, hi)];
    for (i64 r = threadIdx.x; r < len; r += blockDim.x) {
      const uint32_t lo = p.lo_list[cbase + r];
      double dg = 0.0;
      for (int b = 0; b < sb.n_hi; ++b) {
        const uint32_t m = sb.hi_mask[b], v = hi & m;
        dg += ((v != 0) && (v != m)) ? -sb.hi_dz[b] : sb.hi_dz[b];
      }
      for (int b = 0; b < sb.n_lo; ++b) {
        const uint32_t m = sb.lo_mask[b], v = lo & m;
        dg += ((v != 0) && (v != m)) ? -sb.lo_dz[b] : sb.lo_dz[b];
      }
      for (int b = 0; b < sb.n_mix; ++b) {
        const bool bl = (lo & sb.mix_lo[b]) != 0, bh = (hi & sb.mix_hi[b]) != 0;
        dg += (bl != bh) ? -sb.mix_dz[b] : sb.mix_dz[b];
      }
      diag[base + r] = dg;
    }
  }
}

struct HeisenbergOp : cmpy_op_s {
  int num_sites = 0, n_up = 0, lo_bits = 0, hi_bits = 0, all_states = 0;
  i64 nhi = 0;
  i64 max_len = 0;
  i64* d_off = nullptr; uint32_t* d_lo_list = nullptr; i64* d_cls_off = nullptr;
  uint16_t* d_lo_rank = nullptr; HeisBonds* d_bonds = nullptr;
  int threads = 256, blocks_per_sm = 1;
  // fast path for more than 16 sites: sub-row launches of the class-major kernel (hubbard_cls.cuh)
  LongTables lng;
  std::vector<int> slow_pt;   // popcounts of the high part left to heis_row_kernel
  SpinDiag sd;
  double w_hop = 0.0;
  uint32_t* d_zero_u32 = nullptr;
  double* d_zero_f64 = nullptr;
  bool fast_ok = false;

  ~HeisenbergOp() override {
    cudaFree(d_off); cudaFree(d_lo_list); cudaFree(d_cls_off); cudaFree(d_lo_rank); cudaFree(d_bonds);
    cudaFree(d_zero_u32); cudaFree(d_zero_f64);
    lng.release();
  }

  HeisParams params() const {
    HeisParams p;
    p.pt_filter = -1;
    p.lo_bits = lo_bits; p.n_up = n_up; p.all_states = all_states; p.nhi = nhi;
    p.off = d_off; p.lo_list = d_lo_list; p.cls_off = d_cls_off; p.lo_rank = d_lo_rank;
    p.bonds = d_bonds; p.x = nullptr; p.y = nullptr;
    p.lz.enabled = 0; p.lz.iter = nullptr; p.lz.beta = nullptr; p.lz.alpha = nullptr;
    p.lz.partials = d_partials; p.lz.ticket = d_ticket;
    return p;
  }

  int build(int N, int nup, int npairs, const int32_t* pairs, double jj, double jz) {
    num_sites = N; all_states = nup < 0; n_up = all_states ? 0 : nup;
    ARG_CHECK(N >= 1 && N <= 32, "heisenberg: 1 <= num_sites <= 32");
    ARG_CHECK(all_states || nup <= N, "heisenberg: n_up out of range");
    lo_bits = N / 2; if (lo_bits < 1) lo_bits = 1; if (lo_bits > 16) lo_bits = 16;
    if (N > LONG_RBITS && !all_states) lo_bits = LONG_RBITS;  // same split as the class-major fast path
    if (lo_bits > N) lo_bits = N;
    hi_bits = N - lo_bits;
    ARG_CHECK(hi_bits <= 16, "heisenberg: internal split");
    nhi = 1ll << hi_bits;
    const u64* B = host_binom();
    const int nlo = 1 << lo_bits;
    // lo tables
    std::vector<i64> cls_off(lo_bits + 2, 0);
    std::vector<uint32_t> lo_list(nlo);
    std::vector<uint16_t> lo_rank(nlo);
    if (all_states) {
      for (int v = 0; v < nlo; ++v) { lo_list[v] = v; lo_rank[v] = (uint16_t)v; }
    } else {
      for (int k = 0; k <= lo_bits; ++k) cls_off[k + 1] = cls_off[k] + (i64)B[lo_bits * BINOM_N + k];
      std::vector<i64> fill(lo_bits + 1, 0);
      for (int v = 0; v < nlo; ++v) {
        int k = __builtin_popcount(v);
        lo_rank[v] = (uint16_t)fill[k];
        lo_list[cls_off[k] + fill[k]] = v;
        ++fill[k];
      }
    }
    std::vector<i64> off(nhi + 1, 0);
    max_len = 0;
    for (i64 h = 0; h < nhi; ++h) {
      i64 len;
      if (all_states) len = nlo;
      else {
        int k = n_up - __builtin_popcount((unsigned)h);
        len = (k >= 0 && k <= lo_bits) ? (i64)B[lo_bits * BINOM_N + k] : 0;
      }
      off[h + 1] = off[h] + len;
      if (len > max_len) max_len = len;
    }
    size = off[nhi];
    ARG_CHECK(size >= 1, "heisenberg: empty sector");
    // bonds: merge directed pairs into undirected bonds with multiplicity
    std::vector<int> bi, bj, mult;
    for (int q = 0; q < npairs; ++q) {
      int a = pairs[2 * q], b = pairs[2 * q + 1];
      ARG_CHECK(a >= 0 && a < N && b >= 0 && b < N && a != b, "heisenberg: bad neighbor pair");
      int lo = a < b ? a : b, hi = a < b ? b : a;
      size_t k = 0;
      for (; k < bi.size(); ++k) if (bi[k] == lo && bj[k] == hi) break;
      if (k == bi.size()) { bi.push_back(lo); bj.push_back(hi); mult.push_back(0); }
      ++mult[k];
    }
    HeisBonds hb;
    memset(&hb, 0, sizeof(hb));
    for (size_t k = 0; k < bi.size(); ++k) {
      const double w = mult[k] * (0.25 * jj / 2), dz = mult[k] * (0.25 * jz);
      const int a = bi[k], b = bj[k];
      if (b < lo_bits) {
        ARG_CHECK(hb.n_lo < HEIS_MAX_BONDS, "too many bonds");
        hb.lo_mask[hb.n_lo] = (1u << a) | (1u << b); hb.lo_w[hb.n_lo] = w; hb.lo_dz[hb.n_lo] = dz; ++hb.n_lo;
      } else if (a >= lo_bits) {
        ARG_CHECK(hb.n_hi < HEIS_MAX_BONDS, "too many bonds");
        hb.hi_mask[hb.n_hi] = (1u << (a - lo_bits)) | (1u << (b - lo_bits));
        hb.hi_w[hb.n_hi] = w; hb.hi_dz[hb.n_hi] = dz; ++hb.n_hi;
      } else {
        ARG_CHECK(hb.n_mix < HEIS_MAX_BONDS, "too many bonds");
        hb.mix_lo[hb.n_mix] = 1u << a; hb.mix_hi[hb.n_mix] = 1u << (b - lo_bits);
        hb.mix_w[hb.n_mix] = w; hb.mix_dz[hb.n_mix] = dz; ++hb.n_mix;
      }
    }
    CU_CHECK(cudaMalloc(&d_off, sizeof(i64) * (nhi + 1)));
    CU_CHECK(cudaMemcpy(d_off, off.data(), sizeof(i64) * (nhi + 1), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&d_lo_list, sizeof(uint32_t) * nlo));
    CU_CHECK(cudaMemcpy(d_lo_list, lo_list.data(), sizeof(uint32_t) * nlo, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&d_lo_rank, sizeof(uint16_t) * nlo));
    CU_CHECK(cudaMemcpy(d_lo_rank, lo_rank.data(), sizeof(uint16_t) * nlo, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&d_cls_off, sizeof(i64) * (lo_bits + 2)));
    CU_CHECK(cudaMemcpy(d_cls_off, cls_off.data(), sizeof(i64) * (lo_bits + 2), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&d_bonds, sizeof(HeisBonds)));
    CU_CHECK(cudaMemcpy(d_bonds, &hb, sizeof(HeisBonds), cudaMemcpyHostToDevice));
    int rcf = configure_fast(bi, bj, mult, jj, jz);
    if (rcf) return rcf;
    // launch config
    size_t smem = sizeof(double) * (size_t)max_len;
    i64 t = ((max_len + 3) / 4 + 31) / 32 * 32;
    if (t < 64) t = 64; if (t > 512) t = 512;
    threads = (int)t;
    int rcs = raise_smem_limit(heis_row_kernel<false>, smem_optin);
    if (!rcs) rcs = raise_smem_limit(heis_row_kernel<true>, smem_optin);
    if (rcs) return rcs;
    int nb = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, heis_row_kernel<true>, threads, smem));
    ARG_CHECK(nb >= 1, "heisenberg: row does not fit shared memory");
    blocks_per_sm = nb > 8 ? 8 : nb;
    return CMPY_OK;
  }

  // Class-major fast path: uniform bond weights, fixed magnetisation, more than 16 sites.
  int configure_fast(const std::vector<int>& bi, const std::vector<int>& bj, const std::vector<int>& mult,
                     double jj, double jz) {
    fast_ok = false;
    if (all_states || num_sites <= LONG_RBITS || bi.empty()) return CMPY_OK;
    for (size_t k = 1; k < mult.size(); ++k) if (mult[k] != mult[0]) return CMPY_OK;
    w_hop = mult[0] * (0.25 * jj / 2);
    const double dz = mult[0] * (0.25 * jz);
    memset(&sd, 0, sizeof(sd));
    for (size_t k = 0; k < bi.size(); ++k) {
      const int delta = bj[k] - bi[k];
      int i = 0;
      for (; i < sd.ndelta; ++i) if (sd.delta[i] == delta) break;
      if (i == sd.ndelta) {
        if (sd.ndelta == 4) return CMPY_OK;
        sd.delta[sd.ndelta++] = delta;
      }
      sd.dmask[i] |= 1u << bi[k];
    }
    sd.e0 = dz * (double)bi.size();
    sd.escale = -2.0 * dz;
    const double eps0[1] = {0.0};
    slow_pt.clear();
    int rc = build_long_tables(lng, num_sites, n_up, size, (int)bi.size(), bi.data(), bj.data(), 0, eps0,
                               smem_optin, &slow_pt);
    if (rc) return rc;
    if (!lng.ok) return CMPY_OK;
    CU_CHECK(cudaMalloc(&d_zero_u32, 16)); CU_CHECK(cudaMemset(d_zero_u32, 0, 16));
    CU_CHECK(cudaMalloc(&d_zero_f64, 16)); CU_CHECK(cudaMemset(d_zero_f64, 0, 16));
    rc = raise_smem_limit(hub_cls_kernel<false, 1024, 8, true, true>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<true, 1024, 8, true, true>, smem_optin);
    if (rc) return rc;
    fast_ok = true;
    return CMPY_OK;
  }

  int apply_fast(const double* x, double* y, const LzCtx& lz, cudaStream_t st) {
    HubParams hp;
    memset(&hp, 0, sizeof(hp));
    hp.num_up = 1; hp.num_dn = size; hp.up_states = d_zero_u32; hp.e_up = d_zero_f64;
    hp.num_sites = num_sites; hp.u0 = 0.0; hp.hop0 = w_hop;
    hp.row0 = 0; hp.nrows = 1; hp.with_up = 0; hp.accumulate = 0;
    hp.x = x; hp.y = y;
    hp.lz.enabled = 0; hp.lz.partials = d_partials; hp.lz.ticket = d_ticket;
    int launches = 0;
    for (auto& S : lng.sets) {
      ClsParams cp;
      cp.hp = hp; cp.lay = S.cls.lay; cp.blob = S.cls.d_blob; cp.pair_seg = S.shift ? S.cls.d_pair_seg1 : S.cls.d_pair_seg;
      cp.e_dn_const = 0.0; cp.sd = sd;
      cp.lg.ntop = S.ntop; cp.lg.row_len = S.row_len; cp.lg.nsb = lng.nsb; cp.lg.shift = S.shift;
      cp.lg.top_val = S.d_top_val; cp.lg.sub_off = S.d_sub_off; cp.lg.tb_ptr = S.d_tb_ptr;
      cp.lg.tb_ent = S.d_tb_ent; cp.lg.sb_src = S.d_sb_src; cp.lg.sb_map = S.d_sb_map;
      i64 g = sm_count;
      if (g > S.ntop) g = S.ntop;
      if (lz.enabled) {
        cp.hp.lz = lz; cp.hp.lz.partials = d_partials; cp.hp.lz.ticket = d_ticket;
        cp.hp.lz.enabled = launches == 0 ? 1 : 2;
        hub_cls_kernel<true, 1024, 8, true, true><<<(int)g, 1024, S.cls.smem, st>>>(cp);
      } else {
        hub_cls_kernel<false, 1024, 8, true, true><<<(int)g, 1024, S.cls.smem, st>>>(cp);
      }
      KERNEL_CHECK();
      ++launches;
    }
    // the classes the class-major kernel cannot take (sub-rows of odd length): generic row kernel
    for (int pt : slow_pt) {
      HeisParams p = params();
      p.x = x; p.y = y; p.pt_filter = pt;
      size_t smem = sizeof(double) * (size_t)max_len;
      i64 g = (i64)sm_count * blocks_per_sm;
      if (g > nhi) g = nhi;
      if (lz.enabled) {
        p.lz = lz; p.lz.partials = d_partials; p.lz.ticket = d_ticket;
        p.lz.enabled = launches == 0 ? 1 : 2;
        heis_row_kernel<true><<<(int)g, threads, smem, st>>>(p);
      } else {
        heis_row_kernel<false><<<(int)g, threads, smem, st>>>(p);
      }
      KERNEL_CHECK();
      ++launches;
    }
    return CMPY_OK;
  }

  int apply(const double* x, double* y, const LzCtx& lz, cudaStream_t st) override {
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (variant == 5 && !(fast_ok && aligned16))
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "class-major variant not available for this spin sector");
    if (fast_ok && aligned16 && variant != 1) return apply_fast(x, y, lz, st);
    HeisParams p = params();
    p.x = x; p.y = y;
    size_t smem = sizeof(double) * (size_t)max_len;
    i64 g = (i64)sm_count * blocks_per_sm;
    if (g > nhi) g = nhi;
    if (lz.enabled) {
      p.lz = lz; p.lz.partials = d_partials; p.lz.ticket = d_ticket;
      heis_row_kernel<true><<<(int)g, threads, smem, st>>>(p);
    } else {
      heis_row_kernel<false><<<(int)g, threads, smem, st>>>(p);
    }
    KERNEL_CHECK();
    return CMPY_OK;
  }

  int diagonal(double* d_diag, cudaStream_t st) override {
    HeisParams p = params();
    i64 g = (i64)sm_count * 4;
    if (g > nhi) g = nhi;
    heis_diag_kernel<<<(int)g, 256, 0, st>>>(p, d_diag);
    KERNEL_CHECK();
    return CMPY_OK;
  }
};
