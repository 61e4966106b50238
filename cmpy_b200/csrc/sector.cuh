// sector.cuh -- K1 (enumerate / rank), K2 (hop tables), K3 (energies).
// ref: cmpy/basis.py:655-666 (generate_states), cmpy/operators.py:226-299,425-460.
#pragma once
#include "common.cuh"

// ---- K1 --------------------------------------------------------------------------
__global__ void sector_enumerate_kernel(int num_sites, int n, i64 count, i64* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < count;
       i += (i64)gridDim.x * blockDim.x)
    out[i] = (i64)colex_unrank(i, n, num_sites);
}

__global__ void sector_rank_kernel(const i64* __restrict__ states, i64 m, i64* __restrict__ idx) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m;
       i += (i64)gridDim.x * blockDim.x)
    idx[i] = colex_rank((u64)states[i]);
}

static inline int grid_for(i64 n, int threads, int cap = 148 * 16) {
  i64 g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// ---- K2 --------------------------------------------------------------------------
// mask of the sites strictly between site1 and site2, limited to bits < width
__host__ __device__ __forceinline__ u64 between_mask(int site1, int site2, int width) {
  u64 m = 0;
  for (int i = site1 + 1; i < site2; ++i)
    if (i < width) m |= (1ull << i);
  return m;
}

__global__ void species_hops_kernel(const i64* __restrict__ states, i64 num, int fixed_popcount,
                                    int width, int site1, int site2,
                                    int32_t* __restrict__ target, int8_t* __restrict__ sign) {
  const u64 b1 = 1ull << site1, b2 = 1ull << site2;
  const u64 bm = between_mask(site1, site2, width);
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < num;
       i += (i64)gridDim.x * blockDim.x) {
    u64 s = (u64)states[i];
    bool o1 = (s & b1) != 0, o2 = (s & b2) != 0;
    int32_t t = -1;
    if (o1 != o2) {
      u64 ns = s ^ b1 ^ b2;
      t = (int32_t)(fixed_popcount ? colex_rank(ns) : bsearch_left(states, num, (i64)ns));
    }
    target[i] = t;
    sign[i] = (__popcll(s & bm) & 1) ? -1 : 1;
  }
}

// Packed ELL entry of the H.v tables: [31] sign (1 = negative), [30:25] bond id,
// [24:0] target string index.
#define ELL_TGT_BITS 25
#define ELL_TGT_MASK ((1u << ELL_TGT_BITS) - 1u)
#define ELL_MAX_BONDS 64

struct BondList {
  int n;
  int s1[ELL_MAX_BONDS];
  int s2[ELL_MAX_BONDS];
};

// cnt[i] = number of hoppable bonds of string i; ell[k*num + i] = k-th packed entry.
__global__ void species_ell_kernel(const i64* __restrict__ states, i64 num, int fixed_popcount,
                                   int width, BondList bonds, int ell_width,
                                   uint32_t* __restrict__ ell, uint8_t* __restrict__ cnt) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < num;
       i += (i64)gridDim.x * blockDim.x) {
    u64 s = (u64)states[i];
    int c = 0;
    for (int b = 0; b < bonds.n; ++b) {
      const u64 b1 = 1ull << bonds.s1[b], b2 = 1ull << bonds.s2[b];
      bool o1 = (s & b1) != 0, o2 = (s & b2) != 0;
      if (o1 == o2) continue;
      u64 ns = s ^ b1 ^ b2;
      i64 t = fixed_popcount ? colex_rank(ns) : bsearch_left(states, num, (i64)ns);
      u64 bm = between_mask(bonds.s1[b], bonds.s2[b], width);
      uint32_t neg = (uint32_t)(__popcll(s & bm) & 1);
      if (c < ell_width)
        ell[(i64)c * num + i] = (neg << 31) | ((uint32_t)b << ELL_TGT_BITS) | (uint32_t)t;
      ++c;
    }
    cnt[i] = (uint8_t)c;
  }
}

// ---- K3 --------------------------------------------------------------------------
struct SiteValues {
  int n;
  double v[64];
};

__host__ __device__ __forceinline__ double weighted_element_dev(u64 state, const SiteValues& sv) {
  double value = 0.0;
  for (int i = 0; i < sv.n; ++i)
    if (state & (1ull << i)) value += sv.v[i];  // ascending site order, plain adds
  return value;
}

__global__ void weighted_elements_kernel(const i64* __restrict__ states, i64 num, SiteValues sv,
                                         double* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < num;
       i += (i64)gridDim.x * blockDim.x)
    out[i] = weighted_element_dev((u64)states[i], sv);
}

__global__ void inter_elements_kernel(const i64* __restrict__ up, i64 num_up,
                                      const i64* __restrict__ dn, i64 num_dn, SiteValues sv,
                                      double* __restrict__ out) {
  const i64 total = num_up * num_dn;
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < total;
       i += (i64)gridDim.x * blockDim.x) {
    i64 a = i / num_dn, b = i - a * num_dn;
    out[i] = weighted_element_dev((u64)(up[a] & dn[b]), sv);
  }
}
