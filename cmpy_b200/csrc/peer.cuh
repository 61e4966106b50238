// peer.cuh -- K9: slab transposes over NVLink peer memory (up-string-sharded H.v).
//
// One process per GPU; every rank owns a slab of up-rows of X (row-major nrows x num_dn) and,
// in the transposed phase, a slab of dn-columns XT (ncols x num_up).  The XT / YT slabs live in
// symmetric memory (torch.distributed._symmetric_memory: CUDA VMM handles exchanged between the
// processes of one box), so a kernel on one GPU can address the slabs of every peer directly:
//   push : X[r, c] (local read, coalesced along c)  ->  XT_q[c - cb[q], row0 + r]  (remote stores,
//          256-byte pieces along r) for the owner q of column c
//   pull : y[r, c] += YT_q[c - cb[q], row0 + r]     (remote loads along r, local RMW along c)
// This replaces pack kernels + NCCL all-to-all + unpack kernels (and their two staging slabs) by
// one kernel per direction whose NVLink traffic overlaps its own local reads tile by tile.
// The reference has no distributed path; layout: cmpy/operators.py:33-90.
#pragma once
#include "common.cuh"

#define PEER_MAX 16

struct PeerTable {
  double* base[PEER_MAX];   // XT / YT slab of every rank (peer-mapped device pointers)
  i64 cb[PEER_MAX + 1];     // column bounds: rank q owns columns [cb[q], cb[q+1])
  int world;
};

__device__ __forceinline__ int peer_owner(const PeerTable& t, i64 c) {
  int q = 0;
#pragma unroll 1
  while (q + 1 < t.world && c >= t.cb[q + 1]) ++q;
  return q;
}

// PULL = false: push X -> XT_q ; PULL = true: y += YT_q^T
template <bool PULL>
__global__ void __launch_bounds__(256) peer_transpose_kernel(double* __restrict__ loc, i64 nrows,
                                                             i64 nd, i64 row0, i64 ld_t,
                                                             PeerTable pt) {
  __shared__ double tile[32][33];
  const i64 tiles_c = (nd + 31) / 32, tiles_r = (nrows + 31) / 32;
  const i64 ntiles = tiles_c * tiles_r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // column-tile fastest: consecutive CTAs stream consecutive pieces of the same local rows
    const i64 tr = t / tiles_c, tc = t - tr * tiles_c;
    const i64 r0 = tr * 32, c0 = tc * 32;
    __syncthreads();
    if (!PULL) {
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        const i64 r = r0 + ty + k, c = c0 + tx;
        if (r < nrows && c < nd) tile[ty + k][tx] = loc[r * nd + c];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        const i64 c = c0 + ty + k, r = r0 + tx;
        if (r < nrows && c < nd) {
          const int q = peer_owner(pt, c);
          pt.base[q][(c - pt.cb[q]) * ld_t + row0 + r] = tile[tx][ty + k];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        const i64 c = c0 + ty + k, r = r0 + tx;
        if (r < nrows && c < nd) {
          const int q = peer_owner(pt, c);
          tile[tx][ty + k] = pt.base[q][(c - pt.cb[q]) * ld_t + row0 + r];
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 32; k += 8) {
        const i64 r = r0 + ty + k, c = c0 + tx;
        if (r < nrows && c < nd) loc[r * nd + c] += tile[ty + k][tx];
      }
    }
  }
}
