// peer.cuh -- K9: slab transposes over NVLink peer memory (up-string-sharded H.v).
//
// One process per GPU; every rank owns a slab of up-rows of X (row-major nrows x num_dn) and,
// in the transposed phase, a slab of dn-columns XT (ncols x num_up).  The XT / YT slabs live in
// symmetric memory (torch.distributed._symmetric_memory: CUDA VMM handles exchanged between the
// processes of one box), so a kernel on one GPU can address the slabs of every peer directly:
//   push : X[r, c] (local read, coalesced along c)  ->  XT_q[c - cb[q], row0 + r]  (remote stores,
//          1 KB runs along r) for the owner q of column c
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
// Tile = PEER_TR rows of the local slab x 32 columns: the remote side of a tile is 32 runs of
// PEER_TR contiguous doubles (1 KB), the local side PEER_TR runs of 256 bytes.
// Measured on 2 x B200 (C4 slabs): push 505 -> 560 GB/s with 128-row tiles, pull 465 -> 395 GB/s, so
// the pull keeps 32-row tiles.
template <bool PULL, int PEER_TR>
__global__ void __launch_bounds__(256) peer_transpose_kernel(double* __restrict__ loc, i64 nrows,
                                                             i64 nd, i64 row0, i64 ld_t,
                                                             PeerTable pt, const double* __restrict__ scale = nullptr) {
  const double sc = (PULL && scale) ? scale[0] : 1.0;   // pull: y += sc * YT^T (device scalar of the sharded Lanczos)
  __shared__ double tile[32][PEER_TR + 1];  // [column][row]
  const i64 tiles_c = (nd + 31) / 32, tiles_r = (nrows + PEER_TR - 1) / PEER_TR;
  const i64 ntiles = tiles_c * tiles_r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // column-tile fastest: consecutive CTAs stream consecutive pieces of the same local rows
    const i64 tr = t / tiles_c, tc = t - tr * tiles_c;
    const i64 r0 = tr * PEER_TR, c0 = tc * 32;
    __syncthreads();
    if (!PULL) {
#pragma unroll 4
      for (int k = ty; k < PEER_TR; k += 8) {      // local rows, lanes along the columns
        const i64 r = r0 + k, c = c0 + tx;
        if (r < nrows && c < nd) tile[tx][k] = loc[r * nd + c];
      }
      __syncthreads();
#pragma unroll
      for (int k = ty; k < 32; k += 8) {           // one column per warp pass, lanes along the rows
        const i64 c = c0 + k;
        if (c < nd) {
          const int q = peer_owner(pt, c);
          double* dst = pt.base[q] + (c - pt.cb[q]) * ld_t + row0 + r0;
#pragma unroll
          for (int j = tx; j < PEER_TR; j += 32)
            if (r0 + j < nrows) dst[j] = tile[k][j];
        }
      }
    } else {
#pragma unroll
      for (int k = ty; k < 32; k += 8) {
        const i64 c = c0 + k;
        if (c < nd) {
          const int q = peer_owner(pt, c);
          const double* src = pt.base[q] + (c - pt.cb[q]) * ld_t + row0 + r0;
#pragma unroll
          for (int j = tx; j < PEER_TR; j += 32)
            if (r0 + j < nrows) tile[k][j] = src[j];
        }
      }
      __syncthreads();
#pragma unroll 4
      for (int k = ty; k < PEER_TR; k += 8) {
        const i64 r = r0 + k, c = c0 + tx;
        if (r < nrows && c < nd) loc[r * nd + c] += sc * tile[tx][k];
      }
    }
  }
}
