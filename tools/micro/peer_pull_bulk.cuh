// EXPERIMENT RECORD (not built, not shipped): the pull-accumulate of the sharded H.v fed by TMA bulk copies,
// and the column-chunked ("pipelined") second half it was written for.  Measured on 2 x B200, 4x4 sector
// (profiles/r2_pull_pipeline_ab.txt): slower than the LDG pull at full grid (2.72 vs 2.61 ms per H.v) and,
// like the LDG pull, proportional to the number of SMs it runs on (~3.2-4 GB/s per SM): the limit is the
// outstanding remote reads one SM may hold, whichever unit issues them.  Depends on PeerTable /
// peer_chunk_range of an intermediate version of cmpy_b200/csrc/peer.cuh (git history: "pipelined pull").
// Pull-accumulate fed by the TMA unit (round 2).  The LDG pull above is bound by the outstanding
// remote loads one SM can hold: its bandwidth is proportional to the SMs it runs on (measured: ~3.2 GB/s
// per SM, 450-470 GB/s on 148), so it cannot run beside the up pass on a few reserved SMs.  Here the remote
// side of a tile -- 32 runs of PB_ROWS contiguous doubles, one per column -- is fetched by 1-D bulk
// copies (cp.async.bulk, mbarrier complete_tx) into an NS-stage ring of shared-memory tiles, issued by the
// lanes of warp 0, one run each; all 256 threads then do the local read-modify-write of a landed tile.
// Runs start at an EVEN global row (16-byte alignment of the copies: ld_t even, slabs 16-byte aligned --
// checked by the host, `peer_bulk_ok`); rows outside [row0, row0 + nrows) are fetched but not used.
#define PB_ROWS 64
#define PB_PITCH 66   // doubles per column of a tile: 528 B, a multiple of 16

__device__ __forceinline__ uint32_t pb_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
               :: "r"(pb_saddr(bar)), "r"(parity) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(256) peer_pull_bulk_kernel(double* __restrict__ loc, i64 nrows, i64 nd, i64 row0,
                                                             i64 ld_t, PeerTable pt, const double* __restrict__ scale,
                                                             int ck, int cK) {
  extern __shared__ __align__(128) unsigned char pb_raw[];
  double* const ring = reinterpret_cast<double*>(pb_raw);     // NS tiles of 32 x PB_PITCH doubles
  __shared__ uint64_t full[NS];
  __shared__ int s_first[PEER_MAX], s_pref[PEER_MAX + 1];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  if (tid == 0) {
    int acc = 0;
    for (int q = 0; q < pt.world; ++q) {
      i64 lo, hi;
      peer_chunk_range(pt, q, ck, cK, lo, hi);
      s_pref[q] = acc;
      s_first[q] = (int)(lo / 32);
      acc += hi > lo ? (int)((hi - 1) / 32 - lo / 32 + 1) : 0;
    }
    s_pref[pt.world] = acc;
    for (int s = 0; s < NS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(pb_saddr(&full[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double sc = scale ? scale[0] : 1.0;
  const i64 g_lo = row0 & ~(i64)1, g_end = (row0 + nrows + 1) & ~(i64)1;   // even bounds of the fetched rows
  const int tiles_c = s_pref[pt.world];
  const int ntiles = tiles_c * (int)((g_end - g_lo + PB_ROWS - 1) / PB_ROWS);
  const int mine = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // tile i of this CTA -> owner q, first column c0, active columns [lo, hi) relative to c0, first global row
  auto decode = [&](int i, int& q, int& c0, int& lo, int& hi, i64& gr0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    const int tr = t / tiles_c, j = t - tr * tiles_c;
    q = 0;
    while (q + 1 < pt.world && j >= s_pref[q + 1]) ++q;
    i64 lo64, hi64;
    peer_chunk_range(pt, q, ck, cK, lo64, hi64);
    c0 = (s_first[q] + (j - s_pref[q])) * 32;
    lo = (int)lo64 - c0; hi = (int)hi64 - c0;
    if (lo < 0) lo = 0;
    if (hi > 32) hi = 32;
    gr0 = g_lo + (i64)tr * PB_ROWS;
  };
  auto issue = [&](int i) {          // warp 0: lane k fetches the run of column c0 + k
    int q, c0, lo, hi; i64 gr0;
    decode(i, q, c0, lo, hi, gr0);
    const int s = i % NS;
    const uint32_t bytes = (uint32_t)(g_end - gr0 < PB_ROWS ? g_end - gr0 : PB_ROWS) * 8u;
    if (tx == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(pb_saddr(&full[s])), "r"(bytes * (uint32_t)(hi - lo)) : "memory");
    __syncwarp();
    if (tx >= lo && tx < hi) {
      const double* src = pt.base[q] + ((i64)(c0 + tx) - pt.cb[q]) * ld_t + gr0;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(pb_saddr(ring + ((size_t)s * 32 + tx) * PB_PITCH)), "l"(src), "r"(bytes), "r"(pb_saddr(&full[s])) : "memory");
    }
  };
  if (ty == 0)
    for (int i = 0; i < NS && i < mine; ++i) issue(i);
  for (int i = 0; i < mine; ++i) {
    int q, c0, lo, hi; i64 gr0;
    decode(i, q, c0, lo, hi, gr0);
    const int s = i % NS;
    pb_wait(&full[s], (uint32_t)((i / NS) & 1));
    const double* tl = ring + (size_t)s * 32 * PB_PITCH;
    const i64 rl0 = gr0 - row0;                      // local row of tile row 0 (may be -1)
    if (tx >= lo && tx < hi) {
      double* const lt = loc + c0 + tx;
#pragma unroll
      for (int k = ty; k < PB_ROWS; k += 8) {
        const i64 r = rl0 + k;
        if (r >= 0 && r < nrows) lt[r * nd] += sc * tl[tx * PB_PITCH + k];
      }
    }
    __syncthreads();                                 // every thread is done with stage s
    if (ty == 0 && i + NS < mine) issue(i + NS);
  }
}

// true when the bulk-copy pull may be used on this table
static inline bool peer_bulk_ok(const PeerTable& pt, i64 ld_t) {
  if (ld_t & 1) return false;
  for (int q = 0; q < pt.world; ++q)
    if (reinterpret_cast<uintptr_t>(pt.base[q]) & 15) return false;
  return true;
}

#define PB_STAGES 4
#define PB_CTAS_PER_SM 3   // 4 stages x 16.5 KB = 66 KB of shared memory per CTA

// y[r, c] += sc * YT_q[c - cb[q], row0 + r] over column chunk ck of cK, on at most max_sms SMs (0: all):
// the bulk-copy kernel when the slabs allow it, else the LDG kernel (odd num_up).
static inline cudaError_t peer_launch_pull(double* y, i64 nrows, i64 nd, i64 row0, i64 ld_t, const PeerTable& pt,
                                           const double* scale, int ck, int cK, int max_sms, int sm_count,
                                           cudaStream_t st) {
  if (nrows == 0 || nd == 0) return cudaSuccess;
  const int sms = (max_sms > 0 && max_sms < sm_count) ? max_sms : sm_count;
  const i64 tiles_c = (nd / cK + 31) / 32 + pt.world + 1;
  if (peer_bulk_ok(pt, ld_t)) {
    static bool raised = false;
    const int smem = PB_STAGES * 32 * PB_PITCH * (int)sizeof(double);
    if (!raised) {
      cudaError_t e = cudaFuncSetAttribute(peer_pull_bulk_kernel<PB_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      raised = true;
    }
    const i64 ntiles = ((nrows + 1 + PB_ROWS - 1) / PB_ROWS + 1) * tiles_c;
    const i64 cap = (i64)PB_CTAS_PER_SM * sms;
    peer_pull_bulk_kernel<PB_STAGES><<<(int)(ntiles < cap ? ntiles : cap), 256, smem, st>>>(y, nrows, nd, row0, ld_t, pt, scale, ck, cK);
  } else {
    const i64 ntiles = ((nrows + 63) / 64) * tiles_c;
    const i64 cap = (i64)8 * sms;
    peer_transpose_kernel<true, 64><<<(int)(ntiles < cap ? ntiles : cap), 256, 0, st>>>(y, nrows, nd, row0, ld_t, pt, scale, ck, cK);
  }
  return cudaGetLastError();
}
