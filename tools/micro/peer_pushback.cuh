// EXPERIMENT RECORD (not built, not shipped): push-style way back of the sharded H.v -- the owner of a dn-major
// slab adds its transposed tiles into the peers' row-major result slabs with remote red.add.f64 (result slabs
// registered as peer-mapped), beside the up pass of the next column chunk.  Bit-identical to the pull; measured
// on 2 and 8 x B200 (profiles/r2_pushback_ab.txt): at par with the pull (N = 8: 0.83 vs 0.85 ms per H.v; N = 2:
// 2.65-2.9 vs 2.61-2.68), so the pull stayed.  Depends on PeerTable of cmpy_b200/csrc/peer.cuh; the host side
// (cmpy_dist_register_slab, way_back_push with a high-priority stream for the chunked up pass) is in the git
// history ("push-style way back").
// Push-style way back (round 2): the owner of a dn-major slab adds its transposed tiles into the peers'
// row-major result slabs with remote reductions,
//   y_p[u - rb[p], cb[me] + j] += sc * YT_me[j, u]      for the local columns j in [j0, j1), all up-rows u,
// (p = owner of row u).  Every element of y receives exactly one such addition per H.v (on top of the value
// its own rank's dn pass stored before the barrier), so the result does not depend on arrival order.
// Why not a pull: remote READS are capped per SM (~3.2-4 GB/s per SM, LDG and TMA alike: 450-470 GB/s need all
// 148 SMs), posted remote reductions are not -- red.add.f64 reaches 520-530 GB/s from 16 SMs
// (profiles/r2_peer_red_microbench.txt) -- so this kernel runs beside the up pass of the next column chunk.
// Tile = 32 up-rows x 128 columns: local reads are runs of 256 B along u, remote reductions runs of 1 KB along c.
// `rows`: base[p] = y slab of rank p, cb[] = ROW bounds rb[] (up-rows of rank p).
#define PBK_TC 128                      // column granularity the host sizes grids with
#define PBK_ELEMS 4096                  // amplitudes per tile (TU x TC)
#define PBK_CTAS_PER_SM 3               // two tiles (66 KB) per CTA
// The local reads of tile i + 1 (8-byte cp.async straight into shared memory, no registers) are in flight while
// tile i is pushed (first version, load phase and push phase one after the other with 4 loads per thread in
// flight: 15 GB/s per SM).  TU x TC = 32 x 128 (local runs of 256 B, remote runs of 1 KB) or 64 x 64 (512 B both).
template <int TU, int TC>
__global__ void __launch_bounds__(256) peer_pushback_kernel(const double* __restrict__ yt_loc, i64 ld_t, i64 nd, i64 col0,
                                                            i64 j0, i64 j1, PeerTable rows, int me,
                                                            const double* __restrict__ scale) {
  static_assert(TU * TC == PBK_ELEMS && TU % 32 == 0 && TC % 32 == 0, "tile shape");
  constexpr int PITCH = TU + 1, TILE = TC * PITCH;
  extern __shared__ __align__(16) double pbk_tiles[];   // 2 x TILE, [column][row]
  __shared__ int s_pref[PEER_MAX + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int p = 0; p < rows.world; ++p) { s_pref[p] = acc; acc += (int)((rows.cb[p + 1] - rows.cb[p] + TU - 1) / TU); }
    s_pref[rows.world] = acc;
  }
  __syncthreads();
  const double sc = scale ? scale[0] : 1.0;
  const int tiles_u = s_pref[rows.world];
  const int tiles_c = (int)((j1 - j0 + TC - 1) / TC);
  const int ntiles = tiles_u * tiles_c;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8

  // tile t -> target rank p, first up-row u0 (global), rows nu, first local column jt, columns nc.
  // up-tile fastest: at any time the CTAs work on all targets (both NVLink directions and the local part busy),
  // and consecutive tiles read consecutive pieces of the same local columns
  auto decode = [&](int t, int& p, i64& u0, int& nu, i64& jt, int& nc) {
    const int tc = t / tiles_u, tu = t - tc * tiles_u;
    p = 0;
    while (p + 1 < rows.world && tu >= s_pref[p + 1]) ++p;
    u0 = rows.cb[p] + (i64)(tu - s_pref[p]) * TU;
    nu = (int)(rows.cb[p + 1] - u0 < TU ? rows.cb[p + 1] - u0 : TU);
    jt = j0 + (i64)tc * TC;
    nc = (int)(j1 - jt < TC ? j1 - jt : TC);
  };
  auto fetch = [&](int t, double* tile) {     // one column per warp pass, lanes along the up-rows
    int p, nu, nc; i64 u0, jt;
    decode(t, p, u0, nu, jt, nc);
    const double* src = yt_loc + jt * ld_t + u0 + tx;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + tx);
#pragma unroll
    for (int c = ty; c < TC; c += 8)
#pragma unroll
      for (int ui = 0; ui < TU / 32; ++ui)
        if (c < nc && ui * 32 + tx < nu)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + (uint32_t)(c * PITCH + ui * 32) * 8u), "l"(src + (i64)c * ld_t + ui * 32) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int t = blockIdx.x, buf = 0;
  if (t < ntiles) fetch(t, pbk_tiles);
  for (; t < ntiles; t += gridDim.x, buf ^= 1) {
    const double* tile = pbk_tiles + buf * TILE;
    if (t + (int)gridDim.x < ntiles) {
      fetch(t + (int)gridDim.x, pbk_tiles + (buf ^ 1) * TILE);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int p, nu, nc; i64 u0, jt;
    decode(t, p, u0, nu, jt, nc);
    double* dst = rows.base[p] + (u0 - rows.cb[p]) * nd + col0 + jt + tx;
    if (p == me) {
      // own rows: plain read-modify-write, all loads issued before the stores (nobody else touches these
      // elements; a local red.f64 issues at ~1 element per clock per SM)
      double old[TU / 8][TC / 32];
#pragma unroll
      for (int ri = 0; ri < TU / 8; ++ri)
#pragma unroll
        for (int i = 0; i < TC / 32; ++i) {
          const int r = ty + 8 * ri;
          old[ri][i] = (r < nu && i * 32 + tx < nc) ? dst[(i64)r * nd + i * 32] : 0.0;
        }
#pragma unroll
      for (int ri = 0; ri < TU / 8; ++ri)
#pragma unroll
        for (int i = 0; i < TC / 32; ++i) {
          const int r = ty + 8 * ri;
          if (r < nu && i * 32 + tx < nc)
            dst[(i64)r * nd + i * 32] = __dadd_rn(old[ri][i], __dmul_rn(sc, tile[(i * 32 + tx) * PITCH + r]));
        }
    } else {
#pragma unroll
      for (int ri = 0; ri < TU / 8; ++ri) {       // one up-row per warp pass, lanes along the columns
        const int r = ty + 8 * ri;
        if (r < nu) {
          double v[TC / 32];
#pragma unroll
          for (int i = 0; i < TC / 32; ++i) v[i] = i * 32 + tx < nc ? __dmul_rn(sc, tile[(i * 32 + tx) * PITCH + r]) : 0.0;
#pragma unroll
          for (int i = 0; i < TC / 32; ++i)
            if (i * 32 + tx < nc)   // posted remote reduction
              asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(dst + (i64)r * nd + i * 32), "d"(v[i]) : "memory");
        }
      }
    }
    __syncthreads();    // the tile is free for the fetch after next
  }
}

// launches the push-accumulate of local columns [j0, j1) on at most max_sms SMs (0: all)
template <int TU, int TC>
static inline cudaError_t peer_launch_pushback_t(const double* yt_loc, i64 ld_t, i64 nd, i64 col0, i64 j0, i64 j1,
                                                 const PeerTable& rows, int me, const double* scale, int max_sms,
                                                 int sm_count, cudaStream_t st) {
  static bool raised = false;
  const int smem = 2 * TC * (TU + 1) * (int)sizeof(double);
  if (!raised) {
    cudaError_t e = cudaFuncSetAttribute(peer_pushback_kernel<TU, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    raised = true;
  }
  const int sms = (max_sms > 0 && max_sms < sm_count) ? max_sms : sm_count;
  const i64 ntiles = ((ld_t + TU - 1) / TU + rows.world) * ((j1 - j0 + TC - 1) / TC);
  const i64 cap = (i64)PBK_CTAS_PER_SM * sms;
  peer_pushback_kernel<TU, TC><<<(int)(ntiles < cap ? ntiles : cap), 256, smem, st>>>(yt_loc, ld_t, nd, col0, j0, j1, rows, me, scale);
  return cudaGetLastError();
}
static inline cudaError_t peer_launch_pushback(const double* yt_loc, i64 ld_t, i64 nd, i64 col0, i64 j0, i64 j1,
                                               const PeerTable& rows, int me, const double* scale, int max_sms,
                                               int sm_count, cudaStream_t st) {
  if (j1 <= j0 || ld_t == 0) return cudaSuccess;
  static int shape = -1;
  if (shape < 0) { const char* e = getenv("CMPY_PBK_SHAPE"); shape = e ? atoi(e) : 0; }
  if (shape == 1) return peer_launch_pushback_t<64, 64>(yt_loc, ld_t, nd, col0, j0, j1, rows, me, scale, max_sms, sm_count, st);
  return peer_launch_pushback_t<32, 128>(yt_loc, ld_t, nd, col0, j0, j1, rows, me, scale, max_sms, sm_count, st);
}
