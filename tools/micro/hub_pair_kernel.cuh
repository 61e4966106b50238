// EXPERIMENT RECORD (not built, not shipped): the full single-GPU H.v as ONE kernel with two roles per CTA --
// threads 0..511 the row engine (diagonal + dn hops through shared memory), threads 512..1023 the up hops of the
// same row as streaming 16-byte row gathers, a named-barrier rendezvous per row (y of the row is stored by the
// gather role, then updated by the engine's phase B).  Correct (the whole H.v / Lanczos GPU parity suite passed with
// it as the default), but 4.37 ms per H.v on the 4x4 sector against 4.00 ms for hub_seg_kernel (16-site chain 3.21
// vs 2.56): 512 threads x 8 x 16 bytes = 64 KB of gathers in flight per SM is half of what the stand-alone gather
// kernel needs for its 2.1 ms, and the register file (2 x 512 x 64) leaves no room for more.  See DESIGN.md 5.3.
// This is the body of cmpy_b200/csrc/hubbard_eng.cuh between eng_cp_async8 and the host part at the time
// (git history: "fused pair kernel").
// CTA-wide barrier of the engine's threads: the whole CTA, or (PAIR) named barrier 1 over the NT engine threads of
// the fused pair kernel
template <int NT, bool PAIR>
__device__ __forceinline__ void eng_sync() {
  if (PAIR) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
}
#define ENG_PAIR_THREADS 1024   // engine role + gather role of hub_pair_kernel
// rendezvous of the two roles of hub_pair_kernel: the up-hop gathers of the row are stored in y
__device__ __forceinline__ void eng_pair_rendezvous() { asm volatile("bar.sync 2, %0;" ::"n"(ENG_PAIR_THREADS) : "memory"); }

// smem: [xs: xs_elems + ENG_ZREG doubles][ys: xs_elems + ENG_ZREG doubles] behind the tables
// The rows of one CTA, NT threads (tid in [0, NT)).  PAIR: the engine role of hub_pair_kernel -- y already holds
// the up hops (and the caller's accumulate / Lanczos terms) of the row when the rendezvous before phase B returns;
// phase B adds this role's part.
template <bool LZ, bool WITH_UP, int NT, int NLH, bool PAIR>
__device__ __forceinline__ void eng_rows(const EngConst& C, const EngArgs& A, unsigned char* smem_rows, int tid,
                                         double s1, double s2, bool has_prev, double& dot) {
  __shared__ double s_dg[ENG_MAX_Q + ENG_MAX_SEG];
  __shared__ i64 s_up_off[WITH_UP ? ELL_MAX_BONDS : 1];
  __shared__ double s_up_coef[WITH_UP ? ELL_MAX_BONDS : 1];
  const HubParams& p = A.hp;
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for ptxas: tables walk the uniform datapath
  double* xs = reinterpret_cast<double*>(smem_rows);
  const int xs_total = C.xs_elems + ENG_ZREG;
  const uint32_t xs_a = (uint32_t)__cvta_generic_to_shared(xs);
  const uint32_t dg_a = (uint32_t)__cvta_generic_to_shared(s_dg);
  const uint32_t ydelta = (uint32_t)xs_total * 8u;
  for (int i = tid; i < 2 * xs_total; i += NT) xs[i] = 0.0;   // slack slots and the zero region stay 0
  const i64 nd = p.num_dn, nu = p.num_up;
  const double inv_hop = 1.0 / p.hop0;
  const double u0_s = p.u0 * inv_hop;
  EngEpi E;
  E.hop0 = p.hop0; E.accumulate = p.accumulate; E.s1 = s1; E.s2 = s2; E.has_prev = has_prev;
  E.c1 = 1.0; E.c2 = 0.0;
  if (p.acc_scale) { E.accumulate = 2; E.c1 = p.acc_scale[0]; E.c2 = p.acc_scale[1]; }
  if (PAIR) { E.accumulate = 2; E.c2 = 1.0; E.lz_acc = true; }   // y = c1 * (this part) + y;  LZ: w = y + s1 * a
  E.up_off = s_up_off; E.up_coef = s_up_coef; E.cu = 0;
  const uint32_t dst_l = xs_a + (uint32_t)lane * 8u;
  eng_sync<NT, PAIR>();
  for (i64 row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    const i64 u = p.row0 + row;
    const double* __restrict__ xr = p.x + row * nd;
    E.xr = xr; E.yr = p.y + row * nd;
    const double* src_l = xr + lane;
    // ---- stage the row: natural order -> class-major padded layout, one segment per warp pass ----
    for (int it = C.sptr[warp]; it < C.sptr[warp + 1]; ++it) {
      const uint32_t ts = C.task_s[it];
      const int k = (int)(ts & 15u), jA = (int)((ts >> 8) & 255u), jB = jA + (int)((ts >> 16) & 255u);
      const int sk = C.S[k], sgb = C.hoff[k];
      const uint32_t pk8 = (uint32_t)C.P8[k], dst_k = dst_l + (uint32_t)C.xb8[k];
      const bool l0 = lane < sk, l1 = lane + 32 < sk, l2 = lane + 64 < sk;
#pragma unroll 1
      for (int jj = jA; jj < jB; ++jj) {
        const uint32_t dst = dst_k + (uint32_t)jj * pk8;
        const double* src = src_l + (int)C.goff_cm[sgb + jj];
        if (l0) eng_cp_async8(dst, src);
        if (l1) eng_cp_async8(dst + 256u, src + 32);
        if (l2) eng_cp_async8(dst + 512u, src + 64);
      }
    }
    if (row + gridDim.x < p.nrows) {   // pull the next row of this CTA into L2 while this one is processed
      const char* nxt = reinterpret_cast<const char*>(xr + (i64)gridDim.x * nd);
      for (int b = tid * 128; b < (int)(nd * 8); b += NT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + b));
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
    eng_fill_diag(C, A.ln, s_dg, tid, p.up_states[u], (p.e_up[u] + A.e_dn_const) * inv_hop, u0_s);
    asm volatile("cp.async.wait_all;" ::: "memory");
    eng_sync<NT, PAIR>();
    eng_run_a<NLH>(C, A.ln, warp, xs_a, ydelta, dg_a, lane);
    if (PAIR) eng_pair_rendezvous();   // also orders phase A before phase B among the engine threads
    else eng_sync<NT, PAIR>();
    eng_run_b<LZ, WITH_UP>(C, warp, xs_a, ydelta, E, dot, lane);
    eng_sync<NT, PAIR>();   // xs / ys / s_dg of this row fully consumed
  }
}

// TAB_SMEM: the tables are copied to shared memory once per CTA and walked with (broadcast) LDS instead
// of constant loads -- the constant caches of an SM hold ~5 KB and the 4x4 lattice needs more (see EngConst)
template <bool LZ, bool WITH_UP, int NT, int NLH, bool TAB_SMEM = true>
__global__ void __launch_bounds__(NT, 1) hub_eng_kernel(const __grid_constant__ EngConst Cc, const EngArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TAB_BYTES = TAB_SMEM ? (int)((sizeof(EngConst) + 15) & ~(size_t)15) : 0;
  if (TAB_SMEM) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&Cc);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(EngConst) / 4); i += NT) dst[i] = src[i];
    __syncthreads();
  }
  const EngConst& C = TAB_SMEM ? *reinterpret_cast<const EngConst*>(smem_raw) : Cc;
  __shared__ double red[32];
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(A.hp.lz, j, s1, s2, has_prev);
  double dot = 0.0;
  eng_rows<LZ, WITH_UP, NT, NLH, false>(C, A, smem_raw + TAB_BYTES, threadIdx.x, s1, s2, has_prev, dot);
  lz_finish<LZ>(A.hp.lz, j, dot, red);
}

// ---------------------------------------------------------------------------------------------------
// K4 "pair": the full single-GPU H.v as ONE kernel with two roles per CTA (round 2).
//   threads   0..511  : the row engine above (diagonal + dn hops through shared memory), 16 warps
//   threads 512..1023 : the up hops of the same row as streaming row gathers, y[u, :] = sum_k +-hop x[u'_k, :]
//                       (8 x 16-byte loads per thread in flight), stored to y with the caller's accumulate /
//                       Lanczos terms; then the rendezvous (named barrier over all 1024 threads); the engine
//                       role adds its part to y in its phase B while the gather role starts on the next row.
// The up hops are bound by the L2 -> SM path and need bytes in flight, the engine by the shared-memory pipe: as
// ONE instruction stream (the engine's WITH_UP phase B, 4 x 8 bytes per thread in flight) they take 5.1 ms per H.v
// on the 4x4 lattice, as hub_seg_kernel 4.0 ms.  Per row the two roles take about the same time, so the lock step
// costs little.  Requirements (host): whole-vector operator, even num_dn (16-byte rows), uniform hop.
// Reference semantics: cmpy/operators.py:463-527 (project_hopping: up hops at stride num_dn).
#define UPG_G 8     // gathers of 16 bytes in flight per thread
template <bool LZ>
__device__ __forceinline__ void upg_rows(const HubParams& p, int tid, double s1, double s2, bool has_prev) {
  constexpr int NT = ENG_PAIR_THREADS / 2;
  __shared__ i64 s_off[ELL_MAX_BONDS];
  __shared__ double s_coef[ELL_MAX_BONDS];
  // value stored for the row (G = gathered sum): plain y = G; accumulate y = y + G; LZ y = s1 G - s2 y
  double c1 = 1.0, c2 = p.accumulate ? 1.0 : 0.0;
  if (LZ) { c1 = s1; c2 = has_prev ? -s2 : 0.0; }
  const bool rmw = LZ ? has_prev : (p.accumulate != 0);
  const i64 nd = p.num_dn, nu = p.num_up;
  const int npair = (int)(nd >> 1);   // nd even (host)
  for (i64 row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    const i64 u = p.row0 + row;
    const int cu = (int)p.cnt_up[u];
    for (int q = tid; q < cu; q += NT) {
      const uint32_t e = p.ell_up[(i64)q * nu + u];
      s_off[q] = (i64)(e & ELL_TGT_MASK) * nd;
      s_coef[q] = (e >> 31) ? -p.hop0 : p.hop0;
    }
    asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory");   // table of this row complete (gather threads only)
    double* __restrict__ yr = p.y + row * nd;
    for (int i = tid; i < npair; i += NT) {
      const double* __restrict__ xc = p.x + 2 * i;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
      for (int q0 = 0; q0 < cu; q0 += UPG_G) {
        double2 g[UPG_G];
#pragma unroll
        for (int q = 0; q < UPG_G; ++q)
          g[q] = q0 + q < cu ? __ldg(reinterpret_cast<const double2*>(xc + s_off[q0 + q])) : make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < UPG_G; ++q)
          if (q0 + q < cu) { const double c = s_coef[q0 + q]; a0 += c * g[q].x; a1 += c * g[q].y; }
      }
      double2* yp = reinterpret_cast<double2*>(yr + 2 * i);
      double2 w = make_double2(c1 * a0, c1 * a1);
      if (rmw) { const double2 o = *yp; w.x += c2 * o.x; w.y += c2 * o.y; }
      *yp = w;
    }
    eng_pair_rendezvous();   // the row of y is stored (also: nobody reads s_off / s_coef of this row any more)
  }
}

template <bool LZ, int NLH>
__global__ void __launch_bounds__(ENG_PAIR_THREADS, 1) hub_pair_kernel(const __grid_constant__ EngConst Cc, const EngArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TAB_BYTES = (int)((sizeof(EngConst) + 15) & ~(size_t)15);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&Cc);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(EngConst) / 4); i += ENG_PAIR_THREADS) dst[i] = src[i];
    __syncthreads();
  }
  const EngConst& C = *reinterpret_cast<const EngConst*>(smem_raw);
  __shared__ double red[32];
  int j; double s1, s2; bool has_prev;
  lz_scalars<LZ>(A.hp.lz, j, s1, s2, has_prev);
  double dot = 0.0;
  if (threadIdx.x < ENG_PAIR_THREADS / 2)
    eng_rows<LZ, false, ENG_PAIR_THREADS / 2, NLH, true>(C, A, smem_raw + TAB_BYTES, threadIdx.x, s1, s2, has_prev, dot);
  else
    upg_rows<LZ>(A.hp, threadIdx.x - ENG_PAIR_THREADS / 2, s1, s2, has_prev);
  lz_finish<LZ>(A.hp.lz, j, dot, red);
}
#endif  // __CUDACC__
