// hubbard_op.cuh -- host-side operator object for the Hubbard / Anderson H.v (tables,
// kernel selection, launches).
#pragma once
#include <stdlib.h>
#include "hubbard.cuh"
#include "hubbard_seg.cuh"
#include "hubbard_cls.cuh"
#include "hubbard_eng.cuh"

// ---------------------------------------------------------------------------------
struct SpeciesTables {
  i64 num = 0;
  int width = 0;
  i64* d_states = nullptr;
  uint32_t* d_states32 = nullptr;
  uint32_t* d_ell = nullptr;
  uint8_t* d_cnt = nullptr;
  double* d_energy = nullptr;
  void release() {
    cudaFree(d_states); cudaFree(d_states32); cudaFree(d_ell); cudaFree(d_cnt); cudaFree(d_energy);
    d_states = nullptr; d_states32 = nullptr; d_ell = nullptr; d_cnt = nullptr; d_energy = nullptr;
  }
};

__global__ void narrow_states_kernel(const i64* __restrict__ in, i64 n, uint32_t* __restrict__ out) {
  for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    out[i] = (uint32_t)in[i];
}

// device side of the generation-3 row engine: the constant-bank tables travel as a kernel
// parameter, only the per-lane tables live in global memory
struct EngTables {
  EngHost H;
  uint16_t* d_dh_cm = nullptr;
  uint16_t* d_dl_q = nullptr;
  uint32_t* d_lh_lane = nullptr;
  bool ok = false;
  void release() {
    cudaFree(d_dh_cm); cudaFree(d_dl_q); cudaFree(d_lh_lane);
    d_dh_cm = nullptr; d_dl_q = nullptr; d_lh_lane = nullptr; ok = false;
  }
};

struct HubbardOp : cmpy_op_s {
  int num_sites = 0, nbonds = 0, sign_width = 0;
  SpeciesTables up, dn;
  double* d_hop = nullptr;
  double* d_u = nullptr;
  bool uniform = false;
  double u0 = 0.0, hop0 = 0.0;
  bool eps_uniform = false;
  int row_threads = 256;
  int row_blocks_per_sm = 1;
  bool row_ok = false;
  SegTables seg;
  int seg_threads = 256;
  int seg_blocks_per_sm = 1;
  bool seg_wide = false;     // 1024-thread CTAs (one CTA per SM, long rows)
  LongTables lng;            // rows of more than 16 sites: sub-row launches of the class-major kernel
  ClsTables cls;             // class-major two-phase kernel (uniform models, long rows)
  bool cls_default = false;  // variant 0 picks it
  int grid_limit = 0;        // > 0: cap on the CTAs of the persistent row kernels (leaves SMs to a concurrent kernel)
  int cls_shape = 0;         // 0: 1024 threads x 8 up-hop loads in flight, 1: 512 x 16, 2: 768 x 12
  EngTables eng;             // generation-3 row engine (constant-bank hop lists)
  bool eng_full = false;     // variant 0 uses the engine for the full H.v too (up hops as row gathers)
  int eng_threads = 1024;    // CTA size of the engine: 1024 threads x 64 registers (measured: 1.84 ms vs 2.02 ms
                             // with 512 x 128 on the 4x4 sector, dn-only pass)

  ~HubbardOp() override {
    up.release(); dn.release(); seg.release(); cls.release(); lng.release();
    eng.release();
    cudaFree(d_hop); cudaFree(d_u);
  }

  HubParams base_params() const {
    HubParams p;
    p.num_up = up.num; p.num_dn = dn.num;
    p.up_states = up.d_states32; p.dn_states = dn.d_states32;
    p.ell_up = up.d_ell; p.cnt_up = up.d_cnt;
    p.ell_dn = dn.d_ell; p.cnt_dn = dn.d_cnt;
    p.e_up = up.d_energy; p.e_dn = dn.d_energy;
    p.hop = d_hop; p.u = d_u; p.num_sites = num_sites;
    p.u0 = u0; p.hop0 = hop0;
    p.row0 = 0; p.nrows = up.num; p.with_up = 1; p.accumulate = 0;
    p.x = nullptr; p.y = nullptr; p.acc_scale = nullptr;
    p.lz.enabled = 0; p.lz.iter = nullptr; p.lz.beta = nullptr; p.lz.alpha = nullptr;
    p.lz.partials = d_partials; p.lz.ticket = d_ticket;
    return p;
  }

  int build_species(SpeciesTables& t, const i64* h_states, i64 num, int fixed_popcount,
                    const BondList& bl, const SiteValues& eps) {
    t.num = num;
    ARG_CHECK(num >= 1, "empty string list");
    ARG_CHECK(num <= (i64)ELL_TGT_MASK, "string list too long for the packed hop table");
    CU_CHECK(cudaMalloc(&t.d_states, sizeof(i64) * num));
    CU_CHECK(cudaMemcpy(t.d_states, h_states, sizeof(i64) * num, cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&t.d_states32, sizeof(uint32_t) * num));
    CU_CHECK(cudaMalloc(&t.d_cnt, num));
    CU_CHECK(cudaMalloc(&t.d_energy, sizeof(double) * num));
    narrow_states_kernel<<<grid_for(num, 256), 256>>>(t.d_states, num, t.d_states32);
    KERNEL_CHECK();
    weighted_elements_kernel<<<grid_for(num, 256), 256>>>(t.d_states, num, eps, t.d_energy);
    KERNEL_CHECK();
    // pass 1: counts only (ell_width = 0), then size the ELL table to the max count
    species_ell_kernel<<<grid_for(num, 128), 128>>>(t.d_states, num, fixed_popcount, sign_width,
                                                    bl, 0, nullptr, t.d_cnt);
    KERNEL_CHECK();
    std::vector<uint8_t> h_cnt(num);
    CU_CHECK(cudaMemcpy(h_cnt.data(), t.d_cnt, num, cudaMemcpyDeviceToHost));
    int w = 0;
    for (i64 i = 0; i < num; ++i) w = h_cnt[i] > w ? h_cnt[i] : w;
    t.width = w;
    CU_CHECK(cudaMalloc(&t.d_ell, sizeof(uint32_t) * (size_t)(w > 0 ? w : 1) * num));
    if (w > 0) {
      species_ell_kernel<<<grid_for(num, 128), 128>>>(t.d_states, num, fixed_popcount,
                                                      sign_width, bl, w, t.d_ell, t.d_cnt);
      KERNEL_CHECK();
    }
    CU_CHECK(cudaDeviceSynchronize());
    return CMPY_OK;
  }

  template <bool UNI, bool LZ>
  int launch(HubParams& p, int use_variant, cudaStream_t st) {
    const i64 total = p.nrows * p.num_dn;
    if (total == 0) return CMPY_OK;
    if (use_variant == 3 && !seg.ok)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "segment variant not available for this sector");
    if (use_variant == 4 && !seg.ok)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "segment variant not available for this sector");
    if (use_variant >= 5 && use_variant <= 7 && !(cls.ok && UNI))
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "class-major variant not available for this sector");
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y)) & 15) == 0;
    if (use_variant == 8 && !(lng.ok && UNI && aligned16 && !p.with_up && !LZ && (p.num_dn % 2 == 0)))
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "long-row variant not available for this call");
    if (use_variant == 9 || use_variant == 10)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "variants 9 / 10 (round-1 chunked-task engine) were removed");
    if (use_variant == 11 && !(eng.ok && UNI))
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "row-engine variant not available for this sector");
    if (p.acc_scale) {   // scaled accumulation: row engine, long-row or plain class-major kernel (dn pass only)
      const bool a16 = ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y)) & 15) == 0;
      if (LZ || p.with_up || !UNI || use_variant != 0)
        return cmpy_fail(CMPY_ERR_UNSUPPORTED, "scaled accumulation: dn pass of a uniform model only");
      if (eng.ok) return launch_eng<LZ>(p, st);
      if (lng.ok && a16 && (p.num_dn % 2 == 0)) return launch_long(p, st, lng);
      if (cls.ok && a16) return launch_cls<LZ>(p, st);
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "scaled accumulation needs the row engine or the class-major kernel");
    }
    if (use_variant == 11 || (use_variant == 0 && eng.ok && UNI && (!p.with_up || eng_full)))
      return launch_eng<LZ>(p, st);
    if (use_variant == 8 || (use_variant == 0 && lng.ok && UNI && aligned16 && !p.with_up && !LZ &&
                             (p.num_dn % 2 == 0))) {
      return launch_long(p, st, lng);
    }
    if (use_variant >= 5 && use_variant <= 7 && !aligned16)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "class-major variant needs 16-byte aligned vectors");
    // default: the segment kernel for the full H.v (its up-hop gathers overlap the shared-memory
    // work: 4.18 vs 4.74 ms on the 4x4 sector), the class-major kernel for row slabs without up
    // hops (sharded operator: 2.1 vs 2.97 ms)
    if ((use_variant >= 5 && use_variant <= 7) ||
        (use_variant == 0 && cls.ok && cls_default && UNI && aligned16 && !p.with_up)) {
      const int saved = cls_shape;
      if (use_variant >= 5) cls_shape = use_variant - 5;
      int rc = launch_cls<LZ>(p, st);
      cls_shape = saved;
      return rc;
    }
    if (use_variant == 3 || use_variant == 4 || (use_variant == 0 && seg.ok)) {
      const bool saved = seg_wide;
      if (use_variant == 3) seg_wide = false;
      if (use_variant == 4) seg_wide = true;
      int rc = launch_seg<LZ>(p, st);
      seg_wide = saved;
      return rc;
    }
    bool use_row = (use_variant == 2) || (use_variant == 0 && row_ok);
    if (use_variant == 2 && !row_ok)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "row variant: the dn row does not fit shared memory");
    if (use_row) {
      size_t smem = sizeof(double) * (size_t)p.num_dn;
      i64 g = (i64)sm_count * row_blocks_per_sm;
      if (g > p.nrows) g = p.nrows;
      if (LZ && g > max_blocks) g = max_blocks;
      hub_row_kernel<UNI, LZ><<<(int)g, row_threads, smem, st>>>(p);
    } else {
      int g = grid_for(total, 256, sm_count * 8);
      hub_flat_kernel<UNI, LZ><<<g, 256, 0, st>>>(p);
    }
    KERNEL_CHECK();
    return CMPY_OK;
  }

  size_t seg_smem() const {
    return (size_t)seg.lay.bytes + (((size_t)dn.num * 4 + 15) & ~(size_t)15) + sizeof(double) * (size_t)dn.num;
  }

  template <bool LZ>
  int launch_seg(HubParams& p, cudaStream_t st) {
    SegParams sp;
    sp.hp = p; sp.lay = seg.lay; sp.blob = seg.d_blob; sp.e_dn_const = seg.e_dn_const;
    const size_t smem = seg_smem();
    i64 g = (i64)sm_count * seg_blocks_per_sm;
    if (g > p.nrows) g = p.nrows;
    const bool uni = uniform && eps_uniform;
    // 16-byte vector path: even row length and 16-byte aligned slab pointers
    const bool vec = ((dn.num & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
    if (seg_wide && vec && uni) {
      // 896 threads (73 registers, 5 gathers in flight per lane) measured 4.00 vs 4.19 ms for 1024
      // threads on the 4x4 sector (640: 4.38, 768: 4.08)
      hub_seg_kernel<true, LZ, true, 896><<<(int)g, 896, smem, st>>>(sp);
    } else if (seg_wide) {
      const int nt = 1024;
      if (vec) {
        if (uni) hub_seg_kernel<true, LZ, true, 1024><<<(int)g, nt, smem, st>>>(sp);
        else hub_seg_kernel<false, LZ, true, 1024><<<(int)g, nt, smem, st>>>(sp);
      } else {
        if (uni) hub_seg_kernel<true, LZ, false, 1024><<<(int)g, nt, smem, st>>>(sp);
        else hub_seg_kernel<false, LZ, false, 1024><<<(int)g, nt, smem, st>>>(sp);
      }
    } else if (vec) {
      if (uni) hub_seg_kernel<true, LZ, true, 512><<<(int)g, seg_threads, smem, st>>>(sp);
      else hub_seg_kernel<false, LZ, true, 512><<<(int)g, seg_threads, smem, st>>>(sp);
    } else {
      if (uni) hub_seg_kernel<true, LZ, false, 512><<<(int)g, seg_threads, smem, st>>>(sp);
      else hub_seg_kernel<false, LZ, false, 512><<<(int)g, seg_threads, smem, st>>>(sp);
    }
    KERNEL_CHECK();
    return CMPY_OK;
  }

  template <bool LZ, int NLH>
  void launch_eng_n(const EngArgs& A, bool with_up, int g, cudaStream_t st) {
    if (eng_threads == 512) {
      if (with_up) hub_eng_kernel<LZ, true, 512, NLH><<<g, 512, eng.H.smem, st>>>(eng.H.C, A);
      else hub_eng_kernel<LZ, false, 512, NLH><<<g, 512, eng.H.smem, st>>>(eng.H.C, A);
    } else {
      if (with_up) hub_eng_kernel<LZ, true, 1024, NLH><<<g, 1024, eng.H.smem, st>>>(eng.H.C, A);
      else hub_eng_kernel<LZ, false, 1024, NLH><<<g, 1024, eng.H.smem, st>>>(eng.H.C, A);
    }
  }

  template <int NLH>
  int raise_eng_limits() {
    int rc = raise_smem_limit(hub_eng_kernel<false, false, 512, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<true, false, 512, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<false, true, 512, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<true, true, 512, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<false, false, 1024, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<true, false, 1024, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<false, true, 1024, NLH>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_eng_kernel<true, true, 1024, NLH>, smem_optin);
    return rc;
  }

  template <bool LZ>
  int launch_eng(HubParams& p, cudaStream_t st) {
    EngArgs A;
    A.hp = p; A.ln.dh_cm = eng.d_dh_cm; A.ln.dl_of_q = eng.d_dl_q; A.ln.lh_lane = reinterpret_cast<const uint4*>(eng.d_lh_lane); A.e_dn_const = eng.H.e_dn_const;
    i64 g = sm_count;
    if (grid_limit > 0 && g > grid_limit) g = grid_limit;
    if (g > p.nrows) g = p.nrows;
    switch (eng_nlh_bound(eng.H.C.nlh)) {
      case 1: launch_eng_n<LZ, 1>(A, p.with_up != 0, (int)g, st); break;
      case 2: launch_eng_n<LZ, 2>(A, p.with_up != 0, (int)g, st); break;
      default: launch_eng_n<LZ, 4>(A, p.with_up != 0, (int)g, st); break;
    }
    KERNEL_CHECK();
    return CMPY_OK;
  }

  // Generation-3 row engine: uniform hop / U / eps, hop != 0, complete dn sector of <= 16 sites.
  int configure_eng(int n_dn, const int* s1, const int* s2, const double* eps) {
    if (!(uniform && eps_uniform) || hop0 == 0.0 || dn.num < 64) return CMPY_OK;
    if (const char* e = getenv("CMPY_ENG_THREADS")) eng_threads = atoi(e) == 512 ? 512 : 1024;
    int rc = build_eng_host(eng.H, num_sites, n_dn, dn.num, nbonds, s1, s2, sign_width, eps, smem_optin,
                            eng_threads / 32, false);
    if (rc || !eng.H.ok) return rc;
    const int nb_lh = eng_nlh_bound(eng.H.C.nlh);
    rc = nb_lh == 1 ? raise_eng_limits<1>() : nb_lh == 2 ? raise_eng_limits<2>() : raise_eng_limits<4>();
    if (rc) return rc;
    int nb = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hub_eng_kernel<true, true, 1024, 4>, 1024, eng.H.smem));
    if (nb < 1) return CMPY_OK;
    CU_CHECK(cudaMalloc(&eng.d_dh_cm, sizeof(uint16_t) * eng.H.dh_cm.size()));
    CU_CHECK(cudaMalloc(&eng.d_dl_q, sizeof(uint16_t) * eng.H.dl_q.size()));
    CU_CHECK(cudaMemcpy(eng.d_dl_q, eng.H.dl_q.data(), sizeof(uint16_t) * eng.H.dl_q.size(), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMalloc(&eng.d_lh_lane, sizeof(uint32_t) * eng.H.lh_lane.size()));
    CU_CHECK(cudaMemcpy(eng.d_dh_cm, eng.H.dh_cm.data(), sizeof(uint16_t) * eng.H.dh_cm.size(), cudaMemcpyHostToDevice));
    CU_CHECK(cudaMemcpy(eng.d_lh_lane, eng.H.lh_lane.data(), sizeof(uint32_t) * eng.H.lh_lane.size(), cudaMemcpyHostToDevice));
    eng.ok = true;
    if (const char* e = getenv("CMPY_ENG_FULL")) eng_full = atoi(e) != 0;
    return CMPY_OK;
  }

  template <bool LZ>
  int launch_cls(HubParams& p, cudaStream_t st) {
    const ClsTables& T = cls;
    ClsParams cp;
    cp.hp = p; cp.lay = T.lay; cp.blob = T.d_blob; cp.pair_seg = T.d_pair_seg; cp.e_dn_const = T.e_dn_const;
    i64 g = sm_count;
    if (grid_limit > 0 && g > grid_limit) g = grid_limit;
    if (g > p.nrows) g = p.nrows;
    // the up-hop offsets of this kernel are 32-bit element offsets relative to the row
    if (p.with_up && (double)(up.num - 1) * (double)dn.num >= 2147483647.0)
      return cmpy_fail(CMPY_ERR_UNSUPPORTED, "class-major variant with up hops: sector too large for 32-bit row offsets");
    // (896 / 768 threads measured 2.36 / 2.43 ms vs 2.12 ms for 1024 on the 4x4 sector: issue-bound)
    if (!p.with_up) hub_cls_kernel<LZ, 1024, 0><<<(int)g, 1024, cls.smem, st>>>(cp);
    else if (cls_shape == 1) hub_cls_kernel<LZ, 512, 16><<<(int)g, 512, cls.smem, st>>>(cp);
    else if (cls_shape == 2) hub_cls_kernel<LZ, 768, 12><<<(int)g, 768, cls.smem, st>>>(cp);
    else hub_cls_kernel<LZ, 1024, 8><<<(int)g, 1024, cls.smem, st>>>(cp);
    KERNEL_CHECK();
    return CMPY_OK;
  }

  // dn-only pass over rows longer than shared memory: one launch per popcount of the top bits
  int launch_long(HubParams& p, cudaStream_t st, LongTables& lt) {
    for (auto& S : lt.sets) {
      ClsParams cp;
      cp.hp = p; cp.lay = S.cls.lay; cp.blob = S.cls.d_blob; cp.pair_seg = S.shift ? S.cls.d_pair_seg1 : S.cls.d_pair_seg;
      cp.e_dn_const = lt.e_dn_const;
      cp.lg.ntop = S.ntop; cp.lg.row_len = S.row_len; cp.lg.nsb = lt.nsb; cp.lg.shift = S.shift;
      cp.lg.top_val = S.d_top_val; cp.lg.sub_off = S.d_sub_off; cp.lg.tb_ptr = S.d_tb_ptr;
      cp.lg.tb_ent = S.d_tb_ent; cp.lg.sb_src = S.d_sb_src; cp.lg.sb_map = S.d_sb_map;
      i64 g = sm_count;
      if (grid_limit > 0 && g > grid_limit) g = grid_limit;
      const i64 items = p.nrows * S.ntop;
      if (g > items) g = items;
      hub_cls_kernel<false, 1024, 8, true><<<(int)g, 1024, S.cls.smem, st>>>(cp);
      KERNEL_CHECK();
    }
    return CMPY_OK;
  }

  int configure_long(int n_dn, const int* s1, const int* s2, const double* eps) {
    if (!(uniform && eps_uniform) || num_sites <= LONG_RBITS) return CMPY_OK;
    int rc = build_long_tables(lng, num_sites, n_dn, dn.num, nbonds, s1, s2, sign_width, eps, smem_optin);
    if (rc || !lng.ok) return rc;
    rc = raise_smem_limit(hub_cls_kernel<false, 1024, 8, true>, smem_optin);
    if (rc) return rc;
    for (auto& S : lng.sets) {
      int nb = 0;
      CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hub_cls_kernel<false, 1024, 8, true>, 1024, S.cls.smem));
      if (nb < 1) { lng.release(); return CMPY_OK; }
    }
    return CMPY_OK;
  }

  // Class-major tables + launch shape (uniform hop / U / eps, complete dn sector).
  int configure_cls(int n_dn, const int* s1, const int* s2, const double* eps) {
    if (!(uniform && eps_uniform)) return CMPY_OK;
    int rc = build_cls_tables(cls, num_sites, n_dn, dn.num, nbonds, s1, s2, sign_width, eps, smem_optin);
    if (rc || !cls.ok) return rc;
    rc = raise_smem_limit(hub_cls_kernel<false, 1024, 8>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<false, 1024, 0>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<true, 1024, 0>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<true, 1024, 8>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<false, 512, 16>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<true, 512, 16>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<false, 768, 12>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_cls_kernel<true, 768, 12>, smem_optin);
    if (rc) return rc;
    int nb = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hub_cls_kernel<true, 1024, 8>, 1024, cls.smem));
    if (nb < 1) { cls.ok = false; return CMPY_OK; }
    cls_default = dn.num >= 2048;  // long rows: one CTA per SM anyway
    return CMPY_OK;
  }

  template <bool UNI, bool VEC>
  int configure_seg_inst(size_t smem, int& nb_out) {
    int rc = raise_smem_limit(hub_seg_kernel<UNI, false, VEC, 512>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_seg_kernel<UNI, true, VEC, 512>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_seg_kernel<UNI, false, VEC, 1024>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_seg_kernel<UNI, true, VEC, 1024>, smem_optin);
    if (!rc && UNI && VEC) rc = raise_smem_limit(hub_seg_kernel<true, false, true, 896>, smem_optin);
    if (!rc && UNI && VEC) rc = raise_smem_limit(hub_seg_kernel<true, true, true, 896>, smem_optin);
    if (rc) return rc;
    int nb = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hub_seg_kernel<UNI, true, VEC, 512>,
                                                           seg_threads, smem));
    nb_out = nb;
    return CMPY_OK;
  }

  // Two-level tables + launch shape of the segment kernel (dn species must be a complete
  // fixed-popcount sector whose row + tables fit shared memory).
  int configure_seg(int n_dn, const int* s1, const int* s2, const double* eps) {
    int rc = build_seg_tables(seg, num_sites, n_dn, dn.num, nbonds, s1, s2, sign_width, eps, uniform);
    if (rc || !seg.ok) return rc;
    const size_t smem = seg_smem();
    if ((i64)smem > smem_optin - 4096) { seg.ok = false; return CMPY_OK; }
    // two columns per lane: enough threads to cover a row once, at most 512
    i64 t = ((dn.num + 1) / 2 + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > 512) t = 512;
    if (smem < 40 * 1024 && t > 256) t = 256;  // small rows: several CTAs per SM
    seg_threads = (int)t;
    const bool uni = uniform && eps_uniform;
    int nb1 = 0, nb2 = 0;
    if (uni) { rc = configure_seg_inst<true, true>(smem, nb1); if (!rc) rc = configure_seg_inst<true, false>(smem, nb2); }
    else { rc = configure_seg_inst<false, true>(smem, nb1); if (!rc) rc = configure_seg_inst<false, false>(smem, nb2); }
    if (rc) return rc;
    int nb = nb1 < nb2 ? nb1 : nb2;
    if (nb < 1) { seg.ok = false; return CMPY_OK; }
    seg_blocks_per_sm = nb > 8 ? 8 : nb;
    // long rows (one CTA per SM): 1024-thread CTAs hide the shared-memory latency chains
    // better than 512 x 128 registers (measured on B200: 4.19 vs 4.93 ms on the 4x4 sector)
    seg_wide = (seg_blocks_per_sm == 1 && dn.num >= 2048);
    return CMPY_OK;
  }

  template <bool UNI>
  int configure_row() {
    size_t smem = sizeof(double) * (size_t)dn.num;
    row_ok = false;
    if (dn.num < 64 || (i64)smem > smem_optin - 2048) return CMPY_OK;
    i64 t = (dn.num + 3) / 4;
    t = ((t + 31) / 32) * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    row_threads = (int)t;
    int rc = raise_smem_limit(hub_row_kernel<UNI, false>, smem_optin);
    if (!rc) rc = raise_smem_limit(hub_row_kernel<UNI, true>, smem_optin);
    if (rc) return rc;
    int nb = 0;
    CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hub_row_kernel<UNI, true>,
                                                           row_threads, smem));
    if (nb < 1) return CMPY_OK;
    row_blocks_per_sm = nb > 8 ? 8 : nb;
    row_ok = true;
    return CMPY_OK;
  }

  int apply_slab(const double* x, double* y, i64 row0, i64 nrows, int with_up, int accumulate,
                 const LzCtx& lz, cudaStream_t st, const double* acc_scale = nullptr) {
    HubParams p = base_params();
    p.x = x; p.y = y; p.row0 = row0; p.nrows = nrows; p.with_up = with_up;
    p.accumulate = accumulate; p.acc_scale = acc_scale;
    if (lz.enabled) {
      p.lz = lz; p.lz.partials = d_partials; p.lz.ticket = d_ticket;
      return uniform ? launch<true, true>(p, variant, st) : launch<false, true>(p, variant, st);
    }
    return uniform ? launch<true, false>(p, variant, st) : launch<false, false>(p, variant, st);
  }

  int apply(const double* x, double* y, const LzCtx& lz, cudaStream_t st) override {
    return apply_slab(x, y, 0, up.num, 1, 0, lz, st);
  }

  int diagonal(double* d_diag, cudaStream_t st) override {
    HubParams p = base_params();
    int g = grid_for(size, 256, sm_count * 8);
    if (uniform) hub_diag_kernel<true><<<g, 256, 0, st>>>(p, d_diag);
    else hub_diag_kernel<false><<<g, 256, 0, st>>>(p, d_diag);
    KERNEL_CHECK();
    return CMPY_OK;
  }
};
