// eng_emu.cu -- TEST INFRASTRUCTURE (not part of the product): runs the staging map and the two
// shared-memory phases of the generation-3 row engine (cmpy_b200/csrc/hubbard_eng.cuh) lane by lane
// on the CPU, from the library's own table builder.  Inside one phase every (warp, lane) only reads
// what the previous phase wrote and writes its own slots, so executing the lanes one after the other
// is what the barrier-separated kernel computes.  tests/test_eng_emulation.py compares the result
// with a direct evaluation of (D + T_dn) x on one row (ref: cmpy/operators.py:305-527).
//
// Built by the test itself:  nvcc -O1 -std=c++17 -shared -Xcompiler -fPIC eng_emu.cu
#include "../../cmpy_b200/csrc/hubbard_eng.cuh"
#include <cmath>

// returns 0 = ran, 1 = sector not supported by the engine, < 0 = error
extern "C" int eng_emu_row(int num_sites, int n_dn, int nbonds, const int* s1, const int* s2, int sign_width,
                           double eps0, double u0, double hop0, unsigned ups, double e_up, int nwarps,
                           int accumulate, const double* x_row, double* y_row, int* info) {
  EngHost T;
  const u64* B = host_binom();
  const i64 num_dn = (i64)B[num_sites * BINOM_N + n_dn];
  std::vector<double> eps(num_sites, eps0);
  int rc = build_eng_host(T, num_sites, n_dn, num_dn, nbonds, s1, s2, sign_width, eps.data(), 232448, nwarps, false);
  if (rc) return -1;
  if (!T.ok) return 1;
  const EngConst& C = T.C;
  const int xs_total = C.xs_elems + ENG_ZREG;
  std::vector<double> buf(2 * (size_t)xs_total, 0.0);
  double* xs = buf.data();
  const eng_addr xs_a = (eng_addr)(uintptr_t)xs, ydelta = (eng_addr)xs_total * 8u;
  EngLane ln;
  ln.dh_cm = T.dh_cm.data();
  ln.dl_of_q = T.dl_q.data();
  ln.lh_lane = reinterpret_cast<const uint4*>(T.lh_lane.data());
  // staging (the cp.async loop of the kernel)
  std::vector<int> hit(C.xs_elems, 0);
  for (int warp = 0; warp < nwarps; ++warp)
    for (int it = C.sptr[warp]; it < C.sptr[warp + 1]; ++it) {
      const uint32_t ts = C.task_s[it];
      const int k = (int)(ts & 15u), jA = (int)((ts >> 8) & 255u), jB = jA + (int)((ts >> 16) & 255u);
      const int sk = C.S[k];
      for (int jj = jA; jj < jB; ++jj)
        for (int lane = 0; lane < 32; ++lane)
          for (int t = 0; t < 3; ++t)
            if (lane + 32 * t < sk) {
              const int slot = (C.xb8[k] + jj * C.P8[k]) / 8 + lane + 32 * t;
              xs[slot] = x_row[C.goff_cm[C.hoff[k] + jj] + lane + 32 * t];
              hit[slot] += 1;
            }
    }
  int staged = 0;
  for (int v : hit) { if (v > 1) return -2; staged += v; }
  if (staged != (int)num_dn) return -3;
  const double inv_hop = 1.0 / hop0;
  double e_dn_const = T.e_dn_const;
  const double eu_s = (e_up + e_dn_const) * inv_hop, u0_s = u0 * inv_hop;
  std::vector<double> dg(ENG_MAX_Q + ENG_MAX_SEG, 0.0);
  for (int i = 0; i < 1024; ++i) eng_fill_diag(C, ln, dg.data(), i, ups, eu_s, u0_s);
  const eng_addr dg_a = (eng_addr)(uintptr_t)dg.data();
  for (int warp = 0; warp < nwarps; ++warp)
    for (int lane = 0; lane < 32; ++lane) {
      switch (eng_nlh_bound(C.nlh)) {
        case 1: eng_run_a<1>(C, ln, warp, xs_a, ydelta, dg_a, lane); break;
        case 2: eng_run_a<2>(C, ln, warp, xs_a, ydelta, dg_a, lane); break;
        default: eng_run_a<4>(C, ln, warp, xs_a, ydelta, dg_a, lane); break;
      }
    }
  EngEpi E;
  E.xr = x_row; E.yr = y_row; E.hop0 = hop0; E.accumulate = accumulate; E.cu = 0; E.c1 = 1.0; E.c2 = 0.0;
  E.up_off = nullptr; E.up_coef = nullptr; E.s1 = 1.0; E.s2 = 0.0; E.has_prev = false;
  double dot = 0.0;
  for (int warp = 0; warp < nwarps; ++warp)
    for (int lane = 0; lane < 32; ++lane) eng_run_b<false, false>(C, warp, xs_a, ydelta, E, dot, lane);
  // slack slots and the zero region must still be zero (no phase may write them)
  for (int i = C.xs_elems; i < xs_total; ++i) if (xs[i] != 0.0) return -4;
  if (info) {
    info[0] = C.aptr[nwarps]; info[1] = C.bptr[nwarps]; info[2] = C.sptr[nwarps]; info[3] = C.xs_elems;
    info[4] = (int)sizeof(EngConst); info[5] = C.nlh;
  }
  return 0;
}
