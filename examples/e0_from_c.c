/* e0_from_c.c -- the C ABI of libcmpy_b200.so used from plain C (no Python, no PyTorch):
 * ground-state energy of the half-filled 8-site Hubbard chain (BASELINE config C1) by the
 * fused GPU Lanczos.  Expected output: E0 = -20.235806999130.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/e0_from_c.c \
 *       -L cmpy_b200 -lcmpy_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cmpy_b200 -o e0_from_c
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cmpy_b200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != 0) {                                                        \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cmpy_last_error());    \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  enum { L = 8, N = 4 };
  int64_t num = 0;
  CHECK(cmpy_binomial(L, N, &num));                 /* 70 strings per species */

  /* sector strings: enumerated on the device, copied back for the operator constructor */
  int64_t* d_states = NULL;
  if (cudaMalloc((void**)&d_states, sizeof(int64_t) * num) != cudaSuccess) return 2;
  CHECK(cmpy_sector_enumerate(L, N, d_states, NULL));
  int64_t* states = (int64_t*)malloc(sizeof(int64_t) * num);
  cudaMemcpy(states, d_states, sizeof(int64_t) * num, cudaMemcpyDeviceToHost);

  /* open chain, U = 4, mu = 2 (eps - mu = -2 on every site), t = +1 */
  int32_t bonds[2 * (L - 1)];
  double hop[L - 1], eps[L], u[L];
  for (int i = 0; i < L - 1; ++i) { bonds[2 * i] = i; bonds[2 * i + 1] = i + 1; hop[i] = 1.0; }
  for (int i = 0; i < L; ++i) { eps[i] = -2.0; u[i] = 4.0; }
  cmpy_op_t op = NULL;
  CHECK(cmpy_hubbard_create(L, states, num, states, num, 1, L - 1, bonds, hop, eps, u, L, &op));
  int64_t dim = 0;
  CHECK(cmpy_op_size(op, &dim));

  /* start vector and work space on the device */
  double* h_v = (double*)malloc(sizeof(double) * dim);
  for (int64_t i = 0; i < dim; ++i) h_v[i] = cos(0.37 * (double)i);
  double *d_v0, *d_w0, *d_w1;
  cudaMalloc((void**)&d_v0, sizeof(double) * dim);
  cudaMalloc((void**)&d_w0, sizeof(double) * dim);
  cudaMalloc((void**)&d_w1, sizeof(double) * dim);
  cudaMemcpy(d_v0, h_v, sizeof(double) * dim, cudaMemcpyHostToDevice);

  enum { MAXIT = 400 };
  double alpha[MAXIT], beta[MAXIT + 1], e0 = 0.0, resid = 0.0;
  int nit = 0;
  CHECK(cmpy_lanczos_run(op, d_v0, d_w0, d_w1, MAXIT, 1e-12, 1e-9, 10, 1, alpha, beta, &nit, &e0, &resid,
                         NULL, NULL));
  printf("dim = %lld, E0 = %.12f after %d Lanczos steps (residual estimate %.1e)\n", (long long)dim, e0, nit,
         resid);

  CHECK(cmpy_op_destroy(op));
  cudaFree(d_states); cudaFree(d_v0); cudaFree(d_w0); cudaFree(d_w1);
  free(states); free(h_v);
  return fabs(e0 - (-20.235806999130)) < 1e-9 ? 0 : 3;
}
