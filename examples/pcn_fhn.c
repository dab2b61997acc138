/*
 * pcn_fhn.c -- the drop-in boundary used from plain C: the loop of test/partialbridgenuH.jl:155-198 /
 * project_partialbridge/partialbridge_fitzhugh.jl:125-176 for an ensemble of chains, written against
 * include/bridge_b200.h only (no Python, no torch).  Build:
 *   gcc -std=c11 -O2 -Iinclude examples/pcn_fhn.c -o examples/pcn_fhn -Lbridge.jl_b200/lib -lbridge_b200 \
 *       -Wl,-rpath,$PWD/bridge.jl_b200/lib -lm
 * Usage: pcn_fhn [chains] [grid points] [iterations]; prints one line "chains N iters acc ll_sum xend_sum" that
 * tests/test_c_example.py compares with the same run through the Python mirror.  Without a CUDA device it prints the
 * library's error (BB_ERR_NODEVICE) and exits with status 3: the library has no CPU path.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bridge_b200.h"

#define CHECK(call)                                                                      \
  do {                                                                                   \
    int st__ = (call);                                                                   \
    if (st__ != BB_OK) {                                                                 \
      fprintf(stderr, "%s -> %d: %s %s\n", #call, st__, bb_strerror(st__), bb_last_cuda_error()); \
      exit(st__ == BB_ERR_NODEVICE ? 3 : 1);                                             \
    }                                                                                    \
  } while (0)

int main(int argc, char** argv) {
  const int64_t P = argc > 1 ? atoll(argv[1]) : 1000;
  const int N = argc > 2 ? atoi(argv[2]) : 129;
  const int iters = argc > 3 ? atoi(argv[3]) : 5;
  enum { S = 2, D = 2 };
  const double obs_t[S + 1] = {0.0, 0.5, 1.0}, obs_v[S] = {-1.0, -0.5};
  /* target: FitzhughDiffusion(0.1, 0.0, 1.5, 0.8, 0.3), partialbridge_fitzhugh.jl:48 */
  bb_model model;
  memset(&model, 0, sizeof(model));
  model.id = BB_MODEL_FHN_HYPO; model.d = 2; model.dprime = 1;
  const double par[5] = {0.1, 0.0, 1.5, 0.8, 0.3};
  memcpy(model.par, par, sizeof(par));
  const double L[2] = {1.0, 0.0}, Sigma[1] = {1e-10}, eps = 1e-3, x0[2] = {-0.5, -0.6}, rho = 0.99;

  bb_ctx* ctx = NULL;
  CHECK(bb_ctx_create(0, &ctx));
  bb_ens* ens = NULL;
  CHECK(bb_ens_create(ctx, P, S, N, 2, 1, BB_ENS_DOUBLE_BUFFER, &ens));

  /* time grids tau(t) = t (2 - t/T) per segment (:13-14), backward chain right to left (bolus3.jl:162-180) */
  double* tt = malloc(sizeof(double) * S * N);
  for (int s = 0; s < S; s++)
    for (int i = 0; i < N; i++) {
      const double T = obs_t[s + 1] - obs_t[s], u = T * i / (N - 1);
      tt[s * N + i] = obs_t[s] + u * (2.0 - u / T);
    }
  double nu[D] = {0, 0}, Hp[D * D] = {1.0 / eps, 0, 0, 1.0 / eps};
  CHECK(bb_gpupdate_nuH(ctx, 2, 1, nu, Hp, L, Sigma, &obs_v[S - 1]));
  bb_guide* guides[S];
  double* nut = malloc(sizeof(double) * N * D);
  double* Ht = malloc(sizeof(double) * N * D * D);
  for (int s = S - 1; s >= 0; s--) {
    const double v = obs_v[s];
    /* "matching" auxiliary process, :106-108 */
    const double Bt[4] = {1 / par[0], -1 / par[0], par[2], -1.0};
    const double bt[2] = {par[1] / par[0] - (v * v * v) / par[0], par[3]};
    const double at[4] = {0, 0, 0, par[4] * par[4]};
    bb_aux aux = {2, 1, Bt, bt, at, NULL};
    double nul[D], Hpl[D * D], C;
    CHECK(bb_backward_nuH(ctx, BB_ODE_LYAP, N, 2, tt + s * N, &aux, nu, Hp, 0.0, nut, Ht, nul, Hpl, &C));
    CHECK(bb_guide_create(ctx, BB_GUIDE_NUH, N, 2, 0, tt + s * N, Ht, nut, NULL, NULL, Bt, bt, 1, &guides[s]));
    memcpy(nu, nul, sizeof(nu)); memcpy(Hp, Hpl, sizeof(Hp));
    if (s > 0) CHECK(bb_gpupdate_nuH(ctx, 2, 1, nu, Hp, L, Sigma, &obs_v[s - 1]));
    CHECK(bb_ens_set_grid(ens, s, tt + s * N, N));
  }
  CHECK(bb_ens_set_start(ens, x0, 2, 1));
  CHECK(bb_wiener_sample(ens, 44, 0xFFFFFFFEu));                     /* W = sample(tt, Wiener()) */
  CHECK(bb_guided_euler_ll(ens, &model, guides, 0, BB_RUN_STORE_X)); /* solve!(Euler(), X, x0, W, Po); ll = llikelihood(...) */
  for (int it = 0; it < iters; it++) CHECK(bb_pcn_step(ens, &model, guides, rho, 44, (uint32_t)it, 0, BB_RUN_STORE_X));
  int64_t acc = 0;
  CHECK(bb_ens_get_acc(ens, &acc));
  double* ll = malloc(sizeof(double) * P);
  double* xend = malloc(sizeof(double) * P * D);
  CHECK(bb_ens_get_f64(ens, BB_F_LL, 0, P, ll));
  CHECK(bb_ens_get_f64(ens, BB_F_XEND, 0, P, xend));
  double sll = 0, sx = 0;
  for (int64_t p = 0; p < P; p++) { sll += ll[p]; sx += xend[2 * p] + xend[2 * p + 1]; }
  printf("chains %lld N %d iters %d acc %lld ll_sum %.17g xend_sum %.17g\n", (long long)P, N, iters, (long long)acc, sll, sx);
  for (int s = 0; s < S; s++) bb_guide_destroy(guides[s]);
  bb_ens_destroy(ens);
  bb_ctx_destroy(ctx);
  free(tt); free(nut); free(Ht); free(ll); free(xend);
  return 0;
}
