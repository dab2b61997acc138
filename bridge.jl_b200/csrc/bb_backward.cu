/*
 * bb_backward.cu -- the constructors of the guided proposals: backward ODEs for the guiding tables.
 *
 *   updateνH⁺C                         src/partialbridgenuH.jl:1-17
 *   partialbridgeodeνH!(R3 / Lyap)     src/partialbridgenuH.jl:21-55, 86-103; src/lyap.jl:2-6
 *   partialbridgeodeHνH!(R3)           src/partialbridgenuH.jl:64-81
 *   gpHinv! / gpV! (GuidedBridge)      src/guip.jl:172-180, src/gode.jl:2-3,13,21, src/ode.jl:88-97
 *   partialbridgeode!(R3) (L, M, mu)   src/partialbridge.jl:1-22
 *   kernelr3                           src/ode.jl:44-49
 *   gpupdate                           src/guip.jl:221-243, partialbridge_bolus3.jl:128-137
 *
 * Each call integrates ONE d x d system over the N grid points: O(N d^3) strictly sequential work,
 * done by a single thread with the state in registers (the path kernels spend P*N steps for every
 * one of these).  The auxiliary process arrives as values at the Ralston stage times (bb_aux).
 * Operation order is the oracle's (ORACLE_GPU_ORDER build), see bb_device.cuh.
 */
#include <math.h>
#include <string.h>

#include <vector>

#include "bb_backward.cuh"
#include "bb_host.h"

/* d > 3: block-parallel generic-d kernels (bb_backward_gen.cu); all pointers are device pointers */
cudaError_t bb_gen_backward_nuH(cudaStream_t st, int method, int N, int d, const double* tt, const double* B,
                                const double* beta, const double* a, const double* a_left, int is_const,
                                const double* nu_end, const double* Hplus_end, double C0, double* nu, double* H,
                                double* out_left, int* status);
cudaError_t bb_gen_update_nuHC(cudaStream_t st, int d, int m, const double* in, double* out, int* status);

namespace {
using namespace bbk;

/* ---- partialbridgeodeνH!  (R3: partialbridgenuH.jl:21-55; Lyap: :86-103) */
template <int d>
__global__ void k_backward_nuH(int method, int N, const double* __restrict__ tt, aux_dev A,
                               const double* nu_end, const double* Hplus_end, double C0, double* nu, double* H,
                               double* out_left /* nu_left[d], Hplus_left[d*d], C */, int* status) {
  double Hp[d * d], Hc[d * d], v[d];
  for (int q = 0; q < d * d; q++) Hp[q] = Hplus_end[q];
  for (int q = 0; q < d; q++) v[q] = nu_end[q];
  if (minv<d>(Hp, Hc)) { *status = BB_ERR_SINGULAR; return; }
  for (int q = 0; q < d * d; q++) H[(size_t)(N - 1) * d * d + q] = Hc[q];
  for (int q = 0; q < d; q++) nu[(size_t)(N - 1) * d + q] = v[q];
  double Cc = C0;
  for (int i = N - 2; i >= 0; i--) {
    const double dt = tt[i] - tt[i + 1];
    if (nuH_step<d>(method, A, i, dt, Hp, Hc, v, Cc)) { *status = BB_ERR_SINGULAR; return; }
    for (int q = 0; q < d; q++) nu[(size_t)i * d + q] = v[q];
    for (int q = 0; q < d * d; q++) H[(size_t)i * d * d + q] = Hc[q];
  }
  for (int q = 0; q < d; q++) out_left[q] = v[q];
  for (int q = 0; q < d * d; q++) out_left[d + q] = Hp[q];
  out_left[d + d * d] = Cc;
  *status = BB_OK;
}

/* ---- partialbridgeodeHνH!  partialbridgenuH.jl:64-81 */
template <int d>
__global__ void k_backward_FH(int N, const double* __restrict__ tt, aux_dev A, const double* F_end,
                              const double* H_end, double C0, double* F, double* H, double* out_C, int* status) {
  double Hc[d * d], Fc[d];
  for (int q = 0; q < d * d; q++) Hc[q] = H_end[q];
  for (int q = 0; q < d; q++) Fc[q] = F_end[q];
  for (int q = 0; q < d * d; q++) H[(size_t)(N - 1) * d * d + q] = Hc[q];
  for (int q = 0; q < d; q++) F[(size_t)(N - 1) * d + q] = Fc[q];
  double Cc = C0;
  for (int i = N - 2; i >= 0; i--) {
    const double dt = tt[i] - tt[i + 1];
    aux_at<d> s0(A, i, 0);
    double aF[d];
    mvec<d, d>(s0.a, Fc, aF);
    Cc += (vdot<d>(s0.beta, Fc) * dt + 0.5 * vdot<d>(Fc, aF) * dt) - 0.5 * trprod<d>(Hc, s0.a) * dt;
    r3_step<d * d>(rhs_dH<d>{A, i}, Hc, dt);
    r3_step<d>(rhs_dF<d>{A, i, Hc}, Fc, dt);
    for (int q = 0; q < d; q++) F[(size_t)i * d + q] = Fc[q];
    for (int q = 0; q < d * d; q++) H[(size_t)i * d * d + q] = Hc[q];
  }
  *out_C = Cc;
  *status = BB_OK;
}

/* ---- gpHinv! / gpV!  guip.jl:172-180 */
template <int d>
__global__ void k_backward_HV(int N, const double* __restrict__ tt, aux_dev A, const double* v,
                              const double* hdia_end, double* Hdia, double* V, int* status) {
  double K[d * d], Vc[d];
  for (int q = 0; q < d * d; q++) K[q] = hdia_end ? hdia_end[q] : 0.0;
  for (int q = 0; q < d; q++) Vc[q] = v[q];
  for (int q = 0; q < d * d; q++) Hdia[(size_t)(N - 1) * d * d + q] = K[q];
  for (int q = 0; q < d; q++) V[(size_t)(N - 1) * d + q] = Vc[q];
  for (int i = N - 2; i >= 0; i--) {
    const double dt = tt[i] - tt[i + 1];
    r3_step<d * d>(rhs_dHinv<d>{A, i}, K, dt);
    r3_step<d>(rhs_btilde<d>{A, i}, Vc, dt);
    for (int q = 0; q < d * d; q++) Hdia[(size_t)i * d * d + q] = K[q];
    for (int q = 0; q < d; q++) V[(size_t)i * d + q] = Vc[q];
  }
  *status = BB_OK;
}

/* ---- partialbridgeode!  partialbridge.jl:1-22 */
template <int d, int m>
__global__ void k_backward_LMmu(int N, const double* __restrict__ tt, aux_dev A, const double* L,
                                const double* Sigma, double* Lt, double* Mt, double* mut, int* status) {
  double Lc[m * d], Mp[m * m], Mi[m * m], mu[m];
  for (int q = 0; q < m * d; q++) Lc[q] = L[q];
  for (int q = 0; q < m * m; q++) Mp[q] = Sigma[q];
  for (int q = 0; q < m; q++) mu[q] = 0.0;
  if (minv<m>(Mp, Mi)) { *status = BB_ERR_SINGULAR; return; }
  for (int q = 0; q < m * d; q++) Lt[(size_t)(N - 1) * m * d + q] = Lc[q];
  for (int q = 0; q < m * m; q++) Mt[(size_t)(N - 1) * m * m + q] = Mi[q];
  for (int q = 0; q < m; q++) mut[(size_t)(N - 1) * m + q] = mu[q];
  for (int i = N - 2; i >= 0; i--) {
    const double dt = tt[i] - tt[i + 1];
    r3_step<m * d>(rhs_dL<d, m>{A, i}, Lc, dt);
    r3_step<m * m>(rhs_dMplus<d, m>{A, i, Lc}, Mp, dt);
    r3_step<m>(rhs_dmu<d, m>{A, i, Lc}, mu, dt);
    if (minv<m>(Mp, Mi)) { *status = BB_ERR_SINGULAR; return; }
    for (int q = 0; q < m * d; q++) Lt[(size_t)i * m * d + q] = Lc[q];
    for (int q = 0; q < m * m; q++) Mt[(size_t)i * m * m + q] = Mi[q];
    for (int q = 0; q < m; q++) mut[(size_t)i * m + q] = mu[q];
  }
  *status = BB_OK;
}

/* ---- updateνH⁺C  partialbridgenuH.jl:1-17.  io: in = L[m*d], Sigma[m*m], v[m], eps ; out = nu[d], Hplus[d*d], C */
template <int d, int m>
__global__ void k_update_nuHC(const double* in, double* out, int* status) {
  const double *L = in, *Sigma = in + m * d, *v = Sigma + m * m;
  const double eps = v[m];
  double Si[m * m], Lt[d * m], LtSi[d * m], H[d * d], Hplus[d * d], Siv[m], nu[d];
  if (minv<m>(Sigma, Si)) { *status = BB_ERR_SINGULAR; return; }
  mtr<m, d>(L, Lt);
  mmul<d, m, m>(Lt, Si, LtSi);
  mmul<d, m, d>(LtSi, L, H);
  for (int i = 0; i < d; i++) H[i * d + i] += eps;
  if (minv<d>(H, Hplus)) { *status = BB_ERR_SINGULAR; return; }
  double HpLt[d * m], HpLtSi[d * m];
  mmul<d, d, m>(Hplus, Lt, HpLt);
  mmul<d, m, m>(HpLt, Si, HpLtSi);
  mvec<d, m>(HpLtSi, v, nu);
  mvec<m, m>(Si, v, Siv);
  double c = 0.0;
  c += 0.5 * vdot<m>(v, Siv);
  double ld;
  if (m == 1) ld = log(Sigma[0]);
  else {
    double Lc[m * m];
    if (chol_lower<m>(Sigma, Lc)) { *status = BB_ERR_SINGULAR; return; }
    ld = 0;
    for (int i = 0; i < m; i++) ld += 2 * log(Lc[i * m + i]);
  }
  c += m / 2.0 * log(2 * 3.14159265358979323846) + 0.5 * ld;
  for (int q = 0; q < d; q++) out[q] = nu[q];
  for (int q = 0; q < d * d; q++) out[d + q] = Hplus[q];
  out[d + d * d] = c;
  *status = BB_OK;
}

/* ---- observation update  Z = I - H⁺L'(Σ + LH⁺L')⁻¹L;  ν <- Z H⁺L'Σ⁻¹v + Zν;  H⁺ <- Z H⁺
 * io: in = nu[d], Hplus[d*d], L[m*d], Sigma[m*m], v[m]; out = nu[d], Hplus[d*d] */
template <int d, int m>
__global__ void k_gpupdate(const double* in, double* out, int* status) {
  double nu[d], Hplus[d * d];
  const double *L = in + d + d * d, *Sigma = L + m * d, *v = Sigma + m * m;
  for (int q = 0; q < d; q++) nu[q] = in[q];
  for (int q = 0; q < d * d; q++) Hplus[q] = in[d + q];
  if (gpupdate_dev<d, m>(nu, Hplus, L, Sigma, v)) { *status = BB_ERR_SINGULAR; return; }
  for (int q = 0; q < d; q++) out[q] = nu[q];
  for (int q = 0; q < d * d; q++) out[d + q] = Hplus[q];
  *status = BB_OK;
}

/* ---------------------------------------------------------------- host plumbing */
struct dev_buf { /* work space from the context's recycled buffers */
  double* p = nullptr;
  bb_ctx* ctx = nullptr;
  ~dev_buf() { if (p) bb_pool_release(ctx, p); }
};
struct dev_pack {
  /* one device allocation holding several host arrays back to back */
  std::vector<double> host;
  std::vector<size_t> off;
  dev_buf dev;
  size_t add(const double* src, size_t n) {
    size_t o = host.size();
    off.push_back(o);
    host.resize(o + n);
    if (src) memcpy(host.data() + o, src, n * sizeof(double));
    else memset(host.data() + o, 0, n * sizeof(double));
    return o;
  }
};

static int pack_upload(bb_ctx* ctx, dev_pack& pk, size_t extra_out) {
  const size_t n = pk.host.size() + extra_out + 2;
  pk.dev.ctx = ctx;
  BB_CUDA(bb_pool_alloc(ctx, n * sizeof(double), (void**)&pk.dev.p));
  BB_CUDA(cudaMemcpyAsync(pk.dev.p, pk.host.data(), pk.host.size() * sizeof(double), cudaMemcpyHostToDevice,
                          ctx->stream));
  BB_CUDA(cudaMemsetAsync(pk.dev.p + pk.host.size(), 0, (extra_out + 2) * sizeof(double), ctx->stream));
  return BB_OK;
}

static int fill_aux(dev_pack& pk, const bb_aux* aux, int N, int d, size_t* oB, size_t* obeta, size_t* oa,
                    size_t* oal) {
  if (!aux || aux->d != d || !aux->B || !aux->beta || !aux->a) return BB_ERR_ARG;
  if (aux->is_const) {
    *oB = pk.add(aux->B, d * d);
    *obeta = pk.add(aux->beta, d);
    *oa = pk.add(aux->a, d * d);
    *oal = *oa;
  } else {
    const size_t ne = (size_t)(N - 1) * 3;
    *oB = pk.add(aux->B, ne * d * d);
    *obeta = pk.add(aux->beta, ne * d);
    *oa = pk.add(aux->a, ne * d * d);
    *oal = aux->a_left ? pk.add(aux->a_left, (size_t)(N - 1) * d * d) : *oa;
  }
  return BB_OK;
}

}  // namespace

#define BB_D_SWITCH(d, CALL) \
  switch (d) {               \
    case 1: { constexpr int D = 1; CALL; } break; \
    case 2: { constexpr int D = 2; CALL; } break; \
    case 3: { constexpr int D = 3; CALL; } break; \
    default: return BB_ERR_UNSUPPORTED; \
  }
#define BB_DM_SWITCH(d, m, CALL)                                         \
  switch ((d) * 4 + (m)) {                                               \
    case 1 * 4 + 1: { constexpr int D = 1, M = 1; CALL; } break;         \
    case 2 * 4 + 1: { constexpr int D = 2, M = 1; CALL; } break;         \
    case 2 * 4 + 2: { constexpr int D = 2, M = 2; CALL; } break;         \
    case 3 * 4 + 1: { constexpr int D = 3, M = 1; CALL; } break;         \
    case 3 * 4 + 2: { constexpr int D = 3, M = 2; CALL; } break;         \
    case 3 * 4 + 3: { constexpr int D = 3, M = 3; CALL; } break;         \
    default: return BB_ERR_UNSUPPORTED;                                  \
  }

/* runs `launch`, then copies `nout` doubles starting at device offset `oout` plus the status word back */
template <class Launch>
static int run_small(bb_ctx* ctx, dev_pack& pk, size_t oout, size_t nout, std::vector<double>& out, Launch launch) {
  int* dstatus = reinterpret_cast<int*>(pk.dev.p + pk.host.size() + (oout - pk.host.size()) + nout);
  bb_time_begin(ctx);
  launch(dstatus);
  bb_time_end(ctx);
  BB_CUDA(cudaGetLastError());
  ctx->launches++;
  out.resize(nout + 1);
  BB_CUDA(cudaMemcpyAsync(out.data(), pk.dev.p + oout, (nout + 1) * sizeof(double), cudaMemcpyDeviceToHost,
                          ctx->stream));
  BB_CUDA(cudaStreamSynchronize(ctx->stream));
  int st;
  memcpy(&st, &out[nout], sizeof(int));
  return st;
}

extern "C" int bb_update_nuHC(bb_ctx* ctx, int32_t d, int32_t m, const double* L, const double* Sigma,
                              const double* v, double eps, double* nu, double* Hplus, double* C) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!L || !Sigma || !v || !nu || !Hplus || !C) return BB_ERR_ARG;
  if (m < 1 || m > d) return BB_ERR_ASSERT_M;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  pk.add(L, m * d); pk.add(Sigma, m * m); pk.add(v, m); pk.add(&eps, 1);
  const size_t nout = d + d * d + 1, oout = pk.host.size();
  int rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  std::vector<double> out;
  if (d > 3) {
    if (d > BB_MAXD_WIDE) return BB_ERR_UNSUPPORTED;
    rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
      bb_gen_update_nuHC(ctx->stream, d, m, pk.dev.p, pk.dev.p + oout, st);
    });
  } else
  BB_DM_SWITCH(d, m, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                 k_update_nuHC<D, M><<<1, 1, 0, ctx->stream>>>(pk.dev.p, pk.dev.p + oout, st);
               }));
  if (rc) return rc;
  memcpy(nu, out.data(), sizeof(double) * d);
  memcpy(Hplus, out.data() + d, sizeof(double) * d * d);
  *C = out[d + d * d];
  return BB_OK;
}

static int gpupdate_impl(bb_ctx* ctx, int d, int m, double* nu, double* Hplus, const double* L,
                         const double* Sigma, const double* v) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!L || !Sigma || !v || !nu || !Hplus) return BB_ERR_ARG;
  if (m < 1 || m > d) return BB_ERR_ASSERT_M;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  pk.add(nu, d); pk.add(Hplus, d * d); pk.add(L, m * d); pk.add(Sigma, m * m); pk.add(v, m);
  const size_t nout = d + d * d, oout = pk.host.size();
  int rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  std::vector<double> out;
  BB_DM_SWITCH(d, m, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                 k_gpupdate<D, M><<<1, 1, 0, ctx->stream>>>(pk.dev.p, pk.dev.p + oout, st);
               }));
  if (rc) return rc;
  memcpy(nu, out.data(), sizeof(double) * d);
  memcpy(Hplus, out.data() + d, sizeof(double) * d * d);
  return BB_OK;
}
extern "C" int bb_gpupdate_nuH(bb_ctx* ctx, int32_t d, int32_t m, double* nu, double* Hplus, const double* L,
                               const double* Sigma, const double* v) {
  return gpupdate_impl(ctx, d, m, nu, Hplus, L, Sigma, v);
}
extern "C" int bb_gpupdate_HV(bb_ctx* ctx, int32_t d, int32_t m, double* Hdia, double* V, const double* L,
                              const double* Sigma, const double* v) {
  return gpupdate_impl(ctx, d, m, V, Hdia, L, Sigma, v);
}

extern "C" int bb_backward_nuH(bb_ctx* ctx, int32_t method, int32_t N, int32_t d, const double* tt,
                               const bb_aux* aux, const double* nu_end, const double* Hplus_end, double C0,
                               double* nu, double* H, double* nu_left, double* Hplus_left, double* C) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!tt || !nu_end || !Hplus_end || !nu || !H || N < 2) return BB_ERR_ARG;
  if (method != BB_ODE_R3 && method != BB_ODE_LYAP) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  size_t oB, ob, oa, oal;
  const size_t ott = pk.add(tt, N);
  int rc = fill_aux(pk, aux, N, d, &oB, &ob, &oa, &oal);
  if (rc) return rc;
  const size_t one = pk.add(nu_end, d), ohe = pk.add(Hplus_end, d * d);
  const size_t nout = (size_t)N * d + (size_t)N * d * d + d + d * d + 1, oout = pk.host.size();
  rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  aux_dev A{D0 + oB, D0 + ob, D0 + oa, D0 + oal, aux->is_const};
  double* dnu = D0 + oout;
  double* dH = dnu + (size_t)N * d;
  double* dleft = dH + (size_t)N * d * d;
  std::vector<double> out;
  if (d > 3) {
    if (d > BB_MAXD_WIDE) return BB_ERR_UNSUPPORTED;
    rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
      bb_gen_backward_nuH(ctx->stream, method, N, d, D0 + ott, A.B, A.beta, A.a, A.a_left, A.is_const, D0 + one, D0 + ohe,
                          C0, dnu, dH, dleft, st);
    });
  } else
  BB_D_SWITCH(d, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                k_backward_nuH<D><<<1, 1, 0, ctx->stream>>>(method, N, D0 + ott, A, D0 + one, D0 + ohe, C0, dnu, dH,
                                                          dleft, st);
              }));
  if (rc) return rc;
  memcpy(nu, out.data(), sizeof(double) * N * d);
  memcpy(H, out.data() + (size_t)N * d, sizeof(double) * N * d * d);
  const double* left = out.data() + (size_t)N * d + (size_t)N * d * d;
  if (nu_left) memcpy(nu_left, left, sizeof(double) * d);
  if (Hplus_left) memcpy(Hplus_left, left + d, sizeof(double) * d * d);
  if (C) *C = left[d + d * d];
  return BB_OK;
}

extern "C" int bb_backward_FH(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const bb_aux* aux,
                              const double* F_end, const double* H_end, double C0, double* F, double* H,
                              double* C) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!tt || !F_end || !H_end || !F || !H || !C || N < 2) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  size_t oB, ob, oa, oal;
  const size_t ott = pk.add(tt, N);
  int rc = fill_aux(pk, aux, N, d, &oB, &ob, &oa, &oal);
  if (rc) return rc;
  const size_t ofe = pk.add(F_end, d), ohe = pk.add(H_end, d * d);
  const size_t nout = (size_t)N * d + (size_t)N * d * d + 1, oout = pk.host.size();
  rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  aux_dev A{D0 + oB, D0 + ob, D0 + oa, D0 + oal, aux->is_const};
  double* dF = D0 + oout;
  double* dH = dF + (size_t)N * d;
  double* dC = dH + (size_t)N * d * d;
  std::vector<double> out;
  BB_D_SWITCH(d, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                k_backward_FH<D><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + ofe, D0 + ohe, C0, dF, dH, dC, st);
              }));
  if (rc) return rc;
  memcpy(F, out.data(), sizeof(double) * N * d);
  memcpy(H, out.data() + (size_t)N * d, sizeof(double) * N * d * d);
  *C = out[(size_t)N * d + (size_t)N * d * d];
  return BB_OK;
}

extern "C" int bb_backward_HV(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const bb_aux* aux,
                              const double* v, const double* hdia_end, double* Hdia, double* V) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!tt || !v || !Hdia || !V || N < 2) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  size_t oB, ob, oa, oal;
  const size_t ott = pk.add(tt, N);
  int rc = fill_aux(pk, aux, N, d, &oB, &ob, &oa, &oal);
  if (rc) return rc;
  const size_t ov = pk.add(v, d);
  const size_t ohe = hdia_end ? pk.add(hdia_end, d * d) : 0;
  const size_t nout = (size_t)N * d * d + (size_t)N * d, oout = pk.host.size();
  rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  aux_dev A{D0 + oB, D0 + ob, D0 + oa, D0 + oal, aux->is_const};
  double* dHd = D0 + oout;
  double* dV = dHd + (size_t)N * d * d;
  std::vector<double> out;
  BB_D_SWITCH(d, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                k_backward_HV<D><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + ov, hdia_end ? D0 + ohe : nullptr, dHd,
                                                         dV, st);
              }));
  if (rc) return rc;
  memcpy(Hdia, out.data(), sizeof(double) * N * d * d);
  memcpy(V, out.data() + (size_t)N * d * d, sizeof(double) * N * d);
  return BB_OK;
}

extern "C" int bb_backward_LMmu(bb_ctx* ctx, int32_t N, int32_t d, int32_t m, const double* tt,
                                const bb_aux* aux, const double* L, const double* Sigma, double* Lt, double* Mt,
                                double* mut) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (!tt || !L || !Sigma || !Lt || !Mt || !mut || N < 2) return BB_ERR_ARG;
  if (m < 1 || m > d) return BB_ERR_ASSERT_M;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  size_t oB, ob, oa, oal;
  const size_t ott = pk.add(tt, N);
  int rc = fill_aux(pk, aux, N, d, &oB, &ob, &oa, &oal);
  if (rc) return rc;
  const size_t oL = pk.add(L, m * d), oS = pk.add(Sigma, m * m);
  const size_t nout = (size_t)N * (m * d + m * m + m), oout = pk.host.size();
  rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  aux_dev A{D0 + oB, D0 + ob, D0 + oa, D0 + oal, aux->is_const};
  double* dL = D0 + oout;
  double* dM = dL + (size_t)N * m * d;
  double* dmu = dM + (size_t)N * m * m;
  std::vector<double> out;
  BB_DM_SWITCH(d, m, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                 k_backward_LMmu<D, M><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + oL, D0 + oS, dL, dM, dmu, st);
               }));
  if (rc) return rc;
  memcpy(Lt, out.data(), sizeof(double) * N * m * d);
  memcpy(Mt, out.data() + (size_t)N * m * d, sizeof(double) * N * m * m);
  memcpy(mut, out.data() + (size_t)N * (m * d + m * m), sizeof(double) * N * m);
  return BB_OK;
}
