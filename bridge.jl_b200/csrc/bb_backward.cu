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
 * Operation order is the oracle's: reference arithmetic by default (bit-identical to liboracle_ref.so), explicit
 * fused multiply-adds with BB_ARITH_FUSED (liboracle_fma.so); see bb_ctx_set_arith.  d > 3 (bb_backward_gen.cu) is
 * fused order only.
 */
#include <math.h>
#include <string.h>

#include <vector>

#include "bb_backward.cuh"
#include "bb_host.h"

/* the kernels in both arithmetic flavours: bbk:: (explicit fma, the per-chain kernels' rounding order) and bbk_ref::
 * (reference arithmetic, bit-identical to liboracle_ref.so); bb_ctx::arith selects, see bb_ctx_set_arith */
#define BBK_NS bbk
#include "bb_backward_kernels.inc"
#undef BBK_NS
#define BBK_NS bbk_ref
#define BBK_MA(a, b, c) (((a) * (b)) + (c))
#include "bb_backward_body.inc"
#include "bb_backward_kernels.inc"
#undef BBK_NS
#undef BBK_MA

/* runs the launch statement with the kernels of the flavour the context asks for */
#define BB_ARITH_LAUNCH(ctx, ...)                                  \
  do {                                                             \
    if ((ctx)->arith == BB_ARITH_FUSED) { using namespace bbk; __VA_ARGS__; } \
    else { using namespace bbk_ref; __VA_ARGS__; }                 \
  } while (0)

/* d > 3: block-parallel generic-d kernels (bb_backward_gen.cu); all pointers are device pointers */
cudaError_t bb_gen_backward_nuH(cudaStream_t st, int method, int N, int d, const double* tt, const double* B,
                                const double* beta, const double* a, const double* a_left, int is_const,
                                const double* nu_end, const double* Hplus_end, double C0, double* nu, double* H,
                                double* out_left, int* status);
cudaError_t bb_gen_update_nuHC(cudaStream_t st, int d, int m, const double* in, double* out, int* status);

namespace {
using bbk_common::aux_dev;

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
  if (d < 1) return BB_ERR_ARG;
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
                 BB_ARITH_LAUNCH(ctx, k_update_nuHC<D, M><<<1, 1, 0, ctx->stream>>>(pk.dev.p, pk.dev.p + oout, st));
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
  if (d < 1) return BB_ERR_ARG;
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
                 BB_ARITH_LAUNCH(ctx, k_gpupdate<D, M><<<1, 1, 0, ctx->stream>>>(pk.dev.p, pk.dev.p + oout, st));
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
  if (d < 1) return BB_ERR_ARG;
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
                BB_ARITH_LAUNCH(ctx, k_backward_nuH<D><<<1, 1, 0, ctx->stream>>>(method, N, D0 + ott, A, D0 + one, D0 + ohe, C0, dnu, dH,
                                                          dleft, st));
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
  if (d < 1) return BB_ERR_ARG;
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
                BB_ARITH_LAUNCH(ctx, k_backward_FH<D><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + ofe, D0 + ohe, C0, dF, dH, dC, st));
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
  if (d < 1) return BB_ERR_ARG;
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
                BB_ARITH_LAUNCH(ctx, k_backward_HV<D><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + ov, hdia_end ? D0 + ohe : nullptr, dHd,
                                                         dV, st));
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
  if (d < 1) return BB_ERR_ARG;
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
                 BB_ARITH_LAUNCH(ctx, k_backward_LMmu<D, M><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, A, D0 + oL, D0 + oS, dL, dM, dmu, st));
               }));
  if (rc) return rc;
  memcpy(Lt, out.data(), sizeof(double) * N * m * d);
  memcpy(Mt, out.data() + (size_t)N * m * d, sizeof(double) * N * m * m);
  memcpy(mut, out.data() + (size_t)N * (m * d + m * m), sizeof(double) * N * m);
  return BB_OK;
}

/* lptilde: see the kernels in bb_backward_kernels.inc */
extern "C" int bb_lptilde_nuH(bb_ctx* ctx, int32_t d, const double* nu0, const double* H0, double C, const double* x,
                              double* out) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (d < 1) return BB_ERR_ARG;
  if (!nu0 || !H0 || !x || !out) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  pk.add(nu0, d); pk.add(H0, d * d); pk.add(&C, 1); pk.add(x, d);
  const size_t oout = pk.host.size();
  int rc = pack_upload(ctx, pk, 1);
  if (rc) return rc;
  std::vector<double> o;
  BB_D_SWITCH(d, rc = run_small(ctx, pk, oout, 1, o, [&](int* st) {
                BB_ARITH_LAUNCH(ctx, k_lptilde_nuH<D><<<1, 1, 0, ctx->stream>>>(pk.dev.p, pk.dev.p + oout, st));
              }));
  if (rc) return rc;
  *out = o[0];
  return BB_OK;
}
extern "C" int bb_lptilde_HV(bb_ctx* ctx, int32_t N, int32_t d, const double* tt, const double* trB, int32_t trB_const,
                             const double* V0, const double* Hdia0, const double* u, double* out) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (d < 1) return BB_ERR_ARG;
  if (!tt || !trB || !V0 || !Hdia0 || !u || !out || N < 2) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  dev_pack pk;
  const size_t ott = pk.add(tt, N);
  const size_t otr = pk.add(trB, trB_const ? 1 : (size_t)(N - 1) * 3);
  const size_t oin = pk.add(V0, d);
  pk.add(Hdia0, d * d); pk.add(u, d);
  const size_t oout = pk.host.size();
  int rc = pack_upload(ctx, pk, 1);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  std::vector<double> o;
  const double log2pi = log(2 * M_PI);
  BB_D_SWITCH(d, rc = run_small(ctx, pk, oout, 1, o, [&](int* st) {
                BB_ARITH_LAUNCH(ctx, k_lptilde_HV<D><<<1, 1, 0, ctx->stream>>>(N, D0 + ott, D0 + otr, trB_const, D0 + oin,
                                                                             log2pi, D0 + oout, st));
              }));
  if (rc) return rc;
  *out = o[0];
  return BB_OK;
}

/* ---- one launch for the backward pass of a whole chain of segments; tables written in place on the device */
extern "C" int bb_guides_chain_nuH(bb_ctx* ctx, int32_t method, int32_t S, int32_t N, int32_t d, int32_t m,
                                   const double* tt, const bb_aux* aux, const double* L, const double* Sigma,
                                   const double* v, double eps, bb_guide** guides, double* nu_left,
                                   double* Hplus_left, double* C) {
  if (!ctx) return BB_ERR_NODEVICE;
  if (d < 1) return BB_ERR_ARG;
  if (!tt || !aux || !L || !Sigma || !v || !guides || S < 1 || S > BB_MAXSEG || N < 2) return BB_ERR_ARG;
  if (method != BB_ODE_R3 && method != BB_ODE_LYAP) return BB_ERR_ARG;
  if (m < 1 || m > d) return BB_ERR_ASSERT_M;
  if (d > 3) return BB_ERR_UNSUPPORTED;
  for (int s = 0; s < S; s++)
    if (aux[s].d != d || !aux[s].is_const || !aux[s].B || !aux[s].beta || !aux[s].a) return BB_ERR_UNSUPPORTED;
  BB_CUDA(cudaSetDevice(ctx->device));
  /* guides that do not exist yet are created (time grid rows, auxiliary drift); existing ones must match */
  std::vector<double> zero((size_t)N * d * d, 0.0);
  for (int s = 0; s < S; s++) {
    if (!guides[s]) {
      int rc = bb_guide_create(ctx, BB_GUIDE_NUH, N, d, 0, tt + (size_t)s * N, zero.data(), zero.data(), nullptr, nullptr,
                               aux[s].B, aux[s].beta, 1, &guides[s]);
      if (rc) return rc;
    } else {
      bb_guide* g = guides[s];
      if (g->ctx != ctx || g->kind != BB_GUIDE_NUH || g->N != N || g->d != d || g->auxc != 1) return BB_ERR_ARG;
      memcpy(g->segc, aux[s].B, sizeof(double) * d * d);          /* the auxiliary drift of this parameter value */
      memcpy(g->segc + d * d, aux[s].beta, sizeof(double) * d);
      g->tt.assign(tt + (size_t)s * N, tt + (size_t)(s + 1) * N);
    }
  }
  dev_pack pk;
  const size_t ott = pk.add(tt, (size_t)S * N);
  const size_t oin = pk.add(L, m * d);
  pk.add(Sigma, m * m); pk.add(&eps, 1);
  for (int s = 0; s < S; s++) {
    pk.add(v + (size_t)s * m, m); pk.add(aux[s].B, d * d); pk.add(aux[s].beta, d); pk.add(aux[s].a, d * d);
  }
  const size_t nout = d + d * d + 1, oout = pk.host.size();
  int rc = pack_upload(ctx, pk, nout);
  if (rc) return rc;
  double* D0 = pk.dev.p;
  bbk_ref::chain_tabs T;
  bbk::chain_tabs Tf;
  memset(&T, 0, sizeof(T)); memset(&Tf, 0, sizeof(Tf));
  for (int s = 0; s < S; s++) { T.tab[s] = guides[s]->tab; Tf.tab[s] = guides[s]->tab; }
  const int rec = guides[0]->rec;
  std::vector<double> out;
  BB_DM_SWITCH(d, m, rc = run_small(ctx, pk, oout, nout, out, [&](int* st) {
                 if (ctx->arith == BB_ARITH_FUSED)
                   bbk::k_chain_nuH<D, M><<<1, 1, 0, ctx->stream>>>(method, S, N, rec, D0 + ott, D0 + oin, Tf, D0 + oout, st);
                 else
                   bbk_ref::k_chain_nuH<D, M><<<1, 1, 0, ctx->stream>>>(method, S, N, rec, D0 + ott, D0 + oin, T, D0 + oout, st);
               }));
  if (rc) return rc;
  if (nu_left) memcpy(nu_left, out.data(), sizeof(double) * d);
  if (Hplus_left) memcpy(Hplus_left, out.data() + d, sizeof(double) * d * d);
  if (C) *C = out[d + d * d];
  return BB_OK;
}

/* ν[N][d], H[N][d][d] (NUH) / V, H♢⁻¹ ... as the path kernels hold them: values on the grid read back from the device
 * table of a guide (rows 1 .. N-1 carry the values at grid points 0 .. N-2; the terminal values are not stored) */
extern "C" int bb_guide_download_nuH(bb_guide* g, double* nu, double* H) {
  if (!g || !nu || !H) return BB_ERR_ARG;
  if (g->kind != BB_GUIDE_NUH) return BB_ERR_UNSUPPORTED;
  bb_ctx* ctx = g->ctx;
  BB_CUDA(cudaSetDevice(ctx->device));
  const int N = g->N, d = g->d, rec = g->rec;
  std::vector<double> tab((size_t)N * rec);
  BB_CUDA(cudaMemcpyAsync(tab.data(), g->tab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < N - 1; i++) {
    const double* R = tab.data() + (size_t)(i + 1) * rec;
    memcpy(nu + (size_t)i * d, R + 2, sizeof(double) * d);
    memcpy(H + (size_t)i * d * d, R + 2 + d, sizeof(double) * d * d);
  }
  for (int q = 0; q < d; q++) nu[(size_t)(N - 1) * d + q] = NAN;
  for (int q = 0; q < d * d; q++) H[(size_t)(N - 1) * d * d + q] = NAN;
  return BB_OK;
}
