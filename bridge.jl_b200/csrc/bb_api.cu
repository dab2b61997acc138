/*
 * bb_api.cu -- the C ABI of include/bridge_b200.h: contexts, ensembles, guiding tables and the
 * launches of the path kernel (bb_chain.cuh).  Host code only prepares tables and launches; there
 * is no CPU implementation of any compute entry point (BB_ERR_NODEVICE without a GPU).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "bb_host.h"

/* ------------------------------------------------------------------------------------------------ errors */
static thread_local char g_cuda_err[512] = "";
void bb_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", cudaGetErrorName(e), cudaGetErrorString(e), where);
}
extern "C" const char* bb_last_cuda_error(void) { return g_cuda_err; }
extern "C" int bb_abi_version(void) { return BB_ABI_VERSION; }
extern "C" const char* bb_strerror(int status) {
  switch (status) {
    case BB_OK: return "ok";
    case BB_ERR_LENGTH: return "Y and W differ in length.";
    case BB_ERR_TIMEAXIS: return "Time axis mismatch between bridge P and driving W.";
    case BB_ERR_STARTPOINT: return "Starting point has wrong length.";
    case BB_ERR_DIM: return "DimensionMismatch(\"length(tt) != size(yy, 2)\")";
    case BB_ERR_ASSERT_M: return "AssertionError: m == length(v)";
    case BB_ERR_MODEL: return "unknown model id, or model and ensemble dimensions differ";
    case BB_ERR_ARG: return "invalid argument";
    case BB_ERR_CUDA: return "CUDA runtime error (see bb_last_cuda_error)";
    case BB_ERR_NOMEM: return "device memory allocation failed";
    case BB_ERR_NODEVICE: return "no CUDA device: libbridge_b200 has no CPU path";
    case BB_ERR_UNSUPPORTED: return "combination of model, guide and dimensions is not instantiated";
    case BB_ERR_SINGULAR: return "singular matrix";
    case BB_ERR_STALE: return "X holds rejected proposals for some chains: call bb_ens_refresh_x first";
    case BB_ERR_COMM: return "NCCL error (see bb_comm_last_error)";
    case BB_ERR_USERSRC: return "the source of the user-defined model does not compile (see bb_user_model_log)";
    default: return "unknown status";
  }
}

/* ------------------------------------------------------------------------------------------------ context */
extern "C" int bb_ctx_create(int device, bb_ctx** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    if (e != cudaSuccess) bb_set_cuda_error(e, "cudaGetDeviceCount");
    return BB_ERR_NODEVICE;
  }
  if (device < 0 || device >= n) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(device));
  if (const char* g = getenv("BB_L2_FETCH")) { /* tuning aid: L2 -> DRAM fetch granularity hint (32/64/128) */
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "[bb] cudaLimitMaxL2FetchGranularity = %zu\n", got);
  }
  bb_ctx* c = new (std::nothrow) bb_ctx();
  if (!c) return BB_ERR_NOMEM;
  c->device = device;
  cudaError_t ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev1);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (ce != cudaSuccess) {
    bb_set_cuda_error(ce, "bb_ctx_create");
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return ce == cudaErrorMemoryAllocation ? BB_ERR_NOMEM : BB_ERR_CUDA;
  }
  *out = c;
  return BB_OK;
}
extern "C" int bb_ctx_destroy(bb_ctx* c) {
  if (!c) return BB_ERR_ARG;
  /* guides and ensembles keep a pointer to their context (and guides borrow its recycled buffers): destroy them first */
  if (c->live > 0) return BB_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->stage) cudaFree(c->stage);
  for (auto& kv : c->pool_size) cudaFree(kv.first); /* recycled buffers, free or still lent out to live guides */
  c->pool_size.clear();
  c->pool_free.clear();
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c;
  return BB_OK;
}
extern "C" int bb_ctx_synchronize(bb_ctx* c) {
  if (!c) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(c->device));
  BB_CUDA(cudaStreamSynchronize(c->stream));
  return BB_OK;
}
extern "C" int bb_ctx_set_stream(bb_ctx* c, void* s) {
  if (!c) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(c->device));
  BB_CUDA(cudaStreamSynchronize(c->stream));
  if (s) {
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
  } else if (!c->own_stream) {
    BB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return BB_OK;
}
extern "C" void* bb_ctx_get_stream(bb_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int64_t bb_ctx_launch_count(bb_ctx* c) { return c ? c->launches : -1; }
extern "C" int bb_ctx_set_timing(bb_ctx* c, int on) {
  if (!c) return BB_ERR_ARG;
  c->timing = on != 0;
  c->ev_valid = false;
  return BB_OK;
}
extern "C" int bb_ctx_set_arith(bb_ctx* c, int arith) {
  if (!c || (arith != BB_ARITH_REFERENCE && arith != BB_ARITH_FUSED)) return BB_ERR_ARG;
  c->arith = arith;
  return BB_OK;
}
#ifndef BB_AUTO_WS2
#define BB_AUTO_WS2 0 /* automatic selection prefers the two-chains-per-dynamics-thread kernel for small ensembles */
#endif
extern "C" int bb_ctx_set_pcn_kernel(bb_ctx* c, int mode) {
  if (!c || mode < BB_PCN_AUTO || mode > BB_PCN_WARP_SPECIALISED_2) return BB_ERR_ARG;
  c->pcn_kernel = mode;
  return BB_OK;
}
extern "C" double bb_ctx_last_kernel_ms(bb_ctx* c) {
  if (!c || !c->ev_valid) return -1.0;
  cudaSetDevice(c->device);
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}
void bb_time_begin(bb_ctx* c) {
  if (c->timing) cudaEventRecord(c->ev0, c->stream);
}
void bb_time_end(bb_ctx* c) {
  if (c->timing) {
    cudaEventRecord(c->ev1, c->stream);
    c->ev_valid = true;
  }
}

/* ---- recycled small device buffers */
static const size_t BB_POOL_MAX_BLOCK = (size_t)32 << 20; /* larger requests are not kept */
cudaError_t bb_pool_alloc(bb_ctx* c, size_t bytes, void** out) {
  const size_t want = (bytes + 4095) & ~(size_t)4095;
  for (size_t i = 0; i < c->pool_free.size(); i++) {
    if (c->pool_free[i].first >= want && c->pool_free[i].first <= 2 * want) {
      *out = c->pool_free[i].second;
      c->pool_free.erase(c->pool_free.begin() + i);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, want);
  if (e == cudaSuccess) c->pool_size[*out] = want;
  return e;
}
void bb_pool_release(bb_ctx* c, void* p) {
  if (!p) return;
  auto it = c->pool_size.find(p);
  if (it == c->pool_size.end()) { cudaFree(p); return; }
  if (it->second > BB_POOL_MAX_BLOCK || c->pool_free.size() >= 64) {
    cudaFree(p);
    c->pool_size.erase(it);
    return;
  }
  c->pool_free.push_back({it->second, p});
}

static int ctx_stage(bb_ctx* c, size_t bytes) {
  if (c->stage_bytes >= bytes) return BB_OK;
  if (c->stage) {
    BB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->stage);
    c->stage = nullptr;
    c->stage_bytes = 0;
  }
  BB_CUDA(cudaMalloc(&c->stage, bytes));
  c->stage_bytes = bytes;
  return BB_OK;
}

/* ------------------------------------------------------------------------------------------------ host algebra
 * (oracle operation order: plain products/sums, closed-form inverses for d <= 3, Gauss-Jordan above) */
int bb_h_inv(int d, const double* A, double* Ai) {
  if (d == 1) {
    if (A[0] == 0.0) return -1;
    Ai[0] = 1.0 / A[0];
    return 0;
  }
  if (d == 2) {
    double det = A[0] * A[3] - A[1] * A[2];
    if (det == 0.0) return -1;
    double id = 1.0 / det;
    double r0 = A[3] * id, r1 = -A[1] * id, r2 = -A[2] * id, r3 = A[0] * id;
    Ai[0] = r0; Ai[1] = r1; Ai[2] = r2; Ai[3] = r3;
    return 0;
  }
  if (d == 3) {
    double c00 = A[4] * A[8] - A[5] * A[7];
    double c01 = A[5] * A[6] - A[3] * A[8];
    double c02 = A[3] * A[7] - A[4] * A[6];
    double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    if (det == 0.0) return -1;
    double id = 1.0 / det;
    double T[9];
    T[0] = c00 * id;
    T[1] = (A[2] * A[7] - A[1] * A[8]) * id;
    T[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    T[3] = c01 * id;
    T[4] = (A[0] * A[8] - A[2] * A[6]) * id;
    T[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    T[6] = c02 * id;
    T[7] = (A[1] * A[6] - A[0] * A[7]) * id;
    T[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    memcpy(Ai, T, sizeof(T));
    return 0;
  }
  if (d > BB_MAXD) return -1;
  double M[BB_MAXD][2 * BB_MAXD];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      M[i][j] = A[i * d + j];
      M[i][d + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < d; c++) {
    int p = c;
    for (int i = c + 1; i < d; i++)
      if (fabs(M[i][c]) > fabs(M[p][c])) p = i;
    if (M[p][c] == 0.0) return -1;
    if (p != c)
      for (int j = 0; j < 2 * d; j++) {
        double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t;
      }
    double ip = 1.0 / M[c][c];
    for (int j = 0; j < 2 * d; j++) M[c][j] *= ip;
    for (int i = 0; i < d; i++)
      if (i != c) {
        double f = M[i][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * d; j++) M[i][j] -= f * M[c][j];
      }
  }
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) Ai[i * d + j] = M[i][d + j];
  return 0;
}

/* ------------------------------------------------------------------------------------------------ models */
static bool model_shape_ok(const bb_model* m) {
  const int d = m->d, dp = m->dprime;
  switch (m->id) {
    case BB_MODEL_WIENER: return d == dp && ((d >= 1 && d <= 3) || d == 8);
    case BB_MODEL_LANDMARKS: return d == 16 && dp == 8;
    case BB_MODEL_BOLUS: return d == 2 && dp == 2; /* per-chain-parameter path only: no shared-table kernel */
    case BB_MODEL_OU: return d == 1 && dp == 1;
    case BB_MODEL_LINPRO: return d == dp && d >= 1 && d <= 3;
    case BB_MODEL_FHN_DIAG: return d == 2 && dp == 2;
    case BB_MODEL_FHN_HYPO: return d == 2 && dp == 1;
    case BB_MODEL_INTDIFF: return d == 2 && dp == 1;
    case BB_MODEL_NCLAR3: return d == 3 && dp == 1;
    case BB_MODEL_LORENZ: return d == 3 && dp == 3;
    case BB_MODEL_USER: return d >= 1 && d <= 3 && dp >= 1 && dp <= d && bb_user_lookup(m->reserved) != nullptr;
    default: return false;
  }
}
/* sigma as a dense d x d' matrix (oracle model_sigma) */
static void model_sigma_host(const bb_model* P, double* S) {
  const int d = P->d, dp = P->dprime;
  const double* p = P->par;
  memset(S, 0, sizeof(double) * d * dp);
  switch (P->id) {
    case BB_MODEL_WIENER: for (int i = 0; i < d; i++) S[i * dp + i] = 1.0; break;
    case BB_MODEL_OU: S[0] = p[1]; break;
    case BB_MODEL_LINPRO: memcpy(S, p + d * d + d, sizeof(double) * d * d); break;
    case BB_MODEL_FHN_DIAG: S[0] = p[4]; S[3] = p[5]; break;
    case BB_MODEL_FHN_HYPO: S[1] = p[4]; break;
    case BB_MODEL_INTDIFF: S[1] = p[0]; break;
    case BB_MODEL_NCLAR3: S[2] = p[2]; break;
    case BB_MODEL_LORENZ: S[0] = p[3]; S[4] = p[4]; S[8] = p[5]; break;
    case BB_MODEL_BOLUS: S[0] = p[4]; S[3] = p[4]; break;
  }
}
void bb_prepare_model(const bb_model* m, bb_model_dev* o) {
  memset(o, 0, sizeof(*o));
  memcpy(o->par, m->par, sizeof(o->par));
  if (m->id == BB_MODEL_LANDMARKS) { /* partialbridge_landmarks.jl:47,96-98 (oracle model_b) */
    const double a = m->par[0];
    o->der[0] = 1.0 / ((2 * M_PI) * a);
    o->der[1] = 1.0 / (2 * a);
    o->der[2] = 0.0 + m->par[1] * m->par[1];
    o->der[3] = (-m->par[2]) * 0.5;
    return;
  }
  if (m->d > BB_MAXD) return; /* the d' = 8 Wiener process: nothing to derive */
  if (m->id == BB_MODEL_FHN_DIAG || m->id == BB_MODEL_FHN_HYPO) o->der[0] = 1.0 / m->par[0];
  double S[BB_MAXD * BB_MAXD];
  model_sigma_host(m, S);
  const int d = m->d, dp = m->dprime;
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      double s = 0.0;
      for (int l = 0; l < dp; l++) s += S[i * dp + l] * S[j * dp + l];
      o->der[8 + i * d + j] = s;
    }
  if (d == dp) { /* inv(sigma) for innovations!  (src/euler.jl:372); stays zero if sigma is singular */
    double Si[BB_MAXD * BB_MAXD];
    if (bb_h_inv(d, S, Si) == 0) memcpy(o->der + 24, Si, sizeof(double) * d * d);
  }
}
static bool model_sigma_invertible(const bb_model* m) {
  if (m->d != m->dprime || m->d > BB_MAXD) return false;
  double S[BB_MAXD * BB_MAXD], Si[BB_MAXD * BB_MAXD];
  model_sigma_host(m, S);
  return bb_h_inv(m->d, S, Si) == 0;
}
static bb_chain_launch_fn lookup_second(const bb_model* m, int gk, int gm, int auxc, int mode) {
  switch (m->id) {
    case BB_MODEL_WIENER: return gk == 0 ? bb_lookup2_wiener(m->d, mode) : nullptr;
    case BB_MODEL_OU: return bb_lookup2_ou(gk, gm, auxc, mode);
    case BB_MODEL_LINPRO:
      return m->d == 1 ? bb_lookup2_linpro1(gk, gm, auxc, mode)
                       : (m->d == 2 ? bb_lookup2_linpro2(gk, gm, auxc, mode) : bb_lookup2_linpro3(gk, gm, auxc, mode));
    case BB_MODEL_FHN_DIAG: return bb_lookup2_fhn_diag(gk, gm, auxc, mode);
    case BB_MODEL_FHN_HYPO: return bb_lookup2_fhn_hypo(gk, gm, auxc, mode);
    case BB_MODEL_INTDIFF: return bb_lookup2_intdiff(gk, gm, auxc, mode);
    case BB_MODEL_NCLAR3: return bb_lookup2_nclar3(gk, gm, auxc, mode);
    case BB_MODEL_LORENZ: return bb_lookup2_lorenz(gk, gm, auxc, mode);
    default: return nullptr;
  }
}
static bb_chain_launch_fn lookup_kernel(const bb_model* m, int gk, int gm, int auxc, int rng) {
  switch (m->id) {
    case BB_MODEL_WIENER:
      if (gk != 0) return nullptr;
      return m->d > 3 ? bb_lookup_wiener_wide(m->d, rng) : bb_lookup_wiener(m->d, rng);
    case BB_MODEL_LANDMARKS: return bb_lookup_landmarks(gk, gm, auxc, rng);
    case BB_MODEL_OU: return bb_lookup_ou(gk, gm, auxc, rng);
    case BB_MODEL_LINPRO:
      return m->d == 1 ? bb_lookup_linpro1(gk, gm, auxc, rng)
                       : (m->d == 2 ? bb_lookup_linpro2(gk, gm, auxc, rng) : bb_lookup_linpro3(gk, gm, auxc, rng));
    case BB_MODEL_FHN_DIAG: return bb_lookup_fhn_diag(gk, gm, auxc, rng);
    case BB_MODEL_FHN_HYPO: return bb_lookup_fhn_hypo(gk, gm, auxc, rng);
    case BB_MODEL_INTDIFF: return bb_lookup_intdiff(gk, gm, auxc, rng);
    case BB_MODEL_NCLAR3: return bb_lookup_nclar3(gk, gm, auxc, rng);
    case BB_MODEL_LORENZ: return bb_lookup_lorenz(gk, gm, auxc, rng);
    default: return nullptr;
  }
}

/* ------------------------------------------------------------------------------------------------ ensemble */
static inline int64_t ens_rows(const bb_ens* e) { return (int64_t)e->S * e->NC * e->P; }

extern "C" int bb_ens_destroy(bb_ens* e) {
  if (!e) return BB_ERR_ARG;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(e->ctx->stream);
  if (e->W[0]) cudaFree(e->W[0]);
  if (e->X) cudaFree(e->X);
  if (e->mc_sum) cudaFree(e->mc_sum);
  if (e->cmc_m) cudaFree(e->cmc_m);
  bb_theta_free(e);
  void* ptrs[] = {e->par, e->accepted, e->xstale, e->ll, e->llprop, e->logu, e->xend, e->xendprop, e->acc, e->start};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (double* g : e->gridtab)
    if (g) cudaFree(g);
  e->ctx->live--;
  delete e;
  return BB_OK;
}

template <class T>
static int dev_alloc(bb_ens* e, T** p, size_t count) {
  BB_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  BB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), e->ctx->stream));
  e->bytes += (int64_t)(count * sizeof(T));
  return BB_OK;
}

extern "C" int bb_ens_create(bb_ctx* ctx, int64_t P, int32_t S, int32_t N, int32_t d, int32_t dprime,
                             uint32_t flags, bb_ens** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  if (!ctx) return BB_ERR_NODEVICE;
  if (P <= 0 || S <= 0 || S > BB_MAXSEG || N < 2 || d < 1 || d > BB_MAXD_WIDE || dprime < 1 || dprime > d) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(ctx->device));
  bb_ens* e = new (std::nothrow) bb_ens();
  if (!e) return BB_ERR_NOMEM;
  e->ctx = ctx; e->P = P; e->S = S; e->N = N; e->d = d; e->dp = dprime; e->flags = flags;
  ctx->live++;
  e->NC = (N + BB_TC - 1) / BB_TC;
  const int nbuf = (flags & BB_ENS_DOUBLE_BUFFER) ? 2 : 1;
  const size_t rows = (size_t)ens_rows(e);
  int rc = BB_OK;
  /* the two buffers of a chain are adjacent rows of one allocation (see bb_chain.cuh) */
  e->nbuf = nbuf;
  rc = dev_alloc(e, &e->W[0], rows * nbuf * BB_TC * dprime);
  if (rc == BB_OK && !(flags & BB_ENS_NO_X)) rc = dev_alloc(e, &e->X, rows * BB_TC * d);
  e->W[1] = (nbuf == 2 && e->W[0]) ? e->W[0] + BB_TC * dprime : e->W[0];
  if (rc == BB_OK) rc = dev_alloc(e, &e->par, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->accepted, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->xstale, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->ll, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->llprop, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->logu, (size_t)P);
  if (rc == BB_OK) rc = dev_alloc(e, &e->xend, (size_t)P * d);
  if (rc == BB_OK) rc = dev_alloc(e, &e->xendprop, (size_t)P * d);
  if (rc == BB_OK) rc = dev_alloc(e, &e->acc, 1);
  if (rc == BB_OK) rc = dev_alloc(e, &e->start, (size_t)d);
  e->gridtab.assign(S, nullptr);
  e->tt.assign(S, std::vector<double>());
  if (rc != BB_OK) {
    bb_ens_destroy(e);
    return rc;
  }
  *out = e;
  return BB_OK;
}

extern "C" int bb_ens_set_chain_offset(bb_ens* e, int64_t off) {
  if (!e || off < 0) return BB_ERR_ARG;
  e->chain_offset = off;
  return BB_OK;
}

/* table of a bare time grid: row j = (tt[j]-tt[j-1], sqrt of it), row 0 and padding rows are zero */
static void grid_rows(const double* tt, int N, int NC, int rec, std::vector<double>& tab) {
  tab.assign((size_t)NC * BB_TC * rec, 0.0);
  for (int j = 1; j < N; j++) {
    const double dt = tt[j] - tt[j - 1];
    tab[(size_t)j * rec] = dt;
    tab[(size_t)j * rec + 1] = sqrt(dt);
  }
}

extern "C" int bb_ens_set_grid(bb_ens* e, int32_t seg, const double* tt, int32_t n) {
  if (!e || !tt || seg < 0 || seg >= e->S) return BB_ERR_ARG;
  if (n != e->N) return BB_ERR_LENGTH;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  std::vector<double> tab;
  grid_rows(tt, e->N, e->NC, 2, tab);
  if (!e->gridtab[seg]) {
    BB_CUDA(cudaMalloc(&e->gridtab[seg], tab.size() * sizeof(double)));
    e->bytes += (int64_t)(tab.size() * sizeof(double));
  }
  BB_CUDA(cudaMemcpyAsync(e->gridtab[seg], tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice,
                          e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  e->tt[seg].assign(tt, tt + n);
  bb_theta_invalidate(e);
  return BB_OK;
}
extern "C" int bb_ens_get_grid(bb_ens* e, int32_t seg, double* tt, int32_t n) {
  if (!e || !tt || seg < 0 || seg >= e->S) return BB_ERR_ARG;
  if (n != e->N || (int)e->tt[seg].size() != n) return BB_ERR_LENGTH;
  memcpy(tt, e->tt[seg].data(), sizeof(double) * n);
  return BB_OK;
}

extern "C" int bb_ens_set_start(bb_ens* e, const double* u, int32_t n_u, int32_t broadcast) {
  if (!e || !u) return BB_ERR_ARG;
  bb_theta_invalidate(e); /* the left-end Gaussian term of the per-chain path depends on x0 */
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int d = e->d;
  if (broadcast) {
    if (n_u != d) return BB_ERR_STARTPOINT;
    if (!e->start_bcast) {
      BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
      double* fresh = nullptr;
      BB_CUDA(cudaMalloc(&fresh, sizeof(double) * d)); /* on failure the old per-chain array stays valid */
      cudaFree(e->start);
      e->start = fresh;
      e->start_bcast = 1;
    }
    BB_CUDA(cudaMemcpyAsync(e->start, u, sizeof(double) * d, cudaMemcpyHostToDevice, e->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
    return BB_OK;
  }
  if ((int64_t)n_u != e->P * d) return BB_ERR_STARTPOINT;
  std::vector<double> t((size_t)e->P * d);
  for (int64_t p = 0; p < e->P; p++)
    for (int k = 0; k < d; k++) t[(size_t)k * e->P + p] = u[(size_t)p * d + k];
  if (e->start_bcast) {
    BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
    double* fresh = nullptr;
    BB_CUDA(cudaMalloc(&fresh, sizeof(double) * e->P * d)); /* on failure the broadcast point stays valid */
    cudaFree(e->start);
    e->start = fresh;
    e->bytes += (int64_t)(sizeof(double) * e->P * d);
    e->start_bcast = 0;
  }
  BB_CUDA(cudaMemcpyAsync(e->start, t.data(), sizeof(double) * e->P * d, cudaMemcpyHostToDevice, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return BB_OK;
}

/* ---- host [np][S][N][K] <-> device [S][NC][P][8][K]; which buffer a chain uses is decided per chain */
template <bool TO_DEVICE>
__global__ void __launch_bounds__(256) bb_transpose_kernel(double* __restrict__ buf0, double* __restrict__ buf1,
                                                           double* __restrict__ stage,
                                                           const uint8_t* __restrict__ par,
                                                           const uint8_t* __restrict__ accepted, int which,
                                                           long long P, long long p0, long long np, int S, int N,
                                                           int NC, int K, int nbuf) {
  const long long rowlen = (long long)BB_TC * K;
  const long long total = (long long)S * NC * np * rowlen;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(t % rowlen);
    long long q = t / rowlen;
    const long long pl = q % np;
    q /= np;
    const int c = (int)(q % NC);
    const int s = (int)(q / NC);
    const int slot = e / K, k = e - slot * K;
    const int j = c * BB_TC + slot;
    if (j >= N) continue;
    const long long p = p0 + pl;
    int b = par[p];
    if (which == BB_PROP && !accepted[p]) b = 1 - b;
    double* dev = (b ? buf1 : buf0) + (((long long)s * NC + c) * P + p) * (rowlen * nbuf) + e;
    double* hst = stage + ((pl * S + s) * (long long)N + j) * K + k;
    if (TO_DEVICE) *dev = *hst;
    else *hst = *dev;
  }
}

static int ens_transfer(bb_ens* e, int what, int which, int64_t p0, int64_t np, double* host, bool to_device) {
  if (!e || !host || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (what != BB_W && what != BB_X) return BB_ERR_ARG;
  if (which != BB_CUR && which != BB_PROP) return BB_ERR_ARG;
  if (what == BB_X && !e->X) return BB_ERR_ARG;
  /* X is single buffered and holds the last proposal of every chain; the CURRENT path of a chain whose
   * proposal was rejected has to be recomputed first (bb_ens_refresh_x) */
  if (what == BB_X && which == BB_CUR && !to_device && e->x_maybe_stale) return BB_ERR_STALE;
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  const int K = what == BB_W ? e->dp : e->d;
  double* b0 = what == BB_W ? e->W[0] : e->X;
  double* b1 = what == BB_W ? e->W[1] : e->X;
  const int nb = what == BB_W ? e->nbuf : 1;
  const size_t per_chain = (size_t)e->S * e->N * K;
  int64_t slab = (int64_t)((size_t)(256u << 20) / (per_chain * sizeof(double)));
  if (slab < 1) slab = 1;
  if (slab > np) slab = np;
  if (np == 0) return BB_OK;
  int rc = ctx_stage(c, (size_t)slab * per_chain * sizeof(double));
  if (rc != BB_OK) return rc;
  for (int64_t q0 = 0; q0 < np; q0 += slab) {
    const int64_t n = (np - q0 < slab) ? np - q0 : slab;
    double* h = host + (size_t)q0 * per_chain;
    const long long total = (long long)e->S * e->NC * n * BB_TC * K;
    const unsigned grid = (unsigned)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    if (to_device) {
      BB_CUDA(cudaMemcpyAsync(c->stage, h, (size_t)n * per_chain * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      bb_transpose_kernel<true><<<grid, 256, 0, c->stream>>>(b0, b1, c->stage, e->par, e->accepted, which, e->P,
                                                            p0 + q0, n, e->S, e->N, e->NC, K, nb);
    } else {
      bb_transpose_kernel<false><<<grid, 256, 0, c->stream>>>(b0, b1, c->stage, e->par, e->accepted, which, e->P,
                                                             p0 + q0, n, e->S, e->N, e->NC, K, nb);
      BB_CUDA(cudaMemcpyAsync(h, c->stage, (size_t)n * per_chain * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    BB_CUDA(cudaGetLastError());
    c->launches++;
    /* the staging buffer is reused by the next slab */
    BB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return BB_OK;
}
extern "C" int bb_ens_upload(bb_ens* e, int what, int which, int64_t p0, int64_t np, const double* host) {
  const int rc = ens_transfer(e, what, which, p0, np, const_cast<double*>(host), true);
  if (rc == BB_OK && what == BB_X && np > 0) {
    /* a path supplied by the caller IS the chains' path now: it must not be recomputed from W by a later refresh
     * (sample! followed by llikelihood on an uploaded X would otherwise evaluate a different path) */
    BB_CUDA(cudaMemsetAsync(e->xstale + p0, 0, (size_t)np, e->ctx->stream));
    if (p0 == 0 && np == e->P) e->x_maybe_stale = false;
  }
  return rc;
}
extern "C" int bb_ens_download(bb_ens* e, int what, int which, int64_t p0, int64_t np, double* host) {
  return ens_transfer(e, what, which, p0, np, host, false);
}

extern "C" int bb_ens_get_f64(bb_ens* e, int field, int64_t p0, int64_t np, double* host) {
  if (!e || !host || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  cudaStream_t st = e->ctx->stream;
  const double* src = nullptr;
  switch (field) {
    case BB_F_LL: src = e->ll; break;
    case BB_F_LL_PROP: src = e->llprop; break;
    case BB_F_LOGU: src = e->logu; break;
    case BB_F_XEND: src = e->xend; break;
    case BB_F_XEND_PROP: src = e->xendprop; break;
    default: return BB_ERR_ARG;
  }
  if (field == BB_F_XEND || field == BB_F_XEND_PROP) {
    const int d = e->d;
    std::vector<double> t((size_t)np * d);
    for (int k = 0; k < d; k++)
      BB_CUDA(cudaMemcpyAsync(t.data() + (size_t)k * np, src + (size_t)k * e->P + p0, sizeof(double) * np,
                              cudaMemcpyDeviceToHost, st));
    BB_CUDA(cudaStreamSynchronize(st));
    for (int64_t p = 0; p < np; p++)
      for (int k = 0; k < d; k++) host[(size_t)p * d + k] = t[(size_t)k * np + p];
    return BB_OK;
  }
  BB_CUDA(cudaMemcpyAsync(host, src + p0, sizeof(double) * np, cudaMemcpyDeviceToHost, st));
  BB_CUDA(cudaStreamSynchronize(st));
  return BB_OK;
}
extern "C" int bb_ens_set_ll(bb_ens* e, int64_t p0, int64_t np, const double* host) {
  if (!e || !host || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  BB_CUDA(cudaMemcpyAsync(e->ll + p0, host, sizeof(double) * np, cudaMemcpyHostToDevice, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return BB_OK;
}
extern "C" int bb_ens_get_accepted(bb_ens* e, int64_t p0, int64_t np, uint8_t* host) {
  if (!e || !host || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  BB_CUDA(cudaMemcpyAsync(host, e->accepted + p0, (size_t)np, cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return BB_OK;
}
extern "C" int bb_ens_get_acc(bb_ens* e, int64_t* acc) {
  if (!e || !acc) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  unsigned long long v = 0;
  BB_CUDA(cudaMemcpyAsync(&v, e->acc, sizeof(v), cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  *acc = (int64_t)v;
  return BB_OK;
}
extern "C" int bb_ens_reset_acc(bb_ens* e) {
  if (!e) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  BB_CUDA(cudaMemsetAsync(e->acc, 0, sizeof(unsigned long long), e->ctx->stream));
  return BB_OK;
}
extern "C" void* bb_ens_acc_device_ptr(bb_ens* e) { return e ? (void*)e->acc : nullptr; }
extern "C" int64_t bb_ens_bytes(bb_ens* e) { return e ? e->bytes : -1; }

/* ------------------------------------------------------------------------------------------------ guiding tables */
extern "C" int bb_guide_destroy(bb_guide* g) {
  if (!g) return BB_ERR_ARG;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream); /* no launch of this context still reads the table */
  bb_pool_release(g->ctx, g->tab);
  g->ctx->live--;
  delete g;
  return BB_OK;
}

static int guide_create_impl(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                             const double* A, const double* b, const double* Mm, const double* v, const double* Bt,
                             const double* betat, int32_t aux_const, const double* Adiff, int32_t adiff_const,
                             bb_guide** out);
extern "C" int bb_guide_create(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                               const double* A, const double* b, const double* Mm, const double* v,
                               const double* Bt, const double* betat, int32_t aux_const, bb_guide** out) {
  return guide_create_impl(ctx, kind, N, d, m, tt, A, b, Mm, v, Bt, betat, aux_const, nullptr, 1, out);
}
extern "C" int bb_guide_create_ncd(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                                   const double* A, const double* b, const double* Mm, const double* v,
                                   const double* Bt, const double* betat, int32_t aux_const, const double* Adiff,
                                   int32_t adiff_const, bb_guide** out) {
  if (!Adiff) return BB_ERR_ARG;
  return guide_create_impl(ctx, kind, N, d, m, tt, A, b, Mm, v, Bt, betat, aux_const, Adiff, adiff_const, out);
}
static int guide_create_impl(bb_ctx* ctx, int32_t kind, int32_t N, int32_t d, int32_t m, const double* tt,
                             const double* A, const double* b, const double* Mm, const double* v, const double* Bt,
                             const double* betat, int32_t aux_const, const double* Adiff, int32_t adiff_const,
                             bb_guide** out) {
  if (!out) return BB_ERR_ARG;
  *out = nullptr;
  if (!ctx) return BB_ERR_NODEVICE;
  if (!tt || !A || !b || !Bt || !betat || N < 2 || d < 1 || d > BB_MAXD_WIDE) return BB_ERR_ARG;
  /* wide models (d > 3): PartialBridgeνH tables with a constant auxiliary drift only */
  if (d > 3 && (kind != BB_GUIDE_NUH || aux_const == 0 || Adiff)) return BB_ERR_UNSUPPORTED;
  if (kind != BB_GUIDE_NUH && kind != BB_GUIDE_HV && kind != BB_GUIDE_LMMU) return BB_ERR_ARG;
  if (kind == BB_GUIDE_LMMU) {
    if (!Mm || !v || m < 1 || m > d) return BB_ERR_ARG;
  } else {
    m = 0;
  }
  BB_CUDA(cudaSetDevice(ctx->device));
  const int auxm = Adiff ? 2 : (aux_const != 0 ? 1 : 0);
  const bool auxc = auxm == 1;
  const int rec = bb_rec_len(kind, d, m, auxm);
  const int NC = (N + BB_TC - 1) / BB_TC;
  const int nc = bb_rec_nc(kind, d, m), na1 = bb_rec_na1(kind, d, m), na2 = bb_rec_na2(kind, d, m);
  const int off_c = 2, off_a1 = off_c + nc, off_a2 = off_a1 + na1, off_bt = off_a2 + na2, off_be = off_bt + d * d,
            off_tr = off_be + d, off_ad = off_tr + 1;
  std::vector<double> tab;
  grid_rows(tt, N, NC, rec, tab);
  if (d > 3) { /* B~, beta~ follow the rows (they do not fit the per-segment kernel constants): bb_wide.cuh */
    tab.insert(tab.end(), Bt, Bt + (size_t)d * d);
    tab.insert(tab.end(), betat, betat + d);
  }
  for (int j = 1; j < N; j++) {
    const int i = j - 1;
    double* R = tab.data() + (size_t)j * rec;
    if (kind == BB_GUIDE_NUH) {
      memcpy(R + off_c, b + (size_t)i * d, sizeof(double) * d);
      memcpy(R + off_a2, A + (size_t)i * d * d, sizeof(double) * d * d);
    } else if (kind == BB_GUIDE_HV) {
      memcpy(R + off_c, b + (size_t)i * d, sizeof(double) * d);
      /* r = H♢[i] \ (V[i] - x)  (src/guip.jl:193) evaluated as inv(H♢[i]) (V[i] - x) */
      if (bb_h_inv(d, A + (size_t)i * d * d, R + off_a2)) return BB_ERR_SINGULAR;
    } else {
      const double* L = A + (size_t)i * m * d;
      const double* mu = b + (size_t)i * m;
      const double* Mi = Mm + (size_t)i * m * m;
      for (int k = 0; k < m; k++) R[off_c + k] = v[k] - mu[k];
      memcpy(R + off_a1, L, sizeof(double) * m * d);
      /* L[i]' M[i]  (d x m), products accumulated as the oracle's mat_mul */
      for (int r = 0; r < d; r++)
        for (int c2 = 0; c2 < m; c2++) {
          double s = L[0 * d + r] * Mi[0 * m + c2];
          for (int l = 1; l < m; l++) s = fma(L[l * d + r], Mi[l * m + c2], s);
          R[off_a2 + r * m + c2] = s;
        }
    }
    if (!auxc) { /* tabulated auxiliary drift (a constant one is replicated when the record carries it) */
      memcpy(R + off_bt, aux_const ? Bt : Bt + (size_t)i * d * d, sizeof(double) * d * d);
      memcpy(R + off_be, aux_const ? betat : betat + (size_t)i * d, sizeof(double) * d);
    }
    if (auxm == 2) {
      /* H((i,s),x,P°): H[i] (νH), inv(H♢[i]) (GuidedBridge), L[i]'M[i]L[i] (PartialBridge  src/partialbridge.jl:58) */
      const double* Ad = adiff_const ? Adiff : Adiff + (size_t)i * d * d;
      double Hm[BB_MAXD * BB_MAXD];
      if (kind == BB_GUIDE_LMMU) {
        const double* L = A + (size_t)i * m * d;
        for (int r = 0; r < d; r++)
          for (int c2 = 0; c2 < d; c2++) {
            double s2 = R[off_a2 + r * m] * L[c2];
            for (int l = 1; l < m; l++) s2 = fma(R[off_a2 + r * m + l], L[l * d + c2], s2);
            Hm[r * d + c2] = s2;
          }
      } else {
        memcpy(Hm, R + off_a2, sizeof(double) * d * d);
      }
      /* tr(A H), products accumulated as the oracle's mat_trace_prod */
      double tr = 0.0;
      for (int r = 0; r < d; r++) {
        double s2 = Ad[r * d] * Hm[r];
        for (int l = 1; l < d; l++) s2 = fma(Ad[r * d + l], Hm[l * d + r], s2);
        tr = (r == 0) ? s2 : tr + s2;
      }
      R[off_tr] = tr;
      memcpy(R + off_ad, Ad, sizeof(double) * d * d);
    }
  }
  double segc[BB_SEGC];
  memset(segc, 0, sizeof(segc));
  if (auxc && d <= 3) {
    memcpy(segc, Bt, sizeof(double) * d * d);
    memcpy(segc + d * d, betat, sizeof(double) * d);
  }
  if (kind == BB_GUIDE_HV) {
    /* endpoint(y, P) = norm(P.H♢[end], 1) < eps() ? P.V[end] : y   src/euler.jl:241-242 */
    const double* K = A + (size_t)(N - 1) * d * d;
    double n1 = 0;
    for (int k = 0; k < d * d; k++) n1 += fabs(K[k]);
    if (n1 < 2.220446049250313e-16) {
      segc[d * d + d] = 1.0;
      memcpy(segc + d * d + d + 1, b + (size_t)(N - 1) * d, sizeof(double) * d);
    }
  }
  bb_guide* g = new (std::nothrow) bb_guide();
  if (!g) return BB_ERR_NOMEM;
  g->ctx = ctx; g->kind = kind; g->N = N; g->d = d; g->m = m; g->auxc = auxm; g->NC = NC; g->rec = rec;
  ctx->live++;
  g->tt.assign(tt, tt + N);
  memcpy(g->segc, segc, sizeof(segc));
  cudaError_t e1 = bb_pool_alloc(ctx, tab.size() * sizeof(double), (void**)&g->tab);
  if (e1 != cudaSuccess) {
    bb_set_cuda_error(e1, "cudaMalloc(guide)");
    bb_guide_destroy(g);
    return BB_ERR_NOMEM;
  }
  cudaError_t e2 = cudaMemcpyAsync(g->tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
  if (e2 != cudaSuccess) {
    bb_set_cuda_error(e2, "guide table upload");
    bb_guide_destroy(g);
    return BB_ERR_CUDA;
  }
  *out = g;
  return BB_OK;
}

/* ------------------------------------------------------------------------------------------------ launches */
struct run_spec {
  int rng;            /* 0 read W, 1 pCN, 2 sample */
  bool store_x, do_ll, write_end;
  int skip;
  double rho;
  uint64_t seed;
  uint32_t stream;
  bool only_stale = false;
  int64_t p_begin = 0, p_end = -1; /* chain sub-range (default: all) */
  const double* tab1 = nullptr;    /* Mdb: per-step noise factors (device), passed as a.tab[1] (S = 1) */
};

/* rs.rng: 0/1/2 = path kernel modes; 10 = llikelihood on stored X; 11 = innovations!; 12 = StochasticHeun; 13 = Mdb */
static int run_chain(bb_ens* e, const bb_model* model, bb_guide* const* guides, const run_spec& rs) {
  if (!e || !model) return BB_ERR_ARG;
  bb_ctx* c = e->ctx;
  if (!model_shape_ok(model)) return BB_ERR_MODEL;
  if (model->d != e->d || model->dprime != e->dp) return BB_ERR_MODEL;
  BB_CUDA(cudaSetDevice(c->device));
  bb_chain_args a;
  memset(&a, 0, sizeof(a));
  int gk = 0, gm = 0, auxc = 1; /* auxc carries the guide's auxiliary mode (1 const, 0 tabulated, 2 non-constdiff) */
  for (int s = 0; s < e->S; s++) {
    if (guides) {
      const bb_guide* g = guides[s];
      if (!g || g->ctx->device != c->device) return BB_ERR_ARG; /* tables live on the guide's device */
      if (g->N != e->N) return BB_ERR_LENGTH; /* "Y and W differ in length." src/euler.jl:251 */
      if (g->d != e->d) return BB_ERR_MODEL;
      if (s == 0) { gk = g->kind; gm = g->m; auxc = g->auxc; }
      else if (g->kind != gk || g->m != gm || g->auxc != auxc) return BB_ERR_UNSUPPORTED;
      a.tab[s] = g->tab;
      memcpy(a.segc[s], g->segc, sizeof(a.segc[s]));
    } else {
      if (!e->gridtab[s]) return BB_ERR_ARG; /* bb_ens_set_grid first */
      a.tab[s] = e->gridtab[s];
    }
  }
  const int krng = (rs.rng == 1 && !rs.store_x) ? 3 : rs.rng; /* pCN without X° is its own instantiation */
  bb_user_model* um = model->id == BB_MODEL_USER ? bb_user_lookup(model->reserved) : nullptr;
  if (model->id == BB_MODEL_USER && (rs.rng == 11 || (rs.rng != 12 && rs.rng >= 10 && !guides))) return BB_ERR_UNSUPPORTED;
  bb_chain_launch_fn fn = nullptr;
  if (krng == 1 && guides && model->dprime == 1 && model->d <= 3) {
    /* a SMALL ensemble (the strong-scaling share of a GPU: fewer than two 128-chain CTAs per SM) runs the pCN
     * iteration as the warp-specialised kernel (two threads per chain, bb_chain_ws.cuh): 5 % faster there, equal at
     * full size.  Same results bit for bit.  bb_ctx_set_pcn_kernel (or BB_PCN_WS=0 / 1) forces one or the other. */
    static const int env = []() {
      const char* v = getenv("BB_PCN_WS");
      return v ? (v[0] == '0' ? BB_PCN_ONE_THREAD : (v[0] == '2' ? BB_PCN_WARP_SPECIALISED_2 : BB_PCN_WARP_SPECIALISED))
               : BB_PCN_AUTO;
    }();
    const int mode = c->pcn_kernel != BB_PCN_AUTO ? c->pcn_kernel : env;
    const long long nch = (rs.p_end < 0 ? e->P : rs.p_end) - rs.p_begin;
    const bool small = (nch + 127) / 128 < 2ll * c->sm_count;
    if (mode == BB_PCN_WARP_SPECIALISED_2 || (mode == BB_PCN_AUTO && small && BB_AUTO_WS2))
      fn = lookup_kernel(model, gk, gm, auxc, 6); /* two chains per dynamics thread (d <= 2) */
    if (!fn && (mode == BB_PCN_WARP_SPECIALISED || mode == BB_PCN_WARP_SPECIALISED_2 || (mode == BB_PCN_AUTO && small)))
      fn = lookup_kernel(model, gk, gm, auxc, 5);
  }
  if (!fn && !um) fn = rs.rng >= 10 ? lookup_second(model, gk, gm, auxc, rs.rng - 10) : lookup_kernel(model, gk, gm, auxc, krng);
  if (!fn && !um) return BB_ERR_UNSUPPORTED;
  if (rs.rng >= 10 && !e->X) return BB_ERR_ARG;
  if (rs.rng == 11 && !model_sigma_invertible(model)) return BB_ERR_SINGULAR;
  if (rs.rng == 1 && !(e->flags & BB_ENS_DOUBLE_BUFFER)) return BB_ERR_ARG;
  if (rs.store_x && !e->X) return BB_ERR_ARG;
  if (rs.skip < 0) return BB_ERR_ARG;
  if (rs.tab1) a.tab[1] = rs.tab1;
  a.W[0] = e->W[0]; a.W[1] = e->W[1]; a.X = e->X;
  a.xstale = e->xstale; a.only = rs.only_stale ? e->xstale : nullptr;
  a.nbuf = e->nbuf;
  a.par = e->par; a.start = e->start; a.start_bcast = e->start_bcast;
  a.ll = e->ll; a.llprop = e->llprop; a.logu = e->logu; a.xend = e->xend; a.xendprop = e->xendprop;
  a.accepted = e->accepted; a.acc = e->acc;
  a.P = e->P; a.p_begin = rs.p_begin; a.p_end = rs.p_end < 0 ? e->P : rs.p_end;
  if (a.p_begin < 0 || a.p_end > e->P || a.p_begin >= a.p_end) return BB_ERR_ARG;
  a.chain_offset = e->chain_offset; a.S = e->S; a.N = e->N; a.NC = e->NC;
  a.jll = e->N - 1 - rs.skip;
  a.store_x = rs.store_x ? 1 : 0; a.do_ll = rs.do_ll ? 1 : 0; a.write_end = rs.write_end ? 1 : 0;
  bb_philox_key_schedule(rs.seed, a.keys);
  a.stream = rs.stream;
  a.rho = rs.rho;
  a.rho2 = sqrt(1 - rs.rho * rs.rho); /* sqrt(1-ρ^2)  test/partialbridgenuH.jl:178 */
  bb_prepare_model(model, &a.model);
  bb_time_begin(c);
  if (um) { /* run-time compiled model: the same kernels, instantiated for the user's b and sigma (bb_user.cu) */
    const int rcu = rs.rng == 13 ? BB_ERR_UNSUPPORTED
                  : rs.rng == 12 ? bb_user_launch(um, 2, 0, 0, 1, 0, a, c->stream)
                  : rs.rng >= 10 ? bb_user_launch(um, 1, gk, gm, auxc, rs.rng - 10, a, c->stream)
                                 : bb_user_launch(um, 0, gk, gm, auxc, krng, a, c->stream);
    if (rcu != BB_OK) return rcu;
  } else {
    cudaError_t err = fn(a, c->stream);
    if (err == cudaErrorInvalidConfiguration) return BB_ERR_UNSUPPORTED; /* the staging does not fit shared memory */
    if (err != cudaSuccess) {
      bb_set_cuda_error(err, "bb_chain_kernel launch");
      return BB_ERR_CUDA;
    }
  }
  bb_time_end(c);
  c->launches++;
  if (rs.rng == 1) e->x_maybe_stale = true;
  else if ((rs.rng < 10 || rs.rng == 12 || rs.rng == 13) && rs.store_x && !rs.only_stale) e->x_maybe_stale = false;
  else if (rs.rng < 10 && !rs.store_x && e->X) e->x_maybe_stale = true; /* X no longer belongs to the chains' state */
  if (rs.only_stale) e->x_maybe_stale = false;
  return BB_OK;
}

extern "C" int bb_wiener_sample(bb_ens* e, uint64_t seed, uint32_t stream) {
  if (!e) return BB_ERR_ARG;
  /* the Wiener process of dimension d' as its own target: X is not touched */
  bb_model w;
  memset(&w, 0, sizeof(w));
  w.id = BB_MODEL_WIENER; w.d = e->dp; w.dprime = e->dp;
  bb_ens view = *e; /* same device buffers, state dimension d'; X is neither stored nor returned */
  view.d = e->dp;
  run_spec rs{2, false, false, false, 0, 0.0, seed, stream};
  const int rc = run_chain(&view, &w, nullptr, rs);
  if (rc == BB_OK && e->X) e->x_maybe_stale = true; /* X no longer belongs to the new W */
  return rc;
}
extern "C" int bb_euler(bb_ens* e, const bb_model* model) {
  run_spec rs{0, true, false, true, 0, 0.0, 0, 0};
  return run_chain(e, model, nullptr, rs);
}
extern "C" int bb_solve_scheme(bb_ens* e, const bb_model* model, int32_t scheme) {
  if (!e || !model) return BB_ERR_ARG;
  switch (scheme) {
    case BB_SCHEME_EULER:
    case BB_SCHEME_STRATONOVICH: /* ½(σ + σ) = σ exactly for the registry's constant σ  (src/euler.jl:82-83) */
      return bb_euler(e, model);
    case BB_SCHEME_SRK: /* T <: Number only (src/euler.jl:330); its correction term is exactly 0 for constant σ */
      if (model->d != 1) return BB_ERR_UNSUPPORTED;
      return bb_euler(e, model);
    case BB_SCHEME_HEUN: {
      if (e->S != 1) return BB_ERR_UNSUPPORTED;
      run_spec rs{12, true, false, true, 0, 0.0, 0, 0};
      return run_chain(e, model, nullptr, rs);
    }
    case BB_SCHEME_MDB: return BB_ERR_UNSUPPORTED;
    default: return BB_ERR_ARG;
  }
}
/* solve!(Mdb(), Y, u, W, P°)  src/euler.jl:308-327 on a guided proposal (one segment): X_cur <- the path, xend <- yy[N] */
extern "C" int bb_guided_mdb(bb_ens* e, const bb_model* model, bb_guide* const* guides) {
  if (!e || !model || !guides || !guides[0]) return BB_ERR_ARG;
  if (e->S != 1) return BB_ERR_UNSUPPORTED;
  const bb_guide* g = guides[0];
  if (g->N != e->N) return BB_ERR_LENGTH;
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  const int N = e->N;
  std::vector<double> sc((size_t)e->NC * BB_TC, 0.0);
  const double T = g->tt[N - 1];
  for (int j = 1; j < N; j++) sc[j] = sqrt((T - g->tt[j]) / (T - g->tt[j - 1])); /* step j-1 -> j (i = j-1 in the reference) */
  double* dsc = nullptr;
  BB_CUDA(bb_pool_alloc(c, sc.size() * sizeof(double), (void**)&dsc));
  cudaError_t ce = cudaMemcpyAsync(dsc, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
  int rc = BB_ERR_CUDA;
  if (ce == cudaSuccess) {
    run_spec rs{13, true, false, true, 0, 0.0, 0, 0};
    rs.tab1 = dsc;
    rc = run_chain(e, model, guides, rs);
  } else {
    bb_set_cuda_error(ce, "bb_guided_mdb");
  }
  cudaStreamSynchronize(c->stream); /* the pageable copy above and the launch have consumed sc / dsc */
  bb_pool_release(c, dsc);
  return rc;
}
extern "C" int bb_sample_euler(bb_ens* e, const bb_model* model, uint64_t seed, uint32_t stream) {
  run_spec rs{2, true, false, true, 0, 0.0, seed, stream};
  return run_chain(e, model, nullptr, rs);
}
extern "C" int bb_guided_euler_ll(bb_ens* e, const bb_model* model, bb_guide* const* guides, int32_t skip,
                                  uint32_t flags) {
  if (!guides) return BB_ERR_ARG;
  run_spec rs{0, (flags & BB_RUN_STORE_X) != 0, (flags & BB_RUN_NO_LL) == 0, true, skip, 0.0, 0, 0};
  return run_chain(e, model, guides, rs);
}
extern "C" int bb_pcn_step(bb_ens* e, const bb_model* model, bb_guide* const* guides, double rho, uint64_t seed,
                           uint32_t iter, int32_t skip, uint32_t flags) {
  if (!guides) return BB_ERR_ARG;
  if (!(rho >= -1.0 && rho <= 1.0)) return BB_ERR_ARG;
  run_spec rs{1, (flags & BB_RUN_STORE_X) != 0, true, true, skip, rho, seed, iter};
  return run_chain(e, model, guides, rs);
}
extern "C" int bb_llikelihood(bb_ens* e, const bb_model* model, bb_guide* const* guides, int32_t skip) {
  if (!guides) return BB_ERR_ARG;
  run_spec rs{10, false, true, false, skip, 0.0, 0, 0};
  return run_chain(e, model, guides, rs);
}
extern "C" int bb_innovations(bb_ens* e, const bb_model* model, bb_guide* const* guides) {
  run_spec rs{11, false, false, false, 0, 0.0, 0, 0};
  return run_chain(e, model, guides, rs);
}
extern "C" int bb_ens_refresh_x(bb_ens* e, const bb_model* model, bb_guide* const* guides) {
  if (!e) return BB_ERR_ARG;
  if (!e->x_maybe_stale) return BB_OK;
  run_spec rs{0, true, false, false, 0, 0.0, 0, 0};
  rs.only_stale = true;
  return run_chain(e, model, guides, rs);
}

/* ------------------------------------------------------------------------------------------------ host-buffer pCN
 * The reference loop keeps W, X in host memory.  This entry point runs one pCN iteration on HOST buffers:
 * W (current) goes up, W°, X°, ll°, accept flags come back, in slabs of chains that are pipelined over three
 * streams (H2D of slab k+1 | transpose + path kernel of slab k | D2H of slab k-1), so the PCIe link is busy in
 * both directions while the GPU computes.  The chains' device state is updated as by bb_pcn_step. */
__global__ void __launch_bounds__(256) bb_slab_in_kernel(double* __restrict__ W0, const double* __restrict__ stage,
                                                         const uint8_t* __restrict__ par, long long P, long long p0,
                                                         long long np, int S, int N, int NC, int K, int nbuf) {
  const long long rowlen = (long long)BB_TC * K, total = (long long)S * NC * np * rowlen;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(t % rowlen);
    long long q = t / rowlen;
    const long long pl = q % np;
    q /= np;
    const int c = (int)(q % NC), s = (int)(q / NC);
    const int slot = e / K, k = e - slot * K, j = c * BB_TC + slot;
    if (j >= N) continue;
    const long long p = p0 + pl;
    W0[((((long long)s * NC + c) * P + p) * nbuf + par[p]) * rowlen + e] = stage[((pl * S + s) * (long long)N + j) * K + k];
  }
}
/* proposal W° (buffer the chain did NOT read from in this step = par_before ^ 1; par may have flipped on accept) and X° */
__global__ void __launch_bounds__(256) bb_slab_out_kernel(const double* __restrict__ W0, const double* __restrict__ X,
                                                          double* __restrict__ stageW, double* __restrict__ stageX,
                                                          const uint8_t* __restrict__ par,
                                                          const uint8_t* __restrict__ accepted, long long P,
                                                          long long p0, long long np, int S, int N, int NC, int dp,
                                                          int d, int nbuf) {
  const long long rw = (long long)BB_TC * dp, rx = (long long)BB_TC * d;
  const long long totw = (long long)S * NC * np * rw, totx = stageX ? (long long)S * NC * np * rx : 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < totw + totx;
       t += (long long)gridDim.x * blockDim.x) {
    const bool isx = t >= totw;
    const long long u = isx ? t - totw : t, rowlen = isx ? rx : rw;
    const int K = isx ? d : dp;
    const int e = (int)(u % rowlen);
    long long q = u / rowlen;
    const long long pl = q % np;
    q /= np;
    const int c = (int)(q % NC), s = (int)(q / NC);
    const int slot = e / K, k = e - slot * K, j = c * BB_TC + slot;
    if (j >= N) continue;
    const long long p = p0 + pl;
    const long long h = ((pl * S + s) * (long long)N + j) * K + k;
    if (isx) {
      stageX[h] = X[(((long long)s * NC + c) * P + p) * rowlen + e];
    } else {
      const int b = accepted[p] ? par[p] : 1 - par[p];
      stageW[h] = W0[((((long long)s * NC + c) * P + p) * nbuf + b) * rowlen + e];
    }
  }
}

/* BB_RUN_SKIP_REJECTED with mapped host buffers: the rows of the chains that accepted go from the slab's staging
 * buffer (already in the host layout [chain][segment][N][k]: bb_slab_out_kernel) straight into the caller's host arrays,
 * one CTA per chain, front to back; chains that rejected are skipped.  Both sides stream: gathering a chain's rows from
 * the chunked device layout inside this kernel (64 MB between consecutive chunk rows of a chain) ran at 29-40 GB/s over
 * the link, a streaming kernel reaches the copy engine's 48-52 GB/s (tools/pciebench.cu, profiles/r02_pciebench.txt). */
/* list of the chains of a slab that accepted (ascending), count in list[np] */
__global__ void __launch_bounds__(1024) bb_accepted_list_kernel(const uint8_t* __restrict__ accepted, long long np,
                                                                int* __restrict__ list) {
  __shared__ int wsum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (long long t0 = 0; t0 < np; t0 += 1024) {
    const long long t = t0 + threadIdx.x;
    const bool f = t < np && accepted[t] != 0;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; i++) off += wsum[i];
    if (f) list[off + __popc(m & ((1u << lane) - 1))] = (int)t;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < 32; i++) tot += wsum[i];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) list[np] = base;
}
__global__ void __launch_bounds__(256) bb_rows_to_host_kernel(const double* __restrict__ stageW, double* __restrict__ hostW,
                                                              long long rowW, const double* __restrict__ stageX,
                                                              double* __restrict__ hostX, long long rowX,
                                                              const int* __restrict__ list, long long np) {
  /* the CTAs walk the accepted rows TOGETHER, 2 KB per CTA and step, so that the stores over the link form one dense
   * moving window (one private row per CTA = ~600 interleaved streams reached 40 GB/s, this order 44-45); W rows first,
   * then X rows (hostX may be NULL), in one launch */
  const long long pw = (rowW + 255) / 256, px = hostX ? (rowX + 255) / 256 : 0; /* pieces per row */
  const long long cnt = list[np], totw = cnt * pw, total = totw + cnt * px;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const bool isx = t >= totw;
    const long long u = isx ? t - totw : t, ppr = isx ? px : pw, rowlen = isx ? rowX : rowW;
    const long long r = u / ppr, i = (u - r * ppr) * 256 + threadIdx.x;
    if (i < rowlen) {
      const long long pl = list[r];
      if (isx) hostX[pl * rowlen + i] = stageX[pl * rowlen + i];
      else hostW[pl * rowlen + i] = stageW[pl * rowlen + i];
    }
  }
}

static double* mapped_alias(const void* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return at.type == cudaMemoryTypeHost ? (double*)at.devicePointer : nullptr;
}

extern "C" int bb_pcn_step_host(bb_ens* e, const bb_model* model, bb_guide* const* guides, double rho, uint64_t seed,
                                uint32_t iter, int32_t skip, uint32_t flags, const double* W_host, double* Wo_host,
                                double* Xo_host, double* llo_host, uint8_t* accepted_host) {
  if (!e || !guides || !W_host || !Wo_host) return BB_ERR_ARG;
  if (!(rho >= -1.0 && rho <= 1.0)) return BB_ERR_ARG;
  const bool want_x = Xo_host != nullptr;
  if (want_x && !e->X) return BB_ERR_ARG;
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  if (!c->s_h2d) {
    BB_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    BB_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      BB_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
      BB_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
      BB_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
  }
  const size_t wpc = (size_t)e->S * e->N * e->dp, xpc = (size_t)e->S * e->N * e->d; /* doubles per chain */
  /* accepted rows straight into mapped host memory (see the header) */
  double* Wo_map = (flags & BB_RUN_SKIP_REJECTED) ? mapped_alias(Wo_host) : nullptr;
  double* Xo_map = (flags & BB_RUN_SKIP_REJECTED) && want_x ? mapped_alias(Xo_host) : nullptr;
  const bool direct = Wo_map && (!want_x || Xo_map);
  /* slab of chains per pipeline stage: 192 MB of W for the copy engine, 384 MB for the kernel write-back (measured:
   * 2.60 / 2.65 / 2.73 / 2.55e9 path-steps/s at 96 / 192 / 384 / 768 MB; the copy-engine path prefers 96-192) */
  int64_t slab = (int64_t)((size_t)((direct ? 384u : 192u) << 20) / (wpc * sizeof(double)));
  slab = slab < 256 ? 256 : (slab / 256) * 256; /* whole CTAs */
  if (slab > e->P) slab = e->P;
  /* Wo_host == W_host: the host array is updated in place for the chains that accept (the loop's swap of W and Wo);
   * only the direct path leaves the rejecting chains' rows alone */
  if (Wo_host == W_host && !direct) return BB_ERR_ARG;
  const size_t per_slab = (size_t)slab * (2 * wpc + (want_x ? xpc : 0));
  const size_t list_doubles = ((size_t)slab + 1 + 1) / 2 + 16; /* an int list of accepted chains per staging buffer */
  int rc = ctx_stage(c, (2 * per_slab + 2 * list_doubles) * sizeof(double));
  if (rc != BB_OK) return rc;
  double* st[2] = {c->stage, c->stage + per_slab};
  int* acc_list[2] = {reinterpret_cast<int*>(c->stage + 2 * per_slab),
                      reinterpret_cast<int*>(c->stage + 2 * per_slab + list_doubles)};
  BB_CUDA(cudaStreamSynchronize(c->stream)); /* earlier work on the context's stream is complete */
  run_spec rs{1, want_x && (flags & BB_RUN_STORE_X) != 0, true, true, skip, rho, seed, iter};
  if (want_x) rs.store_x = true;
  int k = 0;
  for (int64_t p0 = 0; p0 < e->P; p0 += slab, k++) {
    const int64_t n = (e->P - p0 < slab) ? e->P - p0 : slab;
    const int b = k & 1;
    double *sWin = st[b], *sWo = st[b] + (size_t)slab * wpc, *sXo = want_x ? st[b] + 2 * (size_t)slab * wpc : nullptr;
    /* staging buffers of parity b are free once the D2H of slab k-2 has finished */
    if (k >= 2) BB_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_out[b], 0));
    BB_CUDA(cudaMemcpyAsync(sWin, W_host + (size_t)p0 * wpc, (size_t)n * wpc * sizeof(double), cudaMemcpyHostToDevice,
                            c->s_h2d));
    BB_CUDA(cudaEventRecord(c->ev_in[b], c->s_h2d));
    BB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[b], 0));
    const long long tin = (long long)e->S * e->NC * n * BB_TC * e->dp;
    bb_slab_in_kernel<<<(unsigned)((tin + 255) / 256 > 148 * 16 ? 148 * 16 : (tin + 255) / 256), 256, 0, c->stream>>>(
        e->W[0], sWin, e->par, e->P, p0, n, e->S, e->N, e->NC, e->dp, e->nbuf);
    rs.p_begin = p0; rs.p_end = p0 + n;
    rc = run_chain(e, model, guides, rs);
    if (rc != BB_OK) return rc;
    const long long tout = (long long)e->S * e->NC * n * BB_TC * (e->dp + (want_x ? e->d : 0));
    bb_slab_out_kernel<<<(unsigned)((tout + 255) / 256 > 148 * 16 ? 148 * 16 : (tout + 255) / 256), 256, 0, c->stream>>>(
        e->W[0], e->X, sWo, sXo, e->par, e->accepted, e->P, p0, n, e->S, e->N, e->NC, e->dp, e->d, e->nbuf);
    BB_CUDA(cudaGetLastError());
    c->launches += 2;
    BB_CUDA(cudaEventRecord(c->ev_comp[b], c->stream));
    BB_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[b], 0));
    if (direct) {
      /* on the copy stream (the next slab's kernels do not wait for the link); a small grid keeps the link busy */
      bb_accepted_list_kernel<<<1, 1024, 0, c->s_d2h>>>(e->accepted + p0, n, acc_list[b]);
      bb_rows_to_host_kernel<<<148 * 4, 256, 0, c->s_d2h>>>(sWo, Wo_map + (size_t)p0 * wpc, (long long)wpc, sXo,
                                                             want_x ? Xo_map + (size_t)p0 * xpc : nullptr, (long long)xpc,
                                                             acc_list[b], n);
      BB_CUDA(cudaGetLastError());
      c->launches += 2;
    } else {
    BB_CUDA(cudaMemcpyAsync(Wo_host + (size_t)p0 * wpc, sWo, (size_t)n * wpc * sizeof(double), cudaMemcpyDeviceToHost,
                            c->s_d2h));
    if (want_x)
      BB_CUDA(cudaMemcpyAsync(Xo_host + (size_t)p0 * xpc, sXo, (size_t)n * xpc * sizeof(double), cudaMemcpyDeviceToHost,
                              c->s_d2h));
    }
    if (llo_host)
      BB_CUDA(cudaMemcpyAsync(llo_host + p0, e->llprop + p0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->s_d2h));
    if (accepted_host)
      BB_CUDA(cudaMemcpyAsync(accepted_host + p0, e->accepted + p0, (size_t)n, cudaMemcpyDeviceToHost, c->s_d2h));
    BB_CUDA(cudaEventRecord(c->ev_out[b], c->s_d2h));
    /* the next slab's input transpose must not overwrite sWin of parity b^1 ... it uses the other buffer; the
     * compute stream itself orders slab k+1's kernels after slab k's */
  }
  BB_CUDA(cudaStreamSynchronize(c->s_d2h));
  BB_CUDA(cudaStreamSynchronize(c->stream));
  return BB_OK;
}
