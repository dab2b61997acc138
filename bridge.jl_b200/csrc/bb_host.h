/*
 * bb_host.h -- host-side structures of libbridge_b200.so (not part of the public ABI).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bridge_b200.h"
#include "bb_chain.cuh"

struct bb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int64_t launches = 0;
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  int sm_count = 0;
  int live = 0;  /* ensembles + guides created on this context and not yet destroyed (bb_ctx_destroy refuses while > 0) */
  int pcn_kernel = 0; /* BB_PCN_AUTO / BB_PCN_ONE_THREAD / BB_PCN_WARP_SPECIALISED */
  int arith = 0; /* BB_ARITH_REFERENCE / BB_ARITH_FUSED: rounding order of the shared-table constructors */
  /* staging for host <-> device transposes */
  double* stage = nullptr;
  size_t stage_bytes = 0;
  /* small device buffers (constructor work space, guiding tables) are recycled instead of cudaMalloc / cudaFree per
   * call: cudaFree synchronises the whole device, which serialised table rebuilds with the running path kernel */
  std::vector<std::pair<size_t, void*>> pool_free;
  std::unordered_map<void*, size_t> pool_size;
  /* host-buffer pipeline (bb_pcn_step_host) */
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

struct bb_guide {
  bb_ctx* ctx = nullptr;
  int kind = 0, N = 0, d = 0, m = 0, auxc = 1, NC = 0, rec = 0;
  double* tab = nullptr;  /* device [NC*8][rec] */
  double segc[BB_SEGC] = {0}; /* per-segment constants handed to the kernels by value */
  std::vector<double> tt;
};

struct bb_ens {
  bb_ctx* ctx = nullptr;
  int64_t P = 0;
  int S = 0, N = 0, d = 0, dp = 0, NC = 0, nbuf = 1;
  uint32_t flags = 0;
  int64_t chain_offset = 0;
  double* W[2] = {nullptr, nullptr};
  double* X = nullptr;
  uint8_t* xstale = nullptr;
  bool x_maybe_stale = false;
  uint8_t* par = nullptr;
  uint8_t* accepted = nullptr;
  double *ll = nullptr, *llprop = nullptr, *logu = nullptr, *xend = nullptr, *xendprop = nullptr;
  unsigned long long* acc = nullptr;
  double* start = nullptr; /* [d] or [d][P] */
  int start_bcast = 1;
  std::vector<double*> gridtab; /* per segment: device [NC*8][2] (dt, sqrt dt) */
  std::vector<std::vector<double>> tt;
  int64_t bytes = 0;
  /* per-chain parameters and guiding tables (bb_theta.cu) */
  struct bb_theta* th = nullptr;
  /* pooled online statistics (bb_stats.cu) */
  double *mc_sum = nullptr, *mc_sq = nullptr, *mc_pivot = nullptr;
  int64_t mc_n = 0;
  /* per-chain online statistics (bb_stats.cu): m in the layout of X, m2 with d*d entries per grid point */
  double *cmc_m = nullptr, *cmc_m2 = nullptr;
  int64_t cmc_k = 0;
};

void bb_theta_free(bb_ens* e); /* bb_theta.cu */
void bb_theta_invalidate(bb_ens* e); /* grids or starting points changed: tables / left-end values must be rebuilt */
/* recycled small device buffers of a context (bb_api.cu) */
cudaError_t bb_pool_alloc(bb_ctx* c, size_t bytes, void** out);
void bb_pool_release(bb_ctx* c, void* p);

/* run-time compiled user models (bb_user.cu) */
struct bb_user_model;
bb_user_model* bb_user_lookup(int handle);
int bb_user_launch(bb_user_model* um, int kind, int gk, int gm, int auxm, int rng, const bb_chain_args& a, cudaStream_t st);

/* thread-local error text for bb_last_cuda_error */
void bb_set_cuda_error(cudaError_t e, const char* where);

#define BB_CUDA(call)                          \
  do {                                         \
    cudaError_t e__ = (call);                  \
    if (e__ != cudaSuccess) {                  \
      bb_set_cuda_error(e__, #call);           \
      return e__ == cudaErrorMemoryAllocation ? BB_ERR_NOMEM : BB_ERR_CUDA; \
    }                                          \
  } while (0)

/* per-model kernel lookup (bb_inst_*.cu) */
bb_chain_launch_fn bb_lookup_wiener(int d, int rng);
bb_chain_launch_fn bb_lookup_ou(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_linpro1(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_linpro2(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_linpro3(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_fhn_diag(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_fhn_hypo(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_intdiff(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_nclar3(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_lorenz(int gk, int gm, int auxc, int rng);
bb_chain_launch_fn bb_lookup_landmarks(int gk, int gm, int auxc, int rng); /* bb_wide.cuh */
bb_chain_launch_fn bb_lookup_wiener_wide(int d, int rng);

bb_chain_launch_fn bb_lookup2_wiener(int d, int mode);
bb_chain_launch_fn bb_lookup2_ou(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_linpro1(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_linpro2(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_linpro3(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_fhn_diag(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_fhn_hypo(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_intdiff(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_nclar3(int gk, int gm, int auxc, int mode);
bb_chain_launch_fn bb_lookup2_lorenz(int gk, int gm, int auxc, int mode);

/* timing brackets around compute calls */
void bb_time_begin(bb_ctx* ctx);
void bb_time_end(bb_ctx* ctx);

/* small host algebra in the oracle's operation order (bb_linalg.cpp-style helpers in bb_api.cu) */
int bb_h_inv(int d, const double* A, double* Ai);
