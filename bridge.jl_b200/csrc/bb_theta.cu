/*
 * bb_theta.cu -- per-chain parameters: parameter-update Metropolis-Hastings with the guiding tables of every
 * chain built on the device.
 *
 * Follows the `updateparams` branch of the script loop project_partialbridge/partialbridge_bolus3.jl:248-365 for
 * P independent chains, each with its own θ (see include/bridge_b200.h for the step-by-step correspondence):
 *   bb_theta_backward_kernel   propose(σ, P) (:239-242), then per chain: ν = 0, H⁺ = I/ϵ, gpupdate with the last
 *                              observation (:162-165), and for s = S-1 .. 0 the Lyapunov backward step of
 *                              partialbridgeνH (src/partialbridgenuH.jl:86-103,148-155, src/lyap.jl:2-6) followed by
 *                              gpupdate with v[s-1] (:284-291, :128-137); logpdfnormal(x0 - ν(0), H⁺(0))
 *                              (src/gaussian.jl:66-75), the trace term and logπ of :319,:336.
 *   bb_theta_forward_kernel    solve!(Euler(), X°, x0, W, Q°) + llikelihood (:324-333) with the chain's own θ and
 *                              tables, and the accept test (:340); the same kernel runs the pCN iteration
 *                              (test/partialbridgenuH.jl:176-191) for chains with their own tables.
 *
 * Layout: the tables are T [S][N][K][P (padded to 16)] doubles, K = d + d*d (ν[i], H[i]), chain-minor, so that the 32 chains of a
 * warp read/write 256 contiguous bytes per value (no staging needed); they are written once by the backward
 * kernel and read once by the forward kernel: 8K bytes per path-step each way.  The forward kernel moves a chain's
 * rows through a shared-memory ring with 8-byte cp.async, BB_TDEPTH-1 steps ahead of their use; X° and W° leave as
 * whole 128-byte lines.  W / X keep the chunked layout of bb_chain.cuh (shared with the other kernels).
 * The per-step arithmetic is bb_chain<...>::drift / bb_em_update / nuH_step, i.e. the same instruction sequence
 * as the shared-table kernels and the one-system constructors.
 */
#include <math.h>
#include <string.h>

#include <new>
#include <vector>

#include "bb_backward.cuh"
#include "bb_host.h"

#ifndef BB_TWPF
#define BB_TWPF 1 /* W pieces are loaded into registers one group of 4 steps ahead */
#endif
#ifndef BB_TPAIR
#define BB_TPAIR 1 /* table rows are copied 16 bytes (two chains) at a time, see t_issue */
#endif
#ifndef BB_TDEPTH
#define BB_TDEPTH 4 /* table rows a chain keeps in flight / in shared memory */
#endif

struct bb_theta {
  bb_model model;
  bb_theta_spec spec;
  int K = 0, NL = 0;
  long long PT = 0; /* chains per table row, padded to a multiple of 16 (whole 128-byte lines) */
  double* theta[2] = {nullptr, nullptr}; /* [BB_NTHETA][P]: current, last proposal */
  double* T = nullptr;                   /* [S][N][K][P] */
  double* left[2] = {nullptr, nullptr};  /* [NL][P]: ν(0), H⁺(0), C, logpdfnormal, trace term, log prior */
  double* tt = nullptr;                  /* [S][N] */
  double* startprop = nullptr;           /* [d][P]: x0° of the last parameter proposal (start moves only) */
  unsigned long long* acc = nullptr;
  double* blkv = nullptr;  /* [BB_NBLK][P]: the numbers of the last blocked update (bb_theta_get_block) */
  double* Xprop = nullptr; /* X° of a blocked update (same layout as X); committed to X for the chains that accept */
  bool ll_stale = false; /* block updates moved W without maintaining the running ll of the whole-path steps */
  int tstate = 0; /* 0: T invalid; 1: T belongs to the current θ; 2: T belongs to the last proposal */
};

struct bb_theta_args {
  double* W[2];
  double* X;
  uint8_t* par;
  double* start;
  double* startprop; /* x0° (start moves) */
  double *ll, *llprop, *logu, *xend, *xendprop;
  uint8_t *accepted, *xstale;
  unsigned long long *acc, *acc_theta;
  long long P, chain_offset;
  long long PT; /* chains per table row, padded to whole 128-byte lines */
  int S, N, NC, nbuf, jll, start_bcast, store_x;
  const double* gridtab[BB_MAXSEG]; /* rows (dt, sqrt dt) */
  const double* tt;
  double* theta[2];
  double* T;
  double* left[2];
  int which;   /* θ / left block the kernel works with: 0 current, 1 proposal */
  int propose; /* backward: draw θ° = θ + rw_sd ξ first */
  int mode;    /* forward without pCN: 0 = ll <- ll° (initialisation), 2 = parameter accept test,
                * 3 = recompute X of the chains whose X is stale (rejected proposals), nothing else */
  double rw_sd[BB_NTHETA];
  bb_philox_keys keys;
  uint32_t stream;
  double rho, rho2;
  bb_theta_spec spec;
  double log2pi;
  double prior_c0[BB_NTHETA];
  double seg_len[BB_MAXSEG];
  /* blocked update of segments s_lo .. s_hi-1 */
  int blk, s_lo, s_hi;
  double hzero;
  double* blkv;
  double* Xprop;
};

/* rows of blkv */
enum { BLK_LPN = 0, BLK_LPNO = 1, BLK_LLT = 2, BLK_LLO = 3, BLK_DIFF = 4, BLK_SEG = 5, BB_NBLK = 5 + 2 * BB_MAXSEG };

namespace {
using namespace bbk;

template <class M>
__device__ __forceinline__ void theta_model(const double* th, bb_model_dev& m) {
  constexpr int D = M::D;
  static_assert(M::SPARSE, "per-chain parameters: sparse-sigma models only");
#pragma unroll
  for (int k = 0; k < M::NTH; k++) m.par[k] = th[k];
  if (M::ID == BB_MODEL_FHN_DIAG || M::ID == BB_MODEL_FHN_HYPO) m.der[0] = 1.0 / m.par[0];
  /* a = sigma sigma' (bb_prepare_model) */
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++)
      m.der[8 + i * D + j] = (M::col(i) >= 0 && M::col(i) == M::col(j)) ? M::sig(m, i) * M::sig(m, j) : 0.0;
}

/* B~, beta~ of a segment from (θ, v) at time t: partialbridge_fitzhugh.jl:98-108 (x^2, x^3 are products, as Julia's
 * literal_pow), partialbridge_bolus3.jl:66-67 */
template <class M, int AUXK>
__device__ __forceinline__ void theta_aux(const bb_model_dev& m, double v, double t, double* Bt, double* be) {
  if constexpr (AUXK == BB_AUX_BOLUS) {
    static_assert(M::ID == BB_MODEL_BOLUS, "auxiliary registry: DiffusionAux belongs to the bolus model");
    const double alpha = m.par[0], beta = m.par[1], lam = m.par[2], mu = m.par[3];
    Bt[0] = -lam - beta; Bt[1] = mu; Bt[2] = lam; Bt[3] = -mu;
    be[0] = alpha * bb_dose(t);
    be[1] = 0.0;
  } else {
    static_assert(M::ID == BB_MODEL_FHN_DIAG || M::ID == BB_MODEL_FHN_HYPO, "auxiliary registry: FitzHugh-Nagumo");
    const double eps = m.par[0], s = m.par[1], gam = m.par[2], beta = m.par[3];
    const double ie = 1.0 / eps;
    if (AUXK == BB_AUX_FHN_MATCHING) {
      Bt[0] = ie; Bt[1] = -ie; Bt[2] = gam; Bt[3] = -1.0;
      be[0] = s / eps - (v * v * v) / eps;
      be[1] = beta;
    } else {
      Bt[0] = ie - (3.0 * (v * v)) / eps; Bt[1] = -ie; Bt[2] = gam; Bt[3] = -1.0;
      be[0] = s / eps + (2.0 * (v * v * v)) / eps;
      be[1] = beta;
    }
  }
}
/* a~: the target's a for the FitzHugh-Nagumo pairs; diag(σ1², σ2²) for DiffusionAux (bolus3.jl:68,71) */
template <class M, int AUXK>
__device__ __forceinline__ void theta_aux_a(const bb_model_dev& m, double* at) {
  if constexpr (AUXK == BB_AUX_BOLUS) {
    at[0] = m.par[4] * m.par[4]; at[1] = 0.0; at[2] = 0.0; at[3] = m.par[5] * m.par[5];
  } else {
#pragma unroll
    for (int k = 0; k < M::D * M::D; k++) at[k] = m.der[8 + k];
  }
}

__device__ __forceinline__ double theta_logu(const bb_philox_keys& k, uint32_t stream, uint64_t chain) {
  uint32_t o[4];
  bb_philox4x32_10(0xFFFFFFFDu, stream, (uint32_t)chain, (uint32_t)(chain >> 32), k, o);
  return (double)bb_logf(bb_unif(o[0]));
}

/* ------------------------------------------------------------------------------------------------ backward */
template <class M, int AUXK, int MOBS>
__global__ void __launch_bounds__(128) bb_theta_backward_kernel(const __grid_constant__ bb_theta_args a) {
  constexpr int D = M::D, K = D + D * D, NTH = M::NTH;
  const long long P = a.P, PT = a.PT;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const unsigned long long chain = (unsigned long long)(a.chain_offset + p);
  double th[NTH];
#pragma unroll
  for (int k = 0; k < NTH; k++) th[k] = a.theta[a.propose ? 0 : a.which][(long long)k * P + p];
  if (a.propose) {
    /* propose(σ, P): θ° = θ + σ .* randn  (bolus3.jl:239-242); the n-th updated parameter takes the n-th normal */
    float z[4];
    bb_normal_quad(a.keys, a.stream, (uint32_t)chain, (uint32_t)(chain >> 32), 0xFFFFFFFEu, z);
    int n = 0;
#pragma unroll
    for (int k = 0; k < NTH; k++) {
      if (a.rw_sd[k] != 0.0) {
        const float zz = n == 0 ? z[0] : (n == 1 ? z[1] : (n == 2 ? z[2] : z[3]));
        th[k] = th[k] + a.rw_sd[k] * (double)zz;
        n++;
      }
      a.theta[1][(long long)k * P + p] = th[k];
    }
  }
  /* starting point this pass refers to: x0, or the proposal x0° (bolus3.jl:311-318) */
  double xs[D];
#pragma unroll
  for (int k = 0; k < D; k++) xs[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + p];
  if (a.propose && a.spec.start_sd != 0.0) {
    float z[4];
    bb_normal_quad(a.keys, a.stream, (uint32_t)chain, (uint32_t)(chain >> 32), 0xFFFFFFFEu, z);
    uint32_t o[4];
    bb_philox4x32_10(0xFFFFFFFCu, a.stream, (uint32_t)chain, (uint32_t)(chain >> 32), a.keys, o);
    if (o[0] & 1u) {
#pragma unroll
      for (int k = 0; k < D; k++) xs[k] = xs[k] + (a.spec.start_sd * (double)z[3]) * a.spec.start_dir[k];
    }
#pragma unroll
    for (int k = 0; k < D; k++) a.startprop[(long long)k * P + p] = xs[k];
  } else if (a.which == 1 && a.spec.start_sd != 0.0) {
#pragma unroll
    for (int k = 0; k < D; k++) xs[k] = a.startprop[(long long)k * P + p];
  }
  bb_model_dev m;
  theta_model<M>(th, m);
  constexpr bool TDEP = (AUXK == BB_AUX_BOLUS); /* beta~ depends on t */
  double at[D * D];
  theta_aux_a<M, AUXK>(m, at);
  bool bad = false;
  double nu[D], Hp[D * D], Hc[D * D];
  /* νend = 0, Hend⁺ = I/ϵ, then the update with the last observation  (bolus3.jl:162-165) */
#pragma unroll
  for (int i = 0; i < D; i++) {
    nu[i] = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) Hp[i * D + j] = (i == j) ? 1.0 / a.spec.eps : 0.0;
  }
  const int S = a.S, N = a.N;
  /* a blocked update works on segments slo .. shi-1 (bolus3.jl:258-275) */
  const int shi = a.blk ? a.s_hi : S, slo = a.blk ? a.s_lo : 0;
  if (shi == S) {
    if (gpupdate_dev<D, MOBS>(nu, Hp, a.spec.L, a.spec.Sigma, a.spec.v[S - 1])) bad = true;
  } else {
    /* νend = XX[ind[1]].yy[end], Hend⁺ = Hzero⁺ (:272-273): the chain's own path at the block's right end */
    const double* xr = a.X + (((long long)(shi - 1) * a.NC + (N - 1) / BB_TC) * P + p) * (BB_TC * D) + ((N - 1) % BB_TC) * D;
#pragma unroll
    for (int i = 0; i < D; i++) {
      nu[i] = xr[i];
#pragma unroll
      for (int j = 0; j < D; j++) Hp[i * D + j] = (i == j) ? a.hzero : 0.0;
    }
  }
  double Cc = 0.0, trsum = 0.0;
  for (int s = shi - 1; s >= slo; s--) {
    double Bt[D * D], be[D];
    theta_aux<M, AUXK>(m, a.spec.v[s][0], 0.0, Bt, be);
    const aux_dev A{Bt, be, at, at, 1};
    if (minv<D>(Hp, Hc)) bad = true;
    double* Trow = a.T + ((long long)s * N + (N - 1)) * K * PT + p;
#pragma unroll
    for (int k = 0; k < D; k++) Trow[(long long)k * PT] = nu[k];
#pragma unroll
    for (int k = 0; k < D * D; k++) Trow[(long long)(D + k) * PT] = Hc[k];
    const double* tt = a.tt + (long long)s * N;
    for (int i = N - 2; i >= 0; i--) {
      const double dt = tt[i] - tt[i + 1];
      if constexpr (TDEP) {
        /* values at the Ralston stage times t, t + dt/2, t + 3dt/4 of the step that starts at t = tt[i+1]
         * (src/ode.jl:44-49); B~ and a~ are constant, a~(tt[i]) = a~ */
        double Bs[3][D * D], bs[3][D], as[3][D * D];
        const double tq[3] = {tt[i + 1], tt[i + 1] + 0.5 * dt, tt[i + 1] + 0.75 * dt};
#pragma unroll
        for (int q = 0; q < 3; q++) {
          theta_aux<M, AUXK>(m, a.spec.v[s][0], tq[q], Bs[q], bs[q]);
#pragma unroll
          for (int k = 0; k < D * D; k++) as[q][k] = at[k];
        }
        const aux_dev As{&Bs[0][0], &bs[0][0], &as[0][0], at, 0};
        if (nuH_step<D>(BB_ODE_LYAP, As, 0, dt, Hp, Hc, nu, Cc)) bad = true;
      } else {
        if (nuH_step<D>(BB_ODE_LYAP, A, i, dt, Hp, Hc, nu, Cc)) bad = true;
      }
      Trow -= (long long)K * PT;
#pragma unroll
      for (int k = 0; k < D; k++) Trow[(long long)k * PT] = nu[k];
#pragma unroll
      for (int k = 0; k < D * D; k++) Trow[(long long)(D + k) * PT] = Hc[k];
    }
    double tr = Bt[0];
#pragma unroll
    for (int i = 1; i < D; i++) tr += Bt[i * D + i];
    trsum += a.seg_len[s] * tr;
    if (s > slo) /* "on the left endpoint we don't need to do gpupdate" (:288) */
      if (gpupdate_dev<D, MOBS>(nu, Hp, a.spec.L, a.spec.Sigma, a.spec.v[s - 1])) bad = true;
  }
  if (a.blk) {
    /* start term of a block that contains the first segment (:311-320): x0° = x0 + (sd u) dir, always proposed */
    double lpn = 0.0, lpno = 0.0;
    if (slo == 0) {
      double x[D];
#pragma unroll
      for (int k = 0; k < D; k++) x[k] = xs[k] - nu[k];
      lpn = logpdfnormal_dev<D>(x, Hp, a.log2pi);
      lpno = lpn;
      if (a.spec.start_sd != 0.0) {
        float z[4];
        bb_normal_quad(a.keys, a.stream, (uint32_t)chain, (uint32_t)(chain >> 32), 0xFFFFFFFEu, z);
#pragma unroll
        for (int k = 0; k < D; k++) {
          xs[k] = xs[k] + (a.spec.start_sd * (double)z[3]) * a.spec.start_dir[k];
          a.startprop[(long long)k * P + p] = xs[k];
          x[k] = xs[k] - nu[k];
        }
        lpno = logpdfnormal_dev<D>(x, Hp, a.log2pi);
      }
      if (bad) lpn = lpno = nan("");
    } else if (bad) {
      lpno = nan("");
    }
    a.blkv[(long long)BLK_LPN * P + p] = lpn;
    a.blkv[(long long)BLK_LPNO * P + p] = lpno;
    return;
  }
  /* left end: ν(0), H⁺(0), C, logpdfnormal(x0 - ν(0), symmetrize(H⁺(0)))  (bolus3.jl:319), trace term, logπ(θ) */
  double x[D];
#pragma unroll
  for (int k = 0; k < D; k++) x[k] = xs[k] - nu[k];
  double lpn = logpdfnormal_dev<D>(x, Hp, a.log2pi);
  if (bad) lpn = nan("");
  double lpri = 0.0;
#pragma unroll
  for (int k = 0; k < NTH; k++) {
    if (a.spec.prior_kind[k] == BB_PRIOR_GAMMA) {
      const double xk = th[k], sh = a.spec.prior_a[k] - 1.0;
      double t = a.prior_c0[k];
      if (sh != 0.0) t += sh * log(xk);
      t -= xk / a.spec.prior_b[k];
      lpri += (xk > 0.0) ? t : -INFINITY;
    }
  }
  double* Lw = a.left[a.which] + p;
#pragma unroll
  for (int k = 0; k < D; k++) Lw[(long long)k * P] = nu[k];
#pragma unroll
  for (int k = 0; k < D * D; k++) Lw[(long long)(D + k) * P] = Hp[k];
  Lw[(long long)(K + 0) * P] = Cc;
  Lw[(long long)(K + 1) * P] = lpn;
  Lw[(long long)(K + 2) * P] = trsum;
  Lw[(long long)(K + 3) * P] = lpri;
}

/* ------------------------------------------------------------------------------------------------ forward */
template <class M, int AUXK, bool PCN>
__global__ void __launch_bounds__(BB_THREADS, 2) bb_theta_forward_kernel(const __grid_constant__ bb_theta_args a) {
  using CH = bb_chain<M, BB_GUIDE_NUH, 0, 1, 0>;
  constexpr int D = M::D, DP = M::DP, K = D + D * D, NTH = M::NTH, REC = CH::REC;
  constexpr int NPIECE = BB_TC * DP / 4;
  const long long P = a.P;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pc = p < P ? p : P - 1;
  const bool act = p < P && (PCN || a.mode != 3 || a.xstale[pc] != 0);
  const int lane = threadIdx.x & 31;
  const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
  const int S = a.S, N = a.N, NC = a.NC;

  double th[NTH];
#pragma unroll
  for (int k = 0; k < NTH; k++) th[k] = a.theta[a.which][(long long)k * P + pc];
  bb_model_dev m;
  theta_model<M>(th, m);

  const int par = a.par[pc];
  const int wbuf = PCN ? 1 - par : par;
  const bool sx = a.store_x != 0;
  const double* wr = a.W[par] + pc * (a.nbuf * BB_TC * DP);
  double* ww = a.W[wbuf] + pc * (a.nbuf * BB_TC * DP);
  double* xw = sx ? a.X + pc * (BB_TC * D) : nullptr;
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);

  /* X° (and W° of a pCN proposal) leave through shared memory as whole 128-byte lines written back to back:
   * 32-byte stores that drip out while the chain computes cost 15-20 % of DRAM bandwidth (tools/membench.cu).
   * Rows are XOR-swizzled by the chain index as in bb_chain.cuh (conflict-free 128-bit accesses). */
  extern __shared__ __align__(128) unsigned char th_smem[];
  constexpr bool XBUF = (D == 2);
  constexpr int XWIN = 16 / D; /* steps per 128-byte window */
  double* xbuf = reinterpret_cast<double*>(th_smem) + (size_t)threadIdx.x * 16;
  /* W° rows are completed in shared memory only for scalar noise: with d' = 2 the 64 KB they need would leave room
   * for ONE CTA per SM (8 warps); their pieces are stored directly instead and two CTAs fit */
  constexpr bool WBUF = PCN && (DP == 1);
  double* wrow = reinterpret_cast<double*>(th_smem) + (size_t)BB_THREADS * 16 + (size_t)threadIdx.x * (BB_TC * DP);

  double y[D], wprev[DP], w2[DP];
#pragma unroll
  for (int k = 0; k < D; k++)
    y[k] = (!PCN && a.mode == 2 && a.spec.start_sd != 0.0)
               ? a.startprop[(long long)k * P + pc]
               : (a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc]);
  double lltot = 0.0;

  /* The chain's table rows in the order they are used (g = s N + i, i = 0 .. N-2; row N-1 of a segment drives no
   * step) travel through a BB_TDEPTH-deep shared-memory ring: every thread copies its own K values of a row with
   * 8-byte cp.async (a warp's copies of one value are 256 contiguous bytes), BB_TDEPTH-1 steps before it reads them
   * back, so no warp waits on a global load and no register is tied up.  (Register double-buffering + L2 prefetch
   * left the kernel latency bound: long-scoreboard stalls of 11 warps per issue, 18 % of the lines fetched twice.) */
  double* tring = reinterpret_cast<double*>(th_smem) + (size_t)BB_THREADS * 16 +
                  (WBUF ? (size_t)BB_THREADS * (BB_TC * DP) : 0) + threadIdx.x;
#if !BB_TPAIR
  const double* Tp = a.T + pc;
#endif
  const long long PT = a.PT;
  const long long rowstride = (long long)K * PT;
  const long long nrows = (long long)S * (N - 1); /* rows that drive a step */
  const bool pair_ok = (p & ~1ll) < P; /* at least one chain of this lane pair exists (rows are padded to 16 chains) */
  long long gi = 0;                                /* next row to request: global index and position in its segment */
  int gi_i = 0;
  long long gi_n = 0;
  auto t_issue = [&]() { /* request the next row (if any); always commits a group */
    if (gi_n < nrows) {
      double* dst = tring + (size_t)(gi_n % BB_TDEPTH) * (K * BB_THREADS);
#if BB_TPAIR
      /* two neighbouring chains share the copies: the even lane moves the even values, the odd lane the odd ones,
       * 16 bytes (both chains) at a time -> cp.async.cg (L2 only: the small L1 that remains next to the shared-memory
       * carve-out stays with the time-grid table) and half the instructions */
      if (pair_ok) {
#pragma unroll
        for (int k = 0; k < K; k += 2)
          bb_cp_async16(dst + (k + (int)(threadIdx.x & 1)) * BB_THREADS - (threadIdx.x & 1),
                        a.T + gi * rowstride + (long long)(k + (int)(threadIdx.x & 1)) * PT + (p & ~1ll));
      }
#else
      const double* src = Tp + gi * rowstride;
#pragma unroll
      for (int k = 0; k < K; k++)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(bb_smem_u32(dst + k * BB_THREADS)),
                     "l"(src + (long long)k * PT)
                     : "memory");
#endif
      gi_n++;
      if (++gi_i == N - 1) { gi_i = 0; gi += 2; } else { gi += 1; }
    }
    bb_cp_async_commit();
  };
#pragma unroll 1
  for (int g = 0; g < BB_TDEPTH - 1; g++) t_issue();
  long long rcons = 0; /* rows consumed */

  double wq[4] = {0.0, 0.0, 0.0, 0.0};
#if BB_TWPF
  double wnx[DP][4]; /* the pieces of W of the next group of 4 steps */
#pragma unroll
  for (int i = 0; i < DP; i++) {
#pragma unroll
    for (int l = 0; l < 4; l++) wnx[i][l] = 0.0;
    if (act) bb_ld4(wr + 4 * i, wnx[i]);
  }
#endif
  for (int s = 0; s < S; s++) {
    const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
    const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
    double sc[D * D + D];
    theta_aux<M, AUXK>(m, a.spec.v[s][0], 0.0, sc, sc + D * D);
    constexpr bool TDEP = (AUXK == BB_AUX_BOLUS);
    const double* tts = a.tt + (long long)s * N;
    const double* gt = a.gridtab[s];
    double som = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) w2[k] = 0.0;
    for (int c = 0; c < NC; c++) {
      bb_rowout<D> xo;
      if (act && c + 1 < NC) { /* the chain's next row of W (whole lines) */
#pragma unroll
        for (int l = 0; l < (BB_TC * DP * 8 + 127) / 128; l++)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(wr + wstride + l * 16));
      }
#pragma unroll 1
      for (int h = 0; h < BB_TC / 4; h++) {
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
          const int slot = 4 * h + s4;
          const int j = c * BB_TC + slot;
          double wj[DP];
#pragma unroll
          for (int k = 0; k < DP; k++) {
            const int mm = s4 * DP + k;
            if ((mm & 3) == 0) {
              const int q = h * DP + (mm >> 2);
#if BB_TWPF
              /* the piece was requested one group of 4 steps ago; request the same piece of the next group now */
#pragma unroll
              for (int i = 0; i < 4; i++) wq[i] = wnx[mm >> 2][i];
              {
                const bool last = (s == S - 1) && (c == NC - 1) && (h == BB_TC / 4 - 1);
                const double* nxt = (h == BB_TC / 4 - 1) ? wr + wstride + 4 * (mm >> 2) : wr + 4 * (q + DP);
                if (act && !last) bb_ld4(nxt, wnx[mm >> 2]);
              }
#else
              if (act) bb_ld4(wr + 4 * q, wq);
#endif
              if constexpr (PCN) {
                float z[4];
                bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + q), z);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  const int sl = 4 * h + (mm + i) / DP, kk = (mm + i) % DP;
                  const int jj = c * BB_TC + sl;
                  const double rootdt = gt[2 * jj + 1];
                  /* W2[j] = W2[j-1] + sqrt(dt) xi ;  W°[j] = rho W[j] + sqrt(1-rho^2) W2[j] */
                  if (jj != 0) w2[kk] = fma(rootdt, (double)z[i], w2[kk]);
                  wq[i] = fma(a.rho2, w2[kk], a.rho * wq[i]);
                }
                if constexpr (WBUF) bb_sts4_swz(wrow, 2 * q, threadIdx.x & 7, wq);
                else if (act) bb_st4(ww + 4 * q, wq[0], wq[1], wq[2], wq[3]);
              }
            }
            wj[k] = wq[mm & 3];
          }
          if (j == 0) {
#pragma unroll
            for (int k = 0; k < DP; k++) wprev[k] = wj[k];
          } else if (j < N) {
            double R[REC];
            R[0] = gt[2 * j];
            R[1] = 0.0;
            bb_cp_async_wait<BB_TDEPTH - 2>(); /* this step's row has landed */
#if BB_TPAIR
            __syncwarp(); /* ... including the half the neighbouring lane copied */
#endif
            const double* Ts = tring + (size_t)(rcons % BB_TDEPTH) * (K * BB_THREADS);
#pragma unroll
            for (int k = 0; k < K; k++) R[2 + k] = Ts[k * BB_THREADS];
            rcons++;
            t_issue(); /* refill the stage read one step ago */
            double dw[DP], bd[D];
#pragma unroll
            for (int k = 0; k < DP; k++) {
              dw[k] = wj[k] - wprev[k];
              wprev[k] = wj[k];
            }
            if constexpr (TDEP) { /* b(tt[i], x) and b~(tt[i], x) at the left end of the step, i = j - 1 */
              m.der[1] = m.par[0] * bb_dose(tts[j - 1]);
              sc[D * D] = m.der[1];
            }
            CH::drift(m, R, sc, y, R[0], j <= a.jll, som, bd);
            bb_em_update<M>(m, bd, R[0], dw, y);
          }
          if (sx) {
            if constexpr (XBUF) {
              const int ls = (4 * h + s4) % XWIN;
              *reinterpret_cast<double2*>(xbuf + 2 * ((ls ^ threadIdx.x) & 7)) = make_double2(y[0], y[1]);
            } else {
              xo.put(xw + 4 * h * D, s4, y, act);
            }
          }
        }
        if constexpr (XBUF) {
          if (sx && act && ((4 * h + 3) % XWIN) == XWIN - 1) { /* a 128-byte window of X° is complete */
            double* xdst = xw + (4 * h + 4 - XWIN) * D;
#pragma unroll
            for (int q = 0; q < 4; q++) {
              double v[4];
              bb_lds4_swz(xbuf, 2 * q, threadIdx.x & 7, v);
              bb_st4(xdst + 4 * q, v[0], v[1], v[2], v[3]);
            }
          }
        }
      }
      if constexpr (WBUF) {
        if (act) { /* the chain's row of W° is complete: write its 128 d' bytes back to back */
#pragma unroll
          for (int q = 0; q < NPIECE; q++) {
            double v[4];
            bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, v);
            bb_st4(ww + 4 * q, v[0], v[1], v[2], v[3]);
          }
        }
      }
      wr += wstride;
      ww += wstride;
      if (sx) xw += xstride;
    }
    lltot += som;
  }

  if constexpr (PCN) {
    /* accept iff log(U) <= ll° - ll   (test/partialbridgenuH.jl:183) */
    const double logu = bb_accept_logu(a.keys, a.stream, chain);
    const bool ok = act && (logu <= lltot - a.ll[pc]);
    if (act) {
      a.llprop[p] = lltot;
      a.logu[p] = logu;
      a.accepted[p] = ok ? 1 : 0;
      a.xstale[p] = sx ? (ok ? 0 : 1) : (uint8_t)(a.xstale[p] | (ok ? 1 : 0));
#pragma unroll
      for (int k = 0; k < D; k++) a.xendprop[(long long)k * P + p] = y[k];
      if (ok) {
        a.ll[p] = lltot;
        a.par[p] = (uint8_t)(1 - par);
#pragma unroll
        for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
      }
    }
    const unsigned mk = __ballot_sync(0xFFFFFFFFu, ok);
    if (lane == 0 && mk) atomicAdd(a.acc, (unsigned long long)__popc(mk));
  } else if (a.mode == 2) {
    /* diffll (bolus3.jl:319,331-336) and the MH step (:340): θ, the left-end values, ll follow on accept */
    const double logu = theta_logu(a.keys, a.stream, chain);
    const double* Lc = a.left[0] + pc;
    const double* Lo = a.left[1] + pc;
    double diff = Lo[(long long)(K + 1) * P] - Lc[(long long)(K + 1) * P];
    diff += lltot - a.ll[pc];
    diff += ((Lo[(long long)(K + 2) * P] - Lc[(long long)(K + 2) * P]) + Lo[(long long)(K + 3) * P]) -
            Lc[(long long)(K + 3) * P];
    const bool ok = act && (logu <= diff);
    if (act) {
      a.llprop[p] = lltot;
      a.logu[p] = logu;
      a.accepted[p] = ok ? 1 : 0;
      a.xstale[p] = sx ? (ok ? 0 : 1) : (uint8_t)(a.xstale[p] | (ok ? 1 : 0));
#pragma unroll
      for (int k = 0; k < D; k++) a.xendprop[(long long)k * P + p] = y[k];
      if (ok) {
        a.ll[p] = lltot;
#pragma unroll
        for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
#pragma unroll
        for (int k = 0; k < NTH; k++) a.theta[0][(long long)k * P + p] = th[k];
#pragma unroll
        for (int k = 0; k < K + 4; k++) a.left[0][(long long)k * P + p] = Lo[(long long)k * P];
        if (a.spec.start_sd != 0.0) {
#pragma unroll
          for (int k = 0; k < D; k++) a.start[(long long)k * P + p] = a.startprop[(long long)k * P + p];
        }
      }
    }
    const unsigned mk = __ballot_sync(0xFFFFFFFFu, ok);
    if (lane == 0 && mk) atomicAdd(a.acc_theta, (unsigned long long)__popc(mk));
  } else if (a.mode == 3) {
    if (act) a.xstale[p] = 0;
  } else if (act) {
    a.ll[p] = lltot;
    a.xstale[p] = sx ? 0 : 1;
#pragma unroll
    for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
  }
}


/* ------------------------------------------------------------------------------------------------ blocked update
 * The path-update branch of the multi-segment sampler (partialbridge_bolus3.jl:258-355, `updateparams == false`) for
 * the block of segments s_lo .. s_hi-1 after bb_theta_backward_kernel built the block's tables:
 *   W°[i] = ρ W[i] + sqrt(1-ρ²) W2 (:304-305); XXtemp[i] = solve!(Euler(), xstart, W[i], Q[i]) and
 *   XXᵒ[i] = solve!(Euler(), xstartᵒ, W°[i], Qᵒ[i]) with Qᵒ = Q (:324-325) advance side by side on the same table
 *   rows; diffll = start term + Σ_{i in ind} (ll°[i] - ll_temp[i]) in the script's (descending) order (:331-333);
 *   log(rand()) <= diffll (:340).  W° goes to the chain's other W buffer and X° to Xprop; the commit kernel below
 *   moves both into place for the chains that accept (the script's swap of XX[i], WW[i] for i in ind, :342-345). */
template <class M, int AUXK>
__global__ void __launch_bounds__(BB_THREADS) bb_theta_block_kernel(const __grid_constant__ bb_theta_args a) {
  using CH = bb_chain<M, BB_GUIDE_NUH, 0, 1, 0>;
  constexpr int D = M::D, DP = M::DP, K = D + D * D, NTH = M::NTH, REC = CH::REC;
  constexpr int NPIECE = BB_TC * DP / 4;
  const long long P = a.P;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pc = p < P ? p : P - 1;
  const bool act = p < P;
  const int lane = threadIdx.x & 31;
  const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
  const int S = a.S, N = a.N, NC = a.NC, slo = a.s_lo, shi = a.s_hi;

  double th[NTH];
#pragma unroll
  for (int k = 0; k < NTH; k++) th[k] = a.theta[0][(long long)k * P + pc];
  bb_model_dev m;
  theta_model<M>(th, m);

  const int par = a.par[pc];
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
  const double* wr = a.W[par] + pc * (a.nbuf * BB_TC * DP) + (long long)slo * NC * wstride;
  double* ww = a.W[1 - par] + pc * (a.nbuf * BB_TC * DP) + (long long)slo * NC * wstride;
  double* xw = a.Xprop + pc * (BB_TC * D) + (long long)slo * NC * xstride;

  /* both passes start from the current path's value at the block's left end; a block that contains the first segment
   * starts the proposal at x0° (bb_theta_backward_kernel drew it) */
  double yt[D], yo[D];
  if (slo == 0) {
#pragma unroll
    for (int k = 0; k < D; k++) {
      yt[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc];
      yo[k] = a.spec.start_sd != 0.0 ? a.startprop[(long long)k * P + pc] : yt[k];
    }
  } else {
    const double* xl = a.X + (((long long)(slo - 1) * NC + (N - 1) / BB_TC) * P + pc) * (BB_TC * D) + ((N - 1) % BB_TC) * D;
#pragma unroll
    for (int k = 0; k < D; k++) yt[k] = yo[k] = xl[k];
  }

  /* table rows through the shared-memory ring of bb_theta_forward_kernel */
  extern __shared__ __align__(128) unsigned char th_smem[];
  double* tring = reinterpret_cast<double*>(th_smem) + threadIdx.x;
  const long long PT = a.PT;
  const long long rowstride = (long long)K * PT;
  const long long nrows = (long long)(shi - slo) * (N - 1);
  const bool pair_ok = (p & ~1ll) < P;
  long long gi = (long long)slo * N;
  int gi_i = 0;
  long long gi_n = 0;
  auto t_issue = [&]() {
    if (gi_n < nrows) {
      double* dst = tring + (size_t)(gi_n % BB_TDEPTH) * (K * BB_THREADS);
      if (pair_ok) {
#pragma unroll
        for (int k = 0; k < K; k += 2)
          bb_cp_async16(dst + (k + (int)(threadIdx.x & 1)) * BB_THREADS - (threadIdx.x & 1),
                        a.T + gi * rowstride + (long long)(k + (int)(threadIdx.x & 1)) * PT + (p & ~1ll));
      }
      gi_n++;
      if (++gi_i == N - 1) { gi_i = 0; gi += 2; } else { gi += 1; }
    }
    bb_cp_async_commit();
  };
#pragma unroll 1
  for (int g = 0; g < BB_TDEPTH - 1; g++) t_issue();
  long long rcons = 0;

  double dseg[BB_MAXSEG];
  double llt = 0.0, llo = 0.0;
  double wq[4] = {0.0, 0.0, 0.0, 0.0}, wqo[4] = {0.0, 0.0, 0.0, 0.0};
  double wprev_t[DP], wprev_o[DP], w2[DP];
  for (int s = slo; s < shi; s++) {
    const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
    const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
    double sc[D * D + D];
    theta_aux<M, AUXK>(m, a.spec.v[s][0], 0.0, sc, sc + D * D);
    constexpr bool TDEP = (AUXK == BB_AUX_BOLUS);
    const double* tts = a.tt + (long long)s * N;
    const double* gt = a.gridtab[s];
    double som_t = 0.0, som_o = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) w2[k] = 0.0;
    for (int c = 0; c < NC; c++) {
      bb_rowout<D> xo;
#pragma unroll 1
      for (int h = 0; h < BB_TC / 4; h++) {
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
          const int j = c * BB_TC + 4 * h + s4;
          double wjt[DP], wjo[DP];
#pragma unroll
          for (int k = 0; k < DP; k++) {
            const int mm = s4 * DP + k;
            if ((mm & 3) == 0) {
              const int q = h * DP + (mm >> 2);
              if (act) bb_ld4(wr + 4 * q, wq);
              float z[4];
              bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + q), z);
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int sl = 4 * h + (mm + i) / DP, kk = (mm + i) % DP;
                const int jj = c * BB_TC + sl;
                const double rootdt = gt[2 * jj + 1];
                if (jj != 0) w2[kk] = fma(rootdt, (double)z[i], w2[kk]);
                wqo[i] = fma(a.rho2, w2[kk], a.rho * wq[i]);
              }
              if (act) bb_st4(ww + 4 * q, wqo[0], wqo[1], wqo[2], wqo[3]);
            }
            wjt[k] = wq[mm & 3];
            wjo[k] = wqo[mm & 3];
          }
          if (j == 0) {
#pragma unroll
            for (int k = 0; k < DP; k++) { wprev_t[k] = wjt[k]; wprev_o[k] = wjo[k]; }
          } else if (j < N) {
            double R[REC];
            R[0] = gt[2 * j];
            R[1] = 0.0;
            bb_cp_async_wait<BB_TDEPTH - 2>();
            __syncwarp();
            const double* Ts = tring + (size_t)(rcons % BB_TDEPTH) * (K * BB_THREADS);
#pragma unroll
            for (int k = 0; k < K; k++) R[2 + k] = Ts[k * BB_THREADS];
            rcons++;
            t_issue();
            double dwt[DP], dwo[DP], bd[D];
#pragma unroll
            for (int k = 0; k < DP; k++) {
              dwt[k] = wjt[k] - wprev_t[k];
              dwo[k] = wjo[k] - wprev_o[k];
              wprev_t[k] = wjt[k];
              wprev_o[k] = wjo[k];
            }
            if constexpr (TDEP) {
              m.der[1] = m.par[0] * bb_dose(tts[j - 1]);
              sc[D * D] = m.der[1];
            }
            CH::drift(m, R, sc, yt, R[0], j <= a.jll, som_t, bd);
            bb_em_update<M>(m, bd, R[0], dwt, yt);
            CH::drift(m, R, sc, yo, R[0], j <= a.jll, som_o, bd);
            bb_em_update<M>(m, bd, R[0], dwo, yo);
          }
          xo.put(xw + 4 * h * D, s4, yo, act);
        }
      }
      wr += wstride;
      ww += wstride;
      xw += xstride;
    }
    dseg[s] = som_o - som_t;
    llt += som_t;
    llo += som_o;
    if (act) {
      a.blkv[(long long)(BLK_SEG + 2 * s) * P + p] = som_t;
      a.blkv[(long long)(BLK_SEG + 2 * s + 1) * P + p] = som_o;
    }
  }

  const double logu = bb_accept_logu(a.keys, a.stream, chain);
  double diff = a.blkv[(long long)BLK_LPNO * P + pc] - a.blkv[(long long)BLK_LPN * P + pc];
  for (int s = shi - 1; s >= slo; s--) diff += dseg[s];
  const bool ok = act && (logu <= diff);
  if (act) {
    a.blkv[(long long)BLK_LLT * P + p] = llt;
    a.blkv[(long long)BLK_LLO * P + p] = llo;
    a.blkv[(long long)BLK_DIFF * P + p] = diff;
    a.llprop[p] = llo;
    a.logu[p] = logu;
    a.accepted[p] = ok ? 1 : 0;
#pragma unroll
    for (int k = 0; k < D; k++) a.xendprop[(long long)k * P + p] = yo[k];
  }
  const unsigned mk = __ballot_sync(0xFFFFFFFFu, ok);
  if (lane == 0 && mk) atomicAdd(a.acc, (unsigned long long)__popc(mk));
}

/* the accepted chains take their proposal: W°, X° of the block's segments move into the current buffers */
template <int D, int DP>
__global__ void __launch_bounds__(BB_THREADS) bb_theta_block_commit_kernel(const __grid_constant__ bb_theta_args a) {
  const long long P = a.P;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P || !a.accepted[p]) return;
  const int NC = a.NC, slo = a.s_lo, shi = a.s_hi;
  const int par = a.par[p];
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
  const double* wsrc = a.W[1 - par] + p * (a.nbuf * BB_TC * DP) + (long long)slo * NC * wstride;
  double* wdst = a.W[par] + p * (a.nbuf * BB_TC * DP) + (long long)slo * NC * wstride;
  const double* xsrc = a.Xprop + p * (BB_TC * D) + (long long)slo * NC * xstride;
  double* xdst = a.X + p * (BB_TC * D) + (long long)slo * NC * xstride;
  for (int r = 0; r < (shi - slo) * NC; r++) {
#pragma unroll
    for (int q = 0; q < BB_TC * DP / 4; q++) {
      double v[4];
      bb_ld4(wsrc + 4 * q, v);
      bb_st4(wdst + 4 * q, v[0], v[1], v[2], v[3]);
    }
#pragma unroll
    for (int q = 0; q < BB_TC * D / 4; q++) {
      double v[4];
      bb_ld4(xsrc + 4 * q, v);
      bb_st4(xdst + 4 * q, v[0], v[1], v[2], v[3]);
    }
    wsrc += wstride; wdst += wstride; xsrc += xstride; xdst += xstride;
  }
  if (slo == 0 && a.spec.start_sd != 0.0) {
#pragma unroll
    for (int k = 0; k < D; k++) a.start[(long long)k * P + p] = a.startprop[(long long)k * P + p];
  }
  if (shi == a.S) {
#pragma unroll
    for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = a.xendprop[(long long)k * P + p];
  }
}

typedef void (*theta_kernel_fn)(const bb_theta_args);

template <class M>
static theta_kernel_fn lookup_backward(int auxk, int mobs) {
  if constexpr (M::ID == BB_MODEL_BOLUS) {
    if (auxk != BB_AUX_BOLUS) return nullptr;
    if (mobs == 1) return &bb_theta_backward_kernel<M, BB_AUX_BOLUS, 1>;
    if (mobs == 2) return &bb_theta_backward_kernel<M, BB_AUX_BOLUS, 2>;
    return nullptr;
  } else {
  if (auxk == BB_AUX_FHN_MATCHING) {
    if (mobs == 1) return &bb_theta_backward_kernel<M, BB_AUX_FHN_MATCHING, 1>;
    if (mobs == 2) return &bb_theta_backward_kernel<M, BB_AUX_FHN_MATCHING, 2>;
  }
  if (auxk == BB_AUX_FHN_LINEARISED_END) {
    if (mobs == 1) return &bb_theta_backward_kernel<M, BB_AUX_FHN_LINEARISED_END, 1>;
    if (mobs == 2) return &bb_theta_backward_kernel<M, BB_AUX_FHN_LINEARISED_END, 2>;
  }
  return nullptr;
  }
}
template <class M>
static theta_kernel_fn lookup_forward(int auxk, bool pcn) {
  if constexpr (M::ID == BB_MODEL_BOLUS) {
    if (auxk != BB_AUX_BOLUS) return nullptr;
    return pcn ? &bb_theta_forward_kernel<M, BB_AUX_BOLUS, true> : &bb_theta_forward_kernel<M, BB_AUX_BOLUS, false>;
  } else {
  if (auxk == BB_AUX_FHN_MATCHING)
    return pcn ? &bb_theta_forward_kernel<M, BB_AUX_FHN_MATCHING, true>
               : &bb_theta_forward_kernel<M, BB_AUX_FHN_MATCHING, false>;
  if (auxk == BB_AUX_FHN_LINEARISED_END)
    return pcn ? &bb_theta_forward_kernel<M, BB_AUX_FHN_LINEARISED_END, true>
               : &bb_theta_forward_kernel<M, BB_AUX_FHN_LINEARISED_END, false>;
  return nullptr;
  }
}

template <class M>
static theta_kernel_fn lookup_block(int auxk) {
  if constexpr (M::ID == BB_MODEL_BOLUS) {
    return auxk == BB_AUX_BOLUS ? &bb_theta_block_kernel<M, BB_AUX_BOLUS> : nullptr;
  } else {
    if (auxk == BB_AUX_FHN_MATCHING) return &bb_theta_block_kernel<M, BB_AUX_FHN_MATCHING>;
    if (auxk == BB_AUX_FHN_LINEARISED_END) return &bb_theta_block_kernel<M, BB_AUX_FHN_LINEARISED_END>;
    return nullptr;
  }
}

static int fill_args(bb_ens* e, bb_theta_args& a) {
  bb_theta* t = e->th;
  memset(&a, 0, sizeof(a));
  a.W[0] = e->W[0]; a.W[1] = e->W[1]; a.X = e->X;
  a.par = e->par; a.start = e->start; a.start_bcast = e->start_bcast; a.startprop = t->startprop;
  a.ll = e->ll; a.llprop = e->llprop; a.logu = e->logu; a.xend = e->xend; a.xendprop = e->xendprop;
  a.accepted = e->accepted; a.xstale = e->xstale; a.acc = e->acc; a.acc_theta = t->acc;
  a.P = e->P; a.PT = t->PT; a.chain_offset = e->chain_offset; a.S = e->S; a.N = e->N; a.NC = e->NC; a.nbuf = e->nbuf;
  for (int s = 0; s < e->S; s++) {
    if (!e->gridtab[s] || (int)e->tt[s].size() != e->N) return BB_ERR_ARG; /* bb_ens_set_grid first */
    a.gridtab[s] = e->gridtab[s];
    a.seg_len[s] = e->tt[s][e->N - 1] - e->tt[s][0];
  }
  a.tt = t->tt;
  a.theta[0] = t->theta[0]; a.theta[1] = t->theta[1];
  a.T = t->T;
  a.left[0] = t->left[0]; a.left[1] = t->left[1];
  a.blkv = t->blkv; a.Xprop = t->Xprop;
  a.spec = t->spec;
  a.log2pi = log(2 * M_PI);
  for (int k = 0; k < BB_NTHETA; k++)
    if (t->spec.prior_kind[k] == BB_PRIOR_GAMMA) /* -lgamma(a) - a log(b) */
      a.prior_c0[k] = -lgamma(t->spec.prior_a[k]) - t->spec.prior_a[k] * log(t->spec.prior_b[k]);
  return BB_OK;
}

static int upload_grids(bb_ens* e) {
  bb_theta* t = e->th;
  std::vector<double> h((size_t)e->S * e->N);
  for (int s = 0; s < e->S; s++) {
    if ((int)e->tt[s].size() != e->N) return BB_ERR_ARG;
    memcpy(h.data() + (size_t)s * e->N, e->tt[s].data(), sizeof(double) * e->N);
  }
  BB_CUDA(cudaMemcpyAsync(t->tt, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return BB_OK;
}

struct blk_spec {
  int s_lo, s_hi;
  double hzero;
};

static int run_backward(bb_ens* e, int which, bool propose, const double* rw_sd, uint64_t seed, uint32_t stream,
                        const blk_spec* blk = nullptr) {
  bb_theta* t = e->th;
  bb_ctx* c = e->ctx;
  bb_theta_args a;
  int rc = fill_args(e, a);
  if (rc != BB_OK) return rc;
  rc = upload_grids(e); /* the grids may have changed since attach */
  if (rc != BB_OK) return rc;
  a.which = which;
  a.propose = propose ? 1 : 0;
  if (blk) {
    a.blk = 1; a.s_lo = blk->s_lo; a.s_hi = blk->s_hi; a.hzero = blk->hzero;
  }
  if (rw_sd) memcpy(a.rw_sd, rw_sd, sizeof(a.rw_sd));
  bb_philox_key_schedule(seed, a.keys);
  a.stream = stream;
  theta_kernel_fn fn = t->model.id == BB_MODEL_FHN_HYPO
                           ? lookup_backward<MFhnHypo>(t->spec.aux_kind, t->spec.m)
                           : (t->model.id == BB_MODEL_FHN_DIAG ? lookup_backward<MFhnDiag>(t->spec.aux_kind, t->spec.m)
                                                               : lookup_backward<MBolus>(t->spec.aux_kind, t->spec.m));
  if (!fn) return BB_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((e->P + 127) / 128);
  fn<<<grid, 128, 0, c->stream>>>(a);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    bb_set_cuda_error(err, "bb_theta_backward_kernel launch");
    return BB_ERR_CUDA;
  }
  c->launches++;
  t->tstate = blk ? 0 : (which == 0 ? 1 : 2); /* a block's tables are conditioned on the chains' own paths */
  return BB_OK;
}

static int run_block(bb_ens* e, const blk_spec& blk, int skip, double rho, uint64_t seed, uint32_t stream) {
  bb_theta* t = e->th;
  bb_ctx* c = e->ctx;
  bb_theta_args a;
  int rc = fill_args(e, a);
  if (rc != BB_OK) return rc;
  a.blk = 1; a.s_lo = blk.s_lo; a.s_hi = blk.s_hi; a.hzero = blk.hzero;
  a.jll = e->N - 1 - skip;
  a.store_x = 1;
  a.rho = rho;
  a.rho2 = sqrt(1 - rho * rho);
  bb_philox_key_schedule(seed, a.keys);
  a.stream = stream;
  theta_kernel_fn fn = t->model.id == BB_MODEL_FHN_HYPO
                           ? lookup_block<MFhnHypo>(t->spec.aux_kind)
                           : (t->model.id == BB_MODEL_FHN_DIAG ? lookup_block<MFhnDiag>(t->spec.aux_kind)
                                                               : lookup_block<MBolus>(t->spec.aux_kind));
  if (!fn) return BB_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((e->P + BB_THREADS - 1) / BB_THREADS);
  const size_t smem = (size_t)BB_TDEPTH * t->K * BB_THREADS * 8;
  BB_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fn<<<grid, BB_THREADS, smem, c->stream>>>(a);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) {
    c->launches++;
    if (e->dp == 1) bb_theta_block_commit_kernel<2, 1><<<grid, BB_THREADS, 0, c->stream>>>(a);
    else bb_theta_block_commit_kernel<2, 2><<<grid, BB_THREADS, 0, c->stream>>>(a);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) {
    bb_set_cuda_error(err, "bb_theta_block_kernel launch");
    return BB_ERR_CUDA;
  }
  c->launches++;
  e->x_maybe_stale = false; /* X holds the current path of every chain */
  return BB_OK;
}

static int run_forward(bb_ens* e, bool pcn, int which, int mode, int skip, bool store_x, double rho, uint64_t seed,
                       uint32_t stream) {
  bb_theta* t = e->th;
  bb_ctx* c = e->ctx;
  if (skip < 0) return BB_ERR_ARG;
  if (store_x && !e->X) return BB_ERR_ARG;
  if (pcn && !(e->flags & BB_ENS_DOUBLE_BUFFER)) return BB_ERR_ARG;
  bb_theta_args a;
  int rc = fill_args(e, a);
  if (rc != BB_OK) return rc;
  a.which = which;
  a.mode = mode;
  a.jll = e->N - 1 - skip;
  a.store_x = store_x ? 1 : 0;
  a.rho = rho;
  a.rho2 = sqrt(1 - rho * rho);
  bb_philox_key_schedule(seed, a.keys);
  a.stream = stream;
  theta_kernel_fn fn = t->model.id == BB_MODEL_FHN_HYPO
                           ? lookup_forward<MFhnHypo>(t->spec.aux_kind, pcn)
                           : (t->model.id == BB_MODEL_FHN_DIAG ? lookup_forward<MFhnDiag>(t->spec.aux_kind, pcn)
                                                               : lookup_forward<MBolus>(t->spec.aux_kind, pcn));
  if (!fn) return BB_ERR_UNSUPPORTED;
  const unsigned grid = (unsigned)((e->P + BB_THREADS - 1) / BB_THREADS);
  const size_t smem = (size_t)BB_THREADS * 128 + ((pcn && e->dp == 1) ? (size_t)BB_THREADS * BB_TC * e->dp * 8 : 0) +
                      (size_t)BB_TDEPTH * t->K * BB_THREADS * 8;
  BB_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fn<<<grid, BB_THREADS, smem, c->stream>>>(a);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    bb_set_cuda_error(err, "bb_theta_forward_kernel launch");
    return BB_ERR_CUDA;
  }
  c->launches++;
  if (pcn || mode == 2 || !store_x) e->x_maybe_stale = true;
  else e->x_maybe_stale = false; /* modes 0 and 3 with X stored leave X current for every chain */
  return BB_OK;
}

}  // namespace

void bb_theta_invalidate(bb_ens* e) {
  if (e->th) e->th->tstate = 0;
}

void bb_theta_free(bb_ens* e) {
  bb_theta* t = e->th;
  if (!t) return;
  void* ptrs[] = {t->theta[0], t->theta[1], t->T, t->left[0], t->left[1], t->tt, t->acc, t->startprop, t->blkv, t->Xprop};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete t;
  e->th = nullptr;
}

template <class T>
static int th_alloc(bb_ens* e, T** p, size_t count) {
  BB_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  BB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), e->ctx->stream));
  e->bytes += (int64_t)(count * sizeof(T));
  return BB_OK;
}

extern "C" int bb_theta_attach(bb_ens* e, const bb_model* model, const bb_theta_spec* spec) {
  if (!e || !model || !spec) return BB_ERR_ARG;
  if (e->th) return BB_ERR_ARG;
  if (model->id != BB_MODEL_FHN_HYPO && model->id != BB_MODEL_FHN_DIAG && model->id != BB_MODEL_BOLUS)
    return BB_ERR_UNSUPPORTED;
  if (model->d != e->d || model->dprime != e->dp) return BB_ERR_MODEL;
  if (model->d != 2 || model->dprime != (model->id == BB_MODEL_FHN_HYPO ? 1 : 2)) return BB_ERR_MODEL;
  if (spec->m < 1 || spec->m > e->d) return BB_ERR_ASSERT_M;
  if (model->id == BB_MODEL_BOLUS ? spec->aux_kind != BB_AUX_BOLUS
                                  : (spec->aux_kind != BB_AUX_FHN_MATCHING && spec->aux_kind != BB_AUX_FHN_LINEARISED_END))
    return BB_ERR_UNSUPPORTED;
  if (!(spec->start_sd >= 0.0)) return BB_ERR_ARG;
  for (int k = 0; k < BB_NTHETA; k++) {
    if (spec->prior_kind[k] != BB_PRIOR_FLAT && spec->prior_kind[k] != BB_PRIOR_GAMMA) return BB_ERR_ARG;
    if (spec->prior_kind[k] == BB_PRIOR_GAMMA && !(spec->prior_a[k] > 0 && spec->prior_b[k] > 0)) return BB_ERR_ARG;
  }
  BB_CUDA(cudaSetDevice(e->ctx->device));
  bb_theta* t = new (std::nothrow) bb_theta();
  if (!t) return BB_ERR_NOMEM;
  e->th = t;
  t->model = *model;
  t->spec = *spec;
  const int d = e->d;
  t->K = d + d * d;
  t->NL = t->K + 4;
  const size_t P = (size_t)e->P;
  int rc = th_alloc(e, &t->theta[0], (size_t)BB_NTHETA * P);
  if (rc == BB_OK) rc = th_alloc(e, &t->theta[1], (size_t)BB_NTHETA * P);
  t->PT = (e->P + 15) & ~15ll;
  if (rc == BB_OK) rc = th_alloc(e, &t->T, (size_t)e->S * e->N * t->K * (size_t)t->PT);
  if (rc == BB_OK) rc = th_alloc(e, &t->left[0], (size_t)t->NL * P);
  if (rc == BB_OK) rc = th_alloc(e, &t->left[1], (size_t)t->NL * P);
  if (rc == BB_OK) rc = th_alloc(e, &t->tt, (size_t)e->S * e->N);
  if (rc == BB_OK) rc = th_alloc(e, &t->acc, 1);
  if (rc == BB_OK && spec->start_sd != 0.0) rc = th_alloc(e, &t->startprop, (size_t)d * P);
  if (rc != BB_OK) {
    bb_theta_free(e);
    return rc;
  }
  std::vector<double> h((size_t)BB_NTHETA * P);
  for (int k = 0; k < BB_NTHETA; k++)
    for (size_t p = 0; p < P; p++) h[(size_t)k * P + p] = model->par[k];
  BB_CUDA(cudaMemcpyAsync(t->theta[0], h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  return BB_OK;
}

extern "C" int bb_theta_set(bb_ens* e, int64_t p0, int64_t np, const double* theta) {
  if (!e || !e->th || !theta || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  std::vector<double> h((size_t)np);
  for (int k = 0; k < BB_NTHETA; k++) {
    for (int64_t p = 0; p < np; p++) h[(size_t)p] = theta[(size_t)p * BB_NTHETA + k];
    BB_CUDA(cudaMemcpyAsync(e->th->theta[0] + (size_t)k * e->P + p0, h.data(), sizeof(double) * np,
                            cudaMemcpyHostToDevice, e->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  }
  e->th->tstate = 0;
  return BB_OK;
}

extern "C" int bb_theta_get(bb_ens* e, int which, int64_t p0, int64_t np, double* theta) {
  if (!e || !e->th || !theta || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (which != BB_CUR && which != BB_PROP) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  std::vector<double> h((size_t)np * BB_NTHETA);
  for (int k = 0; k < BB_NTHETA; k++)
    BB_CUDA(cudaMemcpyAsync(h.data() + (size_t)k * np, e->th->theta[which] + (size_t)k * e->P + p0,
                            sizeof(double) * np, cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  for (int64_t p = 0; p < np; p++)
    for (int k = 0; k < BB_NTHETA; k++) theta[(size_t)p * BB_NTHETA + k] = h[(size_t)k * np + p];
  return BB_OK;
}

extern "C" int bb_theta_get_start(bb_ens* e, int which, int64_t p0, int64_t np, double* x0) {
  if (!e || !e->th || !x0 || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (which != BB_CUR && which != BB_PROP) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int d = e->d;
  if (which == BB_CUR && e->start_bcast) {
    std::vector<double> u(d);
    BB_CUDA(cudaMemcpyAsync(u.data(), e->start, sizeof(double) * d, cudaMemcpyDeviceToHost, e->ctx->stream));
    BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
    for (int64_t p = 0; p < np; p++) memcpy(x0 + (size_t)p * d, u.data(), sizeof(double) * d);
    return BB_OK;
  }
  const double* src = which == BB_CUR ? e->start : e->th->startprop;
  if (!src) return BB_ERR_ARG;
  std::vector<double> h((size_t)np * d);
  for (int k = 0; k < d; k++)
    BB_CUDA(cudaMemcpyAsync(h.data() + (size_t)k * np, src + (size_t)k * e->P + p0, sizeof(double) * np,
                            cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  for (int64_t p = 0; p < np; p++)
    for (int k = 0; k < d; k++) x0[(size_t)p * d + k] = h[(size_t)k * np + p];
  return BB_OK;
}

extern "C" int bb_theta_guides(bb_ens* e) {
  if (!e || !e->th) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  bb_time_begin(e->ctx);
  const int rc = run_backward(e, 0, false, nullptr, 0, 0);
  bb_time_end(e->ctx);
  return rc;
}

extern "C" int bb_theta_get_left(bb_ens* e, int which, int64_t p0, int64_t np, double* out) {
  if (!e || !e->th || !out || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (which != BB_CUR && which != BB_PROP) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int NL = e->th->NL;
  std::vector<double> h((size_t)np * NL);
  for (int k = 0; k < NL; k++)
    BB_CUDA(cudaMemcpyAsync(h.data() + (size_t)k * np, e->th->left[which] + (size_t)k * e->P + p0, sizeof(double) * np,
                            cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  for (int64_t p = 0; p < np; p++)
    for (int k = 0; k < NL; k++) out[(size_t)p * NL + k] = h[(size_t)k * np + p];
  return BB_OK;
}

extern "C" int bb_theta_get_tables(bb_ens* e, int64_t p, double* nu, double* H) {
  if (!e || !e->th || !nu || !H || p < 0 || p >= e->P) return BB_ERR_ARG;
  if (e->th->tstate == 0) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int K = e->th->K, d = e->d;
  const size_t rows = (size_t)e->S * e->N;
  std::vector<double> h(rows * K);
  BB_CUDA(cudaMemcpy2DAsync(h.data(), sizeof(double), e->th->T + p, sizeof(double) * e->th->PT, sizeof(double), rows * K,
                            cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  for (size_t g = 0; g < rows; g++) {
    memcpy(nu + g * d, h.data() + g * K, sizeof(double) * d);
    memcpy(H + g * d * d, h.data() + g * K + d, sizeof(double) * d * d);
  }
  return BB_OK;
}

/* the tables must belong to the chains' CURRENT θ */
static int ensure_current_tables(bb_ens* e) {
  if (e->th->tstate == 1) return BB_OK;
  return run_backward(e, 0, false, nullptr, 0, 0);
}
/* whole-path steps compare against the running ll of the current (W, θ) under the full backward chain; after block
 * updates it is re-established by one forward pass (which also puts X back on that chain's guide) */
static int ensure_running_ll(bb_ens* e, int skip) {
  if (!e->th->ll_stale) return BB_OK;
  int rc = ensure_current_tables(e);
  if (rc == BB_OK) rc = run_forward(e, false, 0, 0, skip, e->X != nullptr, 0.0, 0, 0);
  if (rc == BB_OK) e->th->ll_stale = false;
  return rc;
}

extern "C" int bb_theta_guided_euler_ll(bb_ens* e, int32_t skip, uint32_t flags) {
  if (!e || !e->th) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  bb_time_begin(e->ctx);
  int rc = ensure_current_tables(e);
  if (rc == BB_OK) rc = run_forward(e, false, 0, 0, skip, (flags & BB_RUN_STORE_X) != 0, 0.0, 0, 0);
  if (rc == BB_OK) e->th->ll_stale = false;
  bb_time_end(e->ctx);
  return rc;
}

extern "C" int bb_theta_pcn_step(bb_ens* e, double rho, uint64_t seed, uint32_t iter, int32_t skip, uint32_t flags) {
  if (!e || !e->th) return BB_ERR_ARG;
  if (!(rho >= -1.0 && rho <= 1.0)) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  bb_time_begin(e->ctx);
  int rc = ensure_running_ll(e, skip);
  if (rc == BB_OK) rc = ensure_current_tables(e);
  if (rc == BB_OK) rc = run_forward(e, true, 0, 0, skip, (flags & BB_RUN_STORE_X) != 0, rho, seed, iter);
  bb_time_end(e->ctx);
  return rc;
}

extern "C" int bb_theta_param_step(bb_ens* e, const double* rw_sd, uint64_t seed, uint32_t iter, int32_t skip,
                                   uint32_t flags) {
  if (!e || !e->th || !rw_sd) return BB_ERR_ARG;
  int nz = 0;
  for (int k = 0; k < BB_NTHETA; k++) {
    if (!(rw_sd[k] >= 0.0)) return BB_ERR_ARG;
    if (rw_sd[k] != 0.0) nz++;
  }
  if (nz > (e->th->spec.start_sd != 0.0 ? 3 : 4)) return BB_ERR_ARG; /* normal 3 of the quad drives the start move */
  BB_CUDA(cudaSetDevice(e->ctx->device));
  if (e->th->spec.start_sd != 0.0 && e->start_bcast) return BB_ERR_STARTPOINT; /* per-chain starting points: bb_ens_set_start(..., broadcast = 0) */
  bb_time_begin(e->ctx);
  int rc = ensure_running_ll(e, skip);
  /* the left-end values of the current θ (logpdfnormal, trace term, prior) enter the accept test */
  if (rc == BB_OK && e->th->tstate == 0) rc = run_backward(e, 0, false, nullptr, 0, 0);
  if (rc == BB_OK) rc = run_backward(e, 1, true, rw_sd, seed, iter);
  if (rc == BB_OK) rc = run_forward(e, false, 1, 2, skip, (flags & BB_RUN_STORE_X) != 0, 0.0, seed, iter);
  bb_time_end(e->ctx);
  return rc;
}

extern "C" int bb_theta_refresh_x(bb_ens* e) {
  if (!e || !e->th) return BB_ERR_ARG;
  if (!e->x_maybe_stale) return BB_OK;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  int rc = ensure_current_tables(e);
  if (rc == BB_OK) rc = run_forward(e, false, 0, 3, 0, true, 0.0, 0, 0);
  return rc;
}

extern "C" int bb_theta_block_step(bb_ens* e, int32_t s_lo, int32_t s_hi, double rho, double hzero, uint64_t seed,
                                   uint32_t iter, int32_t skip) {
  if (!e || !e->th) return BB_ERR_ARG;
  if (s_lo < 0 || s_hi <= s_lo || s_hi > e->S || skip < 0) return BB_ERR_ARG;
  if (!(rho >= -1.0 && rho <= 1.0) || !(hzero > 0.0)) return BB_ERR_ARG;
  if (!e->X || !(e->flags & BB_ENS_DOUBLE_BUFFER)) return BB_ERR_ARG; /* the block is conditioned on the stored path */
  bb_theta* t = e->th;
  if (s_lo == 0 && t->spec.start_sd != 0.0 && e->start_bcast) return BB_ERR_STARTPOINT;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  int rc = BB_OK;
  if (!t->blkv) rc = th_alloc(e, &t->blkv, (size_t)BB_NBLK * (size_t)e->P);
  if (rc == BB_OK && !t->Xprop) rc = th_alloc(e, &t->Xprop, (size_t)e->S * e->NC * (size_t)e->P * BB_TC * e->d);
  if (rc != BB_OK) return rc;
  bb_time_begin(e->ctx);
  if (e->x_maybe_stale) rc = bb_theta_refresh_x(e); /* rejected proposals of earlier whole-path steps */
  const blk_spec blk{s_lo, s_hi, hzero};
  if (rc == BB_OK) rc = run_backward(e, 0, false, nullptr, seed, iter, &blk);
  if (rc == BB_OK) rc = run_block(e, blk, skip, rho, seed, iter);
  if (rc == BB_OK) t->ll_stale = true;
  bb_time_end(e->ctx);
  return rc;
}

extern "C" int bb_theta_get_block(bb_ens* e, int64_t p0, int64_t np, double* out) {
  if (!e || !e->th || !e->th->blkv || !out || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int NB = BLK_SEG + 2 * e->S;
  std::vector<double> h((size_t)np * NB);
  for (int k = 0; k < NB; k++)
    BB_CUDA(cudaMemcpyAsync(h.data() + (size_t)k * np, e->th->blkv + (size_t)k * e->P + p0, sizeof(double) * np,
                            cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  for (int64_t p = 0; p < np; p++)
    for (int k = 0; k < NB; k++) out[(size_t)p * NB + k] = h[(size_t)k * np + p];
  return BB_OK;
}

extern "C" int bb_theta_get_acc(bb_ens* e, int64_t* acc) {
  if (!e || !e->th || !acc) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  unsigned long long v = 0;
  BB_CUDA(cudaMemcpyAsync(&v, e->th->acc, sizeof(v), cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  *acc = (int64_t)v;
  return BB_OK;
}
extern "C" void* bb_theta_acc_device_ptr(bb_ens* e) { return (e && e->th) ? (void*)e->th->acc : nullptr; }
