/*
 * bb_second.cuh -- the two passes over a STORED path that the reference also has:
 *   MODE 0: llikelihood(LeftRule(), X, P°; skip)   src/partialbridgenuH.jl:171-189, guip.jl:429-446,
 *                                                  partialbridge.jl:67-87, guip!.jl:5-31
 *   MODE 1: innovations!(EulerMaruyama(), W, X, P) src/euler.jl:357-376   (W <- the increments that drive X)
 * (the fused kernel of bb_chain.cuh computes the log-likelihood while it builds the path; these exist
 * because callers may hold a path and ask for either quantity).  One thread per chain, X read once with
 * 256-bit loads, tables read by warp-uniform loads (L1 broadcast).
 */
#pragma once
#include "bb_chain.cuh"

template <class M, int GK, int GM, int AUXM, int MODE>
__device__ __forceinline__ void bb_second_body(const bb_chain_args& a) {
  using CH = bb_chain<M, GK, GM, AUXM, 0>;
  constexpr int D = M::D, DP = M::DP, REC = CH::REC;
  const long long P = a.P;
  const long long p = a.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.p_end) return;
  const int par = a.par[p];
  const double* xr = a.X + p * (BB_TC * D);
  double* ww = a.W[par] + p * (a.nbuf * BB_TC * DP);
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
  double lltot = 0.0;
  for (int s = 0; s < a.S; s++) {
    const double* tab = a.tab[s];
    const double* sc = a.segc[s];
    double xprev[D], w[DP];
    double som = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) w[k] = 0.0;
    for (int c = 0; c < a.NC; c++) {
      bb_rowout<DP> wo;
#pragma unroll 1
      for (int h = 0; h < BB_TC / 4; h++) {
        double x[4 * D];
#pragma unroll
        for (int q = 0; q < D; q++) bb_ld4(xr + 4 * h * D + 4 * q, x + 4 * q);
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
          const int j = c * BB_TC + 4 * h + s4;
          if (j > 0 && j < a.N) {
            const double* R = tab + (size_t)j * REC;
            const double dt = R[0];
            double bd[D];
            CH::drift(a.model, R, sc, xprev, dt, MODE == 0 && j <= a.jll, som, bd);
            if constexpr (MODE == 1) {
              double e[D], de[D];
#pragma unroll
              for (int k = 0; k < D; k++) e[k] = (x[s4 * D + k] - xprev[k]) - bd[k] * dt;
              bb_matvec<D, D>(a.model.der + 24, e, de);
#pragma unroll
              for (int k = 0; k < DP; k++) w[k] = w[k] + de[k];
            }
          }
#pragma unroll
          for (int k = 0; k < D; k++) xprev[k] = x[s4 * D + k];
          if constexpr (MODE == 1) wo.put(ww + 4 * h * DP, s4, w, true);
        }
      }
      xr += xstride;
      ww += wstride;
    }
    lltot += som;
  }
  if (MODE == 0) a.ll[p] = lltot;
}

/* solve!(StochasticHeun(), Y, u, W, P)  src/euler.jl:178-198 for a plain target, one segment:
 *   y2 = y + b(y) dt;  y = y + 0.5 (b(y2) + b(y)) dt + sigma dw   for the steps 0 .. N-3 ("for i in 1:N-2 # fix me");
 *   yy[N-1] = endpoint(y);  yy[N] keeps its old value, as in the reference. */
template <class M, int GK, int GM, int AUXM, int MODE>
__global__ void __launch_bounds__(BB_THREADS) bb_second_kernel(const __grid_constant__ bb_chain_args a) {
  bb_second_body<M, GK, GM, AUXM, MODE>(a);
}

template <class M>
__device__ __forceinline__ void bb_heun_body(const bb_chain_args& a) {
  constexpr int D = M::D, DP = M::DP;
  const long long P = a.P;
  const long long p = a.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.p_end) return;
  const int par = a.par[p], N = a.N;
  const double* wr = a.W[par] + p * (a.nbuf * BB_TC * DP);
  double* xw = a.X + p * (BB_TC * D);
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
  const double* tab = a.tab[0];
  double y[D], wprev[DP], xlast[D];
#pragma unroll
  for (int k = 0; k < D; k++) y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + p];
  { /* the last grid point is not written by this scheme: carry its old value through the row store */
    const int jl = N - 1;
    const double* q = a.X + ((long long)(jl / BB_TC) * P + p) * (BB_TC * D) + (jl % BB_TC) * D;
#pragma unroll
    for (int k = 0; k < D; k++) xlast[k] = q[k];
  }
  for (int c = 0; c < a.NC; c++) {
    bb_rowout<D> xo;
#pragma unroll 1
    for (int h = 0; h < BB_TC / 4; h++) {
      double w[4 * DP];
#pragma unroll
      for (int q = 0; q < DP; q++) bb_ld4(wr + 4 * h * DP + 4 * q, w + 4 * q);
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++) {
        const int j = c * BB_TC + 4 * h + s4;
        if (j >= 1 && j <= N - 2) {
          const double dt = tab[2 * j];
          double b1[D], b2[D], y2[D], hb[D], dw[DP];
          M::b(a.model, y, b1);
#pragma unroll
          for (int k = 0; k < D; k++) y2[k] = fma(b1[k], dt, y[k]);
          M::b(a.model, y2, b2);
#pragma unroll
          for (int k = 0; k < D; k++) hb[k] = 0.5 * (b2[k] + b1[k]);
#pragma unroll
          for (int k = 0; k < DP; k++) dw[k] = w[s4 * DP + k] - wprev[k];
          bb_em_update<M>(a.model, hb, dt, dw, y);
        }
#pragma unroll
        for (int k = 0; k < DP; k++) wprev[k] = w[s4 * DP + k];
        double o[D];
#pragma unroll
        for (int k = 0; k < D; k++) o[k] = (j == N - 1) ? xlast[k] : y[k];
        xo.put(xw + 4 * h * D, s4, o, true);
      }
    }
    wr += wstride;
    xw += xstride;
  }
  a.xstale[p] = 0;
  if (a.write_end) {
#pragma unroll
    for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
  }
}
template <class M>
__global__ void __launch_bounds__(BB_THREADS) bb_heun_kernel(const __grid_constant__ bb_chain_args a) {
  bb_heun_body<M>(a);
}

/* solve!(Mdb(), Y, u, W, P°)  src/euler.jl:308-327, the "modified diffusion bridge" step on a guided proposal, one segment:
 *   y = y + _b((i, tt[i]), y, P°) dt + (σ sqrt((tt[end] - tt[i+1]) / (tt[end] - tt[i]))) dw ;  yy[N] = endpoint(y, P°)
 * i.e. the guided Euler step with the noise damped towards the end point.  The per-step factor comes from a second
 * table, a.tab[1][j] = sqrt((T - tt[j]) / (T - tt[j-1])) for the step j-1 -> j (0 at the last step). */
template <class M, int GK, int GM, int AUXM>
__device__ __forceinline__ void bb_mdb_body(const bb_chain_args& a) {
  using CH = bb_chain<M, GK, GM, AUXM, 0>;
  constexpr int D = M::D, DP = M::DP, REC = CH::REC;
  const long long P = a.P;
  const long long p = a.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.p_end) return;
  const int par = a.par[p], N = a.N;
  const double* wr = a.W[par] + p * (a.nbuf * BB_TC * DP);
  double* xw = a.X + p * (BB_TC * D);
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
  const double* tab = a.tab[0];
  const double* scale = a.tab[1];
  const double* sc = a.segc[0];
  double y[D], wprev[DP], som = 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + p];
  for (int c = 0; c < a.NC; c++) {
    bb_rowout<D> xo;
#pragma unroll 1
    for (int h = 0; h < BB_TC / 4; h++) {
      double w[4 * DP];
#pragma unroll
      for (int q = 0; q < DP; q++) bb_ld4(wr + 4 * h * DP + 4 * q, w + 4 * q);
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++) {
        const int j = c * BB_TC + 4 * h + s4;
        if (j >= 1 && j <= N - 1) {
          const double* R = tab + (size_t)j * REC;
          const double dt = R[0], ns = scale[j];
          double bd[D], dw[DP];
#pragma unroll
          for (int k = 0; k < DP; k++) dw[k] = w[s4 * DP + k] - wprev[k];
          CH::drift(a.model, R, sc, y, dt, false, som, bd);
          if constexpr (M::SPARSE) {
#pragma unroll
            for (int i = 0; i < D; i++) {
              double t1 = fma(bd[i], dt, y[i]);
              if (M::col(i) >= 0) t1 = fma(M::sig(a.model, i) * ns, dw[M::col(i) < 0 ? 0 : M::col(i)], t1);
              y[i] = t1;
            }
          } else {
            double Ss[D * DP], sd[D];
#pragma unroll
            for (int q = 0; q < D * DP; q++) Ss[q] = MLinPro<D>::sigma(a.model)[q] * ns;
            bb_matvec<D, DP>(Ss, dw, sd);
#pragma unroll
            for (int i = 0; i < D; i++) y[i] = fma(bd[i], dt, y[i]) + sd[i];
          }
          if (GK == BB_GUIDE_HV && j == N - 1 && sc[D * D + D] != 0.0) { /* endpoint(y, P::GuidedBridge)  src/euler.jl:241-242 */
#pragma unroll
            for (int k = 0; k < D; k++) y[k] = sc[D * D + D + 1 + k];
          }
        }
#pragma unroll
        for (int k = 0; k < DP; k++) wprev[k] = w[s4 * DP + k];
        xo.put(xw + 4 * h * D, s4, y, true);
      }
    }
    wr += wstride;
    xw += xstride;
  }
  a.xstale[p] = 0;
  if (a.write_end) {
#pragma unroll
    for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
  }
}
template <class M, int GK, int GM, int AUXM>
__global__ void __launch_bounds__(BB_THREADS) bb_mdb_kernel(const __grid_constant__ bb_chain_args a) {
  bb_mdb_body<M, GK, GM, AUXM>(a);
}
#ifndef __CUDACC_RTC__ /* host-side launch + lookup (not part of a run-time compiled user-model kernel) */
template <class M>
static cudaError_t bb_heun_launch(const bb_chain_args& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + BB_THREADS - 1) / BB_THREADS);
  bb_heun_kernel<M><<<grid, BB_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

template <class M, int GK, int GM, int AUXM>
static cudaError_t bb_mdb_launch(const bb_chain_args& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + BB_THREADS - 1) / BB_THREADS);
  bb_mdb_kernel<M, GK, GM, AUXM><<<grid, BB_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

template <class M, int GK, int GM, int AUXM, int MODE>
static cudaError_t bb_second_launch(const bb_chain_args& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + BB_THREADS - 1) / BB_THREADS);
  bb_second_kernel<M, GK, GM, AUXM, MODE><<<grid, BB_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

template <class M, int GK, int GM>
static bb_chain_launch_fn bb_lookup_second_guide(int auxm, int mode) {
  if (mode == 0) return auxm == 1 ? &bb_second_launch<M, GK, GM, 1, 0>
                                  : (auxm == 0 ? &bb_second_launch<M, GK, GM, 0, 0> : &bb_second_launch<M, GK, GM, 2, 0>);
  if constexpr (M::D == M::DP) {
    if (mode == 1) return auxm == 1 ? &bb_second_launch<M, GK, GM, 1, 1>
                                    : (auxm == 0 ? &bb_second_launch<M, GK, GM, 0, 1> : &bb_second_launch<M, GK, GM, 2, 1>);
  }
  if (mode == 3) return auxm == 1 ? &bb_mdb_launch<M, GK, GM, 1> : (auxm == 0 ? &bb_mdb_launch<M, GK, GM, 0> : &bb_mdb_launch<M, GK, GM, 2>);
  return nullptr;
}
/* mode 0 = llikelihood (needs a guide), 1 = innovations (guide optional, needs d' = d), 2 = StochasticHeun (no guide),
 * 3 = Mdb on a guided proposal */
template <class M>
static bb_chain_launch_fn bb_lookup_second(int gk, int gm, int auxc, int mode) {
  if (gk == 0) {
    if (mode == 2) return &bb_heun_launch<M>;
    if constexpr (M::D == M::DP) {
      if (mode == 1) return &bb_second_launch<M, 0, 0, true, 1>;
    }
    return nullptr;
  }
  if (gk == BB_GUIDE_NUH) return bb_lookup_second_guide<M, BB_GUIDE_NUH, 0>(auxc, mode);
  if (gk == BB_GUIDE_HV) return bb_lookup_second_guide<M, BB_GUIDE_HV, 0>(auxc, mode);
  if (gk == BB_GUIDE_LMMU) {
    if (gm == 1) return bb_lookup_second_guide<M, BB_GUIDE_LMMU, 1>(auxc, mode);
    if constexpr (M::D >= 2)
      if (gm == 2) return bb_lookup_second_guide<M, BB_GUIDE_LMMU, 2>(auxc, mode);
    if constexpr (M::D >= 3)
      if (gm == 3) return bb_lookup_second_guide<M, BB_GUIDE_LMMU, 3>(auxc, mode);
  }
  return nullptr;
}
#endif /* !__CUDACC_RTC__ */
