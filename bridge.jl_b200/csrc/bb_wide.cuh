/*
 * bb_wide.cuh -- the path kernel for WIDE models (state dimension up to 16: BB_MODEL_LANDMARKS, and the
 * d'-dimensional Wiener process that drives it): BASELINE config 5.
 *
 * Same operations and the same data layout as bb_chain.cuh (sample!, W° = rho W + sqrt(1-rho^2) W2, plain or guided
 * solve!, llikelihood, accept/reject; W [S][NC][P][nbuf][16][d'], X [S][NC][P][16][d]), but a plain design: one thread
 * per chain, 64 chains per CTA.  With d = 16 a chain already moves whole 128-byte lines per step (X) and the
 * ensembles of this configuration are small (1e4 chains = 2 warps per SM), so the kernel is bound by the latency of
 * its ~1000 fp64 operations per step, not by bandwidth; the chunk pipeline of bb_chain.cuh (256 chains x 16 steps x
 * 8 d' bytes of staging per CTA) does not fit shared memory at d' = 8.
 * The per-step table row (dt, sqrt dt, nu, H: 2 + d + d*d doubles, the same for every chain) is streamed through a
 * BB_WIDE_RING-deep shared-memory ring: the CTA copies the row BB_WIDE_RING-1 steps ahead with 16-byte cp.async and
 * reads the current one by broadcast LDS (read straight from global memory every value was an L2 round trip: with
 * one or two warps per SM nothing is reused in L1 -- 10 ms per 1000 steps instead of ~2).  The constant auxiliary
 * drift is copied to shared memory once.
 * The per-step arithmetic is bb_chain<...>::drift / bb_em_update: the oracle's operation order.
 * A constant auxiliary drift (B~ [d*d], beta~ [d]) does not fit the per-segment constants of bb_chain_args; it is
 * appended to the segment's table (after the NC*16 rows).
 */
#pragma once
#include "bb_chain.cuh"

#define BB_WIDE_THREADS 64
#define BB_WIDE_RING 4

/* RNG: 0 read W, 1 pCN + X°, 2 fresh Wiener path, 3 pCN without X° (as bb_chain_kernel); GK = 0 or BB_GUIDE_NUH */
template <class M, int GK, int RNG>
__global__ void __launch_bounds__(BB_WIDE_THREADS) bb_wide_kernel(const __grid_constant__ bb_chain_args a) {
  using CH = bb_chain<M, GK, 0, 1, 0>;
  constexpr int D = M::D, DP = M::DP, REC = CH::REC;
  constexpr bool PCN = (RNG == 1 || RNG == 3);
  constexpr int NPIECE = BB_TC * DP / 4;
  const long long P = a.P;
  const long long p = a.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pc = p < a.p_end ? p : a.p_end - 1;
  const bool act = p < a.p_end && (!a.only || a.only[pc] != 0);
  const int lane = threadIdx.x & 31;
  const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
  const int S = a.S, N = a.N, NC = a.NC;
  const bool sx = RNG == 1 ? true : (RNG == 3 ? false : a.store_x != 0);

  const int par = a.par[pc];
  const int wbuf = PCN ? 1 - par : par;
  const double* wr = a.W[par] + pc * (a.nbuf * BB_TC * DP);
  double* ww = a.W[wbuf] + pc * (a.nbuf * BB_TC * DP);
  double* xw = sx ? a.X + pc * (BB_TC * D) : nullptr;
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);

  double y[D], wprev[DP], w2[DP];
#pragma unroll
  for (int k = 0; k < D; k++) y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc];
  double lltot = 0.0;
  double wq[4] = {0.0, 0.0, 0.0, 0.0};

  /* shared memory: the ring of table rows, then B~, beta~ */
  constexpr int NAUX = (GK != 0) ? D * D + D : 0;
  constexpr int PIECES = REC / 2; /* 16-byte pieces per row */
  __shared__ __align__(16) double ring[BB_WIDE_RING * REC + NAUX + 2];
  double* sc_s = ring + BB_WIDE_RING * REC;
  const int rows_per_seg = NC * BB_TC;
  const int total_rows = S * rows_per_seg;
  int issue_row = 0; /* next row (over all segments) to request */
  auto row_issue = [&]() { /* the CTA requests one row; every thread commits a (possibly empty) group */
    if (issue_row < total_rows) {
      const double* src = a.tab[issue_row / rows_per_seg] + (size_t)(issue_row % rows_per_seg) * REC;
      double* dst = ring + (issue_row % BB_WIDE_RING) * REC;
      for (int q = threadIdx.x; q < PIECES; q += BB_WIDE_THREADS) bb_cp_async16(dst + 2 * q, src + 2 * q);
    }
    issue_row++;
    bb_cp_async_commit();
  };
#pragma unroll 1
  for (int r = 0; r < BB_WIDE_RING - 1; r++) row_issue();
  int cons_row = 0;

  for (int s = 0; s < S; s++) {
    const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
    const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
    const double* tab = a.tab[s];
    if constexpr (GK != 0) { /* B~, beta~ of this segment (they follow the rows of its table) */
      __syncthreads();
      for (int q = threadIdx.x; q < NAUX; q += BB_WIDE_THREADS) sc_s[q] = tab[(size_t)NC * BB_TC * REC + q];
      __syncthreads();
    }
    const double* sc = sc_s;
    double som = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) w2[k] = 0.0;
    for (int c = 0; c < NC; c++) {
#pragma unroll 1
      for (int slot = 0; slot < BB_TC; slot++) {
        const int j = c * BB_TC + slot;
        /* this step's row has landed for every thread of the CTA; the stage read one step ago is free again */
        bb_cp_async_wait<BB_WIDE_RING - 2>();
        __syncthreads();
        const double* R = ring + (cons_row % BB_WIDE_RING) * REC;
        cons_row++;
        row_issue();
        double wj[DP];
#pragma unroll
        for (int k = 0; k < DP; k++) {
          const int mm = slot * DP + k; /* element of the chunk row; DP is a multiple of 4 or divides 4 */
          if ((mm & 3) == 0) {
            const int q = mm >> 2;
            if (RNG != 2 || j == 0) {
              if (act) bb_ld4(wr + 4 * q, wq);
            }
            if constexpr (RNG != 0) {
              float z[4];
              bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + q), z);
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int sl = (mm + i) / DP, kk = (mm + i) % DP;
                const int jj = c * BB_TC + sl;
                /* a piece never leaves its grid point when d' is a multiple of 4: sqrt(dt) of the staged row */
                const double rootdt = (DP % 4 == 0) ? R[1] : tab[(size_t)jj * REC + 1];
                if constexpr (PCN) {
                  /* W2[j] = W2[j-1] + sqrt(dt) xi ;  W°[j] = rho W[j] + sqrt(1-rho^2) W2[j] */
                  if (jj != 0) w2[kk] = fma(rootdt, (double)z[i], w2[kk]);
                  wq[i] = fma(a.rho2, w2[kk], a.rho * wq[i]);
                } else {
                  /* W[j] = W[j-1] + sqrt(dt) xi, W[0] kept   (src/wiener.jl:24-35) */
                  w2[kk] = (jj == 0) ? wq[i] : fma(rootdt, (double)z[i], w2[kk]);
                  wq[i] = w2[kk];
                }
              }
              if (act) bb_st4(ww + 4 * q, wq[0], wq[1], wq[2], wq[3]);
            }
          }
          wj[k] = wq[mm & 3];
        }
        if (j == 0) {
#pragma unroll
          for (int k = 0; k < DP; k++) wprev[k] = wj[k];
        } else if (j < N) {
          const double dt = R[0];
          double dw[DP], bd[D];
#pragma unroll
          for (int k = 0; k < DP; k++) {
            dw[k] = wj[k] - wprev[k];
            wprev[k] = wj[k];
          }
          CH::drift(a.model, R, sc, y, dt, j <= a.jll, som, bd);
          bb_em_update<M>(a.model, bd, dt, dw, y);
        }
        if (sx && act) {
#pragma unroll
          for (int q = 0; q < D / 4; q++) bb_st4(xw + slot * D + 4 * q, y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
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
  } else if (act) {
    if (a.do_ll) a.ll[p] = lltot;
    a.xstale[p] = sx ? 0 : 1;
    if (a.write_end) {
#pragma unroll
      for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = y[k];
    }
  }
}

template <class M, int GK, int RNG>
static cudaError_t bb_wide_launch(const bb_chain_args& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + BB_WIDE_THREADS - 1) / BB_WIDE_THREADS);
  bb_wide_kernel<M, GK, RNG><<<grid, BB_WIDE_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

template <class M>
static bb_chain_launch_fn bb_lookup_wide(int gk, int auxm, int rng) {
  if (gk == 0) {
    if (rng == 0) return &bb_wide_launch<M, 0, 0>;
    if (rng == 2) return &bb_wide_launch<M, 0, 2>;
    return nullptr;
  }
  if (gk == BB_GUIDE_NUH && auxm == 1) {
    if (rng == 0) return &bb_wide_launch<M, BB_GUIDE_NUH, 0>;
    if (rng == 1) return &bb_wide_launch<M, BB_GUIDE_NUH, 1>;
    if (rng == 3) return &bb_wide_launch<M, BB_GUIDE_NUH, 3>;
  }
  return nullptr;
}
