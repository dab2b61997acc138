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

/* ------------------------------------------------------------------------------------------------------------
 * Landmarks, FOUR lanes per chain.  At the sizes of config 5 (1e4 chains) one thread per chain leaves half of the
 * warp schedulers empty and every warp alone with a chain of ~1300 instructions per step.  Here lane g of a group
 * of four owns landmark g, i.e. components 4g .. 4g+3 of the state: its rows of the two 16x16 mat-vecs, its part of
 * the pairwise drift, its two noise columns, its Euler update; the new state is all-gathered by 16 shuffles and
 * the log-likelihood dot product runs through the four lanes in order.  Every component sees exactly the rounding
 * sequence of bb_wide_kernel (and of the oracle), so the two kernels agree bit for bit; the work per lane drops to
 * about a quarter and there are four times as many warps.  GK = 0 or BB_GUIDE_NUH; RNG = 0, 1, 3 as above.
 * ---------------------------------------------------------------------------------------------------------- */
#define BB_W4_THREADS 128

__device__ __forceinline__ void bb_ld2(const double* p, double* v) {
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
__device__ __forceinline__ void bb_st2(double* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

/* one row of a 16-column mat-vec from shared memory with 128-bit loads (the row is 16-byte aligned); the summation
 * order is bb_matvec's */
__device__ __forceinline__ double bb_rowdot16(const double* row, const double* x) {
  const double2* r2 = reinterpret_cast<const double2*>(row);
  double2 v = r2[0];
  double s = v.x * x[0];
  s = fma(v.y, x[1], s);
#pragma unroll
  for (int l = 1; l < 8; l++) {
    v = r2[l];
    s = fma(v.x, x[2 * l], s);
    s = fma(v.y, x[2 * l + 1], s);
  }
  return s;
}

template <int GK, int RNG>
__global__ void __launch_bounds__(BB_W4_THREADS) bb_wide4_kernel(const __grid_constant__ bb_chain_args a) {
  using M = MLandmarks;
  using CH = bb_chain<M, GK, 0, 1, 0>;
  constexpr int D = M::D, DP = M::DP, REC = CH::REC, NL = M::NL;
  constexpr bool PCN = (RNG == 1 || RNG == 3);
  constexpr int NPIECE = BB_TC * DP / 4;
  constexpr int OFF_C = 2, OFF_A2 = 2 + D;
  const long long P = a.P;
  const int g = threadIdx.x & 3; /* the landmark of this lane */
  const long long p = a.p_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2);
  const long long pc = p < a.p_end ? p : a.p_end - 1;
  const bool act = p < a.p_end && (!a.only || a.only[pc] != 0);
  const int lane = threadIdx.x & 31;
  const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
  const int S = a.S, N = a.N, NC = a.NC;
  const bool sx = RNG == 1 ? true : (RNG == 3 ? false : a.store_x != 0);
  const bb_model_dev& m = a.model;
  const double c0 = m.der[0], c1 = m.der[1], akk = m.der[2], nlh = m.der[3], sig = m.par[1];

  const int par = a.par[pc];
  const int wbuf = PCN ? 1 - par : par;
  const double* wr = a.W[par] + pc * (a.nbuf * BB_TC * DP) + 2 * g;
  double* ww = a.W[wbuf] + pc * (a.nbuf * BB_TC * DP) + 2 * g;
  double* xw = sx ? a.X + pc * (BB_TC * D) + 4 * g : nullptr;
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);

  double y[D], yo[4], wprev[2], w2[2];
#pragma unroll
  for (int k = 0; k < D; k++) y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc];
  /* own components 4g .. 4g+3 */
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double v = y[k];
#pragma unroll
    for (int i = 1; i < NL; i++) v = (g == i) ? y[4 * i + k] : v;
    yo[k] = v;
  }
  double lltot = 0.0;
  double wnx[2] = {0.0, 0.0};
  if (act) bb_ld2(wr, wnx);

  /* every WARP runs its own ring of table rows (its 8 chains are at the same step; warps of a CTA are not), so a
   * step needs a __syncwarp only and a warp that waits for its W never holds the others back */
  constexpr int NAUX = (GK != 0) ? D * D + D : 0;
  constexpr int PIECES = REC / 2;
  constexpr int NWARP = BB_W4_THREADS / 32;
  /* the four lanes of a group read rows 4g+k of H (and of B~) at the same time: 128 bytes apart they would sit in
   * the same banks (4-way conflicts, measured: shared-memory bound at 1e4 chains), so the rows of landmark g are
   * shifted by 2g doubles: the four 16-byte pieces a group reads are contiguous */
  constexpr int RECS = REC + 8, NAUXS = NAUX + 8;
  __shared__ __align__(16) double ring_all[NWARP * BB_WIDE_RING * RECS + NAUXS + 2];
  double* ring = ring_all + (threadIdx.x >> 5) * (BB_WIDE_RING * RECS);
  double* sc_s = ring_all + NWARP * BB_WIDE_RING * RECS;
  const int rows_per_seg = NC * BB_TC;
  const int total_rows = S * rows_per_seg;
  int issue_row = 0;
  auto row_issue = [&]() {
    if (issue_row < total_rows) {
      const double* src = a.tab[issue_row / rows_per_seg] + (size_t)(issue_row % rows_per_seg) * REC;
      double* dst = ring + (issue_row % BB_WIDE_RING) * RECS;
      for (int q = lane; q < PIECES; q += 32) {
        const int e = 2 * q - OFF_A2; /* element of H (row-major d x d) if >= 0 */
        bb_cp_async16(dst + 2 * q + (e >= 0 ? 2 * (e >> 6) : 0), src + 2 * q);
      }
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
    if constexpr (GK != 0) {
      __syncthreads();
      for (int q = threadIdx.x; q < NAUX; q += BB_W4_THREADS)
        sc_s[q + 2 * (q >> 6)] = tab[(size_t)NC * BB_TC * REC + q]; /* rows of B~ shifted like those of H; beta~ after */
      __syncthreads();
    }
    double som = 0.0;
    w2[0] = w2[1] = 0.0;
    for (int c = 0; c < NC; c++) {
#pragma unroll 1
      for (int slot = 0; slot < BB_TC; slot++) {
        const int j = c * BB_TC + slot;
        bb_cp_async_wait<BB_WIDE_RING - 2>();
        __syncwarp();
        const double* R = ring + (cons_row % BB_WIDE_RING) * RECS;
        cons_row++;
        row_issue();
        /* ---- own two noise columns 2g, 2g+1: elements 2(g&1), 2(g&1)+1 of piece 2 slot + (g >> 1); they were
         * requested one step ago, the next step's are requested now */
        double wj[2] = {wnx[0], wnx[1]};
        {
          const bool lastrow = (s == S - 1) && (c == NC - 1) && (slot == BB_TC - 1);
          const double* nxt = (slot == BB_TC - 1) ? wr + wstride : wr + (slot + 1) * DP;
          if (act && !lastrow) bb_ld2(nxt, wnx);
        }
        if constexpr (PCN) {
          float z[4];
          bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + 2 * slot + (g >> 1)), z);
          const float z0 = (g & 1) ? z[2] : z[0], z1 = (g & 1) ? z[3] : z[1];
          const double rootdt = R[1];
          if (j != 0) {
            w2[0] = fma(rootdt, (double)z0, w2[0]);
            w2[1] = fma(rootdt, (double)z1, w2[1]);
          }
          wj[0] = fma(a.rho2, w2[0], a.rho * wj[0]);
          wj[1] = fma(a.rho2, w2[1], a.rho * wj[1]);
          if (act) bb_st2(ww + slot * DP, wj[0], wj[1]);
        }
        if (j == 0) {
          wprev[0] = wj[0]; wprev[1] = wj[1];
        } else if (j < N) {
          const double dt = R[0];
          const double dw0 = wj[0] - wprev[0], dw1 = wj[1] - wprev[1];
          wprev[0] = wj[0]; wprev[1] = wj[1];
          /* ---- b(t, x) for landmark g  (MLandmarks::b, rows 4g .. 4g+3) */
          /* k(q_g - q_j): the lane evaluates the pairs (g, g+1) and (g, g+2) (mod 4) and takes (g-1, g) from its
           * neighbour -- k is symmetric and (-dx)^2 = dx^2, so these are the values MLandmarks::b computes */
          double kA, kB, kC;
          {
            const double ax = __shfl_sync(0xFFFFFFFFu, yo[0], (g + 1) & 3, 4), ay = __shfl_sync(0xFFFFFFFFu, yo[1], (g + 1) & 3, 4);
            const double bx = __shfl_sync(0xFFFFFFFFu, yo[0], (g + 2) & 3, 4), by = __shfl_sync(0xFFFFFFFFu, yo[1], (g + 2) & 3, 4);
            const double dxa = yo[0] - ax, dya = yo[1] - ay, dxb = yo[0] - bx, dyb = yo[1] - by;
            kA = c0 * bb_exp(-(fma(dya, dya, dxa * dxa) * c1));
            kB = c0 * bb_exp(-(fma(dyb, dyb, dxb * dxb) * c1));
            kC = __shfl_sync(0xFFFFFFFFu, kA, (g + 3) & 3, 4);
          }
          double bg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
          for (int jl = 0; jl < NL; jl++) {
            const int rel = (jl - g) & 3;
            const double kij = rel == 0 ? c0 : (rel == 1 ? kA : (rel == 2 ? kB : kC));
            const double dot = fma(yo[3], y[4 * jl + 3], yo[2] * y[4 * jl + 2]);
#pragma unroll
            for (int k = 0; k < 2; k++) {
              bg[k] += (0.5 * y[4 * jl + 2 + k]) * kij;
              const double t1 = (nlh * y[4 * jl + 2 + k]) * kij;
              const double t2 = ((c1 * dot) * (yo[k] - y[4 * jl + k])) * kij;
              bg[2 + k] += t1 + t2;
            }
          }
          if constexpr (GK != 0) {
            /* r rows 4g .. 4g+3 of H (nu - x);  b~ rows of B~ x + beta~;  <b - b~, r> through the four lanes in order */
            double e[D], rg[4];
#pragma unroll
            for (int k = 0; k < D; k += 2) {
              const double2 nu2 = *reinterpret_cast<const double2*>(R + OFF_C + k);
              e[k] = nu2.x - y[k];
              e[k + 1] = nu2.y - y[k + 1];
            }
#pragma unroll
            for (int k = 0; k < 4; k++) rg[k] = bb_rowdot16(R + OFF_A2 + (4 * g + k) * D + 2 * g, e);
            if (j <= a.jll) {
              double ee[4];
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const double bt = bb_rowdot16(sc_s + (4 * g + k) * D + 2 * g, y);
                ee[k] = bg[k] - (bt + sc_s[D * D + 8 + 4 * g + k]);
              }
              double sdot = 0.0;
#pragma unroll
              for (int st = 0; st < NL; st++) {
                double t = (st == 0) ? ee[0] * rg[0] : fma(ee[0], rg[0], sdot);
                t = fma(ee[1], rg[1], t);
                t = fma(ee[2], rg[2], t);
                t = fma(ee[3], rg[3], t);
                sdot = __shfl_sync(0xFFFFFFFFu, t, st, 4); /* the running sum after landmark st */
              }
              som = fma(sdot, dt, som);
            }
            /* _b = b + a r on the momentum rows */
            bg[2] = fma(akk, rg[2], bg[2]);
            bg[3] = fma(akk, rg[3], bg[3]);
          }
          /* ---- Euler-Maruyama update of the own components (bb_em_update) */
          yo[0] = fma(bg[0], dt, yo[0]);
          yo[1] = fma(bg[1], dt, yo[1]);
          {
            double t2 = fma(bg[2], dt, yo[2]), t3 = fma(bg[3], dt, yo[3]);
            if (sig != 0.0) {
              t2 = fma(sig, dw0, t2);
              t3 = fma(sig, dw1, t3);
            }
            yo[2] = t2; yo[3] = t3;
          }
          /* ---- all-gather the new state */
#pragma unroll
          for (int k = 0; k < D; k++) y[k] = __shfl_sync(0xFFFFFFFFu, yo[k & 3], k >> 2, 4);
        }
        if (sx && act) bb_st4(xw + slot * D, yo[0], yo[1], yo[2], yo[3]);
      }
      wr += wstride;
      ww += wstride;
      if (sx) xw += xstride;
    }
    lltot += som;
  }

  /* per-chain epilogue: lane 0 of the group writes */
  if constexpr (PCN) {
    const double logu = bb_accept_logu(a.keys, a.stream, chain);
    const bool ok = act && (logu <= lltot - a.ll[pc]);
    __syncwarp();
    if (act) {
      if (g == 0) {
        a.llprop[p] = lltot;
        a.logu[p] = logu;
        a.accepted[p] = ok ? 1 : 0;
        a.xstale[p] = sx ? (ok ? 0 : 1) : (uint8_t)(a.xstale[p] | (ok ? 1 : 0));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a.xendprop[(long long)(4 * g + k) * P + p] = yo[k];
        if (ok) a.xend[(long long)(4 * g + k) * P + p] = yo[k];
      }
    }
    __syncwarp(); /* every lane of the group has read ll and par before lane 0 changes them */
    if (act && ok && g == 0) {
      a.ll[p] = lltot;
      a.par[p] = (uint8_t)(1 - par);
    }
    const unsigned mk = __ballot_sync(0xFFFFFFFFu, ok && g == 0);
    if (lane == 0 && mk) atomicAdd(a.acc, (unsigned long long)__popc(mk));
  } else if (act) {
    if (g == 0) {
      if (a.do_ll) a.ll[p] = lltot;
      a.xstale[p] = sx ? 0 : 1;
    }
    if (a.write_end) {
#pragma unroll
      for (int k = 0; k < 4; k++) a.xend[(long long)(4 * g + k) * P + p] = yo[k];
    }
  }
}

template <int GK, int RNG>
static cudaError_t bb_wide4_launch(const bb_chain_args& a, cudaStream_t st) {
  const long long threads = 4 * (a.p_end - a.p_begin);
  const unsigned grid = (unsigned)((threads + BB_W4_THREADS - 1) / BB_W4_THREADS);
  bb_wide4_kernel<GK, RNG><<<grid, BB_W4_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}

/* Landmarks: the four-lane kernels; BB_WIDE_LANES=1 in the environment selects the one-thread-per-chain kernels
 * (tests compare the two bit for bit) */
static bb_chain_launch_fn bb_lookup_landmarks4(int gk, int auxm, int rng) {
  if (gk == 0) return rng == 0 ? &bb_wide4_launch<0, 0> : nullptr;
  if (gk == BB_GUIDE_NUH && auxm == 1) {
    if (rng == 0) return &bb_wide4_launch<BB_GUIDE_NUH, 0>;
    if (rng == 1) return &bb_wide4_launch<BB_GUIDE_NUH, 1>;
    if (rng == 3) return &bb_wide4_launch<BB_GUIDE_NUH, 3>;
  }
  return nullptr;
}
