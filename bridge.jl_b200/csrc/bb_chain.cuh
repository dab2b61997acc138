/*
 * bb_chain.cuh -- the path kernel: one thread owns one chain and walks its S segments in time.
 *
 * Replaces, fused in ONE pass over the data (reference call stacks A, C, D, E of SURVEY.md section 3):
 *   sample!(W2, Wiener())                      src/wiener.jl:50-58
 *   Wo.yy .= rho*W.yy + sqrt(1-rho^2)*W2.yy    test/partialbridgenuH.jl:178
 *   solve!(Euler(), Xo, x0, Wo, Po)            src/euler.jl:247-268 (guided) / :135-152 (plain)
 *   llikelihood(LeftRule(), Xo, Po; skip)      src/partialbridgenuH.jl:171-189, guip.jl:429-446,
 *                                              partialbridge.jl:67-87
 *   if log(rand()) <= llo - ll ... end         test/partialbridgenuH.jl:183-190
 *
 * Data layout in HBM ("chunked AoSoA"): the N grid points of a segment are cut into chunks of
 * BB_TC = 8; W is [S][NC][P][nbuf][8][d'] and X is [S][NC][P][8][d] doubles.  One chain therefore
 * owns 64*d' (64*d) contiguous, 64-byte aligned bytes per chunk: whole DRAM bursts.
 * W is double buffered per chain: a chain reads buffer par[p] and writes its proposal to buffer
 * 1-par[p]; accepting flips par[p] instead of copying the path.  The two buffers of a chain are
 * ADJACENT, so a warp reads half and writes the other half of the same contiguous 4 KB (d'=1)
 * whatever the accept/reject history of its 32 chains was.
 * X is single buffered: it always holds the path of the chain's last proposal (dense 128 B rows, no
 * per-chain scatter); xstale[p] marks chains whose proposal was rejected, and bb_ens_refresh_x
 * recomputes their current path from W on demand.  (Measured on B200: half-dense access -- two
 * separate double-buffer arrays, or an interleaved double-buffered X -- costs 25-35 % of the DRAM
 * bandwidth.)  All accesses are 256-bit (LDG.E.256 / STG.E.256).
 *
 * Per-step tables (dt, sqrt(dt), guiding term, auxiliary drift) are common to all chains: thread 0
 * streams them with 1-D TMA bulk copies into a BB_STAGES-deep shared-memory ring (full/empty
 * mbarriers), BB_LOOKAHEAD chunks ahead of the consumers; every thread reads them by broadcast LDS.
 *
 * State (y, W at the previous grid point, log-likelihood sum) lives in registers; noise comes from
 * Philox4x32-10 keyed by (seed; quad, iteration, global row) so results do not depend on the launch
 * geometry or on how chains are sharded over GPUs.
 */
#pragma once
#ifndef __CUDACC_RTC__
#include <stdlib.h>
#include <atomic>
#endif

#include "bb_device.cuh"

struct bb_chain_args {
  double* W[2];
  double* X;                       /* [S][NC][P][8][d]: the path of the LAST solve / proposal of every chain */
  uint8_t* par;                    /* [P] which buffer holds the chain's current state */
  const double* tab[BB_MAXSEG];    /* per-segment step tables, [NC*8][REC] */
  const double* start;             /* [d] (broadcast) or [d][P] */
  double* ll;                      /* [P] */
  double* llprop;                  /* [P] */
  double* logu;                    /* [P] */
  double* xend;                    /* [d][P] */
  double* xendprop;                /* [d][P] */
  uint8_t* accepted;               /* [P] */
  uint8_t* xstale;                 /* [P] 1: X is not the path of the chain's current W (rejected proposal) */
  const uint8_t* only;             /* if set, only chains with only[p] != 0 are processed (X refresh) */
  unsigned long long* acc;
  long long P;
  long long p_begin, p_end;        /* chains [p_begin, p_end) are processed by this launch */
  long long chain_offset;
  int S, N, NC;
  int jll;                         /* steps j <= jll (1-based end index) enter the log-likelihood */
  int start_bcast, store_x, do_ll, write_end;
  int nbuf;                        /* 1, or 2 when the ensemble is double buffered */
  int cpc;                         /* chains per CTA (<= threads that step chains; set by the launch code, see bb_pick_cpc) */
  bb_philox_keys keys;             /* Philox round keys of the seed */
  uint32_t stream;
  double rho, rho2;
  bb_model_dev model;
  /* per-segment constants: Bt[d*d], betat[d] (constant auxiliary drift), endflag, vend[d] (GuidedBridge end
   * point).  They sit in the kernel parameter bank, indexed by the warp-uniform segment number, so the
   * compiler keeps them in uniform registers / constant operands instead of per-thread registers. */
  double segc[BB_MAXSEG][BB_SEGC];
};

/* collects the K doubles a chain produces per grid point and writes them as 256-bit stores */
template <int K>
struct bb_rowout {
  double pend[3];
  __device__ __forceinline__ void put(double* row, int slot, const double* v, bool act) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      int e = slot * K + k;
      switch (e & 3) {
        case 0: pend[0] = v[k]; break;
        case 1: pend[1] = v[k]; break;
        case 2: pend[2] = v[k]; break;
        default:
          if (act) bb_st4(row + (e - 3), pend[0], pend[1], pend[2], v[k]);
      }
    }
  }
};

/* RNG: 0 = driving path W is read (solve!), 1 = pCN proposal (read W, write W°), 2 = fresh Wiener path
 * (sample! fused with solve!), 3 = pCN proposal without storing X°.  GK: 0 = plain Euler-Maruyama, else
 * bb_guide_kind; GM = rows of L (LMMU). */
template <class M, int GK, int GM, int AUXM, int RNG>
struct bb_chain {
  static constexpr int D = M::D, DP = M::DP;
  /* AUXM: 1 = constant auxiliary drift (per-segment constants), 0 = B~, beta~ tabulated per grid point,
   * 2 = tabulated + the terms of a non-constant-diffusion pair (a != a~): tr((a-a~)H) and a-a~ per grid point */
  static constexpr bool AUXC = (AUXM == 1), NCD = (AUXM == 2);
  /* RNG: 0 read W, 2 fresh Wiener path, 1 / 3 pCN proposal with X° stored / not stored (compile-time, so that the
   * compute-bound no-X launch carries none of the write-back staging code) */
  static constexpr bool PCN = (RNG == 1 || RNG == 3);
  static __device__ __forceinline__ bool sx(const bb_chain_args& a) {
    return RNG == 1 ? true : (RNG == 3 ? false : a.store_x != 0);
  }
  static constexpr int REC = bb_rec_len(GK, D, GM, AUXM);
  static constexpr int NCC = bb_rec_nc(GK, D, GM), NA1 = bb_rec_na1(GK, D, GM), NA2 = bb_rec_na2(GK, D, GM);
  static constexpr int OFF_C = 2, OFF_A1 = OFF_C + NCC, OFF_A2 = OFF_A1 + NA1, OFF_BT = OFF_A2 + NA2,
                       OFF_BE = OFF_BT + D * D, OFF_TR = OFF_BE + D, OFF_AD = OFF_TR + 1;

  /* Stores that trickle out 32 bytes at a time while the chain computes leave partially written lines in L2 for
   * microseconds and cost 15-20 % of DRAM bandwidth (tools/membench.cu: 5.8 TB/s in bursts vs 4.8 TB/s dripped),
   * so rows are completed in shared memory and written back to back.  The X° window needs 128 B per chain:
   * enabled where two CTAs per SM still fit (scalar noise, d <= 2). */
  static constexpr bool XBUF = BB_XFLUSH && (D <= 2) && (DP == 1);
  static constexpr int XWIN = 16 / D; /* steps per 128-byte window */

  struct state {
    double y[D];
    double wprev[DP];
    double w2[DP];
    double som;
  };

  /* guided drift _b((i,t),x,P°) at x = y (plain b for GK = 0) and, if in_ll, the log-likelihood term */
  static __device__ __forceinline__ void drift(const bb_model_dev& model, const double* __restrict__ R,
                                               const double* __restrict__ sc, const double* y, double dt,
                                               bool in_ll, double& som, double* bd) {
    M::b(model, y, bd);
    if constexpr (GK != 0) {
      /* r((i,t),x,P°) = A2 (c - A1 x)  -- partialbridgenuH.jl:161, guip.jl:193, partialbridge.jl:57 */
      double e[NCC], r[D];
      if constexpr (GK == BB_GUIDE_LMMU) {
        double Lx[GM];
        bb_matvec<GM, D>(R + OFF_A1, y, Lx);
#pragma unroll
        for (int k = 0; k < GM; k++) e[k] = R[OFF_C + k] - Lx[k];
        bb_matvec<D, GM>(R + OFF_A2, e, r);
      } else {
#pragma unroll
        for (int k = 0; k < D; k++) e[k] = R[OFF_C + k] - y[k];
        bb_matvec<D, D>(R + OFF_A2, e, r);
      }
      /* llikelihood term  <b - b~, r> dt  with b~ = B~ x + beta~  (partialbridgenuH.jl:176-181) */
      if (in_ll) {
        const double* Bt = AUXC ? sc : R + OFF_BT;
        const double* be = AUXC ? sc + D * D : R + OFF_BE;
        double bt[D], ee[D];
        bb_matvec<D, D>(Bt, y, bt);
#pragma unroll
        for (int k = 0; k < D; k++) ee[k] = bd[k] - (bt[k] + be[k]);
        som = fma(bb_vdot<D>(ee, r), dt, som);
        if constexpr (NCD) {
          /* if !constdiff(P°):  som -= 0.5 tr((a - a~) H) dt;  som += 0.5 r'(a - a~) r dt   src/partialbridge.jl:79-84 */
          som = fma(-(0.5 * R[OFF_TR]), dt, som);
          double rA[D];
#pragma unroll
          for (int j = 0; j < D; j++) {
            double q = r[0] * R[OFF_AD + j];
#pragma unroll
            for (int i = 1; i < D; i++) q = fma(r[i], R[OFF_AD + i * D + j], q);
            rA[j] = q;
          }
          som = fma(0.5 * bb_vdot<D>(rA, r), dt, som);
        }
      }
      /* _b((i,t),x,P°) = b + a r   (partialbridgenuH.jl:157-159); a = sigma sigma' from der[8..] */
      if constexpr (M::SPARSE) {
#pragma unroll
        for (int k = 0; k < D; k++)
          if (M::col(k) >= 0) bd[k] = fma(bb_adiag<M>(model, k), r[k], bd[k]);
      } else {
        double ar[D];
        bb_matvec<D, D>(model.der + 8, r, ar);
#pragma unroll
        for (int k = 0; k < D; k++) bd[k] = bd[k] + ar[k];
      }
    }
  }

  /* one grid point j >= 1: the Euler step j-1 -> j */
  static __device__ __forceinline__ void step(const bb_chain_args& a, const double* __restrict__ R,
                                              const double* __restrict__ sc, state& st, const double* wj,
                                              bool in_ll) {
    const double dt = R[0];
    double dw[DP];
#pragma unroll
    for (int k = 0; k < DP; k++) {
      dw[k] = wj[k] - st.wprev[k];
      st.wprev[k] = wj[k];
    }
    double bd[D];
    drift(a.model, R, sc, st.y, dt, in_ll, st.som, bd);
    bb_em_update<M>(a.model, bd, dt, dw, st.y);
  }

  /* The 4 d' driving values of group h (4 grid points) of chunk c -> out[4 d'] (element m = slot m / d', component
   * m % d'): fetch the d' pieces of W from the staged row (unless the path is sampled afresh), turn them into the values
   * that drive the Euler steps (pCN: W° = rho W + sqrt(1-rho^2) W2; sampling: W itself) and put them back (staged row,
   * or a direct 256-bit store).  None of this depends on the state y. */
  template <bool GENERIC>
  static __device__ __forceinline__ void gen_group(const bb_chain_args& a, const double* __restrict__ rec, state& st,
                                                   double* wq, double* wrow, double* wout_row, int c, int h,
                                                   uint32_t row_lo, uint32_t row_hi, bool wact, double* out) {
    constexpr int NPIECE = BB_TC * DP / 4; /* pieces of 4 doubles per chunk row */
#pragma unroll
    for (int pp = 0; pp < DP; pp++) {
      const int q = h * DP + pp, m = 4 * pp;
      if constexpr (RNG != 2) bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, wq);
      if constexpr (RNG != 0) {
        float z[4];
        bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + q), z);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int sl = 4 * h + (m + i) / DP, kk = (m + i) % DP; /* slot / component of element i */
          const bool first = GENERIC && (c * BB_TC + sl == 0);
          const double rootdt = rec[sl * REC + 1];
          if constexpr (PCN) {
            /* W2[j] = W2[j-1] + sqrt(dt) xi ;  W°[j] = rho W[j] + sqrt(1-rho^2) W2[j] */
            if (!first) st.w2[kk] = fma(rootdt, (double)z[i], st.w2[kk]);
            wq[i] = fma(a.rho2, st.w2[kk], a.rho * wq[i]);
          } else {
            /* W[j] = W[j-1] + sqrt(dt) xi, W[0] kept   (src/wiener.jl:50-58); w2 carries W[j-1] */
            st.w2[kk] = first ? wq[i] : fma(rootdt, (double)z[i], st.w2[kk]);
            wq[i] = st.w2[kk];
          }
        }
        /* memory-bound launches (X° stored too) complete the W° row in the staged row and write it as a whole
         * line at the end of the chunk; compute-bound ones store the piece directly */
        if (BB_WFLUSH && sx(a)) bb_sts4_swz(wrow, 2 * q, threadIdx.x & 7, wq);
        else if (wact) bb_st4(wout_row + 4 * q, wq[0], wq[1], wq[2], wq[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) out[m + i] = wq[i];
    }
  }

  /* One chunk of BB_TC grid points.  GENERIC = false is the steady state: every slot is a full step that
   * enters the log-likelihood; GENERIC = true also handles j = 0 (no step), j >= N (padding), steps that
   * `skip` excludes from the log-likelihood and the GuidedBridge end-point rule.
   * The chain's row of the driving path (16 d' doubles) is already in shared memory (`wrow`, requested BB_WSTAGES-1
   * chunks earlier).  SOFTWARE PIPELINE: a chain has two instruction streams per group of 4 steps -- the noise stream
   * (Philox, Box-Muller, the running sum W2, W°; ~half of the instructions, independent of y) and the dependent fp64
   * chain of the Euler steps.  The loop body holds the noise of group h+1 next to the steps of group h, so that the
   * compiler interleaves the two and a warp has twice the independent work in flight (the kernel is latency bound when
   * a GPU holds few chains: the strong-scaling share of BASELINE config 4 is 6.6 warps per SM). */
  template <bool GENERIC>
  static __device__ __forceinline__ void chunk(const bb_chain_args& a, const double* __restrict__ rec,
                                               const double* __restrict__ sc, state& st, double* wq,
                                               double* wrow, double* wout_row, double* xout_row, double* xbuf,
                                               int c, uint32_t row_lo, uint32_t row_hi, bool wact, bool xact) {
    constexpr int NPIECE = BB_TC * DP / 4;
    bb_rowout<D> xo;
    const int N = a.N;
    /* pipelined where the launch is memory bound (paths stored); the compute-bound pCN without X° (RNG 3) keeps the
     * noise of a group next to its own steps (measured: 4.3 ms vs 5.0 ms pipelined at 2.5e5 chains) */
    constexpr bool PIPE = (RNG != 3) && (DP <= BB_PIPE_MAXDP);
    double wg[4 * DP], wn[4 * DP];
    if constexpr (PIPE) gen_group<GENERIC>(a, rec, st, wq, wrow, wout_row, c, 0, row_lo, row_hi, wact, wg);
    /* groups of 4 grid points: the body is unrolled over one group only, which bounds code size and the
     * registers the scheduler spends on hoisted shared-memory loads */
#pragma unroll 1
    for (int h = 0; h < BB_TC / 4; h++) {
      if constexpr (PIPE) {
        if (h + 1 < BB_TC / 4) gen_group<GENERIC>(a, rec, st, wq, wrow, wout_row, c, h + 1, row_lo, row_hi, wact, wn);
      } else {
        gen_group<GENERIC>(a, rec, st, wq, wrow, wout_row, c, h, row_lo, row_hi, wact, wg);
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++) {
        const int slot = 4 * h + s4;
        const int j = c * BB_TC + slot;
        const double* R = rec + slot * REC;
        double wj[DP];
#pragma unroll
        for (int k = 0; k < DP; k++) wj[k] = wg[s4 * DP + k];
        if (GENERIC && j == 0) {
#pragma unroll
          for (int k = 0; k < DP; k++) st.wprev[k] = wj[k];
        } else if (!GENERIC || j < N) {
          step(a, R, sc, st, wj, !GENERIC || j <= a.jll);
          if (GENERIC && GK == BB_GUIDE_HV && j == N - 1 && sc[D * D + D] != 0.0) {
            /* endpoint(y, P::GuidedBridge) = V[end] when H♢[end] = 0   src/euler.jl:241-242 */
#pragma unroll
            for (int k = 0; k < D; k++) st.y[k] = sc[D * D + D + 1 + k];
          }
        }
        if constexpr (XBUF) {
          if (sx(a)) {
            const int ls = (4 * h + s4) % XWIN;
            if constexpr (D == 2)
              *reinterpret_cast<double2*>(xbuf + 2 * ((ls ^ threadIdx.x) & 7)) = make_double2(st.y[0], st.y[1]);
            else
              xbuf[2 * (((ls >> 1) ^ threadIdx.x) & 7) + (ls & 1)] = st.y[0];
          }
        } else {
          xo.put(xout_row + 4 * h * D, s4, st.y, xact);
        }
      }
      if constexpr (PIPE) {
#pragma unroll
        for (int i = 0; i < 4 * DP; i++) wg[i] = wn[i];
      }
      if constexpr (XBUF) {
        /* a 128-byte window of X° is complete: write it back to back (whole line) */
        if (((4 * h + 3) % XWIN) == XWIN - 1 && xact) {
          double* xdst = xout_row + (4 * h + 4 - XWIN) * D;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            double v[4];
            bb_lds4_swz(xbuf, 2 * q, threadIdx.x & 7, v);
            bb_st4(xdst + 4 * q, v[0], v[1], v[2], v[3]);
          }
        }
      }
    }
    if constexpr (RNG != 0 && BB_WFLUSH) {
      /* the chain's row of W° is complete in shared memory: write its 128 d' bytes back to back */
      if (wact && sx(a)) {
#pragma unroll
        for (int q = 0; q < NPIECE; q++) {
          double v[4];
          bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, v);
          bb_st4(wout_row + 4 * q, v[0], v[1], v[2], v[3]);
        }
      }
    }
  }

  static __device__ __forceinline__ void run(const bb_chain_args& a) {
    /* the table ring: BB_STAGES stages of BB_TSTAGE chunks (the last stage of a segment may be shorter) */
    constexpr uint32_t CHUNK_DOUBLES = BB_TC * REC;
    constexpr uint32_t STAGE_DOUBLES = BB_TSTAGE * CHUNK_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + BB_STAGES * STAGE_DOUBLES);
    uint64_t* empty = full + BB_STAGES;
    /* staging of the driving path: BB_WSTAGES x (one row per chain of the CTA).  The 16-byte pieces of every
     * 128-byte block of a row are XOR-swizzled with the chain's index (piece k of chain t sits at k ^ (t & 7)),
     * which keeps the 128-bit reads of a quarter-warp conflict free without padding. */
    constexpr int WROWP = BB_TC * DP;
    double* wstage = reinterpret_cast<double*>(empty + BB_STAGES);
    double* xbuf = wstage + (size_t)BB_WSTAGES * BB_THREADS * (BB_TC * DP) + (size_t)threadIdx.x * 16;

    const int S = a.S, NC = a.NC;
    const int NST = (NC + BB_TSTAGE - 1) / BB_TSTAGE; /* stages per segment */
    const int T = S * NST;                            /* stages this CTA walks through */
    const int lane = threadIdx.x & 31;
    /* warps that hold no chain of this CTA (chains per CTA below the CTA size, or the ragged last CTA) leave at once:
     * they are not counted in the consumer barriers and issue nothing */
    const long long cta_p0 = a.p_begin + (long long)blockIdx.x * a.cpc;
    const long long cta_n = (a.p_end - cta_p0 < a.cpc) ? a.p_end - cta_p0 : a.cpc;
    const int nwarps = (int)((cta_n + 31) >> 5);

    if (threadIdx.x == 0) {
      for (int i = 0; i < BB_STAGES; i++) {
        bb_mbar_init(&full[i], 1);
        bb_mbar_init(&empty[i], nwarps);
      }
      bb_mbar_fence_init();
    }
    __syncthreads();
    if ((int)(threadIdx.x >> 5) >= nwarps) return;

    const long long P = a.P;
    const long long p = a.p_begin + (long long)blockIdx.x * a.cpc + threadIdx.x;
    const long long pc = p < a.p_end ? p : a.p_end - 1;
    const bool act = (int)threadIdx.x < a.cpc && p < a.p_end && (!a.only || a.only[pc] != 0);
    const int par = a.par[pc];
    const int rbuf = par;                        /* where the chain's current W (and X) live */
    const int wbuf = PCN ? 1 - par : par; /* where this launch writes */
    const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
    const bool xact = act && sx(a);

    state st;
#pragma unroll
    for (int k = 0; k < D; k++) st.y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc];
    st.som = 0.0;
    double lltot = 0.0; /* sum over segments of the per-segment log-likelihoods (bolus3.jl:331-333) */

    /* a chain's two buffers are adjacent: W[1] = W[0] + 8 d', slot of chain p = p * nbuf * 8 d' */
    const double* wr = a.W[rbuf] + pc * (a.nbuf * BB_TC * DP);
    double* ww = a.W[wbuf] + pc * (a.nbuf * BB_TC * DP);
    double* xw = sx(a) ? a.X + pc * (BB_TC * D) : nullptr;
    const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);

    /* producer state (thread 0 only): next stage to request */
    int issued = 0, iseg = 0, ist_in_seg = 0;
    int stage = 0, gs = 0; /* consumer: ring slot and global stage index */
    bool tab_ready = false;
    uint32_t phase = 0;

    double wq[4] = {0.0, 0.0, 0.0, 0.0};
    /* ---- driving path: the warp copies the rows of its 32 chains for one chunk with 8 d' cp.async
     * instructions (LDGSTS.128, L1 bypassed): lane l moves 16-byte piece t % (8 d') of chain t / (8 d'),
     * t = 32 r + l, so every chain's 128 d'-byte row is requested by ONE instruction -- whole L2 lines even
     * though neighbouring chains read different buffers -- and lands in the chain's swizzled row of the stage.
     * Rows are requested BB_WSTAGES-1 chunks before they are used; no registers are tied up meanwhile. */
    constexpr int NCP = BB_TC * DP / 2; /* 16-byte pieces per row = cp.async instructions per chunk and warp */
    const int warp = threadIdx.x >> 5;
    const unsigned amask = __ballot_sync(0xFFFFFFFFu, act);
    const unsigned pmask = __ballot_sync(0xFFFFFFFFu, par != 0);
    const long long warp_p0 = a.p_begin + (long long)blockIdx.x * a.cpc + warp * 32;
    const uint32_t wrow_d = (uint32_t)(a.nbuf * BB_TC * DP); /* doubles between the slots of consecutive chains */
    constexpr bool COOP_FAST = (32 % NCP == 0); /* an instruction covers CPI = 32 / NCP whole chains */
    constexpr int CPI = COOP_FAST ? 32 / NCP : 1;
    /* instruction r, lane l: chain cw = CPI r + l / NCP, piece = l % NCP (fast mapping) */
    const uint32_t l_cw = (uint32_t)lane / NCP, l_pc = (uint32_t)lane % NCP;
    uint32_t parbits = 0, actbits = 0; /* bit r: buffer parity / activity of the chain instruction r serves */
#pragma unroll
    for (int r = 0; r < NCP; r++) {
      const uint32_t cw = COOP_FAST ? CPI * r + l_cw : (32u * r + lane) / NCP;
      parbits |= ((pmask >> cw) & 1u) << r;
      actbits |= ((amask >> cw) & 1u) << r;
    }
    /* per-lane constants: source offset within a chunk (doubles) and swizzled smem offsets (doubles) for the
     * 8 / CPI distinct values of (chain index & 7) that the instructions of this lane meet */
    const uint32_t src_l = l_cw * wrow_d + 2 * l_pc;
    constexpr int NSW = COOP_FAST ? (8 / CPI > 0 ? 8 / CPI : 1) : 1;
    uint32_t dsw[NSW];
#pragma unroll
    for (int i = 0; i < NSW; i++) {
      const uint32_t cw = CPI * i + l_cw;
      dsw[i] = (warp * 32 + l_cw) * WROWP + 2 * ((l_pc & ~7u) | ((l_pc ^ cw) & 7u));
    }
    const double* wsrc = a.W[0] + warp_p0 * wrow_d; /* + chunk * wstride */
    const int TW = S * NC;                          /* rows this chain walks through */
    int gcur = 0;                                   /* global chunk index of the row consumed next */
    auto w_issue = [&](int gc) { /* request chunk gc for the whole warp; always commits a (possibly empty) group */
      if (RNG != 2 && gc < TW) {
        double* dst = wstage + (size_t)(gc % BB_WSTAGES) * BB_THREADS * WROWP;
        if constexpr (COOP_FAST) {
          const double* src = wsrc + (long long)gc * wstride + src_l;
#pragma unroll
          for (int r = 0; r < NCP; r++) {
            const uint32_t poff = ((parbits >> r) & 1u) * (uint32_t)(BB_TC * DP);
            if ((actbits >> r) & 1u)
              bb_cp_async16(dst + dsw[r % NSW] + r * (CPI * WROWP), src + (r * CPI * wrow_d + poff));
          }
        } else {
          const double* src = wsrc + (long long)gc * wstride;
#pragma unroll
          for (int r = 0; r < NCP; r++) {
            const uint32_t t = 32u * r + lane, cw = t / NCP, piece = t % NCP;
            if ((amask >> cw) & 1u)
              bb_cp_async16(dst + (warp * 32 + cw) * WROWP + 2 * ((piece & ~7u) | ((piece ^ cw) & 7u)),
                            src + (cw * wrow_d + ((pmask >> cw) & 1u) * (BB_TC * DP) + 2 * piece));
          }
        }
      }
      bb_cp_async_commit();
    };
    double* wslot = wstage + (size_t)threadIdx.x * WROWP; /* + stage * BB_THREADS * WROWP */
#pragma unroll 1
    for (int i = 0; i < BB_WSTAGES - 1; i++) w_issue(i);

    for (int s = 0; s < S; s++) {
      const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
      const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
      const double* sc = a.segc[s];
#pragma unroll
      for (int k = 0; k < DP; k++) st.w2[k] = 0.0;
      st.som = 0.0;
      if constexpr (RNG == 2) { if (act) bb_ld4(wr, wq); } /* only W[0] of the segment is read (it is kept) */
      const double* rec = ring;
      for (int c = 0; c < NC; c++) {
        if ((c % BB_TSTAGE) == 0) {
          if (threadIdx.x == 0) {
            while (issued < T && issued <= gs + BB_LOOKAHEAD) {
              const int ist = issued % BB_STAGES;
              if (issued >= BB_STAGES) bb_mbar_wait(&empty[ist], ((issued / BB_STAGES) - 1) & 1);
              const int c0 = ist_in_seg * BB_TSTAGE;
              const int nch = (NC - c0 < BB_TSTAGE) ? NC - c0 : BB_TSTAGE;
              const uint32_t bytes = (uint32_t)nch * CHUNK_DOUBLES * 8;
              bb_mbar_expect_tx(&full[ist], bytes);
              bb_tma_load_1d(ring + ist * STAGE_DOUBLES, a.tab[iseg] + (size_t)c0 * CHUNK_DOUBLES, bytes,
                             &full[ist]);
              issued++;
              if (++ist_in_seg == NST) { ist_in_seg = 0; iseg++; }
            }
          }
          __syncwarp();
          /* the stage was requested BB_LOOKAHEAD stages ago; `tab_ready` is the answer of a non-blocking probe
           * made one chunk earlier, so the barrier's round trip is normally off the critical path */
          if (!tab_ready) bb_mbar_wait(&full[stage], phase);
          rec = ring + stage * STAGE_DOUBLES;
        }
        double* wrow = wslot + (size_t)(gcur % BB_WSTAGES) * BB_THREADS * WROWP;
        if constexpr (RNG != 2) {
#if BB_WSTAGES == 1
          /* single stage (d' >= 2: a second one would halve the chains an SM can hold): every lane has written the
           * previous chunk's row back, the warp copies this chunk's rows -- which the lanes asked L2 for a whole
           * chunk ago -- and waits for them; the other warps of the SM cover the L2 round trip */
          __syncwarp();
          w_issue(gcur);
          bb_cp_async_wait<0>();
          __syncwarp();
          if (act && gcur + 1 < TW) {
#pragma unroll
            for (int l = 0; l < (BB_TC * DP * 8 + 127) / 128; l++)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(wr + wstride + l * 16));
          }
#else
          /* rows of chunk gcur were requested a whole chunk ago: wait for them first (normally no wait at all),
           * then refill the other stage -- which every lane has finished reading -- with chunk gcur+1 */
          bb_cp_async_wait<BB_WSTAGES - 2>();
          __syncwarp();
          w_issue(gcur + BB_WSTAGES - 1);
#endif
        }
        if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
          /* probe the next table stage now; its result is needed only after this chunk */
          const int nst = (stage + 1 == BB_STAGES) ? 0 : stage + 1;
          const uint32_t nph = (stage + 1 == BB_STAGES) ? phase ^ 1 : phase;
          tab_ready = (gs + 1 < T) ? bb_mbar_test(&full[nst], nph) : true;
        }
        const bool generic = (c == 0) || (c == NC - 1) || (c * BB_TC + BB_TC - 1 > a.jll);
        if (generic)
          chunk<true>(a, rec, sc, st, wq, wrow, ww, xw, xbuf, c, row_lo, row_hi, act, xact);
        else
          chunk<false>(a, rec, sc, st, wq, wrow, ww, xw, xbuf, c, row_lo, row_hi, act, xact);
        rec += CHUNK_DOUBLES;
        gcur++;
        if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
          __syncwarp();
          if (lane == 0) bb_mbar_arrive(&empty[stage]);
          if (++stage == BB_STAGES) { stage = 0; phase ^= 1; }
          gs++;
        }
        wr += wstride;
        ww += wstride;
        xw += xstride;
      }
      lltot += st.som;
    }

    /* ---- per-chain epilogue */
    if constexpr (PCN) {
      /* accept iff log(U) <= ll° - ll   (test/partialbridgenuH.jl:183) */
      const double logu = bb_accept_logu(a.keys, a.stream, chain);
      const double llc = a.ll[pc];
      const bool ok = act && (logu <= lltot - llc);
      if (act) {
        a.llprop[p] = lltot;
        a.logu[p] = logu;
        a.accepted[p] = ok ? 1 : 0;
        /* X now holds this proposal's path (if stored): it is the chain's current path iff accepted */
        a.xstale[p] = sx(a) ? (ok ? 0 : 1) : (uint8_t)(a.xstale[p] | (ok ? 1 : 0));
#pragma unroll
        for (int k = 0; k < D; k++) a.xendprop[(long long)k * P + p] = st.y[k];
        if (ok) {
          a.ll[p] = lltot;
          a.par[p] = (uint8_t)(1 - par);
#pragma unroll
          for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = st.y[k];
        }
      }
      const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
      if (lane == 0 && m) atomicAdd(a.acc, (unsigned long long)__popc(m));
    } else {
      if (act) {
        if (a.do_ll) a.ll[p] = lltot;
        a.xstale[p] = sx(a) ? 0 : 1; /* a solve that does not store X leaves X behind the chain's current state */
        if (a.write_end) {
#pragma unroll
          for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = st.y[k];
        }
      }
    }
  }
};

/* Resident CTAs per SM the kernel is compiled for.  With d' >= 2 the staging of the driving path alone takes
 * 2 x 256 x 128 d' bytes >= 128 KB of shared memory, so only ONE CTA fits an SM anyway: compiling those kernels for two
 * (a 128-register cap) bought nothing and made them spill 70-420 bytes per thread (config 3, LinPro d = 3:
 * 2.75 ms -> 0.83 ms for guided Euler + ll once the cap is lifted). */
template <class M>
constexpr int bb_min_ctas() { return M::DP >= 2 ? BB_MINB_WIDE : BB_MINB; }

template <class M, int GK, int GM, int AUXM, int RNG>
__global__ void __launch_bounds__(BB_THREADS, bb_min_ctas<M>()) bb_chain_kernel(const __grid_constant__ bb_chain_args a) {
  bb_chain<M, GK, GM, AUXM, RNG>::run(a);
}

#ifndef __CUDACC_RTC__ /* host-side launch + lookup (not part of a run-time compiled user-model kernel) */
template <class M, int GK, int GM, int AUXM, int RNG>
static inline size_t bb_chain_smem(int S) {
  (void)S;
  return (size_t)BB_STAGES * BB_TSTAGE * BB_TC * bb_rec_len(GK, M::D, GM, AUXM) * 8 + 2 * BB_STAGES * 8 +
         (size_t)BB_WSTAGES * BB_THREADS * (BB_TC * M::DP) * 8 +
         (bb_chain<M, GK, GM, AUXM, RNG>::XBUF ? (size_t)BB_THREADS * 128 : 0);
}

/* ---- host-side launch + lookup, one translation unit per model (bb_inst_*.cu) */
typedef cudaError_t (*bb_chain_launch_fn)(const bb_chain_args&, cudaStream_t);

/* Chains per CTA.  All chains cost the same, so a launch is a number of "waves" of resident CTAs; with full CTAs the last
 * wave is partial (2.5e5 chains = 977 CTAs of 256 on 296 slots = 3.3 waves) and its CTAs, alone on the GPU, run at their
 * latency floor instead of the memory system's rate -- or, for one wave, some SMs hold one CTA more than others.  Slightly
 * smaller CTAs that fill a whole number of waves keep every SM equally loaded to the end.  It pays for one or two waves
 * (6.25e4 chains: 1.715 -> 1.549 ms, 1.25e5: 3.21 -> 3.06 ms, warp-specialised kernel at 31 250: 0.880 -> 0.856 ms); from
 * three (one CTA per SM: four) waves on the idle lanes cost more than the partial wave (fewer active warps per SM).
 * BB_CPC overrides (tuning). */
static inline int bb_pick_cpc(long long n, int cmax, int ctas_per_sm) {
  static const int env = []() { const char* v = getenv("BB_CPC"); return v ? atoi(v) : 0; }();
  if (env > 0) return env < cmax ? env : cmax;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long slots = (long long)sms * ctas_per_sm;
  const long long waves = (n + slots * cmax - 1) / (slots * cmax);
  /* measured: 2.5e5 chains (2 CTAs per SM) as 4 waves of 212 run 6.13 ms, as 3.3 waves of 256 5.82 ms; config 3 (1e5 chains,
   * one CTA per SM) as 3 waves of 228 0.800 ms, as 2.6 waves of 256 0.834 ms; 6.6 -> 7 waves (FHN diagonal) 8.5 -> 9.1 ms */
  if (waves > (ctas_per_sm == 1 ? 3 : 2)) return cmax;
  long long cpc = (n + waves * slots - 1) / (waves * slots);
  cpc = (cpc + 3) & ~3ll;
  if (cpc < cmax / 4) cpc = cmax / 4; /* tiny ensembles: do not stream the tables for a handful of chains per CTA */
  return (int)(cpc < cmax ? cpc : cmax);
}

template <class M, int GK, int GM, int AUXM, int RNG>
static cudaError_t bb_chain_launch(const bb_chain_args& a, cudaStream_t st) {
  const size_t smem = bb_chain_smem<M, GK, GM, AUXM, RNG>(a.S);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration; /* d = d' = 3 with the widest step records: BB_ERR_UNSUPPORTED */
  /* per instantiation: bit i = attribute set on device i (it is per device).  Atomic: contexts of different host
   * threads launch concurrently; setting the attribute twice is harmless.  Devices >= 64 always set it. */
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((attr_done.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(bb_chain_kernel<M, GK, GM, AUXM, RNG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_release);
  }
  bb_chain_args b = a;
  b.cpc = bb_pick_cpc(a.p_end - a.p_begin, BB_THREADS, bb_min_ctas<M>());
  const unsigned grid = (unsigned)((a.p_end - a.p_begin + b.cpc - 1) / b.cpc);
  bb_chain_kernel<M, GK, GM, AUXM, RNG><<<grid, BB_THREADS, smem, st>>>(b);
  return cudaGetLastError();
}

template <class M, int GK, int GM, int AUXM, int RNG>
static cudaError_t bb_chain_ws_launch(const bb_chain_args& a, cudaStream_t st); /* bb_chain_ws.cuh */
template <class M, int GK, int GM, int AUXM>
static cudaError_t bb_chain_ws2_launch(const bb_chain_args& a, cudaStream_t st); /* bb_chain_ws.cuh */

template <class M, int GK, int GM>
static bb_chain_launch_fn bb_lookup_guide(int auxm, int rng) {
  if (rng == 0) return auxm == 1 ? &bb_chain_launch<M, GK, GM, 1, 0>
                                 : (auxm == 0 ? &bb_chain_launch<M, GK, GM, 0, 0> : &bb_chain_launch<M, GK, GM, 2, 0>);
  if (rng == 1) return auxm == 1 ? &bb_chain_launch<M, GK, GM, 1, 1>
                                 : (auxm == 0 ? &bb_chain_launch<M, GK, GM, 0, 1> : &bb_chain_launch<M, GK, GM, 2, 1>);
  if (rng == 3) return auxm == 1 ? &bb_chain_launch<M, GK, GM, 1, 3>
                                 : (auxm == 0 ? &bb_chain_launch<M, GK, GM, 0, 3> : &bb_chain_launch<M, GK, GM, 2, 3>);
  /* 5: pCN with X° stored (mode 1) as the warp-specialised kernel (bb_chain_ws.cuh), scalar-noise models */
  if constexpr (M::DP == 1) {
    if (rng == 5) return auxm == 1 ? &bb_chain_ws_launch<M, GK, GM, 1, 1>
                                   : (auxm == 0 ? &bb_chain_ws_launch<M, GK, GM, 0, 1> : &bb_chain_ws_launch<M, GK, GM, 2, 1>);
    /* 6: the same with two chains per dynamics thread (state dimension <= 2: X° through the shared-memory window) */
    if constexpr (M::D <= 2) {
      if (rng == 6) return auxm == 1 ? &bb_chain_ws2_launch<M, GK, GM, 1>
                                     : (auxm == 0 ? &bb_chain_ws2_launch<M, GK, GM, 0> : &bb_chain_ws2_launch<M, GK, GM, 2>);
    }
  }
  return nullptr;
}
template <class M>
static bb_chain_launch_fn bb_lookup_unguided(int rng) {
  if (rng == 0) return &bb_chain_launch<M, 0, 0, true, 0>;
  if (rng == 2) return &bb_chain_launch<M, 0, 0, true, 2>;
  return nullptr;
}
template <class M>
static bb_chain_launch_fn bb_lookup_model(int gk, int gm, int auxc, int rng) {
  if (gk == 0) return bb_lookup_unguided<M>(rng);
  if (gk == BB_GUIDE_NUH) return bb_lookup_guide<M, BB_GUIDE_NUH, 0>(auxc, rng);
  if (gk == BB_GUIDE_HV) return bb_lookup_guide<M, BB_GUIDE_HV, 0>(auxc, rng);
  if (gk == BB_GUIDE_LMMU) {
    if (gm == 1) return bb_lookup_guide<M, BB_GUIDE_LMMU, 1>(auxc, rng);
    if constexpr (M::D >= 2)
      if (gm == 2) return bb_lookup_guide<M, BB_GUIDE_LMMU, 2>(auxc, rng);
    if constexpr (M::D >= 3)
      if (gm == 3) return bb_lookup_guide<M, BB_GUIDE_LMMU, 3>(auxc, rng);
  }
  return nullptr;
}

#include "bb_chain_ws.cuh" /* the warp-specialised pCN kernel (defined in terms of bb_chain) */
#endif /* !__CUDACC_RTC__ */
