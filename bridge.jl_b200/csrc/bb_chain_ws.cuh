/*
 * bb_chain_ws.cuh -- the pCN iteration as a WARP-SPECIALISED kernel: two threads per chain.
 *
 * One pCN step of a chain is two instruction streams of about equal length:
 *   noise     sample!(W2, Wiener()) and Wo.yy .= rho*W.yy + sqrt(1-rho^2)*W2.yy   (src/wiener.jl:50-58,
 *             test/partialbridgenuH.jl:178): Philox4x32-10, float32 Box-Muller, the running sum W2, W° -- INT / FP32 /
 *             a few fp64 operations, independent of the chain's state;
 *   dynamics  solve!(Euler(), Xo, x0, Wo, Po) + llikelihood(LeftRule(), Xo, Po)    (src/euler.jl:247-268,
 *             src/partialbridgenuH.jl:171-189): the dependent fp64 chain of the guided Euler steps.
 * In bb_chain_kernel one thread runs both, and a warp is stalled on fixed-latency dependencies most of the time
 * (ncu: 52-58 % of the issue slots used, `wait` the top stall): at full size the memory system hides it, at the
 * strong-scaling share of a GPU (31 250 chains = 6.6 warps per SM) nothing does.  Here the upper half of a CTA's warps
 * are NOISE warps and the lower half DYNAMICS warps; warp pair k serves the same 32 chains.  The noise warp brings the
 * chains' rows of W into shared memory (cp.async, BB_WS_PF chunks ahead), turns them into W° in place, hands the row to
 * its dynamics warp through an mbarrier (full / empty per ring slot, BB_WS_WST slots) and writes the finished row back as
 * whole 128-byte lines; the dynamics warp consumes the row, steps the chain and writes X°.  Twice the warps, each with
 * half the instructions: the data, the arithmetic and every rounding are those of bb_chain_kernel (bit-identical
 * results; the GPU tests run both).
 *
 * Measured (B200, FitzHugh-Nagumo config 4, profiles/r02_ws_ab.txt): 31 250 chains 0.929 -> 0.880 ms; 250 000 chains
 * 5.93 ms either way (memory bound); without X° and for d' >= 2 the one-thread kernel is faster.  The library therefore
 * launches this kernel for scalar-noise models with X° stored when the ensemble is SMALL (fewer than 2 CTAs of 128
 * chains per SM), and bb_chain_kernel otherwise.  What still limits the small case is the dynamics warp itself: a
 * serial chain of ~32 dependent fp64 operations per step at ~8 cycles each (no-noise experiment: 0.82 ms).
 */
#pragma once
#include "bb_chain.cuh"

#ifndef BB_WS_WST
#define BB_WS_WST 4     /* slots of the W ring of a warp pair */
#endif
#ifndef BB_WS_PF
#define BB_WS_PF 2      /* chunks the noise warp requests its rows ahead (a chunk of noise work is shorter than the DRAM latency) */
#endif
#define BB_WS_MAXPAIR 4 /* warp pairs per CTA (256 threads) */

template <class M, int GK, int GM, int AUXM, int RNG>
struct bb_chain_ws {
  static_assert(RNG == 1 || RNG == 3, "pCN modes only");
  using Dyn = bb_chain<M, GK, GM, AUXM, 0>; /* the dynamics warp runs the read-W chunk of the one-thread kernel */
  static constexpr int D = M::D, DP = M::DP, REC = Dyn::REC;
  static constexpr bool SX = (RNG == 1);
  static constexpr int WROWP = BB_TC * DP, NPIECE = BB_TC * DP / 4;
  /* mbarriers: table ring full / empty, then W ring full / empty per (pair, slot); padded to whole 128-byte lines */
  static constexpr int BAR_BYTES = ((2 * BB_STAGES + 2 * BB_WS_MAXPAIR * BB_WS_WST) * 8 + 127) / 128 * 128;

  static __host__ __device__ constexpr size_t smem_bytes(int nt) {
    return (size_t)BB_STAGES * BB_TSTAGE * BB_TC * REC * 8 + BAR_BYTES +
           (size_t)BB_WS_WST * (nt / 2) * WROWP * 8 + (Dyn::XBUF ? (size_t)(nt / 2) * 128 : 0);
  }

  /* W° of one chunk: every piece of the chain's staged row is replaced in place (pieces in time order: the running
   * sum W2 is sequential) */
  template <bool FIRST>
  static __device__ __forceinline__ void noise_chunk(const bb_chain_args& a, const double* __restrict__ rec, double* w2,
                                                     double* wrow, int c, uint32_t row_lo, uint32_t row_hi) {
#pragma unroll 1
    for (int h = 0; h < BB_TC / 4; h++) {
#pragma unroll
      for (int pp = 0; pp < DP; pp++) {
        const int q = h * DP + pp, m = 4 * pp;
        double wq[4];
        bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, wq);
        float z[4];
        bb_normal_quad(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + q), z);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int sl = 4 * h + (m + i) / DP, kk = (m + i) % DP;
          const bool first = FIRST && (sl == 0);
          const double rootdt = rec[sl * REC + 1];
          /* W2[j] = W2[j-1] + sqrt(dt) xi ;  W°[j] = rho W[j] + sqrt(1-rho^2) W2[j] */
          if (!first) w2[kk] = fma(rootdt, (double)z[i], w2[kk]);
          wq[i] = fma(a.rho2, w2[kk], a.rho * wq[i]);
        }
        bb_sts4_swz(wrow, 2 * q, threadIdx.x & 7, wq);
      }
    }
  }

  static __device__ __forceinline__ void run(const bb_chain_args& a) {
    constexpr uint32_t CHUNK_DOUBLES = BB_TC * REC;
    constexpr uint32_t STAGE_DOUBLES = BB_TSTAGE * CHUNK_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NW = blockDim.x >> 5, NPAIR = NW >> 1, CH = blockDim.x >> 1;
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + BB_STAGES * STAGE_DOUBLES);
    uint64_t* empty = full + BB_STAGES;
    uint64_t* wfull = empty + BB_STAGES;                  /* [pair][slot] */
    uint64_t* wempty = wfull + BB_WS_MAXPAIR * BB_WS_WST;
    double* wstage = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(full) + BAR_BYTES);
    double* xbuf_all = wstage + (size_t)BB_WS_WST * CH * WROWP;

    const int S = a.S, NC = a.NC;
    const int NST = (NC + BB_TSTAGE - 1) / BB_TSTAGE;
    const int T = S * NST;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool noise = warp >= NPAIR;
    const int pair = noise ? warp - NPAIR : warp;
    const int ci = pair * 32 + lane; /* chain of this thread within the CTA */

    /* warp pairs that hold no chain of this CTA (chains per CTA below 32 NPAIR, or the ragged last CTA) leave at once */
    const long long cta_p0 = a.p_begin + (long long)blockIdx.x * a.cpc;
    const long long cta_n = (a.p_end - cta_p0 < a.cpc) ? a.p_end - cta_p0 : a.cpc;
    const int npair_act = (int)((cta_n + 31) >> 5);
    if (threadIdx.x == 0) {
      for (int i = 0; i < BB_STAGES; i++) {
        bb_mbar_init(&full[i], 1);
        bb_mbar_init(&empty[i], 2 * npair_act);
      }
      for (int i = 0; i < BB_WS_MAXPAIR * BB_WS_WST; i++) {
        bb_mbar_init(&wfull[i], 1);
        bb_mbar_init(&wempty[i], 1);
      }
      bb_mbar_fence_init();
    }
    __syncthreads();
    if (pair >= npair_act) return;

    const long long P = a.P;
    const long long p = a.p_begin + (long long)blockIdx.x * a.cpc + ci;
    const long long pc = p < a.p_end ? p : a.p_end - 1;
    const bool act = ci < a.cpc && p < a.p_end && (!a.only || a.only[pc] != 0);
    const int par = a.par[pc];
    const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
    const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
    const int TW = S * NC;
    uint64_t* my_full = wfull + pair * BB_WS_WST;
    uint64_t* my_empty = wempty + pair * BB_WS_WST;
    double* wslot = wstage + (size_t)ci * WROWP; /* + slot * CH * WROWP */

    /* table ring: thread 0 produces (1-D TMA bulk copies), every warp of the CTA consumes */
    int issued = 0, iseg = 0, ist_in_seg = 0;
    int stage = 0, gs = 0;
    bool tab_ready = false;
    uint32_t phase = 0;
    const double* rec = ring;
    auto tab_acquire = [&](int c) {
      if ((c % BB_TSTAGE) == 0) {
        if (threadIdx.x == 0) {
          while (issued < T && issued <= gs + BB_LOOKAHEAD) {
            const int ist = issued % BB_STAGES;
            if (issued >= BB_STAGES) bb_mbar_wait(&empty[ist], ((issued / BB_STAGES) - 1) & 1);
            const int c0 = ist_in_seg * BB_TSTAGE;
            const int nch = (NC - c0 < BB_TSTAGE) ? NC - c0 : BB_TSTAGE;
            const uint32_t bytes = (uint32_t)nch * CHUNK_DOUBLES * 8;
            bb_mbar_expect_tx(&full[ist], bytes);
            bb_tma_load_1d(ring + ist * STAGE_DOUBLES, a.tab[iseg] + (size_t)c0 * CHUNK_DOUBLES, bytes, &full[ist]);
            issued++;
            if (++ist_in_seg == NST) { ist_in_seg = 0; iseg++; }
          }
        }
        __syncwarp();
        if (!tab_ready) bb_mbar_wait(&full[stage], phase);
        rec = ring + stage * STAGE_DOUBLES;
      }
    };
    auto tab_probe = [&](int c) { /* non-blocking look at the next stage, one chunk before it is needed */
      if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
        const int nst = (stage + 1 == BB_STAGES) ? 0 : stage + 1;
        const uint32_t nph = (stage + 1 == BB_STAGES) ? phase ^ 1 : phase;
        tab_ready = (gs + 1 < T) ? bb_mbar_test(&full[nst], nph) : true;
      }
    };
    auto tab_release = [&](int c) {
      rec += CHUNK_DOUBLES;
      if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
        __syncwarp();
        if (lane == 0) bb_mbar_arrive(&empty[stage]);
        if (++stage == BB_STAGES) { stage = 0; phase ^= 1; }
        gs++;
      }
    };

    if (noise) {
      /* ================================================================== NOISE warp */
      /* cooperative copy of the warp's 32 rows, mapping as in bb_chain::run: one cp.async instruction requests whole
       * 128 d'-byte rows (whole L2 lines even though neighbouring chains read different buffers) */
      constexpr int NCP = BB_TC * DP / 2;
      const unsigned amask = __ballot_sync(0xFFFFFFFFu, act);
      const unsigned pmask = __ballot_sync(0xFFFFFFFFu, par != 0);
      const long long warp_p0 = a.p_begin + (long long)blockIdx.x * a.cpc + pair * 32;
      const uint32_t wrow_d = (uint32_t)(a.nbuf * BB_TC * DP);
      constexpr bool COOP_FAST = (32 % NCP == 0);
      constexpr int CPI = COOP_FAST ? 32 / NCP : 1;
      const uint32_t l_cw = (uint32_t)lane / NCP, l_pc = (uint32_t)lane % NCP;
      uint32_t parbits = 0, actbits = 0;
#pragma unroll
      for (int r = 0; r < NCP; r++) {
        const uint32_t cw = COOP_FAST ? CPI * r + l_cw : (32u * r + lane) / NCP;
        parbits |= ((pmask >> cw) & 1u) << r;
        actbits |= ((amask >> cw) & 1u) << r;
      }
      const uint32_t src_l = l_cw * wrow_d + 2 * l_pc;
      constexpr int NSW = COOP_FAST ? (8 / CPI > 0 ? 8 / CPI : 1) : 1;
      uint32_t dsw[NSW];
#pragma unroll
      for (int i = 0; i < NSW; i++) {
        const uint32_t cw = CPI * i + l_cw;
        dsw[i] = (pair * 32 + l_cw) * WROWP + 2 * ((l_pc & ~7u) | ((l_pc ^ cw) & 7u));
      }
      const double* wsrc = a.W[0] + warp_p0 * wrow_d;
      auto w_issue = [&](int gc) {
        if (gc < TW) {
          double* dst = wstage + (size_t)(gc % BB_WS_WST) * CH * WROWP;
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
                bb_cp_async16(dst + (pair * 32 + cw) * WROWP + 2 * ((piece & ~7u) | ((piece ^ cw) & 7u)),
                              src + (cw * wrow_d + ((pmask >> cw) & 1u) * (BB_TC * DP) + 2 * piece));
            }
          }
        }
        bb_cp_async_commit();
      };
      /* the proposal goes to the buffer the chain does NOT read from */
      double* ww = a.W[1 - par] + pc * (a.nbuf * BB_TC * DP);
      double w2[DP];
#pragma unroll 1
      for (int i = 0; i < BB_WS_PF; i++) w_issue(i);
      int g = 0;
      for (int s = 0; s < S; s++) {
        const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
        const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
#pragma unroll
        for (int k = 0; k < DP; k++) w2[k] = 0.0;
        for (int c = 0; c < NC; c++, g++) {
          tab_acquire(c);
          /* request the rows of chunk g+PF: their slot was last used by chunk g+PF-WST, which the dynamics warp must have
           * consumed and this warp must have written back (program order + the warp barrier) */
          __syncwarp();
          if (g + BB_WS_PF < TW && g + BB_WS_PF >= BB_WS_WST)
            bb_mbar_wait(&my_empty[(g + BB_WS_PF) % BB_WS_WST], (((g + BB_WS_PF) / BB_WS_WST) - 1) & 1);
          w_issue(g + BB_WS_PF);
          bb_cp_async_wait<BB_WS_PF>(); /* chunk g has landed (everything but the newest PF groups) */
          __syncwarp();
          tab_probe(c);
          double* wrow = wslot + (size_t)(g % BB_WS_WST) * CH * WROWP;
          if (c == 0) noise_chunk<true>(a, rec, w2, wrow, c, row_lo, row_hi);
          else noise_chunk<false>(a, rec, w2, wrow, c, row_lo, row_hi);
          __syncwarp();
          if (lane == 0) bb_mbar_arrive(&my_full[g % BB_WS_WST]); /* release: the dynamics warp may read the row */
          if (act) { /* the chain's row of W° is complete: write its 128 d' bytes back to back (whole lines) */
#pragma unroll
            for (int q = 0; q < NPIECE; q++) {
              double v[4];
              bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, v);
              bb_st4(ww + 4 * q, v[0], v[1], v[2], v[3]);
            }
          }
          ww += wstride;
          tab_release(c);
        }
      }
      return;
    }

    /* ==================================================================== DYNAMICS warp */
    typename Dyn::state st;
#pragma unroll
    for (int k = 0; k < D; k++) st.y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc];
    st.som = 0.0;
    double lltot = 0.0;
    const bool xact = act && SX;
    double* xw = SX ? a.X + pc * (BB_TC * D) : nullptr;
    double* xbuf = xbuf_all + (size_t)ci * 16;
    double wq[4] = {0.0, 0.0, 0.0, 0.0};
    int g = 0;
    for (int s = 0; s < S; s++) {
      const double* sc = a.segc[s];
      st.som = 0.0;
      for (int c = 0; c < NC; c++, g++) {
        tab_acquire(c);
        bb_mbar_wait(&my_full[g % BB_WS_WST], (g / BB_WS_WST) & 1); /* acquire: W° of chunk g is in the slot */
        tab_probe(c);
        double* wrow = wslot + (size_t)(g % BB_WS_WST) * CH * WROWP;
        const bool generic = (c == 0) || (c == NC - 1) || (c * BB_TC + BB_TC - 1 > a.jll);
        if (generic)
          Dyn::template chunk<true>(a, rec, sc, st, wq, wrow, nullptr, xw, xbuf, c, 0u, 0u, false, xact);
        else
          Dyn::template chunk<false>(a, rec, sc, st, wq, wrow, nullptr, xw, xbuf, c, 0u, 0u, false, xact);
        __syncwarp();
        if (lane == 0) bb_mbar_arrive(&my_empty[g % BB_WS_WST]);
        xw += xstride;
        tab_release(c);
      }
      lltot += st.som;
    }
    /* ---- per-chain epilogue: accept iff log(U) <= ll° - ll   (test/partialbridgenuH.jl:183) */
    const double logu = bb_accept_logu(a.keys, a.stream, chain);
    const double llc = a.ll[pc];
    const bool ok = act && (logu <= lltot - llc);
    if (act) {
      a.llprop[p] = lltot;
      a.logu[p] = logu;
      a.accepted[p] = ok ? 1 : 0;
      a.xstale[p] = SX ? (ok ? 0 : 1) : (uint8_t)(a.xstale[p] | (ok ? 1 : 0));
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
  }
};

template <class M, int GK, int GM, int AUXM, int RNG>
__global__ void __launch_bounds__(BB_THREADS, bb_min_ctas<M>()) bb_chain_ws_kernel(const __grid_constant__ bb_chain_args a) {
  bb_chain_ws<M, GK, GM, AUXM, RNG>::run(a);
}

/* ------------------------------------------------------------------------------------------------------------
 * bb_chain_ws2: the same split with TWO chains per dynamics thread.
 *
 * What limits the split above on a small ensemble is the dynamics warp: one dependent chain of fp64 operations per
 * thread, which leaves most issue slots of its scheduler empty (experiment: noise warps without RNG work 0.82 ms,
 * dynamics warps idle 0.44 ms, at 31 250 chains).  Here a CTA is 3 warps for 64 chains -- two noise warps (32 chains
 * each, exactly the noise branch above) feeding ONE dynamics warp whose lane l steps chains l and 32 + l side by side on
 * the same table rows: two independent dependency chains per thread, half the dynamics warps.  Arithmetic per chain is
 * unchanged (bb_chain::step), results are bit-identical.
 */
template <class M, int GK, int GM, int AUXM>
struct bb_chain_ws2 {
  using Dyn = bb_chain<M, GK, GM, AUXM, 0>;
  static constexpr int D = M::D, DP = M::DP, REC = Dyn::REC, XWIN = Dyn::XWIN;
  static_assert(DP == 1 && D <= 2, "scalar noise, X° through the 128-byte shared-memory window");
  static constexpr int WROWP = BB_TC * DP, NPIECE = BB_TC * DP / 4;
  static constexpr int NT = 96, CH = 64;
  static constexpr int BAR_BYTES = ((2 * BB_STAGES + 2 * 2 * BB_WS_WST) * 8 + 127) / 128 * 128;
  static __host__ __device__ constexpr size_t smem_bytes() {
    return (size_t)BB_STAGES * BB_TSTAGE * BB_TC * REC * 8 + BAR_BYTES + (size_t)BB_WS_WST * CH * WROWP * 8 +
           (size_t)CH * 128;
  }

  template <bool GENERIC>
  static __device__ __forceinline__ void dyn_chunk2(const bb_chain_args& a, const double* __restrict__ rec,
                                                    const double* __restrict__ sc, typename Dyn::state& s0,
                                                    typename Dyn::state& s1, const double* wrow0, const double* wrow1,
                                                    double* xout0, double* xout1, double* xbuf0, double* xbuf1, int c,
                                                    bool xact0, bool xact1) {
    const int N = a.N;
#pragma unroll 1
    for (int h = 0; h < BB_TC / 4; h++) {
      double w0[4], w1[4];
      bb_lds4_swz(wrow0, 2 * h, threadIdx.x & 7, w0);
      bb_lds4_swz(wrow1, 2 * h, threadIdx.x & 7, w1);
#pragma unroll
      for (int s4 = 0; s4 < 4; s4++) {
        const int slot = 4 * h + s4;
        const int j = c * BB_TC + slot;
        const double* R = rec + slot * REC;
        if (GENERIC && j == 0) {
          s0.wprev[0] = w0[s4];
          s1.wprev[0] = w1[s4];
        } else if (!GENERIC || j < N) {
          const bool in_ll = !GENERIC || j <= a.jll;
          Dyn::step(a, R, sc, s0, &w0[s4], in_ll);
          Dyn::step(a, R, sc, s1, &w1[s4], in_ll);
          if (GENERIC && GK == BB_GUIDE_HV && j == N - 1 && sc[D * D + D] != 0.0) {
#pragma unroll
            for (int k = 0; k < D; k++) s0.y[k] = s1.y[k] = sc[D * D + D + 1 + k];
          }
        }
        const int ls = slot % XWIN;
        if constexpr (D == 2) {
          *reinterpret_cast<double2*>(xbuf0 + 2 * ((ls ^ threadIdx.x) & 7)) = make_double2(s0.y[0], s0.y[1]);
          *reinterpret_cast<double2*>(xbuf1 + 2 * ((ls ^ threadIdx.x) & 7)) = make_double2(s1.y[0], s1.y[1]);
        } else {
          xbuf0[2 * (((ls >> 1) ^ threadIdx.x) & 7) + (ls & 1)] = s0.y[0];
          xbuf1[2 * (((ls >> 1) ^ threadIdx.x) & 7) + (ls & 1)] = s1.y[0];
        }
      }
      if (((4 * h + 3) % XWIN) == XWIN - 1) { /* a 128-byte window of X° is complete: write it back to back */
        const int off = (4 * h + 4 - XWIN) * D;
#pragma unroll
        for (int q = 0; q < 4; q++) {
          double v[4];
          if (xact0) {
            bb_lds4_swz(xbuf0, 2 * q, threadIdx.x & 7, v);
            bb_st4(xout0 + off + 4 * q, v[0], v[1], v[2], v[3]);
          }
          if (xact1) {
            bb_lds4_swz(xbuf1, 2 * q, threadIdx.x & 7, v);
            bb_st4(xout1 + off + 4 * q, v[0], v[1], v[2], v[3]);
          }
        }
      }
    }
  }

  static __device__ __forceinline__ void run(const bb_chain_args& a) {
    using WS = bb_chain_ws<M, GK, GM, AUXM, 1>;
    constexpr uint32_t CHUNK_DOUBLES = BB_TC * REC;
    constexpr uint32_t STAGE_DOUBLES = BB_TSTAGE * CHUNK_DOUBLES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + BB_STAGES * STAGE_DOUBLES);
    uint64_t* empty = full + BB_STAGES;
    uint64_t* wfull = empty + BB_STAGES; /* [noise warp][slot] */
    uint64_t* wempty = wfull + 2 * BB_WS_WST;
    double* wstage = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(full) + BAR_BYTES);
    double* xbuf_all = wstage + (size_t)BB_WS_WST * CH * WROWP;

    const int S = a.S, NC = a.NC;
    const int NST = (NC + BB_TSTAGE - 1) / BB_TSTAGE;
    const int T = S * NST;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool noise = warp >= 1;
    const int nw = warp - 1; /* noise warp 0 / 1 serves chains [0, 32) / [32, 64) of the CTA */

    if (threadIdx.x == 0) {
      for (int i = 0; i < BB_STAGES; i++) {
        bb_mbar_init(&full[i], 1);
        bb_mbar_init(&empty[i], 3);
      }
      for (int i = 0; i < 2 * BB_WS_WST; i++) {
        bb_mbar_init(&wfull[i], 1);
        bb_mbar_init(&wempty[i], 1);
      }
      bb_mbar_fence_init();
    }
    __syncthreads();

    const long long P = a.P;
    const long long cta_p0 = a.p_begin + (long long)blockIdx.x * CH; /* always full CTAs (a.cpc is not used) */
    const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);
    const int TW = S * NC;

    int issued = 0, iseg = 0, ist_in_seg = 0;
    int stage = 0, gs = 0;
    bool tab_ready = false;
    uint32_t phase = 0;
    const double* rec = ring;
    auto tab_acquire = [&](int c) {
      if ((c % BB_TSTAGE) == 0) {
        if (threadIdx.x == 0) {
          while (issued < T && issued <= gs + BB_LOOKAHEAD) {
            const int ist = issued % BB_STAGES;
            if (issued >= BB_STAGES) bb_mbar_wait(&empty[ist], ((issued / BB_STAGES) - 1) & 1);
            const int c0 = ist_in_seg * BB_TSTAGE;
            const int nch = (NC - c0 < BB_TSTAGE) ? NC - c0 : BB_TSTAGE;
            const uint32_t bytes = (uint32_t)nch * CHUNK_DOUBLES * 8;
            bb_mbar_expect_tx(&full[ist], bytes);
            bb_tma_load_1d(ring + ist * STAGE_DOUBLES, a.tab[iseg] + (size_t)c0 * CHUNK_DOUBLES, bytes, &full[ist]);
            issued++;
            if (++ist_in_seg == NST) { ist_in_seg = 0; iseg++; }
          }
        }
        __syncwarp();
        if (!tab_ready) bb_mbar_wait(&full[stage], phase);
        rec = ring + stage * STAGE_DOUBLES;
      }
    };
    auto tab_probe = [&](int c) {
      if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
        const int nst = (stage + 1 == BB_STAGES) ? 0 : stage + 1;
        const uint32_t nph = (stage + 1 == BB_STAGES) ? phase ^ 1 : phase;
        tab_ready = (gs + 1 < T) ? bb_mbar_test(&full[nst], nph) : true;
      }
    };
    auto tab_release = [&](int c) {
      rec += CHUNK_DOUBLES;
      if ((c % BB_TSTAGE) == BB_TSTAGE - 1 || c == NC - 1) {
        __syncwarp();
        if (lane == 0) bb_mbar_arrive(&empty[stage]);
        if (++stage == BB_STAGES) { stage = 0; phase ^= 1; }
        gs++;
      }
    };

    if (noise) {
      /* ================================================================== NOISE warp (as in bb_chain_ws) */
      const int ci = nw * 32 + lane;
      const long long p = cta_p0 + ci;
      const long long pc = p < a.p_end ? p : a.p_end - 1;
      const bool act = p < a.p_end && (!a.only || a.only[pc] != 0);
      const int par = a.par[pc];
      const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
      uint64_t* my_full = wfull + nw * BB_WS_WST;
      uint64_t* my_empty = wempty + nw * BB_WS_WST;
      double* wslot = wstage + (size_t)ci * WROWP;
      constexpr int NCP = BB_TC * DP / 2;
      static_assert(32 % NCP == 0, "one cp.async instruction covers whole chains");
      constexpr int CPI = 32 / NCP;
      const unsigned amask = __ballot_sync(0xFFFFFFFFu, act);
      const unsigned pmask = __ballot_sync(0xFFFFFFFFu, par != 0);
      const long long warp_p0 = cta_p0 + nw * 32;
      const uint32_t wrow_d = (uint32_t)(a.nbuf * BB_TC * DP);
      const uint32_t l_cw = (uint32_t)lane / NCP, l_pc = (uint32_t)lane % NCP;
      uint32_t parbits = 0, actbits = 0;
#pragma unroll
      for (int r = 0; r < NCP; r++) {
        const uint32_t cw = CPI * r + l_cw;
        parbits |= ((pmask >> cw) & 1u) << r;
        actbits |= ((amask >> cw) & 1u) << r;
      }
      const uint32_t src_l = l_cw * wrow_d + 2 * l_pc;
      constexpr int NSW = 8 / CPI > 0 ? 8 / CPI : 1;
      uint32_t dsw[NSW];
#pragma unroll
      for (int i = 0; i < NSW; i++) {
        const uint32_t cw = CPI * i + l_cw;
        dsw[i] = (nw * 32 + l_cw) * WROWP + 2 * ((l_pc & ~7u) | ((l_pc ^ cw) & 7u));
      }
      const double* wsrc = a.W[0] + warp_p0 * wrow_d;
      auto w_issue = [&](int gc) {
        if (gc < TW) {
          double* dst = wstage + (size_t)(gc % BB_WS_WST) * CH * WROWP;
          const double* src = wsrc + (long long)gc * wstride + src_l;
#pragma unroll
          for (int r = 0; r < NCP; r++) {
            const uint32_t poff = ((parbits >> r) & 1u) * (uint32_t)(BB_TC * DP);
            if ((actbits >> r) & 1u)
              bb_cp_async16(dst + dsw[r % NSW] + r * (CPI * WROWP), src + (r * CPI * wrow_d + poff));
          }
        }
        bb_cp_async_commit();
      };
      double* ww = a.W[1 - par] + pc * (a.nbuf * BB_TC * DP);
      double w2[DP];
#pragma unroll 1
      for (int i = 0; i < BB_WS_PF; i++) w_issue(i);
      int g = 0;
      for (int s = 0; s < S; s++) {
        const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
        const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
#pragma unroll
        for (int k = 0; k < DP; k++) w2[k] = 0.0;
        for (int c = 0; c < NC; c++, g++) {
          tab_acquire(c);
          __syncwarp();
          if (g + BB_WS_PF < TW && g + BB_WS_PF >= BB_WS_WST)
            bb_mbar_wait(&my_empty[(g + BB_WS_PF) % BB_WS_WST], (((g + BB_WS_PF) / BB_WS_WST) - 1) & 1);
          w_issue(g + BB_WS_PF);
          bb_cp_async_wait<BB_WS_PF>();
          __syncwarp();
          tab_probe(c);
          double* wrow = wslot + (size_t)(g % BB_WS_WST) * CH * WROWP;
          if (c == 0) WS::template noise_chunk<true>(a, rec, w2, wrow, c, row_lo, row_hi);
          else WS::template noise_chunk<false>(a, rec, w2, wrow, c, row_lo, row_hi);
          __syncwarp();
          if (lane == 0) bb_mbar_arrive(&my_full[g % BB_WS_WST]);
          if (act) {
#pragma unroll
            for (int q = 0; q < NPIECE; q++) {
              double v[4];
              bb_lds4_swz(wrow, 2 * q, threadIdx.x & 7, v);
              bb_st4(ww + 4 * q, v[0], v[1], v[2], v[3]);
            }
          }
          ww += wstride;
          tab_release(c);
        }
      }
      return;
    }

    /* ==================================================================== DYNAMICS warp: chains lane and 32 + lane */
    const long long p0 = cta_p0 + lane, p1 = cta_p0 + 32 + lane;
    const long long pc0 = p0 < a.p_end ? p0 : a.p_end - 1, pc1 = p1 < a.p_end ? p1 : a.p_end - 1;
    const bool act0 = p0 < a.p_end && (!a.only || a.only[pc0] != 0), act1 = p1 < a.p_end && (!a.only || a.only[pc1] != 0);
    typename Dyn::state st0, st1;
#pragma unroll
    for (int k = 0; k < D; k++) {
      st0.y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc0];
      st1.y[k] = a.start_bcast ? a.start[k] : a.start[(long long)k * P + pc1];
    }
    st0.som = st1.som = 0.0;
    double ll0 = 0.0, ll1 = 0.0;
    double* xw0 = a.X + pc0 * (BB_TC * D);
    double* xw1 = a.X + pc1 * (BB_TC * D);
    double* xbuf0 = xbuf_all + (size_t)lane * 16;
    double* xbuf1 = xbuf_all + (size_t)(32 + lane) * 16;
    const double* wslot0 = wstage + (size_t)lane * WROWP;
    const double* wslot1 = wstage + (size_t)(32 + lane) * WROWP;
    int g = 0;
    for (int s = 0; s < S; s++) {
      const double* sc = a.segc[s];
      st0.som = st1.som = 0.0;
      for (int c = 0; c < NC; c++, g++) {
        tab_acquire(c);
        const int slot = g % BB_WS_WST;
        const uint32_t wph = (g / BB_WS_WST) & 1;
        bb_mbar_wait(&wfull[slot], wph);              /* acquire: W° of chunk g, chains [0, 32) */
        bb_mbar_wait(&wfull[BB_WS_WST + slot], wph);  /* ... and chains [32, 64) */
        tab_probe(c);
        const double* wrow0 = wslot0 + (size_t)slot * CH * WROWP;
        const double* wrow1 = wslot1 + (size_t)slot * CH * WROWP;
        const bool generic = (c == 0) || (c == NC - 1) || (c * BB_TC + BB_TC - 1 > a.jll);
        if (generic) dyn_chunk2<true>(a, rec, sc, st0, st1, wrow0, wrow1, xw0, xw1, xbuf0, xbuf1, c, act0, act1);
        else dyn_chunk2<false>(a, rec, sc, st0, st1, wrow0, wrow1, xw0, xw1, xbuf0, xbuf1, c, act0, act1);
        __syncwarp();
        if (lane == 0) {
          bb_mbar_arrive(&wempty[slot]);
          bb_mbar_arrive(&wempty[BB_WS_WST + slot]);
        }
        xw0 += xstride;
        xw1 += xstride;
        tab_release(c);
      }
      ll0 += st0.som;
      ll1 += st1.som;
    }
    /* ---- per-chain epilogue: accept iff log(U) <= ll° - ll   (test/partialbridgenuH.jl:183) */
    unsigned nacc = 0;
#pragma unroll
    for (int w = 0; w < 2; w++) {
      const long long p = w ? p1 : p0, pc = w ? pc1 : pc0;
      const bool act = w ? act1 : act0;
      const double lltot = w ? ll1 : ll0;
      const typename Dyn::state& st = w ? st1 : st0;
      const unsigned long long chain = (unsigned long long)(a.chain_offset + pc);
      const double logu = bb_accept_logu(a.keys, a.stream, chain);
      const double llc = a.ll[pc];
      const int par = a.par[pc];
      const bool ok = act && (logu <= lltot - llc);
      if (act) {
        a.llprop[p] = lltot;
        a.logu[p] = logu;
        a.accepted[p] = ok ? 1 : 0;
        a.xstale[p] = ok ? 0 : 1;
#pragma unroll
        for (int k = 0; k < D; k++) a.xendprop[(long long)k * P + p] = st.y[k];
        if (ok) {
          a.ll[p] = lltot;
          a.par[p] = (uint8_t)(1 - par);
#pragma unroll
          for (int k = 0; k < D; k++) a.xend[(long long)k * P + p] = st.y[k];
        }
      }
      nacc += __popc(__ballot_sync(0xFFFFFFFFu, ok));
    }
    if (lane == 0 && nacc) atomicAdd(a.acc, (unsigned long long)nacc);
  }
};

template <class M, int GK, int GM, int AUXM>
__global__ void __launch_bounds__(96, 4) bb_chain_ws2_kernel(const __grid_constant__ bb_chain_args a) {
  bb_chain_ws2<M, GK, GM, AUXM>::run(a);
}

template <class M, int GK, int GM, int AUXM>
static cudaError_t bb_chain_ws2_launch(const bb_chain_args& a, cudaStream_t st) {
  using K = bb_chain_ws2<M, GK, GM, AUXM>;
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((attr_done.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(bb_chain_ws2_kernel<M, GK, GM, AUXM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)K::smem_bytes());
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_release);
  }
  const long long n = a.p_end - a.p_begin;
  bb_chain_ws2_kernel<M, GK, GM, AUXM><<<(unsigned)((n + K::CH - 1) / K::CH), K::NT, K::smem_bytes(), st>>>(a);
  return cudaGetLastError();
}

/* CTA size: 256 threads (128 chains) when that still gives every SM two CTAs, else 128 threads (64 chains) so that a
 * small ensemble -- the strong-scaling share of a GPU -- spreads over all SMs */
template <class M, int GK, int GM, int AUXM, int RNG>
static cudaError_t bb_chain_ws_launch(const bb_chain_args& a, cudaStream_t st) {
  using K = bb_chain_ws<M, GK, GM, AUXM, RNG>;
  static std::atomic<unsigned long long> attr_done{0};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 64 || !((attr_done.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(bb_chain_ws_kernel<M, GK, GM, AUXM, RNG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::smem_bytes(BB_THREADS));
    if (e != cudaSuccess) return e;
    if (dev < 64) attr_done.fetch_or(1ull << dev, std::memory_order_release);
  }
  const long long n = a.p_end - a.p_begin;
  const int nt = (n + 127) / 128 >= 2ll * sms ? 256 : 128;
  bb_chain_args b = a;
  b.cpc = bb_pick_cpc(n, nt / 2, nt == 256 ? 2 : 4); /* resident CTAs per SM: registers (128 per thread) and shared memory */
  const unsigned grid = (unsigned)((n + b.cpc - 1) / b.cpc);
  bb_chain_ws_kernel<M, GK, GM, AUXM, RNG><<<grid, nt, K::smem_bytes(nt), st>>>(b);
  return cudaGetLastError();
}
