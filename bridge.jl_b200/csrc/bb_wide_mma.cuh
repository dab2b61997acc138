/*
 * bb_wide_mma.cuh -- Landmarks (d = 16, d' = 8) guided path kernel with the two 16 x 16 mat-vecs on the fp64 TENSOR
 * cores: BASELINE config 5 ("sigma dW / guiding term as a tensor-core tile").
 *
 * A guided Euler step of this model needs r = H[i] (nu[i] - x) and b~ = B~ x + beta~ -- two 16 x 16 matrix-vector
 * products whose MATRICES are the same for every chain.  Eight chains of a warp make it a matrix-matrix product:
 *     R^T (8 chains x 16) = E^T (8 x 16) . H^T (16 x 16)
 * = 2 column tiles x 4 k-tiles of mma.sync.aligned.m8n8k4.row.col.f64 (DMMA; tcgen05 has no f64 kind).  Thread
 * T = 4 n + g of a warp (chain n of the warp, landmark g, as in bb_wide4_kernel) supplies A[n][k = g] = e_n[4 g + kt] --
 * a component it owns -- and B[k = g][col n] = H[rowsel][4 g + kt]; with the row permutation
 *     column 8 nt + 2 gg + b  <->  matrix row 4 gg + 2 nt + b
 * the accumulator fragment of thread (n, g) is exactly r_n[4 g .. 4 g + 3], the rows of its own landmark: no shuffle, no
 * shared memory.  16 DMMA per warp and step replace 4 lanes x 128 dependent DFMA (8 chains x 2 x 256 multiply-adds either
 * way; what changes is the instruction count -- 880 -> ~400 per warp and step -- and the length of the dependent chains).
 * B~ is constant over a segment: its eight fragments live in registers.  H[i] changes every step: every warp streams the
 * table rows through its own shared-memory ring (16-byte cp.async, three steps ahead) and a lane reads its eight fragment
 * elements with four 128-bit loads, as it does nu[4 g .. 4 g + 3], dt, sqrt(dt).
 *
 * Lane g owns landmark g throughout: it evaluates the pairwise drift of its landmark (the other three landmarks' (q, p)
 * arrive by 12 shuffles; no lane ever holds the whole state), regrouped as dq = S/2, dp = -lambda/2 S + sum_j w_j (q_g - q_j)
 * with S = sum_j p_j k_gj, its two noise columns (one Philox call, ONE Box-Muller pair) and its Euler update.  The
 * Girsanov sum <b - b~, r> is reduced by a butterfly over the four lanes.
 *
 * Rounding: a DMMA accumulates its four products in its own order and the butterfly sums in tree order, so results
 * differ from the oracle's sequential sums at rounding level: this kernel is compared with the reference-arithmetic
 * oracle at the contract tolerance (|dX| <= 1e-9 (1+|X|), |dll| <= 1e-6 |ll|), not bit for bit; bb_wide4_kernel
 * (BB_WIDE_MMA=0 or bb_ctx_set_wide_kernel) remains the bit-exact cross-check.
 */
#pragma once
#include <type_traits>

#include "bb_wide.cuh"

__device__ __forceinline__ void bb_dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
/* table values: read-only for the whole launch and shared by all warps of the SM -> the non-coherent path (LDG.CONSTANT,
 * cached in L1).  A plain ld.global.ca inside `asm volatile` becomes LDG.STRONG.SM, which waited behind the thread's
 * own outstanding stores (X°, W°): one full step of latency (ncu: 32 % of all stall samples on the first use). */
__device__ __forceinline__ void bb_ldg2_ca(const double* p, double& v0, double& v1) {
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "l"(p));
}
__device__ __forceinline__ void bb_lds2(const double* p, double& v0, double& v1) { /* LDS.128 */
  const double2 v = *reinterpret_cast<const double2*>(p);
  v0 = v.x; v1 = v.y;
}
__device__ __forceinline__ void bb_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* the two normals of columns 2g, 2g+1 at one grid point: words (0,1) or (2,3) of quad q (bb_normal_quad's layout) */
__device__ __forceinline__ void bb_normal_pair(const bb_philox_keys& k, uint32_t stream, uint32_t row_lo, uint32_t row_hi,
                                               uint32_t q, int half, float& z0, float& z1) {
  uint32_t o[4];
  bb_philox4x32_10(q, stream, row_lo, row_hi, k, o);
  bb_box_muller(half ? o[2] : o[0], half ? o[3] : o[1], z0, z1);
}

template <int RNG>
__global__ void __launch_bounds__(BB_W4_THREADS, 3) bb_wide4m_kernel(const __grid_constant__ bb_chain_args a) {
  using M = MLandmarks;
  using CH = bb_chain<M, BB_GUIDE_NUH, 0, 1, 0>;
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

  /* B operand of this lane: matrix row rowsel(nt) = 4 (n >> 1) + 2 nt + (n & 1), columns 4 g + kt   (n = lane >> 2).
   * The order of the 16 products inside the contraction is free as long as A and B agree: k-tile kt pairs the lane's OWN
   * component kt, e[4 g + kt], with column 4 g + kt -- the lane needs no other lane's state for the A operand, and its
   * four B elements are contiguous in the table (two 128-bit loads per column tile) */
  const int ncol = lane >> 2;
  const int frag0 = (4 * (ncol >> 1) + (ncol & 1)) * D + 4 * g; /* + 2 nt D + kt */

  const int par = a.par[pc];
  const int wbuf = PCN ? 1 - par : par;
  const double* wr = a.W[par] + pc * (a.nbuf * BB_TC * DP) + 2 * g;
  double* ww = a.W[wbuf] + pc * (a.nbuf * BB_TC * DP) + 2 * g;
  double* xw = sx ? a.X + pc * (BB_TC * D) + 4 * g : nullptr;
  const long long wstride = P * (a.nbuf * BB_TC * DP), xstride = P * (BB_TC * D);

  double yo[4], wprev[2], w2[2]; /* the own landmark's (q, p): components 4g .. 4g+3 of the state */
#pragma unroll
  for (int k = 0; k < 4; k++) yo[k] = a.start_bcast ? a.start[4 * g + k] : a.start[(long long)(4 * g + k) * P + pc];
  double lltot = 0.0;
  double wnx[2] = {0.0, 0.0};
  if (act) bb_ld2(wr, wnx);

  /* every warp streams the table rows (dt, sqrt dt, nu, H: 2.2 KB per step, the same for all chains) through its own
   * BB_WIDE_RING-deep shared-memory ring with 16-byte cp.async, BB_WIDE_RING-1 steps ahead of their use (read straight
   * from global memory, one step ahead, the first use of a row stalled for a third of the step: ncu long_scoreboard).
   * Matrix row r of H is shifted by 2 r doubles, which makes the 128-bit fragment reads of a quarter-warp (two chains x
   * four landmarks) conflict free. */
  constexpr int PIECES = REC / 2, RECS = REC + 2 * D, NWARP = BB_W4_THREADS / 32;
  __shared__ __align__(16) double ring_all[NWARP * BB_WIDE_RING * RECS];
  double* ring = ring_all + (threadIdx.x >> 5) * (BB_WIDE_RING * RECS);
  const int rows_per_seg = NC * BB_TC;
  const int total_rows = S * rows_per_seg;
  int issue_row = 0, cons_row = 0, issue_seg = 0, issue_in_seg = 0;
  /* piece q = lane + 32 i of a row goes to a fixed place: offsets computed once */
  constexpr int NISS = (PIECES + 31) / 32;
  int dsto[NISS];
#pragma unroll
  for (int i = 0; i < NISS; i++) {
    const int q = lane + 32 * i, e = 2 * q - OFF_A2; /* e: element of H (row-major d x d) if >= 0 */
    dsto[i] = 2 * q + (e >= 0 ? 2 * (e >> 4) : 0);
  }
  auto row_issue = [&]() {
    if (issue_row < total_rows) {
      const double* src = a.tab[issue_seg] + (size_t)issue_in_seg * REC + 2 * lane;
      double* dst = ring + (issue_row & (BB_WIDE_RING - 1)) * RECS;
#pragma unroll
      for (int i = 0; i < NISS; i++)
        if (i < NISS - 1 || lane + 32 * i < PIECES) bb_cp_async16(dst + dsto[i], src + 64 * i);
      if (++issue_in_seg == rows_per_seg) { issue_in_seg = 0; issue_seg++; }
    }
    issue_row++;
    bb_cp_async_commit();
  };
#pragma unroll 1
  for (int r = 0; r < BB_WIDE_RING - 1; r++) row_issue();
  const int fragS = (4 * (ncol >> 1) + (ncol & 1)) * (D + 2) + 4 * g; /* the lane's H fragment in a staged row (+ 2 nt (D+2)) */

  for (int s = 0; s < S; s++) {
    const unsigned long long row = chain * (unsigned long long)S + (unsigned long long)s;
    const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)(row >> 32);
    const double* tab = a.tab[s];
    /* B~ fragments and the own rows of beta~ (they follow the rows of the segment's table) */
    const double* aux = tab + (size_t)NC * BB_TC * REC;
    double btf[2][4], beo[4];
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      bb_ldg2_ca(aux + frag0 + 2 * nt * D, btf[nt][0], btf[nt][1]);
      bb_ldg2_ca(aux + frag0 + 2 * nt * D + 2, btf[nt][2], btf[nt][3]);
    }
    bb_ldg2_ca(aux + D * D + 4 * g, beo[0], beo[1]);
    bb_ldg2_ca(aux + D * D + 4 * g + 2, beo[2], beo[3]);
    double som = 0.0;
    w2[0] = w2[1] = 0.0;
    for (int c = 0; c < NC; c++) {
      if (act) { /* the chain's rows of W (1 KB per chunk) go into L2 BB_W4M_AHEAD chunks ahead (DRAM latency is longer than a step) */
#ifndef BB_W4M_AHEAD
#define BB_W4M_AHEAD 2
#endif
        const int c0 = (c == 0) ? 1 : BB_W4M_AHEAD;
        for (int ca = c0; ca <= BB_W4M_AHEAD; ca++)
          if (c + ca < NC) {
            const char* nx = reinterpret_cast<const char*>(wr - 2 * g + ca * wstride) + 256 * g;
            bb_prefetch_l2(nx); bb_prefetch_l2(nx + 128);
          }
      }
      /* GEN = false is the steady state: every slot is a full step that enters the log-likelihood and is followed by
       * another row; GEN = true also handles j = 0 (no step), j >= N (padding), `skip` and the very last row */
      auto chunk = [&](auto gen_tag) {
      constexpr bool GEN = decltype(gen_tag)::value;
#pragma unroll 1
      for (int slot = 0; slot < BB_TC; slot++) {
        const int j = c * BB_TC + slot;
        /* this step's table row has landed in the warp's ring; the slot read one step ago is requested again */
        bb_cp_async_wait<BB_WIDE_RING - 2>();
        __syncwarp();
        const double* R = ring + (cons_row & (BB_WIDE_RING - 1)) * RECS;
        cons_row++;
        row_issue();
        double dt, rootdt;
        bb_lds2(R, dt, rootdt);
        /* ---- own two noise columns 2g, 2g+1 */
        double wj[2] = {wnx[0], wnx[1]};
        {
          const bool lastrow = GEN && (s == S - 1) && (c == NC - 1) && (slot == BB_TC - 1);
          const double* nxt = (slot == BB_TC - 1) ? wr + wstride : wr + (slot + 1) * DP;
          if (act && !lastrow) bb_ld2(nxt, wnx);
        }
        if constexpr (PCN) {
          float z0, z1;
          bb_normal_pair(a.keys, a.stream, row_lo, row_hi, (uint32_t)(NPIECE * c + 2 * slot + (g >> 1)), g & 1, z0, z1);
          if (!GEN || j != 0) {
            w2[0] = fma(rootdt, (double)z0, w2[0]);
            w2[1] = fma(rootdt, (double)z1, w2[1]);
          }
          wj[0] = fma(a.rho2, w2[0], a.rho * wj[0]);
          wj[1] = fma(a.rho2, w2[1], a.rho * wj[1]);
          if (act) bb_st2(ww + slot * DP, wj[0], wj[1]);
        }
        if (GEN && j == 0) {
          wprev[0] = wj[0]; wprev[1] = wj[1];
        } else if (!GEN || j < N) {
          const double dw0 = wj[0] - wprev[0], dw1 = wj[1] - wprev[1];
          wprev[0] = wj[0]; wprev[1] = wj[1];
          /* ---- r = H (nu - x) and B~ x on the tensor cores: rows 4g .. 4g+3 of both land in this lane */
          double rg[4] = {0.0, 0.0, 0.0, 0.0}, bt[4] = {0.0, 0.0, 0.0, 0.0};
          {
            double nug[4], hf[2][4];
            bb_lds2(R + OFF_C + 4 * g, nug[0], nug[1]);
            bb_lds2(R + OFF_C + 4 * g + 2, nug[2], nug[3]);
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
              bb_lds2(R + OFF_A2 + fragS + 2 * nt * (D + 2), hf[nt][0], hf[nt][1]);
              bb_lds2(R + OFF_A2 + fragS + 2 * nt * (D + 2) + 2, hf[nt][2], hf[nt][3]);
            }
#pragma unroll
            for (int kt = 0; kt < 4; kt++) {
              const double yk = yo[kt];
              const double ek = nug[kt] - yk;
              bb_dmma884(rg[0], rg[1], ek, hf[0][kt], rg[0], rg[1]);
              bb_dmma884(rg[2], rg[3], ek, hf[1][kt], rg[2], rg[3]);
              bb_dmma884(bt[0], bt[1], yk, btf[0][kt], bt[0], bt[1]);
              bb_dmma884(bt[2], bt[3], yk, btf[1][kt], bt[2], bt[3]);
            }
          }
          /* the other three landmarks' (q, p) come from the lanes g+1, g+2, g+3 of the chain's group by shuffles: no lane
           * ever holds the whole state */
          double qx[3], qy[3], px[3], py[3];
#pragma unroll
          for (int rel = 1; rel < NL; rel++) {
            qx[rel - 1] = __shfl_sync(0xFFFFFFFFu, yo[0], (g + rel) & 3, 4);
            qy[rel - 1] = __shfl_sync(0xFFFFFFFFu, yo[1], (g + rel) & 3, 4);
            px[rel - 1] = __shfl_sync(0xFFFFFFFFu, yo[2], (g + rel) & 3, 4);
            py[rel - 1] = __shfl_sync(0xFFFFFFFFu, yo[3], (g + rel) & 3, 4);
          }
          double kr[3];
          {
            const double dxa = yo[0] - qx[0], dya = yo[1] - qy[0], dxb = yo[0] - qx[1], dyb = yo[1] - qy[1];
            kr[0] = c0 * bb_exp(-(fma(dya, dya, dxa * dxa) * c1));
            kr[1] = c0 * bb_exp(-(fma(dyb, dyb, dxb * dxb) * c1));
            kr[2] = __shfl_sync(0xFFFFFFFFu, kr[0], (g + 3) & 3, 4); /* k(q_g - q_{g+3}) = k(q_{g+3} - q_{(g+3)+1}) */
          }
          double S0 = c0 * yo[2], S1 = c0 * yo[3], T0 = 0.0, T1 = 0.0;
#pragma unroll
          for (int r = 0; r < 3; r++) {
            S0 = fma(px[r], kr[r], S0);
            S1 = fma(py[r], kr[r], S1);
            const double w = (c1 * fma(yo[3], py[r], yo[2] * px[r])) * kr[r];
            T0 = fma(w, yo[0] - qx[r], T0);
            T1 = fma(w, yo[1] - qy[r], T1);
          }
          double bg[4];
          bg[0] = 0.5 * S0;
          bg[1] = 0.5 * S1;
          bg[2] = fma(nlh, S0, T0);
          bg[3] = fma(nlh, S1, T1);
          if (!GEN || j <= a.jll) { /* <b - b~, r> dt: own four terms, then a butterfly over the four lanes of the chain */
            double part = (bg[0] - (bt[0] + beo[0])) * rg[0];
            part = fma(bg[1] - (bt[1] + beo[1]), rg[1], part);
            part = fma(bg[2] - (bt[2] + beo[2]), rg[2], part);
            part = fma(bg[3] - (bt[3] + beo[3]), rg[3], part);
            part += __shfl_xor_sync(0xFFFFFFFFu, part, 1, 4);
            part += __shfl_xor_sync(0xFFFFFFFFu, part, 2, 4);
            som = fma(part, dt, som);
          }
          /* _b = b + a r on the momentum rows; Euler-Maruyama update of the own components (bb_em_update) */
          bg[2] = fma(akk, rg[2], bg[2]);
          bg[3] = fma(akk, rg[3], bg[3]);
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
        }
        if (sx && act) bb_st4(xw + slot * D, yo[0], yo[1], yo[2], yo[3]);
      }
      };
      if (c == 0 || c == NC - 1 || c * BB_TC + BB_TC - 1 > a.jll) chunk(std::true_type{});
      else chunk(std::false_type{});
      wr += wstride;
      ww += wstride;
      if (sx) xw += xstride;
    }
    lltot += som;
  }

  /* per-chain epilogue: lane 0 of the group writes (as bb_wide4_kernel) */
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
    __syncwarp();
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

template <int RNG>
static cudaError_t bb_wide4m_launch(const bb_chain_args& a, cudaStream_t st) {
  const long long threads = 4 * (a.p_end - a.p_begin);
  const unsigned grid = (unsigned)((threads + BB_W4_THREADS - 1) / BB_W4_THREADS);
  bb_wide4m_kernel<RNG><<<grid, BB_W4_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}
static bb_chain_launch_fn bb_lookup_landmarks4m(int gk, int auxm, int rng) {
  if (gk == BB_GUIDE_NUH && auxm == 1) {
    if (rng == 0) return &bb_wide4m_launch<0>;
    if (rng == 1) return &bb_wide4m_launch<1>;
    if (rng == 3) return &bb_wide4m_launch<3>;
  }
  return nullptr;
}
