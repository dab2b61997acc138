/*
 * bb_backward.cuh -- device building blocks of the proposal constructors (backward ODEs), shared by the
 * one-system kernels of bb_backward.cu and the per-chain kernels of bb_theta.cu:
 *   small dense algebra in the oracle's operation order, kernelr3 (src/ode.jl:44-49), the right-hand sides of
 *   src/partialbridgenuH.jl, src/gode.jl, src/partialbridge.jl, one backward step of partialbridgeodeνH!
 *   (R3: src/partialbridgenuH.jl:21-55; Lyap: :86-103, src/lyap.jl:2-6) and the observation update
 *   (src/guip.jl:221-243, partialbridge_bolus3.jl:128-137).
 */
#pragma once
#include <math.h>

#include "../../include/bridge_b200.h"

namespace bbk {

template <int n, int k, int m>
__device__ __forceinline__ void mmul(const double* A, const double* B, double* C) {
  double T[n * m];
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < m; j++) {
      double s = A[i * k] * B[j];
#pragma unroll
      for (int l = 1; l < k; l++) s = fma(A[i * k + l], B[l * m + j], s);
      T[i * m + j] = s;
    }
#pragma unroll
  for (int i = 0; i < n * m; i++) C[i] = T[i];
}
template <int n, int k>
__device__ __forceinline__ void mvec(const double* A, const double* x, double* y) {
  double T[n];
#pragma unroll
  for (int i = 0; i < n; i++) {
    double s = A[i * k] * x[0];
#pragma unroll
    for (int l = 1; l < k; l++) s = fma(A[i * k + l], x[l], s);
    T[i] = s;
  }
#pragma unroll
  for (int i = 0; i < n; i++) y[i] = T[i];
}
template <int n, int m>
__device__ __forceinline__ void mtr(const double* A, double* At) {
  double T[n * m];
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < m; j++) T[j * n + i] = A[i * m + j];
#pragma unroll
  for (int i = 0; i < n * m; i++) At[i] = T[i];
}
template <int n>
__device__ __forceinline__ double vdot(const double* a, const double* b) {
  double s = a[0] * b[0];
#pragma unroll
  for (int i = 1; i < n; i++) s = fma(a[i], b[i], s);
  return s;
}
template <int d>
__device__ __forceinline__ int minv(const double* A, double* Ai) {
  if constexpr (d == 1) {
    if (A[0] == 0.0) return -1;
    Ai[0] = 1.0 / A[0];
    return 0;
  } else if constexpr (d == 2) {
    double det = A[0] * A[3] - A[1] * A[2];
    if (det == 0.0) return -1;
    double id = 1.0 / det;
    double r0 = A[3] * id, r1 = -A[1] * id, r2 = -A[2] * id, r3 = A[0] * id;
    Ai[0] = r0; Ai[1] = r1; Ai[2] = r2; Ai[3] = r3;
    return 0;
  } else {
    static_assert(d == 3, "d <= 3");
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
#pragma unroll
    for (int i = 0; i < 9; i++) Ai[i] = T[i];
    return 0;
  }
}
template <int d>
__device__ __forceinline__ double trprod(const double* A, const double* B) {
  double Pm[d * d];
  mmul<d, d, d>(A, B, Pm);
  double s = Pm[0];
#pragma unroll
  for (int i = 1; i < d; i++) s += Pm[i * d + i];
  return s;
}
template <int d>
__device__ __forceinline__ int chol_lower(const double* A, double* Lc) {
  for (int i = 0; i < d * d; i++) Lc[i] = 0.0;
  for (int j = 0; j < d; j++) {
    double s = A[j * d + j];
    for (int k = 0; k < j; k++) s -= Lc[j * d + k] * Lc[j * d + k];
    if (!(s > 0.0)) return -1;
    Lc[j * d + j] = sqrt(s);
    for (int i = j + 1; i < d; i++) {
      double t = A[i * d + j];
      for (int k = 0; k < j; k++) t -= Lc[i * d + k] * Lc[j * d + k];
      Lc[i * d + j] = t / Lc[j * d + j];
    }
  }
  return 0;
}

/* auxiliary process values of interval i, Ralston stage k */
struct aux_dev {
  const double* B;
  const double* beta;
  const double* a;
  const double* a_left;
  int is_const;
};
template <int d>
struct aux_at {
  const double *B, *beta, *a;
  __device__ aux_at(const aux_dev& A, int i, int k) {
    if (A.is_const) {
      B = A.B; beta = A.beta; a = A.a;
    } else {
      size_t e = (size_t)3 * i + k;
      B = A.B + e * d * d; beta = A.beta + e * d; a = A.a + e * d * d;
    }
  }
};

/* kernelr3(f, t, y, dt)  src/ode.jl:44-49 */
template <int n, class F>
__device__ __forceinline__ void r3_step(F f, double* y, double h) {
  double k1[n], k2[n], k3[n], yt[n];
  f(0, y, k1);
  const double c2 = 1.0 / 2 * h;
#pragma unroll
  for (int i = 0; i < n; i++) yt[i] = fma(c2, k1[i], y[i]);
  f(1, yt, k2);
  const double c3 = 3.0 / 4 * h;
#pragma unroll
  for (int i = 0; i < n; i++) yt[i] = fma(c3, k2[i], y[i]);
  f(2, yt, k3);
  const double w1 = 2.0 / 9, w2 = 1.0 / 3, w3 = 4.0 / 9;
#pragma unroll
  for (int i = 0; i < n; i++) {
    double s = w1 * k1[i];
    s = fma(w2, k2[i], s);
    s = fma(w3, k3[i], s);
    y[i] = fma(h, s, y[i]);
  }
}

template <int d>
struct rhs_btilde { /* B y + beta */
  const aux_dev& A; int i;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    mvec<d, d>(s.B, y, k);
#pragma unroll
    for (int q = 0; q < d; q++) k[q] += s.beta[q];
  }
};
template <int d>
struct rhs_dHplus { /* B y + (B y)' - a */
  const aux_dev& A; int i;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double BY[d * d];
    mmul<d, d, d>(s.B, y, BY);
#pragma unroll
    for (int r = 0; r < d; r++)
#pragma unroll
      for (int c = 0; c < d; c++) k[r * d + c] = (BY[r * d + c] + BY[c * d + r]) - s.a[r * d + c];
  }
};
template <int d>
struct rhs_dHinv { /* B K + K B' - a */
  const aux_dev& A; int i;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double BK[d * d], Bt[d * d], KBt[d * d];
    mmul<d, d, d>(s.B, y, BK);
    mtr<d, d>(s.B, Bt);
    mmul<d, d, d>(y, Bt, KBt);
#pragma unroll
    for (int q = 0; q < d * d; q++) k[q] = (BK[q] + KBt[q]) - s.a[q];
  }
};
template <int d>
struct rhs_dH { /* -B'y - yB + y a y' */
  const aux_dev& A; int i;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double Bt[d * d], nBt[d * d], T1[d * d], T2[d * d], T3[d * d], yt[d * d];
    mtr<d, d>(s.B, Bt);
#pragma unroll
    for (int q = 0; q < d * d; q++) nBt[q] = -Bt[q];
    mmul<d, d, d>(nBt, y, T1);
    mmul<d, d, d>(y, s.B, T2);
    mmul<d, d, d>(y, s.a, T3);
    mtr<d, d>(y, yt);
    mmul<d, d, d>(T3, yt, T3);
#pragma unroll
    for (int q = 0; q < d * d; q++) k[q] = (T1[q] - T2[q]) + T3[q];
  }
};
template <int d>
struct rhs_dF { /* -B'y + H a y + H beta */
  const aux_dev& A; int i; const double* H;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double Bt[d * d], nBt[d * d], Ha[d * d], t1[d], t2[d], t3[d];
    mtr<d, d>(s.B, Bt);
#pragma unroll
    for (int q = 0; q < d * d; q++) nBt[q] = -Bt[q];
    mvec<d, d>(nBt, y, t1);
    mmul<d, d, d>(H, s.a, Ha);
    mvec<d, d>(Ha, y, t2);
    mvec<d, d>(H, s.beta, t3);
#pragma unroll
    for (int q = 0; q < d; q++) k[q] = (t1[q] + t2[q]) + t3[q];
  }
};
template <int d, int m>
struct rhs_dL { /* -y B */
  const aux_dev& A; int i;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double ny[m * d];
#pragma unroll
    for (int q = 0; q < m * d; q++) ny[q] = -y[q];
    mmul<m, d, d>(ny, s.B, k);
  }
};
template <int d, int m>
struct rhs_dMplus { /* -(L a L') */
  const aux_dev& A; int i; const double* L;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double La[m * d], Lt[d * m], T[m * m];
    mmul<m, d, d>(L, s.a, La);
    mtr<m, d>(L, Lt);
    mmul<m, d, m>(La, Lt, T);
#pragma unroll
    for (int q = 0; q < m * m; q++) k[q] = -T[q];
  }
};
template <int d, int m>
struct rhs_dmu { /* -L beta */
  const aux_dev& A; int i; const double* L;
  __device__ void operator()(int st, const double* y, double* k) const {
    aux_at<d> s(A, i, st);
    double nL[m * d];
#pragma unroll
    for (int q = 0; q < m * d; q++) nL[q] = -L[q];
    mvec<m, d>(nL, s.beta, k);
  }
};

/* one backward step i+1 -> i of partialbridgeodeνH! (R3: partialbridgenuH.jl:36-52; Lyap: :91-101).
 * in/out: Hp = H⁺, Hc = H = inv(H⁺) (old value on entry, new on exit), v = ν, Cc = C.  dt = tt[i]-tt[i+1] < 0.
 * Returns -1 if a matrix is singular. */
template <int d>
__device__ __forceinline__ int nuH_step(int method, const aux_dev& A, int i, double dt, double* Hp, double* Hc,
                                        double* v, double& Cc) {
  aux_at<d> s0(A, i, 0);
  if (method == BB_ODE_R3) {
    r3_step<d * d>(rhs_dHplus<d>{A, i}, Hp, dt);
    double F[d], aF[d];
    mvec<d, d>(Hc, v, F);
    mvec<d, d>(s0.a, F, aF);
    double dC = (vdot<d>(s0.beta, F) + 0.5 * vdot<d>(F, aF)) - 0.5 * trprod<d>(Hc, s0.a);
    Cc += dC * dt;
    r3_step<d>(rhs_btilde<d>{A, i}, v, dt);
  } else {
    double F[d], aF[d];
    mvec<d, d>(Hc, v, F);
    r3_step<d>(rhs_btilde<d>{A, i}, v, dt);
    /* lyapunovpsdbackward_step(t, H⁺, h, P): ϕ (H⁺ + ½h a(t-h)) ϕ' + ½h a(t), ϕ = (I+½hB)\(I-½hB), B at t-h/2 */
    const double h = -dt;
    aux_at<d> s1(A, i, 1);
    double Pm[d * d], Mm[d * d], Pi[d * d], phi[d * d], phit[d * d], Y[d * d], T[d * d];
    const double hh = 1.0 / 2 * h;
#pragma unroll
    for (int r = 0; r < d; r++)
#pragma unroll
      for (int c = 0; c < d; c++) {
        const double id = (r == c) ? 1.0 : 0.0;
        Pm[r * d + c] = id + hh * s1.B[r * d + c];
        Mm[r * d + c] = id - hh * s1.B[r * d + c];
      }
    if (minv<d>(Pm, Pi)) return -1;
    mmul<d, d, d>(Pi, Mm, phi);
    const double* al = A.is_const ? A.a : A.a_left + (size_t)i * d * d;
#pragma unroll
    for (int q = 0; q < d * d; q++) Y[q] = fma(hh, al[q], Hp[q]);
    mmul<d, d, d>(phi, Y, T);
    mtr<d, d>(phi, phit);
    mmul<d, d, d>(T, phit, T);
#pragma unroll
    for (int q = 0; q < d * d; q++) Hp[q] = fma(hh, s0.a[q], T[q]);
    mvec<d, d>(s0.a, F, aF);
    Cc += (vdot<d>(s0.beta, F) * dt + 0.5 * vdot<d>(F, aF) * dt) - 0.5 * trprod<d>(Hc, s0.a) * dt;
  }
  if (minv<d>(Hp, Hc)) return -1;
  return 0;
}

/* observation update  Z = I - H⁺L'(Σ + LH⁺L')⁻¹L;  ν <- Z H⁺L'Σ⁻¹v + Zν;  H⁺ <- Z H⁺  (in place)
 * partialbridge_bolus3.jl:128-137 / src/guip.jl:221-243 (with the H⁺ = Inf diag branch).  Returns -1 if singular. */
template <int d, int m>
__device__ __forceinline__ int gpupdate_dev(double* nu, double* Hplus, const double* L, const double* Sigma,
                                            const double* v) {
  bool allinf = true;
#pragma unroll
  for (int i = 0; i < d; i++)
    if (!(isinf(Hplus[i * d + i]) && Hplus[i * d + i] > 0)) allinf = false;
  double Si[m * m], Lt[d * m];
  if (minv<m>(Sigma, Si)) return -1;
  mtr<m, d>(L, Lt);
  if (allinf) {
    double LtSi[d * m], Am[d * d], Ai[d * d], rhs[d];
    mmul<d, m, m>(Lt, Si, LtSi);
    mmul<d, m, d>(LtSi, L, Am);
    if (minv<d>(Am, Ai)) return -1;
    mvec<d, m>(LtSi, v, rhs);
    mvec<d, d>(Ai, rhs, nu);
#pragma unroll
    for (int q = 0; q < d * d; q++) Hplus[q] = Ai[q];
    return 0;
  }
  double HLt[d * m], LHLt[m * m], G[m * m], Gi[m * m], T[d * m], Z[d * d], ZH[d * d], t1[d], t2[d];
  mmul<d, d, m>(Hplus, Lt, HLt);
  mmul<m, d, m>(L, HLt, LHLt);
#pragma unroll
  for (int q = 0; q < m * m; q++) G[q] = Sigma[q] + LHLt[q];
  if (minv<m>(G, Gi)) return -1;
  mmul<d, m, m>(HLt, Gi, T);
  mmul<d, m, d>(T, L, Z);
#pragma unroll
  for (int i = 0; i < d; i++)
#pragma unroll
    for (int j = 0; j < d; j++) Z[i * d + j] = ((i == j) ? 1.0 : 0.0) - Z[i * d + j];
  mmul<d, d, d>(Z, Hplus, ZH);
  double ZHLt[d * m], ZHLtSi[d * m];
  mmul<d, d, m>(ZH, Lt, ZHLt);
  mmul<d, m, m>(ZHLt, Si, ZHLtSi);
  mvec<d, m>(ZHLtSi, v, t1);
  mvec<d, d>(Z, nu, t2);
#pragma unroll
  for (int q = 0; q < d; q++) nu[q] = t1[q] + t2[q];
#pragma unroll
  for (int q = 0; q < d * d; q++) Hplus[q] = ZH[q];
  return 0;
}

}  // namespace bbk
