/*
 * bb_backward_gen.cu -- the PartialBridgeνH constructors for state dimensions 4 .. 16 (BB_MODEL_LANDMARKS, d = 16):
 *   updateνH⁺C                      src/partialbridgenuH.jl:1-17
 *   partialbridgeodeνH!(R3 / Lyap)  src/partialbridgenuH.jl:21-55, 86-103; src/lyap.jl:2-6; kernelr3 src/ode.jl:44-49
 * One CTA of 256 threads integrates the one d x d system: thread (i, j) owns element (i, j) of every matrix, all
 * matrices live in shared memory.  Every element is computed by the same sequence of roundings as the oracle's
 * serial loops (products accumulated left to right with fma, Gauss-Jordan inverse with partial pivoting and plain
 * multiply/subtract), so the tables are bit-identical to a serial evaluation in the same rounding order (tests/); only the work of one step is spread over
 * the threads.  d <= 3 keeps the one-thread closed-form kernels of bb_backward.cu.
 */
#include <math.h>
#include <string.h>

#include <vector>

#include "bb_host.h"

namespace {

constexpr int GD = 16;        /* matrices are stored with row stride GD */
constexpr int GT = GD * GD;   /* threads per CTA */

struct gaux {
  const double *B, *beta, *a, *a_left;
  int is_const;
};
__device__ __forceinline__ const double* gaux_B(const gaux& A, int d, int i, int k) {
  return A.is_const ? A.B : A.B + ((size_t)3 * i + k) * d * d;
}
__device__ __forceinline__ const double* gaux_beta(const gaux& A, int d, int i, int k) {
  return A.is_const ? A.beta : A.beta + ((size_t)3 * i + k) * d;
}
__device__ __forceinline__ const double* gaux_a(const gaux& A, int d, int i, int k) {
  return A.is_const ? A.a : A.a + ((size_t)3 * i + k) * d * d;
}

/* C[n][m] = A[n][k] B[k][m]; operands in shared memory with row stride GD or in global memory with their own stride */
__device__ __forceinline__ void g_mmul(int n, int k, int m, const double* A, int lda, const double* B, int ldb, double* C) {
  const int i = threadIdx.x / GD, j = threadIdx.x % GD;
  double s = 0.0;
  if (i < n && j < m) {
    s = A[i * lda] * B[j];
    for (int l = 1; l < k; l++) s = fma(A[i * lda + l], B[l * ldb + j], s);
  }
  __syncthreads(); /* C may alias A or B */
  if (i < n && j < m) C[i * GD + j] = s;
  __syncthreads();
}
/* y[n] = A[n][k] x[k] */
__device__ __forceinline__ void g_mvec(int n, int k, const double* A, int lda, const double* x, double* y) {
  const int i = threadIdx.x;
  double s = 0.0;
  if (i < n) {
    s = A[i * lda] * x[0];
    for (int l = 1; l < k; l++) s = fma(A[i * lda + l], x[l], s);
  }
  __syncthreads();
  if (i < n) y[i] = s;
  __syncthreads();
}
/* Ai = inv(A), d x d, Gauss-Jordan with partial pivoting on the augmented matrix W [GD][2 GD] (oracle mat_inv, d > 3).
 * Returns 0, or -1 if singular (uniform over the CTA). */
__device__ int g_inv(int d, const double* A, double* Ai, double* W, int* piv) {
  const int t = threadIdx.x;
  for (int e = t; e < d * 2 * d; e += GT) {
    const int i = e / (2 * d), j = e % (2 * d);
    W[i * 2 * GD + j] = j < d ? A[i * GD + j] : ((j - d) == i ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int c = 0; c < d; c++) {
    if (t == 0) {
      int p = c;
      for (int i = c + 1; i < d; i++)
        if (fabs(W[i * 2 * GD + c]) > fabs(W[p * 2 * GD + c])) p = i;
      *piv = (W[p * 2 * GD + c] == 0.0) ? -1 : p;
    }
    __syncthreads();
    const int p = *piv;
    if (p < 0) return -1;
    if (p != c && t < 2 * d) {
      const double u = W[c * 2 * GD + t];
      W[c * 2 * GD + t] = W[p * 2 * GD + t];
      W[p * 2 * GD + t] = u;
    }
    __syncthreads();
    const double ip = 1.0 / W[c * 2 * GD + c];
    __syncthreads();
    if (t < 2 * d) W[c * 2 * GD + t] *= ip;
    __syncthreads();
    /* rows i != c: W[i][j] -= f W[c][j], f = W[i][c] (read before anything in the row changes) */
    double f[2];
    for (int q = 0; q < 2; q++) {
      const int e = t + q * GT, i = e / (2 * d);
      f[q] = (e < d * 2 * d) ? W[i * 2 * GD + c] : 0.0;
    }
    __syncthreads();
    for (int q = 0; q < 2; q++) {
      const int e = t + q * GT, i = e / (2 * d), j = e % (2 * d);
      if (e < d * 2 * d && i != c && f[q] != 0.0) W[i * 2 * GD + j] -= f[q] * W[c * 2 * GD + j];
    }
    __syncthreads();
  }
  for (int e = t; e < d * d; e += GT) Ai[(e / d) * GD + e % d] = W[(e / d) * 2 * GD + d + e % d];
  __syncthreads();
  return 0;
}

/* elementwise helper: thread (i, j) with i, j < d */
#define G_EL(d) const int gi_ = threadIdx.x / GD, gj_ = threadIdx.x % GD; if (gi_ < (d) && gj_ < (d))
#define G_IX (gi_ * GD + gj_)

struct gshared {
  double Hp[GT], Hc[GT], K1[GT], K2[GT], K3[GT], Yt[GT], T1[GT], T2[GT], W[GD * 2 * GD];
  double v[GD], k1[GD], k2[GD], k3[GD], yt[GD], F[GD], aF[GD], dg[GD];
  double Cc;
  int piv;
};

/* k = B Y + (B Y)' - a   (rhs of dH⁺, src/partialbridgenuH.jl:40) */
__device__ __forceinline__ void g_rhs_dHplus(int d, const double* B, const double* a, const double* Y, double* K, double* tmp) {
  g_mmul(d, d, d, B, d, Y, GD, tmp);
  { G_EL(d) K[G_IX] = (tmp[gi_ * GD + gj_] + tmp[gj_ * GD + gi_]) - a[gi_ * d + gj_]; }
  __syncthreads();
}
/* k = B y + beta */
__device__ __forceinline__ void g_rhs_btilde(int d, const double* B, const double* beta, const double* y, double* k) {
  g_mvec(d, d, B, d, y, k);
  if (threadIdx.x < d) k[threadIdx.x] += beta[threadIdx.x];
  __syncthreads();
}

__global__ void __launch_bounds__(GT) k_backward_nuH_gen(int method, int N, int d, const double* __restrict__ tt, gaux A,
                                                         const double* nu_end, const double* Hplus_end, double C0,
                                                         double* nu, double* H, double* out_left, int* status) {
  extern __shared__ __align__(16) unsigned char gsm[];
  gshared& S = *reinterpret_cast<gshared*>(gsm);
  const int t = threadIdx.x;
  { G_EL(d) S.Hp[G_IX] = Hplus_end[gi_ * d + gj_]; }
  if (t < d) S.v[t] = nu_end[t];
  if (t == 0) S.Cc = C0;
  __syncthreads();
  if (g_inv(d, S.Hp, S.Hc, S.W, &S.piv)) { if (t == 0) *status = BB_ERR_SINGULAR; return; }
  { G_EL(d) H[(size_t)(N - 1) * d * d + gi_ * d + gj_] = S.Hc[G_IX]; }
  if (t < d) nu[(size_t)(N - 1) * d + t] = S.v[t];
  const double w1 = 2.0 / 9, w2 = 1.0 / 3, w3 = 4.0 / 9;
  for (int i = N - 2; i >= 0; i--) {
    const double dt = tt[i] - tt[i + 1];
    const double c2 = 1.0 / 2 * dt, c3 = 3.0 / 4 * dt;
    const double *B0 = gaux_B(A, d, i, 0), *be0 = gaux_beta(A, d, i, 0), *a0 = gaux_a(A, d, i, 0);
    /* F = H ν with the OLD H and ν  (:45 / :97) */
    g_mvec(d, d, S.Hc, GD, S.v, S.F);
    if (method == BB_ODE_R3) {
      /* H⁺ <- kernelr3(dH⁺)  (:40) */
      g_rhs_dHplus(d, B0, a0, S.Hp, S.K1, S.T1);
      { G_EL(d) S.Yt[G_IX] = fma(c2, S.K1[G_IX], S.Hp[G_IX]); }
      __syncthreads();
      g_rhs_dHplus(d, gaux_B(A, d, i, 1), gaux_a(A, d, i, 1), S.Yt, S.K2, S.T1);
      { G_EL(d) S.Yt[G_IX] = fma(c3, S.K2[G_IX], S.Hp[G_IX]); }
      __syncthreads();
      g_rhs_dHplus(d, gaux_B(A, d, i, 2), gaux_a(A, d, i, 2), S.Yt, S.K3, S.T1);
      {
        G_EL(d) {
          double s = w1 * S.K1[G_IX];
          s = fma(w2, S.K2[G_IX], s);
          s = fma(w3, S.K3[G_IX], s);
          S.Hp[G_IX] = fma(dt, s, S.Hp[G_IX]);
        }
      }
      __syncthreads();
    }
    /* C += (β·F + ½ F'aF - ½ tr(H a)) dt   (R3 :31,46; Lyap :98 distributes dt) */
    g_mvec(d, d, a0, d, S.F, S.aF);
    if (t < d) { /* diagonal of H a, accumulated as mat_mul does */
      double s = S.Hc[t * GD] * a0[t];
      for (int l = 1; l < d; l++) s = fma(S.Hc[t * GD + l], a0[l * d + t], s);
      S.dg[t] = s;
    }
    __syncthreads();
    if (t == 0) {
      double bF = be0[0] * S.F[0], FaF = S.F[0] * S.aF[0], tr = S.dg[0];
      for (int l = 1; l < d; l++) {
        bF = fma(be0[l], S.F[l], bF);
        FaF = fma(S.F[l], S.aF[l], FaF);
        tr += S.dg[l];
      }
      if (method == BB_ODE_R3) S.Cc += ((bF + 0.5 * FaF) - 0.5 * tr) * dt;
      else S.Cc += (bF * dt + 0.5 * FaF * dt) - 0.5 * tr * dt;
    }
    /* ν <- kernelr3(B ν + β)  (:49 / :95) */
    g_rhs_btilde(d, B0, be0, S.v, S.k1);
    if (t < d) S.yt[t] = fma(c2, S.k1[t], S.v[t]);
    __syncthreads();
    g_rhs_btilde(d, gaux_B(A, d, i, 1), gaux_beta(A, d, i, 1), S.yt, S.k2);
    if (t < d) S.yt[t] = fma(c3, S.k2[t], S.v[t]);
    __syncthreads();
    g_rhs_btilde(d, gaux_B(A, d, i, 2), gaux_beta(A, d, i, 2), S.yt, S.k3);
    if (t < d) {
      double s = w1 * S.k1[t];
      s = fma(w2, S.k2[t], s);
      s = fma(w3, S.k3[t], s);
      S.v[t] = fma(dt, s, S.v[t]);
    }
    __syncthreads();
    if (method != BB_ODE_R3) {
      /* lyapunovpsdbackward_step: ϕ (H⁺ + ½h a(t-h)) ϕ' + ½h a(t), ϕ = (I + ½hB)\(I - ½hB), B at t - h/2  (src/lyap.jl:2-6) */
      const double hh = 1.0 / 2 * (-dt);
      const double* B1 = gaux_B(A, d, i, 1);
      const double* al = A.is_const ? A.a : A.a_left + (size_t)i * d * d;
      {
        G_EL(d) {
          const double id = (gi_ == gj_) ? 1.0 : 0.0;
          S.K1[G_IX] = id + hh * B1[gi_ * d + gj_];
          S.K2[G_IX] = id - hh * B1[gi_ * d + gj_];
          S.Yt[G_IX] = fma(hh, al[gi_ * d + gj_], S.Hp[G_IX]);
        }
      }
      __syncthreads();
      if (g_inv(d, S.K1, S.K3, S.W, &S.piv)) { if (t == 0) *status = BB_ERR_SINGULAR; return; }
      g_mmul(d, d, d, S.K3, GD, S.K2, GD, S.T1);  /* ϕ */
      g_mmul(d, d, d, S.T1, GD, S.Yt, GD, S.T2);  /* ϕ Y */
      { G_EL(d) S.K1[G_IX] = S.T1[gj_ * GD + gi_]; } /* ϕ' */
      __syncthreads();
      g_mmul(d, d, d, S.T2, GD, S.K1, GD, S.T2);
      { G_EL(d) S.Hp[G_IX] = fma(hh, a0[gi_ * d + gj_], S.T2[G_IX]); }
      __syncthreads();
    }
    if (g_inv(d, S.Hp, S.Hc, S.W, &S.piv)) { if (t == 0) *status = BB_ERR_SINGULAR; return; }
    { G_EL(d) H[(size_t)i * d * d + gi_ * d + gj_] = S.Hc[G_IX]; }
    if (t < d) nu[(size_t)i * d + t] = S.v[t];
  }
  { G_EL(d) out_left[d + gi_ * d + gj_] = S.Hp[G_IX]; }
  if (t < d) out_left[t] = S.v[t];
  __syncthreads();
  if (t == 0) {
    out_left[d + d * d] = S.Cc;
    *status = BB_OK;
  }
}

/* updateνH⁺C  partialbridgenuH.jl:1-17.  in = L[m*d], Sigma[m*m], v[m], eps; out = nu[d], Hplus[d*d], C */
__global__ void __launch_bounds__(GT) k_update_nuHC_gen(int d, int m, const double* in, double* out, int* status) {
  extern __shared__ __align__(16) unsigned char gsm[];
  gshared& S = *reinterpret_cast<gshared*>(gsm);
  const int t = threadIdx.x;
  const double *L = in, *Sigma = in + m * d, *v = Sigma + m * m;
  const double eps = v[m];
  /* Si = inv(Σ) -> K1;  Lt -> K2 (d x m);  LtSi -> K3;  H -> Hp;  H⁺ -> Hc */
  { G_EL(m) S.Yt[G_IX] = Sigma[gi_ * m + gj_]; }
  __syncthreads();
  if (m <= 3) { /* closed forms as the oracle's mat_inv for m <= 3 are not reproduced here: m >= 4 only */
    if (t == 0) *status = BB_ERR_UNSUPPORTED;
    return;
  }
  if (g_inv(m, S.Yt, S.K1, S.W, &S.piv)) { if (t == 0) *status = BB_ERR_SINGULAR; return; }
  {
    const int i = t / GD, j = t % GD;
    if (i < d && j < m) S.K2[i * GD + j] = L[j * d + i];
  }
  __syncthreads();
  g_mmul(d, m, m, S.K2, GD, S.K1, GD, S.K3);
  g_mmul(d, m, d, S.K3, GD, L, d, S.Hp);
  if (t < d) S.Hp[t * GD + t] += eps;
  __syncthreads();
  if (g_inv(d, S.Hp, S.Hc, S.W, &S.piv)) { if (t == 0) *status = BB_ERR_SINGULAR; return; }
  g_mmul(d, d, m, S.Hc, GD, S.K2, GD, S.T1);   /* H⁺ L' */
  g_mmul(d, m, m, S.T1, GD, S.K1, GD, S.T2);   /* H⁺ L' Σ⁻¹ */
  g_mvec(d, m, S.T2, GD, v, S.v);              /* ν */
  g_mvec(m, m, S.K1, GD, v, S.F);              /* Σ⁻¹ v */
  if (t == 0) {
    double q = v[0] * S.F[0];
    for (int l = 1; l < m; l++) q = fma(v[l], S.F[l], q);
    double c = 0.0;
    c += 0.5 * q;
    /* logdet Σ by Cholesky (oracle chol_lower) */
    double* Lc = S.K3;
    double ld = 0.0;
    bool bad = false;
    for (int i = 0; i < m * GD; i++) Lc[i] = 0.0;
    for (int j = 0; j < m && !bad; j++) {
      double s = Sigma[j * m + j];
      for (int k = 0; k < j; k++) s -= Lc[j * GD + k] * Lc[j * GD + k];
      if (!(s > 0.0)) { bad = true; break; }
      Lc[j * GD + j] = sqrt(s);
      for (int i = j + 1; i < m; i++) {
        double u = Sigma[i * m + j];
        for (int k = 0; k < j; k++) u -= Lc[i * GD + k] * Lc[j * GD + k];
        Lc[i * GD + j] = u / Lc[j * GD + j];
      }
    }
    for (int i = 0; i < m; i++) ld += 2 * log(Lc[i * GD + i]);
    c += m / 2.0 * log(2 * 3.14159265358979323846) + 0.5 * ld;
    out[d + d * d] = c;
    *status = bad ? BB_ERR_SINGULAR : BB_OK;
  }
  if (t < d) out[t] = S.v[t];
  { G_EL(d) out[d + gi_ * d + gj_] = S.Hc[G_IX]; }
}

}  // namespace

/* host entry points used by bb_backward.cu for d > 3; all pointers are device pointers */
cudaError_t bb_gen_backward_nuH(cudaStream_t st, int method, int N, int d, const double* tt, const double* B,
                                const double* beta, const double* a, const double* a_left, int is_const,
                                const double* nu_end, const double* Hplus_end, double C0, double* nu, double* H,
                                double* out_left, int* status) {
  /* the attribute is per device; the working set (sizeof(gshared) ~ 21 KB) is below the default limit anyway */
  cudaError_t e = cudaFuncSetAttribute(k_backward_nuH_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(gshared));
  if (e != cudaSuccess) return e;
  gaux A{B, beta, a, a_left, is_const};
  k_backward_nuH_gen<<<1, GT, sizeof(gshared), st>>>(method, N, d, tt, A, nu_end, Hplus_end, C0, nu, H, out_left, status);
  return cudaGetLastError();
}
cudaError_t bb_gen_update_nuHC(cudaStream_t st, int d, int m, const double* in, double* out, int* status) {
  cudaError_t e = cudaFuncSetAttribute(k_update_nuHC_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(gshared));
  if (e != cudaSuccess) return e;
  k_update_nuHC_gen<<<1, GT, sizeof(gshared), st>>>(d, m, in, out, status);
  return cudaGetLastError();
}
