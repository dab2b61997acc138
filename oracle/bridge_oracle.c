/*
 * bridge_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C, single-path, double-precision CPU restatement of the Bridge.jl loops
 * that libbridge_b200.so replaces.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's library.
 * The product (bridge.jl_b200/, libbridge_b200.so) never links, imports or calls it.
 *
 * Why a restatement: the reference is Julia and no Julia runtime exists in this
 * image or on the GPU box, so the reference cannot be run.
 * Every function cites the reference file:line it follows (paths relative to the
 * Bridge.jl checkout, commit b09488fe).  Parity pins: see oracle/README.md and
 * tests/test_oracle_pins.py (docs golden vector, closed-form LinPro answers, the
 * reference tests' own tolerances).  Bit-level parity with Julia is UNPINNED
 * (StaticArrays/LinearAlgebra versions are not locked by the reference);
 * tolerance-level parity is pinned.
 *
 * Two builds of this file exist (oracle/Makefile):
 *   liboracle_ref.so  -O2 -ffp-contract=off            reference arithmetic: no fused
 *                     multiply-add anywhere (Julia does not contract a*b+c), division
 *                     where the reference divides.
 *   liboracle_fma.so  -DORACLE_GPU_ORDER -ffp-contract=off   same algorithm, but MA(a,b,c)
 *                     is a true fma() and x/eps becomes x*(1/eps) exactly where the CUDA
 *                     kernels do so; used to show the kernels are bit-identical to a CPU
 *                     evaluation of the same rounding sequence.
 * Random numbers: Julia's MersenneTwister/Xoshiro streams cannot be reproduced, so
 * the noise source is Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123
 * constants) with a documented counter layout, identical here and in the kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/bridge_b200.h"

#define DM 20          /* max state dimension handled by the oracle (test/lyap.jl uses d=20) */
#define DM2 (DM * DM)

#ifdef ORACLE_GPU_ORDER
#define MA(a, b, c) fma((a), (b), (c))
#else
#define MA(a, b, c) ((a) * (b) + (c))
#endif

int bbo_gpu_order(void) {
#ifdef ORACLE_GPU_ORDER
  return 1;
#else
  return 0;
#endif
}

/* ======================================================================= small dense algebra (row-major) */
static void mat_mul(int n, int k, int m, const double* A, const double* B, double* C) {
  /* C[n][m] = A[n][k] B[k][m]; plain left-to-right accumulation as StaticArrays' unrolled products */
  double T[DM2];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      double s = A[i * k] * B[j];
      for (int l = 1; l < k; l++) s = MA(A[i * k + l], B[l * m + j], s);
      T[i * m + j] = s;
    }
  memcpy(C, T, sizeof(double) * n * m);
}
static void mat_vec(int n, int k, const double* A, const double* x, double* y) {
  double T[DM];
  for (int i = 0; i < n; i++) {
    double s = A[i * k] * x[0];
    for (int l = 1; l < k; l++) s = MA(A[i * k + l], x[l], s);
    T[i] = s;
  }
  memcpy(y, T, sizeof(double) * n);
}
static void mat_tr(int n, int m, const double* A, double* At) {
  double T[DM2];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) T[j * n + i] = A[i * m + j];
  memcpy(At, T, sizeof(double) * n * m);
}
static double vdot(int n, const double* a, const double* b) {
  double s = a[0] * b[0];
  for (int i = 1; i < n; i++) s = MA(a[i], b[i], s);
  return s;
}
/* inverse: closed forms for d <= 3 (as StaticArrays' inv), Gauss-Jordan with partial pivoting above */
static int mat_inv(int d, const double* A, double* Ai) {
  if (d == 1) {
    if (A[0] == 0.0) return -1;
    Ai[0] = 1.0 / A[0];
    return 0;
  }
  if (d == 2) {
    double det = A[0] * A[3] - A[1] * A[2];
    if (det == 0.0) return -1;
    double id = 1.0 / det;
    double r0 = A[3] * id, r1 = -A[1] * id, r2 = -A[2] * id, r3 = A[0] * id;
    Ai[0] = r0; Ai[1] = r1; Ai[2] = r2; Ai[3] = r3;
    return 0;
  }
  if (d == 3) {
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
    memcpy(Ai, T, sizeof(T));
    return 0;
  }
  double M[DM][2 * DM];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      M[i][j] = A[i * d + j];
      M[i][d + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < d; c++) {
    int p = c;
    for (int i = c + 1; i < d; i++)
      if (fabs(M[i][c]) > fabs(M[p][c])) p = i;
    if (M[p][c] == 0.0) return -1;
    if (p != c)
      for (int j = 0; j < 2 * d; j++) {
        double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t;
      }
    double ip = 1.0 / M[c][c];
    for (int j = 0; j < 2 * d; j++) M[c][j] *= ip;
    for (int i = 0; i < d; i++)
      if (i != c) {
        double f = M[i][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * d; j++) M[i][j] -= f * M[c][j];
      }
  }
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) Ai[i * d + j] = M[i][d + j];
  return 0;
}
static double mat_trace_prod(int d, const double* A, const double* B) { /* tr(A*B) */
  double P[DM2];
  mat_mul(d, d, d, A, B, P);
  double s = P[0];
  for (int i = 1; i < d; i++) s += P[i * d + i];
  return s;
}
/* lower Cholesky factor; returns -1 if not positive definite */
static int chol_lower(int d, const double* A, double* Lc) {
  memset(Lc, 0, sizeof(double) * d * d);
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

/* logpdfnormal(x, Σ): src/gaussian.jl:66-75 */
double bbo_logpdfnormal(int d, const double* x, const double* Sigma) {
  if (d == 1) return -(x[0] * x[0] / Sigma[0] + log(Sigma[0]) + log(2 * M_PI)) / 2;
  double S[DM2], Ssym[DM2], y[DM];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) Ssym[i * d + j] = 0.5 * (Sigma[i * d + j] + Sigma[j * d + i]);
  if (chol_lower(d, Ssym, S)) return NAN;
  double n2 = 0, sld = 0;
  for (int i = 0; i < d; i++) { /* forward substitution S y = x */
    double t = x[i];
    for (int k = 0; k < i; k++) t -= S[i * d + k] * y[k];
    y[i] = t / S[i * d + i];
    n2 += y[i] * y[i];
    sld += log(S[i * d + i]);
  }
  return -(n2 + 2 * sld + d * log(2 * M_PI)) / 2;
}

/* exp(x), x <= 0, of the landmarks kernel.  Reference arithmetic: libm exp (what Julia computes).  GPU order: the
 * kernels' own exponential (bb_device.cuh bb_exp: +, *, fma, rint only), so that paths can be compared bit for bit. */
static double bb_exp(double x) {
#ifdef ORACLE_GPU_ORDER
  const double kd = rint(x * 0x1.71547652b82fep+0);
  double r = fma(-kd, 0x1.62e42fee00000p-1, x);
  r = fma(-kd, 0x1.a39ef35793c76p-33, r);
  double p = 0x1.6124613a86d09p-33;
  p = fma(p, r, 0x1.1eed8eff8d898p-29);
  p = fma(p, r, 0x1.ae64567f544e4p-26);
  p = fma(p, r, 0x1.27e4fb7789f5cp-22);
  p = fma(p, r, 0x1.71de3a556c734p-19);
  p = fma(p, r, 0x1.a01a01a01a01ap-16);
  p = fma(p, r, 0x1.a01a01a01a01ap-13);
  p = fma(p, r, 0x1.6c16c16c16c17p-10);
  p = fma(p, r, 0x1.1111111111111p-7);
  p = fma(p, r, 0x1.5555555555555p-5);
  p = fma(p, r, 0x1.5555555555555p-3);
  p = fma(p, r, 0x1.0000000000000p-1);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int k = (int)kd;
  k = k < -1022 ? -1022 : (k > 1023 ? 1023 : k);
  union { uint64_t u; double d; } s;
  s.u = (uint64_t)(k + 1023) << 52;
  return x < -708.0 ? 0.0 : p * s.d;
#else
  return exp(x);
#endif
}
double bbo_exp(double x) { return bb_exp(x); }

/* ======================================================================= random numbers */
/* Philox4x32-10, Random123 (Salmon et al. 2011), constants PHILOX_M4x32_0/1, PHILOX_W32_0/1 */
void bbo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- float32 Box-Muller built from +, *, fma and IEEE sqrt only, so that this CPU evaluation and
 * the kernels' (bridge.jl_b200/csrc/rng.cuh) are BIT-IDENTICAL.  Coefficients: tools/gen_rng_poly.py
 * (log: max rel. error 9e-8; sinpi/cospi: max abs. error 9e-8).  A normal therefore carries a 24-bit
 * mantissa and |z| <= 6.77; it is widened to double before use. */
static inline float bb_logf(float u) { /* u in [2^-33, 1] */
  uint32_t ix;
  memcpy(&ix, &u, 4);
  ix += 0x3F800000u - 0x3F3504F3u;
  int e = (int)(ix >> 23) - 127;
  ix = (ix & 0x007FFFFFu) + 0x3F3504F3u;
  float m;
  memcpy(&m, &ix, 4); /* m in [sqrt(1/2), sqrt(2)) */
  float f = m - 1.0f;
  float p = -0x1.4bde76p-4f;
  p = fmaf(p, f, 0x1.045b0cp-3f);
  p = fmaf(p, f, -0x1.09ab66p-3f);
  p = fmaf(p, f, 0x1.22dbfcp-3f);
  p = fmaf(p, f, -0x1.54d552p-3f);
  p = fmaf(p, f, 0x1.99a15p-3f);
  p = fmaf(p, f, -0x1.0000c6p-2f);
  p = fmaf(p, f, 0x1.555552p-2f);
  float f2 = f * f;
  float t = p * f;
  t = fmaf(t, f2, -0.5f * f2); /* f^3 P(f) - f^2/2 */
  float r = t + f;
  return fmaf((float)e, 0x1.62e43p-1f, r);
}
static inline float bb_unif(uint32_t w) { /* (w + 1/2) / 2^32 rounded to float, in (0, 1] */
  return fmaf((float)w, 0x1p-32f, 0x1p-33f);
}
/* one Box-Muller pair from two 32-bit words: radius from wu, angle pi*t with t = (int32)wa / 2^31 */
static inline void bb_box_muller(uint32_t wu, uint32_t wa, float* z0, float* z1) {
  float rad = sqrtf(-2.0f * bb_logf(bb_unif(wu)));
  float t = (float)(int32_t)wa * 0x1p-31f; /* [-1, 1] */
  float q = rintf(t * 2.0f);               /* -2..2 */
  float r = fmaf(q, -0.5f, t);             /* exact, |r| <= 1/4 */
  float s2 = r * r;
  float ps = -0x1.2d9b7cp-1f;
  ps = fmaf(ps, s2, 0x1.465ec4p+1f);
  ps = fmaf(ps, s2, -0x1.4abbbap+2f);
  ps = fmaf(ps, s2, 0x1.921fb6p+1f);
  float sr = ps * r; /* sin(pi r) */
  float pc = 0x1.d9c326p-3f;
  pc = fmaf(pc, s2, -0x1.55c57ap+0f);
  pc = fmaf(pc, s2, 0x1.03c1dcp+2f);
  pc = fmaf(pc, s2, -0x1.3bd3ccp+2f);
  float cr = fmaf(pc, s2, 1.0f); /* cos(pi r) */
  float sn, cs;
  switch (((int)q) & 3) {
    case 0: sn = sr;  cs = cr;  break;
    case 1: sn = cr;  cs = -sr; break;
    case 2: sn = -sr; cs = -cr; break;
    default: sn = -cr; cs = sr; break;
  }
  *z0 = rad * sn;
  *z1 = rad * cs;
}

/* Counter layout shared with the kernels:
 *   key = (seed lo, seed hi);  counter = (q, stream, id lo, id hi)
 *   normals: id = global row = chain*S + seg ; one Philox call gives FOUR normals: words (0,1) -> elements
 *            0,1 and words (2,3) -> elements 2,3 of quad q.  Normal n = j*d' + k of the row (time index j,
 *            noise component k, the reference's draw order src/wiener.jl:28-33) is element n & 3 of quad
 *            q = n >> 2.  Index j = 0 is drawn but unused.
 *   accept : id = global chain, q = 0xFFFFFFFF, log U = (double) bb_logf(bb_unif(word 0)). */
void bbo_normal_quad(uint64_t seed, uint32_t stream, uint64_t row, uint32_t q, double z[4]) {
  uint32_t ctr[4] = {q, stream, (uint32_t)row, (uint32_t)(row >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  bbo_philox4x32_10(ctr, key, o);
  float a, b, c, d;
  bb_box_muller(o[0], o[1], &a, &b);
  bb_box_muller(o[2], o[3], &c, &d);
  z[0] = (double)a; z[1] = (double)b; z[2] = (double)c; z[3] = (double)d;
}
double bbo_normal(uint64_t seed, uint32_t stream, uint64_t row, uint64_t n) {
  double z[4];
  bbo_normal_quad(seed, stream, row, (uint32_t)(n >> 2), z);
  return z[n & 3];
}
double bbo_accept_logu(uint64_t seed, uint32_t stream, uint64_t chain) {
  uint32_t ctr[4] = {0xFFFFFFFFu, stream, (uint32_t)chain, (uint32_t)(chain >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  bbo_philox4x32_10(ctr, key, o);
  return (double)bb_logf(bb_unif(o[0]));
}

/* log U of the parameter-update accept test: the same generator with q = 0xFFFFFFFD (the pCN accept test of the
 * same iteration keeps 0xFFFFFFFF; the parameter proposal's normals are quad 0xFFFFFFFE of row = chain) */
double bbo_logu_q(uint64_t seed, uint32_t stream, uint64_t chain, uint32_t q) {
  uint32_t ctr[4] = {q, stream, (uint32_t)chain, (uint32_t)(chain >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  bbo_philox4x32_10(ctr, key, o);
  return (double)bb_logf(bb_unif(o[0]));
}

/* ======================================================================= A1: sample!(W, Wiener{T}())
 * src/wiener.jl:50-58 (scalar), :24-35 (SVector), :37-48 (VSamplePath):
 *   yy[1] = y1 (kept);  yy[i] = yy[i-1] + sqrt(tt[i]-tt[i-1]) * randn()   component-minor draw order.
 * W is [N][dp]. */
void bbo_wiener_sample(int N, int dp, const double* tt, uint64_t seed, uint32_t stream,
                       uint64_t row, double* W) {
  for (int j = 1; j < N; j++) {
    double rootdt = sqrt(tt[j] - tt[j - 1]);
    for (int k = 0; k < dp; k++) {
      double xi = bbo_normal(seed, stream, row, (uint64_t)j * dp + k);
      W[j * dp + k] = MA(rootdt, xi, W[(j - 1) * dp + k]);
    }
  }
}

/* ======================================================================= A9: target models (registry) */
/* drift b(t,x,P) */
static void model_b(const bb_model* P, double t, const double* x, double* b) {
  (void)t;
  const double* p = P->par;
  switch (P->id) {
    case BB_MODEL_WIENER: /* src/wiener.jl:145-147 */
      for (int i = 0; i < P->d; i++) b[i] = 0.0;
      break;
    case BB_MODEL_OU: /* docs/src/manual.md:44  b = -P.β * x */
      b[0] = (-p[0]) * x[0];
      break;
    case BB_MODEL_LINPRO: { /* src/linpro.jl:80  b = P.B*(x .- P.μ) */
      int d = P->d;
      const double *B = p, *mu = p + d * d;
      double y[DM];
      for (int i = 0; i < d; i++) y[i] = x[i] - mu[i];
      mat_vec(d, d, B, y, b);
      break;
    }
    case BB_MODEL_FHN_DIAG: { /* src/Models.jl:18  (P.ϵ\(x1 - x1^3 - x2 + P.s), P.γ*x1 - x2 + P.β) */
      double x1 = x[0], x2 = x[1];
      double c = x1 * x1;
#ifdef ORACLE_GPU_ORDER
      double u = fma(-c, x1, x1);
      b[0] = ((u - x2) + p[1]) * (1.0 / p[0]);
#else
      double u = x1 - c * x1;
      b[0] = ((u - x2) + p[1]) / p[0];
#endif
      b[1] = MA(p[2], x1, -x2) + p[3];
      break;
    }
    case BB_MODEL_FHN_HYPO: { /* partialbridge_fitzhugh.jl:44  ((x1-x2-x1^3+P.s)/P.ϵ, P.γ*x1-x2+P.β) */
      double x1 = x[0], x2 = x[1];
      double c = x1 * x1;
      double u = x1 - x2;
#ifdef ORACLE_GPU_ORDER
      u = fma(-c, x1, u);
      b[0] = (u + p[1]) * (1.0 / p[0]);
#else
      u = u - c * x1;
      b[0] = (u + p[1]) / p[0];
#endif
      b[1] = MA(p[2], x1, -x2) + p[3];
      break;
    }
    case BB_MODEL_INTDIFF: /* test/partialbridge.jl:25-26  (x2, -(x2+sin(x2)) + 1/2) */
      b[0] = x[1];
      b[1] = -(x[1] + sin(x[1])) + 0.5;
      break;
    case BB_MODEL_NCLAR3: /* partialbridge_nclar.jl:58  (x2, x3, -P.α*sin(P.ω*x3)) */
      b[0] = x[1];
      b[1] = x[2];
      b[2] = (-p[0]) * sin(p[1] * x[2]);
      break;
    case BB_MODEL_LORENZ: /* src/Models.jl:45  (θ1(x2-x1), x1(θ2-x3)-x2, x1x2-θ3x3) */
      b[0] = p[0] * (x[1] - x[0]);
      b[1] = MA(x[0], (p[1] - x[2]), -x[1]);
#ifdef ORACLE_GPU_ORDER
      b[2] = fma(x[0], x[1], -(p[2] * x[2]));
#else
      b[2] = x[0] * x[1] - p[2] * x[2];
#endif
      break;
    case BB_MODEL_BOLUS: { /* partialbridge_bolus3.jl:49,73  (α dose(t) - (λ+β) x1 + μ x2,  λ x1 - μ x2), dose(t) = 2(t/2)/(1+(t/2)^2) */
      double u = t / 2;
      double dose = (2 * u) / (1 + u * u);
      b[0] = (p[0] * dose - (p[2] + p[1]) * x[0]) + p[3] * x[1];
      b[1] = p[2] * x[0] - p[3] * x[1];
      break;
    }
    case BB_MODEL_LANDMARKS: { /* partialbridge_landmarks.jl:47 (kernel), :90-101 (b!), state = fll(Vector{Point}) */
      const int n = 4;
      const double a = p[0], lam = p[2];
      const double c0 = 1.0 / ((2 * M_PI) * a); /* 1/(2*π*P.a)^(length(x)/2), length(x) = 2 */
      const double c1 = 1.0 / (2 * a);
      const double nlh = (-lam) * 0.5;
      for (int i = 0; i < 4 * n; i++) b[i] = 0.0;
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          const double *qi = x + 4 * i, *pi = x + 4 * i + 2, *qj = x + 4 * j, *pj = x + 4 * j + 2;
          double dx = qi[0] - qj[0], dy = qi[1] - qj[1];
#ifdef ORACLE_GPU_ORDER
          double kij = c0 * bb_exp(-(fma(dy, dy, dx * dx) * c1)); /* |x|^2 directly, product with 1/(2a): as the kernels */
#else
          double nrm = sqrt(dy * dy + dx * dx); /* norm(x) */
          double kij = c0 * bb_exp(-(nrm * nrm) / (2 * a));
#endif
          double dot = MA(pi[1], pj[1], pi[0] * pj[0]);
          for (int k = 0; k < 2; k++) {
            b[4 * i + k] += (0.5 * pj[k]) * kij;
            double t1 = (nlh * pj[k]) * kij;
            double t2 = ((c1 * dot) * (qi[k] - qj[k])) * kij;
            b[4 * i + 2 + k] += t1 + t2;
          }
        }
      break;
    }
    default:
      for (int i = 0; i < P->d; i++) b[i] = NAN;
  }
}
/* sigma(t,x,P) as a dense d x d' matrix (all registry models have constant sigma) */
static void model_sigma(const bb_model* P, double* S) {
  int d = P->d, dp = P->dprime;
  const double* p = P->par;
  memset(S, 0, sizeof(double) * d * dp);
  switch (P->id) {
    case BB_MODEL_WIENER: for (int i = 0; i < d; i++) S[i * dp + i] = 1.0; break;
    case BB_MODEL_OU: S[0] = p[1]; break;
    case BB_MODEL_LINPRO: memcpy(S, p + d * d + d, sizeof(double) * d * d); break;
    case BB_MODEL_FHN_DIAG: S[0] = p[4]; S[3] = p[5]; break;
    case BB_MODEL_FHN_HYPO: S[1] = p[4]; break;
    case BB_MODEL_INTDIFF: S[1] = p[0]; break;
    case BB_MODEL_NCLAR3: S[2] = p[2]; break;
    case BB_MODEL_LORENZ: S[0] = p[3]; S[4] = p[4]; S[8] = p[5]; break;
    case BB_MODEL_BOLUS: S[0] = p[4]; S[3] = p[4]; break; /* σ = [σ1 0; 0 σ1]  bolus3.jl:50 */
    case BB_MODEL_LANDMARKS: /* noise on the momenta: component 4i+2+k <- column 2i+k  (partialbridge_landmarks.jl:111-118) */
      for (int i = 0; i < 4; i++)
        for (int k = 0; k < 2; k++) S[(4 * i + 2 + k) * dp + 2 * i + k] = p[1];
      break;
  }
}
static int model_sigma_is_sparse(const bb_model* P) { return P->id != BB_MODEL_LINPRO; }
/* a = sigma sigma'  (fallback a(t,x,P) = outer(σ), src/types.jl:32; LinPro stores P.a = σσ', linpro.jl:73) */
static void model_a(const bb_model* P, double* A) {
  double S[DM2], St[DM2];
  model_sigma(P, S);
  mat_tr(P->d, P->dprime, S, St);
  /* products with exact zeros: plain sums so that sparse sigma gives exact sigma_i^2 entries */
  int d = P->d, dp = P->dprime;
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) {
      double s = 0.0;
      for (int l = 0; l < dp; l++) s += S[i * dp + l] * St[l * d + j];
      A[i * d + j] = s;
    }
}
void bbo_model_a(const bb_model* P, double* A) { model_a(P, A); }
void bbo_model_b(const bb_model* P, double t, const double* x, double* b) { model_b(P, t, x, b); }

/* one Euler-Maruyama update  y <- (y + b*dt) + sigma*dw   src/euler.jl:148, _scale src/euler.jl:3-4 */
static void em_update(const bb_model* P, const double* S, const double* bdrift, double dt,
                      const double* dw, double* y) {
  int d = P->d, dp = P->dprime;
  if (model_sigma_is_sparse(P)) {
    /* scalar / UniformScaling / SDiagonal / column-vector sigma: one product per component */
    for (int i = 0; i < d; i++) {
      double t1 = MA(bdrift[i], dt, y[i]);
      double s = 0.0;
      int used = 0;
      for (int l = 0; l < dp; l++)
        if (S[i * dp + l] != 0.0) { /* the single structural non-zero of this row */
          t1 = MA(S[i * dp + l], dw[l], t1);
          used = 1;
        }
      (void)s;
      if (!used) t1 = t1 + 0.0 * dw[0]; /* (0, σ)' * dw: adding 0*dw keeps NaN/Inf propagation of the reference */
      y[i] = t1;
    }
  } else {
    double sd[DM];
    mat_vec(d, dp, S, dw, sd);
    for (int i = 0; i < d; i++) y[i] = MA(bdrift[i], dt, y[i]) + sd[i];
  }
}

/* ======================================================================= A2: solve!(EulerMaruyama(), Y, u, W, P)
 * src/euler.jl:135-152.  W [N][dp], X [N][d] (out).  Returns yy[N] in X[N-1]. */
void bbo_euler(const bb_model* P, int N, const double* tt, const double* u, const double* W,
               double* X) {
  int d = P->d, dp = P->dprime;
  double y[DM], b[DM], dw[DM], S[DM2];
  model_sigma(P, S);
  memcpy(y, u, sizeof(double) * d);
  for (int i = 0; i < N - 1; i++) {
    memcpy(X + (size_t)i * d, y, sizeof(double) * d);
    model_b(P, tt[i], y, b);
    for (int l = 0; l < dp; l++) dw[l] = W[(size_t)(i + 1) * dp + l] - W[(size_t)i * dp + l];
    em_update(P, S, b, tt[i + 1] - tt[i], dw, y);
  }
  memcpy(X + (size_t)(N - 1) * d, y, sizeof(double) * d); /* endpoint(y,P) = y  src/euler.jl:65 */
}

/* solve!(StochasticHeun(), Y, u, W, P)   src/euler.jl:178-198
 *   for i in 1:N-2 ("fix me" in the reference): yy[i] = y; B = b(t_i, y); y2 = y + B dt;
 *       y = y + 0.5 (b(t_{i+1}, y2) + B) dt + sigma (w_{i+1} - w_i)
 *   yy[N-1] = endpoint(y, P);  yy[N] is NOT written (X[N-1] keeps the caller's value). */
void bbo_heun(const bb_model* P, int N, const double* tt, const double* u, const double* W,
              double* X) {
  int d = P->d, dp = P->dprime;
  double y[DM], y2[DM], b[DM], b2[DM], hb[DM], dw[DM], S[DM2];
  model_sigma(P, S);
  memcpy(y, u, sizeof(double) * d);
  for (int i = 0; i < N - 2; i++) {
    memcpy(X + (size_t)i * d, y, sizeof(double) * d);
    double dt = tt[i + 1] - tt[i];
    model_b(P, tt[i], y, b);
    for (int k = 0; k < d; k++) y2[k] = MA(b[k], dt, y[k]);
    model_b(P, tt[i + 1], y2, b2);
    for (int k = 0; k < d; k++) hb[k] = 0.5 * (b2[k] + b[k]);
    for (int l = 0; l < dp; l++) dw[l] = W[(size_t)(i + 1) * dp + l] - W[(size_t)i * dp + l];
    em_update(P, S, hb, dt, dw, y);
  }
  if (N >= 2) memcpy(X + (size_t)(N - 2) * d, y, sizeof(double) * d);
}

/* ======================================================================= auxiliary process access */
static void aux_stage(const bb_aux* A, int i, int k, const double** B, const double** beta,
                      const double** a) {
  int d = A->d;
  if (A->is_const) {
    *B = A->B; *beta = A->beta; *a = A->a;
  } else {
    size_t e = (size_t)3 * i + k;
    *B = A->B + e * d * d; *beta = A->beta + e * d; *a = A->a + e * d * d;
  }
}
static const double* aux_a_left(const bb_aux* A, int i) {
  return A->is_const ? A->a : A->a_left + (size_t)i * A->d * A->d;
}

/* kernelr3 skeleton (src/ode.jl:44-49): caller supplies f evaluated through a callback taking the stage */
typedef void (*rhs_fn)(void* ctx, int stage, const double* y, double* k);
static void r3_step(int n, rhs_fn f, void* ctx, double* y, double h) {
  double k1[DM2], k2[DM2], k3[DM2], yt[DM2];
  f(ctx, 0, y, k1);
  double c2 = 1.0 / 2 * h; /* 1/2*dt */
  for (int i = 0; i < n; i++) yt[i] = MA(c2, k1[i], y[i]);
  f(ctx, 1, yt, k2);
  double c3 = 3.0 / 4 * h; /* 3/4*dt */
  for (int i = 0; i < n; i++) yt[i] = MA(c3, k2[i], y[i]);
  f(ctx, 2, yt, k3);
  const double w1 = 2.0 / 9, w2 = 1.0 / 3, w3 = 4.0 / 9;
  for (int i = 0; i < n; i++) {
    double s = w1 * k1[i];
    s = MA(w2, k2[i], s);
    s = MA(w3, k3[i], s);
    y[i] = MA(h, s, y[i]); /* y + dt*(2/9*k1 + 1/3*k2 + 4/9*k3) */
  }
}

typedef struct {
  const bb_aux* A;
  int i;          /* interval */
  const double* H; /* extra (FH variant) */
  const double* L; /* extra (LMmu variant) */
  int m;
} rhs_ctx;

/* b̃(t,y) = B y + β   src/partialbridgenuH.jl:27, src/gode.jl:2 */
static void rhs_btilde(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d;
  mat_vec(d, d, B, y, k);
  for (int i = 0; i < d; i++) k[i] += be[i];
}
/* dH⁺(t,y) = B y + (B y)' - a   src/partialbridgenuH.jl:28 */
static void rhs_dHplus(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d;
  double BY[DM2];
  mat_mul(d, d, d, B, y, BY);
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) k[i * d + j] = (BY[i * d + j] + BY[j * d + i]) - a[i * d + j];
}
/* _dHinv(t,K) = B K + K B' - a   src/gode.jl:3 */
static void rhs_dHinv(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d;
  double BK[DM2], Bt[DM2], KBt[DM2];
  mat_mul(d, d, d, B, y, BK);
  mat_tr(d, d, B, Bt);
  mat_mul(d, d, d, y, Bt, KBt);
  for (int i = 0; i < d * d; i++) k[i] = (BK[i] + KBt[i]) - a[i];
}
/* dH(t,y) = -B'y - yB + y a y'   src/partialbridgenuH.jl:68 */
static void rhs_dH(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d;
  double Bt[DM2], nBt[DM2], T1[DM2], T2[DM2], T3[DM2], yt[DM2];
  mat_tr(d, d, B, Bt);
  for (int i = 0; i < d * d; i++) nBt[i] = -Bt[i];
  mat_mul(d, d, d, nBt, y, T1); /* -B'*y */
  mat_mul(d, d, d, y, B, T2);   /* y*B */
  mat_mul(d, d, d, y, a, T3);   /* y*a */
  mat_tr(d, d, y, yt);
  mat_mul(d, d, d, T3, yt, T3); /* y*a*y' */
  for (int i = 0; i < d * d; i++) k[i] = (T1[i] - T2[i]) + T3[i];
}
/* dF(t,y,(H,P)) = -B'y + H a y + H β   src/partialbridgenuH.jl:69 */
static void rhs_dF(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d;
  double Bt[DM2], nBt[DM2], Ha[DM2], t1[DM], t2[DM], t3[DM];
  mat_tr(d, d, B, Bt);
  for (int i = 0; i < d * d; i++) nBt[i] = -Bt[i];
  mat_vec(d, d, nBt, y, t1);
  mat_mul(d, d, d, c->H, a, Ha);
  mat_vec(d, d, Ha, y, t2);
  mat_vec(d, d, c->H, be, t3);
  for (int i = 0; i < d; i++) k[i] = (t1[i] + t2[i]) + t3[i];
}
/* L' = -y*B   src/partialbridge.jl:13 */
static void rhs_dL(void* c_, int st, const double* y, double* k) {
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d, m = c->m;
  double ny[DM2];
  for (int i = 0; i < m * d; i++) ny[i] = -y[i];
  mat_mul(m, d, d, ny, B, k);
}
/* M⁺' = -outer(L*σ(t,P)) = -(Lσ)(Lσ)' = -L a L'   src/partialbridge.jl:14 (a = σσ' is what the ABI carries) */
static void rhs_dMplus(void* c_, int st, const double* y, double* k) {
  (void)y;
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d, m = c->m;
  double La[DM2], Lt[DM2], T[DM2];
  mat_mul(m, d, d, c->L, a, La);
  mat_tr(m, d, c->L, Lt);
  mat_mul(m, d, m, La, Lt, T);
  for (int i = 0; i < m * m; i++) k[i] = -T[i];
}
/* μ' = -L*β   src/partialbridge.jl:15 */
static void rhs_dmu(void* c_, int st, const double* y, double* k) {
  (void)y;
  rhs_ctx* c = (rhs_ctx*)c_;
  const double *B, *be, *a;
  aux_stage(c->A, c->i, st, &B, &be, &a);
  int d = c->A->d, m = c->m;
  double nL[DM2];
  for (int i = 0; i < m * d; i++) nL[i] = -c->L[i];
  mat_vec(m, d, nL, be, k);
}

/* ======================================================================= A5: updateνH⁺C  src/partialbridgenuH.jl:1-17 */
int bbo_update_nuHC(int d, int m, const double* L, const double* Sigma, const double* v,
                    double eps, double* nu, double* Hplus, double* C) {
  double Si[DM2], Lt[DM2], LtSi[DM2], H[DM2], Siv[DM], t[DM];
  if (mat_inv(m, Sigma, Si)) return BB_ERR_SINGULAR;
  mat_tr(m, d, L, Lt);
  mat_mul(d, m, m, Lt, Si, LtSi);  /* L'*inv(Σ) */
  mat_mul(d, m, d, LtSi, L, H);    /* L'*inv(Σ)*L */
  for (int i = 0; i < d; i++) H[i * d + i] += eps; /* + ϵ*I */
  if (mat_inv(d, H, Hplus)) return BB_ERR_SINGULAR;
  double HpLt[DM2], HpLtSi[DM2];
  mat_mul(d, d, m, Hplus, Lt, HpLt);
  mat_mul(d, m, m, HpLt, Si, HpLtSi);
  mat_vec(d, m, HpLtSi, v, nu);    /* ν = H⁺*L'*inv(Σ)*v */
  /* updateC: C = 0.5*dot(v, Σ\v) + m/2*log(2π) + 0.5*logdet(Σ)   :9-14 */
  mat_vec(m, m, Si, v, Siv);
  double c = 0.0;
  c += 0.5 * vdot(m, v, Siv);
  double ld;
  if (m == 1) ld = log(Sigma[0]);
  else {
    double Lc[DM2];
    if (chol_lower(m, Sigma, Lc)) return BB_ERR_SINGULAR;
    ld = 0;
    for (int i = 0; i < m; i++) ld += 2 * log(Lc[i * m + i]);
  }
  c += m / 2.0 * log(2 * M_PI) + 0.5 * ld;
  *C = c;
  (void)t;
  return 0;
}
/* updateFHC(L, Σ, v, F, H, ϵ, C)   src/partialbridgenuH.jl:57-62 */
int bbo_update_FHC(int d, int m, const double* L, const double* Sigma, const double* v,
                   double* F, double* H, double eps, double* C) {
  double Si[DM2], Lt[DM2], LtSi[DM2], T[DM2], t[DM];
  if (mat_inv(m, Sigma, Si)) return BB_ERR_SINGULAR;
  mat_tr(m, d, L, Lt);
  mat_mul(d, m, m, Lt, Si, LtSi);
  mat_mul(d, m, d, LtSi, L, T);
  for (int i = 0; i < d; i++) T[i * d + i] += eps;
  for (int i = 0; i < d * d; i++) H[i] += T[i];
  mat_vec(d, m, LtSi, v, t);
  for (int i = 0; i < d; i++) F[i] += t[i];
  double nu[DM], Hp[DM2], c;
  int rc = bbo_update_nuHC(d, m, L, Sigma, v, 1.0, nu, Hp, &c); /* only for updateC's value */
  if (rc) return rc;
  *C += c;
  return 0;
}

/* observation update of (ν, H⁺)   project_partialbridge/partialbridge_bolus3.jl:128-137 */
int bbo_gpupdate_nuH(int d, int m, double* nu, double* Hplus, const double* L,
                     const double* Sigma, const double* v) {
  int allinf = 1;
  for (int i = 0; i < d; i++)
    if (!(isinf(Hplus[i * d + i]) && Hplus[i * d + i] > 0)) allinf = 0;
  double Si[DM2], Lt[DM2];
  if (mat_inv(m, Sigma, Si)) return BB_ERR_SINGULAR;
  mat_tr(m, d, L, Lt);
  if (allinf) {
    double LtSi[DM2], A[DM2], Ai[DM2], rhs[DM];
    mat_mul(d, m, m, Lt, Si, LtSi);
    mat_mul(d, m, d, LtSi, L, A);
    if (mat_inv(d, A, Ai)) return BB_ERR_SINGULAR;
    mat_vec(d, m, LtSi, v, rhs);
    mat_vec(d, d, Ai, rhs, nu);
    memcpy(Hplus, Ai, sizeof(double) * d * d);
    return 0;
  }
  double HLt[DM2], LHLt[DM2], G[DM2], Gi[DM2], T[DM2], Z[DM2], ZH[DM2], t1[DM], t2[DM];
  mat_mul(d, d, m, Hplus, Lt, HLt);          /* H⁺L' */
  mat_mul(m, d, m, L, HLt, LHLt);            /* L H⁺ L' */
  for (int i = 0; i < m * m; i++) G[i] = Sigma[i] + LHLt[i];
  if (mat_inv(m, G, Gi)) return BB_ERR_SINGULAR;
  mat_mul(d, m, m, HLt, Gi, T);              /* H⁺L' inv(Σ + LH⁺L') */
  mat_mul(d, m, d, T, L, Z);
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) Z[i * d + j] = ((i == j) ? 1.0 : 0.0) - Z[i * d + j];
  mat_mul(d, d, d, Z, Hplus, ZH);            /* Z H⁺ */
  double ZHLt[DM2], ZHLtSi[DM2];
  mat_mul(d, d, m, ZH, Lt, ZHLt);
  mat_mul(d, m, m, ZHLt, Si, ZHLtSi);
  mat_vec(d, m, ZHLtSi, v, t1);              /* Z H⁺ L' inv(Σ) v */
  mat_vec(d, d, Z, nu, t2);                  /* Z ν */
  for (int i = 0; i < d; i++) nu[i] = t1[i] + t2[i];
  memcpy(Hplus, ZH, sizeof(double) * d * d);
  return 0;
}
/* Bridge.gpupdate(H♢, V, L, Σ, v)  src/guip.jl:221-243 -- same algebra on (H♢, V) */
int bbo_gpupdate_HV(int d, int m, double* Hdia, double* V, const double* L, const double* Sigma,
                    const double* v) {
  return bbo_gpupdate_nuH(d, m, V, Hdia, L, Sigma, v);
}

/* ======================================================================= A5: partialbridgeodeνH!
 * R3 variant src/partialbridgenuH.jl:21-55; Lyap variant :86-103 with src/lyap.jl:2-6.
 * nu [N][d], H [N][d][d] out; (nu_left, Hplus_left, C) the state at tt[0]. */
int bbo_backward_nuH(int method, int N, int d, const double* tt, const bb_aux* A,
                     const double* nu_end, const double* Hplus_end, double C0, double* nu,
                     double* H, double* nu_left, double* Hplus_left, double* C) {
  double Hp[DM2], Hc[DM2], v[DM];
  memcpy(Hp, Hplus_end, sizeof(double) * d * d);
  memcpy(v, nu_end, sizeof(double) * d);
  if (mat_inv(d, Hp, Hc)) return BB_ERR_SINGULAR;
  memcpy(H + (size_t)(N - 1) * d * d, Hc, sizeof(double) * d * d); /* Ht[end] = H = inv(H⁺) */
  memcpy(nu + (size_t)(N - 1) * d, v, sizeof(double) * d);         /* νt[end] = ν */
  double Cc = C0;
  rhs_ctx c;
  c.A = A; c.H = 0; c.L = 0; c.m = 0;
  for (int i = N - 2; i >= 0; i--) {
    double dt = tt[i] - tt[i + 1];
    c.i = i;
    const double *B0, *be0, *a0;
    aux_stage(A, i, 0, &B0, &be0, &a0); /* values at t[i+1] */
    if (method == BB_ODE_R3) {
      r3_step(d * d, rhs_dHplus, &c, Hp, dt);            /* :40 */
      double F[DM], aF[DM];
      mat_vec(d, d, Hc, v, F);                           /* :45  F = H*ν (old H, old ν) */
      mat_vec(d, d, a0, F, aF);
      double dC = (vdot(d, be0, F) + 0.5 * vdot(d, F, aF)) - 0.5 * mat_trace_prod(d, Hc, a0); /* :31 */
      Cc += dC * dt;                                     /* :46 */
      r3_step(d, rhs_btilde, &c, v, dt);                 /* :49 */
    } else {
      double F[DM], aF[DM];
      mat_vec(d, d, Hc, v, F);                           /* :97  F = Ht[i+1]*νt[i+1] */
      r3_step(d, rhs_btilde, &c, v, dt);                 /* :95 */
      /* lyapunovpsdbackward_step(t[i+1], H⁺, -dt, P)    src/lyap.jl:2-6 */
      double h = -dt;
      const double *B1, *be1, *a1;
      aux_stage(A, i, 1, &B1, &be1, &a1);                /* B(t - h/2) */
      double Pm[DM2], Mm[DM2], Pi[DM2], phi[DM2], phit[DM2], Y[DM2], T[DM2];
      double hh = 1.0 / 2 * h;
      for (int r = 0; r < d; r++)
        for (int s = 0; s < d; s++) {
          double id = (r == s) ? 1.0 : 0.0;
          Pm[r * d + s] = id + hh * B1[r * d + s];
          Mm[r * d + s] = id - hh * B1[r * d + s];
        }
      if (mat_inv(d, Pm, Pi)) return BB_ERR_SINGULAR;
      mat_mul(d, d, d, Pi, Mm, phi);                     /* ϕ = (I + ½hB)\(I - ½hB) */
      const double* al = aux_a_left(A, i);               /* a(t - h) */
      for (int r = 0; r < d * d; r++) Y[r] = MA(hh, al[r], Hp[r]);
      mat_mul(d, d, d, phi, Y, T);
      mat_tr(d, d, phi, phit);
      mat_mul(d, d, d, T, phit, T);
      for (int r = 0; r < d * d; r++) Hp[r] = MA(hh, a0[r], T[r]); /* ... + ½h a(t) */
      mat_vec(d, d, a0, F, aF);
      Cc += (vdot(d, be0, F) * dt + 0.5 * vdot(d, F, aF) * dt) - 0.5 * mat_trace_prod(d, Hc, a0) * dt; /* :98 */
    }
    if (mat_inv(d, Hp, Hc)) return BB_ERR_SINGULAR;      /* :50 / :100 */
    memcpy(nu + (size_t)i * d, v, sizeof(double) * d);
    memcpy(H + (size_t)i * d * d, Hc, sizeof(double) * d * d);
  }
  if (nu_left) memcpy(nu_left, v, sizeof(double) * d);
  if (Hplus_left) memcpy(Hplus_left, Hp, sizeof(double) * d * d);
  if (C) *C = Cc;
  return 0;
}

/* partialbridgeodeHνH!(R3(), t, Ft, Ht, P, (F, H, C))   src/partialbridgenuH.jl:64-81 */
int bbo_backward_FH(int N, int d, const double* tt, const bb_aux* A, const double* F_end,
                    const double* H_end, double C0, double* F, double* H, double* C) {
  double Hc[DM2], Fc[DM];
  memcpy(Hc, H_end, sizeof(double) * d * d);
  memcpy(Fc, F_end, sizeof(double) * d);
  memcpy(H + (size_t)(N - 1) * d * d, Hc, sizeof(double) * d * d);
  memcpy(F + (size_t)(N - 1) * d, Fc, sizeof(double) * d);
  double Cc = C0;
  rhs_ctx c;
  c.A = A; c.L = 0; c.m = 0;
  for (int i = N - 2; i >= 0; i--) {
    double dt = tt[i] - tt[i + 1];
    c.i = i;
    const double *B0, *be0, *a0;
    aux_stage(A, i, 0, &B0, &be0, &a0);
    double aF[DM];
    mat_vec(d, d, a0, Fc, aF);
    Cc += (vdot(d, be0, Fc) * dt + 0.5 * vdot(d, Fc, aF) * dt) - 0.5 * mat_trace_prod(d, Hc, a0) * dt; /* :73 */
    r3_step(d * d, rhs_dH, &c, Hc, dt); /* :74 */
    c.H = Hc;
    r3_step(d, rhs_dF, &c, Fc, dt);     /* :75 (fresh H) */
    memcpy(F + (size_t)i * d, Fc, sizeof(double) * d);
    memcpy(H + (size_t)i * d * d, Hc, sizeof(double) * d * d);
  }
  *C = Cc;
  return 0;
}

/* A4: GuidedBridge tables: gpHinv!(H♢, Pt, h♢), gpV!(V, Pt, v)   src/guip.jl:172-180, src/gode.jl:13,21,
 * _solvebackward! src/ode.jl:88-97 */
int bbo_backward_HV(int N, int d, const double* tt, const bb_aux* A, const double* v,
                    const double* hdia_end, double* Hdia, double* V) {
  double K[DM2], Vc[DM];
  if (hdia_end) memcpy(K, hdia_end, sizeof(double) * d * d);
  else memset(K, 0, sizeof(double) * d * d);
  memcpy(Vc, v, sizeof(double) * d);
  memcpy(Hdia + (size_t)(N - 1) * d * d, K, sizeof(double) * d * d);
  memcpy(V + (size_t)(N - 1) * d, Vc, sizeof(double) * d);
  rhs_ctx c;
  c.A = A; c.H = 0; c.L = 0; c.m = 0;
  for (int i = N - 2; i >= 0; i--) {
    double dt = tt[i] - tt[i + 1];
    c.i = i;
    r3_step(d * d, rhs_dHinv, &c, K, dt);
    r3_step(d, rhs_btilde, &c, Vc, dt);
    memcpy(Hdia + (size_t)i * d * d, K, sizeof(double) * d * d);
    memcpy(V + (size_t)i * d, Vc, sizeof(double) * d);
  }
  return 0;
}

/* A6: partialbridgeode!(R3(), t, L, Σ, Lt, Mt, μt, P)   src/partialbridge.jl:1-22 */
int bbo_backward_LMmu(int N, int d, int m, const double* tt, const bb_aux* A, const double* L,
                      const double* Sigma, double* Lt, double* Mt, double* mut) {
  double Lc[DM2], Mp[DM2], Mi[DM2], mu[DM];
  memcpy(Lc, L, sizeof(double) * m * d);
  memcpy(Mp, Sigma, sizeof(double) * m * m);
  memset(mu, 0, sizeof(double) * m);
  if (mat_inv(m, Sigma, Mi)) return BB_ERR_SINGULAR;
  memcpy(Lt + (size_t)(N - 1) * m * d, Lc, sizeof(double) * m * d);
  memcpy(Mt + (size_t)(N - 1) * m * m, Mi, sizeof(double) * m * m);
  memcpy(mut + (size_t)(N - 1) * m, mu, sizeof(double) * m);
  rhs_ctx c;
  c.A = A; c.H = 0; c.m = m;
  for (int i = N - 2; i >= 0; i--) {
    double dt = tt[i] - tt[i + 1];
    c.i = i;
    r3_step(m * d, rhs_dL, &c, Lc, dt);       /* :13 */
    c.L = Lc;                                  /* the freshly updated L  :14-15 */
    r3_step(m * m, rhs_dMplus, &c, Mp, dt);
    r3_step(m, rhs_dmu, &c, mu, dt);
    if (mat_inv(m, Mp, Mi)) return BB_ERR_SINGULAR;
    memcpy(Lt + (size_t)i * m * d, Lc, sizeof(double) * m * d);
    memcpy(Mt + (size_t)i * m * m, Mi, sizeof(double) * m * m);
    memcpy(mut + (size_t)i * m, mu, sizeof(double) * m);
  }
  return 0;
}

/* lptilde(x, P::PartialBridgeνH) in the form the reference tests: -0.5*(x'*H[1]*x - 2*x'*H[1]*ν[1]) - C
 * (test/partialbridgenuH.jl:124; src/partialbridgenuH.jl:169 has a typo -- P.ν instead of P.ν[1] -- and does not run) */
double bbo_lptilde_nuH(int d, const double* nu0, const double* H0, double C, const double* x) {
  double Hx[DM], Hnu[DM];
  mat_vec(d, d, H0, x, Hx);
  mat_vec(d, d, H0, nu0, Hnu);
  return -0.5 * (vdot(d, x, Hx) - 2 * vdot(d, x, Hnu)) - C;
}
/* lptilde(P::GuidedBridge, u) = logpdfnormal(P.V[1] - u, P.H♢[1]) - traceB(P.tt, P.Pt)   src/guip.jl:203-206
 * traceB(tt, P) = solve(R3(), _traceB, tt, 0.0, P), _traceB(t, x, P) = tr(B(t, P))         src/ode.jl:178-184
 * trB: tr B at the stage times tt[i], tt[i] + h/2, tt[i] + 3h/4 of every forward interval, or one constant */
typedef struct { const double* trB; int is_const; int i; } trace_ctx;
static void rhs_trace(void* c_, int st, const double* y, double* k) {
  (void)y;
  trace_ctx* c = (trace_ctx*)c_;
  k[0] = c->is_const ? c->trB[0] : c->trB[3 * c->i + st];
}
double bbo_lptilde_HV(int N, int d, const double* tt, const double* trB, int trB_const, const double* V0,
                      const double* Hdia0, const double* u) {
  double y = 0.0, e[DM];
  trace_ctx c = {trB, trB_const, 0};
  for (int i = 0; i < N - 1; i++) {
    c.i = i;
    r3_step(1, rhs_trace, &c, &y, tt[i + 1] - tt[i]);
  }
  for (int q = 0; q < d; q++) e[q] = V0[q] - u[q];
  return bbo_logpdfnormal(d, e, Hdia0) - y;
}

/* ======================================================================= guided proposal (forward) */
typedef struct {
  int32_t kind, N, d, m;
  const double* tt;                /* [N] */
  const double* A;                 /* NUH: H; HV: H♢; LMMU: L */
  const double* b;                 /* NUH: ν; HV: V;  LMMU: μ */
  const double* Mm;                /* LMMU: M */
  const double* v;                 /* LMMU: v */
  const double* Bt; const double* betat; /* auxiliary drift on the grid */
  int32_t aux_const;
  const double* Adiff;             /* a(Target) - a~ ([d*d] or [N][d*d]); NULL: constdiff(P°)  */
  int32_t adiff_const;
} bbo_guide;

/* r((i,t),x,P°):  νH src/partialbridgenuH.jl:161; GuidedBridge src/guip.jl:193; PartialBridge src/partialbridge.jl:57 */
static void guide_r(const bbo_guide* G, int i, const double* x, double* r) {
  int d = G->d, m = G->m;
  double e[DM];
  if (G->kind == BB_GUIDE_NUH) {
    const double *H = G->A + (size_t)i * d * d, *nu = G->b + (size_t)i * d;
    for (int k = 0; k < d; k++) e[k] = nu[k] - x[k];
    mat_vec(d, d, H, e, r);
  } else if (G->kind == BB_GUIDE_HV) {
    const double *K = G->A + (size_t)i * d * d, *V = G->b + (size_t)i * d;
    for (int k = 0; k < d; k++) e[k] = V[k] - x[k];
#ifdef ORACLE_GPU_ORDER
    double Ki[DM2];
    mat_inv(d, K, Ki);
    mat_vec(d, d, Ki, e, r);
#else
    /* H♢[i] \ (V[i]-x): StaticArrays solves d<=3 in closed form */
    if (d == 1) r[0] = e[0] / K[0];
    else if (d == 2) {
      double det = K[0] * K[3] - K[1] * K[2];
      r[0] = (K[3] * e[0] - K[1] * e[1]) / det;
      r[1] = (K[0] * e[1] - K[2] * e[0]) / det;
    } else {
      double Ki[DM2];
      mat_inv(d, K, Ki);
      mat_vec(d, d, Ki, e, r);
    }
#endif
  } else {
    const double *L = G->A + (size_t)i * m * d, *mu = G->b + (size_t)i * m,
                 *M = G->Mm + (size_t)i * m * m;
    double Lx[DM], Lt[DM2], LtM[DM2];
    mat_vec(m, d, L, x, Lx);
    for (int k = 0; k < m; k++) e[k] = (G->v[k] - mu[k]) - Lx[k]; /* P.v - P.μ[i] - P.L[i]*x */
    mat_tr(m, d, L, Lt);
    mat_mul(d, m, m, Lt, M, LtM);                                  /* P.L[i]'*P.M[i] */
    mat_vec(d, m, LtM, e, r);
  }
}
/* a(t,x,Target)*r with the structural zeros of the registry models */
static void model_a_r(const bb_model* P, const double* Amat, const double* r, double* ar) {
  int d = P->d;
  if (model_sigma_is_sparse(P)) {
    for (int i = 0; i < d; i++) ar[i] = Amat[i * d + i] * r[i]; /* a is diagonal for these models */
  } else {
    mat_vec(d, d, Amat, r, ar);
  }
}
/* _b((i,t),x,P°) = b(t,x,Target) + a(t,x,Target)*r   src/partialbridgenuH.jl:157-159, guip.jl:192, partialbridge.jl:53-55 */
static void guided_b(const bb_model* P, const double* Amat, const bbo_guide* G, int i, double t,
                     const double* x, double* bo, double* r) {
  int d = P->d;
  double b[DM], ar[DM];
  model_b(P, t, x, b);
  guide_r(G, i, x, r);
  if (model_sigma_is_sparse(P)) {
    for (int k = 0; k < d; k++) bo[k] = MA(Amat[k * d + k], r[k], b[k]);
  } else {
    model_a_r(P, Amat, r, ar);
    for (int k = 0; k < d; k++) bo[k] = b[k] + ar[k];
  }
}

/* solve!(Euler(), Y, u, W, P°)   src/euler.jl:247-268.  xend = yy[N] (the return value). */
void bbo_guided_euler(const bb_model* P, const bbo_guide* G, const double* u, const double* W,
                      double* X, double* xend) {
  int d = P->d, dp = P->dprime, N = G->N;
  const double* tt = G->tt; /* tt[:] = P.tt  :256 */
  double y[DM], bo[DM], r[DM], dw[DM], S[DM2], Amat[DM2];
  model_sigma(P, S);
  model_a(P, Amat);
  memcpy(y, u, sizeof(double) * d);
  for (int i = 0; i < N - 1; i++) {
    if (X) memcpy(X + (size_t)i * d, y, sizeof(double) * d);
    guided_b(P, Amat, G, i, tt[i], y, bo, r);
    for (int l = 0; l < dp; l++) dw[l] = W[(size_t)(i + 1) * dp + l] - W[(size_t)i * dp + l];
    em_update(P, S, bo, tt[i + 1] - tt[i], dw, y);
  }
  /* endpoint(y, P::GuidedBridge) = norm(P.H♢[end],1) < eps() ? P.V[end] : y   src/euler.jl:241-242 */
  if (G->kind == BB_GUIDE_HV) {
    const double* K = G->A + (size_t)(N - 1) * d * d;
    double n1 = 0; /* opnorm-1 for matrices: max column sum; norm(.,1) of an SMatrix is the entrywise 1-norm */
    for (int k = 0; k < d * d; k++) n1 += fabs(K[k]);
    if (n1 < 2.220446049250313e-16) memcpy(y, G->b + (size_t)(N - 1) * d, sizeof(double) * d);
  }
  if (X) memcpy(X + (size_t)(N - 1) * d, y, sizeof(double) * d);
  if (xend) memcpy(xend, y, sizeof(double) * d);
}

/* solve!(Mdb(), Y, u, W, P°)   src/euler.jl:308-327 with P° a guided proposal:
 *   y = y + _b((i,tt[i]), y, P)*(tt[i+1]-tt[i]) + _scale(ww[i+1]-ww[i], σ(tt[i], y, P)*sqrt((tt[end]-tt[i+1])/(tt[end]-tt[i]))) */
void bbo_guided_mdb(const bb_model* P, const bbo_guide* G, const double* u, const double* W, double* X, double* xend) {
  int d = P->d, dp = P->dprime, N = G->N;
  const double* tt = G->tt;
  double y[DM], bo[DM], r[DM], dw[DM], S[DM2], Ss[DM2], Amat[DM2];
  model_sigma(P, S);
  model_a(P, Amat);
  memcpy(y, u, sizeof(double) * d);
  for (int i = 0; i < N - 1; i++) {
    if (X) memcpy(X + (size_t)i * d, y, sizeof(double) * d);
    guided_b(P, Amat, G, i, tt[i], y, bo, r);
    for (int l = 0; l < dp; l++) dw[l] = W[(size_t)(i + 1) * dp + l] - W[(size_t)i * dp + l];
    const double ns = sqrt((tt[N - 1] - tt[i + 1]) / (tt[N - 1] - tt[i]));
    for (int q = 0; q < d * dp; q++) Ss[q] = S[q] * ns; /* σ*sqrt(...): exact zeros stay zero */
    em_update(P, Ss, bo, tt[i + 1] - tt[i], dw, y);
  }
  if (G->kind == BB_GUIDE_HV) {
    const double* K = G->A + (size_t)(N - 1) * d * d;
    double n1 = 0;
    for (int k = 0; k < d * d; k++) n1 += fabs(K[k]);
    if (n1 < 2.220446049250313e-16) memcpy(y, G->b + (size_t)(N - 1) * d, sizeof(double) * d);
  }
  if (X) memcpy(X + (size_t)(N - 1) * d, y, sizeof(double) * d);
  if (xend) memcpy(xend, y, sizeof(double) * d);
}

/* llikelihood(LeftRule(), X, P°; skip)   src/partialbridgenuH.jl:171-189, guip.jl:429-446, partialbridge.jl:67-87
 * (constdiff branch; b̃ = B̃(tt[i]) x + β̃(tt[i])) */
double bbo_llikelihood(const bb_model* P, const bbo_guide* G, const double* X, int skip) {
  int d = P->d, N = G->N;
  const double* tt = G->tt;
  double som = 0.0;
  for (int i = 0; i < N - 1 - skip; i++) {
    const double* x = X + (size_t)i * d;
    double r[DM], b[DM], bt[DM];
    guide_r(G, i, x, r);
    model_b(P, tt[i], x, b);
    const double* Bt = G->aux_const ? G->Bt : G->Bt + (size_t)i * d * d;
    const double* be = G->aux_const ? G->betat : G->betat + (size_t)i * d;
    mat_vec(d, d, Bt, x, bt);
    double e[DM];
    for (int k = 0; k < d; k++) e[k] = b[k] - (bt[k] + be[k]);
    som = MA(vdot(d, e, r), tt[i + 1] - tt[i], som);
    if (G->Adiff) {
      /* if !constdiff(Po): H = H((i,s),x,Po); A = a(target) - a(aux);
       *   som -= 0.5*tr(A*H)*dt;  som += 0.5*(r'*A*r)*dt          src/partialbridge.jl:79-84 */
      const double* A = G->adiff_const ? G->Adiff : G->Adiff + (size_t)i * d * d;
      double dt = tt[i + 1] - tt[i];
      double H[DM2], rA[DM];
      int m = G->m;
      if (G->kind == BB_GUIDE_NUH) memcpy(H, G->A + (size_t)i * d * d, sizeof(double) * d * d);
      else if (G->kind == BB_GUIDE_HV) mat_inv(d, G->A + (size_t)i * d * d, H);
      else {
        const double *L = G->A + (size_t)i * m * d, *M = G->Mm + (size_t)i * m * m;
        double Lt[DM2], LtM[DM2];
        mat_tr(m, d, L, Lt);
        mat_mul(d, m, m, Lt, M, LtM);
        mat_mul(d, m, d, LtM, L, H); /* L'*M*L   src/partialbridge.jl:58 */
      }
      double trAH;
      {
        double s2 = 0;
        for (int rr = 0; rr < d; rr++) {
          double s3 = A[rr * d] * H[rr];
          for (int l = 1; l < d; l++) s3 = MA(A[rr * d + l], H[l * d + rr], s3);
          s2 = (rr == 0) ? s3 : s2 + s3;
        }
        trAH = s2;
      }
      for (int j = 0; j < d; j++) { /* r'*A */
        double q = r[0] * A[j];
        for (int l = 1; l < d; l++) q = MA(r[l], A[l * d + j], q);
        rA[j] = q;
      }
#ifdef ORACLE_GPU_ORDER
      som = fma(-(0.5 * trAH), dt, som);
      som = fma(0.5 * vdot(d, rA, r), dt, som);
#else
      som -= 0.5 * trAH * dt;
      som += 0.5 * vdot(d, rA, r) * dt;
#endif
    }
  }
  return som;
}

/* innovations!(EulerMaruyama(), W, Y, P)   src/euler.jl:358-376; G may be NULL (unguided) */
int bbo_innovations(const bb_model* P, const bbo_guide* G, int N, const double* tt,
                    const double* X, double* W) {
  int d = P->d, dp = P->dprime;
  if (d != dp) return BB_ERR_UNSUPPORTED;
  double S[DM2], Si[DM2], Amat[DM2], w[DM], b[DM], r[DM], e[DM], de[DM];
  model_sigma(P, S);
  model_a(P, Amat);
  if (mat_inv(d, S, Si)) return BB_ERR_SINGULAR;
  memset(w, 0, sizeof(double) * d);
  for (int i = 0; i < N - 1; i++) {
    memcpy(W + (size_t)i * d, w, sizeof(double) * d);
    const double* x = X + (size_t)i * d;
    const double* xn = X + (size_t)(i + 1) * d;
    if (G) guided_b(P, Amat, G, i, tt[i], x, b, r);
    else model_b(P, tt[i], x, b);
    double dt = tt[i + 1] - tt[i];
    for (int k = 0; k < d; k++) e[k] = (xn[k] - x[k]) - b[k] * dt;
    mat_vec(d, d, Si, e, de);
    for (int k = 0; k < d; k++) w[k] = w[k] + de[k];
  }
  memcpy(W + (size_t)(N - 1) * d, w, sizeof(double) * d);
  return 0;
}

/* The noise half of a pCN proposal for ONE segment (test/partialbridgenuH.jl:176-178, partialbridge_bolus3.jl:304-305):
 * sample!(W2, Wiener()) with W2.yy[1] = 0 drawn from noise row `row`, then Wo.yy .= ρ*W.yy + sqrt(1-ρ^2)*W2.yy. */
void bbo_pcn_combine(int N, int dp, const double* tt, const double* wc, double rho, uint64_t seed, uint32_t iter,
                     uint64_t row, double* wo) {
  double rho2 = sqrt(1 - rho * rho);
  double w2[DM];
  for (int k = 0; k < dp; k++) {
    w2[k] = 0.0;
    wo[k] = MA(rho2, w2[k], rho * wc[k]);
  }
  for (int j = 1; j < N; j++) {
    double rootdt = sqrt(tt[j] - tt[j - 1]);
    for (int k = 0; k < dp; k++) {
      double xi = bbo_normal(seed, iter, row, (uint64_t)j * dp + k);
      w2[k] = MA(rootdt, xi, w2[k]);
      wo[j * dp + k] = MA(rho2, w2[k], rho * wc[j * dp + k]);
    }
  }
}

/* ======================================================================= A8: one pCN iteration of one chain
 * test/partialbridgenuH.jl:176-191; multi-segment block with one accept: bolus3.jl:300-355.
 * Wc  [S][N][dp]  current driving paths      (in)
 * Wo  [S][N][dp]  proposal W° = ρW + sqrt(1-ρ²)W2   (out)
 * Xo  [S][N][d]   proposal path (out, may be NULL)
 * returns ll° (sum over segments), logU; accept decision is  logU <= ll° - ll  (caller). */
double bbo_pcn_propose(const bb_model* P, const bbo_guide* const* G, int S, const double* u,
                       const double* Wc, double rho, uint64_t seed, uint32_t iter,
                       uint64_t chain, int skip, double* Wo, double* Xo, double* xend,
                       double* logU) {
  int d = P->d, dp = P->dprime;
  double start[DM], end[DM];
  memcpy(start, u, sizeof(double) * d);
  double ll = 0.0;
  for (int s = 0; s < S; s++) {
    int N = G[s]->N;
    const double* tt = G[s]->tt;
    const double* wc = Wc + (size_t)s * N * dp;
    double* wo = Wo + (size_t)s * N * dp;
    bbo_pcn_combine(N, dp, tt, wc, rho, seed, iter, chain * (uint64_t)S + s, wo);
    double* xo = Xo ? Xo + (size_t)s * N * d : NULL;
    double* tmp = NULL;
    if (!xo) { tmp = (double*)malloc(sizeof(double) * N * d); xo = tmp; }
    bbo_guided_euler(P, G[s], start, wo, xo, end);
    ll += bbo_llikelihood(P, G[s], xo, skip);
    if (tmp) free(tmp);
    memcpy(start, end, sizeof(double) * d);
  }
  if (xend) memcpy(xend, end, sizeof(double) * d);
  *logU = bbo_accept_logu(seed, iter, chain);
  return ll;
}

/* ======================================================================= CPU baseline driver (bench.py only)
 * Runs `iters` pCN iterations of `P` chains with OpenMP over chains, exactly the reference's
 * loop structure (sample!, combine, solve!, llikelihood, accept with copy-free swap).
 * Returns the number of accepted proposals; *seconds gets the wall time of the loop. */
long long bbo_pcn_bench(const bb_model* Pm, const bbo_guide* const* G, int S, long long P,
                        const double* u, double rho, uint64_t seed, int iters, int skip,
                        int nthreads, double* seconds, double* ll_out) {
  int d = Pm->d, dp = Pm->dprime, N = G[0]->N;
  size_t wsz = (size_t)S * N * dp, xsz = (size_t)S * N * d;
  double* W = (double*)calloc((size_t)P * 2 * wsz, sizeof(double));
  double* X = (double*)calloc((size_t)P * 2 * xsz, sizeof(double));
  double* ll = (double*)calloc((size_t)P, sizeof(double));
  unsigned char* par = (unsigned char*)calloc((size_t)P, 1);
  long long acc = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  /* initial state: W ~ Wiener, X = guided Euler, ll */
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < P; c++) {
    double start[DM], end[DM];
    memcpy(start, u, sizeof(double) * d);
    double l = 0;
    for (int s = 0; s < S; s++) {
      double* w = W + ((size_t)c * 2) * wsz + (size_t)s * N * dp;
      double* x = X + ((size_t)c * 2) * xsz + (size_t)s * N * d;
      bbo_wiener_sample(N, dp, G[s]->tt, seed, 0xFFFFFFFEu, (uint64_t)c * S + s, w);
      bbo_guided_euler(Pm, G[s], start, w, x, end);
      l += bbo_llikelihood(Pm, G[s], x, skip);
      memcpy(start, end, sizeof(double) * d);
    }
    ll[c] = l;
  }
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  for (int it = 0; it < iters; it++) {
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (long long c = 0; c < P; c++) {
      int p = par[c];
      double logU, xe[DM];
      double llo = bbo_pcn_propose(Pm, G, S, u, W + ((size_t)c * 2 + p) * wsz, rho, seed,
                                   (uint32_t)it, (uint64_t)c, skip,
                                   W + ((size_t)c * 2 + (1 - p)) * wsz,
                                   X + ((size_t)c * 2 + (1 - p)) * xsz, xe, &logU);
      if (logU <= llo - ll[c]) {
        par[c] = (unsigned char)(1 - p);
        ll[c] = llo;
        acc += 1;
      }
    }
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (seconds) *seconds = t1 - t0;
  if (ll_out) memcpy(ll_out, ll, sizeof(double) * (size_t)P);
  free(W); free(X); free(ll); free(par);
  return acc;
}

/* ======================================================================= tuned CPU baseline driver (bench.py only)
 * The SAME algorithm and pass structure as bbo_pcn_bench / the reference loop (sample!(W2), Wo .= ρW + √(1-ρ²)W2,
 * solve!(Euler(), Xo, x0, Wo, Po), llikelihood(LeftRule(), Xo, Po), accept with a buffer swap), written the way Julia
 * specialises it for this workload -- FitzHugh-Nagumo hypoelliptic target (SVector{2}, scalar Wiener), PartialBridgeνH
 * with a constant auxiliary drift: no generic-d loops or DM-sized temporaries, model constants hoisted, every Philox
 * call yields its four normals (the generic driver uses one of four), the normals of a segment drawn in a batch
 * (vectorisable).  Operation ORDER is that of the generic code, so in the contraction-free builds the two drivers
 * agree bit for bit (tests/test_oracle_pins.py); the timed `fast` build may contract differently. */
static inline void bm_pair(uint32_t wu, uint32_t wa, float* z0, float* z1) { /* bb_box_muller without the switch */
  float rad = sqrtf(-2.0f * bb_logf(bb_unif(wu)));
  float t = (float)(int32_t)wa * 0x1p-31f;
  float q = rintf(t * 2.0f);
  float r = fmaf(q, -0.5f, t);
  float s2 = r * r;
  float ps = -0x1.2d9b7cp-1f;
  ps = fmaf(ps, s2, 0x1.465ec4p+1f);
  ps = fmaf(ps, s2, -0x1.4abbbap+2f);
  ps = fmaf(ps, s2, 0x1.921fb6p+1f);
  float sr = ps * r;
  float pc = 0x1.d9c326p-3f;
  pc = fmaf(pc, s2, -0x1.55c57ap+0f);
  pc = fmaf(pc, s2, 0x1.03c1dcp+2f);
  pc = fmaf(pc, s2, -0x1.3bd3ccp+2f);
  float cr = fmaf(pc, s2, 1.0f);
  int qi = ((int)q) & 3;
  float a = (qi & 1) ? cr : sr, b = (qi & 1) ? sr : cr;
  *z0 = rad * ((qi & 2) ? -a : a);
  *z1 = rad * ((qi == 1 || qi == 2) ? -b : b);
}
/* normals 0 .. 4*nq-1 of row `row` (quads 0 .. nq-1), as bbo_normal_quad gives them */
static void normals_batch(uint64_t seed, uint32_t stream, uint64_t row, int nq, float* z) {
  const uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32), r0 = (uint32_t)row, r1 = (uint32_t)(row >> 32);
#pragma omp simd
  for (int q = 0; q < nq; q++) {
    uint32_t c0 = (uint32_t)q, c1 = stream, c2 = r0, c3 = r1, k0 = s0, k1 = s1;
    for (int r = 0; r < 10; r++) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    bm_pair(c0, c1, &z[4 * q], &z[4 * q + 1]);
    bm_pair(c2, c3, &z[4 * q + 2], &z[4 * q + 3]);
  }
}
typedef struct { double eps, s, gam, beta, sig, a11; } fhn_const;
static inline void fhn_b(const fhn_const* c, double x1, double x2, double* b0, double* b1) { /* model_b, FHN_HYPO */
  double cc = x1 * x1, u = x1 - x2;
#ifdef ORACLE_GPU_ORDER
  u = fma(-cc, x1, u);
  *b0 = (u + c->s) * (1.0 / c->eps);
#else
  u = u - cc * x1;
  *b0 = (u + c->s) / c->eps;
#endif
  *b1 = MA(c->gam, x1, -x2) + c->beta;
}
/* one pCN proposal of one chain: returns ll°, writes W° and X° */
static double fhn_tuned_propose(const fhn_const* c, const bbo_guide* const* G, int S, const double* u, const double* Wc,
                                double rho, double rho2, uint64_t seed, uint32_t iter, uint64_t chain, double* Wo,
                                double* Xo, float* z) {
  double y0 = u[0], y1 = u[1], ll = 0.0;
  for (int s = 0; s < S; s++) {
    const int N = G[s]->N;
    const double* restrict tt = G[s]->tt;
    const double* restrict H = G[s]->A;
    const double* restrict nu = G[s]->b;
    const double* Bt = G[s]->Bt; const double* be = G[s]->betat;
    const double* restrict wc = Wc + (size_t)s * N;
    double* restrict wo = Wo + (size_t)s * N;
    double* restrict xo = Xo + (size_t)s * N * 2;
    /* sample!(W2, Wiener()); Wo.yy .= ρ*W.yy + sqrt(1-ρ^2)*W2.yy */
    normals_batch(seed, iter, chain * (uint64_t)S + s, (N + 3) >> 2, z);
    double w2 = 0.0;
    wo[0] = MA(rho2, w2, rho * wc[0]);
    for (int j = 1; j < N; j++) {
      w2 = MA(sqrt(tt[j] - tt[j - 1]), (double)z[j], w2);
      wo[j] = MA(rho2, w2, rho * wc[j]);
    }
    /* solve!(Euler(), Xo, x0, Wo, Po)   src/euler.jl:247-268 with _b = b + a*H[i]*(ν[i] - x) */
    for (int i = 0; i < N - 1; i++) {
      xo[2 * i] = y0; xo[2 * i + 1] = y1;
      double b0, b1;
      fhn_b(c, y0, y1, &b0, &b1);
      const double e0 = nu[2 * i] - y0, e1 = nu[2 * i + 1] - y1;
      const double r0 = MA(H[4 * i + 1], e1, H[4 * i] * e0), r1 = MA(H[4 * i + 3], e1, H[4 * i + 2] * e0);
      b0 = MA(0.0, r0, b0);
      b1 = MA(c->a11, r1, b1);
      const double dt = tt[i + 1] - tt[i], dw = wo[i + 1] - wo[i];
      y0 = MA(b0, dt, y0) + 0.0 * dw;
      y1 = MA(c->sig, dw, MA(b1, dt, y1));
    }
    xo[2 * (N - 1)] = y0; xo[2 * (N - 1) + 1] = y1;
    /* llikelihood(LeftRule(), Xo, Po)   src/partialbridgenuH.jl:171-189 */
    double som = 0.0;
    for (int i = 0; i < N - 1; i++) {
      const double x0 = xo[2 * i], x1 = xo[2 * i + 1];
      const double e0 = nu[2 * i] - x0, e1 = nu[2 * i + 1] - x1;
      const double r0 = MA(H[4 * i + 1], e1, H[4 * i] * e0), r1 = MA(H[4 * i + 3], e1, H[4 * i + 2] * e0);
      double b0, b1;
      fhn_b(c, x0, x1, &b0, &b1);
      const double bt0 = MA(Bt[1], x1, Bt[0] * x0), bt1 = MA(Bt[3], x1, Bt[2] * x0);
      const double d0 = b0 - (bt0 + be[0]), d1 = b1 - (bt1 + be[1]);
      som = MA(MA(d1, r1, d0 * r0), tt[i + 1] - tt[i], som);
    }
    ll += som;
  }
  return ll;
}
long long bbo_pcn_bench_fhn_tuned(const bb_model* Pm, const bbo_guide* const* G, int S, long long P, const double* u,
                                  double rho, uint64_t seed, int iters, int nthreads, double* seconds,
                                  double* ll_out) {
  if (Pm->id != BB_MODEL_FHN_HYPO) return -1;
  for (int s = 0; s < S; s++)
    if (G[s]->kind != BB_GUIDE_NUH || !G[s]->aux_const || G[s]->Adiff || G[s]->N != G[0]->N) return -1;
  const int N = G[0]->N;
  const fhn_const c = {Pm->par[0], Pm->par[1], Pm->par[2], Pm->par[3], Pm->par[4], 0.0 + Pm->par[4] * Pm->par[4]};
  const double rho2 = sqrt(1 - rho * rho);
  size_t wsz = (size_t)S * N, xsz = (size_t)S * N * 2;
  double* W = (double*)calloc((size_t)P * 2 * wsz, sizeof(double));
  double* X = (double*)calloc((size_t)P * 2 * xsz, sizeof(double));
  double* ll = (double*)calloc((size_t)P, sizeof(double));
  unsigned char* par = (unsigned char*)calloc((size_t)P, 1);
  long long acc = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (long long ch = 0; ch < P; ch++) { /* initial state exactly as bbo_pcn_bench */
    double start[2] = {u[0], u[1]}, end[2], l = 0;
    for (int s = 0; s < S; s++) {
      double* w = W + ((size_t)ch * 2) * wsz + (size_t)s * N;
      double* x = X + ((size_t)ch * 2) * xsz + (size_t)s * N * 2;
      bbo_wiener_sample(N, 1, G[s]->tt, seed, 0xFFFFFFFEu, (uint64_t)ch * S + s, w);
      bbo_guided_euler(Pm, G[s], start, w, x, end);
      l += bbo_llikelihood(Pm, G[s], x, 0);
      start[0] = end[0]; start[1] = end[1];
    }
    ll[ch] = l;
  }
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
#pragma omp parallel reduction(+ : acc)
  {
    float* z = (float*)malloc(sizeof(float) * (size_t)(N + 8));
    for (int it = 0; it < iters; it++) {
#pragma omp for schedule(static)
      for (long long ch = 0; ch < P; ch++) {
        const int p = par[ch];
        const double llo = fhn_tuned_propose(&c, G, S, u, W + ((size_t)ch * 2 + p) * wsz, rho, rho2, seed, (uint32_t)it,
                                             (uint64_t)ch, W + ((size_t)ch * 2 + (1 - p)) * wsz,
                                             X + ((size_t)ch * 2 + (1 - p)) * xsz, z);
        if (bbo_accept_logu(seed, (uint32_t)it, (uint64_t)ch) <= llo - ll[ch]) {
          par[ch] = (unsigned char)(1 - p);
          ll[ch] = llo;
          acc += 1;
        }
      }
    }
    free(z);
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (seconds) *seconds = t1 - t0;
  if (ll_out) memcpy(ll_out, ll, sizeof(double) * (size_t)P);
  free(W); free(X); free(ll); free(par);
  return acc;
}

/* ======================================================================= CPU baseline of the parameter-update step
 * (timing only: tools/cpu_theta_baseline.py).  The `updateparams` branch of partialbridge_bolus3.jl:268-355 for P
 * independent chains of the hypoelliptic FitzHugh-Nagumo model with the "matching" auxiliary process
 * (partialbridge_fitzhugh.jl:44-46,106-108), flat priors, fixed starting point; OpenMP over chains.  Per chain and
 * iteration: propose θ°, backward chain (gpupdate + Lyapunov step per segment), solve! + llikelihood with W fixed,
 * logpdfnormal / trace terms, accept.  Returns the number of accepted proposals. */
static void theta_fhn_backward(const double* par, int S, int N, const double* grids, const double* x0, const double* L,
                               double Sigma, double eps, const double* obs_v, double* nu_t, double* H_t,
                               double* lpn, double* trsum) {
  double nu[2] = {0, 0}, Hp[4] = {1.0 / eps, 0, 0, 1.0 / eps}, C = 0.0;
  double at[4] = {0, 0, 0, par[4] * par[4]};
  bbo_gpupdate_nuH(2, 1, nu, Hp, L, &Sigma, &obs_v[S - 1]);
  *trsum = 0.0;
  for (int s = S - 1; s >= 0; s--) {
    double ie = 1.0 / par[0], v = obs_v[s];
    double Bt[4] = {ie, -ie, par[2], -1.0}, bt[2] = {par[1] / par[0] - (v * v * v) / par[0], par[3]};
    bb_aux A = {2, 1, Bt, bt, at, NULL};
    double nul[2], Hpl[4];
    bbo_backward_nuH(BB_ODE_LYAP, N, 2, grids + (size_t)s * N, &A, nu, Hp, C, nu_t + (size_t)s * N * 2,
                     H_t + (size_t)s * N * 4, nul, Hpl, &C);
    memcpy(nu, nul, sizeof(nu)); memcpy(Hp, Hpl, sizeof(Hp));
    *trsum += (grids[(size_t)s * N + N - 1] - grids[(size_t)s * N]) * (Bt[0] + Bt[3]);
    if (s > 0) bbo_gpupdate_nuH(2, 1, nu, Hp, L, &Sigma, &obs_v[s - 1]);
  }
  double x[2] = {x0[0] - nu[0], x0[1] - nu[1]};
  *lpn = bbo_logpdfnormal(2, x, Hp);
}
static double theta_fhn_forward(const double* par, int S, int N, const double* grids, const double* x0, const double* obs_v,
                                const double* nu_t, const double* H_t, const double* W, double* X) {
  bb_model P;
  memset(&P, 0, sizeof(P));
  P.id = BB_MODEL_FHN_HYPO; P.d = 2; P.dprime = 1;
  memcpy(P.par, par, sizeof(double) * 5);
  double start[2] = {x0[0], x0[1]}, end[2], ll = 0.0;
  for (int s = 0; s < S; s++) {
    double ie = 1.0 / par[0], v = obs_v[s];
    double Bt[4] = {ie, -ie, par[2], -1.0}, bt[2] = {par[1] / par[0] - (v * v * v) / par[0], par[3]};
    bbo_guide G;
    memset(&G, 0, sizeof(G));
    G.kind = BB_GUIDE_NUH; G.N = N; G.d = 2; G.tt = grids + (size_t)s * N; G.A = H_t + (size_t)s * N * 4;
    G.b = nu_t + (size_t)s * N * 2; G.Bt = Bt; G.betat = bt; G.aux_const = 1; G.adiff_const = 1;
    bbo_guided_euler(&P, &G, start, W + (size_t)s * N, X + (size_t)s * N * 2, end);
    ll += bbo_llikelihood(&P, &G, X + (size_t)s * N * 2, 0);
    start[0] = end[0]; start[1] = end[1];
  }
  return ll;
}
long long bbo_theta_param_bench(const double* par0, int S, int N, const double* grids, const double* x0, const double* L,
                                double Sigma, double eps, const double* obs_v, const double* rw_sd, long long P,
                                uint64_t seed, int iters, int nthreads, double* seconds) {
  size_t wsz = (size_t)S * N;
  double* W = (double*)calloc((size_t)P * wsz, sizeof(double));
  double* th = (double*)calloc((size_t)P * 5, sizeof(double));
  double* ll = (double*)calloc((size_t)P, sizeof(double));
  double* lp = (double*)calloc((size_t)P * 2, sizeof(double));
  long long acc = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* nu_t = (double*)malloc(sizeof(double) * wsz * 2);
    double* H_t = (double*)malloc(sizeof(double) * wsz * 4);
    double* X = (double*)malloc(sizeof(double) * wsz * 2);
#pragma omp for schedule(static)
    for (long long c = 0; c < P; c++) {
      memcpy(th + c * 5, par0, sizeof(double) * 5);
      for (int s = 0; s < S; s++)
        bbo_wiener_sample(N, 1, grids + (size_t)s * N, seed, 0xFFFFFFFEu, (uint64_t)c * S + s, W + c * wsz + (size_t)s * N);
      theta_fhn_backward(th + c * 5, S, N, grids, x0, L, Sigma, eps, obs_v, nu_t, H_t, &lp[2 * c], &lp[2 * c + 1]);
      ll[c] = theta_fhn_forward(th + c * 5, S, N, grids, x0, obs_v, nu_t, H_t, W + c * wsz, X);
    }
    free(nu_t); free(H_t); free(X);
  }
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  for (int it = 0; it < iters; it++) {
#pragma omp parallel reduction(+ : acc)
    {
      double* nu_t = (double*)malloc(sizeof(double) * wsz * 2);
      double* H_t = (double*)malloc(sizeof(double) * wsz * 4);
      double* X = (double*)malloc(sizeof(double) * wsz * 2);
#pragma omp for schedule(static)
      for (long long c = 0; c < P; c++) {
        double tp[5], z[4], lpn, trs;
        bbo_normal_quad(seed, (uint32_t)it, (uint64_t)c, 0xFFFFFFFEu, z);
        int n = 0;
        for (int k = 0; k < 5; k++) {
          tp[k] = th[c * 5 + k];
          if (rw_sd[k] != 0.0) tp[k] = tp[k] + rw_sd[k] * z[n++ & 3];
        }
        theta_fhn_backward(tp, S, N, grids, x0, L, Sigma, eps, obs_v, nu_t, H_t, &lpn, &trs);
        double llo = theta_fhn_forward(tp, S, N, grids, x0, obs_v, nu_t, H_t, W + c * wsz, X);
        double diff = lpn - lp[2 * c];
        diff += llo - ll[c];
        diff += trs - lp[2 * c + 1];
        if (bbo_logu_q(seed, (uint32_t)it, (uint64_t)c, 0xFFFFFFFDu) <= diff) {
          memcpy(th + c * 5, tp, sizeof(tp));
          ll[c] = llo; lp[2 * c] = lpn; lp[2 * c + 1] = trs;
          acc += 1;
        }
      }
      free(nu_t); free(H_t); free(X);
    }
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (seconds) *seconds = t1 - t0;
  free(W); free(th); free(ll); free(lp);
  return acc;
}

int bbo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
