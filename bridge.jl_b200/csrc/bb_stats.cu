/*
 * bb_stats.cu -- online statistics of the chains' current paths, pooled over chains and over calls: the ensemble
 * analogue of mcstart / mcnext! / mcstats (src/mclog.jl:22-56, 88-93).  The reference keeps, for ONE chain,
 * m = running mean path and m2 = sum of outer products of deviations over the iterations; a sampler that runs P
 * chains at once wants the same two moments over all chains and all recorded iterations without downloading
 * P paths per iteration (project_partialbridge/partialbridge_fitzhugh.jl:169-189 stores every 1000th path).
 *
 * Accumulators: sum[S][N][d], sq[S][N][d][d] (double), n -- of the DEVIATIONS from a pivot path c[S][N][d] (the path of
 * chain 0 at the first update after a reset), as Welford's update in the reference works with deviations from the running
 * mean: near a conditioned end point the variance is tiny against the mean (Σ = 1e-10 in the FitzHugh-Nagumo configs) and
 * raw moments sum(x^2) - n mean^2 would cancel.  mean = c + sum/n, cov = (sq - n (sum/n)(sum/n)') / (n-1).
 * One CTA reduces the 16 grid points of one chunk over 256 chains in shared memory and issues d + d*d atomic adds per
 * grid point; X is read once (8 d bytes per path-step).  The atomic adds make the last bits depend on the order in which
 * CTAs arrive (the only non-deterministic results of the library); the mcband semantics differ from the reference's in
 * that moments pool P chains per call (src/mclog.jl:22-56 tracks one chain over iterations).
 */
#include <string.h>

#include <vector>

#include "bb_host.h"

/* pivot[s][j][a] = X of chain 0 */
__global__ void bb_mc_pivot_kernel(const double* __restrict__ X, double* __restrict__ pivot, long long P, int S, int N, int NC,
                                   int D) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S * N * D) return;
  const int a = t % D, j = (t / D) % N, s = t / (D * N);
  const int c = j / BB_TC, slot = j % BB_TC;
  pivot[t] = X[(((long long)s * NC + c) * P) * (BB_TC * D) + slot * D + a];
}

template <int D>
__global__ void __launch_bounds__(256) bb_mc_update_kernel(const double* __restrict__ X, const double* __restrict__ pivot,
                                                           double* __restrict__ sum, double* __restrict__ sq, long long P,
                                                           int N, int NC) {
  constexpr int ROW = BB_TC * D, SUB = 64; /* chains per shared-memory tile; a CTA covers 256 chains in 4 tiles */
  constexpr int NM = BB_TC * (D + D * D);  /* moments per chunk: thread t < NM owns one of them */
  __shared__ double tile[SUB][ROW + 1];
  const int chunk = blockIdx.y; /* s * NC + c */
  const int s = chunk / NC, c = chunk - s * NC;
  int slot = 0, a = 0, b = -1;
  if ((int)threadIdx.x < BB_TC * D) { slot = threadIdx.x / D; a = threadIdx.x % D; }
  else if ((int)threadIdx.x < NM) {
    const int u = threadIdx.x - BB_TC * D;
    slot = u / (D * D); a = (u % (D * D)) / D; b = u % D;
  }
  double acc = 0.0;
  for (int sub = 0; sub < 256 / SUB; sub++) {
    /* 4 threads per chain load a quarter row each */
    const int lc = threadIdx.x >> 2, part = threadIdx.x & 3;
    const long long p = (long long)blockIdx.x * 256 + sub * SUB + lc;
    constexpr int Q4 = ROW / 4; /* 256-bit pieces per row */
    for (int q = part; q < Q4; q += 4) {
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      if (p < P) bb_ld4(X + ((long long)chunk * P + p) * ROW + 4 * q, v);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int e = 4 * q + i, jj = c * BB_TC + e / D;
        /* deviation from the pivot path (0 for padding and missing chains: they add nothing) */
        tile[lc][e] = (p < P && jj < N) ? v[i] - pivot[((long long)s * N + jj) * D + e % D] : 0.0;
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < NM) {
      if (b < 0) {
        for (int q = 0; q < SUB; q++) acc += tile[q][slot * D + a];
      } else {
        for (int q = 0; q < SUB; q++) acc = fma(tile[q][slot * D + a], tile[q][slot * D + b], acc);
      }
    }
    __syncthreads();
  }
  const int j = c * BB_TC + slot;
  if ((int)threadIdx.x < NM && j < N) {
    if (b < 0) atomicAdd(&sum[((long long)s * N + j) * D + a], acc);
    else atomicAdd(&sq[(((long long)s * N + j) * D + a) * D + b], acc);
  }
}

extern "C" int bb_ens_mc_reset(bb_ens* e) {
  if (!e) return BB_ERR_ARG;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const size_t n1 = (size_t)e->S * e->N * e->d, n2 = n1 * e->d;
  if (!e->mc_sum) {
    BB_CUDA(cudaMalloc(&e->mc_sum, (n1 + n2 + n1) * sizeof(double)));
    e->mc_sq = e->mc_sum + n1;
    e->mc_pivot = e->mc_sq + n2;
    e->bytes += (int64_t)((n1 + n2 + n1) * sizeof(double));
  }
  BB_CUDA(cudaMemsetAsync(e->mc_sum, 0, (n1 + n2 + n1) * sizeof(double), e->ctx->stream));
  e->mc_n = 0;
  return BB_OK;
}

extern "C" int bb_ens_mc_update(bb_ens* e) {
  if (!e || !e->X) return BB_ERR_ARG;
  if (e->x_maybe_stale) return BB_ERR_STALE; /* statistics are over CURRENT paths: bb_ens_refresh_x first */
  if (!e->mc_sum) {
    int rc = bb_ens_mc_reset(e);
    if (rc != BB_OK) return rc;
  }
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  if (e->d < 1 || e->d > 3) return BB_ERR_UNSUPPORTED;
  const dim3 grid((unsigned)((e->P + 255) / 256), (unsigned)(e->S * e->NC));
  bb_time_begin(c);
  if (e->mc_n == 0) { /* first update after a reset: the pivot is the path of chain 0 */
    const int tot = e->S * e->N * e->d;
    bb_mc_pivot_kernel<<<(tot + 255) / 256, 256, 0, c->stream>>>(e->X, e->mc_pivot, e->P, e->S, e->N, e->NC, e->d);
    c->launches++;
  }
  switch (e->d) {
    case 1: bb_mc_update_kernel<1><<<grid, 256, 0, c->stream>>>(e->X, e->mc_pivot, e->mc_sum, e->mc_sq, e->P, e->N, e->NC); break;
    case 2: bb_mc_update_kernel<2><<<grid, 256, 0, c->stream>>>(e->X, e->mc_pivot, e->mc_sum, e->mc_sq, e->P, e->N, e->NC); break;
    default: bb_mc_update_kernel<3><<<grid, 256, 0, c->stream>>>(e->X, e->mc_pivot, e->mc_sum, e->mc_sq, e->P, e->N, e->NC); break;
  }
  bb_time_end(c);
  BB_CUDA(cudaGetLastError());
  c->launches++;
  e->mc_n += e->P;
  return BB_OK;
}

/* mcstats(mc) = (m, m2/(k - 1))  src/mclog.jl:88-93;  mean [S][N][d], cov [S][N][d][d] (either may be NULL) */
extern "C" int bb_ens_mc_stats(bb_ens* e, double* mean, double* cov, int64_t* n) {
  if (!e) return BB_ERR_ARG;
  if (n) *n = e->mc_n;
  if (!e->mc_sum || e->mc_n == 0) return (mean || cov) ? BB_ERR_ARG : BB_OK;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const int d = e->d;
  const size_t n1 = (size_t)e->S * e->N * d, n2 = n1 * d;
  std::vector<double> h(n1 + n2 + n1);
  BB_CUDA(cudaMemcpyAsync(h.data(), e->mc_sum, (n1 + n2 + n1) * sizeof(double), cudaMemcpyDeviceToHost, e->ctx->stream));
  BB_CUDA(cudaStreamSynchronize(e->ctx->stream));
  const double k = (double)e->mc_n;
  for (size_t g = 0; g < (size_t)e->S * e->N; g++) {
    double m[BB_MAXD];
    for (int a = 0; a < d; a++) {
      m[a] = h[g * d + a] / k; /* mean deviation from the pivot */
      if (mean) mean[g * d + a] = h[n1 + n2 + g * d + a] + m[a];
    }
    if (cov)
      for (int a = 0; a < d; a++)
        for (int b = 0; b < d; b++) {
          double v = (h[n1 + (g * d + a) * d + b] - k * m[a] * m[b]) / (k - 1.0);
          if (a == b && v < 0.0) v = 0.0; /* rounding at (numerically) constant grid points */
          cov[(g * d + a) * d + b] = v;
        }
  }
  return BB_OK;
}


/* ------------------------------------------------------------------------------------------ per chain, over iterations
 * mcstart / mcnext! / mcstats / mcband with the reference's own semantics (src/mclog.jl:22-24, 47-56, 75-93): every chain p
 * keeps ITS OWN state (m, m2, k) -- running mean path and sum of outer products of deviations over the recorded
 * iterations -- as the scripts do with `mcstate = [mcnext!(mcstate[i], XX[i].yy) for i in ...]`
 * (project_partialbridge/partialbridge_fitzhugh.jl:169-189).  The update is Welford's, element by element in the
 * reference's order (no fused operations: the library is compiled with -fmad=false):
 *     delta = x - m;   m += delta / (k + 1);   m2 += delta (x - m)'
 * m has the layout of X ([S][NC][P][16 d]), m2 that of X with d*d entries per grid point: the kernel is element-wise
 * with contiguous accesses, 8 (3 d + 2 d^2) bytes per path-step (112 B for d = 2), HBM-bound and deterministic.
 * The state costs (1 + d) times the memory of X and is allocated on first use only. */
template <int D>
__global__ void __launch_bounds__(256) bb_chain_mc_update_kernel(const double* __restrict__ X, double* __restrict__ m,
                                                                 double* __restrict__ m2, long long P, int N, int NC,
                                                                 int chunks, double kp1) {
  const long long per_chunk = P * BB_TC; /* grid points of all chains in one chunk (s, c): contiguous in X */
  for (int chunk = blockIdx.y; chunk < chunks; chunk += gridDim.y) {
    const int c = chunk % NC;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < per_chunk; u += (long long)gridDim.x * blockDim.x) {
      const int slot = (int)(u & (BB_TC - 1));
      if (c * BB_TC + slot >= N) continue; /* padding of the last chunk */
      const long long t = (long long)chunk * per_chunk + u;
      double x[D], mm[D], delta[D];
#pragma unroll
      for (int a = 0; a < D; a++) {
        x[a] = X[t * D + a];
        mm[a] = m[t * D + a];
        delta[a] = x[a] - mm[a];
        mm[a] = mm[a] + delta[a] / kp1; /* m[i] += delta/(n+1)  src/mclog.jl:52 */
        m[t * D + a] = mm[a];
      }
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int b = 0; b < D; b++) {
          const long long q = (t * D + a) * D + b;
          m2[q] = m2[q] + delta[a] * (x[b] - mm[b]); /* m2[i] += outer(delta, x[i] - m[i])  src/mclog.jl:53 */
        }
    }
  }
}

/* stage[pl][s][j][k] <- src[s][c][p0 + pl][slot][k]  (K = d for m, d*d for m2) */
__global__ void __launch_bounds__(256) bb_chain_mc_gather_kernel(const double* __restrict__ src, double* __restrict__ stage,
                                                                 long long total, long long P, long long p0, int S, int N,
                                                                 int NC, int K) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % K);
    long long q = t / K;
    const int j = (int)(q % N);
    q /= N;
    const int s = (int)(q % S);
    const long long p = p0 + q / S;
    const int c = j / BB_TC, slot = j - c * BB_TC;
    stage[t] = src[((((long long)s * NC + c) * P + p) * BB_TC + slot) * K + k];
  }
}

extern "C" int bb_ens_chain_mc_reset(bb_ens* e) {
  if (!e || !e->X) return BB_ERR_ARG;
  if (e->d < 1 || e->d > 3) return BB_ERR_UNSUPPORTED;
  BB_CUDA(cudaSetDevice(e->ctx->device));
  const size_t n1 = (size_t)e->S * e->NC * e->P * BB_TC * e->d, n2 = n1 * e->d;
  if (!e->cmc_m) {
    BB_CUDA(cudaMalloc(&e->cmc_m, (n1 + n2) * sizeof(double)));
    e->cmc_m2 = e->cmc_m + n1;
    e->bytes += (int64_t)((n1 + n2) * sizeof(double));
  }
  BB_CUDA(cudaMemsetAsync(e->cmc_m, 0, (n1 + n2) * sizeof(double), e->ctx->stream)); /* mcstart: zeros, n = 0 */
  e->cmc_k = 0;
  return BB_OK;
}

extern "C" int bb_ens_chain_mc_update(bb_ens* e) {
  if (!e || !e->X) return BB_ERR_ARG;
  if (e->d < 1 || e->d > 3) return BB_ERR_UNSUPPORTED;
  if (e->x_maybe_stale) return BB_ERR_STALE; /* statistics are over CURRENT paths: bb_ens_refresh_x first */
  if (!e->cmc_m) {
    int rc = bb_ens_chain_mc_reset(e);
    if (rc != BB_OK) return rc;
  }
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  static_assert((BB_TC & (BB_TC - 1)) == 0, "BB_TC is a power of two");
  const int chunks = e->S * e->NC;
  const long long want = (e->P * BB_TC + 255) / 256;
  const dim3 grid((unsigned)(want < 4096 ? want : 4096), (unsigned)(chunks < 65535 ? chunks : 65535));
  const double kp1 = (double)(e->cmc_k + 1);
  bb_time_begin(c);
  switch (e->d) {
    case 1: bb_chain_mc_update_kernel<1><<<grid, 256, 0, c->stream>>>(e->X, e->cmc_m, e->cmc_m2, e->P, e->N, e->NC, chunks, kp1); break;
    case 2: bb_chain_mc_update_kernel<2><<<grid, 256, 0, c->stream>>>(e->X, e->cmc_m, e->cmc_m2, e->P, e->N, e->NC, chunks, kp1); break;
    default: bb_chain_mc_update_kernel<3><<<grid, 256, 0, c->stream>>>(e->X, e->cmc_m, e->cmc_m2, e->P, e->N, e->NC, chunks, kp1); break;
  }
  bb_time_end(c);
  BB_CUDA(cudaGetLastError());
  c->launches++;
  e->cmc_k++;
  return BB_OK;
}

/* raw state of chains p0 .. p0+np-1 in the caller's layout: m -> mean [np][S][N][d], m2 -> m2out [np][S][N][d][d] */
static int chain_mc_fetch(bb_ens* e, int64_t p0, int64_t np, double* mean, double* m2out) {
  bb_ctx* c = e->ctx;
  BB_CUDA(cudaSetDevice(c->device));
  for (int pass = 0; pass < 2; pass++) {
    double* host = pass == 0 ? mean : m2out;
    if (!host) continue;
    const double* src = pass == 0 ? e->cmc_m : e->cmc_m2;
    const int K = pass == 0 ? e->d : e->d * e->d;
    const size_t per_chain = (size_t)e->S * e->N * K;
    int64_t slab = (int64_t)((size_t)(64u << 20) / (per_chain * sizeof(double)));
    if (slab < 1) slab = 1;
    if (slab > np) slab = np;
    double* stage = nullptr;
    BB_CUDA(cudaMalloc(&stage, (size_t)slab * per_chain * sizeof(double)));
    for (int64_t q0 = 0; q0 < np; q0 += slab) {
      const int64_t n = np - q0 < slab ? np - q0 : slab;
      const long long total = (long long)n * (long long)per_chain;
      const long long want = (total + 255) / 256;
      const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
      bb_chain_mc_gather_kernel<<<grid, 256, 0, c->stream>>>(src, stage, total, e->P, p0 + q0, e->S, e->N, e->NC, K);
      cudaError_t err = cudaGetLastError();
      if (err == cudaSuccess)
        err = cudaMemcpyAsync(host + (size_t)q0 * per_chain, stage, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost,
                              c->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(c->stream); /* the staging buffer is reused by the next slab */
      c->launches++;
      if (err != cudaSuccess) {
        cudaFree(stage);
        BB_CUDA(err);
      }
    }
    BB_CUDA(cudaFree(stage));
  }
  return BB_OK;
}

/* mcstats(mc_p) = (m, m2/(k - 1))  src/mclog.jl:88-93 for chains p0 .. p0+np-1 */
extern "C" int bb_ens_chain_mc_stats(bb_ens* e, int64_t p0, int64_t np, double* mean, double* cov, int64_t* k) {
  if (!e) return BB_ERR_ARG;
  if (k) *k = e->cmc_k;
  if (!mean && !cov) return BB_OK;
  if (np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (!e->cmc_m || e->cmc_k == 0) return BB_ERR_ARG;
  if (np == 0) return BB_OK;
  const int rc = chain_mc_fetch(e, p0, np, mean, cov);
  if (rc != BB_OK) return rc;
  if (cov) {
    const double km1 = (double)(e->cmc_k - 1); /* k = 1: 0/0 = NaN, as in the reference */
    const size_t n2 = (size_t)np * e->S * e->N * e->d * e->d;
    for (size_t i = 0; i < n2; i++) cov[i] = cov[i] / km1;
  }
  return BB_OK;
}

/* mcband(mc_p) = (m - Q std, m + Q std), std = sqrt(diag(m2) * (1/(k - 1))), Q = sqrt(2.) * erfinv(0.95)
 * src/mclog.jl:75-85; lower, upper [np][S][N][d] */
extern "C" int bb_ens_chain_mc_band(bb_ens* e, int64_t p0, int64_t np, double* lower, double* upper) {
  if (!e || !lower || !upper || np < 0 || p0 < 0 || p0 + np > e->P) return BB_ERR_ARG;
  if (!e->cmc_m || e->cmc_k == 0) return BB_ERR_ARG;
  if (np == 0) return BB_OK;
  const int d = e->d;
  const size_t n1 = (size_t)np * e->S * e->N * d;
  std::vector<double> m2;
  try {
    m2.resize(n1 * d);
  } catch (...) {
    return BB_ERR_NOMEM;
  }
  const int rc = chain_mc_fetch(e, p0, np, lower, m2.data());
  if (rc != BB_OK) return rc;
  const double Q = BB_MCBAND_Q;
  const double s = 1.0 / (double)(e->cmc_k - 1);
  for (size_t g = 0; g < n1 / d; g++)
    for (int a = 0; a < d; a++) {
      const double mean = lower[g * d + a];
      const double sd = sqrt(m2[(g * d + a) * d + a] * s);
      lower[g * d + a] = mean - Q * sd;
      upper[g * d + a] = mean + Q * sd;
    }
  return BB_OK;
}
