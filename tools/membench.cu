// membench.cu -- what the B200 memory system delivers for the path kernel's ACCESS PATTERN with no arithmetic at all
// (tuning aid; run under gpurun:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu && /tmp/membench)
//   copy      : plain streaming copy (the MEASURED_PEAKS.json denominator, 50 % reads)
//   pcn_burst : per chain and 16-step chunk: read its 128-B W row, write the adjacent 128-B W° row and a 256-B X° row,
//               every row moved by back-to-back 256-bit accesses of the owning thread (25 % reads)
//   pcn_drip  : same bytes, but the 32-byte stores of a row are issued one at a time with a delay in between
//               (as the path kernel does while it computes 2-4 Euler steps between stores)
//   guided    : read W row (other half untouched), write X° row (33 % reads), burst
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld4(const double* p, double* v) {
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, const double* v) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
// W: [chunk][P][2][16] doubles, X: [chunk][P][32] doubles
template <int MODE>  // 0 burst pcn, 1 drip pcn, 2 guided burst
__global__ void __launch_bounds__(256, 2) k_pcn(double* W, double* X, long long P, int nchunk, int delay) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int par = (p * 2654435761u >> 13) & 1;  // pseudo-random buffer parity per chain
  double acc = 0;
  for (int c = 0; c < nchunk; c++) {
    double* wrow = W + ((long long)c * P + p) * 32;
    double* xrow = X + ((long long)c * P + p) * 32;
    double v[16];
#pragma unroll
    for (int q = 0; q < 4; q++) ld4(wrow + par * 16 + 4 * q, v + 4 * q);
#pragma unroll
    for (int q = 0; q < 16; q++) acc += v[q];
    if (MODE != 2) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        st4(wrow + (1 - par) * 16 + 4 * q, v + 4 * q);
        if (MODE == 1) { long long t0 = clock64(); while (clock64() - t0 < delay * 4) {} }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
      st4(xrow + 4 * q, v + 4 * (q & 3));
      if (MODE == 1) { long long t0 = clock64(); while (clock64() - t0 < delay * 2) {} }
    }
  }
  if (acc == 1.2345e300) W[0] = acc;
}
template <class F>
float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main(int argc, char** argv) {
  const long long P = argc > 1 ? atoll(argv[1]) : 250000;
  const int nchunk = 252;  // 4 segments x 63 chunks
  size_t wbytes = (size_t)nchunk * P * 32 * 8, xbytes = wbytes;
  double *W, *X; cudaMalloc(&W, wbytes); cudaMalloc(&X, xbytes); cudaMemset(W, 0, wbytes); cudaMemset(X, 0, xbytes);
  {
    size_t n = wbytes / 32;
    float ms = timeit([&] { k_copy<<<148 * 16, 512>>>((const double4*)W, (double4*)X, n); });
    printf("copy       %8.3f ms  %7.1f GB/s (read+write)\n", ms, 2.0 * wbytes / ms * 1e-6);
  }
  const unsigned grid = (unsigned)((P + 255) / 256);
  const double pcn_bytes = (double)nchunk * P * (128 + 128 + 256), g_bytes = (double)nchunk * P * (128 + 256);
  float ms = timeit([&] { k_pcn<0><<<grid, 256>>>(W, X, P, nchunk, 0); });
  printf("pcn_burst  %8.3f ms  %7.1f GB/s\n", ms, pcn_bytes / ms * 1e-6);
  for (int d : {100, 300, 600}) {
    ms = timeit([&] { k_pcn<1><<<grid, 256>>>(W, X, P, nchunk, d); }, 3);
    printf("pcn_drip %4d cyc/step %8.3f ms  %7.1f GB/s\n", d, ms, pcn_bytes / ms * 1e-6);
  }
  ms = timeit([&] { k_pcn<2><<<grid, 256>>>(W, X, P, nchunk, 0); });
  printf("guided     %8.3f ms  %7.1f GB/s\n", ms, g_bytes / ms * 1e-6);
  return 0;
}
