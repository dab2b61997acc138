// tuning helper: how fast can a kernel write device data straight into page-locked, device-mapped host memory
// (the BB_RUN_SKIP_REJECTED path of bb_pcn_step_host), against cudaMemcpyAsync of the same bytes?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pciebench.cu -o gpurun_out/pciebench
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void wr(const double* __restrict__ src, double* __restrict__ dst, size_t n, size_t run, size_t stride) {
  // rows of `run` doubles are contiguous on both sides; every other row is skipped when stride == 2 (rejected chains)
  size_t nv = n / V;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nv; t += (size_t)gridDim.x * blockDim.x) {
    size_t e = t * V, row = e / run;
    if (stride == 2 && (row & 1)) continue;
    if (V == 1) dst[e] = src[e];
    if (V == 2) *reinterpret_cast<double2*>(dst + e) = *reinterpret_cast<const double2*>(src + e);
    if (V == 4) {
      double4 v = *reinterpret_cast<const double4*>(src + e);
      asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst + e), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    }
  }
}
int main() {
  const size_t bytes = (size_t)2 << 30, n = bytes / 8;
  double *d, *h;
  cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(b); cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemcpyAsync D2H            %.1f GB/s\n", bytes / ms * 1e-6);
  }
  int grids[] = {148, 148 * 2, 148 * 8, 148 * 32};
  for (int g : grids) {
    for (int v = 1; v <= 4; v *= 2) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        if (v == 1) wr<1><<<g, 256>>>(d, h, n, 4096, 1);
        if (v == 2) wr<2><<<g, 256>>>(d, h, n, 4096, 1);
        if (v == 4) wr<4><<<g, 256>>>(d, h, n, 4096, 1);
        cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
      }
      printf("kernel store %2d B/thread grid %5d  %.1f GB/s   (%s)\n", 8 * v, g, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  for (int v = 1; v <= 4; v *= 4) {
    cudaEventRecord(a);
    if (v == 1) wr<1><<<148 * 8, 256>>>(d, h, n, 4096, 2); else wr<4><<<148 * 8, 256>>>(d, h, n, 4096, 2);
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    printf("every other 32 KB row, %2d B/thread     %.1f GB/s of the bytes written\n", 8 * v, bytes / 2 / ms * 1e-6);
  }
  // both directions at once: H2D by the copy engine on a second stream, D2H by (a) the copy engine, (b) the kernel
  double *d2, *h2;
  cudaMalloc(&d2, bytes); cudaHostAlloc(&h2, bytes, cudaHostAllocDefault);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t c0, c1; cudaEventCreate(&c0); cudaEventCreate(&c1);
  for (int mode = 0; mode < 3; mode++) {
    cudaDeviceSynchronize();
    cudaEventRecord(a, s1); cudaEventRecord(c0, s2);
    for (int r = 0; r < 2; r++) cudaMemcpyAsync(d2, h2, bytes, cudaMemcpyHostToDevice, s2);
    for (int r = 0; r < 2; r++) {
      if (mode == 0) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s1);
      if (mode == 1) wr<1><<<148 * 4, 256, 0, s1>>>(d, h, n, 4096, 1);
      if (mode == 2) wr<4><<<148 * 4, 256, 0, s1>>>(d, h, n, 4096, 1);
    }
    cudaEventRecord(b, s1); cudaEventRecord(c1, s2);
    cudaDeviceSynchronize();
    float m1, m2; cudaEventElapsedTime(&m1, a, b); cudaEventElapsedTime(&m2, c0, c1);
    printf("duplex, D2H by %s: D2H %.1f GB/s, concurrent H2D (copy engine) %.1f GB/s\n",
           mode == 0 ? "copy engine" : (mode == 1 ? "kernel, 8 B stores" : "kernel, 32 B stores"), 2 * bytes / m1 * 1e-6,
           2 * bytes / m2 * 1e-6);
  }
  return 0;
}
