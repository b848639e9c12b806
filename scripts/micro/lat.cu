// Latency microbenchmarks (one warp, dependent chains) for the ops on the Brent round's critical path.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double *out, long long *cyc, double x0, double y0) {
  double x = x0 + threadIdx.x * 1e-9, y = y0;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x * y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x + y;
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = y / x + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // shuffle of a double + multiply (one reduction level)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) x = x * __shfl_xor_sync(0xffffffffu, x, 1) + 0.5;
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // log10 (libm)
  t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; i++) x = log10(x + 2.0) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // barrier (block of blockDim threads)
  t0 = clock64();
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 8 independent DFMA chains (throughput of one warp)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9); a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  x += a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  // rcp approx + 2 Newton steps
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r);
    x = r + 1.5;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main() {
  double *out; long long *cyc, h[9];
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 9 * 8);
  const char *names[9] = {"dfma_dep", "dmul_dep", "dadd_dep", "ddiv_dep(+dadd)", "shfl+dfma_dep", "log10_dep(+2dadd)", "bar", "dfma_x8_indep(per 8)", "rcp_newton2(+dadd)"};
  for (int threads : {32, 64, 128}) {
    k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999);
    k<<<1, threads>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("threads=%d:", threads);
    for (int i = 0; i < 9; i++) printf(" %s=%.1f", names[i], (double)h[i] / N);
    printf("\n");
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
