// fp64_lat.cu — dependent-issue latencies of the FP64 scalar chain of the panel kernels on B200 (one warp, one SM):
// DFMA, DADD, double division, double sqrt, rsqrt-based replacement, warp shuffle sum, __syncthreads with 512 threads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_lat tools/fp64_lat.cu
#include <cuda_runtime.h>
#include <cstdio>
#define N 512
__global__ void k(double* out, long long* cyc, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fma(a, b, 1e-3);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = a + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = 1.0 + b / (a + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // sqrt chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) a = sqrt(a + 3.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // the reflector scalars as written in k_panel_*: sqrt, then two divisions (independent of each other)
  double tau = 0, scale = 0;
  t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    const double alpha = a, xn2 = b + tau * 1e-30 + scale * 1e-30;
    const double h = sqrt(fma(alpha, alpha, xn2));
    const double beta = (alpha >= 0.0) ? -h : h;
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
    a = a + tau * 1e-30;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // rsqrt-based: one rsqrt + one reciprocal
  t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    const double alpha = a, xn2 = b + tau * 1e-30 + scale * 1e-30;
    const double s2 = fma(alpha, alpha, xn2);
    const double rh = rsqrt(s2);
    const double h = s2 * rh;
    const double beta = (alpha >= 0.0) ? -h : h;
    const double rbeta = (alpha >= 0.0) ? -rh : rh;
    tau = (beta - alpha) * rbeta;
    scale = 1.0 / (alpha - beta);
    a = a + tau * 1e-30;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // warp sum (5 shuffles + adds)
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // __syncthreads
  t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = a + tau + scale;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 512); cudaMalloc(&cyc, 8 * 8);
  for (int threads : {32, 512}) {
    k<<<1, threads>>>(out, cyc, 1.0, 0.999);
    k<<<1, threads>>>(out, cyc, 1.0, 0.999);
    cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* names[8] = {"DFMA", "DADD", "div(+2 adds)", "sqrt(+add)", "scalars sqrt+2div", "scalars rsqrt+1rcp", "warp_sum", "__syncthreads"};
    for (int i = 0; i < 8; ++i) printf("%d threads: %-20s %.1f cycles per iteration\n", threads, names[i], (double)h[i] / N);
  }
  return 0;
}
