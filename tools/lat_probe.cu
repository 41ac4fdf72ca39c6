// Latency probes on sm_100a: dependent DFMA / DADD / SHFL+DADD / sqrt / div chains, single warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fma(x, b, a);
  }
  long long t1 = clock64();
  double y = x;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) y = y + a;
  }
  long long t2 = clock64();
  double z = y;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) z += __shfl_xor_sync(0xffffffffu, z, 1 << u);
  }
  long long t3 = clock64();
  double s = fabs(z) + 2.0;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) s = sqrt(s) + a;
  long long t4 = clock64();
  double d = s;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) d = a / d + b;
  long long t5 = clock64();
  double h = d;
#pragma unroll 1
  for (int i = 0; i < 256; ++i) h = hypot(h, a);
  long long t6 = clock64();
  float f = (float)h;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) f = fmaf(f, 1.0001f, 0.5f);
  }
  long long t7 = clock64();
  out[threadIdx.x] = x + y + z + s + d + h + f;
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = t6 - t5; cyc[6] = t7 - t6;
  }
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 64);
  for (int threads = 32; threads <= 512; threads *= 4) {
    probe<<<1, threads>>>(out, cyc, 1.000001, 0.999999);
    cudaDeviceSynchronize();
    printf("threads %3d: DFMA %.1f cyc/op  DADD %.1f  SHFL+DADD %.1f  sqrt+add %.1f  div+add %.1f  hypot %.1f  FFMA %.1f\n", threads,
           cyc[0] / 1024.0, cyc[1] / 1024.0, cyc[2] / 256.0, cyc[3] / 256.0, cyc[4] / 256.0, cyc[5] / 256.0, cyc[6] / 1024.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
