// Microbenchmark: FP64 peak of the DFMA pipe vs the DMMA (mma.sync f64) tensor path on sm_100a.
// Decides which instruction the trailing-update kernel (K6) is built on (BASELINE.json north_star:
// "uses FP64 DMMA tensor-core MMA where ncu shows it beats the FP64 FMA pipe").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void __launch_bounds__(1024) dmma884_kernel(double* out, int iters, double a, double b) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4], af[4] = {a, a, b, b}, bf[2] = {b, a};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(c[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a, double b) {
  double c[ILP][4], af[8] = {a, a, b, b, a, b, a, b}, bf[4] = {b, a, a, b};
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma16816(c[i], af, bf);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d MHz\n", p.name, sms, p.clockRate / 1000);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = sms * bps;
    { constexpr int ILP = 8;
      float ms = time_it([&] { dfma_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      printf("DFMA      ILP%-2d blocks/SM %d: %8.2f TFLOP/s\n", ILP, bps, 2.0 * ILP * iters * 256.0 * grid / ms / 1e9); }
    { constexpr int ILP = 8;
      float ms = time_it([&] { dmma884_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      printf("DMMA 884  ILP%-2d blocks/SM %d: %8.2f TFLOP/s\n", ILP, bps, 2.0 * 256 * ILP * iters * 8.0 * grid / ms / 1e9); }
    { constexpr int ILP = 8;
      float ms = time_it([&] { dmma1688_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      printf("DMMA 1688 ILP%-2d blocks/SM %d: %8.2f TFLOP/s\n", ILP, bps, 2.0 * 1024 * ILP * iters * 8.0 * grid / ms / 1e9); }
    { constexpr int ILP = 8;
      float ms = time_it([&] { dmma16816_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
      printf("DMMA 16816 ILP%-2d blocks/SM %d: %8.2f TFLOP/s\n", ILP, bps, 2.0 * 2048 * ILP * iters * 8.0 * grid / ms / 1e9); }
  }
  // how many warps per SM sub-partition does DMMA need?  (threads per block, 1 block per SM)
  for (int threads = 128; threads <= 1024; threads *= 2) {
    constexpr int ILP = 8;
    float ms = time_it([&] { dmma884_kernel<ILP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA 884  ILP8 %4d threads/SM (%d warps/SMSP): %8.2f TFLOP/s\n", threads, threads / 128,
           2.0 * 256 * ILP * iters * (threads / 32.0) * sms / ms / 1e9);
  }
  for (int threads = 128; threads <= 512; threads *= 2) {
    constexpr int ILP = 16;
    float ms = time_it([&] { dmma884_kernel<ILP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("DMMA 884 ILP16 %4d threads/SM (%d warps/SMSP): %8.2f TFLOP/s\n", threads, threads / 128,
           2.0 * 256 * ILP * iters * (threads / 32.0) * sms / ms / 1e9);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return 0;
}
