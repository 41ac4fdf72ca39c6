// prio_probe.cu — how does the B200 block scheduler treat a HIGH-priority kernel whose CTAs need a whole SM while a
// LOW-priority grid of small, short CTAs (two per SM) keeps every SM half busy?  (Look-ahead design question: can the
// panel / Gram kernels of the main stream get their SMs while the side stream's rank-k CTAs stream through?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prio_probe tools/prio_probe.cu ; run: ./prio_probe
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned s;
  asm volatile("mov.u32 %0, %smid;" : "=r"(s));
  return s;
}

struct Rec { unsigned long long t0, t1; unsigned sm; unsigned pad; };

// low priority: thin CTA (128 threads, ~104 KB smem => two per SM), spins for `ns`
__global__ void __launch_bounds__(128) k_low(Rec* rec, long long ns) {
  extern __shared__ double sm[];
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) sm[0] = 1.0;
  while ((long long)(gtime() - t0) < ns) { }
  if (threadIdx.x == 0) { rec[blockIdx.x].t0 = t0; rec[blockIdx.x].t1 = gtime(); rec[blockIdx.x].sm = smid(); }
}
// high priority: fat CTA (512 threads, 200 KB smem => a whole SM), spins for `ns`; `coop`: grid barrier first
__global__ void __launch_bounds__(512, 1) k_high(Rec* rec, long long ns, int coop) {
  extern __shared__ double sm[];
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) sm[0] = 1.0;
  if (coop) cooperative_groups::this_grid().sync();
  const unsigned long long tb = gtime();
  while ((long long)(gtime() - tb) < ns) { }
  if (threadIdx.x == 0) { rec[blockIdx.x].t0 = t0; rec[blockIdx.x].t1 = gtime(); rec[blockIdx.x].sm = smid(); }
}
__global__ void k_mark(unsigned long long* t) { *t = gtime(); }

int main() {
  int lo_p, hi_p;
  cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
  printf("priority range: least %d greatest %d\n", lo_p, hi_p);
  cudaStream_t s_low, s_high;
  const int LOW_SMEM = 104 * 1024, HIGH_SMEM = 200 * 1024;
  cudaFuncSetAttribute(k_low, cudaFuncAttributeMaxDynamicSharedMemorySize, LOW_SMEM);
  cudaFuncSetAttribute(k_high, cudaFuncAttributeMaxDynamicSharedMemorySize, HIGH_SMEM);
  const int NLOW = 296 * 40, NHIGH = 120;
  Rec *d_low, *d_high;
  unsigned long long* d_mark;
  cudaMalloc(&d_low, sizeof(Rec) * NLOW);
  cudaMalloc(&d_high, sizeof(Rec) * NHIGH);
  cudaMalloc(&d_mark, 8);
  std::vector<Rec> low(NLOW), high(NHIGH);
  for (int variant = 0; variant < 4; ++variant) {
    const bool prio = variant & 1, coop = variant & 2;
    cudaStreamCreateWithPriority(&s_low, cudaStreamNonBlocking, lo_p);
    cudaStreamCreateWithPriority(&s_high, cudaStreamNonBlocking, prio ? hi_p : lo_p);
    cudaMemset(d_low, 0, sizeof(Rec) * NLOW);
    cudaMemset(d_high, 0, sizeof(Rec) * NHIGH);
    cudaDeviceSynchronize();
    // low grid: 40 waves of 296 CTAs x 5 us = ~200 us of work; the high kernel is launched ~while it runs
    k_low<<<NLOW, 128, LOW_SMEM, s_low>>>(d_low, 5000);
    k_mark<<<1, 1, 0, s_high>>>(d_mark);
    long long ns = 50000;
    int cflag = coop ? 1 : 0;
    Rec* dh = d_high;
    void* args[] = {&dh, &ns, &cflag};
    cudaError_t e;
    if (coop) e = cudaLaunchCooperativeKernel((void*)k_high, dim3(NHIGH), dim3(512), args, HIGH_SMEM, s_high);
    else { k_high<<<NHIGH, 512, HIGH_SMEM, s_high>>>(d_high, ns, 0); e = cudaGetLastError(); }
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("variant %d: launch %s sync %s\n", variant, cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
    unsigned long long mark;
    cudaMemcpy(&mark, d_mark, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(low.data(), d_low, sizeof(Rec) * NLOW, cudaMemcpyDeviceToHost);
    cudaMemcpy(high.data(), d_high, sizeof(Rec) * NHIGH, cudaMemcpyDeviceToHost);
    unsigned long long l0 = ~0ull, l1 = 0, h0min = ~0ull, h0max = 0, h1max = 0;
    for (auto& r : low) { l0 = std::min(l0, r.t0); l1 = std::max(l1, r.t1); }
    for (auto& r : high) { h0min = std::min(h0min, r.t0); h0max = std::max(h0max, r.t0); h1max = std::max(h1max, r.t1); }
    // low CTAs that STARTED while the high kernel was fully resident (between its last CTA start and its end)
    int low_during = 0;
    std::vector<char> smset(256, 0);
    for (auto& r : low) if (r.t0 > h0max && r.t0 < h1max - 50000 / 2) { ++low_during; smset[r.sm & 255] = 1; }
    int sms_during = 0;
    for (char c : smset) sms_during += c;
    printf("variant %d (high stream priority %s, %s launch): low grid %.1f us .. %.1f us; mark at %.1f us; high CTAs start %.1f .. %.1f us, end %.1f us; "
           "low CTAs started while high fully resident: %d on %d SMs\n",
           variant, prio ? "HIGH" : "same", coop ? "cooperative" : "plain", 0.0, (l1 - l0) / 1e3, ((long long)mark - (long long)l0) / 1e3,
           ((long long)h0min - (long long)l0) / 1e3, ((long long)h0max - (long long)l0) / 1e3, ((long long)h1max - (long long)l0) / 1e3, low_during, sms_during);
    cudaStreamDestroy(s_low);
    cudaStreamDestroy(s_high);
  }
  return 0;
}
