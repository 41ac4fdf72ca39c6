// Cross-CTA latency probe: LL-packet ping-pong between CTA 0 and CTA b (different SMs), and an
// all-to-one gather like the panel kernel's reduce step.
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(16) LLPacket { unsigned lo, tag0, hi, tag1; };
__device__ __forceinline__ void ll_store(LLPacket* p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ double ll_load(const LLPacket* p, unsigned tag) {
  unsigned lo, t0, hi, t1;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p) : "memory");
  } while (t0 != tag || t1 != tag);
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ void ll_store_relaxed(LLPacket* p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"((unsigned)b), "r"(tag), "r"((unsigned)(b >> 32)), "r"(tag) : "memory");
}
__device__ __forceinline__ double ll_load_relaxed(const LLPacket* p, unsigned tag) {
  unsigned lo, t0, hi, t1;
  do {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p) : "memory");
  } while (t0 != tag || t1 != tag);
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
template <bool RELAXED>
__global__ void pingpong(LLPacket* buf, long long* cyc, int partner, int iters, unsigned base) {
  if (threadIdx.x != 0) return;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (RELAXED) { ll_store_relaxed(&buf[0], 1.0, base + i); ll_load_relaxed(&buf[1], base + i); }
      else { ll_store(&buf[0], 1.0, base + i); ll_load(&buf[1], base + i); }
    }
    cyc[0] = clock64() - t0;
  } else if (blockIdx.x == partner) {
    for (int i = 0; i < iters; ++i) {
      if (RELAXED) { ll_load_relaxed(&buf[0], base + i); ll_store_relaxed(&buf[1], 2.0, base + i); }
      else { ll_load(&buf[0], base + i); ll_store(&buf[1], 2.0, base + i); }
    }
  }
}
// gather: every CTA publishes one packet per step; CTA 0 waits for all G, then publishes a total that all wait for
__global__ void gather(LLPacket* part, LLPacket* bc, long long* cyc, int iters, unsigned base) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const unsigned tag = base + i;
    if (tid == 0) ll_store(&part[(i & 1) * 256 + b], 1.0, tag);
    if (b == 0) {
      double v = 0.0;
      if (tid < G) v = ll_load(&part[(i & 1) * 256 + tid], tag);
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __shared__ double sr[8];
      if ((tid & 31) == 0) sr[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) { double t = 0; for (int w = 0; w < 8; ++w) t += sr[w]; ll_store(&bc[i & 1], t, tag); }
    }
    if (tid == 0) ll_load(&bc[i & 1], tag);
    __syncthreads();
  }
  if (b == 0 && tid == 0) cyc[0] = clock64() - t0;
}
int main() {
  LLPacket* buf; long long* cyc;
  cudaMalloc(&buf, 16 * 1024); cudaMemset(buf, 0, 16 * 1024); cudaMallocManaged(&cyc, 64);
  const int iters = 2000;
  unsigned base = 1000;
  for (int partner : {1, 2, 37, 74, 111, 147}) {
    pingpong<false><<<148, 32>>>(buf, cyc, partner, iters, base); cudaDeviceSynchronize(); base += iters + 10;
    long long v = cyc[0];
    pingpong<true><<<148, 32>>>(buf, cyc, partner, iters, base); cudaDeviceSynchronize(); base += iters + 10;
    printf("pingpong CTA0 <-> CTA%-3d: volatile %.0f cycles/round trip, relaxed.gpu %.0f\n", partner, (double)v / iters, (double)cyc[0] / iters);
  }
  for (int G : {16, 64, 128}) {
    gather<<<G, 256>>>(buf + 16, buf + 600, cyc, iters, base); cudaDeviceSynchronize(); base += iters + 10;
    printf("gather+broadcast over %3d CTAs: %.0f cycles/step\n", G, (double)cyc[0] / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
