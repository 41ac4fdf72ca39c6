// rt.cu — CUDA runtime wrappers with C linkage so the host driver (dgeqrdm_host.c) stays plain C,
// plus the FP64 peak micro-benchmark bench.py uses as the roofline denominator of K6.
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

long long g_qrdm_launches = 0;
static int g_qrdm_dev_gen = 0;

extern "C" {

int qrdm_rt_malloc(void** ptr, size_t bytes) { return (int)cudaMalloc(ptr, bytes); }
int qrdm_rt_free(void* ptr) { return (int)cudaFree(ptr); }
int qrdm_rt_host_alloc(void** ptr, size_t bytes) { return (int)cudaMallocHost(ptr, bytes); }
int qrdm_rt_host_free(void* ptr) { return (int)cudaFreeHost(ptr); }
int qrdm_rt_memset(void* ptr, int v, size_t bytes, void* stream) {
  return (int)cudaMemsetAsync(ptr, v, bytes, (cudaStream_t)stream);
}
int qrdm_rt_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
int qrdm_rt_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
int qrdm_rt_h2d_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream) {
  return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
int qrdm_rt_d2h_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, void* stream) {
  return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
int qrdm_rt_stream_create(void** stream) { return (int)cudaStreamCreateWithFlags((cudaStream_t*)stream, cudaStreamNonBlocking); }
// prio > 0: the device's greatest priority, prio <= 0: its least (B200: 0 = least, also the default; -5 = greatest).
// The block scheduler serves pending CTAs of the highest-priority stream first and keeps SMs free for them even when a
// lower-priority CTA would fit (tools/prio_probe.cu): the look-ahead's side stream lives on that.
int qrdm_rt_stream_create_prio(void** stream, int prio) {
  int least = 0, greatest = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamCreateWithPriority((cudaStream_t*)stream, cudaStreamNonBlocking, prio > 0 ? greatest : least);
}
int qrdm_rt_is_pinned(const void* ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
  return at.type == cudaMemoryTypeHost ? 1 : 0;
}
int qrdm_rt_stream_wait_event(void* stream, void* ev) { return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0); }
int qrdm_rt_sync(void* stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }
int qrdm_rt_event_create(void** ev) { return (int)cudaEventCreate((cudaEvent_t*)ev); }
int qrdm_rt_event_destroy(void* ev) { return ev ? (int)cudaEventDestroy((cudaEvent_t)ev) : 0; }
int qrdm_rt_event_record(void* ev, void* stream) { return (int)cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream); }
int qrdm_rt_event_sync(void* ev) { return (int)cudaEventSynchronize((cudaEvent_t)ev); }
double qrdm_rt_event_ms(void* ev0, void* ev1) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, (cudaEvent_t)ev0, (cudaEvent_t)ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}
int qrdm_rt_device_info(int* sm_count, size_t* free_bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return (int)e;
  if (prop.major < 10) {
    fprintf(stderr, "qrdm_b200: device %s is sm_%d%d; this library is built for sm_100a only\n", prop.name, prop.major, prop.minor);
    return (int)cudaErrorInvalidDevice;
  }
  *sm_count = prop.multiProcessorCount;
  size_t tot = 0;
  e = cudaMemGetInfo(free_bytes, &tot);
  return (int)e;
}
int qrdm_rt_set_device(int dev) { return (int)cudaSetDevice(dev); }
int qrdm_rt_get_device(int* dev) { return (int)cudaGetDevice(dev); }
int qrdm_rt_device_generation(void) { return g_qrdm_dev_gen; }
void qrdm_rt_new_device_generation(void) { ++g_qrdm_dev_gen; }
int qrdm_rt_stream_destroy(void* stream) { return stream ? (int)cudaStreamDestroy((cudaStream_t)stream) : 0; }
const char* qrdm_rt_errstr(int code) { return cudaGetErrorString((cudaError_t)code); }
long long qrdm_rt_launch_count(void) { return g_qrdm_launches; }

}  // extern "C"

// ---- FP64 peak: dependent-free chains of DFMA or DMMA.8x8x4 on every SM ----
template <bool DMMA>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (DMMA) dmma884(c[i][0], c[i][1], a, b);
      else { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" double qrdm_rt_fp64_peak(int use_dmma, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * sms * 2 * 256) != cudaSuccess) return -1.0;
  const int iters = 8000, grid = sms * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, s);
    if (use_dmma) k_fp64_peak<true><<<grid, 256, 0, s>>>(out, iters, 1.0000001, 1e-9);
    else k_fp64_peak<false><<<grid, 256, 0, s>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const double flops = use_dmma ? 2.0 * 256 * 8 * (double)iters * 8.0 * grid   /* 8 warps x 8 mma x 256 fma */
                                : 2.0 * 16 * (double)iters * 256.0 * grid;
  return flops / (best * 1e-3) / 1e12;
}

// ---- device-to-device copy bandwidth (read + written bytes per second): the HBM roofline denominator bench.py measures
// live next to MEASURED_PEAKS.json's hbm_gbs ----
__global__ void __launch_bounds__(256) k_copy_bw(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
extern "C" double qrdm_rt_copy_gbs(size_t bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (bytes < ((size_t)1 << 20)) bytes = (size_t)1 << 20;
  bytes &= ~(size_t)15;
  double2 *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, bytes) != cudaSuccess) { cudaGetLastError(); return -1.0; }
  if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaGetLastError(); cudaFree(a); return -1.0; }
  cudaMemsetAsync(a, 0, bytes, s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, s);
    k_copy_bw<<<sms * 8, 256, 0, s>>>(a, b, bytes / 16);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  return 2.0 * (double)bytes / (best * 1e-3) / 1e9;
}

// ---- NCCL, bound at run time (dlopen) so that libqrdm_b200.so has no link-time dependency and a
// process that already loaded torch's bundled libnccl.so.2 shares that copy ----
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1;
} g_nccl;

bool nccl_load() {
  if (g_nccl.handle) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) { fprintf(stderr, "qrdm_b200: cannot dlopen libnccl.so.2: %s\n", dlerror()); return false; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) {
    fprintf(stderr, "qrdm_b200: libnccl.so.2 lacks the expected symbols\n");
    return false;
  }
  return true;
}
int nccl_check(ncclResult_t r, const char* what) {
  if (r == ncclSuccess) return 0;
  fprintf(stderr, "qrdm_b200: NCCL error in %s: %s\n", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return -1;
}
}  // namespace

extern "C" {
int qrdm_rt_comm_unique_id(char* out128) {
  if (!nccl_load()) return -1;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId")) return -1;
  memcpy(out128, &id, 128);
  return 0;
}
int qrdm_rt_comm_init(int rank, int nranks, const char* id128) {
  if (!nccl_load()) return -1;
  if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (nccl_check(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank), "ncclCommInitRank")) return -1;
  g_nccl.nranks = nranks;
  return 0;
}
int qrdm_rt_comm_destroy(void) {
  if (g_nccl.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_nccl.comm);
  g_nccl.comm = nullptr;
  g_nccl.nranks = 1;
  return 0;
}
int qrdm_rt_allreduce(double* buf, size_t count, void* stream) {
  if (!g_nccl.comm) { fprintf(stderr, "qrdm_b200: all-reduce without a communicator (call qrdm_b200_comm_init)\n"); return -1; }
  return nccl_check(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, g_nccl.comm, (cudaStream_t)stream), "ncclAllReduce");
}
}  // extern "C"
