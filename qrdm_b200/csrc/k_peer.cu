// k_peer.cu — peer memory of the row-sharded path (SURVEY.md §8e "fusion with the collective") and the
// one-shot LL all-reduce that replaces ncclAllReduce for the latency-bound vectors.
//
// The reference has no counterpart (single process, src/dgeqrdm_work.c); this is the transport under the
// per-column reductions of its panel loop (src/dgeqr2.c:148-189: dnrm2 at src/dlarfg.c:126, dgemv at
// src/dlarf.c:173-184) and under the small per-iteration vectors (candidate Gram src/dgeqrdm_work.c:376-379,
// V'C of src/dlarfb.c:130-143, the norm downdate :81-108) when the rows of A live on several GPUs.
//
// One process per GPU.  Every rank allocates ONE receive buffer, exports it with cudaIpcGetMemHandle, the
// launcher all-gathers the 64-byte handles (qrdm_b200/sharded.py), and qrdm_rt_peer_open maps the peers'
// buffers (cudaIpcOpenMemHandle, which also enables peer access).  From then on kernels exchange 16-byte LL
// packets {lo, tag, hi, tag} by plain stores into the PEER's buffer over NVLink and polls of their OWN buffer:
//
//   k_peer_allreduce   y[i] = sum_r x_r[i] for up to QRDM_PEER_CAP doubles in ONE kernel: thread i stores its
//                      value into slot [parity][rank][i] of every peer, then polls its own slots [parity][r][i]
//                      and adds them in rank order — the same order on every rank, so all ranks hold bit-identical
//                      sums (the replicated-decision scheme of the sharded driver needs exactly that).
//                      Latency = one NVLink store + one local L2 poll; no barrier, no fence, no host round trip.
//   k_panel_tall<true> (k_panel.cu) does the same exchange from INSIDE the persistent panel kernel, once per column.
//
// Slot reuse: two parities (generic: sequence number & 1; panel: a device-side exchange counter & 1).  A rank can only send for exchange e+2 after it has finished exchange e+1, which needed
// every peer's e+1 contribution, which each peer sends only after it has finished READING exchange e (stream / program
// order) — so a slot is never overwritten before its last reader is done.  This needs consecutive PERFORMED exchanges
// to alternate parity: a kernel that draws a sequence number (qrdm_peer_next_gen) must perform the exchange even when
// it has nothing to send (k_sub_w2<true> exchanges one dummy element for a dead sub-panel).  Tags never repeat within 2^32
// exchanges and the buffers start zeroed (tag 0 is never used).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ll.cuh"

namespace {
struct PeerState {
  bool exported = false, open = false;
  PeerCtx ctx{};
  void* mapped[QRDM_PEER_MAXR] = {};  // IPC mappings to close (NULL for the local buffer)
  unsigned gen_seq = 0;               // generic all-reduce sequence number (tag), identical on every rank
  LLPacket* local = nullptr;
  unsigned* xseq = nullptr;
} g_peer;
}  // namespace

const PeerCtx* qrdm_peer_ctx() { return g_peer.open ? &g_peer.ctx : nullptr; }
void qrdm_peer_next_gen(unsigned* tag, int* parity) {
  g_peer.gen_seq = g_peer.gen_seq + 1 == 0 ? 1u : g_peer.gen_seq + 1;
  *tag = g_peer.gen_seq;
  *parity = (int)(g_peer.gen_seq & 1u);
}

__global__ void __launch_bounds__(256) k_peer_allreduce(double* __restrict__ buf, int count, PeerCtx pc, unsigned tag, int parity) {
  const int me = pc.rank, N = pc.nranks;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const double v = buf[i];
    for (int r = 0; r < N; ++r)
      if (r != me) ll_store(peer_gen_slot(pc.recv[r], parity, me, (size_t)i), v, tag);
    // all polls in flight together, then a fixed-order sum (rank 0 first): bit-identical on every rank
    unsigned lo[QRDM_PEER_MAXR], t0[QRDM_PEER_MAXR], hi[QRDM_PEER_MAXR], t1[QRDM_PEER_MAXR];
#pragma unroll
    for (int r = 0; r < QRDM_PEER_MAXR; ++r) {
      if (r < N && r != me) {
        const LLPacket* p = peer_gen_slot(pc.recv[me], parity, r, (size_t)i);
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo[r]), "=r"(t0[r]), "=r"(hi[r]), "=r"(t1[r]) : "l"(p) : "memory");
      }
    }
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < QRDM_PEER_MAXR; ++r) {
      if (r >= N) break;
      if (r == me) { s += v; continue; }
      unsigned spins = 0;
      while (t0[r] != tag || t1[r] != tag) {
        const LLPacket* p = peer_gen_slot(pc.recv[me], parity, r, (size_t)i);
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo[r]), "=r"(t0[r]), "=r"(hi[r]), "=r"(t1[r]) : "l"(p) : "memory");
        if (++spins > LL_SPIN_LIMIT) __trap();
      }
      s += __longlong_as_double((long long)(((unsigned long long)hi[r] << 32) | lo[r]));
    }
    buf[i] = s;
  }
}

extern "C" {

int qrdm_rt_peer_available(void) { return g_peer.open ? g_peer.ctx.nranks : 0; }

// In-place sum over the ranks of buf[0..count), any count (pieces of QRDM_PEER_CAP).  Returns 0, or -1 without peers.
int qrdm_k_peer_allreduce(double* buf, size_t count, void* stream) {
  if (!g_peer.open) return -1;
  if (g_peer.ctx.nranks == 1) return 0;
  for (size_t off = 0; off < count; off += QRDM_PEER_CAP) {
    const int cnt = (int)(count - off < (size_t)QRDM_PEER_CAP ? count - off : (size_t)QRDM_PEER_CAP);
    unsigned tag = 0;
    int parity = 0;
    qrdm_peer_next_gen(&tag, &parity);
    int grid = (cnt + 255) / 256;
    if (grid > 592) grid = 592;
    k_peer_allreduce<<<grid, 256, 0, (cudaStream_t)stream>>>(buf + off, cnt, g_peer.ctx, tag, parity);
    QRDM_LAUNCH_CHECK();
  }
  return 0;
}

// Allocate (once) the local receive buffer and return its 64-byte IPC handle.
int qrdm_rt_peer_export(char* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!g_peer.local) {
    if (cudaMalloc((void**)&g_peer.local, QRDM_PEER_BYTES) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (cudaMalloc((void**)&g_peer.xseq, 64) != cudaSuccess) { cudaGetLastError(); return -1; }
  }
  if (cudaMemset(g_peer.xseq, 0, 64) != cudaSuccess) return -1;
  if (cudaMemset(g_peer.local, 0, QRDM_PEER_BYTES) != cudaSuccess) return -1;  // synchronous: zeroed before anybody can write
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, g_peer.local) != cudaSuccess) {
    fprintf(stderr, "qrdm_b200: cudaIpcGetMemHandle failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  memcpy(out64, &h, 64);
  g_peer.exported = true;
  return 0;
}

static void peer_unmap() {
  for (int r = 0; r < QRDM_PEER_MAXR; ++r) {
    if (g_peer.mapped[r]) cudaIpcCloseMemHandle(g_peer.mapped[r]);
    g_peer.mapped[r] = nullptr;
  }
  g_peer.open = false;
}

// handles64: nranks x 64 bytes, entry r = the handle rank r exported.  Every rank must call this after ALL ranks have
// exported (the all-gather of the handles guarantees it), and nobody may start a sharded call before all ranks returned
// from it (the launcher's barrier).
int qrdm_rt_peer_open(int rank, int nranks, const char* handles64) {
  if (nranks < 1 || nranks > QRDM_PEER_MAXR || rank < 0 || rank >= nranks) {
    fprintf(stderr, "qrdm_b200: peer exchange supports 1..%d ranks (one node)\n", QRDM_PEER_MAXR);
    return -1;
  }
  if (!g_peer.exported || !g_peer.local) { fprintf(stderr, "qrdm_b200: qrdm_b200_peer_handle must be called first\n"); return -1; }
  peer_unmap();
  memset(&g_peer.ctx, 0, sizeof(g_peer.ctx));
  g_peer.ctx.rank = rank;
  g_peer.ctx.nranks = nranks;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { g_peer.ctx.recv[r] = g_peer.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles64 + (size_t)r * 64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      fprintf(stderr, "qrdm_b200: cudaIpcOpenMemHandle(rank %d) failed: %s\n", r, cudaGetErrorString(e));
      cudaGetLastError();
      peer_unmap();
      return -1;
    }
    g_peer.mapped[r] = p;
    g_peer.ctx.recv[r] = (LLPacket*)p;
  }
  g_peer.ctx.xseq = g_peer.xseq;
  g_peer.gen_seq = 0;
  g_peer.open = true;
  return 0;
}

int qrdm_rt_peer_destroy(void) {
  peer_unmap();
  if (g_peer.local) cudaFree(g_peer.local);
  if (g_peer.xseq) cudaFree(g_peer.xseq);
  g_peer.local = nullptr;
  g_peer.xseq = nullptr;
  g_peer.exported = false;
  return 0;
}

}  // extern "C"
