// ll.cuh — "LL" packets: the barrier-free exchange primitive of the panel kernels (between the CTAs of one GPU,
// k_panel.cu) and of the row-sharded path (between GPUs, over NVLink peer memory: k_peer.cu, k_panel_tall<true>).
#pragma once
#include "common.cuh"

// Cross-CTA exchange without barriers: every value travels as a 16-byte "LL" packet
// {lo32, tag, hi32, tag} (the scheme NCCL's low-latency protocol uses): an aligned 8-byte store is
// atomic, so a reader that sees the expected tag in both halves has the payload — no fence, no
// counter, one L2 round trip.  tag = (panel launch epoch << 8) + step + 1 is never reused.
// Per column the reduction is a reduce-scatter + broadcast: CTA (jj mod G) gathers the G partials of
// column jj, sums them in a fixed order (deterministic) and publishes the total; everybody then
// polls the 64 totals + the 64 entries of the pivot row published by the CTA that owns that row.
// Traffic per column ~ G*64 packets instead of the G*G*64 of an all-to-all read, latency two
// round trips, and identical inputs on every CTA => bit-identical decisions (stop test, tau).
struct __align__(16) LLPacket { unsigned lo, tag0, hi, tag1; };
#define LL_SPIN_LIMIT (1u << 27)  // polls (~0.5 us each) before a waiting thread gives up with a trap

__device__ __forceinline__ void ll_store(LLPacket* p, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"((unsigned)b), "r"(tag),
               "r"((unsigned)(b >> 32)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ double ll_load(const LLPacket* p, unsigned tag) {
  unsigned lo, t0, hi, t1, spins = 0;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p) : "memory");
    if (++spins > LL_SPIN_LIMIT) __trap();  // a partner CTA never showed up: fail loudly instead of hanging the GPU
  } while (t0 != tag || t1 != tag);
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// Sum over the partials of all G CTAs for one column, lane l taking CTAs l, l+32, ...  A lane has up to
// ceil(148/32) = 5 packets to read: all loads are issued first and only then checked (re-polling the ones that
// had not arrived), so they cost ONE L2 round trip instead of one each — the blocking ll_load in a loop
// serialised them, which is where the 0.012 us per CTA of the panel's per-column cost came from.
// The packets are added in the same order as before: results are bit-identical.
__device__ __forceinline__ double ll_gather_sum(const LLPacket* base, size_t stride, int lane, int G, unsigned tag) {
  constexpr int MAXU = (QRDM_PANEL_MAXCTA + 31) / 32;
  unsigned lo[MAXU], t0[MAXU], hi[MAXU], t1[MAXU];
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int c = lane + 32 * u;
    if (c < G)
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                   : "=r"(lo[u]), "=r"(t0[u]), "=r"(hi[u]), "=r"(t1[u]) : "l"(base + (size_t)c * stride) : "memory");
  }
  double v = 0.0;
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int c = lane + 32 * u;
    if (c < G) {
      unsigned spins = 0;
      while (t0[u] != tag || t1[u] != tag) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(lo[u]), "=r"(t0[u]), "=r"(hi[u]), "=r"(t1[u]) : "l"(base + (size_t)c * stride) : "memory");
        if (++spins > LL_SPIN_LIMIT) __trap();
      }
      v += __longlong_as_double((long long)(((unsigned long long)hi[u] << 32) | lo[u]));
    }
  }
  return v;
}

// all-gather variant: partials of CTAs start, start+step, ... (<= 5 of them), loads issued together
__device__ __forceinline__ double ll_gather_sum_strided(const LLPacket* base, size_t stride, int start, int step, int G, unsigned tag) {
  constexpr int MAXU = 5;
  unsigned lo[MAXU], t0[MAXU], hi[MAXU], t1[MAXU];
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int c = start + step * u;
    if (c < G)
      asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                   : "=r"(lo[u]), "=r"(t0[u]), "=r"(hi[u]), "=r"(t1[u]) : "l"(base + (size_t)c * stride) : "memory");
  }
  double v = 0.0;
#pragma unroll
  for (int u = 0; u < MAXU; ++u) {
    const int c = start + step * u;
    if (c < G) {
      unsigned spins = 0;
      while (t0[u] != tag || t1[u] != tag) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(lo[u]), "=r"(t0[u]), "=r"(hi[u]), "=r"(t1[u]) : "l"(base + (size_t)c * stride) : "memory");
        if (++spins > LL_SPIN_LIMIT) __trap();
      }
      v += __longlong_as_double((long long)(((unsigned long long)hi[u] << 32) | lo[u]));
    }
  }
  return v;
}

// ---- peer memory of the row-sharded path (one process per GPU; buffers exported/mapped with CUDA IPC) ----
// Every rank owns one receive buffer; rank r's buffer is mapped at recv[r] in every process (recv[rank] is the
// local allocation).  A sender stores LL packets straight into its peers' buffers over NVLink; a receiver polls
// its OWN buffer only, so a wait is a local L2 access.  Layout of a receive buffer (LLPacket units):
//   panel region  [2 parity][QRDM_PEER_MAXR sender][128]          per panel column: 64 sums + 64 pivot-row entries
//   generic region[2 parity][QRDM_PEER_MAXR sender][QRDM_PEER_CAP] one-shot all-reduce of up to CAP doubles
#define QRDM_PEER_MAXR 8
#define QRDM_PEER_CAP 49152
#define QRDM_PEER_PANEL_PKTS (2 * QRDM_PEER_MAXR * 128)
#define QRDM_PEER_GEN_PKTS ((size_t)2 * QRDM_PEER_MAXR * QRDM_PEER_CAP)
#define QRDM_PEER_BYTES ((QRDM_PEER_PANEL_PKTS + QRDM_PEER_GEN_PKTS) * sizeof(LLPacket))
struct PeerCtx {
  int rank, nranks;
  LLPacket* recv[QRDM_PEER_MAXR];
  unsigned* xseq;  // LOCAL device counter of the in-kernel exchanges performed so far (same value on every rank)
};
__device__ __forceinline__ LLPacket* peer_panel_slot(LLPacket* base, int parity, int sender, int e) {
  return base + ((size_t)parity * QRDM_PEER_MAXR + sender) * 128 + e;
}
__device__ __forceinline__ LLPacket* peer_gen_slot(LLPacket* base, int parity, int sender, size_t e) {
  return base + QRDM_PEER_PANEL_PKTS + ((size_t)parity * QRDM_PEER_MAXR + sender) * QRDM_PEER_CAP + e;
}
// Stores that cross NVLink use the very same instructions as the on-GPU exchange — st/ld.volatile.global.v4.u32,
// the 16-byte {data, flag, data, flag} line of NCCL's LL protocol: each 8-byte half carries its own tag, so a
// reader that sees the tag in both halves has the payload whatever the arrival order of the halves.
// host side (k_peer.cu)
const PeerCtx* qrdm_peer_ctx();          // NULL until qrdm_rt_peer_open succeeded
void qrdm_peer_next_gen(unsigned* tag, int* parity);  // next exchange of the generic region (same sequence on every rank)
