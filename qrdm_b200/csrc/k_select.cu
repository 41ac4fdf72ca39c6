// k_select.cu — K3a: iteration prologue + stable top-k_max selection of the partial column norms.
//
// Replaces: the qsort of (norm, index) pairs and the tau_-filter of DM_perm, reference
// src/dgeqrdm_work.c:23-26, 126-142, 345-351, plus the loop bookkeeping of :694-703 and the
// max-norm of norm_update (:65,106-107) that the stop rule (:782-785) needs.
//
// Only the first k_max <= 64 entries of the reference's full descending sort are ever consumed,
// so a top-k selection suffices: radix-select on the IEEE bit pattern (norms are >= +0, so the
// unsigned order of the bits is the order of the values) narrows the active columns to <= 1024
// survivors, which are then ranked exactly with the reference's tie rule — glibc's qsort is a
// stable merge sort, i.e. equal norms keep ascending column index.
#include "select_body.cuh"

#define SEL_THREADS 1024

__global__ void __launch_bounds__(SEL_THREADS) k_select(qrdm_prob P) {
  __shared__ SelShared S;
  qrdm_select_body<SEL_THREADS>(P, S);
}

extern "C" int qrdm_k_select(const qrdm_prob* p, void* stream) {
  k_select<<<1, SEL_THREADS, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
