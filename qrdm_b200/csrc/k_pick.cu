// k_pick.cu — K3c (greedy delta-orthogonality pick) + the column-permutation plan, and K3d (the
// batched full-height column moves).
//
// Replaces: the greedy loop of DM_perm (reference src/dgeqrdm_work.c:382-403) and
// permute_marked (src/dgeqrdm_work.c:149-262, nz = 0).  The plan reproduces the reference's
// exchange sequence exactly — including its quirks: selected columns already inside the leading
// fjb slots stay put, the result is NOT in norm order, "pointless" exchanges of two selected
// columns still happen, and vn2 is never exchanged — because jpvt parity depends on all of them.
// The exchanges are then composed into disjoint cycles so that K3d moves every affected column
// exactly once (bytes = 16*m per moved column) with (row-chunks x cycles) parallelism.
#include "pick_body.cuh"

__global__ void __launch_bounds__(256) k_pick(qrdm_prob P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  qrdm_pick_body(P, *reinterpret_cast<PickShared*>(smem_raw));
  // statistics for the HBM roofline of K3d: columns moved so far (every one is read once and written once)
  __syncthreads();
  if (threadIdx.x == 0) P.ctrl->stat_perm_cols += P.ctrl->cyc_start[P.ctrl->ncyc];
}

// K3d: rotate the full-height columns along each cycle.  grid = (row chunks, cycles).
__global__ void __launch_bounds__(256) k_permute(qrdm_prob P) {
  const qrdm_ctrl* ctrl = P.ctrl;
  const int cyc = blockIdx.y;
  if (cyc >= ctrl->ncyc) return;
  const int j = ctrl->j;
  const int b = ctrl->cyc_start[cyc], e = ctrl->cyc_start[cyc + 1];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.m) return;
  double* a = P.a + (size_t)j * P.lda + r;
  const double first = a[(size_t)ctrl->cyc_pos[b] * P.lda];
  double nxt = a[(size_t)ctrl->cyc_pos[b + 1] * P.lda];
  for (int k = b; k < e - 1; ++k) {
    const double v = nxt;
    if (k + 2 < e) nxt = a[(size_t)ctrl->cyc_pos[k + 2] * P.lda];
    a[(size_t)ctrl->cyc_pos[k] * P.lda] = v;
  }
  a[(size_t)ctrl->cyc_pos[e - 1] * P.lda] = first;
}

extern "C" int qrdm_k_pick(const qrdm_prob* p, void* stream) {
  static int attr_gen = -1;  // per-device attribute, see qrdm_rt_device_generation
  if (attr_gen != qrdm_rt_device_generation()) {
    cudaFuncSetAttribute(k_pick, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PickShared));
    attr_gen = qrdm_rt_device_generation();
  }
  k_pick<<<1, 256, sizeof(PickShared), (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}

extern "C" int qrdm_k_permute(const qrdm_prob* p, void* stream) {
  // at most 2*nb exchanges -> at most 2*nb cycles of length >= 2
  const int maxcyc = 2 * (p->nb < QRDM_KMAX ? p->nb : QRDM_KMAX);
  dim3 grid((p->m + 255) / 256, maxcyc);
  k_permute<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
  QRDM_LAUNCH_CHECK();
  return 0;
}
