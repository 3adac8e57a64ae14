// Occupancy evaluation on the device: the confusion
// matrices of Metric_mIoU.add_batch (mmdet3d/datasets/occ_metrics.py:93-157)
// accumulated in place, so the eval loop no longer needs a .cpu().numpy()
// round trip per sample.  Integer counting: bit-exact against the oracle.
#include "common.cuh"
#include "../../include/preworld_b200.h"

namespace {

constexpr int MAX_CL = 32;

// hist[gt*n_cl + pred] += 1 for voxels with mask != 0 (if a mask is given) and
// gt < n_cl (255 = unlabelled is skipped, occ_metrics.py:106-113);
// occ_hist[(gt != free)*2 + (pred != free)] += 1 for every masked voxel
// (compute_IoU on the binarised grids, occ_metrics.py:147-151: no label test,
// so gt == 255 counts as occupied there).
__global__ void __launch_bounds__(256)
occ_confusion_kernel(const unsigned char* __restrict__ pred, const unsigned char* __restrict__ gt,
                     const unsigned char* __restrict__ mask, long long n, int n_cl, int free_idx,
                     unsigned long long* __restrict__ hist, unsigned long long* __restrict__ occ) {
  __shared__ unsigned int s_hist[MAX_CL * MAX_CL];
  __shared__ unsigned int s_occ[4];
  for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x < 4) s_occ[threadIdx.x] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (mask != nullptr && mask[i] == 0) continue;
    const int g = gt[i], p = pred[i];
    if (g < n_cl && p < n_cl) atomicAdd(&s_hist[g * n_cl + p], 1u);
    atomicAdd(&s_occ[(g != free_idx ? 2 : 0) + (p != free_idx ? 1 : 0)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
  if (threadIdx.x < 4 && s_occ[threadIdx.x])
    atomicAdd(occ + threadIdx.x, (unsigned long long)s_occ[threadIdx.x]);
}

}  // namespace

PW_API int pw_occ_confusion(const unsigned char* pred, const unsigned char* gt,
                            const unsigned char* mask, long long n, int n_cl, int free_idx,
                            long long* hist, long long* occ_hist, void* stream) {
  PW_REQUIRE(n >= 0 && n_cl > 0 && n_cl <= MAX_CL && free_idx >= 0);
  if (n == 0) return 0;
  PW_REQUIRE(pred && gt && hist && occ_hist);
  // one CTA sees at most 2^32-1 voxels per bin: grid-stride chunks stay far below
  int blocks = (int)min((long long)148 * 4, (n + 255) / 256);
  occ_confusion_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      pred, gt, mask, n, n_cl, free_idx, reinterpret_cast<unsigned long long*>(hist),
      reinterpret_cast<unsigned long long*>(occ_hist));
  PW_LAUNCH_CHECK(); pw_count_launch(1);
  return 0;
}
