// Evaluation counts for sm_100a (evaluate.py:72-98).  The reference compares every (prediction id, ground-truth
// id) pair with four full-image passes; the same numbers come out of ONE pass: a contingency table of the
// two label images gives every intersection, its row / column sums give the areas, and
// union = area_p + area_g - intersection.  The O(P x G) float tail stays on the host (same numpy expressions
// as the reference, so F1 and SEG are bit-identical).
//   cb200_label_presence : present[v] = 1 for every label value v that occurs (the np.unique of :73-76)
//   cb200_contingency    : table[rank_p(pred[i])][rank_g(gt[i])]++ , warp-aggregated atomics
#include "common.cuh"

namespace cb200 {

template <typename T>
__global__ void __launch_bounds__(256)
label_presence_kernel(const T* __restrict__ labels, int64_t n, int max_value, uint8_t* __restrict__ present) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  int last = -1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int v = (int)labels[i];
    if (v == last || v < 0 || v > max_value) continue;
    last = v;
    if (!present[v]) present[v] = 1;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
contingency_kernel(const T* __restrict__ pred, const T* __restrict__ gt, int64_t n, const int32_t* __restrict__ rank_p,
                   const int32_t* __restrict__ rank_g, int cols, unsigned int* __restrict__ table) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // whole warps iterate together (the tail lanes carry an impossible key)
  for (int64_t base = first - (threadIdx.x & 31); base < n; base += gs) {
    const int64_t i = base + (threadIdx.x & 31);
    unsigned key = 0xffffffffu;
    if (i < n) key = (unsigned)rank_p[(int)pred[i]] * (unsigned)cols + (unsigned)rank_g[(int)gt[i]];
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key != 0xffffffffu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1))
      atomicAdd(table + key, (unsigned)__popc(peers));
  }
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_label_presence(const void* labels, int dtype, int64_t n, int max_value, uint8_t* present, void* stream) {
  if (!labels || !present || n < 0 || max_value < 0) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(present, 0, (size_t)max_value + 1, st));
  if (n == 0) return CB200_OK;
  const int blocks = grid_for(n, 256, 8, 8);
  if (dtype == CB200_U16) label_presence_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)labels, n, max_value, present);
  else if (dtype == CB200_I32) label_presence_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)labels, n, max_value, present);
  else return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_contingency(const void* pred, const void* gt, int dtype, int64_t n, const int32_t* rank_pred,
                      const int32_t* rank_gt, int rows, int cols, unsigned int* table, void* stream) {
  if (!pred || !gt || !rank_pred || !rank_gt || !table || n < 0 || rows <= 0 || cols <= 0) return CB200_EINVAL;
  if ((int64_t)rows * cols >= 0xffffffffll) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(table, 0, sizeof(unsigned int) * (size_t)rows * cols, st));
  if (n == 0) return CB200_OK;
  const int blocks = grid_for(n, 256, 8, 8);
  if (dtype == CB200_U16)
    contingency_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)pred, (const uint16_t*)gt, n, rank_pred, rank_gt, cols, table);
  else if (dtype == CB200_I32)
    contingency_kernel<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)pred, (const int32_t*)gt, n, rank_pred, rank_gt, cols, table);
  else return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
