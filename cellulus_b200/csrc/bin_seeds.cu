// Grid-binned seeds: scikit-learn's get_bin_seeds (sklearn _mean_shift.py:254-297, MeanShift(bin_seeding=True),
// min_bin_freq = 1) on the device -- the "grid-binned" seeding of BASELINE configs[3].
//
//   binned = np.round(point / bin_size)          float64 division, round half to even   -> one integer per axis
//   seeds  = unique bins, as float32, * bin_size  (NumPy: a float32 array times a Python float stays float32)
//   if every point has its own bin the points themselves are the seeds (sklearn warns and returns X)
//
// Keys = the D bin indices packed into 63 bits -> cub radix sort -> heads of equal-key runs compacted -> decode.
// sklearn keeps the bins in first-seen order (a dict); the order of the seeds has no influence on the fitted
// centres (they are re-sorted by (count, coordinates), sklearn:530-534), so the seeds come out in key order.
#include <cub/device/device_radix_sort.cuh>

#include "compact.cuh"

namespace cb200 {

constexpr int BIN_BITS = 21;                 // per axis
constexpr long long BIN_OFFSET = 1ll << 20;  // |bin| < 2^20

template <int D>
__global__ void __launch_bounds__(256)
bin_keys_kernel(const double* __restrict__ points, int64_t n, int64_t stride, double bin_size,
                unsigned long long* __restrict__ keys, int* __restrict__ overflow) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    unsigned long long key = 0;
#pragma unroll
    for (int k = D - 1; k >= 0; --k) {
      const double b = rint(__ddiv_rn(__ldg(points + k * stride + i), bin_size));  // np.round(point / bin_size)
      if (!(fabs(b) < (double)BIN_OFFSET)) *overflow = 1;
      key = (key << BIN_BITS) | (unsigned long long)((long long)b + BIN_OFFSET);
    }
    keys[i] = key;
  }
}

struct KeyHeadPred {
  const unsigned long long* keys;
  __device__ __forceinline__ bool operator()(int64_t i) const { return i == 0 || keys[i] != keys[i - 1]; }
};
template <int D>
struct SeedEmit {
  const unsigned long long* keys;
  double* seeds;
  int64_t seed_stride;
  float bin_size;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const {
    unsigned long long key = keys[i];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const long long b = (long long)(key & ((1ull << BIN_BITS) - 1)) - BIN_OFFSET;
      key >>= BIN_BITS;
      seeds[k * seed_stride + d] = (double)__fmul_rn((float)b, bin_size);  // float32 bins * float32(bin_size)
    }
  }
};

template <int D>
__global__ void __launch_bounds__(256)
bin_fallback_kernel(const double* __restrict__ points, int64_t n, int64_t stride, double* __restrict__ seeds,
                    int64_t seed_stride, const long long* __restrict__ n_bins) {
  if (*n_bins != n) return;  // binning failed to merge anything: the points are the seeds (sklearn:288-293)
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
#pragma unroll
    for (int k = 0; k < D; ++k) seeds[k * seed_stride + i] = __ldg(points + k * stride + i);
  }
}

static size_t bin_align(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace cb200

using namespace cb200;

extern "C" int64_t cb200_bin_seeds_workspace_bytes(int64_t n_points) {
  if (n_points <= 0) return 512;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 (int)n_points);
  return (int64_t)(bin_align(sort_bytes) + 2 * bin_align(8 * (size_t)n_points) + bin_align((size_t)CompactWorkspace::bytes(n_points)) +
                   512);
}

extern "C" int cb200_bin_seeds(const double* points, int64_t n_points, int64_t pts_stride, int num_dims, double bin_size,
                               double* seeds, int64_t seed_stride, long long* n_out, int* overflow, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  if (!points || !seeds || !n_out || !overflow || !workspace || n_points < 0 || !(bin_size > 0.0)) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  if (n_points > INT32_MAX) return CB200_EUNSUPPORTED;
  if (seed_stride < n_points || workspace_bytes < cb200_bin_seeds_workspace_bytes(n_points)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(overflow, 0, sizeof(int), st));
  if (n_points == 0) {
    CB200_CUDA_TRY(cudaMemsetAsync(n_out, 0, sizeof(long long), st));
    return CB200_OK;
  }
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 (int)n_points);
  char* w = reinterpret_cast<char*>(bin_align(reinterpret_cast<size_t>(workspace)));
  void* sort_ws = w;                                        w += bin_align(sort_bytes);
  auto* keys_in = reinterpret_cast<unsigned long long*>(w); w += bin_align(8 * (size_t)n_points);
  auto* keys_out = reinterpret_cast<unsigned long long*>(w); w += bin_align(8 * (size_t)n_points);
  void* compact_ws = w;
  const int blocks = grid_for(n_points, 256, 2, 16);
  if (num_dims == 2) bin_keys_kernel<2><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, bin_size, keys_in, overflow);
  else bin_keys_kernel<3><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, bin_size, keys_in, overflow);
  CB200_LAUNCH_CHECK();
  CB200_CUDA_TRY(cub::DeviceRadixSort::SortKeys(sort_ws, sort_bytes, keys_in, keys_out, (int)n_points, 0,
                                                BIN_BITS * num_dims, st));
  const KeyHeadPred pred{keys_out};
  int rc;
  if (num_dims == 2)
    rc = run_compaction(pred, SeedEmit<2>{keys_out, seeds, seed_stride, (float)bin_size}, n_points, seed_stride, n_out,
                        compact_ws, st);
  else
    rc = run_compaction(pred, SeedEmit<3>{keys_out, seeds, seed_stride, (float)bin_size}, n_points, seed_stride, n_out,
                        compact_ws, st);
  if (rc != CB200_OK) return rc;
  if (num_dims == 2) bin_fallback_kernel<2><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, seeds, seed_stride, n_out);
  else bin_fallback_kernel<3><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, seeds, seed_stride, n_out);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}
