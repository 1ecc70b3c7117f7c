// Evaluation counts for sm_100a (evaluate.py:72-98).  The reference compares every (prediction id, ground-truth
// id) pair with four full-image passes; the same numbers come out of ONE pass: a contingency table of the
// two label images gives every intersection, its row / column sums give the areas, and
// union = area_p + area_g - intersection.  The O(P x G) float tail stays on the host (same numpy expressions
// as the reference, so F1 and SEG are bit-identical).
//   cb200_label_presence : present[v] = 1 for every label value v that occurs (the np.unique of :73-76)
//   cb200_contingency    : table[rank_p(pred[i])][rank_g(gt[i])]++ , warp-aggregated atomics
#include "common.cuh"

namespace cb200 {

// Both kernels read 16 bytes (8 uint16 / 4 int32 labels) per thread and work on RUNS of equal labels inside
// those 16 bytes -- label images are piecewise constant, so nearly every thread sees a single run.
template <typename T>
struct LabelVec {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};
template <typename T>
__device__ __forceinline__ LabelVec<T> load_labels(const T* p, int64_t vec_index) {
  LabelVec<T> r;
  *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p) + vec_index);
  return r;
}

// Label images are mostly one value (background) and a few thousand objects: marking `present[v]` straight in
// global memory makes every warp store to the same few bytes (measured 230 us for an 8 MB image: 36 GB/s).  Each
// block therefore collects the values it sees in a shared-memory bitmap -- a broadcast read per label, an atomicOr
// only the first time a bit is seen -- and writes the set bits out once at the end.  Values beyond the bitmap
// (int32 labels >= PRESENCE_BITS) take the direct path.
constexpr int PRESENCE_BITS = 1 << 16;  // covers every uint16 label
template <typename T>
__global__ void __launch_bounds__(256)
label_presence_kernel(const T* __restrict__ labels, int64_t n, int64_t nvec, int max_value, uint8_t* __restrict__ present) {
  constexpr int N = LabelVec<T>::N;
  __shared__ unsigned bits[PRESENCE_BITS / 32];
  const int n_words = (min(max_value, PRESENCE_BITS - 1) >> 5) + 1;
  for (int w = threadIdx.x; w < n_words; w += blockDim.x) bits[w] = 0u;
  __syncthreads();
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int last = -1;
  auto mark = [&](int v) {
    if (v == last || v < 0 || v > max_value) return;
    last = v;
    if (v < PRESENCE_BITS) {
      const unsigned bit = 1u << (v & 31);
      if (!(bits[v >> 5] & bit)) atomicOr(&bits[v >> 5], bit);
    } else if (!present[v]) {
      present[v] = 1;
    }
  };
  for (int64_t i = tid; i < nvec; i += gs) {
    const LabelVec<T> x = load_labels<T>(labels, i);
#pragma unroll
    for (int k = 0; k < N; ++k) mark((int)x.v[k]);
  }
  for (int64_t i = nvec * N + tid; i < n; i += gs) mark((int)labels[i]);  // tail
  __syncthreads();
  for (int w = threadIdx.x; w < n_words; w += blockDim.x) {
    unsigned m = bits[w];
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      present[w * 32 + b] = 1;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
contingency_kernel(const T* __restrict__ pred, const T* __restrict__ gt, int64_t n, int64_t nvec,
                   const int32_t* __restrict__ rank_p, const int32_t* __restrict__ rank_g, int max_value, int cols,
                   unsigned int* __restrict__ table) {
  constexpr int N = LabelVec<T>::N;
  constexpr unsigned NONE = 0xffffffffu;
  // a label outside [0, max_value] has no entry in the rank tables: it counts as background
  auto key_of = [&](T p, T g) {
    const int pv = (int)p, gv = (int)g;
    const unsigned rp = (pv >= 0 && pv <= max_value) ? (unsigned)rank_p[pv] : 0u;
    const unsigned rg = (gv >= 0 && gv <= max_value) ? (unsigned)rank_g[gv] : 0u;
    return rp * (unsigned)cols + rg;
  };
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const unsigned lane = threadIdx.x & 31;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // whole warps iterate together: the first run of every lane is added with one warp-aggregated atomic
  for (int64_t base = first - lane; base < nvec; base += gs) {
    const int64_t i = base + lane;
    unsigned key0 = NONE;
    int count0 = 0;
    if (i < nvec) {
      const LabelVec<T> p = load_labels<T>(pred, i), g = load_labels<T>(gt, i);
      unsigned key = key_of(p.v[0], g.v[0]);
      int count = 1;
      bool in_first = true;
      key0 = key;
#pragma unroll
      for (int k = 1; k < N; ++k) {
        if (p.v[k] == p.v[k - 1] && g.v[k] == g.v[k - 1]) {
          ++count;
          continue;
        }
        if (in_first) count0 = count;
        else atomicAdd(table + key, (unsigned)count);
        in_first = false;
        key = key_of(p.v[k], g.v[k]);
        count = 1;
      }
      if (in_first) count0 = count;
      else atomicAdd(table + key, (unsigned)count);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key0);
    const int total = __reduce_add_sync(peers, count0);
    if (key0 != NONE && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(table + key0, (unsigned)total);
  }
  for (int64_t i = nvec * N + first; i < n; i += gs)  // tail
    atomicAdd(table + key_of(pred[i], gt[i]), 1u);
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_label_presence(const void* labels, int dtype, int64_t n, int max_value, uint8_t* present, void* stream) {
  if (!labels || !present || n < 0 || max_value < 0) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(present, 0, (size_t)max_value + 1, st));
  if (n == 0) return CB200_OK;
  const bool aligned = reinterpret_cast<uintptr_t>(labels) % 16 == 0;
  if (dtype == CB200_U16) {
    const int64_t nvec = aligned ? n / 8 : 0;
    label_presence_kernel<uint16_t><<<grid_for(nvec + 256, 256, 4, 4), 256, 0, st>>>((const uint16_t*)labels, n, nvec, max_value, present);
  } else if (dtype == CB200_I32) {
    const int64_t nvec = aligned ? n / 4 : 0;
    label_presence_kernel<int32_t><<<grid_for(nvec + 256, 256, 4, 4), 256, 0, st>>>((const int32_t*)labels, n, nvec, max_value, present);
  } else {
    return CB200_EUNSUPPORTED;
  }
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_contingency(const void* pred, const void* gt, int dtype, int64_t n, const int32_t* rank_pred,
                      const int32_t* rank_gt, int max_value, int rows, int cols, unsigned int* table, void* stream) {
  if (!pred || !gt || !rank_pred || !rank_gt || !table || n < 0 || max_value < 0 || rows <= 0 || cols <= 0)
    return CB200_EINVAL;
  if ((int64_t)rows * cols >= 0xffffffffll) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(table, 0, sizeof(unsigned int) * (size_t)rows * cols, st));
  if (n == 0) return CB200_OK;
  const bool aligned = reinterpret_cast<uintptr_t>(pred) % 16 == 0 && reinterpret_cast<uintptr_t>(gt) % 16 == 0;
  if (dtype == CB200_U16) {
    const int64_t nvec = aligned ? n / 8 : 0;
    contingency_kernel<uint16_t><<<grid_for(nvec + 256, 256, 2, 8), 256, 0, st>>>(
        (const uint16_t*)pred, (const uint16_t*)gt, n, nvec, rank_pred, rank_gt, max_value, cols, table);
  } else if (dtype == CB200_I32) {
    const int64_t nvec = aligned ? n / 4 : 0;
    contingency_kernel<int32_t><<<grid_for(nvec + 256, 256, 2, 8), 256, 0, st>>>(
        (const int32_t*)pred, (const int32_t*)gt, n, nvec, rank_pred, rank_gt, max_value, cols, table);
  } else {
    return CB200_EUNSUPPORTED;
  }
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
