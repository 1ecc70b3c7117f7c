// "cell" post-processing for sm_100a (segment.py:42-52):
//
//   distance_foreground = dtedt(segmentation == 0)
//   expanded_mask       = distance_foreground < grow_distance
//   distance_background = dtedt(expanded_mask)
//   segmentation[distance_background < shrink_distance] = 0
//
// Only the comparison `edt < radius` is ever used, so the exact Euclidean distance transform is evaluated
// as a separable lower envelope restricted to a window of ceil(radius) pixels per axis: pass 1 finds the
// squared distance to the nearest zero element in the row, every further pass takes
// min_d (g(p + d * stride) + d^2).  The squared distances are exact integers, the final test is
// sqrt((double)d2) < radius -- the value scipy compares.  An input WITHOUT any zero element follows scipy's
// behaviour: distances are measured to a virtual zero at index (-1, 0, ..., 0).
// All passes stream the volume once (the +-R taps are L1/L2 hits): HBM-bound.
#include "common.cuh"

namespace cb200 {

constexpr unsigned EDT_INF = 0x3fffffffu;

struct MaskInside {  // EDT input = a uint8 mask, nonzero = inside
  const uint8_t* mask;
  __device__ __forceinline__ bool operator()(int64_t i) const { return mask[i] != 0; }
};
struct BackgroundInside {  // EDT input = (segmentation == 0)
  const int32_t* seg;
  __device__ __forceinline__ bool operator()(int64_t i) const { return seg[i] == 0; }
};

template <typename Inside>
__global__ void __launch_bounds__(256)
edt_row_kernel(Inside inside, int64_t n, int ex, int R, unsigned* __restrict__ g, int* any_zero) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  bool saw_zero = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    if (!inside(i)) {
      g[i] = 0u;
      saw_zero = true;
      continue;
    }
    const int x = (int)(i % ex);
    unsigned best = EDT_INF;
    for (int d = 1; d <= R; ++d) {
      if ((x - d >= 0 && !inside(i - d)) || (x + d < ex && !inside(i + d))) {
        best = (unsigned)d * (unsigned)d;
        break;
      }
    }
    g[i] = best;
  }
  if (__any_sync(0xffffffffu, saw_zero) && (threadIdx.x & 31) == 0) *any_zero = 1;
}

__device__ __forceinline__ unsigned edt_envelope(const unsigned* __restrict__ g, int64_t i, int pos, int len,
                                                 int64_t inner, int R) {
  unsigned best = g[i];
  for (int d = 1; d <= R; ++d) {
    const unsigned dd = (unsigned)d * (unsigned)d;
    if (dd >= best) break;  // no tap further out can win
    if (pos - d >= 0) best = min(best, g[i - d * inner] + dd);
    if (pos + d < len) best = min(best, g[i + d * inner] + dd);
  }
  return best;
}

__global__ void __launch_bounds__(256)
edt_axis_kernel(const unsigned* __restrict__ g_in, int64_t n, int64_t inner, int len, int R,
                unsigned* __restrict__ g_out) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int pos = (int)((i / inner) % len);
    g_out[i] = edt_envelope(g_in, i, pos, len, inner, R);
  }
}

struct EdtShape {
  int ext[3];  // extents, slowest axis first; unused leading entries are 1
};

// last axis pass fused with the comparison; CLEAR = zero `seg` where the test holds instead of writing a mask
template <bool CLEAR>
__global__ void __launch_bounds__(256)
edt_final_kernel(const unsigned* __restrict__ g_in, int64_t n, int64_t inner, int len, int R, EdtShape shape,
                 int num_dims, double radius, const int* __restrict__ any_zero, uint8_t* __restrict__ out, int32_t* __restrict__ seg) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const bool has_zero = *any_zero != 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    bool within;
    if (has_zero) {
      const int pos = (int)((i / inner) % len);
      const unsigned d2 = edt_envelope(g_in, i, pos, len, inner, R);
      within = d2 < EDT_INF && sqrt((double)d2) < radius;
    } else {  // scipy: no zero element anywhere -> distances to the virtual element (-1, 0, ..., 0)
      const int64_t x = i % shape.ext[2];
      const int64_t y = (i / shape.ext[2]) % shape.ext[1];
      const int64_t z = i / ((int64_t)shape.ext[2] * shape.ext[1]);
      // the first axis of the array carries the +1
      const long long d2 = num_dims == 3 ? (z + 1) * (z + 1) + y * y + x * x : (y + 1) * (y + 1) + x * x;
      within = sqrt((double)d2) < radius;
    }
    if constexpr (CLEAR) {
      if (within) seg[i] = 0;
    } else {
      out[i] = within ? 1 : 0;
    }
  }
}

struct EdtWorkspace {
  unsigned* g0;
  unsigned* g1;
  uint8_t* mask;
  int* flags;  // [0], [1]: any-zero flag of the first / second transform
  static int64_t bytes(int64_t n) { return 2 * ((n * 4 + 255) / 256 * 256) + (n + 255) / 256 * 256 + 256; }
  EdtWorkspace(void* ws, int64_t n) {
    uint8_t* p = static_cast<uint8_t*>(ws);
    const int64_t gbytes = (n * 4 + 255) / 256 * 256;
    g0 = reinterpret_cast<unsigned*>(p);
    g1 = reinterpret_cast<unsigned*>(p + gbytes);
    mask = p + 2 * gbytes;
    flags = reinterpret_cast<int*>(p + 2 * gbytes + (n + 255) / 256 * 256);
  }
};

template <typename Inside, bool CLEAR>
static int edt_within_impl(Inside inside, int num_dims, const int64_t* spatial, double radius, uint8_t* out,
                           int32_t* seg, EdtWorkspace& w, int* flag, cudaStream_t st) {
  int64_t n = 1;
  EdtShape shape = {{1, 1, 1}};
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0 || spatial[k] > INT32_MAX) return CB200_EINVAL;
    shape.ext[3 - num_dims + k] = (int)spatial[k];
    n *= spatial[k];
  }
  if (!(radius < 16384.0)) return CB200_EUNSUPPORTED;  // squared distances are kept in 30 bits
  const int R = radius > 0.0 ? (int)ceil(radius) : 0;
  CB200_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
  const int blocks = grid_for(n, 256, 2, 16);
  const int ex = shape.ext[2], ey = shape.ext[1], ez = shape.ext[0];
  edt_row_kernel<Inside><<<blocks, 256, 0, st>>>(inside, n, ex, R, w.g0, flag);
  CB200_LAUNCH_CHECK();
  const unsigned* src = w.g0;
  if (num_dims == 3) {
    edt_axis_kernel<<<blocks, 256, 0, st>>>(w.g0, n, ex, ey, R, w.g1);
    CB200_LAUNCH_CHECK();
    src = w.g1;
  }
  // last pass: the slowest axis (y in 2-D, z in 3-D)
  const int64_t inner = num_dims == 3 ? (int64_t)ex * ey : ex;
  const int len = num_dims == 3 ? ez : ey;
  edt_final_kernel<CLEAR><<<blocks, 256, 0, st>>>(src, n, inner, len, R, shape, num_dims, radius, flag, out, seg);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_edt_workspace_bytes(int64_t n_pix) { return n_pix < 0 ? 0 : EdtWorkspace::bytes(n_pix); }

int cb200_edt_within(const uint8_t* input, int num_dims, const int64_t* spatial, double radius, uint8_t* out,
                     void* workspace, void* stream) {
  if (!input || !spatial || !out || !workspace) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  int64_t n = 1;
  for (int k = 0; k < num_dims; ++k) n *= spatial[k];
  EdtWorkspace w(workspace, n);
  return edt_within_impl<MaskInside, false>(MaskInside{input}, num_dims, spatial, radius, out, nullptr, w, w.flags,
                                            (cudaStream_t)stream);
}

int cb200_grow_shrink(int32_t* seg, int num_dims, const int64_t* spatial, double grow_distance,
                      double shrink_distance, void* workspace, void* stream) {
  if (!seg || !spatial || !workspace) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  int64_t n = 1;
  for (int k = 0; k < num_dims; ++k) n *= spatial[k];
  EdtWorkspace w(workspace, n);
  cudaStream_t st = (cudaStream_t)stream;
  // expanded_mask = dtedt(segmentation == 0) < grow_distance                      (segment.py:47-48)
  int rc = edt_within_impl<BackgroundInside, false>(BackgroundInside{seg}, num_dims, spatial, grow_distance, w.mask,
                                                    nullptr, w, w.flags, st);
  if (rc != CB200_OK) return rc;
  // segmentation[dtedt(expanded_mask) < shrink_distance] = 0                      (segment.py:49-50)
  return edt_within_impl<MaskInside, true>(MaskInside{w.mask}, num_dims, spatial, shrink_distance, nullptr, seg, w,
                                           w.flags + 1, st);
}

}  // extern "C"
