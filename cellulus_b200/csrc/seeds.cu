// Seed finder of detect.py:128-132 for sm_100a:
//   mag    = np.linalg.norm(centred[:-1], axis=0)            (float64, sequential sum of squares, sqrt)
//   smooth = scipy.ndimage.gaussian_filter(mag, sigma)       (separable, mode="reflect", truncate 4)
//   peaks  = skimage.feature.peak_local_max(-smooth)         (3^D maximum filter equality, > global min,
//                                                             1-px border excluded)
// The blur reproduces scipy's correlate1d arithmetic exactly (symmetric kernel: centre term first, then
// (left + right) * w from the outermost pair inwards, no FMA), so the equality test of the peak finder sees
// bit-identical values.  All three passes are streaming, HBM-bound.
#include "common.cuh"
#include "compact.cuh"

namespace cb200 {

template <typename T>
__global__ void __launch_bounds__(256)
channel_norm_kernel(const T* __restrict__ emb, int D, int64_t n, double* __restrict__ out) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) {
      const double v = load_as_double<T>(emb, (int64_t)k * n + i);
      s = __dadd_rn(s, __dmul_rn(v, v));
    }
    out[i] = sqrt(s);
  }
}

constexpr int BLUR_MAX_RADIUS = 64;
struct BlurWeights {
  double w[BLUR_MAX_RADIUS + 1];  // w[0] = centre, w[j] = weight at distance j
  int radius;
};

// scipy "reflect": (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_index(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// one axis of the separable filter: `inner` = product of the extents after the axis, `len` = extent of the axis
__global__ void __launch_bounds__(256)
blur_axis_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t total, int64_t inner, int len,
                 BlurWeights bw, bool negate) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gs) {
    const int64_t in_idx = i % inner;
    const int64_t rest = i / inner;
    const int pos = (int)(rest % len);
    const int64_t base = (rest / len) * len * inner + in_idx;
    double tmp = __dmul_rn(in[base + (int64_t)pos * inner], bw.w[0]);
    for (int j = bw.radius; j >= 1; --j) {
      const double l = in[base + (int64_t)reflect_index(pos - j, len) * inner];
      const double r = in[base + (int64_t)reflect_index(pos + j, len) * inner];
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(l, r), bw.w[j]));
    }
    out[i] = negate ? -tmp : tmp;
  }
}

// peak mask: image == max over the 3^D neighbourhood (edge-replicated), image > threshold, not on the 1-px border
template <int D>
__global__ void __launch_bounds__(256)
peak_mask_kernel(const double* __restrict__ img, int ex, int ey, int ez, double threshold, uint8_t* __restrict__ mask) {
  const int64_t n = (int64_t)ex * ey * ez;
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int x = (int)(i % ex);
    const int64_t r = i / ex;
    const int y = (int)(D == 2 ? r : r % ey);
    const int z = (int)(D == 2 ? 0 : r / ey);
    const double v = img[i];
    bool peak = v > threshold;
    peak = peak && x >= 1 && x < ex - 1 && y >= 1 && y < ey - 1 && (D == 2 || (z >= 1 && z < ez - 1));
    if (peak) {  // interior: all 3^D neighbours exist
      for (int dz = (D == 3 ? -1 : 0); dz <= (D == 3 ? 1 : 0) && peak; ++dz)
        for (int dy = -1; dy <= 1 && peak; ++dy)
          for (int dx = -1; dx <= 1; ++dx)
            if (img[((int64_t)(z + dz) * ey + (y + dy)) * ex + (x + dx)] > v) { peak = false; break; }
    }
    mask[i] = peak ? 1 : 0;
  }
}

struct MaskPred {
  const uint8_t* mask;
  __device__ __forceinline__ bool operator()(int64_t i) const { return mask[i] != 0; }
};
struct PeakEmit {  // linear index + value, raster order
  const double* img;
  int32_t* index;
  double* value;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const {
    index[d] = (int32_t)i;
    value[d] = img[i];
  }
};

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_channel_norm(const void* emb, int dtype, int num_channels, int64_t n_pix, double* out, void* stream) {
  if (!emb || !out || num_channels <= 0 || n_pix < 0) return CB200_EINVAL;
  if (n_pix == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = grid_for(n_pix, 256, 2, 16);
  if (dtype == CB200_F32) channel_norm_kernel<float><<<blocks, 256, 0, st>>>((const float*)emb, num_channels, n_pix, out);
  else if (dtype == CB200_F64) channel_norm_kernel<double><<<blocks, 256, 0, st>>>((const double*)emb, num_channels, n_pix, out);
  else return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_gaussian_blur(const double* in, double* out, double* scratch, int num_dims, const int64_t* spatial,
                        const double* weights /* host, radius + 1 */, int radius, int negate, void* stream) {
  if (!in || !out || !scratch || !spatial || !weights || radius < 0 || radius > BLUR_MAX_RADIUS) return CB200_EINVAL;
  if (num_dims < 1 || num_dims > 3) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  BlurWeights bw;
  bw.radius = radius;
  for (int j = 0; j <= radius; ++j) bw.w[j] = weights[j];
  int64_t total = 1;
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0 || spatial[k] > INT32_MAX) return CB200_EINVAL;
    total *= spatial[k];
  }
  // scipy filters axis 0 first, then 1, ...; ping-pong so that the LAST pass lands in `out`
  const double* src = in;
  double* bufs[2] = {(num_dims & 1) ? out : scratch, (num_dims & 1) ? scratch : out};
  const int blocks = grid_for(total, 256, 1, 16);
  for (int axis = 0; axis < num_dims; ++axis) {
    int64_t inner = 1;
    for (int k = axis + 1; k < num_dims; ++k) inner *= spatial[k];
    double* dst = bufs[axis & 1];
    blur_axis_kernel<<<blocks, 256, 0, st>>>(src, dst, total, inner, (int)spatial[axis], bw,
                                             negate && axis == num_dims - 1);
    CB200_LAUNCH_CHECK();
    src = dst;
  }
  return CB200_OK;
}

int64_t cb200_peaks_workspace_bytes(int64_t n_pix) { return n_pix + CompactWorkspace::bytes(n_pix) + 512; }

int cb200_local_peaks(const double* img, int num_dims, const int64_t* spatial, double threshold, int32_t* peak_index,
                      double* peak_value, int64_t capacity, long long* n_out, void* workspace, void* stream) {
  if (!img || !spatial || !peak_index || !peak_value || !n_out || !workspace) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  int64_t n = 1;
  for (int k = 0; k < num_dims; ++k) n *= spatial[k];
  if (n > INT32_MAX) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* mask = static_cast<uint8_t*>(workspace);
  void* cws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(mask + n) + 255) / 256 * 256);
  const int ex = (int)spatial[num_dims - 1], ey = (int)spatial[num_dims - 2], ez = num_dims == 3 ? (int)spatial[0] : 1;
  const int blocks = grid_for(n, 256, 2, 16);
  if (num_dims == 2) peak_mask_kernel<2><<<blocks, 256, 0, st>>>(img, ex, ey, ez, threshold, mask);
  else peak_mask_kernel<3><<<blocks, 256, 0, st>>>(img, ex, ey, ez, threshold, mask);
  CB200_LAUNCH_CHECK();
  MaskPred pred{mask};
  PeakEmit emit{img, peak_index, peak_value};
  return run_compaction(pred, emit, n, capacity, n_out, cws, st);
}

}  // extern "C"
