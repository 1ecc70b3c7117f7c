// "nucleus" post-processing for sm_100a (segment.py:52-101): per instance, an Otsu threshold of the raw
// intensities under the instance, `mask = instance & (raw > threshold)`, holes of the mask filled inside the
// instance's bounding box (scipy.ndimage.binary_fill_holes: background not reachable from the box border
// through face neighbours), instances written in ascending id order.
//
// Three device stages, all streaming / atomics-bound:
//   1. cb200_label_stats      per-label raw min / max and bounding box (atomics with a read-before-write test)
//   2. cb200_label_histogram  per-label histograms: one bin per value for integer images (skimage's bincount
//                             path), numpy-exact 256 bins over [min, max] for float images
//      (the O(bins) Otsu tail per label runs on the host, like detect's)
//   3. cb200_nucleus_fill     union-find over the NON-mask voxels of every bounding box, each box with a
//                             virtual "outside" node joined to its border; voxels of the mask and voxels not
//                             connected to the outside get the id (atomicMax = "later ids overwrite")
#include "common.cuh"
#include "unionfind.cuh"

namespace cb200 {

template <typename T>
__device__ __forceinline__ double raw_as_double(const T* p, int64_t i) {
  return (double)p[i];
}

// order-preserving map double -> uint64 (so that atomicMin / atomicMax work on the bits)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct LabelStats {
  unsigned long long* raw_min;  // ordered bits
  unsigned long long* raw_max;
  int* box;  // [label][6]: lo z, y, x, hi z, y, x (inclusive)
};

template <typename T>
__global__ void __launch_bounds__(256)
label_stats_kernel(const int32_t* __restrict__ seg, const T* __restrict__ raw, int64_t n, int ex, int ey, int max_label,
                   LabelStats s) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int32_t l = seg[i];
    if (l <= 0 || l > max_label) continue;
    const unsigned long long v = ordered_bits(raw_as_double<T>(raw, i));
    if (v < s.raw_min[l]) atomicMin(s.raw_min + l, v);
    if (v > s.raw_max[l]) atomicMax(s.raw_max + l, v);
    const int x = (int)(i % ex);
    const int64_t r = i / ex;
    const int c[3] = {(int)(r / ey), (int)(r % ey), x};
    int* b = s.box + (int64_t)l * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (c[k] < b[k]) atomicMin(b + k, c[k]);
      if (c[k] > b[3 + k]) atomicMax(b + 3 + k, c[k]);
    }
  }
}

__global__ void label_stats_init_kernel(int max_label, LabelStats s) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > max_label) return;
  s.raw_min[l] = ~0ull;
  s.raw_max[l] = 0ull;
  for (int k = 0; k < 3; ++k) {
    s.box[(int64_t)l * 6 + k] = INT32_MAX;
    s.box[(int64_t)l * 6 + 3 + k] = -1;
  }
}

__global__ void label_stats_decode_kernel(int max_label, LabelStats s, double* __restrict__ out_min,
                                          double* __restrict__ out_max) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > max_label) return;
  auto decode = [](unsigned long long o) {
    const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
  };
  const bool present = s.box[(int64_t)l * 6 + 3] >= 0;
  out_min[l] = present ? decode(s.raw_min[l]) : 0.0;
  out_max[l] = present ? decode(s.raw_max[l]) : 0.0;
}

// integer images: hist[offset[l] + (v - min[l])] += 1
template <typename T>
__global__ void __launch_bounds__(256)
label_bincount_kernel(const int32_t* __restrict__ seg, const T* __restrict__ raw, int64_t n, int max_label,
                      const double* __restrict__ raw_min, const int64_t* __restrict__ offset,
                      unsigned int* __restrict__ hist) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int32_t l = seg[i];
    if (l <= 0 || l > max_label) continue;
    const int64_t bin = (int64_t)raw[i] - (int64_t)raw_min[l];
    atomicAdd(hist + offset[l] + bin, 1u);
  }
}

// float images: np.histogram(values, nbins) over [min, max] of the label; edges[l][nbins + 1] as numpy
// builds them (host linspace in the image's dtype).  The bin is the one with edges[b] <= v < edges[b+1]
// (last bin closed), which is what numpy's index arithmetic plus its two corrections arrive at.
template <typename T>
__global__ void __launch_bounds__(256)
label_histogram_kernel(const int32_t* __restrict__ seg, const T* __restrict__ raw, int64_t n, int max_label,
                       const double* __restrict__ edges, int nbins, unsigned int* __restrict__ hist) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int32_t l = seg[i];
    if (l <= 0 || l > max_label) continue;
    const double v = raw_as_double<T>(raw, i);
    const double* e = edges + (int64_t)l * (nbins + 1);
    const double first = e[0], last = e[nbins];
    if (!(v >= first) || !(v <= last)) continue;  // NaN
    int idx = 0;
    if (last > first) {
      idx = (int)(((v - first) / (last - first)) * (double)nbins);
      idx = min(max(idx, 0), nbins - 1);
      while (idx > 0 && v < e[idx]) --idx;
      while (idx < nbins - 1 && v >= e[idx + 1]) ++idx;
    }
    atomicAdd(hist + (int64_t)l * nbins + idx, 1u);
  }
}

// ------------------------------------------------------------------ Otsu tail per label
// scikit-image's threshold_otsu after the histogram, one thread per label, in numpy's own arithmetic:
//   counts float32; weight1 = cumsum(counts), weight2 = cumsum(counts[::-1])[::-1]            (float32, sequential)
//   mean1 = cumsum(counts * centres) / weight1, mean2 likewise from the far end                (A)
//   variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2 ; first argmax
// A = float for float32 images (float32 * float32), double for integer / float64 images (float32 * int64 or
// float64 promotes to float64).  np.cumsum accumulates sequentially, so a sequential loop reproduces it bit
// for bit.  scratch_w / scratch_c hold the reversed cumulative sums (one entry per bin).
template <typename A>
__device__ __forceinline__ A otsu_mul(float c, double centre);
template <>
__device__ __forceinline__ float otsu_mul<float>(float c, double centre) { return __fmul_rn(c, (float)centre); }
template <>
__device__ __forceinline__ double otsu_mul<double>(float c, double centre) { return __dmul_rn((double)c, centre); }
__device__ __forceinline__ float otsu_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double otsu_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float otsu_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double otsu_div(double a, float b) { return __ddiv_rn(a, (double)b); }
__device__ __forceinline__ float otsu_var(float w1, float w2, float m1, float m2) {
  const float d = __fsub_rn(m1, m2);
  return __fmul_rn(__fmul_rn(w1, w2), __fmul_rn(d, d));
}
__device__ __forceinline__ double otsu_var(float w1, float w2, double m1, double m2) {
  const double d = __dsub_rn(m1, m2);
  return __dmul_rn((double)__fmul_rn(w1, w2), __dmul_rn(d, d));
}

template <typename A>
__global__ void __launch_bounds__(64)
label_otsu_kernel(const unsigned int* __restrict__ hist, const int64_t* __restrict__ offset,
                  const int64_t* __restrict__ nbins, const double* __restrict__ centre0,
                  const double* __restrict__ centres, int n_labels, float* __restrict__ scratch_w,
                  A* __restrict__ scratch_c, double* __restrict__ threshold) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_labels) return;
  const int64_t off = offset[l], nb = nbins[l];
  if (nb <= 0) return;
  // integer images: centre i = centre0[l] + i; float images: centres[off + i]
  auto centre = [&](int64_t i) { return centres ? centres[off + i] : centre0[l] + (double)i; };
  if (nb == 1) {
    threshold[l] = centre(0);
    return;
  }
  float w = 0.f;
  A cs = (A)0;
  for (int64_t i = nb - 1; i >= 0; --i) {  // reversed cumulative sums
    const float c = (float)hist[off + i];
    w = otsu_add(w, c);
    cs = otsu_add(cs, otsu_mul<A>(c, centre(i)));
    scratch_w[off + i] = w;
    scratch_c[off + i] = cs;
  }
  w = 0.f;
  cs = (A)0;
  A best = (A)0;
  int64_t best_i = 0;
  for (int64_t i = 0; i + 1 < nb; ++i) {
    const float c = (float)hist[off + i];
    w = otsu_add(w, c);
    cs = otsu_add(cs, otsu_mul<A>(c, centre(i)));
    const A m1 = otsu_div(cs, w);
    const float w2 = scratch_w[off + i + 1];
    const A m2 = otsu_div(scratch_c[off + i + 1], w2);
    const A v = otsu_var(w, w2, m1, m2);
    if (i == 0 || v > best) {  // np.argmax: first maximum (NaN cannot occur: both end bins are occupied)
      best = v;
      best_i = i;
    }
  }
  threshold[l] = centre(best_i);
}

// Integer images (one bin per value, integer bin centres, float64 arithmetic): every cumulative sum above is a sum of
// INTEGERS -- exact in float32 while a label has fewer than 2^24 pixels, exact in float64 always -- so the order of
// summation cannot change a bit and the sequential recurrences can be replaced by block-wide scans: one BLOCK per
// label, each thread a contiguous strip of bins, weight2 / the reversed sums as (total - prefix).  Labels with
// 2^24 pixels or more (float32 rounding would depend on the order) fall back to the sequential recurrence on
// thread 0.  Same thresholds, bit for bit, as label_otsu_kernel<double>.
constexpr int OTSU_BLOCK = 256;

template <typename T>
__device__ __forceinline__ T otsu_block_exclusive_scan(T v, T* s_warp, T& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T u = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  T base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < OTSU_BLOCK / 32; ++k) {
    const T t = s_warp[k];
    if (k < warp) base += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(OTSU_BLOCK)
label_otsu_block_kernel(const unsigned int* __restrict__ hist, const int64_t* __restrict__ offset,
                        const int64_t* __restrict__ nbins, const double* __restrict__ centre0, int n_labels,
                        float* __restrict__ scratch_w, double* __restrict__ scratch_c, double* __restrict__ threshold) {
  __shared__ double s_d[OTSU_BLOCK / 32];
  __shared__ unsigned long long s_u[OTSU_BLOCK / 32];
  __shared__ double s_best[OTSU_BLOCK / 32];
  __shared__ long long s_best_i[OTSU_BLOCK / 32];
  const int l = blockIdx.x;
  if (l >= n_labels) return;
  const int64_t off = offset[l], nb = nbins[l];
  if (nb <= 0) return;
  const double c0 = centre0[l];
  if (nb == 1) {
    if (threadIdx.x == 0) threshold[l] = c0;
    return;
  }
  const int64_t strip = (nb + OTSU_BLOCK - 1) / OTSU_BLOCK;
  const int64_t i0 = min((int64_t)threadIdx.x * strip, nb), i1 = min(i0 + strip, nb);
  unsigned long long w_loc = 0;
  double cs_loc = 0.0;
  for (int64_t i = i0; i < i1; ++i) {
    const unsigned c = hist[off + i];
    w_loc += c;
    cs_loc += (double)c * (c0 + (double)i);  // integer-valued: exact
  }
  unsigned long long w_tot;
  double cs_tot;
  unsigned long long w = otsu_block_exclusive_scan<unsigned long long>(w_loc, s_u, w_tot);
  double cs = otsu_block_exclusive_scan<double>(cs_loc, s_d, cs_tot);
  if (w_tot >= (1ull << 24)) {  // float32 weights would round: the order matters, keep numpy's sequential order
    if (threadIdx.x == 0) {
      float fw = 0.f;
      double fcs = 0.0;
      for (int64_t i = nb - 1; i >= 0; --i) {
        const float c = (float)hist[off + i];
        fw = otsu_add(fw, c);
        fcs = otsu_add(fcs, otsu_mul<double>(c, c0 + (double)i));
        scratch_w[off + i] = fw;
        scratch_c[off + i] = fcs;
      }
      fw = 0.f;
      fcs = 0.0;
      double best = 0.0;
      int64_t best_i = 0;
      for (int64_t i = 0; i + 1 < nb; ++i) {
        const float c = (float)hist[off + i];
        fw = otsu_add(fw, c);
        fcs = otsu_add(fcs, otsu_mul<double>(c, c0 + (double)i));
        const float w2 = scratch_w[off + i + 1];
        const double v = otsu_var(fw, w2, otsu_div(fcs, fw), otsu_div(scratch_c[off + i + 1], w2));
        if (i == 0 || v > best) {
          best = v;
          best_i = i;
        }
      }
      threshold[l] = c0 + (double)best_i;
    }
    return;
  }
  // second pass over the strip: variance12[i] for i in [i0, min(i1, nb - 1)), first maximum
  double best = -1.0;
  long long best_i = 0x7fffffffffffffffll;
  for (int64_t i = i0; i < i1 && i + 1 < nb; ++i) {
    const unsigned c = hist[off + i];
    w += c;
    cs += (double)c * (c0 + (double)i);
    const float w1 = (float)w, w2 = (float)(w_tot - w);  // exact: below 2^24
    const double v = otsu_var(w1, w2, otsu_div(cs, w1), otsu_div(cs_tot - cs, w2));
    if (v > best) {  // strict: the first maximum of the strip (variances are >= 0, NaN cannot occur)
      best = v;
      best_i = i;
    }
  }
  // block argmax: larger value wins, equal values keep the smaller index (np.argmax: first maximum)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_down_sync(FULL, best, o);
    const long long oi = __shfl_down_sync(FULL, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_best[threadIdx.x >> 5] = best;
    s_best_i[threadIdx.x >> 5] = best_i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < OTSU_BLOCK / 32; ++k) {
      if (s_best[k] > best || (s_best[k] == best && s_best_i[k] < best_i)) {
        best = s_best[k];
        best_i = s_best_i[k];
      }
    }
    threshold[l] = c0 + (double)best_i;
  }
}

// ------------------------------------------------------------------ hole filling
struct Boxes {
  int n;                      // instances
  const int32_t* ids;         // [n] ascending
  const double* thresholds;   // [n]
  const int32_t* box;         // [n][6] lo z,y,x, hi z,y,x (inclusive)
  const int64_t* offset;      // [n + 1] prefix of box volumes
};

struct BoxVoxel {
  int k;          // instance
  int z, y, x;    // position inside the box
  int bz, by, bx; // box extents
  int64_t pixel;  // linear index in the image
};

__device__ __forceinline__ BoxVoxel locate(const Boxes& b, int64_t j, int ex, int ey) {
  int lo = 0, hi = b.n - 1;  // last k with offset[k] <= j
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.offset[mid] <= j) lo = mid;
    else hi = mid - 1;
  }
  BoxVoxel v;
  v.k = lo;
  const int32_t* bb = b.box + (int64_t)lo * 6;
  v.bz = bb[3] - bb[0] + 1;
  v.by = bb[4] - bb[1] + 1;
  v.bx = bb[5] - bb[2] + 1;
  const int64_t local = j - b.offset[lo];
  v.x = (int)(local % v.bx);
  const int64_t r = local / v.bx;
  v.y = (int)(r % v.by);
  v.z = (int)(r / v.by);
  v.pixel = ((int64_t)(bb[0] + v.z) * ey + (bb[1] + v.y)) * ex + (bb[2] + v.x);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
fill_init_kernel(const int32_t* __restrict__ seg, const T* __restrict__ raw, Boxes b, int64_t total, int ex, int ey,
                 int* __restrict__ parent) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total + b.n; j += gs) {
    if (j >= total) {  // the virtual "outside" node of instance j - total
      parent[j] = (int)j;
      continue;
    }
    const BoxVoxel v = locate(b, j, ex, ey);
    const bool in_mask = seg[v.pixel] == b.ids[v.k] && raw_as_double<T>(raw, v.pixel) > b.thresholds[v.k];
    parent[j] = in_mask ? -1 : (int)j;
  }
}

__global__ void __launch_bounds__(256)
fill_merge_kernel(Boxes b, int64_t total, int ex, int ey, int num_dims, int* parent) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gs) {
    if (parent[j] < 0) continue;
    const BoxVoxel v = locate(b, j, ex, ey);
    // links already implied by the left neighbour's own links are skipped (see cc_merge_kernel)
    const bool left = v.x > 0 && parent[j - 1] >= 0;
    const bool edge_yz = v.y == 0 || v.y == v.by - 1 || (num_dims == 3 && (v.z == 0 || v.z == v.bz - 1));
    const bool border = v.x == 0 || v.x == v.bx - 1 || edge_yz;
    if (border && !(left && edge_yz)) uf_union(parent, (int)j, (int)(total + v.k));
    if (left) uf_union(parent, (int)j, (int)(j - 1));
    if (v.y > 0 && parent[j - v.bx] >= 0 && !(left && parent[j - v.bx - 1] >= 0))
      uf_union(parent, (int)j, (int)(j - v.bx));
    const int64_t slab = (int64_t)v.bx * v.by;
    if (v.z > 0 && parent[j - slab] >= 0 && !(left && parent[j - slab - 1] >= 0))
      uf_union(parent, (int)j, (int)(j - slab));
  }
}

__global__ void __launch_bounds__(256)
fill_write_kernel(Boxes b, int64_t total, int ex, int ey, const int* __restrict__ parent, int32_t* __restrict__ out) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gs) {
    const BoxVoxel v = locate(b, j, ex, ey);
    bool covered = parent[j] < 0;  // in the thresholded mask
    if (!covered) covered = uf_find(parent, (int)j) != uf_find(parent, (int)(total + v.k));  // a hole
    if (covered) atomicMax(out + v.pixel, b.ids[v.k]);
  }
}

struct StatsWorkspace {
  static int64_t bytes(int max_label) { return (int64_t)(max_label + 1) * (8 + 8) + 256; }
};

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_label_stats(const int32_t* seg, const void* raw, int raw_dtype, int num_dims, const int64_t* spatial,
                      int max_label, double* raw_min, double* raw_max, int32_t* box, void* workspace, void* stream) {
  if (!seg || !raw || !spatial || !raw_min || !raw_max || !box || !workspace || max_label < 0) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  int64_t n = 1;
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0 || spatial[k] > INT32_MAX) return CB200_EINVAL;
    n *= spatial[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int ex = (int)spatial[num_dims - 1], ey = (int)spatial[num_dims - 2];
  LabelStats s;
  s.raw_min = static_cast<unsigned long long*>(workspace);
  s.raw_max = s.raw_min + (max_label + 1);
  s.box = box;
  const int tb = (max_label + 1 + 255) / 256;
  label_stats_init_kernel<<<tb, 256, 0, st>>>(max_label, s);
  CB200_LAUNCH_CHECK();
  const int blocks = grid_for(n, 256, 4, 16);
#define CB200_STATS(T) label_stats_kernel<T><<<blocks, 256, 0, st>>>(seg, (const T*)raw, n, ex, ey, max_label, s)
  if (raw_dtype == CB200_F32) CB200_STATS(float);
  else if (raw_dtype == CB200_F64) CB200_STATS(double);
  else if (raw_dtype == CB200_U8) CB200_STATS(uint8_t);
  else if (raw_dtype == CB200_U16) CB200_STATS(uint16_t);
  else return CB200_EUNSUPPORTED;
#undef CB200_STATS
  CB200_LAUNCH_CHECK();
  label_stats_decode_kernel<<<tb, 256, 0, st>>>(max_label, s, raw_min, raw_max);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_label_stats_workspace_bytes(int max_label) { return max_label < 0 ? 0 : StatsWorkspace::bytes(max_label); }

int cb200_label_histogram(const int32_t* seg, const void* raw, int raw_dtype, int64_t n_pix, int max_label,
                          const double* raw_min, const int64_t* hist_offset, const double* edges, int nbins,
                          unsigned int* hist, void* stream) {
  if (!seg || !raw || !hist || n_pix < 0 || max_label < 0) return CB200_EINVAL;
  if (n_pix == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = grid_for(n_pix, 256, 4, 16);
  if (raw_dtype == CB200_U8 || raw_dtype == CB200_U16) {
    if (!raw_min || !hist_offset) return CB200_EINVAL;
    if (raw_dtype == CB200_U8)
      label_bincount_kernel<uint8_t><<<blocks, 256, 0, st>>>(seg, (const uint8_t*)raw, n_pix, max_label, raw_min,
                                                            hist_offset, hist);
    else
      label_bincount_kernel<uint16_t><<<blocks, 256, 0, st>>>(seg, (const uint16_t*)raw, n_pix, max_label, raw_min,
                                                             hist_offset, hist);
  } else if (raw_dtype == CB200_F32 || raw_dtype == CB200_F64) {
    if (!edges || nbins <= 0) return CB200_EINVAL;
    if (raw_dtype == CB200_F32)
      label_histogram_kernel<float><<<blocks, 256, 0, st>>>(seg, (const float*)raw, n_pix, max_label, edges, nbins, hist);
    else
      label_histogram_kernel<double><<<blocks, 256, 0, st>>>(seg, (const double*)raw, n_pix, max_label, edges, nbins,
                                                             hist);
  } else {
    return CB200_EUNSUPPORTED;
  }
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_label_otsu_workspace_bytes(int64_t total_bins) { return total_bins < 0 ? 0 : total_bins * 12 + 512; }

int cb200_label_otsu(const unsigned int* hist, const int64_t* hist_offset, const int64_t* num_bins,
                     const double* centre0, const double* centres, int arithmetic_dtype, int n_labels,
                     int64_t total_bins, double* thresholds, void* workspace, void* stream) {
  if (!hist || !hist_offset || !num_bins || !thresholds || !workspace || n_labels < 0 || total_bins < 0) return CB200_EINVAL;
  if (!centre0 && !centres) return CB200_EINVAL;
  if (n_labels == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* p = static_cast<uint8_t*>(workspace);
  float* sw = reinterpret_cast<float*>(p);
  void* sc = p + ((total_bins * 4 + 255) / 256) * 256;
  const int blocks = (n_labels + 63) / 64;
  if (arithmetic_dtype == CB200_F32)
    label_otsu_kernel<float><<<blocks, 64, 0, st>>>(hist, hist_offset, num_bins, centre0, centres, n_labels, sw,
                                                    static_cast<float*>(sc), thresholds);
  else if (arithmetic_dtype == CB200_F64 && !centres)  // integer image: exact sums, one block per label
    label_otsu_block_kernel<<<n_labels, OTSU_BLOCK, 0, st>>>(hist, hist_offset, num_bins, centre0, n_labels, sw,
                                                             static_cast<double*>(sc), thresholds);
  else if (arithmetic_dtype == CB200_F64)
    label_otsu_kernel<double><<<blocks, 64, 0, st>>>(hist, hist_offset, num_bins, centre0, centres, n_labels, sw,
                                                     static_cast<double*>(sc), thresholds);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_nucleus_fill_workspace_bytes(int64_t total_box_voxels, int n_instances) {
  if (total_box_voxels < 0 || n_instances < 0) return 0;
  return (total_box_voxels + n_instances) * 4 + 256;
}

int cb200_nucleus_fill(const int32_t* seg, const void* raw, int raw_dtype, int num_dims, const int64_t* spatial,
                       int n_instances, const int32_t* ids, const double* thresholds, const int32_t* boxes,
                       const int64_t* box_offset, int64_t total_box_voxels, int32_t* out, void* workspace,
                       void* stream) {
  if (!seg || !raw || !spatial || !out || n_instances < 0) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  int64_t n = 1;
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0 || spatial[k] > INT32_MAX) return CB200_EINVAL;
    n *= spatial[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  CB200_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int32_t) * (size_t)n, st));
  if (n_instances == 0 || total_box_voxels == 0) return CB200_OK;
  if (!ids || !thresholds || !boxes || !box_offset || !workspace) return CB200_EINVAL;
  if (total_box_voxels + n_instances > INT32_MAX) return CB200_EUNSUPPORTED;  // int parents
  const int ex = (int)spatial[num_dims - 1], ey = (int)spatial[num_dims - 2];
  Boxes b{n_instances, ids, thresholds, boxes, box_offset};
  int* parent = static_cast<int*>(workspace);
  const int blocks = grid_for(total_box_voxels + n_instances, 256, 2, 16);
#define CB200_FILL_INIT(T) \
  fill_init_kernel<T><<<blocks, 256, 0, st>>>(seg, (const T*)raw, b, total_box_voxels, ex, ey, parent)
  if (raw_dtype == CB200_F32) CB200_FILL_INIT(float);
  else if (raw_dtype == CB200_F64) CB200_FILL_INIT(double);
  else if (raw_dtype == CB200_U8) CB200_FILL_INIT(uint8_t);
  else if (raw_dtype == CB200_U16) CB200_FILL_INIT(uint16_t);
  else return CB200_EUNSUPPORTED;
#undef CB200_FILL_INIT
  CB200_LAUNCH_CHECK();
  fill_merge_kernel<<<blocks, 256, 0, st>>>(b, total_box_voxels, ex, ey, num_dims, parent);
  CB200_LAUNCH_CHECK();
  fill_write_kernel<<<blocks, 256, 0, st>>>(b, total_box_voxels, ex, ey, parent, out);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
