// Detect preamble for sm_100a: min/max + numpy-exact histogram (Otsu input),
// foreground thresholding and order-preserving stream compaction of the
// foreground pixels into a float64 SoA point set with coordinates added
// (detect.py:88-94, utils/mean_shift.py:15-36,85,94).  All of it is one- or
// two-touch streaming over the (D+1, *S) embedding volume: HBM-bound.
#include "common.cuh"
#include "compact.cuh"

namespace cb200 {

// ---------------------------------------------------------------------------
// min / max
// ---------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = CB200_SM_COUNT * 8;

struct ReduceWorkspace {
  unsigned int ticket;
  unsigned int pad;
  double lo[RED_MAX_BLOCKS];
  double hi[RED_MAX_BLOCKS];
};

template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
minmax_kernel(const T* __restrict__ x, int64_t n, const long long* __restrict__ n_dev, double* __restrict__ out2,
              ReduceWorkspace* ws) {
  if (n_dev) n = min(n, (int64_t)__ldg(n_dev));  // the element count lives on the device: n is only its upper bound
  double lo = INFINITY, hi = -INFINITY;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = load_as_double<T>(x, i);
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
    hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
  }
  __shared__ double s_lo[RED_THREADS / 32], s_hi[RED_THREADS / 32];
  __shared__ bool s_last;
  if (lane_id() == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < RED_THREADS / 32; ++i) {
      lo = fmin(lo, s_lo[i]);
      hi = fmax(hi, s_hi[i]);
    }
    ws->lo[blockIdx.x] = lo;
    ws->hi[blockIdx.x] = hi;
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    lo = INFINITY;
    hi = -INFINITY;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      lo = fmin(lo, *((volatile double*)&ws->lo[i]));
      hi = fmax(hi, *((volatile double*)&ws->hi[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
      hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
    }
    __syncthreads();
    if (lane_id() == 0) {
      s_lo[threadIdx.x >> 5] = lo;
      s_hi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < RED_THREADS / 32; ++i) {
        lo = fmin(lo, s_lo[i]);
        hi = fmax(hi, s_hi[i]);
      }
      out2[0] = lo;
      out2[1] = hi;
      ws->ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// np.histogram(x, nbins, range=(edges[0], edges[nbins])) -- uniform-bin fast
// path of numpy/lib/_histograms_impl.py, bit for bit, in float64:
//   f = ((x - first) / (last - first)) * nbins; idx = trunc(f); idx == nbins -> nbins-1;
//   x < edges[idx] -> idx-1;  x >= edges[idx+1] and idx != nbins-1 -> idx+1
// ---------------------------------------------------------------------------
constexpr int HIST_THREADS = 256;
constexpr int HIST_MAX_BINS = 1024;

template <typename T>
__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const T* __restrict__ x, int64_t n, const double* __restrict__ edges, int nbins,
                 unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned char smem_raw[];
  double* s_edges = reinterpret_cast<double*>(smem_raw);                         // nbins + 1
  unsigned int* s_hist = reinterpret_cast<unsigned int*>(s_edges + nbins + 1);   // warps x nbins
  const int warps = HIST_THREADS / 32;
  for (int i = threadIdx.x; i <= nbins; i += blockDim.x) s_edges[i] = edges[i];
  for (int i = threadIdx.x; i < warps * nbins; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const double first = s_edges[0], last = s_edges[nbins];
  const double denom = last - first;
  const double numer = (double)nbins;
  unsigned int* my_hist = s_hist + (threadIdx.x >> 5) * nbins;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = load_as_double<T>(x, i);
    if (!(v >= first) || !(v <= last)) continue;
    const double f = __dmul_rn(__ddiv_rn(__dsub_rn(v, first), denom), numer);
    int idx = (int)f;
    if (idx == nbins) idx -= 1;
    if (v < s_edges[idx]) idx -= 1;
    if (v >= s_edges[idx + 1] && idx != nbins - 1) idx += 1;
    atomicAdd(my_hist + idx, 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
    unsigned long long c = 0;
    for (int w = 0; w < warps; ++w) c += s_hist[w * nbins + i];
    if (c) atomicAdd(counts + i, c);
  }
}

// ---------------------------------------------------------------------------
// centring of detect.py:97-119: per offset channel, the mean over the NON-ZERO entries of mask * channel
// ---------------------------------------------------------------------------
struct CentreWorkspace {
  unsigned int ticket;
  unsigned int pad;
  double sum[3][RED_MAX_BLOCKS];
  double cnt[3][RED_MAX_BLOCKS];
};

template <typename T, int D>
__global__ void __launch_bounds__(RED_THREADS)
masked_channel_mean_kernel(const T* __restrict__ emb, int64_t n, double threshold, double* __restrict__ means,
                           CentreWorkspace* ws) {
  const T* std_channel = emb + (int64_t)D * n;
  double s[D], c[D];
#pragma unroll
  for (int k = 0; k < D; ++k) s[k] = c[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (load_as_double<T>(std_channel, i) < threshold) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double v = load_as_double<T>(emb, (int64_t)k * n + i);
        if (v != 0.0) {
          s[k] += v;
          c[k] += 1.0;
        }
      }
    }
  }
  __shared__ double sh[2 * D][RED_THREADS / 32];
  __shared__ bool s_last;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double a = warp_sum_down(s[k]), b = warp_sum_down(c[k]);
    if (lane_id() == 0) {
      sh[k][threadIdx.x >> 5] = a;
      sh[D + k][threadIdx.x >> 5] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) {
      double a = 0.0, b = 0.0;
      for (int w = 0; w < RED_THREADS / 32; ++w) {
        a += sh[k][w];
        b += sh[D + k][w];
      }
      ws->sum[k][blockIdx.x] = a;
      ws->cnt[k][blockIdx.x] = b;
    }
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < D) {  // fixed-order final sum: deterministic
    __threadfence();
    const int k = threadIdx.x;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < (int)gridDim.x; ++i) {
      a += *((volatile double*)&ws->sum[k][i]);
      b += *((volatile double*)&ws->cnt[k][i]);
    }
    means[k] = a / b;  // numpy: mean of an empty selection is nan (0/0) -- same here
    if (k == 0) ws->ticket = 0u;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
centre_kernel(const T* __restrict__ emb, int D, int64_t n, const double* __restrict__ means, T* __restrict__ out) {
  const int64_t total = (int64_t)(D + 1) * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = (int)(i / n);
    const double v = load_as_double<T>(emb, i);
    out[i] = (T)(ch < D ? v - means[ch] : v);  // the std channel is copied unchanged
  }
}

// foreground predicate: (double) std < threshold   (detect.py:94 / utils/mean_shift.py:34)
template <typename T>
struct FgPred {
  const T* std_channel;
  double threshold;
  __device__ __forceinline__ bool operator()(int64_t i) const { return load_as_double<T>(std_channel, i) < threshold; }
};

// point emit: X[k][dst] = emb[k][pix] + coordinate_k in float64 (exact), channel 0 = x = last axis
template <typename T, int D>
struct FgEmit {
  const T* emb;
  int64_t n_pix;
  int ext[3];  // (x, y, z) extents
  double* points;
  int64_t capacity;
  int32_t* pix_index;
  __device__ __forceinline__ void operator()(int64_t i, long long dst) const {
    const unsigned pix = (unsigned)i;
    const unsigned x = pix % (unsigned)ext[0];
    const unsigned r = pix / (unsigned)ext[0];
    unsigned c[3];
    c[0] = x;
    if constexpr (D == 2) {
      c[1] = r;
    } else {
      c[1] = r % (unsigned)ext[1];
      c[2] = r / (unsigned)ext[1];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) points[k * capacity + dst] = load_as_double<T>(emb, k * n_pix + i) + (double)c[k];
    if (pix_index) pix_index[dst] = (int32_t)pix;
  }
};

template <typename T, typename M>
__global__ void __launch_bounds__(256)
mask_kernel(const T* __restrict__ std_channel, double threshold, int64_t n, M* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    mask[i] = load_as_double<T>(std_channel, i) < threshold ? (M)1 : (M)0;
}

struct FlagPred {
  const uint8_t* flags;
  const long long* n_dev;  // optional: only the first *n_dev flags are meaningful
  __device__ __forceinline__ bool operator()(int64_t i) const {
    if (n_dev && i >= __ldg(n_dev)) return false;
    return __ldg(flags + i) != 0;
  }
};
struct SelectEmit {
  const double* src;
  int64_t src_stride;
  double* dst;
  int64_t dst_stride;
  int D;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const {
    for (int k = 0; k < D; ++k) dst[k * dst_stride + d] = __ldg(src + k * src_stride + i);
  }
};

__global__ void __launch_bounds__(256) bernoulli_kernel(uint8_t* __restrict__ flags, int64_t n, double p, uint64_t seed) {
  const Philox rng(seed);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // one Philox call yields two 53-bit uniforms -> two flags
  const int64_t n2 = (n + 1) / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const uint4 r = rng((uint64_t)i, 0x42524e4cull);
    const double u0 = (double)((((uint64_t)r.x << 32) | r.y) >> 11) * (1.0 / 9007199254740992.0);
    const double u1 = (double)((((uint64_t)r.z << 32) | r.w) >> 11) * (1.0 / 9007199254740992.0);
    flags[2 * i] = u0 < p;
    if (2 * i + 1 < n) flags[2 * i + 1] = u1 < p;
  }
}

template <typename T>
static int fg_compact_typed(const void* emb_v, int num_dims, const int64_t* spatial, double threshold, double* points,
                            int32_t* pix_index, int64_t capacity, long long* n_out, void* mask_out, int mask_dtype,
                            void* workspace, cudaStream_t st) {
  const T* emb = static_cast<const T*>(emb_v);
  int64_t n_pix = 1;
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0) return CB200_EINVAL;
    n_pix *= spatial[k];
  }
  if (n_pix > INT32_MAX) return CB200_EUNSUPPORTED;
  const T* std_channel = emb + (int64_t)num_dims * n_pix;
  if (mask_out) {
    const int blocks = grid_for(n_pix, 256, 4, 16);
    if (mask_dtype == CB200_U8)
      mask_kernel<T, uint8_t><<<blocks, 256, 0, st>>>(std_channel, threshold, n_pix, (uint8_t*)mask_out);
    else if (mask_dtype == CB200_U16)
      mask_kernel<T, uint16_t><<<blocks, 256, 0, st>>>(std_channel, threshold, n_pix, (uint16_t*)mask_out);
    else
      return CB200_EUNSUPPORTED;
    CB200_LAUNCH_CHECK();
  }
  FgPred<T> pred{std_channel, threshold};
  if (num_dims == 2) {
    FgEmit<T, 2> emit{emb, n_pix, {(int)spatial[1], (int)spatial[0], 1}, points, capacity, pix_index};
    return run_compaction(pred, emit, n_pix, capacity, n_out, workspace, st);
  }
  FgEmit<T, 3> emit{emb, n_pix, {(int)spatial[2], (int)spatial[1], (int)spatial[0]}, points, capacity, pix_index};
  return run_compaction(pred, emit, n_pix, capacity, n_out, workspace, st);
}

// min / max of the first min(n_max, *n_dev) elements (n_dev == NULL: of n_max elements); the launch is sized for n_max
int minmax_counted(const void* x, int dtype, int64_t n_max, const long long* n_dev, double* out2, void* workspace,
                   cudaStream_t st) {
  if (!x || !out2 || !workspace || n_max <= 0) return CB200_EINVAL;
  auto* ws = static_cast<ReduceWorkspace*>(workspace);
  const int blocks = grid_for(n_max, RED_THREADS, 8, 8);
  if (dtype == CB200_F32)
    minmax_kernel<float><<<blocks, RED_THREADS, 0, st>>>((const float*)x, n_max, n_dev, out2, ws);
  else if (dtype == CB200_F64)
    minmax_kernel<double><<<blocks, RED_THREADS, 0, st>>>((const double*)x, n_max, n_dev, out2, ws);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

// cb200_select_points over the first min(n_max, *n_dev) source points
int select_points_counted(const double* src, int64_t n_max, const long long* n_dev, int64_t src_stride, int num_dims,
                          const uint8_t* flags, double* dst, int64_t dst_stride, long long* n_out, void* workspace,
                          cudaStream_t st) {
  if (!src || !flags || !dst || !n_out || !workspace || n_max < 0 || num_dims < 1 || num_dims > 3) return CB200_EINVAL;
  FlagPred pred{flags, n_dev};
  SelectEmit emit{src, src_stride, dst, dst_stride, num_dims};
  return run_compaction(pred, emit, n_max, dst_stride, n_out, workspace, st);
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_reduce_workspace_bytes(void) { return (int64_t)sizeof(ReduceWorkspace); }

int cb200_minmax(const void* x, int dtype, int64_t n, double* out2, void* workspace, void* stream) {
  return minmax_counted(x, dtype, n, nullptr, out2, workspace, (cudaStream_t)stream);
}

int cb200_histogram(const void* x, int dtype, int64_t n, const double* edges, int nbins, unsigned long long* counts,
                    void* stream) {
  if (!x || !edges || !counts || n < 0 || nbins <= 0 || nbins > HIST_MAX_BINS) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(double) * (nbins + 1) + sizeof(unsigned int) * (HIST_THREADS / 32) * nbins;
  const int blocks = grid_for(n, HIST_THREADS, 16, 4);
  if (dtype == CB200_F32)
    histogram_kernel<float><<<blocks, HIST_THREADS, smem, st>>>((const float*)x, n, edges, nbins, counts);
  else if (dtype == CB200_F64)
    histogram_kernel<double><<<blocks, HIST_THREADS, smem, st>>>((const double*)x, n, edges, nbins, counts);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_centre_workspace_bytes(void) { return (int64_t)sizeof(CentreWorkspace); }

int cb200_centre_embeddings(const void* emb, int dtype, int num_dims, int64_t n_pix, double threshold, double* means,
                            void* out, void* workspace, void* stream) {
  if (!emb || !means || !workspace || n_pix <= 0) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  auto* ws = static_cast<CentreWorkspace*>(workspace);
  const int blocks = grid_for(n_pix, RED_THREADS, 8, 8);
  const int cblocks = grid_for((int64_t)(num_dims + 1) * n_pix, 256, 4, 16);
#define CB200_CENTRE(T, DD)                                                                                \
  masked_channel_mean_kernel<T, DD><<<blocks, RED_THREADS, 0, st>>>((const T*)emb, n_pix, threshold, means, ws); \
  if (out) centre_kernel<T><<<cblocks, 256, 0, st>>>((const T*)emb, DD, n_pix, means, (T*)out);
  if (dtype == CB200_F32 && num_dims == 2) { CB200_CENTRE(float, 2) }
  else if (dtype == CB200_F32) { CB200_CENTRE(float, 3) }
  else if (dtype == CB200_F64 && num_dims == 2) { CB200_CENTRE(double, 2) }
  else if (dtype == CB200_F64) { CB200_CENTRE(double, 3) }
  else return CB200_EUNSUPPORTED;
#undef CB200_CENTRE
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_compact_workspace_bytes(int64_t n_pix) { return CompactWorkspace::bytes(n_pix); }

int cb200_fg_compact(const void* emb, int dtype, int num_dims, const int64_t* spatial, double threshold, double* points,
                     int32_t* pix_index, int64_t capacity, long long* n_out, void* mask_out, int mask_dtype,
                     void* workspace, void* stream) {
  if (!emb || !spatial || !points || !n_out || !workspace || capacity < 0) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CB200_F32)
    return fg_compact_typed<float>(emb, num_dims, spatial, threshold, points, pix_index, capacity, n_out, mask_out,
                                   mask_dtype, workspace, st);
  if (dtype == CB200_F64)
    return fg_compact_typed<double>(emb, num_dims, spatial, threshold, points, pix_index, capacity, n_out, mask_out,
                                    mask_dtype, workspace, st);
  return CB200_EUNSUPPORTED;
}

int cb200_select_points(const double* src, int64_t n, int64_t src_stride, int num_dims, const uint8_t* flags,
                        double* dst, int64_t dst_stride, long long* n_out, void* workspace, void* stream) {
  return select_points_counted(src, n, nullptr, src_stride, num_dims, flags, dst, dst_stride, n_out, workspace,
                               (cudaStream_t)stream);
}

int cb200_bernoulli_flags(uint8_t* flags, int64_t n, double p, uint64_t seed, void* stream) {
  if (!flags || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  bernoulli_kernel<<<grid_for((n + 1) / 2, 256, 2, 16), 256, 0, (cudaStream_t)stream>>>(flags, n, p, seed);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
