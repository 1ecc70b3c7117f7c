// Connected-component labelling + size filter for sm_100a (utils/misc.py:11-25;
// skimage.measure.label semantics: full connectivity, regions of EQUAL value,
// 0 = background, ids 1.. in raster order of each region's first pixel).
//
// Lock-free union-find over the pixel lattice: every link points a root at a
// SMALLER linear index, so the final root of a component is its first pixel in
// raster order -- exactly the key skimage numbers components by.  Roots are
// then ranked with the order-preserving compaction scan.
#include "common.cuh"
#include "compact.cuh"
#include "unionfind.cuh"

namespace cb200 {

__global__ void __launch_bounds__(256) cc_init_kernel(const int32_t* __restrict__ seg, int64_t n, int* __restrict__ parent) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) parent[i] = seg[i] != 0 ? (int)i : -1;
}

// Each pixel links to the "earlier" half of its 3^D - 1 neighbours -- but only where the link is not already
// implied by a link another pixel makes: inside a region nearly every pixel has an equal left neighbour whose
// own upward link covers it, so unions (atomics) happen along run boundaries only.
//   in-plane:  left always; up unless (left and up-left are equal too: the left pixel links upward);
//              up-left / up-right only when up differs (otherwise they hang on `up` through their own row)
//   3-D:       the voxel in front (z-1) like `up`; the other eight of the z-1 plane only when it differs
template <int D>
__global__ void __launch_bounds__(256)
cc_merge_kernel(const int32_t* __restrict__ seg, int64_t n, int ex, int ey, int ez, int* parent) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  const int64_t plane = (int64_t)ex * ey;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int32_t v = seg[i];
    if (v == 0) continue;
    const int x = (int)(i % ex);
    const int64_t r = i / ex;
    const int y = (int)(D == 2 ? r : r % ey);
    const int z = (int)(D == 2 ? 0 : r / ey);
    const bool has_l = x > 0, has_r = x < ex - 1, has_u = y > 0;
    const bool L = has_l && seg[i - 1] == v;
    const bool U = has_u && seg[i - ex] == v;
    if (L) uf_union(parent, (int)i, (int)(i - 1));
    if (U) {
      if (!(L && seg[i - ex - 1] == v)) uf_union(parent, (int)i, (int)(i - ex));
    } else if (has_u) {
      if (has_l && !L && seg[i - ex - 1] == v) uf_union(parent, (int)i, (int)(i - ex - 1));
      if (has_r && seg[i - ex + 1] == v) uf_union(parent, (int)i, (int)(i - ex + 1));
    }
    if constexpr (D == 3) {
      if (z == 0) continue;
      const int64_t f = i - plane;
      if (seg[f] == v) {
        if (!(L && seg[f - 1] == v)) uf_union(parent, (int)i, (int)f);
      } else {
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (dx == 0 && dy == 0) continue;
            const int xx = x + dx, yy = y + dy;
            if (xx < 0 || xx >= ex || yy < 0 || yy >= ey) continue;
            const int64_t j = f + (int64_t)dy * ex + dx;
            if (seg[j] == v) uf_union(parent, (int)i, (int)j);
          }
      }
    }
  }
}

__global__ void __launch_bounds__(256) cc_flatten_kernel(int* parent, int64_t n, int* __restrict__ sizes) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    if (parent[i] < 0) continue;
    const int r = uf_find(parent, (int)i);
    parent[i] = r;
    if (sizes) atomicAdd(sizes + r, 1);
  }
}

struct RootPred {
  const int* parent;
  __device__ __forceinline__ bool operator()(int64_t i) const { return parent[i] == (int)i; }
};
struct RootEmit {
  int* root_label;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const { root_label[i] = (int)d + 1; }
};

__global__ void __launch_bounds__(256)
cc_write_labels_kernel(const int* __restrict__ parent, const int* __restrict__ root_label, int64_t n,
                       int32_t* __restrict__ labels, const long long* __restrict__ n_roots, int* __restrict__ n_labels) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int p = parent[i];
    labels[i] = p < 0 ? 0 : root_label[p];
  }
  if (n_labels && blockIdx.x == 0 && threadIdx.x == 0) *n_labels = (int)*n_roots;
}

__global__ void __launch_bounds__(256)
cc_filter_kernel(int32_t* __restrict__ seg, const int* __restrict__ parent, const int* __restrict__ sizes, int64_t n,
                 int64_t min_size) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int p = parent[i];
    if (p >= 0 && (int64_t)sizes[p] < min_size) seg[i] = 0;
  }
}

struct CcBuffers {
  int* parent;
  int* aux;  // sizes, then root labels
  long long* n_roots;
  void* compact_ws;
};
static size_t cc_align(size_t v) { return (v + 255) / 256 * 256; }
static CcBuffers cc_carve(void* workspace, int64_t n) {
  char* w = reinterpret_cast<char*>(cc_align(reinterpret_cast<size_t>(workspace)));
  CcBuffers b;
  b.parent = (int*)w;      w += cc_align(4 * (size_t)n);
  b.aux = (int*)w;         w += cc_align(4 * (size_t)n);
  b.n_roots = (long long*)w; w += 256;
  b.compact_ws = w;
  return b;
}

static int cc_label(const int32_t* seg, int D, const int64_t* spatial, int64_t n, const CcBuffers& b, bool want_sizes,
                    cudaStream_t st) {
  const int ex = (int)spatial[D - 1], ey = (int)spatial[D - 2], ez = D == 3 ? (int)spatial[0] : 1;
  const int blocks = grid_for(n, 256, 2, 16);
  cc_init_kernel<<<blocks, 256, 0, st>>>(seg, n, b.parent);
  CB200_LAUNCH_CHECK();
  if (D == 2) cc_merge_kernel<2><<<blocks, 256, 0, st>>>(seg, n, ex, ey, ez, b.parent);
  else cc_merge_kernel<3><<<blocks, 256, 0, st>>>(seg, n, ex, ey, ez, b.parent);
  CB200_LAUNCH_CHECK();
  if (want_sizes) CB200_CUDA_TRY(cudaMemsetAsync(b.aux, 0, 4 * (size_t)n, st));
  cc_flatten_kernel<<<blocks, 256, 0, st>>>(b.parent, n, want_sizes ? b.aux : nullptr);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

static int cc_number(int64_t n, const CcBuffers& b, int32_t* labels_out, int* n_labels, cudaStream_t st) {
  RootPred pred{b.parent};
  RootEmit emit{b.aux};
  const int rc = run_compaction(pred, emit, n, n, b.n_roots, b.compact_ws, st);
  if (rc != CB200_OK) return rc;
  cc_write_labels_kernel<<<grid_for(n, 256, 2, 16), 256, 0, st>>>(b.parent, b.aux, n, labels_out, b.n_roots, n_labels);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

static bool cc_shape(int D, const int64_t* spatial, int64_t& n) {
  if ((D != 2 && D != 3) || !spatial) return false;
  n = 1;
  for (int k = 0; k < D; ++k) {
    if (spatial[k] <= 0) return false;
    n *= spatial[k];
  }
  return n <= INT32_MAX;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_cc_workspace_bytes(int64_t n_pix) {
  return (int64_t)(2 * cc_align(4 * (size_t)n_pix) + 256 + (size_t)CompactWorkspace::bytes(n_pix) + 512);
}

int cb200_label_components(const int32_t* seg, int num_dims, const int64_t* spatial, int32_t* labels_out, int* n_labels,
                           void* workspace, void* stream) {
  int64_t n;
  if (!seg || !labels_out || !workspace || !cc_shape(num_dims, spatial, n)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const CcBuffers b = cc_carve(workspace, n);
  int rc = cc_label(seg, num_dims, spatial, n, b, false, st);
  if (rc != CB200_OK) return rc;
  return cc_number(n, b, labels_out, n_labels, st);
}

int cb200_size_filter(int32_t* seg, int num_dims, const int64_t* spatial, int64_t min_size, int32_t* labels_out,
                      int* n_labels, void* workspace, void* stream) {
  int64_t n;
  if (!seg || !labels_out || !workspace || !cc_shape(num_dims, spatial, n)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const CcBuffers b = cc_carve(workspace, n);
  int rc = cc_label(seg, num_dims, spatial, n, b, true, st);  // filter_labels = measure.label(segmentation)
  if (rc != CB200_OK) return rc;
  cc_filter_kernel<<<grid_for(n, 256, 2, 16), 256, 0, st>>>(seg, b.parent, b.aux, n, min_size);  // segmentation[mask] = 0
  CB200_LAUNCH_CHECK();
  rc = cc_label(seg, num_dims, spatial, n, b, false, st);  // return measure.label(segmentation)
  if (rc != CB200_OK) return rc;
  return cc_number(n, b, labels_out, n_labels, st);
}

}  // extern "C"
