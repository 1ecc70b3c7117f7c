// Centre post-processing and label assignment for sm_100a
// (sklearn _mean_shift.py:511-547 and :563-579; utils/mean_shift.py:101-104,57).
//
// sklearn's duplicate removal is a SEQUENTIAL greedy pass over the centres in
// (count, coords)-descending order.  Its result is the lexicographically-first
// maximal independent set of the "within bandwidth" graph in that order, which
// has an exact parallel formulation: a centre is KEPT once every earlier
// neighbour is REMOVED, and REMOVED once any earlier neighbour is KEPT.  The
// rounds below reach the same fix-point the sequential loop does.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "compact.cuh"

namespace cb200 {

// order-preserving map double -> uint64 (after folding -0.0 onto +0.0, as == does)
__device__ __forceinline__ unsigned long long sortable(double v) {
  v = v + 0.0;
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) iota_kernel(int* __restrict__ p, int64_t n) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) p[i] = (int)i;
}
__global__ void __launch_bounds__(256)
key_from_coord_kernel(const double* __restrict__ col, const int* __restrict__ perm, int64_t n,
                      unsigned long long* __restrict__ keys) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) keys[i] = sortable(col[perm[i]]);
}
__global__ void __launch_bounds__(256)
key_from_count_kernel(const int* __restrict__ counts, const int* __restrict__ perm, int64_t n,
                      unsigned long long* __restrict__ keys) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs)
    keys[i] = (unsigned long long)(unsigned)max(counts[perm[i]], 0);
}

// gather centres into priority order; an entry is a candidate iff count > 0 and it is not an
// exact duplicate of its predecessor (sklearn's dict keyed by the coordinate tuple)
template <int D>
__global__ void __launch_bounds__(256)
gather_sorted_kernel(const double* __restrict__ modes, int64_t seed_stride, const int* __restrict__ counts,
                     const int* __restrict__ perm, int64_t n, double* __restrict__ sorted, uint8_t* __restrict__ cand) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int s = perm[i];
    bool dup = i > 0;
    const int sp = i > 0 ? perm[i - 1] : s;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double v = modes[k * seed_stride + s];
      sorted[k * n + i] = v;
      dup = dup && (v == modes[k * seed_stride + sp]);
    }
    dup = dup && counts[sp] > 0;
    cand[i] = (counts[s] > 0 && !dup) ? 1 : 0;
  }
}

struct NmsGrid {
  double origin[3];
  double inv_cell;
  int dims[3];
};

template <int D>
__device__ __forceinline__ void cell_coords(const double (&x)[D], const NmsGrid& g, int (&c)[3]) {
  c[0] = c[1] = c[2] = 0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double f = floor((x[k] - g.origin[k]) * g.inv_cell);
    f = fmin(fmax(f, 0.0), (double)(g.dims[k] - 1));
    c[k] = (int)f;
  }
}

template <int D>
__global__ void __launch_bounds__(256)
nms_cell_ids_kernel(const double* __restrict__ sorted, int64_t n, NmsGrid g, unsigned* __restrict__ keys,
                    int* __restrict__ idx) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = sorted[k * n + i];
    int c[3];
    cell_coords<D>(x, g, c);
    keys[i] = (unsigned)((c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0]);
    idx[i] = (int)i;
  }
}

__global__ void __launch_bounds__(256)
nms_cell_start_kernel(const unsigned* __restrict__ sorted_keys, int64_t n, int64_t n_cells, int* __restrict__ cell_start) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= n_cells; c += gs) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)sorted_keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = (int)lo;
  }
}

enum : uint8_t { NMS_UNDECIDED = 0, NMS_KEEP = 1, NMS_REMOVED = 2 };

__global__ void __launch_bounds__(256)
nms_init_kernel(const uint8_t* __restrict__ cand, int64_t n, uint8_t* __restrict__ state, int* __restrict__ rounds) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs)
    state[i] = cand[i] ? NMS_UNDECIDED : NMS_REMOVED;
  if (blockIdx.x == 0 && threadIdx.x == 0) rounds[0] = 1;
}

// state is indexed by PRIORITY index i; cell_order lists priority indices sorted by cell (stable:
// ascending priority index inside a cell).
template <int D>
__global__ void __launch_bounds__(256)
nms_round_kernel(const double* __restrict__ sorted, int64_t n, NmsGrid g, const int* __restrict__ cell_start,
                 const int* __restrict__ cell_order, double r2, volatile uint8_t* state, int* __restrict__ rounds,
                 int round) {
  if (rounds[round] == 0) {  // nothing was undecided after the previous round
    if (blockIdx.x == 0 && threadIdx.x == 0) rounds[round + 1] = 0;
    return;
  }
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  int undecided = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    if (state[i] != NMS_UNDECIDED) continue;
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = sorted[k * n + i];
    int c[3];
    cell_coords<D>(x, g, c);
    bool removed = false, blocked = false;
    const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.dims[0] - 1);
    const int y0 = max(c[1] - 1, 0), y1 = min(c[1] + 1, g.dims[1] - 1);
    const int z0 = D == 3 ? max(c[2] - 1, 0) : 0, z1 = D == 3 ? min(c[2] + 1, g.dims[2] - 1) : 0;
    for (int z = z0; z <= z1 && !removed; ++z)
      for (int y = y0; y <= y1 && !removed; ++y) {
        const int64_t row = ((int64_t)z * g.dims[1] + y) * g.dims[0];
        const int beg = cell_start[row + x0], end = cell_start[row + x1 + 1];
        for (int q = beg; q < end; ++q) {
          const int j = cell_order[q];
          if (j >= i) continue;  // only earlier (higher-priority) centres matter
          const uint8_t sj = state[j];
          if (sj == NMS_REMOVED) continue;
          double d = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const double t = __dsub_rn(x[k], sorted[k * n + j]);
            d = __dadd_rn(d, __dmul_rn(t, t));
          }
          if (d <= r2) {
            if (sj == NMS_KEEP) { removed = true; break; }
            blocked = true;
          }
        }
      }
    if (removed) state[i] = NMS_REMOVED;
    else if (!blocked) state[i] = NMS_KEEP;
    else ++undecided;
  }
  undecided = warp_sum(undecided);
  if (lane_id() == 0 && undecided) atomicAdd(&rounds[round + 1], undecided);
}

struct KeepPred {
  const uint8_t* state;
  __device__ __forceinline__ bool operator()(int64_t i) const { return state[i] == NMS_KEEP; }
};
struct CentreEmit {
  const double* sorted;
  int64_t n;
  double* out;
  int D;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const {
    for (int k = 0; k < D; ++k) out[k * n + d] = sorted[k * n + i];
  }
};
__global__ void nms_finish_kernel(const long long* __restrict__ n_keep, const int* __restrict__ rounds, int last_round,
                                  int* __restrict__ out2) {
  out2[0] = (int)*n_keep;
  out2[1] = rounds[last_round];
}

// ------------------------------------------------------------------ label assignment
constexpr int ASG_THREADS = 256;
constexpr int ASG_TILE = 1024;

template <int D, typename L>
__global__ void __launch_bounds__(ASG_THREADS)
assign_labels_kernel(const double* __restrict__ points, int64_t n, int64_t pts_stride,
                     const double* __restrict__ centres, int64_t centre_stride, int K,
                     const int32_t* __restrict__ pix_index, L* __restrict__ labels) {
  __shared__ double s_c[D][ASG_TILE];
  const int64_t i = (int64_t)blockIdx.x * ASG_THREADS + threadIdx.x;
  double x[D];
#pragma unroll
  for (int k = 0; k < D; ++k) x[k] = i < n ? __ldg(points + k * pts_stride + i) : 0.0;
  double best = INFINITY;
  int best_k = 0;
  for (int k0 = 0; k0 < K; k0 += ASG_TILE) {
    const int cnt = min(ASG_TILE, K - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += ASG_THREADS) {
#pragma unroll
      for (int k = 0; k < D; ++k) s_c[k][j] = centres[k * centre_stride + k0 + j];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      double d = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double t = x[k] - s_c[k][j];
        d = fma(t, t, d);
      }
      if (d < best) {  // strict: ties keep the lowest index (pairwise_distances_argmin)
        best = d;
        best_k = k0 + j;
      }
    }
  }
  if (i < n) {
    const int64_t dst = pix_index ? (int64_t)pix_index[i] : i;
    labels[dst] = (L)(best_k + 1);  // +1: 0 is background (utils/mean_shift.py:57)
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct NmsLayout {
  size_t sort64_bytes, sort32_bytes;
  size_t off_sort, off_keys_a, off_keys_b, off_perm_a, off_perm_b, off_sorted, off_cand, off_state, off_ckeys_a,
      off_ckeys_b, off_cidx_a, off_cidx_b, off_cell_start, off_rounds, off_nkeep, off_compact, total;
};
constexpr int NMS_ROUNDS = 48;

static NmsLayout nms_layout(int64_t n, int D, int64_t n_cells) {
  NmsLayout L{};
  cub::DeviceRadixSort::SortPairsDescending(nullptr, L.sort64_bytes, (const unsigned long long*)nullptr,
                                            (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, L.sort32_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
  L.off_sort = take(L.sort64_bytes > L.sort32_bytes ? L.sort64_bytes : L.sort32_bytes);
  L.off_keys_a = take(8 * (size_t)n);
  L.off_keys_b = take(8 * (size_t)n);
  L.off_perm_a = take(4 * (size_t)n);
  L.off_perm_b = take(4 * (size_t)n);
  L.off_sorted = take(8 * (size_t)n * D);
  L.off_cand = take((size_t)n);
  L.off_state = take((size_t)n);
  L.off_ckeys_a = take(4 * (size_t)n);
  L.off_ckeys_b = take(4 * (size_t)n);
  L.off_cidx_a = take(4 * (size_t)n);
  L.off_cidx_b = take(4 * (size_t)n);
  L.off_cell_start = take(4 * (size_t)(n_cells + 1));
  L.off_rounds = take(4 * (NMS_ROUNDS + 2));
  L.off_nkeep = take(8);
  L.off_compact = take((size_t)CompactWorkspace::bytes(n));
  L.total = o + 256;
  return L;
}

template <int D>
static int nms_run(const double* modes, int64_t seed_stride, const int* counts, int64_t n, double bandwidth,
                   const cb200_grid* grid, double* centres_out, int* out2, void* workspace, int64_t workspace_bytes,
                   cudaStream_t st) {
  const NmsLayout L = nms_layout(n, D, grid->n_cells);
  if ((int64_t)L.total > workspace_bytes) return CB200_EINVAL;
  char* w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  void* sort_ws = w + L.off_sort;
  auto* keys_a = (unsigned long long*)(w + L.off_keys_a);
  auto* keys_b = (unsigned long long*)(w + L.off_keys_b);
  int* perm_a = (int*)(w + L.off_perm_a);
  int* perm_b = (int*)(w + L.off_perm_b);
  double* sorted = (double*)(w + L.off_sorted);
  uint8_t* cand = (uint8_t*)(w + L.off_cand);
  uint8_t* state = (uint8_t*)(w + L.off_state);
  unsigned* ckeys_a = (unsigned*)(w + L.off_ckeys_a);
  unsigned* ckeys_b = (unsigned*)(w + L.off_ckeys_b);
  int* cidx_a = (int*)(w + L.off_cidx_a);
  int* cidx_b = (int*)(w + L.off_cidx_b);
  int* cell_start = (int*)(w + L.off_cell_start);
  int* rounds = (int*)(w + L.off_rounds);
  long long* n_keep = (long long*)(w + L.off_nkeep);
  void* compact_ws = w + L.off_compact;

  const int blocks = grid_for(n, 256, 2, 16);
  iota_kernel<<<blocks, 256, 0, st>>>(perm_a, n);
  CB200_LAUNCH_CHECK();
  // LSD: least-significant key first -> last coordinate ... first coordinate, then the count
  size_t sb = L.sort64_bytes;
  for (int k = D - 1; k >= 0; --k) {
    key_from_coord_kernel<<<blocks, 256, 0, st>>>(modes + k * seed_stride, perm_a, n, keys_a);
    CB200_LAUNCH_CHECK();
    CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(sort_ws, sb, keys_a, keys_b, perm_a, perm_b, (int)n, 0, 64, st));
    int* t = perm_a; perm_a = perm_b; perm_b = t;
  }
  key_from_count_kernel<<<blocks, 256, 0, st>>>(counts, perm_a, n, keys_a);
  CB200_LAUNCH_CHECK();
  CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(sort_ws, sb, keys_a, keys_b, perm_a, perm_b, (int)n, 0, 32, st));
  { int* t = perm_a; perm_a = perm_b; perm_b = t; }
  gather_sorted_kernel<D><<<blocks, 256, 0, st>>>(modes, seed_stride, counts, perm_a, n, sorted, cand);
  CB200_LAUNCH_CHECK();

  NmsGrid g;
  for (int k = 0; k < 3; ++k) { g.origin[k] = grid->origin[k]; g.dims[k] = grid->dims[k]; }
  g.inv_cell = grid->inv_cell;
  nms_cell_ids_kernel<D><<<blocks, 256, 0, st>>>(sorted, n, g, ckeys_a, cidx_a);
  CB200_LAUNCH_CHECK();
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < grid->n_cells) ++bits;
  size_t sb32 = L.sort32_bytes;
  CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_ws, sb32, ckeys_a, ckeys_b, cidx_a, cidx_b, (int)n, 0, bits, st));
  nms_cell_start_kernel<<<grid_for(grid->n_cells + 1, 256, 1, 16), 256, 0, st>>>(ckeys_b, n, grid->n_cells, cell_start);
  CB200_LAUNCH_CHECK();

  CB200_CUDA_TRY(cudaMemsetAsync(rounds, 0, 4 * (NMS_ROUNDS + 2), st));
  nms_init_kernel<<<blocks, 256, 0, st>>>(cand, n, state, rounds);
  CB200_LAUNCH_CHECK();
  const double r2 = bandwidth * bandwidth;
  for (int r = 0; r < NMS_ROUNDS; ++r) {
    nms_round_kernel<D><<<grid_for(n, 256, 1, 16), 256, 0, st>>>(sorted, n, g, cell_start, cidx_b, r2, state, rounds, r);
    CB200_LAUNCH_CHECK();
  }
  KeepPred pred{state};
  CentreEmit emit{sorted, n, centres_out, D};
  const int rc = run_compaction(pred, emit, n, n, n_keep, compact_ws, st);
  if (rc != CB200_OK) return rc;
  nms_finish_kernel<<<1, 1, 0, st>>>(n_keep, rounds, NMS_ROUNDS, out2);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_nms_workspace_bytes(int64_t n_seeds, int num_dims, int64_t n_cells) {
  if (n_seeds <= 0) return 512;
  return (int64_t)nms_layout(n_seeds, num_dims, n_cells).total;
}

int cb200_nms_centres(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                      double bandwidth, const cb200_grid* grid, double* centres_out, int* n_centres_and_undecided,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  if (!modes || !counts || !grid || !centres_out || !n_centres_and_undecided || !workspace || n_seeds <= 0)
    return CB200_EINVAL;
  if (n_seeds > INT32_MAX || grid->n_cells >= ((int64_t)1 << 31) || !(grid->cell >= bandwidth)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_dims == 2)
    return nms_run<2>(modes, seed_stride, counts, n_seeds, bandwidth, grid, centres_out, n_centres_and_undecided,
                      workspace, workspace_bytes, st);
  if (num_dims == 3)
    return nms_run<3>(modes, seed_stride, counts, n_seeds, bandwidth, grid, centres_out, n_centres_and_undecided,
                      workspace, workspace_bytes, st);
  return CB200_EUNSUPPORTED;
}

int cb200_assign_labels(const double* points, int64_t n_points, int64_t pts_stride, int num_dims,
                        const double* centres, int64_t centre_stride, int n_centres, const int32_t* pix_index,
                        void* labels_out, int label_dtype, void* stream) {
  if (!points || !centres || !labels_out || n_points < 0 || n_centres <= 0) return CB200_EINVAL;
  if (n_points == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)((n_points + ASG_THREADS - 1) / ASG_THREADS);
#define CB200_ASSIGN(DD, LT)                                                                                        \
  assign_labels_kernel<DD, LT><<<blocks, ASG_THREADS, 0, st>>>(points, n_points, pts_stride, centres, centre_stride, \
                                                               n_centres, pix_index, (LT*)labels_out)
  if (num_dims == 2 && label_dtype == CB200_I32) CB200_ASSIGN(2, int32_t);
  else if (num_dims == 2 && label_dtype == CB200_U16) CB200_ASSIGN(2, uint16_t);
  else if (num_dims == 3 && label_dtype == CB200_I32) CB200_ASSIGN(3, int32_t);
  else if (num_dims == 3 && label_dtype == CB200_U16) CB200_ASSIGN(3, uint16_t);
  else return CB200_EUNSUPPORTED;
#undef CB200_ASSIGN
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
