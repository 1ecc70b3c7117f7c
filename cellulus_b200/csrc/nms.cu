// Centre post-processing and label assignment for sm_100a
// (sklearn _mean_shift.py:511-547 and :563-579; utils/mean_shift.py:101-104,57).
//
// sklearn's duplicate removal is a SEQUENTIAL greedy pass over the centres in
// (count, coords)-descending order.  Its result is the lexicographically-first
// maximal independent set of the "within bandwidth" graph in that order, which
// has an exact parallel formulation: a centre is KEPT once every neighbour of
// higher priority is REMOVED, and REMOVED once any neighbour of higher priority
// is KEPT.  Priority is evaluated pairwise with sklearn's sort key -- no global
// sort of the (often > 10^5) converged modes is needed; only the K survivors are
// sorted to reproduce the order of `cluster_centers_` (label numbering).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "compact.cuh"

namespace cb200 {

struct NmsGrid {
  double origin[3];
  double inv_cell;
  int dims[3];
  int n_cells;
};

static NmsGrid to_nms_grid(const cb200_grid& g) {
  NmsGrid d;
  for (int k = 0; k < 3; ++k) {
    d.origin[k] = g.origin[k];
    d.dims[k] = g.dims[k];
  }
  d.inv_cell = g.inv_cell;
  d.n_cells = (int)g.n_cells;
  return d;
}

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

// sklearn's sort key, descending: (count, (x, y[, z])); exact ties (duplicates) fall to the lower index
template <int D>
__device__ __forceinline__ bool has_priority(int cnt_j, const double (&xj)[D], int j, int cnt_i, const double (&xi)[D],
                                             int i) {
  if (cnt_j != cnt_i) return cnt_j > cnt_i;
#pragma unroll
  for (int k = 0; k < D; ++k)
    if (xj[k] != xi[k]) return xj[k] > xi[k];
  return j < i;
}

// ---- fine grid for the suppression: edge = bw/2 (1+1e-6) ------------------------------------------
// Two modes in the same fine cell are always within the bandwidth of each other (cell diagonal
// 0.87 bw in 3-D, 0.71 bw in 2-D), so AT MOST ONE mode per cell can survive and only the
// highest-priority live mode of a cell -- its "top" -- can be the next survivor.  That turns the
// O(modes x neighbours) fix-point into O(non-empty cells x neighbours) per round.
static bool fine_grid_from(const cb200_grid& g, double bandwidth, NmsGrid& f) {
  const double edge = 0.5 * bandwidth * (1.0 + 1e-6);
  int64_t n = 1;
  for (int k = 0; k < 3; ++k) {
    f.origin[k] = g.origin[k];
    int64_t d = 1;
    if (k < g.num_dims) d = (int64_t)floor(((double)g.dims[k] * g.cell) / edge) + 1;
    if (d > INT32_MAX) return false;
    f.dims[k] = (int)d;
    n *= d;
    if (n >= ((int64_t)1 << 31) - 2) return false;
  }
  f.inv_cell = 1.0 / edge;
  f.n_cells = (int)n;
  return true;
}

// cell id per mode; modes with an empty window (count == 0, sklearn:511-513) go to a sentinel cell past the grid
template <int D>
__global__ void __launch_bounds__(256)
nms_cell_ids_kernel(const double* __restrict__ modes, int64_t stride, const int* __restrict__ counts, int64_t n,
                    NmsGrid g, unsigned* __restrict__ keys, int* __restrict__ idx) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    unsigned key = (unsigned)g.n_cells;
    if (counts[i] > 0) {
      double x[D];
#pragma unroll
      for (int k = 0; k < D; ++k) x[k] = modes[k * stride + i];
      int c[3];
      cell_coords<D>(x, g, c);
      key = (unsigned)((c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0]);
    }
    keys[i] = key;
    idx[i] = (int)i;
  }
}

__global__ void __launch_bounds__(256)
cell_start_kernel(const unsigned* __restrict__ sorted_keys, int64_t n, int64_t n_cells, int* __restrict__ cell_start) {
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
nms_init_kernel(const int* __restrict__ counts, int64_t n, uint8_t* __restrict__ state, int* __restrict__ undecided) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs)
    state[i] = counts[i] > 0 ? NMS_UNDECIDED : NMS_REMOVED;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    undecided[0] = 1;  // "previous round left work"
    undecided[1] = 0;
  }
}

// first sorted position of every non-empty real cell (the compaction predicate / emit)
struct CellHeadPred {
  const unsigned* keys;
  unsigned n_cells;
  __device__ __forceinline__ bool operator()(int64_t q) const {
    const unsigned k = keys[q];
    return k < n_cells && (q == 0 || keys[q - 1] != k);
  }
};
struct CellHeadEmit {
  const unsigned* keys;
  int* cells;
  __device__ __forceinline__ void operator()(int64_t q, long long d) const { cells[d] = (int)keys[q]; }
};

// One WARP per non-empty fine cell per round:
//   1. top = highest-priority live mode of the cell (none -> the cell is finished);
//   2. scan the modes of the 5^D neighbour cells that lie within the bandwidth of top:
//      a KEPT one removes top, a live one of higher priority makes top wait;
//   3. otherwise top is KEPT and every live mode within its bandwidth is REMOVED.
template <int D>
__global__ void __launch_bounds__(256)
nms_round_kernel(const double* __restrict__ modes, int64_t stride, const int* __restrict__ counts, NmsGrid g,
                 const int* __restrict__ cell_start, const int* __restrict__ cell_order,
                 const int* __restrict__ cells, const long long* __restrict__ n_cells_live, uint8_t* cell_done,
                 double r2, volatile uint8_t* state, const int* __restrict__ prev_undecided,
                 int* __restrict__ undecided) {
  if (*prev_undecided == 0) return;  // fix-point already reached
  const int lane = lane_id();
  const int64_t warp_gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_live = *n_cells_live;
  int left = 0;
  for (int64_t w = warp_gid; w < n_live; w += n_warps) {
    if (cell_done[w]) continue;
    const int cell = cells[w];
    // ---- 1. the cell's top live mode (lane-local best, then a shuffle tournament)
    int bi = -1, bc = 0;
    double bx[D];
#pragma unroll
    for (int k = 0; k < D; ++k) bx[k] = 0.0;
    for (int q = cell_start[cell] + lane; q < cell_start[cell + 1]; q += 32) {
      const int j = cell_order[q];
      if (state[j] != NMS_UNDECIDED) continue;
      double xj[D];
#pragma unroll
      for (int k = 0; k < D; ++k) xj[k] = modes[k * stride + j];
      const int cj = counts[j];
      if (bi < 0 || has_priority<D>(cj, xj, j, bc, bx, bi)) {
        bi = j;
        bc = cj;
#pragma unroll
        for (int k = 0; k < D; ++k) bx[k] = xj[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int oi = __shfl_xor_sync(FULL, bi, o);
      const int oc = __shfl_xor_sync(FULL, bc, o);
      double ox[D];
#pragma unroll
      for (int k = 0; k < D; ++k) ox[k] = __shfl_xor_sync(FULL, bx[k], o);
      if (oi >= 0 && (bi < 0 || has_priority<D>(oc, ox, oi, bc, bx, bi))) {
        bi = oi;
        bc = oc;
#pragma unroll
        for (int k = 0; k < D; ++k) bx[k] = ox[k];
      }
    }
    if (bi < 0) {  // nothing live: finished for good
      if (lane == 0) cell_done[w] = 1;
      continue;
    }
    // ---- 2. is top dominated?
    const int cz = cell / (g.dims[0] * g.dims[1]);
    const int cy = (cell / g.dims[0]) % g.dims[1];
    const int cx = cell % g.dims[0];
    const int x0 = max(cx - 2, 0), x1 = min(cx + 2, g.dims[0] - 1);
    const int y0 = max(cy - 2, 0), y1 = min(cy + 2, g.dims[1] - 1);
    const int z0 = D == 3 ? max(cz - 2, 0) : 0, z1 = D == 3 ? min(cz + 2, g.dims[2] - 1) : 0;
    bool removed = false, blocked = false;
    for (int z = z0; z <= z1 && !removed; ++z)
      for (int y = y0; y <= y1 && !removed; ++y) {
        const int64_t row = ((int64_t)z * g.dims[1] + y) * g.dims[0];
        const int beg = cell_start[row + x0], end = cell_start[row + x1 + 1];
        for (int q0 = beg; q0 < end && !removed; q0 += 32) {
          const int q = q0 + lane;
          bool rem = false, blk = false;
          if (q < end) {
            const int j = cell_order[q];
            const uint8_t sj = state[j];
            if (j != bi && sj != NMS_REMOVED) {
              double xj[D];
#pragma unroll
              for (int k = 0; k < D; ++k) xj[k] = modes[k * stride + j];
              double d = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) {
                const double t = __dsub_rn(bx[k], xj[k]);
                d = __dadd_rn(d, __dmul_rn(t, t));
              }
              if (d <= r2) {
                if (sj == NMS_KEEP) rem = true;
                else blk = has_priority<D>(counts[j], xj, j, bc, bx, bi);
              }
            }
          }
          removed = __any_sync(FULL, rem);
          blocked = blocked || __any_sync(FULL, blk);
        }
      }
    if (removed) {
      if (lane == 0) state[bi] = NMS_REMOVED;
      ++left;  // the cell may hold further live modes
      continue;
    }
    if (blocked) {
      ++left;
      continue;
    }
    // ---- 3. top survives: remove every live mode within its bandwidth
    if (lane == 0) state[bi] = NMS_KEEP;
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const int64_t row = ((int64_t)z * g.dims[1] + y) * g.dims[0];
        const int beg = cell_start[row + x0], end = cell_start[row + x1 + 1];
        for (int q = beg + lane; q < end; q += 32) {
          const int j = cell_order[q];
          if (j == bi || state[j] != NMS_UNDECIDED) continue;
          double d = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const double t = __dsub_rn(bx[k], modes[k * stride + j]);
            d = __dadd_rn(d, __dmul_rn(t, t));
          }
          if (d <= r2) state[j] = NMS_REMOVED;
        }
      }
    if (lane == 0) cell_done[w] = 1;  // everything else in this cell was within the bandwidth of top
  }
  if (lane == 0 && left) atomicAdd(undecided, left);
}

struct KeepPred {
  const uint8_t* state;
  __device__ __forceinline__ bool operator()(int64_t i) const { return state[i] == NMS_KEEP; }
};
struct IndexEmit {
  int* out;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const { out[d] = (int)i; }
};
__global__ void nms_counts_kernel(const long long* __restrict__ n_keep, const int* __restrict__ undecided,
                                  int* __restrict__ out2) {
  out2[0] = (int)*n_keep;
  out2[1] = *undecided;
}

// rank sort of the K survivors by priority (K is small: K^2 comparisons, one warp per survivor, beat 26 radix passes)
template <int D>
__global__ void __launch_bounds__(256)
nms_rank_emit_kernel(const double* __restrict__ modes, int64_t stride, const int* __restrict__ counts,
                     const int* __restrict__ keep_idx, int k, double* __restrict__ centres, int64_t centre_stride) {
  const int lane = lane_id();
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (a >= k) return;
  const int i = keep_idx[a];
  double xi[D];
#pragma unroll
  for (int d = 0; d < D; ++d) xi[d] = modes[d * stride + i];
  const int ci = counts[i];
  int rank = 0;
  for (int b = lane; b < k; b += 32) {
    const int j = keep_idx[b];
    double xj[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xj[d] = modes[d * stride + j];
    rank += (j != i && has_priority<D>(counts[j], xj, j, ci, xi, i)) ? 1 : 0;
  }
  rank = warp_sum(rank);
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) centres[d * centre_stride + rank] = xi[d];
  }
}

// ---- large-K path: stable LSD radix sorts over the survivors (least-significant key first) ----
__device__ __forceinline__ unsigned long long sortable(double v) {
  v = v + 0.0;  // fold -0.0 onto +0.0, as == does
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
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
template <int D>
__global__ void __launch_bounds__(256)
gather_centres_kernel(const double* __restrict__ modes, int64_t stride, const int* __restrict__ perm, int64_t k,
                      double* __restrict__ centres, int64_t centre_stride) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < k; a += gs) {
    const int i = perm[a];
#pragma unroll
    for (int d = 0; d < D; ++d) centres[d * centre_stride + a] = modes[d * stride + i];
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr int NMS_RANK_SORT_MAX = 8192;

struct NmsLayout {
  size_t sort64_bytes, sort32_bytes;
  size_t off_sort, off_state, off_ckeys_a, off_ckeys_b, off_cidx_a, off_cidx_b, off_cell_start, off_undecided,
      off_nkeep, off_ncells_live, off_cells, off_cell_done, off_keep_idx, off_keys_a, off_keys_b, off_perm_b,
      off_compact, total;
};

static NmsLayout nms_layout(int64_t n, int64_t n_fine_cells) {
  NmsLayout L{};
  cub::DeviceRadixSort::SortPairsDescending(nullptr, L.sort64_bytes, (const unsigned long long*)nullptr,
                                            (unsigned long long*)nullptr, (const int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairs(nullptr, L.sort32_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 256); return at; };
  L.off_sort = take(L.sort64_bytes > L.sort32_bytes ? L.sort64_bytes : L.sort32_bytes);
  L.off_state = take((size_t)n);
  L.off_ckeys_a = take(4 * (size_t)n);
  L.off_ckeys_b = take(4 * (size_t)n);
  L.off_cidx_a = take(4 * (size_t)n);
  L.off_cidx_b = take(4 * (size_t)n);
  L.off_cell_start = take(4 * (size_t)(n_fine_cells + 2));
  L.off_undecided = take(4 * 64);
  L.off_nkeep = take(8);
  L.off_ncells_live = take(8);
  L.off_cells = take(4 * (size_t)n);
  L.off_cell_done = take((size_t)n);
  L.off_keep_idx = take(4 * (size_t)n);
  L.off_keys_a = take(8 * (size_t)n);
  L.off_keys_b = take(8 * (size_t)n);
  L.off_perm_b = take(4 * (size_t)n);
  L.off_compact = take((size_t)CompactWorkspace::bytes(n));
  L.total = o + 256;
  return L;
}

constexpr int NMS_MAX_ROUNDS_PER_CALL = 16;

template <int D>
static int nms_suppress(const double* modes, int64_t stride, const int* counts, int64_t n, double bandwidth,
                        const cb200_grid* grid, int rounds, int resume, int* out2, void* workspace,
                        int64_t workspace_bytes, cudaStream_t st) {
  NmsGrid g;
  if (!fine_grid_from(*grid, bandwidth, g)) return CB200_EUNSUPPORTED;
  const NmsLayout L = nms_layout(n, g.n_cells);
  if ((int64_t)L.total > workspace_bytes) return CB200_EINVAL;
  if (rounds < 1 || rounds > NMS_MAX_ROUNDS_PER_CALL) return CB200_EINVAL;
  char* w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  void* sort_ws = w + L.off_sort;
  uint8_t* state = (uint8_t*)(w + L.off_state);
  unsigned* ckeys_a = (unsigned*)(w + L.off_ckeys_a);
  unsigned* ckeys_b = (unsigned*)(w + L.off_ckeys_b);
  int* cidx_a = (int*)(w + L.off_cidx_a);
  int* cidx_b = (int*)(w + L.off_cidx_b);
  int* cell_start = (int*)(w + L.off_cell_start);
  int* undecided = (int*)(w + L.off_undecided);  // slot (r & 1) = after round r-1; ping-pong
  long long* n_keep = (long long*)(w + L.off_nkeep);
  long long* n_cells_live = (long long*)(w + L.off_ncells_live);
  int* cells = (int*)(w + L.off_cells);
  uint8_t* cell_done = (uint8_t*)(w + L.off_cell_done);
  int* keep_idx = (int*)(w + L.off_keep_idx);
  void* compact_ws = w + L.off_compact;
  const int blocks = grid_for(n, 256, 2, 16);

  if (!resume) {
    // fine grid hash of the modes: cell id -> stable radix sort -> cell_start; list of non-empty cells
    nms_cell_ids_kernel<D><<<blocks, 256, 0, st>>>(modes, stride, counts, n, g, ckeys_a, cidx_a);
    CB200_LAUNCH_CHECK();
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) < (int64_t)g.n_cells + 1) ++bits;
    size_t sb32 = L.sort32_bytes;
    CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_ws, sb32, ckeys_a, ckeys_b, cidx_a, cidx_b, (int)n, 0, bits, st));
    cell_start_kernel<<<grid_for((int64_t)g.n_cells + 2, 256, 1, 16), 256, 0, st>>>(ckeys_b, n, (int64_t)g.n_cells + 1,
                                                                                    cell_start);
    CB200_LAUNCH_CHECK();
    CellHeadPred hp{ckeys_b, (unsigned)g.n_cells};
    CellHeadEmit he{ckeys_b, cells};
    const int rc0 = run_compaction(hp, he, n, n, n_cells_live, compact_ws, st);
    if (rc0 != CB200_OK) return rc0;
    CB200_CUDA_TRY(cudaMemsetAsync(cell_done, 0, (size_t)n, st));
    nms_init_kernel<<<blocks, 256, 0, st>>>(counts, n, state, undecided);
    CB200_LAUNCH_CHECK();
  }
  const double r2 = bandwidth * bandwidth;
  const int round_blocks = (int)std::min<int64_t>((n + 7) / 8, (int64_t)CB200_SM_COUNT * 8);  // warp per cell
  for (int r = 0; r < rounds; ++r) {
    int* prev = undecided + (r & 1);
    int* cur = undecided + ((r + 1) & 1);
    CB200_CUDA_TRY(cudaMemsetAsync(cur, 0, sizeof(int), st));
    nms_round_kernel<D><<<round_blocks, 256, 0, st>>>(modes, stride, counts, g, cell_start, cidx_b, cells, n_cells_live,
                                                      cell_done, r2, state, prev, cur);
    CB200_LAUNCH_CHECK();
  }
  int* last = undecided + (rounds & 1);
  if (rounds & 1) {  // keep the invariant "slot 0 = state after the last round" for a resumed call
    CB200_CUDA_TRY(cudaMemcpyAsync(undecided, last, sizeof(int), cudaMemcpyDeviceToDevice, st));
    last = undecided;
  }
  KeepPred pred{state};
  IndexEmit emit{keep_idx};
  const int rc = run_compaction(pred, emit, n, n, n_keep, compact_ws, st);
  if (rc != CB200_OK) return rc;
  nms_counts_kernel<<<1, 1, 0, st>>>(n_keep, last, out2);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

template <int D>
static int nms_emit(const double* modes, int64_t stride, const int* counts, int64_t n, const cb200_grid* grid,
                    double bandwidth, int k, double* centres, int64_t centre_stride, void* workspace,
                    int64_t workspace_bytes, cudaStream_t st) {
  NmsGrid g;
  if (!fine_grid_from(*grid, bandwidth, g)) return CB200_EUNSUPPORTED;
  const NmsLayout L = nms_layout(n, g.n_cells);
  if ((int64_t)L.total > workspace_bytes || k < 0 || k > n || centre_stride < k) return CB200_EINVAL;
  if (k == 0) return CB200_OK;
  char* w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  int* keep_idx = (int*)(w + L.off_keep_idx);
  if (k <= NMS_RANK_SORT_MAX) {
    nms_rank_emit_kernel<D><<<(k + 7) / 8, 256, 0, st>>>(modes, stride, counts, keep_idx, k, centres, centre_stride);
    CB200_LAUNCH_CHECK();
    return CB200_OK;
  }
  // K is large: LSD radix over the survivors -- last coordinate ... first coordinate, then the count.
  // keep_idx is ascending, so the stable sorts also resolve exact ties to the lower index.
  void* sort_ws = w + L.off_sort;
  auto* keys_a = (unsigned long long*)(w + L.off_keys_a);
  auto* keys_b = (unsigned long long*)(w + L.off_keys_b);
  int* perm_a = keep_idx;
  int* perm_b = (int*)(w + L.off_perm_b);
  const int blocks = grid_for(k, 256, 2, 16);
  size_t sb = L.sort64_bytes;
  for (int d = D - 1; d >= 0; --d) {
    key_from_coord_kernel<<<blocks, 256, 0, st>>>(modes + d * stride, perm_a, k, keys_a);
    CB200_LAUNCH_CHECK();
    CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(sort_ws, sb, keys_a, keys_b, perm_a, perm_b, k, 0, 64, st));
    int* t = perm_a; perm_a = perm_b; perm_b = t;
  }
  key_from_count_kernel<<<blocks, 256, 0, st>>>(counts, perm_a, k, keys_a);
  CB200_LAUNCH_CHECK();
  CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(sort_ws, sb, keys_a, keys_b, perm_a, perm_b, k, 0, 32, st));
  gather_centres_kernel<D><<<blocks, 256, 0, st>>>(modes, stride, perm_b, k, centres, centre_stride);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

// ------------------------------------------------------------------ label assignment
constexpr int ASG_THREADS = 256;
constexpr int ASG_TILE = 1024;

template <int D>
__device__ __forceinline__ double sqdist_fma(const double (&x)[D], const double (&c)[D]) {
  double d = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double t = x[k] - c[k];
    d = fma(t, t, d);
  }
  return d;
}

// brute force: every centre, tiles through shared memory.  `list` (optional) restricts the work to the
// listed point indices (the orphans of the grid search).
template <int D, typename L>
__global__ void __launch_bounds__(ASG_THREADS)
assign_brute_kernel(const double* __restrict__ points, int64_t n, int64_t pts_stride,
                    const double* __restrict__ centres, int64_t centre_stride, int K,
                    const int32_t* __restrict__ pix_index, L* __restrict__ labels, const int* __restrict__ list,
                    const int* __restrict__ list_count) {
  __shared__ double s_c[D][ASG_TILE];
  const int64_t total = list ? (int64_t)*list_count : n;
  for (int64_t base = (int64_t)blockIdx.x * ASG_THREADS; base < total; base += (int64_t)gridDim.x * ASG_THREADS) {
    const int64_t t = base + threadIdx.x;
    const bool active = t < total;
    const int64_t i = active ? (list ? (int64_t)list[t] : t) : 0;
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = active ? __ldg(points + k * pts_stride + i) : 0.0;
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
        double c[D];
#pragma unroll
        for (int k = 0; k < D; ++k) c[k] = s_c[k][j];
        const double d = sqdist_fma<D>(x, c);
        if (d < best) {  // strict: ties keep the lowest index (pairwise_distances_argmin)
          best = d;
          best_k = k0 + j;
        }
      }
    }
    if (active) {
      const int64_t dst = pix_index ? (int64_t)pix_index[i] : i;
      labels[dst] = (L)(best_k + 1);  // +1: 0 is background (utils/mean_shift.py:57)
    }
  }
}

// grid of the centres: per-cell singly linked lists (insertion order is irrelevant: the search keeps the
// lexicographic minimum of (distance, index))
template <int D>
__global__ void __launch_bounds__(256)
centre_grid_insert_kernel(const double* __restrict__ centres, int64_t centre_stride, int K, NmsGrid g,
                          int* __restrict__ head, int* __restrict__ next) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= K) return;
  double x[D];
#pragma unroll
  for (int k = 0; k < D; ++k) x[k] = centres[k * centre_stride + a];
  int c[3];
  cell_coords<D>(x, g, c);
  const int cell = (c[2] * g.dims[1] + c[1]) * g.dims[0] + c[0];
  next[a] = atomicExch(head + cell, a);
}

// A centre found within one cell edge (>= bandwidth) in the 3^D block is the global nearest: everything
// outside the block is strictly farther than one edge.  Points with no such centre are listed as orphans
// and finished by the brute-force kernel (`predict` labels them too, sklearn:563-579).
template <int D, typename L>
__global__ void __launch_bounds__(256)
assign_grid_kernel(const double* __restrict__ points, int64_t n, int64_t pts_stride,
                   const double* __restrict__ centres, int64_t centre_stride, NmsGrid g, double cell2,
                   const int* __restrict__ head, const int* __restrict__ next, const int32_t* __restrict__ pix_index,
                   L* __restrict__ labels, int* __restrict__ orphans, int* __restrict__ orphan_count) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = __ldg(points + k * pts_stride + i);
    int c[3];
    // unclamped cell: a point outside the centres' grid simply finds fewer neighbour cells
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double f = floor((x[k] - g.origin[k]) * g.inv_cell);
      if (!(f >= -1.0) || !(f <= (double)g.dims[k])) inside = false;
      c[k] = inside ? (int)f : 0;
    }
    double best = INFINITY;
    int best_k = 0x7fffffff;
    if (inside) {
      const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.dims[0] - 1);
      const int y0 = max(c[1] - 1, 0), y1 = min(c[1] + 1, g.dims[1] - 1);
      const int z0 = D == 3 ? max(c[2] - 1, 0) : 0, z1 = D == 3 ? min(c[2] + 1, g.dims[2] - 1) : 0;
      for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y)
          for (int xx = x0; xx <= x1; ++xx) {
            for (int a = head[(z * g.dims[1] + y) * g.dims[0] + xx]; a >= 0; a = next[a]) {
              double cc[D];
#pragma unroll
              for (int k = 0; k < D; ++k) cc[k] = __ldg(centres + k * centre_stride + a);
              const double d = sqdist_fma<D>(x, cc);
              if (d < best || (d == best && a < best_k)) {
                best = d;
                best_k = a;
              }
            }
          }
    }
    if (best <= cell2) {
      const int64_t dst = pix_index ? (int64_t)pix_index[i] : i;
      labels[dst] = (L)(best_k + 1);
    } else {
      orphans[atomicAdd(orphan_count, 1)] = (int)i;
    }
  }
}

// ---- exact duplicates first ------------------------------------------------------------------------
// A flat-kernel hill climb has finitely many fixed points: seeds that end in the same window of points end in the SAME
// mean, bit for bit (same points, same summation order).  At BASELINE configs[2] 148 228 converged seeds are 395 distinct
// modes; on the 272^3 sharded volume 3.23 M seeds are 963.  scikit-learn merges them in a dict (sklearn:514-521); the
// suppression rounds above treat them as candidates that remove each other.  This pass keeps ONE copy of every distinct
// mode -- the one the suppression would keep anyway: highest count, then lowest index (has_priority) -- so that the rounds
// run on hundreds of modes instead of hundreds of thousands.  A removed copy can never influence another mode (it is
// removed by its own kept copy, or by whatever removes that), so the surviving centres and their order are unchanged.
// Open-addressing table keyed by the modes' bit patterns; a slot holds the best (count, index) seen for its mode.
__device__ __forceinline__ uint64_t mode_hash(const double* __restrict__ modes, int64_t stride, int D, int64_t i) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < D; ++k) {
    uint64_t x = (uint64_t)__double_as_longlong(modes[k * stride + i]);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    h = (h ^ x) * 0x9E3779B97F4A7C15ull;
  }
  return h ^ (h >> 32);
}
__device__ __forceinline__ bool same_mode(const double* __restrict__ modes, int64_t stride, int D, int64_t i, int64_t j) {
  for (int k = 0; k < D; ++k)
    if (__double_as_longlong(modes[k * stride + i]) != __double_as_longlong(modes[k * stride + j])) return false;
  return true;
}
// packed priority: larger = kept; 0 = empty slot
__device__ __forceinline__ unsigned long long mode_priority(int count, int64_t i) {
  return ((unsigned long long)(unsigned)(count + 1) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
}
__device__ __forceinline__ int64_t priority_index(unsigned long long p) { return (int64_t)(0xffffffffu - (unsigned)(p & 0xffffffffull)); }

__global__ void __launch_bounds__(256)
unique_insert_kernel(const double* __restrict__ modes, int64_t stride, int D, const int* __restrict__ counts, int64_t n,
                     unsigned long long* __restrict__ table, uint64_t mask) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int cnt = counts[i];
    if (cnt <= 0) continue;  // a seed without neighbours is dropped (sklearn:511-513)
    const unsigned long long mine = mode_priority(cnt, i);
    uint64_t slot = mode_hash(modes, stride, D, i) & mask;
    while (true) {
      unsigned long long cur = *((volatile unsigned long long*)&table[slot]);
      if (cur == 0ull) {
        cur = atomicCAS(&table[slot], 0ull, mine);
        if (cur == 0ull) break;  // claimed an empty slot for this mode
      }
      if (same_mode(modes, stride, D, i, priority_index(cur))) {  // the occupant is a copy of this mode (copies only
        if (mine > cur) atomicMax(&table[slot], mine);             // replace copies, so the test stays valid)
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

// a mode is kept iff it is the best copy recorded for its bit pattern
struct UniquePred {
  const double* modes;
  int64_t stride;
  int D;
  const int* counts;
  const unsigned long long* table;
  uint64_t mask;
  __device__ __forceinline__ bool operator()(int64_t i) const {
    if (counts[i] <= 0) return false;
    uint64_t slot = mode_hash(modes, stride, D, i) & mask;
    while (true) {
      const unsigned long long cur = table[slot];
      if (cur == 0ull) return false;  // cannot happen: every live mode was inserted
      const int64_t j = priority_index(cur);
      if (j == i) return true;
      if (same_mode(modes, stride, D, i, j)) return false;
      slot = (slot + 1) & mask;
    }
  }
};
struct UniqueEmit {
  const double* modes;
  int64_t stride;
  int D;
  const int* counts;
  double* out;
  int64_t out_stride;
  int* counts_out;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const {
    for (int k = 0; k < D; ++k) out[k * out_stride + d] = modes[k * stride + i];
    counts_out[d] = counts[i];
  }
};

// ---- distinct trajectories ---------------------------------------------------------------------------
// Two seeds whose means are bit-identical after the same number of iterations follow the same trajectory from there on.
// After ONE window evaluation most seeds of an object already share their mean (the window of every seed near the
// centre holds the same points); cb200_ms_grid_modes_distinct climbs only one representative (the lowest seed index) of
// every distinct unfinished mean and drops the copies (count = 0: they would end as copies of the representative's
// mode, which the suppression merges anyway).
constexpr int MS_DROPPED = INT_MIN;  // iters[] of a seed that was merged into its representative
constexpr int MS_MAX_MERGE_ROUNDS = 30;

// tests / window evaluations of the launches after the first, summed into the caller's second statistics block
__global__ void ms_stats_sum_kernel(const int* __restrict__ blocks, int n_blocks, int* __restrict__ out8) {
  if (threadIdx.x == 0) {
    unsigned long long tests = 0, steps = 0;
    for (int b = 0; b < n_blocks; ++b) {
      tests += *reinterpret_cast<const unsigned long long*>(blocks + 8 * b + 2);
      steps += *reinterpret_cast<const unsigned long long*>(blocks + 8 * b + 4);
    }
    *reinterpret_cast<unsigned long long*>(out8 + 2) = tests;
    *reinterpret_cast<unsigned long long*>(out8 + 4) = steps;
  }
}

__global__ void __launch_bounds__(256)
unfinished_insert_kernel(const double* __restrict__ means, int64_t stride, int D, const int* __restrict__ iters, int64_t n,
                         unsigned long long* __restrict__ table, uint64_t mask) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    if (iters[i] >= 0 || iters[i] == MS_DROPPED) continue;  // finished, or merged into another seed earlier
    const unsigned long long mine = (unsigned long long)i + 1ull;
    uint64_t slot = mode_hash(means, stride, D, i) & mask;
    while (true) {
      unsigned long long cur = *((volatile unsigned long long*)&table[slot]);
      if (cur == 0ull) {
        cur = atomicCAS(&table[slot], 0ull, mine);
        if (cur == 0ull) break;
      }
      if (same_mode(means, stride, D, i, (int64_t)(cur - 1ull))) {
        if (mine < cur) atomicMin(&table[slot], mine);
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

// representatives go to the worklist; the copies are dropped (idempotent side effect: the predicate runs twice)
struct RepresentativePred {
  const double* means;
  int64_t stride;
  int D;
  int* iters;
  int* counts;
  const unsigned long long* table;
  uint64_t mask;
  __device__ __forceinline__ bool operator()(int64_t i) const {
    if (iters[i] >= 0 || iters[i] == MS_DROPPED) return false;
    uint64_t slot = mode_hash(means, stride, D, i) & mask;
    while (true) {
      const unsigned long long cur = table[slot];
      if (cur == 0ull) return false;  // cannot happen
      const int64_t j = (int64_t)(cur - 1ull);
      if (j == i) return true;
      if (same_mode(means, stride, D, i, j)) {
        counts[i] = 0;
        iters[i] = MS_DROPPED;
        return false;
      }
      slot = (slot + 1) & mask;
    }
  }
};
struct WorklistEmit {
  int* worklist;
  __device__ __forceinline__ void operator()(int64_t i, long long d) const { worklist[d] = (int)i; }
};

static uint64_t unique_table_slots(int64_t n) {
  uint64_t slots = 1024;
  while (slots < (uint64_t)(2 * n)) slots <<= 1;
  return slots;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_nms_workspace_bytes(int64_t n_seeds, const cb200_grid* grid, double bandwidth) {
  if (n_seeds <= 0 || !grid) return 512;
  NmsGrid g;
  if (!fine_grid_from(*grid, bandwidth, g)) return -1;
  return (int64_t)nms_layout(n_seeds, g.n_cells).total;
}

int cb200_nms_suppress(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                       double bandwidth, const cb200_grid* grid, int rounds, int resume,
                       int* n_keep_and_undecided, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!modes || !counts || !grid || !n_keep_and_undecided || !workspace || n_seeds <= 0) return CB200_EINVAL;
  if (n_seeds > INT32_MAX || !(grid->cell >= bandwidth)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_dims == 2)
    return nms_suppress<2>(modes, seed_stride, counts, n_seeds, bandwidth, grid, rounds, resume, n_keep_and_undecided,
                           workspace, workspace_bytes, st);
  if (num_dims == 3)
    return nms_suppress<3>(modes, seed_stride, counts, n_seeds, bandwidth, grid, rounds, resume, n_keep_and_undecided,
                           workspace, workspace_bytes, st);
  return CB200_EUNSUPPORTED;
}

int cb200_nms_emit(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                   double bandwidth, const cb200_grid* grid, int n_keep, double* centres_out, int64_t centre_stride,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  if (!modes || !counts || !grid || !centres_out || !workspace || n_seeds <= 0) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_dims == 2)
    return nms_emit<2>(modes, seed_stride, counts, n_seeds, grid, bandwidth, n_keep, centres_out, centre_stride,
                       workspace, workspace_bytes, st);
  if (num_dims == 3)
    return nms_emit<3>(modes, seed_stride, counts, n_seeds, grid, bandwidth, n_keep, centres_out, centre_stride,
                       workspace, workspace_bytes, st);
  return CB200_EUNSUPPORTED;
}

int64_t cb200_assign_workspace_bytes(int64_t n_points, int n_centres, int64_t n_cells) {
  return (int64_t)(align_up(4 * (size_t)std::max<int64_t>(n_cells, 1), 256) + align_up(4 * (size_t)std::max(n_centres, 1), 256) +
                   align_up(4 * (size_t)std::max<int64_t>(n_points, 1), 256) + 512);
}

int cb200_assign_labels(const double* points, int64_t n_points, int64_t pts_stride, int num_dims,
                        const double* centres, int64_t centre_stride, int n_centres, const cb200_grid* grid,
                        const int32_t* pix_index, void* labels_out, int label_dtype, void* workspace, void* stream) {
  if (!points || !centres || !labels_out || n_points < 0 || n_centres <= 0) return CB200_EINVAL;
  if (n_points == 0) return CB200_OK;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  if (label_dtype != CB200_I32 && label_dtype != CB200_U16) return CB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int brute_blocks = (int)std::min<int64_t>((n_points + ASG_THREADS - 1) / ASG_THREADS, (int64_t)CB200_SM_COUNT * 16);
#define CB200_BRUTE(DD, LT, LIST, COUNT, BLOCKS)                                                                     \
  assign_brute_kernel<DD, LT><<<BLOCKS, ASG_THREADS, 0, st>>>(points, n_points, pts_stride, centres, centre_stride,   \
                                                             n_centres, pix_index, (LT*)labels_out, LIST, COUNT)
  const bool use_grid = grid && workspace && grid->n_cells < ((int64_t)1 << 31) && n_points <= INT32_MAX;
  if (!use_grid) {
    if (num_dims == 2 && label_dtype == CB200_I32) CB200_BRUTE(2, int32_t, nullptr, nullptr, brute_blocks);
    else if (num_dims == 2) CB200_BRUTE(2, uint16_t, nullptr, nullptr, brute_blocks);
    else if (label_dtype == CB200_I32) CB200_BRUTE(3, int32_t, nullptr, nullptr, brute_blocks);
    else CB200_BRUTE(3, uint16_t, nullptr, nullptr, brute_blocks);
    CB200_LAUNCH_CHECK();
    return CB200_OK;
  }
  char* w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  int* head = (int*)w;                w += align_up(4 * (size_t)grid->n_cells, 256);
  int* next = (int*)w;                w += align_up(4 * (size_t)n_centres, 256);
  int* orphans = (int*)w;             w += align_up(4 * (size_t)n_points, 256);
  int* orphan_count = (int*)w;
  const NmsGrid g = to_nms_grid(*grid);
  CB200_CUDA_TRY(cudaMemsetAsync(head, 0xff, 4 * (size_t)grid->n_cells, st));
  CB200_CUDA_TRY(cudaMemsetAsync(orphan_count, 0, sizeof(int), st));
  const double cell2 = grid->cell * grid->cell;
  const int gblocks = grid_for(n_points, 256, 1, 16);
#define CB200_GRID(DD, LT)                                                                                            \
  centre_grid_insert_kernel<DD><<<(n_centres + 255) / 256, 256, 0, st>>>(centres, centre_stride, n_centres, g, head,   \
                                                                          next);                                       \
  assign_grid_kernel<DD, LT><<<gblocks, 256, 0, st>>>(points, n_points, pts_stride, centres, centre_stride, g, cell2,   \
                                                      head, next, pix_index, (LT*)labels_out, orphans, orphan_count); \
  CB200_BRUTE(DD, LT, orphans, orphan_count, CB200_SM_COUNT * 2)
  if (num_dims == 2 && label_dtype == CB200_I32) { CB200_GRID(2, int32_t); }
  else if (num_dims == 2) { CB200_GRID(2, uint16_t); }
  else if (label_dtype == CB200_I32) { CB200_GRID(3, int32_t); }
  else { CB200_GRID(3, uint16_t); }
#undef CB200_GRID
#undef CB200_BRUTE
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int64_t cb200_unique_modes_workspace_bytes(int64_t n_seeds) {
  if (n_seeds < 0) return -1;
  return (int64_t)(unique_table_slots(n_seeds) * sizeof(unsigned long long)) + CompactWorkspace::bytes(n_seeds) + 512;
}

int cb200_unique_modes(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                       double* modes_out, int64_t out_stride, int* counts_out, long long* n_out, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  if (!modes || !counts || !modes_out || !counts_out || !n_out || !workspace || n_seeds < 0 || num_dims < 1 || num_dims > 3)
    return CB200_EINVAL;
  if (n_seeds >= ((int64_t)1 << 32) - 1 || workspace_bytes < cb200_unique_modes_workspace_bytes(n_seeds)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_seeds == 0) {
    CB200_CUDA_TRY(cudaMemsetAsync(n_out, 0, sizeof(long long), st));
    return CB200_OK;
  }
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  const uint64_t slots = unique_table_slots(n_seeds);
  auto* table = reinterpret_cast<unsigned long long*>(w);
  void* compact_ws = w + slots * sizeof(unsigned long long);
  CB200_CUDA_TRY(cudaMemsetAsync(table, 0, slots * sizeof(unsigned long long), st));
  unique_insert_kernel<<<grid_for(n_seeds, 256, 1, 16), 256, 0, st>>>(modes, seed_stride, num_dims, counts, n_seeds, table,
                                                                      slots - 1);
  CB200_LAUNCH_CHECK();
  UniquePred pred{modes, seed_stride, num_dims, counts, table, slots - 1};
  UniqueEmit emit{modes, seed_stride, num_dims, counts, modes_out, out_stride, counts_out};
  return run_compaction(pred, emit, n_seeds, out_stride, n_out, compact_ws, st);
}

int64_t cb200_ms_distinct_workspace_bytes(int64_t n_seeds) {
  if (n_seeds < 0) return -1;
  return (int64_t)(unique_table_slots(n_seeds) * sizeof(unsigned long long)) + ((4 * n_seeds + 255) / 256 * 256) +
         CompactWorkspace::bytes(n_seeds) + 2048;
}

int cb200_ms_grid_modes_distinct(const double* points_sorted, int64_t n_points, int64_t sorted_stride, const cb200_grid* grid,
                                 const int* cell_start, double* means, int64_t seed_stride, int64_t n_seeds,
                                 double bandwidth, int max_iter, int merge_rounds, int* counts, int* iters, int* work,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  if (!workspace || !work || n_seeds < 0 || workspace_bytes < cb200_ms_distinct_workspace_bytes(n_seeds)) return CB200_EINVAL;
  if (merge_rounds < 1 || merge_rounds > MS_MAX_MERGE_ROUNDS) return CB200_EINVAL;
  if (n_seeds == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = grid ? grid->num_dims : 0;
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  const uint64_t slots = unique_table_slots(n_seeds);
  auto* table = reinterpret_cast<unsigned long long*>(w);
  w += slots * sizeof(unsigned long long);
  int* worklist = reinterpret_cast<int*>(w);
  w += (4 * n_seeds + 255) / 256 * 256;
  long long* n_work = reinterpret_cast<long long*>(w);
  w += 256;
  int* blocks = reinterpret_cast<int*>(w);  // claim counter + statistics of every launch after the first: 8 ints each
  w += 1024;
  CB200_CUDA_TRY(cudaMemsetAsync(blocks, 0, 1024, st));
  // first evaluation of every seed; unconverged seeds are left with iters = -2
  int rc = ms_grid_modes_launch(points_sorted, n_points, sorted_stride, grid, cell_start, means, seed_stride, n_seeds, bandwidth,
                                max_iter, counts, iters, work, nullptr, nullptr, 1, st);
  if (rc != CB200_OK) return rc;
  for (int r = 1; r <= merge_rounds; ++r) {
    // merge the unfinished seeds by the bit pattern of their mean (all of them have done r iterations)
    CB200_CUDA_TRY(cudaMemsetAsync(table, 0, slots * sizeof(unsigned long long), st));
    unfinished_insert_kernel<<<grid_for(n_seeds, 256, 1, 16), 256, 0, st>>>(means, seed_stride, D, iters, n_seeds, table, slots - 1);
    CB200_LAUNCH_CHECK();
    RepresentativePred pred{means, seed_stride, D, iters, counts, table, slots - 1};
    WorklistEmit emit{worklist};
    rc = run_compaction(pred, emit, n_seeds, n_seeds, n_work, w, st);
    if (rc != CB200_OK) return rc;
    // the representatives: one more evaluation each, or -- after the last merge -- to convergence
    rc = ms_grid_modes_launch(points_sorted, n_points, sorted_stride, grid, cell_start, means, seed_stride, n_seeds, bandwidth,
                              max_iter, counts, iters, blocks + 8 * (r - 1), worklist, n_work, r < merge_rounds ? 1 : 0, st);
    if (rc != CB200_OK) return rc;
  }
  ms_stats_sum_kernel<<<1, 32, 0, st>>>(blocks, merge_rounds, work + 8);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
