// Flat-kernel mean-shift mode seeking for sm_100a (sklearn _mean_shift.py:108-128).
//
// The embedding dimension is 2-3, so this is n-body work on the FP64/FP32
// pipes, not a tensor-core contraction.  Two forms:
//
//  * brute force: block = 128 seeds (one per thread) x one chunk of points;
//    point tiles are streamed global -> shared with the bulk async-copy engine
//    (cp.async.bulk + mbarrier, the 1-D TMA path; UBLKCP in SASS), double
//    buffered, and every thread reads each point as a shared-memory broadcast.
//    Per-chunk partial sums are written out and combined in a FIXED order by
//    the update kernel, so results are run-to-run deterministic.
//  * grid hash: points sorted by cell (edge >= bandwidth); one warp climbs one
//    seed all the way to convergence in a single launch, lanes striding over
//    the points of the 3^D neighbour cells, warp-shuffle reduction per step.
//
// Distances are evaluated in float64 exactly as the reference's KD-tree does
// (sequential mul/add, no FMA, inclusive <= bw^2) so that in/out decisions --
// which is what mode parity hinges on -- are the reference's.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace cb200 {

// reduced distance exactly as sklearn's KD-tree: d = 0; for k: t = m_k - x_k; d += t*t  (no FMA)
template <int D>
__device__ __forceinline__ double rdist(const double (&m)[D], const double (&x)[D]) {
  double d = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double t = __dsub_rn(m[k], x[k]);
    d = __dadd_rn(d, __dmul_rn(t, t));
  }
  return d;
}

// ------------------------------------------------------------------ brute force
constexpr int MSB_THREADS = 128;   // seeds per block
constexpr int MSB_TILE = 512;      // points per pipeline stage (3 stages x 3 dims x 4 KB = 36 KB static smem)
constexpr int MSB_STAGES = 3;
constexpr int MSB_TARGET_BLOCKS = CB200_SM_COUNT * 8;

struct BrutePlan {
  int64_t seed_tiles, chunks, chunk_points;
};
static BrutePlan brute_plan(int64_t n_active, int64_t n_points) {
  BrutePlan p;
  p.seed_tiles = (n_active + MSB_THREADS - 1) / MSB_THREADS;
  if (p.seed_tiles < 1) p.seed_tiles = 1;
  const int64_t tiles = (n_points + MSB_TILE - 1) / MSB_TILE;
  int64_t chunks = (MSB_TARGET_BLOCKS + p.seed_tiles - 1) / p.seed_tiles;
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  const int64_t tiles_per_chunk = (tiles + chunks - 1) / chunks;
  p.chunk_points = tiles_per_chunk * MSB_TILE;
  p.chunks = (n_points + p.chunk_points - 1) / p.chunk_points;
  if (p.chunks < 1) p.chunks = 1;
  return p;
}

// partial layout: [chunk][k = 0..D (D = count)][active index] doubles
template <int D>
__global__ void __launch_bounds__(MSB_THREADS)
ms_brute_accumulate_kernel(const double* __restrict__ points, int64_t n_points, int64_t pts_stride,
                           const double* __restrict__ means, int64_t seed_stride, const int* __restrict__ active,
                           int64_t n_active, double r2, int64_t chunk_points, double* __restrict__ partial) {
  __shared__ __align__(128) double s_pts[MSB_STAGES][D][MSB_TILE];
  __shared__ __align__(8) uint64_t s_full[MSB_STAGES];

  const int64_t a = (int64_t)blockIdx.x * MSB_THREADS + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * chunk_points;
  const int64_t c1 = min(c0 + chunk_points, n_points);
  const int n_tiles = (int)((c1 - c0 + MSB_TILE - 1) / MSB_TILE);

  double m[D];
  if (a < n_active) {
    const int s = active[a];
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = means[k * seed_stride + s];
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = __longlong_as_double(0x7ff8000000000000ll);  // NaN: never in range
  }

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < MSB_STAGES; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int tile) {
    const int st = tile % MSB_STAGES;
    const int64_t p0 = c0 + (int64_t)tile * MSB_TILE;
    int cnt = (int)min((int64_t)MSB_TILE, c1 - p0);
    cnt = (cnt + 1) & ~1;  // 16-byte granules; the pad element (if any) lies inside the even stride
    const uint32_t bytes = (uint32_t)cnt * 8u;
    mbar_expect_tx(&s_full[st], bytes * D);
#pragma unroll
    for (int k = 0; k < D; ++k) bulk_g2s(&s_pts[st][k][0], points + k * pts_stride + p0, bytes, &s_full[st]);
  };

  if (threadIdx.x == 0) {
    for (int t = 0; t < MSB_STAGES && t < n_tiles; ++t) issue(t);
  }

  double sum[D];
#pragma unroll
  for (int k = 0; k < D; ++k) sum[k] = 0.0;
  int cnt_in = 0;

  for (int t = 0; t < n_tiles; ++t) {
    const int st = t % MSB_STAGES;
    mbar_wait(&s_full[st], (uint32_t)((t / MSB_STAGES) & 1));
    const int64_t p0 = c0 + (int64_t)t * MSB_TILE;
    const int cnt = (int)min((int64_t)MSB_TILE, c1 - p0);
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      double x[D];
#pragma unroll
      for (int k = 0; k < D; ++k) x[k] = s_pts[st][k][j];  // broadcast read
      const bool in = rdist<D>(m, x) <= r2;
      if (__any_sync(FULL, in)) {  // rare for brute force: most (seed, point) pairs are far apart
        if (in) {
#pragma unroll
          for (int k = 0; k < D; ++k) sum[k] += x[k];
          ++cnt_in;
        }
      }
    }
    __syncthreads();  // every thread is done with stage st before it is refilled
    if (threadIdx.x == 0 && t + MSB_STAGES < n_tiles) issue(t + MSB_STAGES);
  }

  if (a < n_active) {
    double* out = partial + (int64_t)blockIdx.y * (D + 1) * n_active + a;
#pragma unroll
    for (int k = 0; k < D; ++k) out[k * n_active] = sum[k];
    out[D * n_active] = (double)cnt_in;
  }
}

// one hill-climb step bookkeeping per active seed (sklearn:113-127)
template <int D>
__global__ void __launch_bounds__(256)
ms_update_kernel(double* __restrict__ means, int64_t seed_stride, int* __restrict__ counts, int* __restrict__ iters,
                 const int* __restrict__ active, int64_t n_active, int64_t chunks, const double* __restrict__ partial,
                 double stop, int max_iter, int* __restrict__ next_active, int* __restrict__ n_next) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active) return;
  const int s = active[a];
  double sum[D];
#pragma unroll
  for (int k = 0; k < D; ++k) sum[k] = 0.0;
  double cnt = 0.0;
  for (int64_t c = 0; c < chunks; ++c) {  // fixed order -> deterministic
    const double* p = partial + c * (D + 1) * n_active + a;
#pragma unroll
    for (int k = 0; k < D; ++k) sum[k] += p[k * n_active];
    cnt += p[D * n_active];
  }
  const int n = (int)cnt;
  counts[s] = n;
  if (n == 0) return;  // empty window: the seed stops where it is and is dropped later (:115-116, :511-513)
  double shift2 = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double old = means[k * seed_stride + s];
    const double nw = sum[k] / cnt;  // np.mean(points_within, axis=0)
    means[k * seed_stride + s] = nw;
    const double d = nw - old;
    shift2 += d * d;
  }
  const int it = iters[s];
  if (sqrt(shift2) <= stop || it == max_iter) return;  // converged or out of iterations (:120-125)
  iters[s] = it + 1;
  next_active[atomicAdd(n_next, 1)] = s;
}

// ------------------------------------------------------------------ grid hash
struct GridDev {
  double origin[3];
  double inv_cell;
  int dims[3];
};

static GridDev to_dev(const cb200_grid& g) {
  GridDev d;
  for (int k = 0; k < 3; ++k) {
    d.origin[k] = g.origin[k];
    d.dims[k] = g.dims[k];
  }
  d.inv_cell = g.inv_cell;
  return d;
}

template <int D>
__device__ __forceinline__ unsigned cell_of(const double (&x)[D], const GridDev& g) {
  unsigned id = 0;
#pragma unroll
  for (int k = D - 1; k >= 0; --k) {
    double f = floor((x[k] - g.origin[k]) * g.inv_cell);
    f = fmin(fmax(f, 0.0), (double)(g.dims[k] - 1));
    id = id * (unsigned)g.dims[k] + (unsigned)(int)f;
  }
  return id;
}

template <int D>
__global__ void __launch_bounds__(256)
grid_cell_ids_kernel(const double* __restrict__ points, int64_t n, int64_t stride, GridDev g,
                     unsigned* __restrict__ keys, int* __restrict__ idx) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    double x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = __ldg(points + k * stride + i);
    keys[i] = cell_of<D>(x, g);
    idx[i] = (int)i;
  }
}

template <int D>
__global__ void __launch_bounds__(256)
grid_gather_kernel(const double* __restrict__ points, int64_t n, int64_t stride, const int* __restrict__ order,
                   double* __restrict__ sorted, int64_t sorted_stride) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int src = order[i];
#pragma unroll
    for (int k = 0; k < D; ++k) sorted[k * sorted_stride + i] = __ldg(points + k * stride + src);
  }
}

// cell_start[c] = first sorted position whose key >= c  (c in [0, n_cells])
__global__ void __launch_bounds__(256)
grid_cell_start_kernel(const unsigned* __restrict__ sorted_keys, int64_t n, int64_t n_cells,
                       int* __restrict__ cell_start) {
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

#ifndef CB200_MSG_UNROLL
#define CB200_MSG_UNROLL 2
#endif
constexpr int MSG_UNROLL = CB200_MSG_UNROLL;  // trips of a row unrolled together (loads of later trips in flight)
#ifndef CB200_MSG_THREADS
#define CB200_MSG_THREADS 128
#endif
constexpr int MSG_THREADS = CB200_MSG_THREADS;

// One warp per seed, climbing to convergence.  Seeds are claimed dynamically so that
// slow climbers do not hold up a whole block.
#ifndef CB200_MSG_MINBLOCKS
#define CB200_MSG_MINBLOCKS 8  // 64 registers: 32 resident warps per SM (measured: 0.506 ms at 72 regs / 24 warps -> 0.452 ms)
#endif
template <int D>
__global__ void __launch_bounds__(MSG_THREADS, CB200_MSG_MINBLOCKS)
ms_grid_modes_kernel(const double* __restrict__ pts, int64_t pts_stride, GridDev g,
                     const int* __restrict__ cell_start, double* __restrict__ means, int64_t seed_stride,
                     int64_t n_seeds, double r2, double stop, int max_iter, int* __restrict__ counts,
                     int* __restrict__ iters, int* __restrict__ work_counter, const int* __restrict__ worklist,
                     const long long* __restrict__ n_work_dev, int eval_limit) {
  // Two optional modes for the distinct-trajectory form (cb200_ms_grid_modes_distinct):
  //   eval_limit > 0   a seed that has not converged after `eval_limit` window evaluations is left UNFINISHED: its
  //                    current mean in `means`, iters[s] = -1 - (iterations done);
  //   worklist != NULL only the seeds listed there (their number is read on the device) are climbed, each resuming
  //                    from such an unfinished state.
  const int lane = lane_id();
  unsigned long long tests = 0;  // distance tests this lane made (statistics: the kernel's algorithmic work)
  unsigned long long steps = 0;  // window evaluations (iterations + 1 per seed), counted by lane 0
  const int64_t n_claims = worklist ? (int64_t)__ldg(n_work_dev) : n_seeds;
  while (true) {
    int s = 0;
    if (lane == 0) s = atomicAdd(work_counter, 1);
    s = __shfl_sync(FULL, s, 0);
    if (s >= n_claims) break;
    int it = 0, n_within = 0;
    if (worklist) {
      s = __ldg(worklist + s);
      it = -1 - iters[s];
    }
    const int it_first = it;
    bool unfinished = false;
    double m[D];
#pragma unroll
    for (int k = 0; k < D; ++k) m[k] = means[k * seed_stride + s];
    while (true) {
      // cell of the current mean; neighbour block clipped to the grid
      int c[3] = {0, 0, 0};
      bool any = true;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const double f = floor((m[k] - g.origin[k]) * g.inv_cell);
        if (!(f >= -1.0) || !(f <= (double)g.dims[k])) any = false;  // also catches NaN
        c[k] = any ? (int)f : 0;
      }
      double sum[D];
#pragma unroll
      for (int k = 0; k < D; ++k) sum[k] = 0.0;
      int cnt = 0;
      if (any) {
        const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.dims[0] - 1);
        const int y0 = max(c[1] - 1, 0), y1 = min(c[1] + 1, g.dims[1] - 1);
        const int z0 = D == 3 ? max(c[2] - 1, 0) : 0, z1 = D == 3 ? min(c[2] + 1, g.dims[2] - 1) : 0;
        if (x0 <= x1) {
          // the (up to 9) rows of x-adjacent cells: lane r fetches the bounds of row r, so that the two
          // cell_start loads of every row are in flight together instead of one dependent pair per row
          const int ny = y1 - y0 + 1, n_rows = ny * (z1 - z0 + 1);
          int my_beg = 0, my_end = 0;
          if (lane < n_rows) {
            const int rz = lane / ny, ry = lane - rz * ny;
            const int64_t row = ((int64_t)(z0 + rz) * g.dims[1] + (y0 + ry)) * g.dims[0];
            my_beg = __ldg(cell_start + row + x0);
            my_end = __ldg(cell_start + row + x1 + 1);
          }
          tests += (unsigned)(my_end - my_beg);  // candidates of this lane's row; summed over the warp at the end
          for (int r = 0; r < n_rows; ++r) {
            const int beg = __shfl_sync(FULL, my_beg, r);
            const int end = __shfl_sync(FULL, my_end, r);
#pragma unroll MSG_UNROLL
            for (int i = beg + lane; i < end; i += 32) {
              double x[D];
#pragma unroll
              for (int k = 0; k < D; ++k) x[k] = __ldg(pts + k * pts_stride + i);
              if (rdist<D>(m, x) <= r2) {
#pragma unroll
                for (int k = 0; k < D; ++k) sum[k] += x[k];
                ++cnt;
              }
            }
          }
        }
      }
      cnt = warp_sum(cnt);
      n_within = cnt;
      if (cnt == 0) break;  // sklearn:115-116
      double shift2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        double t = warp_sum_down(sum[k]);
        t = __shfl_sync(FULL, t, 0);
        const double nw = t / (double)cnt;
        const double d = nw - m[k];
        shift2 += d * d;
        m[k] = nw;
      }
      if (sqrt(shift2) <= stop || it == max_iter) break;  // sklearn:120-125
      ++it;
      if (eval_limit > 0 && it - it_first >= eval_limit) {
        unfinished = true;
        break;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) means[k * seed_stride + s] = m[k];
      counts[s] = n_within;
      iters[s] = unfinished ? -1 - it : it;
      steps += (unsigned)(it - it_first) + (unfinished ? 0u : 1u);
    }
  }
  // one statistics update per warp for the whole launch
  tests += __shfl_xor_sync(FULL, tests, 16);
  tests += __shfl_xor_sync(FULL, tests, 8);
  tests += __shfl_xor_sync(FULL, tests, 4);
  tests += __shfl_xor_sync(FULL, tests, 2);
  tests += __shfl_xor_sync(FULL, tests, 1);
  if (lane == 0 && tests) {
    atomicAdd(reinterpret_cast<unsigned long long*>(work_counter + 2), tests);
    atomicAdd(reinterpret_cast<unsigned long long*>(work_counter + 4), steps);
  }
}

// the warp-per-seed climb with its optional modes (see the kernel); `n_seeds` sizes the launch in every mode
int ms_grid_modes_launch(const double* points_sorted, int64_t n_points, int64_t sorted_stride, const cb200_grid* grid,
                         const int* cell_start, double* means, int64_t seed_stride, int64_t n_seeds, double bandwidth,
                         int max_iter, int* counts, int* iters, int* work_counter, const int* worklist,
                         const long long* n_work_dev, int eval_limit, cudaStream_t st) {
  if (!points_sorted || !grid || !cell_start || !means || !counts || !iters || !work_counter || n_seeds < 0)
    return CB200_EINVAL;
  if (n_seeds == 0) return CB200_OK;
  if (n_seeds > INT32_MAX || !(grid->cell >= bandwidth)) return CB200_EINVAL;
  (void)n_points;
  const GridDev g = to_dev(*grid);
  const int64_t warps_needed = n_seeds;
  int64_t blocks = (warps_needed + MSG_THREADS / 32 - 1) / (MSG_THREADS / 32);
  const int64_t cap = (int64_t)CB200_SM_COUNT * 8;  // persistent grid, dynamic seed claiming
  if (blocks > cap) blocks = cap;
  const double r2 = bandwidth * bandwidth, stop = 1e-3 * bandwidth;
  if (grid->num_dims == 2)
    ms_grid_modes_kernel<2><<<(int)blocks, MSG_THREADS, 0, st>>>(points_sorted, sorted_stride, g, cell_start, means,
                                                                 seed_stride, n_seeds, r2, stop, max_iter, counts,
                                                                 iters, work_counter, worklist, n_work_dev, eval_limit);
  else
    ms_grid_modes_kernel<3><<<(int)blocks, MSG_THREADS, 0, st>>>(points_sorted, sorted_stride, g, cell_start, means,
                                                                 seed_stride, n_seeds, r2, stop, max_iter, counts,
                                                                 iters, work_counter, worklist, n_work_dev, eval_limit);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_ms_brute_partial_bytes(int64_t n_active, int64_t n_points, int num_dims) {
  if (n_active <= 0 || n_points <= 0) return 16;
  const BrutePlan p = brute_plan(n_active, n_points);
  return p.chunks * (int64_t)(num_dims + 1) * n_active * (int64_t)sizeof(double);
}

int cb200_ms_brute_accumulate(const double* points, int64_t n_points, int64_t pts_stride, int num_dims,
                              const double* means, int64_t seed_stride, const int* active, int64_t n_active,
                              double bandwidth, void* partial, void* stream) {
  if (!points || !means || !active || !partial || n_points <= 0 || n_active < 0) return CB200_EINVAL;
  if (n_active == 0) return CB200_OK;
  // bulk async copies move 16-byte granules: even stride and an aligned base
  if ((pts_stride & 1) || (reinterpret_cast<uintptr_t>(points) & 15) || pts_stride < n_points) return CB200_EINVAL;
  const BrutePlan p = brute_plan(n_active, n_points);
  if (p.chunks > 65535) return CB200_EUNSUPPORTED;
  dim3 grid((unsigned)p.seed_tiles, (unsigned)p.chunks);
  cudaStream_t st = (cudaStream_t)stream;
  const double r2 = bandwidth * bandwidth;
  if (num_dims == 2)
    ms_brute_accumulate_kernel<2><<<grid, MSB_THREADS, 0, st>>>(points, n_points, pts_stride, means, seed_stride, active,
                                                                n_active, r2, p.chunk_points, (double*)partial);
  else if (num_dims == 3)
    ms_brute_accumulate_kernel<3><<<grid, MSB_THREADS, 0, st>>>(points, n_points, pts_stride, means, seed_stride, active,
                                                                n_active, r2, p.chunk_points, (double*)partial);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_ms_update(double* means, int64_t seed_stride, int num_dims, int* counts, int* iters, const int* active,
                    int64_t n_active, int64_t n_points, const void* partial, double bandwidth, int max_iter,
                    int* next_active, int* n_next, void* stream) {
  if (!means || !counts || !iters || !active || !partial || !next_active || !n_next || n_active < 0)
    return CB200_EINVAL;
  if (n_active == 0) return CB200_OK;
  const BrutePlan p = brute_plan(n_active, n_points);
  const int blocks = (int)((n_active + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const double stop = 1e-3 * bandwidth;
  if (num_dims == 2)
    ms_update_kernel<2><<<blocks, 256, 0, st>>>(means, seed_stride, counts, iters, active, n_active, p.chunks,
                                                (const double*)partial, stop, max_iter, next_active, n_next);
  else if (num_dims == 3)
    ms_update_kernel<3><<<blocks, 256, 0, st>>>(means, seed_stride, counts, iters, active, n_active, p.chunks,
                                                (const double*)partial, stop, max_iter, next_active, n_next);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_grid_plan(const double* lo, const double* hi, int num_dims, double bandwidth, int64_t max_cells,
                    cb200_grid* grid) {
  if (!lo || !hi || !grid || (num_dims != 2 && num_dims != 3) || !(bandwidth > 0.0)) return CB200_EINVAL;
  if (max_cells <= 0) max_cells = (int64_t)1 << 26;
  // a hair above the bandwidth so that |x - m| <= bw always lands within +-1 cell after rounding
  double cell = bandwidth * (1.0 + 1e-6);
  for (int attempt = 0; attempt < 64; ++attempt) {
    int64_t n = 1;
    bool ok = true;
    for (int k = 0; k < 3; ++k) {
      grid->origin[k] = k < num_dims ? lo[k] : 0.0;
      int64_t d = 1;
      if (k < num_dims) {
        if (!(hi[k] >= lo[k])) return CB200_EINVAL;
        d = (int64_t)floor((hi[k] - lo[k]) / cell) + 1;
      }
      if (d > INT32_MAX) ok = false;
      grid->dims[k] = (int32_t)d;
      n *= d;
      if (n > max_cells) ok = false;
    }
    if (ok) {
      grid->cell = cell;
      grid->inv_cell = 1.0 / cell;
      grid->num_dims = num_dims;
      grid->n_cells = n;
      return CB200_OK;
    }
    cell *= 1.5;  // coarser cells are still correct (edge >= bandwidth), just less selective
  }
  return CB200_EINVAL;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int64_t cb200_grid_build_workspace_bytes(int64_t n_points, int64_t n_cells) {
  if (n_points <= 0) return 256;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n_points);
  size_t total = align_up(sort_bytes, 256);
  total += 2 * align_up(sizeof(unsigned) * (size_t)n_points, 256);  // keys in/out
  total += 2 * align_up(sizeof(int) * (size_t)n_points, 256);       // idx in/out
  (void)n_cells;
  return (int64_t)total + 256;
}

int cb200_grid_build(const double* points, int64_t n_points, int64_t pts_stride, const cb200_grid* grid,
                     double* points_sorted, int64_t sorted_stride, int* order, int* cell_start, void* workspace,
                     int64_t workspace_bytes, void* stream) {
  if (!points || !grid || !points_sorted || !cell_start || !workspace || n_points < 0) return CB200_EINVAL;
  if (n_points > INT32_MAX || grid->n_cells >= ((int64_t)1 << 32)) return CB200_EUNSUPPORTED;
  if (workspace_bytes < cb200_grid_build_workspace_bytes(n_points, grid->n_cells)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDev g = to_dev(*grid);
  const int D = grid->num_dims;
  if (n_points == 0) {
    CB200_CUDA_TRY(cudaMemsetAsync(cell_start, 0, sizeof(int) * (size_t)(grid->n_cells + 1), st));
    return CB200_OK;
  }
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n_points);
  char* w = static_cast<char*>(workspace);
  w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(w), 256));
  void* sort_ws = w;                 w += align_up(sort_bytes, 256);
  unsigned* keys_in = (unsigned*)w;  w += align_up(sizeof(unsigned) * (size_t)n_points, 256);
  unsigned* keys_out = (unsigned*)w; w += align_up(sizeof(unsigned) * (size_t)n_points, 256);
  int* idx_in = (int*)w;             w += align_up(sizeof(int) * (size_t)n_points, 256);
  int* idx_out = (int*)w;
  const int blocks = grid_for(n_points, 256, 2, 16);
  if (D == 2) grid_cell_ids_kernel<2><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, g, keys_in, idx_in);
  else grid_cell_ids_kernel<3><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, g, keys_in, idx_in);
  CB200_LAUNCH_CHECK();
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < grid->n_cells) ++bits;
  CB200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_ws, sort_bytes, keys_in, keys_out, idx_in, idx_out,
                                                 (int)n_points, 0, bits, st));
  if (D == 2) grid_gather_kernel<2><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, idx_out, points_sorted, sorted_stride);
  else grid_gather_kernel<3><<<blocks, 256, 0, st>>>(points, n_points, pts_stride, idx_out, points_sorted, sorted_stride);
  CB200_LAUNCH_CHECK();
  grid_cell_start_kernel<<<grid_for(grid->n_cells + 1, 256, 1, 16), 256, 0, st>>>(keys_out, n_points, grid->n_cells,
                                                                                  cell_start);
  CB200_LAUNCH_CHECK();
  if (order) CB200_CUDA_TRY(cudaMemcpyAsync(order, idx_out, sizeof(int) * (size_t)n_points, cudaMemcpyDeviceToDevice, st));
  return CB200_OK;
}

int cb200_ms_grid_modes(const double* points_sorted, int64_t n_points, int64_t sorted_stride, const cb200_grid* grid,
                        const int* cell_start, double* means, int64_t seed_stride, int64_t n_seeds, double bandwidth,
                        int max_iter, int* counts, int* iters, int* work_counter, void* stream) {
  return ms_grid_modes_launch(points_sorted, n_points, sorted_stride, grid, cell_start, means, seed_stride, n_seeds, bandwidth,
                              max_iter, counts, iters, work_counter, nullptr, nullptr, 0, (cudaStream_t)stream);
}

}  // extern "C"
