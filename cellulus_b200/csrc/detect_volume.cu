// cb200_detect_volume: the whole per-bandwidth detect sequence behind ONE call.
//
// The step-by-step entry points (cb200_fg_compact ... cb200_assign_labels) neither allocate nor synchronise;
// a caller that drives them from an interpreter pays for ~11 calls, ~30 small allocations and 4 blocking count
// reads per volume, which costs more wall time than the ~45 small kernels between the big ones.  This
// composite runs the identical sequence from C++: scratch comes from a library-owned arena that is kept
// between calls, the four data-dependent counts (foreground, fit subset, bounding box, centres) are read
// with a stream synchronise each, everything else is enqueued back to back.  Same kernels, same results.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cb200 {

// Library-owned scratch: a per-device list of cudaMalloc'ed chunks that is bump-allocated per call and kept
// for the next one (cudaMalloc only while the working set is still growing -- the same idea as a caching
// allocator; cudaMallocAsync was measured 2-4x slower here because blocks freed a few kernels ago are not yet
// reusable and the pool keeps mapping new memory).  Calls are serialised per device; a call on another
// stream first waits for the previous call's last kernel.
struct Chunk {
  char* base;
  size_t size, used;
};
struct DeviceArena {
  std::mutex lock;
  std::vector<Chunk> chunks;
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used_before = false;
};
static DeviceArena g_arena[64];

struct ArenaScope {  // one call's view of the arena
  DeviceArena& a;
  cudaStream_t st;
  std::unique_lock<std::mutex> guard;
  ArenaScope(DeviceArena& arena, cudaStream_t s) : a(arena), st(s), guard(arena.lock) {
    for (Chunk& c : a.chunks) c.used = 0;
    if (a.used_before && a.last_stream != st) cudaStreamWaitEvent(st, a.last_use, 0);
  }
  ~ArenaScope() {
    if (!a.last_use) cudaEventCreateWithFlags(&a.last_use, cudaEventDisableTiming);
    cudaEventRecord(a.last_use, st);
    a.last_stream = st;
    a.used_before = true;
  }
  template <typename T>
  cudaError_t get(T** out, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) / 256 * 256;
    for (Chunk& c : a.chunks) {
      if (c.size - c.used >= bytes) {
        *out = reinterpret_cast<T*>(c.base + c.used);
        c.used += bytes;
        return cudaSuccess;
      }
    }
    const size_t want = bytes > ((size_t)64 << 20) ? bytes : ((size_t)64 << 20);
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    a.chunks.push_back(Chunk{static_cast<char*>(p), want, bytes});
    *out = static_cast<T*>(p);
    return cudaSuccess;
  }
};

}  // namespace cb200

using namespace cb200;

extern "C" int cb200_release_scratch(void) {
  int device = 0;
  CB200_CUDA_TRY(cudaGetDevice(&device));
  if (device < 0 || device >= 64) return CB200_EUNSUPPORTED;
  DeviceArena& a = g_arena[device];
  std::lock_guard<std::mutex> guard(a.lock);
  CB200_CUDA_TRY(cudaDeviceSynchronize());  // the last call's kernels may still be reading the arena
  for (Chunk& c : a.chunks) cudaFree(c.base);
  a.chunks.clear();
  return CB200_OK;
}

#define CB200_TRY_RC(expr)        \
  do {                            \
    const int _rc = (expr);       \
    if (_rc != CB200_OK) return _rc; \
  } while (0)

extern "C" int cb200_detect_volume(const void* emb, int dtype, int num_dims, const int64_t* spatial, double threshold,
                                   double bandwidth, double reduction_probability, uint64_t philox_seed, int max_iter,
                                   void* labels_out, int label_dtype, void* mask_out, int mask_dtype,
                                   double* centres_out, int64_t centre_capacity, cb200_detect_info* info,
                                   void* stream) {
  if (!emb || !spatial || !labels_out || !info || !(bandwidth > 0.0)) return CB200_EINVAL;
  if (num_dims != 2 && num_dims != 3) return CB200_EUNSUPPORTED;
  if (label_dtype != CB200_I32 && label_dtype != CB200_U16) return CB200_EUNSUPPORTED;
  const int D = num_dims;
  int64_t n_pix = 1;
  for (int k = 0; k < D; ++k) {
    if (spatial[k] <= 0) return CB200_EINVAL;
    n_pix *= spatial[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  *info = cb200_detect_info{};
  static const bool trace = getenv("CB200_DETECT_TRACE") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[detect_volume] %-10s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - t_prev).count());
    t_prev = now;
  };
  int device = 0;
  CB200_CUDA_TRY(cudaGetDevice(&device));
  if (device < 0 || device >= 64) return CB200_EUNSUPPORTED;
  ArenaScope pool(g_arena[device], st);
  CB200_CUDA_TRY(cudaMemsetAsync(labels_out, 0, (size_t)n_pix * (label_dtype == CB200_I32 ? 4 : 2), st));

  // ---- foreground points (utils/mean_shift.py:15-36,85,94)
  const int64_t cap = (n_pix + 1) & ~(int64_t)1;  // even stride (16-byte granules)
  double* pts;
  int32_t* pix;
  long long* counts_dev;  // [0] foreground, [1] fit subset
  uint8_t* compact_ws;
  CB200_CUDA_TRY(pool.get(&pts, (size_t)D * cap));
  CB200_CUDA_TRY(pool.get(&pix, (size_t)cap));
  CB200_CUDA_TRY(pool.get(&counts_dev, 2));
  CB200_CUDA_TRY(pool.get(&compact_ws, (size_t)cb200_compact_workspace_bytes(n_pix)));
  CB200_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(long long), st));
  CB200_TRY_RC(cb200_fg_compact(emb, dtype, D, spatial, threshold, pts, pix, cap, counts_dev, mask_out, mask_dtype,
                                compact_ws, st));
  long long n = 0;
  CB200_CUDA_TRY(cudaMemcpyAsync(&n, counts_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CB200_CUDA_TRY(cudaStreamSynchronize(st));
  info->n_foreground = n;
  lap("compact");
  if (n == 0) return CB200_OK;  // all background (utils/mean_shift.py:83-84)

  // ---- fit subset (utils/mean_shift.py:67-70), Bernoulli flags from the device Philox stream
  const double* fit = pts;
  int64_t fit_stride = cap;
  long long n_fit = n;
  if (reduction_probability < 1.0) {
    uint8_t* flags;
    double* subset;
    const int64_t sub_cap = (n + 1) & ~(int64_t)1;
    CB200_CUDA_TRY(pool.get(&flags, (size_t)n));
    CB200_CUDA_TRY(pool.get(&subset, (size_t)D * sub_cap));
    CB200_TRY_RC(cb200_bernoulli_flags(flags, n, reduction_probability, philox_seed, st));
    CB200_TRY_RC(cb200_select_points(pts, n, cap, D, flags, subset, sub_cap, counts_dev + 1, compact_ws, st));
    CB200_CUDA_TRY(cudaMemcpyAsync(&n_fit, counts_dev + 1, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CB200_CUDA_TRY(cudaStreamSynchronize(st));
    fit = subset;
    fit_stride = sub_cap;
  }
  info->n_fit = n_fit;
  lap("subset");
  if (n_fit == 0) return CB200_ENOFIT;  // sklearn: "Found array with 0 sample(s)"

  // ---- bounding box of the fit points -> cell grid
  double* box_dev;
  uint8_t* reduce_ws;
  CB200_CUDA_TRY(pool.get(&box_dev, 6));
  CB200_CUDA_TRY(pool.get(&reduce_ws, (size_t)cb200_reduce_workspace_bytes()));
  CB200_CUDA_TRY(cudaMemsetAsync(reduce_ws, 0, (size_t)cb200_reduce_workspace_bytes(), st));
  for (int k = 0; k < D; ++k)
    CB200_TRY_RC(cb200_minmax(fit + (size_t)k * fit_stride, CB200_F64, n_fit, box_dev + 2 * k, reduce_ws, st));
  double box[6];
  CB200_CUDA_TRY(cudaMemcpyAsync(box, box_dev, sizeof(double) * 2 * D, cudaMemcpyDeviceToHost, st));
  CB200_CUDA_TRY(cudaStreamSynchronize(st));
  double lo[3], hi[3];
  for (int k = 0; k < D; ++k) {
    lo[k] = box[2 * k];
    hi[k] = box[2 * k + 1];
  }
  cb200_grid grid;
  CB200_TRY_RC(cb200_grid_plan(lo, hi, D, bandwidth, (int64_t)1 << 26, &grid));
  info->grid = grid;
  lap("bbox");

  // ---- grid hash of the fit points, every fit point climbs (sklearn:491-496, :108-128)
  const int64_t fit_cap = (n_fit + 1) & ~(int64_t)1;
  double *sorted, *modes;
  int *cell_start, *counts, *iters, *work;
  uint8_t* build_ws;
  const int64_t build_bytes = cb200_grid_build_workspace_bytes(n_fit, grid.n_cells);
  CB200_CUDA_TRY(pool.get(&sorted, (size_t)D * fit_cap));
  CB200_CUDA_TRY(pool.get(&modes, (size_t)D * fit_cap));
  CB200_CUDA_TRY(pool.get(&cell_start, (size_t)grid.n_cells + 1));
  CB200_CUDA_TRY(pool.get(&counts, (size_t)n_fit));
  CB200_CUDA_TRY(pool.get(&iters, (size_t)n_fit));
  CB200_CUDA_TRY(pool.get(&work, 8));
  CB200_CUDA_TRY(pool.get(&build_ws, (size_t)build_bytes));
  CB200_TRY_RC(cb200_grid_build(fit, n_fit, fit_stride, &grid, sorted, fit_cap, nullptr, cell_start, build_ws,
                                build_bytes, st));
  for (int k = 0; k < D; ++k)
    CB200_CUDA_TRY(cudaMemcpyAsync(modes + (size_t)k * fit_cap, fit + (size_t)k * fit_stride, sizeof(double) * n_fit,
                                   cudaMemcpyDeviceToDevice, st));
  CB200_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * n_fit, st));
  CB200_CUDA_TRY(cudaMemsetAsync(iters, 0, sizeof(int) * n_fit, st));
  CB200_CUDA_TRY(cudaMemsetAsync(work, 0, 8 * sizeof(int), st));
  CB200_TRY_RC(cb200_ms_grid_modes(sorted, n_fit, fit_cap, &grid, cell_start, modes, fit_cap, n_fit, bandwidth,
                                   max_iter > 0 ? max_iter : 300, counts, iters, work, st));
  info->n_seeds = n_fit;
  lap("modes");

  // ---- centres: dedupe + greedy suppression (sklearn:511-547)
  const int64_t nms_bytes = cb200_nms_workspace_bytes(n_fit, &grid, bandwidth);
  if (nms_bytes < 0) return CB200_EUNSUPPORTED;
  uint8_t* nms_ws;
  int* keep_dev;
  CB200_CUDA_TRY(pool.get(&nms_ws, (size_t)nms_bytes));
  CB200_CUDA_TRY(pool.get(&keep_dev, 2));
  CB200_CUDA_TRY(cudaMemsetAsync(keep_dev, 0, 2 * sizeof(int), st));
  int keep[2] = {0, 1};
  long long stats[2] = {0, 0};  // the hill climb's work statistics ride on the first count read
  CB200_CUDA_TRY(cudaMemcpyAsync(stats, work + 2, sizeof(stats), cudaMemcpyDeviceToHost, st));
  for (int call = 0; call < 64 && keep[1] != 0; ++call) {
    CB200_TRY_RC(cb200_nms_suppress(modes, fit_cap, D, counts, n_fit, bandwidth, &grid, 4, call > 0, keep_dev, nms_ws,
                                    nms_bytes, st));
    CB200_CUDA_TRY(cudaMemcpyAsync(keep, keep_dev, sizeof(keep), cudaMemcpyDeviceToHost, st));
    CB200_CUDA_TRY(cudaStreamSynchronize(st));
    ++info->suppress_calls;
  }
  lap("suppress");
  info->distance_tests = stats[0];
  info->climb_steps = stats[1];
  if (keep[1] != 0) return CB200_ENOCONVERGE;
  const int k_centres = keep[0];
  info->n_centres = k_centres;
  if (k_centres == 0) return CB200_ENOCENTRE;  // sklearn: "No point was within bandwidth ... of any seed"
  const int64_t c_cap = ((int64_t)k_centres + 1) & ~(int64_t)1;
  double* centres;
  int64_t c_stride = c_cap;
  if (centres_out && centre_capacity >= k_centres) {
    centres = centres_out;
    c_stride = centre_capacity;
  } else {
    CB200_CUDA_TRY(pool.get(&centres, (size_t)D * c_cap));
  }
  CB200_TRY_RC(cb200_nms_emit(modes, fit_cap, D, counts, n_fit, bandwidth, &grid, k_centres, centres, c_stride, nms_ws,
                              nms_bytes, st));

  // ---- predict on ALL foreground points, scatter, +1 (utils/mean_shift.py:74,101-104,57)
  uint8_t* assign_ws;
  CB200_CUDA_TRY(pool.get(&assign_ws, (size_t)cb200_assign_workspace_bytes(n, k_centres, grid.n_cells)));
  CB200_TRY_RC(cb200_assign_labels(pts, n, cap, D, centres, c_stride, k_centres, &grid, pix, labels_out, label_dtype,
                                   assign_ws, st));
  lap("assign");
  return CB200_OK;
}
