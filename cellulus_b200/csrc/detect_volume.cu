// cb200_detect_volume: the whole per-bandwidth detect sequence behind ONE call.
//
// The step-by-step entry points (cb200_fg_compact ... cb200_assign_labels) neither allocate nor synchronise;
// a caller that drives them from an interpreter pays for ~11 calls, ~30 small allocations and 4 blocking count
// reads per volume, which costs more wall time than the ~45 small kernels between the big ones.  This
// composite runs the identical sequence from C++ out of ONE caller-provided workspace.  The data-dependent counts
// are read with THREE stream synchronisations: foreground count + fit count + bounding box together (the subset and
// box kernels are enqueued before the counts are known, sized for the capacity, and read the counts on the
// device), the number of distinct modes, the number of centres; everything else is enqueued back to back.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace cb200 {

// Scratch of one call: bump-allocated from the CALLER's workspace (no cudaMalloc in here).  The sizes depend on
// counts that are only known on the device, so a call that runs out of room stops, reports the bytes the counts it
// has read so far call for (info->workspace_needed) and returns CB200_ENOSPACE; the caller grows its buffer and
// calls again (the torch shim keeps one buffer per device and does this once per new high-water mark).
struct Bump {
  char* base;
  size_t size, used = 0, wanted = 0;
  Bump(void* p, int64_t n) : base(static_cast<char*>(p)), size(n > 0 ? (size_t)n : 0) {
    const size_t mis = reinterpret_cast<uintptr_t>(base) & 255;
    if (mis) {
      const size_t skip = 256 - mis;
      base += skip;
      size = size > skip ? size - skip : 0;
    }
  }
  template <typename T>
  bool get(T** out, size_t count) {
    const size_t bytes = ((count ? count : 1) * sizeof(T) + 255) / 256 * 256;
    wanted += bytes;
    if (used + bytes > size) {
      *out = nullptr;
      return false;
    }
    *out = reinterpret_cast<T*>(base + used);
    used += bytes;
    return true;
  }
};

// bytes the sequence needs for n_pix pixels of which n_fg are foreground and n_fit take part in the fit
static int64_t detect_bytes(int D, int64_t n_pix, int64_t n_fg, int64_t n_fit, int64_t n_cells, int64_t nms_bytes) {
  auto al = [](int64_t b) { return (b + 255) / 256 * 256 + 256; };
  int64_t t = 512;
  t += al((int64_t)D * 8 * ((n_fg + 1) & ~(int64_t)1)) + al(4 * n_fg) + al(16) + al(cb200_compact_workspace_bytes(n_pix));
  t += al(n_fg) + al((int64_t)D * 8 * ((n_fit + 4096 + 1) & ~(int64_t)1));                // flags, fit subset (with headroom)
  t += al(48) + al(cb200_reduce_workspace_bytes());                                          // bounding box
  t += 2 * al((int64_t)D * 8 * ((n_fit + 1) & ~(int64_t)1)) + al(4 * (n_cells + 1)) + 2 * al(4 * n_fit) + al(32);
  t += al(cb200_grid_build_workspace_bytes(n_fit, n_cells)) + al(cb200_ms_distinct_workspace_bytes(n_fit)) + al(64);
  t += al(nms_bytes) + al(8) + al((int64_t)D * 8 * ((n_fit + 1) & ~(int64_t)1));            // suppression, centres
  t += al((int64_t)D * 8 * ((n_fit + 1) & ~(int64_t)1)) + al(4 * n_fit) + al(8) + al(cb200_unique_modes_workspace_bytes(n_fit));  // distinct modes
  t += al(cb200_assign_workspace_bytes(n_fg, (int)std::min<int64_t>(n_fit, INT32_MAX), n_cells));
  return t;
}

}  // namespace cb200

using namespace cb200;

extern "C" int64_t cb200_detect_volume_workspace_bytes(int num_dims, const int64_t* spatial, int64_t expected_foreground,
                                                       double reduction_probability) {
  if (!spatial || (num_dims != 2 && num_dims != 3)) return -1;
  int64_t n_pix = 1;
  for (int k = 0; k < num_dims; ++k) n_pix *= spatial[k] > 0 ? spatial[k] : 1;
  const int64_t n_fg = expected_foreground > 0 && expected_foreground < n_pix ? expected_foreground : n_pix;
  const double p = reduction_probability < 1.0 ? reduction_probability : 1.0;
  const int64_t n_fit = std::min<int64_t>(n_fg, (int64_t)(1.1 * p * (double)n_fg) + 4096);
  const int64_t n_cells = std::min<int64_t>((int64_t)1 << 26, 4 * n_pix + 64);  // cells of edge >= bandwidth over the points
  // suppression: a fine grid (8 cells per cell in 3-D) + a few arrays over the seeds; bounded as in cb200_nms_workspace_bytes
  const int64_t nms = 64 * n_fit + 8 * 8 * std::min<int64_t>(n_cells, n_fit * 27 + 64) + 4096;
  return detect_bytes(num_dims, n_pix, n_fg, n_fit, std::min<int64_t>(n_cells, n_fit * 27 + 64), nms);
}

// out of workspace: report what the counts known so far call for and stop
#define POOL_GET(ptr, count)                                                                        \
  do {                                                                                              \
    if (!pool.get(ptr, count)) {                                                                    \
      info->workspace_needed = detect_bytes(D, n_pix, known_fg, known_fit, known_cells, known_nms); \
      return CB200_ENOSPACE;                                                                        \
    }                                                                                               \
  } while (0)

#define CB200_TRY_RC(expr)        \
  do {                            \
    const int _rc = (expr);       \
    if (_rc != CB200_OK) return _rc; \
  } while (0)

extern "C" int cb200_detect_volume(const void* emb, int dtype, int num_dims, const int64_t* spatial, double threshold,
                                   double bandwidth, double reduction_probability, uint64_t philox_seed, int max_iter,
                                   void* labels_out, int label_dtype, void* mask_out, int mask_dtype,
                                   double* centres_out, int64_t centre_capacity, void* workspace,
                                   int64_t workspace_bytes, int64_t foreground_capacity, cb200_detect_info* info,
                                   void* stream) {
  if (!emb || !spatial || !labels_out || !info || !workspace || !(bandwidth > 0.0)) return CB200_EINVAL;
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
  Bump pool(workspace, workspace_bytes);
  // what the workspace must hold, refined as the counts become known (for the CB200_ENOSPACE report)
  int64_t known_fg = foreground_capacity > 0 ? foreground_capacity : n_pix;
  int64_t known_fit = reduction_probability < 1.0 ? (int64_t)(1.1 * reduction_probability * (double)known_fg) + 4096 : known_fg;
  int64_t known_cells = 1 << 16, known_nms = 64 * known_fit + (1 << 20);
  CB200_CUDA_TRY(cudaMemsetAsync(labels_out, 0, (size_t)n_pix * (label_dtype == CB200_I32 ? 4 : 2), st));

  // ---- foreground points (utils/mean_shift.py:15-36,85,94)
  // room for `foreground_capacity` points (<= 0: every pixel); even stride (16-byte granules)
  const int64_t cap = ((foreground_capacity > 0 && foreground_capacity < n_pix ? foreground_capacity : n_pix) + 1) & ~(int64_t)1;
  double* pts;
  int32_t* pix;
  long long* counts_dev;  // [0] foreground, [1] fit subset
  uint8_t* compact_ws;
  POOL_GET(&pts, (size_t)D * cap);
  POOL_GET(&pix, (size_t)cap);
  POOL_GET(&counts_dev, 2);
  POOL_GET(&compact_ws, (size_t)cb200_compact_workspace_bytes(n_pix));
  CB200_CUDA_TRY(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(long long), st));
  CB200_TRY_RC(cb200_fg_compact(emb, dtype, D, spatial, threshold, pts, pix, cap, counts_dev, mask_out, mask_dtype,
                                compact_ws, st));

  // ---- fit subset (utils/mean_shift.py:67-70; Bernoulli flags from the device Philox stream) and its bounding box,
  // enqueued BEFORE the foreground count is known on the host: the kernels are sized for the capacity and read the
  // counts where they are (the flag of point i depends on i only).  ONE blocking read then returns the foreground
  // count, the fit count and the box together.
  const bool subsample = reduction_probability < 1.0;
  const double* fit = pts;
  int64_t fit_stride = cap;
  int64_t sub_cap = 0;
  const long long* n_fit_dev = counts_dev;  // all foreground points take part
  if (subsample) {
    uint8_t* flags;
    double* subset;
    sub_cap = (std::min<int64_t>(cap, (int64_t)(1.1 * reduction_probability * (double)cap) + 4096) + 1) & ~(int64_t)1;
    POOL_GET(&flags, (size_t)cap);
    POOL_GET(&subset, (size_t)D * sub_cap);
    CB200_TRY_RC(cb200_bernoulli_flags(flags, cap, reduction_probability, philox_seed, st));
    CB200_TRY_RC(select_points_counted(pts, cap, counts_dev, cap, D, flags, subset, sub_cap, counts_dev + 1, compact_ws, st));
    fit = subset;
    fit_stride = sub_cap;
    n_fit_dev = counts_dev + 1;
  }
  double* box_dev;
  uint8_t* reduce_ws;
  POOL_GET(&box_dev, 6);
  POOL_GET(&reduce_ws, (size_t)cb200_reduce_workspace_bytes());
  CB200_CUDA_TRY(cudaMemsetAsync(reduce_ws, 0, (size_t)cb200_reduce_workspace_bytes(), st));
  for (int k = 0; k < D; ++k)
    CB200_TRY_RC(minmax_counted(fit + (size_t)k * fit_stride, CB200_F64, subsample ? sub_cap : cap, n_fit_dev,
                                box_dev + 2 * k, reduce_ws, st));
  struct {
    long long counts[2];
    double box[6];
  } head;
  CB200_CUDA_TRY(cudaMemcpyAsync(head.counts, counts_dev, sizeof(head.counts), cudaMemcpyDeviceToHost, st));
  CB200_CUDA_TRY(cudaMemcpyAsync(head.box, box_dev, sizeof(double) * 2 * D, cudaMemcpyDeviceToHost, st));
  CB200_CUDA_TRY(cudaStreamSynchronize(st));
  const long long n = head.counts[0];
  const long long n_fit = subsample ? head.counts[1] : n;
  info->n_foreground = n;
  known_fg = n;
  known_fit = subsample ? std::max<int64_t>(n_fit, (int64_t)(1.1 * reduction_probability * (double)n) + 4096) : n;
  known_nms = 64 * known_fit + (1 << 20);
  lap("compact+subset+bbox");
  if (n > cap || (subsample && n_fit > sub_cap)) {  // more points than the caller made room for: nothing beyond was written
    info->workspace_needed = detect_bytes(D, n_pix, known_fg + (known_fg >> 3), known_fit, known_cells, known_nms);
    return CB200_ENOSPACE;
  }
  if (n == 0) return CB200_OK;  // all background (utils/mean_shift.py:83-84)
  info->n_fit = n_fit;
  known_fit = n_fit;
  known_nms = 64 * known_fit + (1 << 20);
  if (n_fit == 0) return CB200_ENOFIT;  // sklearn: "Found array with 0 sample(s)"

  // ---- cell grid over the bounding box of the fit points
  double lo[3], hi[3];
  for (int k = 0; k < D; ++k) {
    lo[k] = head.box[2 * k];
    hi[k] = head.box[2 * k + 1];
  }
  cb200_grid grid;
  CB200_TRY_RC(cb200_grid_plan(lo, hi, D, bandwidth, (int64_t)1 << 26, &grid));
  info->grid = grid;
  known_cells = grid.n_cells;
  lap("bbox");

  // ---- grid hash of the fit points, every fit point climbs (sklearn:491-496, :108-128)
  const int64_t fit_cap = (n_fit + 1) & ~(int64_t)1;
  double *sorted, *modes;
  int *cell_start, *counts, *iters, *work;
  uint8_t* build_ws;
  const int64_t build_bytes = cb200_grid_build_workspace_bytes(n_fit, grid.n_cells);
  POOL_GET(&sorted, (size_t)D * fit_cap);
  POOL_GET(&modes, (size_t)D * fit_cap);
  POOL_GET(&cell_start, (size_t)grid.n_cells + 1);
  POOL_GET(&counts, (size_t)n_fit);
  POOL_GET(&iters, (size_t)n_fit);
  POOL_GET(&work, 16);
  POOL_GET(&build_ws, (size_t)build_bytes);
  CB200_TRY_RC(cb200_grid_build(fit, n_fit, fit_stride, &grid, sorted, fit_cap, nullptr, cell_start, build_ws,
                                build_bytes, st));
  for (int k = 0; k < D; ++k)
    CB200_CUDA_TRY(cudaMemcpyAsync(modes + (size_t)k * fit_cap, fit + (size_t)k * fit_stride, sizeof(double) * n_fit,
                                   cudaMemcpyDeviceToDevice, st));
  CB200_CUDA_TRY(cudaMemsetAsync(work, 0, 16 * sizeof(int), st));
  // CB200_DETECT_CLIMB=all: every seed climbs to convergence on its own (A/B switch).  Default: one window evaluation
  // per seed, then one representative of every distinct unfinished mean (cb200_ms_grid_modes_distinct) -- the centres
  // only need the distinct modes, and copies of a trajectory end in copies of its mode.
  static const bool climb_all = [] { const char* e = getenv("CB200_DETECT_CLIMB"); return e && e[0] == 'a'; }();
  if (climb_all || n_fit < 20000) {  // the merge pays from a few ten thousand seeds on
    CB200_TRY_RC(cb200_ms_grid_modes(sorted, n_fit, fit_cap, &grid, cell_start, modes, fit_cap, n_fit, bandwidth,
                                     max_iter > 0 ? max_iter : 300, counts, iters, work, st));
  } else {
    uint8_t* distinct_ws;
    const int64_t distinct_bytes = cb200_ms_distinct_workspace_bytes(n_fit);
    POOL_GET(&distinct_ws, (size_t)distinct_bytes);
    // one merge for ordinary seed counts; millions of dense seeds keep meeting as they climb (CB200_MS_ROUNDS overrides)
    static const int rounds_env = [] { const char* e = getenv("CB200_MS_ROUNDS"); return e ? atoi(e) : 0; }();
    const int merge_rounds = rounds_env >= 1 && rounds_env <= 30 ? rounds_env : (n_fit < 1000000 ? 1 : 4);
    CB200_TRY_RC(cb200_ms_grid_modes_distinct(sorted, n_fit, fit_cap, &grid, cell_start, modes, fit_cap, n_fit, bandwidth,
                                              max_iter > 0 ? max_iter : 300, merge_rounds, counts, iters, work, distinct_ws,
                                              distinct_bytes, st));
  }
  info->n_seeds = n_fit;
  lap("modes");

  // ---- centres: merge bit-identical modes, then greedy suppression (sklearn:511-547)
  long long stats[6] = {0, 0, 0, 0, 0, 0};  // the hill climb's work statistics ride on the first count read:
  CB200_CUDA_TRY(cudaMemcpyAsync(stats, work + 2, sizeof(stats), cudaMemcpyDeviceToHost, st));  // tests, steps, -, -, tests, steps
  // CB200_DETECT_DEDUPE=0: suppression over all converged seeds (A/B switch).  Default: seeds that end in the same
  // window end in the SAME mean, so one exact pass leaves hundreds of candidates out of hundreds of thousands.
  static const bool dedupe = [] { const char* e = getenv("CB200_DETECT_DEDUPE"); return !(e && e[0] == '0'); }();
  const double* cand = modes;
  const int* cand_counts = counts;
  long long n_cand = n_fit;
  if (dedupe && n_fit >= 4096) {
    double* umodes;
    int* ucounts;
    long long* n_unique_dev;
    uint8_t* unique_ws;
    const int64_t unique_bytes = cb200_unique_modes_workspace_bytes(n_fit);
    POOL_GET(&umodes, (size_t)D * fit_cap);
    POOL_GET(&ucounts, (size_t)n_fit);
    POOL_GET(&n_unique_dev, 1);
    POOL_GET(&unique_ws, (size_t)unique_bytes);
    CB200_TRY_RC(cb200_unique_modes(modes, fit_cap, D, counts, n_fit, umodes, fit_cap, ucounts, n_unique_dev, unique_ws,
                                    unique_bytes, st));
    CB200_CUDA_TRY(cudaMemcpyAsync(&n_cand, n_unique_dev, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CB200_CUDA_TRY(cudaStreamSynchronize(st));
    cand = umodes;
    cand_counts = ucounts;
    lap("unique");
  }
  info->n_distinct_modes = n_cand;
  int keep[2] = {0, n_cand > 0 ? 1 : 0};
  uint8_t* nms_ws = nullptr;
  int64_t nms_bytes = 0;
  if (n_cand > 0) {
    nms_bytes = cb200_nms_workspace_bytes(n_cand, &grid, bandwidth);
    if (nms_bytes < 0) return CB200_EUNSUPPORTED;
    known_nms = nms_bytes;
    int* keep_dev;
    POOL_GET(&nms_ws, (size_t)nms_bytes);
    POOL_GET(&keep_dev, 2);
    CB200_CUDA_TRY(cudaMemsetAsync(keep_dev, 0, 2 * sizeof(int), st));
    for (int call = 0; call < 64 && keep[1] != 0; ++call) {
      CB200_TRY_RC(cb200_nms_suppress(cand, fit_cap, D, cand_counts, n_cand, bandwidth, &grid, 4, call > 0, keep_dev,
                                      nms_ws, nms_bytes, st));
      CB200_CUDA_TRY(cudaMemcpyAsync(keep, keep_dev, sizeof(keep), cudaMemcpyDeviceToHost, st));
      CB200_CUDA_TRY(cudaStreamSynchronize(st));
      ++info->suppress_calls;
    }
  } else {
    CB200_CUDA_TRY(cudaStreamSynchronize(st));  // the statistics copy
  }
  lap("suppress");
  info->distance_tests = stats[0] + stats[4];
  info->climb_steps = stats[1] + stats[5];
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
    POOL_GET(&centres, (size_t)D * c_cap);
  }
  CB200_TRY_RC(cb200_nms_emit(cand, fit_cap, D, cand_counts, n_cand, bandwidth, &grid, k_centres, centres, c_stride,
                              nms_ws, nms_bytes, st));

  // ---- predict on ALL foreground points, scatter, +1 (utils/mean_shift.py:74,101-104,57)
  uint8_t* assign_ws;
  POOL_GET(&assign_ws, (size_t)cb200_assign_workspace_bytes(n, k_centres, grid.n_cells));
  CB200_TRY_RC(cb200_assign_labels(pts, n, cap, D, centres, c_stride, k_centres, &grid, pix, labels_out, label_dtype,
                                   assign_ws, st));
  lap("assign");
  return CB200_OK;
}
