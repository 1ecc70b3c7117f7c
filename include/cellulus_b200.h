/*
 * cellulus_b200 -- C ABI of the B200-native embedding-space hot path.
 *
 * One shared object (cellulus_b200/libcellulus_b200.so), plain C symbols, raw
 * DEVICE pointers + explicit sizes + a cudaStream_t passed as void*.  No torch
 * types.  The reference (funkelab/cellulus) is pure Python and has no FFI of
 * its own, so each entry point cites the reference function it replaces
 * (paths relative to the reference root; "sklearn:" = scikit-learn 1.9.0
 * sklearn/cluster/_mean_shift.py, to which the reference delegates).
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - return value: 0 = ok; CB200_EINVAL (-1) bad argument; CB200_EUNSUPPORTED
 *     (-2) unsupported dtype/dims; otherwise a positive cudaError_t.
 *   - nothing here allocates device memory: every scratch buffer is passed in
 *     (sizes from the *_workspace_bytes queries).  No entry synchronises the
 *     stream EXCEPT the one composite whose control flow depends on counts
 *     that only exist on the device, and which says so at its declaration:
 *     cb200_detect_volume (three blocking count reads per call).
 *   - re-entrant across streams as long as workspaces are not shared.
 *   - the staged loss entries (cb200_oce_loss_*_staged with staging scratch)
 *     launch ONE resident wave whose thread blocks wait for each other inside
 *     the launch; everything else is ordinary fire-and-forget.
 *   - spatial shapes are passed in tensor-axis order ([z,] y, x); coordinate /
 *     channel columns are in the reference's (x, y[, z]) order: column 0
 *     indexes the LAST tensor axis (models/unet.py:114-118).
 */
#ifndef CELLULUS_B200_H
#define CELLULUS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CB200_API __attribute__((visibility("default")))
#else
#define CB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_OK 0
#define CB200_EINVAL (-1)
#define CB200_EUNSUPPORTED (-2)
#define CB200_ENOFIT (-3)      /* the fit subset is empty (sklearn: "Found array with 0 sample(s)") */
#define CB200_ENOCENTRE (-4)   /* no seed had a neighbour (sklearn: "No point was within bandwidth ...") */
#define CB200_ENOCONVERGE (-5) /* centre suppression did not reach its fix-point */
#define CB200_ENOSPACE (-6)    /* the workspace is too small for the counts found on the device; see the call */

/* element types */
#define CB200_F32 0
#define CB200_BF16 1
#define CB200_F64 2
#define CB200_I64 3
#define CB200_I32 4
#define CB200_I16 5
#define CB200_U8 6
#define CB200_U16 7

/* memory layout of a (B, D, *spatial) tensor */
#define CB200_LAYOUT_PLANAR 0
#define CB200_LAYOUT_CHANNELS_LAST 1

/* library / build information: returns the sm arch the kernels were built for (100) */
CB200_API int cb200_version(int* major, int* minor, int* sm_arch);
/* last CUDA error string for a positive return code */
CB200_API const char* cb200_error_string(int code);

/* Measurement aid (bench.py): runs blocks_per_sm * 148 blocks of 256 threads, 8 independent FMA chains per
 * thread, `iters` rounds, in CB200_F32 or CB200_F64; *flop (host out) = the floating-point operations issued.
 * Timed with CUDA events by the caller it gives the device's FMA peak, the roofline denominator of the
 * mean-shift kernels.  `out`: 8 bytes of device scratch. */
CB200_API int cb200_fma_peak(int dtype, int iters, int blocks_per_sm, void* out, int64_t* flop /* host out */,
                   void* stream);

/* ===================================================================== *
 *  Loss slice (training)                                                *
 * ===================================================================== */

/* bytes of scratch needed by cb200_oce_loss_fwd_bwd (accumulators, ticket, arrival counters) */
CB200_API int64_t cb200_oce_loss_workspace_bytes(void);
/* bytes of OPTIONAL staging scratch for cb200_oce_loss_fwd_bwd_staged: non-zero for 2-D CB200_LAYOUT_PLANAR
 * offsets (one copy of the offsets tensor), 0 where staging is not used (channels-last, 3-D) */
CB200_API int64_t cb200_oce_loss_staging_bytes(int offsets_dtype, int offsets_layout, int batch, int num_dims,
                                     const int64_t* spatial /* host, num_dims */);

/*
 * Fused neighbour gather + OCE loss forward + backward in ONE pass.
 * Replaces, in cellulus/train.py:169-178,
 *     select_and_add_coordinates(offsets, anchors)   models/unet.py:108-124
 *     select_and_add_coordinates(offsets, refs)      models/unet.py:108-124
 *     OCELoss.forward(ea, er)                        criterions/oce_loss.py:53-63
 *     loss.backward()  (gather backward = scatter-add onto the anchor pixel)
 *
 *  offsets      CB200_F32 or CB200_BF16; offsets_layout says how the (B, D, *spatial) tensor lies in memory:
 *               CB200_LAYOUT_PLANAR = contiguous NCHW / NCDHW (what the reference's model emits),
 *               CB200_LAYOUT_CHANNELS_LAST = contiguous (B, *spatial, D) (torch channels_last): one
 *               vector load fetches a whole embedding, which halves the gather traffic
 *  anchors/refs (B, P, D) contiguous, CB200_I64 (as the reference delivers
 *               them), CB200_I32 or CB200_I16; columns (x, y[, z])
 *  grad         fp32, same layout as offsets, OVERWRITTEN with d loss / d offsets
 *               (zero-filled by this call, then accumulated); may be NULL to
 *               run the forward only
 *  out          4 floats: loss, oce_loss, regularization_loss, number of
 *               pairs skipped because a coordinate was out of range
 *  workspace    cb200_oce_loss_workspace_bytes() bytes, zero-initialised ONCE
 *               by the caller (the kernel leaves it zeroed again)
 */
CB200_API int cb200_oce_loss_fwd_bwd(const void* offsets, int offsets_dtype, int offsets_layout,
                           const void* anchors, const void* refs, int coord_dtype,
                           int batch, int num_dims, const int64_t* spatial /* host, num_dims */,
                           int64_t pairs_per_sample,
                           float temperature, float regularization_weight,
                           float* grad, float* out, void* workspace, void* stream);
/*
 * Same call with caller-provided staging scratch (`staging_bytes` >= cb200_oce_loss_staging_bytes(...), 16-byte
 * aligned; contents undefined before and after).  For planar 2-D offsets -- what the reference's model emits,
 * models/unet.py:69-71 -- the kernel first writes a channels-last copy of the offsets into the scratch (its
 * thread blocks transpose their sample between them) and gathers from the copy, so that a scattered reference
 * pixel costs one memory transaction instead of one per channel; gradient and results are identical in layout
 * and value to cb200_oce_loss_fwd_bwd.  staging == NULL (or too small) behaves exactly like cb200_oce_loss_fwd_bwd.
 * The gradient is cleared inside the same launch: ONE kernel per call.
 */
CB200_API int cb200_oce_loss_fwd_bwd_staged(const void* offsets, int offsets_dtype, int offsets_layout,
                           const void* anchors, const void* refs, int coord_dtype,
                           int batch, int num_dims, const int64_t* spatial /* host, num_dims */,
                           int64_t pairs_per_sample,
                           float temperature, float regularization_weight,
                           float* grad, float* out, void* workspace,
                           void* staging, int64_t staging_bytes, void* stream);

/* grad *= *scale (device scalar); returns immediately on the device when *scale == 1.
 * Used by the autograd shim to apply the upstream gradient of `loss`. */
CB200_API int cb200_scale_inplace(float* grad, int64_t n, const float* scale, void* stream);

/*
 * Unfused drop-ins, for callers that keep the reference's three-call shape.
 *
 * cb200_gather_add_coords: models/unet.py:108-124 forward.
 *   out (B, P, D) fp32 = offsets[b, :, coord] + coord
 * cb200_scatter_add_coords: its backward; grad_offsets (B, D, *spatial) fp32
 *   is zero-filled then accumulated from grad_out (B, P, D).
 */
CB200_API int cb200_gather_add_coords(const void* offsets, int offsets_dtype, const void* coords, int coord_dtype,
                            int batch, int num_dims, const int64_t* spatial, int64_t pairs_per_sample,
                            float* out, void* stream);
CB200_API int cb200_scatter_add_coords(const float* grad_out, const void* coords, int coord_dtype,
                             int batch, int num_dims, const int64_t* spatial, int64_t pairs_per_sample,
                             float* grad_offsets, void* stream);
/*
 * cb200_oce_pair_loss: criterions/oce_loss.py:53-63 on materialised embeddings.
 *   ea, er (n_pairs, D) fp32; grad_ea (n_pairs, D) fp32 or NULL; out = 4 floats
 *   as above; workspace as for cb200_oce_loss_fwd_bwd.
 */
CB200_API int cb200_oce_pair_loss(const float* ea, const float* er, int64_t n_pairs, int num_dims,
                        float temperature, float regularization_weight,
                        float* grad_ea, float* out, void* workspace, void* stream);

/*
 * Device pair sampler (Philox4x32-10), the B200-native replacement of
 * ZarrDataset.sample_coordinates / sample_offsets_within_radius
 * (datasets/zarr_dataset.py:177-251): anchors uniform in [kappa, extent-kappa]
 * per column, each repeated num_references times consecutively; offsets
 * uniform over the integer points of the open ball sum o^2 < kappa^2 minus
 * the origin (one bounded index into the table of admissible offsets -- the
 * distribution the reference's rejection filter produces).  Same distribution
 * as the reference, a counter-based stream instead of numpy's global generator;
 * the stream is specified in csrc/pair_stream.cuh and restated in numpy in
 * oracle/device_sampler.py.
 *  extent       host, num_dims ints in COLUMN order (x, y[, z]) -- the
 *               reference draws column d from output_shape[d] (quirk Q4)
 *  anchors/refs (B, num_anchors*num_references, D), CB200_I64 / I32 / I16
 *  kappa        trunc(kappa) <= 127 and the offset table must fit in shared memory
 *               (kappa <= 120 in 2-D, <= 21 in 3-D), else CB200_EUNSUPPORTED
 */
CB200_API int cb200_sample_pairs(void* anchors, void* refs, int coord_dtype, int batch, int num_dims,
                       const int64_t* extent, double kappa, int64_t num_anchors, int num_references,
                       uint64_t seed, uint64_t sequence, void* stream);

/*
 * The loss slice with the sampling INSIDE the kernel: draws the pairs of the stream (seed, sequence) --
 * exactly the lists cb200_sample_pairs would write -- gathers, evaluates OCELoss forward + backward and
 * never materialises a coordinate list.  Replaces, per training step, the DataLoader-side sampler
 * (datasets/zarr_dataset.py:198-242), its two host->device list copies (train.py:162-166) and
 * train.py:169-178.  One lane owns one anchor: its num_references pairs accumulate their gradient in
 * registers and issue a single reduction.  HBM traffic = offsets once + gradient once.
 *  offsets/grad/out/workspace  as for cb200_oce_loss_fwd_bwd
 *  spatial      host, tensor-axis order ([z,] y, x);  extent: host, sampling extents in COLUMN order
 *               (normally the reversed `spatial`; the reference's quirk Q4 feeds output_shape unreversed).
 *               Pairs that fall outside the tensor are skipped and counted in out[3].
 *  dump_anchors/dump_refs  optional (both or neither): (B, num_anchors*num_references, D) lists of
 *               dump_dtype (CB200_I64 / I32 / I16) receiving the pairs this call used (test / debug)
 */
CB200_API int cb200_oce_loss_sampled(const void* offsets, int offsets_dtype, int offsets_layout, int batch, int num_dims,
                           const int64_t* spatial, const int64_t* extent, double kappa, int64_t num_anchors,
                           int num_references, uint64_t seed, uint64_t sequence, float temperature,
                           float regularization_weight, float* grad, float* out, void* workspace,
                           void* dump_anchors, void* dump_refs, int dump_dtype, void* stream);
/* Same call with optional staging scratch (cb200_oce_loss_staging_bytes): planar 2-D offsets are gathered from a
 * channels-last copy that the kernel writes at the start of its own launch (see cb200_oce_loss_fwd_bwd_staged);
 * staging == NULL behaves exactly like cb200_oce_loss_sampled. */
CB200_API int cb200_oce_loss_sampled_staged(const void* offsets, int offsets_dtype, int offsets_layout, int batch, int num_dims,
                           const int64_t* spatial, const int64_t* extent, double kappa, int64_t num_anchors,
                           int num_references, uint64_t seed, uint64_t sequence, float temperature,
                           float regularization_weight, float* grad, float* out, void* workspace,
                           void* dump_anchors, void* dump_refs, int dump_dtype,
                           void* staging, int64_t staging_bytes, void* stream);

/* ===================================================================== *
 *  Detect slice (inference)                                             *
 * ===================================================================== */

/*
 * TTA aggregate: models/unet.py:90-98.
 *   stack (T, C, n) fp32 -> out (C+1, n) fp32: per-channel mean over T, then
 *   the per-channel POPULATION std summed over channels.
 */
CB200_API int cb200_tta_aggregate(const float* stack, int num_passes, int channels, int64_t n, float* out, void* stream);
/* Streaming form for an on-device TTA loop (no T-deep stack in HBM):
 *   state (2*C, n) fp32 = running mean, running M2 (Welford); pass index t is 0-based. */
CB200_API int cb200_tta_accumulate(float* state, const float* prediction, int t, int channels, int64_t n, void* stream);
CB200_API int cb200_tta_finalize(const float* state, int num_passes, int channels, int64_t n, float* out, void* stream);

/* Salt / pepper noise of the infer-mode forward, models/unet.py:80-82: out = (u <= p) ? value : raw with
 * u ~ U[0,1) drawn on the device (Philox; `sequence` distinguishes the passes). */
CB200_API int cb200_salt_pepper(const float* raw, int64_t n, float p, float value, uint64_t seed, uint64_t sequence,
                      float* out, void* stream);
/* Same draw with the seed read from DEVICE memory (one uint64) when the kernel runs: a CUDA graph of the whole
 * test-time-augmentation loop of models/unet.py:73-89 then draws fresh noise on every replay. */
CB200_API int cb200_salt_pepper_device_seed(const float* raw, int64_t n, float p, float value, const uint64_t* seed,
                      uint64_t sequence, float* out, void* stream);

/*
 * Foreground threshold, detect.py:88-94 (+ skimage threshold_otsu, np.histogram).
 * cb200_minmax: out2 = {min, max} as doubles.            workspace: cb200_reduce_workspace_bytes()
 * cb200_histogram: np.histogram(x, nbins, range=(edges[0], edges[nbins])) with
 *   numpy's exact uniform-bin index arithmetic incl. the +-1 edge corrections;
 *   edges = np.linspace(min, max, nbins+1) (device, doubles); counts (nbins) uint64, accumulated.
 */
CB200_API int64_t cb200_reduce_workspace_bytes(void);
CB200_API int cb200_minmax(const void* x, int dtype, int64_t n, double* out2, void* workspace, void* stream);
CB200_API int cb200_histogram(const void* x, int dtype, int64_t n, const double* edges, int nbins,
                    unsigned long long* counts, void* stream);

/*
 * Centring, detect.py:97-119: means[k] = mean over the NON-ZERO entries of (std < threshold) * emb[k]
 * (float64, fixed summation order); out (optional, same dtype/shape as emb) = emb with means subtracted
 * from the D offset channels, std channel copied (the `centered-embeddings` dataset, detect.py:119).
 *   workspace: cb200_centre_workspace_bytes(), zero-initialised once.
 */
CB200_API int64_t cb200_centre_workspace_bytes(void);
CB200_API int cb200_centre_embeddings(const void* emb, int dtype, int num_dims, int64_t n_pix, double threshold,
                            double* means /* device, num_dims */, void* out, void* workspace, void* stream);

/*
 * Seed finder of the use_seeds branch, detect.py:128-132:
 *   cb200_channel_norm:   out[i] = sqrt(sum_k emb[k][i]^2) in float64 (np.linalg.norm(centred[:-1], axis=0))
 *   cb200_gaussian_blur:  scipy.ndimage.gaussian_filter (separable, mode="reflect"), bit-exact: per axis
 *                         tmp = x0*w0; for j = radius..1: tmp += (x[-j] + x[+j])*w[j]; weights (host, radius+1,
 *                         w[0] = centre) from scipy's _gaussian_kernel1d; `negate` flips the sign of the result
 *                         (peaks of -smooth).  scratch: one more n_pix doubles.
 *   cb200_local_peaks:    skimage.feature.peak_local_max defaults: pixel equals the 3^D maximum, is strictly above
 *                         `threshold` (the image minimum), not on the 1-px border; raster-ordered linear indices
 *                         and values (sort by value on the host: the list is short).  n_out: device int64.
 */
CB200_API int cb200_channel_norm(const void* emb, int dtype, int num_channels, int64_t n_pix, double* out, void* stream);
CB200_API int cb200_gaussian_blur(const double* in, double* out, double* scratch, int num_dims, const int64_t* spatial,
                        const double* weights, int radius, int negate, void* stream);
CB200_API int64_t cb200_peaks_workspace_bytes(int64_t n_pix);
CB200_API int cb200_local_peaks(const double* img, int num_dims, const int64_t* spatial, double threshold,
                      int32_t* peak_index, double* peak_value, int64_t capacity, long long* n_out,
                      void* workspace, void* stream);

/*
 * Foreground compaction, utils/mean_shift.py:15-36,85,94 (+ detect.py:94):
 *   mask = std < threshold (compared in float64); foreground pixels, in raster
 *   order, become points X[k][i] = emb[k][pix] + coordinate_k (float64, SoA:
 *   points + k*capacity), pix_index[i] = linear pixel index (int32: n_pix < 2^31).
 *   emb (D+1, *spatial) CB200_F32 / CB200_F64, channel D = std.
 *   mask_out: optional (n_pix) CB200_U8/CB200_U16 0/1 image (binary-segmentation, detect.py:95)
 *   n_out: device int64, number of foreground pixels (written before points are).
 *   If the count exceeds `capacity`, nothing beyond capacity is written and
 *   *n_out still holds the true count (call again with a larger buffer).
 *   workspace: cb200_compact_workspace_bytes(n_pix)
 */
CB200_API int64_t cb200_compact_workspace_bytes(int64_t n_pix);
CB200_API int cb200_fg_compact(const void* emb, int dtype, int num_dims, const int64_t* spatial, double threshold,
                     double* points, int32_t* pix_index, int64_t capacity, long long* n_out,
                     void* mask_out, int mask_dtype, void* workspace, void* stream);

/* Row subset of an SoA point set: dst[k][j] = src[k][i] for the j-th i with flags[i] != 0
 * (the `X[np.random.rand(N) < p]` of utils/mean_shift.py:68-70; flags from the host RNG for
 * parity, or from cb200_bernoulli_flags).  n_out: device int64. */
CB200_API int cb200_select_points(const double* src, int64_t n, int64_t src_stride, int num_dims, const uint8_t* flags,
                        double* dst, int64_t dst_stride, long long* n_out, void* workspace, void* stream);
CB200_API int cb200_bernoulli_flags(uint8_t* flags, int64_t n, double p, uint64_t seed, void* stream);

/*
 * Flat-kernel mean-shift hill climb, sklearn:108-128, for every seed:
 *   loop { nbrs = {x : sum_k (m_k - x_k)^2 <= bw^2}; none -> stop;
 *          m = mean(nbrs); stop if ||m - m_old|| <= 1e-3 bw or it == max_iter; ++it }
 * All arithmetic float64, distance evaluated as the KD-tree does (sequential
 * mul/add, no FMA) so the in/out decisions are those of the reference.
 *
 * Brute-force n-body form: one iteration over the ACTIVE seeds per call.
 *   points SoA (D x n, stride pts_stride), means SoA (D x n_seeds, stride seed_stride)
 *   active: indices of seeds still climbing (n_active of them)
 *   partial: scratch of cb200_ms_brute_partial_bytes(n_active, n, num_dims) bytes
 * cb200_ms_brute_accumulate fills the scratch; cb200_ms_update finishes the iteration:
 * updates means/counts/iters, appends still-active seeds to next_active, *n_next (device int).
 */
CB200_API int64_t cb200_ms_brute_partial_bytes(int64_t n_active, int64_t n_points, int num_dims);
CB200_API int cb200_ms_brute_accumulate(const double* points, int64_t n_points, int64_t pts_stride, int num_dims,
                              const double* means, int64_t seed_stride, const int* active, int64_t n_active,
                              double bandwidth, void* partial, void* stream);
CB200_API int cb200_ms_update(double* means, int64_t seed_stride, int num_dims, int* counts, int* iters,
                    const int* active, int64_t n_active, int64_t n_points, const void* partial,
                    double bandwidth, int max_iter, int* next_active, int* n_next, void* stream);

/*
 * Grid-hash pruned form: cells of edge >= bandwidth; each warp climbs one seed
 * to convergence inside a single launch, visiting only the 3^D neighbour cells.
 *   cb200_grid_plan: host-side helper, fills `grid` (origin, cell edge, dims) from a bounding box.
 *   cb200_grid_build: cell id per point -> sort -> points_sorted SoA + cell_start (n_cells+1 ints).
 *   cb200_ms_grid_modes: modes (D x n_seeds SoA, in: seeds, out: modes), counts, iters.
 */
typedef struct cb200_grid {
  double origin[3];
  double cell;      /* edge length, >= bandwidth */
  double inv_cell;
  int32_t dims[3];  /* cells per column (x, y, z); unused = 1 */
  int32_t num_dims;
  int64_t n_cells;
} cb200_grid;

CB200_API int cb200_grid_plan(const double* lo, const double* hi, int num_dims, double bandwidth,
                    int64_t max_cells, cb200_grid* grid /* host out */);
CB200_API int64_t cb200_grid_build_workspace_bytes(int64_t n_points, int64_t n_cells);
CB200_API int cb200_grid_build(const double* points, int64_t n_points, int64_t pts_stride, const cb200_grid* grid,
                     double* points_sorted, int64_t sorted_stride, int* order /* n_points, may be NULL */,
                     int* cell_start /* n_cells + 1 */, void* workspace, int64_t workspace_bytes, void* stream);
CB200_API int cb200_ms_grid_modes(const double* points_sorted, int64_t n_points, int64_t sorted_stride,
                        const cb200_grid* grid, const int* cell_start,
                        double* means, int64_t seed_stride, int64_t n_seeds,
                        double bandwidth, int max_iter, int* counts, int* iters,
                        int* work_counter /* device, 32 bytes, zeroed by the caller: [0] seed claim
                                             counter; bytes 8..15: uint64 statistic, distance tests made;
                                             bytes 16..23: uint64 sum over seeds of (iterations + 1) */,
                        void* stream);
/*
 * The same climb over DISTINCT trajectories only.  Two seeds whose means are bit-identical after the same number of
 * iterations follow the same trajectory from there on, and after one window evaluation most seeds of an object already
 * share their mean.  Pass 1 gives every seed ONE evaluation; the unconverged seeds are merged by the bit pattern of their
 * mean; pass 2 climbs one representative (the lowest seed index) of every distinct mean to convergence.  With
 * merge_rounds > 1 the representatives get ONE more evaluation per round and are merged again (trajectories keep
 * meeting as they approach the same window; worth it for millions of dense seeds), to convergence after the last.  On return
 * every finished seed and every representative holds exactly what cb200_ms_grid_modes gives it (mean, count,
 * iterations); the merged copies have count 0 (they would end as copies of their representative's mode, which the
 * centre post-processing merges anyway), so cb200_unique_modes / cb200_nms_suppress see the same distinct modes.
 *   work       16 ints, zeroed by the caller: [0] claims, [2..5] tests / climb steps of pass 1 (two uint64); [8], [10..13]
 *              the same for pass 2
 *   workspace  cb200_ms_distinct_workspace_bytes(n_seeds)
 */
CB200_API int64_t cb200_ms_distinct_workspace_bytes(int64_t n_seeds);
CB200_API int cb200_ms_grid_modes_distinct(const double* points_sorted, int64_t n_points, int64_t sorted_stride,
                        const cb200_grid* grid, const int* cell_start,
                        double* means, int64_t seed_stride, int64_t n_seeds,
                        double bandwidth, int max_iter, int merge_rounds /* 1..30 */, int* counts, int* iters, int* work,
                        void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Grid-binned seeds: scikit-learn's get_bin_seeds (sklearn:254-297; MeanShift(bin_seeding=True), min_bin_freq = 1),
 * the "grid-binned" seeding of BASELINE configs[3] (the reference itself always seeds with every fit point):
 *   bins = unique np.round(point / bin_size) (float64 division, round half to even), seeds = float32(bin) *
 *   float32(bin_size) widened to float64, written as an SoA (D x seed_stride, seed_stride >= n_points) in key
 *   order (the order of seeds does not influence the fitted centres); if every point has its own bin the points
 *   themselves are the seeds, as in scikit-learn.  *n_out (device int64) = number of seeds; *overflow (device
 *   int) is set when a bin index does not fit 21 bits (|point / bin_size| >= 2^20): the result is then invalid.
 */
CB200_API int64_t cb200_bin_seeds_workspace_bytes(int64_t n_points);
CB200_API int cb200_bin_seeds(const double* points, int64_t n_points, int64_t pts_stride, int num_dims, double bin_size,
                    double* seeds, int64_t seed_stride, long long* n_out, int* overflow, void* workspace,
                    int64_t workspace_bytes, void* stream);

/*
 * Centre post-processing, sklearn:511-547: drop empty seeds, order by
 * (count, coords) descending, merge exact duplicates, greedy suppression of
 * every centre within `bandwidth` (inclusive) of an earlier survivor --
 * reproduced exactly as the lexicographically-first maximal independent set
 * (parallel rounds; priority = sklearn's sort key evaluated pairwise).
 *
 * cb200_nms_suppress: runs `rounds` (1..16) rounds over a FINE grid hash of the modes (edge bw/2:
 *   modes sharing a cell are always within the bandwidth, so only the best live mode of a
 *   cell can survive next -- one warp per non-empty cell per round).
 *   grid must cover every mode with count > 0 (the bounding box of the fit points does);
 *   resume = 0 starts from scratch, resume = 1 continues on the same workspace;
 *   n_keep_and_undecided (device int[2]): [0] = survivors so far, [1] = modes still
 *   undecided -- the result is final when [1] == 0 (2-3 rounds in practice).
 * cb200_nms_emit: writes the n_keep survivors in sklearn's `cluster_centers_` order
 *   into centres_out SoA (D x centre_stride).  n_keep is the host copy of [0].
 */
/* Exact duplicates first (optional, in front of cb200_nms_suppress): a flat-kernel hill climb has finitely many fixed
 * points -- 148 228 converged seeds of BASELINE configs[2] are 395 distinct modes -- and scikit-learn merges them in a
 * dict (sklearn:514-521).  Keeps ONE copy of every bit-identical mode with count > 0: the copy the suppression would
 * keep (highest count, then lowest index), in input order; the surviving centres and their order are the same with or
 * without this pass.  modes_out SoA (D x out_stride, out_stride >= n_seeds may alias nothing), counts_out (n_seeds),
 * *n_out (device int64) = number of distinct modes. */
CB200_API int64_t cb200_unique_modes_workspace_bytes(int64_t n_seeds);
CB200_API int cb200_unique_modes(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                       double* modes_out, int64_t out_stride, int* counts_out, long long* n_out, void* workspace,
                       int64_t workspace_bytes, void* stream);
CB200_API int64_t cb200_nms_workspace_bytes(int64_t n_seeds, const cb200_grid* grid, double bandwidth);
CB200_API int cb200_nms_suppress(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                       double bandwidth, const cb200_grid* grid, int rounds, int resume,
                       int* n_keep_and_undecided, void* workspace, int64_t workspace_bytes, void* stream);
CB200_API int cb200_nms_emit(const double* modes, int64_t seed_stride, int num_dims, const int* counts, int64_t n_seeds,
                   double bandwidth, const cb200_grid* grid, int n_keep, double* centres_out, int64_t centre_stride,
                   void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Label assignment, sklearn:563-579 (`predict` = nearest centre, ties -> lowest
 * index, orphans labelled too) fused with the scatter of utils/mean_shift.py:101-104,57:
 *   labels_out[pix_index[i]] = 1 + argmin_k ||X_i - c_k||  (label volume pre-zeroed by caller)
 *   label_dtype CB200_I32 (mean_shift_segmentation's return) or CB200_U16 (detect.py:30,161).
 *   pix_index may be NULL: labels are then written densely, labels_out[i].
 *   grid + workspace (cb200_assign_workspace_bytes) switch on the pruned search: centres are
 *   hashed into cells of edge >= bandwidth, a centre found within one edge in the 3^D block is
 *   the global nearest, the remaining points (orphans) are finished by brute force.
 *   grid == NULL or workspace == NULL: brute force over all centres.
 */
CB200_API int64_t cb200_assign_workspace_bytes(int64_t n_points, int n_centres, int64_t n_cells);
CB200_API int cb200_assign_labels(const double* points, int64_t n_points, int64_t pts_stride, int num_dims,
                        const double* centres, int64_t centre_stride, int n_centres, const cb200_grid* grid,
                        const int32_t* pix_index, void* labels_out, int label_dtype, void* workspace, void* stream);

/*
 * The whole per-bandwidth sequence of `mean_shift_segmentation` (utils/mean_shift.py:6-45 + sklearn fit /
 * predict) behind one call, for callers that do not need the intermediate results:
 *   threshold -> foreground points -> fit subset (Bernoulli(reduction_probability) from the device Philox
 *   stream; 1.0 = all points) -> cell grid -> modes -> centres -> labels (+1, 0 = background).
 * Same kernels and results as the step-by-step entry points above.  It allocates nothing: all scratch comes from the
 * caller's `workspace`.  Its sizes depend on counts that only exist on the device (foreground pixels, fit subset,
 * cells, centres), so unlike the other entry points this one SYNCHRONISES the stream to read them (three small
 * device -> host reads per volume) and is therefore not graph-capturable.  Protocol:
 *   workspace_bytes = cb200_detect_volume_workspace_bytes(num_dims, spatial, expected_foreground, reduction_probability)
 *   foreground_capacity = the expected_foreground the workspace was sized for (<= 0: every pixel)
 *   return CB200_ENOSPACE: more foreground (or cells, or seeds) than there was room for; nothing useful was written,
 *   info->n_foreground holds the count found and info->workspace_needed the bytes to call again with.
 * labels_out (n_pix, CB200_I32 / CB200_U16) is cleared by the call; mask_out optional
 * (CB200_U8 / CB200_U16); centres_out optional device SoA (D x centre_capacity) receiving `cluster_centers_`.
 * Returns CB200_ENOFIT / CB200_ENOCENTRE where scikit-learn raises ValueError.
 */
typedef struct cb200_detect_info {
  int64_t n_foreground;
  int64_t n_fit;
  int64_t n_seeds;
  int32_t n_centres;
  int32_t suppress_calls;
  cb200_grid grid;
  int64_t distance_tests; /* seed x candidate evaluations of the hill climb (its algorithmic work) */
  int64_t climb_steps;    /* sum over seeds of (iterations + 1): window evaluations */
  int64_t workspace_needed; /* set with CB200_ENOSPACE */
  int64_t n_distinct_modes; /* bit-distinct converged modes with count > 0: what the suppression ran on */
} cb200_detect_info;

CB200_API int64_t cb200_detect_volume_workspace_bytes(int num_dims, const int64_t* spatial, int64_t expected_foreground,
                                            double reduction_probability);
CB200_API int cb200_detect_volume(const void* emb, int dtype, int num_dims, const int64_t* spatial, double threshold,
                        double bandwidth, double reduction_probability, uint64_t philox_seed, int max_iter,
                        void* labels_out, int label_dtype, void* mask_out, int mask_dtype, double* centres_out,
                        int64_t centre_capacity, void* workspace, int64_t workspace_bytes, int64_t foreground_capacity,
                        cb200_detect_info* info /* host out */, void* stream);

/*
 * Greedy seed-and-grow clustering, utils/greedy_cluster.py:46-120,176-253 (clustering = "greedy",
 * detect.py:162-192): one persistent cooperative kernel runs the whole sequential loop on the device.
 *   cb200_greedy_prepare: foreground pixels (fg_mask != 0, raster order) -> emb_masked SoA (D x capacity) =
 *       emb[k] + coordinate_k and seed_masked = (std - max) / (min - max), both in compute_dtype (the
 *       reference computes 2-D in fp32 and 3-D in the dtype of the stored embeddings); pix_index for the scatter.
 *   cb200_greedy_cluster: instance_masked (int16, n_points) = 1.. per accepted proposal, 0 otherwise;
 *       n_objects_and_iterations (device int[2], optional).
 *   cb200_scatter_i16: dst[pix_index[i]] = src[i].
 */
CB200_API int64_t cb200_greedy_workspace_bytes(int64_t n_points);
CB200_API int cb200_greedy_prepare(const void* emb, int dtype, int num_dims, const int64_t* spatial, const uint8_t* fg_mask,
                         int compute_dtype, double seed_min, double seed_max, void* emb_masked, int64_t capacity,
                         void* seed_masked, int32_t* pix_index, long long* n_out, void* workspace, void* stream);
CB200_API int cb200_greedy_cluster(const void* emb_masked, int64_t stride, const void* seed_masked, int64_t n_points,
                         int num_dims, int compute_dtype, double bandwidth, int min_object_size, double seed_thresh,
                         long long min_unclustered_sum, short* instance_masked, int* n_objects_and_iterations,
                         void* workspace, void* stream);
CB200_API int cb200_scatter_i16(const short* src, const int32_t* pix_index, int64_t n, short* dst, void* stream);

/*
 * Connected-component size filter, utils/misc.py:11-25 (skimage.measure.label:
 * full connectivity, equal-valued regions, 0 = background, raster-order ids).
 *   cb200_label_components: labels_out int32 (n_pix); *n_labels device int.
 *   cb200_size_filter: in-place zeroing of components smaller than min_size in
 *   `seg` (int32), then relabel into labels_out.
 */
CB200_API int64_t cb200_cc_workspace_bytes(int64_t n_pix);
CB200_API int cb200_label_components(const int32_t* seg, int num_dims, const int64_t* spatial,
                           int32_t* labels_out, int* n_labels, void* workspace, void* stream);
CB200_API int cb200_size_filter(int32_t* seg, int num_dims, const int64_t* spatial, int64_t min_size,
                      int32_t* labels_out, int* n_labels, void* workspace, void* stream);

/*
 * "cell" post-processing, segment.py:42-52 (scipy.ndimage.distance_transform_edt, unit sampling):
 *   cb200_edt_within : out[p] = (dtedt(input)[p] < radius); `input` uint8, nonzero = inside.  An input without
 *                      any zero element measures distances to the virtual element (-1, 0, ..., 0), as scipy does.
 *   cb200_grow_shrink: in place on int32 labels,
 *                        expanded = dtedt(seg == 0) < grow_distance        (:47-48)
 *                        seg[dtedt(expanded) < shrink_distance] = 0        (:49-50)
 *   radius < 16384; num_dims 2 or 3.
 */
CB200_API int64_t cb200_edt_workspace_bytes(int64_t n_pix);
CB200_API int cb200_edt_within(const uint8_t* input, int num_dims, const int64_t* spatial, double radius, uint8_t* out,
                     void* workspace, void* stream);
CB200_API int cb200_grow_shrink(int32_t* seg, int num_dims, const int64_t* spatial, double grow_distance,
                      double shrink_distance, void* workspace, void* stream);

/*
 * "nucleus" post-processing, segment.py:52-101: per instance an Otsu threshold of the raw intensities under
 * the instance (skimage.filters.threshold_otsu), mask = instance & (raw > threshold), holes filled inside the
 * instance's bounding box (scipy.ndimage.binary_fill_holes), ids written in ascending order.
 *   raw dtype: CB200_U8 / CB200_U16 (skimage's one-bin-per-value histogram) or CB200_F32 / CB200_F64
 *   (np.histogram, 256 bins over the instance's [min, max]).
 *   cb200_label_stats     : per label id in [0, max_label]: raw_min / raw_max (double) and box[id][6] =
 *                           lo z,y,x, hi z,y,x (inclusive; hi < 0 = label absent; z = 0 in 2-D)
 *   cb200_label_histogram : integer raw: hist[hist_offset[id] + (v - raw_min[id])]++ ;
 *                           float raw  : hist[id * nbins + b]++ with edges[id][nbins + 1] as numpy builds them
 *                           (`hist` zeroed by the caller; the O(bins) Otsu tail runs on the host)
 *   cb200_nucleus_fill    : out (int32, zeroed by the call) gets ids[k] on mask_k and on its holes; boxes[k][6]
 *                           and box_offset[k] = prefix sum of the box volumes (n_instances + 1 entries)
 */
CB200_API int64_t cb200_label_stats_workspace_bytes(int max_label);
CB200_API int cb200_label_stats(const int32_t* seg, const void* raw, int raw_dtype, int num_dims, const int64_t* spatial,
                      int max_label, double* raw_min, double* raw_max, int32_t* box, void* workspace, void* stream);
CB200_API int cb200_label_histogram(const int32_t* seg, const void* raw, int raw_dtype, int64_t n_pix, int max_label,
                          const double* raw_min, const int64_t* hist_offset, const double* edges, int nbins,
                          unsigned int* hist, void* stream);
/*   cb200_label_otsu      : the O(bins) tail of threshold_otsu for every label at once in numpy's arithmetic
 *                           (float32 cumulative weights; products, means and variance in `arithmetic_dtype` =
 *                           CB200_F32 for float32 images, CB200_F64 for integer and float64 images).  Integer
 *                           images (`centres` == NULL): one BLOCK per label with block-wide scans -- all partial
 *                           sums are integers, exact in any order; float images: one thread per label in
 *                           numpy's sequential order.  Label l owns hist[hist_offset[l] .. + num_bins[l]); bin centre
 *                           i is centres[hist_offset[l] + i] when `centres` is given, else centre0[l] + i.
 *                           thresholds[l] = the centre at the first maximum; labels with num_bins <= 0 untouched. */
CB200_API int64_t cb200_label_otsu_workspace_bytes(int64_t total_bins);
CB200_API int cb200_label_otsu(const unsigned int* hist, const int64_t* hist_offset, const int64_t* num_bins,
                     const double* centre0, const double* centres, int arithmetic_dtype, int n_labels,
                     int64_t total_bins, double* thresholds, void* workspace, void* stream);
CB200_API int64_t cb200_nucleus_fill_workspace_bytes(int64_t total_box_voxels, int n_instances);
CB200_API int cb200_nucleus_fill(const int32_t* seg, const void* raw, int raw_dtype, int num_dims, const int64_t* spatial,
                       int n_instances, const int32_t* ids, const double* thresholds, const int32_t* boxes,
                       const int64_t* box_offset, int64_t total_box_voxels, int32_t* out, void* workspace,
                       void* stream);

/*
 * Evaluation counts, evaluate.py:72-98 (compute_pairwise_IoU): one pass instead of four per id pair.
 *   cb200_label_presence: present[v] = 1 for every value v in [0, max_value] that occurs in `labels`
 *                         (CB200_U16 or CB200_I32; values outside the range are ignored)
 *   cb200_contingency   : table[rank_pred[pred[i]] * cols + rank_gt[gt[i]]]++ over all pixels; the rank
 *                         tables (max_value + 1 entries each) map a label value to its row / column
 *                         (0 = background; values outside [0, max_value] count as background); `table` is zeroed
 *                         by the call.  Intersections are the entries, areas the row / column sums.
 */
CB200_API int cb200_label_presence(const void* labels, int dtype, int64_t n, int max_value, uint8_t* present, void* stream);
CB200_API int cb200_contingency(const void* pred, const void* gt, int dtype, int64_t n, const int32_t* rank_pred,
                      const int32_t* rank_gt, int max_value, int rows, int cols, unsigned int* table, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CELLULUS_B200_H */
