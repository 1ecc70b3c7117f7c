// OCE loss with the pair sampling INSIDE the kernel (sm_100a): sampling, neighbour gather, pair terms and the
// backward in one pass, with no coordinate list in HBM at all.
//
// Replaces, per training step (cellulus/train.py:160-180 + datasets/zarr_dataset.py:177-251):
//     ZarrDataset.sample_coordinates()                      (DataLoader workers, two int64 lists over PCIe)
//     select_and_add_coordinates(offsets, anchors / refs)   models/unet.py:108-124
//     OCELoss.forward + loss.backward()                     criterions/oce_loss.py:53-63
// The pairs are those of the device pair stream (pair_stream.cuh): the explicit-list kernel of oce_loss.cu fed
// cb200_sample_pairs(seed, sequence) computes the same loss, which is how parity is tested.
//
// Work mapping: ONE LANE OWNS ONE ANCHOR.  The stream repeats every anchor R times (np.repeat, :236), so the
// anchor's coordinates, its gathered embedding, ||ea|| and the whole regulariser term are computed once per
// R pairs and the anchor's gradient accumulates in registers: no shuffles, no segmented reduction, exactly one
// reduction into the gradient tensor per anchor.  A warp takes 32 consecutive anchors of one sample; every
// trip draws one offset per lane from the shared-memory table of admissible offsets (one Philox block feeds
// four trips) and gathers one reference pixel.
//
// HBM traffic: the offsets tensor once + the dense gradient once (31.5 MB at BASELINE configs[1], against
// 211 MB with int64 lists); the gathers and the per-anchor reductions are L2 traffic.
#include "loss_common.cuh"
#include "pair_stream.cuh"

namespace cb200 {

constexpr int SMP_THREADS = 256;
constexpr int SMP_MIN_BLOCKS = 5;  // 709 blocks of 8 anchor-warps at configs[1]: one wave needs 5 per SM

// LY_STAGED (planar 2-D offsets + caller scratch): every CTA first writes its share of a channels-last copy of the
// offsets into `staged` and counts itself in; the gathers read the copy (one 8-byte access per pixel instead of one
// wavefront per lane PER CHANNEL), the gradient stays planar.  The grid is one resident wave, so the wait for the
// other CTAs cannot starve.
template <int D, typename OT, int LY, bool BWD, bool DUMP>
__global__ void __launch_bounds__(SMP_THREADS, SMP_MIN_BLOCKS)
oce_loss_sampled_kernel(const OT* __restrict__ offsets, PairStreamParams p, Shape<D> shape, unsigned batch,
                        unsigned blocks_per_sample /* of 32 anchors */, float neg_log2e_over_t, float two_over_t, float w,
                        float* __restrict__ grad, LossWorkspace* ws, float* out, void* dump_anchors, void* dump_refs,
                        int dump_dtype, OT* staged, unsigned stage_vec) {
  extern __shared__ uint32_t s_table[];
  if constexpr (LY == LY_STAGED && D == 2) {
    stage_interleaved_slice<OT>(offsets, staged, batch, (unsigned)shape.npix, stage_vec != 0);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ws->arrived[0]) : "memory");
  }
  build_offset_table<D>(s_table, p);  // overlaps the zero-fill grid in front of us
  if constexpr (LY == LY_STAGED && D == 2) {
    if (threadIdx.x == 0) {
      unsigned seen;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&ws->arrived[0]) : "memory");
        if (seen >= gridDim.x) break;
        __nanosleep(40);
      }
    }
    __syncthreads();
  }
  const OT* gather_base = LY == LY_STAGED ? staged : offsets;
  const Philox rng(p.seed);
  const unsigned lane = lane_id();
  const unsigned warps_per_block = SMP_THREADS / 32;
  const unsigned n_tasks = batch * blocks_per_sample;
  const unsigned task_stride = gridDim.x * warps_per_block;
  const unsigned npix = (unsigned)shape.npix;

  float acc_oce = 0.f, acc_nrm = 0.f;
  int bad = 0;
  bool waited = !BWD;

  for (unsigned task = blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < n_tasks; task += task_stride) {
    const unsigned b = task / blocks_per_sample;
    const unsigned a_raw = (task - b * blocks_per_sample) * 32u + lane;
    const bool owns = a_raw < p.num_anchors;
    const unsigned a = owns ? a_raw : p.num_anchors - 1;  // idle lanes shadow the last anchor, results masked
    const unsigned first = b * npix;

    int anc[D];
    stream_anchor<D>(rng, p, b, a, anc);
    bool ok_a = owns;
#pragma unroll
    for (int k = 0; k < D; ++k) ok_a = ok_a && ((unsigned)anc[k] < (unsigned)shape.ext[k]);
    // an anchor outside the tensor (sampling extent != tensor extent): all of its pairs are skipped and counted
    if (owns && !ok_a) bad += (int)p.num_refs;
    const unsigned pa = ok_a ? (unsigned)pixel_of<D>(anc, shape) : 0u;
    float oa[D], ea[D], g[D];
    if (ok_a) {
      if constexpr (LY == LY_STAGED) {
        if constexpr (D == 2) gather_staged<D, OT>(gather_base, first, pa, oa);
      } else {
        gather_pixel<D, OT, LY == LY_CL>(gather_base, npix, first, pa, oa);
      }
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) oa[k] = 0.f;
    }
    float n2 = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      ea[k] = __fadd_rn(oa[k], (float)anc[k]);  // selection += coordinate (models/unet.py:120)
      n2 = fmaf(ea[k], ea[k], n2);
      g[k] = 0.f;
    }
    int n_ok = 0;

    float fa[D];  // the anchor coordinate as a float: (float)(anc + off) == fa + (float)off exactly (small integers)
#pragma unroll
    for (int k = 0; k < D; ++k) fa[k] = (float)anc[k];

    // Software pipeline over ROUNDS of four pairs (the four words of a Philox block; a block gives `draws`
    // rounds).  Four slots stay in flight per lane -- packed offset + D gathered floats each -- and every slot
    // is refilled with the next round's gather right after its math, so a warp always has 3-4 gathers
    // outstanding.  (Left to the compiler, the gathers of a round were serialised behind each other's math.)
    uint32_t pk[4];
    float orf[4][D];
    unsigned okm = 0;  // bit j: slot j holds a live pair
    unsigned t_next = 0;
    auto issue = [&](int j, uint32_t word) {
      const unsigned t = t_next + j;
      pk[j] = s_table[bounded(word, p.n_table)];
      int ref[D];
      unpack_offset<D>(pk[j], ref);
      const bool want = ok_a && (t < p.num_refs);
      bool ok = want;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        ref[k] += anc[k];
        ok = ok && ((unsigned)ref[k] < (unsigned)shape.ext[k]);
      }
      if (want && !ok) ++bad;
      okm = ok ? (okm | (1u << j)) : (okm & ~(1u << j));
      if (ok) {
        if constexpr (LY == LY_STAGED) {
          if constexpr (D == 2) gather_staged<D, OT>(gather_base, first, (unsigned)pixel_of<D>(ref, shape), orf[j]);
        } else {
          gather_pixel<D, OT, LY == LY_CL>(gather_base, npix, first, (unsigned)pixel_of<D>(ref, shape), orf[j]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) orf[j][k] = 0.f;
      }
      if constexpr (DUMP) {  // debug / test mode: the lists this call used
        if (owns && t < p.num_refs) {
          const size_t pair = ((size_t)b * p.num_anchors + a) * p.num_refs + t;
          store_coord_dyn<D>(dump_anchors, dump_dtype, pair, anc);
          store_coord_dyn<D>(dump_refs, dump_dtype, pair, ref);
        }
      }
    };
    auto math = [&](int j) {
      int off[D];
      unpack_offset<D>(pk[j], off);
      float diff[D], d2 = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const float er = __fadd_rn(orf[j][k], fa[k] + (float)off[k]);  // reference embedding + its coordinate
        diff[k] = ea[k] - er;
        d2 = fmaf(diff[k], diff[k], d2);
      }
      const float e = ex2_approx(d2 * neg_log2e_over_t);  // exp(-d^2 / T)
      if (okm & (1u << j)) {
        acc_oce += 1.0f - e;
        ++n_ok;
        if constexpr (BWD) {
          const float ge = two_over_t * e;
#pragma unroll
          for (int k = 0; k < D; ++k) g[k] = fmaf(ge, diff[k], g[k]);
        }
      }
    };

    uint4 ro = stream_offset_block(rng, p, b, a, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) issue(j, pick_word(ro, j));
    const unsigned n_rounds = p.n_tg * p.draws;
    unsigned h = 1, tg = 0;  // position of the round being issued: draw h of block tg
#pragma unroll 1
    for (unsigned r = 1; r < n_rounds; ++r, ++h) {
      t_next += 4;
      if (h == p.draws) {
        h = 0;
        ro = stream_offset_block(rng, p, b, a, ++tg);
      } else {  // second draw of every word: the low half of the first product is the next uniform word
        ro.x *= p.n_table; ro.y *= p.n_table; ro.z *= p.n_table; ro.w *= p.n_table;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        math(j);
        issue(j, pick_word(ro, j));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) math(j);
    // the regulariser depends on the anchor only: n_ok pairs contribute ||ea|| each (criterions/oce_loss.py:59-61)
    const float rs = n2 > 0.f ? rsqrt_approx(n2) : 0.f;  // 1 / ||ea||, 0 at the origin (torch's norm backward)
    const float fn = (float)n_ok;
    acc_nrm = fmaf(fn, n2 * rs, acc_nrm);
    if constexpr (BWD) {
      if (!waited) {  // the gradient tensor is zero-filled by the grid in front (programmatic dependent launch)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        waited = true;
      }
      if (n_ok > 0) {
        const float gr = fn * w * rs;
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = fmaf(gr, ea[k], g[k]);
        scatter_pixel<D, LY == LY_CL>(grad, npix, first, pa, g);
      }
    }
  }
  if constexpr (BWD) {
    if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  block_reduce_to_workspace(acc_oce, acc_nrm, bad, ws, w, out);
}

template <int D, typename OT, int LY, bool BWD, bool DUMP>
static int launch_sampled_variant(const void* offsets, const PairStreamParams& p, const Shape<D>& shape, int batch,
                                  float T, float w, float* grad, float* out, LossWorkspace* ws, void* dump_anchors,
                                  void* dump_refs, int dump_dtype, void* staged, cudaStream_t st) {
  auto kernel = oce_loss_sampled_kernel<D, OT, LY, BWD, DUMP>;
  const size_t table_bytes = (size_t)p.n_table * sizeof(uint32_t);
  static int occupancy = 0;  // per template instantiation
  static size_t occupancy_for = 0;
  if (occupancy == 0 || occupancy_for != table_bytes) {
    if (table_bytes > 48 * 1024)
      CB200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
    int occ = 0;
    CB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, SMP_THREADS, table_bytes));
    occupancy = occ > 0 ? occ : 1;
    occupancy_for = table_bytes;
  }
  const unsigned blocks_per_sample = (p.num_anchors + 31) / 32;
  const uint64_t n_tasks = (uint64_t)batch * blocks_per_sample;
  if (n_tasks >= ((uint64_t)1 << 31)) return CB200_EUNSUPPORTED;
  const unsigned warps = SMP_THREADS / 32;
  uint64_t blocks = (n_tasks + warps - 1) / warps;
  const uint64_t resident = (uint64_t)CB200_SM_COUNT * occupancy;  // persistent grid: one wave
  if (blocks > resident) blocks = resident;
  if (blocks < 1) blocks = 1;
  const float log2e = 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(SMP_THREADS);
  cfg.dynamicSmemBytes = table_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = (BWD && g_loss_pdl) ? 1 : 0;  // only behind our own zero-fill grid
  const unsigned stage_vec = (LY == LY_STAGED && shape.npix % 2 == 0 && reinterpret_cast<uintptr_t>(offsets) % 8 == 0 &&
                              reinterpret_cast<uintptr_t>(staged) % 16 == 0) ? 1u : 0u;
  CB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, (const OT*)offsets, p, shape, (unsigned)batch, blocks_per_sample,
                                    -log2e / T, 2.0f / T, w, grad, ws, out, dump_anchors, dump_refs, dump_dtype,
                                    (OT*)staged, stage_vec));
  return CB200_OK;
}

template <int D, typename OT>
static int launch_sampled(const void* offsets, int layout, int batch, const int64_t* spatial, const int64_t* extent,
                          double kappa, int64_t num_anchors, int num_refs, uint64_t seed, uint64_t sequence, float T,
                          float w, float* grad, float* out, void* workspace, void* dump_anchors, void* dump_refs,
                          int dump_dtype, void* staging, int64_t staging_bytes, cudaStream_t st) {
  Shape<D> shape;
  if (!make_shape<D>(spatial, shape)) return CB200_EINVAL;
  if ((int64_t)batch * D * shape.npix > INT32_MAX || batch > 65535) return CB200_EUNSUPPORTED;  // 32-bit element indices
  PairStreamParams p;
  if (!pair_stream_plan(p, D, extent, kappa, num_anchors, num_refs)) return CB200_EINVAL;
  if ((size_t)p.n_table * sizeof(uint32_t) > PAIR_TABLE_MAX_BYTES) return CB200_EUNSUPPORTED;
  p.seed = seed;
  p.sequence = sequence;
  auto* ws = static_cast<LossWorkspace*>(workspace);
  if (grad) {
    const int rc = zero_fill(grad, (int64_t)batch * D * shape.npix, st);
    if (rc != CB200_OK) return rc;
  }
  if (num_anchors == 0 || num_refs == 0) {  // nothing to draw: loss 0, gradient 0
    CB200_CUDA_TRY(cudaMemsetAsync(out, 0, 4 * sizeof(float), st));
    return CB200_OK;
  }
  const bool il = layout == CB200_LAYOUT_CHANNELS_LAST;
  if (il && D == 2 && (reinterpret_cast<uintptr_t>(offsets) % (2 * sizeof(OT)) || reinterpret_cast<uintptr_t>(grad) % 8))
    return CB200_EINVAL;  // vector gathers / vector reductions need pixel-aligned bases
#define CB200_SAMPLED(BWD, LY)                                                                                       \
  (dump_anchors ? launch_sampled_variant<D, OT, LY, BWD, true>(offsets, p, shape, batch, T, w, grad, out, ws,            \
                                                                dump_anchors, dump_refs, dump_dtype, staging, st)        \
                : launch_sampled_variant<D, OT, LY, BWD, false>(offsets, p, shape, batch, T, w, grad, out, ws, nullptr, \
                                                                 nullptr, 0, staging, st))
  if constexpr (D == 2) {
    // planar tensors + caller scratch: gather from a channels-last copy made inside the kernel
    static const bool staged_ok = [] { const char* e = getenv("CB200_LOSS_STAGED"); return !(e && e[0] == '0'); }();
    if (!il && staged_ok && staging && staging_bytes >= (int64_t)sizeof(OT) * batch * D * shape.npix &&
        reinterpret_cast<uintptr_t>(staging) % (2 * sizeof(OT)) == 0)
      return grad ? CB200_SAMPLED(true, LY_STAGED) : CB200_SAMPLED(false, LY_STAGED);
  }
  if (grad) return il ? CB200_SAMPLED(true, LY_CL) : CB200_SAMPLED(true, LY_PLANAR);
  return il ? CB200_SAMPLED(false, LY_CL) : CB200_SAMPLED(false, LY_PLANAR);
#undef CB200_SAMPLED
}

}  // namespace cb200

using namespace cb200;

extern "C" int cb200_oce_loss_sampled(const void* offsets, int offsets_dtype, int offsets_layout, int batch, int num_dims,
                                      const int64_t* spatial, const int64_t* extent, double kappa, int64_t num_anchors,
                                      int num_references, uint64_t seed, uint64_t sequence, float temperature,
                                      float regularization_weight, float* grad, float* out, void* workspace,
                                      void* dump_anchors, void* dump_refs, int dump_dtype, void* stream) {
  return cb200_oce_loss_sampled_staged(offsets, offsets_dtype, offsets_layout, batch, num_dims, spatial, extent, kappa,
                                       num_anchors, num_references, seed, sequence, temperature, regularization_weight,
                                       grad, out, workspace, dump_anchors, dump_refs, dump_dtype, nullptr, 0, stream);
}

extern "C" int cb200_oce_loss_sampled_staged(const void* offsets, int offsets_dtype, int offsets_layout, int batch,
                                             int num_dims, const int64_t* spatial, const int64_t* extent, double kappa,
                                             int64_t num_anchors, int num_references, uint64_t seed, uint64_t sequence,
                                             float temperature, float regularization_weight, float* grad, float* out,
                                             void* workspace, void* dump_anchors, void* dump_refs, int dump_dtype,
                                             void* staging, int64_t staging_bytes, void* stream) {
  if (!offsets || !spatial || !extent || !out || !workspace) return CB200_EINVAL;
  if (batch <= 0 || num_anchors < 0 || num_references < 0 || !(temperature != 0.f)) return CB200_EINVAL;
  if ((dump_anchors == nullptr) != (dump_refs == nullptr)) return CB200_EINVAL;
  if (dump_anchors && dump_dtype != CB200_I64 && dump_dtype != CB200_I32 && dump_dtype != CB200_I16) return CB200_EUNSUPPORTED;
  if (dump_anchors && num_dims == 2) {  // vector stores of (x, y)
    const size_t align = dump_dtype == CB200_I64 ? 16 : dump_dtype == CB200_I32 ? 8 : 4;
    if ((reinterpret_cast<uintptr_t>(dump_anchors) | reinterpret_cast<uintptr_t>(dump_refs)) % align) return CB200_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_DISPATCH(DD)                                                                                               \
  if (offsets_dtype == CB200_F32)                                                                                        \
    return launch_sampled<DD, float>(offsets, offsets_layout, batch, spatial, extent, kappa, num_anchors, num_references, \
                                     seed, sequence, temperature, regularization_weight, grad, out, workspace,          \
                                     dump_anchors, dump_refs, dump_dtype, staging, staging_bytes, st);                   \
  if (offsets_dtype == CB200_BF16)                                                                                       \
    return launch_sampled<DD, __nv_bfloat16>(offsets, offsets_layout, batch, spatial, extent, kappa, num_anchors,        \
                                             num_references, seed, sequence, temperature, regularization_weight, grad,  \
                                             out, workspace, dump_anchors, dump_refs, dump_dtype, staging,              \
                                             staging_bytes, st);                                                        \
  return CB200_EUNSUPPORTED;
  if (num_dims == 2) { CB200_DISPATCH(2) }
  if (num_dims == 3) { CB200_DISPATCH(3) }
#undef CB200_DISPATCH
  return CB200_EUNSUPPORTED;
}
