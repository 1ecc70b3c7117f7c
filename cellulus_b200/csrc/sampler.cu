// Device pair sampler for sm_100a -- the B200-native counterpart of
// ZarrDataset.sample_coordinates / sample_offsets_within_radius
// (datasets/zarr_dataset.py:177-251).  Counter-based Philox4x32-10: every pair
// derives its anchor from (sample, anchor index) and its offset from the pair
// index, so no state is shared and the lists never exist on the host.
//
// Distribution (identical to the reference's):
//   anchor column k ~ U{trunc(kappa) .. extent_k - trunc(kappa)}  (np.random.randint(kappa, out-kappa+1))
//   offset ~ uniform over integer o in [-trunc(kappa), trunc(kappa)]^D with sum o^2 < kappa^2, o != 0
//            (rejection sampling = what `in_circle` / `not_zero` filtering of i.i.d. draws yields)
//   each anchor repeated num_references times consecutively (np.repeat, :236)
#include "common.cuh"

namespace cb200 {

// one pair's coordinates as a single store where the type allows it
template <int D, typename CT>
__device__ __forceinline__ void store_coord(CT* __restrict__ base, size_t pair, const int (&c)[D]) {
  if constexpr (D == 2 && sizeof(CT) == 8) {
    reinterpret_cast<longlong2*>(base)[pair] = make_longlong2(c[0], c[1]);
  } else if constexpr (D == 2 && sizeof(CT) == 4) {
    reinterpret_cast<int2*>(base)[pair] = make_int2(c[0], c[1]);
  } else if constexpr (D == 2 && sizeof(CT) == 2) {
    reinterpret_cast<short2*>(base)[pair] = make_short2((short)c[0], (short)c[1]);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) base[pair * D + k] = (CT)c[k];
  }
}

// The kernel is issue-bound (Philox rounds and the rejection loop run once per WARP until its slowest lane
// accepts), so: index arithmetic in 32 bits whenever the pair count allows (IT = unsigned), every Philox block
// of the offset stream serves as many attempts as it has words for (two in 2-D), one vector store per pair.
template <int D, typename CT, typename IT>
__global__ void __launch_bounds__(256)
sample_pairs_kernel(CT* __restrict__ anchors, CT* __restrict__ refs, int batch, IT num_anchors, IT num_refs,
                    int lo, int ext0, int ext1, int ext2, int kap, double kappa2, uint64_t seed, uint64_t sequence) {
  const Philox rng(seed);
  const IT P = num_anchors * num_refs;
  const IT total = (IT)batch * P;
  const IT gs = (IT)gridDim.x * blockDim.x;
  const int ext[3] = {ext0, ext1, ext2};
  const int k2 = (int)ceil(kappa2) - 1;  // integer s2 < kappa^2  <=>  s2 <= ceil(kappa^2) - 1
  constexpr int ATTEMPTS = 4 / D;        // attempts served by one 4-word Philox block
  for (IT g = (IT)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gs) {
    const IT b = g / P;
    const IT a = (g - b * P) / num_refs;
    // anchor: one Philox block per (sample, anchor)
    const uint4 ra = rng((uint64_t)b * (uint64_t)num_anchors + (uint64_t)a, sequence * 2);
    const uint32_t rr[4] = {ra.x, ra.y, ra.z, ra.w};
    int anc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int span = ext[k] - 2 * lo + 1;  // inclusive range [lo, ext - lo]
      anc[k] = lo + (int)bounded(rr[k], (uint32_t)span);
    }
    // offset: rejection sampling over the words of successive Philox blocks
    int off[D];
    bool accepted = false;
    for (uint32_t block = 0; !accepted; ++block) {
      const uint4 ro = rng((uint64_t)g, sequence * 2 + 1 + ((uint64_t)block << 32));
      const uint32_t r4[4] = {ro.x, ro.y, ro.z, ro.w};
#pragma unroll
      for (int t = 0; t < ATTEMPTS; ++t) {
        int cand[D], s2 = 0, s1 = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          cand[k] = (int)bounded(r4[t * D + k], (uint32_t)(2 * kap + 1)) - kap;
          s2 += cand[k] * cand[k];
          s1 |= cand[k];
        }
        if (!accepted && s2 <= k2 && s1 != 0) {
          accepted = true;
#pragma unroll
          for (int k = 0; k < D; ++k) off[k] = cand[k];
        }
      }
    }
    int ref[D];
#pragma unroll
    for (int k = 0; k < D; ++k) ref[k] = anc[k] + off[k];
    store_coord<D, CT>(anchors, (size_t)g, anc);
    store_coord<D, CT>(refs, (size_t)g, ref);
  }
}

template <int D, typename CT>
static int launch_sampler(void* anchors, void* refs, int batch, const int64_t* extent, double kappa,
                          int64_t num_anchors, int num_refs, uint64_t seed, uint64_t sequence, cudaStream_t st) {
  const int kap = (int)kappa;  // numpy truncates the float bounds
  int ext[3] = {1, 1, 1};
  for (int k = 0; k < D; ++k) {
    if (extent[k] - 2 * (int64_t)kap + 1 <= 0 || extent[k] > INT32_MAX) return CB200_EINVAL;
    ext[k] = (int)extent[k];
  }
  if (kap < 1 || !(kappa * kappa > 1.0)) return CB200_EINVAL;  // the ball must contain a non-zero offset
  const int64_t total = (int64_t)batch * num_anchors * num_refs;
  if (total == 0) return CB200_OK;
  const bool aligned = (reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(refs)) % (2 * sizeof(CT)) == 0;
  if (D == 2 && !aligned) return CB200_EINVAL;  // vector stores of (x, y)
  const int blocks = grid_for(total, 256, 2, 16);
  if (total < ((int64_t)1 << 31))
    sample_pairs_kernel<D, CT, unsigned><<<blocks, 256, 0, st>>>(
        (CT*)anchors, (CT*)refs, batch, (unsigned)num_anchors, (unsigned)num_refs, kap, ext[0], ext[1], ext[2], kap,
        kappa * kappa, seed, sequence);
  else
    sample_pairs_kernel<D, CT, unsigned long long><<<blocks, 256, 0, st>>>(
        (CT*)anchors, (CT*)refs, batch, (unsigned long long)num_anchors, (unsigned long long)num_refs, kap, ext[0],
        ext[1], ext[2], kap, kappa * kappa, seed, sequence);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" int cb200_sample_pairs(void* anchors, void* refs, int coord_dtype, int batch, int num_dims,
                                  const int64_t* extent, double kappa, int64_t num_anchors, int num_references,
                                  uint64_t seed, uint64_t sequence, void* stream) {
  if (!anchors || !refs || !extent || batch <= 0 || num_anchors < 0 || num_references < 0) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_SAMPLE(DD)                                                                                                   \
  switch (coord_dtype) {                                                                                                   \
    case CB200_I64: return launch_sampler<DD, long long>(anchors, refs, batch, extent, kappa, num_anchors, num_references, seed, sequence, st); \
    case CB200_I32: return launch_sampler<DD, int>(anchors, refs, batch, extent, kappa, num_anchors, num_references, seed, sequence, st);       \
    case CB200_I16: return launch_sampler<DD, short>(anchors, refs, batch, extent, kappa, num_anchors, num_references, seed, sequence, st);     \
  }                                                                                                                        \
  return CB200_EUNSUPPORTED;
  if (num_dims == 2) { CB200_SAMPLE(2) }
  if (num_dims == 3) { CB200_SAMPLE(3) }
#undef CB200_SAMPLE
  return CB200_EUNSUPPORTED;
}
