// Device pair sampler for sm_100a: writes the pair stream of pair_stream.cuh out as coordinate lists -- the
// B200-native counterpart of ZarrDataset.sample_coordinates / sample_offsets_within_radius
// (datasets/zarr_dataset.py:177-251) for callers that want the lists (the fused kernel of oce_sampled.cu
// consumes the same stream without them).
//
// Distribution (identical to the reference's):
//   anchor column k ~ U{trunc(kappa) .. extent_k - trunc(kappa)}  (np.random.randint(kappa, out-kappa+1))
//   offset ~ uniform over integer o in [-trunc(kappa), trunc(kappa)]^D with sum o^2 < kappa^2, o != 0
//            (what `in_circle` / `not_zero` filtering of i.i.d. draws yields), drawn by ONE bounded index into
//            the table of admissible offsets -- no rejection loop, one Philox block per four (eight) pairs
//   each anchor repeated num_references times consecutively (np.repeat, :236)
#include "pair_stream.cuh"

namespace cb200 {

// One thread per pair, coalesced vector stores.  The pairs that share an offset block recompute it (the
// kernel writes 8 - 32 bytes per pair and is bound by those stores once the rejection loop is gone).
template <int D, typename CT, typename IT>
__global__ void __launch_bounds__(256)
sample_pairs_kernel(CT* __restrict__ anchors, CT* __restrict__ refs, PairStreamParams p) {
  extern __shared__ uint32_t s_table[];
  build_offset_table<D>(s_table, p);
  const Philox rng(p.seed);
  const unsigned b = blockIdx.y;  // one grid row per sample: pair indices inside a sample fit IT
  const IT P = (IT)p.num_anchors * p.num_refs;
  const IT gs = (IT)gridDim.x * blockDim.x;
  for (IT q = (IT)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gs) {
    const unsigned a = (unsigned)(q / p.num_refs);
    const unsigned t = (unsigned)(q - (IT)a * p.num_refs);
    int anc[D], off[D], ref[D];
    stream_anchor<D>(rng, p, b, a, anc);
    const unsigned per_block = 4 * p.draws, tg = t / per_block;
    const uint4 ro = stream_offset_block(rng, p, b, a, tg);
    unpack_offset<D>(stream_offset_packed(s_table, p, ro, t - tg * per_block), off);
#pragma unroll
    for (int k = 0; k < D; ++k) ref[k] = anc[k] + off[k];
    const size_t g = (size_t)b * (size_t)P + (size_t)q;
    store_coord<D, CT>(anchors, g, anc);
    store_coord<D, CT>(refs, g, ref);
  }
}

template <int D, typename CT>
static int launch_sampler(void* anchors, void* refs, int batch, const int64_t* extent, double kappa,
                          int64_t num_anchors, int num_refs, uint64_t seed, uint64_t sequence, cudaStream_t st) {
  PairStreamParams p;
  if (!pair_stream_plan(p, D, extent, kappa, num_anchors, num_refs)) return CB200_EINVAL;
  const size_t table_bytes = (size_t)p.n_table * sizeof(uint32_t);
  if (table_bytes > PAIR_TABLE_MAX_BYTES) return CB200_EUNSUPPORTED;
  p.seed = seed;
  p.sequence = sequence;
  const int64_t total = (int64_t)batch * num_anchors * num_refs;
  if (total == 0) return CB200_OK;
  const bool aligned = (reinterpret_cast<uintptr_t>(anchors) | reinterpret_cast<uintptr_t>(refs)) % (2 * sizeof(CT)) == 0;
  if (D == 2 && !aligned) return CB200_EINVAL;  // vector stores of (x, y)
  if (batch > 65535) return CB200_EUNSUPPORTED;
  const int64_t P = num_anchors * (int64_t)num_refs;
  // every block builds the table once: few, long-lived blocks, about one wave of them over the batch
  int blocks = grid_for(P, 256, 8, 8);
  const int per_sample_cap = (CB200_SM_COUNT * 8 + batch - 1) / batch;
  if (blocks > per_sample_cap) blocks = per_sample_cap;
  const dim3 grid((unsigned)blocks, (unsigned)batch);
  if (P < ((int64_t)1 << 32)) {
    auto kernel = sample_pairs_kernel<D, CT, unsigned>;
    if (table_bytes > 48 * 1024)
      CB200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
    kernel<<<grid, 256, table_bytes, st>>>((CT*)anchors, (CT*)refs, p);
  } else {
    auto kernel = sample_pairs_kernel<D, CT, unsigned long long>;
    if (table_bytes > 48 * 1024)
      CB200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)table_bytes));
    kernel<<<grid, 256, table_bytes, st>>>((CT*)anchors, (CT*)refs, p);
  }
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
