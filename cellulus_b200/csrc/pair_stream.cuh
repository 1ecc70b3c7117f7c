// The device pair stream: which (anchor, reference) pairs a (seed, sequence) names.
//
// B200-native counterpart of ZarrDataset.sample_coordinates / sample_offsets_within_radius
// (datasets/zarr_dataset.py:177-251).  Same DISTRIBUTION as the reference, a counter-based stream instead of
// numpy's global generator, so any pair can be produced anywhere without state:
//
//   anchor (b, a), a < A:    r = Philox4x32-10(key = seed, counter = b*A + a, stream = 2*sequence)
//                            column k = trunc(kappa) + mulhi(r[k], extent_k - 2 trunc(kappa) + 1)
//                            = np.random.randint(kappa, extent_k - kappa + 1)            (:203-218)
//   offset of pair (b, a, t), t < R:
//                            r = Philox4x32-10(key = seed, counter = (b*A + a)*ceil(R/4) + t/4, stream = 2*sequence + 1)
//                            offset = TABLE[mulhi(r[t % 4], |TABLE|)]
//   TABLE = the integer points o of [-trunc(kappa), trunc(kappa)]^D with sum o^2 < kappa^2 and o != 0,
//           enumerated with column 0 fastest.  A uniform draw from TABLE is exactly what the reference's
//           rejection filter (`in_circle`, `not_zero`, :179-196) leaves of its i.i.d. uniform candidates.
//   pair index inside the sample = a*R + t (np.repeat of the anchors, :236); reference = anchor + offset.
//
// Two kernels consume the stream: sample_pairs_kernel (sampler.cu) writes it out as coordinate lists, and
// oce_loss_sampled_kernel (oce_sampled.cu) evaluates the loss on it without the lists ever existing.
// oracle/device_sampler.py restates the stream in numpy; tests hold both kernels to it bit for bit.
#pragma once

#include "common.cuh"

namespace cb200 {

struct PairStreamParams {
  unsigned num_anchors;   // A
  unsigned num_refs;      // R
  unsigned n_tg;          // ceil(R / 4): Philox blocks per anchor in the offset stream
  unsigned n_table;       // admissible offsets
  unsigned n_cand;        // (2 kap + 1)^D candidates the table is filtered from
  int kap;                // trunc(kappa)
  int k2;                 // integer s2 < kappa^2  <=>  s2 <= k2
  int span[3];            // extent_k - 2 kap + 1 per COLUMN (x, y[, z])
  uint64_t seed, sequence;
};

// host: fill everything but seed / sequence; returns false when the parameters cannot be sampled
inline bool pair_stream_plan(PairStreamParams& p, int num_dims, const int64_t* extent, double kappa, int64_t num_anchors,
                             int64_t num_refs) {
  const int kap = (int)kappa;  // numpy truncates the float bounds
  if (kap < 1 || kap > 127 || !(kappa * kappa > 1.0)) return false;  // the ball must hold a non-zero offset; int8 table
  if (num_anchors < 0 || num_refs < 0 || num_anchors >= ((int64_t)1 << 31) || num_refs >= ((int64_t)1 << 24)) return false;
  p.kap = kap;
  p.k2 = (int)ceil(kappa * kappa) - 1;
  for (int k = 0; k < 3; ++k) p.span[k] = 1;
  for (int k = 0; k < num_dims; ++k) {
    const int64_t span = extent[k] - 2 * (int64_t)kap + 1;
    if (span <= 0 || extent[k] > INT32_MAX) return false;
    p.span[k] = (int)span;
  }
  const int side = 2 * kap + 1;
  int64_t n_cand = 1;
  for (int k = 0; k < num_dims; ++k) n_cand *= side;
  int64_t n_table = 0;
  for (int64_t c = 0; c < n_cand; ++c) {
    int64_t rest = c;
    int s2 = 0, any = 0;
    for (int k = 0; k < num_dims; ++k) {
      const int o = (int)(rest % side) - kap;
      rest /= side;
      s2 += o * o;
      any |= o;
    }
    if (s2 <= p.k2 && any != 0) ++n_table;
  }
  if (n_table <= 0) return false;
  p.n_cand = (unsigned)n_cand;
  p.n_table = (unsigned)n_table;
  p.num_anchors = (unsigned)num_anchors;
  p.num_refs = (unsigned)num_refs;
  p.n_tg = (unsigned)((num_refs + 3) / 4);
  return true;
}

constexpr size_t PAIR_TABLE_MAX_BYTES = 160 * 1024;  // dynamic shared memory the table may take

// ordered compaction of the admissible offsets into shared memory (every thread of the block calls this)
template <int D>
__device__ __forceinline__ void build_offset_table(uint32_t* __restrict__ table, const PairStreamParams& p) {
  __shared__ int s_warp_total[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = (blockDim.x + 31) >> 5;
  const int side = 2 * p.kap + 1;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (unsigned c0 = 0; c0 < p.n_cand; c0 += blockDim.x) {
    const unsigned c = c0 + tid;
    bool ok = false;
    uint32_t packed = 0;
    if (c < p.n_cand) {
      unsigned rest = c;
      int s2 = 0, any = 0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const int o = (int)(rest % (unsigned)side) - p.kap;
        rest /= (unsigned)side;
        s2 += o * o;
        any |= o;
        packed |= ((uint32_t)o & 0xffu) << (8 * k);
      }
      ok = s2 <= p.k2 && any != 0;
    }
    const unsigned vote = __ballot_sync(FULL, ok);
    if (lane == 0) s_warp_total[warp] = __popc(vote);
    __syncthreads();
    int base = s_base;
    for (int w = 0; w < warp; ++w) base += s_warp_total[w];
    if (ok) table[base + __popc(vote & ((1u << lane) - 1u))] = packed;
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int w = 0; w < n_warps; ++w) total += s_warp_total[w];
      s_base += total;
    }
    __syncthreads();
  }
}

template <int D>
__device__ __forceinline__ void stream_anchor(const Philox& rng, const PairStreamParams& p, unsigned b, unsigned a,
                                              int (&anc)[D]) {
  const uint4 r = rng((uint64_t)b * p.num_anchors + a, p.sequence * 2);
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < D; ++k) anc[k] = p.kap + (int)bounded(rr[k], (uint32_t)p.span[k]);
}

// the Philox block that holds the offset words of pairs t = 4 tg .. 4 tg + 3 of anchor (b, a)
__device__ __forceinline__ uint4 stream_offset_block(const Philox& rng, const PairStreamParams& p, unsigned b, unsigned a,
                                                     unsigned tg) {
  return rng(((uint64_t)b * p.num_anchors + a) * p.n_tg + tg, p.sequence * 2 + 1);
}

__device__ __forceinline__ uint32_t pick_word(const uint4& r, unsigned j) {
  return j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
}

template <int D>
__device__ __forceinline__ void stream_offset(const uint32_t* __restrict__ table, const PairStreamParams& p,
                                              uint32_t word, int (&off)[D]) {
  const uint32_t packed = table[bounded(word, p.n_table)];
#pragma unroll
  for (int k = 0; k < D; ++k) off[k] = (int)(signed char)(packed >> (8 * k));
}

// one coordinate tuple as a single store where the type allows it
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
template <int D>
__device__ __forceinline__ void store_coord_dyn(void* base, int dtype, size_t pair, const int (&c)[D]) {
  if (dtype == CB200_I64) store_coord<D, long long>((long long*)base, pair, c);
  else if (dtype == CB200_I32) store_coord<D, int>((int*)base, pair, c);
  else store_coord<D, short>((short*)base, pair, c);
}

}  // namespace cb200
