// The device pair stream: which (anchor, reference) pairs a (seed, sequence) names.
//
// B200-native counterpart of ZarrDataset.sample_coordinates / sample_offsets_within_radius
// (datasets/zarr_dataset.py:177-251).  Same DISTRIBUTION as the reference, a counter-based stream instead of
// numpy's global generator, so any pair can be produced anywhere without state:
//
//   anchor (b, a), a < A:    r = Philox4x32-10(key = seed, counter = b*A + a, stream = 2*sequence)
//                            column k = trunc(kappa) + mulhi(r[k], extent_k - 2 trunc(kappa) + 1)
//                            = np.random.randint(kappa, extent_k - kappa + 1)            (:203-218)
//   offset of pair (b, a, t), t < R: one Philox block serves 4q pairs, q = 2 when |TABLE| <= 1024, else 1:
//                            r = Philox4x32-10(key = seed, counter = (b*A + a)*ceil(R/(4q)) + t/(4q), stream = 2*sequence + 1)
//                            u = t % (4q); w = r[u % 4]; if u >= 4: w = lo32(w * |TABLE|)
//                            offset = TABLE[mulhi(w, |TABLE|)]
//                            (the low half of the first product is a fresh uniform word up to a bias of
//                             |TABLE|^2 / 2^32 < 2.5e-4: two bounded draws from one 32-bit word)
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
  unsigned n_tg;          // ceil(R / (4 q)): Philox blocks per anchor in the offset stream
  unsigned draws;         // q: bounded draws taken from one 32-bit word (2 for small tables, else 1)
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
  p.draws = n_table <= 1024 ? 2u : 1u;
  p.n_tg = (unsigned)((num_refs + 4 * p.draws - 1) / (4 * p.draws));
  return true;
}

constexpr size_t PAIR_TABLE_MAX_BYTES = 160 * 1024;  // dynamic shared memory the table may take

// ordered compaction of the admissible offsets into shared memory (every thread of the block calls this):
// every warp owns a contiguous range of candidates, counts its admissible ones, and after one block-wide
// exchange of the counts writes them behind those of the warps in front
template <int D>
__device__ __forceinline__ bool offset_candidate(unsigned c, const PairStreamParams& p, uint32_t& packed) {
  const unsigned side = 2u * (unsigned)p.kap + 1u;
  unsigned rest = c;
  int s2 = 0, any = 0;
  packed = 0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int o = (int)(rest % side) - p.kap;
    rest /= side;
    s2 += o * o;
    any |= o;
    packed |= ((uint32_t)o & 0xffu) << (8 * k);
  }
  return c < p.n_cand && s2 <= p.k2 && any != 0;
}
template <int D>
__device__ __noinline__ void build_offset_table(uint32_t* __restrict__ table, const PairStreamParams& p) {
  __shared__ int s_warp_total[32];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, n_warps = (blockDim.x + 31u) >> 5;
  const unsigned per_warp = ((p.n_cand + n_warps - 1u) / n_warps + 31u) & ~31u;
  const unsigned c_begin = warp * per_warp, c_end = min(c_begin + per_warp, (p.n_cand + 31u) & ~31u);
  uint32_t packed;
  int count = 0;
#pragma unroll 1
  for (unsigned c = c_begin + lane; c < c_end; c += 32u)
    count += __popc(__ballot_sync(FULL, offset_candidate<D>(c, p, packed)));
  if (lane == 0) s_warp_total[warp] = count;
  __syncthreads();
  int base = 0;
#pragma unroll 1
  for (unsigned w = 0; w < warp; ++w) base += s_warp_total[w];
#pragma unroll 1
  for (unsigned c = c_begin + lane; c < c_end; c += 32u) {
    const bool ok = offset_candidate<D>(c, p, packed);
    const unsigned vote = __ballot_sync(FULL, ok);
    if (ok) table[base + __popc(vote & ((1u << lane) - 1u))] = packed;
    base += __popc(vote);
  }
  __syncthreads();
}

template <int D>
__device__ __forceinline__ void stream_anchor(const Philox& rng, const PairStreamParams& p, unsigned b, unsigned a,
                                              int (&anc)[D]) {
  const uint4 r = rng((uint64_t)b * p.num_anchors + a, p.sequence * 2);
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < D; ++k) anc[k] = p.kap + (int)bounded(rr[k], (uint32_t)p.span[k]);
}

// the Philox block that holds the offset words of pairs t = 4q tg .. 4q tg + 4q - 1 of anchor (b, a)
__device__ __forceinline__ uint4 stream_offset_block(const Philox& rng, const PairStreamParams& p, unsigned b, unsigned a,
                                                     unsigned tg) {
  return rng(((uint64_t)b * p.num_anchors + a) * p.n_tg + tg, p.sequence * 2 + 1);
}

__device__ __forceinline__ uint32_t pick_word(const uint4& r, unsigned j) {
  return j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w;
}

// packed table entry of pair t (u = t % (4q) inside its block): second draws reuse the low product half
__device__ __forceinline__ uint32_t stream_offset_packed(const uint32_t* __restrict__ table, const PairStreamParams& p,
                                                         const uint4& block, unsigned u) {
  uint32_t w = pick_word(block, u & 3);
  if (u >= 4) w *= p.n_table;
  return table[bounded(w, p.n_table)];
}
template <int D>
__device__ __forceinline__ void unpack_offset(uint32_t packed, int (&off)[D]) {
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
