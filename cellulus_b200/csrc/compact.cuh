// Order-preserving stream compaction skeleton shared by the detect kernels:
// count per tile -> single-block scan of the tile counts -> emit in raster order.
#pragma once
#include "common.cuh"

namespace cb200 {

// ---------------------------------------------------------------------------
// order-preserving compaction: count per tile -> scan -> emit
// ---------------------------------------------------------------------------
constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = 8;
constexpr int CMP_TILE = CMP_THREADS * CMP_ITEMS;

template <class Pred>
__global__ void __launch_bounds__(CMP_THREADS)
compact_count_kernel(Pred pred, int64_t n, int* __restrict__ tile_counts) {
  const int64_t base = (int64_t)blockIdx.x * CMP_TILE;
  int c = 0;
#pragma unroll
  for (int j = 0; j < CMP_ITEMS; ++j) {
    const int64_t i = base + j * CMP_THREADS + threadIdx.x;
    c += (i < n && pred(i)) ? 1 : 0;
  }
  c = warp_sum(c);
  __shared__ int s[CMP_THREADS / 32];
  if (lane_id() == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < CMP_THREADS / 32; ++w) t += s[w];
    tile_counts[blockIdx.x] = t;
  }
}

// single-block exclusive scan of the tile counts (a few hundred thousand at most)
static __global__ void __launch_bounds__(1024)
compact_scan_kernel(const int* __restrict__ tile_counts, int64_t n_tiles, long long* __restrict__ tile_offsets,
                    long long* __restrict__ total) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < n_tiles ? tile_counts[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(FULL, inc, o);
      if (lane_id() >= o) inc += t;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(FULL, w, o);
        if (lane_id() >= o) w += t;
      }
      s_warp[threadIdx.x] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
    const long long carry = s_carry;
    if (i < n_tiles) tile_offsets[i] = carry + warp_off + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// All CMP_ITEMS predicates of a thread are evaluated first (independent loads in flight), the per-warp
// ballots of every item go to shared memory, ONE warp scans the CMP_ITEMS x warps counts, and only then are
// the selected elements emitted -- two block barriers per 2048-element tile instead of sixteen.
template <class Pred, class Emit>
__global__ void __launch_bounds__(CMP_THREADS)
compact_emit_kernel(Pred pred, Emit emit, int64_t n, const long long* __restrict__ tile_offsets, int64_t capacity) {
  constexpr int WARPS = CMP_THREADS / 32;
  const int64_t base = (int64_t)blockIdx.x * CMP_TILE;
  __shared__ int s_cnt[CMP_ITEMS * WARPS];  // [item][warp] -> exclusive prefix after the scan
  const int warp = threadIdx.x >> 5, lane = lane_id();
  unsigned ballots[CMP_ITEMS];
  bool p[CMP_ITEMS];
#pragma unroll
  for (int j = 0; j < CMP_ITEMS; ++j) {
    const int64_t i = base + j * CMP_THREADS + threadIdx.x;
    p[j] = (i < n) && pred(i);
  }
#pragma unroll
  for (int j = 0; j < CMP_ITEMS; ++j) {
    ballots[j] = __ballot_sync(FULL, p[j]);
    if (lane == 0) s_cnt[j * WARPS + warp] = __popc(ballots[j]);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of CMP_ITEMS * WARPS (= 64) counts: two per lane
    static_assert(CMP_ITEMS * WARPS == 64, "scan below assumes 64 counts");
    const int a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
    int inc = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    const int excl = inc - a - b;
    s_cnt[2 * lane] = excl;
    s_cnt[2 * lane + 1] = excl + a;
  }
  __syncthreads();
  const long long tile_off = tile_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < CMP_ITEMS; ++j) {
    if (p[j]) {
      const int64_t i = base + j * CMP_THREADS + threadIdx.x;
      const long long dst = tile_off + s_cnt[j * WARPS + warp] + __popc(ballots[j] & ((1u << lane) - 1u));
      if (dst < capacity) emit(i, dst);
    }
  }
}

struct CompactWorkspace {  // layout inside the caller's buffer
  static int64_t bytes(int64_t n) {
    const int64_t tiles = (n + CMP_TILE - 1) / CMP_TILE;
    return tiles * (int64_t)(sizeof(long long) + sizeof(int)) + 64;
  }
};

template <class Pred, class Emit>
static int run_compaction(Pred pred, Emit emit, int64_t n, int64_t capacity, long long* n_out, void* workspace,
                          cudaStream_t st) {
  const int64_t tiles = (n + CMP_TILE - 1) / CMP_TILE;
  if (tiles == 0) {
    CB200_CUDA_TRY(cudaMemsetAsync(n_out, 0, sizeof(long long), st));
    return CB200_OK;
  }
  if (tiles > INT32_MAX) return CB200_EUNSUPPORTED;
  long long* tile_offsets = static_cast<long long*>(workspace);
  int* tile_counts = reinterpret_cast<int*>(tile_offsets + tiles);
  compact_count_kernel<Pred><<<(int)tiles, CMP_THREADS, 0, st>>>(pred, n, tile_counts);
  CB200_LAUNCH_CHECK();
  compact_scan_kernel<<<1, 1024, 0, st>>>(tile_counts, tiles, tile_offsets, n_out);
  CB200_LAUNCH_CHECK();
  compact_emit_kernel<Pred, Emit><<<(int)tiles, CMP_THREADS, 0, st>>>(pred, emit, n, tile_offsets, capacity);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}


}  // namespace cb200
