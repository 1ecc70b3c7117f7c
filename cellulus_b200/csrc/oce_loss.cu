// Fused object-centric-embedding loss for sm_100a.
//
// One pass over the pair lists does what the reference does in ~25 eager
// kernels (cellulus/train.py:169-178): gather offsets at the anchor and the
// reference pixel and add the integer coordinate (models/unet.py:108-124),
// the pair terms of OCELoss.forward (criterions/oce_loss.py:53-63), and the
// whole backward: d loss / d ea = (2/T) e^{-d^2/T} (ea - er) + w ea/||ea||
// scatter-added onto the anchor pixel (the reference side is detached).
//
// HBM traffic is the two coordinate lists (read once, streaming, 16-byte
// vector loads for 2-D int64) + the offsets tensor + the dense gradient; the
// pixel gathers and the gradient atomics are L2 traffic.  Consecutive pairs
// usually share their anchor (np.repeat in zarr_dataset.py:236), so gradients
// are first combined by a warp-level SEGMENTED reduction over runs of equal
// anchor pixels and only run heads issue red.global.add -- ~16x fewer atomics.
// Loss terms are summed per thread in fp32 (tens of terms), then in fp64
// through warp shuffles, shared memory and one fp64 atomic per block.
#include <cstdlib>

#include "loss_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace cb200 {

#define CB200_LOSS_UNROLL 1
#ifndef CB200_LOSS_MIN_BLOCKS
#define CB200_LOSS_MIN_BLOCKS 6  // <= 40 registers: 48 resident warps per SM (measured best, tools/loss_sweep.py)
#endif
constexpr int LOSS_THREADS = 256;
constexpr int LOSS_UNROLL = CB200_LOSS_UNROLL;
constexpr int LOSS_MIN_BLOCKS = CB200_LOSS_MIN_BLOCKS;

// ---- raw coordinate registers: what the load returns, converted only at first use, so that the
// software prefetch of the NEXT iteration's coordinates never stalls on its own loads ----------
template <int D, typename CT>
struct RawCoord;
template <>
struct RawCoord<2, long long> {
  longlong2 v;
  __device__ __forceinline__ void load(const long long* base, unsigned pair) { v = ld_stream_ll2(base + (size_t)pair * 2); }
  __device__ __forceinline__ void zero() { v.x = v.y = 0; }
  __device__ __forceinline__ int get(int k) const { return (int)(k == 0 ? v.x : v.y); }
};
template <>
struct RawCoord<2, int> {
  int2 v;
  __device__ __forceinline__ void load(const int* base, unsigned pair) { v = __ldg(reinterpret_cast<const int2*>(base) + pair); }
  __device__ __forceinline__ void zero() { v.x = v.y = 0; }
  __device__ __forceinline__ int get(int k) const { return k == 0 ? v.x : v.y; }
};
template <>
struct RawCoord<2, short> {
  short2 v;
  __device__ __forceinline__ void load(const short* base, unsigned pair) { v = __ldg(reinterpret_cast<const short2*>(base) + pair); }
  __device__ __forceinline__ void zero() { v.x = v.y = 0; }
  __device__ __forceinline__ int get(int k) const { return k == 0 ? v.x : v.y; }
};
template <typename CT>
struct RawCoord<3, CT> {
  CT v[3];
  __device__ __forceinline__ void load(const CT* base, unsigned pair) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if constexpr (sizeof(CT) == 8) v[k] = (CT)ld_stream_ll(base + (size_t)pair * 3 + k);
      else v[k] = __ldg(base + (size_t)pair * 3 + k);
    }
  }
  __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = 0; }
  __device__ __forceinline__ int get(int k) const { return (int)v[k]; }
};

// ---- one chunk = 32 consecutive pairs of one sample, one pair per lane -----------------------------
template <int D>
struct Chunk {
  int ca[D], cr[D];  // anchor / reference coordinates as delivered
  float oa[D], orf[D];  // gathered offsets
  int pix_a;            // anchor pixel (run key); negative for dead lanes
  bool live;
};

// Stage 1: range-check (fast path: one unsigned compare per coordinate; the rare negative index takes the
// wrapping slow path, as torch advanced indexing would) and ISSUE the gathers.  Nothing here waits.
template <int D, typename OT, bool IL>
__device__ __forceinline__ void chunk_gather(Chunk<D>& c, const OT* __restrict__ offsets, unsigned first,
                                             const Shape<D>& shape, int& bad) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k)
    ok = ok && ((unsigned)c.ca[k] < (unsigned)shape.ext[k]) && ((unsigned)c.cr[k] < (unsigned)shape.ext[k]);
  int pa, pr;
  if (__builtin_expect(__any_sync(FULL, c.live && !ok), 0)) {
    int wa[D], wr[D];
    ok = wrap_and_check<D>(c.ca, wa, shape) & wrap_and_check<D>(c.cr, wr, shape);
    pa = pixel_of<D>(wa, shape);
    pr = pixel_of<D>(wr, shape);
    if (c.live && !ok) ++bad;
  } else {
    pa = pixel_of<D>(c.ca, shape);
    pr = pixel_of<D>(c.cr, shape);
  }
  c.live = c.live && ok;
  c.pix_a = c.live ? pa : -1;  // dead lanes never merge with a live neighbour
  if (c.live) {
    gather_pixel<D, OT, IL>(offsets, (unsigned)shape.npix, first, (unsigned)pa, c.oa);
    gather_pixel<D, OT, IL>(offsets, (unsigned)shape.npix, first, (unsigned)pr, c.orf);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) c.oa[k] = c.orf[k] = 0.f;
  }
}

// Stage 2: the pair terms, the segmented warp sum of the gradients over runs of equal anchor pixel, and
// one reduction per run.
template <int D, bool BWD, bool IL>
__device__ __forceinline__ void chunk_finish(const Chunk<D>& c, float* __restrict__ grad, unsigned first,
                                             const Shape<D>& shape, float neg_log2e_over_t, float two_over_t, float w, unsigned lane,
                                             float& acc_oce, float& acc_nrm) {
  float g[D], diff[D], ea[D];
  float d2 = 0.f, n2 = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    ea[k] = __fadd_rn(c.oa[k], (float)c.ca[k]);  // selection += coordinate (models/unet.py:120)
    const float er = __fadd_rn(c.orf[k], (float)c.cr[k]);
    diff[k] = ea[k] - er;
    d2 = fmaf(diff[k], diff[k], d2);
    n2 = fmaf(ea[k], ea[k], n2);
  }
  const float e = ex2_approx(d2 * neg_log2e_over_t);   // exp(-d^2 / T)
  const float rs = n2 > 0.f ? rsqrt_approx(n2) : 0.f;  // 1 / ||ea||, 0 at the origin (torch's norm backward)
  if (c.live) {
    acc_oce += 1.0f - e;
    acc_nrm = fmaf(n2, rs, acc_nrm);  // ||ea||
  }
  if constexpr (BWD) {
    const float ge = two_over_t * e;
    const float gr = w * rs;
#pragma unroll
    for (int k = 0; k < D; ++k) g[k] = c.live ? fmaf(ge, diff[k], gr * ea[k]) : 0.f;
    const int key = c.pix_a;
    const int prev = __shfl_up_sync(FULL, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const unsigned heads = __ballot_sync(FULL, head);
    const unsigned above = heads & (0xfffffffeu << lane);
    const unsigned reach = (above ? (unsigned)(__ffs(above) - 1) : 32u) - lane;  // lanes left in this run, >= 1
#pragma unroll
    for (unsigned o = 1; o < 32; o <<= 1) {
      const bool take = o < reach;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const float v = __shfl_down_sync(FULL, g[k], o);
        if (take) g[k] += v;
      }
    }
    if (head && c.live) scatter_pixel<D, IL>(grad, (unsigned)shape.npix, first, (unsigned)key, g);
  }
}

// ---- the fused kernel ------------------------------------------------------------------------------
// One warp owns one chunk per iteration.  The coordinates of the chunk after next are always in flight from
// HBM (streaming 16-byte loads into a raw register buffer that is converted only when taken), so the list
// latency hides behind the gathers and the math; the L2 gather latency is covered by the 48 resident warps
// per SM.  (Issuing the next chunk's gathers before the current chunk's math, and a warp-specialised
// cp.async.bulk ring for the lists, were both built and measured SLOWER: see profiles/README.md.)
template <int D, typename CT, typename OT, bool BWD, bool IL>
__global__ void __launch_bounds__(LOSS_THREADS, LOSS_MIN_BLOCKS)
oce_loss_fused_kernel(const OT* __restrict__ offsets, const CT* __restrict__ anchors, const CT* __restrict__ refs,
                      unsigned P, unsigned chunks_per_sample, Shape<D> shape, float neg_log2e_over_t, float two_over_t,
                      float w, float* __restrict__ grad, LossWorkspace* ws, float* out) {
  const unsigned lane = lane_id();
  const unsigned b = blockIdx.y;
  const unsigned stride = gridDim.x * (LOSS_THREADS / 32);
  const CT* __restrict__ a_base = anchors + (size_t)b * P * D;
  const CT* __restrict__ r_base = refs + (size_t)b * P * D;
  const unsigned first = b * (unsigned)shape.npix;  // the sample's first pixel in the whole tensor

  float acc_oce = 0.f, acc_nrm = 0.f;
  int bad = 0;

  RawCoord<D, CT> na, nr;  // coordinates of the chunk after next, as loaded
  auto fetch = [&](unsigned c) {
    const unsigned p = (c << 5) + lane;
    if (c < chunks_per_sample && p < P) {  // otherwise stale values, never used: the lane is not live
      na.load(a_base, p);
      nr.load(r_base, p);
    }
  };
  auto take = [&](Chunk<D>& ch, unsigned c) {
    ch.live = c < chunks_per_sample && ((c << 5) + lane) < P;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      ch.ca[k] = na.get(k);
      ch.cr[k] = nr.get(k);
    }
  };

  unsigned c0 = blockIdx.x * (LOSS_THREADS / 32) + (threadIdx.x >> 5);
  if (c0 < chunks_per_sample) {
    Chunk<D> cur, nxt;
    na.zero();
    nr.zero();
    fetch(c0);
    take(cur, c0);
    fetch(c0 + stride);
    chunk_gather<D, OT, IL>(cur, offsets, first, shape, bad);
    // the gradient tensor is zero-filled by the preceding grid (programmatic dependent launch): everything
    // above overlapped with it, nothing below may.  (Clearing it inside this kernel behind a grid barrier
    // was built and measured slower: 51.8 vs 47.4 us.)
    if constexpr (BWD) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (; c0 < chunks_per_sample; c0 += stride) {
      chunk_finish<D, BWD, IL>(cur, grad, first, shape, neg_log2e_over_t, two_over_t, w, lane, acc_oce, acc_nrm);
      take(nxt, c0 + stride);
      fetch(c0 + 2 * stride);
      if (c0 + stride < chunks_per_sample) chunk_gather<D, OT, IL>(nxt, offsets, first, shape, bad);
      cur = nxt;
    }
  }
  block_reduce_to_workspace(acc_oce, acc_nrm, bad, ws, w, out);
}

// ---- the fused kernel for PLANAR 2-D tensors with staging scratch -------------------------------------
// The reference's model emits planar NCHW offsets (models/unet.py:69-71); gathered in place, a scattered reference
// pixel costs one L1 wavefront per lane PER CHANNEL (57 us per configs[1] step against 47 us channels-last).  This
// form prepares its own operands inside the launch: the CTAs of a sample write a channels-last copy of that
// sample's offsets into caller scratch (`staged`; prep & 4: two pixels per thread, 8-byte loads, 16-byte stores)
// and clear the sample's gradient (prep & 3: 1 = scalar stores, 2 = 16-byte stores) between them; each CTA then
// counts itself in and, with its first list loads in flight, waits until every CTA of its sample slot has
// arrived.  All CTAs of the grid are co-resident by construction (one wave, checked by the launcher), so the wait
// cannot starve.  From there on it is the kernel above with 8-byte gathers from the copy; the gradient -- one
// reduction per run of equal anchors -- stays planar.  ONE launch per step.
// (The same in-kernel clearing for the other layouts was built and measured: 49.5 us against 49.4 us behind the
// zero-fill grid with an identical body -- programmatic dependent launch already hides the zero-fill.)
template <typename OT>
__device__ __forceinline__ void chunk_gather_staged(Chunk<2>& c, const OT* staged, unsigned first, const Shape<2>& shape,
                                                    int& bad) {
  constexpr int D = 2;
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k)
    ok = ok && ((unsigned)c.ca[k] < (unsigned)shape.ext[k]) && ((unsigned)c.cr[k] < (unsigned)shape.ext[k]);
  int pa, pr;
  if (__builtin_expect(__any_sync(FULL, c.live && !ok), 0)) {
    int wa[D], wr[D];
    ok = wrap_and_check<D>(c.ca, wa, shape) & wrap_and_check<D>(c.cr, wr, shape);
    pa = pixel_of<D>(wa, shape);
    pr = pixel_of<D>(wr, shape);
    if (c.live && !ok) ++bad;
  } else {
    pa = pixel_of<D>(c.ca, shape);
    pr = pixel_of<D>(c.cr, shape);
  }
  c.live = c.live && ok;
  c.pix_a = c.live ? pa : -1;
  if (c.live) {
    gather_staged<D, OT>(staged, first, (unsigned)pa, c.oa);
    gather_staged<D, OT>(staged, first, (unsigned)pr, c.orf);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) c.oa[k] = c.orf[k] = 0.f;
  }
}

template <typename CT, typename OT, bool BWD>
__global__ void __launch_bounds__(LOSS_THREADS, LOSS_MIN_BLOCKS)
oce_loss_staged_kernel(const OT* __restrict__ offsets, const CT* __restrict__ anchors, const CT* __restrict__ refs,
                       unsigned P, unsigned chunks_per_sample, Shape<2> shape, float neg_log2e_over_t, float two_over_t,
                       float w, float* __restrict__ grad, LossWorkspace* ws, float* out, unsigned prep, OT* staged) {
  constexpr int D = 2;
  const unsigned lane = lane_id();
  const unsigned b = blockIdx.y;
  const unsigned stride = gridDim.x * (LOSS_THREADS / 32);
  const unsigned npix = (unsigned)shape.npix;
  if constexpr (BWD) {
    const unsigned n = npix * D;  // floats per sample (< 2^31: make_shape)
    float* __restrict__ g = grad + (size_t)b * n;
    const unsigned per = ((n + gridDim.x - 1) / gridDim.x + 3u) & ~3u;
    const unsigned lo = min(n, blockIdx.x * per), hi = min(n, lo + per);
    if ((prep & 3u) == 2u) {
      float4* __restrict__ g4 = reinterpret_cast<float4*>(g + lo);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (unsigned i = threadIdx.x; i < ((hi - lo) >> 2); i += LOSS_THREADS) g4[i] = z;
    } else {
      for (unsigned i = lo + threadIdx.x; i < hi; i += LOSS_THREADS) g[i] = 0.f;
    }
  }
  {
    // planar (2, npix) -> interleaved (npix, 2) for this CTA's slice of the sample's pixels
    const OT* __restrict__ src = offsets + (size_t)b * npix * 2;
    OT* __restrict__ dst = staged + (size_t)b * npix * 2;
    const unsigned per = ((npix + gridDim.x - 1) / gridDim.x + 1u) & ~1u;
    const unsigned lo = min(npix, blockIdx.x * per), hi = min(npix, lo + per);
    bool done = false;
    if constexpr (sizeof(OT) == 4) {
      if ((prep & 4u) != 0) {  // even npix, aligned bases
        for (unsigned i = lo + 2 * threadIdx.x; i + 1 < hi; i += 2 * LOSS_THREADS) {
          const float2 x = __ldg(reinterpret_cast<const float2*>(src + i));
          const float2 y = __ldg(reinterpret_cast<const float2*>(src + npix + i));
          *reinterpret_cast<float4*>(dst + 2 * (size_t)i) = make_float4(x.x, y.x, x.y, y.y);
        }
        done = true;
      }
    }
    if (!done) {
      for (unsigned i = lo + threadIdx.x; i < hi; i += LOSS_THREADS) {
        dst[2 * (size_t)i] = src[i];
        dst[2 * (size_t)i + 1] = src[npix + i];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&ws->arrived[(b % LOSS_ZERO_SLOTS) * LOSS_ZERO_PITCH]) : "memory");

  const CT* __restrict__ a_base = anchors + (size_t)b * P * D;
  const CT* __restrict__ r_base = refs + (size_t)b * P * D;
  const unsigned first = b * npix;
  float acc_oce = 0.f, acc_nrm = 0.f;
  int bad = 0;
  RawCoord<D, CT> na, nr;
  auto fetch = [&](unsigned c) {
    const unsigned p = (c << 5) + lane;
    if (c < chunks_per_sample && p < P) {
      na.load(a_base, p);
      nr.load(r_base, p);
    }
  };
  auto take = [&](Chunk<D>& ch, unsigned c) {
    ch.live = c < chunks_per_sample && ((c << 5) + lane) < P;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      ch.ca[k] = na.get(k);
      ch.cr[k] = nr.get(k);
    }
  };
  unsigned c0 = blockIdx.x * (LOSS_THREADS / 32) + (threadIdx.x >> 5);
  const bool has_work = c0 < chunks_per_sample;
  Chunk<D> cur, nxt;
  na.zero();
  nr.zero();
  if (has_work) {
    fetch(c0);
    take(cur, c0);
    fetch(c0 + stride);
  }
  // every CTA of the sample slot has written its slice: one poller per CTA, the other warps park on the barrier
  if (threadIdx.x == 0) {
    const unsigned slot = b % LOSS_ZERO_SLOTS;
    const unsigned target = gridDim.x * ((gridDim.y - slot + LOSS_ZERO_SLOTS - 1) / LOSS_ZERO_SLOTS);
    const unsigned* flag = &ws->arrived[slot * LOSS_ZERO_PITCH];
    unsigned seen;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
      if (seen >= target) break;
      __nanosleep(40);
    }
  }
  __syncthreads();
  if (has_work) {
    chunk_gather_staged<OT>(cur, staged, first, shape, bad);
    for (; c0 < chunks_per_sample; c0 += stride) {
      chunk_finish<D, BWD, false>(cur, grad, first, shape, neg_log2e_over_t, two_over_t, w, lane, acc_oce, acc_nrm);
      take(nxt, c0 + stride);
      fetch(c0 + 2 * stride);
      if (c0 + stride < chunks_per_sample) chunk_gather_staged<OT>(nxt, staged, first, shape, bad);
      cur = nxt;
    }
  }
  block_reduce_to_workspace(acc_oce, acc_nrm, bad, ws, w, out);
}

// zero-fill of the gradient tensor: 4 independent 16-byte stores per thread per trip, one wave
__global__ void __launch_bounds__(256) zero_fill_kernel(float4* __restrict__ p, int64_t n4, float* __restrict__ tail,
                                                        int n_tail) {
  // programmatic dependent launch: the fused loss kernel may start its prologue (list prefetch, first
  // gathers) while this grid is still writing zeros; it waits (griddepcontrol.wait) before its first reduction
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    p[i] = z;
    p[i + stride] = z;
    p[i + 2 * stride] = z;
    p[i + 3 * stride] = z;
  }
  for (; i < n4; i += stride) p[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] = 0.f;
}
int zero_fill(float* p, int64_t n, cudaStream_t st) {
  if (n <= 0) return CB200_OK;
  if (reinterpret_cast<uintptr_t>(p) & 15) {
    CB200_CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(float) * (size_t)n, st));
    return CB200_OK;
  }
  const int64_t n4 = n / 4;
  zero_fill_kernel<<<grid_for(n4, 256, 4, 4), 256, 0, st>>>(reinterpret_cast<float4*>(p), n4, p + n4 * 4, (int)(n - n4 * 4));
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

// grad *= *scale, skipping all work when the upstream gradient is exactly 1
__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ g, int64_t n4, int64_t n,
                                                            const float* __restrict__ scale) {
  const float s = __ldg(scale);
  if (s == 1.0f) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = g4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    g4[i] = v;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] *= s;
}

// ---------------------------------------------------------------------------
// Unfused drop-ins (three-call shape of the reference)
// ---------------------------------------------------------------------------
template <int D, typename CT, typename OT>
__global__ void __launch_bounds__(256)
gather_add_kernel(const OT* __restrict__ offsets, const CT* __restrict__ coords, int batch, int64_t P, Shape<D> shape,
                  float* __restrict__ out) {
  const int64_t total = (int64_t)batch * P;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c[D];
    load_coord<D, CT>(coords, i, c);
    int cw[D];
    const bool ok = wrap_and_check<D>(c, cw, shape);
    const int b = (int)(i / P);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      float v = __int_as_float(0x7fc00000);  // NaN marks an invalid coordinate
      if (ok) v = __fadd_rn(load_as_float<OT>(offsets, ((int64_t)b * D + k) * shape.npix + pixel_of<D>(cw, shape)),
                            (float)c[k]);
      out[i * D + k] = v;
    }
  }
}

template <int D, typename CT>
__global__ void __launch_bounds__(256)
scatter_add_kernel(const float* __restrict__ grad_out, const CT* __restrict__ coords, int batch, int64_t P,
                   Shape<D> shape, float* __restrict__ grad) {
  const int lane = lane_id();
  const int64_t total = (int64_t)batch * P;
  const int64_t n_chunks = (total + 31) >> 5;
  const int64_t warp_gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warp_stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t chunk = warp_gid; chunk < n_chunks; chunk += warp_stride) {
    const int64_t i = (chunk << 5) + lane;
    int c[D], cw[D];
    float g[D];
    int64_t key = -1 - lane;
    bool live = i < total;
    if (live) {
      load_coord<D, CT>(coords, i, c);
      live = wrap_and_check<D>(c, cw, shape);
    }
    if (live) {
      const int b = (int)(i / P);
      key = (int64_t)b * D * shape.npix + pixel_of<D>(cw, shape);
#pragma unroll
      for (int k = 0; k < D; ++k) g[k] = __ldg(grad_out + i * D + k);
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) g[k] = 0.f;
    }
    const int64_t prev = __shfl_up_sync(FULL, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const unsigned heads = __ballot_sync(FULL, head);
    const unsigned above = (lane == 31) ? 0u : (heads & (0xfffffffeu << lane));
    const int limit = above ? (__ffs(above) - 1) : 32;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const float v = __shfl_down_sync(FULL, g[k], o);
        if (lane + o < limit) g[k] += v;
      }
    }
    if (head && live) {
#pragma unroll
      for (int k = 0; k < D; ++k) atomicAdd(grad + key + k * shape.npix, g[k]);
    }
  }
}

template <int D, bool BWD>
__global__ void __launch_bounds__(LOSS_THREADS)
pair_loss_kernel(const float* __restrict__ ea_p, const float* __restrict__ er_p, int64_t n_pairs, float inv_t, float w,
                 float* __restrict__ grad_ea, LossWorkspace* ws, float* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float two_inv_t = 2.0f * inv_t;
  float acc_oce = 0.f, acc_nrm = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += stride) {
    float ea[D], diff[D], d2 = 0.f, n2 = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      ea[k] = ld_stream_f(ea_p + i * D + k);
      diff[k] = ea[k] - ld_stream_f(er_p + i * D + k);
      d2 = fmaf(diff[k], diff[k], d2);
      n2 = fmaf(ea[k], ea[k], n2);
    }
    const float e = expf(-d2 * inv_t);
    const float nrm = sqrtf(n2);
    acc_oce += 1.0f - e;
    acc_nrm += nrm;
    if constexpr (BWD) {
      const float ge = two_inv_t * e;
      const float gr = nrm > 0.f ? w / nrm : 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) grad_ea[i * D + k] = fmaf(ge, diff[k], gr * ea[k]);
    }
  }
  block_reduce_to_workspace(acc_oce, acc_nrm, 0, ws, w, out);
}

// programmatic dependent launch of the fused kernel behind the zero-fill (A/B switch: CB200_LOSS_PDL=0)
const bool g_loss_pdl = [] {
  const char* e = getenv("CB200_LOSS_PDL");
  return !(e && e[0] == '0');
}();
template <int D, typename CT, typename OT, bool BWD, bool IL>
static int launch_fused_variant(const void* offsets, const void* anchors, const void* refs, int batch,
                                const Shape<D>& shape, int64_t P, float T, float w, float* grad, float* out,
                                LossWorkspace* ws, cudaStream_t st) {
  auto kernel = oce_loss_fused_kernel<D, CT, OT, BWD, IL>;
  // persistent grid: exactly the CTAs that are co-resident (one wave), split evenly over the samples
  static int occupancy = 0;  // per template instantiation
  if (occupancy == 0) {
    int occ = 0;
    CB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, LOSS_THREADS, 0));
    occupancy = occ > 0 ? occ : 1;
  }
  const unsigned cps = (unsigned)((P + 31) >> 5);
  const unsigned per_block = (LOSS_THREADS / 32) * LOSS_UNROLL;
  unsigned blocks_x = (unsigned)((CB200_SM_COUNT * occupancy) / batch);  // floor: never spill into a second wave
  const unsigned needed = (cps + per_block - 1) / per_block;
  if (blocks_x > needed) blocks_x = needed;
  if (blocks_x < 1) blocks_x = 1;
  const float log2e = 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks_x, (unsigned)batch);
  cfg.blockDim = dim3(LOSS_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = (BWD && g_loss_pdl) ? 1 : 0;  // only behind our own zero-fill grid
  CB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, (const OT*)offsets, (const CT*)anchors, (const CT*)refs, (unsigned)P,
                                    cps, shape, -log2e / T, 2.0f / T, w, grad, ws, out));
  return CB200_OK;
}

// CB200_LOSS_STAGED=0: planar tensors are gathered in place even when the caller passes staging scratch;
// CB200_LOSS_COOP=1: cooperative instead of plain launch of the staged form (+3 us per launch, measured).  A/B switches.
static const bool g_staged = [] { const char* e = getenv("CB200_LOSS_STAGED"); return !(e && e[0] == '0'); }();
static const bool g_coop = [] { const char* e = getenv("CB200_LOSS_COOP"); return e && e[0] == '1'; }();

// returns CB200_OK with *launched = false when the staged form does not apply (grid larger than one resident wave)
template <typename CT, typename OT, bool BWD>
static int launch_staged_variant(const void* offsets, const void* anchors, const void* refs, int batch, const Shape<2>& shape,
                                 int64_t P, float T, float w, float* grad, float* out, LossWorkspace* ws, void* staged,
                                 cudaStream_t st, bool* launched) {
  auto kernel = oce_loss_staged_kernel<CT, OT, BWD>;
  static int occupancy = 0;  // per template instantiation
  if (occupancy == 0) {
    int occ = 0;
    CB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, LOSS_THREADS, 0));
    occupancy = occ > 0 ? occ : 1;
  }
  const unsigned cps = (unsigned)((P + 31) >> 5);
  unsigned blocks_x = (unsigned)((CB200_SM_COUNT * occupancy) / batch);
  const unsigned needed = (cps + (LOSS_THREADS / 32) - 1) / (LOSS_THREADS / 32);
  if (blocks_x > needed) blocks_x = needed;
  if (blocks_x < 1) blocks_x = 1;
  *launched = false;
  // the CTAs wait for each other: every one of them must be resident at once
  if ((uint64_t)blocks_x * batch > (uint64_t)CB200_SM_COUNT * occupancy) return CB200_OK;
  unsigned prep = ((shape.npix % 2 == 0) && reinterpret_cast<uintptr_t>(offsets) % 8 == 0 &&
                   reinterpret_cast<uintptr_t>(staged) % 16 == 0) ? 4u : 8u;
  if (BWD) prep |= ((reinterpret_cast<uintptr_t>(grad) % 16 == 0) && ((shape.npix * 2) % 4 == 0)) ? 2u : 1u;
  const float log2e = 1.4426950408889634f;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks_x, (unsigned)batch);
  cfg.blockDim = dim3(LOSS_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.numAttrs = g_coop ? 1 : 0;
  CB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, (const OT*)offsets, (const CT*)anchors, (const CT*)refs, (unsigned)P,
                                    cps, shape, -log2e / T, 2.0f / T, w, grad, ws, out, prep, (OT*)staged));
  *launched = true;
  return CB200_OK;
}

template <int D, typename CT, typename OT>
static int launch_fused(const void* offsets, int layout, const void* anchors, const void* refs, int batch,
                        const int64_t* spatial, int64_t P, float T, float w, float* grad, float* out, void* workspace,
                        void* staging, int64_t staging_bytes, cudaStream_t st) {
  Shape<D> shape;
  if (!make_shape<D>(spatial, shape)) return CB200_EINVAL;
  if (P >= ((int64_t)1 << 31) - 64 || batch > 65535) return CB200_EUNSUPPORTED;
  if ((int64_t)batch * D * shape.npix > INT32_MAX) return CB200_EUNSUPPORTED;  // 32-bit element indices
  auto* ws = static_cast<LossWorkspace*>(workspace);
  if constexpr (D == 2) {
    // planar tensors + caller scratch: gather from a channels-last copy made inside the kernel (one-wave grids)
    if (layout == CB200_LAYOUT_PLANAR && g_staged && staging &&
        staging_bytes >= (int64_t)sizeof(OT) * batch * D * shape.npix && reinterpret_cast<uintptr_t>(staging) % (2 * sizeof(OT)) == 0) {
      bool launched = false;
      const int rc = grad ? launch_staged_variant<CT, OT, true>(offsets, anchors, refs, batch, shape, P, T, w, grad, out, ws,
                                                               staging, st, &launched)
                          : launch_staged_variant<CT, OT, false>(offsets, anchors, refs, batch, shape, P, T, w, grad, out, ws,
                                                                staging, st, &launched);
      if (rc != CB200_OK || launched) return rc;
    }
  }
  if (grad) {
    const int rc = zero_fill(grad, (int64_t)batch * D * shape.npix, st);
    if (rc != CB200_OK) return rc;
  }
  const bool il = layout == CB200_LAYOUT_CHANNELS_LAST;
  if (il && D == 2 && (reinterpret_cast<uintptr_t>(offsets) % (2 * sizeof(OT)) || reinterpret_cast<uintptr_t>(grad) % 8))
    return CB200_EINVAL;  // vector gathers / vector reductions need pixel-aligned bases
#define CB200_FUSED(BWD, IL) \
  launch_fused_variant<D, CT, OT, BWD, IL>(offsets, anchors, refs, batch, shape, P, T, w, grad, out, ws, st)
  if (grad) return il ? CB200_FUSED(true, true) : CB200_FUSED(true, false);
  return il ? CB200_FUSED(false, true) : CB200_FUSED(false, false);
#undef CB200_FUSED
}

template <int D, typename CT>
static int dispatch_offsets(const void* offsets, int odt, int layout, const void* a, const void* r, int batch,
                            const int64_t* spatial, int64_t P, float T, float w, float* grad, float* out, void* ws,
                            void* staging, int64_t staging_bytes, cudaStream_t st) {
  if (odt == CB200_F32)
    return launch_fused<D, CT, float>(offsets, layout, a, r, batch, spatial, P, T, w, grad, out, ws, staging,
                                      staging_bytes, st);
  if (odt == CB200_BF16)
    return launch_fused<D, CT, __nv_bfloat16>(offsets, layout, a, r, batch, spatial, P, T, w, grad, out, ws, staging,
                                              staging_bytes, st);
  return CB200_EUNSUPPORTED;
}

template <int D>
static int dispatch_coords(const void* offsets, int odt, int layout, const void* a, const void* r, int cdt, int batch,
                           const int64_t* spatial, int64_t P, float T, float w, float* grad, float* out, void* ws,
                           void* sg, int64_t sgb, cudaStream_t st) {
  switch (cdt) {
    case CB200_I64: return dispatch_offsets<D, long long>(offsets, odt, layout, a, r, batch, spatial, P, T, w, grad, out, ws, sg, sgb, st);
    case CB200_I32: return dispatch_offsets<D, int>(offsets, odt, layout, a, r, batch, spatial, P, T, w, grad, out, ws, sg, sgb, st);
    case CB200_I16: return dispatch_offsets<D, short>(offsets, odt, layout, a, r, batch, spatial, P, T, w, grad, out, ws, sg, sgb, st);
  }
  return CB200_EUNSUPPORTED;
}

template <int D, typename CT>
static int launch_gather(const void* offsets, int odt, const void* coords, int batch, const int64_t* spatial, int64_t P,
                         float* out, cudaStream_t st) {
  Shape<D> shape;
  if (!make_shape<D>(spatial, shape)) return CB200_EINVAL;
  const int blocks = grid_for((int64_t)batch * P, 256, 2);
  if (odt == CB200_F32)
    gather_add_kernel<D, CT, float><<<blocks, 256, 0, st>>>((const float*)offsets, (const CT*)coords, batch, P, shape, out);
  else if (odt == CB200_BF16)
    gather_add_kernel<D, CT, __nv_bfloat16>
        <<<blocks, 256, 0, st>>>((const __nv_bfloat16*)offsets, (const CT*)coords, batch, P, shape, out);
  else
    return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

template <int D, typename CT>
static int launch_scatter(const float* grad_out, const void* coords, int batch, const int64_t* spatial, int64_t P,
                          float* grad, cudaStream_t st) {
  Shape<D> shape;
  if (!make_shape<D>(spatial, shape)) return CB200_EINVAL;
  CB200_CUDA_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)batch * D * shape.npix, st));
  const int blocks = grid_for((int64_t)batch * P, 256, 2);
  scatter_add_kernel<D, CT><<<blocks, 256, 0, st>>>(grad_out, (const CT*)coords, batch, P, shape, grad);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

// 2-D coordinate pairs are fetched with one vector load: the list must be aligned to a whole pair
static bool coords_aligned(const void* p, int coord_dtype, int num_dims) {
  const size_t esize = coord_dtype == CB200_I64 ? 8 : coord_dtype == CB200_I32 ? 4 : 2;
  const size_t align = num_dims == 2 ? 2 * esize : esize;
  return (reinterpret_cast<uintptr_t>(p) % align) == 0;
}

extern "C" {

int64_t cb200_oce_loss_workspace_bytes(void) { return (int64_t)sizeof(LossWorkspace); }


int64_t cb200_oce_loss_staging_bytes(int offsets_dtype, int offsets_layout, int batch, int num_dims, const int64_t* spatial) {
  if (!spatial || batch <= 0 || num_dims != 2 || offsets_layout != CB200_LAYOUT_PLANAR) return 0;
  const int64_t esize = offsets_dtype == CB200_F32 ? 4 : offsets_dtype == CB200_BF16 ? 2 : 0;
  int64_t n = (int64_t)batch * num_dims * esize;
  for (int k = 0; k < num_dims; ++k) {
    if (spatial[k] <= 0) return 0;
    n *= spatial[k];
  }
  return n;
}

int cb200_oce_loss_fwd_bwd(const void* offsets, int offsets_dtype, int offsets_layout, const void* anchors,
                           const void* refs, int coord_dtype, int batch, int num_dims, const int64_t* spatial, int64_t pairs_per_sample,
                           float temperature, float regularization_weight, float* grad, float* out, void* workspace,
                           void* stream) {
  return cb200_oce_loss_fwd_bwd_staged(offsets, offsets_dtype, offsets_layout, anchors, refs, coord_dtype, batch, num_dims,
                                       spatial, pairs_per_sample, temperature, regularization_weight, grad, out,
                                       workspace, nullptr, 0, stream);
}

int cb200_oce_loss_fwd_bwd_staged(const void* offsets, int offsets_dtype, int offsets_layout, const void* anchors,
                                  const void* refs, int coord_dtype, int batch, int num_dims, const int64_t* spatial,
                                  int64_t pairs_per_sample, float temperature, float regularization_weight, float* grad,
                                  float* out, void* workspace, void* staging, int64_t staging_bytes, void* stream) {
  if (!offsets || !spatial || !out || !workspace) return CB200_EINVAL;
  if (batch <= 0 || pairs_per_sample < 0 || !(temperature != 0.f)) return CB200_EINVAL;
  if (pairs_per_sample > 0 && (!anchors || !refs)) return CB200_EINVAL;  // an empty list may be a null pointer
  if (!coords_aligned(anchors, coord_dtype, num_dims) || !coords_aligned(refs, coord_dtype, num_dims)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_dims == 2)
    return dispatch_coords<2>(offsets, offsets_dtype, offsets_layout, anchors, refs, coord_dtype, batch, spatial, pairs_per_sample,
                              temperature, regularization_weight, grad, out, workspace, staging, staging_bytes, st);
  if (num_dims == 3)
    return dispatch_coords<3>(offsets, offsets_dtype, offsets_layout, anchors, refs, coord_dtype, batch, spatial, pairs_per_sample,
                              temperature, regularization_weight, grad, out, workspace, nullptr, 0, st);
  return CB200_EUNSUPPORTED;
}

int cb200_scale_inplace(float* grad, int64_t n, const float* scale, void* stream) {
  if (!grad || !scale || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  if ((reinterpret_cast<uintptr_t>(grad) & 15) != 0) return CB200_EINVAL;
  scale_inplace_kernel<<<grid_for(n / 4 + 1, 256, 4), 256, 0, (cudaStream_t)stream>>>(grad, n / 4, n, scale);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_gather_add_coords(const void* offsets, int offsets_dtype, const void* coords, int coord_dtype, int batch,
                            int num_dims, const int64_t* spatial, int64_t pairs_per_sample, float* out, void* stream) {
  if (!offsets || !spatial || batch <= 0 || pairs_per_sample < 0) return CB200_EINVAL;
  if (pairs_per_sample == 0) return CB200_OK;
  if (!coords || !out) return CB200_EINVAL;
  if (!coords_aligned(coords, coord_dtype, num_dims)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_GATHER(DD)                                                                                          \
  switch (coord_dtype) {                                                                                          \
    case CB200_I64: return launch_gather<DD, long long>(offsets, offsets_dtype, coords, batch, spatial, pairs_per_sample, out, st); \
    case CB200_I32: return launch_gather<DD, int>(offsets, offsets_dtype, coords, batch, spatial, pairs_per_sample, out, st);       \
    case CB200_I16: return launch_gather<DD, short>(offsets, offsets_dtype, coords, batch, spatial, pairs_per_sample, out, st);     \
  }                                                                                                               \
  return CB200_EUNSUPPORTED;
  if (num_dims == 2) { CB200_GATHER(2) }
  if (num_dims == 3) { CB200_GATHER(3) }
#undef CB200_GATHER
  return CB200_EUNSUPPORTED;
}

int cb200_scatter_add_coords(const float* grad_out, const void* coords, int coord_dtype, int batch, int num_dims,
                             const int64_t* spatial, int64_t pairs_per_sample, float* grad_offsets, void* stream) {
  if (!spatial || !grad_offsets || batch <= 0 || pairs_per_sample < 0) return CB200_EINVAL;
  if (pairs_per_sample > 0 && (!grad_out || !coords)) return CB200_EINVAL;
  if (!coords_aligned(coords, coord_dtype, num_dims)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_SCATTER(DD)                                                                                     \
  switch (coord_dtype) {                                                                                      \
    case CB200_I64: return launch_scatter<DD, long long>(grad_out, coords, batch, spatial, pairs_per_sample, grad_offsets, st); \
    case CB200_I32: return launch_scatter<DD, int>(grad_out, coords, batch, spatial, pairs_per_sample, grad_offsets, st);       \
    case CB200_I16: return launch_scatter<DD, short>(grad_out, coords, batch, spatial, pairs_per_sample, grad_offsets, st);     \
  }                                                                                                           \
  return CB200_EUNSUPPORTED;
  if (num_dims == 2) { CB200_SCATTER(2) }
  if (num_dims == 3) { CB200_SCATTER(3) }
#undef CB200_SCATTER
  return CB200_EUNSUPPORTED;
}

int cb200_oce_pair_loss(const float* ea, const float* er, int64_t n_pairs, int num_dims, float temperature,
                        float regularization_weight, float* grad_ea, float* out, void* workspace, void* stream) {
  if (!out || !workspace || n_pairs < 0 || !(temperature != 0.f)) return CB200_EINVAL;
  if (n_pairs > 0 && (!ea || !er)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  auto* ws = static_cast<LossWorkspace*>(workspace);
  const int blocks = grid_for(n_pairs, LOSS_THREADS, 4);
  const float inv_t = 1.0f / temperature;
#define CB200_PAIR(DD)                                                                                               \
  if (grad_ea)                                                                                                       \
    pair_loss_kernel<DD, true><<<blocks, LOSS_THREADS, 0, st>>>(ea, er, n_pairs, inv_t, regularization_weight, grad_ea, ws, out); \
  else                                                                                                               \
    pair_loss_kernel<DD, false><<<blocks, LOSS_THREADS, 0, st>>>(ea, er, n_pairs, inv_t, regularization_weight, nullptr, ws, out);
  if (num_dims == 2) { CB200_PAIR(2) }
  else if (num_dims == 3) { CB200_PAIR(3) }
  else return CB200_EUNSUPPORTED;
#undef CB200_PAIR
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
