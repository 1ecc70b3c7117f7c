// Device helpers shared by the OCE-loss kernels (oce_loss.cu: explicit pair lists; oce_sampled.cu: pairs drawn
// inside the kernel): shapes, pixel gathers / gradient reductions for both memory layouts, the block -> workspace
// reduction that finishes loss / oce / reg in the same launch.
#pragma once

#include "common.cuh"

namespace cb200 {

struct LossWorkspace {
  double acc[2];            // sum(1 - exp(-d^2/T)), sum(||ea||)
  unsigned long long bad;   // pairs skipped: coordinate out of range
  unsigned int ticket;      // blocks finished
  unsigned int pad;
  // in-kernel operand preparation (oce_loss.cu): CTAs that have finished their slice, per sample slot (b % 32);
  // one counter per 128-byte line so that the pollers of different samples hit different L2 slices
  unsigned int arrived[32 * 32];
};
constexpr unsigned LOSS_ZERO_SLOTS = 32, LOSS_ZERO_PITCH = 32;

template <int D>
struct Shape {
  int ext[D];      // extent per COLUMN (x, y[, z]) = reversed tensor axes
  int64_t npix;    // product
};

template <int D, typename CT>
__device__ __forceinline__ void load_coord(const CT* __restrict__ base, int64_t pair, int (&c)[D]) {
  if constexpr (D == 2 && sizeof(CT) == 8) {
    const longlong2 v = ld_stream_ll2(base + pair * 2);
    c[0] = (int)v.x;
    c[1] = (int)v.y;
  } else if constexpr (D == 2 && sizeof(CT) == 4) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(base) + pair);
    c[0] = v.x;
    c[1] = v.y;
  } else if constexpr (D == 2 && sizeof(CT) == 2) {
    const short2 v = __ldg(reinterpret_cast<const short2*>(base) + pair);
    c[0] = v.x;
    c[1] = v.y;
  } else if constexpr (sizeof(CT) == 8) {
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = (int)ld_stream_ll(base + pair * D + k);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = (int)__ldg(base + pair * D + k);
  }
}

// torch advanced indexing wraps negative indices once; anything else is an error
template <int D>
__device__ __forceinline__ bool wrap_and_check(const int (&c)[D], int (&wrapped)[D], const Shape<D>& s) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const int v = c[k];
    const int w = v < 0 ? v + s.ext[k] : v;
    ok = ok && (w >= 0) && (w < s.ext[k]);
    wrapped[k] = w;
  }
  return ok;
}

template <int D>
__device__ __forceinline__ int pixel_of(const int (&c)[D], const Shape<D>& s) {
  if constexpr (D == 2) return c[1] * s.ext[0] + c[0];
  return (c[2] * s.ext[1] + c[1]) * s.ext[0] + c[0];
}

constexpr int LOSS_MAX_WARPS = 16;
__device__ __forceinline__ void block_reduce_to_workspace(float oce, float nrm, int bad, LossWorkspace* ws,
                                                          float w, float* out) {
  __shared__ double s_acc[2][LOSS_MAX_WARPS];
  __shared__ int s_bad[LOSS_MAX_WARPS];
  __shared__ bool s_last;
  double a = warp_sum((double)oce), b = warp_sum((double)nrm);
  int c = warp_sum(bad);
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 0) {
    s_acc[0][warp] = a;
    s_acc[1][warp] = b;
    s_bad[warp] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    int tc = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      ta += s_acc[0][i];
      tb += s_acc[1][i];
      tc += s_bad[i];
    }
    atomicAdd(&ws->acc[0], ta);
    atomicAdd(&ws->acc[1], tb);
    if (tc) atomicAdd(&ws->bad, (unsigned long long)tc);
    __threadfence();
    const unsigned t = atomicAdd(&ws->ticket, 1u);
    s_last = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    const double oce_sum = atomicAdd(&ws->acc[0], 0.0);
    const double nrm_sum = atomicAdd(&ws->acc[1], 0.0);
    const unsigned long long nbad = atomicAdd(&ws->bad, 0ull);
    const float oce_f = (float)oce_sum;
    const float reg_f = w * (float)nrm_sum;  // criterions/oce_loss.py:59-61
    out[0] = oce_f + reg_f;                  // :62
    out[1] = oce_f;
    out[2] = reg_f;
    out[3] = (float)nbad;
    ws->acc[0] = 0.0;  // leave the workspace zeroed for the next call
    ws->acc[1] = 0.0;
    ws->bad = 0ull;
    ws->ticket = 0u;
#pragma unroll
    for (unsigned i = 0; i < LOSS_ZERO_SLOTS; ++i) ws->arrived[i * LOSS_ZERO_PITCH] = 0u;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// pixel gather: PLANAR = (B, D, *S) contiguous; interleaved = channels-last (B, *S, D) contiguous.
// `first` = index of the sample's first pixel in the whole tensor (b * npix); all element indices are 32-bit
// (the launcher refuses tensors of 2^31 elements or more), so one address costs one IMAD.WIDE on the
// kernel-parameter base instead of a rebuilt 64-bit per-sample pointer.
template <int D, typename OT, bool IL>
__device__ __forceinline__ void gather_pixel(const OT* __restrict__ base, unsigned npix, unsigned first, unsigned pix,
                                             float (&o)[D]) {
  if constexpr (!IL) {
    const unsigned e = first * D + pix;
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = load_as_float<OT>(base, e + k * npix);
  } else if constexpr (D == 2 && sizeof(OT) == 4) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(base) + (first + pix));
    o[0] = v.x;
    o[1] = v.y;
  } else if constexpr (D == 2 && sizeof(OT) == 2) {
    const __nv_bfloat162 v = __ldg(reinterpret_cast<const __nv_bfloat162*>(base) + (first + pix));
    o[0] = __low2float(v);
    o[1] = __high2float(v);
  } else {
    const unsigned e = (first + pix) * D;
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = load_as_float<OT>(base, e + k);
  }
}

template <int D, bool IL>
__device__ __forceinline__ void scatter_pixel(float* __restrict__ gbase, unsigned npix, unsigned first, unsigned pix,
                                              const float (&g)[D]) {
  if constexpr (!IL) {
    const unsigned e = first * D + pix;
#pragma unroll
    for (int k = 0; k < D; ++k) atomicAdd(gbase + (e + k * npix), g[k]);
  } else if constexpr (D == 2) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(reinterpret_cast<float2*>(gbase) + (first + pix)),
                 "f"(g[0]), "f"(g[1])
                 : "memory");
  } else {
    const unsigned e = (first + pix) * D;
#pragma unroll
    for (int k = 0; k < D; ++k) atomicAdd(gbase + (e + k), g[k]);
  }
}

// Gather from the channels-last STAGED copy of planar 2-D offsets that a loss kernel writes at the start of its own
// launch (oce_loss.cu: oce_loss_staged_kernel; oce_sampled.cu: LY_STAGED).
template <int D, typename OT>
__device__ __forceinline__ void gather_staged(const OT* base, unsigned first, unsigned pix, float (&o)[D]) {
  static_assert(D == 2, "staged gathers are built for 2-D embeddings");
  // the copy was written by other CTAs of THIS launch: plain (coherent) loads, not ld.global.nc
  if constexpr (sizeof(OT) == 4) {
    asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(o[0]), "=f"(o[1]) : "l"(reinterpret_cast<const float2*>(base) + (first + pix)));
  } else {
    unsigned v;
    asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(reinterpret_cast<const unsigned*>(base) + (first + pix)));
    o[0] = __uint_as_float(v << 16);
    o[1] = __uint_as_float(v & 0xffff0000u);
  }
}

// how a loss kernel sees the (B, D, *S) offsets / gradient tensors
constexpr int LY_PLANAR = 0, LY_CL = 1, LY_STAGED = 2;  // staged: gathers from the copy, gradient stays planar

// planar (2, npix) per sample -> interleaved (npix, 2): this CTA's share of ALL batch * npix pixels.
// vec: npix even and both bases aligned (8-byte loads, 16-byte stores, two pixels per thread).
template <typename OT>
__device__ __forceinline__ void stage_interleaved_slice(const OT* __restrict__ offsets, OT* __restrict__ staged, unsigned batch,
                                                        unsigned npix, bool vec) {
  const unsigned total = batch * npix;  // < 2^31 (launcher)
  const unsigned per = ((total + gridDim.x - 1) / gridDim.x + 1u) & ~1u;
  const unsigned lo = min(total, blockIdx.x * per), hi = min(total, lo + per);
  if constexpr (sizeof(OT) == 4) {
    if (vec) {
      for (unsigned g = lo + 2 * threadIdx.x; g + 1 < hi; g += 2 * blockDim.x) {
        const unsigned b = g / npix, i = g - b * npix;
        const OT* src = offsets + (size_t)b * npix * 2 + i;
        const float2 x = __ldg(reinterpret_cast<const float2*>(src));
        const float2 y = __ldg(reinterpret_cast<const float2*>(src + npix));
        *reinterpret_cast<float4*>(staged + 2 * (size_t)g) = make_float4(x.x, y.x, x.y, y.y);
      }
      return;
    }
  }
  for (unsigned g = lo + threadIdx.x; g < hi; g += blockDim.x) {
    const unsigned b = g / npix, i = g - b * npix;
    const OT* src = offsets + (size_t)b * npix * 2 + i;
    staged[2 * (size_t)g] = src[0];
    staged[2 * (size_t)g + 1] = src[npix];
  }
}

template <int D>
static bool make_shape(const int64_t* spatial, Shape<D>& s) {
  int64_t npix = 1;
  for (int k = 0; k < D; ++k) {
    const int64_t e = spatial[D - 1 - k];  // column k = tensor axis D-1-k
    if (e <= 0 || e > INT32_MAX) return false;
    s.ext[k] = (int)e;
    npix *= e;
  }
  if (npix * D > INT32_MAX) return false;  // element offsets are 32-bit per sample
  s.npix = npix;
  return true;
}


// gradient zero-fill that releases its dependent grid early (programmatic dependent launch); oce_loss.cu
int zero_fill(float* p, int64_t n, cudaStream_t st);
extern const bool g_loss_pdl;

}  // namespace cb200
