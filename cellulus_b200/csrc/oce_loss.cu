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
#include "common.cuh"

namespace cb200 {

struct LossWorkspace {
  double acc[2];            // sum(1 - exp(-d^2/T)), sum(||ea||)
  unsigned long long bad;   // pairs skipped: coordinate out of range
  unsigned int ticket;      // blocks finished
  unsigned int pad;
};

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

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_UNROLL = 4;

__device__ __forceinline__ void block_reduce_to_workspace(float oce, float nrm, int bad, LossWorkspace* ws,
                                                          float w, float* out) {
  __shared__ double s_acc[2][LOSS_THREADS / 32];
  __shared__ int s_bad[LOSS_THREADS / 32];
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
#pragma unroll
    for (int i = 0; i < LOSS_THREADS / 32; ++i) {
      ta += s_acc[0][i];
      tb += s_acc[1][i];
      tc += s_bad[i];
    }
    atomicAdd(&ws->acc[0], ta);
    atomicAdd(&ws->acc[1], tb);
    if (tc) atomicAdd(&ws->bad, (unsigned long long)tc);
    __threadfence();
    const unsigned t = atomicAdd(&ws->ticket, 1u);
    s_last = (t == gridDim.x - 1);
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
  }
}

template <int D, typename CT, typename OT, bool BWD>
__global__ void __launch_bounds__(LOSS_THREADS)
oce_loss_fused_kernel(const OT* __restrict__ offsets, const CT* __restrict__ anchors, const CT* __restrict__ refs,
                      int batch, int64_t P, Shape<D> shape, float inv_t, float w, float* __restrict__ grad,
                      LossWorkspace* ws, float* out) {
  const int lane = lane_id();
  const int warps_per_block = LOSS_THREADS / 32;
  const int64_t chunks_per_sample = (P + 31) >> 5;
  const int64_t n_chunks = chunks_per_sample * batch;
  const int64_t warp_gid = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * warps_per_block;
  const float two_inv_t = 2.0f * inv_t;

  float acc_oce = 0.f, acc_nrm = 0.f;
  int bad = 0;

  for (int64_t chunk0 = warp_gid * LOSS_UNROLL; chunk0 < n_chunks; chunk0 += warp_stride * LOSS_UNROLL) {
    int ca[LOSS_UNROLL][D], cr[LOSS_UNROLL][D];
    int b[LOSS_UNROLL];
    bool live[LOSS_UNROLL];
    // ---- phase 1: coordinate loads (2 * UNROLL independent 16-byte requests in flight) ----
#pragma unroll
    for (int u = 0; u < LOSS_UNROLL; ++u) {
      const int64_t chunk = chunk0 + u;
      const int bb = (int)(chunk / chunks_per_sample);
      const int64_t p = ((chunk - (int64_t)bb * chunks_per_sample) << 5) + lane;
      b[u] = bb;
      live[u] = (chunk < n_chunks) && (p < P);
      if (live[u]) {
        const int64_t pair = (int64_t)bb * P + p;
        load_coord<D, CT>(anchors, pair, ca[u]);
        load_coord<D, CT>(refs, pair, cr[u]);
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) ca[u][k] = cr[u][k] = 0;
      }
    }
    // ---- phase 2: pixel gathers (L2-resident offsets) ----
    float oa[LOSS_UNROLL][D], orf[LOSS_UNROLL][D];
    int pix_a[LOSS_UNROLL];
#pragma unroll
    for (int u = 0; u < LOSS_UNROLL; ++u) {
      int wa[D], wr[D];
      if (live[u]) {
        const bool ok = wrap_and_check<D>(ca[u], wa, shape) & wrap_and_check<D>(cr[u], wr, shape);
        if (!ok) {
          live[u] = false;
          ++bad;
        }
      }
      pix_a[u] = -1 - u;  // dead lanes never merge with a live neighbour
      if (live[u]) {
        const int pa = pixel_of<D>(wa, shape);
        const int pr = pixel_of<D>(wr, shape);
        const int64_t base = (int64_t)b[u] * D * shape.npix;
        pix_a[u] = pa;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          oa[u][k] = load_as_float<OT>(offsets, base + k * shape.npix + pa);
          orf[u][k] = load_as_float<OT>(offsets, base + k * shape.npix + pr);
        }
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) oa[u][k] = orf[u][k] = 0.f;
      }
    }
    // ---- phase 3: pair terms, segmented warp reduction, scatter ----
#pragma unroll
    for (int u = 0; u < LOSS_UNROLL; ++u) {
      float g[D];
      float d2 = 0.f, n2 = 0.f, diff[D], ea[D];
#pragma unroll
      for (int k = 0; k < D; ++k) {
        ea[k] = __fadd_rn(oa[u][k], (float)ca[u][k]);  // selection += coordinate (models/unet.py:120)
        const float er = __fadd_rn(orf[u][k], (float)cr[u][k]);
        diff[k] = ea[k] - er;
        d2 = fmaf(diff[k], diff[k], d2);
        n2 = fmaf(ea[k], ea[k], n2);
      }
      const float e = expf(-d2 * inv_t);
      const float nrm = sqrtf(n2);
      if (live[u]) {
        acc_oce += 1.0f - e;
        acc_nrm += nrm;
      }
      if constexpr (BWD) {
        const float ge = two_inv_t * e;
        const float gr = nrm > 0.f ? w / nrm : 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) g[k] = live[u] ? fmaf(ge, diff[k], gr * ea[k]) : 0.f;
        // runs of equal anchor pixel (same sample: a chunk never straddles samples)
        const int key = pix_a[u];
        const int prev = __shfl_up_sync(FULL, key, 1);
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
        if (head && live[u]) {
          float* gp = grad + (int64_t)b[u] * D * shape.npix + key;
#pragma unroll
          for (int k = 0; k < D; ++k) atomicAdd(gp + k * shape.npix, g[k]);
        }
      }
    }
  }
  block_reduce_to_workspace(acc_oce, acc_nrm, bad, ws, w, out);
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

template <int D>
static bool make_shape(const int64_t* spatial, Shape<D>& s) {
  int64_t npix = 1;
  for (int k = 0; k < D; ++k) {
    const int64_t e = spatial[D - 1 - k];  // column k = tensor axis D-1-k
    if (e <= 0 || e > INT32_MAX) return false;
    s.ext[k] = (int)e;
    npix *= e;
  }
  if (npix > INT32_MAX) return false;  // pixel indices are 32-bit per sample
  s.npix = npix;
  return true;
}

template <int D, typename CT, typename OT>
static int launch_fused(const void* offsets, const void* anchors, const void* refs, int batch, const int64_t* spatial,
                        int64_t P, float T, float w, float* grad, float* out, void* workspace, cudaStream_t st) {
  Shape<D> shape;
  if (!make_shape<D>(spatial, shape)) return CB200_EINVAL;
  const int64_t n_chunks = ((P + 31) >> 5) * batch;
  const int warps = LOSS_THREADS / 32;
  int64_t blocks = (n_chunks + (int64_t)warps * LOSS_UNROLL - 1) / ((int64_t)warps * LOSS_UNROLL);
  const int64_t cap = (int64_t)CB200_SM_COUNT * 8;  // persistent: 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  auto* ws = static_cast<LossWorkspace*>(workspace);
  if (grad) {
    CB200_CUDA_TRY(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)batch * D * shape.npix, st));
    oce_loss_fused_kernel<D, CT, OT, true><<<(int)blocks, LOSS_THREADS, 0, st>>>(
        (const OT*)offsets, (const CT*)anchors, (const CT*)refs, batch, P, shape, 1.0f / T, w, grad, ws, out);
  } else {
    oce_loss_fused_kernel<D, CT, OT, false><<<(int)blocks, LOSS_THREADS, 0, st>>>(
        (const OT*)offsets, (const CT*)anchors, (const CT*)refs, batch, P, shape, 1.0f / T, w, nullptr, ws, out);
  }
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

template <int D, typename CT>
static int dispatch_offsets(const void* offsets, int odt, const void* a, const void* r, int batch,
                            const int64_t* spatial, int64_t P, float T, float w, float* grad, float* out, void* ws,
                            cudaStream_t st) {
  if (odt == CB200_F32) return launch_fused<D, CT, float>(offsets, a, r, batch, spatial, P, T, w, grad, out, ws, st);
  if (odt == CB200_BF16)
    return launch_fused<D, CT, __nv_bfloat16>(offsets, a, r, batch, spatial, P, T, w, grad, out, ws, st);
  return CB200_EUNSUPPORTED;
}

template <int D>
static int dispatch_coords(const void* offsets, int odt, const void* a, const void* r, int cdt, int batch,
                           const int64_t* spatial, int64_t P, float T, float w, float* grad, float* out, void* ws,
                           cudaStream_t st) {
  switch (cdt) {
    case CB200_I64: return dispatch_offsets<D, long long>(offsets, odt, a, r, batch, spatial, P, T, w, grad, out, ws, st);
    case CB200_I32: return dispatch_offsets<D, int>(offsets, odt, a, r, batch, spatial, P, T, w, grad, out, ws, st);
    case CB200_I16: return dispatch_offsets<D, short>(offsets, odt, a, r, batch, spatial, P, T, w, grad, out, ws, st);
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

int cb200_oce_loss_fwd_bwd(const void* offsets, int offsets_dtype, const void* anchors, const void* refs,
                           int coord_dtype, int batch, int num_dims, const int64_t* spatial, int64_t pairs_per_sample,
                           float temperature, float regularization_weight, float* grad, float* out, void* workspace,
                           void* stream) {
  if (!offsets || !anchors || !refs || !spatial || !out || !workspace) return CB200_EINVAL;
  if (batch <= 0 || pairs_per_sample < 0 || !(temperature != 0.f)) return CB200_EINVAL;
  if (!coords_aligned(anchors, coord_dtype, num_dims) || !coords_aligned(refs, coord_dtype, num_dims)) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_dims == 2)
    return dispatch_coords<2>(offsets, offsets_dtype, anchors, refs, coord_dtype, batch, spatial, pairs_per_sample,
                              temperature, regularization_weight, grad, out, workspace, st);
  if (num_dims == 3)
    return dispatch_coords<3>(offsets, offsets_dtype, anchors, refs, coord_dtype, batch, spatial, pairs_per_sample,
                              temperature, regularization_weight, grad, out, workspace, st);
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
  if (!offsets || !coords || !spatial || !out || batch <= 0 || pairs_per_sample < 0) return CB200_EINVAL;
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
  if (!grad_out || !coords || !spatial || !grad_offsets || batch <= 0 || pairs_per_sample < 0) return CB200_EINVAL;
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
  if (!ea || !er || !out || !workspace || n_pairs < 0 || !(temperature != 0.f)) return CB200_EINVAL;
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
