// Test-time-augmentation aggregate for sm_100a (models/unet.py:90-98).
//
// stack (T, C, n) fp32 -> out (C+1, n): per-channel mean over the T passes and
// the per-channel POPULATION std summed over channels.  Pure streaming: every
// input byte is read once with 16-byte no-allocate loads, T/UNROLL batches of
// independent requests in flight per thread; the statistics are carried in
// registers with Welford's update (the same recurrence torch.std_mean uses),
// so nothing but the (C+1, n) result is written.
#include "common.cuh"

namespace cb200 {

constexpr int TTA_THREADS = 128;
constexpr int TTA_MAXC = 4;

template <int V>
struct Vec {
  float v[V];
};
template <int V>
__device__ __forceinline__ Vec<V> ld_stream_vec(const float* p);
template <>
__device__ __forceinline__ Vec<4> ld_stream_vec<4>(const float* p) {
  const float4 t = ld_stream_f4(p);
  return Vec<4>{{t.x, t.y, t.z, t.w}};
}
template <>
__device__ __forceinline__ Vec<2> ld_stream_vec<2>(const float* p) {
  Vec<2> r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
  return r;
}
template <int V>
__device__ __forceinline__ void st_stream_vec(float* p, const Vec<V>& x) {
  if constexpr (V == 4) st_stream_f4(p, make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
  else asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x.v[0]), "f"(x.v[1]));
}

template <int V>
struct WelfordV {
  Vec<V> mean, m2;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < V; ++j) mean.v[j] = m2.v[j] = 0.f;
  }
  __device__ __forceinline__ void push(const Vec<V>& x, const float inv_count) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float d = x.v[j] - mean.v[j];
      mean.v[j] = fmaf(d, inv_count, mean.v[j]);
      m2.v[j] = fmaf(d, x.v[j] - mean.v[j], m2.v[j]);
    }
  }
};

// One thread owns V consecutive pixels of every channel (V = 4, or 2 when the block being aggregated is so
// small -- a 496x496 scan block is 246 k pixels -- that 16-byte granules would leave most of the chip without
// a thread).  The T passes are consumed in batches of U; batch b+1 is loaded (U*C independent requests) BEFORE
// batch b is folded into the Welford state, so the HBM stream never drains.
template <int C, int V>
__global__ void __launch_bounds__(TTA_THREADS)
tta_aggregate_kernel(const float* __restrict__ stack, int T, int64_t n, int64_t nv, float* __restrict__ out) {
  constexpr int U = (8 / C > 0 ? 8 / C : 1) * (V == 2 ? 2 : 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_T = 1.0f / (float)T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    WelfordV<V> w[C];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c].init();
    Vec<V> cur[U][C], nxt[U][C];
    auto load = [&](Vec<V> (&v)[U][C], int t0) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (t0 + u < T) v[u][c] = ld_stream_vec<V>(stack + ((int64_t)(t0 + u) * C + c) * n + i * V);
    };
    load(cur, 0);
    for (int t = 0; t < T; t += U) {
      if (t + U < T) load(nxt, t + U);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < T) {
          const float inv = 1.0f / (float)(t + u + 1);
#pragma unroll
          for (int c = 0; c < C; ++c) w[c].push(cur[u][c], inv);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) cur[u][c] = nxt[u][c];
    }
    Vec<V> s;
#pragma unroll
    for (int j = 0; j < V; ++j) s.v[j] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      st_stream_vec<V>(out + (int64_t)c * n + i * V, w[c].mean);
#pragma unroll
      for (int j = 0; j < V; ++j) s.v[j] += sqrtf(w[c].m2.v[j] * inv_T);
    }
    st_stream_vec<V>(out + (int64_t)C * n + i * V, s);
  }
}

// Small blocks (a 496 x 496 scan block is 246 k pixels): even 8-byte granules leave the chip half empty and
// every thread walks all T passes one after the other.  Here TWO adjacent lanes share a granule, each folds
// one half of the passes (ceil(T/2) / floor(T/2)), and the halves are merged with the pairwise update of
// Chan et al. (mean and M2 of the union from those of the parts) through one shuffle -- twice the threads,
// half the dependent chain, and every warp load still covers two contiguous 128-byte runs.
template <int C, int V>
__global__ void __launch_bounds__(TTA_THREADS)
tta_aggregate_split_kernel(const float* __restrict__ stack, int T, int64_t n, int64_t nv, float* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = tid >> 1;  // granule
  const int half = (int)(tid & 1);
  const bool live = i < nv;    // nv is padded so that whole warps stay together for the shuffle
  const int t_split = (T + 1) / 2;
  const int t0 = half ? t_split : 0, t1 = half ? T : t_split;
  WelfordV<V> w[C];
#pragma unroll
  for (int c = 0; c < C; ++c) w[c].init();
  if (live) {
    constexpr int U = 8;
    for (int t = t0; t < t1; t += U) {
      Vec<V> x[U][C];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (t + u < t1) x[u][c] = ld_stream_vec<V>(stack + ((int64_t)(t + u) * C + c) * n + i * V);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (t + u < t1) {
          const float inv = 1.0f / (float)(t + u - t0 + 1);
#pragma unroll
          for (int c = 0; c < C; ++c) w[c].push(x[u][c], inv);
        }
    }
  }
  // merge: the even lane receives the odd lane's half
  const float na = (float)t_split, nb = (float)(T - t_split), inv_T = 1.0f / (float)T;
  Vec<V> s;
#pragma unroll
  for (int j = 0; j < V; ++j) s.v[j] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float mean_b = __shfl_down_sync(0xffffffffu, w[c].mean.v[j], 1);
      const float m2_b = __shfl_down_sync(0xffffffffu, w[c].m2.v[j], 1);
      const float delta = mean_b - w[c].mean.v[j];
      w[c].mean.v[j] = fmaf(delta, nb * inv_T, w[c].mean.v[j]);
      w[c].m2.v[j] = w[c].m2.v[j] + m2_b + delta * delta * (na * nb * inv_T);
      s.v[j] += sqrtf(w[c].m2.v[j] * inv_T);
    }
    if (live && !half) st_stream_vec<V>(out + (int64_t)c * n + i * V, w[c].mean);
  }
  if (live && !half) st_stream_vec<V>(out + (int64_t)C * n + i * V, s);
}

// Generic scalar fallback: any C <= TTA_MAXC, any n / alignment; `first` = first pixel handled.
__global__ void __launch_bounds__(TTA_THREADS)
tta_aggregate_scalar_kernel(const float* __restrict__ stack, int T, int C, int64_t n, int64_t first,
                            float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_T = 1.0f / (float)T;
  for (int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      float mean = 0.f, m2 = 0.f;
      for (int t = 0; t < T; ++t) {
        const float x = ld_stream_f(stack + ((int64_t)t * C + c) * n + i);
        const float d = x - mean;
        mean = fmaf(d, 1.0f / (float)(t + 1), mean);
        m2 = fmaf(d, x - mean, m2);
      }
      out[(int64_t)c * n + i] = mean;
      s += sqrtf(m2 * inv_T);
    }
    out[(int64_t)C * n + i] = s;
  }
}

// Streaming form: state = [mean (C,n) | M2 (C,n)], one prediction folded in per call.
__global__ void __launch_bounds__(TTA_THREADS)
tta_accumulate_kernel(float* __restrict__ state, const float* __restrict__ pred, int t, int64_t cn) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv = 1.0f / (float)(t + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cn; i += stride) {
    const float x = ld_stream_f(pred + i);
    float mean = t == 0 ? 0.f : state[i];
    float m2 = t == 0 ? 0.f : state[cn + i];
    const float d = x - mean;
    mean = fmaf(d, inv, mean);
    m2 = fmaf(d, x - mean, m2);
    state[i] = mean;
    state[cn + i] = m2;
  }
}

__global__ void __launch_bounds__(TTA_THREADS)
tta_finalize_kernel(const float* __restrict__ state, int T, int C, int64_t n, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_T = 1.0f / (float)T;
  const int64_t cn = (int64_t)C * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      out[(int64_t)c * n + i] = state[(int64_t)c * n + i];
      s += sqrtf(state[cn + (int64_t)c * n + i] * inv_T);
    }
    out[cn + i] = s;
  }
}

// Salt / pepper noise of the infer-mode forward (models/unet.py:80-82): out = (u <= p) ? value : raw,
// u ~ U[0,1) from Philox (the reference draws u on the CPU and copies it over for every pass).
// `seed_dev` (optional): the seed is read from device memory, so that a CUDA graph of the whole test-time-augmentation
// loop draws fresh noise on every replay (the graph bumps the seed itself).
__global__ void __launch_bounds__(256)
salt_pepper_kernel(const float* __restrict__ raw, int64_t n, float p, float value, uint64_t seed,
                   const uint64_t* __restrict__ seed_dev, uint64_t sequence, float* __restrict__ out) {
  const Philox rng(seed_dev ? __ldg(seed_dev) : seed);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (n + 3) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 r = rng((uint64_t)i, sequence);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = i * 4 + j;
      if (e < n) {
        const float u = (float)(rr[j] >> 8) * (1.0f / 16777216.0f);  // 24-bit uniform in [0, 1)
        out[e] = u <= p ? value : raw[e];
      }
    }
  }
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int cb200_salt_pepper(const float* raw, int64_t n, float p, float value, uint64_t seed, uint64_t sequence, float* out,
                      void* stream) {
  if (!raw || !out || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  salt_pepper_kernel<<<grid_for((n + 3) / 4, 256, 2, 16), 256, 0, (cudaStream_t)stream>>>(raw, n, p, value, seed,
                                                                                          nullptr, sequence, out);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_salt_pepper_device_seed(const float* raw, int64_t n, float p, float value, const uint64_t* seed, uint64_t sequence,
                                  float* out, void* stream) {
  if (!raw || !out || !seed || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  salt_pepper_kernel<<<grid_for((n + 3) / 4, 256, 2, 16), 256, 0, (cudaStream_t)stream>>>(raw, n, p, value, 0, seed,
                                                                                          sequence, out);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_tta_aggregate(const float* stack, int num_passes, int channels, int64_t n, float* out, void* stream) {
  if (!stack || !out || num_passes <= 0 || channels <= 0 || channels > TTA_MAXC || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(stack) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  int64_t first_scalar = 0;
  if (vec_ok) {
    // 16-byte granules when they still give every SM ~1k threads, 8-byte granules for small blocks
    const bool wide = n / 4 >= (int64_t)CB200_SM_COUNT * 1024;
    const int64_t nv = wide ? n / 4 : n / 2;
    const int blocks = grid_for(nv, TTA_THREADS, 1, 16);
    const bool split = !wide && num_passes >= 4;  // small block: two lanes per granule, half the passes each
    const int split_threads = 64;
    const int64_t split_blocks = (2 * nv + split_threads - 1) / split_threads;
#define CB200_TTA(CC)                                                                                      \
  if (wide) tta_aggregate_kernel<CC, 4><<<blocks, TTA_THREADS, 0, st>>>(stack, num_passes, n, nv, out);    \
  else if (split) tta_aggregate_split_kernel<CC, 2><<<(unsigned)split_blocks, split_threads, 0, st>>>(stack, num_passes, n, nv, out); \
  else tta_aggregate_kernel<CC, 2><<<blocks, TTA_THREADS, 0, st>>>(stack, num_passes, n, nv, out);
    switch (channels) {
      case 1: CB200_TTA(1) break;
      case 2: CB200_TTA(2) break;
      case 3: CB200_TTA(3) break;
      default: CB200_TTA(4) break;
    }
#undef CB200_TTA
    CB200_LAUNCH_CHECK();
    first_scalar = n;
  }
  if (first_scalar < n) {
    tta_aggregate_scalar_kernel<<<grid_for(n - first_scalar, TTA_THREADS, 1, 16), TTA_THREADS, 0, st>>>(
        stack, num_passes, channels, n, first_scalar, out);
    CB200_LAUNCH_CHECK();
  }
  return CB200_OK;
}

int cb200_tta_accumulate(float* state, const float* prediction, int t, int channels, int64_t n, void* stream) {
  if (!state || !prediction || t < 0 || channels <= 0 || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  const int64_t cn = (int64_t)channels * n;
  tta_accumulate_kernel<<<grid_for(cn, TTA_THREADS, 2, 16), TTA_THREADS, 0, (cudaStream_t)stream>>>(state, prediction,
                                                                                                  t, cn);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

int cb200_tta_finalize(const float* state, int num_passes, int channels, int64_t n, float* out, void* stream) {
  if (!state || !out || num_passes <= 0 || channels <= 0 || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  tta_finalize_kernel<<<grid_for(n, TTA_THREADS, 2, 16), TTA_THREADS, 0, (cudaStream_t)stream>>>(state, num_passes,
                                                                                               channels, n, out);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
