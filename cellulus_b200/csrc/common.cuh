// Shared device/host helpers for the cellulus_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cellulus_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cellulus_b200 kernels are written for sm_100a (B200) only"
#endif

#define CB200_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#define CB200_CUDA_TRY(expr)                    \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

#define CB200_LAUNCH_CHECK()                    \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

namespace cb200 {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- streaming (read-once) loads: bypass L1 allocation --------------------
__device__ __forceinline__ longlong2 ld_stream_ll2(const void* p) {
  longlong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ long long ld_stream_ll(const void* p) {
  long long r;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f(const void* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

// ---- warp reductions -------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// fixed-order (lane 0 .. 31 tree) sum that is identical on every call: shfl_down tree
__device__ __forceinline__ double warp_sum_down(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  return v;  // valid in lane 0
}

// ---- element loads with conversion ----------------------------------------
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) {
  return __ldg(p + i);
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(__ldg(p + i));
}

template <typename T>
__device__ __forceinline__ double load_as_double(const T* p, int64_t i);
template <>
__device__ __forceinline__ double load_as_double<float>(const float* p, int64_t i) {
  return (double)__ldg(p + i);
}
template <>
__device__ __forceinline__ double load_as_double<double>(const double* p, int64_t i) {
  return __ldg(p + i);
}

// ---- Philox4x32-10 counter RNG (for the device sampler / Bernoulli flags) ---
struct Philox {
  uint32_t key[2];
  __device__ Philox(uint64_t seed) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
  }
  __device__ __forceinline__ uint4 operator()(uint64_t counter, uint64_t sequence) const {
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
    uint32_t c2 = (uint32_t)sequence, c3 = (uint32_t)(sequence >> 32);
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// unbiased integer in [0, n) from a 32-bit draw (Lemire multiply-shift; the
// residual bias is < n / 2^32, far below what any test can resolve)
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) { return __umulhi(r, n); }

inline int grid_for(int64_t work_items, int threads, int per_thread = 1, int max_waves = 8) {
  int64_t blocks = (work_items + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
  int64_t cap = (int64_t)CB200_SM_COUNT * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// detect.cu: forms of cb200_minmax / cb200_select_points whose element count is read ON THE DEVICE (*n_dev), so that a
// composite can enqueue them before it has read that count (n_max sizes the launch; n_dev == NULL: exactly n_max)
int minmax_counted(const void* x, int dtype, int64_t n_max, const long long* n_dev, double* out2, void* workspace,
                   cudaStream_t st);
int select_points_counted(const double* src, int64_t n_max, const long long* n_dev, int64_t src_stride, int num_dims,
                          const uint8_t* flags, double* dst, int64_t dst_stride, long long* n_out, void* workspace,
                          cudaStream_t st);

// meanshift.cu: cb200_ms_grid_modes with its optional modes (unfinished hand-over after `eval_limit` evaluations; resume
// of the seeds in `worklist`, their number read on the device)
int ms_grid_modes_launch(const double* points_sorted, int64_t n_points, int64_t sorted_stride, const cb200_grid* grid,
                         const int* cell_start, double* means, int64_t seed_stride, int64_t n_seeds, double bandwidth,
                         int max_iter, int* counts, int* iters, int* work_counter, const int* worklist,
                         const long long* n_work_dev, int eval_limit, cudaStream_t st);

// ------------------------------------------------------------------ mbarrier / 1-D TMA helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace cb200
