// Greedy seed-and-grow clustering for sm_100a (cellulus/utils/greedy_cluster.py:46-120, 176-253; the
// `clustering = "greedy"` branch of detect.py:162-192).
//
// The reference is a host loop with ~10 eager kernels and two `.item()` synchronisations per object:
//   while unclustered.sum() > min_unclustered_sum:
//       seed = argmax(seed_map * unclustered); stop if its score < seed_thresh
//       proposal = exp(-sum_k (e_k - c_k)^2 / (2 bw^2)) > 0.5
//       accept if |proposal| > min_object_size and more than half of it is still unclustered
//       unclustered[proposal] = 0
// Here it is ONE persistent cooperative kernel: all CTAs are co-resident, the whole loop runs on the device
// with two grid barriers per object (global argmax; proposal statistics), the point set stays in L2.
// Arithmetic follows the reference's tensor dtype (fp32 in 2-D; the dtype of the stored embeddings in 3-D).
#include "common.cuh"
#include "compact.cuh"

namespace cb200 {

struct GreedyState {           // device scratch, zero-initialised by the launcher
  unsigned int barrier;        // monotonically increasing arrival counter
  int count;                   // next instance id - 1
  long long unclustered;       // points not yet clustered
  int n_prop[2];               // ping-pong proposal statistics
  int n_prop_unclustered[2];
  int n_objects;               // result: number of instances written
  int iterations;              // result: seeds tried
};

__device__ __forceinline__ void grid_barrier(GreedyState* st, unsigned& generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++generation;
    __threadfence();
    atomicAdd(&st->barrier, 1u);
    const unsigned target = generation * gridDim.x;
    while (*((volatile unsigned*)&st->barrier) < target) { __nanosleep(20); }
    __threadfence();
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }

#ifndef CB200_GREEDY_BPS
#define CB200_GREEDY_BPS 2  // thread blocks per SM of the persistent grid (4 x 256: 11.0 ms, 2 x 512: 9.7 ms, 1 x 1024: 9.9 ms per configs[2] volume)
#endif
#ifndef CB200_GREEDY_THREADS
#define CB200_GREEDY_THREADS 512
#endif
constexpr int GREEDY_THREADS = CB200_GREEDY_THREADS;

template <typename T, int D>
__global__ void __launch_bounds__(GREEDY_THREADS)
greedy_cluster_kernel(const T* __restrict__ emb, int64_t stride, const T* __restrict__ seed_map, int n, T two_bw2,
                      int min_object_size, double seed_thresh, long long min_unclustered_sum, uint8_t* unclustered,
                      uint8_t* proposal, short* __restrict__ instance, T* slot_val, int* slot_idx, GreedyState* st) {
  __shared__ T s_val[GREEDY_THREADS / 32];
  __shared__ int s_idx[GREEDY_THREADS / 32];
  __shared__ int s_cnt[2][GREEDY_THREADS / 32];
  __shared__ T s_best_val;
  __shared__ int s_best_idx;
  unsigned generation = 0;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  int parity = 0;

  for (int i = tid; i < n; i += nthreads) {
    unclustered[i] = 1;
    instance[i] = 0;
  }
  if (tid == 0) st->unclustered = n;
  grid_barrier(st, generation);

  while (true) {
    // ---- A. global argmax of seed_map * unclustered (first index among equals, like torch.argmax)
    T best = (T)-1;
    int best_i = 0x7fffffff;
    for (int i = tid; i < n; i += nthreads) {
      const T score = seed_map[i] * (T)unclustered[i];
      if (score > best) {  // ascending i within a thread: ties keep the earlier index
        best = score;
        best_i = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < GREEDY_THREADS / 32; ++w)
        if (s_val[w] > best || (s_val[w] == best && s_idx[w] < best_i)) { best = s_val[w]; best_i = s_idx[w]; }
      slot_val[blockIdx.x] = best;
      slot_idx[blockIdx.x] = best_i;
      if (blockIdx.x == 0) {  // reset the statistics slot this iteration will fill
        st->n_prop[parity] = 0;
        st->n_prop_unclustered[parity] = 0;
      }
    }
    grid_barrier(st, generation);
    if (warp == 0) {  // every CTA reduces the per-CTA candidates itself: no broadcast barrier needed
      best = (T)-1;
      best_i = 0x7fffffff;
      for (int b = lane; b < (int)gridDim.x; b += 32) {
        const T v = *((volatile T*)&slot_val[b]);
        const int ix = *((volatile int*)&slot_idx[b]);
        if (v > best || (v == best && ix < best_i)) { best = v; best_i = ix; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T ov = __shfl_xor_sync(FULL, best, o);
        const int oi = __shfl_xor_sync(FULL, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      }
      if (lane == 0) { s_best_val = best; s_best_idx = best_i; }
    }
    __syncthreads();
    const T seed_score = s_best_val;
    const int seed = s_best_idx;
    const long long still = *((volatile long long*)&st->unclustered);
    if (!(still > min_unclustered_sum) || (double)seed_score < seed_thresh || seed >= n) break;  // uniform across the grid

    // ---- B. proposal around the seed's embedding and its statistics
    T c[D];
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = emb[k * stride + seed];
    int np = 0, npu = 0;
    for (int i = tid; i < n; i += nthreads) {
      T s = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const T d = emb[k * stride + i] - c[k];
        s = s + (d * d) / two_bw2;  // torch.pow(e - c, 2) / (2 bw^2), summed over the channels in order
      }
      const bool p = t_exp<T>((T)-1 * s) > (T)0.5;
      proposal[i] = p;
      if (p) {
        ++np;
        npu += (i != seed && unclustered[i]) ? 1 : 0;  // the seed itself was cleared before the test (:89)
      }
    }
    np = warp_sum(np);
    npu = warp_sum(npu);
    if (lane == 0) { s_cnt[0][warp] = np; s_cnt[1][warp] = npu; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int a = 0, b = 0;
      for (int w = 0; w < GREEDY_THREADS / 32; ++w) { a += s_cnt[0][w]; b += s_cnt[1][w]; }
      if (a) atomicAdd(&st->n_prop[parity], a);
      if (b) atomicAdd(&st->n_prop_unclustered[parity], b);
    }
    grid_barrier(st, generation);

    // ---- C. accept / reject, clear the proposal
    const int n_prop = *((volatile int*)&st->n_prop[parity]);
    const int n_pu = *((volatile int*)&st->n_prop_unclustered[parity]);
    const bool accept = n_prop > min_object_size && ((float)n_pu / (float)n_prop > 0.5f);
    const int count = *((volatile int*)&st->count) + 1;
    for (int i = tid; i < n; i += nthreads) {
      if (i == seed) unclustered[i] = 0;
      if (proposal[i]) {
        if (accept) instance[i] = (short)count;
        unclustered[i] = 0;
      }
    }
    grid_barrier(st, generation);  // everybody has read `count` / `unclustered` before they change
    if (tid == 0) {
      st->unclustered = still - 1 - n_pu;
      if (accept) st->count = count;
      st->iterations += 1;
    }
    parity ^= 1;
    grid_barrier(st, generation);
  }
  if (tid == 0) st->n_objects = st->count;
}

// (emb_k + coordinate_k) in the reference's tensor dtype, and the normalised seed map, for foreground pixels
template <typename TI, typename T, int D>
struct GreedyEmit {
  const TI* emb;
  int64_t n_pix;
  int ext[3];
  T smin, smax;
  T* out_emb;
  int64_t capacity;
  T* out_seed;
  int32_t* pix_index;
  __device__ __forceinline__ void operator()(int64_t i, long long dst) const {
    const unsigned pix = (unsigned)i;
    unsigned c[3];
    c[0] = pix % (unsigned)ext[0];
    const unsigned r = pix / (unsigned)ext[0];
    if constexpr (D == 2) { c[1] = r; c[2] = 0; } else { c[1] = r % (unsigned)ext[1]; c[2] = r / (unsigned)ext[1]; }
#pragma unroll
    for (int k = 0; k < D; ++k) out_emb[k * capacity + dst] = (T)emb[k * n_pix + i] + (T)(float)c[k];
    out_seed[dst] = ((T)emb[(int64_t)D * n_pix + i] - smax) / (smin - smax);
    pix_index[dst] = (int32_t)pix;
  }
};
struct U8Pred {
  const uint8_t* mask;
  __device__ __forceinline__ bool operator()(int64_t i) const { return mask[i] != 0; }
};

__global__ void __launch_bounds__(256)
scatter_i16_kernel(const short* __restrict__ src, const int32_t* __restrict__ pix_index, int64_t n, short* __restrict__ dst) {
  const int64_t gs = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) dst[pix_index[i]] = src[i];
}

template <typename TI, typename T, int D>
static int greedy_prepare_typed(const void* emb, const int64_t* spatial, const uint8_t* mask, double smin, double smax,
                                void* out_emb, int64_t capacity, void* out_seed, int32_t* pix_index, long long* n_out,
                                void* workspace, cudaStream_t st) {
  int64_t n_pix = 1;
  for (int k = 0; k < D; ++k) n_pix *= spatial[k];
  if (n_pix > INT32_MAX) return CB200_EUNSUPPORTED;
  GreedyEmit<TI, T, D> emit{(const TI*)emb, n_pix,
                            {(int)spatial[D - 1], (int)spatial[D - 2], D == 3 ? (int)spatial[0] : 1},
                            (T)smin, (T)smax, (T*)out_emb, capacity, (T*)out_seed, pix_index};
  U8Pred pred{mask};
  return run_compaction(pred, emit, n_pix, capacity, n_out, workspace, st);
}

template <typename T, int D>
static int greedy_run(const void* emb, int64_t stride, const void* seed_map, int n, double bandwidth,
                      int min_object_size, double seed_thresh, long long min_unclustered_sum, short* instance,
                      void* workspace, int* result2, cudaStream_t st) {
  auto kernel = greedy_cluster_kernel<T, D>;
  int occ = 0;
  CB200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, GREEDY_THREADS, 0));
  if (occ < 1) return CB200_EINVAL;
  int blocks = std::min(CB200_SM_COUNT * std::min(occ, CB200_GREEDY_BPS), std::max(1, (n + GREEDY_THREADS - 1) / GREEDY_THREADS));
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  GreedyState* state = (GreedyState*)w;                 w += 256;
  uint8_t* unclustered = (uint8_t*)w;                   w += ((size_t)n + 255) / 256 * 256;
  uint8_t* proposal = (uint8_t*)w;                      w += ((size_t)n + 255) / 256 * 256;
  T* slot_val = (T*)w;                                  w += 8 * (size_t)CB200_SM_COUNT * 4 + 256;
  int* slot_idx = (int*)w;
  CB200_CUDA_TRY(cudaMemsetAsync(state, 0, sizeof(GreedyState), st));
  const T* e = (const T*)emb;
  const T* s = (const T*)seed_map;
  T two_bw2 = (T)(2.0 * (bandwidth * bandwidth));  // a python float in the reference, cast to the tensor dtype
  double thr = seed_thresh;  // `.item() < seed_thresh`: a python-float comparison in the reference
  void* args[] = {(void*)&e, (void*)&stride, (void*)&s, (void*)&n, (void*)&two_bw2, (void*)&min_object_size,
                  (void*)&thr, (void*)&min_unclustered_sum, (void*)&unclustered, (void*)&proposal, (void*)&instance,
                  (void*)&slot_val, (void*)&slot_idx, (void*)&state};
  CB200_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kernel, dim3(blocks), dim3(GREEDY_THREADS), args, 0, st));
  if (result2)
    CB200_CUDA_TRY(cudaMemcpyAsync(result2, &state->n_objects, 2 * sizeof(int), cudaMemcpyDeviceToDevice, st));
  return CB200_OK;
}

}  // namespace cb200

using namespace cb200;

extern "C" {

int64_t cb200_greedy_workspace_bytes(int64_t n_points) {
  return 256 + 2 * (((int64_t)n_points + 255) / 256 * 256) + 8 * (int64_t)CB200_SM_COUNT * 4 + 256 +
         4 * (int64_t)CB200_SM_COUNT * 4 + 1024;
}

int cb200_greedy_prepare(const void* emb, int dtype, int num_dims, const int64_t* spatial, const uint8_t* fg_mask,
                         int compute_dtype, double seed_min, double seed_max, void* emb_masked, int64_t capacity,
                         void* seed_masked, int32_t* pix_index, long long* n_out, void* workspace, void* stream) {
  if (!emb || !spatial || !fg_mask || !emb_masked || !seed_masked || !pix_index || !n_out || !workspace) return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_GP(TI, T, DD) \
  return greedy_prepare_typed<TI, T, DD>(emb, spatial, fg_mask, seed_min, seed_max, emb_masked, capacity, seed_masked, pix_index, n_out, workspace, st)
  if (num_dims == 2) {
    if (dtype == CB200_F32 && compute_dtype == CB200_F32) CB200_GP(float, float, 2);
    if (dtype == CB200_F64 && compute_dtype == CB200_F32) CB200_GP(double, float, 2);
    if (dtype == CB200_F64 && compute_dtype == CB200_F64) CB200_GP(double, double, 2);
    if (dtype == CB200_F32 && compute_dtype == CB200_F64) CB200_GP(float, double, 2);
  } else if (num_dims == 3) {
    if (dtype == CB200_F32 && compute_dtype == CB200_F32) CB200_GP(float, float, 3);
    if (dtype == CB200_F64 && compute_dtype == CB200_F32) CB200_GP(double, float, 3);
    if (dtype == CB200_F64 && compute_dtype == CB200_F64) CB200_GP(double, double, 3);
    if (dtype == CB200_F32 && compute_dtype == CB200_F64) CB200_GP(float, double, 3);
  }
#undef CB200_GP
  return CB200_EUNSUPPORTED;
}

int cb200_greedy_cluster(const void* emb_masked, int64_t stride, const void* seed_masked, int64_t n_points,
                         int num_dims, int compute_dtype, double bandwidth, int min_object_size, double seed_thresh,
                         long long min_unclustered_sum, short* instance_masked, int* n_objects_and_iterations,
                         void* workspace, void* stream) {
  if (!emb_masked || !seed_masked || !instance_masked || !workspace || n_points <= 0 || n_points > INT32_MAX)
    return CB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define CB200_GR(T, DD)                                                                                           \
  return greedy_run<T, DD>(emb_masked, stride, seed_masked, (int)n_points, bandwidth, min_object_size, seed_thresh, \
                           min_unclustered_sum, instance_masked, workspace, n_objects_and_iterations, st)
  if (num_dims == 2 && compute_dtype == CB200_F32) CB200_GR(float, 2);
  if (num_dims == 2 && compute_dtype == CB200_F64) CB200_GR(double, 2);
  if (num_dims == 3 && compute_dtype == CB200_F32) CB200_GR(float, 3);
  if (num_dims == 3 && compute_dtype == CB200_F64) CB200_GR(double, 3);
#undef CB200_GR
  return CB200_EUNSUPPORTED;
}

int cb200_scatter_i16(const short* src, const int32_t* pix_index, int64_t n, short* dst, void* stream) {
  if (!src || !pix_index || !dst || n < 0) return CB200_EINVAL;
  if (n == 0) return CB200_OK;
  scatter_i16_kernel<<<grid_for(n, 256, 2, 16), 256, 0, (cudaStream_t)stream>>>(src, pix_index, n, dst);
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}

}  // extern "C"
