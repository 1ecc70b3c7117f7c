// Measurement aid (not on the hot path): the FMA issue peak of the FP32 / FP64 pipes of THIS device, the
// denominator of the mean-shift kernels' roofline ("FP32 non-tensor peak to be measured by an FMA
// microbenchmark").  Every thread runs 8 independent FMA chains; flop = threads * 8 * iters * 2.
#include "common.cuh"

namespace cb200 {

template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, T seed, T* __restrict__ out) {
  T a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = seed + (T)(threadIdx.x + k);
  const T m = (T)0.999, c = (T)1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = a[k] * m + c;  // contracted to one FMA per chain
  }
  T s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == (T)-1) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace cb200

using namespace cb200;

extern "C" int cb200_fma_peak(int dtype, int iters, int blocks_per_sm, void* out, int64_t* flop, void* stream) {
  if (!out || !flop || iters <= 0 || blocks_per_sm <= 0) return CB200_EINVAL;
  const int blocks = CB200_SM_COUNT * blocks_per_sm;
  *flop = (int64_t)blocks * 256 * 8 * 2 * (int64_t)iters;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CB200_F32) fma_peak_kernel<float><<<blocks, 256, 0, st>>>(iters, 1.0f, (float*)out);
  else if (dtype == CB200_F64) fma_peak_kernel<double><<<blocks, 256, 0, st>>>(iters, 1.0, (double*)out);
  else return CB200_EUNSUPPORTED;
  CB200_LAUNCH_CHECK();
  return CB200_OK;
}
