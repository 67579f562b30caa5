// Per-clip MSE: eval_mse = tf.reduce_mean((SR-H)**2, axis=[2,3,4])  (model/pfnl.py:90).
// Two deterministic passes: kMseChunks partial sums per clip (double), then one finalize
// thread block per clip.  HBM-bound: bytes = 2 * N * per_clip * 4.
#include "common.cuh"
#include "kernels.h"

namespace pfnl {

__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ sr, const float* __restrict__ hr,
                                                          long long per_clip, double* __restrict__ partial) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const float* a = sr + (long long)n * per_clip;
  const float* b = hr + (long long)n * per_clip;
  const long long per_chunk = (per_clip + kMseChunks - 1) / kMseChunks;
  const long long beg = (long long)chunk * per_chunk;
  long long end = beg + per_chunk;
  if (end > per_clip) end = per_clip;
  double acc = 0.0;
  for (long long i = beg + threadIdx.x; i < end; i += 256) {
    const float d = a[i] - b[i];
    acc += (double)d * (double)d;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long long)n * kMseChunks + chunk] = red[0];
}

__global__ void mse_final_kernel(const double* __restrict__ partial, long long per_clip, float* __restrict__ mse,
                                 int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double s = 0.0;
  for (int c = 0; c < kMseChunks; ++c) s += partial[(long long)n * kMseChunks + c];
  mse[n] = (float)(s / (double)per_clip);
}

int launch_mse(const float* sr, const float* hr, int N, long long per_clip, double* partial, float* mse,
               cudaStream_t s) {
  if (N <= 0) return PFNL_OK;
  dim3 grid(kMseChunks, N);
  mse_partial_kernel<<<grid, 256, 0, s>>>(sr, hr, per_clip, partial);
  PFNL_LAUNCH_CHECK();
  mse_final_kernel<<<ceil_div(N, 128), 128, 0, s>>>(partial, per_clip, mse, N);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

}  // namespace pfnl
