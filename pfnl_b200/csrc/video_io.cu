// The steps immediately around the hot path in test_video_truth / test_video_lr:
//   downsample4_kernel  DownSample_4D (utils.py:169-192): REFLECT pad 6, depthwise 13x13 Gaussian
//                       (utils.py:95-105), stride 4, VALID  ->  HR [F,H,W,3] -> LR [F,ceil(H/4),ceil(W/4),3]
//   gather_windows      the sliding 7-frame window with edge clamping (model/pfnl.py:236-242, 294-300)
//                       done on the device: LR frames are uploaded once instead of 7 times
//   quantize_u8         round(clip(sr*255, 0, 255)).astype(uint8) (model/pfnl.py:255-257); np.round is
//                       round-half-to-even == cvt.rni
// All HBM-bound (bytes = in + out).
#include "common.cuh"
#include "kernels.h"

namespace pfnl {

__device__ __forceinline__ int reflect_idx(int i, int n) {  // tf.pad mode REFLECT (no edge repeat)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) downsample4_kernel(const float* __restrict__ hr, int F, int H, int W,
                                                          const float* __restrict__ blur /*[13*13]*/,
                                                          float* __restrict__ lr, int h, int w) {
  __shared__ float bs[169];
  if (threadIdx.x < 169) bs[threadIdx.x] = blur[threadIdx.x];
  __syncthreads();
  const long long total = (long long)F * h * w * 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % 3);
    long long r = e / 3;
    const int ox = (int)(r % w);
    r /= w;
    const int oy = (int)(r % h);
    const int f = (int)(r / h);
    const float* img = hr + (long long)f * H * W * 3 + c;
    float acc = 0.f;
    for (int i = 0; i < 13; ++i) {
      const int yy = reflect_idx(4 * oy + i - 6, H);
      const float* row = img + (long long)yy * W * 3;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int xx = reflect_idx(4 * ox + j - 6, W);
        acc = fmaf(row[xx * 3], bs[i * 13 + j], acc);
      }
    }
    lr[e] = acc;
  }
}

__global__ void __launch_bounds__(256) gather_windows_kernel(const float* __restrict__ frames, int F, long long fsz,
                                                             int first, int count, float* __restrict__ clips) {
  // clips[k][t] = frames[clamp(first + k + t - 3, 0, F-1)], float4 granularity when fsz % 4 == 0
  const long long total = (long long)count * kFrames * fsz;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long o = e % fsz;
    const long long kt = e / fsz;
    const int t = (int)(kt % kFrames);
    const int k = (int)(kt / kFrames);
    int src = first + k + t - kFrames / 2;
    src = src < 0 ? 0 : (src > F - 1 ? F - 1 : src);
    clips[e] = frames[(long long)src * fsz + o];
  }
}

__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ in, long long n,
                                                          unsigned char* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float v = in[e] * 255.f;
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[e] = (unsigned char)__float2int_rn(v);
  }
}

static int blocks_for(long long total) {
  long long bl = (total + 255) / 256;
  if (bl > 148LL * 16) bl = 148LL * 16;
  return (int)bl;
}

int launch_downsample4(const float* hr, int F, int H, int W, const float* blur_dev, float* lr, cudaStream_t s) {
  const int h = (H - 1) / 4 + 1, w = (W - 1) / 4 + 1;
  const long long total = (long long)F * h * w * 3;
  if (total <= 0) return PFNL_OK;
  downsample4_kernel<<<blocks_for(total), 256, 0, s>>>(hr, F, H, W, blur_dev, lr, h, w);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_gather_windows(const float* frames, int F, long long frame_elems, int first, int count, float* clips,
                          cudaStream_t s) {
  const long long total = (long long)count * kFrames * frame_elems;
  if (total <= 0) return PFNL_OK;
  gather_windows_kernel<<<blocks_for(total), 256, 0, s>>>(frames, F, frame_elems, first, count, clips);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_quantize_u8(const float* in, long long n, unsigned char* out, cudaStream_t s) {
  if (n <= 0) return PFNL_OK;
  quantize_u8_kernel<<<blocks_for(n), 256, 0, s>>>(in, n, out);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

}  // namespace pfnl
