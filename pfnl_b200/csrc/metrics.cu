// Quality metrics of the evaluation step that follows the hot path:
//   luma_kernel      to_uint8 + _rgb2ycbcr(...)[:,:,0] (utils.py:194-216): RGB fp32 -> Y (double),
//                    optionally rounded to uint8 as MATLAB's rgb2ycbcr does for uint8 images
//                    (matlab/compute_psnr.m:2-5, matlab/SSIM.m: "org=rgb2ycbcr(img1)")
//   ysq_partial      per-frame mean squared Y difference over the spatially cropped frame
//                    (utils.py AVG_PSNR: diff[.., sp_border:-sp_border, sp_border:-sp_border]; compute_psnr.m)
//   ssim_partial     Wang et al. SSIM as matlab/SSIM.m computes it: 11x11 Gaussian (sigma 1.5) window,
//                    'valid' filtering, K = (0.01, 0.03), L = 255, double precision, mean of the map
// All deterministic (fixed partial-sum layout, no atomics).  HBM-bound except ssim_partial, which does
// 5 x 121 double FMAs per output pixel out of a shared-memory tile.
#include "common.cuh"
#include "kernels.h"

namespace pfnl {

namespace {

constexpr int kSsimWin = 11;
constexpr int kSsimTileX = 32, kSsimTileY = 8;

__constant__ double c_ssim_window[kSsimWin * kSsimWin];

// to_uint8 (utils.py:211-214): x.astype(float32); (x-vmin)/(vmax-vmin)*255; clip(round(x),0,255) - fp32
// arithmetic, np.round = round-half-to-even
__device__ __forceinline__ float to_uint8_f(float x, float vmin, float vmax) {
  float v = (x - vmin) / (vmax - vmin) * 255.f;
  v = rintf(v);
  return fminf(fmaxf(v, 0.f), 255.f);
}

__global__ void __launch_bounds__(256) luma_kernel(const float* __restrict__ rgb, long long npix, float vmin, float vmax,
                                                   int round_y, double* __restrict__ y) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const double r = to_uint8_f(rgb[p * 3 + 0], vmin, vmax);
    const double g = to_uint8_f(rgb[p * 3 + 1], vmin, vmax);
    const double b = to_uint8_f(rgb[p * 3 + 2], vmin, vmax);
    // first row of T (utils.py:198) in np.dot's order, then + O[0] = 16 (utils.py:205)
    double v = r * 0.256788235294118 + g * 0.504129411764706 + b * 0.097905882352941;
    v += 16.0;
    if (round_y) v = floor(v + 0.5);  // MATLAB uint8(): round half away from zero (v > 0), saturate
    if (round_y) v = fmin(fmax(v, 0.0), 255.0);
    y[p] = v;
  }
}

// fixed-order tree sum over the 256 threads of a block; tid = linear thread index
__device__ __forceinline__ double block_sum_256(double v, double* red, int tid) {
  red[tid] = v;
  __syncthreads();
  for (int o = 128; o >= 1; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  return red[0];
}

// grid (kMetricChunks, F): rows of the cropped frame are dealt to chunks round-robin
__global__ void __launch_bounds__(256) ysq_partial_kernel(const double* __restrict__ ya, const double* __restrict__ yb,
                                                          int H, int W, int border, double* __restrict__ partial) {
  __shared__ double red[256];
  const int f = blockIdx.y, chunk = blockIdx.x;
  const double* a = ya + (long long)f * H * W;
  const double* b = yb + (long long)f * H * W;
  double acc = 0.0;
  for (int y = border + chunk; y < H - border; y += gridDim.x)
    for (int x = border + threadIdx.x; x < W - border; x += 256) {
      const double d = a[(long long)y * W + x] - b[(long long)y * W + x];
      acc += d * d;
    }
  const double s = block_sum_256(acc, red, threadIdx.x);
  if (threadIdx.x == 0) partial[(long long)f * gridDim.x + chunk] = s;
}

// grid (tiles_x, tiles_y, F), block 32x8: one thread per pixel of the 'valid' SSIM map
__global__ void __launch_bounds__(256) ssim_partial_kernel(const double* __restrict__ ya, const double* __restrict__ yb,
                                                           int H, int W, double C1, double C2,
                                                           double* __restrict__ partial) {
  constexpr int TW = kSsimTileX + kSsimWin - 1, TH = kSsimTileY + kSsimWin - 1;
  __shared__ double ta[TH][TW], tb[TH][TW];
  __shared__ double red[256];
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * kSsimTileX, y0 = blockIdx.y * kSsimTileY;
  const double* a = ya + (long long)f * H * W;
  const double* b = yb + (long long)f * H * W;
  const int tid = threadIdx.y * kSsimTileX + threadIdx.x;
  for (int i = tid; i < TH * TW; i += 256) {
    const int ty = i / TW, tx = i % TW;
    const int gy = y0 + ty, gx = x0 + tx;
    const bool in = gy < H && gx < W;
    ta[ty][tx] = in ? a[(long long)gy * W + gx] : 0.0;
    tb[ty][tx] = in ? b[(long long)gy * W + gx] : 0.0;
  }
  __syncthreads();
  const int ox = x0 + threadIdx.x, oy = y0 + threadIdx.y;
  double val = 0.0;
  if (ox < W - kSsimWin + 1 && oy < H - kSsimWin + 1) {
    double mu1 = 0, mu2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int i = 0; i < kSsimWin; ++i)
#pragma unroll
      for (int j = 0; j < kSsimWin; ++j) {
        const double w = c_ssim_window[i * kSsimWin + j];
        const double p = ta[threadIdx.y + i][threadIdx.x + j], q = tb[threadIdx.y + i][threadIdx.x + j];
        mu1 += w * p;
        mu2 += w * q;
        s11 += w * (p * p);
        s22 += w * (q * q);
        s12 += w * (p * q);
      }
    const double mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const double sg1 = s11 - mu1_sq, sg2 = s22 - mu2_sq, sg12 = s12 - mu12;
    val = ((2.0 * mu12 + C1) * (2.0 * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2));
  }
  __syncthreads();
  const double s = block_sum_256(val, red, tid);
  if (tid == 0) partial[((long long)f * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
}

// out[f] = sum(partial[f][0..n)) * scale
__global__ void metric_final_kernel(const double* __restrict__ partial, int n, double scale, double* __restrict__ out,
                                    int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double s = 0.0;
  for (int c = 0; c < n; ++c) s += partial[(long long)f * n + c];
  out[f] = s * scale;
}

int blocks_for(long long total) {
  long long bl = (total + 255) / 256;
  if (bl > 148LL * 16) bl = 148LL * 16;
  return (int)bl;
}

}  // namespace

int init_metrics() {
  // fspecial('gaussian', 11, 1.5) normalised to sum 1 (matlab/SSIM.m: window/sum(sum(window)))
  double w[kSsimWin * kSsimWin], sum = 0.0;
  for (int i = 0; i < kSsimWin; ++i)
    for (int j = 0; j < kSsimWin; ++j) {
      const double y = i - 5, x = j - 5;
      w[i * kSsimWin + j] = exp(-(x * x + y * y) / (2.0 * 1.5 * 1.5));
      sum += w[i * kSsimWin + j];
    }
  for (double& v : w) v /= sum;
  PFNL_CUDA(cudaMemcpyToSymbol(c_ssim_window, w, sizeof(w)));
  return PFNL_OK;
}

size_t metric_partials(int F, int H, int W) {
  const size_t ssim = (size_t)F * ceil_div(W, kSsimTileX) * ceil_div(H, kSsimTileY);
  const size_t psnr = (size_t)F * kMetricChunks;
  return ssim > psnr ? ssim : psnr;
}

int launch_luma(const float* rgb, long long npix, float vmin, float vmax, int round_y, double* y, cudaStream_t s) {
  if (npix <= 0) return PFNL_OK;
  luma_kernel<<<blocks_for(npix), 256, 0, s>>>(rgb, npix, vmin, vmax, round_y, y);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_ysq(const double* ya, const double* yb, int F, int H, int W, int border, double* partial, double* out,
               cudaStream_t s) {
  dim3 grid(kMetricChunks, F);
  ysq_partial_kernel<<<grid, 256, 0, s>>>(ya, yb, H, W, border, partial);
  PFNL_LAUNCH_CHECK();
  const double n = (double)(H - 2 * border) * (double)(W - 2 * border);
  metric_final_kernel<<<ceil_div(F, 128), 128, 0, s>>>(partial, kMetricChunks, 1.0 / n, out, F);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_ssim(const double* ya, const double* yb, int F, int H, int W, double* partial, double* out,
                cudaStream_t s) {
  dim3 grid(ceil_div(W, kSsimTileX), ceil_div(H, kSsimTileY), F);
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  ssim_partial_kernel<<<grid, dim3(kSsimTileX, kSsimTileY), 0, s>>>(ya, yb, H, W, C1, C2, partial);
  PFNL_LAUNCH_CHECK();
  const double n = (double)(H - kSsimWin + 1) * (double)(W - kSsimWin + 1);
  metric_final_kernel<<<ceil_div(F, 128), 128, 0, s>>>(partial, (int)(grid.x * grid.y), 1.0 / n, out, F);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

}  // namespace pfnl
