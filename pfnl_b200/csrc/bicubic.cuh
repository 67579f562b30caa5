// TF-1.12 legacy ResizeBicubic (align_corners=False, no half-pixel centres), x4 only.
// tf.image.resize_images(x[:,3],[4H,4W],method=2), model/pfnl.py:63.
#pragma once
#include "common.cuh"

namespace pfnl {

// Keys cubic A=-0.75 taps at phase p/4: the values of TF's 1024-entry coefficient table at
// offsets 0,256,512,768 (all exactly representable in fp32; each row sums to 1).
static __constant__ float kBicubicTaps[4][4] = {{0.f, 1.f, 0.f, 0.f},
                                                {-0.10546875f, 0.87890625f, 0.26171875f, -0.03515625f},
                                                {-0.09375f, 0.59375f, 0.59375f, -0.09375f},
                                                {-0.03515625f, 0.26171875f, 0.87890625f, -0.10546875f}};

// Output pixel (Y,X), channel c of the x4 resize of img [H,W,C].  Output index o=4k+p reads
// input indices clamp(k-1..k+2) with phase-p taps; x first then y, left-to-right sums
// (TF's Interpolate1D order).
__device__ __forceinline__ float bicubic4_at(const float* __restrict__ img, int H, int W, int C, int Y, int X, int c) {
  const int kx = X >> 2, px = X & 3, ky = Y >> 2, py = Y & 3;
  const int x0 = max(kx - 1, 0), x1 = kx, x2 = min(kx + 1, W - 1), x3 = min(kx + 2, W - 1);
  float rowv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int yy = min(max(ky - 1 + i, 0), H - 1);
    const float* row = img + (long long)yy * W * C + c;
    float v = row[(long long)x0 * C] * kBicubicTaps[px][0];
    v += row[(long long)x1 * C] * kBicubicTaps[px][1];
    v += row[(long long)x2 * C] * kBicubicTaps[px][2];
    v += row[(long long)x3 * C] * kBicubicTaps[px][3];
    rowv[i] = v;
  }
  float acc = rowv[0] * kBicubicTaps[py][0];
  acc += rowv[1] * kBicubicTaps[py][1];
  acc += rowv[2] * kBicubicTaps[py][2];
  acc += rowv[3] * kBicubicTaps[py][3];
  return acc;
}

}  // namespace pfnl
