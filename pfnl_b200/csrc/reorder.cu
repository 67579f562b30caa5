// Index-exact reorder kernels and the legacy-TF bicubic x4 resize.
//   pack_tokens     : tf.concat(frames,-1) + tf.space_to_depth(.,2)      model/pfnl.py:55-57
//   depth_to_space  : tf.depth_to_space (DCR) == modules/ps.py:_PS        model/pfnl.py:59,76,78
//   space_to_depth  : tf.space_to_depth                                   model/pfnl.py:57
//   bicubic4        : tf.image.resize_images(.,[4H,4W],method=2)          model/pfnl.py:63
// All HBM-bound; bytes = in + out.  The DCR reorder is a copy of contiguous b*Co-float
// segments (row (n,h*b+dy) of the output is the concatenation over w of
// in[n,h,w,dy*b*Co:(dy+1)*b*Co]), so it is vectorised by the widest of 16/8/4 bytes that
// divides the segment, with fully coalesced accesses on the contiguous side.
#include "common.cuh"
#include "bicubic.cuh"
#include "kernels.h"

namespace pfnl {

__global__ void pack_tokens_kernel(const float* __restrict__ lr, int N, int H, int W, float* __restrict__ tok) {
  const int W2 = W >> 1, H2 = H >> 1;
  const long long total = (long long)N * H2 * W2 * kNL;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(e % kNL);
    long long r = e / kNL;
    int w2 = (int)(r % W2);
    r /= W2;
    int h2 = (int)(r % H2);
    int n = (int)(r / H2);
    int q = ch / 21, rr = ch % 21;
    int dy = q >> 1, dx = q & 1;
    int t = rr / 3, c = rr % 3;
    tok[e] = lr[((((long long)n * kFrames + t) * H + (2 * h2 + dy)) * W + (2 * w2 + dx)) * 3 + c];
  }
}

// dir 0: depth_to_space (packed side = output), dir 1: space_to_depth (packed side = input).
// "packed" tensor P is [N,Hs*b,Ws*b,Co] viewed as segments of seg=b*Co floats; "deep" tensor D
// is [N,Hs,Ws,b*b*Co].  Vector v of the packed tensor maps to
//   D[((n*Hs+h)*Ws+w)*C + dy*seg + j*V],  v = ((((n*Hs+h)*b+dy)*Ws+w)*seg)/V + j.
template <typename VT, int V>
__global__ void dcr_reorder_kernel(const VT* __restrict__ src, VT* __restrict__ dst, long long nvec, int Hs, int Ws,
                                   int b, int segv /*seg/V*/, int Cv /*C/V*/, int dir) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
       v += (long long)gridDim.x * blockDim.x) {
    int j = (int)(v % segv);
    long long sidx = v / segv;
    int w = (int)(sidx % Ws);
    long long r = sidx / Ws;
    int dy = (int)(r % b);
    long long nh = r / b;  // n*Hs + h
    long long deep = (nh * Ws + w) * Cv + (long long)dy * segv + j;
    if (dir == 0)
      dst[v] = src[deep];
    else
      dst[deep] = src[v];
  }
}

static int launch_dcr(const float* src, float* dst, int N, int Hs, int Ws, int C, int b, int dir, cudaStream_t s) {
  // C = channels of the deep tensor = b*b*Co; segment = b*Co = C/b floats
  const int seg = C / b;
  const long long total = (long long)N * Hs * Ws * C;
  if (total == 0) return PFNL_OK;
  const bool a16 = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  const bool a8 = (((uintptr_t)src | (uintptr_t)dst) & 7) == 0;
  const int threads = 256;
  auto blocks_for = [&](long long nvec) {
    long long bl = (nvec + threads - 1) / threads;
    if (bl > 148LL * 16) bl = 148LL * 16;
    return (int)bl;
  };
  if (seg % 4 == 0 && a16) {
    long long nvec = total / 4;
    dcr_reorder_kernel<float4, 4><<<blocks_for(nvec), threads, 0, s>>>((const float4*)src, (float4*)dst, nvec, Hs, Ws,
                                                                     b, seg / 4, C / 4, dir);
  } else if (seg % 2 == 0 && a8) {
    long long nvec = total / 2;
    dcr_reorder_kernel<float2, 2><<<blocks_for(nvec), threads, 0, s>>>((const float2*)src, (float2*)dst, nvec, Hs, Ws,
                                                                     b, seg / 2, C / 2, dir);
  } else {
    dcr_reorder_kernel<float, 1><<<blocks_for(total), threads, 0, s>>>(src, dst, total, Hs, Ws, b, seg, C, dir);
  }
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_depth_to_space(const float* in, int N, int H, int W, int C, int b, float* out, cudaStream_t s) {
  return launch_dcr(in, out, N, H, W, C, b, 0, s);
}
int launch_space_to_depth(const float* in, int N, int H, int W, int C, int b, float* out, cudaStream_t s) {
  // deep tensor = output [N,H/b,W/b,C*b*b]
  return launch_dcr(in, out, N, H / b, W / b, C * b * b, b, 1, s);
}

int launch_pack_tokens(const float* lr, int N, int H, int W, float* tokens, cudaStream_t s) {
  long long total = (long long)N * (H / 2) * (W / 2) * kNL;
  if (total == 0) return PFNL_OK;
  int threads = 256;
  long long bl = (total + threads - 1) / threads;
  if (bl > 148LL * 16) bl = 148LL * 16;
  pack_tokens_kernel<<<(int)bl, threads, 0, s>>>(lr, N, H, W, tokens);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}


__global__ void bicubic4_kernel(const float* __restrict__ in, int N, int H, int W, int C, float* __restrict__ out) {
  const int H4 = H * 4, W4 = W * 4;
  const long long total = (long long)N * H4 * W4 * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    long long r = e / C;
    int X = (int)(r % W4);
    r /= W4;
    int Y = (int)(r % H4);
    int n = (int)(r / H4);
    out[e] = bicubic4_at(in + (long long)n * H * W * C, H, W, C, Y, X, c);
  }
}

int launch_bicubic4(const float* in, int N, int H, int W, int C, float* out, cudaStream_t s) {
  long long total = (long long)N * H * 4 * W * 4 * C;
  if (total == 0) return PFNL_OK;
  int threads = 256;
  long long bl = (total + threads - 1) / threads;
  if (bl > 148LL * 16) bl = 148LL * 16;
  bicubic4_kernel<<<(int)bl, threads, 0, s>>>(in, N, H, W, C, out);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

}  // namespace pfnl
