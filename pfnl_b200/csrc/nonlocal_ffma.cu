// Non-local block, fp32 FFMA path (NonLocalBlock nltype=1 'gaussian', sub_sample=1; utils.py:18-71):
//   G = X*Wg + bg              (utils.py:26)        nl_linear_kernel
//   S = X*X^T  (theta = phi = X, utils.py:33-34,41-42,53)
//   P = exp(S) / rowsum(exp(S)) (utils.py:57-58)     nl_flash_ffma_kernel: streamed over key
//   Y = P*G                    (utils.py:64)          tiles with an online (max-subtracted)
//                                                     softmax - the L x L matrix is never stored
//   Z = Y*Ww + bw              (utils.py:67)        nl_linear_kernel<scatter>: + depth_to_space(.,2)
//   inp0 += depth_to_space(Z,2) (model/pfnl.py:59-60)   and the residual add, fused in its store
// The max-subtracted softmax equals the reference's naive exp/sum wherever the latter is
// finite in fp32 (it overflows for L >~ 113 on bright flat inputs; that is not reproduced).
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace pfnl {

// Y[rows,84] = X[rows,84] * Wm[84,84] + b.  CTA = 12 * ITER rows, 252 active threads:
// thread -> 4 output columns (21 column groups) x ITER rows (12 row groups).  Every CTA stages the whole
// 28 KB weight matrix, so ITER grows with the row count: 1 (12 rows per CTA) keeps the latency-bound
// small cases (bench: 4096 rows) spread over > 2 waves, 4 (48 rows) amortises the weights for long clips.
template <bool SCATTER, int ITER>
__global__ void __launch_bounds__(256) nl_linear_kernel(const float* __restrict__ X, int rows,
                                                        const float* __restrict__ Wm, const float* __restrict__ b,
                                                        float* __restrict__ Y, const float* __restrict__ lr, int H,
                                                        int W) {
  __shared__ __align__(16) float wsm[kNL * kNL];
  constexpr int kLinRows = 12 * ITER, kLinIter = ITER;
  __shared__ __align__(16) float xsm[kLinRows * kNL];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * kLinRows;
  // staging is the latency of this kernel: 16-byte loads, all issued before the first use
  {
    constexpr int NW4 = kNL * kNL / 4;  // 1764 float4
    const float4* w4 = reinterpret_cast<const float4*>(Wm);
    float4 wv[(NW4 + 255) / 256];
#pragma unroll
    for (int u = 0; u < (NW4 + 255) / 256; ++u) {
      const int i = tid + u * 256;
      wv[u] = i < NW4 ? w4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    constexpr int NX4 = kLinRows * kNL / 4;  // a row is 21 float4
    const float4* x4 = reinterpret_cast<const float4*>(X + (long long)row0 * kNL);
    float4 xv[(NX4 + 255) / 256];
#pragma unroll
    for (int u = 0; u < (NX4 + 255) / 256; ++u) {
      const int i = tid + u * 256;
      const bool ok = i < NX4 && row0 + i / (kNL / 4) < rows;
      xv[u] = ok ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < (NW4 + 255) / 256; ++u) {
      const int i = tid + u * 256;
      if (i < NW4) reinterpret_cast<float4*>(wsm)[i] = wv[u];
    }
#pragma unroll
    for (int u = 0; u < (NX4 + 255) / 256; ++u) {
      const int i = tid + u * 256;
      if (i < NX4) reinterpret_cast<float4*>(xsm)[i] = xv[u];
    }
  }
  __syncthreads();
  if (tid >= 252) return;
  const int cgp = tid % 21, rg = tid / 21;
  float acc[kLinIter][4];
#pragma unroll
  for (int i = 0; i < kLinIter; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < kNL; ++k) {
    const float4 bv = *reinterpret_cast<const float4*>(wsm + k * kNL + cgp * 4);
#pragma unroll
    for (int i = 0; i < kLinIter; ++i) {
      const int rl = rg + 12 * i;
      const float av = xsm[rl * kNL + k];
      acc[i][0] = fmaf(av, bv.x, acc[i][0]);
      acc[i][1] = fmaf(av, bv.y, acc[i][1]);
      acc[i][2] = fmaf(av, bv.z, acc[i][2]);
      acc[i][3] = fmaf(av, bv.w, acc[i][3]);
    }
  }
  const float4 bs = *reinterpret_cast<const float4*>(b + cgp * 4);
#pragma unroll
  for (int i = 0; i < kLinIter; ++i) {
    const int rl = rg + 12 * i;
    const long long row = (long long)row0 + rl;
    if (row >= rows) continue;
    const float v[4] = {acc[i][0] + bs.x, acc[i][1] + bs.y, acc[i][2] + bs.z, acc[i][3] + bs.w};
    if (!SCATTER) {
      *reinterpret_cast<float4*>(Y + row * kNL + cgp * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      // row = n*L + h2*W2 + w2 ; channel ch = (dy*2+dx)*21 + t*3 + c
      const int W2 = W >> 1, L = (H >> 1) * W2;
      const int n = (int)(row / L), tok = (int)(row % L);
      const int h2 = tok / W2, w2 = tok % W2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = cgp * 4 + j;
        const int q = ch / 21, rr = ch % 21;
        const int y = 2 * h2 + (q >> 1), x = 2 * w2 + (q & 1);
        const int t = rr / 3, c = rr % 3;
        const float xin = lr[((((long long)n * kFrames + t) * H + y) * W + x) * 3 + c];
        Y[(((long long)n * H + y) * W + x) * 21 + rr] = xin + v[j];
      }
    }
  }
}

int launch_nl_linear(const float* X, int rows, const float* Wm, const float* b, float* Y, cudaStream_t s) {
  if (rows <= 0) return PFNL_OK;
  if (rows > 8192)
    nl_linear_kernel<false, 4><<<ceil_div(rows, 48), 256, 0, s>>>(X, rows, Wm, b, Y, nullptr, 0, 0);
  else
    nl_linear_kernel<false, 1><<<ceil_div(rows, 12), 256, 0, s>>>(X, rows, Wm, b, Y, nullptr, 0, 0);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

int launch_nl_linear_scatter(const float* Yin, const float* lr, int N, int H, int W, const float* Ww,
                             const float* bw, float* inp21, cudaStream_t s) {
  const int rows = N * (H / 2) * (W / 2);
  if (rows <= 0) return PFNL_OK;
  if (rows > 8192)
    nl_linear_kernel<true, 4><<<ceil_div(rows, 48), 256, 0, s>>>(Yin, rows, Ww, bw, inp21, lr, H, W);
  else
    nl_linear_kernel<true, 1><<<ceil_div(rows, 12), 256, 0, s>>>(Yin, rows, Ww, bw, inp21, lr, H, W);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

// Flash-style streamed attention with Q = K = X, V = G.  CTA = 32 queries of one clip,
// 128 threads; key tiles of 64.
constexpr int kBQ = 32, kBK = 64;
constexpr int kPP = kBK + 1;  // P row pitch
constexpr int kFlashSmemBytes = (kBQ * kNL + 2 * kBK * kNL + kBQ * kPP + 2 * kBQ) * 4;

__global__ void __launch_bounds__(128) nl_flash_ffma_kernel(const float* __restrict__ X, const float* __restrict__ G,
                                                            int L, float* __restrict__ Y) {
  extern __shared__ __align__(16) float nl_smem[];
  float* qs = nl_smem;                 // [kBQ][84]
  float* ks = qs + kBQ * kNL;          // [kBK][84]
  float* vs = ks + kBK * kNL;          // [kBK][84]
  float* ps = vs + kBK * kNL;          // [kBQ][kPP]
  float* row_alpha = ps + kBQ * kPP;   // [kBQ]
  float* row_l = row_alpha + kBQ;      // [kBQ]
  const int tid = threadIdx.x;
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * kBQ;
  const float* Xn = X + (long long)n * L * kNL;
  const float* Gn = G + (long long)n * L * kNL;

  for (int i = tid; i < kBQ * kNL; i += 128) {
    int q = q0 + i / kNL;
    qs[i] = q < L ? Xn[(long long)q0 * kNL + i] : 0.f;
  }
  // S-phase mapping: 8 query groups (4 queries) x 16 key lanes (keys tk, tk+16, tk+32, tk+48)
  const int tq = tid >> 4, tk = tid & 15;
  // O-phase mapping: 8 query groups x 14 column groups of 6 (112 active threads)
  const int oq = tid / 14, oc = tid % 14;
  const bool o_active = tid < 112;
  float m_run[4], l_part[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_part[i] = 0.f;
  }
  float oacc[4][6];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) oacc[i][j] = 0.f;

  for (int k0 = 0; k0 < L; k0 += kBK) {
    __syncthreads();  // previous tile fully consumed (also orders the qs fill on the first pass)
    for (int i = tid; i < kBK * kNL; i += 128) {
      int kk = k0 + i / kNL;
      bool ok = kk < L;
      ks[i] = ok ? Xn[(long long)k0 * kNL + i] : 0.f;
      vs[i] = ok ? Gn[(long long)k0 * kNL + i] : 0.f;
    }
    __syncthreads();
    // S = Q K^T for this thread's 4x4 block
    float sacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[i][j] = 0.f;
#pragma unroll 3
    for (int k4 = 0; k4 < kNL / 4; ++k4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(qs + (tq * 4 + i) * kNL + k4 * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(ks + (tk + 16 * j) * kNL + k4 * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sacc[i][j] = fmaf(qv[i].x, kv[j].x, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].y, kv[j].y, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].z, kv[j].z, sacc[i][j]);
          sacc[i][j] = fmaf(qv[i].w, kv[j].w, sacc[i][j]);
        }
    }
    // online softmax: row max over the 64 keys (4 local + 16 lanes of the half-warp)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mt = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (k0 + tk + 16 * j >= L) sacc[i][j] = -INFINITY;
        mt = fmaxf(mt, sacc[i][j]);
      }
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
      const float m_new = fmaxf(m_run[i], mt);  // finite: every tile has >= 1 valid key
      const float alpha = expf(m_run[i] - m_new);
      m_run[i] = m_new;
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(sacc[i][j] - m_new);
        psum += p;
        ps[(tq * 4 + i) * kPP + tk + 16 * j] = p;
      }
      l_part[i] = l_part[i] * alpha + psum;
      if (tk == 0) row_alpha[tq * 4 + i] = alpha;
    }
    __syncthreads();
    // O = O*alpha + P V
    if (o_active) {
      float al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        al[i] = row_alpha[oq * 4 + i];
#pragma unroll
        for (int j = 0; j < 6; ++j) oacc[i][j] *= al[i];
      }
#pragma unroll 4
      for (int kk = 0; kk < kBK; ++kk) {
        const float2 v0 = *reinterpret_cast<const float2*>(vs + kk * kNL + oc * 6);
        const float2 v1 = *reinterpret_cast<const float2*>(vs + kk * kNL + oc * 6 + 2);
        const float2 v2 = *reinterpret_cast<const float2*>(vs + kk * kNL + oc * 6 + 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = ps[(oq * 4 + i) * kPP + kk];
          oacc[i][0] = fmaf(p, v0.x, oacc[i][0]);
          oacc[i][1] = fmaf(p, v0.y, oacc[i][1]);
          oacc[i][2] = fmaf(p, v1.x, oacc[i][2]);
          oacc[i][3] = fmaf(p, v1.y, oacc[i][3]);
          oacc[i][4] = fmaf(p, v2.x, oacc[i][4]);
          oacc[i][5] = fmaf(p, v2.y, oacc[i][5]);
        }
      }
    }
  }
  // row sums: reduce the 16 key lanes' partial sums
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float l = l_part[i];
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (tk == 0) row_l[tq * 4 + i] = l;
  }
  __syncthreads();
  if (o_active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + oq * 4 + i;
      if (q >= L) continue;
      const float inv = 1.f / row_l[oq * 4 + i];
      float* o = Y + ((long long)n * L + q) * kNL + oc * 6;
#pragma unroll
      for (int j = 0; j < 6; ++j) o[j] = oacc[i][j] * inv;
    }
  }
}

int init_nonlocal_ffma() {
  PFNL_CUDA(cudaFuncSetAttribute(nl_flash_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlashSmemBytes));
  return PFNL_OK;
}

int launch_nl_flash_ffma(const float* X, const float* G, int N, int L, float* Y, cudaStream_t s) {
  if (N <= 0 || L <= 0) return PFNL_OK;
  dim3 grid(ceil_div(L, kBQ), N);
  nl_flash_ffma_kernel<<<grid, 128, kFlashSmemBytes, s>>>(X, G, L, Y);
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

}  // namespace pfnl
