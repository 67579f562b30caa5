// Hardware probe (run once on a B200 through gpurun; results are recorded in DESIGN.md):
//  1. are the hand-built UMMA instruction / shared-memory descriptors right (plain GEMM tile)?
//  2. can a 3x3 conv tap be fed by pointing the A descriptor at a row-shifted window of ONE
//     TMA-loaded halo patch (128B swizzle), for patch row pitches of 16 and 10 pixels, and does
//     the descriptor base-offset field matter?   (decides the conv kernel's smem layout)
//  3. how exact is the fp32 accumulation in TMEM (round-to-nearest vs truncation), and what is
//     the end-to-end error of the 3-pass hi/lo fp16 split against an fp64 convolution?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probes/bin/umma_probe probes/umma_probe.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../pfnl_b200/csrc/tc_ptx.cuh"
#include "../pfnl_b200/csrc/tc_tmap.h"

using namespace pfnl::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int TH = 16, TW = 8;  // output tile: 16 rows x 8 cols = 128 pixels (M)
constexpr int PH = TH + 2;

struct Params {
  int PW;        // patch row pitch in pixels (box width of the TMA load)
  int taps;      // 1: centre tap only, 9: full 3x3, 0: tap (0,0) only (unshifted window)
  int bo_mode;   // 0: base_offset = 0, 1: base_offset = (start>>7)&7
  int nsplit;    // 1: single fp16 pass, 2: hi/lo split, 3 MMAs per k-step
  int y0, x0;    // tile origin in the image
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                    const __grid_constant__ CUtensorMap tm_lo,
                                                    const __half* __restrict__ w_hi, const __half* __restrict__ w_lo,
                                                    float* __restrict__ out, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* patch_hi = smem;                       // 18*16*128 = 36864
  uint8_t* patch_lo = smem + 36864;               // 36864
  uint8_t* wsm_hi = smem + 2 * 36864;             // 9*8192 = 73728
  uint8_t* wsm_lo = wsm_hi + 73728;               // 73728
  uint64_t* bars = (uint64_t*)(wsm_lo + 73728);   // [0] full, [1] mma done
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const uint32_t patch_bytes = PH * p.PW * 128;
  if (tid == 0) {
    uint32_t tx = patch_bytes * p.nsplit + 73728 * p.nsplit;
    mbar_arrive_expect_tx(&bars[0], tx);
    tma_load_4d(patch_hi, &tm_hi, &bars[0], 0, p.x0 - 1, p.y0 - 1, 0);
    bulk_load(wsm_hi, w_hi, 73728, &bars[0]);
    if (p.nsplit == 2) {
      tma_load_4d(patch_lo, &tm_lo, &bars[0], 0, p.x0 - 1, p.y0 - 1, 0);
      bulk_load(wsm_lo, w_lo, 73728, &bars[0]);
    }
  }
  mbar_wait(&bars[0], 0);
  fence_after_sync();

  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, 64);
    const uint32_t sbo_a = p.PW * 128;
    bool first0 = true, first1 = true;
    for (int tap = 0; tap < 9; ++tap) {
      if (p.taps == 1 && tap != 4) continue;
      if (p.taps == 0 && tap != 0) continue;
      const int dy = tap / 3, dx = tap % 3;
      const uint32_t shift = (dy * p.PW + dx) * 128;
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_hi = smem_u32(patch_hi) + shift + k * 32;
        const uint32_t a_lo = smem_u32(patch_lo) + shift + k * 32;
        const uint32_t b_hi = smem_u32(wsm_hi) + tap * 8192 + k * 32;
        const uint32_t b_lo = smem_u32(wsm_lo) + tap * 8192 + k * 32;
        const uint32_t bo_hi = p.bo_mode ? ((a_hi >> 7) & 7) : 0;
        const uint32_t bo_lo = p.bo_mode ? ((a_lo >> 7) & 7) : 0;
        mma_f16(tmem, make_sdesc_sw128(a_hi, sbo_a, bo_hi), make_sdesc_sw128(b_hi, 1024, 0), idesc, first0 ? 0u : 1u);
        first0 = false;
        if (p.nsplit == 2) {
          mma_f16(tmem + 64, make_sdesc_sw128(a_hi, sbo_a, bo_hi), make_sdesc_sw128(b_lo, 1024, 0), idesc,
                  first1 ? 0u : 1u);
          first1 = false;
          mma_f16(tmem + 64, make_sdesc_sw128(a_lo, sbo_a, bo_lo), make_sdesc_sw128(b_hi, 1024, 0), idesc, 1u);
        }
      }
    }
    mma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  fence_after_sync();

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (rows of the tile)
  const int row = warp * 32 + (tid & 31);
  for (int half = 0; half < 2; ++half) {
    uint32_t r0[32], r1[32];
    tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, r0);
    if (p.nsplit == 2) tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + half * 32, r1);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) {
      float v = __uint_as_float(r0[j]);
      if (p.nsplit == 2) v += __uint_as_float(r1[j]) * (1.0f / 2048.0f);
      out[row * 64 + half * 32 + j] = v;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- accumulate-rounding micro test: D = 1.0, then n MMAs each adding exactly 0.75 ulp(1) -------------
__global__ void __launch_bounds__(128) accum_kernel(const __half* __restrict__ a_img, const __half* __restrict__ b_one,
                                                    const __half* __restrict__ b_tiny, float* __restrict__ out,
                                                    int n_adds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_sm = smem;            // 128 x 64 fp16 = 16384
  uint8_t* b1_sm = smem + 16384;   // 64 x 64 = 8192
  uint8_t* b2_sm = b1_sm + 8192;
  uint64_t* bars = (uint64_t*)(b2_sm + 8192);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384 + 8192 * 2);
    bulk_load(a_sm, a_img, 16384, &bars[0]);
    bulk_load(b1_sm, b_one, 8192, &bars[0]);
    bulk_load(b2_sm, b_tiny, 8192, &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, 64);
    // k-step 0 of A holds [1, 0, ...] per row; k-step 1 holds [1.5*2^-12, 0, ...]
    mma_f16(tmem, make_sdesc_sw128(smem_u32(a_sm), 1024, 0), make_sdesc_sw128(smem_u32(b1_sm), 1024, 0), idesc, 0u);
    for (int i = 0; i < n_adds; ++i)
      mma_f16(tmem, make_sdesc_sw128(smem_u32(a_sm) + 32, 1024, 0), make_sdesc_sw128(smem_u32(b2_sm) + 32, 1024, 0),
              idesc, 1u);
    mma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  fence_after_sync();
  uint32_t r[32];
  tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  if ((tid & 31) == 0) out[warp] = __uint_as_float(r[0]);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static void swizzle_tile(const std::vector<__half>& rowmajor /*rows x 64*/, int rows, __half* dst) {
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 8; ++c)
      memcpy((uint8_t*)dst + sw128_offset(r, c), &rowmajor[(size_t)r * 64 + c * 8], 16);
}

int main() {
  const int H = 40, W = 24;
  srand(7);
  std::vector<float> x((size_t)H * W * 64), wt(9 * 64 * 64);  // x[h][w][c], wt[tap][co][ci]
  for (auto& v : x) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : wt) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.06f;
  std::vector<__half> x_hi(x.size()), x_lo(x.size());
  for (size_t i = 0; i < x.size(); ++i) {
    x_hi[i] = __float2half_rn(x[i]);
    x_lo[i] = __float2half_rn((x[i] - __half2float(x_hi[i])) * 2048.f);
  }
  std::vector<__half> w_hi_img(9 * 4096), w_lo_img(9 * 4096);
  std::vector<float> w_hi_f(wt.size()), w_lo_f(wt.size());
  for (int tap = 0; tap < 9; ++tap) {
    std::vector<__half> hi(4096), lo(4096);
    for (int i = 0; i < 4096; ++i) {
      float v = wt[tap * 4096 + i];
      hi[i] = __float2half_rn(v);
      lo[i] = __float2half_rn((v - __half2float(hi[i])) * 2048.f);
      w_hi_f[tap * 4096 + i] = __half2float(hi[i]);
      w_lo_f[tap * 4096 + i] = __half2float(lo[i]);
    }
    swizzle_tile(hi, 64, &w_hi_img[tap * 4096]);
    swizzle_tile(lo, 64, &w_lo_img[tap * 4096]);
  }
  __half *d_xhi, *d_xlo, *d_whi, *d_wlo;
  float* d_out;
  CK(cudaMalloc(&d_xhi, x.size() * 2));
  CK(cudaMalloc(&d_xlo, x.size() * 2));
  CK(cudaMalloc(&d_whi, 9 * 4096 * 2));
  CK(cudaMalloc(&d_wlo, 9 * 4096 * 2));
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaMemcpy(d_xhi, x_hi.data(), x.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_xlo, x_lo.data(), x.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_whi, w_hi_img.data(), 9 * 4096 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_wlo, w_lo_img.data(), 9 * 4096 * 2, cudaMemcpyHostToDevice));
  const int smem_bytes = 2 * 36864 + 2 * 73728 + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));

  auto run = [&](Params p, const char* label) {
    CUtensorMap tmh, tml;
    int r1 = make_act_tmap(&tmh, d_xhi, 1, H, W, p.PW, PH);
    int r2 = make_act_tmap(&tml, d_xlo, 1, H, W, p.PW, PH);
    if (r1 || r2) {
      printf("%s: tensor map encode failed (%d,%d)\n", label, r1, r2);
      return;
    }
    CK(cudaMemset(d_out, 0xff, 128 * 64 * 4));
    probe_kernel<<<1, 128, smem_bytes>>>(tmh, tml, d_whi, d_wlo, d_out, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: kernel failed: %s\n", label, cudaGetErrorString(e));
      exit(3);
    }
    std::vector<float> out(128 * 64);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    // references: (a) double conv of the fp16-rounded hi operands, (b) double conv of the fp32 data
    double err_hi = 0, err_full = 0, ref_max = 0, err_f32 = 0;
    for (int m = 0; m < 128; ++m) {
      int ty = m / TW, tx = m % TW;
      for (int co = 0; co < 64; ++co) {
        double acc_hi = 0, acc_full = 0;
        float acc_f32 = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
          if (p.taps == 1 && tap != 4) continue;
          if (p.taps == 0 && tap != 0) continue;
          int yy = p.y0 + ty + tap / 3 - 1, xx = p.x0 + tx + tap % 3 - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          for (int ci = 0; ci < 64; ++ci) {
            size_t xi = ((size_t)yy * W + xx) * 64 + ci;
            size_t wi = (size_t)tap * 4096 + co * 64 + ci;
            acc_hi += (double)__half2float(x_hi[xi]) * (double)w_hi_f[wi];
            acc_full += (double)x[xi] * (double)wt[wi];
            acc_f32 = fmaf(x[xi], wt[wi], acc_f32);
          }
        }
        double got = out[m * 64 + co];
        err_hi = fmax(err_hi, fabs(got - acc_hi));
        err_full = fmax(err_full, fabs(got - acc_full));
        err_f32 = fmax(err_f32, fabs((double)acc_f32 - acc_full));
        ref_max = fmax(ref_max, fabs(acc_full));
      }
    }
    printf("%-46s PW=%2d taps=%d bo=%d nsplit=%d origin=(%d,%d): |ref|max=%.3f  err_vs_fp16operands=%.3e  "
           "err_vs_fp64(fp32 data)=%.3e  [host fp32 fma chain err=%.3e]\n",
           label, p.PW, p.taps, p.bo_mode, p.nsplit, p.y0, p.x0, ref_max, err_hi, err_full, err_f32);
  };

  printf("== umma_probe ==\n");
  for (int PW : {16, 10}) {
    for (int bo : {0, 1}) {
      run({PW, 0, bo, 1, 8, 8}, "tap(0,0) unshifted window, interior");
      run({PW, 1, bo, 1, 8, 8}, "centre tap (shift PW+1 rows), interior");
      run({PW, 9, bo, 1, 8, 8}, "3x3, interior");
      run({PW, 9, bo, 1, 0, 0}, "3x3, top-left corner (negative TMA coords)");
      run({PW, 9, bo, 1, 32, 16}, "3x3, bottom-right ragged (OOB fill)");
    }
  }
  run({16, 9, 1, 2, 8, 8}, "3x3 hi/lo split x3, PW16 bo1");
  run({16, 9, 0, 2, 8, 8}, "3x3 hi/lo split x3, PW16 bo0");
  run({10, 9, 0, 2, 8, 8}, "3x3 hi/lo split x3, PW10 bo0");

  // accumulate rounding
  {
    std::vector<__half> a(128 * 64, __float2half(0.f)), b1(64 * 64, __float2half(0.f)), b2(64 * 64, __float2half(0.f));
    for (int r = 0; r < 128; ++r) {
      a[r * 64 + 0] = __float2half(1.0f);
      a[r * 64 + 16] = __float2half(1.5f * 0.000244140625f);  // 1.5 * 2^-12, k-step 1
    }
    for (int n = 0; n < 64; ++n) {
      b1[n * 64 + 0] = __float2half(1.0f);
      b2[n * 64 + 16] = __float2half(0.000244140625f);        // 2^-12
    }
    std::vector<__half> ai(128 * 64), b1i(4096), b2i(4096);
    swizzle_tile(a, 128, ai.data());
    swizzle_tile(b1, 64, b1i.data());
    swizzle_tile(b2, 64, b2i.data());
    __half *da, *db1, *db2;
    float* dout;
    CK(cudaMalloc(&da, 16384));
    CK(cudaMalloc(&db1, 8192));
    CK(cudaMalloc(&db2, 8192));
    CK(cudaMalloc(&dout, 16));
    CK(cudaMemcpy(da, ai.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db1, b1i.data(), 8192, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db2, b2i.data(), 8192, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 33000));
    for (int n : {1, 4, 16, 64}) {
      accum_kernel<<<1, 128, 32768 + 64>>>(da, db1, db2, dout, n);
      CK(cudaDeviceSynchronize());
      float o[4];
      CK(cudaMemcpy(o, dout, 16, cudaMemcpyDeviceToHost));
      const double ulp = ldexp(1.0, -23);
      printf("accumulate 1.0 + %2d x 0.75ulp: got 1+%.2f ulp  (exact 1+%.2f ulp; RN-each-step 1+%d ulp; RZ-each-step 1+0 ulp)\n",
             n, (o[0] - 1.0) / ulp, 0.75 * n, n);
    }
  }
  printf("== done ==\n");
  return 0;
}
