// tcgen05.mma throughput probe (one CTA per SM, all 148 SMs busy): cycles per MMA for
//   N in {64, 128, 256}, A descriptors standard (SBO 1024) vs row-shifted halo windows (SBO 1280),
//   same-A-twice (does the hardware reuse an operand?), and cta_group::1 only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/bin/mma_rate_probe probes/mma_rate_probe.cu
#include <stdio.h>
#include <stdlib.h>

#include "../pfnl_b200/csrc/tc_ptx.cuh"

using namespace pfnl::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mode 0: A standard tile (SBO 1024), distinct k-slices round robin
// mode 1: A = 3x3 halo-window descriptors (SBO 1280, shifted starts), like conv_tc
// mode 2: like 0 but every MMA uses the SAME A and B descriptors
// mode 3: alternate N=128 / N=64 MMAs (the x3 hi/lo mix)
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int n_mma, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;              // 48 KB region (patch-like)
  uint8_t* b_sm = smem + 49152;      // up to 9 x 32 KB? keep 9 taps x N*128
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill smem with something finite
  for (int i = tid; i < (49152 + 3 * N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 256);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
    const uint64_t ad_std = make_sdesc_sw128(smem_u32(a_sm), 1024, 0);
    const uint64_t ad_halo = make_sdesc_sw128(smem_u32(a_sm), 1280, 0);
    const uint64_t bd = make_sdesc_sw128(smem_u32(b_sm), 1024, 0);
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n_mma; i += 36) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t a, b;
            if (mode == 1) {
              a = ad_halo + ((((tp / 3) * 10 + (tp % 3)) * 128 + k * 32) >> 4);
              b = bd + (((tp % 3) * N * 128 + k * 32) >> 4);
            } else if (mode == 2) {
              a = ad_std;
              b = bd;
            } else {
              a = ad_std + ((tp * 2048 + k * 32) >> 4);
              b = bd + (((tp % 3) * N * 128 + k * 32) >> 4);
            }
            if (mode == 3 && (k & 1))
              mma_f16(tmem + 128, a, b, idesc64, 1u);
            else
              mma_f16(tmem, a, b, idesc, 1u);
          }
        }
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int N>
void run(const char* label, int mode, long long* d_out, int sms) {
  const int n_mma = 36 * 200;
  const int smem = 1024 + 49152 + 3 * N * 128;
  CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<N><<<sms, 128, smem>>>(d_out, n_mma, mode);
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0, mx = 0;
  for (int i = 0; i < sms; ++i) {
    avg += h[i];
    if (h[i] > mx) mx = h[i];
  }
  avg /= sms;
  const double per = avg / n_mma;
  const double flops_per_mma = 2.0 * 128 * N * 16 * (mode == 3 ? 0.75 : 1.0);
  printf("%-44s N=%3d grid=%3d : %.1f cycles/MMA (max CTA %.1f)  -> %.0f FLOP/cycle/SM\n", label, N, sms, per,
         mx / n_mma, flops_per_mma / per);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  long long* d_out;
  CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  printf("== mma_rate_probe: %s, %d SMs, nominal dense fp16 = 8192 FLOP/cycle/SM ==\n", prop.name, sms);
  for (int grid : {1, sms}) {
    run<64>("standard A tiles", 0, d_out, grid);
    run<64>("halo-window A (SBO 1280, shifted)", 1, d_out, grid);
    run<64>("same A and B every MMA", 2, d_out, grid);
    run<128>("standard A tiles", 0, d_out, grid);
    run<128>("halo-window A (SBO 1280, shifted)", 1, d_out, grid);
    run<128>("alternating N=128 / N=64 (x3 mix)", 3, d_out, grid);
    run<256>("standard A tiles", 0, d_out, grid);
    run<256>("halo-window A (SBO 1280, shifted)", 1, d_out, grid);
    run<256>("same A and B every MMA", 2, d_out, grid);
  }
  printf("== done ==\n");
  return 0;
}
