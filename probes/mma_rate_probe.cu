// tcgen05.mma throughput probe (every SM busy): cycles per MMA instruction, kind::f16, K = 16, for
//   cta_group::1  M in {64, 128}   N in {32 .. 256}
//   cta_group::2  M in {128, 256}  N in {64, 128, 256}   (CTA pairs, one issuing thread per pair)
// and for the instruction mixes the conv kernel issues (N=128 hi pass + N=64 lo pass), with one or several
// accumulators (does a dependent accumulate stall the pipe?).
// Round 1 measured only cta_group::1 / M=128 / N in {64,128,256}: cycles = max(N/2, 40) + 21.6.  This version
// decides whether a 2-CTA (M=256) or an M=64 x N=256 (operands swapped) formulation can beat that law.
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

template <int CG>
__device__ __forceinline__ void mma_cg(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    mma_f16(d, a, b, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
  }
}

// mix 0: every MMA has shape (M, N), one accumulator
// mix 1: alternate (M, N) -> acc0 and (M, N/2) -> acc1   (the fp16x3 hi/lo pattern when N = 128)
// mix 2: shape (M, N), rotating over 4 accumulators (N <= 64 only: 4 x 64 columns)
template <int CG, int M, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int n_mma, int mix) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;          // 48 KB: 128 rows x 64 k (16 KB) per "tap", 3 taps
  uint8_t* b_sm = smem + 49152;  // 3 taps x N rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank = 0;
  if (CG == 2) rank = cluster_ctarank();
  for (int i = tid; i < (49152 + 3 * N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    if (CG == 1) {
      tmem_alloc(&tmem_slot, 256);
      tmem_relinquish();
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                   "r"(256u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_f16(M, N);
    constexpr uint32_t idesc_half = make_idesc_f16(M, N / 2);
    const uint64_t ad = make_sdesc_sw128(smem_u32(a_sm), 1024, 0);
    const uint64_t bd = make_sdesc_sw128(smem_u32(b_sm), 1024, 0);
    t0 = clock64();
    if (rank == 0) {
      if (elect_one()) {
        for (int i = 0; i < n_mma; i += 12) {
#pragma unroll
          for (int tp = 0; tp < 3; ++tp) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a = ad + ((tp * 16384 + k * 32) >> 4);
              const uint64_t b = bd + ((tp * N * 128 + k * 32) >> 4);
              if (mix == 1 && (k & 1))
                mma_cg<CG>(tmem + 128, a, b, idesc_half, 1u);
              else if (mix == 2)
                mma_cg<CG>(tmem + (k & 3) * 64, a, b, idesc, 1u);
              else
                mma_cg<CG>(tmem, a, b, idesc, 1u);
            }
          }
        }
        if (CG == 1)
          mma_commit(&bar);
        else
          asm volatile(
              "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                  smem_u32(&bar)),
              "h"((uint16_t)3)
              : "memory");
      }
      __syncwarp();
    }
    // bounded wait: a protocol mistake must end the probe, not hang the box
    bool ok = false;
    for (long long spin = 0; spin < (1ll << 20) && !ok; ++spin) ok = mbar_try_wait(&bar, 0);
    t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x] = ok ? (t1 - t0) : -1;
  }
  fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1)
      tmem_dealloc(tmem, 256);
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int CG, int M, int N>
void run(const char* label, int mix, long long* d_out, int sms) {
  const int n_mma = 12 * 600;
  const int smem = 1024 + 49152 + 3 * N * 128;
  CK(cudaFuncSetAttribute(rate_kernel<CG, M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, rate_kernel<CG, M, N>, d_out, n_mma, mix));
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, d_out, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0, mx = 0;
  int cnt = 0, bad = 0;
  for (int i = 0; i < sms; i += CG) {  // the issuing CTA of each pair
    if (h[i] < 0) {
      ++bad;
      continue;
    }
    avg += h[i];
    ++cnt;
    if (h[i] > mx) mx = h[i];
  }
  if (cnt == 0) {
    printf("%-34s cg=%d M=%3d N=%3d : TIMEOUT (%d CTAs)\n", label, CG, M, N, bad);
    return;
  }
  avg /= cnt;
  const double per = avg / n_mma;
  // FLOPs one SM executes per instruction: its M/CG rows x N columns x K=16 (mix 1: half of the instructions have N/2)
  const double flops_sm = 2.0 * (M / CG) * N * 16 * (mix == 1 ? 0.75 : 1.0);
  printf("%-34s cg=%d M=%3d N=%3d : %6.1f cycles/MMA (max %6.1f)  %5.0f FLOP/cycle/SM = %4.1f %% of 8192%s\n", label, CG,
         M, N, per, mx / n_mma, flops_sm / per, 100.0 * flops_sm / per / 8192.0, bad ? "  [some CTAs timed out]" : "");
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount & ~1;
  long long* d_out;
  CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  printf("== mma_rate_probe v2: %s, %d SMs, nominal dense fp16 = 8192 FLOP/cycle/SM ==\n", prop.name, sms);
  run<1, 128, 32>("one accumulator", 0, d_out, sms);
  run<1, 128, 48>("one accumulator", 0, d_out, sms);
  run<1, 128, 64>("one accumulator", 0, d_out, sms);
  run<1, 128, 64>("4 rotating accumulators", 2, d_out, sms);
  run<1, 128, 80>("one accumulator", 0, d_out, sms);
  run<1, 128, 96>("one accumulator", 0, d_out, sms);
  run<1, 128, 128>("one accumulator", 0, d_out, sms);
  run<1, 128, 128>("alternating N / N/2 (x3 mix)", 1, d_out, sms);
  run<1, 128, 160>("one accumulator", 0, d_out, sms);
  run<1, 128, 192>("one accumulator", 0, d_out, sms);
  run<1, 128, 256>("one accumulator", 0, d_out, sms);
  run<1, 64, 64>("one accumulator", 0, d_out, sms);
  run<1, 64, 128>("one accumulator", 0, d_out, sms);
  run<1, 64, 256>("one accumulator", 0, d_out, sms);
  run<2, 256, 64>("pair, one accumulator", 0, d_out, sms);
  run<2, 256, 128>("pair, one accumulator", 0, d_out, sms);
  run<2, 256, 128>("pair, alternating N / N/2", 1, d_out, sms);
  run<2, 256, 256>("pair, one accumulator", 0, d_out, sms);
  run<2, 128, 64>("pair, one accumulator", 0, d_out, sms);
  run<2, 128, 128>("pair, one accumulator", 0, d_out, sms);
  run<2, 128, 256>("pair, one accumulator", 0, d_out, sms);
  printf("== done ==\n");
  return 0;
}
