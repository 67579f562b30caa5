// L2 -> shared-memory fill rate of one SM (one CTA per SM): how fast can a CTA pull a block of L2-resident
// data into its shared memory, by path and by the number of CTAs doing it at the same time?
//   path 0: cp.async.bulk (TMA engine), 32 KB chunks          path 1: cp.async 16 B by 256 threads (LSU)
//   path 2: half by each                                      path 3: cp.async.bulk, 4 KB chunks
//   source: the SAME 144 KB for every CTA, or a private 144 KB per CTA
// Motivation: the weight swap of conv_tc.cu (147 KB per phase) takes ~5 K cycles whatever is tried.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probes/bin/fill_rate_probe probes/fill_rate_probe.cu
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

constexpr int kBytes = 144 * 1024;
constexpr int kReps = 24;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(256, 1) fill_kernel(const uint8_t* __restrict__ src, long long cta_stride, int path,
                                                      long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const uint8_t* mine = src + (long long)blockIdx.x * cta_stride;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  __syncthreads();
  long long t0 = 0;
  for (int r = -2; r < kReps; ++r) {  // two warm-up rounds bring the data into L2
    if (r == 0) {
      __syncthreads();
      t0 = clock64();
    }
    const int tma_bytes = path == 1 ? 0 : (path == 2 ? kBytes / 2 : kBytes);
    const int chunk = path == 3 ? 4096 : 32768;
    if (tid == 0 && tma_bytes > 0) {
      mbar_arrive_expect_tx(&bar, tma_bytes);
      for (int off = 0; off < tma_bytes; off += chunk) {
        const int n = (tma_bytes - off) < chunk ? (tma_bytes - off) : chunk;
        bulk_load(sm + off, mine + off, n, &bar);
      }
    }
    for (int off = tma_bytes + tid * 16; off < kBytes; off += 256 * 16) cp_async16(sm + off, mine + off);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    if (tma_bytes > 0) mbar_wait(&bar, (uint32_t)((r + 2) & 1));
    __syncthreads();
  }
  if (tid == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint8_t* src;
  long long* d_out;
  CK(cudaMalloc(&src, (size_t)kBytes * 160));
  CK(cudaMemset(src, 1, (size_t)kBytes * 160));
  CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  const int smem = kBytes + 2048;
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  printf("== fill_rate_probe: %s, %d SMs; %d KB per round, %d timed rounds ==\n", prop.name, sms, kBytes / 1024, kReps);
  const char* pname[4] = {"cp.async.bulk 32 KB chunks", "cp.async 16 B x 256 threads", "half bulk + half cp.async",
                          "cp.async.bulk 4 KB chunks"};
  for (int shared_src = 1; shared_src >= 0; --shared_src)
    for (int path = 0; path < 4; ++path)
      for (int grid : {1, 32, 128, sms}) {
        fill_kernel<<<grid, 256, smem>>>(src, shared_src ? 0 : kBytes, path, d_out);
        CK(cudaDeviceSynchronize());
        long long h[256];
        CK(cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += h[i];
        avg /= grid;
        const double cyc = avg / kReps;
        printf("%-28s %-12s grid=%3d : %7.0f cycles per %d KB  -> %5.1f B/clk/SM, %7.0f B/clk chip\n", pname[path],
               shared_src ? "same source" : "private src", grid, cyc, kBytes / 1024, kBytes / cyc,
               kBytes / cyc * grid);
      }
  printf("== done ==\n");
  return 0;
}
