// Shared helpers for libpfnl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pfnl_b200.h"

namespace pfnl {

constexpr int kFrames = PFNL_NUM_FRAMES;
constexpr int kMF = PFNL_MF;
constexpr int kNL = PFNL_NL_CH;
constexpr float kLReLU = 0.2f;  // tf.nn.leaky_relu default alpha (model/pfnl.py:42)

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define PFNL_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::pfnl::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define PFNL_LAUNCH_CHECK() PFNL_CUDA(cudaGetLastError())

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float lrelu(float v) { return fmaxf(v * kLReLU, v); }

// 16-byte async copy global->shared; src_bytes==0 zero-fills (used for 'same' zero padding).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

}  // namespace pfnl
