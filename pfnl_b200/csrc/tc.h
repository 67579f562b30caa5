// Interface of the tensor-core (tcgen05 / TMA) path: conv_tc.cu, nonlocal_tc.cu, tc_host.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdlib.h>

#include <functional>
#include <vector>

#include "../../include/pfnl_b200.h"
#include "kernels.h"

namespace pfnl {

// Shared-memory halo patch of the 3x3 tensor-core conv: row pitch in pixels (TMA box width) and
// the UMMA descriptor base-offset convention.  Settled by probes/umma_probe.cu on a B200 (see
// DESIGN.md): the swizzle is a function of the absolute smem address, so a pitch of 10 pixels
// (no padding) with base_offset = 0 is exact; base_offset = (start>>7)&7 is WRONG on B200.
constexpr float kTcTruncCompDefault = 0.18f;  // = the model's value; calibration: profiles/r2f_trunc_comp.txt
constexpr int kTcPatchW3 = 10;
constexpr int kTcBaseOffsetMode = 0;

// precision mode -> number of fp16 planes of the conv path / whether the non-local block runs on tcgen05
inline int tc_nsplit(int precision) {
  return (precision == PFNL_PREC_TC_FP16X3 || precision == PFNL_PREC_TC_FP16X3_NLTC) ? 2 : 1;
}
// FP16X3 runs the non-local block on tcgen05 with hi/lo-split operands (PFNL_NL_FFMA=1 in the environment keeps
// the fp32 CUDA-core kernel for A/B runs); FP16 and FP16X3_NLTC use fp16 operands.
inline bool tc_nl_on_tensor_cores(int precision) {
  static const bool ffma = getenv("PFNL_NL_FFMA") != nullptr;
  if (precision == PFNL_PREC_TC_FP16X3) return !ffma;
  return precision == PFNL_PREC_TC_FP16 || precision == PFNL_PREC_TC_FP16X3_NLTC;
}

struct TcRawWeights {  // fp32 HWIO device pointers owned by the handle
  const float *nl_g_w, *nl_g_b, *nl_w_w, *nl_w_b;
  const float *nl_gw_w, *nl_gw_b;  // folded output linear of the non-local block: Wg*Ww, bg*Ww+bw
  const float *conv0_w, *conv0_b;
  const float* conv1_w[PFNL_NUM_BLOCK];
  const float* conv1_b[PFNL_NUM_BLOCK];
  const float* conv10_w[PFNL_NUM_BLOCK];
  const float* conv10_b[PFNL_NUM_BLOCK];
  const float* conv2_w[PFNL_NUM_BLOCK];
  const float* conv2_b[PFNL_NUM_BLOCK];
  const float *merge1_w, *merge1_b;
};

struct TcWeights {
  int precision = 0;
  int nsplit = 1;  // 1: fp16, 2: hi/lo fp16 planes
  TcRawWeights raw{};
  // UMMA-ready fp16 weight images (K-major, 128B swizzle), one per layer:
  //   conv1_i  [nsplit][9 taps][64 co][64 ci]
  //   conv10_i [nsplit][7 slices][64 co][64 ci]
  //   conv2b_i [nsplit][9][64][64]   (base half of conv2, Cin 0-63)
  //   conv2f_i [nsplit][9][64][64]   (frame half of conv2, Cin 64-127)
  //   merge1   [nsplit][7*9][64 co (48 real)][64 ci]
  void* conv1[PFNL_NUM_BLOCK] = {};
  void* conv10[PFNL_NUM_BLOCK] = {};
  void* conv2b[PFNL_NUM_BLOCK] = {};
  void* conv2f[PFNL_NUM_BLOCK] = {};
  void* merge1 = nullptr;
  void* conv0 = nullptr;  // [10 k-chunks][nsplit*64 rows][8 k] fp16, un-swizzled core matrices (conv0_tc_kernel)
  float* zero_bias = nullptr;  // [64] zeros
  bool flow = true;            // PFRB stack as the persistent dataflow kernel (pfnl_set_flow / PFNL_TC_FLOW=0: phase kernels)
  bool pdl = true;             // programmatic dependent launch between the tensor-core kernels (off while profiling)
  // grow-only scratch of the stage-level non-local entry (pfnl_nonlocal): operands, partials, Y
  mutable unsigned char* nl_scratch = nullptr;
  mutable size_t nl_scratch_cap = 0;
  float trunc_comp = 0.f;      // kappa of the TMEM truncation-bias compensation (conv_tc_dev.cuh); PFNL_TC_TRUNC_COMP
  int num_sms = 0;             // SM count of the handle's device (grid size of the persistent kernels)
  // non-local (precision 2): fp16 Wg^T image etc.
  void* nl_priv = nullptr;
};

struct TcWorkspace {
  // activations as fp16 NHWC planes [images,H,W,64]; plane 0 = hi, plane 1 = lo (x3 mode)
  void* actA[2];   // inp0 (residual stream)
  void* actB[2];   // inp1
  void* base[2];   // conv10 output [N,H,W,64]
  float* pbase;    // fp32 partial conv2 over the base half [N,H,W,64]
  void* nl_x16;    // fp16 token matrix [N*L,96]
  void* nl_priv;
  int* flow_flags;  // dependency counters of pfrb_flow.cu (handle-owned, zero between launches)
  int* flow_fault;  // host-mapped {1+kind, cta, a, b} of a wait that timed out (survives the trap)
};

void tc_carve(TcWorkspace& w, int precision, int N, int H, int W, const std::function<char*(size_t)>& take);
int tc_init(TcWeights& tw, int precision, const TcRawWeights& raw, std::vector<void*>& allocs);
void tc_destroy(TcWeights& tw);

// conv0 .. convmerge1 (model/pfnl.py:61-74): inp21 [N,H,W,21] fp32 -> merge [N,H,W,48] fp32
int tc_trunk(const TcWeights& tw, TcWorkspace& w, int precision, const float* inp21, int N, int H, int W,
             float* merge, cudaStream_t s, long long* launches, Profiler* prof);
bool tc_has_nonlocal();
int tc_nl_init();                               // per-device kernel attributes
size_t tc_nl_workspace_bytes(int N, int L);     // X16 + Gt16 staging of the tensor-core non-local block
// tokens [N,L,84] (+ lr for the residual) -> inp21 [N,H,W,21]   (model/pfnl.py:58-60)
int tc_nonlocal(const TcWeights& tw, TcWorkspace& w, const float* tokens, const float* lr, int N, int H, int W,
                float* inp21, cudaStream_t s, long long* launches, Profiler* prof);
// tokens [N,L,84] -> NonLocalBlock output [N,L,84]
int tc_nonlocal_tokens(const TcWeights& tw, const float* tokens, int N, int L, float* out, cudaStream_t s,
                       long long* launches, Profiler* prof);
// The PFRB stack as one persistent dataflow kernel (pfrb_flow.cu): blocks [blk0, blk0+nblk) on the fp16 planes;
// in place on actA.
int* tc_fault_buffer();  // api.cu: host-mapped wait-timeout record (device pointer; may be NULL)
int tc_flow_init();
bool tc_flow_default();                          // false with PFNL_TC_FLOW=0 in the environment
void tc_flow_split(int num_sms, int n_units, int out[4]);  // CTAs per role: conv1, conv10, conv2b, conv2f
size_t tc_flow_flag_ints(int N, int H, int W);   // ints of dependency counters the launch needs (zeroed once)
int tc_pfrb_flow(const TcWeights& tw, TcWorkspace& w, int blk0, int nblk, int N, int H, int W, bool pdl,
                 cudaStream_t s);
// One PFRB with fp32 frames in/out (conversion kernels around the tensor-core block).
int tc_pfrb_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, int blk, const float* frames, int N, int H,
                   int W, float* frames_out, cudaStream_t s, long long* launches);

// conv0 / convmerge1 alone, fp32 tensors either side (stage-level parity of the tcgen05 kernels)
int tc_conv0_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, const float* inp21, int N, int H, int W,
                    float* frames_out, cudaStream_t s, long long* launches);
int tc_merge1_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, const float* frames, int N, int H, int W,
                     float* merge, cudaStream_t s, long long* launches);

}  // namespace pfnl
