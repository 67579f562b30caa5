// Placeholder for the tensor-core path while it is being brought up: every entry point fails
// loudly with PFNL_ERR_UNIMPLEMENTED (never a silent fallback).
#include "common.cuh"
#include "tc.h"

namespace pfnl {

void tc_carve(TcWorkspace&, int, int, int, int, const std::function<char*(size_t)>&) {}
int tc_init(TcWeights&, int precision, const TcRawWeights&, std::vector<void*>&) {
  set_error("precision %d (tensor-core path) is not built into this library", precision);
  return PFNL_ERR_UNIMPLEMENTED;
}
void tc_destroy(TcWeights&) {}
int tc_trunk(const TcWeights&, TcWorkspace&, int, const float*, int, int, int, float*, cudaStream_t, long long*,
             Profiler*) {
  set_error("tensor-core trunk not built");
  return PFNL_ERR_UNIMPLEMENTED;
}
int tc_nonlocal(const TcWeights&, TcWorkspace&, const float*, const float*, int, int, int, float*, cudaStream_t,
                long long*, Profiler*) {
  set_error("tensor-core non-local not built");
  return PFNL_ERR_UNIMPLEMENTED;
}
int tc_nonlocal_tokens(const TcWeights&, const float*, int, int, float*, cudaStream_t, long long*) {
  set_error("tensor-core non-local not built");
  return PFNL_ERR_UNIMPLEMENTED;
}
int tc_pfrb_fp32io(const TcWeights&, TcWorkspace&, int, int, const float*, int, int, int, float*, cudaStream_t,
                   long long*) {
  set_error("tensor-core PFRB not built");
  return PFNL_ERR_UNIMPLEMENTED;
}

}  // namespace pfnl
