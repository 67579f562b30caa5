// tcgen05 non-local block (PFNL_PREC_TC_FP16) - under construction: until it lands the fp16
// precision uses the fp32 FFMA non-local kernels (nonlocal_ffma.cu), never a CPU path.
#include "common.cuh"
#include "tc.h"

namespace pfnl {

bool tc_has_nonlocal() { return false; }

int tc_nonlocal(const TcWeights&, TcWorkspace&, const float*, const float*, int, int, int, float*, cudaStream_t,
                long long*, Profiler*) {
  set_error("tensor-core non-local kernel not built");
  return PFNL_ERR_UNIMPLEMENTED;
}
int tc_nonlocal_tokens(const TcWeights&, const float*, int, int, float*, cudaStream_t, long long*) {
  set_error("tensor-core non-local kernel not built");
  return PFNL_ERR_UNIMPLEMENTED;
}

}  // namespace pfnl
