// Host-side TMA tensor-map construction without linking libcuda: cuTensorMapEncodeTiled is
// fetched through cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfnl {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// fp16 activation plane, channel-chunk-major: [images][8 chunks][H][W][8 ch] - a pixel's 8 channels of one
// chunk are 16 contiguous bytes and the pixels of an image row are adjacent (what the epilogue warps
// write, see conv_tc.cu).  4-D map (w*8+c, h, chunk, img), box (box_w*8, box_h, 8, 1), no swizzle: the
// box lands in shared memory as 8 sub-patches (one per chunk) of [box_h][box_w] 16-byte pixels, in
// which 8 consecutive pixels of a row are exactly one un-swizzled UMMA core matrix (8 rows x 16 B).
// Out-of-bounds elements are zero-filled ('same' padding).
// Returns 0 on success, the CUresult (or -1 if the entry point is missing) otherwise.
inline int make_act_tmap(CUtensorMap* m, const void* base, int images, int H, int W, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, 8, (cuuint64_t)images};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)H * W * 128};
  cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

}  // namespace tc
}  // namespace pfnl
