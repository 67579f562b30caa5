// Internal launcher interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace pfnl {

// Optional per-launch timing (pfnl_profile): CUDA events recorded on the launching stream around
// each kernel class; used by bench.py to measure the dominant kernel's duration live.
enum ProfKind {
  kProfPack = 0,
  kProfNonlocal = 1,
  kProfConv0 = 2,
  kProfConv1 = 3,   // 3x3 64->64
  kProfConv10 = 4,  // 1x1 448->64
  kProfConv2 = 5,   // 3x3 128->64 (+ residual)
  kProfMerge1 = 6,
  kProfTail = 7,
  kProfOther = 8,
  kProfPfrbFlow = 9,  // the whole PFRB stack as one persistent dataflow kernel (pfrb_flow.cu)
  kProfNlKernel = 10,  // nl_tc_kernel alone (stage entry pfnl_nonlocal only; inside kProfNonlocal in the forward)
  kProfKinds = 11
};
struct Profiler {
  bool on = false;
  struct Rec {
    int kind;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void begin(int kind, cudaStream_t s) {
    if (!on) return;
    Rec r;
    r.kind = kind;
    r.a = get();
    r.b = get();
    cudaEventRecord(r.a, s);
    recs.push_back(r);
  }
  void end(cudaStream_t s) {
    if (!on) return;
    cudaEventRecord(recs.back().b, s);
  }
};

// ---- reorder.cu -------------------------------------------------------------------------
int launch_pack_tokens(const float* lr, int N, int H, int W, float* tokens, cudaStream_t s);
int launch_depth_to_space(const float* in, int N, int H, int W, int C, int b, float* out, cudaStream_t s);
int launch_space_to_depth(const float* in, int N, int H, int W, int C, int b, float* out, cudaStream_t s);
int launch_bicubic4(const float* in, int N, int H, int W, int C, float* out, cudaStream_t s);

// ---- conv_ffma.cu -----------------------------------------------------------------------
// One K-slice of the implicit GEMM: `slice_ch` input channels of output image m live at
//   ptr + (m / img_div) * img_stride + pixel * pix_stride + c
// so channel concats (model/pfnl.py:67,69,73) are never materialised.
struct ConvSlice {
  const float* ptr;
  long long img_stride;
  int img_div;
  int pix_stride;
};
struct ConvArgs {
  ConvSlice slice[7];
  int nslices;
  int slice_ch;  // channels per slice, multiple of 16
  int images;    // output images (N or N*7)
  int H, W;
  const float* wpack;  // packed by pack_conv_ffma_weights
  const float* bias;   // [cout]
  int cout;            // real output channels (<=64, multiple of 4)
  int act;             // 1: leaky_relu(0.2)
  const float* residual;  // optional [images,H,W,cout], added after the activation
  float* out;             // [images,H,W,cout]
};
int init_conv_ffma();
int launch_conv_ffma(int ks, const ConvArgs& a, cudaStream_t s);
// HWIO [ks,ks,cin,cout] -> [cin/16][ks*ks][16][64] (cout zero-padded to 64)
size_t conv_ffma_packed_floats(int ks, int cin);
int launch_pack_conv_ffma_weights(const float* hwio_dev, int ks, int cin, int cout, float* packed_dev,
                                  cudaStream_t s);

// Generic direct conv (any Cin/Cout, k in {1,3,5}); HWIO weights on the device.
int launch_conv_direct(const float* in, int N, int H, int W, int Cin, const float* kernel, const float* bias,
                       int k, int Cout, int act, const float* residual, float* out, cudaStream_t s);
// conv0 (model/pfnl.py:48,61-62): inp21 [N,H,W,21] (frame t = channels 3t..3t+2) -> [N*7,H,W,64], 5x5, LReLU.
int launch_conv0(const float* inp21, int N, int H, int W, const float* w_hwio, const float* bias, float* out,
                 cudaStream_t s);
// Upscaler tail (model/pfnl.py:76-80): merge [N,H,W,48] -> d2s -> convmerge2 -> d2s, + bicubic(lr[:,3]).
int launch_tail(const float* merge, const float* lr, int N, int H, int W, const float* w2_hwio, const float* b2,
                float* sr, cudaStream_t s);

// ---- nonlocal_ffma.cu -------------------------------------------------------------------
int init_nonlocal_ffma();
int launch_nl_linear(const float* X, int rows, const float* Wm, const float* b, float* Y, cudaStream_t s);
// Z = Y*Ww+bw, then inp21[n,2h2+dy,2w2+dx,t*3+c] = lr[n,t,..] + Z[n,tok,(dy*2+dx)*21+t*3+c]  (pfnl.py:59-60)
int launch_nl_linear_scatter(const float* Y, const float* lr, int N, int H, int W, const float* Ww, const float* bw,
                             float* inp21, cudaStream_t s);
int launch_nl_flash_ffma(const float* X, const float* G, int N, int L, float* Y, cudaStream_t s);

// ---- video_io.cu ------------------------------------------------------------------------
int launch_downsample4(const float* hr, int F, int H, int W, const float* blur_dev, float* lr, cudaStream_t s);
int launch_gather_windows(const float* frames, int F, long long frame_elems, int first, int count, float* clips,
                          cudaStream_t s);
int launch_quantize_u8(const float* in, long long n, unsigned char* out, cudaStream_t s);

// ---- metrics.cu -------------------------------------------------------------------------
constexpr int kMetricChunks = 64;
int init_metrics();
size_t metric_partials(int F, int H, int W);  // doubles of partial-sum scratch the two metrics need
int launch_luma(const float* rgb, long long npix, float vmin, float vmax, int round_y, double* y, cudaStream_t s);
int launch_ysq(const double* ya, const double* yb, int F, int H, int W, int border, double* partial, double* out,
               cudaStream_t s);
int launch_ssim(const double* ya, const double* yb, int F, int H, int W, double* partial, double* out, cudaStream_t s);

// ---- mse.cu -----------------------------------------------------------------------------
constexpr int kMseChunks = 64;
int launch_mse(const float* sr, const float* hr, int N, long long per_clip, double* partial /*[N*kMseChunks]*/,
               float* mse, cudaStream_t s);

}  // namespace pfnl
