// C ABI of libpfnl_b200.so (include/pfnl_b200.h): handle, workspace, the PFNL.forward launch
// sequence (model/pfnl.py:39-80) and the stage-level entry points.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tc.h"

namespace pfnl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Host-mapped record of a bounded device-side wait that gave up (tc_ptx.cuh wait_timeout_trap): it survives the
// trap that follows, so the error message can say which CTA waited for what.  Process-wide, portable memory.
static int* g_fault_host = nullptr;
int* tc_fault_buffer() {
  static int* dev = nullptr;
  if (!g_fault_host) {
    // 8 ints of fault record + 256 CTAs x 8 ints of progress marks (PFNL_FLOW_DEBUG=1, pfrb_flow.cu)
    if (cudaHostAlloc((void**)&g_fault_host, (8 + 256 * 8) * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) !=
        cudaSuccess) {
      cudaGetLastError();
      g_fault_host = nullptr;
      return nullptr;
    }
    memset(g_fault_host, 0, (8 + 256 * 8) * sizeof(int));
    if (cudaHostGetDevicePointer((void**)&dev, g_fault_host, 0) != cudaSuccess) {
      cudaGetLastError();
      dev = nullptr;
    }
  }
  return dev;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  if (g_fault_host && g_fault_host[0] != 0)
    set_error("CUDA error %d (%s) at %s:%d: %s [device wait timed out: kind %d, CTA %d, detail %d / %d]", (int)e,
              cudaGetErrorString(e), file, line, what, g_fault_host[0] - 1, g_fault_host[1], g_fault_host[2],
              g_fault_host[3]);
  else
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return PFNL_ERR_CUDA;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  // fp32 path
  float* tokens;   // [N,L,84]
  float* g;        // [N,L,84]
  float* yatt;     // [N,L,84]
  float* inp21;    // [N,H,W,21]
  float* framesA;  // [N*7,H,W,64]  inp0 (residual stream)
  float* framesB;  // [N*7,H,W,64]  inp1
  float* base;     // [N,H,W,64]
  float* merge;    // [N,H,W,48]
  TcWorkspace tc;  // tensor-core path buffers (precision 1,2)
  size_t bytes;
};

// Lays the workspace out behind `base` (may be NULL to only size it).
static Workspace carve(char* basep, int precision, int N, int H, int W) {
  Workspace w;
  memset(&w, 0, sizeof(w));
  size_t off = 0;
  auto take = [&](size_t bytes) -> char* {
    char* p = basep ? basep + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  const size_t L = (size_t)(H / 2) * (W / 2);
  const size_t hw = (size_t)H * W;
  w.tokens = (float*)take((size_t)N * L * kNL * 4);
  w.g = (float*)take((size_t)N * L * kNL * 4);
  w.yatt = (float*)take((size_t)N * L * kNL * 4);
  w.inp21 = (float*)take((size_t)N * hw * 21 * 4);
  w.framesA = (float*)take((size_t)N * kFrames * hw * kMF * 4);
  w.framesB = (float*)take((size_t)N * kFrames * hw * kMF * 4);
  w.base = (float*)take((size_t)N * hw * kMF * 4);
  w.merge = (float*)take((size_t)N * hw * 48 * 4);
  if (precision != PFNL_PREC_FP32) tc_carve(w.tc, precision, N, H, W, take);
  w.bytes = off;
  return w;
}

}  // namespace pfnl

using namespace pfnl;

struct pfnl_handle {
  int device = -1;
  int precision = 0;
  // raw HWIO weights on the device (fp32)
  float *nl_g_w = nullptr, *nl_g_b = nullptr, *nl_w_w = nullptr, *nl_w_b = nullptr;
  // g and w folded: w(softmax(S)*(X*Wg+bg)) = softmax(S)*X*(Wg*Ww) + (bg*Ww+bw) because softmax rows sum
  // to 1 and nothing non-linear sits between the two 1x1 convs (utils.py:26,64,67)
  float *nl_gw_w = nullptr, *nl_gw_b = nullptr;
  float *conv0_w = nullptr, *conv0_b = nullptr;
  float* conv1_w[PFNL_NUM_BLOCK] = {};
  float* conv1_b[PFNL_NUM_BLOCK] = {};
  float* conv10_w[PFNL_NUM_BLOCK] = {};
  float* conv10_b[PFNL_NUM_BLOCK] = {};
  float* conv2_w[PFNL_NUM_BLOCK] = {};
  float* conv2_b[PFNL_NUM_BLOCK] = {};
  float *merge1_w = nullptr, *merge1_b = nullptr, *merge2_w = nullptr, *merge2_b = nullptr;
  // FFMA-packed kernels
  float* conv1_p[PFNL_NUM_BLOCK] = {};
  float* conv10_p[PFNL_NUM_BLOCK] = {};
  float* conv2_p[PFNL_NUM_BLOCK] = {};
  float* merge1_p = nullptr;
  TcWeights tcw;  // tensor-core packed kernels (precision 1,2)
  std::vector<void*> allocs;
  // workspace
  char* ws = nullptr;
  size_t ws_cap = 0;
  int* flow_flags = nullptr;  // dependency counters of the PFRB dataflow kernel (zero between launches)
  size_t flow_cap = 0;
  double* mse_partial = nullptr;
  int mse_cap = 0;
  double* metric_buf = nullptr;  // [2][F*H*W] luma planes + partial sums of pfnl_msy / pfnl_ssim_y
  size_t metric_cap = 0;
  // scratch for pfnl_conv2d_nhwc's on-the-fly weight packing
  float* pack_scratch = nullptr;
  size_t pack_cap = 0;
  float* blur_dev = nullptr;  // 13x13 taps of pfnl_downsample4
  // host staging of pfnl_forward_host[_submit]: two slots so that the H2D copy of call k+1 and the D2H copy of
  // call k-1 run (on their own streams) under the forward of call k
  struct HostSlot {
    float *pin_in = nullptr, *pin_out = nullptr, *dev_in = nullptr, *dev_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    cudaEvent_t in_ready = nullptr, fwd_done = nullptr, out_ready = nullptr;
    bool busy = false;        // submitted, not yet waited for
    float* user_out = nullptr;  // pageable destination to fill from pin_out at wait time (NULL: copied directly)
    size_t out_bytes = 0;
  } slot[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  int next_slot = 0;
  long long launches = 0;
  Profiler prof;
  bool graphs = false;
  // CUDA-graph cache of the forward launch sequence, keyed by shape and buffer addresses; a miss for a known
  // shape recycles the least recently used executable of that shape with cudaGraphExecUpdate (same topology,
  // new pointers) instead of instantiating again
  struct GraphEntry {
    int N, H, W;
    const void* lr;
    void* sr;
    cudaGraphExec_t exec;
    long long nodes;
    unsigned long long stamp;
  };
  std::vector<GraphEntry> graph_cache;
  unsigned long long graph_clock = 0;
  long long graph_instantiations = 0, graph_updates = 0;
  cudaStream_t graph_stream = nullptr;  // stands in for the legacy NULL stream, which cannot be captured
  cudaEvent_t graph_ev_in = nullptr, graph_ev_out = nullptr;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

int upload(pfnl_handle* h, const float* host, size_t n, float** out) {
  if (!host) {
    set_error("pfnl_create: a weight pointer is NULL");
    return PFNL_ERR_BAD_ARG;
  }
  float* d = nullptr;
  PFNL_CUDA(cudaMalloc(&d, n * sizeof(float)));
  h->allocs.push_back(d);
  PFNL_CUDA(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
  *out = d;
  return PFNL_OK;
}

int dev_alloc(pfnl_handle* h, size_t bytes, void** out) {
  void* d = nullptr;
  PFNL_CUDA(cudaMalloc(&d, bytes));
  h->allocs.push_back(d);
  *out = d;
  return PFNL_OK;
}

int check_shape(int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) {
    set_error("bad shape N=%d H=%d W=%d (must be positive)", N, H, W);
    return PFNL_ERR_BAD_SHAPE;
  }
  if ((H & 1) || (W & 1)) {
    set_error("bad shape H=%d W=%d: tf.space_to_depth(.,2) needs even H and W (model/pfnl.py:57)", H, W);
    return PFNL_ERR_BAD_SHAPE;
  }
  return PFNL_OK;
}

// carve + the handle-owned pieces of the tensor-core workspace
Workspace carve_h(pfnl_handle* h, int N, int H, int W) {
  Workspace w = carve(h->ws, h->precision, N, H, W);
  w.tc.flow_flags = h->flow_flags;
  w.tc.flow_fault = tc_fault_buffer();
  return w;
}

int ensure_workspace(pfnl_handle* h, int N, int H, int W) {
  if (h->precision != PFNL_PREC_FP32) {
    const size_t ints = tc_flow_flag_ints(N, H, W);
    if (ints > h->flow_cap) {
      if (h->flow_flags) {
        PFNL_CUDA(cudaDeviceSynchronize());
        PFNL_CUDA(cudaFree(h->flow_flags));
        h->flow_flags = nullptr;
        h->flow_cap = 0;
      }
      PFNL_CUDA(cudaMalloc((void**)&h->flow_flags, ints * sizeof(int)));
      PFNL_CUDA(cudaMemset(h->flow_flags, 0, ints * sizeof(int)));
      h->flow_cap = ints;
    }
  }
  const size_t need = carve(nullptr, h->precision, N, H, W).bytes;
  if (need <= h->ws_cap) return PFNL_OK;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (h->ws) {
    PFNL_CUDA(cudaDeviceSynchronize());
    for (auto& e : h->graph_cache) cudaGraphExecDestroy(e.exec);
    h->graph_cache.clear();
    PFNL_CUDA(cudaFree(h->ws));
    h->ws = nullptr;
    h->ws_cap = 0;
  }
  (void)st;
  cudaError_t e = cudaMalloc((void**)&h->ws, need);
  if (e != cudaSuccess) {
    set_error("workspace allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
    cudaGetLastError();
    return PFNL_ERR_NO_MEMORY;
  }
  h->ws_cap = need;
  return PFNL_OK;
}

ConvSlice make_slice(const float* p, long long img_stride, int img_div, int pix_stride) {
  ConvSlice s;
  s.ptr = p;
  s.img_stride = img_stride;
  s.img_div = img_div;
  s.pix_stride = pix_stride;
  return s;
}

// One Progressive Fusion Residual Block on the fp32 path (model/pfnl.py:66-71).
int pfrb_fp32(pfnl_handle* h, int i, const float* in, float* out, Workspace& w, int N, int H, int W,
              cudaStream_t s) {
  const long long hw = (long long)H * W;
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.H = H;
  a.W = W;
  a.act = 1;
  a.cout = kMF;
  // inp1[t] = conv1_i(inp0[t])                              pfnl.py:66
  a.nslices = 1;
  a.slice_ch = kMF;
  a.slice[0] = make_slice(in, hw * kMF, 1, kMF);
  a.images = N * kFrames;
  a.wpack = h->conv1_p[i];
  a.bias = h->conv1_b[i];
  a.residual = nullptr;
  a.out = w.framesB;
  h->prof.begin(kProfConv1, s);
  int rc = launch_conv_ffma(3, a, s);
  h->prof.end(s);
  if (rc) return rc;
  // base = conv10_i(concat_t inp1[t])                       pfnl.py:67-68
  a.nslices = kFrames;
  for (int t = 0; t < kFrames; ++t) a.slice[t] = make_slice(w.framesB + t * hw * kMF, kFrames * hw * kMF, 1, kMF);
  a.images = N;
  a.wpack = h->conv10_p[i];
  a.bias = h->conv10_b[i];
  a.out = w.base;
  h->prof.begin(kProfConv10, s);
  rc = launch_conv_ffma(1, a, s);
  h->prof.end(s);
  if (rc) return rc;
  // inp0[t] += conv2_i(concat[base, inp1[t]])               pfnl.py:69-71
  a.nslices = 2;
  a.slice[0] = make_slice(w.base, hw * kMF, kFrames, kMF);
  a.slice[1] = make_slice(w.framesB, hw * kMF, 1, kMF);
  a.images = N * kFrames;
  a.wpack = h->conv2_p[i];
  a.bias = h->conv2_b[i];
  a.residual = in;
  a.out = out;
  h->prof.begin(kProfConv2, s);
  rc = launch_conv_ffma(3, a, s);
  h->prof.end(s);
  if (rc) return rc;
  h->launches += 3;
  return PFNL_OK;
}

int forward_launches(pfnl_handle* h, const float* lr, int N, int H, int W, float* sr, cudaStream_t s) {
  Workspace w = carve_h(h, N, H, W);
  const int L = (H / 2) * (W / 2);
  const long long hw = (long long)H * W;
  int rc;
  const bool nl_tc = tc_nl_on_tensor_cores(h->precision) && tc_has_nonlocal();
  if (!nl_tc) {
    // tokens = space_to_depth(concat frames)                pfnl.py:55-57
    // (the tensor-core non-local path gathers them inside its operand-preparation kernel instead)
    h->prof.begin(kProfPack, s);
    rc = launch_pack_tokens(lr, N, H, W, w.tokens, s);
    h->prof.end(s);
    if (rc) return rc;
    h->launches += 1;
  }
  if (nl_tc) {
    if ((rc = tc_nonlocal(h->tcw, w.tc, w.tokens, lr, N, H, W, w.inp21, s, &h->launches, &h->prof))) return rc;
  } else {
    h->prof.begin(kProfNonlocal, s);
    // NonLocalBlock                                         pfnl.py:58, utils.py:18-71
    // Q = K = V = X (theta = phi = input, utils.py:33-34,41-42; g folded into the output linear)
    if ((rc = launch_nl_flash_ffma(w.tokens, w.tokens, N, L, w.yatt, s))) return rc;
    // inp0 += depth_to_space(w(y))                          pfnl.py:59-60
    if ((rc = launch_nl_linear_scatter(w.yatt, lr, N, H, W, h->nl_gw_w, h->nl_gw_b, w.inp21, s))) return rc;
    h->prof.end(s);
    h->launches += 2;
  }
  if (h->precision == PFNL_PREC_FP32) {
    // conv0 on each frame                                   pfnl.py:61-62
    h->prof.begin(kProfConv0, s);
    rc = launch_conv0(w.inp21, N, H, W, h->conv0_w, h->conv0_b, w.framesA, s);
    h->prof.end(s);
    if (rc) return rc;
    h->launches += 1;
    for (int i = 0; i < PFNL_NUM_BLOCK; ++i)
      if ((rc = pfrb_fp32(h, i, w.framesA, w.framesA, w, N, H, W, s))) return rc;
    // merge = convmerge1(concat_t inp0[t])                  pfnl.py:73-74
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.H = H;
    a.W = W;
    a.act = 1;
    a.cout = 48;
    a.nslices = kFrames;
    a.slice_ch = kMF;
    for (int t = 0; t < kFrames; ++t) a.slice[t] = make_slice(w.framesA + t * hw * kMF, kFrames * hw * kMF, 1, kMF);
    a.images = N;
    a.wpack = h->merge1_p;
    a.bias = h->merge1_b;
    a.out = w.merge;
    h->prof.begin(kProfMerge1, s);
    rc = launch_conv_ffma(3, a, s);
    h->prof.end(s);
    if (rc) return rc;
    h->launches += 1;
  } else {
    if ((rc = tc_trunk(h->tcw, w.tc, h->precision, w.inp21, N, H, W, w.merge, s, &h->launches, &h->prof))) return rc;
  }
  // depth_to_space -> convmerge2 -> depth_to_space, + bicubic skip    pfnl.py:63,76-80
  h->prof.begin(kProfTail, s);
  rc = launch_tail(w.merge, lr, N, H, W, h->merge2_w, h->merge2_b, sr, s);
  h->prof.end(s);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

}  // namespace

extern "C" {

int pfnl_version(void) { return PFNL_VERSION; }

const char* pfnl_last_error(void) { return g_err; }

int pfnl_device_supported(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
  return p.major == 10 ? 1 : 0;
}

int pfnl_create(pfnl_handle** out, int device, const pfnl_weights* wts, int precision) {
  if (!out || !wts) {
    set_error("pfnl_create: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  *out = nullptr;
  if (precision < PFNL_PREC_FP32 || precision > PFNL_PREC_TC_FP16X3_NLTC) {
    set_error("pfnl_create: unknown precision %d", precision);
    return PFNL_ERR_BAD_ARG;
  }
  int sup = pfnl_device_supported(device);
  if (sup < 0) return sup;
  if (sup == 0) {
    set_error("pfnl_create: device %d is not compute capability 10.x (sm_100a only, no fallback)", device);
    return PFNL_ERR_UNSUPPORTED_ARCH;
  }
  DeviceGuard guard(device);
  if (!guard.ok) return cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
  pfnl_handle* h = new pfnl_handle();
  h->device = device;
  h->precision = precision;
  int rc = PFNL_OK;
#define TRY(x)              \
  do {                      \
    if ((rc = (x))) goto fail; \
  } while (0)
  tc_fault_buffer();  // allocate the host-mapped fault record now: not allowed later inside a stream capture
  TRY(init_conv_ffma());
  TRY(init_nonlocal_ffma());
  TRY(init_metrics());
  TRY(upload(h, wts->nl_g_kernel, kNL * kNL, &h->nl_g_w));
  TRY(upload(h, wts->nl_g_bias, kNL, &h->nl_g_b));
  TRY(upload(h, wts->nl_w_kernel, kNL * kNL, &h->nl_w_w));
  TRY(upload(h, wts->nl_w_bias, kNL, &h->nl_w_b));
  {
    // folded non-local output linear, computed in double on the host
    if (!wts->nl_g_kernel || !wts->nl_g_bias || !wts->nl_w_kernel || !wts->nl_w_bias) {
      set_error("pfnl_create: a weight pointer is NULL");
      rc = PFNL_ERR_BAD_ARG;
      goto fail;
    }
    std::vector<float> gw(kNL * kNL), gb(kNL);
    for (int i = 0; i < kNL; ++i)
      for (int j = 0; j < kNL; ++j) {
        double a = 0.0;
        for (int k = 0; k < kNL; ++k) a += (double)wts->nl_g_kernel[i * kNL + k] * (double)wts->nl_w_kernel[k * kNL + j];
        gw[i * kNL + j] = (float)a;
      }
    for (int j = 0; j < kNL; ++j) {
      double a = (double)wts->nl_w_bias[j];
      for (int k = 0; k < kNL; ++k) a += (double)wts->nl_g_bias[k] * (double)wts->nl_w_kernel[k * kNL + j];
      gb[j] = (float)a;
    }
    TRY(upload(h, gw.data(), kNL * kNL, &h->nl_gw_w));
    TRY(upload(h, gb.data(), kNL, &h->nl_gw_b));
  }
  TRY(upload(h, wts->conv0_kernel, 75 * 64, &h->conv0_w));
  TRY(upload(h, wts->conv0_bias, 64, &h->conv0_b));
  for (int i = 0; i < PFNL_NUM_BLOCK; ++i) {
    TRY(upload(h, wts->conv1_kernel[i], 9 * 64 * 64, &h->conv1_w[i]));
    TRY(upload(h, wts->conv1_bias[i], 64, &h->conv1_b[i]));
    TRY(upload(h, wts->conv10_kernel[i], 448 * 64, &h->conv10_w[i]));
    TRY(upload(h, wts->conv10_bias[i], 64, &h->conv10_b[i]));
    TRY(upload(h, wts->conv2_kernel[i], 9 * 128 * 64, &h->conv2_w[i]));
    TRY(upload(h, wts->conv2_bias[i], 64, &h->conv2_b[i]));
  }
  TRY(upload(h, wts->merge1_kernel, 9 * 448 * 48, &h->merge1_w));
  TRY(upload(h, wts->merge1_bias, 48, &h->merge1_b));
  TRY(upload(h, wts->merge2_kernel, 9 * 12 * 12, &h->merge2_w));
  TRY(upload(h, wts->merge2_bias, 12, &h->merge2_b));
  // FFMA packing (used by precision 0 and by pfnl_pfrb/pfnl_conv2d_nhwc reference launches)
  for (int i = 0; i < PFNL_NUM_BLOCK; ++i) {
    TRY(dev_alloc(h, conv_ffma_packed_floats(3, 64) * 4, (void**)&h->conv1_p[i]));
    TRY(launch_pack_conv_ffma_weights(h->conv1_w[i], 3, 64, 64, h->conv1_p[i], 0));
    TRY(dev_alloc(h, conv_ffma_packed_floats(1, 448) * 4, (void**)&h->conv10_p[i]));
    TRY(launch_pack_conv_ffma_weights(h->conv10_w[i], 1, 448, 64, h->conv10_p[i], 0));
    TRY(dev_alloc(h, conv_ffma_packed_floats(3, 128) * 4, (void**)&h->conv2_p[i]));
    TRY(launch_pack_conv_ffma_weights(h->conv2_w[i], 3, 128, 64, h->conv2_p[i], 0));
  }
  TRY(dev_alloc(h, conv_ffma_packed_floats(3, 448) * 4, (void**)&h->merge1_p));
  TRY(launch_pack_conv_ffma_weights(h->merge1_w, 3, 448, 48, h->merge1_p, 0));
  if (precision != PFNL_PREC_FP32) {
    TcRawWeights raw;
    raw.nl_gw_w = h->nl_gw_w;
    raw.nl_gw_b = h->nl_gw_b;
    raw.nl_g_w = h->nl_g_w;
    raw.nl_g_b = h->nl_g_b;
    raw.nl_w_w = h->nl_w_w;
    raw.nl_w_b = h->nl_w_b;
    raw.conv0_w = h->conv0_w;
    raw.conv0_b = h->conv0_b;
    for (int i = 0; i < PFNL_NUM_BLOCK; ++i) {
      raw.conv1_w[i] = h->conv1_w[i];
      raw.conv1_b[i] = h->conv1_b[i];
      raw.conv10_w[i] = h->conv10_w[i];
      raw.conv10_b[i] = h->conv10_b[i];
      raw.conv2_w[i] = h->conv2_w[i];
      raw.conv2_b[i] = h->conv2_b[i];
    }
    raw.merge1_w = h->merge1_w;
    raw.merge1_b = h->merge1_b;
    TRY(tc_init(h->tcw, precision, raw, h->allocs));
  }
  {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "cudaDeviceSynchronize", __FILE__, __LINE__);
      goto fail;
    }
  }
#undef TRY
  *out = h;
  return PFNL_OK;
fail:
  pfnl_destroy(h);
  return rc;
}

int pfnl_destroy(pfnl_handle* h) {
  if (!h) return PFNL_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (auto& e : h->graph_cache) cudaGraphExecDestroy(e.exec);
  if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
  if (h->graph_ev_in) cudaEventDestroy(h->graph_ev_in);
  if (h->graph_ev_out) cudaEventDestroy(h->graph_ev_out);
  for (auto& sl : h->slot) {
    if (sl.pin_in) cudaFreeHost(sl.pin_in);
    if (sl.pin_out) cudaFreeHost(sl.pin_out);
    if (sl.dev_in) cudaFree(sl.dev_in);
    if (sl.dev_out) cudaFree(sl.dev_out);
    if (sl.in_ready) cudaEventDestroy(sl.in_ready);
    if (sl.fwd_done) cudaEventDestroy(sl.fwd_done);
    if (sl.out_ready) cudaEventDestroy(sl.out_ready);
  }
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  for (auto& r : h->prof.recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : h->prof.pool) cudaEventDestroy(e);
  tc_destroy(h->tcw);
  for (void* p : h->allocs) cudaFree(p);
  if (h->ws) cudaFree(h->ws);
  if (h->flow_flags) cudaFree(h->flow_flags);
  if (h->mse_partial) cudaFree(h->mse_partial);
  if (h->metric_buf) cudaFree(h->metric_buf);
  if (h->pack_scratch) cudaFree(h->pack_scratch);
  if (h->blur_dev) cudaFree(h->blur_dev);
  cudaGetLastError();
  delete h;
  return PFNL_OK;
}

size_t pfnl_workspace_bytes(int precision, int N, int H, int W) {
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  return carve(nullptr, precision, N, H, W).bytes;
}

int pfnl_reserve(pfnl_handle* h, int N, int H, int W) {
  if (!h) {
    set_error("pfnl_reserve: NULL handle");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  return ensure_workspace(h, N, H, W);
}

int pfnl_set_graphs(pfnl_handle* h, int enable) {
  if (!h) {
    set_error("pfnl_set_graphs: NULL handle");
    return PFNL_ERR_BAD_ARG;
  }
  h->graphs = enable != 0;
  return PFNL_OK;
}

int pfnl_set_flow(pfnl_handle* h, int enable) {
  if (!h) {
    set_error("pfnl_set_flow: NULL handle");
    return PFNL_ERR_BAD_ARG;
  }
  if ((enable != 0) != h->tcw.flow) {  // captured graphs hold the other launch sequence
    DeviceGuard guard(h->device);
    PFNL_CUDA(cudaDeviceSynchronize());
    for (auto& e : h->graph_cache) cudaGraphExecDestroy(e.exec);
    h->graph_cache.clear();
  }
  h->tcw.flow = enable != 0;
  return PFNL_OK;
}

int pfnl_debug_fault(int* out4) {
  if (!out4) return PFNL_ERR_BAD_ARG;
  for (int i = 0; i < 4; ++i) out4[i] = g_fault_host ? g_fault_host[i] : 0;
  return PFNL_OK;
}

int pfnl_debug_progress(int* out, int n) {
  if (!out || n < 0 || n > 256 * 8) return PFNL_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) out[i] = g_fault_host ? g_fault_host[8 + i] : 0;
  return PFNL_OK;
}

int pfnl_debug_flow_split(int num_sms, int n_units, int* out4) {
  if (!out4 || num_sms < 8 || n_units < 1) return PFNL_ERR_BAD_ARG;
  tc_flow_split(num_sms, n_units, out4);
  return PFNL_OK;
}

long long pfnl_launch_count(const pfnl_handle* h) { return h ? h->launches : 0; }

int pfnl_profile(pfnl_handle* h, int enable) {
  if (!h) {
    set_error("pfnl_profile: NULL handle");
    return PFNL_ERR_BAD_ARG;
  }
  h->prof.on = enable != 0;
  h->tcw.pdl = !h->prof.on;
  return PFNL_OK;
}

int pfnl_profile_read(pfnl_handle* h, double* ms_by_kind, long long* launches_by_kind) {
  if (!h || !ms_by_kind || !launches_by_kind) {
    set_error("pfnl_profile_read: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  DeviceGuard guard(h->device);
  PFNL_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < kProfKinds; ++k) {
    ms_by_kind[k] = 0.0;
    launches_by_kind[k] = 0;
  }
  for (auto& r : h->prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_by_kind[r.kind] += ms;
      launches_by_kind[r.kind] += 1;
    }
    h->prof.pool.push_back(r.a);
    h->prof.pool.push_back(r.b);
  }
  h->prof.recs.clear();
  cudaGetLastError();
  return PFNL_OK;
}

int pfnl_forward(pfnl_handle* h, const float* lr, int N, int H, int W, float* sr, void* stream) {
  if (!h || !lr || !sr) {
    set_error("pfnl_forward: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  PFNL_CUDA(cudaStreamIsCapturing(s, &cap));
  if (cap == cudaStreamCaptureStatusNone) {
    if ((rc = ensure_workspace(h, N, H, W))) return rc;
  } else if (carve(nullptr, h->precision, N, H, W).bytes > h->ws_cap ||
             (h->precision != PFNL_PREC_FP32 && tc_flow_flag_ints(N, H, W) > h->flow_cap)) {
    set_error("pfnl_forward: stream is capturing and the workspace for (%d,%d,%d) is not reserved", N, H, W);
    return PFNL_ERR_BAD_ARG;
  }
  if (!h->graphs || h->prof.on || cap != cudaStreamCaptureStatusNone) return forward_launches(h, lr, N, H, W, sr, s);

  // The legacy NULL stream cannot be captured: run the graph on a private stream ordered after / before the
  // caller's stream with two events.
  cudaStream_t gs = s;
  if (s == nullptr) {
    if (!h->graph_stream) {
      PFNL_CUDA(cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
      PFNL_CUDA(cudaEventCreateWithFlags(&h->graph_ev_in, cudaEventDisableTiming));
      PFNL_CUDA(cudaEventCreateWithFlags(&h->graph_ev_out, cudaEventDisableTiming));
    }
    gs = h->graph_stream;
  }
  pfnl_handle::GraphEntry* hit = nullptr;
  pfnl_handle::GraphEntry* recycle = nullptr;  // least recently used executable of the same shape
  int same_shape = 0;
  for (auto& e : h->graph_cache) {
    if (e.N != N || e.H != H || e.W != W) continue;
    ++same_shape;
    if (e.lr == (const void*)lr && e.sr == (void*)sr) hit = &e;
    if (!recycle || e.stamp < recycle->stamp) recycle = &e;
  }
  if (!hit) {
    cudaGraph_t graph = nullptr;
    const long long before = h->launches;
    PFNL_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    rc = forward_launches(h, lr, N, H, W, sr, gs);
    cudaError_t e = cudaStreamEndCapture(gs, &graph);
    const long long nodes = h->launches - before;
    h->launches = before;
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    // callers that alternate between a few buffers (the two slots of pfnl_forward_host, ping-pong outputs) keep
    // one executable per address pair; beyond 4 per shape the oldest one is re-pointed in place
    if (same_shape >= 4 && recycle != nullptr) {
      cudaGraphExecUpdateResultInfo info;
      if (cudaGraphExecUpdate(recycle->exec, graph, &info) == cudaSuccess) {
        recycle->lr = lr;
        recycle->sr = sr;
        recycle->nodes = nodes;
        hit = recycle;
        ++h->graph_updates;
      } else {
        cudaGetLastError();
      }
    }
    if (!hit) {
      cudaGraphExec_t exec = nullptr;
      e = cudaGraphInstantiate(&exec, graph, 0);
      if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
      }
      ++h->graph_instantiations;
      if (h->graph_cache.size() >= 32) {  // many shapes: drop the globally oldest
        size_t o = 0;
        for (size_t i = 1; i < h->graph_cache.size(); ++i)
          if (h->graph_cache[i].stamp < h->graph_cache[o].stamp) o = i;
        cudaGraphExecDestroy(h->graph_cache[o].exec);
        h->graph_cache.erase(h->graph_cache.begin() + o);
      }
      h->graph_cache.push_back({N, H, W, lr, sr, exec, nodes, 0});
      hit = &h->graph_cache.back();
    }
    cudaGraphDestroy(graph);
  }
  hit->stamp = ++h->graph_clock;
  if (s == nullptr) {
    PFNL_CUDA(cudaEventRecord(h->graph_ev_in, s));
    PFNL_CUDA(cudaStreamWaitEvent(gs, h->graph_ev_in, 0));
  }
  PFNL_CUDA(cudaGraphLaunch(hit->exec, gs));
  if (s == nullptr) {
    PFNL_CUDA(cudaEventRecord(h->graph_ev_out, gs));
    PFNL_CUDA(cudaStreamWaitEvent(s, h->graph_ev_out, 0));
  }
  h->launches += hit->nodes;
  return PFNL_OK;
}

long long pfnl_graph_stats(const pfnl_handle* h, int what) {
  if (!h) return -1;
  return what == 0 ? h->graph_instantiations : (what == 1 ? h->graph_updates : (long long)h->graph_cache.size());
}

// Staging slot `k` sized for (in_bytes, out_bytes).
static int host_slot_prepare(pfnl_handle* h, int k, size_t in_bytes, size_t out_bytes) {
  pfnl_handle::HostSlot& sl = h->slot[k];
  if (!h->h2d_stream) {
    PFNL_CUDA(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    PFNL_CUDA(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
  }
  if (!sl.in_ready) {
    PFNL_CUDA(cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming));
    PFNL_CUDA(cudaEventCreateWithFlags(&sl.fwd_done, cudaEventDisableTiming));
    PFNL_CUDA(cudaEventCreateWithFlags(&sl.out_ready, cudaEventDisableTiming));
  }
  if (in_bytes > sl.in_cap) {
    if (sl.pin_in) cudaFreeHost(sl.pin_in);
    if (sl.dev_in) cudaFree(sl.dev_in);
    sl.pin_in = nullptr;
    sl.dev_in = nullptr;
    sl.in_cap = 0;
    PFNL_CUDA(cudaMallocHost((void**)&sl.pin_in, in_bytes));
    PFNL_CUDA(cudaMalloc((void**)&sl.dev_in, in_bytes));
    sl.in_cap = in_bytes;
  }
  if (out_bytes > sl.out_cap) {
    if (sl.pin_out) cudaFreeHost(sl.pin_out);
    if (sl.dev_out) cudaFree(sl.dev_out);
    sl.pin_out = nullptr;
    sl.dev_out = nullptr;
    sl.out_cap = 0;
    PFNL_CUDA(cudaMallocHost((void**)&sl.pin_out, out_bytes));
    PFNL_CUDA(cudaMalloc((void**)&sl.dev_out, out_bytes));
    sl.out_cap = out_bytes;
  }
  return PFNL_OK;
}

int pfnl_forward_host_wait(pfnl_handle* h, int ticket) {
  if (!h || ticket < 0 || ticket > 1) {
    set_error("pfnl_forward_host_wait: bad argument");
    return PFNL_ERR_BAD_ARG;
  }
  pfnl_handle::HostSlot& sl = h->slot[ticket];
  if (!sl.busy) return PFNL_OK;
  DeviceGuard guard(h->device);
  PFNL_CUDA(cudaEventSynchronize(sl.out_ready));
  if (sl.user_out) memcpy(sl.user_out, sl.pin_out, sl.out_bytes);
  sl.busy = false;
  sl.user_out = nullptr;
  return PFNL_OK;
}

int pfnl_forward_host_submit(pfnl_handle* h, const void* lr_host, int lr_dtype, int N, int H, int W, float* sr_host,
                             void* stream, int* ticket) {
  if (!h || !lr_host || !sr_host || !ticket || (lr_dtype != 0 && lr_dtype != 1)) {
    set_error("pfnl_forward_host_submit: bad argument");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n_in = (size_t)N * kFrames * H * W * 3;
  const size_t in_bytes = n_in * sizeof(float);
  const size_t out_bytes = (size_t)N * 16 * H * W * 3 * sizeof(float);
  const int k = h->next_slot;
  if ((rc = pfnl_forward_host_wait(h, k))) return rc;  // the slot's previous call (two submits ago)
  if ((rc = host_slot_prepare(h, k, in_bytes, out_bytes))) return rc;
  if ((rc = ensure_workspace(h, N, H, W))) return rc;
  pfnl_handle::HostSlot& sl = h->slot[k];
  // A registered (pinned) float32 caller buffer is copied directly; pageable memory goes through the slot's
  // pinned staging, float64 input (what the reference feeds, model/pfnl.py:209,252) is narrowed on the way in.
  cudaPointerAttributes at;
  const bool in_pinned = lr_dtype == 0 && cudaPointerGetAttributes(&at, lr_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
  const bool out_pinned = cudaPointerGetAttributes(&at, sr_host) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const float* src = (const float*)lr_host;
  if (!in_pinned) {
    if (lr_dtype == 1) {
      const double* d = (const double*)lr_host;
      for (size_t i = 0; i < n_in; ++i) sl.pin_in[i] = (float)d[i];
    } else {
      memcpy(sl.pin_in, lr_host, in_bytes);
    }
    src = sl.pin_in;
  }
  PFNL_CUDA(cudaMemcpyAsync(sl.dev_in, src, in_bytes, cudaMemcpyHostToDevice, h->h2d_stream));
  PFNL_CUDA(cudaEventRecord(sl.in_ready, h->h2d_stream));
  PFNL_CUDA(cudaStreamWaitEvent(s, sl.in_ready, 0));
  if ((rc = pfnl_forward(h, sl.dev_in, N, H, W, sl.dev_out, stream))) return rc;
  PFNL_CUDA(cudaEventRecord(sl.fwd_done, s));
  PFNL_CUDA(cudaStreamWaitEvent(h->d2h_stream, sl.fwd_done, 0));
  PFNL_CUDA(cudaMemcpyAsync(out_pinned ? sr_host : sl.pin_out, sl.dev_out, out_bytes, cudaMemcpyDeviceToHost,
                            h->d2h_stream));
  PFNL_CUDA(cudaEventRecord(sl.out_ready, h->d2h_stream));
  sl.busy = true;
  sl.user_out = out_pinned ? nullptr : sr_host;
  sl.out_bytes = out_bytes;
  h->next_slot = k ^ 1;
  *ticket = k;
  return PFNL_OK;
}

int pfnl_forward_host(pfnl_handle* h, const float* lr_host, int N, int H, int W, float* sr_host, void* stream) {
  if (!h || !lr_host || !sr_host) {
    set_error("pfnl_forward_host: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  int ticket = -1;
  int rc = pfnl_forward_host_submit(h, lr_host, 0, N, H, W, sr_host, stream, &ticket);
  if (rc) return rc;
  return pfnl_forward_host_wait(h, ticket);
}

int pfnl_mse(pfnl_handle* h, const float* sr, const float* hr, int N, int H4, int W4, float* mse, void* stream) {
  if (!h || !sr || !hr || !mse) {
    set_error("pfnl_mse: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || H4 <= 0 || W4 <= 0) {
    set_error("pfnl_mse: bad shape N=%d H4=%d W4=%d", N, H4, W4);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  if (N > h->mse_cap) {
    if (h->mse_partial) {
      PFNL_CUDA(cudaDeviceSynchronize());
      PFNL_CUDA(cudaFree(h->mse_partial));
      h->mse_partial = nullptr;
      h->mse_cap = 0;
    }
    PFNL_CUDA(cudaMalloc((void**)&h->mse_partial, (size_t)N * kMseChunks * sizeof(double)));
    h->mse_cap = N;
  }
  int rc = launch_mse(sr, hr, N, (long long)H4 * W4 * 3, h->mse_partial, mse, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 2;
  return PFNL_OK;
}

int pfnl_pack_tokens(pfnl_handle* h, const float* lr, int N, int H, int W, float* tokens, void* stream) {
  if (!h || !lr || !tokens) {
    set_error("pfnl_pack_tokens: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if ((rc = launch_pack_tokens(lr, N, H, W, tokens, (cudaStream_t)stream))) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_nonlocal(pfnl_handle* h, const float* tokens, int N, int L, float* out, void* stream) {
  if (!h || !tokens || !out) {
    set_error("pfnl_nonlocal: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || L <= 0) {
    set_error("pfnl_nonlocal: bad shape N=%d L=%d", N, L);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (tc_nl_on_tensor_cores(h->precision) && tc_has_nonlocal()) return tc_nonlocal_tokens(h->tcw, tokens, N, L, out, s, &h->launches, h->prof.on ? &h->prof : nullptr);
  // scratch: G and Y live in the workspace sized for an equivalent (N, 2, 2L) frame
  int rc = ensure_workspace(h, N, 2, 2 * L);
  if (rc) return rc;
  Workspace w = carve(h->ws, h->precision, N, 2, 2 * L);
  if ((rc = launch_nl_flash_ffma(tokens, tokens, N, L, w.yatt, s))) return rc;
  if ((rc = launch_nl_linear(w.yatt, N * L, h->nl_gw_w, h->nl_gw_b, out, s))) return rc;
  h->launches += 2;
  return PFNL_OK;
}

int pfnl_depth_to_space(pfnl_handle* h, const float* in, int N, int H, int W, int C, int block, float* out,
                        void* stream) {
  if (!h || !in || !out) {
    set_error("pfnl_depth_to_space: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || block <= 0 || C % (block * block) != 0) {
    set_error("pfnl_depth_to_space: bad shape N=%d H=%d W=%d C=%d block=%d", N, H, W, C, block);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  int rc = launch_depth_to_space(in, N, H, W, C, block, out, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_space_to_depth(pfnl_handle* h, const float* in, int N, int H, int W, int C, int block, float* out,
                        void* stream) {
  if (!h || !in || !out) {
    set_error("pfnl_space_to_depth: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || block <= 0 || H % block != 0 || W % block != 0) {
    set_error("pfnl_space_to_depth: bad shape N=%d H=%d W=%d C=%d block=%d", N, H, W, C, block);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  int rc = launch_space_to_depth(in, N, H, W, C, block, out, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_conv2d_nhwc(pfnl_handle* h, const float* in, int N, int H, int W, int Cin, const float* kernel,
                     const float* bias, int k, int Cout, int act, const float* residual, float* out, void* stream) {
  if (!h || !in || !kernel || !bias || !out) {
    set_error("pfnl_conv2d_nhwc: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (k != 1 && k != 3 && k != 5)) {
    set_error("pfnl_conv2d_nhwc: bad shape N=%d H=%d W=%d Cin=%d Cout=%d k=%d", N, H, W, Cin, Cout, k);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const bool aligned =
      ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)bias) | ((uintptr_t)residual)) & 15) == 0;
  int rc;
  if ((k == 1 || k == 3) && Cin % 16 == 0 && Cout % 4 == 0 && Cout <= 64 && aligned) {
    const size_t need = conv_ffma_packed_floats(k, Cin) * sizeof(float);
    if (need > h->pack_cap) {
      if (h->pack_scratch) {
        PFNL_CUDA(cudaDeviceSynchronize());
        PFNL_CUDA(cudaFree(h->pack_scratch));
        h->pack_scratch = nullptr;
        h->pack_cap = 0;
      }
      PFNL_CUDA(cudaMalloc((void**)&h->pack_scratch, need));
      h->pack_cap = need;
    }
    if ((rc = launch_pack_conv_ffma_weights(kernel, k, Cin, Cout, h->pack_scratch, s))) return rc;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.nslices = 1;
    a.slice_ch = Cin;
    a.slice[0] = make_slice(in, (long long)H * W * Cin, 1, Cin);
    a.images = N;
    a.H = H;
    a.W = W;
    a.wpack = h->pack_scratch;
    a.bias = bias;
    a.cout = Cout;
    a.act = act;
    a.residual = residual;
    a.out = out;
    if ((rc = launch_conv_ffma(k, a, s))) return rc;
    h->launches += 2;
  } else {
    if ((rc = launch_conv_direct(in, N, H, W, Cin, kernel, bias, k, Cout, act, residual, out, s))) return rc;
    h->launches += 1;
  }
  return PFNL_OK;
}

int pfnl_bicubic4(pfnl_handle* h, const float* in, int N, int H, int W, int C, float* out, void* stream) {
  if (!h || !in || !out) {
    set_error("pfnl_bicubic4: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0) {
    set_error("pfnl_bicubic4: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  int rc = launch_bicubic4(in, N, H, W, C, out, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_pfrb(pfnl_handle* h, int blk, const float* frames, int N, int H, int W, float* frames_out, void* stream) {
  if (!h || !frames || !frames_out) {
    set_error("pfnl_pfrb: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (blk < 0 || blk >= PFNL_NUM_BLOCK) {
    set_error("pfnl_pfrb: block index %d out of range", blk);
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h, N, H, W))) return rc;
  Workspace w = carve_h(h, N, H, W);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->precision == PFNL_PREC_FP32) return pfrb_fp32(h, blk, frames, frames_out, w, N, H, W, s);
  return tc_pfrb_fp32io(h->tcw, w.tc, h->precision, blk, frames, N, H, W, frames_out, s, &h->launches);
}

int pfnl_conv0(pfnl_handle* h, const float* inp21, int N, int H, int W, float* frames_out, void* stream) {
  if (!h || !inp21 || !frames_out) {
    set_error("pfnl_conv0: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h, N, H, W))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->precision == PFNL_PREC_FP32) {
    if ((rc = launch_conv0(inp21, N, H, W, h->conv0_w, h->conv0_b, frames_out, s))) return rc;
    h->launches += 1;
    return PFNL_OK;
  }
  Workspace w = carve_h(h, N, H, W);
  return tc_conv0_fp32io(h->tcw, w.tc, h->precision, inp21, N, H, W, frames_out, s, &h->launches);
}

int pfnl_convmerge1(pfnl_handle* h, const float* frames, int N, int H, int W, float* merge, void* stream) {
  if (!h || !frames || !merge) {
    set_error("pfnl_convmerge1: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  int rc = check_shape(N, H, W);
  if (rc) return rc;
  DeviceGuard guard(h->device);
  if ((rc = ensure_workspace(h, N, H, W))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->precision == PFNL_PREC_FP32) {
    const long long hw = (long long)H * W;
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.H = H;
    a.W = W;
    a.act = 1;
    a.cout = 48;
    a.nslices = kFrames;
    a.slice_ch = kMF;
    for (int t = 0; t < kFrames; ++t) a.slice[t] = make_slice(frames + t * hw * kMF, kFrames * hw * kMF, 1, kMF);
    a.images = N;
    a.wpack = h->merge1_p;
    a.bias = h->merge1_b;
    a.out = merge;
    if ((rc = launch_conv_ffma(3, a, s))) return rc;
    h->launches += 1;
    return PFNL_OK;
  }
  Workspace w = carve_h(h, N, H, W);
  return tc_merge1_fp32io(h->tcw, w.tc, h->precision, frames, N, H, W, merge, s, &h->launches);
}

int pfnl_downsample4(pfnl_handle* h, const float* hr, int F, int H, int W, const float* blur_host, float* lr,
                     void* stream) {
  if (!h || !hr || !blur_host || !lr) {
    set_error("pfnl_downsample4: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (F <= 0 || H < 7 || W < 7) {  // REFLECT padding of 6 needs at least 7 pixels
    set_error("pfnl_downsample4: bad shape F=%d H=%d W=%d", F, H, W);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->blur_dev) PFNL_CUDA(cudaMalloc((void**)&h->blur_dev, 169 * sizeof(float)));
  PFNL_CUDA(cudaMemcpyAsync(h->blur_dev, blur_host, 169 * sizeof(float), cudaMemcpyHostToDevice, s));
  int rc = launch_downsample4(hr, F, H, W, h->blur_dev, lr, s);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_gather_windows(pfnl_handle* h, const float* frames, int F, int fh, int fw, int first, int count,
                        float* clips, void* stream) {
  if (!h || !frames || !clips) {
    set_error("pfnl_gather_windows: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (F <= 0 || fh <= 0 || fw <= 0 || count <= 0 || first < 0 || first + count > F) {
    set_error("pfnl_gather_windows: bad shape F=%d h=%d w=%d first=%d count=%d", F, fh, fw, first, count);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  int rc = launch_gather_windows(frames, F, (long long)fh * fw * 3, first, count, clips, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}

int pfnl_quantize_u8(pfnl_handle* h, const float* in, long long n, unsigned char* out, void* stream) {
  if (!h || !in || !out) {
    set_error("pfnl_quantize_u8: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (n <= 0) {
    set_error("pfnl_quantize_u8: bad size %lld", n);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  int rc = launch_quantize_u8(in, n, out, (cudaStream_t)stream);
  if (rc) return rc;
  h->launches += 1;
  return PFNL_OK;
}


// luma planes of both inputs + partial sums live in one lazily grown buffer
static int metric_scratch(pfnl_handle* h, int F, int H, int W, double** ya, double** yb, double** partial) {
  const size_t npix = (size_t)F * H * W;
  const size_t need = 2 * npix + metric_partials(F, H, W);
  if (need > h->metric_cap) {
    if (h->metric_buf) {
      PFNL_CUDA(cudaDeviceSynchronize());
      PFNL_CUDA(cudaFree(h->metric_buf));
      h->metric_buf = nullptr;
      h->metric_cap = 0;
    }
    PFNL_CUDA(cudaMalloc((void**)&h->metric_buf, need * sizeof(double)));
    h->metric_cap = need;
  }
  *ya = h->metric_buf;
  *yb = h->metric_buf + npix;
  *partial = h->metric_buf + 2 * npix;
  return PFNL_OK;
}

int pfnl_msy(pfnl_handle* h, const float* a, const float* b, int F, int H, int W, float vmin, float vmax,
             int sp_border, int round_y, double* out, void* stream) {
  if (!h || !a || !b || !out) {
    set_error("pfnl_msy: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (F <= 0 || H <= 0 || W <= 0 || sp_border < 0 || 2 * sp_border >= H || 2 * sp_border >= W || !(vmax > vmin)) {
    set_error("pfnl_msy: bad shape F=%d H=%d W=%d sp_border=%d (or vmax <= vmin)", F, H, W, sp_border);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  double *ya, *yb, *partial;
  int rc = metric_scratch(h, F, H, W, &ya, &yb, &partial);
  if (rc) return rc;
  const long long npix = (long long)F * H * W;
  if ((rc = launch_luma(a, npix, vmin, vmax, round_y, ya, s))) return rc;
  if ((rc = launch_luma(b, npix, vmin, vmax, round_y, yb, s))) return rc;
  if ((rc = launch_ysq(ya, yb, F, H, W, sp_border, partial, out, s))) return rc;
  h->launches += 4;
  return PFNL_OK;
}

int pfnl_ssim_y(pfnl_handle* h, const float* a, const float* b, int F, int H, int W, float vmin, float vmax,
                double* out, void* stream) {
  if (!h || !a || !b || !out) {
    set_error("pfnl_ssim_y: NULL argument");
    return PFNL_ERR_BAD_ARG;
  }
  if (F <= 0 || H < 11 || W < 11 || !(vmax > vmin)) {  // SSIM.m returns -Inf below 11x11
    set_error("pfnl_ssim_y: bad shape F=%d H=%d W=%d (needs at least 11x11; or vmax <= vmin)", F, H, W);
    return PFNL_ERR_BAD_SHAPE;
  }
  DeviceGuard guard(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  double *ya, *yb, *partial;
  int rc = metric_scratch(h, F, H, W, &ya, &yb, &partial);
  if (rc) return rc;
  const long long npix = (long long)F * H * W;
  if ((rc = launch_luma(a, npix, vmin, vmax, 1, ya, s))) return rc;
  if ((rc = launch_luma(b, npix, vmin, vmax, 1, yb, s))) return rc;
  if ((rc = launch_ssim(ya, yb, F, H, W, partial, out, s))) return rc;
  h->launches += 4;
  return PFNL_OK;
}

// Table-driven CRC-32C, 8 bytes per step (slicing-by-8); host only.
uint32_t pfnl_crc32c(const void* data, size_t n, uint32_t crc) {
  struct Table {
    uint32_t t[8][256];
    Table() {
      for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);
        t[0][i] = c;
      }
      for (uint32_t i = 0; i < 256; ++i)
        for (int k = 1; k < 8; ++k) t[k][i] = (t[k - 1][i] >> 8) ^ t[0][t[k - 1][i] & 0xFFu];
    }
  };
  static const Table table;  // function-local static: initialised once, thread-safely (C++11)
  const uint32_t (*tab)[256] = table.t;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = ~crc;
  while (n >= 8) {
    const uint32_t lo = c ^ ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    c = tab[7][lo & 0xFFu] ^ tab[6][(lo >> 8) & 0xFFu] ^ tab[5][(lo >> 16) & 0xFFu] ^ tab[4][lo >> 24] ^
        tab[3][p[4]] ^ tab[2][p[5]] ^ tab[1][p[6]] ^ tab[0][p[7]];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return ~c;
}

}  // extern "C"
