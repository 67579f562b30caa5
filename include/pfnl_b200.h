/* pfnl_b200.h - C ABI of libpfnl_b200.so: the B200 (sm_100a) implementation of the PFNL
 * 4x multi-frame forward hot path.
 *
 * The reference (psychopa4/PFNL, TensorFlow 1.12, pure Python) has no FFI; the seam this
 * library plugs into is the Python method boundary PFNL.forward(x) and its call sites
 *     sr = sess.run(SR_test, feed_dict={L_test: batch})      model/pfnl.py:252, :309
 *     mse_val = sess.run(self.eval_mse, feed_dict={...})      model/pfnl.py:130 (graph :90)
 * Each entry point below names the reference code it replaces.  Plain pointers and sizes
 * only - no torch / TF types.  All tensors are fp32, channels-last, contiguous.
 *
 * Conventions
 *   - every function returns PFNL_OK (0) or a negative pfnl_status; the message for the
 *     calling thread's last failure is pfnl_last_error().
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are
 *     ordinary host memory.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Calls are asynchronous on `stream` unless stated otherwise.
 *   - the caller owns every I/O buffer and keeps it alive until the stream has drained.
 *   - a handle is bound to one device and is not thread-safe; distinct handles are
 *     independent.  There is NO CPU fallback: on a device that is not sm_100 every compute
 *     entry point returns PFNL_ERR_UNSUPPORTED_ARCH.
 */
#ifndef PFNL_B200_H
#define PFNL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFNL_VERSION 100 /* 0.1.0 */

#define PFNL_NUM_FRAMES 7  /* model/pfnl.py:22 */
#define PFNL_SCALE 4       /* model/pfnl.py:23 */
#define PFNL_NUM_BLOCK 20  /* model/pfnl.py:43 */
#define PFNL_MF 64         /* model/pfnl.py:40 */
#define PFNL_NL_CH 84      /* 3*7*4, model/pfnl.py:58 */

typedef enum pfnl_status {
  PFNL_OK = 0,
  PFNL_ERR_BAD_ARG = -1,          /* NULL pointer, unknown enum value */
  PFNL_ERR_BAD_SHAPE = -2,        /* odd H/W (tf.space_to_depth needs even, pfnl.py:57), N<=0 ... */
  PFNL_ERR_CUDA = -3,             /* a CUDA runtime/driver call failed */
  PFNL_ERR_UNSUPPORTED_ARCH = -4, /* device is not compute capability 10.x */
  PFNL_ERR_UNIMPLEMENTED = -5,
  PFNL_ERR_NO_MEMORY = -6
} pfnl_status;

/* Arithmetic used by the conv / non-local kernels.  Reorders are index-exact in all modes. */
typedef enum pfnl_precision {
  PFNL_PREC_FP32 = 0,       /* FFMA everywhere: the <=1e-3 parity path */
  PFNL_PREC_TC_FP16X3 = 1,  /* tcgen05 convs on hi/lo-split fp16 operands (3 MMAs, fp32 accumulate) */
  PFNL_PREC_TC_FP16 = 2,    /* tcgen05 convs + tcgen05 non-local block on fp16 operands, fp32 accumulate */
  PFNL_PREC_TC_FP16X3_NLTC = 3 /* convs as FP16X3, non-local block on tcgen05 with fp16 operands (BASELINE configs[1]:
                                  "fp16 tensor-core non-local"); meets 1e-3 only for trained-like weights */
} pfnl_precision;

/* Weights, HOST pointers, TF layouts (kernels HWIO [kh,kw,Cin,Cout], biases [Cout]).
 * Field <-> TF variable (scope 'nlvsr', model/pfnl.py:47-53; utils.py:23-26,66-67):
 *   nl_g_*      nlvsr/nlblock_0/g/g/{kernel,bias}   [1,1,84,84],[84]
 *   nl_w_*      nlvsr/nlblock_0/w/w/{kernel,bias}   [1,1,84,84],[84]
 *   conv0_*     nlvsr/conv0/{kernel,bias}           [5,5,3,64],[64]
 *   conv1_*[i]  nlvsr/conv1_{i}/...                 [3,3,64,64],[64]
 *   conv10_*[i] nlvsr/conv10_{i}/...                [1,1,448,64],[64]  (Cin = t*64+c, pfnl.py:67)
 *   conv2_*[i]  nlvsr/conv2_{i}/...                 [3,3,128,64],[64]  (Cin 0-63 base, 64-127 frame, pfnl.py:69)
 *   merge1_*    nlvsr/convmerge1/...                [3,3,448,48],[48]
 *   merge2_*    nlvsr/convmerge2/...                [3,3,12,12],[12]
 * pfnl_create copies and repacks them; the caller may free them afterwards. */
typedef struct pfnl_weights {
  const float* nl_g_kernel;
  const float* nl_g_bias;
  const float* nl_w_kernel;
  const float* nl_w_bias;
  const float* conv0_kernel;
  const float* conv0_bias;
  const float* conv1_kernel[PFNL_NUM_BLOCK];
  const float* conv1_bias[PFNL_NUM_BLOCK];
  const float* conv10_kernel[PFNL_NUM_BLOCK];
  const float* conv10_bias[PFNL_NUM_BLOCK];
  const float* conv2_kernel[PFNL_NUM_BLOCK];
  const float* conv2_bias[PFNL_NUM_BLOCK];
  const float* merge1_kernel;
  const float* merge1_bias;
  const float* merge2_kernel;
  const float* merge2_bias;
} pfnl_weights;

typedef struct pfnl_handle pfnl_handle;

int pfnl_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* pfnl_last_error(void);
/* 1 if `device` is compute capability 10.x, 0 if not, negative status on error. */
int pfnl_device_supported(int device);

/* Replaces: graph construction + variable initialisation/restore in test_video_*
 * (model/pfnl.py:220-232, 281-291): builds the per-device state for PFNL.forward. */
int pfnl_create(pfnl_handle** out, int device, const pfnl_weights* weights, int precision);
int pfnl_destroy(pfnl_handle* h);

/* Bytes of device workspace pfnl_forward needs for a batch of N clips of HxW LR frames. */
size_t pfnl_workspace_bytes(int precision, int N, int H, int W);
/* Pre-allocates the workspace for (N,H,W) so that pfnl_forward does not allocate
 * (and can be captured in a CUDA graph).  Synchronous. */
int pfnl_reserve(pfnl_handle* h, int N, int H, int W);
/* 1: pfnl_forward replays a cached CUDA graph per (N,H,W) shape and buffer pair (a new buffer pair of a known shape
 * re-points an existing executable instead of instantiating again; the legacy NULL stream is served through a
 * private stream ordered with events); 0: plain launches. */
int pfnl_set_graphs(pfnl_handle* h, int enable);
/* Tensor-core precisions only.  1 (default): the 20 PFRBs (model/pfnl.py:65-71) run as ONE persistent dataflow
 * kernel (csrc/pfrb_flow.cu); 0: two launches per block (csrc/conv_tc.cu).  Same arithmetic, bit-identical
 * results; kept switchable for A/B measurements and tests.  PFNL_TC_FLOW=0 in the environment sets the default. */
int pfnl_set_flow(pfnl_handle* h, int enable);
/* Debugging aid.  Every device-side wait of the tensor-core kernels is bounded; the first one that gives up records
 * {1 + kind, CTA, detail, detail} in host-mapped memory before it traps (the launch then fails with a CUDA error).
 * Readable even after the context is lost.  kind 0: mbarrier (detail = smem address, parity); 1-4: a producer of the
 * PFRB dataflow kernel waiting for the inputs of (block, item) in role kind-1; 5: its conv2 epilogue. */
int pfnl_debug_fault(int* out4);
/* With PFNL_FLOW_DEBUG=1 every CTA of the PFRB dataflow kernel keeps 8 ints of progress marks in host-mapped memory
 * ([0..2] producer: state, block, item; [3] MMA warp: tiles issued; [4..6] epilogue: state, block, item; [7] role):
 * copies the first n ints (n <= 2048).  Slow; for post-mortems of a timed-out wait only. */
int pfnl_debug_progress(int* out, int n);
/* CTAs the PFRB dataflow kernel gives to conv1, conv10, conv2b, conv2f on a device of num_sms SMs for a problem of
 * n_units (clips x 16x8 tiles); their sum is num_sms.  Host arithmetic only (no GPU needed). */
int pfnl_debug_flow_split(int num_sms, int n_units, int* out4);

/* Replaces: sess.run(SR_test, feed_dict={L_test: lr}) -> PFNL.forward, model/pfnl.py:39-80.
 *   lr_dev [N,7,H,W,3] -> sr_dev [N,1,4H,4W,3]; H and W even. */
int pfnl_forward(pfnl_handle* h, const float* lr_dev, int N, int H, int W, float* sr_dev,
                 void* stream);
/* Same call with HOST buffers, mirroring the feed/fetch of model/pfnl.py:251-253: copies
 * lr to the device (through pinned staging), runs the forward, copies sr back, and
 * returns when sr_host is complete. */
int pfnl_forward_host(pfnl_handle* h, const float* lr_host, int N, int H, int W, float* sr_host,
                      void* stream);
/* The same feed/fetch, pipelined over consecutive batches (the loop of test_video_*, model/pfnl.py:246-253): submit
 * copies lr to the device on a copy stream, runs the forward on `stream` and starts the copy of sr back on another
 * copy stream, then returns a ticket (0 or 1: two staging slots); wait(ticket) returns when that call's sr_host is
 * complete.  Submitting batch k+1 before waiting for batch k overlaps its H2D (and the D2H of batch k) with compute.
 * A submit first waits for the call that used its slot two submits earlier.  lr_dtype 0: float32, 1: float64 (the
 * reference feeds float64 numpy arrays into its float32 placeholder, model/pfnl.py:209,252) - narrowed on the way
 * into pinned staging.  Pinned float32 inputs / pinned outputs are copied directly. */
int pfnl_forward_host_submit(pfnl_handle* h, const void* lr_host, int lr_dtype, int N, int H, int W, float* sr_host,
                             void* stream, int* ticket);
int pfnl_forward_host_wait(pfnl_handle* h, int ticket);
/* CUDA-graph cache counters: what = 0 executables instantiated, 1 executables re-pointed in place
 * (cudaGraphExecUpdate), 2 executables currently cached. */
long long pfnl_graph_stats(const pfnl_handle* h, int what);

/* Replaces: eval_mse = tf.reduce_mean((SR-H)**2, axis=[2,3,4]), model/pfnl.py:90.
 *   sr_dev, hr_dev [N,1,H4,W4,3] -> mse_dev [N]. */
int pfnl_mse(pfnl_handle* h, const float* sr_dev, const float* hr_dev, int N, int H4, int W4,
             float* mse_dev, void* stream);

/* Kernel launches issued by this handle since creation (graph replays count their nodes). */
long long pfnl_launch_count(const pfnl_handle* h);

/* Per-kernel-class device timing for roofline reporting.  pfnl_profile(h,1) makes every
 * following pfnl_forward bracket its launches with CUDA events on the launching stream (CUDA
 * graphs are bypassed while it is on); pfnl_profile_read synchronises the device, returns the
 * summed milliseconds and launch counts per class since the last read, and resets them.
 * Classes (PFNL_PROF_KINDS = 9): 0 pack_tokens, 1 non-local, 2 conv0, 3 conv1 (3x3 64->64),
 * 4 conv10 (1x1 448->64), 5 conv2 (3x3 128->64 + residual), 6 convmerge1, 7 tail, 8 other. */
#define PFNL_PROF_KINDS 9
int pfnl_profile(pfnl_handle* h, int enable);
int pfnl_profile_read(pfnl_handle* h, double* ms_by_kind, long long* launches_by_kind);

/* ---- stage-level entry points (isolation benchmarks and parity tests) ---------------- */

/* tf.concat(frames,-1) + tf.space_to_depth(.,2), model/pfnl.py:55-57:
 *   lr [N,7,H,W,3] -> tokens [N,(H/2)*(W/2),84], channel = (dy*2+dx)*21 + t*3 + c. */
int pfnl_pack_tokens(pfnl_handle* h, const float* lr_dev, int N, int H, int W, float* tokens_dev,
                     void* stream);
/* NonLocalBlock(nltype=1, sub_sample=1), utils.py:18-71: tokens [N,L,84] -> [N,L,84]. */
int pfnl_nonlocal(pfnl_handle* h, const float* tokens_dev, int N, int L, float* out_dev,
                  void* stream);
/* tf.depth_to_space (DCR), model/pfnl.py:59,76,78 == modules/ps.py:_PS:
 *   in [N,H,W,C] -> out [N,H*b,W*b,C/(b*b)]. */
int pfnl_depth_to_space(pfnl_handle* h, const float* in_dev, int N, int H, int W, int C, int block,
                        float* out_dev, void* stream);
/* tf.space_to_depth, model/pfnl.py:57: in [N,H,W,C] -> out [N,H/b,W/b,C*b*b]. */
int pfnl_space_to_depth(pfnl_handle* h, const float* in_dev, int N, int H, int W, int C, int block,
                        float* out_dev, void* stream);
/* tf.layers.Conv2D(strides=1,padding='same') NHWC/HWIO, model/pfnl.py:48-53:
 *   in [N,H,W,Cin], kernel [k,k,Cin,Cout] (device), bias [Cout] (device),
 *   optional residual [N,H,W,Cout] added after the activation (model/pfnl.py:71),
 *   act: 0 none, 1 leaky_relu(0.2).  k in {1,3,5}.  Always fp32 FFMA. */
int pfnl_conv2d_nhwc(pfnl_handle* h, const float* in_dev, int N, int H, int W, int Cin,
                     const float* kernel_dev, const float* bias_dev, int k, int Cout, int act,
                     const float* residual_dev, float* out_dev, void* stream);
/* tf.image.resize_images(img,[4H,4W],method=2), model/pfnl.py:63: [N,H,W,C] -> [N,4H,4W,C]. */
int pfnl_bicubic4(pfnl_handle* h, const float* in_dev, int N, int H, int W, int C, float* out_dev,
                  void* stream);
/* One Progressive Fusion Residual Block (model/pfnl.py:66-71) with block index `blk`'s
 * weights, in the handle's precision: frames [N*7,H,W,64] fp32 -> frames_out (may alias). */
int pfnl_pfrb(pfnl_handle* h, int blk, const float* frames_dev, int N, int H, int W,
              float* frames_out_dev, void* stream);
/* conv0 = Conv2D(64, 5, 'same', leaky_relu) applied to each of the 7 frames (model/pfnl.py:48,61-62), in the
 * handle's precision: inp21 [N,H,W,21] fp32 (frame t = channels 3t..3t+2) -> frames_out [N*7,H,W,64] fp32. */
int pfnl_conv0(pfnl_handle* h, const float* inp21_dev, int N, int H, int W, float* frames_out_dev, void* stream);
/* convmerge1 = Conv2D(48, 3, 'same', leaky_relu) over the channel concat of the 7 frames (model/pfnl.py:52,73-74),
 * in the handle's precision: frames [N*7,H,W,64] fp32 -> merge [N,H,W,48] fp32. */
int pfnl_convmerge1(pfnl_handle* h, const float* frames_dev, int N, int H, int W, float* merge_dev, void* stream);

/* ---- the steps around the hot path in test_video_truth / test_video_lr ---- */

/* DownSample_4D (utils.py:169-192) with the 13x13 blur of utils.py:95-105: REFLECT pad 6, stride 4.
 *   hr_dev [F,H,W,3] -> lr_dev [F,(H-1)/4+1,(W-1)/4+1,3]; blur_host = the 169 fp32 taps (row-major). */
int pfnl_downsample4(pfnl_handle* h, const float* hr_dev, int F, int H, int W, const float* blur_host,
                     float* lr_dev, void* stream);
/* Sliding 7-frame windows with edge clamping (model/pfnl.py:236-242, 294-300):
 *   frames_dev [F,h,w,3] -> clips_dev [count,7,h,w,3], clip k = frames clamp(first+k-3 .. first+k+3, 0, F-1). */
int pfnl_gather_windows(pfnl_handle* h, const float* frames_dev, int F, int fh, int fw, int first, int count,
                        float* clips_dev, void* stream);
/* round(clip(sr*255,0,255)).astype(uint8) (model/pfnl.py:255-257, round-half-to-even like np.round):
 *   in_dev fp32 [n] -> out_dev uint8 [n]. */
int pfnl_quantize_u8(pfnl_handle* h, const float* in_dev, long long n, unsigned char* out_dev, void* stream);


/* ---- evaluation metrics on the luma channel ---- */

/* Mean squared Y difference per frame: utils.py AVG_PSNR:216-246 (to_uint8 -> _rgb2ycbcr[:,:,0] -> crop
 * sp_border -> mean(diff^2); PSNR = 20*log10(255/sqrt(.)) is left to the host) and, with round_y = 1 and
 * sp_border = 0, matlab/compute_psnr.m (MATLAB's rgb2ycbcr of a uint8 image returns a rounded uint8 Y).
 *   a_dev, b_dev [F,H,W,3] fp32 RGB in [vmin,vmax] -> msy_dev [F] double. */
int pfnl_msy(pfnl_handle* h, const float* a_dev, const float* b_dev, int F, int H, int W, float vmin, float vmax,
             int sp_border, int round_y, double* msy_dev, void* stream);
/* Mean SSIM per frame exactly as matlab/SSIM.m computes it for RGB uint8 images with its default
 * arguments: Y = uint8 rgb2ycbcr, 11x11 Gaussian window (sigma 1.5), 'valid', K = (0.01,0.03), L = 255.
 *   a_dev, b_dev [F,H,W,3] fp32 RGB in [vmin,vmax] (quantised to uint8 first, as a saved PNG is), H,W >= 11
 *   -> ssim_dev [F] double. */
int pfnl_ssim_y(pfnl_handle* h, const float* a_dev, const float* b_dev, int F, int H, int W, float vmin,
                float vmax, double* ssim_dev, void* stream);

/* ---- host utility ---- */

/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the checksum of TensorFlow's
 * tensor-bundle checkpoints (tf.train.Saver, base_model.py:219-243) read by pfnl_b200/tf_checkpoint.py. */
uint32_t pfnl_crc32c(const void* data_host, size_t n, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif /* PFNL_B200_H */
