// Tensor-core convolutions of the PFRB stack (model/pfnl.py:65-74) for sm_100a:
// TMA -> shared memory -> tcgen05.mma (fp16 operands, fp32 accumulators in TMEM) -> fused epilogue.
//
// Data layout in HBM: activations are fp16 planes stored channel-chunk-major,
// [images][8 chunks][H][W][8 ch]: an epilogue warp owns 32 pixels (4 tile rows x 8) x 16 channels, so
// with this layout each of its 16-byte-per-lane accesses covers 4 runs of 128 contiguous bytes (4 L1
// wavefronts per instruction, every byte of every line used) instead of touching 32 different
// 128-byte lines for 32 bytes each as pixel-major NHWC rows do.  The L1/shared-memory data path is
// the shared resource of this kernel - the UMMA operand reads alone keep it ~90 % busy - so epilogue
// wavefronts cost tile time one for one (ncu: l1tex 53 % with the NHWC layout, profiles/r10_*).
// PFNL_PREC_TC_FP16 keeps one plane per tensor;
// PFNL_PREC_TC_FP16X3 keeps two (hi = fp16(v), lo = fp16((v-hi)*2048)) and computes
//   [D0 | D1] += A_hi x [W_hi ; W_lo]   (one N = 128 MMA per k-step: one A read for two products)
//        D1   += A_lo x  W_hi           (N = 64)
// result D0 + D1/2048: ~22 mantissa bits through the fp16 tensor pipe.  TMEM accumulation
// truncates (probes/umma_probe.cu), so the 36 k-steps of a 3x3 tile are split into 2 accumulation
// chains that the epilogue sums round-to-nearest.
//
// Implicit GEMM: M = 128 output pixels (a 16-row x 8-column spatial tile of one image),
// N = 64 (48 for convmerge1) output channels, K = taps x 64 input channels.  Per tile ONE TMA
// box {10 px x 8 ch, 18 rows, 8 chunks} (out-of-bounds zero fill == 'same' zero padding) lands a
// halo patch in smem as 8 sub-patches of [18][10] 16-byte pixels, un-swizzled: 8 consecutive pixels
// of a row are one UMMA core matrix (8 rows x 16 B, contiguous).  The A operand of tap (dy,dx),
// k-step k is sub-patches 2k, 2k+1 (LBO = sub-patch size) addressed through a descriptor whose
// start is shifted by (dy*10+dx) pixels and whose 8-row-group stride (SBO) is one patch row (10
// pixels) - each input byte is fetched from L2 once per tile, not 9 times, and no alignment or
// swizzle phase constrains the shift.
// Weights ([taps][N][64] fp16, pre-swizzled) are resident in smem while a phase runs.
//
// One launch runs several PHASES back to back on the same persistent CTAs, swapping the weight image
// in smem between them:   conv1 -> conv10,   conv2(base half, fp32 partial) -> conv2(frame half),   and the
// 7 frame slices of convmerge1.  Everything that enters shared memory goes through one TMA queue, so the
// first patch of a phase whose inputs do not depend on the previous one is issued BEFORE the weight swap.
// A work unit is one spatial tile x 7 frames, and a CTA owns the same units in both phases, so the
// phase-2 inputs that phase 1 produced are the CTA's own writes (conv10 reads the 7 conv1 tiles of
// its unit; the frame half adds the partial sums its own threads stored).  This halves the launch
// count of the PFRB stack: the 128-tile conv10 / partial-sum kernels were dominated by launch gap,
// weight load and first-load latency, not by their MMAs.
//
// Warp roles (576 threads, 1 CTA / SM, persistent): warp 0 = TMA producer, warp 1 = MMA issuer
// (converged warp, one elected lane issues; descriptors advance by constant adds), warps 2-17 =
// epilogue (16 warps: 4 TMEM lane quarters x 4 sixteen-channel chunks; tcgen05.ld -> bias /
// leaky_relu / partial sums / residual -> fp16 planes (16-byte accesses, 4 full lines per warp
// instruction) or fp32, operands prefetched before the accumulator wait).  TMEM accumulators are double buffered.  Programmatic
// dependent launch overlaps the prologue with the previous kernel's tail.
#include "conv_tc_dev.cuh"

namespace pfnl {

struct TcPhase {
  CUtensorMap tm_hi, tm_lo;  // source planes
  const __half* wimg;        // weight image of this phase
  int frames, n_units;       // unit = one spatial tile x `frames` consecutive output images
  int img_mul, img_add;      // source image coordinate of stage s = out_img*img_mul + img_add + s
  int epi;
  int accumulate;            // kEpiPartialF32: add the previous content of out_f32
  int f32_chunked;           // out_f32 is chunk-major [img][4][H][W][16] (the conv2 partial sums) instead of NHWC
  const float* bias;         // [NOUT] or NULL
  const float* pbase;        // fp32 [out_img/frames][4][H][W][16], channel-chunk-major (kEpiResPlanes)
  __half* out_hi;
  __half* out_lo;
  const __half* res_hi;
  const __half* res_lo;
  float* out_f32;            // [out_images][H][W][NOUT]
};

constexpr int kTcMaxPhases = 7;
struct TcProgram {
  TcPhase ph[kTcMaxPhases];
};

struct TcCommon {
  int H, W, tiles_x, tiles_y;
  int nphases;
  int phase1_reads_phase0;  // phase 1 TMA-loads tensors that phase 0 of the same CTA wrote (conv1 -> conv10)
  float trunc_comp;         // see conv_tc_dev.cuh
  long long* trace;  // debug (PFNL_TC_TRACE=1): clock64 stamps of CTA 0, [role][event]; else NULL
};

#define TC_TRACE(role, ev)                                                                              \
  do {                                                                                                  \
    if (cm.trace != nullptr && blockIdx.x == 0 && (ev) < 64) cm.trace[(role) * 64 + (ev)] = clock64(); \
  } while (0)


template <class P0, class P1, int NSPLIT>
struct KernelCfg {
  static constexpr int TAP0 = NSPLIT * P0::WT_BYTES, TAP1 = NSPLIT * P1::WT_BYTES;  // [W_hi ; W_lo] stacked
  static constexpr int W0 = P0::NTAPS * TAP0, W1 = P1::NTAPS * TAP1;
  static constexpr int W_BYTES = cmax(W0, W1);
  static constexpr int SLOT_BYTES = (cmax(P0::PATCH_BYTES, P1::PATCH_BYTES) + 1023) / 1024 * 1024;
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int CTRL_BYTES = 3072;
  static constexpr int TOTAL_SLOTS = (SMEM_MAX - 1024 - CTRL_BYTES - W_BYTES) / SLOT_BYTES;
  // ONE ring of patch slots shared by the hi and lo planes (loads alternate hi, lo, hi, lo ...)
  static constexpr int NS = TOTAL_SLOTS > 6 ? 6 : TOTAL_SLOTS;
  static constexpr int SMEM_BYTES = 1024 + CTRL_BYTES + W_BYTES + NS * SLOT_BYTES;
  // TMEM per accumulator buffer: NCH chains; a chain is [D0 (NOUT cols) | D1 (NOUT cols, split mode)]
  static constexpr int CH_STRIDE = NSPLIT == 2 ? 128 : 64;
  static constexpr int TMEM_BUF_COLS = cmax(P0::NCH, P1::NCH) * CH_STRIDE;
  static constexpr int TMEM_NEED = 2 * TMEM_BUF_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "TMEM overflow");
  static_assert(NS >= (NSPLIT == 2 ? 2 : 1), "shared memory budget too small");
  static_assert(SMEM_BYTES <= SMEM_MAX, "shared memory overflow");
};

constexpr int kPrefetchAhead = 0;                   // tiles of L2 prefetch distance in the TMA producer (0 = off: the
                                                    // prefetches occupy the same TMA engine as the loads they hide)

struct TcCtrl {
  uint64_t wfull;        // weight image of the current phase has landed
  uint64_t wfree;        // all MMAs of the finished phase are complete (weights may be overwritten)
  uint64_t stores_done;  // all epilogue stores of the finished phase are globally visible
  TcBars bars;
  uint32_t tmem_base;
  float bias[kTcMaxPhases][64];
};

// ---------------------------------------------------------------------------------------------------------
// role bodies, one call per phase (the per-tile bodies live in conv_tc_dev.cuh)
// ---------------------------------------------------------------------------------------------------------
// Issues the patch loads of tiles [first, first + count) of this CTA's tile sequence in phase P (count < 0:
// all remaining).  The first tile of a phase whose inputs do not depend on the previous phase is issued
// BEFORE the weight swap (`first = 0, count = 1`), so that its 46 KB do not queue behind the 147 KB image
// in the TMA engine, which moves only ~40 B/cycle.
template <class PC, class KC, int NSPLIT>
__device__ __forceinline__ void producer_phase(const TcPhase& P, const TcCommon& cm, uint8_t* ring, TcCtrl* ctl,
                                               TcRing& rg, int& tcount, int first = 0, int count = -1) {
  constexpr int PAD = (PC::KS - 1) / 2;
  int seq = 0;
  for (int u = blockIdx.x; u < P.n_units; u += gridDim.x)
    for (int t = 0; t < P.frames; ++t, ++seq) {
      if (seq < first) continue;
      if (count >= 0 && seq >= first + count) return;
      ++tcount;
      const int tx = u % cm.tiles_x;
      const int r = u / cm.tiles_x;
      const int ty = r % cm.tiles_y;
      const int img = (r / cm.tiles_y) * P.frames + t;
      load_tile<PC, NSPLIT, KC::NS, KC::SLOT_BYTES>(&P.tm_hi, &P.tm_lo, ring, &ctl->bars, rg, tx * 8 - PAD,
                                                    ty * 16 - PAD, img * P.img_mul + P.img_add);
      TC_TRACE(0, tcount);
    }
}

template <class PC, class KC, int NSPLIT>
__device__ __forceinline__ void mma_phase(const TcPhase& P, const TcCommon& cm, uint8_t* wsm, uint8_t* ring,
                                          TcCtrl* ctl, uint32_t tmem, TcRing& rg, int& it, int lane) {
  long long* tr = (cm.trace != nullptr && blockIdx.x == 0) ? cm.trace : nullptr;
  for (int u = blockIdx.x; u < P.n_units; u += gridDim.x)
    for (int t = 0; t < P.frames; ++t, ++it)
      mma_tile<PC, NSPLIT, KC::NS, KC::SLOT_BYTES, KC::TMEM_BUF_COLS, KC::CH_STRIDE>(wsm, ring, &ctl->bars, tmem, rg, it,
                                                                                  lane, tr, it);
}

template <class PC, class KC, int NSPLIT>
__device__ __forceinline__ void epilogue_phase(const TcPhase& P, const TcCommon& cm, TcCtrl* ctl, const float* bias_sm,
                                               uint32_t tmem, int& it, int warp, int lane) {
  long long* tr = (cm.trace != nullptr && blockIdx.x == 0) ? cm.trace : nullptr;
  TcEpiArgs E;
  E.epi = P.epi;
  E.accumulate = P.accumulate;
  E.f32_chunked = P.f32_chunked;
  E.coherent_pbase = 0;  // written by the same threads in the previous phase
  E.pbase = P.pbase;
  E.out_hi = P.out_hi;
  E.out_lo = P.out_lo;
  E.res_hi = P.res_hi;
  E.res_lo = P.res_lo;
  E.out_f32 = P.out_f32;
  E.H = cm.H;
  E.W = cm.W;
  E.trunc_comp = cm.trunc_comp;
  E.st_policy = 0;
  U256 pre[2];  // 16 fp32: partial sums (kept across the unit's frames) or previous fp32 content
#pragma unroll
  for (int j = 0; j < 8; ++j) pre[0].w[j] = pre[1].w[j] = 0u;
  TcNoHook nohook;
  for (int u = blockIdx.x; u < P.n_units; u += gridDim.x)
    for (int t = 0; t < P.frames; ++t, ++it) {
      const int tx = u % cm.tiles_x;
      const int r = u / cm.tiles_x;
      const int ty = r % cm.tiles_y;
      const int nimg = r / cm.tiles_y;
      epi_tile<PC, NSPLIT, KC::TMEM_BUF_COLS, KC::CH_STRIDE>(E, &ctl->bars, bias_sm, tmem, it, warp, lane,
                                                             nimg * P.frames + t, nimg, tx, ty, t == 0, pre, tr, nohook, it);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase 0 has shape P0, phases 1..nphases-1 have shape P1 (conv1 -> conv10; partial sums -> conv2;
// the 7 accumulating frame slices of convmerge1).
template <class P0, class P1, int NSPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
    conv_tc_kernel(const __grid_constant__ TcProgram prog, const TcCommon cm) {
  using KC = KernelCfg<P0, P1, NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* wsm = smem;                 // weight image of the running phase
  uint8_t* ring = wsm + KC::W_BYTES;   // [NS][SLOT_BYTES]
  TcCtrl* ctl = reinterpret_cast<TcCtrl*>(ring + KC::NS * KC::SLOT_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TcPhase& ph0 = prog.ph[0];

  // ---- prologue: touches only weights / biases (never written by any kernel); overlaps the previous
  //      kernel's tail under programmatic dependent launch
  if (tid == 0) {
    mbar_init(&ctl->wfull, 1);
    mbar_init(&ctl->wfree, 1);
    mbar_init(&ctl->stores_done, kTcEpiWarps);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&ctl->bars.full[i], 1);
      mbar_init(&ctl->bars.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->bars.tmem_full[i], 1);
      mbar_init(&ctl->bars.tmem_empty[i], kTcEpiWarps);
    }
    fence_mbar_init();
    fence_proxy_async();
    tma_prefetch_desc(&ph0.tm_hi);
    if (NSPLIT == 2) tma_prefetch_desc(&ph0.tm_lo);
    mbar_arrive_expect_tx(&ctl->wfull, KC::W0);  // this CTA receives the whole image
    if (cm.nphases > 1) {  // warm L2 with the next phase's weight image (first touch is a DRAM read)
      const int off = blockIdx.x * 16384;
      if (off < KC::W1)
        l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(prog.ph[1].wimg) + off,
                         (KC::W1 - off) < 16384 ? (KC::W1 - off) : 16384);
    }
  }
  for (int i = tid; i < kTcMaxPhases * 64; i += kTcThreads) {
    const int pi = i >> 6, c = i & 63;
    const float* bp = pi < cm.nphases ? prog.ph[pi].bias : nullptr;
    const int nout = pi == 0 ? P0::NOUT : P1::NOUT;
    ctl->bias[pi][c] = (bp != nullptr && c < nout) ? bp[c] : 0.f;
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, KC::TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 0) load_weights<KC::W0>(wsm, ph0.wimg, &ctl->wfull);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctl->tmem_base;
  pdl_launch_dependents();  // let the next kernel's CTAs start their prologue as SMs free up
  pdl_wait();               // activations written by the previous kernel are visible after this
  if (cm.trace != nullptr && tid == 0) cm.trace[192 + 2 * blockIdx.x] = globaltimer_ns();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      TcRing rg{0, 0};
      int tcount = 0;
      TC_TRACE(0, 0);
      producer_phase<P0, KC, NSPLIT>(ph0, cm, ring, ctl, rg, tcount);
      for (int pi = 1; pi < cm.nphases; ++pi) {
        const TcPhase& ph = prog.ph[pi];
        // swap the weight image once every MMA of the previous phase has completed
        tma_prefetch_desc(&ph.tm_hi);
        if (NSPLIT == 2) tma_prefetch_desc(&ph.tm_lo);
        if (pi + 1 < cm.nphases) {
          const int off = blockIdx.x * 16384;
          if (off < KC::W1)
            l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(prog.ph[pi + 1].wimg) + off,
                             (KC::W1 - off) < 16384 ? (KC::W1 - off) : 16384);
        }
        const bool early = !cm.phase1_reads_phase0;
        if (early) producer_phase<P1, KC, NSPLIT>(ph, cm, ring, ctl, rg, tcount, 0, 1);
        mbar_wait(&ctl->wfree, (pi - 1) & 1);
        TC_TRACE(0, 32 + pi);
        mbar_arrive_expect_tx(&ctl->wfull, KC::W1);
        load_weights<KC::W1>(wsm, ph.wimg, &ctl->wfull);  // needs only the MMAs to have drained
        TC_TRACE(0, 40 + pi);
        if (cm.phase1_reads_phase0) {
          // this phase's patches are the CTA's own outputs of the previous phase (conv10 <- conv1): the
          // generic-proxy stores must be visible to the TMA (async proxy) first
          mbar_wait(&ctl->stores_done, (pi - 1) & 1);
          fence_proxy_async_all();
        }
        producer_phase<P1, KC, NSPLIT>(ph, cm, ring, ctl, rg, tcount, early ? 1 : 0, -1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp) =====================
    TcRing rg{0, 0};
    int it = 0;
    mbar_wait(&ctl->wfull, 0);
    fence_after_sync();
    if (lane == 0) TC_TRACE(1, 0);
    mma_phase<P0, KC, NSPLIT>(ph0, cm, wsm, ring, ctl, tmem, rg, it, lane);
    for (int pi = 1; pi < cm.nphases; ++pi) {
      if (elect_one()) mma_commit(&ctl->wfree);  // arrives when every MMA issued so far has completed
      __syncwarp();
      mbar_wait(&ctl->wfull, pi & 1);
      fence_after_sync();
      if (lane == 0) TC_TRACE(1, 40 + pi);
      mma_phase<P1, KC, NSPLIT>(prog.ph[pi], cm, wsm, ring, ctl, tmem, rg, it, lane);
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    int it = 0;
    epilogue_phase<P0, KC, NSPLIT>(ph0, cm, ctl, ctl->bias[0], tmem, it, warp, lane);
    for (int pi = 1; pi < cm.nphases; ++pi) {
      if (cm.phase1_reads_phase0) {
        // this thread's stores -> visible device-wide and to the async proxy before the next phase reads them
        __threadfence();
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->stores_done);
      }
      epilogue_phase<P1, KC, NSPLIT>(prog.ph[pi], cm, ctl, ctl->bias[pi], tmem, it, warp, lane);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, KC::TMEM_COLS);
  if (cm.trace != nullptr && tid == 0) cm.trace[193 + 2 * blockIdx.x] = globaltimer_ns();
}

// ---- weight images ------------------------------------------------------------------------------------
// HWIO fp32 [taps][cin_total][cout] -> [plane][tap][NOUT rows][64 ci] fp16, rows K-major and
// pre-swizzled (128B) so a linear bulk copy lands them UMMA-ready.  Uses input channels
// ci_off..ci_off+63.  plane 1 (nsplit=2) holds (w - fp16(w)) * 2048.
__global__ void pack_tc_weights_kernel(const float* __restrict__ hwio, int taps, int cin_total, int ci_off, int cout,
                                       int nrows, int nsplit, __half* __restrict__ out, size_t tap_stride,
                                       size_t lo_plane_offset) {
  const int total = taps * nrows * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int ci = e & 63;
    const int row = (e >> 6) % nrows;
    const int tap = e / (64 * nrows);
    const float w = row < cout ? hwio[((long long)tap * cin_total + ci_off + ci) * cout + row] : 0.f;
    const __half hi = __float2half_rn(w);
    const uint32_t off = sw128_offset(row, ci >> 3) + (ci & 7) * 2;
    uint8_t* base = reinterpret_cast<uint8_t*>(out) + (size_t)tap * tap_stride;
    *reinterpret_cast<__half*>(base + off) = hi;
    if (nsplit == 2) {
      const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
      *reinterpret_cast<__half*>(base + lo_plane_offset + off) = lo;
    }
  }
}

// fp32 NHWC [images*H*W, 64] <-> fp16 planes (channel-chunk-major); hw = H*W
__device__ __forceinline__ long long plane_index(long long e, long long hw) {
  const int c = (int)(e & 63);
  const long long pix = e >> 6, img = pix / hw, p = pix - img * hw;
  return ((img * 8 + (c >> 3)) * hw + p) * 8 + (c & 7);
}
__global__ void f32_to_planes_kernel(const float* __restrict__ in, long long n, long long hw, int nsplit,
                                     __half* __restrict__ hi, __half* __restrict__ lo) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const float v = in[e];
    const long long o = plane_index(e, hw);
    if (nsplit == 2) {
      __half h, l;
      split_half(v, h, l);
      hi[o] = h;
      lo[o] = l;
    } else {
      hi[o] = __float2half_rn(v);
    }
  }
}
__global__ void planes_to_f32_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, long long n,
                                     long long hw, int nsplit, float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long o = plane_index(e, hw);
    float v = __half2float(hi[o]);
    if (nsplit == 2) v = fmaf(__half2float(lo[o]), 1.f / 2048.f, v);
    out[e] = v;
  }
}

// conv0 (5x5, 3->64, leaky_relu; model/pfnl.py:48,61-62) writing fp16 planes.  K = 75 is too
// small/odd for the MMA path (0.2 % of the FLOPs): CUDA cores.  Tile = 16x16 pixels of one frame;
// thread = 4 consecutive pixels x 16 output channels (one smem read per 8 FMAs; the four lanes that
// share a pixel quad write one contiguous 128-byte pixel row of the output plane).
template <int NSPLIT>
__global__ void __launch_bounds__(256, 2) conv0_planes_kernel(const float* __restrict__ inp21, int H, int W,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  __shared__ __align__(16) float wsm[75 * 64];
  constexpr int PP = 61;  // patch row pitch (floats): keeps the 8 pixel quads of a warp on 8 different banks
  __shared__ float patch[20 * PP];
  const int tid = threadIdx.x;
  const int tiles_x = ceil_div(W, 16);
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
  const int img = blockIdx.y;
  const int n = img / kFrames, t = img % kFrames;
  const int y0 = ty * 16, x0 = tx * 16;
  // weights permuted to [k][j4][g][4]: the 4 channel groups of a warp read 4 consecutive 16-byte chunks
  for (int i = tid; i < 75 * 64; i += 256) {
    const int kk = i >> 6, co = i & 63;
    wsm[kk * 64 + ((co >> 2) & 3) * 16 + (co >> 4) * 4 + (co & 3)] = w[i];
  }
  for (int i = tid; i < 20 * 20 * 3; i += 256) {
    int c = i % 3, pp = i / 3;
    int py = pp / 20, px = pp % 20;
    int gy = y0 + py - 2, gx = x0 + px - 2;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = inp21[(((long long)n * H + gy) * W + gx) * 21 + t * 3 + c];
    patch[py * PP + px * 3 + c] = v;
  }
  __syncthreads();
  const int g = tid & 3;        // output channels g*16 .. g*16+15
  const int quad = tid >> 2;    // 64 pixel quads: row = quad / 4, columns (quad % 4)*4 .. +3
  const int py = quad >> 2, px0 = (quad & 3) * 4;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll 1
  for (int dy = 0; dy < 5; ++dy) {
#pragma unroll 1
    for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pr = patch + (py + dy) * PP + (px0 + dx) * 3 + c;
        const float a0 = pr[0], a1 = pr[3], a2 = pr[6], a3 = pr[9];
        const float4* wr = reinterpret_cast<const float4*>(wsm + ((dy * 5 + dx) * 3 + c) * 64 + g * 4);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = wr[j4 * 4];
          acc[0][4 * j4 + 0] = fmaf(a0, b.x, acc[0][4 * j4 + 0]);
          acc[0][4 * j4 + 1] = fmaf(a0, b.y, acc[0][4 * j4 + 1]);
          acc[0][4 * j4 + 2] = fmaf(a0, b.z, acc[0][4 * j4 + 2]);
          acc[0][4 * j4 + 3] = fmaf(a0, b.w, acc[0][4 * j4 + 3]);
          acc[1][4 * j4 + 0] = fmaf(a1, b.x, acc[1][4 * j4 + 0]);
          acc[1][4 * j4 + 1] = fmaf(a1, b.y, acc[1][4 * j4 + 1]);
          acc[1][4 * j4 + 2] = fmaf(a1, b.z, acc[1][4 * j4 + 2]);
          acc[1][4 * j4 + 3] = fmaf(a1, b.w, acc[1][4 * j4 + 3]);
          acc[2][4 * j4 + 0] = fmaf(a2, b.x, acc[2][4 * j4 + 0]);
          acc[2][4 * j4 + 1] = fmaf(a2, b.y, acc[2][4 * j4 + 1]);
          acc[2][4 * j4 + 2] = fmaf(a2, b.z, acc[2][4 * j4 + 2]);
          acc[2][4 * j4 + 3] = fmaf(a2, b.w, acc[2][4 * j4 + 3]);
          acc[3][4 * j4 + 0] = fmaf(a3, b.x, acc[3][4 * j4 + 0]);
          acc[3][4 * j4 + 1] = fmaf(a3, b.y, acc[3][4 * j4 + 1]);
          acc[3][4 * j4 + 2] = fmaf(a3, b.z, acc[3][4 * j4 + 2]);
          acc[3][4 * j4 + 3] = fmaf(a3, b.w, acc[3][4 * j4 + 3]);
        }
      }
    }
  }
  const int gy = y0 + py;
  if (gy < H) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gx = x0 + px0 + i;
      if (gx >= W) continue;
      const long long poff = plane_off(img, g * 2, gy, gx, H, W);
      const long long cs = (long long)H * W * 8;
      U256 oh, ol;
      __half* ph = reinterpret_cast<__half*>(&oh);
      __half* pl = reinterpret_cast<__half*>(&ol);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = lrelu(acc[i][j] + bias[g * 16 + j]);
        if (NSPLIT == 2)
          split_half(v, ph[j], pl[j]);
        else
          ph[j] = __float2half_rn(v);
      }
      st_plane16(out_hi + poff, cs, oh);
      if (NSPLIT == 2) st_plane16(out_lo + poff, cs, ol);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// conv0 (5x5, 3->64, leaky_relu; model/pfnl.py:48,61-62) on tcgen05: an explicit im2col tile.  K = 75 taps
// (dy,dx,c) padded to 80 = 5 k-steps; one CTA (128 threads = 128 tile pixels = 128 TMEM lanes) per 16x8
// tile: stage the 20x12x3 fp32 patch, let every thread write its pixel's 80 operand values as hi/lo fp16
// core-matrix rows (un-swizzled K-major: [k-chunk][row][8 k]), 5 x (N=128 + N=64) MMAs, epilogue as the
// other convs.  ~65 KB of smem and 128 TMEM columns per CTA: three persistent CTAs share an SM (each keeps
// its 20 KB weight image for all of its tiles) and overlap each other's phases, which replaces an
// intra-CTA pipeline.
template <int NSPLIT>
struct Conv0Cfg {
  static constexpr int KC = 10;                       // k-chunks of 8
  static constexpr int A_BYTES = KC * 128 * 16;       // one plane of the im2col tile
  static constexpr int B_ROWS = NSPLIT * 64;
  static constexpr int B_BYTES = KC * B_ROWS * 16;
  static constexpr int PATCH_FLOATS = 20 * 12 * 3;
  static constexpr int SMEM = 1024 + NSPLIT * A_BYTES + B_BYTES + PATCH_FLOATS * 4 + 64 * 4 + 64;
};

// HWIO fp32 [75][64] -> [kc][plane*64 + co][8 k] fp16 (k = (dy*5+dx)*3 + c, zero for k >= 75)
__global__ void pack_conv0_tc_kernel(const float* __restrict__ hwio, int nsplit, __half* __restrict__ out) {
  const int rows = nsplit * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 10 * 64 * 8; e += gridDim.x * blockDim.x) {
    const int kk = e & 7, co = (e >> 3) & 63, kc = e >> 9;
    const int k = kc * 8 + kk;
    const float w = k < 75 ? hwio[k * 64 + co] : 0.f;
    const __half hi = __float2half_rn(w);
    out[((size_t)kc * rows + co) * 8 + kk] = hi;
    if (nsplit == 2) out[((size_t)kc * rows + 64 + co) * 8 + kk] = __float2half_rn((w - __half2float(hi)) * 2048.f);
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(128) conv0_tc_kernel(const float* __restrict__ inp21, int H, int W, int tiles_x,
                                                       int tiles_y, int ntiles, const __half* __restrict__ wimg,
                                                       const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                       __half* __restrict__ out_lo, float trunc_comp) {
  using CF = Conv0Cfg<NSPLIT>;
  const float comp = 1.f + trunc_comp * 6.f * 1.1920929e-7f;  // chain of 5 accumulating MMAs (conv_tc_dev.cuh)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;                                  // [NSPLIT][KC][128 rows][16 B]
  uint8_t* b_sm = a_sm + NSPLIT * CF::A_BYTES;           // [KC][B_ROWS][16 B]
  float* patch = reinterpret_cast<float*>(b_sm + CF::B_BYTES);  // [20][12][3]
  float* bias_sm = patch + CF::PATCH_FLOATS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(bias_sm + 64);   // [0]: weights landed, [1]: MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
    fence_proxy_async();
    mbar_arrive_expect_tx(&bar[0], CF::B_BYTES);
    bulk_load(b_sm, wimg, CF::B_BYTES, &bar[0]);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  if (tid < 64) bias_sm[tid] = bias[tid];
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // inp21 is written by the previous kernel
  const int m = tid, my = m >> 3, mx = m & 7;  // this thread's pixel = tile row m = TMEM lane
  uint32_t done_par = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, done_par ^= 1u) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, img = tile / (tiles_x * tiles_y);
    const int n = img / kFrames, t = img % kFrames;
    const int y0 = ty * 16, x0 = tx * 8;
    {  // all loads in flight before the first store: the patch is the latency of this kernel
      constexpr int NL = (CF::PATCH_FLOATS + 127) / 128;
      float pv[NL];
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 128;
        const int c = i % 3, pp = i / 3;
        const int py = pp / 12, px = pp % 12;
        const int gy = y0 + py - 2, gx = x0 + px - 2;
        pv[u] = 0.f;
        if (i < CF::PATCH_FLOATS && gy >= 0 && gy < H && gx >= 0 && gx < W)
          pv[u] = inp21[(((long long)n * H + gy) * W + gx) * 21 + t * 3 + c];
      }
#pragma unroll
      for (int u = 0; u < NL; ++u) {
        const int i = tid + u * 128;
        if (i < CF::PATCH_FLOATS) patch[i] = pv[u];
      }
    }
    __syncthreads();
    // ---- im2col: this thread's 80 operand values, 8 per 16-byte core-matrix row
#pragma unroll
    for (int kc = 0; kc < CF::KC; ++kc) {
      __align__(16) __half hh[8];
      __align__(16) __half hl[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int k = kc * 8 + kk;
        float v = 0.f;
        if (k < 75) {
          const int tap = k / 3, c = k - tap * 3;
          const int dy = tap / 5, dx = tap - dy * 5;
          v = patch[((my + dy) * 12 + mx + dx) * 3 + c];
        }
        if (NSPLIT == 2)
          split_half(v, hh[kk], hl[kk]);
        else
          hh[kk] = __float2half_rn(v);
      }
      *reinterpret_cast<uint4*>(a_sm + (kc * 128 + m) * 16) = *reinterpret_cast<const uint4*>(hh);
      if (NSPLIT == 2)
        *reinterpret_cast<uint4*>(a_sm + CF::A_BYTES + (kc * 128 + m) * 16) = *reinterpret_cast<const uint4*>(hl);
    }
    fence_proxy_async();  // generic-proxy operand writes -> visible to the tensor core (async proxy)
    fence_before_sync();  // the previous tile's tcgen05.ld are ordered before this tile's MMAs
    __syncthreads();
    if (warp == 0) {
      fence_after_sync();
      mbar_wait(&bar[0], 0);  // the weight image (first tile: may still be landing)
      if (elect_one()) {
        constexpr uint32_t idesc_hi = make_idesc_f16(128, NSPLIT * 64);
        constexpr uint32_t idesc_lo = make_idesc_f16(128, 64);
        // K-major, no swizzle: 8-row groups 128 B apart (SBO), the two k-chunks of a K=16 step LBO apart
        const uint64_t ad = make_sdesc_interleave(smem_u32(a_sm), 128 * 16, 128);
        const uint64_t bd = make_sdesc_interleave(smem_u32(b_sm), CF::B_ROWS * 16, 128);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
          mma_f16(tmem, ad + ((ks * 2 * 128 * 16) >> 4), bd + ((ks * 2 * CF::B_ROWS * 16) >> 4), idesc_hi, ks > 0);
        if (NSPLIT == 2) {
          const uint64_t al = make_sdesc_interleave(smem_u32(a_sm + CF::A_BYTES), 128 * 16, 128);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks)
            mma_f16(tmem + 64, al + ((ks * 2 * 128 * 16) >> 4), bd + ((ks * 2 * CF::B_ROWS * 16) >> 4), idesc_lo, 1u);
        }
        mma_commit(&bar[1]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[1], done_par);
    fence_after_sync();
    // ---- epilogue: thread = pixel (TMEM lane), 4 passes of 16 channels
    const int y = y0 + my, x = x0 + mx;
    const bool inb = y < H && x < W;
    const long long cs = (long long)H * W * 8;
    const uint32_t t0 = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t d0[16], d1[16];
      tmem_ld_32x32b_x16(t0 + c0, d0);
      if (NSPLIT == 2) tmem_ld_32x32b_x16(t0 + 64 + c0, d1);
      tmem_ld_wait();
      U256 oh, ol;
      __half* ph = reinterpret_cast<__half*>(&oh);
      __half* pl = reinterpret_cast<__half*>(&ol);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v = __fmul_rn(__uint_as_float(d0[j]), comp);
        if (NSPLIT == 2) v = fmaf(__uint_as_float(d1[j]), 1.f / 2048.f, v);
        v = lrelu(v + bias_sm[c0 + j]);
        if (NSPLIT == 2)
          split_half(v, ph[j], pl[j]);
        else
          ph[j] = __float2half_rn(v);
      }
      if (inb) {
        const long long poff = plane_off(img, c0 >> 3, y, x, H, W);
        st_plane16(out_hi + poff, cs, oh);
        if (NSPLIT == 2) st_plane16(out_lo + poff, cs, ol);
      }
    }
  }
  pdl_launch_dependents();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- host side --------------------------------------------------------------------------------------------
namespace {

// programmatic dependent launch: per handle (pfnl_profile turns it off so that per-kernel event intervals do not
// overlap), PFNL_TC_NO_PDL=1 turns it off for the process
bool pdl_enabled(const TcWeights& tw) {
  static const bool env_off = getenv("PFNL_TC_NO_PDL") != nullptr;
  return tw.pdl && !env_off;
}

// Fills the source-side fields of a phase (tensor maps over fp16 planes [src_images,H,W,64]).
template <class PC>
int phase_sources(TcPhase& ph, const void* src_hi, const void* src_lo, int src_images, int H, int W) {
  int r = make_act_tmap(&ph.tm_hi, src_hi, src_images, H, W, PC::BOX_W, PC::BOX_H);
  if (r == 0) r = make_act_tmap(&ph.tm_lo, src_lo ? src_lo : src_hi, src_images, H, W, PC::BOX_W, PC::BOX_H);
  if (r != 0) {
    set_error("cuTensorMapEncodeTiled failed (%d) for images=%d H=%d W=%d", r, src_images, H, W);
    return PFNL_ERR_CUDA;
  }
  return PFNL_OK;
}

// Launches one persistent kernel running phase a and (optionally) phase b on the same units.
template <class P0, class P1, int NSPLIT>
int launch_tc(const TcProgram& prog, int nphases, bool phase1_reads_phase0, int H, int W, const TcWeights& tw,
              cudaStream_t s) {
  const int num_sms = tw.num_sms;
  const TcPhase& a = prog.ph[0];
  using KC = KernelCfg<P0, P1, NSPLIT>;
  TcCommon cm;
  memset(&cm, 0, sizeof(cm));
  cm.H = H;
  cm.W = W;
  cm.tiles_x = ceil_div(W, 8);
  cm.tiles_y = ceil_div(H, 16);
  cm.nphases = nphases;
  cm.phase1_reads_phase0 = (nphases > 1 && phase1_reads_phase0) ? 1 : 0;
  cm.trunc_comp = tw.trunc_comp;
  if (a.n_units <= 0) return PFNL_OK;
  for (int i = 1; i < nphases; ++i)
    if (prog.ph[i].n_units != a.n_units) {
      set_error("launch_tc: phases must cover the same work units (%d vs %d)", a.n_units, prog.ph[i].n_units);
      return PFNL_ERR_BAD_ARG;
    }
  int grid = a.n_units < num_sms ? a.n_units : num_sms;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = KC::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled(tw)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  static long long* trace_dev = nullptr;
  static const bool tracing = getenv("PFNL_TC_TRACE") != nullptr;
  if (tracing) {
    if (!trace_dev) PFNL_CUDA(cudaMalloc((void**)&trace_dev, (192 + 2 * 256) * sizeof(long long)));
    PFNL_CUDA(cudaMemsetAsync(trace_dev, 0, (192 + 2 * 256) * sizeof(long long), s));
    cm.trace = trace_dev;
  }
  const TcPhase& bb = prog.ph[nphases > 1 ? 1 : 0];
  PFNL_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<P0, P1, NSPLIT>, prog, cm));
  if (tracing) {
    long long t[192 + 2 * 256];
    PFNL_CUDA(cudaStreamSynchronize(s));
    PFNL_CUDA(cudaMemcpy(t, trace_dev, sizeof(t), cudaMemcpyDeviceToHost));
    const long long t0 = t[0];
    fprintf(stderr,
            "[tc-trace] phases=%d P0=(KS%d,NSRC%d,N%d,ch%d,epi%d,frames%d) P1=(KS%d,NSRC%d,epi%d,frames%d) NSPLIT=%d "
            "units=%d grid=%d (cycles since producer start)\n",
            cm.nphases, P0::KS, P0::NSRC, P0::NOUT, P0::NCH, a.epi, a.frames, P1::KS, P1::NSRC, bb.epi, bb.frames, NSPLIT,
            a.n_units, grid);
    fprintf(stderr, "  producer issue done :");
    for (int i = 1; i < 14 && t[i]; ++i) fprintf(stderr, " %lld", t[i] - t0);
    fprintf(stderr, "\n  mma: weights ready %lld ; per tile (data ready, issue done):", t[64] - t0);
    for (int i = 0; i < 12 && t[64 + 1 + 2 * i]; ++i)
      fprintf(stderr, " (%lld,%lld)", t[64 + 1 + 2 * i] - t0, t[64 + 2 + 2 * i] - t0);
    fprintf(stderr, "\n  swap to phase 1: MMAs drained %lld, image issued %lld, image landed %lld", t[33] - t0, t[41] - t0,
            t[64 + 41] - t0);
    fprintf(stderr, "\n  epilogue per tile (acc ready, done):");
    for (int i = 0; i < 12 && t[128 + 2 * i]; ++i)
      fprintf(stderr, " (%lld,%lld)", t[128 + 2 * i] - t0, t[128 + 1 + 2 * i] - t0);
    fprintf(stderr, "\n");
    // per-CTA wall clock (globaltimer, ns): when each CTA passed the dependency wait and when it finished
    long long smin = 0, smax = 0, emin = 0, emax = 0, dmin = 0, dmax = 0, dsum = 0;
    for (int b = 0; b < grid && b < 256; ++b) {
      const long long st = t[192 + 2 * b], en = t[193 + 2 * b], d = en - st;
      if (b == 0 || st < smin) smin = st;
      if (b == 0 || st > smax) smax = st;
      if (b == 0 || en < emin) emin = en;
      if (b == 0 || en > emax) emax = en;
      if (b == 0 || d < dmin) dmin = d;
      if (b == 0 || d > dmax) dmax = d;
      dsum += d;
    }
    fprintf(stderr,
            "  per-CTA: start spread %.2f us, end spread %.2f us, CTA duration min/avg/max %.2f/%.2f/%.2f us, "
            "kernel span (first start -> last end) %.2f us\n",
            (smax - smin) / 1e3, (emax - emin) / 1e3, dmin / 1e3, dsum / 1e3 / (grid < 256 ? grid : 256), dmax / 1e3,
            (emax - smin) / 1e3);
  }
  return PFNL_OK;
}

template <class P0, class P1, int NSPLIT>
int set_attr() {
  PFNL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<P0, P1, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 KernelCfg<P0, P1, NSPLIT>::SMEM_BYTES));
  return PFNL_OK;
}

// phase shapes used by the PFRB stack; the split mode uses 2 accumulation chains for 3x3 kernels
template <int NSPLIT>
struct Shapes {
  static constexpr int CH3 = NSPLIT == 2 ? 2 : 1;
  using C3 = PhaseCfg<3, 1, 64, CH3>;   // conv1, conv2 halves
  using C10 = PhaseCfg<1, 7, 64, 1>;    // conv10 over the 7 frame slices
  using CM = PhaseCfg<3, 1, 48, CH3>;   // convmerge1 slice
};

size_t plane_bytes(int images, int H, int W) { return (size_t)images * H * W * 64 * sizeof(__half); }

int pack_weights(const float* hwio, int taps, int cin_total, int ci_off, int cout, int nrows, int nsplit, void** out,
                 std::vector<void*>& allocs) {
  void* d = nullptr;
  PFNL_CUDA(cudaMalloc(&d, (size_t)nsplit * taps * nrows * 128));
  allocs.push_back(d);
  const int total = taps * nrows * 64;
  pack_tc_weights_kernel<<<ceil_div(total, 256), 256>>>(hwio, taps, cin_total, ci_off, cout, nrows, nsplit,
                                                       (__half*)d, (size_t)nsplit * nrows * 128,
                                                       (size_t)nrows * 128);
  PFNL_LAUNCH_CHECK();
  *out = d;
  return PFNL_OK;
}

}  // namespace

void tc_carve(TcWorkspace& w, int precision, int N, int H, int W, const std::function<char*(size_t)>& take) {
  const int nsplit = tc_nsplit(precision);
  for (int pl = 0; pl < 2; ++pl) {
    w.actA[pl] = w.actB[pl] = w.base[pl] = nullptr;
  }
  for (int pl = 0; pl < nsplit; ++pl) {
    w.actA[pl] = take(plane_bytes(N * kFrames, H, W));
    w.actB[pl] = take(plane_bytes(N * kFrames, H, W));
    w.base[pl] = take(plane_bytes(N, H, W));
  }
  w.pbase = (float*)take((size_t)N * H * W * 64 * sizeof(float));
  w.nl_x16 = nullptr;
  w.nl_priv = nullptr;
  w.flow_flags = nullptr;
  w.flow_fault = nullptr;
  if (tc_nl_on_tensor_cores(precision) && tc_has_nonlocal()) {
    const int L = (H / 2) * (W / 2);
    w.nl_x16 = take(tc_nl_workspace_bytes(N, L));
    w.nl_priv = take((size_t)N * L * kNL * sizeof(float));  // Y = softmax(S) * G, fp32 [N,L,84]
  }
}

int tc_init(TcWeights& tw, int precision, const TcRawWeights& raw, std::vector<void*>& allocs) {
  if (!get_encode_tiled()) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PFNL_ERR_CUDA;
  }
  cudaDeviceProp prop;
  int dev = 0;
  PFNL_CUDA(cudaGetDevice(&dev));
  PFNL_CUDA(cudaGetDeviceProperties(&prop, dev));
  tw.num_sms = prop.multiProcessorCount;
  PFNL_CUDA(tc_apply_wait_limit_from_env());
  tw.flow = tc_flow_default();
  tw.trunc_comp = getenv("PFNL_TC_TRUNC_COMP") != nullptr ? (float)atof(getenv("PFNL_TC_TRUNC_COMP")) : kTcTruncCompDefault;
  tw.precision = precision;
  tw.nsplit = tc_nsplit(precision);
  tw.raw = raw;
  int rc;
  if ((rc = set_attr<Shapes<1>::C3, Shapes<1>::C10, 1>())) return rc;
  if ((rc = set_attr<Shapes<1>::C3, Shapes<1>::C3, 1>())) return rc;
  if ((rc = set_attr<Shapes<1>::CM, Shapes<1>::CM, 1>())) return rc;
  if ((rc = set_attr<Shapes<2>::C3, Shapes<2>::C10, 2>())) return rc;
  if ((rc = set_attr<Shapes<2>::C3, Shapes<2>::C3, 2>())) return rc;
  if ((rc = set_attr<Shapes<2>::CM, Shapes<2>::CM, 2>())) return rc;
  if ((rc = tc_nl_init())) return rc;
  if ((rc = tc_flow_init())) return rc;
  PFNL_CUDA(cudaFuncSetAttribute(conv0_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv0Cfg<1>::SMEM));
  PFNL_CUDA(cudaFuncSetAttribute(conv0_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Conv0Cfg<2>::SMEM));
  const int ns = tw.nsplit;
  {
    void* d = nullptr;
    PFNL_CUDA(cudaMalloc(&d, (size_t)10 * ns * 64 * 16));
    allocs.push_back(d);
    tw.conv0 = d;
    pack_conv0_tc_kernel<<<20, 256>>>(raw.conv0_w, ns, (__half*)d);
    PFNL_LAUNCH_CHECK();
  }
  for (int i = 0; i < PFNL_NUM_BLOCK; ++i) {
    if ((rc = pack_weights(raw.conv1_w[i], 9, 64, 0, 64, 64, ns, &tw.conv1[i], allocs))) return rc;
    // conv10 [1,1,448,64]: "tap" t = frame slice t (input channels t*64..t*64+63)
    void* d = nullptr;
    PFNL_CUDA(cudaMalloc(&d, (size_t)ns * 7 * 64 * 128));
    allocs.push_back(d);
    tw.conv10[i] = d;
    for (int t = 0; t < 7; ++t) {
      // slice t -> [t][hi ; lo]
      pack_tc_weights_kernel<<<16, 256>>>(raw.conv10_w[i], 1, 448, t * 64, 64, 64, ns,
                                          (__half*)((uint8_t*)d + (size_t)t * ns * 8192), (size_t)ns * 8192,
                                          (size_t)8192);
      PFNL_LAUNCH_CHECK();
    }
    if ((rc = pack_weights(raw.conv2_w[i], 9, 128, 0, 64, 64, ns, &tw.conv2b[i], allocs))) return rc;
    if ((rc = pack_weights(raw.conv2_w[i], 9, 128, 64, 64, 64, ns, &tw.conv2f[i], allocs))) return rc;
  }
  // convmerge1 [3,3,448,48]: one weight image per frame slice t (launched 7 times, accumulating)
  {
    void* d = nullptr;
    const size_t per = (size_t)ns * 9 * 48 * 128;
    PFNL_CUDA(cudaMalloc(&d, per * 7));
    allocs.push_back(d);
    tw.merge1 = d;
    for (int t = 0; t < 7; ++t) {
      pack_tc_weights_kernel<<<ceil_div(9 * 48 * 64, 256), 256>>>(raw.merge1_w, 9, 448, t * 64, 48, 48, ns,
                                                                  (__half*)((uint8_t*)d + per * t),
                                                                  (size_t)ns * 48 * 128, (size_t)48 * 128);
      PFNL_LAUNCH_CHECK();
    }
  }
  PFNL_CUDA(cudaDeviceSynchronize());
  return PFNL_OK;
}

void tc_destroy(TcWeights& tw) {
  if (tw.nl_scratch) cudaFree(tw.nl_scratch);
  tw.nl_scratch = nullptr;
  tw.nl_scratch_cap = 0;
}


namespace {

// One PFRB (model/pfnl.py:66-71) on fp16 planes, in place on actA: two persistent launches.
template <int NSPLIT>
int pfrb_tc(const TcWeights& tw, TcWorkspace& w, int i, int N, int H, int W, cudaStream_t s, long long* launches,
            Profiler* prof) {
  using SH = Shapes<NSPLIT>;
  const int tiles = ceil_div(W, 8) * ceil_div(H, 16);
  const int units = N * tiles;
  int rc;
  TcProgram prog;
  TcPhase& a = prog.ph[0];
  TcPhase& b = prog.ph[1];
  // ---- launch A: inp1[t] = conv1_i(inp0[t])  (pfnl.py:66)  ->  base = conv10_i(concat_t inp1[t])  (pfnl.py:67-68)
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  if ((rc = phase_sources<typename SH::C3>(a, w.actA[0], w.actA[1], N * kFrames, H, W))) return rc;
  a.wimg = (const __half*)tw.conv1[i];
  a.frames = kFrames;
  a.n_units = units;
  a.img_mul = 1;
  a.epi = kEpiActPlanes;
  a.bias = tw.raw.conv1_b[i];
  a.out_hi = (__half*)w.actB[0];
  a.out_lo = (__half*)w.actB[1];
  // conv10 reads the 7 conv1 outputs of its own unit as 7 K-stages (no concat copy)
  if ((rc = phase_sources<typename SH::C10>(b, w.actB[0], w.actB[1], N * kFrames, H, W))) return rc;
  b.wimg = (const __half*)tw.conv10[i];
  b.frames = 1;
  b.n_units = units;
  b.img_mul = kFrames;
  b.epi = kEpiActPlanes;
  b.bias = tw.raw.conv10_b[i];
  b.out_hi = (__half*)w.base[0];
  b.out_lo = (__half*)w.base[1];
  if (prof) prof->begin(kProfConv1, s);
  rc = launch_tc<typename SH::C3, typename SH::C10, NSPLIT>(prog, 2, true, H, W, tw, s);
  if (prof) prof->end(s);
  if (rc) return rc;
  // ---- launch B: conv2_i(concat[base, inp1[t]]) = conv(base; W2[:,:,0:64]) + conv(inp1[t]; W2[:,:,64:128]):
  //      the base half is identical for the 7 frames -> computed once per unit (fp32 partial sums), then
  //      inp0[t] += leaky_relu(partial + conv(inp1[t]) + bias)                          (pfnl.py:69-71)
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  if ((rc = phase_sources<typename SH::C3>(a, w.base[0], w.base[1], N, H, W))) return rc;
  a.wimg = (const __half*)tw.conv2b[i];
  a.frames = 1;
  a.n_units = units;
  a.img_mul = 1;
  a.epi = kEpiPartialF32;
  a.f32_chunked = 1;
  a.out_f32 = w.pbase;
  if ((rc = phase_sources<typename SH::C3>(b, w.actB[0], w.actB[1], N * kFrames, H, W))) return rc;
  b.wimg = (const __half*)tw.conv2f[i];
  b.frames = kFrames;
  b.n_units = units;
  b.img_mul = 1;
  b.epi = kEpiResPlanes;
  b.bias = tw.raw.conv2_b[i];
  b.pbase = w.pbase;
  b.res_hi = (const __half*)w.actA[0];
  b.res_lo = (const __half*)w.actA[1];
  b.out_hi = (__half*)w.actA[0];
  b.out_lo = (__half*)w.actA[1];
  if (prof) prof->begin(kProfConv2, s);
  rc = launch_tc<typename SH::C3, typename SH::C3, NSPLIT>(prog, 2, false, H, W, tw, s);
  if (prof) prof->end(s);
  if (rc) return rc;
  *launches += 2;
  return PFNL_OK;
}

// conv0 on each frame (model/pfnl.py:61-62): inp21 [N,H,W,21] fp32 -> fp16 planes of inp0 (actA)
template <int NSPLIT>
int conv0_tc(const TcWeights& tw, TcWorkspace& w, const float* inp21, int N, int H, int W, cudaStream_t s) {
  static const bool conv0_ffma = getenv("PFNL_TC_CONV0_FFMA") != nullptr;  // the CUDA-core version, kept for A/B runs
  if (conv0_ffma) {
    dim3 grid(ceil_div(W, 16) * ceil_div(H, 16), N * kFrames);
    conv0_planes_kernel<NSPLIT><<<grid, 256, 0, s>>>(inp21, H, W, tw.raw.conv0_w, tw.raw.conv0_b, (__half*)w.actA[0],
                                                     (__half*)w.actA[1]);
  } else {
    const int tiles_x = ceil_div(W, 8), tiles_y = ceil_div(H, 16);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int ntiles = tiles_x * tiles_y * N * kFrames;
    cfg.gridDim = dim3(ntiles < 3 * tw.num_sms ? ntiles : 3 * tw.num_sms);  // 3 CTAs per SM, each keeps its weights
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = Conv0Cfg<NSPLIT>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl_enabled(tw)) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    PFNL_CUDA(cudaLaunchKernelEx(&cfg, conv0_tc_kernel<NSPLIT>, inp21, H, W, tiles_x, tiles_y, ntiles,
                                 (const __half*)tw.conv0, tw.raw.conv0_b, (__half*)w.actA[0], (__half*)w.actA[1],
                                 tw.trunc_comp));
  }
  PFNL_LAUNCH_CHECK();
  return PFNL_OK;
}

// merge = convmerge1(concat_t inp0[t])  (model/pfnl.py:73-74) from the fp16 planes in actA: ONE launch, 7 phases -
// frame slice t of the K = 7*576 contraction per phase, fp32 partial sums accumulated in place by the same threads
template <int NSPLIT>
int merge1_tc(const TcWeights& tw, TcWorkspace& w, int N, int H, int W, float* merge, cudaStream_t s) {
  using SH = Shapes<NSPLIT>;
  int rc;
  const int units = N * ceil_div(W, 8) * ceil_div(H, 16);
  const size_t per = (size_t)NSPLIT * 9 * 48 * 128;
  TcProgram prog;
  for (int t = 0; t < kFrames; ++t) {
    TcPhase& a = prog.ph[t];
    memset(&a, 0, sizeof(a));
    if ((rc = phase_sources<typename SH::CM>(a, w.actA[0], w.actA[1], N * kFrames, H, W))) return rc;
    a.wimg = (const __half*)((const uint8_t*)tw.merge1 + per * t);
    a.frames = 1;
    a.n_units = units;
    a.img_mul = kFrames;
    a.img_add = t;
    a.epi = t == kFrames - 1 ? kEpiFinalF32 : kEpiPartialF32;
    a.accumulate = t > 0;
    a.bias = t == kFrames - 1 ? tw.raw.merge1_b : nullptr;
    a.out_f32 = merge;
  }
  return launch_tc<typename SH::CM, typename SH::CM, NSPLIT>(prog, kFrames, false, H, W, tw, s);
}

template <int NSPLIT>
int trunk_tc(const TcWeights& tw, TcWorkspace& w, const float* inp21, int N, int H, int W, float* merge,
             cudaStream_t s, long long* launches, Profiler* prof) {
  int rc;
  if (prof) prof->begin(kProfConv0, s);
  rc = conv0_tc<NSPLIT>(tw, w, inp21, N, H, W, s);
  if (prof) prof->end(s);
  if (rc) return rc;
  *launches += 1;
  if (tw.flow) {
    // the 20 blocks as one persistent dataflow kernel, in place on actA
    if (prof) prof->begin(kProfPfrbFlow, s);
    rc = tc_pfrb_flow(tw, w, 0, PFNL_NUM_BLOCK, N, H, W, pdl_enabled(tw), s);
    if (prof) prof->end(s);
    if (rc) return rc;
    *launches += 1;
  } else {
    for (int i = 0; i < PFNL_NUM_BLOCK; ++i)
      if ((rc = pfrb_tc<NSPLIT>(tw, w, i, N, H, W, s, launches, prof))) return rc;
  }
  if (prof) prof->begin(kProfMerge1, s);
  rc = merge1_tc<NSPLIT>(tw, w, N, H, W, merge, s);
  if (prof) prof->end(s);
  if (rc) return rc;
  *launches += 1;
  return PFNL_OK;
}

}  // namespace

int tc_trunk(const TcWeights& tw, TcWorkspace& w, int precision, const float* inp21, int N, int H, int W,
             float* merge, cudaStream_t s, long long* launches, Profiler* prof) {
  if (tc_nsplit(precision) == 2) return trunk_tc<2>(tw, w, inp21, N, H, W, merge, s, launches, prof);
  return trunk_tc<1>(tw, w, inp21, N, H, W, merge, s, launches, prof);
}

int tc_pfrb_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, int blk, const float* frames, int N, int H,
                   int W, float* frames_out, cudaStream_t s, long long* launches) {
  const int ns = tc_nsplit(precision);
  const long long n = (long long)N * kFrames * H * W * 64;
  f32_to_planes_kernel<<<148 * 8, 256, 0, s>>>(frames, n, (long long)H * W, ns, (__half*)w.actA[0],
                                               (__half*)w.actA[1]);
  PFNL_LAUNCH_CHECK();
  int rc;
  const bool flow = tw.flow;
  if (flow) {  // one block through the dataflow kernel
    rc = tc_pfrb_flow(tw, w, blk, 1, N, H, W, false, s);
    *launches += 1;
  } else {
    rc = ns == 2 ? pfrb_tc<2>(tw, w, blk, N, H, W, s, launches, nullptr)
                 : pfrb_tc<1>(tw, w, blk, N, H, W, s, launches, nullptr);
  }
  if (rc) return rc;
  planes_to_f32_kernel<<<148 * 8, 256, 0, s>>>((const __half*)w.actA[0], (const __half*)w.actA[1], n,
                                               (long long)H * W, ns, frames_out);
  PFNL_LAUNCH_CHECK();
  *launches += 2;
  return PFNL_OK;
}

// Stage-level entries in the handle's precision (pfnl_conv0 / pfnl_convmerge1): the tcgen05 kernels with fp32
// tensors either side.
int tc_conv0_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, const float* inp21, int N, int H, int W,
                    float* frames_out, cudaStream_t s, long long* launches) {
  const int ns = tc_nsplit(precision);
  int rc = ns == 2 ? conv0_tc<2>(tw, w, inp21, N, H, W, s) : conv0_tc<1>(tw, w, inp21, N, H, W, s);
  if (rc) return rc;
  const long long n = (long long)N * kFrames * H * W * 64;
  planes_to_f32_kernel<<<148 * 8, 256, 0, s>>>((const __half*)w.actA[0], (const __half*)w.actA[1], n,
                                               (long long)H * W, ns, frames_out);
  PFNL_LAUNCH_CHECK();
  *launches += 2;
  return PFNL_OK;
}

int tc_merge1_fp32io(const TcWeights& tw, TcWorkspace& w, int precision, const float* frames, int N, int H, int W,
                     float* merge, cudaStream_t s, long long* launches) {
  const int ns = tc_nsplit(precision);
  const long long n = (long long)N * kFrames * H * W * 64;
  f32_to_planes_kernel<<<148 * 8, 256, 0, s>>>(frames, n, (long long)H * W, ns, (__half*)w.actA[0],
                                               (__half*)w.actA[1]);
  PFNL_LAUNCH_CHECK();
  int rc = ns == 2 ? merge1_tc<2>(tw, w, N, H, W, merge, s) : merge1_tc<1>(tw, w, N, H, W, merge, s);
  if (rc) return rc;
  *launches += 2;
  return PFNL_OK;
}

}  // namespace pfnl
