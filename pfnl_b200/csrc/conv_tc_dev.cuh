// Device-side building blocks shared by the tensor-core conv kernels (conv_tc.cu: the phase kernels;
// pfrb_flow.cu: the persistent dataflow kernel of the PFRB stack): phase shapes, ring/TMEM bookkeeping and the
// per-tile bodies of the three warp roles (TMA producer, MMA issuer, epilogue).  See conv_tc.cu for the layout
// and the implicit-GEMM formulation.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc.h"
#include "tc_ptx.cuh"
#include "tc_tmap.h"

namespace pfnl {

using namespace tc;

enum TcEpi {
  kEpiActPlanes = 0,   // v = lrelu(acc + bias)                        -> fp16 planes
  kEpiPartialF32 = 1,  // v = acc (+ previous content if accumulate)   -> fp32
  kEpiResPlanes = 2,   // v = lrelu(acc + pbase + bias) + residual     -> fp16 planes (may alias residual)
  kEpiFinalF32 = 3     // v = lrelu(acc + previous + bias)             -> fp32
};

// compile-time description of one phase
template <int KS_, int NSRC_, int NOUT_, int NCH_>
struct PhaseCfg {
  static constexpr int KS = KS_, NSRC = NSRC_, NOUT = NOUT_, NCH = NCH_;
  static constexpr int TAPS = KS * KS;
  static constexpr int NTAPS = NSRC * TAPS;
  static constexpr int BOX_W = KS == 3 ? kTcPatchW3 : 8;
  static constexpr int BOX_H = KS == 3 ? 18 : 16;
  static constexpr int PATCH_BYTES = BOX_W * BOX_H * 128;
  static constexpr int WT_BYTES = NOUT * 128;  // one plane of one tap: [NOUT rows][64 ci]
};

constexpr int cmax(int a, int b) { return a > b ? a : b; }

constexpr int kTcEpiWarps = 16;                     // 4 TMEM lane quarters x 4 sixteen-channel chunks
constexpr int kTcThreads = (2 + kTcEpiWarps) * 32;  // + producer warp + MMA warp = 576

struct TcRing {  // ring slot cursor (producer and MMA issuer keep identical copies)
  int sl, ph;
};

__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// element offset of (image, 8-channel chunk, y, x) in a channel-chunk-major fp16 plane [img][8][H][W][8]
__device__ __forceinline__ long long plane_off(int img, int chunk8, int y, int x, int H, int W) {
  return ((((long long)img * 8 + chunk8) * H + y) * W + x) * 8;
}
// fp32 partial sums of conv2's base half: [img][4 chunks][H][W][16] (64 contiguous bytes per thread)
__device__ __forceinline__ long long pbase_off(int img, int chunk16, int y, int x, int H, int W) {
  return ((((long long)img * 4 + chunk16) * H + y) * W + x) * 16;
}

// accumulation chain of tap tp (3x3: taps 0-4 -> chain 0, 5-8 -> chain 1 when NCH = 2) or source s
template <int KS, int NCH>
__device__ __forceinline__ constexpr int tc_chain(int tp, int s) {
  return NCH == 1 ? 0 : (KS == 3 ? (NCH == 2 ? (tp >= 5 ? 1 : 0) : tp / 3) : s % NCH);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per lane, half the
// LSU requests of two 16-byte accesses for the thread-per-pixel-row pattern of the epilogue.
struct __align__(32) U256 {
  uint32_t w[8];
};
__device__ __forceinline__ U256 ld256(const void* ptr) {
  U256 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                 "=r"(r.w[7])
               : "l"(ptr)
               : "memory");
  return r;
}
__device__ __forceinline__ void st256(void* ptr, const U256& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}
// one fp16 plane access of an epilogue thread: its 16 channels = two 16-byte pieces, `cs` elements apart
// (cs = H*W*8, the chunk stride of the plane)
__device__ __forceinline__ U256 ld_plane16(const __half* p, long long cs) {
  U256 r;
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *reinterpret_cast<const uint4*>(p + cs);
  r.w[0] = a.x, r.w[1] = a.y, r.w[2] = a.z, r.w[3] = a.w;
  r.w[4] = b.x, r.w[5] = b.y, r.w[6] = b.z, r.w[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_plane16(__half* p, long long cs, const U256& r) {
  *reinterpret_cast<uint4*>(p) = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
  *reinterpret_cast<uint4*>(p + cs) = make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]);
}
__device__ __forceinline__ void st_plane16_hint(__half* p, long long cs, const U256& r, unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "l"(policy)
               : "memory");
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p + cs), "r"(r.w[4]), "r"(r.w[5]),
               "r"(r.w[6]), "r"(r.w[7]), "l"(policy)
               : "memory");
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// generic <-> async proxy ordering for GLOBAL memory only (the dataflow kernel's cross-CTA hand-over: stores of one
// CTA, TMA loads of another); does not have to drain the shared-memory side
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Whole weight image: linear bulk copies L2 -> smem (the image is stored pre-swizzled, UMMA-ready).
template <int W_BYTES>
__device__ __forceinline__ void load_weights(uint8_t* wsm, const __half* wimg, uint64_t* wfull) {
  for (int off = 0; off < W_BYTES; off += 32768) {
    const int n = (W_BYTES - off) < 32768 ? (W_BYTES - off) : 32768;
    bulk_load(wsm + off, reinterpret_cast<const uint8_t*>(wimg) + off, n, wfull);
  }
}

// 2 x 32-byte loads that bypass the (non-coherent) L1: data another CTA of the SAME launch has written
// (LDG.E.ENL2.256.STRONG.GPU: one full sector per lane and instruction - as four 16-byte loads the same 64 bytes
// cost every conv2 tile 1.1 K cycles, profiles/r2z_epi_parts.txt)
__device__ __forceinline__ void ld256x2_coherent(const float* p, U256& a, U256& b) {
  asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.w[0]), "=r"(a.w[1]), "=r"(a.w[2]), "=r"(a.w[3]), "=r"(a.w[4]), "=r"(a.w[5]), "=r"(a.w[6]),
                 "=r"(a.w[7])
               : "l"(p)
               : "memory");
  asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]),
                 "=r"(b.w[7])
               : "l"(p + 8)
               : "memory");
}

// mbarriers of the tile pipeline: patch ring (TMA -> MMA) and the two TMEM accumulator buffers (MMA -> epilogue)
struct TcBars {
  uint64_t full[8], empty[8];
  uint64_t tmem_full[2], tmem_empty[2];
};

// what the epilogue does with a finished accumulator tile
struct TcEpiArgs {
  int epi;                   // TcEpi
  int accumulate;            // kEpiPartialF32: add the previous content of out_f32
  int f32_chunked;           // out_f32 is chunk-major [img][4][H][W][16] (the conv2 partial sums) instead of NHWC
  int coherent_pbase;        // pbase was written by another CTA of this launch: read it past the L1
  const float* pbase;        // fp32 [out_img/frames][4][H][W][16], channel-chunk-major (kEpiResPlanes)
  __half* out_hi;
  __half* out_lo;
  const __half* res_hi;
  const __half* res_lo;
  float* out_f32;            // [out_images][H][W][NOUT]
  int H, W;
  float trunc_comp;          // kappa of the truncation-bias compensation below (0: off)
  unsigned long long st_policy;  // L2 eviction hint of the fp16-plane stores (0: none)
};

// TMEM accumulation truncates toward zero (probes/umma_probe.cu): every accumulating MMA loses on average half an
// ulp of the running sum, in the direction of zero - a bias, not noise.  For a sum built from zero-mean terms
// (Xavier/zero-mean weights) E[s_k | s_K] = (k/K) s_K, and ulp(s)/|s| averages 0.72 * 2^-23 over a binade, so the
// expected loss of a chain of K accumulations is 0.36 * 2^-23 * sum_k s_k = kappa * (K+1) * 2^-23 * s_K with
// kappa = 0.18.  The epilogue multiplies each chain's D0 by 1 + kappa (K+1) 2^-23: what remains of the
// truncation is its zero-mean part, as with round-to-nearest.  (D1 carries the 2^-11-scaled cross terms; its
// truncation is three orders of magnitude below.)  Number of accumulating MMAs of chain c in one tile:
template <class PC>
__device__ __forceinline__ constexpr int tc_chain_len(int c) {
  int n = 0;
  for (int s = 0; s < PC::NSRC; ++s)
    for (int tp = 0; tp < PC::TAPS; ++tp)
      if (tc_chain<PC::KS, PC::NCH>(tp, s) == c) n += 4;
  return n;
}

// ---------------------------------------------------------------------------------------------------------
// per-tile bodies of the three warp roles
// ---------------------------------------------------------------------------------------------------------
// TMA producer (one thread): the NSRC x NSPLIT patch loads of one tile into the next ring slots.  Source image
// of stage s = ic0 + s; (x0, y0) = top-left pixel of the halo patch (may be negative: zero fill = 'same' padding).
template <class PC, int NSPLIT, int NS, int SLOT_BYTES>
__device__ __forceinline__ void load_tile(const CUtensorMap* tm_hi, const CUtensorMap* tm_lo, uint8_t* ring,
                                          TcBars* bars, TcRing& rg, int x0, int y0, int ic0, uint64_t l2_policy = 0,
                                          long long* ts = nullptr) {
  for (int s = 0; s < PC::NSRC; ++s) {
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl) {  // hi plane first (consumed first), then lo
      mbar_wait(&bars->empty[rg.sl], rg.ph ^ 1);
      if (ts != nullptr && s == 0) ts[pl] = clock64();  // (trace) slot free: this load goes out now
      mbar_arrive_expect_tx(&bars->full[rg.sl], PC::PATCH_BYTES);
      if (l2_policy != 0)
        tma_load_4d_hint(ring + rg.sl * SLOT_BYTES, pl == 1 ? tm_lo : tm_hi, &bars->full[rg.sl], x0 * 8, y0, 0, ic0 + s,
                         l2_policy);
      else
        tma_load_4d(ring + rg.sl * SLOT_BYTES, pl == 1 ? tm_lo : tm_hi, &bars->full[rg.sl], x0 * 8, y0, 0, ic0 + s);
      if (++rg.sl == NS) {
        rg.sl = 0;
        rg.ph ^= 1;
      }
    }
  }
}

// MMA issuer (converged warp, one elected lane issues): all MMAs of tile number `it` of this CTA.
//   tr: clock stamps of the traced CTA (or NULL), tri: index of this tile in the trace window
template <class PC, int NSPLIT, int NS, int SLOT_BYTES, int TMEM_BUF_COLS, int CH_STRIDE>
__device__ __forceinline__ void mma_tile(uint8_t* wsm, uint8_t* ring, TcBars* bars, uint32_t tmem, TcRing& rg, int it,
                                         int lane, long long* tr, int tri, long long* ts = nullptr) {
  constexpr int KS = PC::KS, NSRC = PC::NSRC, NOUT = PC::NOUT, NCH = PC::NCH, TAPS = PC::TAPS;
  constexpr int TAP_BYTES = NSPLIT * PC::WT_BYTES;
  constexpr uint32_t idesc_lo = make_idesc_f16(128, NOUT);           // A_lo x W_hi          -> D1
  constexpr uint32_t idesc_hi = make_idesc_f16(128, NSPLIT * NOUT);  // A_hi x [W_hi ; W_lo] -> [D0 | D1]
  // A operand (un-swizzled K-major): pixels of 16 B per 8-channel sub-patch; an 8-pixel tile row is one core
  // matrix, consecutive tile rows are one patch row apart (SBO), the two k-chunks one sub-patch apart (LBO)
  constexpr uint32_t SBO_A = PC::BOX_W * 16;
  constexpr uint32_t SUB_A = PC::BOX_W * PC::BOX_H * 16;  // bytes per sub-patch
  const uint64_t wd = make_sdesc_sw128(smem_u32(wsm), 1024, 0);
  const int buf = it & 1;
  if (ts != nullptr && lane == 0) ts[2] = clock64();
  mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);
  fence_after_sync();
  if (ts != nullptr && lane == 0) ts[3] = clock64();
  const uint32_t dbase = tmem + buf * TMEM_BUF_COLS;
  uint32_t accmask = 0;  // bit c: chain c's block [D0|D1] has been written in this tile
  for (int s = 0; s < NSRC; ++s) {
    const uint64_t wsd = wd + (uint64_t)((s * TAPS * TAP_BYTES) >> 4);
    // ---- hi-plane pass: [D0_c | D1_c] (+)= A_hi x [W_hi ; W_lo]
    mbar_wait(&bars->full[rg.sl], rg.ph);
    fence_after_sync();
    if (tr != nullptr && lane == 0 && s == 0 && 1 + 2 * tri < 64) tr[64 + 1 + 2 * tri] = clock64();
    {
      const uint64_t ad = make_sdesc_interleave(smem_u32(ring + rg.sl * SLOT_BYTES), SUB_A, SBO_A);
      if (elect_one()) {
        uint32_t am = accmask;
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp) {
          const int ch = tc_chain<KS, NCH>(tp, s);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t aoff = (((tp / KS) * PC::BOX_W + (tp % KS)) * 16 + k * 2 * SUB_A) >> 4;
            const uint32_t boff = (tp * TAP_BYTES + k * 32) >> 4;
            mma_f16(dbase + ch * CH_STRIDE, ad + aoff, wsd + boff, idesc_hi, (am >> ch) & 1u);
            am |= 1u << ch;
          }
        }
        mma_commit(&bars->empty[rg.sl]);
        if (NSPLIT == 1 && s == NSRC - 1) mma_commit(&bars->tmem_full[buf]);
      }
      __syncwarp();
#pragma unroll
      for (int tp = 0; tp < TAPS; ++tp) accmask |= 1u << tc_chain<KS, NCH>(tp, s);
      if (++rg.sl == NS) {
        rg.sl = 0;
        rg.ph ^= 1;
      }
    }
    if (NSPLIT == 2) {
      // ---- lo-plane pass: D1 of chain 0 += A_lo x W_hi (chain 0 was initialised by the hi pass:
      //      tap 0 / source 0 always belongs to chain 0), so this always accumulates
      if (ts != nullptr && lane == 0 && s == 0) ts[4] = clock64();
      mbar_wait(&bars->full[rg.sl], rg.ph);
      fence_after_sync();
      if (ts != nullptr && lane == 0 && s == 0) ts[5] = clock64();
      const uint64_t ad = make_sdesc_interleave(smem_u32(ring + rg.sl * SLOT_BYTES), SUB_A, SBO_A);
      if (elect_one()) {
#pragma unroll
        for (int tp = 0; tp < TAPS; ++tp) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t aoff = (((tp / KS) * PC::BOX_W + (tp % KS)) * 16 + k * 2 * SUB_A) >> 4;
            const uint32_t boff = (tp * TAP_BYTES + k * 32) >> 4;
            mma_f16(dbase + NOUT, ad + aoff, wsd + boff, idesc_lo, 1u);
          }
        }
        mma_commit(&bars->empty[rg.sl]);
        if (s == NSRC - 1) mma_commit(&bars->tmem_full[buf]);  // same thread that issued the MMAs
      }
      __syncwarp();
      if (++rg.sl == NS) {
        rg.sl = 0;
        rg.ph ^= 1;
      }
    }
    if (tr != nullptr && lane == 0 && s == NSRC - 1 && 2 + 2 * tri < 64) tr[64 + 2 + 2 * tri] = clock64();
  }
}

// Epilogue (16 warps): tile number `it` of this CTA = output image `img`, spatial tile (tx, ty); `nimg` = image of
// the per-unit partial sums.  `pre` holds the partial sums / previous fp32 content across calls (the phase kernels
// load the partial sums once per unit: load_pbase = first frame of the unit).
// `hook` lets the caller slot work into the two places where the warp has slack: hook.idle() runs when the
// accumulator is not ready yet (the warp would only wait), hook.before_stores() right before this tile's global
// stores (>= 1-2 K cycles after the previous tile's stores were issued).  The dataflow kernel publishes the
// previous tile there: a __threadfence issued directly after the stores waits 3-5 K cycles for their
// acknowledgements (measured), issued one tile later it finds them complete.
struct TcNoHook {
  __device__ __forceinline__ void idle() {}
  __device__ __forceinline__ void before_stores() {}
  __device__ __forceinline__ constexpr bool skip(int) const { return false; }
};
// (instrumented builds) leaves out parts of the epilogue to time the rest: 4 = partial-sum loads, 8 = residual loads,
// 16 = stores.  The results are garbage.
struct TcSkipHook {
  int bits;
  __device__ __forceinline__ void idle() {}
  __device__ __forceinline__ void before_stores() {}
  __device__ __forceinline__ bool skip(int b) const { return (bits & b) != 0; }
};
// EPI0 = index of the first of the 16 epilogue warps (a multiple of 2 with EPI0 % 4 == warp-quarter alignment kept by
// `warp & 3`): 2 in the phase kernels (warps 2-17), 4 in the dataflow kernel (warps 4-19, whole warpgroups).
template <class PC, int NSPLIT, int TMEM_BUF_COLS, int CH_STRIDE, class Hook, int EPI0 = 2>
__device__ __forceinline__ void epi_tile(const TcEpiArgs& P, TcBars* bars, const float* bias_sm, uint32_t tmem, int it,
                                         int warp, int lane, int img, int nimg, int tx, int ty, bool load_pbase,
                                         U256 (&pre)[2], long long* tr, Hook& hook, int tri) {
  constexpr int NOUT = PC::NOUT, NCH = PC::NCH;
  const int q = warp & 3;             // TMEM lane quarter this warp may access (hardware restriction)
  const int c0 = ((warp - EPI0) >> 2) * 16;
  const bool chunk_active = c0 < NOUT;
  const int m = q * 32 + lane;        // row of the tile = TMEM lane
  const int my = m >> 3, mx = m & 7;  // pixel inside the 16x8 tile
  const bool epi_planes = P.epi == kEpiActPlanes || P.epi == kEpiResPlanes;
  const bool epi_res = P.epi == kEpiResPlanes;
  const bool epi_prev = (P.epi == kEpiPartialF32 && P.accumulate) || P.epi == kEpiFinalF32;
  const int buf = it & 1;
  const int y = ty * 16 + my, x = tx * 8 + mx;
  const bool inb = chunk_active && y < P.H && x < P.W;
  const long long pix = ((long long)img * P.H + y) * P.W + x;
  const long long poff = plane_off(img, c0 >> 3, y, x, P.H, P.W);  // this thread's first 8 channels in a plane
  const long long cs = (long long)P.H * P.W * 8;                    // ... the other 8 are one chunk further
  const long long foff = pbase_off(img, c0 >> 4, y, x, P.H, P.W);
  // ---- prefetch (independent of the accumulator) ----
  U256 rh, rl;
#pragma unroll
  for (int j = 0; j < 8; ++j) rh.w[j] = rl.w[j] = 0u;
  if (inb) {
    if (epi_res) {
      if (load_pbase && !hook.skip(4)) {  // the base-half partial sums are shared by the unit's 7 frames
        const float* pb = P.pbase + pbase_off(nimg, c0 >> 4, y, x, P.H, P.W);
        if (P.coherent_pbase) {
          ld256x2_coherent(pb, pre[0], pre[1]);
        } else {
          pre[0] = ld256(pb);
          pre[1] = ld256(pb + 8);
        }
      }
      if (!hook.skip(8)) {
        rh = ld_plane16(P.res_hi + poff, cs);
        if (NSPLIT == 2) rl = ld_plane16(P.res_lo + poff, cs);
      }
    } else if (epi_prev) {
      const float* o = P.out_f32 + (P.f32_chunked ? foff : pix * NOUT + c0);
      pre[0] = ld256(o);
      pre[1] = ld256(o + 8);
    }
  }
  if (!__all_sync(0xffffffffu, mbar_try_wait(&bars->tmem_full[buf], (it >> 1) & 1))) {
    hook.idle();
    mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
  }
  fence_after_sync();
  if (tr != nullptr && warp == EPI0 && lane == 0 && 2 * tri < 64) tr[128 + 2 * tri] = clock64();
  float v[16];
  if (chunk_active) {
    // chain c: D0 at c*CH_STRIDE, D1 (split mode) at c*CH_STRIDE + NOUT.  Chains are summed in
    // fp32 round-to-nearest here; D1 carries the 2^-11-scaled cross terms.
    const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + buf * TMEM_BUF_COLS + c0;
    uint32_t d0[NCH][16], d1[NSPLIT == 2 ? NCH : 1][16];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tmem_ld_32x32b_x16(t0 + c * CH_STRIDE, d0[c]);
      if (NSPLIT == 2) tmem_ld_32x32b_x16(t0 + c * CH_STRIDE + NOUT, d1[c]);
    }
    tmem_ld_wait();
    float comp[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) comp[c] = 1.f + P.trunc_comp * (float)(tc_chain_len<PC>(c) + 1) * 1.1920929e-7f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // __fmul_rn: never contracted with a later add into an FMA - the phase kernels and the dataflow kernel inline
      // this code in different contexts and must round identically (they are compared bit for bit)
      float a = __fmul_rn(__uint_as_float(d0[0][j]), comp[0]);
#pragma unroll
      for (int c = 1; c < NCH; ++c) a = fmaf(__uint_as_float(d0[c][j]), comp[c], a);
      if (NSPLIT == 2) {
        float b = __uint_as_float(d1[0][j]);
#pragma unroll
        for (int c = 1; c < NCH; ++c) b += __uint_as_float(d1[c][j]);
        a = fmaf(b, 1.f / 2048.f, a);
      }
      v[j] = a;
    }
  }
  // this warp's tcgen05.ld are complete: hand the TMEM buffer back before the global stores
  fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
  U256 oa, ob;  // the two 32-byte pieces this thread stores: fp16 planes (hi, lo) or 16 fp32
#pragma unroll
  for (int j = 0; j < 8; ++j) oa.w[j] = ob.w[j] = 0u;
  if (inb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // + partial sums / previous content (zeros otherwise)
      v[j] += __uint_as_float(pre[0].w[j]);
      v[8 + j] += __uint_as_float(pre[1].w[j]);
    }
    if (epi_planes) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = lrelu(v[j] + bias_sm[c0 + j]);
      if (epi_res) {
        const __half* hh = reinterpret_cast<const __half*>(&rh);
        if (NSPLIT == 2) {
          // residual = hi + lo/2048 (exactly representable in fp32), then one rounded add
          const __half* hl = reinterpret_cast<const __half*>(&rl);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += fmaf(__half2float(hl[j]), 1.f / 2048.f, __half2float(hh[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __half2float(hh[j]);
        }
      }
      __half* ph = reinterpret_cast<__half*>(&oa);
      __half* pl = reinterpret_cast<__half*>(&ob);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (NSPLIT == 2)
          split_half(v[j], ph[j], pl[j]);
        else
          ph[j] = __float2half_rn(v[j]);
      }
    } else {
      if (P.epi == kEpiFinalF32) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = lrelu(v[j] + bias_sm[c0 + j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        oa.w[j] = __float_as_uint(v[j]);
        ob.w[j] = __float_as_uint(v[8 + j]);
      }
    }
  }
  hook.before_stores();
  if (inb && !hook.skip(16)) {
    if (epi_planes) {
      if (P.st_policy != 0) {
        st_plane16_hint(P.out_hi + poff, cs, oa, P.st_policy);
        if (NSPLIT == 2) st_plane16_hint(P.out_lo + poff, cs, ob, P.st_policy);
      } else {
        st_plane16(P.out_hi + poff, cs, oa);
        if (NSPLIT == 2) st_plane16(P.out_lo + poff, cs, ob);
      }
    } else {
      float* o = P.out_f32 + (P.f32_chunked ? foff : pix * NOUT + c0);
      st256(o, oa);
      st256(o + 8, ob);
    }
  }
  if (tr != nullptr && warp == EPI0 && lane == 0 && 2 * tri + 1 < 64) tr[128 + 2 * tri + 1] = clock64();
}

}  // namespace pfnl
