// Non-local block on tcgen05 tensor cores: NonLocalBlock nltype=1 'gaussian', sub_sample=1 (utils.py:18-71)
// with theta = phi = X (utils.py:33-34,41-42).  Two operand precisions (NlCfg<NSPLIT>): fp16 (PFNL_PREC_TC_FP16,
// _FP16X3_NLTC) and hi/lo-split fp16 pairs = fp32-grade logits and values (PFNL_PREC_TC_FP16X3).
//
//   nl_prep_kernel   X fp32 [N,L,84] -> X16 [N,Lp,128] fp16 (zero padded; the Q/K operand) and X^T -> Xt16
//                    [N,96,Lp] fp16 (V, stored channel-major so that the PV B-operand is K-major; V = X because
//                    the g linear is folded into the output linear), each as hi [, lo] planes
//   nl_tc_kernel     per (clip, 128-query tile): TMA -> smem, S = Q K^T (tcgen05, fp32 in TMEM),
//                    online softmax on 8 warps (two threads per query row, 64 keys each: row max
//                    exchanged through smem, exp / sum thread-local), P (fp16) -> smem, O += P V
//                    (tcgen05), LAZY running-max correction of the TMEM-resident O (tcgen05.ld/st
//                    only when the row max grows by more than 8: P <= e^8 stays far inside fp16 and
//                    the common scale cancels in Y = O / l).  l itself is an output column: V carries an all-ones
//                    channel, so sum_j P_ij is accumulated in TMEM exactly like the values (same truncation).
//                    The L x L matrix (utils.py:53-58) is never materialised.
//   then             Z = Y*(Wg*Ww)+(bg*Ww+bw), depth_to_space, + input  (nl_linear_scatter, nonlocal_ffma.cu)
//
// fp16 mode, per CTA: Q 32 KB, 2 x (K 32 KB + V 24 KB) ring, 2 x P 32 KB (softmax(j+1) writes one buffer while
// the PV product of tile j still reads the other); TMEM 2 x 128 columns of S (S(j+1) is issued while
// softmax(j) runs) + 96 columns of O.  Split mode: see NlCfg.
#include <cuda_fp16.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "tc.h"
#include "tc_ptx.cuh"
#include "tc_tmap.h"

namespace pfnl {

using namespace tc;

namespace {

constexpr int kQT = 128;   // queries per CTA
constexpr int kKT = 128;   // keys per tile
constexpr int kCP = 128;   // padded channel count of X16 (84 -> 128: two 64-wide K blocks)
constexpr int kVR = 96;    // padded channel count of Gt16 rows (84 -> 96, multiple of 16)

constexpr int kNlSoftmaxWarps = 8;                       // 4 TMEM lane quarters x 2 column halves
constexpr int kNlThreads = (2 + kNlSoftmaxWarps) * 32;  // + TMA warp + MMA warp = 320
constexpr float kNlRescaleThreshold = 8.f;

// NSPLIT = 1: fp16 operands.  NSPLIT = 2: hi/lo-split operands (x = hi + lo/2048, like the convs):
//   [S_hh | S_x] = Q_hi x [K_hi ; K_lo]^T (one N = 256 MMA per k-step),  S_x += Q_lo x K_hi^T,  S = S_hh + S_x/2048
//   [O_h | O_l]  = P x [V_hi ; V_lo]                                       O = O_h + O_l/2048
// i.e. fp32-grade logits and values; only P (in [0, e^8], normalised by the sum of the SAME rounded values)
// stays a single fp16.  The split mode has no room to double-buffer (shared memory: Q 64 + K 64 + V 48 +
// P 32 KB; TMEM: S 256 + O 192 columns), so its stages are single-buffered.
template <int NSPLIT>
struct NlCfg {
  static constexpr int KSTAGES = NSPLIT == 1 ? 2 : 1;  // K/V ring depth
  static constexpr int SBUFS = NSPLIT == 1 ? 2 : 1;    // S accumulators in TMEM
  static constexpr int PBUFS = NSPLIT == 1 ? 2 : 1;    // P tiles in shared memory
  static constexpr int PLANE = kQT * 128;              // one plane of one 64-wide block: 128 rows x 128 B
  static constexpr int BLK_QK = NSPLIT * PLANE;        // one 64-channel block of Q or K: [hi ; lo] rows
  static constexpr int Q_BYTES = 2 * BLK_QK;
  static constexpr int K_BYTES = 2 * BLK_QK;
  static constexpr int VPLANE = kVR * 128;             // one plane of one 64-key block of V^T: 96 rows x 128 B
  static constexpr int BLK_V = NSPLIT * VPLANE;
  static constexpr int V_BYTES = 2 * BLK_V;
  static constexpr int P_BYTES = 2 * kQT * 128;
  static constexpr int SMEM = 1024 + Q_BYTES + KSTAGES * (K_BYTES + V_BYTES) + PBUFS * P_BYTES + 3072;
  static constexpr int S_COLS = NSPLIT * 128;
  static constexpr int O_COLS = NSPLIT * kVR;
  static constexpr int TM_O = SBUFS * S_COLS;          // 256 in both modes
  static_assert(TM_O + O_COLS <= 512, "TMEM overflow");
  static_assert(SMEM <= 227 * 1024, "shared memory overflow");
};

struct NlCtrl {
  float rowmax[2][2][kQT];  // [tile parity][column half][row]: per-tile row maxima of the two half-row threads
  float rowsum[2][kQT];     // [column half][row]: final softmax denominators of the two half-row threads
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_ready, pv_done[2];  // pv_done[b]: the PV product reading P buffer b has completed
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ bool elect_one_nl() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 2-D map over a row-major fp16 matrix [rows, cols]: box = (64 cols, box_rows), 128B swizzle.
int make_mat_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return (int)fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// this thread's 16/32 columns of O at `col`: multiply by alpha in place
__device__ __forceinline__ void rescale_o32(uint32_t taddr, float alpha) {
  uint32_t o[32];
  tmem_ld_32x32b_x32(taddr, o);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
  tmem_st_32x32b_x32(taddr, o);
}
__device__ __forceinline__ void rescale_o16(uint32_t taddr, float alpha) {
  uint32_t o[16];
  tmem_ld_32x32b_x16(taddr, o);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
  tmem_st_32x32b_x16(taddr, o);
}

template <int NSPLIT>
__global__ void __launch_bounds__(kNlThreads, 1)
    nl_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_xl,
                 const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_gl, int L, int Lp,
                 float* __restrict__ Y, int ksplit, float* __restrict__ Ypart, float* __restrict__ ML) {
  using CF = NlCfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_sm = smem;                                  // [2 blocks][NSPLIT planes][128 rows][128 B]
  uint8_t* k_sm = q_sm + CF::Q_BYTES;                    // [KSTAGES] x same
  uint8_t* v_sm = k_sm + CF::KSTAGES * CF::K_BYTES;      // [KSTAGES][2 key blocks][NSPLIT planes][96 rows][128 B]
  uint8_t* p_sm = v_sm + CF::KSTAGES * CF::V_BYTES;      // [PBUFS][2 key blocks][128 rows][128 B]
  NlCtrl* ctl = reinterpret_cast<NlCtrl*>(p_sm + CF::PBUFS * CF::P_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, q0 = blockIdx.x * kQT;
  // key split (large L, few clips): this CTA covers key tiles [j0, j0 + ntiles) of the row and writes an
  // un-normalised partial (O, reference max, sum) that nl_merge_kernel combines (log-sum-exp merge)
  const int split = blockIdx.z;
  const int all_tiles = Lp / kKT;
  const int j0 = (int)((long long)split * all_tiles / ksplit);
  const int ntiles = (int)((long long)(split + 1) * all_tiles / ksplit) - j0;

  if (tid == 0) {
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->k_full[i], 1);
      mbar_init(&ctl->k_empty[i], 1);
      mbar_init(&ctl->v_full[i], 1);
      mbar_init(&ctl->v_empty[i], 1);
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->s_empty[i], kNlSoftmaxWarps);
      mbar_init(&ctl->pv_done[i], 1);
    }
    mbar_init(&ctl->p_ready, kNlSoftmaxWarps);
    fence_mbar_init();
    fence_proxy_async();
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_g);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&tm_xl);
      tma_prefetch_desc(&tm_gl);
    }
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = ctl->tmem_base;
  const uint32_t tm_s0 = tmem, tm_o = tmem + CF::TM_O;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctl->q_full, CF::Q_BYTES);
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl)
          tma_load_2d(q_sm + b * CF::BLK_QK + pl * CF::PLANE, pl ? &tm_xl : &tm_x, &ctl->q_full, b * 64, n * Lp + q0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % CF::KSTAGES, ph = (j / CF::KSTAGES) & 1;
        mbar_wait(&ctl->k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&ctl->k_full[st], CF::K_BYTES);
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int pl = 0; pl < NSPLIT; ++pl)
            tma_load_2d(k_sm + st * CF::K_BYTES + b * CF::BLK_QK + pl * CF::PLANE, pl ? &tm_xl : &tm_x,
                        &ctl->k_full[st], b * 64, n * Lp + (j0 + j) * kKT);
        mbar_wait(&ctl->v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&ctl->v_full[st], CF::V_BYTES);
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int pl = 0; pl < NSPLIT; ++pl)
            tma_load_2d(v_sm + st * CF::V_BYTES + b * CF::BLK_V + pl * CF::VPLANE, pl ? &tm_gl : &tm_g,
                        &ctl->v_full[st], (j0 + j) * kKT + b * 64, n * kVR);
      }
    }
  } else if (warp == 1) {
    // converged warp, one elected lane issues (see conv_tc.cu); descriptors advance by constant adds
    constexpr uint32_t idesc_s = make_idesc_f16(128, NSPLIT * 128);  // Q_hi x [K_hi ; K_lo]
    constexpr uint32_t idesc_sx = make_idesc_f16(128, 128);          // Q_lo x K_hi
    constexpr uint32_t idesc_o = make_idesc_f16(128, NSPLIT * kVR);  // P x [V_hi ; V_lo]
    const uint64_t qd = make_sdesc_sw128(smem_u32(q_sm), 1024, 0);
    auto issue_s = [&](int j) {
      const int st = j % CF::KSTAGES, ph = (j / CF::KSTAGES) & 1;
      const int sb = j % CF::SBUFS, sph = (j / CF::SBUFS) & 1;
      mbar_wait(&ctl->k_full[st], ph);
      mbar_wait(&ctl->s_empty[sb], sph ^ 1);
      fence_after_sync();
      const uint64_t kd = make_sdesc_sw128(smem_u32(k_sm + st * CF::K_BYTES), 1024, 0);
      if (elect_one_nl()) {
        // channels 0..63 (block 0, 4 k-steps) and 64..95 (block 1, 2 k-steps; 96..127 are zero padding)
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
          const uint32_t off = ((ks >> 2) * CF::BLK_QK + (ks & 3) * 32) >> 4;
          mma_f16(tm_s0 + sb * CF::S_COLS, qd + off, kd + off, idesc_s, ks > 0 ? 1u : 0u);
        }
        if (NSPLIT == 2) {
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {
            const uint32_t off = ((ks >> 2) * CF::BLK_QK + (ks & 3) * 32) >> 4;
            mma_f16(tm_s0 + sb * CF::S_COLS + 128, qd + (CF::PLANE >> 4) + off, kd + off, idesc_sx, 1u);
          }
        }
        mma_commit(&ctl->k_empty[st]);
        mma_commit(&ctl->s_full[sb]);
      }
      __syncwarp();
    };
    mbar_wait(&ctl->q_full, 0);
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      if (CF::SBUFS == 2 && j + 1 < ntiles) issue_s(j + 1);
      const int st = j % CF::KSTAGES, ph = (j / CF::KSTAGES) & 1;
      const int pb = j % CF::PBUFS;
      mbar_wait(&ctl->v_full[st], ph);
      mbar_wait(&ctl->p_ready, j & 1);
      fence_after_sync();
      const uint64_t vd = make_sdesc_sw128(smem_u32(v_sm + st * CF::V_BYTES), 1024, 0);
      const uint64_t pd = make_sdesc_sw128(smem_u32(p_sm + pb * CF::P_BYTES), 1024, 0);
      if (elect_one_nl()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t aoff = ((ks >> 2) * (kQT * 128) + (ks & 3) * 32) >> 4;
          const uint32_t boff = ((ks >> 2) * CF::BLK_V + (ks & 3) * 32) >> 4;
          mma_f16(tm_o, pd + aoff, vd + boff, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        mma_commit(&ctl->v_empty[st]);
        mma_commit(&ctl->pv_done[pb]);
      }
      __syncwarp();
      if (CF::SBUFS == 1 && j + 1 < ntiles) issue_s(j + 1);
    }
  } else {
    // softmax / correction / epilogue: two threads per query row (column halves of S and of O)
    const int qd = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;     // 0: keys 0..63 / O columns 0..47 ; 1: keys 64..127 / O columns 48..95
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    float m_run = -INFINITY;  // the reference max the exponentials are taken against
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j % CF::SBUFS, sph = (j / CF::SBUFS) & 1;
      const int pb = j % CF::PBUFS;
      mbar_wait(&ctl->s_full[sb], sph);
      fence_after_sync();
      const int kvalid = L - (j0 + j) * kKT - half * 64;  // keys >= kvalid of this thread's 64 are padding
      float mt = -INFINITY;
      uint32_t sreg[2][32];
      const uint32_t s_addr = tm_s0 + sb * CF::S_COLS + lane_addr + half * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld_32x32b_x32(s_addr + c * 32, sreg[c]);
        if (NSPLIT == 2) {
          uint32_t sx[32];
          tmem_ld_32x32b_x32(s_addr + 128 + c * 32, sx);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sreg[c][i] = __float_as_uint(fmaf(__uint_as_float(sx[i]), 1.f / 2048.f, __uint_as_float(sreg[c][i])));
        }
      }
      tmem_ld_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->s_empty[sb]);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float sv = __uint_as_float(sreg[c][i]);
          if (c * 32 + i >= kvalid) sv = -INFINITY;
          sreg[c][i] = __float_as_uint(sv);
          mt = fmaxf(mt, sv);
        }
      // row max over both halves (the partner thread lives in another warp): smem + named barrier
      ctl->rowmax[j & 1][half][row] = mt;
      asm volatile("bar.sync 1, %0;" ::"n"(kNlSoftmaxWarps * 32) : "memory");
      mt = fmaxf(mt, ctl->rowmax[j & 1][half ^ 1][row]);  // finite: every tile holds at least one valid key
      // lazy rescaling: keep the old reference while the max grew by at most the threshold
      const bool bump = mt > m_run + kNlRescaleThreshold;  // identical in both threads of the row
      const float m_new = bump ? mt : m_run;
      const float alpha = bump ? __expf(m_run - m_new) : 1.f;  // first tile: exp(-inf) = 0, O and l are empty
      m_run = m_new;
      // O may only be rescaled once every PV product issued so far has landed (they complete in order)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(&ctl->pv_done[(j - 1) % CF::PBUFS], ((j - 1) / CF::PBUFS) & 1);
        fence_after_sync();
        // this thread's half of the 96 O columns (32 + 16), in every plane of O
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl) {
          rescale_o32(tm_o + lane_addr + pl * kVR + half * 48, alpha);
          rescale_o16(tm_o + lane_addr + pl * kVR + half * 48 + 32, alpha);
        }
        tmem_st_wait();
      }
      // this tile's P buffer was last read by the PV product of tile j - PBUFS
      if (j >= CF::PBUFS) mbar_wait(&ctl->pv_done[pb], ((j - CF::PBUFS) / CF::PBUFS) & 1);
      uint8_t* pbuf = p_sm + pb * CF::P_BYTES;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          __align__(16) __half hp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float pv = __expf(__uint_as_float(sreg[c][g8 * 8 + i]) - m_new);
            hp[i] = __float2half_rn(pv);
          }
          const int chunk = half * 8 + c * 4 + g8;  // 16-byte chunk index along the 128 keys
          const uint32_t off = (chunk >> 3) * (kQT * 128) + sw128_offset(row, chunk & 7);
          *reinterpret_cast<uint4*>(pbuf + off) = *reinterpret_cast<const uint4*>(hp);
        }
      }
      fence_proxy_async();   // make the generic-proxy P writes visible to the tensor-core (async) proxy
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->p_ready);
    }
    mbar_wait(&ctl->pv_done[(ntiles - 1) % CF::PBUFS], ((ntiles - 1) / CF::PBUFS) & 1);
    fence_after_sync();
    const int q = q0 + row;
    {
      uint32_t o[32];
      uint32_t o2[16];
      tmem_ld_32x32b_x32(tm_o + lane_addr + half * 48, o);
      tmem_ld_32x32b_x16(tm_o + lane_addr + half * 48 + 32, o2);
      if (NSPLIT == 2) {
        uint32_t ol[32];
        uint32_t ol2[16];
        tmem_ld_32x32b_x32(tm_o + lane_addr + kVR + half * 48, ol);
        tmem_ld_32x32b_x16(tm_o + lane_addr + kVR + half * 48 + 32, ol2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          o[i] = __float_as_uint(fmaf(__uint_as_float(ol[i]), 1.f / 2048.f, __uint_as_float(o[i])));
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o2[i] = __float_as_uint(fmaf(__uint_as_float(ol2[i]), 1.f / 2048.f, __uint_as_float(o2[i])));
      }
      tmem_ld_wait();
      // The softmax denominator is taken from the accumulator itself: V's first padding channel (84) is all ones,
      // so O[:,84] = sum_j P_ij went through exactly the same sequence of truncating TMEM accumulations (and lazy
      // rescalings) as the value columns - the toward-zero bias of long accumulation chains (8 per key tile: 256
      // at L = 4096) is common to numerator and denominator and cancels in Y = O / O[:,84], whatever the shape of
      // the attention row.  Column 84 lives in the half-1 thread: o2[84 - 48 - 32].
      if (half == 1) ctl->rowsum[1][row] = __uint_as_float(o2[kNL - 48 - 32]);
      asm volatile("bar.sync 1, %0;" ::"n"(kNlSoftmaxWarps * 32) : "memory");
      const float l_tot = ctl->rowsum[1][row];
      const float inv = ksplit > 1 ? 1.f : 1.f / l_tot;  // partials stay un-normalised
      if (ksplit > 1 && half == 0 && q < L) {
        float* ml = ML + (((long long)split * gridDim.y + n) * L + q) * 2;
        ml[0] = m_run;
        ml[1] = l_tot;
      }
      if (q < L) {
        float* dst = (ksplit > 1 ? Ypart + (long long)split * gridDim.y * L * kNL : Y) +
                     ((long long)n * L + q) * kNL + half * 48;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + i) =
              make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                          __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          if (half * 48 + 32 + i < kNL)
            *reinterpret_cast<float4*>(dst + 32 + i) =
                make_float4(__uint_as_float(o2[i]) * inv, __uint_as_float(o2[i + 1]) * inv,
                            __uint_as_float(o2[i + 2]) * inv, __uint_as_float(o2[i + 3]) * inv);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Log-sum-exp merge of the key-split partials: Y[r] = sum_s w_s O_s[r] / sum_s w_s l_s[r], w_s = exp(m_s - max_s m_s)
// (every split saw at least one valid key, so every m_s is finite).  Thread = one row x 4 channels.
__global__ void __launch_bounds__(256) nl_merge_kernel(const float* __restrict__ Ypart, const float* __restrict__ ML,
                                                       int ksplit, long long rows, float* __restrict__ Y) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * (kNL / 4)) return;
  const long long r = e / (kNL / 4);
  const int c4 = (int)(e - r * (kNL / 4)) * 4;
  float m = -INFINITY;
  for (int s = 0; s < ksplit; ++s) m = fmaxf(m, ML[((long long)s * rows + r) * 2]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float den = 0.f;
  for (int s = 0; s < ksplit; ++s) {
    const float w = __expf(ML[((long long)s * rows + r) * 2] - m);
    den = fmaf(w, ML[((long long)s * rows + r) * 2 + 1], den);
    const float4 o = *reinterpret_cast<const float4*>(Ypart + ((long long)s * rows + r) * kNL + c4);
    acc.x = fmaf(w, o.x, acc.x);
    acc.y = fmaf(w, o.y, acc.y);
    acc.z = fmaf(w, o.z, acc.z);
    acc.w = fmaf(w, o.w, acc.w);
  }
  const float inv = 1.f / den;
  *reinterpret_cast<float4*>(Y + r * kNL + c4) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// Key splits for (N clips, L tokens) on `sms` SMs (one CTA per SM: ~210 KB of shared memory): the k <= 8 that
// minimises waves(k)/k, each split keeping at least 2 key tiles; 1 when the query tiles alone fill the GPU or the
// row is short (L = 256 in the headline config: 2 key tiles, latency-bound either way).
int nl_pick_ksplit(int N, int L, int sms) {
  const int tiles = ceil_div(L, kKT);
  const long long base = (long long)tiles * N;
  int best = 1;
  double best_cost = (double)((base + sms - 1) / sms);
  for (int k = 2; k <= 8 && k * 2 <= tiles; ++k) {
    const double cost = (double)((base * k + sms - 1) / sms) / k + 0.01 * k;  // + a little for the merge
    if (cost < best_cost - 1e-9) {
      best = k;
      best_cost = cost;
    }
  }
  return best;
}

// X fp32 [N,L,84] -> X16 [plane][N,Lp,128] (zero padded; Q and K) and Xt16 [plane][N,96,Lp] = X^T (zero padded; V);
// plane 1 (nsplit = 2) holds (x - fp16(x)) * 2048.
// V = X because the g linear is folded into the output linear: w(P*(X*Wg+bg)) = P*X*(Wg*Ww) + (bg*Ww+bw)
// (softmax rows sum to 1; utils.py:26,64,67).  Pure cast + transpose, CTA = 64 tokens of one clip.
// With `lr` given (the forward path) the tokens are gathered straight from the LR clip - frame concat +
// space_to_depth(.,2), model/pfnl.py:55-57: channel (dy*2+dx)*21 + t*3 + c of token (h2,w2) is
// lr[n,t,2*h2+dy,2*w2+dx,c] - so the fp32 token matrix is never written (X is ignored).
__global__ void __launch_bounds__(256) nl_prep_kernel(const float* __restrict__ X, const float* __restrict__ lr, int H,
                                                      int W, int L, int Lp, int nsplit, __half* __restrict__ X16,
                                                      long long x_plane, __half* __restrict__ Xt16,
                                                      long long xt_plane) {
  __shared__ float xs[64 * (kNL + 1)];
  const int tid = threadIdx.x;
  const int n = blockIdx.y, t0 = blockIdx.x * 64;
  if (lr != nullptr) {
    const int W2 = W >> 1;
    for (int i = tid; i < 64 * kNL; i += 256) {
      const int tl = i / kNL, ch = i - tl * kNL;
      const int tok = t0 + tl;
      float v = 0.f;
      if (tok < L) {
        const int h2 = tok / W2, w2 = tok - h2 * W2;
        const int q = ch / 21, rr = ch - q * 21;
        const int t = rr / 3, c = rr - t * 3;
        v = lr[((((long long)n * kFrames + t) * H + 2 * h2 + (q >> 1)) * W + 2 * w2 + (q & 1)) * 3 + c];
      }
      xs[tl * (kNL + 1) + ch] = v;
    }
  } else {
    const float* Xn = X + (long long)n * L * kNL;
    for (int i = tid; i < 64 * kNL; i += 256) {
      const int tl = i / kNL, c = i - tl * kNL;
      xs[tl * (kNL + 1) + c] = (t0 + tl) < L ? Xn[(long long)t0 * kNL + i] : 0.f;
    }
  }
  __syncthreads();
  for (int i = tid; i < 64 * kCP; i += 256) {   // X16 rows (coalesced 256-byte rows)
    const int tl = i / kCP, c = i % kCP;
    const float v = c < kNL ? xs[tl * (kNL + 1) + c] : 0.f;
    const __half hi = __float2half_rn(v);
    const long long o = ((long long)n * Lp + t0 + tl) * kCP + c;
    X16[o] = hi;
    if (nsplit == 2) X16[x_plane + o] = __float2half_rn((v - __half2float(hi)) * 2048.f);
  }
  for (int i = tid; i < kVR * 64; i += 256) {   // Xt16 rows (token-contiguous)
    const int c = i >> 6, tl = i & 63;
    // row 84 (the first padding channel) is all ones: its output column accumulates sum_j P_ij in TMEM next to
    // the value columns, see the normalisation in nl_tc_kernel
    const float v = c < kNL ? xs[tl * (kNL + 1) + c] : (c == kNL ? 1.f : 0.f);
    const __half hi = __float2half_rn(v);
    const long long o = ((long long)n * kVR + c) * Lp + t0 + tl;
    Xt16[o] = hi;
    if (nsplit == 2) Xt16[xt_plane + o] = __float2half_rn((v - __half2float(hi)) * 2048.f);
  }
}

// bytes of the two operand arrays (always sized for two planes)
size_t nl_x_bytes(int N, size_t Lp) { return (2 * (size_t)N * Lp * kCP * 2 + 1023) / 1024 * 1024; }
size_t nl_g_bytes(int N, size_t Lp) { return (2 * (size_t)N * kVR * Lp * 2 + 1023) / 1024 * 1024; }

}  // namespace

bool tc_has_nonlocal() { return true; }

// hi/lo-split operands where the convs are split too (fp32-grade path), fp16 operands otherwise
int tc_nl_nsplit(int precision) { return precision == PFNL_PREC_TC_FP16X3 ? 2 : 1; }

int tc_nl_init() {
  PFNL_CUDA(tc_apply_wait_limit_from_env());
  PFNL_CUDA(cudaFuncSetAttribute(nl_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NlCfg<1>::SMEM));
  PFNL_CUDA(cudaFuncSetAttribute(nl_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, NlCfg<2>::SMEM));
  return PFNL_OK;
}

// `lr` != NULL: gather the tokens from the LR clip [N,7,H,W,3] (L = H/2 * W/2), `tokens` unused
static int run_nl_tc(const float* tokens, const float* lr, int H, int W, int N, int L, int nsplit, __half* x16,
                     __half* gt16, float* Y, float* part, int sms, cudaStream_t s, int* launches_out,
                     Profiler* kprof = nullptr) {
  const int Lp = ceil_div(L, kKT) * kKT;
  const long long x_plane = (long long)N * Lp * kCP, g_plane = (long long)N * kVR * Lp;
  dim3 pg(Lp / 64, N);
  nl_prep_kernel<<<pg, 256, 0, s>>>(tokens, lr, H, W, L, Lp, nsplit, x16, x_plane, gt16, g_plane);
  PFNL_LAUNCH_CHECK();
  CUtensorMap tmx, tmxl, tmg, tmgl;
  int r = make_mat_tmap(&tmx, x16, (uint64_t)N * Lp, kCP, kQT);
  if (r == 0) r = make_mat_tmap(&tmxl, x16 + (nsplit == 2 ? x_plane : 0), (uint64_t)N * Lp, kCP, kQT);
  if (r == 0) r = make_mat_tmap(&tmg, gt16, (uint64_t)N * kVR, (uint64_t)Lp, kVR);
  if (r == 0) r = make_mat_tmap(&tmgl, gt16 + (nsplit == 2 ? g_plane : 0), (uint64_t)N * kVR, (uint64_t)Lp, kVR);
  if (r != 0) {
    set_error("non-local tensor map encode failed (%d)", r);
    return PFNL_ERR_CUDA;
  }
  const int ksplit = part != nullptr ? nl_pick_ksplit(N, L, sms) : 1;
  float* ypart = part;
  float* ml = part != nullptr ? part + (size_t)ksplit * N * L * kNL : nullptr;
  dim3 grid(Lp / kQT, N, ksplit);
  if (kprof) kprof->begin(kProfNlKernel, s);
  if (nsplit == 2)
    nl_tc_kernel<2><<<grid, kNlThreads, NlCfg<2>::SMEM, s>>>(tmx, tmxl, tmg, tmgl, L, Lp, Y, ksplit, ypart, ml);
  else
    nl_tc_kernel<1><<<grid, kNlThreads, NlCfg<1>::SMEM, s>>>(tmx, tmxl, tmg, tmgl, L, Lp, Y, ksplit, ypart, ml);
  if (kprof) kprof->end(s);
  PFNL_LAUNCH_CHECK();
  *launches_out = 2;
  if (ksplit > 1) {
    const long long rows = (long long)N * L;
    nl_merge_kernel<<<(unsigned)ceil_div(rows * (kNL / 4), 256), 256, 0, s>>>(ypart, ml, ksplit, rows, Y);
    PFNL_LAUNCH_CHECK();
    *launches_out = 3;
  }
  return PFNL_OK;
}

// floats of key-split scratch for (N, L): partial O [k][N,L,84] + (max, sum) [k][N,L,2]
static size_t nl_part_floats(int N, int L, int sms) {
  const int k = nl_pick_ksplit(N, L, sms);
  return k > 1 ? (size_t)k * N * L * (kNL + 2) : 0;
}

size_t tc_nl_workspace_bytes(int N, int L) {
  const size_t Lp = (size_t)ceil_div(L, kKT) * kKT;
  // the split is chosen for the device's SM count at run time; size the scratch for the largest choice
  size_t part = 0;
  for (int sms = 64; sms <= 256; sms += 4) part = std::max(part, nl_part_floats(N, L, sms));
  return nl_x_bytes(N, Lp) + nl_g_bytes(N, Lp) + (part * 4 + 1023) / 1024 * 1024 + 1024;
}

int tc_nonlocal(const TcWeights& tw, TcWorkspace& w, const float* tokens, const float* lr, int N, int H, int W,
                float* inp21, cudaStream_t s, long long* launches, Profiler* prof) {
  const int L = (H / 2) * (W / 2);
  const size_t Lp = (size_t)ceil_div(L, kKT) * kKT;
  __half* x16 = (__half*)w.nl_x16;
  __half* gt16 = (__half*)((uint8_t*)w.nl_x16 + nl_x_bytes(N, Lp));
  float* y = (float*)w.nl_priv;
  float* part = (float*)((uint8_t*)w.nl_x16 + nl_x_bytes(N, Lp) + nl_g_bytes(N, Lp));
  int nl = 0;
  if (prof) prof->begin(kProfNonlocal, s);
  int rc = run_nl_tc(nullptr, lr, H, W, N, L, tc_nl_nsplit(tw.precision), x16, gt16, y, part, tw.num_sms, s, &nl);
  if (rc == PFNL_OK) rc = launch_nl_linear_scatter(y, lr, N, H, W, tw.raw.nl_gw_w, tw.raw.nl_gw_b, inp21, s);
  if (prof) prof->end(s);
  if (rc) return rc;
  *launches += nl + 1;
  return PFNL_OK;
}

int tc_nonlocal_tokens(const TcWeights& tw, const float* tokens, int N, int L, float* out, cudaStream_t s,
                       long long* launches, Profiler* prof) {
  // stage-level entry (pfnl_nonlocal): grow-only scratch owned by the handle (freed by tc_destroy)
  unsigned char*& scratch = tw.nl_scratch;
  size_t& cap = tw.nl_scratch_cap;
  const size_t Lp = (size_t)ceil_div(L, kKT) * kKT;
  const size_t b_x = nl_x_bytes(N, Lp);
  const size_t b_g = nl_g_bytes(N, Lp);
  const size_t b_y = ((size_t)N * L * kNL * 4 + 1023) / 1024 * 1024;
  const size_t b_p = (nl_part_floats(N, L, tw.num_sms) * 4 + 1023) / 1024 * 1024;
  if (b_x + b_g + b_y + b_p > cap) {
    if (scratch) {
      PFNL_CUDA(cudaDeviceSynchronize());
      PFNL_CUDA(cudaFree(scratch));
      scratch = nullptr;
      cap = 0;
    }
    PFNL_CUDA(cudaMalloc((void**)&scratch, b_x + b_g + b_y + b_p));
    cap = b_x + b_g + b_y + b_p;
  }
  __half* x16 = (__half*)scratch;
  __half* gt16 = (__half*)(scratch + b_x);
  float* y = (float*)(scratch + b_x + b_g);
  float* part = b_p ? (float*)(scratch + b_x + b_g + b_y) : nullptr;
  int nl = 0;
  int rc = run_nl_tc(tokens, nullptr, 0, 0, N, L, tc_nl_nsplit(tw.precision), x16, gt16, y, part, tw.num_sms, s, &nl,
                     prof);
  if (rc == PFNL_OK) rc = launch_nl_linear(y, N * L, tw.raw.nl_gw_w, tw.raw.nl_gw_b, out, s);
  if (rc) return rc;
  *launches += nl + 1;
  return PFNL_OK;
}

}  // namespace pfnl
