// fused_fwd.cu -- fused forward of the EGT attention block's N x N part on sm_100a tensor cores.
//
//   reference:  EGT.call_gated (lib/models/egt_layers.py:57-143) fused with the edge projections, the
//   LayerNorm on e and the edge write-back of edge_update_residual
//   (lib/models/graph_xformer_model_base.py:192-218).  The [B,N,N,h] tensors E, G, H_hat, A~ never
//   leave the SM: e is streamed in once by TMA, e' streamed out once by TMA.
//
// One CTA = one graph b and 128 query rows, 16 compute warps: thread (kq, g, t) owns query row l0+t ==
// TMEM lane t, heads 4g..4g+3, and the key pairs p with p % 2 == kq -- four resident compute warps per SM
// sub-partition.  Warp 16 issues TMA and tcgen05.mma (warps 17-19 complete its warpgroup so that setmaxnreg
// can move registers to the compute warps).
//
// Keys are processed in PAIRS, two pairs (one per kq) per pipeline step.  For pair p the tensor core
// produces, in TMEM (columns ordered (g,key,hh4)):
//     S  [128x16] = Qs [128x64] * Kexp^T      Kexp[(g,key,hh4), c] = K[key,c] * [c % 8 == hh]
//     EG [128x32] = e  [128x16] * Wblk        raw edge channels of the two keys x folded-LN weights
// (the per-head dot product is a block-diagonal contraction over the head-innermost channel axis, so it is a
// plain GEMM against an expanded K).  The row's threads apply the LN statistics, clip, masks, exp / sigmoid,
// accumulate the softmax denominator and the gate sum, and write A~ and H_hat back to TMEM as bf16
// A-operands of
//     O  [128x64] += A~ [128x16] * Vexp       Vexp[(g,key,hh4), c] = V[key,c] * [c % 8 == hh]
//     De [128x16]  = H_hat [128x16] * Wr blk  edge write-back of the two keys
// e' = e + De + b_r is formed in place over the e stage and leaves by TMA store.
// Softmax reference (tf.nn.softmax subtracts the row maximum, egt_layers.py:111): every logit satisfies
// |H_hat| <= bound (FusedPrep::bound = clip + sqrt(d_e) ||W'_E|| + |v|, known before the loop).  While the bound is
// below the fp32 exponent budget the reference 0 is exact; above it the MAXONLY pre-pass of this kernel leaves the row
// maximum in lse[0] and the main pass exponentiates H_hat - max.  lse[0] is what the backward adds to the log row sum.
//
// There is no CTA-wide barrier in the main loop: compute threads arrive on an mbarrier when their part of a
// step (4 keys) is done and the issuer waits for it.  The handshake compute -> issuer -> tensor core -> compute
// costs ~1700 cycles (measured), as much as a step's arithmetic, so every dependency spans TWO steps: the
// S/EG products of step it+2 are issued when step it is done, and the e' update of step it (which needs the
// H_hat W_r product issued when step it is done) runs during step it+2.
#include "common.cuh"
#include "fused.h"
#include "umma.cuh"

namespace egt {
using namespace umma;

namespace {

constexpr int NS = 4;                                  // input stages of 8 keys: e | K rows | V rows
constexpr uint32_t SM_Q = 0;                           // [128 x 128B] swizzled, Q pre-scaled by dk^-0.5
constexpr uint32_t SM_STAGE = 16384;
constexpr uint32_t ST_E = 0, ST_K = 16384, ST_V = 17408, STAGE_BYTES = 18432;
constexpr uint32_t SM_KVX = SM_STAGE + NS * STAGE_BYTES;       // 8 slots (pair p & 7) x (Kexp 2048 | Vexp 2048)
constexpr uint32_t SM_W = SM_KVX + 8 * 4096;                   // b_eg 1024 | b_wr 512 | b_eg_lo 1024
constexpr uint32_t SM_CONST = SM_W + 2560;                     // uE vE uG vG br (40 floats)
constexpr uint32_t SM_BAR = SM_CONST + 256;
constexpr uint32_t SM_MASK = SM_BAR + 256;                       // key-valid bytes, zero padded (N <= 4096)
constexpr uint32_t SM_WO = (SM_MASK + 4096 + 16 + 1023) & ~1023u;   // W_O operand image (MN-major, 8 KB) | b_O (64 floats)
constexpr uint32_t SM_TOTAL = SM_WO + 8192 + 256;

constexpr uint32_t TM_O = 0;
constexpr uint32_t TM_IN = 64, TM_IN_COLS = 96, TM_PAIR = 48;  // 2 buffers x 2 pairs: S 16 | EG 32   (step parity)
constexpr uint32_t IN_S = 0, IN_EG = 16;
constexpr uint32_t TM_DE = 256, TM_DE_COLS = 32;               // 2 buffers x 2 pairs: De 16             (step parity)
constexpr uint32_t TM_OUT = 320, TM_OUT_COLS = 32, TM_OPAIR = 16;   // 2 buffers x 2 pairs: A~ 8 | H_hat 8 (bf16 A operands)

constexpr uint32_t ID_N16 = idesc_bf16(128, 16, 0, 0);
constexpr uint32_t ID_N32 = idesc_bf16(128, 32, 0, 0);
constexpr uint32_t ID_PV = idesc_bf16(128, 64, 0, 1);

struct Bars { uint64_t q_full, e_full[NS], mma1[2], mma2[2], step[2], oproj; uint32_t tmem_base; };

}  // namespace

// MAXONLY: the row-maximum pre-pass of the softmax (same scheme as wide_fwd.cu).  tf.nn.softmax (egt_layers.py:111)
// subtracts the row maximum; the main pass needs its reference before the key loop.  While the data-independent bound of
// the logits (FusedPrep::bound) is below the fp32 exponent budget the reference 0 is exact and every CTA of the pre-pass
// returns at once; above it the pre-pass streams e once more (S and [E|G] products, logit arithmetic, no exp, no P V,
// no e') and leaves max_m H_hat[l, m, hh] over the live keys in lse[0].
template <bool RAND, bool MAXONLY>
__global__ void __launch_bounds__(640, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_eo,
                 const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                 const FusedFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // shared address space is kept
  const uint32_t sbase = smem_u32(smem);
  Bars *bars = (Bars *)(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, l0 = blockIdx.x * 128;
  const int N = a.N;
  const int NT = (N + 7) / 8, NQ = (N + 3) / 4;       // 8-key tiles, pipeline steps of 4 keys (two pairs)

  pdl_trigger();
  if (MAXONLY) {
    pdl_wait();
    if (a.prep->bound <= kSoftmaxBudget) return;       // the whole grid agrees: nothing to do
  }
  if (warp == 16) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars->q_full), 1);
      for (int i = 0; i < NS; ++i) mbar_init(smem_u32(&bars->e_full[i]), 1);
      for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->mma1[i]), 1); mbar_init(smem_u32(&bars->mma2[i]), 1); }
      // every compute thread arrives once per step; steps alternate between two barriers because a warp may
      // finish step it+1 before a slower one finishes step it (dependencies span two steps)
      mbar_init(smem_u32(&bars->step[0]), 512); mbar_init(smem_u32(&bars->step[1]), 512);
      mbar_init(smem_u32(&bars->oproj), 1);
      mbar_fence_init();
      tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_eo); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    pdl_wait();
  } else if (warp > 16) {
    if (a.w_o) {   // W_O (float32 [64,64], "x @ W") -> MN-major 128B-swizzled bf16 operand image; b_O
      const int u = tid - 17 * 32;                     // 0 .. 95
      for (int i = u; i < 64 * 8; i += 96) {
        const int k = i >> 3, n = (i & 7) << 3;
        const float4 w0 = *(const float4 *)(a.w_o + k * 64 + n), w1 = *(const float4 *)(a.w_o + k * 64 + n + 4);
        uint4 v;
        v.x = pack_bf16(w0.x, w0.y); v.y = pack_bf16(w0.z, w0.w); v.z = pack_bf16(w1.x, w1.y); v.w = pack_bf16(w1.z, w1.w);
        *(uint4 *)(smem + SM_WO + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u + ((uint32_t)(((n >> 3) ^ k) & 7) << 4)) = v;
      }
      if (u < 64) ((float *)(smem + SM_WO + 8192))[u] = a.b_o[u];
      fence_proxy_async_smem();
    }
  } else if (warp < 16) {
    pdl_wait();                                        // prep / qkv come from the preceding kernel
    if (tid < 64) ((uint4 *)(smem + SM_W))[tid] = ((const uint4 *)a.prep->b_eg)[tid];            // b_eg
    else if (tid < 96) ((uint4 *)(smem + SM_W + 1024))[tid - 64] = ((const uint4 *)a.prep->b_wr)[tid - 64];
    else if (tid < 160) ((uint4 *)(smem + SM_W + 1536))[tid - 96] = ((const uint4 *)a.prep->b_eg_lo)[tid - 96];
    if (tid < 40) ((float *)(smem + SM_CONST))[tid] = a.prep->uE[tid];                            // uE vE uG vG br
    for (int i = tid; i < 4 * NQ; i += 512)                                                        // key-valid bytes
      smem[SM_MASK + i] = i < N ? (a.mask ? (uint8_t)(a.mask[(size_t)blockIdx.y * N + i] != 0) : (uint8_t)1) : (uint8_t)0;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 16) {
    // ====================== issuer warpgroup (warp 16 issues; warps 17-19 only keep the barriers) ======
    reg_dealloc<32>();   // 512 x 112 + 128 x 32 == 640 x 96: setmaxnreg only moves registers inside the CTA's launch allocation
    const bool leader = warp == 16 && lane == 0;
    auto load_tile = [&](int T) {
      const int st = T % NS;
      const uint32_t bar = smem_u32(&bars->e_full[st]);
      const uint32_t dst = sbase + SM_STAGE + st * STAGE_BYTES;
      mbar_expect_tx(bar, STAGE_BYTES);
      tma_load_3d(dst + ST_E, &tm_e, bar, T * 64, l0, b);
      tma_load_3d(dst + ST_K, &tm_kv, bar, FD, T * 8, b);
      tma_load_3d(dst + ST_V, &tm_kv, bar, 2 * FD, T * 8, b);
    };
    // descriptor low words (address | LBO); the high words are compile-time constants
    constexpr uint32_t HI_SW = desc_hi(1024, LAYOUT_SW128), HI_NONE = desc_hi(128, LAYOUT_NONE);
    const uint32_t loQ = desc_lo(sbase + SM_Q, 16), loK = desc_lo(sbase + SM_KVX, 16);
    const uint32_t loV = desc_lo(sbase + SM_KVX + 2048, 2048), loE = desc_lo(sbase + SM_STAGE + ST_E, 16);
    const uint32_t loWeg = desc_lo(sbase + SM_W, 512), loWr = desc_lo(sbase + SM_W + 1024, 256);
    const uint32_t loWegLo = desc_lo(sbase + SM_W + 1536, 512);
    const bool use_lo = a.prep->use_lo != 0;           // W' = hi + lo only when the logits are large (fused.h)
    // tcgen05.mma is issued warp-collectively by warp 16 (umma.cuh: converged warp, one elected lane, k-chains in one
    // asm statement): issuing from inside `if (lane == 0)` costs > 100 cycles per instruction and was most of the
    // ~1700-cycle handshake measured in round 1.  TMA stays with lane 0.
    auto issue_mma1_pair = [&](int p, int st, int buf) {
      const int j = p & 3, slot = p & 7;
      const uint32_t d = tmem + TM_IN + buf * TM_IN_COLS + (p & 1) * TM_PAIR;
      MmaChain<4>::ss(d + IN_S, loQ, HI_SW, loK + slot * 256, HI_SW, ID_N16, 0, 2, 2);   // 4096 B per slot
      MmaChain<1>::ss(d + IN_EG, loE + st * (STAGE_BYTES / 16) + 2 * j, HI_SW, loWeg, HI_NONE, ID_N32, 0, 0, 0);
      if (use_lo) MmaChain<1>::ss(d + IN_EG, loE + st * (STAGE_BYTES / 16) + 2 * j, HI_SW, loWegLo, HI_NONE, ID_N32, 1, 0, 0);
    };
    const uint32_t bar_e0 = smem_u32(&bars->e_full[0]), bar_m1 = smem_u32(&bars->mma1[0]), bar_m2 = smem_u32(&bars->mma2[0]);
    auto issue_mma1 = [&](int q) {                     // both pairs of step q -> input buffer q & 1
      const int T = q >> 1, st = T % NS, buf = q & 1;
      mbar_wait(bar_e0 + 8 * st, (T / NS) & 1);
      tc_fence_after();
      issue_mma1_pair(2 * q, st, buf);
      issue_mma1_pair(2 * q + 1, st, buf);
    };
    auto issue_mma2 = [&](int q) {                     // O += A~ Vexp ; De = H_hat Wr  (operands / result: parity q & 1)
      const int ob = q & 1;
#pragma unroll
      for (int kq = 0; kq < 2; ++kq) {
        const int p = 2 * q + kq;
        const uint32_t ao = tmem + TM_OUT + ob * TM_OUT_COLS + kq * TM_OPAIR;
        MmaChain<1>::ts(tmem + TM_O, ao, loV + (p & 7) * 256, HI_SW, ID_PV, p > 0, 0, 0);
        MmaChain<1>::ts(tmem + TM_DE + ob * TM_DE_COLS + kq * 16, ao + 8, loWr, HI_NONE, ID_N16, 0, 0, 0);
      }
    };
    if (leader) {
      mbar_expect_tx(smem_u32(&bars->q_full), 16384);
      tma_load_3d(sbase + SM_Q, &tm_q, smem_u32(&bars->q_full), 0, l0, b);
      for (int T = 0; T < NT && T < NS; ++T) load_tile(T);
    }
    __syncthreads();                                   // sync #0: Kexp/Vexp of steps 0, 1 are built
    if (warp == 16) {
      tc_fence_after();
      mbar_wait(smem_u32(&bars->q_full), 0);
      issue_mma1(0);
      mma_commit_w(bar_m1);                            // S / EG of step 0
      if (NQ > 1) { issue_mma1(1); mma_commit_w(bar_m1 + 8); }
    }
    const uint32_t bar_step = smem_u32(&bars->step[0]);
    int next_store = 0;                                // tiles [0, next_store) have been handed to the TMA store
    for (int it = 0; it < NQ && warp == 16; ++it) {    // warps 17-19 go straight to the tail barrier
      mbar_wait(bar_step + 8 * (it & 1), (it >> 1) & 1);   // all compute threads finished step it
      tc_fence_after();
      fence_proxy_async_smem();                        // the compute threads' shared-memory writes of step it
      if (!MAXONLY) issue_mma2(it);                    //  (ordered before this point by the mbarrier) -> async proxy
      if (it + 2 < NQ) issue_mma1(it + 2);
      mma_commit_w(bar_m2 + 8 * (it & 1));             // ONE completion per handshake: products of step it and
                                                       // S / EG of step it+2, both consumed during step it+2
      if (lane == 0) {
        // step it contained the e' update of step it-2; tile T (steps 2T, 2T+1) is complete when it == 2T+3
        if (it >= 4 && (it & 1) == 0) {                // tile stored at the previous handshake: recycle its stage
          const int T = (it - 4) >> 1;
          if (!MAXONLY) tma_store_wait_read<0>();
          if (T + NS < NT) load_tile(T + NS);
        }
        if (it >= 3 && (it & 1) == 1) {
          const int T = (it - 3) >> 1;
          if (!MAXONLY) {
            tma_store_3d(&tm_eo, sbase + SM_STAGE + (T % NS) * STAGE_BYTES + ST_E, T * 64, l0, b);
            tma_store_commit();
          }
          next_store = T + 1;
        }
      }
      __syncwarp();
    }
    __syncthreads();                                   // sync #(NQ+1): every e' update is done
    if (leader && !MAXONLY) {
      for (int T = next_store; T < NT; ++T)
        tma_store_3d(&tm_eo, sbase + SM_STAGE + (T % NS) * STAGE_BYTES + ST_E, T * 64, l0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
    __syncthreads();                                   // partial row sums exchanged
    if (a.w_o) {
      __syncthreads();                                 // V_att tile staged (over the Q tile)
      if (warp == 16) {
        tc_fence_after();
        constexpr uint32_t ID_OP = idesc_bf16(128, 64, 0, 1);
        MmaChain<4>::ss(tmem + TM_IN, desc_lo(sbase + SM_Q, 16), HI_SW, desc_lo(sbase + SM_WO, 8192), HI_SW, ID_OP, 0, 2, 128);
        mma_commit_w(smem_u32(&bars->oproj));
      }
      __syncwarp();
    }
    __syncthreads();                                   // final
    if (warp == 16) tmem_dealloc(tmem, 512);
    return;
  }

  // ================================= compute threads (warps 0-15) ================================
  reg_alloc<112>();
  const int kq = tid >> 8, g = (tid >> 7) & 1, t = tid & 127;
  const int l = l0 + t;
  const bool rowvalid = l < N;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float *cst = (const float *)(smem + SM_CONST);
  const uint8_t *smask = smem + SM_MASK;
  const uint32_t bar_mma1 = smem_u32(&bars->mma1[0]), bar_mma2 = smem_u32(&bars->mma2[0]);
  const uint32_t bar_e = smem_u32(&bars->e_full[0]), bar_step = smem_u32(&bars->step[0]);
  const float lo = a.clip_lo, hi = a.clip_hi;
  float psum[4], gsum[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { psum[i] = 0.f; gsum[i] = 0.f; }
  // exponent reference of the softmax: 0 while the logits cannot overflow, else the row maximum of the pre-pass
  const bool use_ref = !MAXONLY && a.prep->bound > kSoftmaxBudget;
  const float ln_eps = a.ln_eps;
  float nsh[4] = {0.f, 0.f, 0.f, 0.f};                 // -ref * log2(e) of this thread's heads
  if (use_ref && rowvalid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) nsh[i] = -a.lse[((size_t)b * N + l) * FH + 4 * g + i] * kLog2e;
  }
  if (MAXONLY) {
#pragma unroll
    for (int i = 0; i < 4; ++i) psum[i] = -INFINITY;   // psum holds the running maxima
  }
  const uint32_t trow = (uint32_t)t * 128u, tx7 = (uint32_t)(t & 7);

  // expanded K / V operands of pair p2 (stage st2) into slot p2 & 3: one 16-byte chunk per thread
  // (thread (kq, g, .) builds the K (g = 0) / V (g = 1) chunk of pair 2q + kq)
  const int b_n = (tid & 127) >> 3, b_dd = tid & 7, b_hh = 4 * (b_n >> 3) + (b_n & 3);
  const uint32_t b_src = (uint32_t)(g ? ST_V : ST_K) + (uint32_t)((b_n >> 2) & 1) * 128u + (uint32_t)(b_dd * 8 + b_hh) * 2u;
  const uint32_t b_dst = SM_KVX + (uint32_t)g * 2048u + (uint32_t)b_n * 128u + ((uint32_t)((b_dd ^ b_n) & 7) << 4);
  auto build = [&](int p2, int st2) {
    const uint32_t val = *(const uint16_t *)(smem + SM_STAGE + st2 * STAGE_BYTES + b_src + (p2 & 3) * 256);
    const uint32_t wv = val << ((b_hh & 1) * 16);
    uint4 ch;
    ch.x = (b_hh >> 1) == 0 ? wv : 0u; ch.y = (b_hh >> 1) == 1 ? wv : 0u;
    ch.z = (b_hh >> 1) == 2 ? wv : 0u; ch.w = (b_hh >> 1) == 3 ? wv : 0u;
    *(uint4 *)(smem + b_dst + (p2 & 7) * 4096) = ch;
  };

  // ---- phase A: pair p, this thread's 4 heads of both keys ------------------------------------------
  auto phase_a = [&](int p, int st) {
    const int j = p & 3, ob = (p >> 1) & 1;
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
    const uint32_t tin = tlane + TM_IN + ob * TM_IN_COLS + kq * TM_PAIR;
    const uint32_t tout = tlane + TM_OUT + ob * TM_OUT_COLS + kq * TM_OPAIR;
    uint32_t sreg[8], egreg[16];
    tmem_ld8(tin + IN_S + g * 8, sreg);
    tmem_ld16(tin + IN_EG + g * 16, egreg);
    float r[2], nrm[2];
    bool kvalid[2];
    uint32_t rb[2][2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ks = 2 * j + kk, m = 2 * p + kk;
      const uint4 ev = *(const uint4 *)(es + ST_E + trow + (((uint32_t)ks ^ tx7) << 4));
      const float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y),
                          bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
      float mu = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
      mu *= 0.125f;
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float dlt = x[c] - mu; var = fmaf(dlt, dlt, var); }
      r[kk] = rsqrtf(fmaf(var, 0.125f, ln_eps));
      nrm[kk] = -r[kk] * mu;
      kvalid[kk] = smask[m] != 0;
      rb[kk][0] = rb[kk][1] = 0u;
    }
    if (RAND) {   // one Philox call = this thread's 2 keys x 4 heads (rng_elem_index, common.cuh)
      const uint64_t qd = rng_elem_index((uint64_t)b, (uint64_t)l, (uint64_t)(2 * p), 4u * (uint32_t)g, (uint64_t)N, FH) >> 3;
      const uint64_t roff = a.offset + (a.offset_dev ? *a.offset_dev : 0ull);
      const Philox4 ph = philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), (uint32_t)roff,
                                       (uint32_t)(roff >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      rb[0][0] = ph.x; rb[0][1] = ph.y; rb[1][0] = ph.z; rb[1][1] = ph.w;   // key kk, head 4g+i: 16-bit lane 4kk+i
    }
    const float4 uE4 = *(const float4 *)(cst + 4 * g), vE4 = *(const float4 *)(cst + 8 + 4 * g);
    const float4 uG4 = *(const float4 *)(cst + 16 + 4 * g), vG4 = *(const float4 *)(cst + 24 + 4 * g);
    const float uE[4] = {uE4.x, uE4.y, uE4.z, uE4.w}, vE[4] = {vE4.x, vE4.y, vE4.z, vE4.w};
    const float uG[4] = {uG4.x, uG4.y, uG4.z, uG4.w}, vG[4] = {vG4.x, vG4.y, vG4.z, vG4.w};
    tmem_ld_wait();
    uint32_t apack[4], hpack[4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      float av[4], hv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float S = __uint_as_float(sreg[kk * 4 + i]);
        const float E = fmaf(r[kk], __uint_as_float(egreg[kk * 8 + i]), fmaf(nrm[kk], uE[i], vE[i]));
        const float G = fmaf(r[kk], __uint_as_float(egreg[kk * 8 + 4 + i]), fmaf(nrm[kk], uG[i], vG[i]));
        const float Hh = fminf(fmaxf(S, lo), hi) + E;                      // egt_layers.py:79-86
        bool live = kvalid[kk];
        if (RAND) {
          const uint32_t w = rb[kk][i >> 1];
          const uint32_t bits = (i & 1) ? (w >> 16) : (w & 0xFFFFu);
          live = live && !(bits < a.rand_thr);                             // :103-108
        }
        if (MAXONLY) {                                                     // row maximum over the live keys only
          psum[i] = fmaxf(psum[i], live ? Hh : -INFINITY);
          av[i] = hv[i] = G;                                               // (unused)
          continue;
        }
        const float pr = live ? ex2_approx(fmaf(Hh, kLog2e, nsh[i])) : 0.f;   // :111 (unnormalised, relative to ref)
        const float gg = live ? sigmoid_fast(G) : 0.f;                     // :112
        psum[i] += pr;
        gsum[i] += gg;
        av[i] = pr * gg;                                                   // :113
        hv[i] = Hh;
      }
      apack[kk * 2 + 0] = pack_bf16(av[0], av[1]); apack[kk * 2 + 1] = pack_bf16(av[2], av[3]);
      hpack[kk * 2 + 0] = pack_bf16(hv[0], hv[1]); hpack[kk * 2 + 1] = pack_bf16(hv[2], hv[3]);
    }
    if (!MAXONLY) {
      tmem_st4(tout + g * 4, apack);
      tmem_st4(tout + 8 + g * 4, hpack);
    }
  };

  // ---- phase B: e' = e + H_hat W_r + b_r for key g of pair p, in place over the e stage ---------------
  auto phase_b = [&](int p, int st) {
    const int j = p & 3, ob = (p >> 1) & 1;
    uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
    const int ks = 2 * j + g;
    uint32_t dr[8];
    tmem_ld8(tlane + TM_DE + ob * TM_DE_COLS + kq * 16 + g * 8, dr);
    uint4 *pe = (uint4 *)(es + ST_E + trow + (((uint32_t)ks ^ tx7) << 4));
    const uint4 ev = *pe;
    const float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y),
                        bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
    tmem_ld_wait();
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] = x[c] + (__uint_as_float(dr[c]) + cst[32 + c]);
    uint4 ov;
    ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]);
    ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
    *pe = ov;
  };

  // ---- pipeline ------------------------------------------------------------------------------------
  // step it covers pairs 2*it (kq = 0) and 2*it + 1 (kq = 1), both in tile it >> 1; running indices
  // instead of divisions: step it -> (stage st_a, TMEM buffer buf_a, parity par_a)
  mbar_wait(bar_e, 0);
  build(kq, 0);
  if (NQ > 1) build(2 + kq, 0);
  fence_proxy_async_smem();
  __syncthreads();                                     // sync #0
  int st_a = 0;                                        // stage of step it
  int st_b = 0;                                        // stage of step it - 2
  int st_n = 1 % NS, par_n = 0;                        // stage / load parity of step it + 2
  for (int it = 0; it < NQ; ++it) {
    const uint32_t ob8 = 8u * (uint32_t)(it & 1);
    if (it >= 2) mbar_wait(bar_mma2 + ob8, ((it - 2) >> 1) & 1);   // handshake it-2: S / EG of this step, products of step it-2
    else mbar_wait(bar_mma1 + ob8, 0);                 // steps 0, 1: issued before the loop
    tc_fence_after();
    phase_a(2 * it + kq, st_a);
    if (it >= 2 && !MAXONLY) phase_b(2 * (it - 2) + kq, st_b);
    if (it + 2 < NQ) {
      if ((it & 1) == 0) mbar_wait(bar_e + 8 * st_n, par_n);        // first step of tile (it+2)>>1
      build(2 * (it + 2) + kq, st_n);
    }
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_step + ob8);                       // step it done by this thread
    if (it & 1) {                                      // steps it+1 and it+3 open new tiles
      if (++st_a == NS) st_a = 0;
      if (++st_n == NS) { st_n = 0; par_n ^= 1; }
      if (it >= 3) { if (++st_b == NS) st_b = 0; }     // step it-1 opens a new tile
    }
  }
  for (int q = (NQ >= 2 ? NQ - 2 : 0); q < NQ; ++q) {  // e' updates of the last two steps
    mbar_wait(bar_mma2 + 8 * (q & 1), (q >> 1) & 1);
    tc_fence_after();
    if (!MAXONLY) phase_b(2 * q + kq, (q >> 1) % NS);
  }
  fence_proxy_async_smem();
  __syncthreads();                                     // sync #(NQ+1)

  // ---- row epilogue: normalise, centrality scaler, V_att, saved statistics ----------------------
  {
    // all tcgen05.mma have completed (the last mma2 commit was waited for): the Kexp/Vexp slots are free
    float *xch = (float *)(smem + SM_KVX) + ((kq * 2 + g) * 128 + t) * 8;
    *(float4 *)xch = make_float4(psum[0], psum[1], psum[2], psum[3]);
    *(float4 *)(xch + 4) = make_float4(gsum[0], gsum[1], gsum[2], gsum[3]);
    __syncthreads();                                   // partial row sums exchanged
    if (MAXONLY) {   // row maxima over both key-pair parities -> lse[0]; a row without a live key keeps the reference 0
      const float *oth = (const float *)(smem + SM_KVX) + (((kq ^ 1) * 2 + g) * 128 + t) * 8;
      const float4 p4 = *(const float4 *)oth;
      const float mx[4] = {fmaxf(psum[0], p4.x), fmaxf(psum[1], p4.y), fmaxf(psum[2], p4.z), fmaxf(psum[3], p4.w)};
      if (kq == 0 && rowvalid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a.lse[((size_t)b * N + l) * FH + 4 * g + i] = mx[i] > -INFINITY ? mx[i] : 0.f;
      }
    } else {
    {
      const float *oth = (const float *)(smem + SM_KVX) + (((kq ^ 1) * 2 + g) * 128 + t) * 8;
      const float4 p4 = *(const float4 *)oth, g4 = *(const float4 *)(oth + 4);
      psum[0] += p4.x; psum[1] += p4.y; psum[2] += p4.z; psum[3] += p4.w;
      gsum[0] += g4.x; gsum[1] += g4.y; gsum[2] += g4.z; gsum[3] += g4.w;
    }
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float inv = psum[i] > 0.f ? __fdividef(1.f, psum[i]) : 0.f;
      float s = 1.f;
      if (a.scale_degree && l >= a.num_virtual_nodes)                      // egt_layers.py:123-135
        s = a.scaler_type == EGT_SCALER_LOG ? log1pf(gsum[i]) : gsum[i];
      f[i] = inv * s;
    }
    uint32_t o[32];
    tmem_ld32(tlane + TM_O + 32 * kq, o);              // this thread stores channel groups dd = 4kq .. 4kq+3
    tmem_ld_wait();
    if (rowvalid) {
      uint2 *dst = (uint2 *)(a.v_att + ((size_t)b * N + l) * FD + 32 * kq + 4 * g);
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lo4 = __uint_as_float(o[dd * 8 + i]), hi4 = __uint_as_float(o[dd * 8 + 4 + i]);
          v[i] = (g ? hi4 : lo4) * f[i];
        }
        dst[dd * 2] = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));   // channels dd*8 + 4g .. +3
      }
      if (a.w_o) {   // same values into the A tile of the output projection (rows l >= N stay zero: TMA OOB fill of Q)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float lo4 = __uint_as_float(o[dd * 8 + i]), hi4 = __uint_as_float(o[dd * 8 + 4 + i]);
            v[i] = (g ? hi4 : lo4) * f[i];
          }
          *(uint2 *)(smem + SM_Q + trow + (((uint32_t)(4 * kq + dd) ^ tx7) << 4) + 8 * g) =
              make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
        }
      }
      if (kq == 0) {
        const size_t ps = ((size_t)b * N + l) * FH + 4 * g, rs = (size_t)a.B * N * FH;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!use_ref) a.lse[ps + i] = 0.f;                               // reference point of the exponent (else: the pre-pass' row maximum)
          a.lse[rs + ps + i] = psum[i] > 0.f ? __logf(psum[i]) : 0.f;
          a.deg[ps + i] = gsum[i];
        }
      }
    }
    }   // !MAXONLY
  }
  if (a.w_o) {   // h' = h + V_att W_O + b_O  (graph_xformer_model_base.py:136-140)
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                                   // V_att tile staged
    mbar_wait(smem_u32(&bars->oproj), 0);
    tc_fence_after();
    uint32_t o[16];
    const int c0 = 16 * (2 * kq + g);                  // this thread's 16 output channels of row l
    tmem_ld16(tlane + TM_IN + c0, o);
    tmem_ld_wait();
    if (rowvalid) {
      const float *bo = (const float *)(smem + SM_WO + 8192) + c0;
      const uint4 *hp = (const uint4 *)(a.h + ((size_t)b * N + l) * FD + c0);
      uint4 *op = (uint4 *)(a.h_out + ((size_t)b * N + l) * FD + c0);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint4 hv = hp[q];
        const float hr[8] = {bf16_lo(hv.x), bf16_hi(hv.x), bf16_lo(hv.y), bf16_hi(hv.y),
                             bf16_lo(hv.z), bf16_hi(hv.z), bf16_lo(hv.w), bf16_hi(hv.w)};
        float y[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) y[c] = __uint_as_float(o[8 * q + c]) + bo[8 * q + c] + hr[c];
        uint4 ov;
        ov.x = pack_bf16(y[0], y[1]); ov.y = pack_bf16(y[2], y[3]); ov.z = pack_bf16(y[4], y[5]); ov.w = pack_bf16(y[6], y[7]);
        op[q] = ov;
      }
    }
  }
  tc_fence_before();
  __syncthreads();                                     // final
}

int fused_fwd_launch(const FusedFwdArgs &a, const void *e, void *e_out, const void *qkv, cudaStream_t st) {
  CUtensorMap tm_e, tm_eo, tm_q, tm_kv;
  const uint64_t N = a.N, B = a.B;
  int rc;
  if ((rc = encode_tmap_3d(&tm_e, e, N * FDE, N, B, N * FDE * 2, N * N * FDE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_eo, e_out, N * FDE, N, B, N * FDE * 2, N * N * FDE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_q, qkv, 3 * FD, N, B, 3 * FD * 2, N * 3 * FD * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_kv, qkv, 3 * FD, N, B, 3 * FD * 2, N * 3 * FD * 2, 64, 8, 1, 0))) return rc;
  const int smem = SM_TOTAL + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.B);
  {   // row-maximum pre-pass: every CTA returns at once unless the logit bound exceeds the exponent budget
    FusedFwdArgs m = a;
    m.w_o = nullptr;
    LaunchScope _ls("fused_fwd_rowmax_kernel", st);
    if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(fused_fwd_kernel<true, true>, grid, dim3(640), smem, st, tm_e, tm_eo, tm_q, tm_kv, m));
    else EGT_CHECK_CUDA(launch_pdl(fused_fwd_kernel<false, true>, grid, dim3(640), smem, st, tm_e, tm_eo, tm_q, tm_kv, m));
  }
  LaunchScope _ls("fused_fwd_kernel", st);
  if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(fused_fwd_kernel<true, false>, grid, dim3(640), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
  else EGT_CHECK_CUDA(launch_pdl(fused_fwd_kernel<false, false>, grid, dim3(640), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
  return EGT_OK;
}

}  // namespace egt
