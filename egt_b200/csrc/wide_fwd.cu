// wide_fwd.cu -- width-generic fused forward of the EGT attention block's N x N part on sm_100a tensor cores.
//
//   reference:  EGT.call_gated (lib/models/egt_layers.py:57-143) fused with the edge projections, the LayerNorm on e
//   and the edge write-back of edge_update_residual (lib/models/graph_xformer_model_base.py:192-218).  The
//   [B,N,N,h] tensors E, G, H_hat, A~ never leave the SM: e is streamed in once by TMA, e' streamed out once by TMA.
//
// One CTA = one graph b and 128 query rows (TMEM lane = row).  NG compute warpgroups; thread (q, t) of group q owns
// query row l0+t for the keys m = q (mod NG): ALL heads of one (row, key) pair, so the LayerNorm statistics of a
// pair are computed once and nothing about a pair is split across threads.  Two more warps: the tcgen05.mma issuer
// and the TMA producer.
//
// Per key the tensor core produces, in the group's private TMEM columns,
//     S  [128 x 16]   = Qs [128 x d] * Kexp^T        Kexp[hh, c] = K[key, c] * [c % h == hh]   (block-diagonal trick:
//                                                    the per-head dot product over the head-innermost channel axis)
//     EG [128 x 2h]   = e_key [128 x d_e] * W'       raw edge channels x LayerNorm-folded weights (wide.h)
// the pair's thread applies the LN statistics, clip, masks, exp / sigmoid, adds to its softmax denominator and gate
// sum, and writes A~ and H_hat back to TMEM as bf16 A operands of
//     O  [128 x d]   += A~ [128 x 16] * Vexp         Vexp[hh, c] = V[key, c] * [c % h == hh]
//     e' [128 x d_e]  = e_key * I + H_hat * W_r + 1 * b_r      (identity and bias as tensor-core operands, so the
//                                                    thread only converts e' to bf16 over the e stage; TMA stores it)
//
// Pipeline: every group runs its OWN handshake with the issuer (mbarriers ready[q] / done[q]); groups drift apart, so
// while one waits for the round trip compute -> issuer -> tensor core -> compute the other groups of the same SM
// sub-partition fill the issue slots.  A tile of TK keys lives in one of NS shared-memory stages; the producer warp
// stores a tile and refills its stage as soon as every group has written its e' (mbarrier tile_done[stage]).
#include "common.cuh"
#include "wide.h"
#include "wide_common.cuh"

namespace egt {
using namespace umma;

namespace {

// MAXONLY: the row-maximum pre-pass of the softmax.  tf.nn.softmax (egt_layers.py:111) subtracts the row maximum; the
// main pass instead exponentiates H_hat - ref with a reference that must be known BEFORE the key loop.  While the
// data-independent bound of the logits (WidePrep::bound) is below the exponent budget, ref = 0 is exact in fp32 and
// the pre-pass returns at once; above it, the pre-pass streams e once more (S and [E|G] products and the logit
// arithmetic only -- no exp, no P V, no e' write) and leaves max_m H_hat[l, m, hh] over the live keys in lse[0].
template <class C, bool RAND, bool MAXONLY>
__global__ void __launch_bounds__(C::THREADS, 1)
wide_fwd_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_eo,
                const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                const WideFwdArgs a) {
  constexpr int H = C::H, DE = C::DE, D = C::D, NG = C::NG, NS = C::NS, TK = C::TK, KPG = C::KPG;
  // ---- shared memory map (bytes from the 1024-aligned base) ----
  constexpr uint32_t SM_Q = 0;                                         // NQA atoms [128 x 128 B], 128B swizzle
  constexpr uint32_t KV_MAT = C::NQA * 2048;                           // one expanded 16 x d operand
  constexpr uint32_t SM_KVX = SM_Q + C::NQA * 16384;                   // per group: Kexp | Vexp slot 0 | Vexp slot 1
  constexpr uint32_t SM_STAGE = SM_KVX + NG * 3 * KV_MAT;
  constexpr uint32_t ST_E = 0, ST_K = C::NBOX * 16384, ST_V = ST_K + C::KV_ROWS;
  constexpr uint32_t STAGE_BYTES = (ST_V + C::KV_ROWS + 1023) & ~1023u;
  constexpr uint32_t TX_BYTES = C::NBOX * 16384 + 2 * TK * D * 2;
  constexpr uint32_t SM_W = SM_STAGE + NS * STAGE_BYTES;               // w_eg[2] | w_r | b_r | i16[2]
  constexpr uint32_t W_EG = 0, W_EG_SZ = 2 * C::DEW * C::EGN * 2, W_R = 2 * W_EG_SZ, W_R_SZ = C::DEP * 32;
  constexpr uint32_t W_B = W_R + W_R_SZ, W_I = W_B + W_R_SZ, W_TOTAL = W_I + 1024;
  constexpr uint32_t SM_ONES = SM_W + W_TOTAL;                         // 4 KB of bf16 1.0 (A operand of the bias product)
  constexpr uint32_t SM_CONST = SM_ONES + 4096;                        // uE vE uG vG (4 x 16 floats)
  constexpr uint32_t SM_BAR = SM_CONST + 256;
  constexpr uint32_t SM_MASK = SM_BAR + 256;                           // key-valid bytes, zero padded
  constexpr uint32_t SM_REF = SM_MASK + 4096 + 64;                     // -ref * log2(e) per (row, head), float32
  constexpr uint32_t SM_TOTAL = SM_REF + 128 * H * 4;
  static_assert(SM_TOTAL + 1024 <= 232448, "shared memory budget");
  static_assert(SM_TOTAL == C::FWD_SMEM, "host-side shared-memory size (wide_common.cuh) out of date");
  static_assert(NG * 128 * 2 * H * 4 <= SM_STAGE, "row-sum exchange must fit below the stages");
  // ---- tensor memory map (columns) ----
  // Q lives in tensor memory as the A operand of Q K^T (no shared-memory A read per key: a [128 x 16] shared-memory A
  // tile is 4 KB per tcgen05.mma whatever N is, and with one S product per KEY that streaming was most of the kernel).
  // A~ and H_hat are written over the S / EG columns the thread has just consumed; the issuer orders the product
  // that reads them before the next key's S / EG by issue order.
  constexpr uint32_t TM_O = 0, TM_Q = D, TM_G = D + D / 2, G_S = 0, G_EG = 16, G_A = 0, G_H = 16, G_EO = 16 + C::EGN;
  constexpr uint32_t GC = 16 + C::EGN + C::DEP;
  static_assert(TM_G + NG * GC <= 512, "tensor memory budget");
  struct Bars { uint64_t q_full, e_full[NS], tile_done[NS], ready[NG], done[NG]; uint32_t tmem_base; };
  static_assert(sizeof(Bars) <= 256, "barrier block");

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  Bars *bars = (Bars *)(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, l0 = blockIdx.x * 128;
  const int N = a.N;
  const int NT = (N + TK - 1) / TK, J = NT * KPG;        // tiles; keys per group (padded keys are masked)
  if (MAXONLY) {
    pdl_wait();
    if (a.prep->bound <= kWideSoftmaxBudget) return;     // the whole grid agrees: nothing to do
  }

  // ---------------------------------------- set-up ----------------------------------------
  if (warp == 4 * NG) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars->q_full), 1);
      for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&bars->e_full[i]), 1); mbar_init(smem_u32(&bars->tile_done[i]), NG * 128); }
      for (int i = 0; i < NG; ++i) { mbar_init(smem_u32(&bars->ready[i]), 1); mbar_init(smem_u32(&bars->done[i]), 128); }
      mbar_fence_init();
      tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_eo); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
  }
  pdl_wait();                                            // prep / qkv come from the preceding kernels
  {
    const int nthr = C::THREADS;
    for (int i = tid; i < (int)(NG * 3 * KV_MAT / 16); i += nthr) ((uint4 *)(smem + SM_KVX))[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 4096 / 16; i += nthr) ((uint4 *)(smem + SM_ONES))[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    const WidePrep *pp = a.prep;
    for (int i = tid; i < (int)(W_EG_SZ / 16); i += nthr) {
      ((uint4 *)(smem + SM_W + W_EG))[i] = ((const uint4 *)pp->w_eg[0])[i];
      ((uint4 *)(smem + SM_W + W_EG + W_EG_SZ))[i] = ((const uint4 *)pp->w_eg[1])[i];
    }
    for (int i = tid; i < (int)(W_R_SZ / 16); i += nthr) {
      ((uint4 *)(smem + SM_W + W_R))[i] = ((const uint4 *)pp->w_r)[i];
      ((uint4 *)(smem + SM_W + W_B))[i] = ((const uint4 *)pp->b_r)[i];
    }
    for (int i = tid; i < 64; i += nthr) ((uint4 *)(smem + SM_W + W_I))[i] = ((const uint4 *)pp->i16[0])[i];   // both variants
    for (int i = tid; i < 64; i += nthr) ((float *)(smem + SM_CONST))[i] = pp->uE[i];                           // uE vE uG vG
    for (int i = tid; i < NT * TK; i += nthr)
      smem[SM_MASK + i] = i < N ? (a.mask ? (uint8_t)(a.mask[(size_t)b * N + i] != 0) : (uint8_t)1) : (uint8_t)0;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();                                       // sync A
  tc_fence_after();
  // A 512-column allocation is the whole tensor memory of the SM: its base is column 0, lane 0.  The issuer uses
  // the literal so that every tcgen05.mma operand is a warp-uniform value for ptxas (a value loaded from shared
  // memory is not, and costs vector -> uniform register moves per instruction).
  if (bars->tmem_base != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  const uint32_t bar_e0 = smem_u32(&bars->e_full[0]), bar_td0 = smem_u32(&bars->tile_done[0]);
  const uint32_t bar_ready0 = smem_u32(&bars->ready[0]), bar_done0 = smem_u32(&bars->done[0]);

  if (warp >= 4 * NG) {
    if (C::USE_SETMAXNREG) reg_dealloc<C::REG_HELPER>();
    auto load_tile = [&](int T) {
      const int st = T % NS;
      const uint32_t bar = bar_e0 + 8 * st, dst = sbase + SM_STAGE + st * STAGE_BYTES;
      mbar_expect_tx(bar, TX_BYTES);
#pragma unroll
      for (int x = 0; x < C::NBOX; ++x) tma_load_3d(dst + ST_E + x * 16384, &tm_e, bar, T * TK * DE + 64 * x, l0, b);
      tma_load_3d(dst + ST_K, &tm_kv, bar, D, T * TK, b);
      tma_load_3d(dst + ST_V, &tm_kv, bar, 2 * D, T * TK, b);
    };
    if (warp == 4 * NG + 1 && lane == 0) {               // TMA producer: Q tile and the first NS key tiles
      mbar_expect_tx(smem_u32(&bars->q_full), C::NQA * 16384);
#pragma unroll
      for (int x = 0; x < C::NQA; ++x) tma_load_3d(sbase + SM_Q + x * 16384, &tm_q, smem_u32(&bars->q_full), 64 * x, l0, b);
      for (int T = 0; T < NT && T < NS; ++T) load_tile(T);
    }
    __syncthreads();                                     // sync B: first expanded operands are built
    // The code can run TWO issuer warps (warp 4NG: even compute groups, warp 4NG+2: odd ones; O starts zeroed and every
    // A~ V product accumulates).  Measured with four groups: C5 forward -5 %, C3 -9 %; with two groups +8-10 %.  It is
    // switched OFF: with two issuers the order in which the groups' products are added into O varies from run to run,
    // so h' is no longer bit-reproducible (tests/test_parity_gpu.py checks permutation equivariance bit for bit).
    constexpr int NI = 1;
    if (warp == 4 * NG || (NI == 2 && warp == 4 * NG + 2)) {
      const int q0 = warp == 4 * NG ? 0 : 1;
      // ====== tcgen05.mma issuer: the whole (converged) warp runs the loop, one elected lane issues (umma.cuh) ======
      constexpr uint32_t HI_SW = desc_hi(1024, LAYOUT_SW128), HI_NONE = desc_hi(128, LAYOUT_NONE);
      constexpr uint32_t ID_S = idesc_bf16(128, 16, 0, 0), ID_EG = idesc_bf16(128, C::EGN, 0, 0);
      constexpr uint32_t ID_PV = idesc_bf16(128, D, 0, 1), ID_EO = idesc_bf16(128, C::DEP, 0, 0);
      const uint32_t loQ = desc_lo(sbase + SM_Q, 16);
      const uint32_t loWeg = desc_lo(sbase + SM_W + W_EG, C::EGN * 16), loWr = desc_lo(sbase + SM_W + W_R, C::DEP * 16);
      const uint32_t loWb = desc_lo(sbase + SM_W + W_B, C::DEP * 16), loI = desc_lo(sbase + SM_W + W_I, 256);
      const uint32_t loOnes = desc_lo(sbase + SM_ONES, 2048);
      // low descriptor word of the e operand of key kt of stage st (K-major, 128B swizzle; k-step s = +32 bytes)
      auto lo_e = [&](int st, int kt) {
        const int ch = DE >= 16 ? kt * DE : (kt & ~1) * DE;          // first channel of the K window inside the tile row
        return desc_lo(sbase + SM_STAGE + st * STAGE_BYTES + ST_E + (uint32_t)(ch >> 6) * 16384u + (uint32_t)(ch & 63) * 2u, 16);
      };
      const bool use_lo = a.prep->use_lo != 0;           // W' = hi + lo only when the logits are large (wide.h)
      auto issue_mma1 = [&](int q, int st, int kt) {     // S and [E|G] of the group's next key
        const uint32_t tg = tmem + TM_G + q * GC;
        const uint32_t loK = desc_lo(sbase + SM_KVX + q * 3 * KV_MAT, 16);
#pragma unroll
        for (int at = 0; at < C::NQA; ++at) {              // k-steps over the channels: runs of <= 4 per 64-channel atom of Kexp
          const int ks = C::DKS - 4 * at < 4 ? C::DKS - 4 * at : 4;
          if (ks == 4) MmaChain<4>::ts(tg + G_S, tmem + TM_Q + at * 32, loK + at * 128, HI_SW, ID_S, at > 0, 8, 2);
          else if (ks == 3) MmaChain<3>::ts(tg + G_S, tmem + TM_Q + at * 32, loK + at * 128, HI_SW, ID_S, at > 0, 8, 2);
          else if (ks == 2) MmaChain<2>::ts(tg + G_S, tmem + TM_Q + at * 32, loK + at * 128, HI_SW, ID_S, at > 0, 8, 2);
          else if (ks == 1) MmaChain<1>::ts(tg + G_S, tmem + TM_Q + at * 32, loK + at * 128, HI_SW, ID_S, at > 0, 8, 2);
        }
        const uint32_t le = lo_e(st, kt);
        const uint32_t lw = loWeg + (DE >= 16 ? 0u : (uint32_t)(kt & 1) * (W_EG_SZ / 16));
        constexpr int EK = C::DEW / 16;                    // W' = hi + lo (wide.h): the e window is multiplied by both
        MmaChain<EK>::ss(tg + G_EG, le, HI_SW, lw, HI_NONE, ID_EG, 0, 2, 2 * C::EGN);
        if (use_lo) MmaChain<EK>::ss(tg + G_EG, le, HI_SW, lw + 2 * EK * C::EGN, HI_NONE, ID_EG, 1, 2, 2 * C::EGN);
      };
      auto issue_mma2 = [&](int q, int st, int kt, int vslot, bool first) {   // O += A~ Vexp ; e' = e I + H_hat W_r + b_r
        const uint32_t tg = tmem + TM_G + q * GC;
        const uint32_t loV = desc_lo(sbase + SM_KVX + (q * 3 + 1 + vslot) * KV_MAT, 2048);
        (void)first;
        MmaChain<1>::ts(tmem + TM_O, tg + G_A, loV, HI_SW, ID_PV, 1u, 0, 0);
        const uint32_t le = lo_e(st, kt);
        const uint32_t li = loI + (DE >= 16 ? 0u : (uint32_t)(kt & 1) * 32u);
#pragma unroll
        for (int s = 0; s < C::DEP / 16; ++s)
          MmaChain<1>::ss(tg + G_EO + 16 * s, le + (DE >= 16 ? 2 * s : 0), HI_SW, li, HI_NONE, ID_S, 0, 0, 0);
        MmaChain<1>::ts(tg + G_EO, tg + G_H, loWr, HI_NONE, ID_EO, 1, 0, 0);
        MmaChain<1>::ss(tg + G_EO, loOnes, HI_NONE, loWb, HI_NONE, ID_EO, 1, 0, 0);
      };
      tc_fence_after();
      mbar_wait(smem_u32(&bars->q_full), 0);
      mbar_wait(bar_e0, 0);
      tc_fence_after();
      for (int q = q0; q < NG; q += NI) { issue_mma1(q, 0, q); mma_commit_w(bar_ready0 + 8 * q); }
      int T = 0, i = 0, st = 0;                          // tile / index inside the tile / stage of key j
      for (int j = 0; j < J; ++j) {
        int T2 = T, i2 = i + 1, st2 = st;                // the same for key j + 1
        if (i2 == KPG) { i2 = 0; ++T2; if (++st2 == NS) st2 = 0; }
        for (int q = q0; q < NG; q += NI) {
          mbar_wait(bar_done0 + 8 * q, j & 1);           // group q finished key j (its A~ / H_hat are in tensor memory)
          tc_fence_after();
          fence_proxy_async_smem();
          if (!MAXONLY) issue_mma2(q, st, i * NG + q, j & 1, j == 0 && q == 0);
          if (j + 1 < J) {
            if (i2 == 0 && q == q0) { mbar_wait(bar_e0 + 8 * st2, (T2 / NS) & 1); tc_fence_after(); }
            issue_mma1(q, st2, i2 * NG + q);
          }
          mma_commit_w(bar_ready0 + 8 * q);
        }
        T = T2; i = i2; st = st2;
      }
    } else if (warp == 4 * NG + 1 && lane == 0) {
      // =============================== TMA producer ===============================
      for (int T = 0; T < NT; ++T) {                     // store tile T when every group has written its e', refill the stage
        const int st = T % NS;
        mbar_wait(bar_td0 + 8 * st, (T / NS) & 1);
        if (!MAXONLY) {
          fence_proxy_async_smem();
#pragma unroll
          for (int x = 0; x < C::NBOX; ++x)
            tma_store_3d(&tm_eo, sbase + SM_STAGE + st * STAGE_BYTES + ST_E + x * 16384, T * TK * DE + 64 * x, l0, b);
          tma_store_commit();
        }
        if (T + NS < NT) { if (!MAXONLY) tma_store_wait_read<0>(); load_tile(T + NS); }
      }
      if (!MAXONLY) tma_store_wait_all<0>();
    }
    __syncwarp();
    __syncthreads();                                     // sync C: everything is stored
    if (warp == 4 * NG) { tc_fence_after(); tmem_dealloc(tmem, 512); }
    return;
  }

  // ================================= compute threads =================================
  if (C::USE_SETMAXNREG) reg_alloc<C::REG_COMPUTE>();
  const int q = warp >> 2, t = tid & 127;
  const int l = l0 + t;
  const bool rowvalid = l < N;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t tg = tlane + TM_G + q * GC;
  const float *cst = (const float *)(smem + SM_CONST);
  const uint8_t *smask = smem + SM_MASK;
  const float lo = a.clip_lo, hi = a.clip_hi, ln_eps = a.ln_eps;
  // exponent reference of the softmax: 0 while the logits cannot overflow, else the row maximum of the pre-pass
  const bool use_ref = !MAXONLY && a.prep->bound > kWideSoftmaxBudget;
  float *sref = (float *)(smem + SM_REF) + t * H;
  const uint32_t trow = (uint32_t)t * 128u, tx7 = (uint32_t)(t & 7);
  const uint32_t bar_ready = bar_ready0 + 8 * q, bar_done = bar_done0 + 8 * q;
  float psum[H], gsum[H];                                // MAXONLY: psum holds the running maxima
#pragma unroll
  for (int i = 0; i < H; ++i) { psum[i] = MAXONLY ? -INFINITY : 0.f; gsum[i] = 0.f; }

  // expanded K / V operands of key kt of stage st: channel c = t of the key's row goes to row c % h of the 16 x d
  // operand, into the 16-byte chunk c / 8 at its own position c % 8; every other element of the operand stays zero
  const uint32_t x_dst = (uint32_t)(t >> 6) * 2048u + (uint32_t)(t % H) * 128u + ((((uint32_t)(t & 63) >> 3) ^ (uint32_t)(t % H)) & 7u) * 16u;
  auto build = [&](int st, int kt, int vslot) {
    if (t < D) {
      const uint8_t *src = smem + SM_STAGE + st * STAGE_BYTES + ST_K + (uint32_t)(kt * D + t) * 2u;
      uint8_t *dst = smem + SM_KVX + q * 3 * KV_MAT + x_dst;
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {
        const uint32_t val = *(const uint16_t *)(src + kv * (ST_V - ST_K));
        const uint32_t wv = val << ((t & 1) * 16);
        uint4 ch;
        ch.x = ((t >> 1) & 3) == 0 ? wv : 0u; ch.y = ((t >> 1) & 3) == 1 ? wv : 0u;
        ch.z = ((t >> 1) & 3) == 2 ? wv : 0u; ch.w = ((t >> 1) & 3) == 3 ? wv : 0u;
        *(uint4 *)(dst + (kv ? (1 + vslot) * KV_MAT : 0u)) = ch;
      }
    }
  };

  // ---- phase A: key m (kt inside the tile of stage st): every head of this (row, key) pair ----
  auto phase_a = [&](int m, int st, int kt) {
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES + ST_E + (uint32_t)((kt * DE) >> 6) * 16384u + trow;
    const uint32_t cb = (uint32_t)((kt * DE) & 63) >> 3;               // first 16-byte chunk of the key inside the row
    // LayerNorm statistics of e[l, m, :]  (two passes over the shared-memory row, nothing kept in registers)
    float mu = 0.f;
#pragma unroll
    for (int j = 0; j < DE / 8; ++j) {
      const uint4 ev = *(const uint4 *)(es + (((cb + j) ^ tx7) << 4));
      mu += ((bf16_lo(ev.x) + bf16_hi(ev.x)) + (bf16_lo(ev.y) + bf16_hi(ev.y))) + ((bf16_lo(ev.z) + bf16_hi(ev.z)) + (bf16_lo(ev.w) + bf16_hi(ev.w)));
    }
    mu *= 1.0f / DE;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < DE / 8; ++j) {
      const uint4 ev = *(const uint4 *)(es + (((cb + j) ^ tx7) << 4));
      const float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y), bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float dlt = x[c] - mu; var = fmaf(dlt, dlt, var); }
    }
    const float r = rsqrtf(fmaf(var, 1.0f / DE, ln_eps));
    const float nrm = -r * mu;
    const bool kvalid = smask[m] != 0;
#pragma unroll
    for (int half = 0; half < H / 8; ++half) {
      uint32_t sreg[8], egreg[16];
      tmem_ld8(tg + G_S + 8 * half, sreg);
      tmem_ld16(tg + G_EG + 16 * half, egreg);
      uint32_t rb[4] = {0u, 0u, 0u, 0u};
      if (RAND) {   // one Philox call = 2 keys x 4 heads (rng_elem_index, common.cuh); this thread uses its key's half
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          const uint64_t qd = rng_elem_index((uint64_t)b, (uint64_t)l, (uint64_t)(m & ~1), (uint32_t)(8 * half + 4 * q4), (uint64_t)N, H) >> 3;
          const uint64_t roff = a.offset + (a.offset_dev ? *a.offset_dev : 0ull);
          const Philox4 ph = philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), (uint32_t)roff, (uint32_t)(roff >> 32),
                                           (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          rb[2 * q4] = (m & 1) ? ph.z : ph.x; rb[2 * q4 + 1] = (m & 1) ? ph.w : ph.y;
        }
      }
      float uE[8], vE[8], uG[8], vG[8];
#pragma unroll
      for (int v4 = 0; v4 < 2; ++v4) {
        const float4 a0 = *(const float4 *)(cst + 8 * half + 4 * v4), a1 = *(const float4 *)(cst + 16 + 8 * half + 4 * v4);
        const float4 a2 = *(const float4 *)(cst + 32 + 8 * half + 4 * v4), a3 = *(const float4 *)(cst + 48 + 8 * half + 4 * v4);
        uE[4 * v4] = a0.x; uE[4 * v4 + 1] = a0.y; uE[4 * v4 + 2] = a0.z; uE[4 * v4 + 3] = a0.w;
        vE[4 * v4] = a1.x; vE[4 * v4 + 1] = a1.y; vE[4 * v4 + 2] = a1.z; vE[4 * v4 + 3] = a1.w;
        uG[4 * v4] = a2.x; uG[4 * v4 + 1] = a2.y; uG[4 * v4 + 2] = a2.z; uG[4 * v4 + 3] = a2.w;
        vG[4 * v4] = a3.x; vG[4 * v4 + 1] = a3.y; vG[4 * v4 + 2] = a3.z; vG[4 * v4 + 3] = a3.w;
      }
      float nsh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};             // -ref * log2(e) of this row's heads
      if (use_ref) {
        const float4 r0 = *(const float4 *)(sref + 8 * half), r1 = *(const float4 *)(sref + 8 * half + 4);
        nsh[0] = r0.x; nsh[1] = r0.y; nsh[2] = r0.z; nsh[3] = r0.w; nsh[4] = r1.x; nsh[5] = r1.y; nsh[6] = r1.z; nsh[7] = r1.w;
      }
      tmem_ld_wait();
      float av[8], hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float S = __uint_as_float(sreg[i]);
        const float E = fmaf(r, __uint_as_float(egreg[i]), fmaf(nrm, uE[i], vE[i]));
        const float G = fmaf(r, __uint_as_float(egreg[8 + i]), fmaf(nrm, uG[i], vG[i]));
        const float Hh = fminf(fmaxf(S, lo), hi) + E;                      // egt_layers.py:79-86
        bool live = kvalid;
        if (RAND) {
          const uint32_t w = rb[i >> 1];
          const uint32_t bits = (i & 1) ? (w >> 16) : (w & 0xFFFFu);
          live = live && !(bits < a.rand_thr);                             // :103-108
        }
        if (MAXONLY) {                                                     // row maximum over the live keys only
          psum[8 * half + i] = fmaxf(psum[8 * half + i], live ? Hh : -INFINITY);
          av[i] = hv[i] = G;                                               // (unused)
          continue;
        }
        const float pr = live ? ex2_approx(fmaf(Hh, kLog2e, nsh[i])) : 0.f;    // :111 (un-normalised, relative to ref)
        const float gg = live ? sigmoid_fast(G) : 0.f;                     // :112
        psum[8 * half + i] += pr;
        gsum[8 * half + i] += gg;
        av[i] = pr * gg;                                                   // :113
        hv[i] = Hh;
      }
      if (MAXONLY) continue;
      uint32_t apack[4], hpack[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { apack[i] = pack_bf16(av[2 * i], av[2 * i + 1]); hpack[i] = pack_bf16(hv[2 * i], hv[2 * i + 1]); }
      // operands over the consumed inputs: A~ half -> S columns 4*half.. (S half 0 is consumed in both cases), H_hat half
      // -> EG columns 4*half.. (EG half 0 likewise); with h = 8 the upper half of the K = 16 operands is written as zeros
      if (H == 8) {
        const uint32_t a8[8] = {apack[0], apack[1], apack[2], apack[3], 0u, 0u, 0u, 0u};
        const uint32_t h8[8] = {hpack[0], hpack[1], hpack[2], hpack[3], 0u, 0u, 0u, 0u};
        tmem_st8(tg + G_A, a8);
        tmem_st8(tg + G_H, h8);
      } else {
        tmem_st4(tg + G_A + 4 * half, apack);
        tmem_st4(tg + G_H + 4 * half, hpack);
      }
    }
  };

  // ---- phase B: e' of key kt of stage st (complete in tensor memory) -> bf16, in place over the e stage ----
  auto phase_b = [&](int st, int kt) {
    uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES + ST_E + (uint32_t)((kt * DE) >> 6) * 16384u + trow;
    const uint32_t cb = (uint32_t)((kt * DE) & 63) >> 3;
#pragma unroll
    for (int j0 = 0; j0 < DE / 8; j0 += 4) {
      constexpr int NCH = DE / 8 < 4 ? DE / 8 : 4;
      uint32_t dr[8 * NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) tmem_ld8(tg + G_EO + 8 * (j0 + j), dr + 8 * j);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        uint4 ov;
        ov.x = pack_bf16(__uint_as_float(dr[8 * j + 0]), __uint_as_float(dr[8 * j + 1]));
        ov.y = pack_bf16(__uint_as_float(dr[8 * j + 2]), __uint_as_float(dr[8 * j + 3]));
        ov.z = pack_bf16(__uint_as_float(dr[8 * j + 4]), __uint_as_float(dr[8 * j + 5]));
        ov.w = pack_bf16(__uint_as_float(dr[8 * j + 6]), __uint_as_float(dr[8 * j + 7]));
        *(uint4 *)(es + (((cb + j0 + j) ^ tx7) << 4)) = ov;
      }
    }
  };

  // ---- pipeline ----
  mbar_wait(bar_e0, 0);                                  // tile 0 (e, K rows, V rows) has landed
  build(0, q, 0);
  {   // Q tile: shared memory (TMA, 128B swizzle) -> tensor memory, packed bf16; this thread copies its share of row t
    mbar_wait(smem_u32(&bars->q_full), 0);
    constexpr int QCH = D / 8 / NG;                      // 16-byte chunks (8 channels) per thread
    static_assert((D / 8) % NG == 0, "Q copy split");
#pragma unroll
    for (int c = 0; c < QCH; ++c) {
      const uint32_t ch = (uint32_t)(q * QCH + c);
      const uint4 v = *(const uint4 *)(smem + SM_Q + (ch >> 3) * 16384u + trow + (((ch & 7u) ^ tx7) << 4));
      const uint32_t r4[4] = {v.x, v.y, v.z, v.w};
      tmem_st4(tlane + TM_Q + 4 * ch, r4);
    }
    tmem_st_wait();
  }
  if (!MAXONLY) {   // the shared accumulator O starts at zero (two issuers: there is no "first" product)
    constexpr int CWZ = D / NG;
    const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < CWZ / 8; ++j) tmem_st8(tlane + TM_O + q * CWZ + 8 * j, z);
    tmem_st_wait();
  }
  if (use_ref && q == 0) {   // row maxima of the pre-pass -> shared memory as -ref * log2(e) (every group reads them)
    const size_t ps = ((size_t)b * N + (rowvalid ? l : 0)) * H;
#pragma unroll
    for (int i = 0; i < H; ++i) sref[i] = rowvalid ? -a.lse[ps + i] * kLog2e : 0.f;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();                                       // sync B
  {
    int T = 0, i = 0, st = 0;                            // tile / index inside the tile / stage of key j
    int pst = 0, pkt = 0, plast = 0;                     // the previous key of this group
    for (int j = 0; j < J; ++j) {
      const int kt = i * NG + q, m = T * TK + kt;
      mbar_wait(bar_ready, j & 1);                       // S / EG of key j are in tensor memory; so is e' of key j-1
      tc_fence_after();
      if (j > 0 && !MAXONLY) {
        phase_b(pst, pkt);
        if (plast) { fence_proxy_async_smem(); mbar_arrive(bar_td0 + 8 * pst); }
      }
      phase_a(m, st, kt);
      if (MAXONLY && i == KPG - 1) mbar_arrive(bar_td0 + 8 * st);   // nothing is written back: the stage is free
      int T2 = T, i2 = i + 1, st2 = st;
      if (i2 == KPG) { i2 = 0; ++T2; if (++st2 == NS) st2 = 0; }
      if (j + 1 < J) {
        if (i2 == 0) mbar_wait(bar_e0 + 8 * st2, (T2 / NS) & 1);
        build(st2, i2 * NG + q, (j + 1) & 1);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_done);
      pst = st; pkt = kt; plast = (i == KPG - 1);
      T = T2; i = i2; st = st2;
    }
    mbar_wait(bar_ready, J & 1);                         // e' of the last key
    tc_fence_after();
    if (!MAXONLY) {
      phase_b(pst, pkt);
      fence_proxy_async_smem();
      mbar_arrive(bar_td0 + 8 * pst);
    }
  }
  tc_fence_before();
  asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");   // every tcgen05.mma of the CTA has completed

  // ---- row epilogue: normalise, centrality scaler, V_att, saved statistics ----
  tc_fence_after();
  {
    float *xch = (float *)smem;                          // Q / Kexp / Vexp are idle now
    float *mine = xch + ((size_t)q * 128 + t) * (2 * H);
#pragma unroll
    for (int i = 0; i < H; i += 4) {
      *(float4 *)(mine + i) = make_float4(psum[i], psum[i + 1], psum[i + 2], psum[i + 3]);
      *(float4 *)(mine + H + i) = make_float4(gsum[i], gsum[i + 1], gsum[i + 2], gsum[i + 3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
    if (MAXONLY) {   // row maxima over the groups' keys -> lse[0]; a row without a live key keeps the reference 0
      if (q == 0) {
#pragma unroll
        for (int g2 = 1; g2 < NG; ++g2) {
          const float *oth = xch + ((size_t)g2 * 128 + t) * (2 * H);
#pragma unroll
          for (int i = 0; i < H; ++i) psum[i] = fmaxf(psum[i], oth[i]);
        }
        if (rowvalid) {
          const size_t ps = ((size_t)b * N + l) * H;
#pragma unroll
          for (int i = 0; i < H; ++i) a.lse[ps + i] = psum[i] > -INFINITY ? psum[i] : 0.f;
        }
      }
      tc_fence_before();
      __syncthreads();                                   // sync C
      return;
    }
#pragma unroll
    for (int g2 = 1; g2 < NG; ++g2) {
      const float *oth = xch + ((size_t)((q + g2) % NG) * 128 + t) * (2 * H);
#pragma unroll
      for (int i = 0; i < H; i += 4) {
        const float4 p4 = *(const float4 *)(oth + i), g4 = *(const float4 *)(oth + H + i);
        psum[i] += p4.x; psum[i + 1] += p4.y; psum[i + 2] += p4.z; psum[i + 3] += p4.w;
        gsum[i] += g4.x; gsum[i + 1] += g4.y; gsum[i + 2] += g4.z; gsum[i + 3] += g4.w;
      }
    }
    float f[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float inv = psum[i] > 0.f ? __fdividef(1.f, psum[i]) : 0.f;
      float s = 1.f;
      if (a.scale_degree && l >= a.num_virtual_nodes)                      // egt_layers.py:123-135
        s = a.scaler_type == EGT_SCALER_LOG ? log1pf(gsum[i]) : gsum[i];
      f[i] = inv * s;
    }
    constexpr int CW = D / NG;                           // this thread stores channels [q*CW, (q+1)*CW) of row l
    static_assert(CW % 8 == 0 && CW % H == 0, "column split of the output");
    uint32_t o[CW];
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) tmem_ld8(tlane + TM_O + q * CW + 8 * j, o + 8 * j);
    tmem_ld_wait();
    if (rowvalid) {
      uint4 *dst = (uint4 *)(a.v_att + ((size_t)b * N + l) * D + q * CW);
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __uint_as_float(o[8 * j + c]) * f[(8 * j + c) % H];   // channel = dd*h + hh
        uint4 ov;
        ov.x = pack_bf16(v[0], v[1]); ov.y = pack_bf16(v[2], v[3]); ov.z = pack_bf16(v[4], v[5]); ov.w = pack_bf16(v[6], v[7]);
        dst[j] = ov;
      }
      if (q == 0) {
        const size_t ps = ((size_t)b * N + l) * H, rs = (size_t)a.B * N * H;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          if (!use_ref) a.lse[ps + i] = 0.f;                                 // reference point of the exponent (else: the pre-pass' row maximum)
          a.lse[rs + ps + i] = psum[i] > 0.f ? __logf(psum[i]) : 0.f;
          a.deg[ps + i] = gsum[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();                                       // sync C
}

template <class C>
int launch_cfg(const WideFwdArgs &a, const void *e, void *e_out, const void *qkv, cudaStream_t st) {
  CUtensorMap tm_e, tm_eo, tm_q, tm_kv;
  const uint64_t N = a.N, B = a.B, DE = C::DE, D = C::D;
  int rc;
  if ((rc = encode_tmap_3d(&tm_e, e, N * DE, N, B, N * DE * 2, N * N * DE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_eo, e_out, N * DE, N, B, N * DE * 2, N * N * DE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_q, qkv, 3 * D, N, B, 3 * D * 2, N * 3 * D * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_kv, qkv, 3 * D, N, B, 3 * D * 2, N * 3 * D * 2, (uint32_t)D, C::TK, 1, 0))) return rc;
  const int smem = C::FWD_SMEM + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_fwd_kernel<C, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_fwd_kernel<C, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_fwd_kernel<C, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_fwd_kernel<C, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.B);
  {   // row-maximum pre-pass: every CTA returns at once unless the logit bound exceeds the exponent budget
    LaunchScope _ls("wide_fwd_rowmax_kernel", st);
    if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(wide_fwd_kernel<C, true, true>, grid, dim3(C::THREADS), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
    else EGT_CHECK_CUDA(launch_pdl(wide_fwd_kernel<C, false, true>, grid, dim3(C::THREADS), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
  }
  LaunchScope _ls("wide_fwd_kernel", st);
  if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(wide_fwd_kernel<C, true, false>, grid, dim3(C::THREADS), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
  else EGT_CHECK_CUDA(launch_pdl(wide_fwd_kernel<C, false, false>, grid, dim3(C::THREADS), smem, st, tm_e, tm_eo, tm_q, tm_kv, a));
  return EGT_OK;
}

}  // namespace

int wide_fwd_launch(const egt_block_cfg_t *cfg, const WideFwdArgs &a, const void *e, void *e_out, const void *qkv,
                    cudaStream_t st) {
  const int h = cfg->attn.h, dk = cfg->attn.dk, de = cfg->d_e;
  if (h == 16 && dk == 8 && de == 32) return launch_cfg<WideFwdC5>(a, e, e_out, qkv, st);
  if (h == 8 && dk == 8 && de == 64) return launch_cfg<WideFwdC1>(a, e, e_out, qkv, st);
  if (h == 8 && dk == 12 && de == 8) return launch_cfg<WideFwdC3>(a, e, e_out, qkv, st);
  EGT_REQUIRE(false, EGT_E_SHAPE, "wide_fwd: no instantiation for h=%d dk=%d d_e=%d", h, dk, de);
}

}  // namespace egt
