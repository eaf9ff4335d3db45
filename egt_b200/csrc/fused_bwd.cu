// fused_bwd.cu -- fused backward of the EGT attention block's N x N part on sm_100a tensor cores.
//
//   reference: TF autodiff of EGT.call_gated (lib/models/egt_layers.py:57-143) together with the edge
//   projections, LayerNorm on e and the residual edge write-back of edge_update_residual
//   (lib/models/graph_xformer_model_base.py:192-218), as derived in SURVEY.md 3.4.  Nothing of shape
//   [B,N,N,h] touches HBM: e and de' stream in once by TMA, de streams out once by TMA, everything
//   else (E, G, H_hat, P, g, dE, dG, dS) is recomputed / consumed on chip.
//
// One CTA = one graph b and 128 query rows, four warpgroups (setmaxnreg 192 / 192 / 64 / 64):
//   warps 0-3, 4-7   "main": thread t of both groups owns query row l0+t == TMEM lane t; group 0 handles
//                    heads 0-3, group 1 heads 4-7 (phase A below, all the per-(l,m,h) arithmetic)
//   warp 8           issuer of TMA and tcgen05.mma (warps 9-11 only complete its warpgroup)
//   warps 12-15      "helper": per-(l,m) work that needs no head split -- LayerNorm backward + residual (de),
//                    the expanded K / V operands of the coming pairs, and the dK / dV read-out
//
// Keys are processed in PAIRS.  For pair p the tensor core produces, in TMEM (columns ordered (g,key,hh4)):
//     S   [128x16] = Qs   [128x64] * Kexp^T       Kexp[(g,key,hh4), c] = K[key,c] * [c % 8 == hh]
//     dA  [128x16] = dO   [128x64] * Vexp^T       dO = dV_att (.) scaler,  Vexp likewise from V
//     EG  [128x32] = e    [128x16] * Wblk         raw edge channels of the two keys x folded-LN weights
//     dHx [128x16] = de'  [128x16] * Wr^T blk     gradient arriving through the edge write-back
// The row's threads recompute p, g, H_hat from the saved row statistics and form
//     dP = dA g ; dH = p (dP - D) + dHx ; dG = (dA p + ddeg) g (1-g) ; dS = dH [lo <= S <= hi]
// which go back to TMEM as bf16 A-operands of
//     dQ  [128x64] += dS [128x16] * Kexp                      (accumulated over all keys)
//     dx^ [128x16]  = [dE|dG] [128x32] * W'^T blk             (then LayerNorm backward + residual -> de)
// and, transposed through shared memory (MN-major A, K = query row), once per 16 keys of
//     dKexp [(key,hh) x 64] = dS^T * Qs ;  dVexp [(key,hh) x 64] = A~^T * dO
// whose block diagonal (c % 8 == hh) is dK / dV.  Weight-gradient sums are kept in registers
// (packed fp32x2 FMAs) and folded by fused_bwd_finalize_kernel.
//
// There is no CTA-wide barrier in the main loop: compute threads arrive on an mbarrier when their part of a
// key pair is done and the issuer waits for it.  Every dependency spans TWO pairs (same structure as
// fused_fwd.cu): the S / dA / EG / dHx products of pair it+2 are issued when pair it is done, and the de
// update of pair it (which needs the d x^ product issued when pair it is done) runs during pair it+2.
#include <stdlib.h>
#include "common.cuh"
#include "fused.h"
#include "umma.cuh"
#include "fused_prep.cuh"

namespace egt {
using namespace umma;

namespace {

constexpr int NS = 3;                                  // input stages of 8 keys: e | de' | K rows | V rows
constexpr uint32_t SM_Q = 0;                           // [128 x 128B] swizzled, Q pre-scaled by dk^-0.5
constexpr uint32_t SM_DO = 16384;                      // [128 x 128B] swizzled, dV_att (.) scaler
constexpr uint32_t SM_STAGE = 32768;
constexpr uint32_t ST_E = 0, ST_DE = 16384, ST_K = 32768, ST_V = 33792, STAGE_BYTES = 34816;
constexpr uint32_t SM_KVX = SM_STAGE + NS * STAGE_BYTES;       // 4 slots x (Kexp 2048 | Vexp 2048)
constexpr uint32_t SM_TR = SM_KVX + 4 * 4096;                  // dS^T 32768 | A~^T 32768  (16 keys x 128 rows)
constexpr uint32_t SM_W = SM_TR + 65536;                       // b_eg 1024 | b_hx 512 | b_de 2 x 512 | b_eg_lo 1024
constexpr uint32_t SM_CONST = SM_W + 3584;                     // uE vE uG vG (32 floats)
constexpr uint32_t SM_BAR = SM_CONST + 256;
constexpr uint32_t SM_MASK = SM_BAR + 256;                       // key-valid bytes, zero padded (N <= 4096)
constexpr uint32_t SM_TOTAL = SM_MASK + 4096 + 16;
static_assert(SM_TOTAL + 1024 <= 232448, "shared memory budget");

constexpr uint32_t TM_DQ = 0, TM_DK = 64, TM_DV = 128;
constexpr uint32_t TM_IN = 192, TM_IN_COLS = 80;               // 2 buffers (pair parity): S 16 | dA 16 | EG 32 | dHx 16
constexpr uint32_t IN_S = 0, IN_DA = 16, IN_EG = 32, IN_HX = 64;
constexpr uint32_t TM_DX = 352, TM_DX_COLS = 16;               // 2 buffers (pair parity): d x^ 16
constexpr uint32_t TM_OUT = 384, TM_OUT_COLS = 32;             // 2 buffers (pair parity): dS 8 | dZ 16 (bf16 A operands)

constexpr uint32_t ID_N16 = idesc_bf16(128, 16, 0, 0);
constexpr uint32_t ID_N32 = idesc_bf16(128, 32, 0, 0);
constexpr uint32_t ID_DQ = idesc_bf16(128, 64, 0, 1);
constexpr uint32_t ID_T = idesc_bf16(128, 64, 1, 1);

struct Bars { uint64_t q_full, e_full[NS], mma1[2], mma2[2], tbar, step[2]; uint32_t tmem_base; };

__device__ __forceinline__ float sel8(const uint32_t *o, int hh) {
  const uint32_t a0 = (hh & 1) ? o[1] : o[0], a1 = (hh & 1) ? o[3] : o[2];
  const uint32_t a2 = (hh & 1) ? o[5] : o[4], a3 = (hh & 1) ? o[7] : o[6];
  const uint32_t b0 = (hh & 2) ? a1 : a0, b1 = (hh & 2) ? a3 : a2;
  return __uint_as_float((hh & 4) ? b1 : b0);
}

}  // namespace

template <bool RAND>
__global__ void __launch_bounds__(512, 1)
fused_bwd_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_dei,
                 const __grid_constant__ CUtensorMap tm_de, const __grid_constant__ CUtensorMap tm_q,
                 const __grid_constant__ CUtensorMap tm_kv, const FusedBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  const uint32_t sbase = smem_u32(smem);
  Bars *bars = (Bars *)(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, l0 = blockIdx.x * 128;
  const int N = a.N;
  // gridDim.z > 1: the keys of a graph are split over gridDim.z CTAs in runs of a multiple of 16 keys (wave quantisation,
  // see wide_bwd.cu).  Pair / tile / key indices below are LOCAL to the run [K0, K0 + NK); K0 turns them into positions
  // in the graph where global memory is addressed (TMA coordinates, mask, dK / dV rows, random-mask counters).
  const int per_split = (((N + (int)gridDim.z - 1) / (int)gridDim.z) + 15) & ~15;
  const int K0 = (int)blockIdx.z * per_split;
  const int NK = N - K0 < per_split ? N - K0 : per_split;          // the launcher guarantees NK >= 1
  const int NT = (NK + 7) / 8, NP = (NK + 1) / 2;     // 8-key tiles, key pairs

  pdl_trigger();
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars->q_full), 1);
      for (int i = 0; i < NS; ++i) mbar_init(smem_u32(&bars->e_full[i]), 1);
      for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->mma1[i]), 1); mbar_init(smem_u32(&bars->mma2[i]), 1); }
      mbar_init(smem_u32(&bars->tbar), 1);
      // every main and helper thread arrives once per key pair; pairs alternate between two barriers because a
      // thread may finish pair it+1 before a slower one finishes pair it (dependencies span two pairs)
      mbar_init(smem_u32(&bars->step[0]), 384); mbar_init(smem_u32(&bars->step[1]), 384);
      mbar_fence_init();
      tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_dei); tma_prefetch_desc(&tm_de);
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    pdl_wait();
  } else if (warp < 8) {
    pdl_wait();                                        // prep / dV_att come from the preceding kernel
    if (tid < 160) ((uint4 *)(smem + SM_W))[tid] = ((const uint4 *)a.prep->b_eg)[tid];   // b_eg | b_hx | b_de
    else if (tid < 224) ((uint4 *)(smem + SM_W + 2560))[tid - 160] = ((const uint4 *)a.prep->b_eg_lo)[tid - 160];
    if (tid < 32) ((float *)(smem + SM_CONST))[tid] = a.prep->uE[tid];                  // uE vE uG vG
    for (int i = tid; i < 2 * ((NK + 1) / 2); i += 256)                                  // key-valid bytes
      smem[SM_MASK + i] = i < NK ? (a.mask ? (uint8_t)(a.mask[(size_t)blockIdx.y * N + K0 + i] != 0) : (uint8_t)1) : (uint8_t)0;
    fence_proxy_async_smem();
  } else if (warp >= 12) {
    pdl_wait();                                        // the helper adds into d_qkv, zeroed earlier on the stream
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 8 && warp < 12) {
    // ====================== issuer warpgroup (warp 8 issues; warps 9-11 only keep the barriers) ======
    reg_dealloc<64>();   // 256 x 192 + 128 x 64 + 128 x 64 == 512 x 128: setmaxnreg only moves registers inside the CTA's launch allocation
    const bool leader = warp == 8 && lane == 0;
    auto load_tile = [&](int T) {
      const int st = T % NS;
      const uint32_t bar = smem_u32(&bars->e_full[st]);
      const uint32_t dst = sbase + SM_STAGE + st * STAGE_BYTES;
      mbar_expect_tx(bar, STAGE_BYTES);
      tma_load_3d(dst + ST_E, &tm_e, bar, K0 * 8 + T * 64, l0, b);
      tma_load_3d(dst + ST_DE, &tm_dei, bar, K0 * 8 + T * 64, l0, b);
      tma_load_3d(dst + ST_K, &tm_kv, bar, FD, K0 + T * 8, b);
      tma_load_3d(dst + ST_V, &tm_kv, bar, 2 * FD, K0 + T * 8, b);
    };
    // descriptor low words (address | LBO); the high words are compile-time constants
    constexpr uint32_t HI_SW = desc_hi(1024, LAYOUT_SW128), HI_NONE = desc_hi(128, LAYOUT_NONE);
    const uint32_t loQ = desc_lo(sbase + SM_Q, 16), loDO = desc_lo(sbase + SM_DO, 16);
    const uint32_t loK = desc_lo(sbase + SM_KVX, 16), loV = desc_lo(sbase + SM_KVX + 2048, 16);
    const uint32_t loKmn = desc_lo(sbase + SM_KVX, 2048);
    const uint32_t loE = desc_lo(sbase + SM_STAGE + ST_E, 16), loDE = desc_lo(sbase + SM_STAGE + ST_DE, 16);
    const uint32_t loWeg = desc_lo(sbase + SM_W, 512), loWhx = desc_lo(sbase + SM_W + 1024, 256);
    const uint32_t loWde = desc_lo(sbase + SM_W + 1536, 256);
    const uint32_t loWegLo = desc_lo(sbase + SM_W + 2560, 512);
    const bool use_lo = a.prep->use_lo != 0;           // W' = hi + lo only when the logits are large (fused.h)
    const uint32_t bar_e0 = smem_u32(&bars->e_full[0]), bar_m1 = smem_u32(&bars->mma1[0]), bar_m2 = smem_u32(&bars->mma2[0]);
    // tcgen05.mma is issued warp-collectively by warp 8 (umma.cuh): converged warp, one elected lane, k-chains in one
    // asm statement.  TMA stays with lane 0.
    auto issue_mma1 = [&](int p) {                     // -> input buffer p & 1 (no commit: see the handshake below)
      const int T = p >> 2, j = p & 3, st = T % NS, slot = p & 3, buf = p & 1;
      mbar_wait(bar_e0 + 8 * st, (T / NS) & 1);
      tc_fence_after();
      const uint32_t d = tmem + TM_IN + buf * TM_IN_COLS;
      const uint32_t k0 = loK + slot * 256, v0 = loV + slot * 256, e0 = st * (STAGE_BYTES / 16) + 2 * j;
      MmaChain<4>::ss(d + IN_S, loQ, HI_SW, k0, HI_SW, ID_N16, 0, 2, 2);
      MmaChain<4>::ss(d + IN_DA, loDO, HI_SW, v0, HI_SW, ID_N16, 0, 2, 2);
      MmaChain<1>::ss(d + IN_EG, loE + e0, HI_SW, loWeg, HI_NONE, ID_N32, 0, 0, 0);
      if (use_lo) MmaChain<1>::ss(d + IN_EG, loE + e0, HI_SW, loWegLo, HI_NONE, ID_N32, 1, 0, 0);
      MmaChain<1>::ss(d + IN_HX, loDE + e0, HI_SW, loWhx, HI_NONE, ID_N16, 0, 0, 0);
    };
    auto issue_mma2 = [&](int p) {                     // dQ += dS Kexp ; d x^ = [dE|dG] W'^T   (operands / result: parity p & 1)
      const int ob = p & 1, slot = p & 3;
      const uint32_t ao = tmem + TM_OUT + ob * TM_OUT_COLS;
      MmaChain<1>::ts(tmem + TM_DQ, ao, loKmn + slot * 256, HI_SW, ID_DQ, p > 0, 0, 0);
      MmaChain<2>::ts(tmem + TM_DX + ob * TM_DX_COLS, ao + 8, loWde, HI_NONE, ID_N16, 0, 8, 32);
    };
    auto issue_t = [&](int p) {
      if ((p & 7) == 7 || p == NP - 1) {                // a 16-key block of dS^T / A~^T is complete
        const uint32_t loT = desc_lo(sbase + SM_TR, 16384), loQm = desc_lo(sbase + SM_Q, 16384);
        const uint32_t loDOm = desc_lo(sbase + SM_DO, 16384);
        MmaChain<8>::ss(tmem + TM_DK, loT, HI_SW, loQm, HI_SW, ID_T, 0, 128, 128);
        MmaChain<8>::ss(tmem + TM_DV, loT + 2048, HI_SW, loDOm, HI_SW, ID_T, 0, 128, 128);
        mma_commit_w(smem_u32(&bars->tbar));
      }
    };
    if (leader) {
      mbar_expect_tx(smem_u32(&bars->q_full), 16384);
      tma_load_3d(sbase + SM_Q, &tm_q, smem_u32(&bars->q_full), 0, l0, b);
      for (int T = 0; T < NT && T < NS; ++T) load_tile(T);
    }
    __syncthreads();                                   // sync #0: dO tile, Kexp/Vexp of pairs 0,1 are built
    if (warp == 8) {
      tc_fence_after();
      mbar_wait(smem_u32(&bars->q_full), 0);
      issue_mma1(0);
      mma_commit_w(bar_m1);
      if (NP > 1) { issue_mma1(1); mma_commit_w(bar_m1 + 8); }
    }
    const uint32_t bar_step = smem_u32(&bars->step[0]);
    int next_store = 0;                                // tiles [0, next_store) have been handed to the TMA store
    for (int it = 0; it < NP && warp == 8; ++it) {     // warps 9-11 go straight to the tail barrier
      mbar_wait(bar_step + 8 * (it & 1), (it >> 1) & 1);   // main and helper threads finished pair it (no CTA-wide barrier)
      tc_fence_after();
      fence_proxy_async_smem();
      issue_mma2(it);
      if (it + 2 < NP) issue_mma1(it + 2);
      mma_commit_w(bar_m2 + 8 * (it & 1));             // ONE completion per handshake: products of pair it and the
                                                       // inputs of pair it+2, both consumed during pair it+2
      issue_t(it);                                     // (the 16-key transposed products have their own barrier)
      if (lane == 0) {
        // pair it contained the de update of pair it-2; tile T (pairs 4T .. 4T+3) is complete when it == 4T+5
        if (it >= 6 && ((it - 6) & 3) == 0) {          // tile stored at the previous handshake: recycle its stage
          const int T = (it - 6) >> 2;
          tma_store_wait_read<0>();
          if (T + NS < NT) load_tile(T + NS);
        }
        if (it >= 5 && ((it - 5) & 3) == 0) {
          const int T = (it - 5) >> 2;
          tma_store_3d(&tm_de, sbase + SM_STAGE + (T % NS) * STAGE_BYTES + ST_DE, K0 * 8 + T * 64, l0, b);
          tma_store_commit();
          next_store = T + 1;
        }
      }
      __syncwarp();
    }
    __syncthreads();                                   // sync #(NP+1): every de update is done
    if (leader) {
      for (int T = next_store; T < NT; ++T)
        tma_store_3d(&tm_de, sbase + SM_STAGE + (T % NS) * STAGE_BYTES + ST_DE, K0 * 8 + T * 64, l0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
    __syncthreads();                                   // reduction scratch zeroed
    __syncthreads();                                   // partial sums complete
    __syncthreads();                                   // final
    if (warp == 8) tmem_dealloc(tmem, 512);
    return;
  }

  const uint32_t bar_mma2 = smem_u32(&bars->mma2[0]);
  const uint32_t bar_e = smem_u32(&bars->e_full[0]), bar_t = smem_u32(&bars->tbar), bar_step = smem_u32(&bars->step[0]);

  if (warp >= 12) {
    // ================================= helper warpgroup (warps 12-15) ================================
    // Per-(query row, key) work that does not split by head: thread t owns query row l0+t for BOTH keys of a pair.
    reg_dealloc<64>();
    const int t = tid & 127;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t trow = (uint32_t)t * 128u, tx7 = (uint32_t)(t & 7);
    float dbr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) dbr[c] = 0.f;

    // expanded K / V operands of pair p2 (stage st2) into slot p2 & 3: one 16-byte chunk of each per thread
    const int b_n = t >> 3, b_dd = t & 7, b_hh = 4 * (b_n >> 3) + (b_n & 3);
    const uint32_t b_src = (uint32_t)ST_K + (uint32_t)((b_n >> 2) & 1) * 128u + (uint32_t)(b_dd * 8 + b_hh) * 2u;
    const uint32_t b_dst = SM_KVX + (uint32_t)b_n * 128u + ((uint32_t)((b_dd ^ b_n) & 7) << 4);
    auto build = [&](int p2, int st2) {
      const uint8_t *src = smem + SM_STAGE + st2 * STAGE_BYTES + b_src + (p2 & 3) * 256;
#pragma unroll
      for (int kv = 0; kv < 2; ++kv) {
        const uint32_t val = *(const uint16_t *)(src + kv * (ST_V - ST_K));
        const uint32_t wv = val << ((b_hh & 1) * 16);
        uint4 ch;
        ch.x = (b_hh >> 1) == 0 ? wv : 0u; ch.y = (b_hh >> 1) == 1 ? wv : 0u;
        ch.z = (b_hh >> 1) == 2 ? wv : 0u; ch.w = (b_hh >> 1) == 3 ? wv : 0u;
        *(uint4 *)(smem + b_dst + kv * 2048 + (p2 & 3) * 4096) = ch;
      }
    };

    // LayerNorm backward + residual for both keys of pair p -> de, in place over de'
    auto phase_b = [&](int p, int st) {
      uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
      uint32_t dr[16];
      tmem_ld16(tlane + TM_DX + (p & 1) * TM_DX_COLS, dr);
      tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t eoff = trow + (((uint32_t)(2 * (p & 3) + kk) ^ tx7) << 4);
        const uint4 ev = *(const uint4 *)(es + ST_E + eoff);
        float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y),
                      bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
        float mu = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
        mu *= 0.125f;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { x[c] -= mu; var = fmaf(x[c], x[c], var); }
        const float r = rsqrtf(fmaf(var, 0.125f, a.ln_eps));
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          x[c] *= r;                                   // x^
          const float dxh = __uint_as_float(dr[8 * kk + c]);
          m1 += dxh;
          m2 = fmaf(dxh, x[c], m2);
        }
        m1 *= 0.125f; m2 *= 0.125f;
        const uint4 dev = *(const uint4 *)(es + ST_DE + eoff);
        const float dp[8] = {bf16_lo(dev.x), bf16_hi(dev.x), bf16_lo(dev.y), bf16_hi(dev.y),
                             bf16_lo(dev.z), bf16_hi(dev.z), bf16_lo(dev.w), bf16_hi(dev.w)};
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          o[c] = fmaf(r, __uint_as_float(dr[8 * kk + c]) - fmaf(x[c], m2, m1), dp[c]);
          dbr[c] += dp[c];
        }
        uint4 ov;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]);
        ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        *(uint4 *)(es + ST_DE + eoff) = ov;
      }
    };

    // dK / dV of the 16-key block kb: lane t = (key, hh) picks the block diagonal
    const bool single_tile = gridDim.x == 1;
    const bool bf_out = single_tile && gridDim.z == 1 && a.d_qkv_bf != nullptr;
    auto t_epilogue = [&](int kb) {
      const int m = K0 + 16 * kb + (t >> 3), hh = t & 7;
      float *dst = a.d_qkv + ((size_t)b * N + (m < N ? m : 0)) * (3 * FD) + FD + hh;
      __nv_bfloat16 *dst_bf = a.d_qkv_bf + ((size_t)b * N + (m < N ? m : 0)) * (3 * FD) + FD + hh;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld32(tlane + (which ? TM_DV : TM_DK) + 32 * half, o);   // warp-collective: never inside a divergent branch
          tmem_ld_wait();
          if (m < N) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v = sel8(o + 8 * q, hh);
              float *pd = dst + which * FD + (half * 4 + q) * 8;
              if (bf_out) dst_bf[which * FD + (half * 4 + q) * 8] = __float2bfloat16_rn(v);
              else if (single_tile) *pd = v;
              else atomicAdd(pd, v);
            }
          }
        }
      }
    };

    mbar_wait(bar_e, 0);
    build(0, 0);
    if (NP > 1) build(1, 0);
    fence_proxy_async_smem();
    __syncthreads();                                   // sync #0
    int st_b = 0;                                      // stage of pair it - 2
    int st_n = 0, par_n = 0;                           // stage / load parity of pair it + 2
    for (int it = 0; it < NP; ++it) {
      if (((it + 2) & 3) == 0 || it == 0) {            // pair it+2 sits in tile (it+2)>>2
        const int T2 = (it + 2) >> 2;
        st_n = T2 % NS; par_n = (T2 / NS) & 1;
      }
      if (it >= 2) {
        // handshake it-2: d x^ of pair it-2 is in tensor memory, and Kexp / Vexp slot (it+2) & 3 has been read
        mbar_wait(bar_mma2 + 8u * (uint32_t)(it & 1), ((it - 2) >> 1) & 1);
        tc_fence_after();
        phase_b(it - 2, st_b);
      }
      if (it >= 1 && ((it - 1) & 7) == 7) {            // dS^T / A~^T block (it-1)/8 went through the tensor core
        const int kb = (it - 1) >> 3;
        mbar_wait(bar_t, kb & 1);
        tc_fence_after();
        t_epilogue(kb);
      }
      if (it + 2 < NP) {
        if (((it + 2) & 3) == 0) mbar_wait(bar_e + 8 * st_n, par_n);   // first pair of a tile
        build(it + 2, st_n);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_step + 8u * (uint32_t)(it & 1));  // this thread's share of pair it is done
      if (it >= 2 && ((it - 1) & 3) == 0) { if (++st_b == NS) st_b = 0; }   // pair it-1 opens a new tile
    }
    for (int q = (NP >= 2 ? NP - 2 : 0); q < NP; ++q) { // de updates of the last two pairs
      mbar_wait(bar_mma2 + 8 * (q & 1), (q >> 1) & 1);
      tc_fence_after();
      phase_b(q, (q >> 2) % NS);
    }
    fence_proxy_async_smem();
    __syncthreads();                                   // sync #(NP+1)
    {
      const int kb = (NP - 1) >> 3;                    // last (possibly partial) 16-key block
      mbar_wait(bar_t, kb & 1);
      tc_fence_after();
      t_epilogue(kb);
    }
    __syncthreads();                                   // reduction scratch zeroed
    {
      float *red = (float *)(smem + SM_TR);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float v = warp_sum(dbr[c]);
        if (lane == 0) atomicAdd(red + 208 + c, v);
      }
    }
    __syncthreads();                                   // partial sums complete
    tc_fence_before();
    __syncthreads();                                   // final
    return;
  }

  // ================================= main compute threads (warps 0-7) ==============================
  reg_alloc<192>();
  const int g = tid >> 7, t = tid & 127;
  const int l = l0 + t;
  const bool rowvalid = l < N;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const float *cst = (const float *)(smem + SM_CONST);
  const uint8_t *smask = smem + SM_MASK;
  const float lo = a.clip_lo, hi = a.clip_hi;
  const uint32_t trow = (uint32_t)t * 128u, tx7 = (uint32_t)(t & 7);
  const uint32_t trbase = (uint32_t)(t >> 3) * 1024u + tx7 * 128u + (uint32_t)g * 8u;
  const uint32_t bar_mma1 = smem_u32(&bars->mma1[0]);

  // ---- per-row quantities: D = sum_dd dV_att * V_att, scaler s, ddeg, log2 row sum; dO tile ---------
  float Dr[4], ddeg[4], l2[4];
  {
    float s8[8], deg8[8];
    const size_t ps = ((size_t)b * N + (rowvalid ? l : 0)) * FH, rs = (size_t)a.B * N * FH;
#pragma unroll
    for (int hh = 0; hh < 8; ++hh) {
      deg8[hh] = a.deg[ps + hh];
      float s = 1.f;
      if (a.scale_degree && l >= a.num_virtual_nodes)                        // egt_layers.py:123-135
        s = a.scaler_type == EGT_SCALER_LOG ? log1pf(deg8[hh]) : deg8[hh];
      s8[hh] = s;
    }
    const uint4 *dvp = (const uint4 *)(a.d_v_att + ((size_t)b * N + (rowvalid ? l : 0)) * FD);
    const uint4 *vp = (const uint4 *)(a.v_att + ((size_t)b * N + (rowvalid ? l : 0)) * FD);
    float Dacc[8];
#pragma unroll
    for (int hh = 0; hh < 8; ++hh) Dacc[hh] = 0.f;
#pragma unroll
    for (int dd = 0; dd < 8; ++dd) {
      uint4 dv = dvp[dd], vv = vp[dd];
      if (!rowvalid) { dv = make_uint4(0, 0, 0, 0); vv = dv; }
      const float d8[8] = {bf16_lo(dv.x), bf16_hi(dv.x), bf16_lo(dv.y), bf16_hi(dv.y),
                           bf16_lo(dv.z), bf16_hi(dv.z), bf16_lo(dv.w), bf16_hi(dv.w)};
      const float v8[8] = {bf16_lo(vv.x), bf16_hi(vv.x), bf16_lo(vv.y), bf16_hi(vv.y),
                           bf16_lo(vv.z), bf16_hi(vv.z), bf16_lo(vv.w), bf16_hi(vv.w)};
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) Dacc[hh] = fmaf(d8[hh], v8[hh], Dacc[hh]);
      if ((dd >> 2) == g) {                                                  // this thread stages chunks 4g..4g+3
        uint4 o;
        o.x = pack_bf16(d8[0] * s8[0], d8[1] * s8[1]); o.y = pack_bf16(d8[2] * s8[2], d8[3] * s8[3]);
        o.z = pack_bf16(d8[4] * s8[4], d8[5] * s8[5]); o.w = pack_bf16(d8[6] * s8[6], d8[7] * s8[7]);
        *(uint4 *)(smem + SM_DO + trow + (((uint32_t)dd ^ tx7) << 4)) = o;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int hh = 4 * g + i;
      // D = s * ds with ds = sum_dd dV_att * O (O = V_att / s, the un-scaled attention output)
      float D = 0.f, dg = 0.f, lsum = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) if (q == hh) { D = Dacc[q]; }
      float s = 1.f, dgv = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) if (q == hh) { s = s8[q]; dgv = deg8[q]; }
      if (a.scale_degree && l >= a.num_virtual_nodes) {
        const float ds = s != 0.f ? D / s : 0.f;
        dg = a.scaler_type == EGT_SCALER_LOG ? ds / (1.f + dgv) : ds;
      }
      if (rowvalid) lsum = a.lse[ps + hh] + a.lse[rs + ps + hh];   // reference point + log row sum
      Dr[i] = rowvalid ? D : 0.f;
      ddeg[i] = rowvalid ? dg : 0.f;
      l2[i] = lsum * kLog2e;
    }
  }

  // ---- weight-gradient partial sums (registers) ---------------------------------------------------
  float2 ME[8][2], MG[8][2], WR[4][4], sE[2], sG[2];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    ME[c][0] = ME[c][1] = MG[c][0] = MG[c][1] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) WR[i][q] = make_float2(0.f, 0.f);
  sE[0] = sE[1] = sG[0] = sG[1] = make_float2(0.f, 0.f);

  auto ln_stats = [&](const uint4 ev, float *x, float &r, float &nrm) {
    x[0] = bf16_lo(ev.x); x[1] = bf16_hi(ev.x); x[2] = bf16_lo(ev.y); x[3] = bf16_hi(ev.y);
    x[4] = bf16_lo(ev.z); x[5] = bf16_hi(ev.z); x[6] = bf16_lo(ev.w); x[7] = bf16_hi(ev.w);
    float mu = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    mu *= 0.125f;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) { const float dlt = x[c] - mu; var = fmaf(dlt, dlt, var); }
    r = rsqrtf(fmaf(var, 0.125f, a.ln_eps));
    nrm = -r * mu;
  };

  // ---- phase A: pair p, this thread's 4 heads of both keys ------------------------------------------
  auto phase_a = [&](int p, int st) {
    const int j = p & 3, ob = p & 1, buf = p & 1;
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
    const uint32_t tin = tlane + TM_IN + buf * TM_IN_COLS;
    const uint32_t tout = tlane + TM_OUT + ob * TM_OUT_COLS;
    uint32_t dsp[4], dzp[8], atp[4];
    const float4 uE4 = *(const float4 *)(cst + 4 * g), vE4 = *(const float4 *)(cst + 8 + 4 * g);
    const float4 uG4 = *(const float4 *)(cst + 16 + 4 * g), vG4 = *(const float4 *)(cst + 24 + 4 * g);
    const float uE[4] = {uE4.x, uE4.y, uE4.z, uE4.w}, vE[4] = {vE4.x, vE4.y, vE4.z, vE4.w};
    const float uG[4] = {uG4.x, uG4.y, uG4.z, uG4.w}, vG[4] = {vG4.x, vG4.y, vG4.z, vG4.w};
    // the tensor-core products of both keys are requested at once: one tcgen05.wait::ld per pair
    uint32_t sreg2[2][4], dareg2[2][4], egreg2[2][8], hxreg2[2][4];
    uint32_t rbw[4] = {0u, 0u, 0u, 0u};
    if (RAND) {   // one Philox call = this thread's 2 keys x 4 heads (rng_elem_index, common.cuh)
      const uint64_t qd = rng_elem_index((uint64_t)b, (uint64_t)l, (uint64_t)(K0 + 2 * p), 4u * (uint32_t)g, (uint64_t)N, FH) >> 3;
      const uint64_t roff = a.offset + (a.offset_dev ? *a.offset_dev : 0ull);
      const Philox4 ph = philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), (uint32_t)roff,
                                       (uint32_t)(roff >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      rbw[0] = ph.x; rbw[1] = ph.y; rbw[2] = ph.z; rbw[3] = ph.w;          // key kk, head 4g+i: 16-bit lane 4kk+i
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      tmem_ld4(tin + IN_S + g * 8 + kk * 4, sreg2[kk]);
      tmem_ld4(tin + IN_DA + g * 8 + kk * 4, dareg2[kk]);
      tmem_ld8(tin + IN_EG + g * 16 + kk * 8, egreg2[kk]);
      tmem_ld4(tin + IN_HX + g * 8 + kk * 4, hxreg2[kk]);
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ks = 2 * j + kk, m = 2 * p + kk;
      const uint32_t *sreg = sreg2[kk], *dareg = dareg2[kk], *egreg = egreg2[kk], *hxreg = hxreg2[kk];
      const uint32_t eoff = trow + (((uint32_t)ks ^ tx7) << 4);
      float x[8], r, nrm;
      ln_stats(*(const uint4 *)(es + ST_E + eoff), x, r, nrm);
      const uint4 dev = *(const uint4 *)(es + ST_DE + eoff);
      const bool kvalid = rowvalid && smask[m] != 0;
      const uint32_t rb0 = rbw[2 * kk], rb1 = rbw[2 * kk + 1];
      if (kk == 0) tmem_ld_wait();
      float dS[4], At[4], dH[4], dGv[4], Hh[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float S = __uint_as_float(sreg[i]), dA = __uint_as_float(dareg[i]);
        const float E = fmaf(r, __uint_as_float(egreg[i]), fmaf(nrm, uE[i], vE[i]));
        const float G = fmaf(r, __uint_as_float(egreg[4 + i]), fmaf(nrm, uG[i], vG[i]));
        const float Sc = fminf(fmaxf(S, lo), hi);                          // egt_layers.py:81-82
        const bool inr = S == Sc;                                          // clip passes gradient inside [lo,hi]
        Hh[i] = Sc + E;                                                    // :85-86
        bool live = kvalid;
        if (RAND) {
          const uint32_t w = i < 2 ? rb0 : rb1;
          const uint32_t bits = (i & 1) ? (w >> 16) : (w & 0xFFFFu);
          live = live && !(bits < a.rand_thr);                             // :103-108
        }
        const float pr = live ? ex2_approx(fmaf(Hh[i], kLog2e, -l2[i])) : 0.f;          // softmax probability
        const float gg = live ? sigmoid_fast(G) : 0.f;                                  // gate
        At[i] = pr * gg;
        const float dP = dA * gg;
        dH[i] = fmaf(pr, dP - Dr[i], __uint_as_float(hxreg[i]));
        const float dg = fmaf(dA, pr, ddeg[i]);
        dGv[i] = dg * fmaf(-gg, gg, gg);
        dS[i] = inr ? dH[i] : 0.f;
      }
      dsp[kk * 2 + 0] = pack_bf16(dS[0], dS[1]); dsp[kk * 2 + 1] = pack_bf16(dS[2], dS[3]);
      dzp[kk * 4 + 0] = pack_bf16(dH[0], dH[1]); dzp[kk * 4 + 1] = pack_bf16(dH[2], dH[3]);
      dzp[kk * 4 + 2] = pack_bf16(dGv[0], dGv[1]); dzp[kk * 4 + 3] = pack_bf16(dGv[2], dGv[3]);
      atp[kk * 2 + 0] = pack_bf16(At[0], At[1]); atp[kk * 2 + 1] = pack_bf16(At[2], At[3]);
      {   // weight-gradient sums: M += x^ (x) [dE|dG] ; sZ += [dE|dG] ; Wr += H^ (x) de'
        const float2 dE0 = make_float2(dH[0], dH[1]), dE1 = make_float2(dH[2], dH[3]);
        const float2 dG0 = make_float2(dGv[0], dGv[1]), dG1 = make_float2(dGv[2], dGv[3]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float xh = fmaf(r, x[c], nrm);
          const float2 xx = make_float2(xh, xh);
          ffma2(ME[c][0], xx, dE0); ffma2(ME[c][1], xx, dE1);
          ffma2(MG[c][0], xx, dG0); ffma2(MG[c][1], xx, dG1);
        }
        fadd2(sE[0], dE0); fadd2(sE[1], dE1); fadd2(sG[0], dG0); fadd2(sG[1], dG1);
        const float2 dp[4] = {make_float2(bf16_lo(dev.x), bf16_hi(dev.x)), make_float2(bf16_lo(dev.y), bf16_hi(dev.y)),
                              make_float2(bf16_lo(dev.z), bf16_hi(dev.z)), make_float2(bf16_lo(dev.w), bf16_hi(dev.w))};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 h2 = make_float2(Hh[i], Hh[i]);
#pragma unroll
          for (int q = 0; q < 4; ++q) ffma2(WR[i][q], h2, dp[q]);
        }
      }
    }
    tmem_st4(tout + g * 4, dsp);
    tmem_st8(tout + 8 + g * 8, dzp);
    // transposed operands last (the previous 16-key block has long left the tensor core by now):
    // row = query t (K index), 16-byte chunk = key, bytes 8g.. = heads 4g..4g+3
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int m = 2 * p + kk;
      const uint32_t k16 = (uint32_t)m & 15u;
      if (k16 == 0 && m > 0) mbar_wait(bar_t, ((m >> 4) - 1) & 1);   // previous block consumed
      const uint32_t off = trbase + (k16 >> 3) * 16384u + (((k16 ^ tx7) & 7u) << 4);
      *(uint2 *)(smem + SM_TR + off) = make_uint2(dsp[kk * 2], dsp[kk * 2 + 1]);
      *(uint2 *)(smem + SM_TR + 32768 + off) = make_uint2(atp[kk * 2], atp[kk * 2 + 1]);
    }
  };

  // ---- pipeline ------------------------------------------------------------------------------------
  mbar_wait(bar_e, 0);                                 // tile 0 (e, de') has landed
  fence_proxy_async_smem();                            // the dO tile written above is a tensor-core operand
  __syncthreads();                                     // sync #0
  int st_a = 0, par_a = 0;                             // stage / load parity of pair it
  for (int it = 0; it < NP; ++it) {
    const uint32_t ob8 = 8u * (uint32_t)(it & 1);
    if (it >= 2) mbar_wait(bar_mma2 + ob8, ((it - 2) >> 1) & 1);   // handshake it-2: inputs of this pair; products of pair it-2 consumed
    else mbar_wait(bar_mma1 + ob8, 0);                 // pairs 0, 1: issued before the loop
    if ((it & 3) == 0) mbar_wait(bar_e + 8 * st_a, par_a);          // first pair of a tile: e / de' are read from shared memory too
    tc_fence_after();
    phase_a(it, st_a);
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_step + ob8);                       // pair it done by this thread
    if (((it + 1) & 3) == 0) { if (++st_a == NS) { st_a = 0; par_a ^= 1; } }
  }
  __syncthreads();                                     // sync #(NP+1)
  mbar_wait(bar_t, ((NP - 1) >> 3) & 1);               // the last transposed products
  tc_fence_after();
  // dQ: all tcgen05.mma of this CTA have completed (tbar was committed last)
  {
    uint32_t o[32];
    tmem_ld32(tlane + TM_DQ + g * 32, o);              // warp-collective: never inside a divergent branch
    tmem_ld_wait();
    if (rowvalid && gridDim.z > 1) {                     // key split: the launcher zero-filled d_qkv
      float *dqf = a.d_qkv + ((size_t)b * N + l) * (3 * FD) + g * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dqf + 4 * q), "f"(__uint_as_float(o[4 * q]) * a.dq_scale),
                     "f"(__uint_as_float(o[4 * q + 1]) * a.dq_scale), "f"(__uint_as_float(o[4 * q + 2]) * a.dq_scale),
                     "f"(__uint_as_float(o[4 * q + 3]) * a.dq_scale) : "memory");
    } else if (rowvalid && gridDim.x == 1 && a.d_qkv_bf) {
      uint4 *dq = (uint4 *)(a.d_qkv_bf + ((size_t)b * N + l) * (3 * FD) + g * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(o[8 * q]) * a.dq_scale, __uint_as_float(o[8 * q + 1]) * a.dq_scale);
        v.y = pack_bf16(__uint_as_float(o[8 * q + 2]) * a.dq_scale, __uint_as_float(o[8 * q + 3]) * a.dq_scale);
        v.z = pack_bf16(__uint_as_float(o[8 * q + 4]) * a.dq_scale, __uint_as_float(o[8 * q + 5]) * a.dq_scale);
        v.w = pack_bf16(__uint_as_float(o[8 * q + 6]) * a.dq_scale, __uint_as_float(o[8 * q + 7]) * a.dq_scale);
        dq[q] = v;
      }
    } else if (rowvalid) {
      float4 *dq = (float4 *)(a.d_qkv + ((size_t)b * N + l) * (3 * FD) + g * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        dq[q] = make_float4(__uint_as_float(o[4 * q]) * a.dq_scale, __uint_as_float(o[4 * q + 1]) * a.dq_scale,
                            __uint_as_float(o[4 * q + 2]) * a.dq_scale, __uint_as_float(o[4 * q + 3]) * a.dq_scale);
    }
  }

  // ---- weight-gradient partial sums: warp shuffle -> shared atomics -> one row per CTA ----------------
  float *red = (float *)(smem + SM_TR);                // the transposed-operand buffers are idle now
  if (tid < FPART) red[tid] = 0.f;
  __syncthreads();                                     // reduction scratch zeroed
  {
    // 104 sums per thread -> 104 sums per warp.  A transpose-reduce over the 32 lanes (5 rounds: lane pairs swap halves
    // of a 32-value group and add) leaves value k of the group, summed over the warp, in lane k: 31 shuffles per 32
    // values instead of 5 per value, and one conflict-free shared atomic per lane instead of 32 from lane 0.
    auto reduce32 = [&](float (&v)[32]) {
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float send = up ? v[i] : v[i + s], keep = up ? v[i + s] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      return v[0];
    };
    float v[32];
    // values 0..63: M[c][j] with c = k >> 3 ; j = k & 7: E heads 4g..4g+3 (j < 4), G heads (j >= 4)
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * grp + cc;
        v[8 * cc + 0] = ME[c][0].x; v[8 * cc + 1] = ME[c][0].y; v[8 * cc + 2] = ME[c][1].x; v[8 * cc + 3] = ME[c][1].y;
        v[8 * cc + 4] = MG[c][0].x; v[8 * cc + 5] = MG[c][0].y; v[8 * cc + 6] = MG[c][1].x; v[8 * cc + 7] = MG[c][1].y;
      }
      const float tot = reduce32(v);
      const int k = 32 * grp + lane, c = k >> 3, j = k & 7;
      atomicAdd(red + c * 16 + ((j & 4) ? 8 : 0) + 4 * g + (j & 3), tot);
    }
    // values 64..95: Wr[i][r8] with i = kk >> 3 (head 4g + i), r8 = kk & 7
    {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[8 * i + 2 * q] = WR[i][q].x; v[8 * i + 2 * q + 1] = WR[i][q].y; }
      const float tot = reduce32(v);
      atomicAdd(red + 144 + (4 * g + (lane >> 3)) * 8 + (lane & 7), tot);
    }
    // the 8 column sums of dZ
    auto put = [&](int idx, float x) {
      x = warp_sum(x);
      if (lane == 0) atomicAdd(red + idx, x);
    };
    put(128 + 4 * g + 0, sE[0].x); put(128 + 4 * g + 1, sE[0].y); put(128 + 4 * g + 2, sE[1].x); put(128 + 4 * g + 3, sE[1].y);
    put(136 + 4 * g + 0, sG[0].x); put(136 + 4 * g + 1, sG[0].y); put(136 + 4 * g + 2, sG[1].x); put(136 + 4 * g + 3, sG[1].y);
  }
  __syncthreads();                                     // partial sums complete
  if (tid < FPART) a.partials[(size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * FPART + tid] = red[tid];
  tc_fence_before();
  __syncthreads();                                     // final
}

__global__ void __launch_bounds__(256) fused_bwd_finalize_kernel(const float *partials, int nparts, egt_block_weights_t w,
                                                                 egt_block_grads_t g) {
  fused_bwd_finalize_body(partials, nparts, w, g, threadIdx.x, blockDim.x);
}

int fused_bwd_launch(const FusedBwdArgs &a, const void *e, const void *de_out, void *de, const void *qkv,
                     cudaStream_t st) {
  CUtensorMap tm_e, tm_dei, tm_de, tm_q, tm_kv;
  const uint64_t N = a.N, B = a.B;
  int rc;
  if ((rc = encode_tmap_3d(&tm_e, e, N * FDE, N, B, N * FDE * 2, N * N * FDE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_dei, de_out, N * FDE, N, B, N * FDE * 2, N * N * FDE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_de, de, N * FDE, N, B, N * FDE * 2, N * N * FDE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_q, qkv, 3 * FD, N, B, 3 * FD * 2, N * 3 * FD * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_kv, qkv, 3 * FD, N, B, 3 * FD * 2, N * 3 * FD * 2, 64, 8, 1, 0))) return rc;
  const int smem = SM_TOTAL + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.B, fused_bwd_key_splits(a.B, a.N));
  LaunchScope _ls("fused_bwd_kernel", st);
  if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(fused_bwd_kernel<true>, grid, dim3(512), smem, st, tm_e, tm_dei, tm_de, tm_q, tm_kv, a));
  else EGT_CHECK_CUDA(launch_pdl(fused_bwd_kernel<false>, grid, dim3(512), smem, st, tm_e, tm_dei, tm_de, tm_q, tm_kv, a));
  return EGT_OK;
}

// CTAs that share the keys of one (graph, row tile), in runs of a multiple of 16 keys (see wide_bwd_key_splits,
// wide_bwd.cu).  Measured with the count the wide kernel's rule picks: S256 (64 keys per CTA) 405 -> 483 us, S512 (128 keys
// per CTA) 742 -> 776 us -- this kernel's per-CTA prologue and epilogue (row statistics, the 104 weight-gradient sums of
// every thread, dQ) cost more than the last wave wastes -- so the split is off unless EGT_FUSED_KSPLIT forces a count
// (the parity tests do).
int fused_bwd_key_splits(int B, int N) {
  (void)B;
  const char *fe = getenv("EGT_FUSED_KSPLIT");
  const int forced = fe ? atoi(fe) : 0;
  auto run = [&](int ks) { return (((N + ks - 1) / ks) + 15) & ~15; };
  if (forced > 1 && forced <= 8 && run(forced) * (forced - 1) < N) return forced;
  return 1;
}

int fused_bwd_finalize_launch(const float *partials, int nparts, const egt_block_weights_t *w,
                              const egt_block_grads_t *g, cudaStream_t st) {
  LaunchScope _ls("fused_bwd_finalize_kernel", st);
  fused_bwd_finalize_kernel<<<1, 256, 0, st>>>(partials, nparts, *w, *g);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
