// wide_bwd.cu -- width-generic fused backward of the EGT attention block's N x N part on sm_100a tensor cores.
//
//   reference: TF autodiff of EGT.call_gated (lib/models/egt_layers.py:57-143) together with the edge projections,
//   LayerNorm on e and the residual edge write-back of edge_update_residual
//   (lib/models/graph_xformer_model_base.py:192-218), as derived in SURVEY.md 3.4.  Nothing of shape [B,N,N,h]
//   touches HBM: e and de' stream in once by TMA, de streams out once by TMA; E, G, H_hat, P, g, dE, dG, dS are
//   recomputed / consumed on chip.
//
// Work split as in wide_fwd.cu: one CTA = one graph and 128 query rows; thread (q, t) of compute group q owns the
// (row l0+t, key m) pairs with m = q (mod 2), every head of a pair.  Per key the tensor core produces in the group's
// tensor-memory columns
//     S   [128x16] = Qs  [128xd] * Kexp^T          dA  [128x16] = dO [128xd] * Vexp^T     (dO = dV_att (.) scaler)
//     EG  [128x2h] = e_key * W' (hi + lo)          dHx [128x16] = de'_key * W_r^T
// the pair's thread recomputes p, g, H_hat from the saved row statistics and forms
//     dP = dA g ; dH = p (dP - D) + dHx ; dG = (dA p + ddeg) g (1-g) ; dS = dH [lo <= S <= hi]
// Everything that is a contraction goes back to the tensor core:
//     dQ  [128xd]  += dS [128x16] * Kexp                                   (A from tensor memory)
//     T   [128xde]  = de'_key * I + (r dZ) [128x2h] * W'^T                 (= r dx^ + de': LayerNorm backward is then
//                                                                           de = T - (r^2 m2) e + const, per channel one FMA)
//     dK^T [d x 16] = Qs^T [d x 128] * dS_img ;  dV^T [d x 16] = dO^T * A~_img       (K = the 128 query rows; the
//                                                  images are written row by row, lane c picks head c % h)
//     W1 [64 x de] += Z_img^T * e_key ;  W2 [64 x de] += Z_img^T * de'_key           (Z_img row = [r dZ | H_hat | 1]:
//                                                  sum x^ (x) dZ, sum H_hat (x) de' and sum de' for the weight gradients)
// so the per-pair thread work is the element-wise chain plus packing.  Column sums of dZ stay in registers.
//
// One handshake per key and group (ready[q] / done[q]); the products a thread writes live in columns whose inputs
// it has consumed, and the issuer orders "consume old operands" before "produce next inputs" by issue order
// (tcgen05.mma executes in issue order).
#include "common.cuh"
#include "wide.h"
#include "wide_common.cuh"
#include "wide_bwd_program.cuh"

namespace egt {
using namespace umma;

namespace {

// ZNONE: Z_img as un-swizzled MN-major core matrices (one 2 KB column block per 8 of its rows) instead of a 16 KB
// swizzled [128 x 128 B] tile -- the same bytes per pair, half the shared memory at h = 8
template <int H_, int DK_, int DE_, int NS_, bool ZNONE_ = false>
struct WideBwdGeo : WideGeo<H_, DK_, DE_, 2, NS_> {
  using G = WideGeo<H_, DK_, DE_, 2, NS_>;
  static constexpr bool ZNONE = ZNONE_;
  static constexpr int ZCH = G::EGN / 8 + H_ / 8 + 1;      // 16-byte chunks of a Z_img row: r dZ | H_hat | 1
  static constexpr int ZBYTES = ZNONE_ ? ZCH * 2048 : 16384;
  static constexpr int WN = G::DEP * (DE_ < 16 ? 2 : 1);    // columns of one weight-gradient accumulator
  static constexpr int PART = G::EGN * DE_ + H_ * DE_ + DE_ + G::EGN;   // floats one CTA contributes
};
using WideBwdC5 = WideBwdGeo<16, 8, 32, 2>;
using WideBwdC3 = WideBwdGeo<8, 12, 8, 2>;
using WideBwdC1 = WideBwdGeo<8, 8, 64, 2, true>;

// value of column `idx` (warp-divergent index) of 16 registers
__device__ __forceinline__ float sel16(const uint32_t *o, int idx) {
  uint32_t a[8], b[4], c[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (idx & 1) ? o[2 * i + 1] : o[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = (idx & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) c[i] = (idx & 4) ? b[2 * i + 1] : b[2 * i];
  return __uint_as_float((idx & 8) ? c[1] : c[0]);
}

template <class C, bool RAND>
__global__ void __launch_bounds__(384, 1)
wide_bwd_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_dei,
                const __grid_constant__ CUtensorMap tm_de, const __grid_constant__ CUtensorMap tm_q,
                const __grid_constant__ CUtensorMap tm_kv, const WideBwdArgs a) {
  constexpr int H = C::H, DE = C::DE, D = C::D, NG = 2, NS = C::NS, TK = C::TK, KPG = C::KPG, EGN = C::EGN, DEP = C::DEP;
  // ---- shared memory map ----
  constexpr uint32_t SM_Q = 0, SM_DO = C::NQA * 16384;                 // [128 x d] tiles, 128B swizzle, 64-channel atoms
  constexpr uint32_t KV_MAT = C::NQA * 2048;
  constexpr uint32_t SM_KVX = 2 * C::NQA * 16384;                      // per group: Kexp slot 0 | Kexp slot 1 | Vexp
  constexpr uint32_t IMG_G = C::ZBYTES + 2 * 4096;                      // per group: Z_img (MN-major) | dS_img | A~_img
  constexpr uint32_t SM_IMG = SM_KVX + NG * 3 * KV_MAT;
  constexpr uint32_t SM_STAGE = SM_IMG + NG * IMG_G;
  constexpr uint32_t ST_E = 0, ST_DE = C::NBOX * 16384, ST_K = 2 * C::NBOX * 16384, ST_V = ST_K + C::KV_ROWS;
  constexpr uint32_t STAGE_BYTES = (ST_V + C::KV_ROWS + 1023) & ~1023u;
  constexpr uint32_t TX_BYTES = 2 * C::NBOX * 16384 + 2 * TK * D * 2;
  constexpr uint32_t SM_W = SM_STAGE + NS * STAGE_BYTES;               // w_eg[0..1] | w_hx[0..1] | w_dx | i16[0..1]
  constexpr uint32_t NVAR = DE < 16 ? 2 : 1;
  constexpr uint32_t W_EG_SZ = 2 * C::DEW * EGN * 2, W_HX_SZ = C::DEW * 32, W_DX_SZ = DEP * EGN * 2;
  constexpr uint32_t W_EG = 0, W_HX = NVAR * W_EG_SZ, W_DX = W_HX + NVAR * W_HX_SZ, W_I = W_DX + W_DX_SZ, W_TOTAL = W_I + 1024;
  constexpr uint32_t SM_CONST = SM_W + W_TOTAL;                        // uE vE uG vG (4 x 16 floats)
  constexpr uint32_t SM_RED = SM_CONST + 256;                          // column sums of dZ (EGN floats)
  constexpr uint32_t SM_BAR = SM_RED + 128;
  constexpr uint32_t SM_MASK = SM_BAR + 256;
  constexpr uint32_t SM_TOTAL = SM_MASK + 4096 + 64;
  static_assert(SM_TOTAL + 1024 <= 232448, "shared memory budget");
  // ---- tensor memory map ----
  constexpr uint32_t TM_DQ = 0, TM_W1 = D, TM_W2 = D + C::WN, TM_G = D + 2 * C::WN;
  constexpr uint32_t G_S = 0, G_DA = 16, G_EG = 32, G_HX = 32 + EGN, G_DX = 48 + EGN, G_T = 48 + EGN + DEP, GC = 80 + EGN + DEP;
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
  // gridDim.z > 1: the keys of a graph are split over gridDim.z CTAs (wave quantisation: 256 CTAs of equal length on 148
  // SMs take two full waves, 1024 quarter-length CTAs take 7 / 4 of one); T, kt below are LOCAL tile / key-in-tile
  // indices, T0 turns them into positions in the graph.  dQ is then accumulated across the CTAs with vector atomics.
  const int NTall = (N + TK - 1) / TK;
  const int per_split = (NTall + (int)gridDim.z - 1) / (int)gridDim.z, T0 = (int)blockIdx.z * per_split;
  const int NT = NTall - T0 < per_split ? NTall - T0 : per_split, J = NT * KPG;   // the launcher guarantees NT >= 1

  // ---------------------------------------- set-up ----------------------------------------
  if (warp == 4 * NG) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars->q_full), 1);
      for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&bars->e_full[i]), 1); mbar_init(smem_u32(&bars->tile_done[i]), NG * 128); }
      for (int i = 0; i < NG; ++i) { mbar_init(smem_u32(&bars->ready[i]), 1); mbar_init(smem_u32(&bars->done[i]), 128); }
      mbar_fence_init();
      tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_dei); tma_prefetch_desc(&tm_de); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
  }
  pdl_wait();
  {
    const int nthr = 384;
    // expanded operands and operand images start as zeros (most of their elements are never written again)
    for (int i = tid; i < (int)((SM_STAGE - SM_KVX) / 16); i += nthr) ((uint4 *)(smem + SM_KVX))[i] = make_uint4(0, 0, 0, 0);
    const WidePrep *pp = a.prep;
    for (uint32_t v = 0; v < NVAR; ++v) {
      for (int i = tid; i < (int)(W_EG_SZ / 16); i += nthr) ((uint4 *)(smem + SM_W + W_EG + v * W_EG_SZ))[i] = ((const uint4 *)pp->w_eg[v])[i];
      for (int i = tid; i < (int)(W_HX_SZ / 16); i += nthr) ((uint4 *)(smem + SM_W + W_HX + v * W_HX_SZ))[i] = ((const uint4 *)pp->w_hx[v])[i];
    }
    for (int i = tid; i < (int)(W_DX_SZ / 16); i += nthr) ((uint4 *)(smem + SM_W + W_DX))[i] = ((const uint4 *)pp->w_dx)[i];
    for (int i = tid; i < 64; i += nthr) ((uint4 *)(smem + SM_W + W_I))[i] = ((const uint4 *)pp->i16[0])[i];
    for (int i = tid; i < 64; i += nthr) ((float *)(smem + SM_CONST))[i] = pp->uE[i];
    for (int i = tid; i < 32; i += nthr) ((float *)(smem + SM_RED))[i] = 0.f;
    for (int i = tid; i < NTall * TK; i += nthr)
      smem[SM_MASK + i] = i < N ? (a.mask ? (uint8_t)(a.mask[(size_t)b * N + i] != 0) : (uint8_t)1) : (uint8_t)0;
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
    reg_dealloc<40>();                                   // 256 x 232 + 128 x 40 == 384 x 168
    auto load_tile = [&](int T) {
      const int st = T % NS;
      const uint32_t bar = bar_e0 + 8 * st, dst = sbase + SM_STAGE + st * STAGE_BYTES;
      mbar_expect_tx(bar, TX_BYTES);
#pragma unroll
      for (int x = 0; x < C::NBOX; ++x) {
        tma_load_3d(dst + ST_E + x * 16384, &tm_e, bar, (T0 + T) * TK * DE + 64 * x, l0, b);
        tma_load_3d(dst + ST_DE + x * 16384, &tm_dei, bar, (T0 + T) * TK * DE + 64 * x, l0, b);
      }
      tma_load_3d(dst + ST_K, &tm_kv, bar, D, (T0 + T) * TK, b);
      tma_load_3d(dst + ST_V, &tm_kv, bar, 2 * D, (T0 + T) * TK, b);
    };
    if (warp == 4 * NG + 1 && lane == 0) {
      mbar_expect_tx(smem_u32(&bars->q_full), C::NQA * 16384);
#pragma unroll
      for (int x = 0; x < C::NQA; ++x) tma_load_3d(sbase + SM_Q + x * 16384, &tm_q, smem_u32(&bars->q_full), 64 * x, l0, b);
      for (int T = 0; T < NT && T < NS; ++T) load_tile(T);
    }
    __syncthreads();                                     // sync B: dO tile, first expanded operands
    if (warp == 4 * NG) {
      // ====== tcgen05.mma issuer: the whole (converged) warp runs the loop, one elected lane issues (umma.cuh) ======
      constexpr uint32_t HI_SW = desc_hi(1024, LAYOUT_SW128), HI_NONE = desc_hi(128, LAYOUT_NONE), HI_TIMG = desc_hi(2048, LAYOUT_NONE);
      constexpr uint32_t ID_N16 = idesc_bf16(128, 16, 0, 0), ID_EG = idesc_bf16(128, EGN, 0, 0), ID_DX = idesc_bf16(128, DEP, 0, 0);
      constexpr uint32_t ID_DQ = idesc_bf16(128, D, 0, 1), ID_T = idesc_bf16(128, 16, 1, 1), ID_W = idesc_bf16(128, DEP, 1, 1);
      const uint32_t loQ = desc_lo(sbase + SM_Q, 16), loDO = desc_lo(sbase + SM_DO, 16);
      const uint32_t loQm = desc_lo(sbase + SM_Q, 16384), loDOm = desc_lo(sbase + SM_DO, 16384);   // MN-major: M = channel
      const uint32_t loWeg = desc_lo(sbase + SM_W + W_EG, EGN * 16), loWhx = desc_lo(sbase + SM_W + W_HX, 256);
      const uint32_t loWdx = desc_lo(sbase + SM_W + W_DX, DEP * 16), loI = desc_lo(sbase + SM_W + W_I, 256);
      // operand windows of key kt of stage st inside the e / de' boxes
      auto win = [&](int st, int kt, uint32_t which) {      // byte address of the key's K window (row 0)
        const int ch = DE >= 16 ? kt * DE : (kt & ~1) * DE;
        return sbase + SM_STAGE + st * STAGE_BYTES + which + (uint32_t)(ch >> 6) * 16384u + (uint32_t)(ch & 63) * 2u;
      };
      // k-steps of a contraction over the channels of a [128 x d] tile: runs of <= 4 steps per 64-channel atom
      auto chain_d = [&](uint32_t dcol, uint32_t loA, uint32_t loB) {
#pragma unroll
        for (int at = 0; at < C::NQA; ++at) {
          constexpr int FULL = 4;
          const int ks = C::DKS - 4 * at < FULL ? C::DKS - 4 * at : FULL;
          if (ks == 4) MmaChain<4>::ss(dcol, loA + at * 1024, HI_SW, loB + at * 128, HI_SW, ID_N16, at > 0, 2, 2);
          else if (ks == 3) MmaChain<3>::ss(dcol, loA + at * 1024, HI_SW, loB + at * 128, HI_SW, ID_N16, at > 0, 2, 2);
          else if (ks == 2) MmaChain<2>::ss(dcol, loA + at * 1024, HI_SW, loB + at * 128, HI_SW, ID_N16, at > 0, 2, 2);
          else if (ks == 1) MmaChain<1>::ss(dcol, loA + at * 1024, HI_SW, loB + at * 128, HI_SW, ID_N16, at > 0, 2, 2);
        }
      };
      // contraction over the query rows of the tile: rows >= N are zero in every operand, so ceil(rows / 16) k-steps do
      const int rows_here = N - l0 < 128 ? N - l0 : 128, KR = (rows_here + 15) >> 4;
      auto chain_rows = [&](uint32_t dcol, uint32_t loA, uint32_t hiA, uint32_t loB, uint32_t hiB, uint32_t idesc, uint32_t acc,
                            uint32_t astep, uint32_t bstep) {
        mma_chain8_n(dcol, loA, hiA, loB, hiB, idesc, acc, astep, bstep, (uint32_t)KR);
      };
      const uint32_t use_lo = a.prep->use_lo != 0;       // W' = hi + lo only when the logits are large (wide.h)
      auto issue_mma1 = [&](int q, int st, int kt, int kslot) {   // inputs of the group's next key
        const uint32_t tg = tmem + TM_G + q * GC;
        chain_d(tg + G_S, loQ, desc_lo(sbase + SM_KVX + (q * 3 + kslot) * KV_MAT, 16));
        chain_d(tg + G_DA, loDO, desc_lo(sbase + SM_KVX + (q * 3 + 2) * KV_MAT, 16));
        const uint32_t le = desc_lo(win(st, kt, ST_E), 16), ld = desc_lo(win(st, kt, ST_DE), 16);
        const uint32_t var = DE >= 16 ? 0u : (uint32_t)(kt & 1);
        constexpr int EK = C::DEW / 16;                    // W' = hi + lo (wide.h): the e window is multiplied by both
        MmaChain<EK>::ss(tg + G_EG, le, HI_SW, loWeg + var * (W_EG_SZ / 16), HI_NONE, ID_EG, 0, 2, 2 * EGN);
        if (use_lo) MmaChain<EK>::ss(tg + G_EG, le, HI_SW, loWeg + var * (W_EG_SZ / 16) + 2 * EK * EGN, HI_NONE, ID_EG, 1, 2, 2 * EGN);
        MmaChain<EK>::ss(tg + G_HX, ld, HI_SW, loWhx + var * (W_HX_SZ / 16), HI_NONE, ID_N16, 0, 2, 32);
      };
      auto issue_mma2 = [&](int q, int st, int kt, int kslot, bool first, bool first_w) {
        const uint32_t tg = tmem + TM_G + q * GC;
        // dQ += dS Kexp
        MmaChain<1>::ts(tmem + TM_DQ, tg + G_DA, desc_lo(sbase + SM_KVX + (q * 3 + kslot) * KV_MAT, 2048), HI_SW, ID_DQ, first ? 0u : 1u, 0, 0);
        // T = de' I + (r dZ) W'^T
        const uint32_t ld = desc_lo(win(st, kt, ST_DE), 16);
        const uint32_t var = DE >= 16 ? 0u : (uint32_t)(kt & 1);
#pragma unroll
        for (int s = 0; s < DEP / 16; ++s)
          MmaChain<1>::ss(tg + G_DX + 16 * s, ld + (DE >= 16 ? 2 * s : 0), HI_SW, loI + var * 32, HI_NONE, ID_N16, 0, 0, 0);
        MmaChain<EGN / 16>::ts(tg + G_DX, tg + G_S, loWdx, HI_NONE, ID_DX, 1, 8, 2 * DEP);
        // dK^T, dV^T of this key: contraction over the 128 query rows
        const uint32_t img = sbase + SM_IMG + q * IMG_G;
        chain_rows(tg + G_T, loQm, HI_SW, desc_lo(img + C::ZBYTES, 128), HI_TIMG, ID_T, 0, 128, 16);
        chain_rows(tg + G_T + 16, loDOm, HI_SW, desc_lo(img + C::ZBYTES + 4096, 128), HI_TIMG, ID_T, 0, 128, 16);
        // weight-gradient accumulators (all keys, both groups)
        const uint32_t loZ = desc_lo(img, C::ZNONE ? 128u : IMG_G);          // rows beyond the image's: whatever follows it (finite)
        constexpr uint32_t HI_Z = C::ZNONE ? HI_TIMG : HI_SW, ZSTEP = C::ZNONE ? 16u : 128u;
        const uint32_t wcol = DE >= 16 ? 0u : (uint32_t)(kt & 1) * DEP;       // d_e = 8: even / odd keys accumulate apart
        chain_rows(tmem + TM_W1 + wcol, loZ, HI_Z, desc_lo(win(st, kt, ST_E), 16384), HI_SW, ID_W, first_w ? 0u : 1u, ZSTEP, 128);
        chain_rows(tmem + TM_W2 + wcol, loZ, HI_Z, desc_lo(win(st, kt, ST_DE), 16384), HI_SW, ID_W, first_w ? 0u : 1u, ZSTEP, 128);
      };
      // one handshake = ONE asm statement (wide_bwd_program.cuh): the products of key (st, kt), the inputs of the
      // group's next key (st2, kt2) when there is one, and the commit
      auto program = [&](int q, int st, int kt, int kslot, bool first, bool first_w, bool has_next, int st2, int kt2) {
        const uint32_t tg = tmem + TM_G + q * GC, img = sbase + SM_IMG + q * IMG_G;
        const uint32_t var = DE >= 16 ? 0u : (uint32_t)(kt & 1), var2 = DE >= 16 ? 0u : (uint32_t)(kt2 & 1);
        const uint32_t wcol = DE >= 16 ? 0u : (uint32_t)(kt & 1) * DEP;
        const uint32_t kx = sbase + SM_KVX + q * 3 * KV_MAT;
        const uint32_t a_tg = tg, a_dq = tmem + TM_DQ, a_w1 = tmem + TM_W1 + wcol, a_w2 = tmem + TM_W2 + wcol;
        const uint32_t a_kc = desc_lo(kx + kslot * KV_MAT, 2048), a_ldc = desc_lo(win(st, kt, ST_DE), 16);
        const uint32_t a_i = loI + var * 32, a_s = desc_lo(img + C::ZBYTES, 128), a_a = desc_lo(img + C::ZBYTES + 4096, 128);
        const uint32_t a_z = desc_lo(img, C::ZNONE ? 128u : IMG_G);
        const uint32_t a_we = desc_lo(win(st, kt, ST_E), 16384), a_wd = desc_lo(win(st, kt, ST_DE), 16384);
        const uint32_t a_kn = desc_lo(kx + (kslot ^ 1) * KV_MAT, 16), a_vn = desc_lo(kx + 2 * KV_MAT, 16);
        const uint32_t a_len = desc_lo(win(st2, kt2, ST_E), 16), a_ldn = desc_lo(win(st2, kt2, ST_DE), 16);
        const uint32_t a_weg = loWeg + var2 * (W_EG_SZ / 16), a_whx = loWhx + var2 * (W_HX_SZ / 16);
        const uint32_t a_bar = bar_ready0 + 8 * q;
#define EGT_BWD_PROGRAM_ARGS a_kc, a_ldc, loWdx, a_i, loQm, loDOm, a_s, a_a, a_z, a_we, a_wd, first, first_w, (uint32_t)KR, has_next, \
                             loQ, loDO, a_kn, a_vn, a_len, a_ldn, a_weg, a_whx, use_lo, a_bar
        (void)a_tg; (void)a_dq; (void)a_w1; (void)a_w2;   // tensor-memory addresses are literals inside the programs
        if constexpr (H == 16) {
          if (q == 0) wide_bwd_program_c5_g0(EGT_BWD_PROGRAM_ARGS); else wide_bwd_program_c5_g1(EGT_BWD_PROGRAM_ARGS);
        } else if constexpr (DE == 8) {
          if (q == 0) wide_bwd_program_c3_g0(wcol, EGT_BWD_PROGRAM_ARGS); else wide_bwd_program_c3_g1(wcol, EGT_BWD_PROGRAM_ARGS);
        } else {
          if (q == 0) wide_bwd_program_c1_g0(EGT_BWD_PROGRAM_ARGS); else wide_bwd_program_c1_g1(EGT_BWD_PROGRAM_ARGS);
        }
#undef EGT_BWD_PROGRAM_ARGS
      };
      const bool use_program = true;
      tc_fence_after();
      mbar_wait(smem_u32(&bars->q_full), 0);
      mbar_wait(bar_e0, 0);
      tc_fence_after();
      for (int q = 0; q < NG; ++q) { issue_mma1(q, 0, q, 0); mma_commit_w(bar_ready0 + 8 * q); }
      int T = 0, i = 0, st = 0;
      for (int j = 0; j < J; ++j) {
        int T2 = T, i2 = i + 1, st2 = st;
        if (i2 == KPG) { i2 = 0; ++T2; if (++st2 == NS) st2 = 0; }
        for (int q = 0; q < NG; ++q) {
          mbar_wait(bar_done0 + 8 * q, j & 1);
          tc_fence_after();
          fence_proxy_async_smem();
          // d_e = 8: the first even and the first odd key start their own accumulators
          const bool first = j == 0 && q == 0, first_w = DE >= 16 ? first : (j == 0);
          if (j + 1 < J && i2 == 0 && q == 0) { mbar_wait(bar_e0 + 8 * st2, (T2 / NS) & 1); tc_fence_after(); }
          if (use_program) {
            program(q, st, i * NG + q, j & 1, first, first_w, j + 1 < J, st2, i2 * NG + q);
          } else {
            issue_mma2(q, st, i * NG + q, j & 1, first, first_w);
            if (j + 1 < J) issue_mma1(q, st2, i2 * NG + q, (j + 1) & 1);
            mma_commit_w(bar_ready0 + 8 * q);
          }
        }
        T = T2; i = i2; st = st2;
      }
    } else if (warp == 4 * NG + 1 && lane == 0) {
      // =============================== TMA producer ===============================
      for (int T = 0; T < NT; ++T) {
        const int st = T % NS;
        mbar_wait(bar_td0 + 8 * st, (T / NS) & 1);
        fence_proxy_async_smem();
#pragma unroll
        for (int x = 0; x < C::NBOX; ++x)
          tma_store_3d(&tm_de, sbase + SM_STAGE + st * STAGE_BYTES + ST_DE + x * 16384, (T0 + T) * TK * DE + 64 * x, l0, b);
        tma_store_commit();
        if (T + NS < NT) { tma_store_wait_read<0>(); load_tile(T + NS); }
      }
      tma_store_wait_all<0>();
    }
    __syncwarp();
    __syncthreads();                                     // sync C
    if (warp == 4 * NG) { tc_fence_after(); tmem_dealloc(tmem, 512); }
    return;
  }

  // ================================= compute threads =================================
  reg_alloc<232>();
  const int q = warp >> 2, t = tid & 127;
  const int l = l0 + t;
  const bool rowvalid = l < N;
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t tg = tlane + TM_G + q * GC;
  const float *cst = (const float *)(smem + SM_CONST);
  const uint8_t *smask = smem + SM_MASK;
  const float lo = a.clip_lo, hi = a.clip_hi, ln_eps = a.ln_eps;
  const uint32_t trow = (uint32_t)t * 128u, tx7 = (uint32_t)(t & 7);
  const uint32_t bar_ready = bar_ready0 + 8 * q, bar_done = bar_done0 + 8 * q;
  uint8_t *zimg = smem + SM_IMG + q * IMG_G;             // Z_img row t: chunks [r dZ (EGN/8) | H_hat (H/8) | ones]
  uint8_t *simg = zimg + C::ZBYTES + (uint32_t)t * 16u;  // dS_img / A~_img: n-chunk c of row t at c * 2048 + t * 16
  // 16-byte chunk c of row t of Z_img
  auto zoff = [&](uint32_t c) { return C::ZNONE ? c * 2048u + (uint32_t)t * 16u : trow + ((c ^ tx7) << 4); };
  const bool single_tile = gridDim.x == 1;

  // ---- per-row quantities: D = sum_dd dV_att V_att, scaler s, ddeg, log2 row sum; dO tile -------------
  float Dr[H], ddeg[H], l2[H];
  {
    float s8[H], deg8[H], Dacc[H];
    const size_t ps = ((size_t)b * N + (rowvalid ? l : 0)) * H, rs = (size_t)a.B * N * H;
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      deg8[hh] = a.deg[ps + hh];
      float s = 1.f;
      if (a.scale_degree && l >= a.num_virtual_nodes)                        // egt_layers.py:123-135
        s = a.scaler_type == EGT_SCALER_LOG ? log1pf(deg8[hh]) : deg8[hh];
      s8[hh] = s;
      Dacc[hh] = 0.f;
    }
    const uint4 *dvp = (const uint4 *)(a.d_v_att + ((size_t)b * N + (rowvalid ? l : 0)) * D);
    const uint4 *vp = (const uint4 *)(a.v_att + ((size_t)b * N + (rowvalid ? l : 0)) * D);
#pragma unroll
    for (int cc = 0; cc < D / 8; ++cc) {                 // chunk cc = channels 8cc .. 8cc+7; channel c = dd*h + hh
      uint4 dv = dvp[cc], vv = vp[cc];
      if (!rowvalid) { dv = make_uint4(0, 0, 0, 0); vv = dv; }
      const float d8[8] = {bf16_lo(dv.x), bf16_hi(dv.x), bf16_lo(dv.y), bf16_hi(dv.y), bf16_lo(dv.z), bf16_hi(dv.z), bf16_lo(dv.w), bf16_hi(dv.w)};
      const float v8[8] = {bf16_lo(vv.x), bf16_hi(vv.x), bf16_lo(vv.y), bf16_hi(vv.y), bf16_lo(vv.z), bf16_hi(vv.z), bf16_lo(vv.w), bf16_hi(vv.w)};
      float o8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int hh = (8 * cc + c) % H;
        Dacc[hh] = fmaf(d8[c], v8[c], Dacc[hh]);
        o8[c] = d8[c] * s8[hh];
      }
      if ((cc & 1) == q) {                               // the two groups share the staging of the dO tile
        uint4 o;
        o.x = pack_bf16(o8[0], o8[1]); o.y = pack_bf16(o8[2], o8[3]); o.z = pack_bf16(o8[4], o8[5]); o.w = pack_bf16(o8[6], o8[7]);
        *(uint4 *)(smem + SM_DO + (uint32_t)(cc >> 3) * 16384u + trow + ((((uint32_t)cc & 7u) ^ tx7) << 4)) = o;
      }
    }
#pragma unroll
    for (int hh = 0; hh < H; ++hh) {
      float dg = 0.f;
      if (a.scale_degree && l >= a.num_virtual_nodes) {
        const float ds = s8[hh] != 0.f ? Dacc[hh] / s8[hh] : 0.f;
        dg = a.scaler_type == EGT_SCALER_LOG ? ds / (1.f + deg8[hh]) : ds;
      }
      Dr[hh] = rowvalid ? Dacc[hh] : 0.f;
      ddeg[hh] = rowvalid ? dg : 0.f;
      l2[hh] = rowvalid ? (a.lse[ps + hh] + a.lse[rs + ps + hh]) * kLog2e : 0.f;   // reference point + log row sum
    }
  }
  float sZ[EGN];                                          // column sums of [dE|dG] (bias gradients), order (eg, hh)
#pragma unroll
  for (int i = 0; i < EGN; ++i) sZ[i] = 0.f;

  // expanded K / V operands (wide_fwd.cu): Kexp slot kslot, the single Vexp slot
  const uint32_t x_dst = (uint32_t)(t >> 6) * 2048u + (uint32_t)(t % H) * 128u + ((((uint32_t)(t & 63) >> 3) ^ (uint32_t)(t % H)) & 7u) * 16u;
  auto build = [&](int st, int kt, int kslot) {
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
        *(uint4 *)(dst + (kv ? 2 : kslot) * KV_MAT) = ch;
      }
    }
  };

  // ---- phase A: key m; returns the two scalars phase B needs for the LayerNorm backward of this pair ----
  auto phase_a = [&](int m, int st, int kt, float &coef, float &k0) {
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES + ST_E + (uint32_t)((kt * DE) >> 6) * 16384u + trow;
    const uint32_t cb = (uint32_t)((kt * DE) & 63) >> 3;
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
    const bool kvalid = rowvalid && smask[m] != 0;
    float m1 = 0.f, m2 = 0.f;                            // sum_c dx^_c and sum_c dx^_c x^_c through the projections
#pragma unroll
    for (int half = 0; half < H / 8; ++half) {
      uint32_t sreg[8], dareg[8], egreg[16], hxreg[8];
      tmem_ld8(tg + G_S + 8 * half, sreg);
      tmem_ld8(tg + G_DA + 8 * half, dareg);
      tmem_ld16(tg + G_EG + 16 * half, egreg);
      tmem_ld8(tg + G_HX + 8 * half, hxreg);
      uint32_t rb[4] = {0u, 0u, 0u, 0u};
      if (RAND) {
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
      tmem_ld_wait();
      float dS[8], At[8], rzE[8], rzG[8], Hh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int hh = 8 * half + i;
        const float S = __uint_as_float(sreg[i]), dA = __uint_as_float(dareg[i]);
        const float Epre = fmaf(r, __uint_as_float(egreg[i]), nrm * uE[i]);        // x^ . W'_E  (without the bias v)
        const float Gpre = fmaf(r, __uint_as_float(egreg[8 + i]), nrm * uG[i]);
        const float Sc = fminf(fmaxf(S, lo), hi);                                    // egt_layers.py:81-82
        const bool inr = S == Sc;                                                    // clip passes gradient inside [lo,hi]
        Hh[i] = Sc + (Epre + vE[i]);                                                 // :85-86
        bool live = kvalid;
        if (RAND) {
          const uint32_t w = rb[i >> 1];
          const uint32_t bits = (i & 1) ? (w >> 16) : (w & 0xFFFFu);
          live = live && !(bits < a.rand_thr);                                       // :103-108
        }
        const float pr = live ? ex2_approx(fmaf(Hh[i], kLog2e, -l2[hh])) : 0.f;    // softmax probability
        const float gg = live ? sigmoid_fast(Gpre + vG[i]) : 0.f;                    // gate
        At[i] = pr * gg;
        const float dP = dA * gg;
        const float dH = fmaf(pr, dP - Dr[hh], __uint_as_float(hxreg[i]));
        const float dg = fmaf(dA, pr, ddeg[hh]);
        const float dGv = dg * fmaf(-gg, gg, gg);
        dS[i] = inr ? dH : 0.f;
        sZ[hh] += dH;
        sZ[H + hh] += dGv;
        m1 = fmaf(dH, uE[i], fmaf(dGv, uG[i], m1));
        m2 = fmaf(dH, Epre, fmaf(dGv, Gpre, m2));
        rzE[i] = r * dH;
        rzG[i] = r * dGv;
      }
      uint32_t zp[8], sp[8], ap[4], hp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        zp[i] = pack_bf16(rzE[2 * i], rzE[2 * i + 1]); zp[4 + i] = pack_bf16(rzG[2 * i], rzG[2 * i + 1]);
        sp[i] = pack_bf16(dS[2 * i], dS[2 * i + 1]); sp[4 + i] = 0u;
        ap[i] = pack_bf16(At[2 * i], At[2 * i + 1]);
        hp[i] = pack_bf16(Hh[2 * i], Hh[2 * i + 1]);
      }
      // operands in tensor memory, over the inputs of this half that were just consumed
      tmem_st8(tg + G_S + 8 * half, zp);                 // r dZ: K index (half, eg, hh%8) as the columns of w_dx
      if (H == 8) tmem_st8(tg + G_DA, sp);               // dS, K = 16 with a zero upper half
      else tmem_st4(tg + G_DA + 4 * half, sp);
      // operand images in shared memory, row t
      *(uint4 *)(zimg + zoff(2 * half)) = make_uint4(zp[0], zp[1], zp[2], zp[3]);
      *(uint4 *)(zimg + zoff(2 * half + 1)) = make_uint4(zp[4], zp[5], zp[6], zp[7]);
      *(uint4 *)(zimg + zoff(EGN / 8 + half)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      *(uint4 *)(simg + half * 2048) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
      *(uint4 *)(simg + 4096 + half * 2048) = make_uint4(ap[0], ap[1], ap[2], ap[3]);
    }
    m1 *= 1.0f / DE; m2 *= 1.0f / DE;
    // de_c = r (dx^_c - m1 - x^_c m2) + de'_c  with x^_c = r e_c + nrm  ==  T_c + coef e_c + k0,  T_c = r dx^_c + de'_c
    coef = -r * r * m2;
    k0 = -r * fmaf(nrm, m2, m1);
  };

  // ---- phase B: key m of stage st: de in place over the de' stage, dK / dV of the key ----
  auto phase_b = [&](int m, int st, int kt, float coef, float k0) {
    uint8_t *stg = smem + SM_STAGE + st * STAGE_BYTES + (uint32_t)((kt * DE) >> 6) * 16384u + trow;
    const uint32_t cb = (uint32_t)((kt * DE) & 63) >> 3;
#pragma unroll
    for (int j0 = 0; j0 < DE / 8; j0 += 4) {
      constexpr int NCH = DE / 8 < 4 ? DE / 8 : 4;
      uint32_t dr[8 * NCH];
#pragma unroll
      for (int j = 0; j < NCH; ++j) tmem_ld8(tg + G_DX + 8 * (j0 + j), dr + 8 * j);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const uint32_t off = ((cb + j0 + j) ^ tx7) << 4;
        const uint4 ev = *(const uint4 *)(stg + ST_E + off);
        const float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y), bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = fmaf(coef, x[c], __uint_as_float(dr[8 * j + c]) + k0);
        uint4 ov;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]); ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        *(uint4 *)(stg + ST_DE + off) = ov;
      }
    }
    // dK[m, c], dV[m, c]: lane t = channel c, column c % h of the key's 16
    uint32_t tk[16], tv[16];
    tmem_ld16(tg + G_T, tk);
    tmem_ld16(tg + G_T + 16, tv);
    tmem_ld_wait();
    const float vk = sel16(tk, t % H), vv = sel16(tv, t % H);
    if (t < D && m < N) {
      float *dst = a.d_qkv + ((size_t)b * N + m) * (3 * D) + D + t;
      if (single_tile) { dst[0] = vk; dst[D] = vv; }
      else { atomicAdd(dst, vk); atomicAdd(dst + D, vv); }
    }
  };

  // ---- pipeline ----
  mbar_wait(bar_e0, 0);
  build(0, q, 0);
  {   // the constant chunk of Z_img: [1, 0, 0, ...] (sum of de' over all pairs = bias gradient of the write-back)
    *(uint4 *)(zimg + zoff(EGN / 8 + H / 8)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();                                       // sync B
  {
    int T = 0, i = 0, st = 0;
    int pst = 0, pkt = 0, pm = 0, plast = 0;
    float pcoef = 0.f, pk0 = 0.f;
    for (int j = 0; j < J; ++j) {
      const int kt = i * NG + q, m = (T0 + T) * TK + kt;
      mbar_wait(bar_ready, j & 1);
      tc_fence_after();
      if (j > 0) {
        phase_b(pm, pst, pkt, pcoef, pk0);
        if (plast) { fence_proxy_async_smem(); mbar_arrive(bar_td0 + 8 * pst); }
      }
      float coef, k0;
      phase_a(m, st, kt, coef, k0);
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
      pst = st; pkt = kt; pm = m; plast = (i == KPG - 1); pcoef = coef; pk0 = k0;
      T = T2; i = i2; st = st2;
    }
    mbar_wait(bar_ready, J & 1);
    tc_fence_after();
    phase_b(pm, pst, pkt, pcoef, pk0);
    fence_proxy_async_smem();
    mbar_arrive(bar_td0 + 8 * pst);
  }
  tc_fence_before();
  asm volatile("bar.sync 1, 256;" ::: "memory");          // every tcgen05.mma of the CTA has completed
  tc_fence_after();

  // ---- dQ ----
  {
    constexpr int CW = D / NG;
    static_assert(CW % 8 == 0, "column split of dQ");
    uint32_t o[CW];
#pragma unroll
    for (int j = 0; j < CW / 8; ++j) tmem_ld8(tlane + TM_DQ + q * CW + 8 * j, o + 8 * j);
    tmem_ld_wait();
    if (rowvalid) {
      float *dqf = a.d_qkv + ((size_t)b * N + l) * (3 * D) + q * CW;
      float4 *dq = (float4 *)dqf;
      if (gridDim.z > 1) {         // key split: the launcher zero-filled d_qkv
#pragma unroll
        for (int j = 0; j < CW / 4; ++j)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dqf + 4 * j), "f"(__uint_as_float(o[4 * j]) * a.dq_scale),
                       "f"(__uint_as_float(o[4 * j + 1]) * a.dq_scale), "f"(__uint_as_float(o[4 * j + 2]) * a.dq_scale),
                       "f"(__uint_as_float(o[4 * j + 3]) * a.dq_scale) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j)
          dq[j] = make_float4(__uint_as_float(o[4 * j]) * a.dq_scale, __uint_as_float(o[4 * j + 1]) * a.dq_scale,
                              __uint_as_float(o[4 * j + 2]) * a.dq_scale, __uint_as_float(o[4 * j + 3]) * a.dq_scale);
      }
    }
  }
  // ---- weight-gradient partial sums of this CTA ----
  float *part = a.partials + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * C::PART;
  {
    // transpose-reduce over the warp (see fused_bwd.cu): value k, summed over the 32 lanes, ends in lane k
    float *red = (float *)(smem + SM_RED);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < EGN ? sZ[i] : 0.f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float send = up ? v[i] : v[i + s], keep = up ? v[i + s] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
      }
    }
    if (lane < EGN) atomicAdd(red + lane, v[0]);
  }
  if (q == 0 && warp < 2) {   // lanes 0 .. 63 of the accumulators: rows [r dZ (EGN) | H_hat (H) | 1] of Z_img
    const int row = t;                                   // warp 0: lanes 0-31, warp 1: lanes 32-63
    uint32_t w1[C::WN], w2[C::WN];
#pragma unroll
    for (int j = 0; j < C::WN / 8; ++j) { tmem_ld8(tlane + TM_W1 + 8 * j, w1 + 8 * j); tmem_ld8(tlane + TM_W2 + 8 * j, w2 + 8 * j); }
    tmem_ld_wait();
    float v1[DE], v2[DE];
#pragma unroll
    for (int c = 0; c < DE; ++c) {
      if (DE >= 16) { v1[c] = __uint_as_float(w1[c]); v2[c] = __uint_as_float(w2[c]); }
      else {          // even keys: columns 0..7 of the first accumulator, odd keys: columns 8..15 of the second
        v1[c] = __uint_as_float(w1[c]) + __uint_as_float(w1[DEP + 8 + c]);
        v2[c] = __uint_as_float(w2[c]) + __uint_as_float(w2[DEP + 8 + c]);
      }
    }
    if (row < EGN) {                                     // Mraw[j][c] = sum r dZ_j e_c
#pragma unroll
      for (int c = 0; c < DE; ++c) part[row * DE + c] = v1[c];
    } else if (row < EGN + H) {                          // sum H_hat_hh de'_c
#pragma unroll
      for (int c = 0; c < DE; ++c) part[EGN * DE + (row - EGN) * DE + c] = v2[c];
    } else if (row == EGN + H) {                         // sum de'_c
#pragma unroll
      for (int c = 0; c < DE; ++c) part[EGN * DE + H * DE + c] = v2[c];
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (tid < EGN) part[EGN * DE + H * DE + DE + tid] = ((const float *)(smem + SM_RED))[tid];
  tc_fence_before();
  __syncthreads();                                       // sync C
}

// Folds the per-CTA partial sums into the weight gradients (the library ADDS into them), see fused_prep.cuh.
//   partial row layout: Mraw[j][c] (j = column of [E|G] in the order (hh/8, eg, hh%8)) | Wr[hh][c] | dbr[c] | sZ[(eg, hh)]
//   sum x^_c dZ_j = Mraw[j][c] - mean_c' Mraw[j][c']        (x^ = r e - r mu, and sum_c' e_c' / d_e = mu)
// Phase 1 (every CTA of the 2-D grid): column sums of a slice of the partial rows, added to `sums`.  Phase 2 (the CTA
// that finishes last, found with a counter behind `sums`): the fold.  `sums` and the counter are zeroed by the caller.
__global__ void __launch_bounds__(128) wide_bwd_finalize_kernel(const float *partials, int nparts, int H, int DE, float *sums,
                                                                egt_block_weights_t w, egt_block_grads_t g) {
  extern __shared__ float s[];
  __shared__ int is_last;
  const int EGN = 2 * H, PART = EGN * DE + H * DE + DE + EGN;
  const int tid = threadIdx.x, nthr = 128;
  {
    const int col = blockIdx.x * 128 + tid;
    if (col < PART) {
      const int per = (nparts + gridDim.y - 1) / gridDim.y, i0 = blockIdx.y * per, i1 = min(nparts, i0 + per);
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      int i = i0;
      for (; i + 8 <= i1; i += 8)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += partials[(size_t)(i + u) * PART + col];
      for (; i < i1; ++i) acc[0] += partials[(size_t)i * PART + col];
      atomicAdd(sums + col, ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])));
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd((unsigned *)(sums + PART), 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int col = tid; col < PART; col += nthr) s[col] = __ldcg(sums + col);
  __syncthreads();
  float *Mraw = s, *Wr = s + EGN * DE, *dbr = Wr + H * DE, *sZ = dbr + DE;
  float *mean = sZ + EGN;                                // [EGN]
  for (int j = tid; j < EGN; j += nthr) {
    float m = 0.f;
    for (int c = 0; c < DE; ++c) m += Mraw[j * DE + c];
    mean[j] = m / DE;
  }
  __syncthreads();
  for (int idx = tid; idx < 2 * DE * H; idx += nthr) {   // dW_E, dW_G
    const int eg = idx / (DE * H), c = (idx / H) % DE, hh = idx % H;
    const int j = (hh >> 3) * 16 + eg * 8 + (hh & 7);
    const float M = Mraw[j * DE + c] - mean[j];
    float *dst = eg ? g.attention_gates_kernel : g.dense_edge_b_kernel;
    dst[c * H + hh] += w.norm_edge_gamma[c] * M + w.norm_edge_beta[c] * sZ[eg * H + hh];
  }
  for (int idx = tid; idx < 2 * H; idx += nthr) {        // db_E, db_G
    const int eg = idx / H, hh = idx % H;
    (eg ? g.attention_gates_bias : g.dense_edge_b_bias)[hh] += sZ[idx];
  }
  for (int idx = tid; idx < H * DE; idx += nthr) g.dense_edge_r_kernel[idx] += Wr[idx];
  for (int c = tid; c < DE; c += nthr) {
    g.dense_edge_r_bias[c] += dbr[c];
    float dg = 0.f, db = 0.f;
    for (int hh = 0; hh < H; ++hh) {
      const int jE = (hh >> 3) * 16 + (hh & 7), jG = jE + 8;
      const float we = w.dense_edge_b_kernel[c * H + hh], wg = w.attention_gates_kernel[c * H + hh];
      dg += we * (Mraw[jE * DE + c] - mean[jE]) + wg * (Mraw[jG * DE + c] - mean[jG]);
      db += we * sZ[hh] + wg * sZ[H + hh];
    }
    g.norm_edge_gamma[c] += dg;
    g.norm_edge_beta[c] += db;
  }
}

template <class C>
int launch_cfg(const WideBwdArgs &a, const void *e, const void *de_out, void *de, const void *qkv, cudaStream_t st) {
  CUtensorMap tm_e, tm_dei, tm_de, tm_q, tm_kv;
  const uint64_t N = a.N, B = a.B, DE = C::DE, D = C::D;
  int rc;
  if ((rc = encode_tmap_3d(&tm_e, e, N * DE, N, B, N * DE * 2, N * N * DE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_dei, de_out, N * DE, N, B, N * DE * 2, N * N * DE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_de, de, N * DE, N, B, N * DE * 2, N * N * DE * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_q, qkv, 3 * D, N, B, 3 * D * 2, N * 3 * D * 2, 64, 128, 1, 1))) return rc;
  if ((rc = encode_tmap_3d(&tm_kv, qkv, 3 * D, N, B, 3 * D * 2, N * 3 * D * 2, (uint32_t)D, C::TK, 1, 0))) return rc;
  const int smem = 232448;                               // the kernel's map is checked against this at compile time
  static bool attr_set = false;
  if (!attr_set) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_bwd_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(wide_bwd_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.B, wide_bwd_key_splits(a.B, a.N, C::TK));
  LaunchScope _ls("wide_bwd_kernel", st);
  if (a.rand_mask) EGT_CHECK_CUDA(launch_pdl(wide_bwd_kernel<C, true>, grid, dim3(384), smem, st, tm_e, tm_dei, tm_de, tm_q, tm_kv, a));
  else EGT_CHECK_CUDA(launch_pdl(wide_bwd_kernel<C, false>, grid, dim3(384), smem, st, tm_e, tm_dei, tm_de, tm_q, tm_kv, a));
  return EGT_OK;
}

}  // namespace

// How many CTAs share the keys of one (graph, row tile): the count in {1, 2, 4, 8} that minimises full waves / count on
// 148 SMs, with enough keys per CTA (its prologue -- Q, dO, the row statistics -- is paid per CTA) and no empty split.
int wide_bwd_key_splits(int B, int N, int TK) {
  const char *fe = getenv("EGT_WIDE_KSPLIT");            // read per call: the parity tests force a split
  const int forced = fe ? atoi(fe) : 0;
  const int base = B * ((N + 127) / 128), NT = (N + TK - 1) / TK;
  int best = 1;
  double cost = (double)((base + 147) / 148);
  for (int ks = 2; ks <= 8; ks *= 2) {
    const int per = (NT + ks - 1) / ks;
    // at least 64 keys per CTA; 32 when every CTA still runs in the first wave (an under-filled grid: the split only adds SMs)
    if (per * (ks - 1) >= NT || per * TK < (base * ks <= 148 ? 32 : 64)) continue;
    const double c = (double)((base * ks + 147) / 148) / ks + 0.01 * ks;
    if (c < cost - 1e-9) { cost = c; best = ks; }
  }
  if (forced > 0) {
    const int per = (NT + forced - 1) / forced;
    if (forced <= 8 && per * (forced - 1) < NT) best = forced;
  }
  return best;
}
static int wide_bwd_tk(const egt_block_cfg_t *cfg) {
  const int h = cfg->attn.h, dk = cfg->attn.dk, de = cfg->d_e;
  if (h == 16 && dk == 8 && de == 32) return WideBwdC5::TK;
  if (h == 8 && dk == 12 && de == 8) return WideBwdC3::TK;
  return WideBwdC1::TK;
}
int wide_bwd_splits(const egt_block_cfg_t *cfg) { return wide_bwd_key_splits(cfg->attn.B, cfg->attn.N, wide_bwd_tk(cfg)); }

bool wide_bwd_supported(const egt_block_cfg_t *cfg) {
  const int h = cfg->attn.h, dk = cfg->attn.dk, de = cfg->d_e;
  return (h == 16 && dk == 8 && de == 32) || (h == 8 && dk == 12 && de == 8) || (h == 8 && dk == 8 && de == 64);
}

size_t wide_bwd_partials_floats(const egt_block_cfg_t *cfg) {
  const size_t H = cfg->attn.h, DE = cfg->d_e, EGN = 2 * H;
  return ((size_t)cfg->attn.B * ((cfg->attn.N + 127) / 128) * wide_bwd_splits(cfg) + 1) * (EGN * DE + H * DE + DE + EGN) + 4;   // + the column sums and a counter
}

int wide_bwd_launch(const egt_block_cfg_t *cfg, const WideBwdArgs &a, const void *e, const void *de_out, void *de,
                    const void *qkv, cudaStream_t st) {
  const int h = cfg->attn.h, dk = cfg->attn.dk, de_w = cfg->d_e;
  if (h == 16 && dk == 8 && de_w == 32) return launch_cfg<WideBwdC5>(a, e, de_out, de, qkv, st);
  if (h == 8 && dk == 12 && de_w == 8) return launch_cfg<WideBwdC3>(a, e, de_out, de, qkv, st);
  if (h == 8 && dk == 8 && de_w == 64) return launch_cfg<WideBwdC1>(a, e, de_out, de, qkv, st);
  EGT_REQUIRE(false, EGT_E_SHAPE, "wide_bwd: no instantiation for h=%d dk=%d d_e=%d", h, dk, de_w);
}

int wide_bwd_finalize_launch(const egt_block_cfg_t *cfg, const float *partials, const egt_block_weights_t *w,
                             const egt_block_grads_t *g, const WidePrep *, cudaStream_t st) {
  const int H = cfg->attn.h, DE = cfg->d_e, EGN = 2 * H;
  const int nparts = cfg->attn.B * ((cfg->attn.N + 127) / 128) * wide_bwd_splits(cfg);
  const int PART = EGN * DE + H * DE + DE + EGN;
  const size_t smem = (size_t)(PART + EGN) * sizeof(float);
  float *sums = const_cast<float *>(partials) + (size_t)nparts * PART;
  EGT_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)(PART + 1) * sizeof(float), st));
  const int ysplit = nparts >= 64 ? 16 : nparts >= 8 ? 4 : 1;
  LaunchScope _ls("wide_bwd_finalize_kernel", st);
  wide_bwd_finalize_kernel<<<dim3((PART + 127) / 128, ysplit), 128, smem, st>>>(partials, nparts, H, DE, sums, *w, *g);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
