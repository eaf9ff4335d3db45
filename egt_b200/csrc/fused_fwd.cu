// fused_fwd.cu -- fused forward of the EGT attention block's N x N part on sm_100a tensor cores.
//
//   reference:  EGT.call_gated (lib/models/egt_layers.py:57-143) fused with the edge projections, the
//   LayerNorm on e and the edge write-back of edge_update_residual
//   (lib/models/graph_xformer_model_base.py:192-218).  The [B,N,N,h] tensors E, G, H_hat, A~ never
//   leave the SM: e is streamed in once by TMA, e' streamed out once by TMA.
//
// One CTA = one graph b and 128 query rows; thread t of warps 0-3 owns query row l0+t == TMEM lane t.
// Keys are processed in sub-tiles of 4 (two key PAIRS).  Per key pair the tensor core produces, in TMEM,
//     S  [128 x 16]  = Q[128x64] * Kexp^T      Kexp[(key,hh), c] = K[key,c] * [c % 8 == hh]
//     EG [128 x 32]  = e_tile[128 x 16] * Wblk (raw edge channels of the two keys x folded-LN weights)
// (the per-head dot product is a block-diagonal contraction over the head-innermost channel axis, so it
// is a plain GEMM against an expanded K).  The row's thread then applies LN statistics, clip, masks,
// exp / sigmoid, accumulates the softmax denominator and the gate sum, and writes A~ and H_hat back
// to TMEM as bf16 A-operands of
//     O  [128 x 64] += A~[128 x 16] * Vexp     Vexp[(key,hh), c] = V[key,c] * [c % 8 == hh]
//     De [128 x 16]  = H_hat[128 x 16] * Wrblk (edge write-back of the two keys)
// Softmax runs without max subtraction: logits are bounded by clip + |E| <= FusedPrep::bound.
//
// Warp 4 issues TMA and tcgen05.mma; all ordering is __syncthreads + completion mbarriers, every async
// operation is issued two sub-tiles ahead of its consumer.
#include "common.cuh"
#include "fused.h"
#include "umma.cuh"

namespace egt {
using namespace umma;

namespace {

constexpr int NS = 4;                                  // e-tile stages (8 keys each)
constexpr uint32_t SM_Q = 0;                           // [128 x 128B] swizzled
constexpr uint32_t SM_STAGE = 16384;                   // NS x (16384 e + 1024 K rows + 1024 V rows)
constexpr uint32_t STAGE_BYTES = 18432;
constexpr uint32_t SM_OST = SM_STAGE + NS * STAGE_BYTES;        // 2 x 16384 e' staging
constexpr uint32_t SM_KVX = SM_OST + 2 * 16384;                 // 4 x 8192: Kexp(2 pairs) | Vexp(2 pairs)
constexpr uint32_t SM_WBLK = SM_KVX + 4 * 8192;                 // 1024
constexpr uint32_t SM_WRBLK = SM_WBLK + 1024;                   // 512
constexpr uint32_t SM_CONST = SM_WRBLK + 512;                   // 40 floats
constexpr uint32_t SM_BAR = SM_CONST + 256;                     // mbarriers
constexpr uint32_t SM_TOTAL = SM_BAR + 256;
constexpr uint32_t TM_O = 0, TM_BUF = 64, TM_BUF_COLS = 96, TM_PAIR_COLS = 48;

constexpr uint32_t ID_S = idesc_bf16(128, 16, 0, 0);
constexpr uint32_t ID_EG = idesc_bf16(128, 32, 0, 0);
constexpr uint32_t ID_PV = idesc_bf16(128, 64, 0, 1);
constexpr uint32_t ID_WR = idesc_bf16(128, 16, 0, 0);

struct Bars { uint64_t q_full, e_full[NS], mma1[3], mma2[3]; uint32_t tmem_base; };

}  // namespace

template <bool RAND>
__global__ void __launch_bounds__(160, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_eo,
                 const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                 const FusedFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
  const uint32_t sbase = smem_u32(smem);
  Bars *bars = (Bars *)(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, l0 = blockIdx.x * 128;
  const int N = a.N;
  const int NT = (N + 7) / 8, NSUB = 2 * NT;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars->q_full), 1);
      for (int i = 0; i < NS; ++i) mbar_init(smem_u32(&bars->e_full[i]), 1);
      for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&bars->mma1[i]), 1); mbar_init(smem_u32(&bars->mma2[i]), 1); }
      mbar_fence_init();
      tma_prefetch_desc(&tm_e); tma_prefetch_desc(&tm_eo); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
  } else {
    // derived weights -> smem (generic proxy writes; made visible to the tensor core by the fence below)
    const uint4 *src = (const uint4 *)a.prep->wblk;
    if (tid < 64) ((uint4 *)(smem + SM_WBLK))[tid] = src[tid];
    if (tid < 32) ((uint4 *)(smem + SM_WRBLK))[tid] = ((const uint4 *)a.prep->wrblk)[tid];
    if (tid < 40) ((float *)(smem + SM_CONST))[tid] = a.prep->uE[tid];   // uE,vE,uG,vG,br are contiguous
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  auto load_tile = [&](int T) {     // issuer lane 0
    const int st = T % NS;
    const uint32_t bar = smem_u32(&bars->e_full[st]);
    const uint32_t dst = sbase + SM_STAGE + st * STAGE_BYTES;
    mbar_expect_tx(bar, STAGE_BYTES);
    tma_load_3d(dst, &tm_e, bar, T * 64, l0, b);
    tma_load_3d(dst + 16384, &tm_kv, bar, FD, T * 8, b);
    tma_load_3d(dst + 17408, &tm_kv, bar, 2 * FD, T * 8, b);
  };

  if (warp == 4) {
    // =============================== issuer warp ===============================================
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bars->q_full), 16384);
      tma_load_3d(sbase + SM_Q, &tm_q, smem_u32(&bars->q_full), 0, l0, b);
      for (int T = 0; T < NT && T < NS; ++T) load_tile(T);
    }
    auto issue_mma1 = [&](int t) {
      const int T = t >> 1, half = t & 1, st = T % NS, buf = t % 3, slot = t & 3;
      mbar_wait(smem_u32(&bars->e_full[st]), (T / NS) & 1);
      tc_fence_after();
      const uint32_t es = sbase + SM_STAGE + st * STAGE_BYTES;
      const uint32_t kx = sbase + SM_KVX + slot * 8192;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t d = tmem + TM_BUF + buf * TM_BUF_COLS + j * TM_PAIR_COLS;
#pragma unroll
        for (int s = 0; s < 4; ++s)
          mma_ss(d, smem_desc(sbase + SM_Q + 32 * s, 16, 1024, LAYOUT_SW128),
                 smem_desc(kx + j * 2048 + 32 * s, 16, 1024, LAYOUT_SW128), ID_S, s > 0);
        mma_ss(d + 16, smem_desc(es + 32 * (half * 2 + j), 16, 1024, LAYOUT_SW128),
               smem_desc(sbase + SM_WBLK, 512, 128, LAYOUT_NONE), ID_EG, 0);
      }
      mma_commit(smem_u32(&bars->mma1[buf]));
    };
    auto issue_mma2 = [&](int t) {
      const int buf = t % 3, slot = t & 3;
      const uint32_t vx = sbase + SM_KVX + slot * 8192 + 4096;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t d = tmem + TM_BUF + buf * TM_BUF_COLS + j * TM_PAIR_COLS;
        mma_ts(tmem + TM_O, d, smem_desc(vx + j * 2048, 2048, 1024, LAYOUT_SW128), ID_PV, (t > 0 || j > 0));
        mma_ts(d + 16, d + 8, smem_desc(sbase + SM_WRBLK, 256, 128, LAYOUT_NONE), ID_WR, 0);
      }
      mma_commit(smem_u32(&bars->mma2[buf]));
    };
    __syncthreads();                                   // sync #0: Kexp/Vexp of sub-tiles 0,1 are built
    if (lane == 0) {
      tc_fence_after();
      mbar_wait(smem_u32(&bars->q_full), 0);
      issue_mma1(0);
      if (NSUB > 1) issue_mma1(1);
    }
    bool store_pending = false;
    for (int t = 0; t < NSUB; ++t) {
      __syncthreads();                                 // sync #(t+1)
      if (lane == 0) {
        tc_fence_after();
        if (store_pending) { tma_store_wait_read<0>(); store_pending = false; }
        issue_mma2(t);
        if (t + 2 < NSUB) issue_mma1(t + 2);
        if (t >= 1 && ((t - 1) & 1)) {                 // tile T is complete: store e', refill its stage
          const int T = (t - 1) >> 1;
          tma_store_3d(&tm_eo, sbase + SM_OST + (T & 1) * 16384, T * 64, l0, b);
          tma_store_commit();
          store_pending = true;
          if (T + NS < NT) load_tile(T + NS);
        }
      }
      __syncwarp();
    }
    __syncthreads();                                   // sync #(NSUB+1): last tile staged
    if (lane == 0) {
      const int T = NT - 1;
      tma_store_3d(&tm_eo, sbase + SM_OST + (T & 1) * 16384, T * 64, l0, b);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
    __syncwarp();
    __syncthreads();                                   // final
    tmem_dealloc(tmem, 512);
    return;
  }

  // ================================= row threads (warps 0-3) =====================================
  const int l = l0 + tid;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  const float *cst = (const float *)(smem + SM_CONST);
  float uE[8], vE[8], uG[8], vG[8], br[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { uE[i] = cst[i]; vE[i] = cst[8 + i]; uG[i] = cst[16 + i]; vG[i] = cst[24 + i]; br[i] = cst[32 + i]; }
  float psum[8], gsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { psum[i] = 0.f; gsum[i] = 0.f; }
  const uint8_t *maskb = a.mask ? a.mask + (size_t)b * N : nullptr;

  auto build = [&](int t2) {    // expanded K / V operands of sub-tile t2 into slot t2 & 3
    const int T2 = t2 >> 1, half2 = t2 & 1, st = T2 % NS;
    const uint8_t *rows = smem + SM_STAGE + st * STAGE_BYTES + 16384;
    uint8_t *kx = smem + SM_KVX + (t2 & 3) * 8192;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 128 * q;
      const int which = idx >> 8, rem = idx & 255, pair = rem >> 7, row = (rem >> 3) & 15, dd = rem & 7;
      const int key = row >> 3, hh = row & 7, ks = half2 * 4 + pair * 2 + key;
      const uint32_t val = *(const uint16_t *)(rows + which * 1024 + ks * 128 + (dd * 8 + hh) * 2);
      const uint32_t wv = val << ((hh & 1) * 16);
      uint4 ch;
      ch.x = (hh >> 1) == 0 ? wv : 0u; ch.y = (hh >> 1) == 1 ? wv : 0u;
      ch.z = (hh >> 1) == 2 ? wv : 0u; ch.w = (hh >> 1) == 3 ? wv : 0u;
      *(uint4 *)(kx + which * 4096 + pair * 2048 + row * 128 + (((dd ^ row) & 7) << 4)) = ch;
    }
  };

  auto phase_a = [&](int t) {
    const int T = t >> 1, half = t & 1, st = T % NS, buf = t % 3;
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
    const uint32_t tb = tlane + TM_BUF + buf * TM_BUF_COLS;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t sreg[16], egreg[32], apack[8], hpack[8];
      tmem_ld16(tb + j * TM_PAIR_COLS, sreg);
      tmem_ld32(tb + j * TM_PAIR_COLS + 16, egreg);
      tmem_ld_wait();
      float av[16], hv[16];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = half * 4 + j * 2 + kk;
        const int m = T * 8 + ks;
        const uint4 ev = *(const uint4 *)(es + sw128_off(tid, ks * 8));
        float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y),
                      bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
        float mu = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) mu += x[c];
        mu *= 0.125f;
        float var = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) { float dlt = x[c] - mu; var = fmaf(dlt, dlt, var); }
        const float r = rsqrtf(fmaf(var, 0.125f, 1e-3f));
        const float nrm = -r * mu;
        bool kvalid = m < N;
        if (maskb && kvalid) kvalid = maskb[m] != 0;
        uint32_t rbits[4] = {0u, 0u, 0u, 0u};
        if (RAND) {
          const uint64_t qd = ((uint64_t)b * N + (uint64_t)l) * N + (uint64_t)m;
          Philox4 ph = philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), (uint32_t)a.offset,
                                     (uint32_t)(a.offset >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          rbits[0] = ph.x; rbits[1] = ph.y; rbits[2] = ph.z; rbits[3] = ph.w;
        }
#pragma unroll
        for (int hh = 0; hh < 8; ++hh) {
          const float S = __uint_as_float(sreg[kk * 8 + hh]);
          const float aE = __uint_as_float(egreg[kk * 16 + hh]);
          const float aG = __uint_as_float(egreg[kk * 16 + 8 + hh]);
          const float E = fmaf(r, aE, fmaf(nrm, uE[hh], vE[hh]));
          const float G = fmaf(r, aG, fmaf(nrm, uG[hh], vG[hh]));
          const float Hh = fminf(fmaxf(S, a.clip_lo), a.clip_hi) + E;      // egt_layers.py:79-86
          bool live = kvalid;
          if (RAND) {
            const uint32_t bits = (hh & 1) ? (rbits[hh >> 1] >> 16) : (rbits[hh >> 1] & 0xFFFFu);
            live = live && !(bits < a.rand_thr);                           // :103-108
          }
          const float p = live ? exp2f(Hh * kLog2e) : 0.f;                 // :111 (unnormalised)
          const float g = live ? __fdividef(1.f, 1.f + exp2f(-G * kLog2e)) : 0.f;   // :112
          psum[hh] += p;
          gsum[hh] += g;
          av[kk * 8 + hh] = p * g;                                         // :113
          hv[kk * 8 + hh] = Hh;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { apack[i] = pack_bf16(av[2 * i], av[2 * i + 1]); hpack[i] = pack_bf16(hv[2 * i], hv[2 * i + 1]); }
      tmem_st8(tb + j * TM_PAIR_COLS, apack);
      tmem_st8(tb + j * TM_PAIR_COLS + 8, hpack);
    }
  };

  auto phase_b = [&](int t) {   // e' = e + H_hat W_r + b_r for the 4 keys of sub-tile t -> staging
    const int T = t >> 1, half = t & 1, st = T % NS, buf = t % 3;
    const uint8_t *es = smem + SM_STAGE + st * STAGE_BYTES;
    uint8_t *os = smem + SM_OST + (T & 1) * 16384;
    const uint32_t tb = tlane + TM_BUF + buf * TM_BUF_COLS;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t dreg[16];
      tmem_ld16(tb + j * TM_PAIR_COLS + 16, dreg);
      tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = half * 4 + j * 2 + kk;
        const uint32_t off = sw128_off(tid, ks * 8);
        const uint4 ev = *(const uint4 *)(es + off);
        const float x[8] = {bf16_lo(ev.x), bf16_hi(ev.x), bf16_lo(ev.y), bf16_hi(ev.y),
                            bf16_lo(ev.z), bf16_hi(ev.z), bf16_lo(ev.w), bf16_hi(ev.w)};
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = x[c] + (__uint_as_float(dreg[kk * 8 + c]) + br[c]);
        uint4 ov;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]);
        ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        *(uint4 *)(os + off) = ov;
      }
    }
  };

  // ---- pipeline ------------------------------------------------------------------------------
  mbar_wait(smem_u32(&bars->e_full[0]), 0);
  build(0);
  if (NSUB > 1) build(1);
  fence_proxy_async_smem();
  __syncthreads();                                     // sync #0
  for (int t = 0; t < NSUB; ++t) {
    mbar_wait(smem_u32(&bars->mma1[t % 3]), (t / 3) & 1);
    tc_fence_after();
    phase_a(t);
    if (t >= 1) {
      mbar_wait(smem_u32(&bars->mma2[(t - 1) % 3]), ((t - 1) / 3) & 1);
      tc_fence_after();
      phase_b(t - 1);
    }
    if (t + 2 < NSUB) {
      const int T2 = (t + 2) >> 1;
      mbar_wait(smem_u32(&bars->e_full[T2 % NS]), (T2 / NS) & 1);
      build(t + 2);
    }
    tmem_st_wait();
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();                                   // sync #(t+1)
  }
  mbar_wait(smem_u32(&bars->mma2[(NSUB - 1) % 3]), ((NSUB - 1) / 3) & 1);
  tc_fence_after();
  phase_b(NSUB - 1);
  fence_proxy_async_smem();
  __syncthreads();                                     // sync #(NSUB+1)

  // ---- row epilogue: normalise, centrality scaler, V_att, saved statistics ----------------------
  {
    float f[8];
#pragma unroll
    for (int hh = 0; hh < 8; ++hh) {
      const float inv = psum[hh] > 0.f ? __fdividef(1.f, psum[hh]) : 0.f;
      float s = 1.f;
      if (a.scale_degree && l >= a.num_virtual_nodes)                      // egt_layers.py:123-135
        s = a.scaler_type == EGT_SCALER_LOG ? log1pf(gsum[hh]) : gsum[hh];
      f[hh] = inv * s;
    }
    uint32_t o[64];
    tmem_ld32(tlane + TM_O, o);
    tmem_ld32(tlane + TM_O + 32, o + 32);
    tmem_ld_wait();
    if (l < N) {
      uint4 *dst = (uint4 *)(a.v_att + ((size_t)b * N + l) * FD);
#pragma unroll
      for (int dd = 0; dd < 8; ++dd) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(o[dd * 8 + 0]) * f[0], __uint_as_float(o[dd * 8 + 1]) * f[1]);
        v.y = pack_bf16(__uint_as_float(o[dd * 8 + 2]) * f[2], __uint_as_float(o[dd * 8 + 3]) * f[3]);
        v.z = pack_bf16(__uint_as_float(o[dd * 8 + 4]) * f[4], __uint_as_float(o[dd * 8 + 5]) * f[5]);
        v.w = pack_bf16(__uint_as_float(o[dd * 8 + 6]) * f[6], __uint_as_float(o[dd * 8 + 7]) * f[7]);
        dst[dd] = v;
      }
      const size_t ps = ((size_t)b * N + l) * FH, rs = (size_t)a.B * N * FH;
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) {
        a.lse[ps + hh] = 0.f;                                              // reference point of the exponent
        a.lse[rs + ps + hh] = psum[hh] > 0.f ? __logf(psum[hh]) : 0.f;
        a.deg[ps + hh] = gsum[hh];
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
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((a.N + 127) / 128, a.B);
  LaunchScope _ls("fused_fwd_kernel", st);
  if (a.rand_mask) fused_fwd_kernel<true><<<grid, 160, smem, st>>>(tm_e, tm_eo, tm_q, tm_kv, a);
  else fused_fwd_kernel<false><<<grid, 160, smem, st>>>(tm_e, tm_eo, tm_q, tm_kv, a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
