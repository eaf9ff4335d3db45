// attn_fast.cu -- shape-specialised staged EGT-layer kernels (lib/models/egt_layers.py:57-143, :145-213).
//
// Same three passes and the same arithmetic per (l, m, head) as attn_staged.cu (which stays as the any-shape
// fallback), restructured for the memory system:
//   * a CTA owns a block of rows (queries in the forward / row pass, keys in the column pass) of one graph; a
//     thread owns one row and HPT (4, 2 or 1) consecutive heads, so E / G / M / H_hat / dE / dG move as one vector
//     access per pair and a warp touches one contiguous span;
//   * the other side's vectors (K,V or Q,dV_att) are staged 32 rows at a time in shared memory as fp32
//     [row][head group][head][dk] (+4 floats per group, which spreads the groups over the banks) and read back
//     as broadcast float4, instead of 2*dk scalar global loads per element;
//   * sizes (h, padded dk) and the rarely used features (attention mask, random mask, dropout) are compile-time,
//     so every per-head vector lives in registers and the common case carries no dead branches.
// With a saved V_att the row pass takes  D = sum_dd dV_att * V_att  (= s * sum_m A~ dA) and skips its first sweep;
// with workspace for dS and s*A~ (block-level backward) the column pass is two plain accumulations.
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include <stdlib.h>

namespace egt {
namespace {
using umma::bf16_hi;
using umma::bf16_lo;
using umma::pack_bf16;

constexpr int AKB = 32;     // staged rows of the other side per block

template <typename T, int V> __device__ __forceinline__ void loadv(const T *p, float *x) {
  if constexpr (sizeof(T) == 4 && V == 4) {
    const float4 v = __ldg((const float4 *)p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  } else if constexpr (sizeof(T) == 4 && V == 2) {
    const float2 v = __ldg((const float2 *)p);
    x[0] = v.x; x[1] = v.y;
  } else if constexpr (V == 4) {
    const uint2 v = __ldg((const uint2 *)p);
    x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
  } else if constexpr (V == 2) {
    const uint32_t v = __ldg((const uint32_t *)p);
    x[0] = bf16_lo(v); x[1] = bf16_hi(v);
  } else {
    x[0] = ldf(p);
  }
}
template <typename T, int V> __device__ __forceinline__ void storev(T *p, const float *x) {
  if constexpr (sizeof(T) == 4 && V == 4) *(float4 *)p = make_float4(x[0], x[1], x[2], x[3]);
  else if constexpr (sizeof(T) == 4 && V == 2) *(float2 *)p = make_float2(x[0], x[1]);
  else if constexpr (V == 4) *(uint2 *)p = make_uint2(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]));
  else if constexpr (V == 2) *(uint32_t *)p = pack_bf16(x[0], x[1]);
  else stf(p, x[0]);
}

// 1/(1+exp(-x)) with the hardware reciprocal; x = -1e9 gives exp -> +inf -> exactly 0 (the masked-gate contract)
__device__ __forceinline__ float sigmoid_rcp(float x) { return umma::rcp_approx(1.0f + __expf(-x)); }

template <int DKP>
__device__ __forceinline__ float dot_s(const float *a, const float *smem_b) {
  float s = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < DKP / 4; ++d4) {
    const float4 v = ((const float4 *)smem_b)[d4];
    s = fmaf(a[4 * d4 + 0], v.x, s); s = fmaf(a[4 * d4 + 1], v.y, s);
    s = fmaf(a[4 * d4 + 2], v.z, s); s = fmaf(a[4 * d4 + 3], v.w, s);
  }
  return s;
}
template <int DKP>
__device__ __forceinline__ void axpy_s(float a, const float *smem_x, float *y) {
#pragma unroll
  for (int d4 = 0; d4 < DKP / 4; ++d4) {
    const float4 v = ((const float4 *)smem_x)[d4];
    y[4 * d4 + 0] = fmaf(a, v.x, y[4 * d4 + 0]); y[4 * d4 + 1] = fmaf(a, v.y, y[4 * d4 + 1]);
    y[4 * d4 + 2] = fmaf(a, v.z, y[4 * d4 + 2]); y[4 * d4 + 3] = fmaf(a, v.w, y[4 * d4 + 3]);
  }
}

// masks in the reference's order (egt_layers.py:89-108), dropout keep factor (:116-117).
// PLAIN: no attention mask, no random mask, no dropout (inference, and training without those features).
struct Elem { float S_raw, Hh, x, gin, keep; };
template <bool PLAIN>
__device__ __forceinline__ Elem eval_elem(const AttnParams &P, float dot, float Ev, float Gv, float negkey, float Mv,
                                          int b, int l, int m, int hh) {
  Elem e;
  float s = dot * P.scale;                                        // :79
  e.S_raw = s;
  if (P.has_clip) s = fminf(fmaxf(s, P.clip_lo), P.clip_hi);      // :81-82
  e.Hh = P.E ? s + Ev : s;                                        // :85-86
  e.x = e.Hh;
  e.gin = P.G ? Gv : 0.f;
  if (P.mask) { e.x += negkey; e.gin += negkey; }                 // :91-94
  e.keep = 1.f;
  if constexpr (!PLAIN) {
    if (P.attn_mask != EGT_MASK_NONE) {                           // :96-101
      const float neg = (Mv - 1.f) * kNegMask;
      e.x += neg; e.gin += neg;
    }
    if (P.rand_mask) {                                            // :103-108
      const float u = rng_uniform(P.seed, P.offset + (P.offset_dev ? *P.offset_dev : 0ull), 0u, rng_elem_index(b, l, m, hh, P.N, P.h));
      const float neg = u < P.random_mask_prob ? -kNegMask : 0.f;
      e.x += neg; e.gin += neg;
    }
    if (P.dropout) {
      const float u = rng_uniform(P.seed, P.offset + (P.offset_dev ? *P.offset_dev : 0ull), 1u, rng_elem_index(b, l, m, hh, P.N, P.h));
      e.keep = u >= P.attn_dropout ? 1.f / (1.f - P.attn_dropout) : 0.f;
    }
  }
  return e;
}

__device__ __forceinline__ float scaler_fast(const AttnParams &P, int l, float deg) {
  if (!P.scale_degree || l < P.num_virtual_nodes) return 1.f;    // egt_layers.py:123-135
  return P.scaler_type == EGT_SCALER_LOG ? log1pf(deg) : deg;
}

// Compile-time geometry of one kernel family.
template <int H_, int DKP_, int HPT_>
struct Geo {
  static constexpr int H = H_, DKP = DKP_, HPT = HPT_;
  static constexpr int HG = H / HPT;                      // head groups == threads per row
  static constexpr int ROWS = (HG >= 16) ? 16 : (HG >= 8) ? 32 : 64;   // rows per CTA
  static constexpr int NT = ROWS * HG;                    // threads per CTA
  // floats per staged head group; GS/4 odd, so the groups a warp reads start in different banks
  static constexpr int GS = HPT * DKP + (((HPT * DKP / 4) & 1) ? 8 : 4);
  static constexpr int RS = HG * GS;                      // floats per staged row
};

// stages rows [r0, r0+AKB) of two [N,*] matrices whose row r starts at pa/pb + r*stride_{a,b}; channel = dd*H + hh
template <typename T, typename G>
__device__ __forceinline__ void stage_pair(float *As, float *Bs, const T *pa, size_t stride_a, const T *pb,
                                           size_t stride_b, int r0, int N, int dk, int tid) {
  for (int idx = tid; idx < AKB * G::DKP * G::H; idx += G::NT) {
    const int rk = idx / (G::DKP * G::H), dd = (idx / G::H) % G::DKP, hh = idx % G::H, r = r0 + rk;
    float av = 0.f, bv = 0.f;
    if (r < N && dd < dk) {
      av = ldf(pa + (size_t)r * stride_a + dd * G::H + hh);
      bv = ldf(pb + (size_t)r * stride_b + dd * G::H + hh);
    }
    const int off = rk * G::RS + (hh / G::HPT) * G::GS + (hh % G::HPT) * G::DKP + dd;
    As[off] = av;
    Bs[off] = bv;
  }
}

template <typename T, int HPT>
__device__ __forceinline__ void load_masks(const AttnParams &P, size_t pair, size_t pe, float *Mv) {
  if (P.attn_mask == EGT_MASK_DENSE) loadv<T, HPT>((const T *)P.M + pe, Mv);
  else if (P.attn_mask == EGT_MASK_ADJ_U8) {
    const float a = (float)((const uint8_t *)P.M)[pair];
#pragma unroll
    for (int i = 0; i < HPT; ++i) Mv[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T, typename G, bool PLAIN>
__global__ void __launch_bounds__(G::NT) attn_fwd_fast(AttnParams P) {
  constexpr int H = G::H, DKP = G::DKP, HPT = G::HPT, HG = G::HG;
  __shared__ __align__(16) float Ks[AKB * G::RS];
  __shared__ __align__(16) float Vs[AKB * G::RS];
  __shared__ float negs[AKB];
  const int tid = threadIdx.x, hg = tid % HG, r = tid / HG;
  const int b = blockIdx.y, N = P.N, dk = P.dk, d = H * dk;
  const int l = blockIdx.x * G::ROWS + r;
  const bool rowvalid = l < N;
  const int lc = rowvalid ? l : N - 1;
  const T *qkv = (const T *)P.qkv + (size_t)b * N * 3 * d;
  float q[HPT][DKP], o[HPT][DKP], mrun[HPT], sum[HPT], deg[HPT];
#pragma unroll
  for (int i = 0; i < HPT; ++i) {
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd) {
      q[i][dd] = dd < dk ? ldf(qkv + (size_t)lc * 3 * d + dd * H + HPT * hg + i) : 0.f;
      o[i][dd] = 0.f;
    }
    mrun[i] = -INFINITY; sum[i] = 0.f; deg[i] = 0.f;
  }
  const size_t rowbase = ((size_t)b * N + lc) * N;
  for (int m0 = 0; m0 < N; m0 += AKB) {
    __syncthreads();
    stage_pair<T, G>(Ks, Vs, qkv + d, (size_t)3 * d, qkv + 2 * d, (size_t)3 * d, m0, N, dk, tid);
    if (tid < AKB) negs[tid] = (P.mask && m0 + tid < N) ? ((float)P.mask[(size_t)b * N + m0 + tid] - 1.f) * kNegMask : 0.f;
    __syncthreads();
    const int kend = N - m0 < AKB ? N - m0 : AKB;
    for (int mk = 0; mk < kend; ++mk) {
      const size_t pair = rowbase + m0 + mk, pe = pair * H + HPT * hg;
      float Ev[HPT], Gv[HPT], Mv[HPT], Hh[HPT];
#pragma unroll
      for (int i = 0; i < HPT; ++i) { Ev[i] = 0.f; Gv[i] = 0.f; Mv[i] = 1.f; }
      if (P.E) loadv<T, HPT>((const T *)P.E + pe, Ev);
      if (P.G) loadv<T, HPT>((const T *)P.G + pe, Gv);
      if constexpr (!PLAIN) load_masks<T, HPT>(P, pair, pe, Mv);
      const float nk = negs[mk];
      const float *kg = Ks + mk * G::RS + hg * G::GS, *vg = Vs + mk * G::RS + hg * G::GS;
#pragma unroll
      for (int i = 0; i < HPT; ++i) {
        const Elem e = eval_elem<PLAIN>(P, dot_s<DKP>(q[i], kg + i * DKP), Ev[i], Gv[i], nk, Mv[i], b, lc, m0 + mk, HPT * hg + i);
        Hh[i] = e.Hh;
        if (e.x > mrun[i]) {                                     // online softmax, as attn_staged.cu
          const float corr = __expf(mrun[i] - e.x);
          sum[i] *= corr;
#pragma unroll
          for (int dd = 0; dd < DKP; ++dd) o[i][dd] *= corr;
          mrun[i] = e.x;
        }
        const float p = __expf(e.x - mrun[i]);
        sum[i] += p;
        float g = 1.f;
        if (P.G) { g = sigmoid_rcp(e.gin); deg[i] += g; }
        axpy_s<DKP>(p * g * e.keep, vg + i * DKP, o[i]);
      }
      if (P.h_hat && rowvalid) storev<T, HPT>((T *)P.h_hat + pe, Hh);
    }
  }
  if (!rowvalid) return;
  T *vo = (T *)P.v_att + ((size_t)b * N + l) * d;
  const size_t rs = (size_t)P.B * N * H;
#pragma unroll
  for (int i = 0; i < HPT; ++i) {
    const int hh = HPT * hg + i;
    const size_t ps = ((size_t)b * N + l) * H + hh;
    const float inv = 1.f / sum[i], s = scaler_fast(P, l, deg[i]);
    P.lse[ps] = mrun[i];                     // (row max, log row sum) are kept apart, see attn_staged.cu
    P.lse[rs + ps] = __logf(sum[i]);
    P.deg[ps] = deg[i];
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd)
      if (dd < dk) stf(vo + dd * H + hh, o[i][dd] * inv * s);
  }
}

// backward, row pass: thread (l, HPT heads) owns dQ; writes dE (= dH_hat), dG, optionally H_hat, and the row
// terms D and s for the column pass                                                       
template <typename T, typename G, bool PLAIN>
__global__ void __launch_bounds__(G::NT) attn_bwd_row_fast(AttnParams P) {
  constexpr int H = G::H, DKP = G::DKP, HPT = G::HPT, HG = G::HG;
  __shared__ __align__(16) float Ks[AKB * G::RS];
  __shared__ __align__(16) float Vs[AKB * G::RS];
  __shared__ float negs[AKB];
  const int tid = threadIdx.x, hg = tid % HG, r = tid / HG;
  const int b = blockIdx.y, N = P.N, dk = P.dk, d = H * dk;
  const int l = blockIdx.x * G::ROWS + r;
  const bool rowvalid = l < N;
  const int lc = rowvalid ? l : N - 1;
  const T *qkv = (const T *)P.qkv + (size_t)b * N * 3 * d;
  const size_t rs = (size_t)P.B * N * H;
  float q[HPT][DKP], dva[HPT][DKP], dq[HPT][DKP], lse[HPT], lsum[HPT], sc[HPT], D[HPT], ddeg[HPT], ds[HPT];
#pragma unroll
  for (int i = 0; i < HPT; ++i) {
    const int hh = HPT * hg + i;
    const size_t ps = ((size_t)b * N + lc) * H + hh;
    lse[i] = P.lse[ps]; lsum[i] = P.lse[rs + ps];
    const float deg = P.deg[ps];
    sc[i] = scaler_fast(P, lc, deg);
    float dot = 0.f;
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd) {
      const bool in = dd < dk;
      q[i][dd] = in ? ldf(qkv + (size_t)lc * 3 * d + dd * H + hh) : 0.f;
      dva[i][dd] = in ? ldf((const T *)P.d_v_att + ((size_t)b * N + lc) * d + dd * H + hh) : 0.f;
      dq[i][dd] = 0.f;
      if (P.v_att && in) dot = fmaf(dva[i][dd], ldf((const T *)P.v_att + ((size_t)b * N + lc) * d + dd * H + hh), dot);
    }
    D[i] = dot;                                                   // = s * sum_m A~ dA when V_att was saved
    ds[i] = sc[i] != 0.f ? dot / sc[i] : 0.f;
    ddeg[i] = deg;                                                // finished below
  }
  const size_t rowbase = ((size_t)b * N + lc) * N;
  for (int pass = P.v_att ? 1 : 0; pass < 2; ++pass) {
    if (pass == 1) {
#pragma unroll
      for (int i = 0; i < HPT; ++i) {
        if (!P.v_att) D[i] = sc[i] * ds[i];
        const float deg = ddeg[i];
        ddeg[i] = 0.f;
        if (P.scale_degree && lc >= P.num_virtual_nodes)
          ddeg[i] = P.scaler_type == EGT_SCALER_LOG ? ds[i] / (1.f + deg) : ds[i];
        if (rowvalid) {
          const size_t ps = ((size_t)b * N + l) * H + HPT * hg + i;
          P.row_ws[ps] = D[i];
          P.row_ws[rs + ps] = sc[i];
        }
      }
    }
    for (int m0 = 0; m0 < N; m0 += AKB) {
      __syncthreads();
      stage_pair<T, G>(Ks, Vs, qkv + d, (size_t)3 * d, qkv + 2 * d, (size_t)3 * d, m0, N, dk, tid);
      if (tid < AKB) negs[tid] = (P.mask && m0 + tid < N) ? ((float)P.mask[(size_t)b * N + m0 + tid] - 1.f) * kNegMask : 0.f;
      __syncthreads();
      const int kend = N - m0 < AKB ? N - m0 : AKB;
      for (int mk = 0; mk < kend; ++mk) {
        const size_t pair = rowbase + m0 + mk, pe = pair * H + HPT * hg;
        float Ev[HPT], Gv[HPT], Mv[HPT], dhh[HPT], Hh[HPT], dEo[HPT], dGo[HPT], dSo[HPT], Aso[HPT];
#pragma unroll
        for (int i = 0; i < HPT; ++i) { Ev[i] = 0.f; Gv[i] = 0.f; Mv[i] = 1.f; dhh[i] = 0.f; }
        if (P.E) loadv<T, HPT>((const T *)P.E + pe, Ev);
        if (P.G) loadv<T, HPT>((const T *)P.G + pe, Gv);
        if constexpr (!PLAIN) load_masks<T, HPT>(P, pair, pe, Mv);
        if (pass == 1 && P.d_h_hat) loadv<T, HPT>((const T *)P.d_h_hat + pe, dhh);
        const float nk = negs[mk];
        const float *kg = Ks + mk * G::RS + hg * G::GS, *vg = Vs + mk * G::RS + hg * G::GS;
#pragma unroll
        for (int i = 0; i < HPT; ++i) {
          const Elem e = eval_elem<PLAIN>(P, dot_s<DKP>(q[i], kg + i * DKP), Ev[i], Gv[i], nk, Mv[i], b, lc, m0 + mk, HPT * hg + i);
          const float p = __expf((e.x - lse[i]) - lsum[i]);
          const float g = P.G ? sigmoid_rcp(e.gin) : 1.f;
          const float dAp = dot_s<DKP>(dva[i], vg + i * DKP);
          if (pass == 0) {
            ds[i] += p * g * e.keep * dAp;
          } else {
            const float dA = sc[i] * dAp * e.keep;
            const float dH = p * (dA * g - D[i]) + dhh[i];
            Hh[i] = e.Hh;
            dEo[i] = dH;
            dGo[i] = (dA * p + ddeg[i]) * g * (1.f - g);
            const bool inside = !P.has_clip || (e.S_raw >= P.clip_lo && e.S_raw <= P.clip_hi);
            dSo[i] = inside ? dH * P.scale : 0.f;
            Aso[i] = p * g * e.keep * sc[i];
            axpy_s<DKP>(dSo[i], kg + i * DKP, dq[i]);
          }
        }
        if (pass == 1 && rowvalid) {
          if (P.h_hat) storev<T, HPT>((T *)P.h_hat + pe, Hh);
          if (P.dG) storev<T, HPT>((T *)P.dG + pe, dGo);
          if (P.dE) storev<T, HPT>((T *)P.dE + pe, dEo);
          if (P.dS_ws) { storev<T, HPT>((T *)P.dS_ws + pe, dSo); storev<T, HPT>((T *)P.As_ws + pe, Aso); }
        }
      }
    }
  }
  if (!rowvalid) return;
  T *dqo = (T *)P.d_qkv + ((size_t)b * N + l) * 3 * d;
#pragma unroll
  for (int i = 0; i < HPT; ++i)
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd)
      if (dd < dk) stf(dqo + dd * H + HPT * hg + i, dq[i][dd] * P.dq_scale);
}

// backward, column pass: thread (m, HPT heads) owns dK[m] and dV[m]
template <typename T, typename G, bool PLAIN>
__global__ void __launch_bounds__(G::NT) attn_bwd_col_fast(AttnParams P) {
  constexpr int H = G::H, DKP = G::DKP, HPT = G::HPT, HG = G::HG;
  __shared__ __align__(16) float Qs[AKB * G::RS];
  __shared__ __align__(16) float Ds[AKB * G::RS];                // dV_att rows
  __shared__ __align__(16) float rst[AKB][H][4];                 // lse, log row sum, D, s
  const int tid = threadIdx.x, hg = tid % HG, r = tid / HG;
  const int b = blockIdx.y, N = P.N, dk = P.dk, d = H * dk;
  const int m = blockIdx.x * G::ROWS + r;
  const bool colvalid = m < N;
  const int mc = colvalid ? m : N - 1;
  const T *qkv = (const T *)P.qkv + (size_t)b * N * 3 * d;
  const T *dvatt = (const T *)P.d_v_att + (size_t)b * N * d;
  const size_t rs = (size_t)P.B * N * H;
  float k[HPT][DKP], v[HPT][DKP], dka[HPT][DKP], dva_[HPT][DKP];
#pragma unroll
  for (int i = 0; i < HPT; ++i)
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd) {
      const bool in = dd < dk;
      const T *kp = qkv + (size_t)mc * 3 * d + d + dd * H + HPT * hg + i;
      k[i][dd] = in ? ldf(kp) : 0.f;
      v[i][dd] = in ? ldf(kp + d) : 0.f;
      dka[i][dd] = 0.f; dva_[i][dd] = 0.f;
    }
  const float nk = P.mask ? ((float)P.mask[(size_t)b * N + mc] - 1.f) * kNegMask : 0.f;
  for (int l0 = 0; l0 < N; l0 += AKB) {
    __syncthreads();
    stage_pair<T, G>(Qs, Ds, qkv, (size_t)3 * d, dvatt, (size_t)d, l0, N, dk, tid);
    for (int idx = tid; idx < AKB * H; idx += G::NT) {
      const int lk = idx / H, hh = idx % H, l = l0 + lk;
      float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l < N) {
        const size_t ps = ((size_t)b * N + l) * H + hh;
        st = make_float4(P.lse[ps], P.lse[rs + ps], P.row_ws[ps], P.row_ws[rs + ps]);
      }
      *(float4 *)&rst[lk][hh][0] = st;
    }
    __syncthreads();
    const int lend = N - l0 < AKB ? N - l0 : AKB;
    for (int lk = 0; lk < lend; ++lk) {
      const size_t pair = ((size_t)b * N + l0 + lk) * N + mc, pe = pair * H + HPT * hg;
      float Ev[HPT], Gv[HPT], Mv[HPT], dhh[HPT];
#pragma unroll
      for (int i = 0; i < HPT; ++i) { Ev[i] = 0.f; Gv[i] = 0.f; Mv[i] = 1.f; dhh[i] = 0.f; }
      if (P.E) loadv<T, HPT>((const T *)P.E + pe, Ev);
      if (P.G) loadv<T, HPT>((const T *)P.G + pe, Gv);
      if constexpr (!PLAIN) load_masks<T, HPT>(P, pair, pe, Mv);
      if (P.d_h_hat) loadv<T, HPT>((const T *)P.d_h_hat + pe, dhh);
      const float *qg = Qs + lk * G::RS + hg * G::GS, *dg = Ds + lk * G::RS + hg * G::GS;
#pragma unroll
      for (int i = 0; i < HPT; ++i) {
        const float *qp = qg + i * DKP, *dp = dg + i * DKP;
        const float4 st = *(const float4 *)&rst[lk][HPT * hg + i][0];
        // q . k accumulates in the same order as in the row pass, so both passes see the same S_raw
        const Elem e = eval_elem<PLAIN>(P, dot_s<DKP>(k[i], qp), Ev[i], Gv[i], nk, Mv[i], b, l0 + lk, mc, HPT * hg + i);
        const float p = __expf((e.x - st.x) - st.y);
        const float g = P.G ? sigmoid_rcp(e.gin) : 1.f;
        const float dAp = dot_s<DKP>(v[i], dp);
        const float dA = st.w * dAp * e.keep;
        const float dH = p * (dA * g - st.z) + dhh[i];
        const bool inside = !P.has_clip || (e.S_raw >= P.clip_lo && e.S_raw <= P.clip_hi);
        axpy_s<DKP>(inside ? dH * P.scale : 0.f, qp, dka[i]);
        axpy_s<DKP>(p * g * e.keep * st.w, dp, dva_[i]);
      }
    }
  }
  if (!colvalid) return;
  T *o = (T *)P.d_qkv + ((size_t)b * N + m) * 3 * d + d;
#pragma unroll
  for (int i = 0; i < HPT; ++i)
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd)
      if (dd < dk) {
        stf(o + dd * H + HPT * hg + i, dka[i][dd]);
        stf(o + d + dd * H + HPT * hg + i, dva_[i][dd]);
      }
}

// column pass when the row pass left dS and s*A~ in the workspace: dK[m] = sum_l dS[l,m] Q[l], dV[m] = sum_l sA~[l,m] dV_att[l]
template <typename T, typename G>
__global__ void __launch_bounds__(G::NT) attn_bwd_col_pre(AttnParams P) {
  constexpr int H = G::H, DKP = G::DKP, HPT = G::HPT, HG = G::HG;
  __shared__ __align__(16) float Qs[AKB * G::RS];
  __shared__ __align__(16) float Ds[AKB * G::RS];                // dV_att rows
  const int tid = threadIdx.x, hg = tid % HG, r = tid / HG;
  const int b = blockIdx.y, N = P.N, dk = P.dk, d = H * dk;
  const int m = blockIdx.x * G::ROWS + r;
  const bool colvalid = m < N;
  const int mc = colvalid ? m : N - 1;
  const T *qkv = (const T *)P.qkv + (size_t)b * N * 3 * d;
  const T *dvatt = (const T *)P.d_v_att + (size_t)b * N * d;
  float dka[HPT][DKP], dva_[HPT][DKP];
#pragma unroll
  for (int i = 0; i < HPT; ++i)
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd) { dka[i][dd] = 0.f; dva_[i][dd] = 0.f; }
  for (int l0 = 0; l0 < N; l0 += AKB) {
    __syncthreads();
    stage_pair<T, G>(Qs, Ds, qkv, (size_t)3 * d, dvatt, (size_t)d, l0, N, dk, tid);
    __syncthreads();
    const int lend = N - l0 < AKB ? N - l0 : AKB;
#pragma unroll 2
    for (int lk = 0; lk < lend; ++lk) {
      const size_t pe = (((size_t)b * N + l0 + lk) * N + mc) * H + HPT * hg;
      float dSv[HPT], Asv[HPT];
      loadv<T, HPT>((const T *)P.dS_ws + pe, dSv);
      loadv<T, HPT>((const T *)P.As_ws + pe, Asv);
      const float *qg = Qs + lk * G::RS + hg * G::GS, *dg = Ds + lk * G::RS + hg * G::GS;
#pragma unroll
      for (int i = 0; i < HPT; ++i) {
        axpy_s<DKP>(dSv[i], qg + i * DKP, dka[i]);
        axpy_s<DKP>(Asv[i], dg + i * DKP, dva_[i]);
      }
    }
  }
  if (!colvalid) return;
  T *o = (T *)P.d_qkv + ((size_t)b * N + m) * 3 * d + d;
#pragma unroll
  for (int i = 0; i < HPT; ++i)
#pragma unroll
    for (int dd = 0; dd < DKP; ++dd)
      if (dd < dk) {
        stf(o + dd * H + HPT * hg + i, dka[i][dd]);
        stf(o + d + dd * H + HPT * hg + i, dva_[i][dd]);
      }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, typename GF, typename GB, bool PLAIN>
int launch3(int kind, const AttnParams &P, cudaStream_t st) {
  if (kind == 0) {
    const dim3 grid((unsigned)((P.N + GF::ROWS - 1) / GF::ROWS), (unsigned)P.B);
    LaunchScope _ls("attn_staged_fwd", st);
    attn_fwd_fast<T, GF, PLAIN><<<grid, GF::NT, 0, st>>>(P);
    EGT_CHECK_CUDA(cudaGetLastError());
  } else {
    const dim3 grid((unsigned)((P.N + GB::ROWS - 1) / GB::ROWS), (unsigned)P.B);
    { LaunchScope _ls("attn_staged_bwd_row", st); attn_bwd_row_fast<T, GB, PLAIN><<<grid, GB::NT, 0, st>>>(P); }
    EGT_CHECK_CUDA(cudaGetLastError());
    {
      LaunchScope _ls("attn_staged_bwd_col", st);
      if (P.dS_ws && P.As_ws) attn_bwd_col_pre<T, GB><<<grid, GB::NT, 0, st>>>(P);
      else attn_bwd_col_fast<T, GB, PLAIN><<<grid, GB::NT, 0, st>>>(P);
    }
    EGT_CHECK_CUDA(cudaGetLastError());
  }
  return EGT_OK;
}

// Heads per thread, measured on B200 (C5 / C3 / C1 widths): as few as the shared-memory budget allows -- more
// warps in flight beat wider accesses for this latency-bound arithmetic: one with 8 heads, two with 16.
// EGT_ATTN_HPT=<f><b> (e.g. 42: four in the forward, two in the backward) overrides for experiments.
template <typename T, int H, int DKP, bool PLAIN>
int launch_hpt(int kind, const AttnParams &P, cudaStream_t st) {
  int hf = H <= 8 ? 1 : 2, hb = H <= 8 ? 1 : 2;
  if (const char *e = getenv("EGT_ATTN_HPT")) { hf = e[0] - '0'; hb = e[1] - '0'; }
  const int hpt = kind == 0 ? hf : hb;
  if constexpr (H <= 8) {                              // (16 heads x 1 would exceed the static shared-memory limit)
    if (hpt == 1) return launch3<T, Geo<H, DKP, 1>, Geo<H, DKP, 1>, PLAIN>(kind, P, st);
  }
  if (hpt == 2) return launch3<T, Geo<H, DKP, 2>, Geo<H, DKP, 2>, PLAIN>(kind, P, st);
  return launch3<T, Geo<H, DKP, 4>, Geo<H, DKP, 4>, PLAIN>(kind, P, st);
}

template <typename T, bool PLAIN>
int launch_shape(int kind, const AttnParams &P, cudaStream_t st) {
  if (P.h == 8 && P.dk <= 8) return launch_hpt<T, 8, 8, PLAIN>(kind, P, st);
  if (P.h == 8 && P.dk <= 12) return launch_hpt<T, 8, 12, PLAIN>(kind, P, st);
  if (P.h == 16 && P.dk <= 8) return launch_hpt<T, 16, 8, PLAIN>(kind, P, st);
  return 1;
}

}  // namespace

// kind: 0 forward, 1 backward (row + column pass).  Returns 1 when the shape / alignment / requested outputs
// have no specialised kernel (the caller then runs attn_staged.cu), EGT_OK or a negative status otherwise.
int attn_fast_launch(int kind, const AttnParams &P, int dtype, cudaStream_t st) {
  if (P.a_tild || P.N < 1 || P.B > 65535) return 1;
  const void *ptrs[] = {P.E, P.G, P.attn_mask == EGT_MASK_DENSE ? P.M : nullptr, P.h_hat, P.d_h_hat, P.dE, P.dG, P.dS_ws, P.As_ws};
  for (const void *q : ptrs)
    if (!aligned16(q)) return 1;
  const bool plain = P.attn_mask == EGT_MASK_NONE && !P.rand_mask && !P.dropout;
  if (dtype == EGT_F32) return plain ? launch_shape<float, true>(kind, P, st) : launch_shape<float, false>(kind, P, st);
  return plain ? launch_shape<__nv_bfloat16, true>(kind, P, st) : launch_shape<__nv_bfloat16, false>(kind, P, st);
}

}  // namespace egt
