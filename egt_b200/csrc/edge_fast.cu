// edge_fast.cu -- shape-specialised staged edge-channel kernels (graph_xformer_model_base.py:149-223).
//
// Same four operations as edge_kernels.cu (which stays as the any-shape fallback), compiled per (d_e, h):
//   * one thread per (b,l,m) pair, the pair's d_e-row moves with 16-byte loads / stores and lives in registers;
//   * the projection weights sit in shared memory and are read as broadcast float4;
//   * CTAs are persistent (grid = a multiple of the SM count), so the weights are staged once per CTA;
//   * the weight-gradient sums  X^T Y  over the pairs of a chunk are 4x4 register tiles fed by float4 shared
//     loads (2 loads per 16 FMA); a CTA keeps them in registers across all of its chunks and adds them to
//     global memory once.
// These kernels serve every configuration the fused tcgen05 path (fused_fwd.cu / fused_bwd.cu) does not take:
// fp32 activations, d_e or h other than 8, the ablation variants.
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace egt {
namespace {
using umma::bf16_hi;
using umma::bf16_lo;
using umma::pack_bf16;

constexpr int FPB = 128;   // pairs per chunk == threads per CTA

template <typename T, int W>
__device__ __forceinline__ void load_row(const T *p, float *x, bool live) {
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < W / 8; ++i) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (live) v = __ldg((const uint4 *)p + i);
      x[8 * i + 0] = bf16_lo(v.x); x[8 * i + 1] = bf16_hi(v.x); x[8 * i + 2] = bf16_lo(v.y); x[8 * i + 3] = bf16_hi(v.y);
      x[8 * i + 4] = bf16_lo(v.z); x[8 * i + 5] = bf16_hi(v.z); x[8 * i + 6] = bf16_lo(v.w); x[8 * i + 7] = bf16_hi(v.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) v = __ldg((const float4 *)p + i);
      x[4 * i + 0] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
  }
}

template <typename T, int W>
__device__ __forceinline__ void store_row(T *p, const float *x) {
  if constexpr (sizeof(T) == 2) {
#pragma unroll
    for (int i = 0; i < W / 8; ++i) {
      uint4 v;
      v.x = pack_bf16(x[8 * i + 0], x[8 * i + 1]); v.y = pack_bf16(x[8 * i + 2], x[8 * i + 3]);
      v.z = pack_bf16(x[8 * i + 4], x[8 * i + 5]); v.w = pack_bf16(x[8 * i + 6], x[8 * i + 7]);
      ((uint4 *)p)[i] = v;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) ((float4 *)p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
  }
}

// rows of a [FPB][W] float staging buffer are W+4 floats apart: the 16-byte row accesses of 8 neighbouring
// threads then fall into 8 different bank groups
template <int W> __device__ __forceinline__ float *srow(float *base, int r) { return base + r * (W + 4); }

template <int W>
__device__ __forceinline__ void stage_row(float *base, int r, const float *x) {
  float4 *d = (float4 *)srow<W>(base, r);
#pragma unroll
  for (int i = 0; i < W / 4; ++i) d[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
}

template <int DE>
__device__ __forceinline__ void ln_stats_reg(const float *x, float eps, float &mu, float &rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < DE; ++c) s += x[c];
  mu = s / DE;
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < DE; ++c) { const float t = x[c] - mu; v += t * t; }
  rstd = rsqrtf(v / DE + eps);
}

// acc[4][4] += sum over the rows q = q0, q0+qs, ... < FPB of  X[q][c0..c0+3] (x) Y[q][j0..j0+3]
template <int WX, int WY>
__device__ __forceinline__ void tile_xty(const float *xs, const float *ys, int c0, int j0, int q0, int qs, float (&acc)[16]) {
  for (int q = q0; q < FPB; q += qs) {
    const float4 xv = *(const float4 *)(xs + q * (WX + 4) + c0);
    const float4 yv = *(const float4 *)(ys + q * (WY + 4) + j0);
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[4 * i + j] = fmaf(xa[i], ya[j], acc[4 * i + j]);
  }
}

// ---------------------------------------------------------------------------------------------------
// e -> LN_e -> G = e^ W_G + b_G ; E = act(e^ W_E + b_E)        (:194-208, :149-162; 'bias' variant: no LN)
template <typename T, int DE, int H>
__global__ void __launch_bounds__(FPB) edge_proj_fwd_fast(EdgeParams p) {
  __shared__ __align__(16) float w[DE][2 * H];      // [c][E heads | G heads]
  __shared__ float gam[DE], bet[DE], bias[2 * H];
  const int tid = threadIdx.x;
  for (int i = tid; i < DE * 2 * H; i += FPB) {
    const int c = i / (2 * H), j = i % (2 * H);
    w[c][j] = j < H ? p.w_e[c * H + j] : (p.gated ? p.w_g[c * H + j - H] : 0.f);
  }
  for (int i = tid; i < DE; i += FPB) { gam[i] = p.has_ln ? p.ln_g[i] : 1.f; bet[i] = p.has_ln ? p.ln_b[i] : 0.f; }
  if (tid < 2 * H) bias[tid] = tid < H ? p.b_e[tid] : (p.gated ? p.b_g[tid - H] : 0.f);
  __syncthreads();
  for (size_t chunk = blockIdx.x; chunk * FPB < p.pairs; chunk += gridDim.x) {
    const size_t pair = chunk * FPB + tid;
    if (pair >= p.pairs) continue;
    float x[DE];
    load_row<T, DE>((const T *)p.e + pair * DE, x, true);
    float mu = 0.f, rstd = 1.f;
    if (p.has_ln) ln_stats_reg<DE>(x, p.ln_eps, mu, rstd);
    float acc[2 * H];
#pragma unroll
    for (int j = 0; j < 2 * H; ++j) acc[j] = bias[j];
#pragma unroll
    for (int c = 0; c < DE; ++c) {
      const float xv = fmaf((x[c] - mu) * rstd, gam[c], bet[c]);
#pragma unroll
      for (int j4 = 0; j4 < (2 * H) / 4; ++j4) {
        if (j4 >= H / 4 && !p.gated) break;
        const float4 wv = *(const float4 *)&w[c][4 * j4];
        acc[4 * j4 + 0] = fmaf(xv, wv.x, acc[4 * j4 + 0]); acc[4 * j4 + 1] = fmaf(xv, wv.y, acc[4 * j4 + 1]);
        acc[4 * j4 + 2] = fmaf(xv, wv.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(xv, wv.w, acc[4 * j4 + 3]);
      }
    }
    if (p.act != EGT_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = edge_act_fwd(p.act, p.act_alpha, acc[j]);
    }
    store_row<T, H>((T *)p.E + pair * H, acc);
    if (p.gated) store_row<T, H>((T *)p.G + pair * H, acc + H);
  }
}

// e' = H_hat W_r + b_r + e                                                                    (:214-218)
template <typename T, int DE, int H>
__global__ void __launch_bounds__(FPB) edge_out_fwd_fast(EdgeParams p) {
  __shared__ __align__(16) float w[H][DE];
  __shared__ __align__(16) float br[DE];
  const int tid = threadIdx.x;
  for (int i = tid; i < H * DE; i += FPB) w[i / DE][i % DE] = p.w_r[i];
  for (int i = tid; i < DE; i += FPB) br[i] = p.b_r[i];
  __syncthreads();
  for (size_t chunk = blockIdx.x; chunk * FPB < p.pairs; chunk += gridDim.x) {
    const size_t pair = chunk * FPB + tid;
    if (pair >= p.pairs) continue;
    float hv[H];
    load_row<T, H>((const T *)p.h_hat + pair * H, hv, true);
    const T *ep = (const T *)p.e + pair * DE;
    T *eo = (T *)p.e_out + pair * DE;
#pragma unroll
    for (int c8 = 0; c8 < DE / 8; ++c8) {           // 8 channels at a time keeps the register count flat in d_e
      float o[8];
      load_row<T, 8>(ep + 8 * c8, o, true);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] += br[8 * c8 + q];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        const float4 w0 = *(const float4 *)&w[hh][8 * c8], w1 = *(const float4 *)&w[hh][8 * c8 + 4];
        o[0] = fmaf(hv[hh], w0.x, o[0]); o[1] = fmaf(hv[hh], w0.y, o[1]); o[2] = fmaf(hv[hh], w0.z, o[2]); o[3] = fmaf(hv[hh], w0.w, o[3]);
        o[4] = fmaf(hv[hh], w1.x, o[4]); o[5] = fmaf(hv[hh], w1.y, o[5]); o[6] = fmaf(hv[hh], w1.z, o[6]); o[7] = fmaf(hv[hh], w1.w, o[7]);
      }
      store_row<T, 8>(eo + 8 * c8, o);
    }
  }
}

// dH_ext = de' W_r^T (when d_h_ext) ; dW_r += H_hat^T de' , db_r += colsum(de') (when g_w_r)
template <typename T, int DE, int H>
__global__ void __launch_bounds__(FPB) edge_out_bwd_fast(EdgeParams p) {
  extern __shared__ __align__(16) float sm[];
  float *wt = sm;                                   // [H][DE]
  float *xs = wt + H * DE;                          // [FPB][H + 4]   H_hat
  float *ys = xs + FPB * (H + 4);                   // [FPB][DE + 4]  de'
  float *red = ys + FPB * (DE + 4);                 // [H * DE + DE]
  const int tid = threadIdx.x;
  for (int i = tid; i < H * DE; i += FPB) wt[i] = p.w_r[i];
  for (int i = tid; i < H * DE + DE; i += FPB) red[i] = 0.f;
  constexpr int TILES = (H / 4) * (DE / 4);         // 4x4 output tiles of dW_r
  constexpr int GROUPS = FPB / TILES > 0 ? FPB / TILES : 1;
  static_assert(TILES <= FPB, "dW_r tile count exceeds the CTA");
  const int tile = tid % TILES, grp = tid / TILES;
  const int t_h0 = 4 * (tile / (DE / 4)), t_c0 = 4 * (tile % (DE / 4));
  float acc[16], accb = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  __syncthreads();
  for (size_t chunk = blockIdx.x; chunk * FPB < p.pairs; chunk += gridDim.x) {
    const size_t pair = chunk * FPB + tid;
    const bool live = pair < p.pairs;
    float g[DE];
    load_row<T, DE>((const T *)p.de_out + (live ? pair : 0) * DE, g, live);
    if (p.d_h_ext && live) {
      float dh[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < DE / 4; ++c4) {
          const float4 wv = *(const float4 *)(wt + hh * DE + 4 * c4);
          a0 = fmaf(g[4 * c4 + 0], wv.x, a0); a1 = fmaf(g[4 * c4 + 1], wv.y, a1);
          a0 = fmaf(g[4 * c4 + 2], wv.z, a0); a1 = fmaf(g[4 * c4 + 3], wv.w, a1);
        }
        dh[hh] = a0 + a1;
      }
      store_row<T, H>((T *)p.d_h_ext + pair * H, dh);
    }
    if (!p.g_w_r) continue;                         // uniform over the grid
    stage_row<DE>(ys, tid, g);
    {
      float hv[H];
      load_row<T, H>((const T *)p.h_hat + (live ? pair : 0) * H, hv, live && p.h_hat != nullptr);
      stage_row<H>(xs, tid, hv);
    }
    __syncthreads();
    if (grp < GROUPS) tile_xty<H, DE>(xs, ys, t_h0, t_c0, grp, GROUPS, acc);
    if (tid < DE) {
      float a = 0.f;
      for (int q = 0; q < FPB; ++q) a += ys[q * (DE + 4) + tid];
      accb += a;
    }
    __syncthreads();
  }
  if (!p.g_w_r) return;
  if (grp < GROUPS) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(red + (t_h0 + i) * DE + t_c0 + j, acc[4 * i + j]);
  }
  if (tid < DE) red[H * DE + tid] = accb;
  __syncthreads();
  for (int i = tid; i < H * DE; i += FPB) atomicAdd(p.g_w_r + i, red[i]);
  for (int i = tid; i < DE; i += FPB) atomicAdd(p.g_b_r + i, red[H * DE + i]);
}

// dE, dG -> d e^ -> LN backward -> de (+ de') ; dW_E, dW_G, db_E, db_G, dgamma_e, dbeta_e
template <typename T, int DE, int H>
__global__ void __launch_bounds__(FPB) edge_proj_bwd_fast(EdgeParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int J = 2 * H;
  float *w = sm;                                    // [DE][J]   (E heads | G heads)
  float *xs = w + DE * J;                           // [FPB][DE + 4]  x^ without the affine (or raw e)
  float *ds = xs + FPB * (DE + 4);                  // [FPB][J + 4]   dE | dG
  float *red = ds + FPB * (J + 4);                  // [DE * J + J]
  float *gam = red + DE * J + J, *bet = gam + DE, *bias = bet + DE;   // [DE], [DE], [H]
  const int tid = threadIdx.x;
  for (int i = tid; i < DE * J; i += FPB) {
    const int c = i / J, j = i % J;
    w[i] = j < H ? p.w_e[c * H + j] : (p.gated ? p.w_g[c * H + j - H] : 0.f);
  }
  for (int i = tid; i < DE * J + J; i += FPB) red[i] = 0.f;
  for (int i = tid; i < DE; i += FPB) { gam[i] = p.has_ln ? p.ln_g[i] : 1.f; bet[i] = p.has_ln ? p.ln_b[i] : 0.f; }
  if (tid < H) bias[tid] = p.b_e[tid];
  constexpr int TILES = (DE / 4) * (J / 4);
  static_assert(TILES <= FPB, "dW tile count exceeds the CTA");
  constexpr int GROUPS = FPB / TILES;
  const int tile = tid % TILES, grp = tid / TILES;
  const int t_c0 = 4 * (tile / (J / 4)), t_j0 = 4 * (tile % (J / 4));
  float acc[16], accb = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  __syncthreads();
  for (size_t chunk = blockIdx.x; chunk * FPB < p.pairs; chunk += gridDim.x) {
    const size_t pair = chunk * FPB + tid;
    const bool live = pair < p.pairs;
    const size_t pr = live ? pair : 0;
    float x[DE];
    load_row<T, DE>((const T *)p.e + pr * DE, x, live);
    float mu = 0.f, rstd = 1.f;
    if (p.has_ln) ln_stats_reg<DE>(x, p.ln_eps, mu, rstd);
#pragma unroll
    for (int c = 0; c < DE; ++c) x[c] = (x[c] - mu) * rstd;      // x^ (raw e when there is no LN)
    stage_row<DE>(xs, tid, x);
    float dz[J];
    load_row<T, H>((const T *)p.dE + pr * H, dz, live);
    if (p.gated) load_row<T, H>((const T *)p.dG + pr * H, dz + H, live);
    else {
#pragma unroll
      for (int j = H; j < J; ++j) dz[j] = 0.f;
    }
    if (p.act != EGT_ACT_NONE) {                    // derivative of the edge activation at the recomputed pre-activation
      float aE[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) aE[hh] = bias[hh];
#pragma unroll
      for (int c = 0; c < DE; ++c) {
        const float xv = fmaf(x[c], gam[c], bet[c]);
#pragma unroll
        for (int j4 = 0; j4 < H / 4; ++j4) {
          const float4 wv = *(const float4 *)(w + c * J + 4 * j4);
          aE[4 * j4 + 0] = fmaf(xv, wv.x, aE[4 * j4 + 0]); aE[4 * j4 + 1] = fmaf(xv, wv.y, aE[4 * j4 + 1]);
          aE[4 * j4 + 2] = fmaf(xv, wv.z, aE[4 * j4 + 2]); aE[4 * j4 + 3] = fmaf(xv, wv.w, aE[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < H; ++hh) dz[hh] *= edge_act_bwd(p.act, p.act_alpha, aE[hh]);
    }
    stage_row<J>(ds, tid, dz);
    if (live) {
      // d x^[c] = gamma[c] * sum_j dz[j] W[c][j]   (kept in place of nothing: x^ is re-read from shared memory below)
      float dx[DE];
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int c = 0; c < DE; ++c) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < J / 4; ++j4) {
          if (j4 >= H / 4 && !p.gated) break;
          const float4 wv = *(const float4 *)(w + c * J + 4 * j4);
          a0 = fmaf(dz[4 * j4 + 0], wv.x, a0); a1 = fmaf(dz[4 * j4 + 1], wv.y, a1);
          a0 = fmaf(dz[4 * j4 + 2], wv.z, a0); a1 = fmaf(dz[4 * j4 + 3], wv.w, a1);
        }
        dx[c] = (a0 + a1) * gam[c];
        m1 += dx[c];
        m2 = fmaf(dx[c], x[c], m2);
      }
      if (p.has_ln) {
        m1 /= DE; m2 /= DE;
#pragma unroll
        for (int c = 0; c < DE; ++c) dx[c] = rstd * (dx[c] - m1 - x[c] * m2);
      }
      if (p.de_out) {
        const T *dp = (const T *)p.de_out + pair * DE;
#pragma unroll
        for (int c8 = 0; c8 < DE / 8; ++c8) {
          float r8[8];
          load_row<T, 8>(dp + 8 * c8, r8, true);
#pragma unroll
          for (int q = 0; q < 8; ++q) dx[8 * c8 + q] += r8[q];
        }
      }
      store_row<T, DE>((T *)p.de + pair * DE, dx);
    }
    __syncthreads();
    if (grp < GROUPS) tile_xty<DE, J>(xs, ds, t_c0, t_j0, grp, GROUPS, acc);
    if (tid < J) {
      float a = 0.f;
      for (int q = 0; q < FPB; ++q) a += ds[q * (J + 4) + tid];
      accb += a;
    }
    __syncthreads();
  }
  if (grp < GROUPS) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(red + (t_c0 + i) * J + t_j0 + j, acc[4 * i + j]);
  }
  if (tid < J) red[DE * J + tid] = accb;
  __syncthreads();
  // dW = gamma*dWn + beta*db ; dgamma[c] += sum_j W[c,j] dWn[c,j] ; dbeta[c] += sum_j W[c,j] db[j]
  const int Jeff = p.gated ? J : H;
  for (int j = tid; j < Jeff; j += FPB) atomicAdd(j < H ? p.g_b_e + j : p.g_b_g + (j - H), red[DE * J + j]);
  for (int o = tid; o < DE * J; o += FPB) {
    const int c = o / J, j = o % J;
    if (j >= Jeff) continue;
    float *gW = j < H ? p.g_w_e : p.g_w_g;
    const int jj = j < H ? j : j - H;
    if (p.has_ln) atomicAdd(gW + c * H + jj, gam[c] * red[o] + bet[c] * red[DE * J + j]);
    else atomicAdd(gW + c * H + jj, red[o]);
  }
  if (p.has_ln) {
    for (int c = tid; c < DE; c += FPB) {
      float dg = 0.f, db = 0.f;
      for (int j = 0; j < Jeff; ++j) { dg = fmaf(w[c * J + j], red[c * J + j], dg); db = fmaf(w[c * J + j], red[DE * J + j], db); }
      atomicAdd(p.g_ln_g + c, dg);
      atomicAdd(p.g_ln_b + c, db);
    }
  }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// persistent grid: as many CTAs as are resident at once (occupancy x SM count), never more than there are chunks
template <typename K>
unsigned resident_ctas(K kernel, size_t smem) {
  int per_sm = 1, sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FPB, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (unsigned)(sms * per_sm);
}
inline unsigned fast_grid(const EdgeParams &p, unsigned cap) {
  const size_t chunks = (p.pairs + FPB - 1) / FPB;
  return (unsigned)(chunks < cap ? chunks : cap);
}

template <typename T, int DE, int H>
int launch_kind(int kind, const EdgeParams &p, cudaStream_t st) {
  switch (kind) {
    case 0: {
      LaunchScope _ls("edge_proj_fwd_kernel", st);
      static const unsigned cap = resident_ctas(edge_proj_fwd_fast<T, DE, H>, 0);
      edge_proj_fwd_fast<T, DE, H><<<fast_grid(p, cap), FPB, 0, st>>>(p);
      break;
    }
    case 1: {
      LaunchScope _ls("edge_out_fwd_kernel", st);
      static const unsigned cap = resident_ctas(edge_out_fwd_fast<T, DE, H>, 0);
      edge_out_fwd_fast<T, DE, H><<<fast_grid(p, cap), FPB, 0, st>>>(p);
      break;
    }
    case 2: {
      const size_t smem = sizeof(float) * (H * DE + FPB * (H + 4) + FPB * (DE + 4) + H * DE + DE);
      static bool attr = false;
      if (!attr) {
        EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_out_bwd_fast<T, DE, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
      }
      LaunchScope _ls("edge_out_bwd_kernel", st);
      static const unsigned cap = resident_ctas(edge_out_bwd_fast<T, DE, H>, smem);
      edge_out_bwd_fast<T, DE, H><<<fast_grid(p, cap), FPB, smem, st>>>(p);
      break;
    }
    default: {
      constexpr int J = 2 * H;
      const size_t smem = sizeof(float) * (DE * J + FPB * (DE + 4) + FPB * (J + 4) + DE * J + J + 2 * DE + H);
      static bool attr = false;
      if (!attr) {
        EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_proj_bwd_fast<T, DE, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
      }
      LaunchScope _ls("edge_proj_bwd_kernel", st);
      static const unsigned cap = resident_ctas(edge_proj_bwd_fast<T, DE, H>, smem);
      edge_proj_bwd_fast<T, DE, H><<<fast_grid(p, cap), FPB, smem, st>>>(p);
      break;
    }
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

template <typename T>
int launch_shape(int kind, const EdgeParams &p, cudaStream_t st) {
#define EGT_SHAPE(DE_, H_) if (p.d_e == DE_ && p.h == H_) return launch_kind<T, DE_, H_>(kind, p, st);
  EGT_SHAPE(8, 8) EGT_SHAPE(16, 8) EGT_SHAPE(32, 8) EGT_SHAPE(48, 8) EGT_SHAPE(64, 8)
  EGT_SHAPE(8, 16) EGT_SHAPE(16, 16) EGT_SHAPE(32, 16) EGT_SHAPE(64, 16)
#undef EGT_SHAPE
  return 1;
}

}  // namespace

// kind: 0 proj_fwd, 1 out_fwd, 2 out_bwd, 3 proj_bwd.  Returns 1 when this (d_e, h) / alignment has no
// specialised kernel (the caller then runs the any-shape kernel), EGT_OK or a negative status otherwise.
int edge_fast_launch(int kind, const EdgeParams &p, int dtype, cudaStream_t st) {
  const void *ptrs[] = {p.e, p.E, p.G, p.h_hat, p.e_out, p.de_out, p.d_h_ext, p.dE, p.dG, p.de};
  for (const void *q : ptrs)
    if (!aligned16(q)) return 1;
  return dtype == EGT_F32 ? launch_shape<float>(kind, p, st) : launch_shape<__nv_bfloat16>(kind, p, st);
}

}  // namespace egt
