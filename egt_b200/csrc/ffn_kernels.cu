// ffn_kernels.cu -- the per-channel feed-forward half of an EGT layer ("next" row 8f-1 of SURVEY.md):
//
//   reference: ffnlr1 / ffnact / ffnlr2 and ffn_block (lib/models/graph_xformer_model_base.py:229-258,
//   :309-324):   y = x + Dense_w( act( Dense_round(w*ffn_multiplier)( LayerNorm(x) ) ) )
//   applied to the node channel (w = model_width, rows = B*N) and, for residual / constrained edge
//   channels, to the edge channel (w = edge_width, rows = B*N*N).  Dropout is 0 in every shipped config.
//
// CUDA-core kernels for any width (weights in shared memory when they fit, else read through L1/L2):
// a CTA walks chunks of RB rows with a grid stride; the backward keeps its weight-gradient sums in shared
// memory across chunks and issues its atomics once at the end.  The edge FFN is point-wise in (l, m) and is
// HBM-bound (read x, write y); fusing it behind the edge write-back of the attention kernel is the next step.
#include <string.h>
#include "common.cuh"
#include "kernels.h"
#include "node_blas.h"

namespace egt {

namespace {

constexpr int FT = 512;   // threads per CTA (one CTA per SM for wide layers: many warps hide the shared-memory latency)

__device__ __forceinline__ float act_fwd(int act, float x) { return edge_act_fwd(act, 0.2f, x); }
__device__ __forceinline__ float act_bwd(int act, float x) { return edge_act_bwd(act, 0.2f, x); }

struct FfnArgs {
  long long rows; int w, hid, act; float eps;
  const float *gamma, *beta, *W1, *b1, *W2, *b2;
  const void *x, *dy; void *y, *dx;
  float *g_gamma, *g_beta, *g_W1, *g_b1, *g_W2, *g_b2;
  int rb;            // rows per chunk
  int w_in_smem;     // W1 | W2 staged in shared memory
  int acc_in_smem;   // backward: weight-gradient sums kept in shared memory across chunks (else atomics per chunk)
};

// stage weights: returns pointers (shared or global) and the float offset of the free area
__device__ __forceinline__ int stage_weights(const FfnArgs &a, float *sm, const float *&W1, const float *&W2,
                                             const float *&b1, const float *&b2, const float *&gam, const float *&bet) {
  int off = 0;
  const int tid = threadIdx.x;
  float *sb1 = sm + off; off += a.hid;
  float *sb2 = sm + off; off += a.w;
  float *sg = sm + off; off += a.w;
  float *sb = sm + off; off += a.w;
  for (int i = tid; i < a.hid; i += FT) sb1[i] = a.b1[i];
  for (int i = tid; i < a.w; i += FT) { sb2[i] = a.b2[i]; sg[i] = a.gamma[i]; sb[i] = a.beta[i]; }
  b1 = sb1; b2 = sb2; gam = sg; bet = sb;
  if (a.w_in_smem) {
    float *s1 = sm + off; off += a.w * a.hid;
    float *s2 = sm + off; off += a.w * a.hid;
    if (((a.w * a.hid) & 3) == 0 && (((uintptr_t)a.W1 | (uintptr_t)a.W2) & 15) == 0 && (off & 3) == 0) {
      for (int i = tid; i < (a.w * a.hid) >> 2; i += FT) {   // 16-byte copies, independent loads in flight
        ((float4 *)s1)[i] = ((const float4 *)a.W1)[i];
        ((float4 *)s2)[i] = ((const float4 *)a.W2)[i];
      }
    } else {
      for (int i = tid; i < a.w * a.hid; i += FT) { s1[i] = a.W1[i]; s2[i] = a.W2[i]; }
    }
    W1 = s1; W2 = s2;
  } else {
    W1 = a.W1; W2 = a.W2;
  }
  return off;
}

// rows [r0, r0+nr) -> xs (raw x), xn (normalised, no affine), xe (gamma*xn + beta); one warp per row
template <typename T>
__device__ __forceinline__ void load_norm_rows(const FfnArgs &a, long long r0, int nr, float *xs, float *xn, float *xe,
                                               float *rstd_s, const float *gam, const float *bet) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int rr = warp; rr < a.rb; rr += FT / 32) {
    if (rr >= nr) {
      for (int c = lane; c < a.w; c += 32) { xs[rr * a.w + c] = 0.f; xn[rr * a.w + c] = 0.f; xe[rr * a.w + c] = 0.f; }
      if (lane == 0) rstd_s[rr] = 0.f;
      continue;
    }
    const T *xp = (const T *)a.x + (size_t)(r0 + rr) * a.w;
    float s = 0.f;
    for (int c = lane; c < a.w; c += 32) { const float v = ldf(xp + c); xs[rr * a.w + c] = v; s += v; }
    s = warp_sum(s);
    const float mu = s / a.w;
    float var = 0.f;
    for (int c = lane; c < a.w; c += 32) { const float d = xs[rr * a.w + c] - mu; var += d * d; }
    var = warp_sum(var);
    const float rstd = rsqrtf(var / a.w + a.eps);
    if (lane == 0) rstd_s[rr] = rstd;
    for (int c = lane; c < a.w; c += 32) {
      const float n = (xs[rr * a.w + c] - mu) * rstd;
      xn[rr * a.w + c] = n;
      xe[rr * a.w + c] = fmaf(n, gam[c], bet[c]);
    }
  }
}

}  // namespace

template <typename T>
__global__ void __launch_bounds__(FT) ffn_fwd_kernel(FfnArgs a) {
  extern __shared__ float sm[];
  const float *W1, *W2, *b1, *b2, *gam, *bet;
  int off = stage_weights(a, sm, W1, W2, b1, b2, gam, bet);
  float *xs = sm + off; off += a.rb * a.w;
  float *xn = sm + off; off += a.rb * a.w;
  float *xe = sm + off; off += a.rb * a.w;
  float *hs = sm + off; off += a.rb * a.hid;
  float *rstd_s = sm + off;
  const int tid = threadIdx.x;
  __syncthreads();
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long r0 = ch * a.rb;
    const int nr = (int)min((long long)a.rb, a.rows - r0);
    load_norm_rows<T>(a, r0, nr, xs, xn, xe, rstd_s, gam, bet);
    __syncthreads();
    for (int o = tid; o < a.rb * a.hid; o += FT) {           // hidden = act(e^ W1 + b1)
      const int rr = o / a.hid, j = o % a.hid;
      float acc = b1[j];
      for (int c = 0; c < a.w; ++c) acc = fmaf(xe[rr * a.w + c], W1[c * a.hid + j], acc);
      hs[o] = act_fwd(a.act, acc);
    }
    __syncthreads();
    for (int o = tid; o < nr * a.w; o += FT) {               // y = x + hidden W2 + b2
      const int rr = o / a.w, c = o % a.w;
      float acc = b2[c];
      for (int j = 0; j < a.hid; ++j) acc = fmaf(hs[rr * a.hid + j], W2[j * a.w + c], acc);
      stf((T *)a.y + (size_t)(r0 + rr) * a.w + c, acc + xs[o]);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(FT) ffn_bwd_kernel(FfnArgs a) {
  extern __shared__ float sm[];
  const float *W1, *W2, *b1, *b2, *gam, *bet;
  int off = stage_weights(a, sm, W1, W2, b1, b2, gam, bet);
  float *xs = sm + off; off += a.rb * a.w;      // raw x, later dy
  float *xn = sm + off; off += a.rb * a.w;
  float *xe = sm + off; off += a.rb * a.w;      // e^, later d e^
  float *hs = sm + off; off += a.rb * a.hid;    // act(pre)
  float *ds = sm + off; off += a.rb * a.hid;    // pre-activation, later d pre
  float *rstd_s = sm + off; off += a.rb;
  float *aW1 = a.acc_in_smem ? sm + off : a.g_W1; off += a.acc_in_smem ? a.w * a.hid : 0;    // weight-gradient sums of this CTA
  float *aW2 = a.acc_in_smem ? sm + off : a.g_W2; off += a.acc_in_smem ? a.w * a.hid : 0;    // (wide layers: global atomics per chunk)
  float *ab1 = sm + off; off += a.hid;
  float *ab2 = sm + off; off += a.w;
  float *ag = sm + off; off += a.w;
  float *ab = sm + off; off += a.w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.acc_in_smem) { for (int i = tid; i < 2 * a.w * a.hid; i += FT) aW1[i] = 0.f; }   // aW1, aW2 are contiguous
  for (int i = tid; i < a.hid + 3 * a.w; i += FT) ab1[i] = 0.f;                        // ab1 .. ab are contiguous
  __syncthreads();
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long r0 = ch * a.rb;
    const int nr = (int)min((long long)a.rb, a.rows - r0);
    load_norm_rows<T>(a, r0, nr, xs, xn, xe, rstd_s, gam, bet);
    __syncthreads();
    for (int o = tid; o < a.rb * a.hid; o += FT) {           // recompute pre-activation and hidden
      const int rr = o / a.hid, j = o % a.hid;
      float acc = b1[j];
      for (int c = 0; c < a.w; ++c) acc = fmaf(xe[rr * a.w + c], W1[c * a.hid + j], acc);
      ds[o] = acc;
      hs[o] = act_fwd(a.act, acc);
    }
    for (int o = tid; o < a.rb * a.w; o += FT) {             // xs <- dy (rows beyond nr: 0)
      const int rr = o / a.w, c = o % a.w;
      xs[o] = rr < nr ? ldf((const T *)a.dy + (size_t)(r0 + rr) * a.w + c) : 0.f;
    }
    __syncthreads();
    // weight gradients that need hs / xe before they are overwritten: dW2 += hs^T dy ; db2 += colsum(dy)
    for (int o = tid; o < a.hid * a.w; o += FT) {
      const int j = o / a.w, c = o % a.w;
      float acc = 0.f;
      for (int rr = 0; rr < a.rb; ++rr) acc = fmaf(hs[rr * a.hid + j], xs[rr * a.w + c], acc);
      if (a.acc_in_smem) aW2[o] += acc; else atomicAdd(aW2 + o, acc);
    }
    for (int c = tid; c < a.w; c += FT) {
      float acc = 0.f;
      for (int rr = 0; rr < a.rb; ++rr) acc += xs[rr * a.w + c];
      ab2[c] += acc;
    }
    __syncthreads();
    for (int o = tid; o < a.rb * a.hid; o += FT) {           // d pre = (dy W2^T) act'(pre)
      const int rr = o / a.hid, j = o % a.hid;
      float acc = 0.f;
      int c = lane % a.w;                                      // rotated start: W2 rows are a.w floats apart, so
      for (int k = 0; k < a.w; ++k) {                          // lanes with consecutive j would hit one bank
        acc = fmaf(xs[rr * a.w + c], W2[j * a.w + c], acc);
        if (++c == a.w) c = 0;
      }
      ds[o] = rr < nr ? acc * act_bwd(a.act, ds[o]) : 0.f;
    }
    __syncthreads();
    // dW1 += e^^T dpre ; db1 += colsum(dpre)
    for (int o = tid; o < a.w * a.hid; o += FT) {
      const int c = o / a.hid, j = o % a.hid;
      float acc = 0.f;
      for (int rr = 0; rr < a.rb; ++rr) acc = fmaf(xe[rr * a.w + c], ds[rr * a.hid + j], acc);
      if (a.acc_in_smem) aW1[o] += acc; else atomicAdd(aW1 + o, acc);
    }
    for (int j = tid; j < a.hid; j += FT) {
      float acc = 0.f;
      for (int rr = 0; rr < a.rb; ++rr) acc += ds[rr * a.hid + j];
      ab1[j] += acc;
    }
    __syncthreads();
    for (int o = tid; o < a.rb * a.w; o += FT) {             // xe <- d e^ = dpre W1^T
      const int rr = o / a.w, c = o % a.w;
      float acc = 0.f;
      int j = lane % a.hid;                                    // rotated start (W1 rows are a.hid floats apart)
      for (int k = 0; k < a.hid; ++k) {
        acc = fmaf(ds[rr * a.hid + j], W1[c * a.hid + j], acc);
        if (++j == a.hid) j = 0;
      }
      xe[o] = acc;
    }
    __syncthreads();
    for (int c = tid; c < a.w; c += FT) {                    // dgamma += colsum(d e^ * xn) ; dbeta += colsum(d e^)
      float sg = 0.f, sb = 0.f;
      for (int rr = 0; rr < a.rb; ++rr) { sg = fmaf(xe[rr * a.w + c], xn[rr * a.w + c], sg); sb += xe[rr * a.w + c]; }
      ag[c] += sg; ab[c] += sb;
    }
    for (int rr = warp; rr < nr; rr += FT / 32) {            // LayerNorm backward + residual, one warp per row
      float m1 = 0.f, m2 = 0.f;
      for (int c = lane; c < a.w; c += 32) {
        const float dxh = xe[rr * a.w + c] * gam[c];
        m1 += dxh;
        m2 = fmaf(dxh, xn[rr * a.w + c], m2);
      }
      m1 = warp_sum(m1) / a.w;
      m2 = warp_sum(m2) / a.w;
      const float rstd = rstd_s[rr];
      for (int c = lane; c < a.w; c += 32) {
        const float dxh = xe[rr * a.w + c] * gam[c];
        const float v = rstd * (dxh - m1 - xn[rr * a.w + c] * m2) + xs[rr * a.w + c];
        stf((T *)a.dx + (size_t)(r0 + rr) * a.w + c, v);
      }
    }
    __syncthreads();
  }
  if (a.acc_in_smem) {
    const int n = a.w * a.hid;
    if ((n & 3) == 0 && (((uintptr_t)a.g_W1 | (uintptr_t)a.g_W2) & 15) == 0 && ((aW1 - sm) & 3) == 0) {
      for (int i = tid; i < n >> 2; i += FT) {         // 16-byte vector reductions (sm_90+)
        const float4 u = ((const float4 *)aW1)[i], v = ((const float4 *)aW2)[i];
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.g_W1 + 4 * i), "f"(u.x), "f"(u.y), "f"(u.z), "f"(u.w) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.g_W2 + 4 * i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    } else {
      for (int i = tid; i < n; i += FT) { atomicAdd(a.g_W1 + i, aW1[i]); atomicAdd(a.g_W2 + i, aW2[i]); }
    }
  }
  for (int i = tid; i < a.hid; i += FT) atomicAdd(a.g_b1 + i, ab1[i]);
  for (int i = tid; i < a.w; i += FT) {
    atomicAdd(a.g_b2 + i, ab2[i]);
    atomicAdd(a.g_gamma + i, ag[i]);
    atomicAdd(a.g_beta + i, ab[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Edge-channel specialisation: width 8, hidden 16 (edge_width = 8 of the MNIST / CLUSTER / PATTERN configs,
// ffn_multiplier = 2).  One thread per (l, m) pair, everything in registers, packed fp32x2 FMAs against
// weights broadcast from shared memory; the row is one 16-byte load / store in bf16.  The backward stages the
// per-pair vectors of a 128-pair chunk in shared memory and every thread owns a fixed subset of the 296
// weight-gradient sums across all chunks of its CTA (atomics once at the end).
namespace {

constexpr int E8W = 8, E8H = 16;

struct F2 { float2 v; };
__device__ __forceinline__ float2 f2ma(float2 acc, float s, float2 w) {
  unsigned long long a = reinterpret_cast<unsigned long long &>(acc);
  float2 ss = make_float2(s, s);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(reinterpret_cast<unsigned long long &>(ss)), "l"(reinterpret_cast<unsigned long long &>(w)));
  return reinterpret_cast<float2 &>(a);
}
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : exp2f(x * kLog2e) - 1.f; }

template <typename T> __device__ __forceinline__ void load_row8(const T *p, float *x);
template <> __device__ __forceinline__ void load_row8<float>(const float *p, float *x) {
  const float4 a = ((const float4 *)p)[0], b = ((const float4 *)p)[1];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <> __device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16 *p, float *x) {
  const uint4 v = *(const uint4 *)p;
  x[0] = __uint_as_float(v.x << 16); x[1] = __uint_as_float(v.x & 0xFFFF0000u);
  x[2] = __uint_as_float(v.y << 16); x[3] = __uint_as_float(v.y & 0xFFFF0000u);
  x[4] = __uint_as_float(v.z << 16); x[5] = __uint_as_float(v.z & 0xFFFF0000u);
  x[6] = __uint_as_float(v.w << 16); x[7] = __uint_as_float(v.w & 0xFFFF0000u);
}
template <typename T> __device__ __forceinline__ void store_row8(T *p, const float *x);
template <> __device__ __forceinline__ void store_row8<float>(float *p, const float *x) {
  ((float4 *)p)[0] = make_float4(x[0], x[1], x[2], x[3]);
  ((float4 *)p)[1] = make_float4(x[4], x[5], x[6], x[7]);
}
template <> __device__ __forceinline__ void store_row8<__nv_bfloat16>(__nv_bfloat16 *p, const float *x) {
  uint4 v;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(x[0], x[1]); v.x = *(uint32_t *)&t;
  t = __floats2bfloat162_rn(x[2], x[3]); v.y = *(uint32_t *)&t;
  t = __floats2bfloat162_rn(x[4], x[5]); v.z = *(uint32_t *)&t;
  t = __floats2bfloat162_rn(x[6], x[7]); v.w = *(uint32_t *)&t;
  *(uint4 *)p = v;
}

// shared weights: W1[8][16] | b1[16] | W2[16][8] | b2[8] | gamma[8] | beta[8] | W2T[8][16] | W1T[16][8]
constexpr int E8_W1 = 0, E8_B1 = 128, E8_W2 = 144, E8_B2 = 272, E8_G = 280, E8_BT = 288, E8_W2T = 296, E8_W1T = 424,
              E8_WTOT = 552;

__device__ __forceinline__ void e8_stage(const FfnArgs &a, float *sw) {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    sw[E8_W1 + i] = a.W1[i];
    sw[E8_W2 + i] = a.W2[i];
    sw[E8_W2T + (i % 8) * 16 + i / 8] = a.W2[i];      // W2[j][c] -> W2T[c][j]
    sw[E8_W1T + (i % 16) * 8 + i / 16] = a.W1[i];     // W1[c][j] -> W1T[j][c]
  }
  if (threadIdx.x < 16) sw[E8_B1 + threadIdx.x] = a.b1[threadIdx.x];
  if (threadIdx.x < 8) {
    sw[E8_B2 + threadIdx.x] = a.b2[threadIdx.x];
    sw[E8_G + threadIdx.x] = a.gamma[threadIdx.x];
    sw[E8_BT + threadIdx.x] = a.beta[threadIdx.x];
  }
}

// LayerNorm + first layer for one row: xn (normalised), xe (affine), pre (16), act (16); returns rstd
__device__ __forceinline__ float e8_forward_row(const float *sw, const float *x, float eps, int act, float *xn, float *xe,
                                                float *pre, float *hid) {
  float mu = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
  mu *= 0.125f;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) { const float d = x[c] - mu; var = fmaf(d, d, var); }
  const float rstd = rsqrtf(fmaf(var, 0.125f, eps));
#pragma unroll
  for (int c = 0; c < 8; ++c) { xn[c] = (x[c] - mu) * rstd; xe[c] = fmaf(xn[c], sw[E8_G + c], sw[E8_BT + c]); }
  float2 a2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a2[q] = *(const float2 *)(sw + E8_B1 + 2 * q);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = *(const float4 *)(sw + E8_W1 + c * 16 + 4 * q);
      a2[2 * q] = f2ma(a2[2 * q], xe[c], make_float2(w.x, w.y));
      a2[2 * q + 1] = f2ma(a2[2 * q + 1], xe[c], make_float2(w.z, w.w));
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    pre[2 * q] = a2[q].x; pre[2 * q + 1] = a2[q].y;
    hid[2 * q] = act == EGT_ACT_ELU ? elu_f(a2[q].x) : act_fwd(act, a2[q].x);
    hid[2 * q + 1] = act == EGT_ACT_ELU ? elu_f(a2[q].y) : act_fwd(act, a2[q].y);
  }
  return rstd;
}

}  // namespace

template <typename T>
__global__ void __launch_bounds__(128) ffn8_fwd_kernel(FfnArgs a) {
  __shared__ __align__(16) float sw[E8_WTOT];
  e8_stage(a, sw);
  __syncthreads();
  for (long long r = (long long)blockIdx.x * 128 + threadIdx.x; r < a.rows; r += (long long)gridDim.x * 128) {
    float x[8], xn[8], xe[8], pre[16], hid[16];
    load_row8<T>((const T *)a.x + r * 8, x);
    e8_forward_row(sw, x, a.eps, a.act, xn, xe, pre, hid);
    float2 o2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) o2[q] = *(const float2 *)(sw + E8_B2 + 2 * q);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 w0 = *(const float4 *)(sw + E8_W2 + j * 8), w1 = *(const float4 *)(sw + E8_W2 + j * 8 + 4);
      o2[0] = f2ma(o2[0], hid[j], make_float2(w0.x, w0.y)); o2[1] = f2ma(o2[1], hid[j], make_float2(w0.z, w0.w));
      o2[2] = f2ma(o2[2], hid[j], make_float2(w1.x, w1.y)); o2[3] = f2ma(o2[3], hid[j], make_float2(w1.z, w1.w));
    }
    float y[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) { y[2 * q] = o2[q].x + x[2 * q]; y[2 * q + 1] = o2[q].y + x[2 * q + 1]; }
    store_row8<T>((T *)a.y + r * 8, y);
  }
}

// staged per-pair record (64 floats): xe[8] | dxe*xn[8] | dxe[8] | dpre[16] | hid[16] | dy[8]
template <typename T>
__global__ void __launch_bounds__(128) ffn8_bwd_kernel(FfnArgs a) {
  __shared__ __align__(16) float sw[E8_WTOT];
  extern __shared__ __align__(8) float rec[];          // [128][RS]
  constexpr int RS = 66;                              // even: the paired column sums use 8-byte loads
  e8_stage(a, sw);
  const int tid = threadIdx.x;
  // this thread's weight-gradient outputs: TWO adjacent products sharing their first factor
  //   tid <  64: dW1[c][j], dW1[c][j+1]   (c = tid / 8,  j = 2 * (tid % 8))    = sum xe[c] * dpre[j..j+1]
  //   tid >= 64: dW2[j][c], dW2[j][c+1]   (j = (tid-64) / 4, c = 2 * ((tid-64) % 4)) = sum hid[j] * dy[c..c+1]
  // and, for tid < 40, one plain column sum: db1[16] | db2[8] | dgamma[8] | dbeta[8]
  const int pa = tid < 64 ? tid / 8 : 40 + (tid - 64) / 4;
  const int pb = tid < 64 ? 24 + 2 * (tid % 8) : 56 + 2 * ((tid - 64) % 4);
  const int ps = tid < 16 ? 24 + tid : tid < 24 ? 56 + (tid - 16) : tid < 32 ? 8 + (tid - 24) : 16 + (tid - 32);
  float2 acc2 = make_float2(0.f, 0.f);
  float accs = 0.f;
  __syncthreads();
  const long long nchunks = (a.rows + 127) / 128;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long r = ch * 128 + tid;
    float *my = rec + tid * RS;
    if (r < a.rows) {
      float x[8], xn[8], xe[8], pre[16], hid[16], dy[8];
      load_row8<T>((const T *)a.x + r * 8, x);
      load_row8<T>((const T *)a.dy + r * 8, dy);
      const float rstd = e8_forward_row(sw, x, a.eps, a.act, xn, xe, pre, hid);
      // d hid = dy W2^T ; d pre = d hid * act'(pre)
      float2 d2[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) d2[q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = *(const float4 *)(sw + E8_W2T + c * 16 + 4 * q);
          d2[2 * q] = f2ma(d2[2 * q], dy[c], make_float2(w.x, w.y));
          d2[2 * q + 1] = f2ma(d2[2 * q + 1], dy[c], make_float2(w.z, w.w));
        }
      }
      float dpre[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float g0 = a.act == EGT_ACT_ELU ? (pre[2 * q] > 0.f ? 1.f : hid[2 * q] + 1.f) : act_bwd(a.act, pre[2 * q]);
        const float g1 = a.act == EGT_ACT_ELU ? (pre[2 * q + 1] > 0.f ? 1.f : hid[2 * q + 1] + 1.f) : act_bwd(a.act, pre[2 * q + 1]);
        dpre[2 * q] = d2[q].x * g0; dpre[2 * q + 1] = d2[q].y * g1;
      }
      // d e^ = d pre W1^T
      float2 e2[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) e2[q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 w0 = *(const float4 *)(sw + E8_W1T + j * 8), w1 = *(const float4 *)(sw + E8_W1T + j * 8 + 4);
        e2[0] = f2ma(e2[0], dpre[j], make_float2(w0.x, w0.y)); e2[1] = f2ma(e2[1], dpre[j], make_float2(w0.z, w0.w));
        e2[2] = f2ma(e2[2], dpre[j], make_float2(w1.x, w1.y)); e2[3] = f2ma(e2[3], dpre[j], make_float2(w1.z, w1.w));
      }
      float dxe[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) { dxe[2 * q] = e2[q].x; dxe[2 * q + 1] = e2[q].y; }
      float m1 = 0.f, m2 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float dxh = dxe[c] * sw[E8_G + c]; m1 += dxh; m2 = fmaf(dxh, xn[c], m2); }
      m1 *= 0.125f; m2 *= 0.125f;
      float dx[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) dx[c] = fmaf(rstd, fmaf(dxe[c], sw[E8_G + c], -fmaf(xn[c], m2, m1)), dy[c]);
      store_row8<T>((T *)a.dx + r * 8, dx);
#pragma unroll
      for (int c = 0; c < 8; ++c) { my[c] = xe[c]; my[8 + c] = dxe[c] * xn[c]; my[16 + c] = dxe[c]; my[56 + c] = dy[c]; }
#pragma unroll
      for (int j = 0; j < 16; ++j) { my[24 + j] = dpre[j]; my[40 + j] = hid[j]; }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) my[i] = 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 128; ++rr) {
      const float fa = rec[rr * RS + pa];
      const float2 fb = *(const float2 *)(rec + rr * RS + pb);
      acc2 = f2ma(acc2, fa, fb);
    }
    if (tid < 40) {
#pragma unroll 8
      for (int rr = 0; rr < 128; ++rr) accs += rec[rr * RS + ps];
    }
    __syncthreads();
  }
  if (tid < 64) {
    atomicAdd(a.g_W1 + (tid / 8) * 16 + 2 * (tid % 8), acc2.x);
    atomicAdd(a.g_W1 + (tid / 8) * 16 + 2 * (tid % 8) + 1, acc2.y);
  } else {
    atomicAdd(a.g_W2 + ((tid - 64) / 4) * 8 + 2 * ((tid - 64) % 4), acc2.x);
    atomicAdd(a.g_W2 + ((tid - 64) / 4) * 8 + 2 * ((tid - 64) % 4) + 1, acc2.y);
  }
  if (tid < 16) atomicAdd(a.g_b1 + tid, accs);
  else if (tid < 24) atomicAdd(a.g_b2 + (tid - 16), accs);
  else if (tid < 32) atomicAdd(a.g_gamma + (tid - 24), accs);
  else if (tid < 40) atomicAdd(a.g_beta + (tid - 32), accs);
}


// ------------------------------------------------------------------------------------------------
static int ffn_plan(FfnArgs &a, int backward, size_t &smem) {
  const size_t w = a.w, h = a.hid;
  const size_t fixed = h + 3 * w;
  const size_t wts = 2 * w * h;
  const size_t accv = backward ? h + 3 * w : 0;
  const size_t limit = (size_t)200 * 1024 / sizeof(float);
  for (int acc_smem = 1; acc_smem >= 0; --acc_smem) {
    const size_t acc = backward && acc_smem ? 2 * w * h : 0;
    for (int rb = 32; rb >= 8; rb >>= 1) {
      const size_t tiles = (size_t)rb * (3 * w + (backward ? 2 : 1) * h) + rb;
      for (int in_smem = 1; in_smem >= 0; --in_smem) {
        const size_t tot = fixed + (in_smem ? wts : 0) + acc + accv + tiles;
        if (tot <= limit) {
          a.rb = rb; a.w_in_smem = in_smem; a.acc_in_smem = acc_smem; smem = tot * sizeof(float);
          return EGT_OK;
        }
      }
    }
  }
  set_error(EGT_E_SHAPE, "ffn: width %d / hidden %d does not fit the shared-memory plan", a.w, a.hid);
  return EGT_E_SHAPE;
}

}  // namespace egt

using namespace egt;

extern "C" int egt_ffn_fwd_ws(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *ws, size_t ws_bytes,
                              void *stream) {
  EGT_REQUIRE(cfg && w && x && y, EGT_E_ARG, "ffn_fwd: NULL argument");
  EGT_REQUIRE(cfg->rows > 0 && cfg->width > 0 && cfg->hidden > 0, EGT_E_SHAPE, "ffn_fwd: rows, width, hidden must be positive");
  EGT_REQUIRE(cfg->dtype == EGT_F32 || cfg->dtype == EGT_BF16, EGT_E_DTYPE, "ffn_fwd: dtype must be EGT_F32 or EGT_BF16");
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = cfg->rows; a.w = cfg->width; a.hid = cfg->hidden; a.act = cfg->activation; a.eps = cfg->ln_eps;
  a.gamma = w->norm_gamma; a.beta = w->norm_beta; a.W1 = w->lr1_kernel; a.b1 = w->lr1_bias; a.W2 = w->lr2_kernel; a.b2 = w->lr2_bias;
  a.x = x; a.y = y;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int rc = ffn_tc_fwd_launch(cfg, w, x, y, st);     // tensor-core path (ffn_tc.cu) when the shape is served
    if (rc != 1) return rc;
  }
  if (ws && ffn_blas_supported(cfg) && ws_bytes >= ffn_blas_workspace_bytes(cfg) && ((uintptr_t)ws & 255) == 0 &&
      ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0)
    return ffn_blas_fwd(cfg, w, x, y, ws, st);             // cuBLAS path (node_blas.cu)
  if (a.w == E8W && a.hid == E8H && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {   // edge channel of the fused widths
    const long long nch = (a.rows + 127) / 128;
    const unsigned grid = (unsigned)(nch < 148 * 16 ? nch : 148 * 16);
    LaunchScope _ls("ffn8_fwd_kernel", st);
    if (cfg->dtype == EGT_F32) ffn8_fwd_kernel<float><<<grid, 128, 0, st>>>(a);
    else ffn8_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(a);
    EGT_CHECK_CUDA(cudaGetLastError());
    return EGT_OK;
  }
  size_t smem = 0;
  int rc = ffn_plan(a, 0, smem);
  if (rc) return rc;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  const long long cap = 148 * 2;
  const unsigned grid = (unsigned)(nchunks < cap ? nchunks : cap);
  LaunchScope _ls("ffn_fwd_kernel", st);
  if (cfg->dtype == EGT_F32) ffn_fwd_kernel<float><<<grid, FT, smem, st>>>(a);
  else ffn_fwd_kernel<__nv_bfloat16><<<grid, FT, smem, st>>>(a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

extern "C" int egt_ffn_bwd_ws(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                              const void *dy, void *dx, void *ws, size_t ws_bytes, void *stream) {
  EGT_REQUIRE(cfg && w && g && x && dy && dx, EGT_E_ARG, "ffn_bwd: NULL argument");
  EGT_REQUIRE(cfg->rows > 0 && cfg->width > 0 && cfg->hidden > 0, EGT_E_SHAPE, "ffn_bwd: rows, width, hidden must be positive");
  EGT_REQUIRE(cfg->dtype == EGT_F32 || cfg->dtype == EGT_BF16, EGT_E_DTYPE, "ffn_bwd: dtype must be EGT_F32 or EGT_BF16");
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = cfg->rows; a.w = cfg->width; a.hid = cfg->hidden; a.act = cfg->activation; a.eps = cfg->ln_eps;
  a.gamma = w->norm_gamma; a.beta = w->norm_beta; a.W1 = w->lr1_kernel; a.b1 = w->lr1_bias; a.W2 = w->lr2_kernel; a.b2 = w->lr2_bias;
  a.g_gamma = g->norm_gamma; a.g_beta = g->norm_beta; a.g_W1 = g->lr1_kernel; a.g_b1 = g->lr1_bias; a.g_W2 = g->lr2_kernel; a.g_b2 = g->lr2_bias;
  a.x = x; a.dy = dy; a.dx = dx;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int rc = ffn_tc_bwd_launch(cfg, w, g, x, dy, dx, st);
    if (rc != 1) return rc;
  }
  if (ws && ffn_blas_supported(cfg) && ws_bytes >= ffn_blas_workspace_bytes(cfg) && ((uintptr_t)ws & 255) == 0 &&
      ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0)
    return ffn_blas_bwd(cfg, w, g, x, dy, dx, ws, st);
  if (a.w == E8W && a.hid == E8H && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0) {
    const long long nch = (a.rows + 127) / 128;
    const unsigned grid = (unsigned)(nch < 148 * 8 ? nch : 148 * 8);
    const size_t rsm = (size_t)128 * 66 * sizeof(float);
    LaunchScope _ls("ffn8_bwd_kernel", st);
    if (cfg->dtype == EGT_F32) ffn8_bwd_kernel<float><<<grid, 128, rsm, st>>>(a);
    else ffn8_bwd_kernel<__nv_bfloat16><<<grid, 128, rsm, st>>>(a);
    EGT_CHECK_CUDA(cudaGetLastError());
    return EGT_OK;
  }
  size_t smem = 0;
  int rc = ffn_plan(a, 1, smem);
  if (rc) return rc;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  const long long cap = 148 * 2;
  const unsigned grid = (unsigned)(nchunks < cap ? nchunks : cap);
  LaunchScope _ls("ffn_bwd_kernel", st);
  if (cfg->dtype == EGT_F32) ffn_bwd_kernel<float><<<grid, FT, smem, st>>>(a);
  else ffn_bwd_kernel<__nv_bfloat16><<<grid, FT, smem, st>>>(a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

extern "C" int egt_ffn_fwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *stream) {
  return egt_ffn_fwd_ws(cfg, w, x, y, nullptr, 0, stream);
}
extern "C" int egt_ffn_bwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                           const void *dy, void *dx, void *stream) {
  return egt_ffn_bwd_ws(cfg, w, g, x, dy, dx, nullptr, 0, stream);
}
// Bytes of device workspace egt_ffn_{fwd,bwd}_ws can use for this shape (0: none needed).
extern "C" size_t egt_ffn_workspace_bytes(const egt_ffn_cfg_t *cfg) {
  if (!cfg) return 0;
  if (ffn_tc_serves(cfg)) return 0;
  return ffn_blas_supported(cfg) ? ffn_blas_workspace_bytes(cfg) : 0;
}
