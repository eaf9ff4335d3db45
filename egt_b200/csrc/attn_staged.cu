// attn_staged.cu -- "staged" EGT-layer kernels: one thread per (graph b, row/col, head).
//
// These implement the reference EGT layer exactly as it is called
// (lib/models/egt_layers.py:57-143 gated, :145-213 ungated): pre-projected E and G come from
// HBM, H_hat goes back to HBM.  They accept any h, dk <= 16, fp32 or bf16 activations, every
// flag of the layer, and are the path behind egt_attn_fwd/egt_attn_bwd.  The block-level fused
// tcgen05 kernel (attn_fused_fwd.cu) replaces them on the headline bf16 shapes.
//
// Softmax is computed online (running max / sum) in fp32; masks are added in fp32 in the
// reference's order so that masked probabilities / gates are exactly zero.
#include "common.cuh"
#include "kernels.h"

namespace egt {

constexpr int DKMAX = 16;

template <typename T>
struct LogitCtx {
  const AttnParams &P;
  int b, hh;
  __device__ LogitCtx(const AttnParams &p, int b_, int hh_) : P(p), b(b_), hh(hh_) {}

  // returns un-masked H_hat; x = masked logit; gin = masked gate logit; keep = dropout scale
  __device__ __forceinline__ float eval(int l, int m, const float *q, const float *k, float &S_raw, float &x,
                                        float &gin, float &keep) const {
    float s = 0.f;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd)
      if (dd < P.dk) s += q[dd] * k[dd];
    s *= P.scale;                                            // egt_layers.py:79
    S_raw = s;
    if (P.has_clip) s = fminf(fmaxf(s, P.clip_lo), P.clip_hi);   // :81-82
    size_t pe = (((size_t)b * P.N + l) * P.N + m) * P.h + hh;
    float Hh = s;
    if (P.E) Hh += ldf((const T *)P.E + pe);                 // :85-86
    x = Hh;
    gin = P.G ? ldf((const T *)P.G + pe) : 0.f;
    if (P.mask) {                                            // :91-94
      float neg = ((float)P.mask[(size_t)b * P.N + m] - 1.f) * kNegMask;
      x += neg; gin += neg;
    }
    if (P.attn_mask == EGT_MASK_DENSE) {                     // :96-101
      float neg = (ldf((const T *)P.M + pe) - 1.f) * kNegMask;
      x += neg; gin += neg;
    } else if (P.attn_mask == EGT_MASK_ADJ_U8) {
      float neg = ((float)((const uint8_t *)P.M)[((size_t)b * P.N + l) * P.N + m] - 1.f) * kNegMask;
      x += neg; gin += neg;
    }
    if (P.rand_mask) {                                       // :103-108 (same noise for H_hat and G)
      float u = rng_uniform(P.seed, P.offset + (P.offset_dev ? *P.offset_dev : 0ull), 0u, rng_elem_index(b, l, m, hh, P.N, P.h));
      float neg = u < P.random_mask_prob ? -kNegMask : 0.f;
      x += neg; gin += neg;
    }
    keep = 1.f;
    if (P.dropout) {                                         // :116-117 tf.nn.dropout
      float u = rng_uniform(P.seed, P.offset + (P.offset_dev ? *P.offset_dev : 0ull), 1u, rng_elem_index(b, l, m, hh, P.N, P.h));
      keep = u >= P.attn_dropout ? 1.f / (1.f - P.attn_dropout) : 0.f;
    }
    return Hh;
  }
};

__device__ __forceinline__ float scaler_of(const AttnParams &P, int l, float deg) {
  if (!P.scale_degree) return 1.f;
  if (l < P.num_virtual_nodes) return 1.f;                   // egt_layers.py:131-135
  return P.scaler_type == EGT_SCALER_LOG ? log1pf(deg) : deg;   // :125-130
}

// --------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_kernel(AttnParams P) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)P.B * P.N * P.h) return;
  int hh = idx % P.h;
  int l = (idx / P.h) % P.N;
  int b = idx / ((size_t)P.h * P.N);
  const int d = P.h * P.dk;
  const T *qkv = (const T *)P.qkv;
  float q[DKMAX], o[DKMAX], k[DKMAX];
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd) {
    q[dd] = dd < P.dk ? ldf(qkv + ((size_t)b * P.N + l) * 3 * d + dd * P.h + hh) : 0.f;
    o[dd] = 0.f;
    k[dd] = 0.f;
  }
  LogitCtx<T> ctx(P, b, hh);
  float mrun = -INFINITY, sum = 0.f, deg = 0.f;
  for (int m = 0; m < P.N; ++m) {
    const T *krow = qkv + ((size_t)b * P.N + m) * 3 * d + d;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd)
      if (dd < P.dk) k[dd] = ldf(krow + dd * P.h + hh);
    float S_raw, x, gin, keep;
    float Hh = ctx.eval(l, m, q, k, S_raw, x, gin, keep);
    size_t pe = (((size_t)b * P.N + l) * P.N + m) * P.h + hh;
    if (P.h_hat) stf((T *)P.h_hat + pe, Hh);
    if (x > mrun) {
      float corr = __expf(mrun - x);
      sum *= corr;
#pragma unroll
      for (int dd = 0; dd < DKMAX; ++dd) o[dd] *= corr;
      mrun = x;
    }
    float p = __expf(x - mrun);
    sum += p;
    float g = 1.f;
    if (P.G) { g = sigmoid_f(gin); deg += g; }
    float a = p * g * keep;
    const T *vrow = krow + d;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd)
      if (dd < P.dk) o[dd] += a * ldf(vrow + dd * P.h + hh);
  }
  const float lsum = __logf(sum);          // (row max, log row sum) are kept apart: a fused lse would
  const float lse = mrun;                  // absorb log(sum) when every key is masked (mrun = -1e9)
  float inv = 1.f / sum;
  float s = scaler_of(P, l, deg);
  size_t ps = ((size_t)b * P.N + l) * P.h + hh;
  P.lse[ps] = mrun;
  P.lse[(size_t)P.B * P.N * P.h + ps] = lsum;
  P.deg[ps] = deg;
  T *vo = (T *)P.v_att + ((size_t)b * P.N + l) * d;
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd)
    if (dd < P.dk) stf(vo + dd * P.h + hh, o[dd] * inv * s);

  if (P.a_tild) {   // third output of the layer (Analysis taps only) -- second sweep with the final lse
    for (int m = 0; m < P.N; ++m) {
      const T *krow = qkv + ((size_t)b * P.N + m) * 3 * d + d;
#pragma unroll
      for (int dd = 0; dd < DKMAX; ++dd)
        if (dd < P.dk) k[dd] = ldf(krow + dd * P.h + hh);
      float S_raw, x, gin, keep;
      ctx.eval(l, m, q, k, S_raw, x, gin, keep);
      float p = __expf((x - lse) - lsum);
      float g = P.G ? sigmoid_f(gin) : 1.f;
      size_t pe = (((size_t)b * P.N + l) * P.N + m) * P.h + hh;
      stf((T *)P.a_tild + pe, p * g * keep);
    }
  }
}

// --------------------------------------------------------------------------------------
// backward, row pass: thread (b,l,hh) owns dQ[l,:,hh]; writes dE (= dH_hat), dG, and the row
// terms D = sum_m dA*A and the scaler s into row_ws for the column pass.   
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_row_kernel(AttnParams P) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)P.B * P.N * P.h) return;
  int hh = idx % P.h;
  int l = (idx / P.h) % P.N;
  int b = idx / ((size_t)P.h * P.N);
  const int d = P.h * P.dk;
  const T *qkv = (const T *)P.qkv;
  float q[DKMAX], dva[DKMAX], k[DKMAX], dq[DKMAX];
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd) {
    bool in = dd < P.dk;
    q[dd] = in ? ldf(qkv + ((size_t)b * P.N + l) * 3 * d + dd * P.h + hh) : 0.f;
    dva[dd] = in ? ldf((const T *)P.d_v_att + ((size_t)b * P.N + l) * d + dd * P.h + hh) : 0.f;
    k[dd] = 0.f;
    dq[dd] = 0.f;
  }
  size_t ps = ((size_t)b * P.N + l) * P.h + hh;
  const float lse = P.lse[ps];
  const float lsum = P.lse[(size_t)P.B * P.N * P.h + ps];
  const float deg = P.deg[ps];
  const float s = scaler_of(P, l, deg);
  LogitCtx<T> ctx(P, b, hh);

  // pass 1: ds = sum_m A_d[m] * (dV_att . V[m])
  float ds = 0.f;
  for (int m = 0; m < P.N; ++m) {
    const T *krow = qkv + ((size_t)b * P.N + m) * 3 * d + d;
    const T *vrow = krow + d;
    float dAp = 0.f;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd)
      if (dd < P.dk) {
        k[dd] = ldf(krow + dd * P.h + hh);
        dAp += dva[dd] * ldf(vrow + dd * P.h + hh);
      }
    float S_raw, x, gin, keep;
    ctx.eval(l, m, q, k, S_raw, x, gin, keep);
    float p = __expf((x - lse) - lsum);
    float g = P.G ? sigmoid_f(gin) : 1.f;
    ds += p * g * keep * dAp;
  }
  const float D = s * ds;
  float ddeg = 0.f;
  if (P.scale_degree && l >= P.num_virtual_nodes)
    ddeg = P.scaler_type == EGT_SCALER_LOG ? ds / (1.f + deg) : ds;
  P.row_ws[ps] = D;
  P.row_ws[(size_t)P.B * P.N * P.h + ps] = s;

  // pass 2
  for (int m = 0; m < P.N; ++m) {
    const T *krow = qkv + ((size_t)b * P.N + m) * 3 * d + d;
    const T *vrow = krow + d;
    float dAp = 0.f;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd)
      if (dd < P.dk) {
        k[dd] = ldf(krow + dd * P.h + hh);
        dAp += dva[dd] * ldf(vrow + dd * P.h + hh);
      }
    float S_raw, x, gin, keep;
    float Hh = ctx.eval(l, m, q, k, S_raw, x, gin, keep);
    float p = __expf((x - lse) - lsum);
    float g = P.G ? sigmoid_f(gin) : 1.f;
    float dA = s * dAp * keep;
    float dP = dA * g;
    float dH = p * (dP - D);
    size_t pe = (((size_t)b * P.N + l) * P.N + m) * P.h + hh;
    if (P.d_h_hat) dH += ldf((const T *)P.d_h_hat + pe);
    if (P.h_hat) stf((T *)P.h_hat + pe, Hh);
    if (P.dG) {
      float dg = dA * p + ddeg;
      stf((T *)P.dG + pe, dg * g * (1.f - g));
    }
    if (P.dE) stf((T *)P.dE + pe, dH);
    bool inside = !P.has_clip || (S_raw >= P.clip_lo && S_raw <= P.clip_hi);
    float dS = inside ? dH * P.scale : 0.f;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd) dq[dd] += dS * k[dd];
  }
  T *dqo = (T *)P.d_qkv + ((size_t)b * P.N + l) * 3 * d;
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd)
    if (dd < P.dk) stf(dqo + dd * P.h + hh, dq[dd] * P.dq_scale);
}

// backward, column pass: thread (b,m,hh) owns dK[m,:,hh] and dV[m,:,hh].
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_col_kernel(AttnParams P) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)P.B * P.N * P.h) return;
  int hh = idx % P.h;
  int m = (idx / P.h) % P.N;
  int b = idx / ((size_t)P.h * P.N);
  const int d = P.h * P.dk;
  const T *qkv = (const T *)P.qkv;
  float k[DKMAX], v[DKMAX], q[DKMAX], dkacc[DKMAX], dvacc[DKMAX];
  const T *krow = qkv + ((size_t)b * P.N + m) * 3 * d + d;
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd) {
    bool in = dd < P.dk;
    k[dd] = in ? ldf(krow + dd * P.h + hh) : 0.f;
    v[dd] = in ? ldf(krow + d + dd * P.h + hh) : 0.f;
    q[dd] = 0.f;
    dkacc[dd] = 0.f;
    dvacc[dd] = 0.f;
  }
  LogitCtx<T> ctx(P, b, hh);
  const size_t rs = (size_t)P.B * P.N * P.h;
  for (int l = 0; l < P.N; ++l) {
    size_t ps = ((size_t)b * P.N + l) * P.h + hh;
    const float lse = P.lse[ps];
    const float lsum = P.lse[rs + ps];
    const float D = P.row_ws[ps];
    const float s = P.row_ws[rs + ps];
    const T *qrow = qkv + ((size_t)b * P.N + l) * 3 * d;
    const T *dvrow = (const T *)P.d_v_att + ((size_t)b * P.N + l) * d;
    float dva[DKMAX];
    float dAp = 0.f;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd) {
      bool in = dd < P.dk;
      q[dd] = in ? ldf(qrow + dd * P.h + hh) : 0.f;
      dva[dd] = in ? ldf(dvrow + dd * P.h + hh) : 0.f;
      dAp += dva[dd] * v[dd];
    }
    float S_raw, x, gin, keep;
    ctx.eval(l, m, q, k, S_raw, x, gin, keep);
    float p = __expf((x - lse) - lsum);
    float g = P.G ? sigmoid_f(gin) : 1.f;
    float a_d = p * g * keep;
    float dA = s * dAp * keep;
    float dH = p * (dA * g - D);
    size_t pe = (((size_t)b * P.N + l) * P.N + m) * P.h + hh;
    if (P.d_h_hat) dH += ldf((const T *)P.d_h_hat + pe);
    bool inside = !P.has_clip || (S_raw >= P.clip_lo && S_raw <= P.clip_hi);
    float dS = inside ? dH * P.scale : 0.f;
    float as = a_d * s;
#pragma unroll
    for (int dd = 0; dd < DKMAX; ++dd) {
      dkacc[dd] += dS * q[dd];
      dvacc[dd] += as * dva[dd];
    }
  }
  T *o = (T *)P.d_qkv + ((size_t)b * P.N + m) * 3 * d + d;
#pragma unroll
  for (int dd = 0; dd < DKMAX; ++dd)
    if (dd < P.dk) {
      stf(o + dd * P.h + hh, dkacc[dd]);
      stf(o + d + dd * P.h + hh, dvacc[dd]);
    }
}

// --------------------------------------------------------------------------------------
template <typename T>
static int launch_fwd(const AttnParams &P, cudaStream_t st) {
  size_t n = (size_t)P.B * P.N * P.h;
  LaunchScope _ls("attn_staged_fwd", st);
  attn_fwd_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}
template <typename T>
static int launch_bwd(const AttnParams &P, cudaStream_t st) {
  size_t n = (size_t)P.B * P.N * P.h;
  unsigned g = (unsigned)((n + 127) / 128);
  { LaunchScope _ls("attn_staged_bwd_row", st); attn_bwd_row_kernel<T><<<g, 128, 0, st>>>(P); }
  EGT_CHECK_CUDA(cudaGetLastError());
  { LaunchScope _ls("attn_staged_bwd_col", st); attn_bwd_col_kernel<T><<<g, 128, 0, st>>>(P); }
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

int attn_staged_fwd(const AttnParams &P, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = attn_fast_launch(0, P, dtype, st); if (rc <= 0) return rc; }
  return dtype == EGT_F32 ? launch_fwd<float>(P, st) : launch_fwd<__nv_bfloat16>(P, st);
}
int attn_staged_bwd(const AttnParams &P, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = attn_fast_launch(1, P, dtype, st); if (rc <= 0) return rc; }
  return dtype == EGT_F32 ? launch_bwd<float>(P, st) : launch_bwd<__nv_bfloat16>(P, st);
}

}  // namespace egt
