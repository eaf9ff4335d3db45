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

namespace egt {

namespace {

constexpr int FT = 128;   // threads per CTA

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
    for (int i = tid; i < a.w * a.hid; i += FT) { s1[i] = a.W1[i]; s2[i] = a.W2[i]; }
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
      for (int c = 0; c < a.w; ++c) acc = fmaf(xs[rr * a.w + c], W2[j * a.w + c], acc);
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
      for (int j = 0; j < a.hid; ++j) acc = fmaf(ds[rr * a.hid + j], W1[c * a.hid + j], acc);
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
  if (a.acc_in_smem)
    for (int i = tid; i < a.w * a.hid; i += FT) { atomicAdd(a.g_W1 + i, aW1[i]); atomicAdd(a.g_W2 + i, aW2[i]); }
  for (int i = tid; i < a.hid; i += FT) atomicAdd(a.g_b1 + i, ab1[i]);
  for (int i = tid; i < a.w; i += FT) {
    atomicAdd(a.g_b2 + i, ab2[i]);
    atomicAdd(a.g_gamma + i, ag[i]);
    atomicAdd(a.g_beta + i, ab[i]);
  }
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

extern "C" int egt_ffn_fwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *stream) {
  EGT_REQUIRE(cfg && w && x && y, EGT_E_ARG, "ffn_fwd: NULL argument");
  EGT_REQUIRE(cfg->rows > 0 && cfg->width > 0 && cfg->hidden > 0, EGT_E_SHAPE, "ffn_fwd: rows, width, hidden must be positive");
  EGT_REQUIRE(cfg->dtype == EGT_F32 || cfg->dtype == EGT_BF16, EGT_E_DTYPE, "ffn_fwd: dtype must be EGT_F32 or EGT_BF16");
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = cfg->rows; a.w = cfg->width; a.hid = cfg->hidden; a.act = cfg->activation; a.eps = cfg->ln_eps;
  a.gamma = w->norm_gamma; a.beta = w->norm_beta; a.W1 = w->lr1_kernel; a.b1 = w->lr1_bias; a.W2 = w->lr2_kernel; a.b2 = w->lr2_bias;
  a.x = x; a.y = y;
  size_t smem = 0;
  int rc = ffn_plan(a, 0, smem);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  const unsigned grid = (unsigned)(nchunks < 148 * 4 ? nchunks : 148 * 4);
  LaunchScope _ls("ffn_fwd_kernel", st);
  if (cfg->dtype == EGT_F32) ffn_fwd_kernel<float><<<grid, FT, smem, st>>>(a);
  else ffn_fwd_kernel<__nv_bfloat16><<<grid, FT, smem, st>>>(a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

extern "C" int egt_ffn_bwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                           const void *dy, void *dx, void *stream) {
  EGT_REQUIRE(cfg && w && g && x && dy && dx, EGT_E_ARG, "ffn_bwd: NULL argument");
  EGT_REQUIRE(cfg->rows > 0 && cfg->width > 0 && cfg->hidden > 0, EGT_E_SHAPE, "ffn_bwd: rows, width, hidden must be positive");
  EGT_REQUIRE(cfg->dtype == EGT_F32 || cfg->dtype == EGT_BF16, EGT_E_DTYPE, "ffn_bwd: dtype must be EGT_F32 or EGT_BF16");
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = cfg->rows; a.w = cfg->width; a.hid = cfg->hidden; a.act = cfg->activation; a.eps = cfg->ln_eps;
  a.gamma = w->norm_gamma; a.beta = w->norm_beta; a.W1 = w->lr1_kernel; a.b1 = w->lr1_bias; a.W2 = w->lr2_kernel; a.b2 = w->lr2_bias;
  a.g_gamma = g->norm_gamma; a.g_beta = g->norm_beta; a.g_W1 = g->lr1_kernel; a.g_b1 = g->lr1_bias; a.g_W2 = g->lr2_kernel; a.g_b2 = g->lr2_bias;
  a.x = x; a.dy = dy; a.dx = dx;
  size_t smem = 0;
  int rc = ffn_plan(a, 1, smem);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EGT_CHECK_CUDA(cudaFuncSetAttribute(ffn_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nchunks = (a.rows + a.rb - 1) / a.rb;
  const unsigned grid = (unsigned)(nchunks < 148 * 2 ? nchunks : 148 * 2);
  LaunchScope _ls("ffn_bwd_kernel", st);
  if (cfg->dtype == EGT_F32) ffn_bwd_kernel<float><<<grid, FT, smem, st>>>(a);
  else ffn_bwd_kernel<__nv_bfloat16><<<grid, FT, smem, st>>>(a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}
