// edge_kernels.cu -- staged edge-channel kernels (one thread per (b,l,m) pair).
//
//   edge_proj_fwd : e -> LN_e -> G = e^ W_G + b_G ; E = act(e^ W_E + b_E)
//                   (graph_xformer_model_base.py:194-208, :149-162; 'bias' variant :173-183 has no LN)
//   edge_out_fwd  : e' = H_hat W_r + b_r + e                          (:214-218)
//   edge_out_bwd  : dH_ext = de' W_r^T ; dW_r, db_r
//   edge_proj_bwd : dE,dG -> d e^ -> LN backward -> de (+ de') ; dW_E, dW_G, db, dgamma, dbeta
// Any d_e, h <= 16, fp32 or bf16 activations.  The fused tcgen05 kernel replaces the forward pair
// on the headline shapes.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace egt {

constexpr int HMAX = 16;
constexpr int EPB = 128;   // pairs per block

template <typename T>
__device__ __forceinline__ void ln_stats(const T *ep, int d_e, float eps, float &mu, float &rstd) {
  float s = 0.f;
  for (int c = 0; c < d_e; ++c) s += ldf(ep + c);
  mu = s / d_e;
  float v = 0.f;
  for (int c = 0; c < d_e; ++c) {
    float t = ldf(ep + c) - mu;
    v += t * t;
  }
  rstd = rsqrtf(v / d_e + eps);
}

template <typename T>
__global__ void __launch_bounds__(EPB) edge_proj_fwd_kernel(EdgeParams p) {
  size_t pair = (size_t)blockIdx.x * EPB + threadIdx.x;
  if (pair >= p.pairs) return;
  const T *ep = (const T *)p.e + pair * p.d_e;
  float mu = 0.f, rstd = 1.f;
  if (p.has_ln) ln_stats(ep, p.d_e, p.ln_eps, mu, rstd);
  float aE[HMAX], aG[HMAX];
#pragma unroll
  for (int hh = 0; hh < HMAX; ++hh) { aE[hh] = 0.f; aG[hh] = 0.f; }
  for (int c = 0; c < p.d_e; ++c) {
    float x = ldf(ep + c);
    if (p.has_ln) x = (x - mu) * rstd * __ldg(p.ln_g + c) + __ldg(p.ln_b + c);
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
      if (hh < p.h) {
        aE[hh] += x * __ldg(p.w_e + c * p.h + hh);
        if (p.gated) aG[hh] += x * __ldg(p.w_g + c * p.h + hh);
      }
  }
#pragma unroll
  for (int hh = 0; hh < HMAX; ++hh)
    if (hh < p.h) {
      stf((T *)p.E + pair * p.h + hh, edge_act_fwd(p.act, p.act_alpha, aE[hh] + __ldg(p.b_e + hh)));
      if (p.gated) stf((T *)p.G + pair * p.h + hh, aG[hh] + __ldg(p.b_g + hh));
    }
}

template <typename T>
__global__ void __launch_bounds__(EPB) edge_out_fwd_kernel(EdgeParams p) {
  size_t pair = (size_t)blockIdx.x * EPB + threadIdx.x;
  if (pair >= p.pairs) return;
  float hv[HMAX];
#pragma unroll
  for (int hh = 0; hh < HMAX; ++hh) hv[hh] = hh < p.h ? ldf((const T *)p.h_hat + pair * p.h + hh) : 0.f;
  const T *ep = (const T *)p.e + pair * p.d_e;
  T *eo = (T *)p.e_out + pair * p.d_e;
  for (int c = 0; c < p.d_e; ++c) {
    float acc = __ldg(p.b_r + c);
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
      if (hh < p.h) acc += hv[hh] * __ldg(p.w_r + hh * p.d_e + c);
    stf(eo + c, acc + ldf(ep + c));
  }
}

// Both backward kernels are persistent: a CTA walks pair-chunks of EPB pairs with a grid stride, keeps its
// share of the weight-gradient sums in registers across chunks and issues its atomics ONCE at the end
// (one atomic per output per CTA instead of one per output per 128 pairs).
constexpr int MAXO_R = 8;     // h * d_e   <= 16 * 64 outputs / EPB threads
constexpr int MAXO_P = 16;    // d_e * J   <= 64 * 32 outputs / EPB threads

// dynamic smem: xs[EPB][h] (H_hat), ys[EPB][d_e] (de')
template <typename T>
__global__ void __launch_bounds__(EPB) edge_out_bwd_kernel(EdgeParams p) {
  extern __shared__ float sm[];
  float *xs = sm;
  float *ys = sm + EPB * p.h;
  float accw[MAXO_R], accb = 0.f;
#pragma unroll
  for (int k = 0; k < MAXO_R; ++k) accw[k] = 0.f;
  for (size_t chunk = blockIdx.x; chunk * EPB < p.pairs; chunk += gridDim.x) {
    size_t pair = chunk * EPB + threadIdx.x;
    bool live = pair < p.pairs;
    float dh[HMAX];
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh) dh[hh] = 0.f;
    for (int c = 0; c < p.d_e; ++c) {
      float g = live ? ldf((const T *)p.de_out + pair * p.d_e + c) : 0.f;
      ys[threadIdx.x * p.d_e + c] = g;
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh)
        if (hh < p.h) dh[hh] += g * __ldg(p.w_r + hh * p.d_e + c);
    }
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
      if (hh < p.h) {
        xs[threadIdx.x * p.h + hh] = (live && p.h_hat) ? ldf((const T *)p.h_hat + pair * p.h + hh) : 0.f;
        if (live && p.d_h_ext) stf((T *)p.d_h_ext + pair * p.h + hh, dh[hh]);
      }
    if (!p.g_w_r) continue;                 // uniform over the grid: no barrier is skipped by a subset of threads
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXO_R; ++k) {
      int o = threadIdx.x + k * EPB;
      if (o < p.h * p.d_e) {
        int hh = o / p.d_e, c = o % p.d_e;
        float acc = 0.f;
        for (int q = 0; q < EPB; ++q) acc += xs[q * p.h + hh] * ys[q * p.d_e + c];
        accw[k] += acc;
      }
    }
    if (threadIdx.x < p.d_e) {
      float acc = 0.f;
      for (int q = 0; q < EPB; ++q) acc += ys[q * p.d_e + threadIdx.x];
      accb += acc;
    }
    __syncthreads();
  }
  if (!p.g_w_r) return;
#pragma unroll
  for (int k = 0; k < MAXO_R; ++k) {
    int o = threadIdx.x + k * EPB;
    if (o < p.h * p.d_e) atomicAdd(p.g_w_r + o, accw[k]);
  }
  if (threadIdx.x < p.d_e) atomicAdd(p.g_b_r + threadIdx.x, accb);
}

// dynamic smem: xs[EPB][d_e] (normalised input without affine, or raw e), ds[EPB][J], dbs[J]
template <typename T>
__global__ void __launch_bounds__(EPB) edge_proj_bwd_kernel(EdgeParams p) {
  extern __shared__ float sm[];
  const int J = p.gated ? 2 * p.h : p.h;
  float *xs = sm;
  float *ds = sm + EPB * p.d_e;
  float *dbs = ds + EPB * J;                 // column sums of ds over ALL chunks of this CTA
  const int tid = threadIdx.x;
  float accw[MAXO_P];
#pragma unroll
  for (int k = 0; k < MAXO_P; ++k) accw[k] = 0.f;
  if (tid < J) dbs[tid] = 0.f;
  for (size_t chunk = blockIdx.x; chunk * EPB < p.pairs; chunk += gridDim.x) {
  size_t pair = chunk * EPB + threadIdx.x;
  bool live = pair < p.pairs;
  float dEp[HMAX], dGv[HMAX];
#pragma unroll
  for (int hh = 0; hh < HMAX; ++hh) { dEp[hh] = 0.f; dGv[hh] = 0.f; }
  if (live) {
    const T *ep = (const T *)p.e + pair * p.d_e;
    float mu = 0.f, rstd = 1.f;
    if (p.has_ln) ln_stats(ep, p.d_e, p.ln_eps, mu, rstd);
    float aE[HMAX];
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh) aE[hh] = 0.f;
    for (int c = 0; c < p.d_e; ++c) {
      float xn = ldf(ep + c);
      float x = xn;
      if (p.has_ln) {
        xn = (xn - mu) * rstd;
        x = xn * __ldg(p.ln_g + c) + __ldg(p.ln_b + c);
      }
      xs[tid * p.d_e + c] = xn;
      if (p.act != EGT_ACT_NONE) {
#pragma unroll
        for (int hh = 0; hh < HMAX; ++hh)
          if (hh < p.h) aE[hh] += x * __ldg(p.w_e + c * p.h + hh);
      }
    }
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
      if (hh < p.h) {
        float g = ldf((const T *)p.dE + pair * p.h + hh);
        if (p.act != EGT_ACT_NONE) g *= edge_act_bwd(p.act, p.act_alpha, aE[hh] + __ldg(p.b_e + hh));
        dEp[hh] = g;
        ds[tid * J + hh] = g;
        if (p.gated) {
          dGv[hh] = ldf((const T *)p.dG + pair * p.h + hh);
          ds[tid * J + p.h + hh] = dGv[hh];
        }
      }
    // d e^[c], LN backward
    float m1 = 0.f, m2 = 0.f;
    if (p.has_ln) {
      for (int c = 0; c < p.d_e; ++c) {
        float de_hat = 0.f;
#pragma unroll
        for (int hh = 0; hh < HMAX; ++hh)
          if (hh < p.h) {
            de_hat += dEp[hh] * __ldg(p.w_e + c * p.h + hh);
            if (p.gated) de_hat += dGv[hh] * __ldg(p.w_g + c * p.h + hh);
          }
        float dxh = de_hat * __ldg(p.ln_g + c);
        m1 += dxh;
        m2 += dxh * xs[tid * p.d_e + c];
      }
      m1 /= p.d_e;
      m2 /= p.d_e;
    }
    T *deo = (T *)p.de + pair * p.d_e;
    for (int c = 0; c < p.d_e; ++c) {
      float de_hat = 0.f;
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh)
        if (hh < p.h) {
          de_hat += dEp[hh] * __ldg(p.w_e + c * p.h + hh);
          if (p.gated) de_hat += dGv[hh] * __ldg(p.w_g + c * p.h + hh);
        }
      float out;
      if (p.has_ln) {
        float dxh = de_hat * __ldg(p.ln_g + c);
        out = rstd * (dxh - m1 - xs[tid * p.d_e + c] * m2);
      } else {
        out = de_hat;
      }
      if (p.de_out) out += ldf((const T *)p.de_out + pair * p.d_e + c);
      stf(deo + c, out);
    }
  } else {
    for (int c = 0; c < p.d_e; ++c) xs[tid * p.d_e + c] = 0.f;
    for (int j = 0; j < J; ++j) ds[tid * J + j] = 0.f;
  }
  __syncthreads();
  for (int j = tid; j < J; j += EPB) {
    float acc = 0.f;
    for (int q = 0; q < EPB; ++q) acc += ds[q * J + j];
    dbs[j] += acc;
  }
  // dWn[c,j] += sum_p xn[p,c]*ds[p,j]
#pragma unroll
  for (int k = 0; k < MAXO_P; ++k) {
    int o = tid + k * EPB;
    if (o < p.d_e * J) {
      int c = o / J, j = o % J;
      float acc = 0.f;
      for (int q = 0; q < EPB; ++q) acc += xs[q * p.d_e + c] * ds[q * J + j];
      accw[k] += acc;
    }
  }
  __syncthreads();
  }   // chunk loop
  __syncthreads();
  for (int j = tid; j < J; j += EPB) atomicAdd(j < p.h ? p.g_b_e + j : p.g_b_g + (j - p.h), dbs[j]);
  // dW = gamma*dWn + beta*db ; dgamma[c] += sum_j W[c,j] dWn[c,j] ; dbeta[c] += sum_j W[c,j] db[j]
#pragma unroll
  for (int k = 0; k < MAXO_P; ++k) {
    int o = tid + k * EPB;
    if (o < p.d_e * J) {
      int c = o / J, j = o % J;
      const float acc = accw[k];
      const float *W = j < p.h ? p.w_e : p.w_g;
      float *gW = j < p.h ? p.g_w_e : p.g_w_g;
      int jj = j < p.h ? j : j - p.h;
      if (p.has_ln) {
        float w = __ldg(W + c * p.h + jj);
        atomicAdd(gW + c * p.h + jj, __ldg(p.ln_g + c) * acc + __ldg(p.ln_b + c) * dbs[j]);
        atomicAdd(p.g_ln_g + c, w * acc);
        atomicAdd(p.g_ln_b + c, w * dbs[j]);
      } else {
        atomicAdd(gW + c * p.h + jj, acc);
      }
    }
  }
}

#define DISPATCH_T(dtype, KERNEL, grid, block, smem, st, arg)                       \
  do {                                                                              \
    LaunchScope _ls(#KERNEL, st);                                                   \
    if ((dtype) == EGT_F32) KERNEL<float><<<grid, block, smem, st>>>(arg);          \
    else KERNEL<__nv_bfloat16><<<grid, block, smem, st>>>(arg);                     \
    EGT_CHECK_CUDA(cudaGetLastError());                                             \
  } while (0)

// EGT_STAGED_GENERIC=1 keeps every staged call on the any-shape kernels (edge_kernels.cu, attn_staged.cu);
// read per call so a test can compare both sets in one process
bool staged_force_generic() {
  const char *e = getenv("EGT_STAGED_GENERIC");
  return e && e[0] == '1';
}

static unsigned pair_grid(const EdgeParams &p) { return (unsigned)((p.pairs + EPB - 1) / EPB); }
static unsigned persistent_grid(const EdgeParams &p) {
  unsigned g = pair_grid(p);
  return g < 148u * 8u ? g : 148u * 8u;
}

int edge_proj_fwd_launch(const EdgeParams &p, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = edge_fast_launch(0, p, dtype, st); if (rc <= 0) return rc; }
  DISPATCH_T(dtype, edge_proj_fwd_kernel, pair_grid(p), EPB, 0, st, p);
  return EGT_OK;
}
int edge_out_fwd_launch(const EdgeParams &p, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = edge_fast_launch(1, p, dtype, st); if (rc <= 0) return rc; }
  DISPATCH_T(dtype, edge_out_fwd_kernel, pair_grid(p), EPB, 0, st, p);
  return EGT_OK;
}
int edge_out_bwd_launch(const EdgeParams &p, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = edge_fast_launch(2, p, dtype, st); if (rc <= 0) return rc; }
  // each thread of the generic kernel owns at most MAXO_R of the h * d_e outputs of dW_r
  EGT_REQUIRE(p.h * p.d_e <= MAXO_R * EPB && p.d_e <= EPB, EGT_E_SHAPE,
              "edge write-back backward: h * d_e = %d exceeds the %d outputs the generic kernel accumulates", p.h * p.d_e,
              MAXO_R * EPB);
  size_t smem = (size_t)EPB * (p.h + p.d_e) * sizeof(float);
  if (smem > 48 * 1024) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_out_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_out_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  DISPATCH_T(dtype, edge_out_bwd_kernel, persistent_grid(p), EPB, smem, st, p);
  return EGT_OK;
}
int edge_proj_bwd_launch(const EdgeParams &p, int dtype, cudaStream_t st) {
  if (!staged_force_generic()) { const int rc = edge_fast_launch(3, p, dtype, st); if (rc <= 0) return rc; }
  int J = p.gated ? 2 * p.h : p.h;
  // each thread of the generic kernel owns at most MAXO_P of the d_e * J outputs of dW_E | dW_G
  EGT_REQUIRE(p.d_e * J <= MAXO_P * EPB, EGT_E_SHAPE,
              "edge projection backward: d_e * %d = %d exceeds the %d outputs the generic kernel accumulates", J, p.d_e * J,
              MAXO_P * EPB);
  size_t smem = ((size_t)EPB * (p.d_e + J) + J) * sizeof(float);
  if (smem > 48 * 1024) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_proj_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(edge_proj_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  DISPATCH_T(dtype, edge_proj_bwd_kernel, persistent_grid(p), EPB, smem, st, p);
  return EGT_OK;
}

}  // namespace egt
