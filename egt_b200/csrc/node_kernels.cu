// node_kernels.cu -- node-channel side of the block: LN_h + QKV projection, output projection +
// residual, and their backward (graph_xformer_model_base.py:107-114,136-140).
// [B*N, d] x [d, 3d] sized work: a few percent of the block's bytes and FLOPs.
#include "common.cuh"
#include "kernels.h"

namespace egt {

constexpr int LROWS = 16;   // rows per CTA in the linear kernel

template <typename T>
__device__ __forceinline__ float ld_any(const void *p, size_t i, int f32) {
  return f32 ? ((const float *)p)[i] : ldf((const T *)p + i);
}
template <typename T>
__device__ __forceinline__ void st_any(void *p, size_t i, int f32, float v) {
  if (f32) ((float *)p)[i] = v;
  else stf((T *)p + i, v);
}

// dynamic smem: xs[LROWS][din]
template <typename T>
__global__ void __launch_bounds__(256) linear_kernel(LinearArgs a) {
  extern __shared__ float xs[];
  const int r0 = blockIdx.x * LROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // stage rows (with optional LayerNorm), one warp per row
  for (int rr = warp; rr < LROWS; rr += 8) {
    int r = r0 + rr;
    if (r >= a.R) {
      for (int k = lane; k < a.din; k += 32) xs[rr * a.din + k] = 0.f;
      continue;
    }
    float s = 0.f;
    for (int k = lane; k < a.din; k += 32) {
      float v = ld_any<T>(a.x, (size_t)r * a.din + k, a.x_f32);
      xs[rr * a.din + k] = v;
      s += v;
    }
    if (a.ln_gamma) {
      s = warp_sum(s);
      float mu = s / a.din;
      float var = 0.f;
      for (int k = lane; k < a.din; k += 32) {
        float t = xs[rr * a.din + k] - mu;
        var += t * t;
      }
      var = warp_sum(var);
      float rstd = rsqrtf(var / a.din + a.ln_eps);
      for (int k = lane; k < a.din; k += 32) {
        float v = (xs[rr * a.din + k] - mu) * rstd * __ldg(a.ln_gamma + k) + __ldg(a.ln_beta + k);
        xs[rr * a.din + k] = v;
        if (a.xn_out) a.xn_out[(size_t)r * a.din + k] = v;
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < a.dout; j += 256) {
    float acc[LROWS];
    float bj = a.bias ? __ldg(a.bias + j) : 0.f;
#pragma unroll
    for (int rr = 0; rr < LROWS; ++rr) acc[rr] = bj;
    for (int k = 0; k < a.din; ++k) {
      float w = a.trans ? __ldg(a.W + (size_t)j * a.din + k) : __ldg(a.W + (size_t)k * a.dout + j);
#pragma unroll
      for (int rr = 0; rr < LROWS; ++rr) acc[rr] += xs[rr * a.din + k] * w;
    }
#pragma unroll
    for (int rr = 0; rr < LROWS; ++rr) {
      int r = r0 + rr;
      if (r < a.R) {
        float v = acc[rr];
        if (j < a.scale_cols) v *= a.scale;
        if (a.res) v += ldf((const T *)a.res + (size_t)r * a.dout + j);
        st_any<T>(a.out, (size_t)r * a.dout + j, a.out_f32, v);
      }
    }
  }
}

int linear_launch(const LinearArgs &a, int dtype, cudaStream_t st) {
  size_t smem = (size_t)LROWS * a.din * sizeof(float);
  unsigned grid = (a.R + LROWS - 1) / LROWS;
  LaunchScope _ls("linear_kernel", st);
  if (dtype == EGT_F32) linear_kernel<float><<<grid, 256, smem, st>>>(a);
  else linear_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

// ---- dW += X^T Y, db += colsum(Y) --------------------------------------------------------
constexpr int XROWS = 32;
// dynamic smem: xs[XROWS][dx], ys[XROWS][dy]
template <typename T>
__global__ void __launch_bounds__(256) xty_kernel(XtyArgs a, int rows_per_cta, int j0, int jn) {
  extern __shared__ float sm[];
  float *xs = sm, *ys = sm + XROWS * a.dx;
  const int tid = threadIdx.x;
  const int rbeg = blockIdx.x * rows_per_cta;
  const int rend = min(a.R, rbeg + rows_per_cta);
  // each thread owns outputs o = tid + t*256 (up to 64 of them, kept in registers)
  constexpr int MAXO = 64;
  float acc[MAXO];
#pragma unroll
  for (int t = 0; t < MAXO; ++t) acc[t] = 0.f;
  float bacc = 0.f;   // column sum for j = tid (+256..) handled below via the same loop
  const int nout = a.dx * jn;
  for (int r0 = rbeg; r0 < rend; r0 += XROWS) {
    int nr = min(XROWS, rend - r0);
    __syncthreads();
    for (int i = tid; i < XROWS * a.dx; i += 256) {
      int rr = i / a.dx, k = i % a.dx;
      xs[i] = rr < nr ? ld_any<T>(a.X, (size_t)(r0 + rr) * a.dx + k, a.x_f32) : 0.f;
    }
    for (int i = tid; i < XROWS * jn; i += 256) {
      int rr = i / jn, k = i % jn;
      ys[i] = rr < nr ? ld_any<T>(a.Y, (size_t)(r0 + rr) * a.dy + j0 + k, a.y_f32) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MAXO; ++t) {
      int o = tid + t * 256;
      if (o < nout) {
        int i = o / jn, j = o % jn;
        float s = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < XROWS; ++rr) s += xs[rr * a.dx + i] * ys[rr * jn + j];
        acc[t] += s;
      }
    }
    if (a.db) {
      for (int j = tid; j < jn; j += 256) {
        float s = 0.f;
        for (int rr = 0; rr < XROWS; ++rr) s += ys[rr * jn + j];
        if (j == tid) bacc += s;
        else atomicAdd(a.db + j0 + j, s);   // chunk wider than 256 columns (rare)
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAXO; ++t) {
    int o = tid + t * 256;
    if (o < nout) atomicAdd(a.dW + (size_t)(o / jn) * a.dy + j0 + (o % jn), acc[t]);
  }
  if (a.db && tid < jn) atomicAdd(a.db + j0 + tid, bacc);
}

int xty_launch(const XtyArgs &a, int dtype, cudaStream_t st) {
  if (a.dx > 16384) {
    set_error(EGT_E_SHAPE, "xty: dx=%d too wide", a.dx);
    return EGT_E_SHAPE;
  }
  int jchunk = 16384 / a.dx;              // outputs per CTA pass <= 64 per thread
  if (jchunk > a.dy) jchunk = a.dy;
  size_t smem = (size_t)XROWS * (a.dx + jchunk) * sizeof(float);
  int ctas = 148 * 2;
  int rows_per_cta = (a.R + ctas - 1) / ctas;
  rows_per_cta = ((rows_per_cta + XROWS - 1) / XROWS) * XROWS;
  unsigned grid = (a.R + rows_per_cta - 1) / rows_per_cta;
  if (smem > 48 * 1024) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  for (int j0 = 0; j0 < a.dy; j0 += jchunk) {
    int jn = a.dy - j0 < jchunk ? a.dy - j0 : jchunk;
    LaunchScope _ls("xty_kernel", st);
    if (dtype == EGT_F32) xty_kernel<float><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
    else xty_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
    EGT_CHECK_CUDA(cudaGetLastError());
  }
  return EGT_OK;
}

// ---- LayerNorm backward + residual gradient ----------------------------------------------
// one warp per row; dgamma/dbeta reduced per CTA in smem then atomically added.
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(LnBwdArgs a, int rows_per_cta) {
  extern __shared__ float sm[];   // dg[D], db[D]
  float *sdg = sm, *sdb = sm + a.D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < 2 * a.D; k += 256) sm[k] = 0.f;
  __syncthreads();
  const int rbeg = blockIdx.x * rows_per_cta;
  const int rend = min(a.R, rbeg + rows_per_cta);
  for (int r = rbeg + warp; r < rend; r += 8) {
    const T *x = (const T *)a.x + (size_t)r * a.D;
    const float *dy = a.dy + (size_t)r * a.D;
    float s = 0.f;
    for (int k = lane; k < a.D; k += 32) s += ldf(x + k);
    s = warp_sum(s);
    float mu = s / a.D;
    float var = 0.f;
    for (int k = lane; k < a.D; k += 32) {
      float t = ldf(x + k) - mu;
      var += t * t;
    }
    var = warp_sum(var);
    float rstd = rsqrtf(var / a.D + a.eps);
    float m1 = 0.f, m2 = 0.f;
    for (int k = lane; k < a.D; k += 32) {
      float xn = (ldf(x + k) - mu) * rstd;
      float g = dy[k];
      float dxh = g * __ldg(a.gamma + k);
      m1 += dxh;
      m2 += dxh * xn;
      atomicAdd(sdg + k, g * xn);
      atomicAdd(sdb + k, g);
    }
    m1 = warp_sum(m1) / a.D;
    m2 = warp_sum(m2) / a.D;
    for (int k = lane; k < a.D; k += 32) {
      float xn = (ldf(x + k) - mu) * rstd;
      float dxh = dy[k] * __ldg(a.gamma + k);
      float v = rstd * (dxh - m1 - xn * m2);
      if (a.dres) v += ldf((const T *)a.dres + (size_t)r * a.D + k);
      stf((T *)a.dx + (size_t)r * a.D + k, v);
    }
  }
  __syncthreads();
  for (int k = tid; k < a.D; k += 256) {
    atomicAdd(a.dgamma + k, sdg[k]);
    atomicAdd(a.dbeta + k, sdb[k]);
  }
}

int ln_bwd_launch(const LnBwdArgs &a, int dtype, cudaStream_t st) {
  int ctas = 148 * 2;
  int rows_per_cta = (a.R + ctas - 1) / ctas;
  if (rows_per_cta < 8) rows_per_cta = 8;
  unsigned grid = (a.R + rows_per_cta - 1) / rows_per_cta;
  size_t smem = 2 * (size_t)a.D * sizeof(float);
  LaunchScope _ls("ln_bwd_kernel", st);
  if (dtype == EGT_F32) ln_bwd_kernel<float><<<grid, 256, smem, st>>>(a, rows_per_cta);
  else ln_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, rows_per_cta);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
