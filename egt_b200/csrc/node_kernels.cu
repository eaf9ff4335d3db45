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

// Register-tiled version for dout % 4 == 0: CTA = 32 rows, thread = 4 rows x 4 columns of a 128-column pass.
// Rows are staged (with the optional LayerNorm) as xs[k][row] so the 4 rows of a thread are one float4; the
// weights go through shared memory 32 k at a time as ws[k][col] (coalesced global reads for both layouts of W),
// so the inner loop is 2 shared float4 loads per 16 FMA.
constexpr int TROWS = 32, TKB = 32, TCOLS = 128, TXS = TROWS + 4;   // TXS: padded row stride of xs
// dynamic smem: xs[din][TXS] | ws[TKB][TCOLS + 4]
template <typename T>
__global__ void __launch_bounds__(256) linear_tiled_kernel(LinearArgs a) {
  extern __shared__ __align__(16) float sm_lin[];
  float *xs = sm_lin, *ws = sm_lin + (size_t)a.din * TXS;
  const int r0 = blockIdx.x * TROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int rr = warp; rr < TROWS; rr += 8) {       // one warp per row
    const int r = r0 + rr;
    if (r >= a.R) {
      for (int k = lane; k < a.din; k += 32) xs[k * TXS + rr] = 0.f;
      continue;
    }
    float s = 0.f;
    for (int k = lane; k < a.din; k += 32) {
      const float v = ld_any<T>(a.x, (size_t)r * a.din + k, a.x_f32);
      xs[k * TXS + rr] = v;
      s += v;
    }
    if (a.ln_gamma) {
      s = warp_sum(s);
      const float mu = s / a.din;
      float var = 0.f;
      for (int k = lane; k < a.din; k += 32) {
        const float t = xs[k * TXS + rr] - mu;
        var += t * t;
      }
      var = warp_sum(var);
      const float rstd = rsqrtf(var / a.din + a.ln_eps);
      for (int k = lane; k < a.din; k += 32) {
        const float v = (xs[k * TXS + rr] - mu) * rstd * __ldg(a.ln_gamma + k) + __ldg(a.ln_beta + k);
        xs[k * TXS + rr] = v;
        if (a.xn_out) a.xn_out[(size_t)r * a.din + k] = v;
      }
    }
  }
  const int cg = tid & 31, rg = tid >> 5;          // columns 4cg.. of the pass, rows 4rg..
  for (int j0 = 0; j0 < a.dout; j0 += TCOLS) {
    const int jn = a.dout - j0 < TCOLS ? a.dout - j0 : TCOLS;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.din; k0 += TKB) {
      const int kn = a.din - k0 < TKB ? a.din - k0 : TKB;
      __syncthreads();                             // xs staged / previous ws block consumed
      if (a.trans) {                               // W[j][k]: consecutive threads walk k
        for (int i = tid; i < TCOLS * TKB; i += 256) {
          const int j = i / TKB, k = i % TKB;
          ws[k * (TCOLS + 4) + j] = (j < jn && k < kn) ? __ldg(a.W + (size_t)(j0 + j) * a.din + k0 + k) : 0.f;
        }
      } else {                                     // W[k][j]: consecutive threads walk j
        for (int i = tid; i < TCOLS * TKB; i += 256) {
          const int k = i / TCOLS, j = i % TCOLS;
          ws[k * (TCOLS + 4) + j] = (j < jn && k < kn) ? __ldg(a.W + (size_t)(k0 + k) * a.dout + j0 + j) : 0.f;
        }
      }
      __syncthreads();
      if (4 * cg < jn) {
#pragma unroll 8
        for (int k = 0; k < kn; ++k) {
          const float4 xv = *(const float4 *)(xs + (size_t)(k0 + k) * TXS + 4 * rg);
          const float4 wv = *(const float4 *)(ws + k * (TCOLS + 4) + 4 * cg);
          const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], wa[j], acc[i][j]);
        }
      }
    }
    if (4 * cg < jn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * rg + i;
        if (r >= a.R) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = j0 + 4 * cg + j;
          float v = acc[i][j] + (a.bias ? __ldg(a.bias + col) : 0.f);
          if (col < a.scale_cols) v *= a.scale;
          if (a.res) v += ldf((const T *)a.res + (size_t)r * a.dout + col);
          st_any<T>(a.out, (size_t)r * a.dout + col, a.out_f32, v);
        }
      }
    }
  }
}

int linear_launch(const LinearArgs &a, int dtype, cudaStream_t st) {
  LaunchScope _ls("linear_kernel", st);
  const size_t tsmem = ((size_t)a.din * TXS + (size_t)TKB * (TCOLS + 4)) * sizeof(float);
  if (a.dout % 4 == 0 && tsmem <= 200 * 1024 && !staged_force_generic()) {
    if (tsmem > 48 * 1024) {
      EGT_CHECK_CUDA(cudaFuncSetAttribute(linear_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
      EGT_CHECK_CUDA(cudaFuncSetAttribute(linear_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
    }
    const unsigned grid = (a.R + TROWS - 1) / TROWS;
    if (dtype == EGT_F32) linear_tiled_kernel<float><<<grid, 256, tsmem, st>>>(a);
    else linear_tiled_kernel<__nv_bfloat16><<<grid, 256, tsmem, st>>>(a);
    EGT_CHECK_CUDA(cudaGetLastError());
    return EGT_OK;
  }
  size_t smem = (size_t)LROWS * a.din * sizeof(float);
  unsigned grid = (a.R + LROWS - 1) / LROWS;
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

// Register-tiled version for dx % 4 == 0 and chunk widths % 4 == 0: a thread owns up to four 4x4 output tiles,
// 2 shared float4 loads per 16 FMA.
template <typename T>
__global__ void __launch_bounds__(256) xty_tiled_kernel(XtyArgs a, int rows_per_cta, int j0, int jn) {
  extern __shared__ __align__(16) float sm_xty[];
  float *xs = sm_xty, *ys = sm_xty + XROWS * a.dx;
  const int tid = threadIdx.x;
  const int rbeg = blockIdx.x * rows_per_cta;
  const int rend = min(a.R, rbeg + rows_per_cta);
  constexpr int MAXT = 4;
  const int tj = jn / 4, ntiles = (a.dx / 4) * tj;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(a.dW) & 15) == 0;   // dy, j0 and the tile origin are multiples of 4
  float acc[MAXT][16];
#pragma unroll
  for (int t = 0; t < MAXT; ++t)
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[t][q] = 0.f;
  float bacc = 0.f;
  for (int r0 = rbeg; r0 < rend; r0 += XROWS) {
    const int nr = min(XROWS, rend - r0);
    __syncthreads();
    for (int i = tid; i < XROWS * a.dx; i += 256) {
      const int rr = i / a.dx, k = i % a.dx;
      xs[i] = rr < nr ? ld_any<T>(a.X, (size_t)(r0 + rr) * a.dx + k, a.x_f32) : 0.f;
    }
    for (int i = tid; i < XROWS * jn; i += 256) {
      const int rr = i / jn, k = i % jn;
      ys[i] = rr < nr ? ld_any<T>(a.Y, (size_t)(r0 + rr) * a.dy + j0 + k, a.y_f32) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      const int tile = tid + t * 256;
      if (tile < ntiles) {
        const float *xp = xs + 4 * (tile / tj), *yp = ys + 4 * (tile % tj);
#pragma unroll 4
        for (int rr = 0; rr < XROWS; ++rr) {
          const float4 xv = *(const float4 *)(xp + rr * a.dx);
          const float4 yv = *(const float4 *)(yp + rr * jn);
          const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][4 * i + j] = fmaf(xa[i], ya[j], acc[t][4 * i + j]);
        }
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
  for (int t = 0; t < MAXT; ++t) {
    const int tile = tid + t * 256;
    if (tile < ntiles) {
      const int i0 = 4 * (tile / tj), jj = 4 * (tile % tj);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float *dst = a.dW + (size_t)(i0 + i) * a.dy + j0 + jj;
        if (vec_ok) {                          // one 16-byte vector reduction per tile row (sm_90+)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc[t][4 * i]), "f"(acc[t][4 * i + 1]),
                       "f"(acc[t][4 * i + 2]), "f"(acc[t][4 * i + 3]) : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(dst + j, acc[t][4 * i + j]);
        }
      }
    }
  }
  if (a.db && tid < jn) atomicAdd(a.db + j0 + tid, bacc);
}

int xty_launch(const XtyArgs &a, int dtype, cudaStream_t st) {
  if (a.dx > 16384) {
    set_error(EGT_E_SHAPE, "xty: dx=%d too wide", a.dx);
    return EGT_E_SHAPE;
  }
  const bool tiled = a.dx % 4 == 0 && a.dy % 4 == 0 && a.dx <= 4096 && !staged_force_generic();
  int jchunk = 16384 / a.dx;              // outputs per CTA pass <= 64 per thread
  if (tiled) jchunk &= ~3;
  if (jchunk > a.dy) jchunk = a.dy;
  size_t smem = (size_t)XROWS * (a.dx + jchunk) * sizeof(float);
  int ctas = 148 * 2;                     // (fewer, longer CTAs measured slower: the staging loop, not the atomics, dominates)
  int rows_per_cta = (a.R + ctas - 1) / ctas;
  rows_per_cta = ((rows_per_cta + XROWS - 1) / XROWS) * XROWS;
  unsigned grid = (a.R + rows_per_cta - 1) / rows_per_cta;
  if (smem > 48 * 1024) {
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EGT_CHECK_CUDA(cudaFuncSetAttribute(xty_tiled_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  for (int j0 = 0; j0 < a.dy; j0 += jchunk) {
    int jn = a.dy - j0 < jchunk ? a.dy - j0 : jchunk;
    LaunchScope _ls("xty_kernel", st);
    if (tiled) {
      if (dtype == EGT_F32) xty_tiled_kernel<float><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
      else xty_tiled_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
    } else {
      if (dtype == EGT_F32) xty_kernel<float><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
      else xty_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(a, rows_per_cta, j0, jn);
    }
    EGT_CHECK_CUDA(cudaGetLastError());
  }
  return EGT_OK;
}

// ---- LayerNorm backward + residual gradient ----------------------------------------------
// One warp per row, lane l owns channels l, l + 32, ... (NPL of them, in registers); dgamma / dbeta are summed per lane
// over the warp's rows in registers, then once per CTA through shared memory and once per CTA into global memory.
template <typename T, int NPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(LnBwdArgs a, int rows_per_cta) {
  extern __shared__ float sm[];   // dg[D], db[D]
  float *sdg = sm, *sdb = sm + a.D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < 2 * a.D; k += 256) sm[k] = 0.f;
  __syncthreads();
  const int rbeg = blockIdx.x * rows_per_cta;
  const int rend = min(a.R, rbeg + rows_per_cta);
  const float inv_d = 1.f / a.D;
  float gam[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int k = lane + 32 * i;
    gam[i] = k < a.D ? __ldg(a.gamma + k) : 0.f;
    dg[i] = 0.f; db[i] = 0.f;
  }
  for (int r = rbeg + warp; r < rend; r += 8) {
    const T *x = (const T *)a.x + (size_t)r * a.D;
    const float *dy = a.dy + (size_t)r * a.D;
    const T *res = a.dres ? (const T *)a.dres + (size_t)r * a.D : nullptr;
    float xv[NPL], g[NPL], rv[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int k = lane + 32 * i;
      const bool ok = k < a.D;
      xv[i] = ok ? ldf(x + k) : 0.f;
      g[i] = ok ? dy[k] : 0.f;
      rv[i] = ok && res ? ldf(res + k) : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) s += xv[i];
    const float mu = warp_sum(s) * inv_d;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      xv[i] = lane + 32 * i < a.D ? xv[i] - mu : 0.f;
      var = fmaf(xv[i], xv[i], var);
    }
    const float rstd = rsqrtf(warp_sum(var) * inv_d + a.eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      xv[i] *= rstd;                                   // x^
      const float dxh = g[i] * gam[i];
      m1 += dxh;
      m2 = fmaf(dxh, xv[i], m2);
      dg[i] = fmaf(g[i], xv[i], dg[i]);
      db[i] += g[i];
    }
    m1 = warp_sum(m1) * inv_d;
    m2 = warp_sum(m2) * inv_d;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const int k = lane + 32 * i;
      if (k < a.D) stf((T *)a.dx + (size_t)r * a.D + k, fmaf(rstd, g[i] * gam[i] - m1 - xv[i] * m2, rv[i]));
    }
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int k = lane + 32 * i;
    if (k < a.D) { atomicAdd(sdg + k, dg[i]); atomicAdd(sdb + k, db[i]); }
  }
  __syncthreads();
  for (int k = tid; k < a.D; k += 256) {
    atomicAdd(a.dgamma + k, sdg[k]);
    atomicAdd(a.dbeta + k, sdb[k]);
  }
}

template <int NPL>
static void ln_bwd_launch_t(const LnBwdArgs &a, int dtype, unsigned grid, size_t smem, int rows_per_cta, cudaStream_t st) {
  if (dtype == EGT_F32) ln_bwd_kernel<float, NPL><<<grid, 256, smem, st>>>(a, rows_per_cta);
  else ln_bwd_kernel<__nv_bfloat16, NPL><<<grid, 256, smem, st>>>(a, rows_per_cta);
}

int ln_bwd_launch(const LnBwdArgs &a, int dtype, cudaStream_t st) {
  EGT_REQUIRE(a.D > 0 && a.D <= 512, EGT_E_SHAPE, "ln_bwd: width %d not in 1..512", a.D);
  int ctas = 148 * 4;
  int rows_per_cta = (a.R + ctas - 1) / ctas;
  if (rows_per_cta < 8) rows_per_cta = 8;
  unsigned grid = (a.R + rows_per_cta - 1) / rows_per_cta;
  size_t smem = 2 * (size_t)a.D * sizeof(float);
  LaunchScope _ls("ln_bwd_kernel", st);
  const int npl = (a.D + 31) / 32;
  if (npl <= 1) ln_bwd_launch_t<1>(a, dtype, grid, smem, rows_per_cta, st);
  else if (npl <= 2) ln_bwd_launch_t<2>(a, dtype, grid, smem, rows_per_cta, st);
  else if (npl <= 3) ln_bwd_launch_t<3>(a, dtype, grid, smem, rows_per_cta, st);
  else if (npl <= 4) ln_bwd_launch_t<4>(a, dtype, grid, smem, rows_per_cta, st);
  else if (npl <= 8) ln_bwd_launch_t<8>(a, dtype, grid, smem, rows_per_cta, st);
  else ln_bwd_launch_t<16>(a, dtype, grid, smem, rows_per_cta, st);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
