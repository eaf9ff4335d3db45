// node_tc.cu -- node-channel side of the block on sm_100a tensor cores (bf16 activations, d = 64, h = 8):
//
//   node_qkv_kernel    qkv  = LN(h) W_qkv + b_qkv, Q third pre-scaled     (graph_xformer_model_base.py:107-114)
//   node_out_kernel    h'   = h + V_att W_O + b_O                         (:136-140)
//   node_bwd1_kernel   dV_att = dh' W_O^T ; dW_O += V_att^T dh' ; db_O += colsum(dh')
//   node_bwd2_kernel   dhn = dqkv W_qkv^T ; dh = LN_bwd(dhn) + dh' ; dgamma, dbeta ;
//                      dW_qkv += LN(h)^T dqkv ; db_qkv += colsum(dqkv)
//
// One CTA = 128 rows of the flattened [B*N, d] node tensor per iteration (grid-stride), thread = row =
// TMEM lane.  Row tiles are written to shared memory as 128B-swizzled tcgen05 operands; the weights
// are converted to bf16 operand images once per CTA.  The row-contracted products (X^T Y, needed for the
// weight gradients) use the tile itself as an MN-major A operand next to a tile of ones, so rows 0-63 of
// the accumulator hold dW and row 64 holds the column sum (the bias gradient); they accumulate in TMEM
// over the CTA's tiles and are flushed once with atomics.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "fused.h"
#include "umma.cuh"
#include "fused_prep.cuh"

namespace egt {
using namespace umma;

namespace {

constexpr int ND = 64;            // model width served by these kernels
constexpr uint32_t TILE = 16384;  // [128 rows x 128 B] swizzled tile

struct NodeBars { uint64_t bar; uint32_t tmem_base; uint32_t pad; uint64_t bar_tma; };

// thread t's row of 64 bf16 (8 x 16 B) -> swizzled tile
__device__ __forceinline__ void put_row(uint8_t *tile, int t, const uint4 *v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) *(uint4 *)(tile + sw128_off(t, 8 * j)) = v[j];
}
__device__ __forceinline__ void unpack8(const uint4 v, float *x) {
  x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
  x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float *x) {
  uint4 v;
  v.x = pack_bf16(x[0], x[1]); v.y = pack_bf16(x[2], x[3]); v.z = pack_bf16(x[4], x[5]); v.w = pack_bf16(x[6], x[7]);
  return v;
}
// B operand images from a float32 Keras kernel W[kdim][ndim] (row-major, "x @ W"); each step converts 8
// consecutive floats of a row of W into one 16-byte chunk of the image.
//   MN-major (n contiguous), 128B swizzle, atoms of 64 n: used for x @ W       (contraction over W's rows)
__device__ __forceinline__ void build_w_mn(uint8_t *img, const float *W, int kdim, int ndim, int tid, int nthr) {
  const int nch = ndim >> 3, total = kdim * nch;
  for (int i0 = tid; i0 < total; i0 += 4 * nthr) {      // 4 chunks per pass: 8 independent 16-byte loads in flight
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < total) {
        const int k = i / nch, n = (i % nch) << 3;
        a[u] = *(const float4 *)(W + (size_t)k * ndim + n);
        b[u] = *(const float4 *)(W + (size_t)k * ndim + n + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < total) {
        const int k = i / nch, n = (i % nch) << 3;
        const float y[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
        const uint32_t off = (uint32_t)(n >> 6) * (uint32_t)(kdim * 128) + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u +
                             ((uint32_t)((((n & 63) >> 3) ^ k) & 7) << 4);
        *(uint4 *)(img + off) = pack8(y);
      }
    }
  }
}
//   K-major image of W^T, i.e. B[n = row of W][k = column of W]: used for x @ W^T (contraction over W's columns)
__device__ __forceinline__ void build_wt_k(uint8_t *img, const float *W, int nrows, int kcols, int tid, int nthr) {
  const int kch = kcols >> 3, total = nrows * kch;
  for (int i0 = tid; i0 < total; i0 += 4 * nthr) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < total) {
        const int n = i / kch, k = (i % kch) << 3;
        a[u] = *(const float4 *)(W + (size_t)n * kcols + k);
        b[u] = *(const float4 *)(W + (size_t)n * kcols + k + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < total) {
        const int n = i / kch, k = (i % kch) << 3;
        const float y[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
        *(uint4 *)(img + (uint32_t)(k >> 6) * (uint32_t)(nrows * 128) + sw128_off(n, k & 63)) = pack8(y);
      }
    }
  }
}
// 16-byte vector reduction into global memory (sm_90+)
__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void fill_ones(uint8_t *tile, int tid, int nthr) {
  for (int i = tid; i < (int)(TILE / 16); i += nthr) ((uint4 *)tile)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
}

__device__ __forceinline__ void node_setup(NodeBars *bars, int tid, uint32_t cols) {
  if (tid < 32) {
    if (tid == 0) { mbar_init(smem_u32(&bars->bar), 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&bars->tmem_base), cols);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
struct NodeQkvArgs {
  const __nv_bfloat16 *h; const float *gamma, *beta; float eps;
  const float *W, *bias; float qscale;
  __nv_bfloat16 *qkv; int R;
  egt_block_weights_t w; float clip_lo, clip_hi; FusedPrep *prep_out;   // extra CTA: fused_prep_body (NULL = off)
};

// 256 threads: two threads per row.  LayerNorm prologue: threads 2r, 2r+1 share row r (32 channels each, partial
// sums exchanged by shuffle); epilogue: thread (lane t, half) reads TMEM lane t and converts output columns
// 96*half .. 96*half+95.
__global__ void __launch_bounds__(256) node_qkv_kernel(const NodeQkvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sA = smem, *sW = smem + TILE;                     // A 16 KB | W image 3 x 8 KB
  float *sgb = (float *)(smem + TILE + 24576);               // gamma, beta, bias(192)
  NodeBars *bars = (NodeBars *)(smem + TILE + 24576 + 1536);
  const int tid = threadIdx.x, t = tid & 127, half = tid >> 7;
  const int nwork = a.prep_out ? gridDim.x - 1 : gridDim.x;
  pdl_trigger();
  if ((int)blockIdx.x == nwork) {
    if (tid < 128) fused_prep_body(a.w, a.clip_lo, a.clip_hi, a.prep_out, tid);
    return;
  }
  node_setup(bars, tid, 256);
  build_w_mn(sW, a.W, ND, 3 * ND, tid, 256);
  if (tid < ND) { sgb[tid] = a.gamma[tid]; sgb[ND + tid] = a.beta[tid]; }
  if (tid < 3 * ND) sgb[2 * ND + tid] = a.bias[tid];
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
  constexpr uint32_t IDESC = idesc_bf16(128, 192, 0, 1);
  uint32_t phase = 0;
  const int pr = tid >> 1, ph = tid & 1;                     // prologue: row within the tile, channel half
  for (int tile = blockIdx.x; tile * 128 < a.R; tile += nwork) {
    {   // LayerNorm of row pr, channels 32 ph .. 32 ph + 31 -> bf16 A tile
      const int r = tile * 128 + pr;
      float x[32];
      const uint4 *src = (const uint4 *)(a.h + (size_t)(r < a.R ? r : 0) * ND + 32 * ph);
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack8(r < a.R ? src[j] : make_uint4(0, 0, 0, 0), x + 8 * j);
      float mu = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) mu += x[c];
      mu += __shfl_xor_sync(0xffffffffu, mu, 1);
      mu *= (1.f / 64.f);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) { const float dlt = x[c] - mu; var = fmaf(dlt, dlt, var); }
      var += __shfl_xor_sync(0xffffffffu, var, 1);
      const float rs = rsqrtf(var * (1.f / 64.f) + a.eps);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          y[c] = fmaf((x[8 * j + c] - mu) * rs, sgb[32 * ph + 8 * j + c], sgb[ND + 32 * ph + 8 * j + c]);
        *(uint4 *)(sA + sw128_off(pr, 32 * ph + 8 * j)) = pack8(y);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < 4; ++s)
        mma_ss(tmem, smem_desc(smem_u32(sA) + 32 * s, 16, 1024, LAYOUT_SW128),
               smem_desc(smem_u32(sW) + 2048 * s, ND * 128, 1024, LAYOUT_SW128), IDESC, s > 0);
      mma_commit(smem_u32(&bars->bar));
    }
    mbar_wait(smem_u32(&bars->bar), phase);
    phase ^= 1;
    tc_fence_after();
    const int r = tile * 128 + t;
#pragma unroll 1
    for (int ch = 3 * half; ch < 3 * half + 3; ++ch) {
      uint32_t o[32];
      tmem_ld32(tlane + 32 * ch, o);
      tmem_ld_wait();
      if (r < a.R) {
        const float sc = ch < 2 ? a.qscale : 1.f;
        uint4 *dst = (uint4 *)(a.qkv + (size_t)r * (3 * ND) + 32 * ch);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float y[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) y[c] = (__uint_as_float(o[8 * q + c]) + sgb[2 * ND + 32 * ch + 8 * q + c]) * sc;
          dst[q] = pack8(y);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (tid < 32) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
struct NodeOutArgs {
  const __nv_bfloat16 *v_att, *h; const float *W, *bias;
  __nv_bfloat16 *h_out; int R;
};

__global__ void __launch_bounds__(128) node_out_kernel(const NodeOutArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sA = smem, *sW = smem + TILE;                     // A 16 KB | W_O image 8 KB
  float *sb = (float *)(smem + TILE + 8192);
  NodeBars *bars = (NodeBars *)(smem + TILE + 8192 + 256);
  const int t = threadIdx.x;
  pdl_trigger();
  node_setup(bars, t, 64);
  build_w_mn(sW, a.W, ND, ND, t, 128);
  if (t < ND) sb[t] = a.bias[t];
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)((t >> 5) * 32) << 16);
  constexpr uint32_t IDESC = idesc_bf16(128, 64, 0, 1);
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile * 128 < a.R; tile += gridDim.x) {
    const int r = tile * 128 + t;
    {
      uint4 v[8];
      const uint4 *src = (const uint4 *)(a.v_att + (size_t)(r < a.R ? r : 0) * ND);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = r < a.R ? src[j] : make_uint4(0, 0, 0, 0);
      put_row(sA, t, v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < 4; ++s)
        mma_ss(tmem, smem_desc(smem_u32(sA) + 32 * s, 16, 1024, LAYOUT_SW128),
               smem_desc(smem_u32(sW) + 2048 * s, ND * 128, 1024, LAYOUT_SW128), IDESC, s > 0);
      mma_commit(smem_u32(&bars->bar));
    }
    mbar_wait(smem_u32(&bars->bar), phase);
    phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t o[32];
      tmem_ld32(tlane + 32 * ch, o);
      tmem_ld_wait();
      if (r < a.R) {
        const uint4 *res = (const uint4 *)(a.h + (size_t)r * ND + 32 * ch);
        uint4 *dst = (uint4 *)(a.h_out + (size_t)r * ND + 32 * ch);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float y[8], hr[8];
          unpack8(res[q], hr);
#pragma unroll
          for (int c = 0; c < 8; ++c) y[c] = __uint_as_float(o[8 * q + c]) + sb[32 * ch + 8 * q + c] + hr[c];
          dst[q] = pack8(y);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (t < 32) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
struct NodeBwd1Args {
  const __nv_bfloat16 *dh_out, *v_att; const float *W;       // W_O [64,64]
  __nv_bfloat16 *d_v_att; float *dW, *db; int R;
  egt_block_weights_t w; float clip_lo, clip_hi; FusedPrep *prep_out;   // extra CTA: fused_prep_body (NULL = off)
  int parts;   // bit 0: dV_att (what the fused backward waits for) ; bit 1: dW_O, db_O (can run next to it on a few SMs)
};

__global__ void __launch_bounds__(128) node_bwd1_kernel(const NodeBwd1Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sX = smem, *sOnes = smem + TILE, *sY = smem + 2 * TILE, *sW = smem + 3 * TILE;   // X | 1 | Y | W_O^T image 8 KB
  NodeBars *bars = (NodeBars *)(smem + 3 * TILE + 8192);
  const int t = threadIdx.x;
  const int nwork = a.prep_out ? gridDim.x - 1 : gridDim.x;
  pdl_trigger();
  if ((int)blockIdx.x == nwork) { fused_prep_body(a.w, a.clip_lo, a.clip_hi, a.prep_out, t); return; }
  const bool do_dx = a.parts & 1, do_dw = a.parts & 2;
  node_setup(bars, t, 128);
  if (do_dx) build_wt_k(sW, a.W, ND, ND, t, 128);
  if (do_dw) fill_ones(sOnes, t, 128);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)((t >> 5) * 32) << 16);
  constexpr uint32_t ID_MAIN = idesc_bf16(128, 64, 0, 0), ID_T = idesc_bf16(128, 64, 1, 1);
  constexpr uint32_t TM_D1 = 0, TM_D2 = 64;
  uint32_t phase = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile * 128 < a.R; tile += nwork) {
    const int r = tile * 128 + t;
    {
      uint4 v[8];
      if (do_dw) {
        const uint4 *sx = (const uint4 *)(a.v_att + (size_t)(r < a.R ? r : 0) * ND);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = r < a.R ? sx[j] : make_uint4(0, 0, 0, 0);
        put_row(sX, t, v);
      }
      const uint4 *sy = (const uint4 *)(a.dh_out + (size_t)(r < a.R ? r : 0) * ND);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = r < a.R ? sy[j] : make_uint4(0, 0, 0, 0);
      put_row(sY, t, v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      if (do_dx) {
#pragma unroll
        for (int s = 0; s < 4; ++s)      // dV_att = dh' W_O^T
          mma_ss(tmem + TM_D1, smem_desc(smem_u32(sY) + 32 * s, 16, 1024, LAYOUT_SW128),
                 smem_desc(smem_u32(sW) + 32 * s, 16, 1024, LAYOUT_SW128), ID_MAIN, s > 0);
      }
      if (do_dw) {
#pragma unroll
        for (int s = 0; s < 8; ++s)      // [V_att | 1]^T dh'
          mma_ss(tmem + TM_D2, smem_desc(smem_u32(sX) + 2048 * s, TILE, 1024, LAYOUT_SW128),
                 smem_desc(smem_u32(sY) + 2048 * s, TILE, 1024, LAYOUT_SW128), ID_T, !(first && s == 0));
      }
      mma_commit(smem_u32(&bars->bar));
    }
    first = false;
    mbar_wait(smem_u32(&bars->bar), phase);
    phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < (do_dx ? 2 : 0); ++ch) {
      uint32_t o[32];
      tmem_ld32(tlane + TM_D1 + 32 * ch, o);
      tmem_ld_wait();
      if (r < a.R) {
        uint4 *dst = (uint4 *)(a.d_v_att + (size_t)r * ND + 32 * ch);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float y[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) y[c] = __uint_as_float(o[8 * q + c]);
          dst[q] = pack8(y);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (!first && do_dw) {   // flush dW_O (rows 0-63) and db_O (row 64)
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t o[32];
      tmem_ld32(tlane + TM_D2 + 32 * ch, o);
      tmem_ld_wait();
      if (t <= ND) {
        float *dst = t < ND ? a.dW + (size_t)t * ND + 32 * ch : a.db + 32 * ch;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          red_add_v4(dst + c, __uint_as_float(o[c]), __uint_as_float(o[c + 1]), __uint_as_float(o[c + 2]), __uint_as_float(o[c + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (t < 32) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
struct NodeBwd2Args {
  const __nv_bfloat16 *h, *dh_out; const float *dqkv;       // dqkv [R,192] float32
  int dqkv_is_bf16;                                          // the tiles come by TMA from a bf16 [R,192] tensor instead
  const float *gamma, *beta; float eps;
  const float *W;                                            // W_qkv [64,192]
  __nv_bfloat16 *dh; float *dW, *db, *dgamma, *dbeta; int R;
  const float *partials; int nparts; egt_block_weights_t w; egt_block_grads_t g;   // extra CTA: finalize (NULL = off)
  int parts;   // bit 0: dh (+ dgamma, dbeta) ; bit 1: dW_qkv, db_qkv.  Launched as two kernels, the halves share an SM.
};

// shared-memory map; the two halves of a split launch only ask for what they use so that their CTAs fit one SM together
__host__ __device__ inline int node_bwd2_smem(int parts) {
  int b = 3 * (int)TILE;                                   // dqkv tiles
  if (parts & 2) b += 2 * (int)TILE;                       // LN(h) | 1
  if (parts & 1) b += 24576 + 128 * 65 * 4;                // W_qkv^T image | transpose scratch
  return b + 2 * ND * 4 + 64;
}

// 256 threads: threads 0-127 own the rows (= TMEM lanes); threads 128-255 only help with the staging (weight image,
// float32 -> bf16 tiles of dqkv), which is 40 % of the kernel's latency chain with 128 threads.
__global__ void __launch_bounds__(256) node_bwd2_kernel(const __grid_constant__ CUtensorMap tm_y, const NodeBwd2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool do_dx = a.parts & 1, do_dw = a.parts & 2;
  uint8_t *p = smem;
  uint8_t *sX = p, *sOnes = p + TILE;                                 // LN(h) | 1   (adjacent: one MN-major operand)
  if (do_dw) p += 2 * TILE;
  uint8_t *sY = p;                                                    // dqkv (3 tiles)
  p += 3 * TILE;
  uint8_t *sW = p;                                                    // W_qkv^T image: 3 k-atoms x 8 KB
  float *sred = (float *)(p + 24576);                                 // [128][65] float32 transpose scratch
  if (do_dx) p += 24576 + 128 * 65 * 4;
  float *sgb = (float *)p;                                            // gamma, beta
  NodeBars *bars = (NodeBars *)(sgb + 2 * ND);
  const int tid = threadIdx.x, t = tid & 127;
  const bool rowthr = tid < 128;
  const int nwork = a.partials ? gridDim.x - 1 : gridDim.x;
  pdl_trigger();
  if ((int)blockIdx.x == nwork) { pdl_wait(); fused_bwd_finalize_body(a.partials, a.nparts, a.w, a.g, tid, 256); return; }
  node_setup(bars, tid, do_dw ? 256 : 64);
  if (tid == 0 && a.dqkv_is_bf16) { mbar_init(smem_u32(&bars->bar_tma), 1); mbar_fence_init(); tma_prefetch_desc(&tm_y); }
  if (do_dx) build_wt_k(sW, a.W, ND, 3 * ND, tid, 256);
  if (do_dw) fill_ones(sOnes, tid, 256);
  if (tid < ND) { sgb[tid] = a.gamma[tid]; sgb[ND + tid] = a.beta[tid]; }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tlane = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
  constexpr uint32_t ID_MAIN = idesc_bf16(128, 64, 0, 0), ID_T = idesc_bf16(128, 192, 1, 1);
  constexpr uint32_t TM_D1 = 0, TM_D2 = 64;
  uint32_t phase = 0;
  bool first = true;
  float dg_acc = 0.f, db_acc = 0.f;     // partial dgamma / dbeta of column tid & 63 over the row quarter tid >> 6
  for (int tile = blockIdx.x; tile * 128 < a.R; tile += nwork) {
    const int r = tile * 128 + t;
    const bool valid = r < a.R;
    float xh[64], rs = 0.f;
    if (rowthr) {   // x^ = (h - mu) rstd ; LN(h) -> X tile
      const uint4 *src = (const uint4 *)(a.h + (size_t)(valid ? r : 0) * ND);
#pragma unroll
      for (int j = 0; j < 8; ++j) unpack8(valid ? src[j] : make_uint4(0, 0, 0, 0), xh + 8 * j);
      float mu = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) mu += xh[c];
      mu *= (1.f / 64.f);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) { xh[c] -= mu; var = fmaf(xh[c], xh[c], var); }
      rs = rsqrtf(var * (1.f / 64.f) + a.eps);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          xh[8 * j + c] *= rs;
          y[c] = valid ? fmaf(xh[8 * j + c], sgb[8 * j + c], sgb[ND + 8 * j + c]) : 0.f;
        }
        if (do_dw) *(uint4 *)(sX + sw128_off(t, 8 * j)) = pack8(y);
      }
    }
    if (a.dqkv_is_bf16) {   // dqkv tile: three [128 x 64] bf16 boxes by TMA (rows beyond R are zero-filled)
      if (tid == 0) {
        const uint32_t bar = smem_u32(&bars->bar_tma);
        mbar_expect_tx(bar, 3 * TILE);
        for (int j = 0; j < 3; ++j) tma_load_3d(smem_u32(sY) + j * TILE, &tm_y, bar, 64 * j, tile * 128, 0);
      }
    } else {   // dqkv tile (float32 [128,192], rows contiguous) -> three bf16 tiles; coalesced: 24 threads per row
      const float *base = a.dqkv + (size_t)tile * 128 * (3 * ND);
      const int rows_here = a.R - tile * 128 < 128 ? a.R - tile * 128 : 128;
#pragma unroll 4
      for (int i = tid; i < 128 * 24; i += 256) {
        const int rr = i / 24, j = i % 24;
        float y[8];
        if (rr < rows_here) {
          const float4 p0 = *(const float4 *)(base + (size_t)rr * (3 * ND) + 8 * j), p1 = *(const float4 *)(base + (size_t)rr * (3 * ND) + 8 * j + 4);
          y[0] = p0.x; y[1] = p0.y; y[2] = p0.z; y[3] = p0.w; y[4] = p1.x; y[5] = p1.y; y[6] = p1.z; y[7] = p1.w;
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) y[c] = 0.f;
        }
        *(uint4 *)(sY + (uint32_t)(j >> 3) * TILE + sw128_off(rr, 8 * (j & 7))) = pack8(y);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      if (a.dqkv_is_bf16) mbar_wait(smem_u32(&bars->bar_tma), phase);     // same parity as the MMA barrier: one use per tile
      tc_fence_after();
      if (do_dx) {
#pragma unroll
        for (int s = 0; s < 12; ++s)     // dhn = dqkv W_qkv^T   (K = 192)
          mma_ss(tmem + TM_D1, smem_desc(smem_u32(sY) + (uint32_t)(s >> 2) * TILE + 32 * (s & 3), 16, 1024, LAYOUT_SW128),
                 smem_desc(smem_u32(sW) + (uint32_t)(s >> 2) * 8192 + 32 * (s & 3), 16, 1024, LAYOUT_SW128), ID_MAIN, s > 0);
      }
      if (do_dw) {
#pragma unroll
        for (int s = 0; s < 8; ++s)      // [LN(h) | 1]^T dqkv
          mma_ss(tmem + TM_D2, smem_desc(smem_u32(sX) + 2048 * s, TILE, 1024, LAYOUT_SW128),
                 smem_desc(smem_u32(sY) + 2048 * s, TILE, 1024, LAYOUT_SW128), ID_T, !(first && s == 0));
      }
      mma_commit(smem_u32(&bars->bar));
    }
    first = false;
    mbar_wait(smem_u32(&bars->bar), phase);
    phase ^= 1;
    tc_fence_after();
    if (do_dx) {   // LayerNorm backward + residual for this thread's row
      float dy[64];
      if (rowthr) {
        uint32_t o[32];
        tmem_ld32(tlane + TM_D1, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) dy[c] = __uint_as_float(o[c]);
        tmem_ld32(tlane + TM_D1 + 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) dy[32 + c] = __uint_as_float(o[c]);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const float dxh = dy[c] * sgb[c];
          m1 += dxh;
          m2 = fmaf(dxh, xh[c], m2);
          sred[t * 65 + c] = dy[c] * xh[c];          // -> dgamma (column sums below)
        }
        m1 *= (1.f / 64.f); m2 *= (1.f / 64.f);
        if (valid) {
          const uint4 *res = (const uint4 *)(a.dh_out + (size_t)r * ND);
          uint4 *dst = (uint4 *)(a.dh + (size_t)r * ND);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float y[8], hr[8];
            unpack8(res[j], hr);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              y[c] = fmaf(rs, fmaf(dy[8 * j + c], sgb[8 * j + c], -fmaf(xh[8 * j + c], m2, m1)), hr[c]);
            dst[j] = pack8(y);
          }
        }
      }
      // column sums over the tile's rows: thread (column tid & 63, row quarter tid >> 6), summed over tiles in registers
      const int cc = tid & 63, r0 = (tid >> 6) * 32;
      __syncthreads();
      {
        float s = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) s += sred[(r0 + rr) * 65 + cc];
        dg_acc += s;
      }
      __syncthreads();
      if (rowthr) {
#pragma unroll
        for (int c = 0; c < 64; ++c) sred[t * 65 + c] = dy[c];   // -> dbeta
      }
      __syncthreads();
      {
        float s = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) s += sred[(r0 + rr) * 65 + cc];
        db_acc += s;
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (!first && do_dx) {
    atomicAdd(a.dgamma + (tid & 63), dg_acc);
    atomicAdd(a.dbeta + (tid & 63), db_acc);
  }
  if (!first && do_dw && rowthr) {
#pragma unroll 1
    for (int ch = 0; ch < 6; ++ch) {   // dW_qkv (rows 0-63), db_qkv (row 64)
      uint32_t o[32];
      tmem_ld32(tlane + TM_D2 + 32 * ch, o);
      tmem_ld_wait();
      if (t <= ND) {
        float *dst = t < ND ? a.dW + (size_t)t * (3 * ND) + 32 * ch : a.db + 32 * ch;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          red_add_v4(dst + c, __uint_as_float(o[c]), __uint_as_float(o[c + 1]), __uint_as_float(o[c + 2]), __uint_as_float(o[c + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, do_dw ? 256 : 64);
}

// ------------------------------------------------------------------------------------------------
static int node_grid(int R) {
  int tiles = (R + 127) / 128;
  static const int cap = getenv("EGT_NODE_GRID") ? atoi(getenv("EGT_NODE_GRID")) : 148;
  return tiles < cap ? tiles : cap;
}
template <typename K>
static int set_smem(K kernel, int bytes) {
  EGT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return EGT_OK;
}

int node_qkv_launch(const void *h, const float *gamma, const float *beta, float eps, const float *W, const float *bias,
                    float qscale, void *qkv, int R, const egt_block_weights_t *w, float clip_lo, float clip_hi,
                    FusedPrep *prep_out, cudaStream_t st) {
  NodeQkvArgs a{(const __nv_bfloat16 *)h, gamma, beta, eps, W, bias, qscale, (__nv_bfloat16 *)qkv, R, *w, clip_lo, clip_hi, prep_out};
  const int smem = TILE + 24576 + 1536 + 64 + 1024;
  static bool once = false;
  if (!once) { int rc = set_smem(node_qkv_kernel, smem); if (rc) return rc; once = true; }
  LaunchScope _ls("node_qkv_kernel", st);
  EGT_CHECK_CUDA(launch_pdl(node_qkv_kernel, dim3(node_grid(R) + (prep_out ? 1 : 0)), dim3(256), smem, st, a));
  return EGT_OK;
}

int node_out_launch(const void *v_att, const void *h, const float *W, const float *bias, void *h_out, int R,
                    cudaStream_t st) {
  NodeOutArgs a{(const __nv_bfloat16 *)v_att, (const __nv_bfloat16 *)h, W, bias, (__nv_bfloat16 *)h_out, R};
  const int smem = TILE + 8192 + 256 + 64 + 1024;
  LaunchScope _ls("node_out_kernel", st);
  EGT_CHECK_CUDA(launch_pdl(node_out_kernel, dim3(node_grid(R)), dim3(128), smem, st, a));
  return EGT_OK;
}

// side != st: the weight-gradient half (dW_O, db_O) runs as a second launch of at most 20 CTAs on `side` -- next to the
// dV_att half and then next to the fused backward kernel, on the SMs its 128-CTA grid leaves idle.  The caller forks
// `side` from `st` before the call and joins it before it returns.
int node_bwd1_launch(const void *dh_out, const void *v_att, const float *W, void *d_v_att, float *dW, float *db, int R,
                     const egt_block_weights_t *w, float clip_lo, float clip_hi, FusedPrep *prep_out, cudaStream_t st,
                     cudaStream_t side) {
  NodeBwd1Args a{(const __nv_bfloat16 *)dh_out, (const __nv_bfloat16 *)v_att, W, (__nv_bfloat16 *)d_v_att, dW, db, R, *w, clip_lo, clip_hi, prep_out, 3};
  const int smem = 3 * TILE + 8192 + 64 + 1024;
  static bool once = false;
  if (!once) { int rc = set_smem(node_bwd1_kernel, smem); if (rc) return rc; once = true; }
  if (side != st) {
    NodeBwd1Args b = a;
    b.parts = 2; b.prep_out = nullptr;
    const int tiles = node_grid(R);
    {
      LaunchScope _ls("node_bwd1w_kernel", side);
      EGT_CHECK_CUDA(launch_pdl(node_bwd1_kernel, dim3(tiles < 20 ? tiles : 20), dim3(128), smem, side, b));
    }
    a.parts = 1;
  }
  LaunchScope _ls("node_bwd1_kernel", st);
  EGT_CHECK_CUDA(launch_pdl(node_bwd1_kernel, dim3(node_grid(R) + (prep_out ? 1 : 0)), dim3(128), smem, st, a));
  return EGT_OK;
}

// side != st: two launches, the dh half on st and the weight-gradient half (+ the finalize CTA) on `side`; their CTAs
// share the SMs (106 + 80 KB of shared memory, 64 + 256 tensor-memory columns), so the two latency chains overlap.
int node_bwd2_launch(const void *h, const void *dh_out, const float *dqkv, const void *dqkv_bf, const float *gamma, const float *beta,
                     float eps, const float *W, void *dh, float *dW, float *db, float *dgamma, float *dbeta, int R,
                     const float *partials, int nparts, const egt_block_weights_t *w, const egt_block_grads_t *g,
                     cudaStream_t st, cudaStream_t side) {
  NodeBwd2Args a{(const __nv_bfloat16 *)h, (const __nv_bfloat16 *)dh_out, dqkv, dqkv_bf ? 1 : 0, gamma, beta, eps, W,
                 (__nv_bfloat16 *)dh, dW, db, dgamma, dbeta, R, partials, nparts, *w, *g, 3};
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (dqkv_bf) {   // [R, 192] bf16, boxes of [128 rows x 64 channels], 128B swizzle
    int rc = encode_tmap_3d(&tm, dqkv_bf, 3 * ND, (uint64_t)R, 1, 3 * ND * 2, (uint64_t)R * 3 * ND * 2, 64, 128, 1, 1);
    if (rc) return rc;
  }
  static bool once = false;
  if (!once) { int rc = set_smem(node_bwd2_kernel, node_bwd2_smem(3) + 1024); if (rc) return rc; once = true; }
  if (side != st) {
    NodeBwd2Args b = a;
    b.parts = 2;
    {
      LaunchScope _ls("node_bwd2w_kernel", side);
      EGT_CHECK_CUDA(launch_pdl(node_bwd2_kernel, dim3(node_grid(R) + (partials ? 1 : 0)), dim3(256), node_bwd2_smem(2) + 1024, side, tm, b));
    }
    a.parts = 1; a.partials = nullptr;
    LaunchScope _ls("node_bwd2_kernel", st);
    EGT_CHECK_CUDA(launch_pdl(node_bwd2_kernel, dim3(node_grid(R)), dim3(256), node_bwd2_smem(1) + 1024, st, tm, a));
    return EGT_OK;
  }
  LaunchScope _ls("node_bwd2_kernel", st);
  EGT_CHECK_CUDA(launch_pdl(node_bwd2_kernel, dim3(node_grid(R) + (partials ? 1 : 0)), dim3(256), node_bwd2_smem(3) + 1024, st, tm, a));
  return EGT_OK;
}

}  // namespace egt
