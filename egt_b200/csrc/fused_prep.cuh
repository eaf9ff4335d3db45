// fused_prep.cuh -- device body of the per-call derived-weight preparation of the fused tcgen05 path.
// It runs either as its own one-CTA kernel (fused_prep_kernel) or as an extra CTA of the node-side
// kernel that precedes the fused kernel on the stream (node_tc.cu), which saves a launch per pass.
#pragma once
#include "common.cuh"
#include "fused.h"

namespace egt {

// Threads 0..127 of one CTA.  B-operand images are K-major without swizzle: element (n, k) of an [N x K]
// matrix lives at  (k/8) * (N*16) + n*16 + (k%8)*2  bytes (8x16-byte core matrices, LBO = N*16, SBO = 128).
// barrier among the 128 threads (warps 0-3) that run fused_prep_body; the hosting CTA may be larger
__device__ __forceinline__ void prep_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void fused_prep_body(const egt_block_weights_t &w, float clip_lo, float clip_hi,
                                                FusedPrep *out, const int tid) {
  __shared__ float wp[2][FDE][FH];   // W' rounded to bf16
  __shared__ float sW[2][FDE][FH], sWr[FH][FDE], sgam[FDE], sbet[FDE], sbias[2][FH], sbr[FDE];
  {   // every global load of the prep is issued here, at once
    int eg = tid / 64, c = (tid / 8) % 8, hh = tid % 8;
    sW[eg][c][hh] = (eg ? w.attention_gates_kernel : w.dense_edge_b_kernel)[c * FH + hh];
    if (tid < 64) sWr[tid / FDE][tid % FDE] = w.dense_edge_r_kernel[tid];
    else if (tid < 72) sgam[tid - 64] = w.norm_edge_gamma[tid - 64];
    else if (tid < 80) sbet[tid - 72] = w.norm_edge_beta[tid - 72];
    else if (tid < 96) sbias[(tid - 80) / FH][(tid - 80) % FH] = ((tid - 80) / FH ? w.attention_gates_bias : w.dense_edge_b_bias)[(tid - 80) % FH];
    else if (tid < 104) sbr[tid - 96] = w.dense_edge_r_bias[tid - 96];
  }
  pdl_wait();                        // parameters were read above; everything below writes global memory
  prep_sync();
  if (tid < FDE) out->br[tid] = sbr[tid];
  __shared__ float wlo[2][FDE][FH], sbound;
  {   // bound of the logits from the full-precision W'_E (decides whether the lo half is needed)
    if (tid < 8) {
      const int hh = tid;
      float n2 = 0.f, v = sbias[0][hh];
      for (int c = 0; c < FDE; ++c) { const float f = sgam[c] * sW[0][c][hh]; n2 += f * f; v += sbet[c] * sW[0][c][hh]; }
      // |LN(e)_c| has l2 norm <= sqrt(d_e)  =>  |E| <= sqrt(d_e) * ||W'[:,hh]|| + |v|
      float bnd = fmaxf(fabsf(clip_lo), fabsf(clip_hi)) + sqrtf((float)FDE * n2) + fabsf(v);
      for (int o = 4; o > 0; o >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffu, bnd, o));
      if (hh == 0) { sbound = bnd; out->bound = bnd; out->use_lo = bnd > kLoThreshold; }
    }
  }
  prep_sync();
  const bool use_lo = sbound > kLoThreshold;
  {
    int eg = tid / 64, c = (tid / 8) % 8, hh = tid % 8;
    const float full = sgam[c] * sW[eg][c][hh];
    const float hi = __bfloat162float(__float2bfloat16_rn(full));
    const float lo = use_lo ? __bfloat162float(__float2bfloat16_rn(full - hi)) : 0.f;
    wp[eg][c][hh] = hi + lo;                           // what the tensor core multiplies by
    wlo[eg][c][hh] = lo;
    out->wp[eg][c][hh] = hi + lo;
  }
  prep_sync();
  if (tid < 16) {
    int eg = tid / 8, hh = tid % 8;
    float u = 0.f, v = sbias[eg][hh];
    for (int c = 0; c < FDE; ++c) {
      u += wp[eg][c][hh];
      v += sbet[c] * sW[eg][c][hh];
    }
    (eg ? out->uG : out->uE)[hh] = u;
    (eg ? out->vG : out->vE)[hh] = v;
  }
  // operand images (head-group ordered, fused.h)
  for (int i = tid; i < 32 * 16; i += 128) {      // b_eg
    int n = i / 16, k = i % 16;
    int g = n / 16, key = (n / 8) % 2, eg = (n / 4) % 2, hh = 4 * g + n % 4, key2 = k / 8, c = k % 8;
    out->b_eg[(k / 8) * (32 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wp[eg][c][hh] - wlo[eg][c][hh] : 0.f);
    out->b_eg_lo[(k / 8) * (32 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wlo[eg][c][hh] : 0.f);
  }
  for (int i = tid; i < 16 * 16; i += 128) {      // b_hx, b_de[g]
    int n = i / 16, k = i % 16;
    {
      int g = n / 8, key = (n / 4) % 2, hh = 4 * g + n % 4, key2 = k / 8, c = k % 8;
      out->b_hx[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
          __float2bfloat16_rn(key == key2 ? sWr[hh][c] : 0.f);
    }
    {
      int key2 = n / 8, c = n % 8, g = k / 8, key = (k / 4) % 2, hh = 4 * g + k % 4;
      out->b_wr[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
          __float2bfloat16_rn(key == key2 ? sWr[hh][c] : 0.f);
    }
    for (int g = 0; g < 2; ++g) {
      int key2 = n / 8, c = n % 8, key = k / 8, eg = (k / 4) % 2, hh = 4 * g + k % 4;
      out->b_de[g][(k / 8) * (16 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wp[eg][c][hh] : 0.f);
    }
  }
}


// Folds the per-CTA partial sums of fused_bwd_kernel into the weight gradients (the library ADDS into them).
//   e^ = gamma (.) x^ + beta ;  [E|G] = e^ W + b ;  e' = e + H^ W_r + b_r
// One CTA of nthr >= 32 threads; runs as its own kernel or as an extra CTA of node_bwd2_kernel.
__device__ __forceinline__ void fused_bwd_finalize_body(const float *partials, int nparts, const egt_block_weights_t &w,
                                                        const egt_block_grads_t &g, const int tid, const int nthr) {
  __shared__ float s[FPART];
  for (int col = tid; col < FPART; col += nthr) {
    float acc[16];                           // 16 independent loads in flight: the loop is pure L2 latency
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    int i = 0;
    for (; i + 16 <= nparts; i += 16)
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] += partials[(size_t)(i + q) * FPART + col];
    for (; i < nparts; ++i) acc[0] += partials[(size_t)i * FPART + col];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += acc[q + 8];
    s[col] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  }
  __syncthreads();
  const float *M = s, *sZ = s + 128, *Wr = s + 144, *dbr = s + 208;
  for (int idx = tid; idx < 224; idx += nthr) {
    if (idx < 128) {                        // dW_E, dW_G
      const int c = idx / 16, j = idx % 16, eg = j / 8, hh = j % 8;
      float *dst = eg ? g.attention_gates_kernel : g.dense_edge_b_kernel;
      dst[c * FH + hh] += w.norm_edge_gamma[c] * M[c * 16 + j] + w.norm_edge_beta[c] * sZ[j];
    } else if (idx < 144) {                 // db_E, db_G
      const int j = idx - 128, eg = j / 8, hh = j % 8;
      (eg ? g.attention_gates_bias : g.dense_edge_b_bias)[hh] += sZ[j];
    } else if (idx < 208) {                 // dW_r
      g.dense_edge_r_kernel[idx - 144] += Wr[idx - 144];
    } else if (idx < 216) {                 // db_r
      g.dense_edge_r_bias[idx - 208] += dbr[idx - 208];
    } else {                                // dgamma_e, dbeta_e
      const int c = idx - 216;
      float dg = 0.f, db = 0.f;
      for (int hh = 0; hh < FH; ++hh) {
        const float we = w.dense_edge_b_kernel[c * FH + hh], wg = w.attention_gates_kernel[c * FH + hh];
        dg += we * M[c * 16 + hh] + wg * M[c * 16 + 8 + hh];
        db += we * sZ[hh] + wg * sZ[8 + hh];
      }
      g.norm_edge_gamma[c] += dg;
      g.norm_edge_beta[c] += db;
    }
  }
}

}  // namespace egt
