// wide_prep.cu -- per-call derived weights of the width-generic fused path (wide.h) and its shape gate.
//
//   reference: the parameters are those of norm_edge / attention_gates / dense_edge_b / dense_edge_r
//   (lib/models/graph_xformer_model_base.py:153-161,195,201-218); the folding of LayerNorm into the projections is
//   the one of fused_prep.cuh.
#include <math.h>
#include "common.cuh"
#include "wide.h"
#include "fused.h"    // kLoThreshold

namespace egt {

bool wide_supported(const egt_block_cfg_t *c, int dtype) {
  const egt_attn_cfg_t &a = c->attn;
  const bool shape = (a.h == 16 && a.dk == 8 && c->d_e == 32) || (a.h == 8 && a.dk == 8 && c->d_e == 64) ||
                     (a.h == 8 && a.dk == 12 && c->d_e == 8);
  return dtype == EGT_BF16 && shape && c->edge_channel_type == EGT_EDGE_RESIDUAL && c->gate_attention && a.has_clip &&
         a.clip_lo <= a.clip_hi && c->edge_act == EGT_ACT_NONE && !(a.training && a.attn_dropout > 0.f) && a.N >= 1 &&
         a.N <= 4096;
}

namespace {

// element offset of (n, k) in an un-swizzled K-major [N x K] operand image (8x16-byte core matrices)
__device__ __forceinline__ int img(int n, int k, int N) { return (k >> 3) * (N * 8) + n * 8 + (k & 7); }

__global__ void __launch_bounds__(256) wide_prep_kernel(egt_block_weights_t w, int H, int DE, float clip_lo, float clip_hi,
                                                        WidePrep *out) {
  __shared__ float wp[2][WMAXDE][WMAXH], sW[2][WMAXDE][WMAXH], sgam[WMAXDE], sbet[WMAXDE], sbias[2][WMAXH];
  const int tid = threadIdx.x;
  // every global load of the kernel is issued in this first phase (the rest works from shared memory: a dependent
  // chain of L2 round trips per output is what made this one-CTA kernel take 13 us)
  for (int i = tid; i < 2 * DE * H; i += 256) {
    const int eg = i / (DE * H), c = (i / H) % DE, hh = i % H;
    sW[eg][c][hh] = (eg ? w.attention_gates_kernel : w.dense_edge_b_kernel)[c * H + hh];
  }
  const bool cta0 = blockIdx.x == 0;                     // every CTA builds W' in shared memory; CTA 0 also writes the scalars
  const int gtid = blockIdx.x * 256 + tid, gstride = gridDim.x * 256;
  for (int c = tid; c < DE; c += 256) {
    sgam[c] = w.norm_edge_gamma[c]; sbet[c] = w.norm_edge_beta[c];
    if (cta0) out->br[c] = w.dense_edge_r_bias[c];
  }
  for (int i = tid; i < 2 * H; i += 256) sbias[i / H][i % H] = (i / H ? w.attention_gates_bias : w.dense_edge_b_bias)[i % H];
  __syncthreads();
  __shared__ float sbound;
  if (tid < 32) {   // bound of the logits from the full-precision W'_E (decides whether the lo half of W' is needed)
    float bnd = 0.f;
    if (tid < H) {
      float n2 = 0.f, v = sbias[0][tid];
      for (int c = 0; c < DE; ++c) { const float f = sgam[c] * sW[0][c][tid]; n2 += f * f; v += sbet[c] * sW[0][c][tid]; }
      // |LN(e)| has l2 norm <= sqrt(d_e)  =>  |E| <= sqrt(d_e) ||W'[:,hh]|| + |v|
      bnd = fmaxf(fabsf(clip_lo), fabsf(clip_hi)) + sqrtf((float)DE * n2) + fabsf(v);
    }
    for (int o = 16; o > 0; o >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffffffffu, bnd, o));
    if (tid == 0) { sbound = bnd; if (cta0) { out->bound = bnd; out->use_lo = bnd > kLoThreshold; } }
  }
  __syncthreads();
  const bool use_lo = sbound > kLoThreshold;
  for (int i = tid; i < 2 * DE * H; i += 256) {
    const int eg = i / (DE * H), c = (i / H) % DE, hh = i % H;
    const float full = sgam[c] * sW[eg][c][hh];
    const float hi = __bfloat162float(__float2bfloat16_rn(full));
    const float v = hi + (use_lo ? __bfloat162float(__float2bfloat16_rn(full - hi)) : 0.f);   // what the tensor core multiplies by
    wp[eg][c][hh] = v;
    if (cta0) out->wp[eg][c][hh] = v;
  }
  __syncthreads();
  if (cta0 && tid < 32) {
    const int eg = tid / 16, hh = tid % 16;
    if (hh < H) {
      float u = 0.f, v = sbias[eg][hh];
      for (int c = 0; c < DE; ++c) {
        u += wp[eg][c][hh];
        v += sbet[c] * sW[eg][c][hh];
      }
      (eg ? out->uG : out->uE)[hh] = u;
      (eg ? out->vG : out->vE)[hh] = v;
    }
  }
  const int EGN = 2 * H, DEP = DE < 16 ? 16 : DE, DEW = DEP;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  // column n of the [E|G] product <-> (eg, hh):  n = (hh/8)*16 + eg*8 + hh%8
  for (int i = gtid; i < 2 * EGN * DEW; i += gstride) {         // w_eg[v]: hi part in k < DEW, lo part in DEW <= k < 2 DEW
    const int v = i / (EGN * DEW), n = (i / DEW) % EGN, k = i % DEW;
    const int eg = (n >> 3) & 1, hh = 8 * (n >> 4) + (n & 7);
    float x;
    if (DE >= 16) x = wp[eg][k][hh];
    else x = (k >> 3) == v ? wp[eg][k & 7][hh] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    out->w_eg[v][img(n, k, EGN)] = hi;
    out->w_eg[v][img(n, DEW + k, EGN)] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
  for (int i = gtid; i < DEP * 16; i += gstride) {              // w_r, b_r
    const int n = i / 16, k = i % 16;
    out->w_r[img(n, k, DEP)] = (k < H && n < DE) ? __float2bfloat16_rn(w.dense_edge_r_kernel[k * DE + n]) : zero;
    float b = 0.f;
    if (n < DE) {
      const float full = w.dense_edge_r_bias[n];
      const float hi = __bfloat162float(__float2bfloat16_rn(full));
      b = k == 0 ? hi : k == 1 ? full - hi : 0.f;
    }
    out->b_r[img(n, k, DEP)] = __float2bfloat16_rn(b);
  }
  for (int i = gtid; i < 2 * 16 * 16; i += gstride) {           // i16[v]
    const int v = i / 256, n = (i / 16) % 16, k = i % 16;
    const bool one = DE >= 16 ? n == k : (n < 8 && k == n + 8 * v);
    out->i16[v][img(n, k, 16)] = __float2bfloat16_rn(one ? 1.f : 0.f);
  }
  for (int i = gtid; i < 2 * 16 * DEW; i += gstride) {          // w_hx[v]: n = head, k = edge channel
    const int v = i / (16 * DEW), n = (i / DEW) % 16, k = i % DEW;
    float x = 0.f;
    if (n < H) {
      if (DE >= 16) x = w.dense_edge_r_kernel[n * DE + k];
      else x = (k >> 3) == v ? w.dense_edge_r_kernel[n * DE + (k & 7)] : 0.f;
    }
    out->w_hx[v][img(n, k, 16)] = __float2bfloat16_rn(x);
  }
  for (int i = gtid; i < DEP * EGN; i += gstride) {             // w_dx: n = edge channel, k = column of [E|G]
    const int n = i / EGN, k = i % EGN;
    const int eg = (k >> 3) & 1, hh = 8 * (k >> 4) + (k & 7);
    out->w_dx[img(n, k, DEP)] = __float2bfloat16_rn(n < DE ? wp[eg][n][hh] : 0.f);
  }
}

}  // namespace

int wide_prep_launch(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, WidePrep *prep, cudaStream_t st) {
  LaunchScope _ls("wide_prep_kernel", st);
  wide_prep_kernel<<<8, 256, 0, st>>>(*w, cfg->attn.h, cfg->d_e, cfg->attn.clip_lo, cfg->attn.clip_hi, prep);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

}  // namespace egt
