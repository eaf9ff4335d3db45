// wide.h -- interface of the width-generic fused tcgen05 path (wide_prep.cu, wide_fwd.cu, wide_bwd.cu).
//
// fused_fwd.cu / fused_bwd.cu are specialised for the narrow edge channel of the reference's MNIST / CLUSTER /
// PATTERN configs (h = 8, dk = 8, d_e = 8).  The kernels declared here serve the other widths BASELINE.json names,
// one template instantiation each:
//     h = 16, dk = 8,  d_e = 32   synthetic roofline sweep, N in {64..512}      (BASELINE config 5)
//     h = 8,  dk = 8,  d_e = 64   ZINC  (configs/main/zinc/500k/egt.json:10-12)  (BASELINE config 1)
//     h = 8,  dk = 12, d_e = 8    CLUSTER at model width 96                      (BASELINE config 3)
// Same contract as the narrow path: bf16 activations, gated residual edge channel with logit clipping, no
// [B,N,N,h] tensor in HBM, e streamed in once and e' out once by TMA.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/egt_b200.h"

namespace egt {

constexpr int WMAXH = 16, WMAXDE = 64;

// Derived weights, rebuilt on the device at the start of every forward / backward call.  LayerNorm_e is folded
// into the projections exactly as in fused.h:  e^ W = r (e W') - r mu u + v.
// tcgen05 B-operand images are bf16, un-swizzled K-major 8x16-byte core matrices: element (n, k) of an [N x K]
// matrix at element offset (k/8) * (N*8) + n*8 + k%8.
struct WidePrep {
  // [E|G] projection of ONE key: N = 2h columns ordered (hh/8, eg, hh%8), K = max(d_e,16) raw edge channels.
  // d_e = 8: the K window covers the key pair (key & ~1, key | 1); image v = key & 1 holds W' in rows 8v..8v+7.
  // W' is kept as bf16 hi + bf16 lo (K doubled: rows K.. hold the part of W' the first rounding lost), so the
  // projections of the bf16 edge channels are exact to ~2^-17 -- large trained weights move logits by hundreds
  // of bf16 ulps otherwise.
  __nv_bfloat16 w_eg[2][2 * (WMAXDE / 8) * 2 * WMAXH * 8];
  __nv_bfloat16 w_r[2 * WMAXDE * 8];        // edge write-back H^ W_r: N = max(d_e,16), K = 16 heads (zero padded)
  __nv_bfloat16 b_r[2 * WMAXDE * 8];        // bias of the write-back as a K = 16 operand: k = 0 bf16(b_r), k = 1 the rest
  __nv_bfloat16 i16[2][2 * 16 * 8];         // 16 x 16 identity (d_e = 8: the two halves of the key-pair window)
  // backward
  __nv_bfloat16 w_hx[2][(WMAXDE / 8) * 16 * 8];       // dH_ext = de' W_r^T of one key: N = 16 heads, K = max(d_e,16)
  __nv_bfloat16 w_dx[(2 * WMAXH / 8) * WMAXDE * 8];   // d x^ = [dE|dG] W'^T: N = max(d_e,16), K = 2h (same order as w_eg's N)
  float uE[WMAXH], vE[WMAXH], uG[WMAXH], vG[WMAXH];
  float br[WMAXDE];
  float wp[2][WMAXDE][WMAXH];               // W'_E, W'_G (float32; hi + lo is what the forward multiplies by)
  float bound;                              // sup |masked logit| given these weights (fused.h)
  int use_lo;                               // bound > kLoThreshold: the lo half of w_eg is multiplied too (else it is zero)
};

struct WideFwdArgs {
  int B, N;
  const uint8_t *mask;                // [B,N] or NULL
  const WidePrep *prep;
  __nv_bfloat16 *v_att;               // [B,N,d]
  float *lse, *deg;                   // [2,B,N,h], [B,N,h]
  float clip_lo, clip_hi, ln_eps;
  int scale_degree, scaler_type, num_virtual_nodes;
  int rand_mask; uint32_t rand_thr;   // mask element when its 16 random bits < rand_thr
  uint64_t seed, offset;
  const uint64_t *offset_dev;         // device word added to offset at run time (NULL: none)
};

struct WideBwdArgs {
  int B, N;
  const uint8_t *mask;
  const WidePrep *prep;
  const __nv_bfloat16 *v_att, *d_v_att;   // [B,N,d]
  const float *lse, *deg;
  float *d_qkv;                       // [B,N,3d] float32: dQ | dK | dV
  float *partials;                    // weight-gradient partial sums, see wide_bwd.cu
  float clip_lo, clip_hi, dq_scale, ln_eps;
  int scale_degree, scaler_type, num_virtual_nodes;
  int rand_mask; uint32_t rand_thr;
  uint64_t seed, offset;
  const uint64_t *offset_dev;         // device word added to offset at run time (NULL: none)
};

// shape gate: which (h, dk, d_e) have an instantiation; flags as fused_supported()
bool wide_supported(const egt_block_cfg_t *cfg, int dtype);
int wide_prep_launch(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, WidePrep *prep, cudaStream_t st);
// qkv: [B,N,3d] bf16, reference channel order, Q third pre-multiplied by dk^-0.5
int wide_fwd_launch(const egt_block_cfg_t *cfg, const WideFwdArgs &a, const void *e, void *e_out, const void *qkv,
                    cudaStream_t st);
int wide_bwd_launch(const egt_block_cfg_t *cfg, const WideBwdArgs &a, const void *e, const void *de_out, void *de,
                    const void *qkv, cudaStream_t st);
bool wide_bwd_supported(const egt_block_cfg_t *cfg);   // shapes wide_bwd.cu instantiates (others pair with the staged backward)
size_t wide_bwd_partials_floats(const egt_block_cfg_t *cfg);
int wide_bwd_key_splits(int B, int N, int TK);
int wide_bwd_splits(const egt_block_cfg_t *cfg);       // CTAs that share the keys of one (graph, row tile); > 1: d_qkv must be zero-filled
int wide_bwd_finalize_launch(const egt_block_cfg_t *cfg, const float *partials, const egt_block_weights_t *w,
                             const egt_block_grads_t *g, const WidePrep *prep, cudaStream_t st);

}  // namespace egt
