// fused.h -- interface of the fused tcgen05 path (fused_prep.cu, fused_fwd.cu, fused_bwd.cu).
//
// Shapes served: bf16 activations, h = 8 heads of dk = 8 (d = 64), d_e = 8, gated residual edge channel
// with logit clipping and no edge activation -- the widths of the reference's MNIST / CLUSTER / PATTERN
// configs (configs/main/{mnist,cluster,pattern}/*/egt_spe.json) and of BASELINE.json's headline metric.
// Everything else runs on the staged kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/egt_b200.h"

namespace egt {

constexpr int FH = 8, FDK = 8, FD = 64, FDE = 8;
// exponent budget of the un-normalised softmax: exp(H_hat - max(bound - budget, 0)) <= e^75, and 4096 of them
// still sum below FLT_MAX
constexpr float kSoftmaxBudget = 75.f;
constexpr float kLoThreshold = 16.f;    // logit bound above which W' is multiplied as hi + lo (FusedPrep::b_eg_lo)

// Derived weights, rebuilt on the device at the start of every forward / backward call (the weights
// change every optimiser step).  LayerNorm_e is folded into the projections:
//   e^ W = r * (e W') - r*mu * u + v,   W' = gamma (.) W,  u = colsum(W'),  v = beta W + b
// so that the tensor core multiplies the RAW edge tile and the per-pair (mu, r) are applied in registers.
struct FusedPrep {
  // tcgen05 B-operand images (bf16, un-swizzled K-major 8x16B core matrices), see fused_prep.cuh.
  // Output columns are ordered by head GROUP g = hh / 4 (fused_bwd.cu: warps 0-3 own heads
  // 0-3, warps 4-7 heads 4-7), hh4 = hh % 4, for a key pair (key, key' in {0,1}):
  __nv_bfloat16 b_eg[2 * 32 * 8];     // [E|G] projection: N = 32 (g,key,eg,hh4), K = 16 (key',c): W'_eg[c,hh]
  __nv_bfloat16 b_hx[2 * 16 * 8];     // dH_ext = de' W_r^T: N = 16 (g,key,hh4), K = 16 (key',c): W_r[hh,c]
  __nv_bfloat16 b_de[2][2 * 16 * 8];  // d x^ = dZ W'^T, one image per g: N = 16 (key',c), K = 16 (key,eg,hh4)
  __nv_bfloat16 b_wr[2 * 16 * 8];     // forward edge write-back H^ W_r: N = 16 (key',c), K = 16 (g,key,hh4): W_r[hh,c]
  // second half of W'_eg = bf16 hi + bf16 lo (same layout as b_eg).  Used (use_lo != 0) only when the logit bound exceeds
  // kLoThreshold: with one bf16 rounding of the weights a logit of magnitude |E| is off by |E| 2^-9, harmless at the
  // usual |E| of a few units and whole units at |E| of several hundred.  wp / uE / uG always describe the weights the
  // tensor core multiplies by (hi, or hi + lo).
  __nv_bfloat16 b_eg_lo[2 * 32 * 8];
  int use_lo;
  float uE[FH], vE[FH], uG[FH], vG[FH], br[FDE];
  float wp[2][FDE][FH];               // W'_E, W'_G as rounded to bf16 (for the weight-gradient epilogue)
  float bound;                        // sup |masked logit| over all inputs given these weights
};

struct FusedFwdArgs {
  int B, N;
  const uint8_t *mask;                // [B,N] or NULL
  const FusedPrep *prep;
  __nv_bfloat16 *v_att;               // [B,N,64]
  float *lse, *deg;                   // [2,B,N,8], [B,N,8]
  float clip_lo, clip_hi, ln_eps;
  int scale_degree, scaler_type, num_virtual_nodes;
  int rand_mask; uint32_t rand_thr;   // mask element when its 16 random bits < rand_thr
  uint64_t seed, offset;
  const uint64_t *offset_dev;         // device word added to offset at run time (NULL: none)
  // output projection fused behind the attention (graph_xformer_model_base.py:136-140); w_o == NULL = off
  const float *w_o, *b_o;             // dense_mha kernel [64,64], bias [64]
  const __nv_bfloat16 *h;             // [B,N,64] residual input
  __nv_bfloat16 *h_out;               // [B,N,64]
};

// Partial weight-gradient sums one backward CTA writes (fused_bwd_finalize_kernel folds them):
//   M[c][j] = sum x^_c dZ_j (j = eg*8 + hh) | sZ[j] = sum dZ_j | Wr[hh][c] = sum H^_hh de'_c | dbr[c] = sum de'_c
constexpr int FPART = 8 * 16 + 16 + 64 + 8;

struct FusedBwdArgs {
  int B, N;
  const uint8_t *mask;                // [B,N] or NULL
  const FusedPrep *prep;
  const __nv_bfloat16 *v_att, *d_v_att;   // [B,N,64] saved forward output / its upstream gradient
  const float *lse, *deg;             // [2,B,N,8] (reference point | log row sum), [B,N,8]
  float *d_qkv;                       // [B,N,192] float32: dQ | dK | dV
  __nv_bfloat16 *d_qkv_bf;            // non-NULL and one row tile per graph: the same as bf16 here instead
  float *partials;                    // [gridDim.x * gridDim.y][FPART]
  float clip_lo, clip_hi, dq_scale, ln_eps;
  int scale_degree, scaler_type, num_virtual_nodes;
  int rand_mask; uint32_t rand_thr;
  uint64_t seed, offset;
  const uint64_t *offset_dev;         // device word added to offset at run time (NULL: none)
};

bool fused_supported(const egt_block_cfg_t *cfg, int dtype);
int fused_prep_launch(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, FusedPrep *prep, cudaStream_t st);   // stand-alone form (tests / tools)
// qkv: [B,N,192] bf16 with the Q third pre-multiplied by dk^-0.5
int fused_fwd_launch(const FusedFwdArgs &a, const void *e, void *e_out, const void *qkv, cudaStream_t st);

// e, de_out -> de (all [B,N,N,8] bf16); d_qkv must be zero-filled by the caller when N > 128 (dK/dV are
// accumulated with atomics across the row tiles of a graph).
int fused_bwd_launch(const FusedBwdArgs &a, const void *e, const void *de_out, void *de, const void *qkv,
                     cudaStream_t st);
int fused_bwd_key_splits(int B, int N);   // CTAs per (graph, row tile) of fused_bwd; > 1: d_qkv must be zero-filled, one partial row per CTA
int fused_bwd_finalize_launch(const float *partials, int nparts, const egt_block_weights_t *w,
                              const egt_block_grads_t *g, cudaStream_t st);

// node side on tensor cores (node_tc.cu); bf16 activations, d = 64
// prep_out != NULL: one extra CTA of the kernel writes the FusedPrep (saves the fused_prep_kernel launch)
int node_qkv_launch(const void *h, const float *gamma, const float *beta, float eps, const float *W, const float *bias,
                    float qscale, void *qkv, int R, const egt_block_weights_t *w, float clip_lo, float clip_hi,
                    FusedPrep *prep_out, cudaStream_t st);
int node_out_launch(const void *v_att, const void *h, const float *W, const float *bias, void *h_out, int R,
                    cudaStream_t st);
int node_bwd1_launch(const void *dh_out, const void *v_att, const float *W, void *d_v_att, float *dW, float *db, int R,
                     const egt_block_weights_t *w, float clip_lo, float clip_hi, FusedPrep *prep_out, cudaStream_t st,
                     cudaStream_t side);   // side != st: dW_O / db_O as a second small launch on `side`
int node_bwd2_launch(const void *h, const void *dh_out, const float *dqkv, const void *dqkv_bf, const float *gamma, const float *beta,
                     float eps, const float *W, void *dh, float *dW, float *db, float *dgamma, float *dbeta, int R,
                     const float *partials, int nparts, const egt_block_weights_t *w, const egt_block_grads_t *g,
                     cudaStream_t st, cudaStream_t side);   // partials != NULL: one extra CTA folds the fused backward's partial
                                                            // sums; side != st: weight-gradient half as a second launch on `side`

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128);

}  // namespace egt
