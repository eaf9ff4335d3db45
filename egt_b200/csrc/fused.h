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

// Derived weights, rebuilt on the device at the start of every forward / backward call (the weights
// change every optimiser step).  LayerNorm_e is folded into the projections:
//   e^ W = r * (e W') - r*mu * u + v,   W' = gamma (.) W,  u = colsum(W'),  v = beta W + b
// so that the tensor core multiplies the RAW edge tile and the per-pair (mu, r) are applied in registers.
struct FusedPrep {
  // tcgen05 B-operand images (bf16, un-swizzled K-major 8x16B core matrices), see fused_prep.cu
  __nv_bfloat16 wblk[2 * 32 * 8];     // [E|G] projection of a key PAIR:   N = 32, K = 16
  __nv_bfloat16 wrblk[2 * 16 * 8];    // edge write-back of a key pair:    N = 16, K = 16
  __nv_bfloat16 wtblk[4 * 16 * 8];    // backward d e^ = dZ W'^T:          N = 16, K = 32
  __nv_bfloat16 wrtblk[2 * 16 * 8];   // backward dH_ext = de' W_r^T:      N = 16, K = 16
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
  float clip_lo, clip_hi;
  int scale_degree, scaler_type, num_virtual_nodes;
  int rand_mask; uint32_t rand_thr;   // mask element when its 16 random bits < rand_thr
  uint64_t seed, offset;
};

struct FusedTensorMaps { CUtensorMap e, e_out, q, kv; };

bool fused_supported(const egt_block_cfg_t *cfg, int dtype);
int fused_prep_launch(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, FusedPrep *prep, cudaStream_t st);
// qkv: [B,N,192] bf16 with the Q third pre-multiplied by dk^-0.5
int fused_fwd_launch(const FusedFwdArgs &a, const void *e, void *e_out, const void *qkv, cudaStream_t st);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128);

}  // namespace egt
