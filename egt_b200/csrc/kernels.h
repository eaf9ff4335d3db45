// kernels.h -- internal launch interfaces between abi.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/egt_b200.h"

namespace egt {

struct AttnParams {
  int B, N, h, dk;
  float scale;                       // dk^-0.5
  int has_clip; float clip_lo, clip_hi;
  int scale_degree, scaler_type, num_virtual_nodes;
  int attn_mask;                     // EGT_MASK_*
  int rand_mask, dropout;            // already gated on `training`
  float random_mask_prob, attn_dropout;
  uint64_t seed, offset;
  const uint64_t *offset_dev;        // device word added to offset at run time (NULL: none)
  const void *qkv, *E, *G, *M;
  const uint8_t *mask;
  void *v_att, *h_hat, *a_tild;
  float *lse, *deg;
  // backward
  const void *d_v_att, *d_h_hat;
  void *d_qkv, *dE, *dG;
  float *row_ws;                     // [2,B,N,h]: D, s
  void *dS_ws, *As_ws;               // optional [B,N,N,h] dtype: dS and s*A~ from the row pass, so that the column pass
                                     // is two plain accumulations (attn_fast.cu); NULL = the column pass recomputes
  float dq_scale;                    // dQ is multiplied by this on store (1, or dk^-0.5 when qkv holds a pre-scaled Q)
};

int attn_staged_fwd(const AttnParams &P, int dtype, cudaStream_t st);
int attn_staged_bwd(const AttnParams &P, int dtype, cudaStream_t st);
// shape-specialised versions (attn_fast.cu); kind 0 forward, 1 backward; returns 1 when the shape is not covered
int attn_fast_launch(int kind, const AttnParams &P, int dtype, cudaStream_t st);
bool staged_force_generic();   // EGT_STAGED_GENERIC=1

// ---- node side (node_kernels.cu) -------------------------------------------------------
// out[r,:] = (LN? LN(x[r,:]) : x[r,:]) @ W (+bias) (+res[r,:]);  W is [din,dout] (trans=0) or
// [dout,din] read transposed (trans=1).  x,res,out have element type `dtype`; xf32 selects a
// float32 x regardless of dtype (workspace intermediates).
struct LinearArgs {
  const void *x; int x_f32;
  const float *W; int trans;
  const float *bias;
  const void *res;
  void *out; int out_f32;
  const float *ln_gamma, *ln_beta; float ln_eps;    // LN prologue when ln_gamma != NULL
  float *xn_out;                                    // optional: write LN(x) as float32 [R,din]
  float scale; int scale_cols;                      // out[:, :scale_cols] *= scale (0 cols = off)
  int R, din, dout;
};
int linear_launch(const LinearArgs &a, int dtype, cudaStream_t st);

// dW[i,j] += sum_r X[r,i] * Y[r,j];  db[j] += sum_r Y[r,j]   (float32 accumulators, atomics)
struct XtyArgs {
  const void *X; int x_f32;
  const void *Y; int y_f32;
  float *dW; float *db;
  int R, dx, dy;
};
int xty_launch(const XtyArgs &a, int dtype, cudaStream_t st);

// LayerNorm backward + residual:  dx = LN_bwd(dy; x, gamma) + dres ;  dgamma,dbeta += ...
struct LnBwdArgs {
  const void *x;           // [R,D] dtype
  const float *dy;         // [R,D] float32 (gradient wrt LN output)
  const void *dres;        // [R,D] dtype (residual-path gradient) or NULL
  const float *gamma; float eps;
  void *dx;                // [R,D] dtype
  float *dgamma, *dbeta;
  int R, D;
};
int ln_bwd_launch(const LnBwdArgs &a, int dtype, cudaStream_t st);

// ---- edge side (edge_kernels.cu) -------------------------------------------------------
struct EdgeParams {
  size_t pairs;            // B*N*N
  int d_e, h;
  int has_ln; float ln_eps;
  int gated;
  int act; float act_alpha;
  const void *e;           // [pairs,d_e]
  const float *ln_g, *ln_b, *w_e, *b_e, *w_g, *b_g, *w_r, *b_r;
  void *E, *G;             // [pairs,h]  (dtype)
  const void *h_hat;       // [pairs,h]
  void *e_out;             // [pairs,d_e]
  // backward
  const void *de_out;      // [pairs,d_e] or NULL
  void *d_h_ext;           // [pairs,h]  = de_out @ W_r^T
  const void *dE, *dG;     // [pairs,h]
  void *de;                // [pairs,d_e]
  float *g_ln_g, *g_ln_b, *g_w_e, *g_b_e, *g_w_g, *g_b_g, *g_w_r, *g_b_r;
};
int edge_proj_fwd_launch(const EdgeParams &p, int dtype, cudaStream_t st);
int edge_out_fwd_launch(const EdgeParams &p, int dtype, cudaStream_t st);
int edge_out_bwd_launch(const EdgeParams &p, int dtype, cudaStream_t st);
int edge_proj_bwd_launch(const EdgeParams &p, int dtype, cudaStream_t st);
// shape-specialised versions (edge_fast.cu); kind 0..3 in the order above; returns 1 when the shape is not covered
int edge_fast_launch(int kind, const EdgeParams &p, int dtype, cudaStream_t st);

// feed-forward half on tensor cores (ffn_tc.cu): bf16, width 8/16/32/64, hidden = 2 width; returns 1 when the shape is
// not served (the caller falls back to the CUDA-core kernels of ffn_kernels.cu); EGT_FFN_TC=0 switches it off
bool ffn_tc_serves(const egt_ffn_cfg_t *cfg);   // shape / dtype / activation served (pointer alignment checked at launch)
int ffn_tc_fwd_launch(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, cudaStream_t st);
int ffn_tc_bwd_launch(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                      const void *dy, void *dx, cudaStream_t st);

}  // namespace egt
