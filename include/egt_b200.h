/* egt_b200.h -- C ABI of the B200-native EGT edge-augmented attention block.
 *
 * This is the drop-in boundary for ONE hot path of shamim-hussain/egt (TF2/Keras):
 *
 *   egt_attn_fwd / egt_attn_bwd    replace  EGT.call_gated / EGT.call_ungated
 *                                  (lib/models/egt_layers.py:57-143, :145-213) and the
 *                                  TF autodiff of them (driven by Model.fit,
 *                                  lib/training/training_base.py:294)
 *   egt_block_fwd / egt_block_bwd  replace  edge_update_{none,bias,residual} around
 *                                  mha_block (lib/models/graph_xformer_model_base.py
 *                                  :106-145, :149-162, :164-223; dispatch :328-339)
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types cross the boundary.
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns and
 *     allocates every buffer (inputs, outputs, saved statistics, workspace).  The library
 *     never allocates or frees device memory and keeps no pointer after it returns.
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t passed as void*), does
 *     not synchronise the host, and returns 0 or a negative egt_status; the message of the
 *     last failure on the calling thread is egt_last_error().
 *   - tensors are dense row-major with the reference's layouts (head axis innermost):
 *       h, h_out, v_att      [B, N, d]            channel c = dd*h + hh   (egt_layers.py:73-76,139-141)
 *       qkv                  [B, N, 3*d]          channel   = s*d + dd*h + hh,  s = 0,1,2 for Q,K,V
 *       e, e_out             [B, N, N, d_e]
 *       E, G, H_hat, A_tild  [B, N, N, h]         (egt_layers.py:79-80)
 *       mask                 [B, N]  uint8, 1 = real node  (Keras mask of h; only KEYS are masked,
 *                                                  egt_layers.py:91-94)
 *       lse                  [2, B, N, h] float32   row max and log row-sum of the masked logits (kept
 *                                                  apart so an all-keys-masked row at -1e9 stays exact)
 *       deg                  [B, N, h] float32      gate sum (centrality)
 *   - `dtype` selects the element type of activations / activation gradients (EGT_F32 or
 *     EGT_BF16).  Weights, weight gradients and saved row statistics are always float32.
 */
#ifndef EGT_B200_H_
#define EGT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGT_ABI_VERSION 1

typedef enum egt_status {
  EGT_OK = 0,
  EGT_E_SHAPE = -1,   /* unsupported / inconsistent sizes  (reference: assert at egt_layers.py:70) */
  EGT_E_DTYPE = -2,
  EGT_E_ALIGN = -3,
  EGT_E_ARCH = -4,    /* not an sm_100 device */
  EGT_E_CUDA = -5,    /* a CUDA runtime call failed */
  EGT_E_ARG = -6      /* bad flag combination      (reference: ValueError at egt_layers.py:20-24) */
} egt_status;

enum { EGT_F32 = 0, EGT_BF16 = 1 };
enum { EGT_SCALER_LOG = 0, EGT_SCALER_LINEAR = 1 };                 /* egt_layers.py:23-24,125-130 */
enum { EGT_EDGE_NONE = 0, EGT_EDGE_BIAS = 1, EGT_EDGE_RESIDUAL = 2, /* graph_xformer_model_base.py:328-334 */
       EGT_EDGE_CONSTRAINED = 3 };
enum { EGT_ACT_NONE = 0, EGT_ACT_LRELU = 1, EGT_ACT_RELU = 2, EGT_ACT_ELU = 3,
       EGT_ACT_TANH = 4, EGT_ACT_SIGMOID = 5 };                     /* graph_xformer_model_base.py:149-162 */
enum { EGT_MASK_NONE = 0,
       EGT_MASK_DENSE = 1,   /* M: [B,N,N,h] of the activation dtype, 0/1 (egt_layers.py:96-101)      */
       EGT_MASK_ADJ_U8 = 2   /* M: [B,N,N] uint8 adjacency, broadcast over heads
                                (what AdjMatModel.get_edge_mask tiles, graph_model_base.py:131-142)  */ };

/* Constructor arguments of the reference `EGT` layer (egt_layers.py:5-16) + sizes. */
typedef struct egt_attn_cfg {
  int32_t B, N, h, dk;          /* d = h*dk;  dk = qkv_channels / (3*h)  (egt_layers.py:70-71) */
  int32_t dtype;                /* EGT_F32 | EGT_BF16 */
  int32_t edge_input;           /* E present            */
  int32_t gate_input;           /* G present            */
  int32_t attn_mask;            /* EGT_MASK_*           */
  int32_t has_clip;             /* clip_logits_value is not None */
  float clip_lo, clip_hi;
  int32_t scale_degree;
  int32_t scaler_type;          /* EGT_SCALER_* */
  int32_t num_virtual_nodes;
  int32_t training;             /* random mask / dropout only when nonzero (egt_layers.py:103,116) */
  float random_mask_prob;
  float attn_dropout;
  uint64_t seed;                /* counter-based RNG (Philox4x32-10), see DESIGN.md "RNG" */
  uint64_t offset;
  const uint64_t *offset_dev;   /* optional DEVICE pointer: the kernels add *offset_dev to offset when they run, so a step
                                 * captured into a CUDA graph draws new noise at every replay (bump the word between replays);
                                 * NULL: offset alone */
} egt_attn_cfg_t;

/* Hyper-parameters of GraphTransformerBase that reach the attention block
 * (graph_xformer_model_base.py:17-45). */
typedef struct egt_block_cfg {
  egt_attn_cfg_t attn;          /* attn.edge_input / gate_input / attn_mask are derived by the library */
  int32_t d_e;                  /* edge_width */
  int32_t edge_channel_type;    /* EGT_EDGE_* */
  int32_t gate_attention;
  int32_t edge_act;             /* EGT_ACT_*  */
  float edge_act_alpha;         /* 'lreluX' -> X/10 (graph_xformer_model_base.py:150-152) */
  float ln_eps;                 /* keras LayerNormalization default 1e-3 */
} egt_block_cfg_t;

/* float32 parameters, Keras layouts: Dense kernel [in,out]; names of SURVEY.md appendix D. */
typedef struct egt_block_weights {
  const float *norm_mha_gamma, *norm_mha_beta;            /* [d]            norm_mha_{tag}        */
  const float *dense_qkv_kernel, *dense_qkv_bias;         /* [d,3d], [3d]   dense_qkv_{tag}       */
  const float *dense_mha_kernel, *dense_mha_bias;         /* [d,d], [d]     dense_mha_{tag}       */
  const float *norm_edge_gamma, *norm_edge_beta;          /* [d_e]          norm_edge_{tag}       */
  const float *attention_gates_kernel, *attention_gates_bias; /* [d_e,h],[h] attention_gates_{tag} */
  const float *dense_edge_b_kernel, *dense_edge_b_bias;   /* [d_e,h], [h]   dense_edge_b_{tag}    */
  const float *dense_edge_r_kernel, *dense_edge_r_bias;   /* [h,d_e], [d_e] dense_edge_r_{tag}    */
} egt_block_weights_t;

/* float32 gradient accumulators (the library ADDS into them); same shapes as the weights.
 * They may all point into one flat buffer so that data-parallel training needs a single
 * all-reduce (MirroredStrategy semantics, lib/training/training_base.py:230-238). */
typedef struct egt_block_grads {
  float *norm_mha_gamma, *norm_mha_beta;
  float *dense_qkv_kernel, *dense_qkv_bias;
  float *dense_mha_kernel, *dense_mha_bias;
  float *norm_edge_gamma, *norm_edge_beta;
  float *attention_gates_kernel, *attention_gates_bias;
  float *dense_edge_b_kernel, *dense_edge_b_bias;
  float *dense_edge_r_kernel, *dense_edge_r_bias;
} egt_block_grads_t;

typedef struct egt_block_fwd_io {
  const void *h;            /* [B,N,d]                     */
  const void *e;            /* [B,N,N,d_e] (NULL for EGT_EDGE_NONE) */
  const uint8_t *mask;      /* [B,N] or NULL               */
  const uint8_t *adj;       /* [B,N,N] uint8, only for EGT_EDGE_CONSTRAINED */
  void *h_out;              /* [B,N,d]                     */
  void *e_out;              /* [B,N,N,d_e]; written only for residual/constrained */
  /* saved for the backward (caller keeps them alive) */
  void *qkv;                /* [B,N,3d]  activation dtype  */
  void *v_att;              /* [B,N,d]   activation dtype  */
  float *lse;               /* [2,B,N,h]                   */
  float *deg;               /* [B,N,h]                     */
  void *workspace;          /* egt_block_workspace_bytes() */
  size_t workspace_bytes;
} egt_block_fwd_io_t;

typedef struct egt_block_bwd_io {
  const void *h, *e;
  const uint8_t *mask, *adj;
  const void *qkv, *v_att;  /* saved by the forward */
  const float *lse, *deg;
  const void *dh_out;       /* [B,N,d]      upstream gradient of h_out */
  const void *de_out;       /* [B,N,N,d_e]  upstream gradient of e_out (NULL = zero) */
  void *dh;                 /* [B,N,d]      */
  void *de;                 /* [B,N,N,d_e]  (NULL for EGT_EDGE_NONE) */
  void *workspace;
  size_t workspace_bytes;
} egt_block_bwd_io_t;

int egt_abi_version(void);
const char *egt_last_error(void);

/* Number of floats in the flat gradient buffer of one block and the offset of each tensor in it,
 * in the order of egt_block_weights_t (absent tensors get offset -1).  offsets_host: int64[14]. */
int64_t egt_block_param_layout(const egt_block_cfg_t *cfg, int64_t *offsets_host);

/* ---- EGT layer: ([QKV, E?, G?, M?], mask) -> (V_att, H_hat, A_tild?) ------------------- */
/* h_hat may be NULL (not materialised); a_tild NULL = skip (only Analysis taps consume it,
 * graph_xformer_model_base.py:134). lse: [2,B,N,h], deg: [B,N,h] float32, written. */
int egt_attn_fwd(const egt_attn_cfg_t *cfg, const void *qkv, const void *E, const void *G,
                 const void *M, const uint8_t *mask, void *v_att, void *h_hat, void *a_tild,
                 float *lse, float *deg, void *stream);

/* d_h_hat may be NULL (treated as zero).  Writes d_qkv [B,N,3d], dE, dG [B,N,N,h]
 * (dE / dG may be NULL when the corresponding input is absent).
 * row_ws: float32 scratch of 2*B*N*h elements (row terms shared by the row and column passes). */
int egt_attn_bwd(const egt_attn_cfg_t *cfg, const void *qkv, const void *E, const void *G,
                 const void *M, const uint8_t *mask, const float *lse, const float *deg,
                 const void *d_v_att, const void *d_h_hat, void *d_qkv, void *dE, void *dG,
                 float *row_ws, void *stream);

/* ---- attention block: (h, e, mask) -> (h', e') ----------------------------------------- */
size_t egt_block_workspace_bytes(const egt_block_cfg_t *cfg, int32_t backward);
int egt_block_fwd(const egt_block_cfg_t *cfg, const egt_block_weights_t *w,
                  const egt_block_fwd_io_t *io, void *stream);
int egt_block_bwd(const egt_block_cfg_t *cfg, const egt_block_weights_t *w,
                  const egt_block_grads_t *g, const egt_block_bwd_io_t *io, void *stream);

/* Which implementation the last egt_block_fwd / egt_block_bwd on this thread dispatched to:
 * 0 = staged kernels (any shape, fp32 or bf16), 1 = fused tcgen05 kernel. */
int egt_last_path(void);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
long egt_launch_count(void);
/* Per-kernel CUDA-event timing of every launch (bench.py's roofline leg).  enable(1) clears the table. */
int egt_profile_enable(int on);
int egt_profile_read(char *names_host, double *ms_host, long *counts_host, int max_entries);

/* ---- feed-forward half of a layer ("next" row) ---------------------------------------- */
/* ffnlr1 / ffnact / ffnlr2 of graph_xformer_model_base.py:229-258 as ffn_block (:309-324) applies them to one
 * channel:  y = x + Dense_w( act( Dense_hidden( LayerNorm(x) ) ) ),  hidden = round(w * ffn_multiplier).
 * rows = B*N with w = model_width for the node channel, rows = B*N*N with w = edge_width for the edge channel.
 * x, y, dy, dx: [rows, width] of the activation dtype; weights / gradients float32, Keras layouts. */
typedef struct egt_ffn_cfg {
  int64_t rows;
  int32_t width, hidden;
  int32_t dtype;            /* EGT_F32 | EGT_BF16 */
  int32_t activation;       /* EGT_ACT_* (config.activation, default 'elu') */
  float ln_eps;             /* keras LayerNormalization default 1e-3 */
} egt_ffn_cfg_t;
typedef struct egt_ffn_weights {
  const float *norm_gamma, *norm_beta;     /* [w]          norm_fnn_{node|edge}_{tag} */
  const float *lr1_kernel, *lr1_bias;      /* [w,hidden]   fnn_lr1_{node|edge}_{tag}  */
  const float *lr2_kernel, *lr2_bias;      /* [hidden,w]   fnn_lr2_{node|edge}_{tag}  */
} egt_ffn_weights_t;
typedef struct egt_ffn_grads {             /* accumulators: the library ADDS into them */
  float *norm_gamma, *norm_beta, *lr1_kernel, *lr1_bias, *lr2_kernel, *lr2_bias;
} egt_ffn_grads_t;
int egt_ffn_fwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *stream);
int egt_ffn_bwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                const void *dy, void *dx, void *stream);
/* The same two calls with a caller-owned device workspace of egt_ffn_workspace_bytes(cfg) bytes (256-byte aligned).
 * With a workspace, channel shapes that neither tensor-core kernel serves (node channel at d = 96 / 128) run as cuBLAS
 * GEMMs with element-wise kernels around them instead of the CUDA-core kernels; egt_ffn_workspace_bytes returns 0 when
 * the shape does not use one (ws may then be NULL). */
size_t egt_ffn_workspace_bytes(const egt_ffn_cfg_t *cfg);
int egt_ffn_fwd_ws(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *ws, size_t ws_bytes,
                   void *stream);
int egt_ffn_bwd_ws(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x,
                   const void *dy, void *dx, void *ws, size_t ws_bytes, void *stream);

/* ---- data parallelism: the single gradient all-reduce of MirroredStrategy (training_base.py:230-238) --- */
/* One-shot SUM all-reduce of `grad` (n float32, n % 4 == 0) over peer-mapped memory (NVLink / NVSwitch).
 * buffer_ptrs_dev / signal_pad_ptrs_dev: device arrays of `world` pointers to every rank's symmetric buffer
 * (>= 2*n floats) and zero-initialised signal pad (>= 1.1 KB), as torch symmetric memory hands out.
 * Stream-ordered, no host synchronisation, CUDA-graph capturable; all ranks must call it equally often. */
int egt_peer_allreduce(const uint64_t *buffer_ptrs_dev, const uint64_t *signal_pad_ptrs_dev, float *grad,
                       int64_t n, int rank, int world, void *stream);
/* The same collective as STORES with the flag inside every 8-byte word (no fence, no flag round trip, no remote load):
 * each rank writes {x0, flag, x1, flag} into slot [rank] of every peer's symmetric buffer and adds the slots of its own
 * buffer in rank order.  buffer_floats = size of every rank's symmetric buffer, at least
 * egt_peer_allreduce_push_floats(n, world) = 4 n world float32, zero-filled before the first call; n % 2 == 0. */
int64_t egt_peer_allreduce_push_floats(int64_t n, int world);
int egt_peer_allreduce_push(const uint64_t *buffer_ptrs_dev, const uint64_t *signal_pad_ptrs_dev, float *grad, int64_t n,
                            int64_t buffer_floats, int rank, int world, void *stream);

/* Counter-based uniform in (0,1) used for the random key mask / dropout (testing hook; host code).
 * stream_id 0 = random mask, 1 = attention dropout.  idx is the RNG element index: Philox call idx >> 3, 16-bit
 * lane idx & 7.  Element (b,l,m,hh) of a [B,N,N,h] tensor uses
 *   idx = ((((b*N + l)*ceil(N/2) + m/2)*ceil(h/4) + hh/4) << 3) | (m & 1) << 2 | (hh & 3)
 * (one call = two consecutive keys x four consecutive heads; csrc/common.cuh rng_elem_index, tests/philox.py). */
float egt_rng_uniform_host(uint64_t seed, uint64_t offset, uint32_t stream_id, uint64_t idx);

/* Testing hook: nonzero routes every shape through the staged kernels (process-wide). */
int egt_debug_force_staged(int on);

/* Bring-up hook: one tcgen05.mma chain D[128,N] = A[128,16*ksteps] * B[N,16*ksteps]^T with A / B laid out
 * in shared memory (or TMEM for A) in the canonical layout `*_mode` selects
 * (0 K-major 128B swizzle, 1 TMEM (A only), 2 MN-major 128B swizzle, 3 K-major no swizzle, 4 MN-major no
 * swizzle) and the given descriptor offsets.  tests/test_umma_probe.py pins every operand form the fused
 * kernels rely on.  A, B: bf16 device pointers; D: float32 [128,N]. */
int egt_debug_umma_probe(int a_mode, int b_mode, int N, int ksteps, uint32_t a_lbo, uint32_t a_sbo,
                         uint32_t b_lbo, uint32_t b_sbo, const uint32_t *a_off_host, const uint32_t *b_off_host,
                         const void *A, const void *B, float *D, void *stream);

/* Measurement hook behind DESIGN.md's cost model of small tcgen05.mma instructions (tools/mma_timing.py): issues
 * `chains` k-chains of `ksteps` instructions (M = 128, K = 16, N as given; a_mode 0 = A from shared memory K-major,
 * 1 = A from tensor memory, 2 = A from shared memory MN-major) rotating over `ndst` accumulators and writes
 * {cycles from first issue to completion, cycles spent issuing} to cycles_host[2].  Allocates 16 bytes of device
 * memory for the duration of the call (the only entry point that does). */
int egt_debug_mma_timing(int a_mode, int N, int ksteps, int chains, int ndst, long long *cycles_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EGT_B200_H_ */
