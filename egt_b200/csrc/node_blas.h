// node_blas.h -- node-channel side of the block on cuBLAS (node_blas.cu): model widths other than 64.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/egt_b200.h"

namespace egt {

bool node_blas_supported(int d);                       // EGT_NODE_BLAS=0 turns it off (staged CUDA-core kernels instead)
size_t node_blas_workspace_bytes(int R, int d);
// qkv [R,3d] bf16 = LN(h) W_qkv + b_qkv, Q third times qscale
int node_blas_qkv(const void *h, const egt_block_weights_t *w, float eps, float qscale, void *qkv, int R, int d, void *ws,
                  cudaStream_t st);
int node_blas_out(const void *v_att, const void *h, const egt_block_weights_t *w, void *h_out, int R, int d, void *ws,
                  cudaStream_t st);
// side: stream for the weight-gradient kernels that feed nothing downstream (pass st to keep everything on one stream);
// the caller makes `side` wait for the inputs before the call and joins it before it returns
int node_blas_bwd1(const void *dh_out, const void *v_att, const egt_block_weights_t *w, const egt_block_grads_t *g, void *d_v_att,
                   int R, int d, void *ws, cudaStream_t st, cudaStream_t side);
int node_blas_bwd2(const void *h, const float *d_qkv, const egt_block_weights_t *w, const egt_block_grads_t *g, float eps,
                   float *dhn, int R, int d, void *ws, cudaStream_t st, cudaStream_t side, cudaEvent_t ev_operands);

// feed-forward half of a layer on cuBLAS (widths / hidden sizes ffn_tc.cu does not serve); EGT_FFN_BLAS=0 turns it off
bool ffn_blas_supported(const egt_ffn_cfg_t *cfg);
size_t ffn_blas_workspace_bytes(const egt_ffn_cfg_t *cfg);
int ffn_blas_fwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const void *x, void *y, void *ws, cudaStream_t st);
int ffn_blas_bwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *w, const egt_ffn_grads_t *g, const void *x, const void *dy,
                 void *dx, void *ws, cudaStream_t st);

}  // namespace egt
