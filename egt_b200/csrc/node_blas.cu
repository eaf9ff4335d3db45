// node_blas.cu -- node-channel side of the attention block for the model widths the tcgen05 node kernels
// (node_tc.cu, d = 64) do not serve: d = 96 (CLUSTER at width 96) and d = 128 (the roofline sweep).
//
//   reference: mha_block (lib/models/graph_xformer_model_base.py:106-145): h^ = LN(h); QKV = h^ W_qkv + b;
//   h' = h + V_att W_O + b_O, and the TF autodiff of those lines.
//
// These are plain GEMMs over the flattened [B*N, d] node tensor (< 7 % of the block's bytes), so they go to cuBLAS
// (bf16 operands, float32 accumulation) with small element-wise kernels around them:
//   * the LayerNorm output is written as [h^ | 1 | 0..] with K = d + 8, and the weight image carries the bias in row
//     d, so the QKV bias -- and the dk^-0.5 scale of the Q third -- are part of the GEMM;
//   * weight gradients accumulate straight into the caller's float32 gradient buffer (beta = 1).
// cuBLAS works in the caller's workspace (cublasSetWorkspace) on the caller's stream; the handle is created on first
// use (outside any stream capture: the first call of a process is never a captured one, see bench.py).
#include <cublas_v2.h>
#include <mutex>
#include "common.cuh"
#include "kernels.h"
#include "node_blas.h"

namespace egt {

namespace {

cublasHandle_t g_handle[2][16] = {};   // [1]: GEMMs issued on the side stream (own workspace)
std::mutex g_handle_mu;

int get_handle(cublasHandle_t *out, int which = 0) {
  int dev = 0;
  EGT_CHECK_CUDA(cudaGetDevice(&dev));
  EGT_REQUIRE(dev >= 0 && dev < 16, EGT_E_ARG, "node_blas: device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_handle_mu);
  if (!g_handle[which][dev]) {
    cublasStatus_t s = cublasCreate(&g_handle[which][dev]);
    EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasCreate failed with status %d", (int)s);
  }
  *out = g_handle[which][dev];
  return EGT_OK;
}

// C[M,N] = alpha op(A) op(B) + beta C, every matrix ROW-major (leading dimensions in elements).
int gemm_rm(cublasHandle_t h, bool ta, bool tb, int M, int N, int K, const void *A, int lda, const void *B, int ldb, void *C,
            cudaDataType ctype, int ldc, float beta) {
  const float alpha = 1.f;
  cublasStatus_t s = cublasGemmEx(h, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &alpha, B,
                                  CUDA_R_16BF, ldb, A, CUDA_R_16BF, lda, &beta, C, ctype, ldc, CUBLAS_COMPUTE_32F,
                                  CUBLAS_GEMM_DEFAULT);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasGemmEx(%d,%d,%d) failed with status %d", M, N, K, (int)s);
  return EGT_OK;
}

// bf16 operand copies of the weights: w_aug [d+8, 3d] = [W_qkv ; b_qkv ; 0] with the Q third times qscale,
// w_qkv [d, 3d], w_o [d, d]
__global__ void __launch_bounds__(256) nb_prep_kernel(const float *Wqkv, const float *bqkv, const float *Wo, int d, float qscale,
                                                      __nv_bfloat16 *w_aug, __nv_bfloat16 *w_qkv, __nv_bfloat16 *w_o) {
  const int n3 = 3 * d, tot_aug = (d + 8) * n3;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < tot_aug; i += gridDim.x * 256) {
    const int r = i / n3, c = i % n3;
    float v = r < d ? Wqkv[i] : r == d ? bqkv[c] : 0.f;
    if (r < d && w_qkv) w_qkv[i] = __float2bfloat16_rn(v);
    if (c < d) v *= qscale;
    w_aug[i] = __float2bfloat16_rn(v);
  }
  if (w_o)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d * d; i += gridDim.x * 256) w_o[i] = __float2bfloat16_rn(Wo[i]);
}

// hn_aug[r, :] = [LN(h[r, :]) | 1 | 0 x 7]; one warp per row
__global__ void __launch_bounds__(256) nb_ln_aug_kernel(const __nv_bfloat16 *h, const float *gamma, const float *beta, float eps,
                                                        int R, int d, __nv_bfloat16 *out) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const __nv_bfloat16 *x = h + (size_t)r * d;
  float v[8], s = 0.f;                                   // d <= 256
  int n = 0;
  for (int c = lane; c < d; c += 32) { v[n] = __bfloat162float(x[c]); s += v[n]; ++n; }
  s = warp_sum(s);
  const float mu = s / d;
  float q = 0.f;
  for (int i = 0; i < n; ++i) { const float t = v[i] - mu; q = fmaf(t, t, q); }
  q = warp_sum(q);
  const float rs = rsqrtf(q / d + eps);
  __nv_bfloat16 *o = out + (size_t)r * (d + 8);
  n = 0;
  for (int c = lane; c < d; c += 32) { o[c] = __float2bfloat16_rn(fmaf((v[n] - mu) * rs, gamma[c], beta[c])); ++n; }
  if (lane < 8) o[d + lane] = __float2bfloat16_rn(lane == 0 ? 1.f : 0.f);
}

// h'[r, c] = h[r, c] + t[r, c] + b[c]
__global__ void __launch_bounds__(256) nb_out_epilogue_kernel(const __nv_bfloat16 *h, const float *t, const float *b, size_t n, int d,
                                                              __nv_bfloat16 *out) {
  for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * 1024) {   // d % 4 == 0
    const float4 tv = *(const float4 *)(t + i);
    const int c = (int)(i % d);
    const uint2 hv = *(const uint2 *)(h + i);
    const float h0 = __uint_as_float(hv.x << 16), h1 = __uint_as_float(hv.x & 0xFFFF0000u);
    const float h2 = __uint_as_float(hv.y << 16), h3 = __uint_as_float(hv.y & 0xFFFF0000u);
    __nv_bfloat162 o0 = __floats2bfloat162_rn(h0 + tv.x + b[c], h1 + tv.y + b[c + 1]);
    __nv_bfloat162 o1 = __floats2bfloat162_rn(h2 + tv.z + b[c + 2], h3 + tv.w + b[c + 3]);
    uint2 ov;
    ov.x = *(uint32_t *)&o0; ov.y = *(uint32_t *)&o1;
    *(uint2 *)(out + i) = ov;
  }
}

// db[c] += sum_r Y[r, c] and, optionally, a bf16 copy of a float32 Y.  CTA = 32 rows x all columns; a thread owns one
// column and keeps 8 row loads in flight (the loop is latency-bound otherwise).
constexpr int kColsumRows = 32;
template <typename T>
__global__ void __launch_bounds__(256) nb_colsum_kernel(const T *Y, int R, int C, float *db, __nv_bfloat16 *copy) {
  const int r0 = blockIdx.x * kColsumRows, r1 = min(R, r0 + kColsumRows);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (int r = r0; r < r1; r += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = r + u < r1 ? ldf(Y + (size_t)(r + u) * C + c) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc += v[u];
        if (copy && r + u < r1) copy[(size_t)(r + u) * C + c] = __float2bfloat16_rn(v[u]);
      }
    }
    atomicAdd(db + c, acc);
  }
}

size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carve {
  __nv_bfloat16 *hn_aug, *w_aug, *w_qkv, *w_o, *dqkv_bf;
  float *tmp;
  void *blas_ws, *blas_ws2;
};
constexpr size_t kBlasWs = 32u << 20;

Carve carve(void *base, int R, int d) {
  char *p = (char *)base;
  Carve c;
  auto take = [&](size_t bytes) { char *q = p; p += al(bytes); return q; };
  c.hn_aug = (__nv_bfloat16 *)take((size_t)R * (d + 8) * 2);
  c.w_aug = (__nv_bfloat16 *)take((size_t)(d + 8) * 3 * d * 2);
  c.w_qkv = (__nv_bfloat16 *)take((size_t)d * 3 * d * 2);
  c.w_o = (__nv_bfloat16 *)take((size_t)d * d * 2);
  c.dqkv_bf = (__nv_bfloat16 *)take((size_t)R * 3 * d * 2);
  c.tmp = (float *)take((size_t)R * d * 4);
  c.blas_ws = take(kBlasWs);
  c.blas_ws2 = take(kBlasWs);
  return c;
}

// which = 1: the handle of the side stream (its own cuBLAS workspace: its GEMMs run next to those of handle 0)
int begin(cublasHandle_t *h, const Carve &c, cudaStream_t st, int which = 0) {
  int rc = get_handle(h, which);
  if (rc) return rc;
  cublasStatus_t s = cublasSetStream(*h, st);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasSetStream failed with status %d", (int)s);
  s = cublasSetWorkspace(*h, which ? c.blas_ws2 : c.blas_ws, kBlasWs);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasSetWorkspace failed with status %d", (int)s);
  return EGT_OK;
}

}  // namespace

bool node_blas_supported(int d) {
  static const bool off = getenv("EGT_NODE_BLAS") && atoi(getenv("EGT_NODE_BLAS")) == 0;
  return !off && d % 8 == 0 && d <= 256;
}

size_t node_blas_workspace_bytes(int R, int d) {
  return al((size_t)R * (d + 8) * 2) + al((size_t)(d + 8) * 3 * d * 2) + al((size_t)d * 3 * d * 2) + al((size_t)d * d * 2) +
         al((size_t)R * 3 * d * 2) + al((size_t)R * d * 4) + 2 * al(kBlasWs);
}

int node_blas_qkv(const void *h, const egt_block_weights_t *w, float eps, float qscale, void *qkv, int R, int d, void *ws,
                  cudaStream_t st) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  {
    LaunchScope _ls("nb_prep_kernel", st);
    nb_prep_kernel<<<64, 256, 0, st>>>(w->dense_qkv_kernel, w->dense_qkv_bias, w->dense_mha_kernel, d, qscale, c.w_aug, nullptr, c.w_o);
  }
  {
    LaunchScope _ls("nb_ln_aug_kernel", st);
    nb_ln_aug_kernel<<<(R + 7) / 8, 256, 0, st>>>((const __nv_bfloat16 *)h, w->norm_mha_gamma, w->norm_mha_beta, eps, R, d, c.hn_aug);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  LaunchScope _ls("cublas_gemm", st);
  return gemm_rm(hd, false, false, R, 3 * d, d + 8, c.hn_aug, d + 8, c.w_aug, 3 * d, qkv, CUDA_R_16BF, 3 * d, 0.f);
}

// h' = h + V_att W_O + b_O   (w_o was converted by node_blas_qkv of the same forward call)
int node_blas_out(const void *v_att, const void *h, const egt_block_weights_t *w, void *h_out, int R, int d, void *ws,
                  cudaStream_t st) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, false, false, R, d, d, v_att, d, c.w_o, d, c.tmp, CUDA_R_32F, d, 0.f))) return rc;
  }
  LaunchScope _ls("nb_out_epilogue_kernel", st);
  const size_t n = (size_t)R * d;
  nb_out_epilogue_kernel<<<(unsigned)((n / 4 + 255) / 256 < 2048 ? (n / 4 + 255) / 256 : 2048), 256, 0, st>>>(
      (const __nv_bfloat16 *)h, c.tmp, w->dense_mha_bias, n, d, (__nv_bfloat16 *)h_out);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

// dV_att = dh' W_O^T ; dW_O += V_att^T dh' ; db_O += colsum(dh').  The two weight-gradient kernels feed nothing in the
// backward pass: they go to `side` (== st: same stream), where they run next to the fused backward kernel.
int node_blas_bwd1(const void *dh_out, const void *v_att, const egt_block_weights_t *w, const egt_block_grads_t *g, void *d_v_att,
                   int R, int d, void *ws, cudaStream_t st, cudaStream_t side) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd, hs;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  if ((rc = begin(&hs, c, side, 1))) return rc;
  {
    LaunchScope _ls("nb_colsum_kernel", side);
    nb_colsum_kernel<__nv_bfloat16><<<(R + kColsumRows - 1) / kColsumRows, 256, 0, side>>>((const __nv_bfloat16 *)dh_out, R, d, g->dense_mha_bias, nullptr);
  }
  {
    LaunchScope _ls("cublas_gemm", side);
    if ((rc = gemm_rm(hs, true, false, d, d, R, v_att, d, dh_out, d, g->dense_mha_kernel, CUDA_R_32F, d, 1.f))) return rc;
  }
  {
    LaunchScope _ls("nb_prep_kernel", st);
    nb_prep_kernel<<<64, 256, 0, st>>>(w->dense_qkv_kernel, w->dense_qkv_bias, w->dense_mha_kernel, d, 1.f, c.w_aug, c.w_qkv, c.w_o);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  LaunchScope _ls("cublas_gemm", st);
  return gemm_rm(hd, false, true, R, d, d, dh_out, d, c.w_o, d, d_v_att, CUDA_R_16BF, d, 0.f);
}

// dW_qkv += LN(h)^T dqkv ; db_qkv += colsum(dqkv) ; dhn = dqkv W_qkv^T (float32, for ln_bwd_kernel)
// (w_qkv was converted by node_blas_bwd1 of the same backward call).  The weight-gradient GEMM goes to `side` once its
// operands exist (fork / join through the two events; side == st: everything on one stream).
int node_blas_bwd2(const void *h, const float *d_qkv, const egt_block_weights_t *w, const egt_block_grads_t *g, float eps,
                   float *dhn, int R, int d, void *ws, cudaStream_t st, cudaStream_t side, cudaEvent_t ev_operands) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd, hs;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  if ((rc = begin(&hs, c, side, 1))) return rc;
  {
    LaunchScope _ls("nb_colsum_kernel", st);
    nb_colsum_kernel<float><<<(R + kColsumRows - 1) / kColsumRows, 256, 0, st>>>(d_qkv, R, 3 * d, g->dense_qkv_bias, c.dqkv_bf);
  }
  {
    LaunchScope _ls("nb_ln_aug_kernel", st);
    nb_ln_aug_kernel<<<(R + 7) / 8, 256, 0, st>>>((const __nv_bfloat16 *)h, w->norm_mha_gamma, w->norm_mha_beta, eps, R, d, c.hn_aug);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  if (side != st) {
    EGT_CHECK_CUDA(cudaEventRecord(ev_operands, st));
    EGT_CHECK_CUDA(cudaStreamWaitEvent(side, ev_operands, 0));
  }
  {
    LaunchScope _ls("cublas_gemm", side);
    if ((rc = gemm_rm(hs, true, false, d, 3 * d, R, c.hn_aug, d + 8, c.dqkv_bf, 3 * d, g->dense_qkv_kernel, CUDA_R_32F, 3 * d, 1.f))) return rc;
  }
  LaunchScope _ls("cublas_gemm", st);
  return gemm_rm(hd, false, true, R, d, 3 * d, c.dqkv_bf, 3 * d, c.w_qkv, 3 * d, dhn, CUDA_R_32F, d, 0.f);
}

// ------------------------------------------------------------------------------------------------------------------
// Feed-forward half of a layer for the channel widths the tcgen05 kernels of ffn_tc.cu do not serve (node channel at
// d = 96 / 128, any hidden width): the same cuBLAS + element-wise-kernel pattern.
//   reference: ffnlr1 / ffnact / ffnlr2, ffn_block (lib/models/graph_xformer_model_base.py:229-258, :309-324)
namespace {

// w1_aug [w+8, hid] = [W1 ; b1 ; 0], w1 [w, hid], w2 [hid, w]   (bf16 operand copies)
__global__ void __launch_bounds__(256) fb_prep_kernel(const float *W1, const float *b1, const float *W2, int w, int hid,
                                                      __nv_bfloat16 *w1_aug, __nv_bfloat16 *w1, __nv_bfloat16 *w2) {
  const int tot_aug = (w + 8) * hid;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < tot_aug; i += gridDim.x * 256) {
    const int r = i / hid, c = i % hid;
    const float v = r < w ? W1[i] : r == w ? b1[c] : 0.f;
    w1_aug[i] = __float2bfloat16_rn(v);
    if (r < w && w1) w1[i] = __float2bfloat16_rn(v);
  }
  for (int i = blockIdx.x * 256 + threadIdx.x; i < hid * w; i += gridDim.x * 256) w2[i] = __float2bfloat16_rn(W2[i]);
}

// hid = act(pre)   (pre already carries the bias)
__global__ void __launch_bounds__(256) fb_act_kernel(const __nv_bfloat16 *pre, __nv_bfloat16 *hid, size_t n, int act) {
  for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * 512) {   // n even
    const __nv_bfloat162 p = *(const __nv_bfloat162 *)(pre + i);
    *(__nv_bfloat162 *)(hid + i) = __floats2bfloat162_rn(edge_act_fwd(act, 0.2f, __bfloat162float(p.x)),
                                                         edge_act_fwd(act, 0.2f, __bfloat162float(p.y)));
  }
}
// dpre = dhid * act'(pre), in place over dhid
__global__ void __launch_bounds__(256) fb_dact_kernel(const __nv_bfloat16 *pre, __nv_bfloat16 *dhid, size_t n, int act) {
  for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * 512) {
    const __nv_bfloat162 p = *(const __nv_bfloat162 *)(pre + i), d = *(const __nv_bfloat162 *)(dhid + i);
    *(__nv_bfloat162 *)(dhid + i) = __floats2bfloat162_rn(__bfloat162float(d.x) * edge_act_bwd(act, 0.2f, __bfloat162float(p.x)),
                                                          __bfloat162float(d.y) * edge_act_bwd(act, 0.2f, __bfloat162float(p.y)));
  }
}

struct FfnCarve {
  __nv_bfloat16 *xe_aug, *pre, *hid, *w1_aug, *w1, *w2;
  float *tmp;
  void *blas_ws;
};
FfnCarve ffn_carve(void *base, size_t R, int w, int hid) {
  char *p = (char *)base;
  FfnCarve c;
  auto take = [&](size_t bytes) { char *q = p; p += al(bytes); return q; };
  c.xe_aug = (__nv_bfloat16 *)take(R * (w + 8) * 2);
  c.pre = (__nv_bfloat16 *)take(R * hid * 2);
  c.hid = (__nv_bfloat16 *)take(R * hid * 2);
  c.w1_aug = (__nv_bfloat16 *)take((size_t)(w + 8) * hid * 2);
  c.w1 = (__nv_bfloat16 *)take((size_t)w * hid * 2);
  c.w2 = (__nv_bfloat16 *)take((size_t)hid * w * 2);
  c.tmp = (float *)take(R * w * 4);
  c.blas_ws = take(kBlasWs);
  return c;
}
int ffn_begin(cublasHandle_t *h, void *blas_ws, cudaStream_t st) {
  int rc = get_handle(h);
  if (rc) return rc;
  cublasStatus_t s = cublasSetStream(*h, st);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasSetStream failed with status %d", (int)s);
  s = cublasSetWorkspace(*h, blas_ws, kBlasWs);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasSetWorkspace failed with status %d", (int)s);
  return EGT_OK;
}
unsigned ew_grid(size_t n, int per_thread) {
  const size_t b = (n / per_thread + 255) / 256;
  return (unsigned)(b < 1 ? 1 : b < 4096 ? b : 4096);
}

}  // namespace

// bf16, smooth activations (the derivative of relu / lrelu flips with the bf16 rounding of a pre-activation near
// zero), widths that are multiples of 8, enough rows for a GEMM to pay
bool ffn_blas_supported(const egt_ffn_cfg_t *cfg) {
  static const bool off = getenv("EGT_FFN_BLAS") && atoi(getenv("EGT_FFN_BLAS")) == 0;
  return !off && cfg->dtype == EGT_BF16 && cfg->activation != EGT_ACT_RELU && cfg->activation != EGT_ACT_LRELU &&
         cfg->width % 8 == 0 && cfg->hidden % 8 == 0 && cfg->width <= 256 && cfg->rows >= 512 && cfg->rows < (1ll << 31);
}

size_t ffn_blas_workspace_bytes(const egt_ffn_cfg_t *cfg) {
  const size_t R = (size_t)cfg->rows, w = cfg->width, hid = cfg->hidden;
  return al(R * (w + 8) * 2) + 2 * al(R * hid * 2) + al((w + 8) * hid * 2) + 2 * al(w * hid * 2) + al(R * w * 4) + al(kBlasWs);
}

int ffn_blas_fwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *wt, const void *x, void *y, void *ws, cudaStream_t st) {
  const int R = (int)cfg->rows, w = cfg->width, hid = cfg->hidden;
  const FfnCarve c = ffn_carve(ws, R, w, hid);
  cublasHandle_t hd;
  int rc = ffn_begin(&hd, c.blas_ws, st);
  if (rc) return rc;
  {
    LaunchScope _ls("fb_prep_kernel", st);
    fb_prep_kernel<<<64, 256, 0, st>>>(wt->lr1_kernel, wt->lr1_bias, wt->lr2_kernel, w, hid, c.w1_aug, nullptr, c.w2);
  }
  {
    LaunchScope _ls("nb_ln_aug_kernel", st);
    nb_ln_aug_kernel<<<(R + 7) / 8, 256, 0, st>>>((const __nv_bfloat16 *)x, wt->norm_gamma, wt->norm_beta, cfg->ln_eps, R, w, c.xe_aug);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, false, false, R, hid, w + 8, c.xe_aug, w + 8, c.w1_aug, hid, c.pre, CUDA_R_16BF, hid, 0.f))) return rc;
  }
  const size_t nh = (size_t)R * hid, nw = (size_t)R * w;
  {
    LaunchScope _ls("fb_act_kernel", st);
    fb_act_kernel<<<ew_grid(nh, 2), 256, 0, st>>>(c.pre, c.pre, nh, cfg->activation);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, false, false, R, w, hid, c.pre, hid, c.w2, w, c.tmp, CUDA_R_32F, w, 0.f))) return rc;
  }
  LaunchScope _ls("nb_out_epilogue_kernel", st);
  nb_out_epilogue_kernel<<<ew_grid(nw, 4) < 2048 ? ew_grid(nw, 4) : 2048, 256, 0, st>>>((const __nv_bfloat16 *)x, c.tmp, wt->lr2_bias, nw, w,
                                                                                          (__nv_bfloat16 *)y);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

int ffn_blas_bwd(const egt_ffn_cfg_t *cfg, const egt_ffn_weights_t *wt, const egt_ffn_grads_t *g, const void *x, const void *dy,
                 void *dx, void *ws, cudaStream_t st) {
  const int R = (int)cfg->rows, w = cfg->width, hid = cfg->hidden;
  const FfnCarve c = ffn_carve(ws, R, w, hid);
  cublasHandle_t hd;
  int rc = ffn_begin(&hd, c.blas_ws, st);
  if (rc) return rc;
  const size_t nh = (size_t)R * hid;
  {
    LaunchScope _ls("fb_prep_kernel", st);
    fb_prep_kernel<<<64, 256, 0, st>>>(wt->lr1_kernel, wt->lr1_bias, wt->lr2_kernel, w, hid, c.w1_aug, c.w1, c.w2);
  }
  {
    LaunchScope _ls("nb_ln_aug_kernel", st);
    nb_ln_aug_kernel<<<(R + 7) / 8, 256, 0, st>>>((const __nv_bfloat16 *)x, wt->norm_gamma, wt->norm_beta, cfg->ln_eps, R, w, c.xe_aug);
  }
  {
    LaunchScope _ls("nb_colsum_kernel", st);
    nb_colsum_kernel<__nv_bfloat16><<<(R + kColsumRows - 1) / kColsumRows, 256, 0, st>>>((const __nv_bfloat16 *)dy, R, w, g->lr2_bias, nullptr);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, false, false, R, hid, w + 8, c.xe_aug, w + 8, c.w1_aug, hid, c.pre, CUDA_R_16BF, hid, 0.f))) return rc;   // pre
  }
  {
    LaunchScope _ls("fb_act_kernel", st);
    fb_act_kernel<<<ew_grid(nh, 2), 256, 0, st>>>(c.pre, c.hid, nh, cfg->activation);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, true, false, hid, w, R, c.hid, hid, dy, w, g->lr2_kernel, CUDA_R_32F, w, 1.f))) return rc;       // dW2 += hid^T dy
    if ((rc = gemm_rm(hd, false, true, R, hid, w, dy, w, c.w2, w, c.hid, CUDA_R_16BF, hid, 0.f))) return rc;              // dhid = dy W2^T
  }
  {
    LaunchScope _ls("fb_dact_kernel", st);
    fb_dact_kernel<<<ew_grid(nh, 2), 256, 0, st>>>(c.pre, c.hid, nh, cfg->activation);                                    // dpre
  }
  {
    LaunchScope _ls("nb_colsum_kernel", st);
    nb_colsum_kernel<__nv_bfloat16><<<(R + kColsumRows - 1) / kColsumRows, 256, 0, st>>>(c.hid, R, hid, g->lr1_bias, nullptr);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  {
    LaunchScope _ls("cublas_gemm", st);
    if ((rc = gemm_rm(hd, true, false, w, hid, R, c.xe_aug, w + 8, c.hid, hid, g->lr1_kernel, CUDA_R_32F, hid, 1.f))) return rc;   // dW1 += LN(x)^T dpre
    if ((rc = gemm_rm(hd, false, true, R, w, hid, c.hid, hid, c.w1, hid, c.tmp, CUDA_R_32F, w, 0.f))) return rc;                  // d LN(x) = dpre W1^T
  }
  LnBwdArgs lb;
  lb.x = x; lb.dy = c.tmp; lb.dres = dy; lb.gamma = wt->norm_gamma; lb.eps = cfg->ln_eps; lb.dx = dx;
  lb.dgamma = g->norm_gamma; lb.dbeta = g->norm_beta; lb.R = R; lb.D = w;
  return ln_bwd_launch(lb, EGT_BF16, st);
}

}  // namespace egt
