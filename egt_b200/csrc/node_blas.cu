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

cublasHandle_t g_handle[16] = {};
std::mutex g_handle_mu;

int get_handle(cublasHandle_t *out) {
  int dev = 0;
  EGT_CHECK_CUDA(cudaGetDevice(&dev));
  EGT_REQUIRE(dev >= 0 && dev < 16, EGT_E_ARG, "node_blas: device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_handle_mu);
  if (!g_handle[dev]) {
    cublasStatus_t s = cublasCreate(&g_handle[dev]);
    EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasCreate failed with status %d", (int)s);
  }
  *out = g_handle[dev];
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

// db[c] += sum_r Y[r, c] and, optionally, a bf16 copy of a float32 Y.  CTA = 64 rows x all columns.
template <typename T>
__global__ void __launch_bounds__(256) nb_colsum_kernel(const T *Y, int R, int C, float *db, __nv_bfloat16 *copy) {
  const int r0 = blockIdx.x * 64, r1 = min(R, r0 + 64);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float v = ldf(Y + (size_t)r * C + c);
      acc += v;
      if (copy) copy[(size_t)r * C + c] = __float2bfloat16_rn(v);
    }
    atomicAdd(db + c, acc);
  }
}

size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carve {
  __nv_bfloat16 *hn_aug, *w_aug, *w_qkv, *w_o, *dqkv_bf;
  float *tmp;
  void *blas_ws;
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
  return c;
}

int begin(cublasHandle_t *h, const Carve &c, cudaStream_t st) {
  int rc = get_handle(h);
  if (rc) return rc;
  cublasStatus_t s = cublasSetStream(*h, st);
  EGT_REQUIRE(s == CUBLAS_STATUS_SUCCESS, EGT_E_CUDA, "cublasSetStream failed with status %d", (int)s);
  s = cublasSetWorkspace(*h, c.blas_ws, kBlasWs);
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
         al((size_t)R * 3 * d * 2) + al((size_t)R * d * 4) + al(kBlasWs);
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

// dV_att = dh' W_O^T ; dW_O += V_att^T dh' ; db_O += colsum(dh')
int node_blas_bwd1(const void *dh_out, const void *v_att, const egt_block_weights_t *w, const egt_block_grads_t *g, void *d_v_att,
                   int R, int d, void *ws, cudaStream_t st) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  {
    LaunchScope _ls("nb_prep_kernel", st);
    nb_prep_kernel<<<64, 256, 0, st>>>(w->dense_qkv_kernel, w->dense_qkv_bias, w->dense_mha_kernel, d, 1.f, c.w_aug, c.w_qkv, c.w_o);
  }
  {
    LaunchScope _ls("nb_colsum_kernel", st);
    nb_colsum_kernel<__nv_bfloat16><<<(R + 63) / 64, 256, 0, st>>>((const __nv_bfloat16 *)dh_out, R, d, g->dense_mha_bias, nullptr);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  LaunchScope _ls("cublas_gemm", st);
  if ((rc = gemm_rm(hd, false, true, R, d, d, dh_out, d, c.w_o, d, d_v_att, CUDA_R_16BF, d, 0.f))) return rc;
  return gemm_rm(hd, true, false, d, d, R, v_att, d, dh_out, d, g->dense_mha_kernel, CUDA_R_32F, d, 1.f);
}

// dW_qkv += LN(h)^T dqkv ; db_qkv += colsum(dqkv) ; dhn = dqkv W_qkv^T (float32, for ln_bwd_kernel)
// (w_qkv was converted by node_blas_bwd1 of the same backward call)
int node_blas_bwd2(const void *h, const float *d_qkv, const egt_block_weights_t *w, const egt_block_grads_t *g, float eps,
                   float *dhn, int R, int d, void *ws, cudaStream_t st) {
  const Carve c = carve(ws, R, d);
  cublasHandle_t hd;
  int rc = begin(&hd, c, st);
  if (rc) return rc;
  {
    LaunchScope _ls("nb_colsum_kernel", st);
    nb_colsum_kernel<float><<<(R + 63) / 64, 256, 0, st>>>(d_qkv, R, 3 * d, g->dense_qkv_bias, c.dqkv_bf);
  }
  {
    LaunchScope _ls("nb_ln_aug_kernel", st);
    nb_ln_aug_kernel<<<(R + 7) / 8, 256, 0, st>>>((const __nv_bfloat16 *)h, w->norm_mha_gamma, w->norm_mha_beta, eps, R, d, c.hn_aug);
  }
  EGT_CHECK_CUDA(cudaGetLastError());
  LaunchScope _ls("cublas_gemm", st);
  if ((rc = gemm_rm(hd, true, false, d, 3 * d, R, c.hn_aug, d + 8, c.dqkv_bf, 3 * d, g->dense_qkv_kernel, CUDA_R_32F, 3 * d, 1.f))) return rc;
  return gemm_rm(hd, false, true, R, d, 3 * d, c.dqkv_bf, 3 * d, c.w_qkv, 3 * d, dhn, CUDA_R_32F, d, 0.f);
}

}  // namespace egt
