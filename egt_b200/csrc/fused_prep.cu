// fused_prep.cu -- per-call derived weights of the fused tcgen05 path, and the host-side helpers
// (shape gate, TMA descriptor encoding).
#include <math.h>
#include <mutex>
#include "common.cuh"
#include "fused.h"

namespace egt {

bool fused_supported(const egt_block_cfg_t *c, int dtype) {
  const egt_attn_cfg_t &a = c->attn;
  return dtype == EGT_BF16 && a.h == FH && a.dk == FDK && c->d_e == FDE &&
         c->edge_channel_type == EGT_EDGE_RESIDUAL && c->gate_attention && a.has_clip &&
         a.clip_lo <= a.clip_hi && c->edge_act == EGT_ACT_NONE && !(a.training && a.attn_dropout > 0.f) &&
         a.N >= 1 && a.N <= 4096;
}

// One CTA of 128 threads.  B-operand images are K-major without swizzle: element (n, k) of an [N x K]
// matrix lives at  (k/8) * (N*16) + n*16 + (k%8)*2  bytes (8x16-byte core matrices, LBO = N*16, SBO = 128).
__global__ void __launch_bounds__(128) fused_prep_kernel(egt_block_weights_t w, float clip_lo, float clip_hi,
                                                         FusedPrep *out) {
  __shared__ float wp[2][FDE][FH];   // W' rounded to bf16
  const int tid = threadIdx.x;
  {
    int eg = tid / 64, c = (tid / 8) % 8, hh = tid % 8;
    const float *W = eg ? w.attention_gates_kernel : w.dense_edge_b_kernel;
    float v = __bfloat162float(__float2bfloat16_rn(w.norm_edge_gamma[c] * W[c * FH + hh]));
    wp[eg][c][hh] = v;
    out->wp[eg][c][hh] = v;
  }
  __syncthreads();
  if (tid < 16) {
    int eg = tid / 8, hh = tid % 8;
    const float *W = eg ? w.attention_gates_kernel : w.dense_edge_b_kernel;
    const float *bias = eg ? w.attention_gates_bias : w.dense_edge_b_bias;
    float u = 0.f, v = bias[hh], n2 = 0.f;
    for (int c = 0; c < FDE; ++c) {
      u += wp[eg][c][hh];
      v += w.norm_edge_beta[c] * W[c * FH + hh];
      n2 += wp[eg][c][hh] * wp[eg][c][hh];
    }
    (eg ? out->uG : out->uE)[hh] = u;
    (eg ? out->vG : out->vE)[hh] = v;
    if (eg == 0) {
      // |LN(e)_c| has l2 norm <= sqrt(d_e)  =>  |E| <= sqrt(d_e) * ||W'[:,hh]|| + |v|
      float bnd = fmaxf(fabsf(clip_lo), fabsf(clip_hi)) + sqrtf((float)FDE * n2) + fabsf(v);
      for (int o = 4; o > 0; o >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffu, bnd, o));
      if (hh == 0) out->bound = bnd;
    }
  }
  if (tid < FDE) out->br[tid] = w.dense_edge_r_bias[tid];
  // wblk: n = key*16 + eg*8 + hh ; k = key'*8 + c            (N = 32, K = 16)
  for (int i = tid; i < 32 * 16; i += 128) {
    int n = i / 16, k = i % 16;
    int key = n / 16, eg = (n / 8) % 2, hh = n % 8, key2 = k / 8, c = k % 8;
    float v = key == key2 ? wp[eg][c][hh] : 0.f;
    out->wblk[(k / 8) * (32 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(v);
  }
  // wrblk: n = key*8 + c ; k = key'*8 + hh                   (N = 16, K = 16)   value W_r[hh][c]
  // wrtblk: n = key*8 + hh ; k = key'*8 + c                  (N = 16, K = 16)   value W_r[hh][c]
  for (int i = tid; i < 16 * 16; i += 128) {
    int n = i / 16, k = i % 16;
    int key = n / 8, a = n % 8, key2 = k / 8, b = k % 8;
    out->wrblk[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
        __float2bfloat16_rn(key == key2 ? w.dense_edge_r_kernel[b * FDE + a] : 0.f);
    out->wrtblk[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
        __float2bfloat16_rn(key == key2 ? w.dense_edge_r_kernel[a * FDE + b] : 0.f);
  }
  // backward images (head-group ordered, fused.h)
  for (int i = tid; i < 32 * 16; i += 128) {      // b_eg
    int n = i / 16, k = i % 16;
    int g = n / 16, key = (n / 8) % 2, eg = (n / 4) % 2, hh = 4 * g + n % 4, key2 = k / 8, c = k % 8;
    out->b_eg[(k / 8) * (32 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wp[eg][c][hh] : 0.f);
  }
  for (int i = tid; i < 16 * 16; i += 128) {      // b_hx, b_de[g]
    int n = i / 16, k = i % 16;
    {
      int g = n / 8, key = (n / 4) % 2, hh = 4 * g + n % 4, key2 = k / 8, c = k % 8;
      out->b_hx[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
          __float2bfloat16_rn(key == key2 ? w.dense_edge_r_kernel[hh * FDE + c] : 0.f);
    }
    {
      int key2 = n / 8, c = n % 8, g = k / 8, key = (k / 4) % 2, hh = 4 * g + k % 4;
      out->b_wr[(k / 8) * (16 * 8) + n * 8 + (k % 8)] =
          __float2bfloat16_rn(key == key2 ? w.dense_edge_r_kernel[hh * FDE + c] : 0.f);
    }
    for (int g = 0; g < 2; ++g) {
      int key2 = n / 8, c = n % 8, key = k / 8, eg = (k / 4) % 2, hh = 4 * g + k % 4;
      out->b_de[g][(k / 8) * (16 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wp[eg][c][hh] : 0.f);
    }
  }
  // wtblk: n = key*8 + c ; k = key'*16 + eg*8 + hh           (N = 16, K = 32)   value W'_eg[c][hh]
  for (int i = tid; i < 16 * 32; i += 128) {
    int n = i / 32, k = i % 32;
    int key = n / 8, c = n % 8, key2 = k / 16, eg = (k / 8) % 2, hh = k % 8;
    out->wtblk[(k / 8) * (16 * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(key == key2 ? wp[eg][c][hh] : 0.f);
  }
}

int fused_prep_launch(const egt_block_cfg_t *cfg, const egt_block_weights_t *w, FusedPrep *prep, cudaStream_t st) {
  LaunchScope _ls("fused_prep_kernel", st);
  fused_prep_kernel<<<1, 128, 0, st>>>(*w, cfg->attn.clip_lo, cfg->attn.clip_hi, prep);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}

// ---- TMA descriptors ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int encode_tmap_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle128) {
  EncodeTiledFn fn = encode_fn();
  EGT_REQUIRE(fn != nullptr, EGT_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  EGT_REQUIRE(((uintptr_t)base & 15) == 0 && stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, EGT_E_ALIGN,
              "TMA needs 16-byte aligned tensors and row strides");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGT_REQUIRE(r == CUDA_SUCCESS, EGT_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EGT_OK;
}

}  // namespace egt
