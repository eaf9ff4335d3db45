// fused_prep.cu -- per-call derived weights of the fused tcgen05 path, and the host-side helpers
// (shape gate, TMA descriptor encoding).
#include <math.h>
#include <mutex>
#include "common.cuh"
#include "fused.h"
#include "fused_prep.cuh"

namespace egt {

bool fused_supported(const egt_block_cfg_t *c, int dtype) {
  const egt_attn_cfg_t &a = c->attn;
  return dtype == EGT_BF16 && a.h == FH && a.dk == FDK && c->d_e == FDE &&
         c->edge_channel_type == EGT_EDGE_RESIDUAL && c->gate_attention && a.has_clip &&
         a.clip_lo <= a.clip_hi && c->edge_act == EGT_ACT_NONE && !(a.training && a.attn_dropout > 0.f) &&
         a.N >= 1 && a.N <= 4096;
}

__global__ void __launch_bounds__(128) fused_prep_kernel(egt_block_weights_t w, float clip_lo, float clip_hi,
                                                         FusedPrep *out) {
  fused_prep_body(w, clip_lo, clip_hi, out, threadIdx.x);
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
