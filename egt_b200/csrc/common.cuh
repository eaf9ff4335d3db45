// common.cuh -- shared device helpers for the EGT attention-block kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/egt_b200.h"

namespace egt {

// Additive mask constant of the reference: (x-1)*1e9 and -1e9 (egt_layers.py:92,99,106).
// The kernels add it in fp32 exactly like TF does, so a masked logit becomes exactly -1e9
// (|x| < 32 is absorbed by the 64-ulp), masked P and g are exactly 0, and the
// all-keys-masked row degenerates to the same uniform softmax.
constexpr float kNegMask = 1e9f;
constexpr float kLog2e = 1.4426950408889634f;

void set_error(int code, const char *fmt, ...);

// Counts every kernel launch of the library and, when egt_profile_enable(1) was called, brackets
// it with CUDA events on the launching stream (bench.py reads the per-kernel times back).
struct LaunchScope {
  int slot; cudaStream_t st; cudaEvent_t e1;
  LaunchScope(const char *name, cudaStream_t stream);
  ~LaunchScope();
};

#define EGT_CHECK_CUDA(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::egt::set_error(EGT_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                             \
      return EGT_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

#define EGT_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::egt::set_error(code, __VA_ARGS__);      \
      return code;                              \
    }                                           \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Kernels of one step run back to back on one stream.  Each kernel lets its successor start early
// (pdl_trigger) and does its own set-up (shared-memory / TMEM / barrier initialisation, weight images)
// before it waits for its predecessor's memory to be visible (pdl_wait).  Rule kept by every kernel of this
// library: no global WRITE and no global read of anything but the block's parameters before pdl_wait().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // EGT_PDL=1 turns the launch attribute on (abi.cu; off by default, it measured slower)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- element access templated on the activation dtype --------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T *p);
template <> __device__ __forceinline__ float ldf<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T *p, float v);
template <> __device__ __forceinline__ void stf<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
  *p = __float2bfloat16_rn(v);
}

__device__ __forceinline__ float sigmoid_f(float x) {
  // 1/(1+exp(-x)); x = -1e9 gives exp -> +inf -> exactly 0 (the masked-gate contract).
  return 1.0f / (1.0f + __expf(-x));
}

// ---- Philox4x32-10 counter RNG (shared by device kernels and the host test hook) -------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c0, hi0, lo0);
    philox_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// One Philox call yields eight 16-bit uniforms; element idx uses call idx>>3, lane idx&7.
// u = (bits16 + 0.5) / 65536  in (0,1).
__host__ __device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t offset, uint32_t stream_id,
                                                      uint64_t idx) {
  uint64_t q = idx >> 3;
  uint32_t lane = (uint32_t)(idx & 7);
  Philox4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset,
                            (uint32_t)(offset >> 32) + stream_id * 0x40000000u, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
  uint32_t word = lane < 2 ? r.x : lane < 4 ? r.y : lane < 6 ? r.z : r.w;
  uint32_t bits = (lane & 1) ? (word >> 16) : (word & 0xFFFFu);
  return ((float)bits + 0.5f) * (1.0f / 65536.0f);
}

// Element (b, l, m, hh) of a [B,N,N,h] tensor -> RNG element index.  The eight uniforms of one Philox call cover
// two consecutive keys x four consecutive heads, so a thread that owns (row, key pair, head quad) -- the work
// split of the fused kernels -- draws all of its bits with a single call.
__host__ __device__ __forceinline__ uint64_t rng_elem_index(uint64_t b, uint64_t l, uint64_t m, uint32_t hh, uint64_t N,
                                                            uint32_t h) {
  const uint64_t np = (N + 1) >> 1, nq = (h + 3u) >> 2;
  const uint64_t call = ((b * N + l) * np + (m >> 1)) * nq + (uint64_t)(hh >> 2);
  return (call << 3) | ((m & 1) << 2) | (uint64_t)(hh & 3u);
}

// activation of edge_channel_contrib (graph_xformer_model_base.py:149-162)
__device__ __forceinline__ float edge_act_fwd(int act, float alpha, float x) {
  switch (act) {
    case EGT_ACT_LRELU: return x >= 0.f ? x : alpha * x;
    case EGT_ACT_RELU: return fmaxf(x, 0.f);
    case EGT_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case EGT_ACT_TANH: return tanhf(x);
    case EGT_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// derivative wrt the pre-activation, given pre-activation x
__device__ __forceinline__ float edge_act_bwd(int act, float alpha, float x) {
  switch (act) {
    case EGT_ACT_LRELU: return x >= 0.f ? 1.f : alpha;
    case EGT_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case EGT_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    case EGT_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case EGT_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace egt
