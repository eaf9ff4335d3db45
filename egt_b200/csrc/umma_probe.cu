// umma_probe.cu -- bring-up / regression probe for every tcgen05.mma operand form the fused kernels use.
//
// egt_debug_umma_probe() lays a logical A[128,K] and B[N,K] (bf16) out in shared memory (or TMEM for A)
// in one of the canonical tcgen05 layouts, issues K/16 tcgen05.mma instructions with the requested
// descriptor fields and returns D[128,N] = A * B^T (fp32).  tests/test_umma_probe.py checks each form
// against a plain matmul, so a wrong descriptor shows up as a failing unit test rather than as a
// wrong attention output.
#include "common.cuh"
#include "umma.cuh"

namespace egt {
using namespace umma;

struct ProbeParams {
  int a_mode, b_mode;        // 0 K-major SW128, 1 TMEM (A only), 2 MN-major SW128, 3 K-major none, 4 MN-major none, 5 K-major SW128 with 16-row atoms (B only)
  int N, ksteps;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_off[8], b_off[8];   // start-address byte offset (smem) or column offset (TMEM) of k-step s
  const __nv_bfloat16 *A, *B;
  float *D;
};

__device__ __forceinline__ uint32_t image_off(int mode, uint32_t mn, uint32_t k, uint32_t lbo, uint32_t sbo) {
  switch (mode) {
    case 0: return (k >> 6) * 16384u + sw128_off(mn, k & 63u);
    case 5: return (k >> 6) * 2048u + sw128_off(mn, k & 63u);          // 16-row atoms (expanded K / V operands)
    case 2: return (mn >> 6) * lbo + (k >> 3) * sbo + (k & 7u) * 128u + (((((mn & 63u) >> 3) ^ k) & 7u) << 4) + ((mn & 7u) << 1);
    case 3: return (k >> 3) * lbo + (mn >> 3) * sbo + (mn & 7u) * 16u + ((k & 7u) << 1);
    case 4: return (mn >> 3) * sbo + (k >> 3) * lbo + (k & 7u) * 16u + ((mn & 7u) << 1);
  }
  return 0;
}
__device__ __forceinline__ uint32_t layout_of(int mode) { return (mode == 0 || mode == 2 || mode == 5) ? LAYOUT_SW128 : LAYOUT_NONE; }

__global__ void __launch_bounds__(128) umma_probe_kernel(ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem, *sB = smem + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = 16 * p.ksteps;
  for (int i = tid; i < 65536 / 16; i += 128) ((uint4 *)smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (p.a_mode != 1)
    for (int i = tid; i < 128 * K; i += 128) {
      int m = i / K, k = i % K;
      *(__nv_bfloat16 *)(sA + image_off(p.a_mode, m, k, p.a_lbo, p.a_sbo)) = p.A[i];
    }
  for (int i = tid; i < p.N * K; i += 128) {
    int n = i / K, k = i % K;
    *(__nv_bfloat16 *)(sB + image_off(p.b_mode, n, k, p.b_lbo, p.b_sbo)) = p.B[i];
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  if (p.a_mode == 1) {   // A row `tid` as packed bf16 into TMEM columns [256, 256 + K/2)
    for (int s = 0; s < p.ksteps; ++s) {
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 v = *(const __nv_bfloat162 *)(p.A + (size_t)tid * K + s * 16 + 2 * j);
        r[j] = *(uint32_t *)&v;
      }
      tmem_st8(lane_base + 256 + p.a_off[s], r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, p.N, p.a_mode == 2 || p.a_mode == 4, p.b_mode == 2 || p.b_mode == 4);
    for (int s = 0; s < p.ksteps; ++s) {
      uint64_t bd = smem_desc(smem_u32(sB) + p.b_off[s], p.b_lbo, p.b_sbo, layout_of(p.b_mode));
      if (p.a_mode == 1) {
        mma_ts(tmem, tmem + 256 + p.a_off[s], bd, idesc, s > 0);
      } else {
        uint64_t ad = smem_desc(smem_u32(sA) + p.a_off[s], p.a_lbo, p.a_sbo, layout_of(p.a_mode));
        mma_ss(tmem, ad, bd, idesc, s > 0);
      }
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < p.N; c += 8) {
    uint32_t r[8];
    tmem_ld8(lane_base + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) p.D[(size_t)tid * p.N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace egt

extern "C" int egt_debug_umma_probe(int a_mode, int b_mode, int N, int ksteps, uint32_t a_lbo, uint32_t a_sbo,
                                    uint32_t b_lbo, uint32_t b_sbo, const uint32_t *a_off_host,
                                    const uint32_t *b_off_host, const void *A, const void *B, float *D, void *stream) {
  using namespace egt;
  EGT_REQUIRE(ksteps >= 1 && ksteps <= 8 && N >= 16 && N <= 256 && N % 16 == 0, EGT_E_ARG, "probe: bad N / ksteps");
  ProbeParams p;
  p.a_mode = a_mode; p.b_mode = b_mode; p.N = N; p.ksteps = ksteps;
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.b_lbo = b_lbo; p.b_sbo = b_sbo;
  for (int i = 0; i < 8; ++i) { p.a_off[i] = i < ksteps ? a_off_host[i] : 0; p.b_off[i] = i < ksteps ? b_off_host[i] : 0; }
  p.A = (const __nv_bfloat16 *)A; p.B = (const __nv_bfloat16 *)B; p.D = D;
  const int smem = 65536 + 1024;
  EGT_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  LaunchScope _ls("umma_probe_kernel", (cudaStream_t)stream);
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  EGT_CHECK_CUDA(cudaGetLastError());
  return EGT_OK;
}
